// Per-thread phases of the batched blind-rotation kernel (the hot loop of every gate bootstrap).
//
// Computes, for each "rotation job", exactly what the reference computes in
//   HomGate linear combination          TFHEpp include/gate.hpp:8-18 (and :236-240 for MUX)
//   BlindRotate<lvl01param>             TFHEpp include/gatebootstrapping.hpp:19-71
//     CMUXFFTwithPolynomialMulByXaiMinusOne   include/detwfa.hpp:36-49
//     trgswfftExternalProduct / Decomposition include/trgsw.hpp:62-78,102-131
//   SampleExtractIndex(.., 0)           TFHEpp include/trlwe.hpp:213-223
// with the polynomial products evaluated EXACTLY modulo 2^32 (see modarith.h).  The prior-art
// GPU formulation is cuFHE's __BlindRotatePreAdd__/Accumulate
// (cuFHE include/gatebootstrapping_gpu.cuh:55-223): one 768-thread block per gate, 64-bit NTT.
//
// B200 mapping: a CTA owns G jobs and 2G warps; warp w = 2g + q owns polynomial q (0 = A, 1 = B)
// of job g: it keeps that accumulator polynomial in registers, produces its three digit
// polynomials in NTT form (phase F), and after the CTA-wide pointwise stage (phase M, which
// streams the NTT-domain bootstrapping key once per CTA and reuses it for all G jobs) turns
// the three limb results back (phase I).  Only two __syncthreads per CMUX step.
//
// Step sequence of br3_kernel (and, identically, of the CPU simulator):
//   br_fwd3_a ; syncwarp ; br_fwd3_b ; syncwarp ; br_fwd3_c      rotated difference, 3 digits, forward NTT x3
//   pw_load(first position)            key words for phase M requested before the barrier
//   syncthreads ; br_pointwise ; syncthreads
//   br_inv3_a ; syncwarp ; br_inv3_b ; syncwarp ; br_inv3_c      inverse NTT x3, lift, recombine, accumulate
// br7_phases.h holds the 16-warp variant of the same protocol (the shape the launch plan prefers).
//
// Every function here is free of intra-phase cross-thread communication: threads talk only
// through shared memory between phases, so a sequential CPU loop over (phase, thread) is an
// exact model of the kernel.  tests/sim/br_sim.cpp relies on that.
#pragma once
#include "fhe_params.h"
#include "hd.h"
#include "modarith.h"
#include "ntt_warp.h"

namespace b200 {

// One blind rotation: c = s0*in0 + s1*in1 + s2*in2 + off (uint16 torus), then BlindRotate + extract.
struct BrJob {
    uint32_t in[3];   // arena slot ids (ignored where the sign is 0)
    int8_t sgn[3];
    int8_t pad;
    uint32_t off;     // added to the b coefficient, modulo 2^16
};

// Shared-memory carve-up of one CTA (G jobs).
template <int G>
struct BrSmem {
    static constexpr int DBUF_WORDS = G * ROWS * TILE_WORDS;  // digit / limb tiles (in place)
    static constexpr int ABAR_HALFS = G * SLOT_STRIDE;        // mod-switched a_i
    static constexpr size_t BYTES =
        (size_t)DBUF_WORDS * 4 + 2 * (size_t)TW2_LEN * sizeof(tw_t) + (size_t)ABAR_HALFS * 2;
    uint32_t* dbuf;
    tw_t* tw2f;
    tw_t* tw2i;
    uint16_t* abar;
    B200_HD void carve(void* base)
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(base);
        dbuf = reinterpret_cast<uint32_t*>(p);
        p += (size_t)DBUF_WORDS * 4;
        tw2f = reinterpret_cast<tw_t*>(p);
        p += (size_t)TW2_LEN * sizeof(tw_t);
        tw2i = reinterpret_cast<tw_t*>(p);
        p += (size_t)TW2_LEN * sizeof(tw_t);
        abar = reinterpret_cast<uint16_t*>(p);
    }
    B200_HD uint32_t* tile(int g, int row) const { return dbuf + (size_t)(g * ROWS + row) * TILE_WORDS; }
    // Natural-order copy of accumulator polynomial q, used only for the rotated read at the start
    // of a CMUX step and by the epilogue.  It lives in the first 1024 words of the warp's own first
    // tile, which is free between phase I of one step and phase F of the next.
    B200_HD uint32_t* acc(int g, int q) const { return tile(g, q * GL); }
};

// lvl0 linear combination of one coefficient (uint16 wrap-around, gate.hpp:14-16)
B200_HD uint32_t br_lincomb(const BrJob& job, const uint16_t* arena, int i)
{
    int32_t c = 0;
    B200_UNROLL
    for (int k = 0; k < 3; k++)
        if (job.sgn[k] != 0) c += (int32_t)job.sgn[k] * (int32_t)arena[(size_t)job.in[k] * SLOT_STRIDE + i];
    if (i == N0) c += (int32_t)job.off;
    return (uint32_t)c & 0xFFFFu;
}

// ---- prologue: mod switch + accumulator init -------------------------------------------
// gatebootstrapping.hpp:26-32 (b: no rounding), :58-65 (a_i: rounded, shift 5 for the uint16 torus)
template <int G>
B200_HD void br_prologue(const BrSmem<G>& sm, const BrJob& job, const uint16_t* arena, int g, int q, int lane,
                         uint32_t (&accr)[32])
{
    for (int i = q * 32 + lane; i < N0; i += 64) {
        const uint32_t c = br_lincomb(job, arena, i);
        sm.abar[g * SLOT_STRIDE + i] = (uint16_t)((c + 16u) >> 5);  // in [0, 2048]
    }
    const uint32_t bbar = 2u * N1 - (br_lincomb(job, arena, N0) >> 5);  // in [1, 2048]
    uint32_t* acc = sm.acc(g, q);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const int n = 32 * a + lane;
        uint32_t v = 0;
        if (q == 1) {  // testvector mu * X^bbar (utils.hpp:113-128)
            const uint32_t m = ((uint32_t)n - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        accr[a] = v;
        acc[n] = v;
    }
}

// ---- phase F: (X^abar - 1) * acc (utils.hpp:130-144), decomposition (trgsw.hpp:62-78), forward NTT.
// The three digit / limb transforms of a polynomial run in lock step inside the owning warp (ct_stage3 in ntt_warp.h).
// F3a: rotated difference -> three digits -> pass 1 x3; returns the three register sets so the
// caller can __syncwarp (the accumulator copy shares tile 3q) before F3b stores them.
template <int G>
B200_HD void br_fwd3_a(const BrSmem<G>& sm, int i, int g, int q, int lane, const uint32_t (&accr)[32],
                       uint32_t (&x0)[32], uint32_t (&x1)[32], uint32_t (&x2)[32])
{
    const uint32_t abar = sm.abar[g * SLOT_STRIDE + i];
    const uint32_t* acc = sm.acc(g, q);
    const uint32_t base = ((uint32_t)lane - abar) & (2u * N1 - 1);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const uint32_t m = (base + 32u * a) & (2u * N1 - 1);
        const uint32_t v = acc[m & (N1 - 1)];
        const uint32_t neg = 0u - ((m >> NBIT) & 1u);
        const uint32_t dv = ((v ^ neg) - neg) - accr[a] + (DEC_OFFSET + DEC_ROUND);
        constexpr uint32_t mask = (1u << BGBIT) - 1, bias = P - (1u << (BGBIT - 1));
        x0[a] = ((dv >> (32 - BGBIT)) & mask) + bias;
        x1[a] = ((dv >> (32 - 2 * BGBIT)) & mask) + bias;
        x2[a] = ((dv >> (32 - 3 * BGBIT)) & mask) + bias;
    }
    fwd_pass1_x3(x0, x1, x2);
}
// F3b (after __syncwarp): three column stores
template <int G>
B200_HD void br_fwd3_b(const BrSmem<G>& sm, int g, int q, int lane, const uint32_t (&x0)[32],
                       const uint32_t (&x1)[32], const uint32_t (&x2)[32])
{
    tile_store_col(sm.tile(g, q * GL + 0), x0, lane);
    tile_store_col(sm.tile(g, q * GL + 1), x1, lane);
    tile_store_col(sm.tile(g, q * GL + 2), x2, lane);
}
// F3c (after __syncwarp): three row loads -> pass 2 x3 -> three row stores
template <int G>
B200_HD void br_fwd3_c(const BrSmem<G>& sm, int g, int q, int lane)
{
    uint32_t x0[32], x1[32], x2[32];
    uint32_t* t = sm.tile(g, q * GL);
    tile_load_row(t, x0, lane);
    tile_load_row(t + TILE_WORDS, x1, lane);
    tile_load_row(t + 2 * TILE_WORDS, x2, lane);
    fwd_pass2_x3(x0, x1, x2, sm.tw2f, lane);
    tile_store_row(t, x0, lane);
    tile_store_row(t + TILE_WORDS, x1, lane);
    tile_store_row(t + 2 * TILE_WORDS, x2, lane);
}
// I3a: three row loads -> inverse pass 1 x3 -> three row stores
template <int G>
B200_HD void br_inv3_a(const BrSmem<G>& sm, int g, int q, int lane)
{
    uint32_t x0[32], x1[32], x2[32];
    uint32_t* t = sm.tile(g, q * LIMBS);
    tile_load_row(t, x0, lane);
    tile_load_row(t + TILE_WORDS, x1, lane);
    tile_load_row(t + 2 * TILE_WORDS, x2, lane);
    inv_pass1_x3(x0, x1, x2, sm.tw2i, lane);
    tile_store_row(t, x0, lane);
    tile_store_row(t + TILE_WORDS, x1, lane);
    tile_store_row(t + 2 * TILE_WORDS, x2, lane);
}
// I3b (after __syncwarp): three column loads -> inverse pass 2 x3 -> lift, recombine, accumulate,
// refresh the accumulator copy (tile 3q is dead by now: all three limbs are in registers)
template <int G>
B200_HD void br_inv3_b(const BrSmem<G>& sm, int g, int q, int lane, uint32_t (&accr)[32])
{
    uint32_t x0[32], x1[32], x2[32];
    uint32_t* t = sm.tile(g, q * LIMBS);
    tile_load_col(t, x0, lane);
    tile_load_col(t + TILE_WORDS, x1, lane);
    tile_load_col(t + 2 * TILE_WORDS, x2, lane);
    inv_pass2_x3(x0, x1, x2);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const uint32_t v = (uint32_t)centered_lift(x0[a]) + ((uint32_t)centered_lift(x1[a]) << LIMB_BITS) +
                           ((uint32_t)centered_lift(x2[a]) << (2 * LIMB_BITS));
        accr[a] += v;
    }
}
// I3c (after __syncwarp: every lane has finished its column loads of tile 3q): refresh the copy
template <int G>
B200_HD void br_inv3_c(const BrSmem<G>& sm, int g, int q, int lane, const uint32_t (&accr)[32])
{
    uint32_t* acc = sm.acc(g, q);
    B200_UNROLL
    for (int a = 0; a < 32; a++) acc[32 * a + lane] = accr[a];
}

// ---- phase M: pointwise multiply-accumulate with the NTT-domain key ----------------------
// bk_i points at bk_ntt[i] : [BK_COLS][ROWS][1024] uint32 in [0,p), pre-scaled by 2^32/N.
// out[g][c][j] = sum_r D[g][r][j] * BK[c][r][j] * 2^-32  (in place over the digit tiles).
// The 36 key words of position j are loaded once per CTA and reused for all G jobs; the load
// is split from the compute so the kernel can issue it a whole phase ahead (L2 latency).
B200_HD void pw_load(const uint32_t* bk_i, int j, uint32_t (&bkv)[BK_COLS][ROWS])
{
    B200_UNROLL
    for (int c = 0; c < BK_COLS; c++) {
        B200_UNROLL
        for (int r = 0; r < ROWS; r++) bkv[c][r] = bk_i[(size_t)(c * ROWS + r) * N1 + j];
    }
}
template <int G>
B200_HD void pw_compute(const BrSmem<G>& sm, int j, const uint32_t (&bkv)[BK_COLS][ROWS])
{
    const int off = tile_of_j(j);
    B200_UNROLL
    for (int g = 0; g < G; g++) {
        uint32_t d[ROWS], o[BK_COLS];
        B200_UNROLL
        for (int r = 0; r < ROWS; r++) d[r] = sm.tile(g, r)[off];
        B200_UNROLL
        for (int c = 0; c < BK_COLS; c++) {
            uint64_t acc = 0;
            B200_UNROLL
            for (int r = 0; r < ROWS; r++) acc += (uint64_t)d[r] * bkv[c][r];
            o[c] = redc64(acc);  // < 4p
        }
        B200_UNROLL
        for (int c = 0; c < BK_COLS; c++) sm.tile(g, c)[off] = o[c];
    }
}
// whole phase for thread tid with a two-deep register pipeline; bk0 holds position j = tid
// (already loaded by the caller, normally before the barrier that opens the phase)
template <int G>
B200_HD void br_pointwise(const BrSmem<G>& sm, const uint32_t* bk_i, int tid, uint32_t (&bk0)[BK_COLS][ROWS])
{
    constexpr int T = 64 * G;
    uint32_t bk1[BK_COLS][ROWS];
    for (int j = tid; j < N1; j += 2 * T) {
        const int j1 = j + T, j2 = j + 2 * T;
        if (j1 < N1) pw_load(bk_i, j1, bk1);
        pw_compute<G>(sm, j, bk0);
        if (j1 < N1) {
            if (j2 < N1) pw_load(bk_i, j2, bk0);
            pw_compute<G>(sm, j1, bk1);
        }
    }
}

// ---- epilogue: SampleExtractIndex(0) (trlwe.hpp:213-223) into the lvl1 scratch buffer -----
template <int G>
B200_HD void br_epilogue(const BrSmem<G>& sm, int g, int q, int lane, uint32_t* u_out)
{
    const uint32_t* acc = sm.acc(g, q);
    if (q == 0) {
        for (int j = lane; j < N1; j += 32) u_out[j] = (j == 0) ? acc[0] : 0u - acc[N1 - j];
    } else if (lane == 0) {
        u_out[N1] = acc[0];
    }
}

// ---- bootstrapping-key precomputation (once per key) ------------------------------------
// Prior art: cuFHE __TRGSW2NTT__ (src/bootstrap_gpu.cu:45-53).  One warp per (i, row, poly, limb):
// centred 11/11/10-bit limb of the raw uint32 key polynomial -> forward NTT -> * 2^32/N -> [0,p).
B200_HD int32_t bk_limb(uint32_t raw, int limb)
{
    // raw = x0 + 2^11*x1 + 2^22*x2 (mod 2^32), x0,x1 in [-1024,1023], x2 in [-512,512]
    uint32_t v = raw;
    const int32_t x0 = (int32_t)((v & 2047u) ^ 1024u) - 1024;
    v = (uint32_t)((int32_t)(v - (uint32_t)x0) >> LIMB_BITS);
    const int32_t x1 = (int32_t)((v & 2047u) ^ 1024u) - 1024;
    v = (uint32_t)((int32_t)(v - (uint32_t)x1) >> LIMB_BITS);
    return limb == 0 ? x0 : (limb == 1 ? x1 : (int32_t)v);
}
B200_HD void bk_prep_a(const uint32_t* raw_poly, int limb, int lane, uint32_t* tile)
{
    uint32_t x[32];
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const int32_t v = bk_limb(raw_poly[32 * a + lane], limb);
        x[a] = (uint32_t)(v + (int32_t)P);  // in (0, 2p)
    }
    fwd_pass1(x);
    tile_store_col(tile, x, lane);
}
B200_HD void bk_prep_b(const uint32_t* tile, const tw_t* tw2f, tw_t scale, int lane, uint32_t* out_poly)
{
    uint32_t x[32];
    tile_load_row(tile, x, lane);
    fwd_pass2(x, tw2f, lane);
    B200_UNROLL
    for (int b = 0; b < 32; b++) out_poly[32 * lane + b] = reduce_full(shoup_mul(x[b], scale));
}

}  // namespace b200
