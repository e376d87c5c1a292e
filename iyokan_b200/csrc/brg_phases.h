// Per-thread phases of brg_kernel: the GENERIC blind rotation, written against fhe_params.h only (any gadget length
// GL, any limb count LIMBS, 16- or 32-bit lvl0 torus).  It is the kernel of the 80-bit flavour (l = 2, five key limbs;
// CGGI16.hpp) and runs at 128 bits too, where the parity tests hold it against the same oracle as the shapes
// specialised for l = 3 (br_phases.h, br7_phases.h) - so the code that serves the 80-bit sweep is pinned twice.
//
// Reference functions: HomGate linear combination (TFHEpp gate.hpp:8-18), BlindRotate (gatebootstrapping.hpp:19-71),
// CMUXFFTwithPolynomialMulByXaiMinusOne (detwfa.hpp:36-49), trgswfftExternalProduct / Decomposition
// (trgsw.hpp:62-131), SampleExtractIndex (trlwe.hpp:213-223); products exact modulo 2^32 (modarith.h).
//
// Mapping: a CTA owns G jobs and 2G warps, warp (g, q) owns accumulator polynomial q of job g in registers and
// transforms its GL digits / LIMBS limbs one after the other (32 points per lane, XOR-swizzled 4 KB tiles); the
// pointwise stage is CTA-wide and reuses the key words of a position for all G jobs.  A job has
// NT = max(ROWS, BK_COLS) tiles (digits in, limb columns out, in place) plus a natural-order copy of the accumulator.
//
// Step sequence:  rotate_diff ; GL x { fwd_a(d) ; syncwarp ; fwd_b(d) } ; syncthreads ; pointwise ; syncthreads ;
//                 LIMBS x { inv_a(l) ; syncwarp ; inv_b(l) } ; acc_update ; syncwarp
#pragma once
#include "br_phases.h"

namespace b200 {

constexpr int BRG_NT = ROWS > BK_COLS ? ROWS : BK_COLS;   // tiles per job

template <int G>
struct BrgSmem {
    static constexpr int DBUF_WORDS = G * BRG_NT * STILE_WORDS;
    static constexpr int ACC_WORDS = G * 2 * N1;
    static constexpr int ABAR_HALFS = G * 640;
    static constexpr size_t BYTES = (size_t)DBUF_WORDS * 4 + (size_t)ACC_WORDS * 4 + 2 * (size_t)TW2_LEN * sizeof(tw_t) +
                                    (size_t)ABAR_HALFS * 2;
    uint32_t* dbuf;
    uint32_t* accb;
    tw_t* tw2f;
    tw_t* tw2i;
    uint16_t* abar;
    B200_HD void carve(void* base)
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(base);
        dbuf = reinterpret_cast<uint32_t*>(p);
        p += (size_t)DBUF_WORDS * 4;
        accb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)ACC_WORDS * 4;
        tw2f = reinterpret_cast<tw_t*>(p);
        p += (size_t)TW2_LEN * sizeof(tw_t);
        tw2i = reinterpret_cast<tw_t*>(p);
        p += (size_t)TW2_LEN * sizeof(tw_t);
        abar = reinterpret_cast<uint16_t*>(p);
    }
    B200_HD uint32_t* tile(int g, int t) const { return dbuf + (size_t)(g * BRG_NT + t) * STILE_WORDS; }
    B200_HD uint32_t* acc(int g, int q) const { return accb + (size_t)(g * 2 + q) * N1; }
};

// lvl0 linear combination of one coefficient, modulo the lvl0 word (gate.hpp:14-16)
B200_HD uint32_t brg_lincomb(const BrJob& job, const torus0_t* arena, int i)
{
    uint32_t c = 0;
    B200_UNROLL
    for (int k = 0; k < 3; k++)
        if (job.sgn[k] != 0) c += (uint32_t)(int32_t)job.sgn[k] * (uint32_t)arena[(size_t)job.in[k] * SLOT_STRIDE + i];
    if (i == N0) c += job.off;
    return c & T0_MASK;
}

// prologue: mod switch (gatebootstrapping.hpp:26-32 b not rounded, :58-65 a rounded; the sum is formed in the lvl0
// word for a 32-bit torus and in an int for the 16-bit one, so abar reaches 2N only at 16 bits) + test vector
template <int G>
B200_HD void brg_prologue(const BrgSmem<G>& sm, const BrJob& job, const torus0_t* arena, int g, int q, int lane,
                          uint32_t (&accr)[32])
{
    for (int i = q * 32 + lane; i < N0; i += 64) {
        const uint32_t c = brg_lincomb(job, arena, i);
        sm.abar[g * 640 + i] = (uint16_t)((uint32_t)(c + MODSW_ROUND) >> MODSW_SHIFT);
    }
    const uint32_t bbar = 2u * N1 - (brg_lincomb(job, arena, N0) >> MODSW_SHIFT);
    uint32_t* acc = sm.acc(g, q);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const int n = 32 * a + lane;
        uint32_t v = 0;
        if (q == 1) {
            const uint32_t m = ((uint32_t)n - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        accr[a] = v;
        acc[n] = v;
    }
}

// F0: (X^abar - 1) * acc + decomposition offsets, coefficient 32a + lane (utils.hpp:130-144, trgsw.hpp:62-78)
template <int G>
B200_HD void brg_rotate_diff(const BrgSmem<G>& sm, int i, int g, int q, int lane, const uint32_t (&accr)[32],
                             uint32_t (&dreg)[32])
{
    const uint32_t abar = sm.abar[g * 640 + i];
    const uint32_t* acc = sm.acc(g, q);
    const uint32_t base = ((uint32_t)lane - abar) & (2u * N1 - 1);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const uint32_t m = (base + 32u * a) & (2u * N1 - 1);
        const uint32_t v = acc[m & (N1 - 1)];
        const uint32_t neg = 0u - ((m >> NBIT) & 1u);
        dreg[a] = ((v ^ neg) - neg) - accr[a] + (DEC_OFFSET + DEC_ROUND);
    }
}
// Fa(d): digit d -> forward pass 1 -> column store
template <int G>
B200_HD void brg_fwd_a(const BrgSmem<G>& sm, int g, int q, int lane, int d, const uint32_t (&dreg)[32])
{
    uint32_t x[32];
    const int sh = 32 - (d + 1) * BGBIT;
    B200_UNROLL
    for (int a = 0; a < 32; a++) x[a] = ((dreg[a] >> sh) & ((1u << BGBIT) - 1)) + (P - (1u << (BGBIT - 1)));
    fwd_pass1(x);
    stile_store_col(sm.tile(g, q * GL + d), x, lane);
}
// Fb(d): row load -> forward pass 2 -> row store (values < 4p)
template <int G>
B200_HD void brg_fwd_b(const BrgSmem<G>& sm, int g, int q, int lane, int d)
{
    uint32_t x[32];
    uint32_t* t = sm.tile(g, q * GL + d);
    stile_load_row(t, x, lane);
    fwd_pass2(x, sm.tw2f, lane);
    stile_store_row(t, x, lane);
}

// pointwise stage: out[g][c][j] = REDC(sum_r D[g][r][j] * BK[c][r][j]), in place over the job's tiles
template <int G>
B200_HD void brg_pw_compute(const BrgSmem<G>& sm, int j, const uint32_t (&bkv)[BK_COLS][ROWS])
{
    const int off = stile_of_j(j);
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
template <int G>
B200_HD void brg_pointwise(const BrgSmem<G>& sm, const uint32_t* bk_i, int tid)
{
    constexpr int T = 64 * G;
    for (int j = tid; j < N1; j += T) {
        uint32_t bkv[BK_COLS][ROWS];
        pw_load(bk_i, j, bkv);
        brg_pw_compute<G>(sm, j, bkv);
    }
}

// Ia(l): row load -> inverse pass 1 -> row store
template <int G>
B200_HD void brg_inv_a(const BrgSmem<G>& sm, int g, int q, int lane, int l)
{
    uint32_t x[32];
    uint32_t* t = sm.tile(g, q * LIMBS + l);
    stile_load_row(t, x, lane);
    inv_pass1(x, sm.tw2i, lane);
    stile_store_row(t, x, lane);
}
// Ib(l): column load -> inverse pass 2 -> centred lift -> sum += v << shift(l)
template <int G>
B200_HD void brg_inv_b(const BrgSmem<G>& sm, int g, int q, int lane, int l, uint32_t (&sum)[32])
{
    uint32_t x[32];
    stile_load_col(sm.tile(g, q * LIMBS + l), x, lane);
    inv_pass2(x);
    const int sh = limb_shift(l);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const uint32_t v = (uint32_t)centered_lift(x[a]);
        sum[a] = (l == 0) ? v : sum[a] + (v << sh);
    }
}
// Ic: acc += external product; refresh the natural-order copy read by the next rotated difference
template <int G>
B200_HD void brg_acc_update(const BrgSmem<G>& sm, int g, int q, int lane, const uint32_t (&sum)[32], uint32_t (&accr)[32])
{
    uint32_t* acc = sm.acc(g, q);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        accr[a] += sum[a];
        acc[32 * a + lane] = accr[a];
    }
}

// epilogue: SampleExtractIndex(0) (trlwe.hpp:213-223)
template <int G>
B200_HD void brg_epilogue(const BrgSmem<G>& sm, int g, int q, int lane, uint32_t* u_out)
{
    const uint32_t* acc = sm.acc(g, q);
    if (q == 0) {
        for (int j = lane; j < N1; j += 32) u_out[j] = (j == 0) ? acc[0] : 0u - acc[N1 - j];
    } else if (lane == 0) {
        u_out[N1] = acc[0];
    }
}

// ---- bootstrapping-key precomputation, generic limb split -------------------------------------------------
// raw = sum_l x_l * 2^shift(l) (mod 2^32) with centred x_l of width(l) bits (the last limb absorbs the carry)
B200_HD int32_t brg_bk_limb(uint32_t raw, int limb)
{
    uint32_t v = raw;
    int32_t x = 0;
    B200_UNROLL
    for (int l = 0; l < LIMBS; l++) {
        const int w = limb_width(l);
        if (l == LIMBS - 1) {
            x = (int32_t)v;  // what is left, sign included
        } else {
            const uint32_t half = 1u << (w - 1), mask = (1u << w) - 1;
            x = (int32_t)((v & mask) ^ half) - (int32_t)half;
            v = (uint32_t)((int32_t)(v - (uint32_t)x) >> w);
        }
        if (l == limb) return x;
    }
    return x;
}
B200_HD void brg_bk_prep_a(const uint32_t* raw_poly, int limb, int lane, uint32_t* tile)
{
    uint32_t x[32];
    B200_UNROLL
    for (int a = 0; a < 32; a++) x[a] = (uint32_t)(brg_bk_limb(raw_poly[32 * a + lane], limb) + (int32_t)P);  // in (0, 2p)
    fwd_pass1(x);
    tile_store_col(tile, x, lane);
}

}  // namespace b200
