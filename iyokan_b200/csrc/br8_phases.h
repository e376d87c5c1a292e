// Blind rotation, quad-cluster shape (br8_kernel): ONE rotation job per 4-CTA thread-block cluster.
//
// Why: the narrow dependency levels of a processor netlist (a few dozen gates) are latency bound - a clock cycle is
// the SUM of single-rotation times.  br6_kernel gives a job two SMs (one accumulator
// polynomial each); this shape gives it four by also cutting every transform in two:  after stage 0 of the merged
// negacyclic Cooley-Tukey transform the positions [0,512) and [512,1024) never meet again, and the Gentleman-Sande
// inverse joins them only in its last stage.  CTA (q, h) of the cluster (rank 2q + h) owns polynomial q, half h:
//   F   rotated difference + digit d of all 1024 coefficients (both CTAs of a polynomial hold a full accumulator
//       copy), stage 0, then stages 1..9 on ITS 512 positions only           3 teams of 64 threads x 8 points
//   A   its three digit half-tiles -> CTA (1-q, h), which needs them for the pointwise stage        (DSMEM, 6.9 KB)
//   M   pointwise stage for the limb columns of polynomial q at its 512 positions (6 rows x 3 columns, key
//       quarter = 36,864 B staged by TMA)
//   I   inverse stages 9..1 of the three limb columns on its half
//   B   the three half results -> CTA (q, 1-h); stage 0 of the inverse joins the halves: CTA (q,0) forms U + V
//       (coefficients 0..511), CTA (q,1) forms (U - V) w (coefficients 512..1023)                  (DSMEM, 6.9 KB)
//   C   lift, recombine, accumulate its 512 coefficients; the updated half -> CTA (q, 1-h)          (DSMEM, 2 KB)
// Every exchange is a stream of st.async remote stores (each performs complete_tx on an mbarrier of the receiving CTA),
// issued by the threads that produce the words, in the pass that produces them.  (First version: one bulk-async DSMEM
// copy per tile after the pass; measured 2.22 ms per rotation against 2.06 ms for br6_kernel - three exposed copy
// latencies per step cost more than halving the arithmetic saved, profiles/r02_br8.md.)  No cluster barrier inside
// the loop: the data dependences order everything (the sender of C(i) has consumed B(i); B(i+1) and C(i+1) cannot be
// produced before C(i) has arrived), except the digit tiles, which are double buffered.
// Reference functions: TFHEpp gatebootstrapping.hpp:19-71, detwfa.hpp:36-49, trgsw.hpp:62-131, trlwe.hpp:213-223.
//
// Half transform, 64 threads x 8 points (position j = 512h + jl, jl local):
//   pass 1  stage 0 on 16 inputs (j = 64a + t, a and a + 8), keep the 8 of half h, stages 1..3 (strides 256, 128, 64)
//   pass 2  stages 4..6: thread (A, c) = (t >> 3, t & 7) holds jl = 64A + 8e + c
//   pass 3  stages 7..9: thread t holds jl = 8t + e
// Stage s uses psi_rev[2^s + (j >> (10 - s))]; the whole forward and inverse tables (2 x 8 KB) sit in shared memory.
// Half tile: word hp(jl) = jl + (jl >> 3) (576 words): pass 2 and pass 3 accesses are bank-conflict free.
#pragma once
#include "br4_phases.h"

// Peer delivery: on the device every result word is also sent to the peer CTA with st.async (a remote shared-memory
// store that performs complete_tx on the peer's mbarrier), fused into the pass that produces it; the simulator delivers
// whole buffers with memcpy at the same points.  `Br8Peer` carries the shared::cluster addresses (0 = no delivery).
#if defined(__CUDA_ARCH__)
#define B200_ST_ASYNC(dst, val, bar) \
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst), "r"(val), "r"(bar) : "memory")
#else
#define B200_ST_ASYNC(dst, val, bar) ((void)(dst), (void)(val), (void)(bar))
#endif

namespace b200 {

struct Br8Peer {
    uint32_t dst = 0;  // shared::cluster address of the peer's copy of the tile being written (word 0)
    uint32_t bar = 0;  // shared::cluster address of the peer's mbarrier
};

constexpr int BR8_TEAM = 64;
constexpr int BR8_THREADS = GL * BR8_TEAM;          // 192
constexpr int H_WORDS = 512 + 64;                   // padded half tile, 2304 B
constexpr int BR8_KEY_WORDS = LIMBS * ROWS * 512;   // 9216 words = 36,864 B: (3 limb columns) x (6 rows) x (512 positions)
B200_HD int hp(int jl) { return jl + (jl >> 3); }

struct Br8Smem {
    static constexpr size_t BYTES = (size_t)BR8_KEY_WORDS * 4 + (size_t)GL * H_WORDS * 4 * 4 + (size_t)LIMBS * H_WORDS * 4 * 2 +
                                    (size_t)N1 * 4 + 2 * 1024 * sizeof(tw_t) + 640 * 2 + 64;
    uint32_t* keyb;   // [LIMBS][ROWS][512]
    uint32_t* dig;    // [2][GL][H_WORDS] own digit half tiles (double buffered: the copy engine may still read step i's)
    uint32_t* peer;   // [2][GL][H_WORDS] digit half tiles of polynomial 1-q (double buffered by step parity)
    uint32_t* outb;   // [LIMBS][H_WORDS] pointwise results / inverse in place
    uint32_t* half2;  // [LIMBS][H_WORDS] the other half's inverse results (from CTA (q, 1-h))
    uint32_t* accb;   // [1024] accumulator polynomial q, natural order; own half updated here, the other half copied in
    tw_t* twf;        // [1024] psi_rev
    tw_t* twi;        // [1024] inverse
    uint16_t* abar;
    uint64_t* mbar;   // [0] key, [1..2] digit tiles (by parity), [3] inverse halves, [4] accumulator half
    B200_HD void carve(void* base)
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(base);
        keyb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)BR8_KEY_WORDS * 4;
        dig = reinterpret_cast<uint32_t*>(p);
        p += (size_t)2 * GL * H_WORDS * 4;
        peer = reinterpret_cast<uint32_t*>(p);
        p += (size_t)2 * GL * H_WORDS * 4;
        outb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)LIMBS * H_WORDS * 4;
        half2 = reinterpret_cast<uint32_t*>(p);
        p += (size_t)LIMBS * H_WORDS * 4;
        accb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)N1 * 4;
        twf = reinterpret_cast<tw_t*>(p);
        p += 1024 * sizeof(tw_t);
        twi = reinterpret_cast<tw_t*>(p);
        p += 1024 * sizeof(tw_t);
        abar = reinterpret_cast<uint16_t*>(p);
        p += 640 * 2;
        mbar = reinterpret_cast<uint64_t*>(p);
    }
};
static_assert((H_WORDS * 4) % 16 == 0 && (BR8_KEY_WORDS * 4) % 16 == 0, "bulk copies move multiples of 16 bytes");

B200_HD void br8_prologue(const Br8Smem& sm, const BrJob& job, const uint16_t* arena, int q, int tid)
{
    for (int i = tid; i < N0; i += BR8_THREADS) {
        const uint32_t c = br_lincomb(job, arena, i);
        sm.abar[i] = (uint16_t)((c + 16u) >> 5);
    }
    const uint32_t bbar = 2u * N1 - (br_lincomb(job, arena, N0) >> 5);
    for (int n = tid; n < N1; n += BR8_THREADS) {
        uint32_t v = 0;
        if (q == 1) {
            const uint32_t m = ((uint32_t)n - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        sm.accb[n] = v;
    }
}

// pass 1: digit d of (X^abar - 1) * acc_q at j = 64a + t (utils.hpp:130-144, trgsw.hpp:62-78), stage 0 across the
// halves, stages 1..3 on half h, store
B200_HD void br8_fwd_p1(const Br8Smem& sm, int i, int h, int d, int t)
{
    const uint32_t abar = sm.abar[i];
    const uint32_t* acc = sm.accb;
    const uint32_t base = ((uint32_t)t - abar) & (2u * N1 - 1);
    const int sh = 32 - (d + 1) * BGBIT;
    uint32_t x[16];
    B200_UNROLL
    for (int a = 0; a < 16; a++) {
        const uint32_t m = (base + 64u * a) & (2u * N1 - 1);
        const uint32_t v = acc[m & (N1 - 1)];
        const uint32_t neg = 0u - ((m >> NBIT) & 1u);
        const uint32_t diff = ((v ^ neg) - neg) - acc[64 * a + t] + (DEC_OFFSET + DEC_ROUND);
        x[a] = ((diff >> sh) & ((1u << BGBIT) - 1)) + (P - (1u << (BGBIT - 1)));
    }
    // stage 0 (single twiddle psi_rev[1]); only the outputs of half h are kept
    uint32_t y[8];
    const tw_t w0 = twf_u(1);
    B200_UNROLL
    for (int a = 0; a < 8; a++) {
        const uint32_t T = shoup_mul(x[a + 8], w0);
        y[a] = h == 0 ? x[a] + T : x[a] - T + P2;
    }
    ct_stage_n<8, 0, 0>(y, [=](int) { return twf_u(2 + h); });
    ct_stage_n<8, 1, 0>(y, [=](int g) { return twf_u(4 + 2 * h + g); });
    ct_stage_n<8, 2, 1>(y, [=](int g) { return twf_u(8 + 4 * h + g); });
    uint32_t* tile = sm.dig + (size_t)((i & 1) * GL + d) * H_WORDS;
    B200_UNROLL
    for (int a = 0; a < 8; a++) tile[hp(64 * a + t)] = y[a];
}
// pass 2: stages 4..6
B200_HD void br8_fwd_p2(uint32_t* tile, const tw_t* twf, int h, int t)
{
    const int A = t >> 3, c = t & 7, Ag = 8 * h + A;
    uint32_t x[8];
    B200_UNROLL
    for (int e = 0; e < 8; e++) x[e] = tile[hp(64 * A + 8 * e + c)];
    ct_stage_n<8, 0, 0>(x, [=](int) { return twf[16 + Ag]; });
    ct_stage_n<8, 1, 1>(x, [=](int g) { return twf[32 + 2 * Ag + g]; });
    ct_stage_n<8, 2, 0>(x, [=](int g) { return twf[64 + 4 * Ag + g]; });
    B200_UNROLL
    for (int e = 0; e < 8; e++) tile[hp(64 * A + 8 * e + c)] = x[e];
}
// pass 3: stages 7..9, output < 4p
B200_HD void br8_fwd_p3(uint32_t* tile, const tw_t* twf, int h, int t, Br8Peer peer = Br8Peer())
{
    const int mg = 64 * h + t;
    uint32_t x[8];
    B200_UNROLL
    for (int e = 0; e < 8; e++) x[e] = tile[9 * t + e];  // hp(8t + e)
    ct_stage_n<8, 0, 1>(x, [=](int) { return twf[128 + mg]; });
    ct_stage_n<8, 1, 0>(x, [=](int g) { return twf[256 + 2 * mg + g]; });
    ct_stage_n<8, 2, 2>(x, [=](int g) { return twf[512 + 4 * mg + g]; });
    B200_UNROLL
    for (int e = 0; e < 8; e++) {
        tile[9 * t + e] = x[e];
        if (peer.dst) B200_ST_ASYNC(peer.dst + 4u * (uint32_t)(9 * t + e), x[e], peer.bar);
    }
}

// pointwise stage over the 512 positions of this CTA: out[l][jl] = REDC(sum_r D[r][jl] * key[l][r][jl]);
// rows of polynomial q are local (sm.dig), rows of polynomial 1-q arrived in sm.peer[parity]
B200_HD void br8_pointwise(const Br8Smem& sm, int q, int parity, int tid)
{
    const uint32_t* own = sm.dig + (size_t)parity * GL * H_WORDS;
    const uint32_t* oth = sm.peer + (size_t)parity * GL * H_WORDS;
    for (int jl = tid; jl < 512; jl += BR8_THREADS) {
        const int off = hp(jl);
        uint32_t dv[ROWS];
        B200_UNROLL
        for (int d = 0; d < GL; d++) {
            dv[q * GL + d] = own[d * H_WORDS + off];
            dv[(q ^ 1) * GL + d] = oth[d * H_WORDS + off];
        }
        B200_UNROLL
        for (int l = 0; l < LIMBS; l++) {
            uint64_t acc = 0;
            B200_UNROLL
            for (int r = 0; r < ROWS; r++) acc += (uint64_t)dv[r] * sm.keyb[(l * ROWS + r) * 512 + jl];
            sm.outb[l * H_WORDS + off] = redc64(acc);
        }
    }
}

// inverse passes on half h of limb column l (Gentleman-Sande, every stage folds the sum below 4p)
B200_HD void br8_inv_pA(uint32_t* tile, const tw_t* twi, int h, int t)  // stages 9..7
{
    const int mg = 64 * h + t;
    uint32_t x[8];
    B200_UNROLL
    for (int e = 0; e < 8; e++) x[e] = tile[9 * t + e];
    gs_stage_n<8, 2, 1>(x, [=](int g) { return twi[512 + 4 * mg + g]; });
    gs_stage_n<8, 1, 1>(x, [=](int g) { return twi[256 + 2 * mg + g]; });
    gs_stage_n<8, 0, 1>(x, [=](int) { return twi[128 + mg]; });
    B200_UNROLL
    for (int e = 0; e < 8; e++) tile[9 * t + e] = x[e];
}
B200_HD void br8_inv_pB(uint32_t* tile, const tw_t* twi, int h, int t)  // stages 6..4
{
    const int A = t >> 3, c = t & 7, Ag = 8 * h + A;
    uint32_t x[8];
    B200_UNROLL
    for (int e = 0; e < 8; e++) x[e] = tile[hp(64 * A + 8 * e + c)];
    gs_stage_n<8, 2, 1>(x, [=](int g) { return twi[64 + 4 * Ag + g]; });
    gs_stage_n<8, 1, 1>(x, [=](int g) { return twi[32 + 2 * Ag + g]; });
    gs_stage_n<8, 0, 1>(x, [=](int) { return twi[16 + Ag]; });
    B200_UNROLL
    for (int e = 0; e < 8; e++) tile[hp(64 * A + 8 * e + c)] = x[e];
}
B200_HD void br8_inv_pC(uint32_t* tile, int h, int t, Br8Peer peer = Br8Peer())  // stages 3..1, result also sent to CTA (q, 1-h)
{
    uint32_t x[8];
    B200_UNROLL
    for (int a = 0; a < 8; a++) x[a] = tile[hp(64 * a + t)];
    gs_stage_n<8, 2, 1>(x, [=](int g) { return twi_u(8 + 4 * h + g); });
    gs_stage_n<8, 1, 1>(x, [=](int g) { return twi_u(4 + 2 * h + g); });
    gs_stage_n<8, 0, 1>(x, [=](int) { return twi_u(2 + h); });
    B200_UNROLL
    for (int a = 0; a < 8; a++) {
        tile[hp(64 * a + t)] = x[a];
        if (peer.dst) B200_ST_ASYNC(peer.dst + 4u * (uint32_t)hp(64 * a + t), x[a], peer.bar);
    }
}
// this CTA's updated accumulator half -> the same place in CTA (q, 1-h)'s copy
B200_HD void br8_send_acc(const Br8Smem& sm, int h, int tid, Br8Peer peer)
{
    for (int n = tid; n < 512; n += BR8_THREADS)
        if (peer.dst) B200_ST_ASYNC(peer.dst + 4u * (uint32_t)n, sm.accb[512 * h + n], peer.bar);
}
// last inverse stage across the halves + lift + recombination into this CTA's 512 accumulator coefficients:
// own = this half's values (limb l), other = the other half's values copied in by CTA (q, 1-h)
B200_HD void br8_inv_join(const Br8Smem& sm, int h, int l, int t)
{
    const uint32_t* own = sm.outb + l * H_WORDS;
    const uint32_t* other = sm.half2 + l * H_WORDS;
    const tw_t w0 = twi_u(1);
    B200_UNROLL
    for (int a = 0; a < 8; a++) {
        const int off = hp(64 * a + t);
        const uint32_t U = h == 0 ? own[off] : other[off], V = h == 0 ? other[off] : own[off];
        const uint32_t r = h == 0 ? fix_lt8p_to_lt4p(U + V) : shoup_mul(U - V + P4, w0);
        const uint32_t v = (uint32_t)centered_lift(r) << (LIMB_BITS * l);
        B200_SMEM_ADD(sm.accb + 512 * h + 64 * a + t, v);
    }
}

B200_HD void br8_epilogue(const Br8Smem& sm, int q, int h, int tid, uint32_t* u_out)
{
    if (h != 0) return;  // both halves of the pair hold the same accumulator copy
    if (q == 0) {
        for (int j = tid; j < N1; j += BR8_THREADS) u_out[j] = (j == 0) ? sm.accb[0] : 0u - sm.accb[N1 - j];
    } else if (tid == 0) {
        u_out[N1] = sm.accb[0];
    }
}

}  // namespace b200
