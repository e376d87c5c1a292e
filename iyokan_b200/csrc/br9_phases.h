// Blind rotation, cluster shape with 4-point threads (br9_kernel): ONE rotation job per 2-CTA cluster,
// 768 threads per CTA = 3 teams of 256 threads x 4 points (ntt_block4.h).
//
// Same protocol as br6_kernel (br6_phases.h: CTA q owns accumulator polynomial q, transforms its three digits, pushes
// the finished digit tiles into the peer CTA with bulk-async DSMEM copies, stages its half of the step's key by TMA,
// inverts its three limbs; reference functions TFHEpp gatebootstrapping.hpp:19-71, detwfa.hpp:36-49, trgsw.hpp:62-131,
// trlwe.hpp:213-223) with the transforms cut into five two-stage passes of 256 threads: 24 warps per SM instead of 12.
// The first forward pass is the table look-up on the digits (no multiplication at all); the pointwise stage and the tile
// layout are br6's.
#pragma once
#include "br6_phases.h"
#include "ntt_block4.h"

namespace b200 {

constexpr int BR9_THREADS = GL * TEAM4_THREADS;  // 768

struct Br9Smem {
    static constexpr size_t BYTES = (size_t)BR6_KEY_WORDS * 4 + (size_t)(ROWS + LIMBS) * B8_WORDS * 4 + (size_t)N1 * 4 +
                                    sizeof(Block4Tw) + (size_t)SLOT_STRIDE * 2 + 16;
    uint32_t* keyb;   // [LIMBS][ROWS][1024] key columns of this CTA's polynomial
    uint32_t* din;    // [ROWS][B8_WORDS]: rows 3q..3q+2 computed here, the other three copied in by the peer
    uint32_t* dout;   // [LIMBS][B8_WORDS]
    uint32_t* accb;   // [1024]
    Block4Tw* tw;
    uint16_t* abar;
    uint64_t* mbar;   // [2]: key stage, incoming digit tiles
    B200_HD void carve(void* base)
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(base);
        keyb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)BR6_KEY_WORDS * 4;
        din = reinterpret_cast<uint32_t*>(p);
        p += (size_t)ROWS * B8_WORDS * 4;
        dout = reinterpret_cast<uint32_t*>(p);
        p += (size_t)LIMBS * B8_WORDS * 4;
        accb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)N1 * 4;
        tw = reinterpret_cast<Block4Tw*>(p);
        p += sizeof(Block4Tw);
        abar = reinterpret_cast<uint16_t*>(p);
        p += (size_t)SLOT_STRIDE * 2;
        mbar = reinterpret_cast<uint64_t*>(p);
    }
    B200_HD uint32_t* in_tile(int r) const { return din + (size_t)r * B8_WORDS; }
    B200_HD uint32_t* out_tile(int l) const { return dout + (size_t)l * B8_WORDS; }
};
static_assert(Br9Smem::BYTES <= 227 * 1024, "one CTA per SM must fit the opt-in shared memory limit");

B200_HD void br9_prologue(const Br9Smem& sm, const BrJob& job, const uint16_t* arena, int q, int tid)
{
    for (int i = tid; i < N0; i += BR9_THREADS) {
        const uint32_t c = br_lincomb(job, arena, i);
        sm.abar[i] = (uint16_t)((c + 16u) >> 5);
    }
    const uint32_t bbar = 2u * N1 - (br_lincomb(job, arena, N0) >> 5);
    for (int n = tid; n < N1; n += BR9_THREADS) {
        uint32_t v = 0;
        if (q == 1) {
            const uint32_t m = ((uint32_t)n - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        sm.accb[n] = v;
    }
}

// digit d of (X^abar - 1) * acc_q (utils.hpp:130-144, trgsw.hpp:62-78), stages 0 and 1 by table look-up
B200_HD void br9_fwd_p1(const Br9Smem& sm, int i, int q, int d, int t)
{
    const uint32_t abar = sm.abar[i];
    const uint32_t* acc = sm.accb;
    const uint32_t base = ((uint32_t)t - abar) & (2u * N1 - 1);
    uint32_t dv[4];
    B200_UNROLL
    for (int e = 0; e < 4; e++) {
        const uint32_t m = (base + 256u * e) & (2u * N1 - 1);
        const uint32_t v = acc[m & (N1 - 1)];
        const uint32_t neg = 0u - ((m >> NBIT) & 1u);
        dv[e] = ((v ^ neg) - neg) - acc[256 * e + t] + (DEC_OFFSET + DEC_ROUND);
    }
    uint32_t* tile = sm.in_tile(q * GL + d);
    // the team's digit is uniform per warp: one instantiation of the table-driven start per bit field
    if (d == 0) blk4_fwd_p1_digits<32 - BGBIT>(tile, sm.tw->b8.r4, dv, t);
    else if (d == 1) blk4_fwd_p1_digits<32 - 2 * BGBIT>(tile, sm.tw->b8.r4, dv, t);
    else blk4_fwd_p1_digits<32 - 3 * BGBIT>(tile, sm.tw->b8.r4, dv, t);
}
B200_HD void br9_fwd_p2(const Br9Smem& sm, int q, int d, int t) { blk4_fwd_p2(sm.in_tile(q * GL + d), t); }
B200_HD void br9_fwd_p3(const Br9Smem& sm, int q, int d, int t) { blk4_fwd_p3(sm.in_tile(q * GL + d), sm.tw->p3f, t); }
B200_HD void br9_fwd_p4(const Br9Smem& sm, int q, int d, int t) { blk4_fwd_p4(sm.in_tile(q * GL + d), sm.tw->b8.q3f, t); }
B200_HD void br9_fwd_p5(const Br9Smem& sm, int q, int d, int t) { blk4_fwd_p5(sm.in_tile(q * GL + d), sm.tw->b8.q4f, t); }

B200_HD void br9_inv_pA(const Br9Smem& sm, int l, int t) { blk4_inv_pA(sm.out_tile(l), sm.tw->b8.q4i, t); }
B200_HD void br9_inv_pB(const Br9Smem& sm, int l, int t) { blk4_inv_pB(sm.out_tile(l), sm.tw->b8.q3i, t); }
B200_HD void br9_inv_pC(const Br9Smem& sm, int l, int t) { blk4_inv_pC(sm.out_tile(l), sm.tw->p3i, t); }
B200_HD void br9_inv_pD(const Br9Smem& sm, int l, int t) { blk4_inv_pD(sm.out_tile(l), t); }
B200_HD void br9_inv_pE(const Br9Smem& sm, int l, int t)
{
    uint32_t x[4];
    blk4_inv_pE(sm.out_tile(l), x, t);
    B200_UNROLL
    for (int e = 0; e < 4; e++) {
        const uint32_t v = (uint32_t)centered_lift(x[e]) << (LIMB_BITS * l);
        B200_SMEM_ADD(sm.accb + 256 * e + t, v);
    }
}

B200_HD void br9_epilogue(const Br9Smem& sm, int q, int tid, uint32_t* u_out)
{
    if (q == 0) {
        for (int j = tid; j < N1; j += BR9_THREADS) u_out[j] = (j == 0) ? sm.accb[0] : 0u - sm.accb[N1 - j];
    } else if (tid == 0) {
        u_out[N1] = sm.accb[0];
    }
}

}  // namespace b200
