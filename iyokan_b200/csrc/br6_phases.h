// Blind rotation, fine-grained cluster shape (br6_kernel): ONE rotation job per 2-CTA cluster,
// 384 threads per CTA = 3 teams of 128 threads x 8 points (ntt_block8.h).
//
// Reference functions: TFHEpp gatebootstrapping.hpp:19-71, detwfa.hpp:36-49, trgsw.hpp:62-131,
// trlwe.hpp:213-223.  CTA q of the cluster owns accumulator polynomial q: it transforms the three digits
// of its polynomial, copies each finished digit tile to the peer CTA with a bulk-async DSMEM copy, stages
// its half of the step's NTT-domain key (73,728 B) by TMA, and inverts the three limbs of its polynomial.  It exists for the latency-bound levels of processor netlists, where a clock cycle
// is the SUM of single-rotation times.
#pragma once
#include "br4_phases.h"
#include "ntt_block8.h"

namespace b200 {

constexpr int BR6_THREADS = GL * TEAM8_THREADS;              // 384
constexpr int BR6_KEY_WORDS = LIMBS * ROWS * N1;             // 18,432 words = 73,728 B: the key columns of one polynomial

struct Br6Smem {
    static constexpr size_t BYTES = (size_t)BR6_KEY_WORDS * 4 + (size_t)(ROWS + LIMBS) * B8_WORDS * 4 + (size_t)N1 * 4 +
                                    sizeof(Block8Tw) + (size_t)SLOT_STRIDE * 2 + 16;
    uint32_t* keyb;   // [LIMBS][ROWS][1024] key columns of this CTA's polynomial
    uint32_t* din;    // [ROWS][B8_WORDS]: rows 3q..3q+2 computed here, the other three copied in by the peer
    uint32_t* dout;   // [LIMBS][B8_WORDS]
    uint32_t* accb;   // [1024]
    Block8Tw* tw;
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
        tw = reinterpret_cast<Block8Tw*>(p);
        p += sizeof(Block8Tw);
        abar = reinterpret_cast<uint16_t*>(p);
        p += (size_t)SLOT_STRIDE * 2;
        mbar = reinterpret_cast<uint64_t*>(p);
    }
    B200_HD uint32_t* in_tile(int r) const { return din + (size_t)r * B8_WORDS; }
    B200_HD uint32_t* out_tile(int l) const { return dout + (size_t)l * B8_WORDS; }
};
static_assert(sizeof(Block8Tw) % 16 == 0, "table block keeps 16-byte alignment");

B200_HD void br6_prologue(const Br6Smem& sm, const BrJob& job, const uint16_t* arena, int q, int tid)
{
    for (int i = tid; i < N0; i += BR6_THREADS) {
        const uint32_t c = br_lincomb(job, arena, i);
        sm.abar[i] = (uint16_t)((c + 16u) >> 5);
    }
    const uint32_t bbar = 2u * N1 - (br_lincomb(job, arena, N0) >> 5);
    for (int n = tid; n < N1; n += BR6_THREADS) {
        uint32_t v = 0;
        if (q == 1) {
            const uint32_t m = ((uint32_t)n - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        sm.accb[n] = v;
    }
}

// digit d of (X^abar - 1) * acc_q (utils.hpp:130-144, trgsw.hpp:62-78), stages 0..2
B200_HD void br6_fwd_p1(const Br6Smem& sm, int i, int q, int d, int t)
{
    const uint32_t abar = sm.abar[i];
    const uint32_t* acc = sm.accb;
    const uint32_t base = ((uint32_t)t - abar) & (2u * N1 - 1);
    uint32_t dv[8], x[8];
    B200_UNROLL
    for (int a = 0; a < 8; a++) {
        const uint32_t m = (base + 128u * a) & (2u * N1 - 1);
        const uint32_t v = acc[m & (N1 - 1)];
        const uint32_t neg = 0u - ((m >> NBIT) & 1u);
        dv[a] = ((v ^ neg) - neg) - acc[128 * a + t] + (DEC_OFFSET + DEC_ROUND);
    }
    // the team's digit is uniform per warp: three instantiations of the table-driven start, one per bit field
    if (d == 0) blk8_fwd_p1_digits<32 - BGBIT>(sm.tw->r4, dv, x);
    else if (d == 1) blk8_fwd_p1_digits<32 - 2 * BGBIT>(sm.tw->r4, dv, x);
    else blk8_fwd_p1_digits<32 - 3 * BGBIT>(sm.tw->r4, dv, x);
    blk8_store_p1(sm.in_tile(q * GL + d), x, t);
}
B200_HD void br6_fwd_p2(const Br6Smem& sm, int q, int d, int t) { blk8_fwd_p2(sm.in_tile(q * GL + d), sm.tw->q2f, t); }
B200_HD void br6_fwd_p3(const Br6Smem& sm, int q, int d, int t) { blk8_fwd_p3(sm.in_tile(q * GL + d), sm.tw->q3f, t); }
B200_HD void br6_fwd_p4(const Br6Smem& sm, int q, int d, int t) { blk8_fwd_p4(sm.in_tile(q * GL + d), sm.tw->q4f, t); }

// Pointwise stage, split around the arrival of the peer's tiles (local rows first, the peer's rows once they have landed).
// A thread (tid < 256) owns quad m = tid of NTT positions for ALL three limb columns of the CTA's
// polynomial, so each digit quad is read once instead of three times: the stage is bound by
// shared-memory wavefronts (key 576 + digits 216 per step), not by the multiplies, and eight busy
// warps already saturate them.
constexpr int BR6_PW_THREADS = N1 / 4;  // 256
template <class SM>  // Br6Smem or Br9Smem: same tiles, same key layout
B200_HD void br6_pw_rows(const SM& sm, int m, int row0, uint64_t (&acc)[LIMBS][4])
{
    const int toff = b8_quad(m);
    B200_UNROLL
    for (int rr = 0; rr < GL; rr++) {
        const u32x4 dv = *reinterpret_cast<const u32x4*>(sm.in_tile(row0 + rr) + toff);
        B200_UNROLL
        for (int l = 0; l < LIMBS; l++) {
            const u32x4 kk = *reinterpret_cast<const u32x4*>(sm.keyb + (size_t)(l * ROWS + row0 + rr) * N1 + 4 * m);
            acc[l][0] += (uint64_t)dv.x * kk.x;
            acc[l][1] += (uint64_t)dv.y * kk.y;
            acc[l][2] += (uint64_t)dv.z * kk.z;
            acc[l][3] += (uint64_t)dv.w * kk.w;
        }
    }
}
template <class SM>
B200_HD void br6_pw_local(const SM& sm, int q, int tid, uint64_t (&acc)[LIMBS][4])
{
    if (tid >= BR6_PW_THREADS) return;
    B200_UNROLL
    for (int l = 0; l < LIMBS; l++) acc[l][0] = acc[l][1] = acc[l][2] = acc[l][3] = 0;
    br6_pw_rows(sm, tid, q * GL, acc);
}
template <class SM>
B200_HD void br6_pw_finish(const SM& sm, int q, int tid, uint64_t (&acc)[LIMBS][4])
{
    if (tid >= BR6_PW_THREADS) return;
    br6_pw_rows(sm, tid, (q ^ 1) * GL, acc);
    B200_UNROLL
    for (int l = 0; l < LIMBS; l++)
        *reinterpret_cast<u32x4*>(sm.out_tile(l) + b8_quad(tid)) =
            u32x4{redc64(acc[l][0]), redc64(acc[l][1]), redc64(acc[l][2]), redc64(acc[l][3])};
}

B200_HD void br6_inv_pA(const Br6Smem& sm, int l, int t) { blk8_inv_pA(sm.out_tile(l), sm.tw->q4i, t); }
B200_HD void br6_inv_pB(const Br6Smem& sm, int l, int t) { blk8_inv_pB(sm.out_tile(l), sm.tw->q3i, t); }
B200_HD void br6_inv_pC(const Br6Smem& sm, int l, int t) { blk8_inv_pC(sm.out_tile(l), sm.tw->q2i, t); }
B200_HD void br6_inv_pD(const Br6Smem& sm, int l, int t)
{
    uint32_t x[8];
    blk8_load_p1(sm.out_tile(l), x, t);
    blk8_inv_pD(x);
    B200_UNROLL
    for (int a = 0; a < 8; a++) {
        const uint32_t v = (uint32_t)centered_lift(x[a]) << (LIMB_BITS * l);
        B200_SMEM_ADD(sm.accb + 128 * a + t, v);
    }
}

B200_HD void br6_epilogue(const Br6Smem& sm, int q, int tid, uint32_t* u_out)
{
    if (q == 0) {
        for (int j = tid; j < N1; j += BR6_THREADS) u_out[j] = (j == 0) ? sm.accb[0] : 0u - sm.accb[N1 - j];
    } else if (tid == 0) {
        u_out[N1] = sm.accb[0];
    }
}

}  // namespace b200
