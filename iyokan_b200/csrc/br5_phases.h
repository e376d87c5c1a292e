// Blind rotation, cluster shape (br5_kernel): ONE rotation job per 2-CTA thread-block cluster.
//
// Same arithmetic and reference functions as br_phases.h / br4_phases.h (TFHEpp
// gatebootstrapping.hpp:19-71, detwfa.hpp:36-49, trgsw.hpp:62-131, trlwe.hpp:213-223); bit-identical
// results.  Used for dependency levels narrow enough to give every job two SMs (<= 74 jobs on a
// 148-SM B200): CTA q of the cluster owns accumulator polynomial q (0 = A, 1 = B):
//   * its three 64-thread teams decompose acc_q and transform the three digit polynomials
//     (ntt_block.h) into the CTA's own tiles; each finished tile (4,352 B) is then copied into the
//     peer CTA's shared memory by ONE bulk-async copy (cp.async.bulk shared::cta -> shared::cluster)
//     that signals an mbarrier in the peer, so after the forward transforms both CTAs hold all six
//     digit polynomials locally without any cluster-scope fence;
//   * each CTA forms only the three output columns of ITS polynomial: pointwise stage against its
//     half of the step's key (73,728 B staged by one bulk-async copy from global memory), first over
//     its own three rows, then - once the mbarrier says the peer's tiles have landed - over the
//     other three; it inverse-transforms the three limbs and accumulates them into acc_q, which
//     never leaves the CTA.
// The only cluster barrier per CMUX step is a relaxed arrive after the pointwise stage ("I no
// longer read the tiles you copied in") waited just before the next copies are issued, i.e. hidden
// behind the inverse transforms.  (Measured alternatives: remote stores + release/acquire cluster
// barrier 2.5 ms per rotation, remote loads 3.2 ms; the single-CTA shape is 3.0 ms.)
// Every function is free of intra-phase cross-thread communication (see br_phases.h): the peer
// tile pointer is plain memory to the lock-step CPU simulator.
#pragma once
#include "br4_phases.h"

namespace b200 {

constexpr int BR5_THREADS = GL * TEAM_THREADS;      // 192
constexpr int BR5_KEY_WORDS = LIMBS * ROWS * N1;    // 18,432 words = 73,728 B: the columns of one polynomial
constexpr int BR5_PW_ITEMS = (N1 / 4) * LIMBS;      // (quad of positions, limb column) = 768 = 4 per thread

struct Br5Smem {
    static constexpr size_t BYTES = (size_t)BR5_KEY_WORDS * 4 + (size_t)(ROWS + LIMBS) * BT_WORDS * 4 + (size_t)N1 * 4 +
                                    sizeof(BlockTw) + (size_t)SLOT_STRIDE * 2 + 16;
    uint32_t* keyb;   // [LIMBS][ROWS][1024] key columns of this CTA's polynomial
    uint32_t* din;    // [ROWS][BT_WORDS]: rows 3q..3q+2 computed here, the other three copied in by the peer
    uint32_t* dout;   // [LIMBS][BT_WORDS]
    uint32_t* accb;   // [1024] accumulator polynomial q, natural order
    BlockTw* tw;
    uint16_t* abar;
    uint64_t* mbar;
    B200_HD void carve(void* base)
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(base);
        keyb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)BR5_KEY_WORDS * 4;
        din = reinterpret_cast<uint32_t*>(p);
        p += (size_t)ROWS * BT_WORDS * 4;
        dout = reinterpret_cast<uint32_t*>(p);
        p += (size_t)LIMBS * BT_WORDS * 4;
        accb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)N1 * 4;
        tw = reinterpret_cast<BlockTw*>(p);
        p += sizeof(BlockTw);
        abar = reinterpret_cast<uint16_t*>(p);
        p += (size_t)SLOT_STRIDE * 2;
        mbar = reinterpret_cast<uint64_t*>(p);
    }
    B200_HD uint32_t* in_tile(int r) const { return din + (size_t)r * BT_WORDS; }
    B200_HD uint32_t* out_tile(int l) const { return dout + (size_t)l * BT_WORDS; }
};

B200_HD void br5_prologue(const Br5Smem& sm, const BrJob& job, const uint16_t* arena, int q, int tid)
{
    for (int i = tid; i < N0; i += BR5_THREADS) {
        const uint32_t c = br_lincomb(job, arena, i);
        sm.abar[i] = (uint16_t)((c + 16u) >> 5);
    }
    const uint32_t bbar = 2u * N1 - (br_lincomb(job, arena, N0) >> 5);
    for (int n = tid; n < N1; n += BR5_THREADS) {
        uint32_t v = 0;
        if (q == 1) {
            const uint32_t m = ((uint32_t)n - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        sm.accb[n] = v;
    }
}

B200_HD void br5_fwd_p1(const Br5Smem& sm, int i, int q, int d, int t)
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
    blk_fwd_p1(x);
    blk_store_p1(sm.in_tile(q * GL + d), x, t);
}
B200_HD void br5_fwd_p2(const Br5Smem& sm, int q, int d, int t)
{
    tw_t w[15];
    blk_load_tw2(sm.tw->p2f, t, w);
    blk_fwd_p2(sm.in_tile(q * GL + d), w, t);
}
B200_HD void br5_fwd_p3(const Br5Smem& sm, int q, int d, int t) { blk_fwd_p3(sm.in_tile(q * GL + d), sm.tw->p3f, t); }

// Pointwise stage, split so that the cluster barrier hides behind arithmetic: a thread owns
// BR5_PW_PER_THREAD items = (quad m of positions, limb column l of this CTA's polynomial).
//   part 1 (before the barrier is waited): rows computed by THIS CTA (3q..3q+2)
//   part 2 (after): rows pushed by the peer, Montgomery reduction, store
constexpr int BR5_PW_PER_THREAD = BR5_PW_ITEMS / BR5_THREADS;  // 4
// rows row0..row0+2 of the external product
B200_HD void br5_pw_rows(const Br5Smem& sm, int tid, int row0, uint64_t (&acc)[BR5_PW_PER_THREAD][4])
{
    u32x4 dv[BR5_PW_PER_THREAD][GL];
    B200_UNROLL
    for (int k = 0; k < BR5_PW_PER_THREAD; k++) {  // all (possibly remote) loads first: their latencies overlap
        const int m = (tid + k * BR5_THREADS) & 255;
        B200_UNROLL
        for (int rr = 0; rr < GL; rr++)
            dv[k][rr] = *reinterpret_cast<const u32x4*>(sm.in_tile(row0 + rr) + 4 * m + 4 * (m >> 4));
    }
    B200_UNROLL
    for (int k = 0; k < BR5_PW_PER_THREAD; k++) {
        const int item = tid + k * BR5_THREADS, m = item & 255, l = item >> 8;
        B200_UNROLL
        for (int rr = 0; rr < GL; rr++) {
            const u32x4 kk = *reinterpret_cast<const u32x4*>(sm.keyb + (size_t)(l * ROWS + row0 + rr) * N1 + 4 * m);
            acc[k][0] += (uint64_t)dv[k][rr].x * kk.x;
            acc[k][1] += (uint64_t)dv[k][rr].y * kk.y;
            acc[k][2] += (uint64_t)dv[k][rr].z * kk.z;
            acc[k][3] += (uint64_t)dv[k][rr].w * kk.w;
        }
    }
}
B200_HD void br5_pw_local(const Br5Smem& sm, int q, int tid, uint64_t (&acc)[BR5_PW_PER_THREAD][4])
{
    B200_UNROLL
    for (int k = 0; k < BR5_PW_PER_THREAD; k++) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0;
    br5_pw_rows(sm, tid, q * GL, acc);
}
B200_HD void br5_pw_finish(const Br5Smem& sm, int q, int tid, uint64_t (&acc)[BR5_PW_PER_THREAD][4])
{
    br5_pw_rows(sm, tid, (q ^ 1) * GL, acc);
    B200_UNROLL
    for (int k = 0; k < BR5_PW_PER_THREAD; k++) {
        const int item = tid + k * BR5_THREADS, m = item & 255, l = item >> 8;
        *reinterpret_cast<u32x4*>(sm.out_tile(l) + 4 * m + 4 * (m >> 4)) =
            u32x4{redc64(acc[k][0]), redc64(acc[k][1]), redc64(acc[k][2]), redc64(acc[k][3])};
    }
}

B200_HD void br5_inv_pA(const Br5Smem& sm, int l, int t) { blk_inv_pA(sm.out_tile(l), sm.tw->p3i, t); }
B200_HD void br5_inv_pB(const Br5Smem& sm, int l, int t)
{
    tw_t w[15];
    blk_load_tw2(sm.tw->p2i, t, w);
    blk_inv_pB(sm.out_tile(l), w, t);
}
B200_HD void br5_inv_pC(const Br5Smem& sm, int l, int t)
{
    uint32_t x[16];
    blk_load_p1(sm.out_tile(l), x, t);
    blk_inv_pC(x);
    B200_UNROLL
    for (int a = 0; a < 16; a++) {
        const uint32_t v = (uint32_t)centered_lift(x[a]) << (LIMB_BITS * l);
        B200_SMEM_ADD(sm.accb + 64 * a + t, v);
    }
}

// SampleExtractIndex(0): CTA 0 holds A (u[0..N-1]), CTA 1 holds B (u[N] = B[0])
B200_HD void br5_epilogue(const Br5Smem& sm, int q, int tid, uint32_t* u_out)
{
    if (q == 0) {
        for (int j = tid; j < N1; j += BR5_THREADS) u_out[j] = (j == 0) ? sm.accb[0] : 0u - sm.accb[N1 - j];
    } else if (tid == 0) {
        u_out[N1] = sm.accb[0];
    }
}

}  // namespace b200
