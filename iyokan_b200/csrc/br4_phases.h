// Blind rotation, latency shape (br4_kernel): ONE rotation job per CTA, 384 threads = 6 teams of 64.
//
// Same arithmetic and same reference functions as br_phases.h (TFHEpp gatebootstrapping.hpp:19-71,
// detwfa.hpp:36-49, trgsw.hpp:62-131, trlwe.hpp:213-223); bit-identical results.  What changes is
// the mapping, chosen for the narrow dependency levels of real netlists (a CAHP processor cycle is
// 40-70 dependent levels of < 148 gates), where the time of ONE rotation is
// what the clock cycle pays:
//   * team m = 3q + d (64 threads, 16 points each, ntt_block.h) owns digit d of accumulator
//     polynomial q on the way in and limb d of polynomial q on the way out, so all 12 transforms of a
//     CMUX step run concurrently on the 12 warps of the SM;
//   * the NTT-domain bootstrapping key of step i (36 polynomials, 147,456 B) is staged into shared
//     memory by ONE bulk-async (TMA) copy issued a phase ahead and awaited on an mbarrier, so the
//     pointwise stage reads key words with 128-bit shared loads instead of L2 round trips;
//   * the accumulator lives in shared memory in natural order; the three limb teams of a polynomial
//     add their contributions with shared-memory atomics (integer adds commute: result is exact).
// Barriers per CMUX step: 2 CTA-wide, 1 per polynomial (192 threads), 4 per team (64 threads).
// Every function below is free of intra-phase cross-thread communication (see br_phases.h), so the
// lock-step CPU simulator (tests/sim/br_sim.cpp) executes this very source.
#pragma once
#include "br_phases.h"
#include "ntt_block.h"

namespace b200 {

constexpr int BR4_TEAMS = 6;
constexpr int BR4_THREADS = BR4_TEAMS * TEAM_THREADS;  // 384
constexpr int BR4_KEY_WORDS = BK_COLS * ROWS * N1;     // 36,864 words = 147,456 B per CMUX step
constexpr int BR4_PW_ITEMS = (N1 / 4) * (BK_COLS / 2); // (quad of positions, column pair) = 768 = 2 per thread

#if defined(__CUDA_ARCH__)
#define B200_SMEM_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define B200_SMEM_ADD(ptr, v) (*(ptr) += (v))
#endif

struct Br4Smem {
    static constexpr size_t BYTES = (size_t)BR4_KEY_WORDS * 4 + 2 * (size_t)ROWS * BT_WORDS * 4 + 2 * (size_t)N1 * 4 +
                                    sizeof(BlockTw) + (size_t)SLOT_STRIDE * 2 + 16;
    uint32_t* keyb;   // [BK_COLS][ROWS][1024] staged key of the current step
    uint32_t* din;    // [ROWS][BT_WORDS]   digit polynomials (NTT in place)
    uint32_t* dout;   // [BK_COLS][BT_WORDS] limb results (inverse NTT in place)
    uint32_t* accb;   // [2][1024] accumulator, natural order
    BlockTw* tw;
    uint16_t* abar;   // [640] mod-switched a_i
    uint64_t* mbar;   // mbarrier of the key stage
    B200_HD void carve(void* base)
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(base);
        keyb = reinterpret_cast<uint32_t*>(p);
        p += (size_t)BR4_KEY_WORDS * 4;
        din = reinterpret_cast<uint32_t*>(p);
        p += (size_t)ROWS * BT_WORDS * 4;
        dout = reinterpret_cast<uint32_t*>(p);
        p += (size_t)BK_COLS * BT_WORDS * 4;
        accb = reinterpret_cast<uint32_t*>(p);
        p += 2 * (size_t)N1 * 4;
        tw = reinterpret_cast<BlockTw*>(p);
        p += sizeof(BlockTw);
        abar = reinterpret_cast<uint16_t*>(p);
        p += (size_t)SLOT_STRIDE * 2;
        mbar = reinterpret_cast<uint64_t*>(p);
    }
    B200_HD uint32_t* in_tile(int r) const { return din + (size_t)r * BT_WORDS; }
    B200_HD uint32_t* out_tile(int c) const { return dout + (size_t)c * BT_WORDS; }
    B200_HD uint32_t* acc(int q) const { return accb + (size_t)q * N1; }
};
static_assert(sizeof(BlockTw) % 16 == 0, "table block keeps 16-byte alignment");
static_assert(Br4Smem::BYTES <= 227 * 1024, "one CTA per SM must fit the opt-in shared memory limit");

// ---- prologue: mod switch (gatebootstrapping.hpp:26-30,58-65) + accumulator init (:31-32) ----
B200_HD void br4_prologue(const Br4Smem& sm, const BrJob& job, const uint16_t* arena, int tid)
{
    for (int i = tid; i < N0; i += BR4_THREADS) {
        const uint32_t c = br_lincomb(job, arena, i);
        sm.abar[i] = (uint16_t)((c + 16u) >> 5);
    }
    const uint32_t bbar = 2u * N1 - (br_lincomb(job, arena, N0) >> 5);
    for (int n = tid; n < 2 * N1; n += BR4_THREADS) {
        uint32_t v = 0;
        if (n >= N1) {  // polynomial B = testvector * X^bbar (utils.hpp:113-128)
            const uint32_t m = ((uint32_t)(n - N1) - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        sm.accb[n] = v;
    }
}

// ---- phase F1: digit d of (X^abar - 1) * acc_q (utils.hpp:130-144, trgsw.hpp:62-78), pass 1 ----
B200_HD void br4_fwd_p1(const Br4Smem& sm, int i, int q, int d, int t)
{
    const uint32_t abar = sm.abar[i];
    const uint32_t* acc = sm.acc(q);
    const uint32_t base = ((uint32_t)t - abar) & (2u * N1 - 1);
    uint32_t dv[16], x[16];
    B200_UNROLL
    for (int a = 0; a < 16; a++) {
        const uint32_t m = (base + 64u * a) & (2u * N1 - 1);
        const uint32_t v = acc[m & (N1 - 1)];
        const uint32_t neg = 0u - ((m >> NBIT) & 1u);
        dv[a] = ((v ^ neg) - neg) - acc[64 * a + t] + (DEC_OFFSET + DEC_ROUND);
    }
    // the team's digit is uniform per warp: one instantiation of the table-driven start per bit field
    if (d == 0) blk_fwd_p1_digits<32 - BGBIT>(sm.tw->r4, dv, x);
    else if (d == 1) blk_fwd_p1_digits<32 - 2 * BGBIT>(sm.tw->r4, dv, x);
    else blk_fwd_p1_digits<32 - 3 * BGBIT>(sm.tw->r4, dv, x);
    blk_store_p1(sm.in_tile(q * GL + d), x, t);
}
B200_HD void br4_fwd_p2(const Br4Smem& sm, int q, int d, int t)
{
    tw_t w[15];
    blk_load_tw2(sm.tw->p2f, t, w);
    blk_fwd_p2(sm.in_tile(q * GL + d), w, t);
}
B200_HD void br4_fwd_p3(const Br4Smem& sm, int q, int d, int t) { blk_fwd_p3(sm.in_tile(q * GL + d), sm.tw->p3f, t); }

// ---- phase M: pointwise multiply-accumulate against the staged key ----
// item = (quad m of NTT positions 4m..4m+3, column pair cp): out[c][j] = REDC(sum_r D[r][j] * BK[c][r][j])
// (measured on B200: interleaving the key loads with the multiply-accumulate of each column beats
// hoisting all 18 loads - every warp is in this phase at once, so bursts serialise LSU and FMA)
B200_HD void br4_pointwise_item(const Br4Smem& sm, int item)
{
    const int m = item & 255, cp = item >> 8;
    const int toff = 4 * m + 4 * (m >> 4);  // bt_pad(4m)
    u32x4 dv[ROWS];
    B200_UNROLL
    for (int r = 0; r < ROWS; r++) dv[r] = *reinterpret_cast<const u32x4*>(sm.in_tile(r) + toff);
    B200_UNROLL
    for (int h = 0; h < 2; h++) {
        const int c = 2 * cp + h;
        uint64_t acc[4] = {0, 0, 0, 0};
        B200_UNROLL
        for (int r = 0; r < ROWS; r++) {
            const u32x4 k = *reinterpret_cast<const u32x4*>(sm.keyb + (size_t)(c * ROWS + r) * N1 + 4 * m);
            acc[0] += (uint64_t)dv[r].x * k.x;
            acc[1] += (uint64_t)dv[r].y * k.y;
            acc[2] += (uint64_t)dv[r].z * k.z;
            acc[3] += (uint64_t)dv[r].w * k.w;
        }
        *reinterpret_cast<u32x4*>(sm.out_tile(c) + toff) = u32x4{redc64(acc[0]), redc64(acc[1]), redc64(acc[2]), redc64(acc[3])};
    }
}

// ---- phase I: inverse NTT of limb l of polynomial q, exact recombination into the accumulator ----
B200_HD void br4_inv_pA(const Br4Smem& sm, int q, int l, int t) { blk_inv_pA(sm.out_tile(q * LIMBS + l), sm.tw->p3i, t); }
B200_HD void br4_inv_pB(const Br4Smem& sm, int q, int l, int t)
{
    tw_t w[15];
    blk_load_tw2(sm.tw->p2i, t, w);
    blk_inv_pB(sm.out_tile(q * LIMBS + l), w, t);
}
B200_HD void br4_inv_pC(const Br4Smem& sm, int q, int l, int t)
{
    uint32_t x[16];
    blk_load_p1(sm.out_tile(q * LIMBS + l), x, t);
    blk_inv_pC(x);
    uint32_t* acc = sm.acc(q);
    B200_UNROLL
    for (int a = 0; a < 16; a++) {
        const uint32_t v = (uint32_t)centered_lift(x[a]) << (LIMB_BITS * l);
        B200_SMEM_ADD(acc + 64 * a + t, v);
    }
}

// ---- epilogue: SampleExtractIndex(0) (trlwe.hpp:213-223) ----
B200_HD void br4_epilogue(const Br4Smem& sm, int tid, uint32_t* u_out)
{
    const uint32_t* a = sm.acc(0);
    for (int j = tid; j < N1; j += BR4_THREADS) u_out[j] = (j == 0) ? a[0] : 0u - a[N1 - j];
    if (tid == 0) u_out[N1] = sm.acc(1)[0];
}

}  // namespace b200
