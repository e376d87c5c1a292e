// Blind rotation, "one warp per transform" variant (br2_kernel<G>): 6 warps per rotation job.
//
// Same arithmetic, same reference functions as br_phases.h (TFHEpp gatebootstrapping.hpp:19-71,
// detwfa.hpp:36-49, trgsw.hpp:62-131, trlwe.hpp:213-223); only the work decomposition differs.
// In br_kernel<G> a warp owns one accumulator polynomial and runs its three digit transforms and
// its three limb transforms one after the other, which caps both the latency of a single job
// (~10 ms per dependency level of a circuit) and the number of warps an SM can keep busy.
// Here warp w = 6g + 3q + d of the CTA owns ONE digit polynomial (forward) / ONE limb polynomial
// (inverse) of accumulator polynomial q of job g:
//   phase F  : rotated difference of acc_q (read from a natural-order shared copy), digit d,
//              forward NTT into tile 3q+d
//   phase M  : pointwise stage, CTA-wide; a lane pair (l, l+16) shares one NTT position and
//              produces output columns {0,1,2} / {3,4,5}; key words are loaded once per CTA and
//              reused for all G jobs
//   phase I  : inverse NTT of limb d, centred lift, << 11d, back into the tile
//   combine  : the three warps of (g, q) meet at a 96-thread named barrier, each adds a third of
//              the coefficients of the three limb tiles into the shared accumulator copy
// Every function is free of intra-phase cross-thread communication (see br_phases.h).
#pragma once
#include "br_phases.h"

namespace b200 {

constexpr int BR2_WARPS_PER_JOB = 6;
constexpr int PW2_ITEMS = 2 * N1;   // (position, column half) pairs
constexpr int PW2_COLS = BK_COLS / 2;

template <int G>
struct Br2Smem {
    static constexpr int DBUF_WORDS = G * ROWS * TILE_WORDS;
    static constexpr int ACC_WORDS = G * 2 * N1;
    static constexpr int ABAR_HALFS = G * SLOT_STRIDE;
    static constexpr size_t BYTES = (size_t)DBUF_WORDS * 4 + (size_t)ACC_WORDS * 4 +
                                    2 * (size_t)TW2_LEN * sizeof(tw_t) + (size_t)ABAR_HALFS * 2;
    uint32_t* dbuf;
    uint32_t* accbuf;
    tw_t* tw2f;
    tw_t* tw2i;
    uint16_t* abar;
    B200_HD void carve(void* base)
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(base);
        dbuf = reinterpret_cast<uint32_t*>(p);
        p += (size_t)DBUF_WORDS * 4;
        accbuf = reinterpret_cast<uint32_t*>(p);
        p += (size_t)ACC_WORDS * 4;
        tw2f = reinterpret_cast<tw_t*>(p);
        p += (size_t)TW2_LEN * sizeof(tw_t);
        tw2i = reinterpret_cast<tw_t*>(p);
        p += (size_t)TW2_LEN * sizeof(tw_t);
        abar = reinterpret_cast<uint16_t*>(p);
    }
    B200_HD uint32_t* tile(int g, int row) const { return dbuf + (size_t)(g * ROWS + row) * TILE_WORDS; }
    B200_HD uint32_t* acc(int g, int q) const { return accbuf + (size_t)(g * 2 + q) * N1; }
};

// ---- prologue: mod switch (all 6 warps of the job share the 636 coefficients) + accumulator init
template <int G>
B200_HD void br2_prologue(const Br2Smem<G>& sm, const BrJob& job, const uint16_t* arena, int g, int q, int d, int lane)
{
    for (int i = (q * GL + d) * 32 + lane; i < N0; i += 32 * BR2_WARPS_PER_JOB) {
        const uint32_t c = br_lincomb(job, arena, i);
        sm.abar[g * SLOT_STRIDE + i] = (uint16_t)((c + 16u) >> 5);
    }
    const uint32_t bbar = 2u * N1 - (br_lincomb(job, arena, N0) >> 5);
    uint32_t* acc = sm.acc(g, q);
    for (int a = d; a < 32; a += GL) {  // the three warps of (g, q) interleave rows
        const int n = 32 * a + lane;
        uint32_t v = 0;
        if (q == 1) {
            const uint32_t m = ((uint32_t)n - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        acc[n] = v;
    }
}

// ---- phase F(a): digit d of (X^abar - 1) * acc_q -> forward pass 1 -> column store
template <int G>
B200_HD void br2_fwd_a(const Br2Smem<G>& sm, int i, int g, int q, int d, int lane)
{
    const uint32_t abar = sm.abar[g * SLOT_STRIDE + i];
    const uint32_t* acc = sm.acc(g, q);
    const uint32_t base = ((uint32_t)lane - abar) & (2u * N1 - 1);
    const int sh = 32 - (d + 1) * BGBIT;
    uint32_t x[32];
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const uint32_t m = (base + 32u * a) & (2u * N1 - 1);
        const uint32_t v = acc[m & (N1 - 1)];
        const uint32_t neg = 0u - ((m >> NBIT) & 1u);
        const uint32_t diff = ((v ^ neg) - neg) - acc[32 * a + lane] + (DEC_OFFSET + DEC_ROUND);
        x[a] = ((diff >> sh) & ((1u << BGBIT) - 1)) + (P - (1u << (BGBIT - 1)));
    }
    fwd_pass1(x);
    tile_store_col(sm.tile(g, q * GL + d), x, lane);
}
template <int G>
B200_HD void br2_fwd_b(const Br2Smem<G>& sm, int g, int q, int d, int lane)
{
    uint32_t x[32];
    uint32_t* t = sm.tile(g, q * GL + d);
    tile_load_row(t, x, lane);
    fwd_pass2(x, sm.tw2f, lane);
    tile_store_row(t, x, lane);
}

// ---- phase M: item = (position j, column half h); lanes l and l+16 of a warp share j
B200_HD void pw2_item(int item, int& j, int& half)
{
    j = ((item >> 5) << 4) | (item & 15);
    half = (item >> 4) & 1;
}
B200_HD void pw2_load(const uint32_t* bk_i, int j, int half, uint32_t (&bkv)[PW2_COLS][ROWS])
{
    B200_UNROLL
    for (int c = 0; c < PW2_COLS; c++) {
        B200_UNROLL
        for (int r = 0; r < ROWS; r++) bkv[c][r] = bk_i[(size_t)((half * PW2_COLS + c) * ROWS + r) * N1 + j];
    }
}
// read + multiply-accumulate for job g (no stores: the lane pair still reads the same words)
template <int G>
B200_HD void pw2_compute(const Br2Smem<G>& sm, int g, int off, const uint32_t (&bkv)[PW2_COLS][ROWS],
                         uint32_t (&o)[PW2_COLS])
{
    uint32_t d[ROWS];
    B200_UNROLL
    for (int r = 0; r < ROWS; r++) d[r] = sm.tile(g, r)[off];
    B200_UNROLL
    for (int c = 0; c < PW2_COLS; c++) {
        uint64_t acc = 0;
        B200_UNROLL
        for (int r = 0; r < ROWS; r++) acc += (uint64_t)d[r] * bkv[c][r];
        o[c] = redc64(acc);
    }
}
template <int G>
B200_HD void pw2_store(const Br2Smem<G>& sm, int g, int half, int off, const uint32_t (&o)[PW2_COLS])
{
    B200_UNROLL
    for (int c = 0; c < PW2_COLS; c++) sm.tile(g, half * PW2_COLS + c)[off] = o[c];
}

// ---- phase I: limb d of polynomial q
template <int G>
B200_HD void br2_inv_a(const Br2Smem<G>& sm, int g, int q, int d, int lane)
{
    uint32_t x[32];
    uint32_t* t = sm.tile(g, q * LIMBS + d);
    tile_load_row(t, x, lane);
    inv_pass1(x, sm.tw2i, lane);
    tile_store_row(t, x, lane);
}
template <int G>
B200_HD void br2_inv_b(const Br2Smem<G>& sm, int g, int q, int d, int lane)
{
    uint32_t x[32];
    uint32_t* t = sm.tile(g, q * LIMBS + d);
    tile_load_col(t, x, lane);
    inv_pass2(x);
    B200_UNROLL
    for (int a = 0; a < 32; a++) x[a] = (uint32_t)centered_lift(x[a]) << (LIMB_BITS * d);
    tile_store_col(t, x, lane);  // same lane wrote/reads each word: no barrier needed in between
}
// ---- combine (after the 96-thread barrier): warp d adds rows a = d, d+3, ... of the three limb tiles
template <int G>
B200_HD void br2_combine(const Br2Smem<G>& sm, int g, int q, int d, int lane)
{
    uint32_t* acc = sm.acc(g, q);
    const uint32_t* t0 = sm.tile(g, q * LIMBS);
    for (int a = d; a < 32; a += LIMBS) {
        const int k = tile_idx(a, lane);
        acc[32 * a + lane] += t0[k] + t0[TILE_WORDS + k] + t0[2 * TILE_WORDS + k];
    }
}

template <int G>
B200_HD void br2_epilogue(const Br2Smem<G>& sm, int g, int q, int d, int lane, uint32_t* u_out)
{
    if (d != 0) return;
    const uint32_t* acc = sm.acc(g, q);
    if (q == 0) {
        for (int j = lane; j < N1; j += 32) u_out[j] = (j == 0) ? acc[0] : 0u - acc[N1 - j];
    } else if (lane == 0) {
        u_out[N1] = acc[0];
    }
}

}  // namespace b200
