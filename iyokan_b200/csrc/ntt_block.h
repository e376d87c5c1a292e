// Team-level 1024-point negacyclic NTT over Z_p: 64 threads x 16 points, three register passes.
//
// Same transform as ntt_warp.h (same stage definition, same twiddles, same lazy-range schedule, so
// every intermediate value is bit-identical to the warp transform); only the work decomposition
// differs.  The warp transform gives one lane 32 points and needs one shared-memory transpose; it
// is the throughput shape.  Here a 64-thread "team" (two warps) shares one polynomial so that a
// single rotation job can occupy a whole SM (6 teams = 12 warps): this is the latency shape used
// for the narrow dependency levels of a netlist (br4_kernel).  It replaces the same reference
// code as ntt_warp.h: TwistIFFT/TwistFFT (TFHEpp include/mulfft.hpp:69-134) and cuFHE's
// NTT1024 (cuFHE include/ntt_gpu/ntt_1024_device.cuh:139-204, 128 threads x 8 points).
//
// Stage s (0..9) pairs (j, j + (512 >> s)) and uses twiddle psi_rev[2^s + (j >> (10 - s))].
//   pass 1 = stages 0..3 : thread t holds j = 64a + t,       a = 0..15  (twiddles thread-independent)
//   pass 2 = stages 4..7 : thread (A, c) = (t >> 2, t & 3) holds j = 64A + 4b + c, b = 0..15
//   pass 3 = stages 8..9 : thread t holds four quads j = 4m + e, m = t + 64k, e = 0..3
// The inverse runs the passes in the opposite order with Gentleman-Sande butterflies.
// Between passes the polynomial lives in a padded shared-memory tile: word bt_pad(j) = j + 4*(j >> 6)
// keeps the pass-1 (stride 1), pass-2 (stride 4 within, 64 across) and pass-3 (128-bit) accesses
// bank-conflict free.  Teams synchronise with a 64-thread named barrier, never __syncthreads.
#pragma once
#include "hd.h"
#include "modarith.h"
#include "ntt_warp.h"

namespace b200 {

constexpr int TEAM_THREADS = 64;
constexpr int BT_WORDS = 1024 + 4 * 16;  // 1088 words = 4352 B per polynomial tile
constexpr int BT_P2_LEN = 16 * 15;       // pass-2 twiddles: [A][15]
constexpr int BT_P3_LEN = 256 * 3;       // pass-3 twiddles: [m][3]

B200_HD int bt_pad(int j) { return j + ((j >> 6) << 2); }

struct BlockTw {
    tw_t p2f[BT_P2_LEN], p2i[BT_P2_LEN];
    tw_t p3f[BT_P3_LEN], p3i[BT_P3_LEN];
    uint32_t r4[R4_WORDS];  // digit x twiddle tables of forward stages 0 and 1 (ntt_warp.h fwd_start_r4_group)
};

// p2[A*15 + (2^ls - 1) + g] = psi_rev[(16 << ls) + (A << ls) + g]      (stage 4 + ls, g < 2^ls)
// p3[m*3] = psi_rev[256 + m],  p3[m*3 + 1 + h] = psi_rev[512 + 2m + h]  (stages 8 and 9)
inline void block_tw_init(const NttTables& t, BlockTw& b)
{
    for (int k = 0; k < R4_WORDS; k++) b.r4[k] = t.r4[k];
    for (int A = 0; A < 16; A++)
        for (int ls = 0; ls < 4; ls++)
            for (int g = 0; g < (1 << ls); g++) {
                const int idx = (16 << ls) + (A << ls) + g, pos = A * 15 + (1 << ls) - 1 + g;
                b.p2f[pos] = t.fwd[idx];
                b.p2i[pos] = t.inv[idx];
            }
    for (int m = 0; m < 256; m++) {
        b.p3f[m * 3] = t.fwd[256 + m];
        b.p3i[m * 3] = t.inv[256 + m];
        for (int h = 0; h < 2; h++) {
            b.p3f[m * 3 + 1 + h] = t.fwd[512 + 2 * m + h];
            b.p3i[m * 3 + 1 + h] = t.inv[512 + 2 * m + h];
        }
    }
}

// radix-2 stages over NPT registers (NPT = 16 or 4); LS = local stage, 2^LS twiddle groups
template <int NPT, int LS, int FIX, class TwFn>
B200_HD void ct_stage_n(uint32_t (&x)[NPT], TwFn tw)
{
    constexpr int half = (NPT / 2) >> LS;
    B200_UNROLL
    for (int g = 0; g < (1 << LS); g++) {
        const tw_t w = tw(g);
        B200_UNROLL
        for (int k = 0; k < half; k++) {
            const int i0 = g * 2 * half + k, i1 = i0 + half;
            const uint32_t X = apply_fix<FIX>(x[i0]);
            const uint32_t T = shoup_mul(x[i1], w);
            x[i0] = X + T;
            x[i1] = X - T + P2;
        }
    }
}
template <int NPT, int LS, int FIX, class TwFn>
B200_HD void gs_stage_n(uint32_t (&x)[NPT], TwFn tw)
{
    constexpr int half = (NPT / 2) >> LS;
    B200_UNROLL
    for (int g = 0; g < (1 << LS); g++) {
        const tw_t w = tw(g);
        B200_UNROLL
        for (int k = 0; k < half; k++) {
            const int i0 = g * 2 * half + k, i1 = i0 + half;
            const uint32_t U = x[i0], V = x[i1];
            x[i0] = apply_fix<FIX>(U + V);
            x[i1] = shoup_mul(U - V + P4, w);
        }
    }
}

// ---- forward (same FIX schedule as fwd_pass1/fwd_pass2 in ntt_warp.h: 0 0 0 1 | 0 1 0 1 | 0 2) ----
B200_HD void blk_fwd_p1(uint32_t (&x)[16])  // x[a] = value at j = 64a + t
{
    ct_stage_n<16, 0, 0>(x, [](int g) { return twf_u(1 + g); });
    ct_stage_n<16, 1, 0>(x, [](int g) { return twf_u(2 + g); });
    ct_stage_n<16, 2, 0>(x, [](int g) { return twf_u(4 + g); });
    ct_stage_n<16, 3, 1>(x, [](int g) { return twf_u(8 + g); });
}
// pass 1 on gadget digits still in their bit field of dv[a] (coefficient 64a + t): stages 0 and 1 by table look-up
// (groups {a0, a0+4, a0+8, a0+12}), stages 2 and 3 as usual
template <int SHIFT>
B200_HD void blk_fwd_p1_digits(const uint32_t* r4, const uint32_t (&dv)[16], uint32_t (&x)[16])
{
    B200_UNROLL
    for (int a0 = 0; a0 < 4; a0++)
        fwd_start_r4_group<SHIFT>(r4, dv[a0], dv[a0 + 4], dv[a0 + 8], dv[a0 + 12], x[a0], x[a0 + 4], x[a0 + 8], x[a0 + 12]);
    ct_stage_n<16, 2, 0>(x, [](int g) { return twf_u(4 + g); });
    ct_stage_n<16, 3, 1>(x, [](int g) { return twf_u(8 + g); });
}
B200_HD void blk_store_p1(uint32_t* tile, const uint32_t (&x)[16], int t)
{
    B200_UNROLL
    for (int a = 0; a < 16; a++) tile[bt_pad(64 * a + t)] = x[a];
}
B200_HD void blk_load_p1(const uint32_t* tile, uint32_t (&x)[16], int t)
{
    B200_UNROLL
    for (int a = 0; a < 16; a++) x[a] = tile[bt_pad(64 * a + t)];
}
// pass 2 in place on the tile: load, stages 4..7, store.  The 15 twiddles of a thread depend only on
// A = t >> 2 and are fetched first (measured: keeping them in registers across steps buys nothing).
B200_HD void blk_load_tw2(const tw_t* p2, int t, tw_t (&tw)[15])
{
    B200_UNROLL
    for (int k = 0; k < 15; k++) tw[k] = p2[(t >> 2) * 15 + k];
}
B200_HD void blk_fwd_p2(uint32_t* tile, const tw_t (&tw)[15], int t)
{
    const int A = t >> 2, c = t & 3;
    uint32_t* base = tile + 68 * A + c;  // bt_pad(64A + 4b + c) = 68A + 4b + c
    uint32_t x[16];
    B200_UNROLL
    for (int b = 0; b < 16; b++) x[b] = base[4 * b];
    ct_stage_n<16, 0, 0>(x, [&](int g) { return tw[0 + g]; });
    ct_stage_n<16, 1, 1>(x, [&](int g) { return tw[1 + g]; });
    ct_stage_n<16, 2, 0>(x, [&](int g) { return tw[3 + g]; });
    ct_stage_n<16, 3, 1>(x, [&](int g) { return tw[7 + g]; });
    B200_UNROLL
    for (int b = 0; b < 16; b++) base[4 * b] = x[b];
}
// pass 3 in place: four quads per thread, stages 8..9, output < 4p
B200_HD void blk_fwd_p3(uint32_t* tile, const tw_t* p3f, int t)
{
    B200_UNROLL
    for (int k = 0; k < 4; k++) {
        const int m = t + 64 * k;
        u32x4* ptr = reinterpret_cast<u32x4*>(tile + 4 * m + 4 * (m >> 4));
        const tw_t* tw = p3f + 3 * m;
        const u32x4 v = *ptr;
        uint32_t x[4] = {v.x, v.y, v.z, v.w};
        ct_stage_n<4, 0, 0>(x, [=](int) { return tw[0]; });
        ct_stage_n<4, 1, 2>(x, [=](int g) { return tw[1 + g]; });
        *ptr = u32x4{x[0], x[1], x[2], x[3]};
    }
}

// ---- inverse (every stage folds the sum below 4p, as inv_pass1/inv_pass2) ----
B200_HD void blk_inv_pA(uint32_t* tile, const tw_t* p3i, int t)
{
    B200_UNROLL
    for (int k = 0; k < 4; k++) {
        const int m = t + 64 * k;
        u32x4* ptr = reinterpret_cast<u32x4*>(tile + 4 * m + 4 * (m >> 4));
        const tw_t* tw = p3i + 3 * m;
        const u32x4 v = *ptr;
        uint32_t x[4] = {v.x, v.y, v.z, v.w};
        gs_stage_n<4, 1, 1>(x, [=](int g) { return tw[1 + g]; });
        gs_stage_n<4, 0, 1>(x, [=](int) { return tw[0]; });
        *ptr = u32x4{x[0], x[1], x[2], x[3]};
    }
}
B200_HD void blk_inv_pB(uint32_t* tile, const tw_t (&tw)[15], int t)
{
    const int A = t >> 2, c = t & 3;
    uint32_t* base = tile + 68 * A + c;
    uint32_t x[16];
    B200_UNROLL
    for (int b = 0; b < 16; b++) x[b] = base[4 * b];
    gs_stage_n<16, 3, 1>(x, [&](int g) { return tw[7 + g]; });
    gs_stage_n<16, 2, 1>(x, [&](int g) { return tw[3 + g]; });
    gs_stage_n<16, 1, 1>(x, [&](int g) { return tw[1 + g]; });
    gs_stage_n<16, 0, 1>(x, [&](int g) { return tw[0 + g]; });
    B200_UNROLL
    for (int b = 0; b < 16; b++) base[4 * b] = x[b];
}
B200_HD void blk_inv_pC(uint32_t (&x)[16])  // x[a] at j = 64a + t; output < 4p
{
    gs_stage_n<16, 3, 1>(x, [](int g) { return twi_u(8 + g); });
    gs_stage_n<16, 2, 1>(x, [](int g) { return twi_u(4 + g); });
    gs_stage_n<16, 1, 1>(x, [](int g) { return twi_u(2 + g); });
    gs_stage_n<16, 0, 1>(x, [](int g) { return twi_u(1 + g); });
}

}  // namespace b200
