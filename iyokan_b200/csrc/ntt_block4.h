// Team-level 1024-point negacyclic NTT over Z_p: 256 threads x 4 points, five register passes of two stages.
//
// Same transform as ntt_warp.h / ntt_block.h / ntt_block8.h (same stages, twiddles and lazy-range schedule: every
// intermediate value is congruent, the outputs identical), cut finer still: a thread's dependency chain is 20 butterflies
// and a CTA of three transforms runs 24 warps, six per scheduler.  The phase counters of br6_kernel
// (profiles/r02_phase_timing.md) showed its eight passes issuing 0.44 instructions per cycle with three warps per
// scheduler and barriers costing 5 %: more independent instruction streams, not fewer barriers, is what a latency-bound
// step lacks.  Replaces the same reference code: TwistIFFT/TwistFFT (TFHEpp include/mulfft.hpp:69-134); cuFHE's NTT1024
// uses 128 x 8 (cuFHE include/ntt_gpu/ntt_1024_device.cuh:139-204).
//
// Stage s (0..9) pairs (j, j + (512 >> s)), twiddle psi_rev[2^s + (j >> (10 - s))].  Thread t of a team:
//   pass 1 = stages 0,1 : j = t + 256e                      (forward: table look-up on the digits, fwd_start_r4_group)
//   pass 2 = stages 2,3 : (B, c) = (t >> 6, t & 63)          : j = 256B + 64e + c   (twiddles warp-uniform, constant memory)
//   pass 3 = stages 4,5 : c = t & 15, B from the warp index  : j = 64B + 16e + c    (B = 4(w >> 1) + (w & 1) + 2 * bit 4 of t:
//                         the two half-warps work 128 positions apart, which keeps the padded tile conflict free)
//   pass 4 = stages 6,7 : (U, c) = (t >> 2, t & 3)           : j = 16U + 4e + c
//   pass 5 = stages 8,9 : quad m = t                         : j = 4m + e           (128-bit accesses)
// Tile layout: that of ntt_block8.h (b8_pad), so the pointwise stage and the DSMEM tile copies are shared with br6.
#pragma once
#include "ntt_block8.h"

namespace b200 {

constexpr int TEAM4_THREADS = 256;
constexpr int B4_P3_LEN = 16 * 3;  // pass-3 twiddles [B][3]: psi_rev[16 + B], psi_rev[32 + 2B], psi_rev[32 + 2B + 1]

struct Block4Tw {
    Block8Tw b8;  // q3 (stages 6,7), q4 (stages 8,9) and the digit tables r4 are shared with the 8-point passes
    tw_t p3f[B4_P3_LEN], p3i[B4_P3_LEN];
};
static_assert(sizeof(Block4Tw) % 16 == 0, "table block keeps 16-byte alignment");

inline void block4_tw_init(const NttTables& t, Block4Tw& b)
{
    block8_tw_init(t, b.b8);
    for (int B = 0; B < 16; B++) {
        b.p3f[B * 3] = t.fwd[16 + B];
        b.p3i[B * 3] = t.inv[16 + B];
        for (int h = 0; h < 2; h++) {
            b.p3f[B * 3 + 1 + h] = t.fwd[32 + 2 * B + h];
            b.p3i[B * 3 + 1 + h] = t.inv[32 + 2 * B + h];
        }
    }
}

B200_HD int blk4_p3_block(int t)
{
    const int w = t >> 5;
    return 4 * (w >> 1) + (w & 1) + 2 * ((t >> 4) & 1);
}

// ---- forward: FIX schedule 0 0 | 0 1 | 0 1 | 0 1 | 0 2 (stage by stage the same as ntt_warp.h) ----
template <int SHIFT>
B200_HD void blk4_fwd_p1_digits(uint32_t* tile, const uint32_t* r4, const uint32_t (&dv)[4], int t)  // dv[e] at j = t + 256e
{
    uint32_t x[4];
    fwd_start_r4_group<SHIFT>(r4, dv[0], dv[1], dv[2], dv[3], x[0], x[1], x[2], x[3]);
    B200_UNROLL
    for (int e = 0; e < 4; e++) tile[b8_pad(t + 256 * e)] = x[e];
}
B200_HD void blk4_fwd_p2(uint32_t* tile, int t)
{
    const int B = t >> 6, c = t & 63;
    uint32_t x[4];
    B200_UNROLL
    for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(256 * B + 64 * e + c)];
    ct_stage_n<4, 0, 0>(x, [=](int) { return twf_u(4 + B); });
    ct_stage_n<4, 1, 1>(x, [=](int h) { return twf_u(8 + 2 * B + h); });
    B200_UNROLL
    for (int e = 0; e < 4; e++) tile[b8_pad(256 * B + 64 * e + c)] = x[e];
}
B200_HD void blk4_fwd_p3(uint32_t* tile, const tw_t* p3f, int t)
{
    const int B = blk4_p3_block(t), c = t & 15;
    const tw_t* tw = p3f + B * 3;
    uint32_t x[4];
    B200_UNROLL
    for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(64 * B + 16 * e + c)];
    ct_stage_n<4, 0, 0>(x, [=](int) { return tw[0]; });
    ct_stage_n<4, 1, 1>(x, [=](int h) { return tw[1 + h]; });
    B200_UNROLL
    for (int e = 0; e < 4; e++) tile[b8_pad(64 * B + 16 * e + c)] = x[e];
}
B200_HD void blk4_fwd_p4(uint32_t* tile, const tw_t* q3f, int t)
{
    const int U = t >> 2, c = t & 3;
    const tw_t* tw = q3f + U * 3;
    uint32_t x[4];
    B200_UNROLL
    for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(16 * U + 4 * e + c)];
    ct_stage_n<4, 0, 0>(x, [=](int) { return tw[0]; });
    ct_stage_n<4, 1, 1>(x, [=](int h) { return tw[1 + h]; });
    B200_UNROLL
    for (int e = 0; e < 4; e++) tile[b8_pad(16 * U + 4 * e + c)] = x[e];
}
B200_HD void blk4_fwd_p5(uint32_t* tile, const tw_t* q4f, int t)  // output < 4p
{
    u32x4* ptr = reinterpret_cast<u32x4*>(tile + b8_quad(t));
    const tw_t* tw = q4f + 3 * t;
    const u32x4 v = *ptr;
    uint32_t x[4] = {v.x, v.y, v.z, v.w};
    ct_stage_n<4, 0, 0>(x, [=](int) { return tw[0]; });
    ct_stage_n<4, 1, 2>(x, [=](int h) { return tw[1 + h]; });
    *ptr = u32x4{x[0], x[1], x[2], x[3]};
}

// ---- inverse: every stage folds the sum below 4p ----
B200_HD void blk4_inv_pA(uint32_t* tile, const tw_t* q4i, int t)
{
    u32x4* ptr = reinterpret_cast<u32x4*>(tile + b8_quad(t));
    const tw_t* tw = q4i + 3 * t;
    const u32x4 v = *ptr;
    uint32_t x[4] = {v.x, v.y, v.z, v.w};
    gs_stage_n<4, 1, 1>(x, [=](int h) { return tw[1 + h]; });
    gs_stage_n<4, 0, 1>(x, [=](int) { return tw[0]; });
    *ptr = u32x4{x[0], x[1], x[2], x[3]};
}
B200_HD void blk4_inv_pB(uint32_t* tile, const tw_t* q3i, int t)
{
    const int U = t >> 2, c = t & 3;
    const tw_t* tw = q3i + U * 3;
    uint32_t x[4];
    B200_UNROLL
    for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(16 * U + 4 * e + c)];
    gs_stage_n<4, 1, 1>(x, [=](int h) { return tw[1 + h]; });
    gs_stage_n<4, 0, 1>(x, [=](int) { return tw[0]; });
    B200_UNROLL
    for (int e = 0; e < 4; e++) tile[b8_pad(16 * U + 4 * e + c)] = x[e];
}
B200_HD void blk4_inv_pC(uint32_t* tile, const tw_t* p3i, int t)
{
    const int B = blk4_p3_block(t), c = t & 15;
    const tw_t* tw = p3i + B * 3;
    uint32_t x[4];
    B200_UNROLL
    for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(64 * B + 16 * e + c)];
    gs_stage_n<4, 1, 1>(x, [=](int h) { return tw[1 + h]; });
    gs_stage_n<4, 0, 1>(x, [=](int) { return tw[0]; });
    B200_UNROLL
    for (int e = 0; e < 4; e++) tile[b8_pad(64 * B + 16 * e + c)] = x[e];
}
B200_HD void blk4_inv_pD(uint32_t* tile, int t)
{
    const int B = t >> 6, c = t & 63;
    uint32_t x[4];
    B200_UNROLL
    for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(256 * B + 64 * e + c)];
    gs_stage_n<4, 1, 1>(x, [=](int h) { return twi_u(8 + 2 * B + h); });
    gs_stage_n<4, 0, 1>(x, [=](int) { return twi_u(4 + B); });
    B200_UNROLL
    for (int e = 0; e < 4; e++) tile[b8_pad(256 * B + 64 * e + c)] = x[e];
}
B200_HD void blk4_inv_pE(const uint32_t* tile, uint32_t (&x)[4], int t)  // x[e] at j = t + 256e; output < 4p
{
    B200_UNROLL
    for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(t + 256 * e)];
    gs_stage_n<4, 1, 1>(x, [](int h) { return twi_u(2 + h); });
    gs_stage_n<4, 0, 1>(x, [](int) { return twi_u(1); });
}

}  // namespace b200
