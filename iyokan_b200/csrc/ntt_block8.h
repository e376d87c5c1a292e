// Team-level 1024-point negacyclic NTT over Z_p: 128 threads x 8 points, four register passes.
//
// Same transform as ntt_warp.h / ntt_block.h (same stages, twiddles and lazy-range schedule: every
// intermediate value is bit-identical), cut finer: the per-thread dependency chain of a transform is
// 40 butterflies instead of 80 (ntt_block.h) or 160 (ntt_warp.h).  This is the shape for the
// cluster kernel (br6_kernel), where two SMs serve ONE rotation job and the time of a CMUX step is
// the length of that chain, not the multiply throughput.  (Measured: on ONE SM the extra pass costs more
// shared-memory traffic than the shorter chain saves - 3.27 ms per rotation with 6 x 128 threads against
// 3.02 ms for br4_kernel's 6 x 64 - so the single-CTA shape keeps ntt_block.h.)  Replaces the same reference code:
// TwistIFFT/TwistFFT (TFHEpp include/mulfft.hpp:69-134); cuFHE's NTT1024 uses the same 128 x 8
// split with a 64-bit modulus (cuFHE include/ntt_gpu/ntt_1024_device.cuh:139-204).
//
// Stage s (0..9) pairs (j, j + (512 >> s)), twiddle psi_rev[2^s + (j >> (10 - s))].
//   pass 1 = stages 0..2 : thread t holds j = 128a + t,            a = 0..7   (twiddles uniform)
//   pass 2 = stages 3..5 : thread (U, c) = (t >> 4, t & 15) holds j = 128U + 16e + c, e = 0..7
//   pass 3 = stages 6..7 : two groups g = t, t + 128; (U, c) = (g >> 2, g & 3): j = 16U + 4e + c, e = 0..3
//   pass 4 = stages 8..9 : two quads m = t, t + 128: j = 4m + e, e = 0..3   (128-bit accesses)
// Tile layout: word b8_pad(j) = j + 4*(j >> 5) (1152 words); found by exhaustive search to make all four
// access patterns and the 128-bit pointwise accesses bank-conflict free.
#pragma once
#include "ntt_block.h"

namespace b200 {

constexpr int TEAM8_THREADS = 128;
constexpr int B8_WORDS = 1024 + 4 * 32;  // 1152 words = 4608 B per polynomial tile
constexpr int B8_Q2_LEN = 8 * 7;         // pass-2 twiddles [U][7]
constexpr int B8_Q3_LEN = 64 * 3;        // pass-3 twiddles [U][3]

B200_HD int b8_pad(int j) { return j + ((j >> 5) << 2); }
B200_HD int b8_quad(int m) { return 4 * m + ((m >> 3) << 2); }  // b8_pad(4m)

struct Block8Tw {
    tw_t q2f[B8_Q2_LEN], q2i[B8_Q2_LEN];
    tw_t q3f[B8_Q3_LEN], q3i[B8_Q3_LEN];
    tw_t q4f[BT_P3_LEN], q4i[BT_P3_LEN];  // [m][3]: psi_rev[256 + m], psi_rev[512 + 2m], psi_rev[512 + 2m + 1]
    uint32_t r4[R4_WORDS];                // digit x twiddle tables of forward stages 0 and 1 (ntt_warp.h fwd_start_r4_group)
};
static_assert(R4_WORDS % 4 == 0, "keeps the table block a multiple of 16 bytes");

inline void block8_tw_init(const NttTables& t, Block8Tw& b)
{
    for (int k = 0; k < R4_WORDS; k++) b.r4[k] = t.r4[k];
    for (int U = 0; U < 8; U++)
        for (int ls = 0; ls < 3; ls++)
            for (int g = 0; g < (1 << ls); g++) {  // stage 3 + ls: psi_rev[(8 << ls) + (U << ls) + g]
                const int idx = (8 << ls) + (U << ls) + g, pos = U * 7 + (1 << ls) - 1 + g;
                b.q2f[pos] = t.fwd[idx];
                b.q2i[pos] = t.inv[idx];
            }
    for (int U = 0; U < 64; U++)
        for (int ls = 0; ls < 2; ls++)
            for (int g = 0; g < (1 << ls); g++) {  // stage 6 + ls: psi_rev[(64 << ls) + (U << ls) + g]
                const int idx = (64 << ls) + (U << ls) + g, pos = U * 3 + (1 << ls) - 1 + g;
                b.q3f[pos] = t.fwd[idx];
                b.q3i[pos] = t.inv[idx];
            }
    for (int m = 0; m < 256; m++) {
        b.q4f[m * 3] = t.fwd[256 + m];
        b.q4i[m * 3] = t.inv[256 + m];
        for (int h = 0; h < 2; h++) {
            b.q4f[m * 3 + 1 + h] = t.fwd[512 + 2 * m + h];
            b.q4i[m * 3 + 1 + h] = t.inv[512 + 2 * m + h];
        }
    }
}

// ---- forward: FIX schedule 0 0 0 | 1 0 1 | 0 1 | 0 2 (stage by stage the same as ntt_warp.h) ----
B200_HD void blk8_fwd_p1(uint32_t (&x)[8])  // x[a] at j = 128a + t
{
    ct_stage_n<8, 0, 0>(x, [](int g) { return twf_u(1 + g); });
    ct_stage_n<8, 1, 0>(x, [](int g) { return twf_u(2 + g); });
    ct_stage_n<8, 2, 0>(x, [](int g) { return twf_u(4 + g); });
}
// pass 1 when the inputs are gadget digits still sitting in their bit field of dv[a] (coefficient 128a + t): stages 0 and 1
// by table look-up (pairs (a, a+4) then (a, a+2): the groups are {a0, a0+2, a0+4, a0+6}), stage 2 as usual
template <int SHIFT>
B200_HD void blk8_fwd_p1_digits(const uint32_t* r4, const uint32_t (&dv)[8], uint32_t (&x)[8])
{
    fwd_start_r4_group<SHIFT>(r4, dv[0], dv[2], dv[4], dv[6], x[0], x[2], x[4], x[6]);
    fwd_start_r4_group<SHIFT>(r4, dv[1], dv[3], dv[5], dv[7], x[1], x[3], x[5], x[7]);
    ct_stage_n<8, 2, 0>(x, [](int g) { return twf_u(4 + g); });
}
B200_HD void blk8_store_p1(uint32_t* tile, const uint32_t (&x)[8], int t)
{
    B200_UNROLL
    for (int a = 0; a < 8; a++) tile[b8_pad(128 * a + t)] = x[a];
}
B200_HD void blk8_load_p1(const uint32_t* tile, uint32_t (&x)[8], int t)
{
    B200_UNROLL
    for (int a = 0; a < 8; a++) x[a] = tile[b8_pad(128 * a + t)];
}
B200_HD void blk8_fwd_p2(uint32_t* tile, const tw_t* q2f, int t)
{
    const int U = t >> 4, c = t & 15;
    const tw_t* tw = q2f + U * 7;
    uint32_t x[8];
    B200_UNROLL
    for (int e = 0; e < 8; e++) x[e] = tile[b8_pad(128 * U + 16 * e + c)];
    ct_stage_n<8, 0, 1>(x, [=](int g) { return tw[0 + g]; });
    ct_stage_n<8, 1, 0>(x, [=](int g) { return tw[1 + g]; });
    ct_stage_n<8, 2, 1>(x, [=](int g) { return tw[3 + g]; });
    B200_UNROLL
    for (int e = 0; e < 8; e++) tile[b8_pad(128 * U + 16 * e + c)] = x[e];
}
B200_HD void blk8_fwd_p3(uint32_t* tile, const tw_t* q3f, int t)
{
    B200_UNROLL
    for (int k = 0; k < 2; k++) {
        const int g = t + 128 * k, U = g >> 2, c = g & 3;
        const tw_t* tw = q3f + U * 3;
        uint32_t x[4];
        B200_UNROLL
        for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(16 * U + 4 * e + c)];
        ct_stage_n<4, 0, 0>(x, [=](int) { return tw[0]; });
        ct_stage_n<4, 1, 1>(x, [=](int h) { return tw[1 + h]; });
        B200_UNROLL
        for (int e = 0; e < 4; e++) tile[b8_pad(16 * U + 4 * e + c)] = x[e];
    }
}
B200_HD void blk8_fwd_p4(uint32_t* tile, const tw_t* q4f, int t)  // output < 4p
{
    B200_UNROLL
    for (int k = 0; k < 2; k++) {
        const int m = t + 128 * k;
        u32x4* ptr = reinterpret_cast<u32x4*>(tile + b8_quad(m));
        const tw_t* tw = q4f + 3 * m;
        const u32x4 v = *ptr;
        uint32_t x[4] = {v.x, v.y, v.z, v.w};
        ct_stage_n<4, 0, 0>(x, [=](int) { return tw[0]; });
        ct_stage_n<4, 1, 2>(x, [=](int h) { return tw[1 + h]; });
        *ptr = u32x4{x[0], x[1], x[2], x[3]};
    }
}

// ---- inverse: every stage folds the sum below 4p ----
B200_HD void blk8_inv_pA(uint32_t* tile, const tw_t* q4i, int t)
{
    B200_UNROLL
    for (int k = 0; k < 2; k++) {
        const int m = t + 128 * k;
        u32x4* ptr = reinterpret_cast<u32x4*>(tile + b8_quad(m));
        const tw_t* tw = q4i + 3 * m;
        const u32x4 v = *ptr;
        uint32_t x[4] = {v.x, v.y, v.z, v.w};
        gs_stage_n<4, 1, 1>(x, [=](int h) { return tw[1 + h]; });
        gs_stage_n<4, 0, 1>(x, [=](int) { return tw[0]; });
        *ptr = u32x4{x[0], x[1], x[2], x[3]};
    }
}
B200_HD void blk8_inv_pB(uint32_t* tile, const tw_t* q3i, int t)
{
    B200_UNROLL
    for (int k = 0; k < 2; k++) {
        const int g = t + 128 * k, U = g >> 2, c = g & 3;
        const tw_t* tw = q3i + U * 3;
        uint32_t x[4];
        B200_UNROLL
        for (int e = 0; e < 4; e++) x[e] = tile[b8_pad(16 * U + 4 * e + c)];
        gs_stage_n<4, 1, 1>(x, [=](int h) { return tw[1 + h]; });
        gs_stage_n<4, 0, 1>(x, [=](int) { return tw[0]; });
        B200_UNROLL
        for (int e = 0; e < 4; e++) tile[b8_pad(16 * U + 4 * e + c)] = x[e];
    }
}
B200_HD void blk8_inv_pC(uint32_t* tile, const tw_t* q2i, int t)
{
    const int U = t >> 4, c = t & 15;
    const tw_t* tw = q2i + U * 7;
    uint32_t x[8];
    B200_UNROLL
    for (int e = 0; e < 8; e++) x[e] = tile[b8_pad(128 * U + 16 * e + c)];
    gs_stage_n<8, 2, 1>(x, [=](int g) { return tw[3 + g]; });
    gs_stage_n<8, 1, 1>(x, [=](int g) { return tw[1 + g]; });
    gs_stage_n<8, 0, 1>(x, [=](int g) { return tw[0 + g]; });
    B200_UNROLL
    for (int e = 0; e < 8; e++) tile[b8_pad(128 * U + 16 * e + c)] = x[e];
}
B200_HD void blk8_inv_pD(uint32_t (&x)[8])  // x[a] at j = 128a + t; output < 4p
{
    gs_stage_n<8, 2, 1>(x, [](int g) { return twi_u(4 + g); });
    gs_stage_n<8, 1, 1>(x, [](int g) { return twi_u(2 + g); });
    gs_stage_n<8, 0, 1>(x, [](int g) { return twi_u(1 + g); });
}

}  // namespace b200
