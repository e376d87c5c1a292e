// Per-thread phases of the batched identity key switch lvl1 -> lvl0 and of the bootstrap-free ops.
//
// Reference: IdentityKeySwitch<lvl10param>, TFHEpp include/keyswitch.hpp:11-52
//   res[n] = (b + 2^15) >> 16;  for each of the 1024 a_i: abar = a_i + 2^17, seven base-4 digits,
//   every non-zero digit g subtracts row ksk[i][j][g-1][0..n] (uint16 wrap).
// plus the MUX tail (gate.hpp:256-260): the two extracted lvl1 samples are added before the key
// switch and mu_lvl0 is added to b afterwards.
// Prior art: cuFHE KeySwitch (include/keyswitch_gpu.cuh:12-61), 1 thread per output coefficient.
//
// B200 mapping: one CTA per gate, thread k owns the 32-bit word holding output coefficients
// 2k and 2k+1; rows are 1280-byte aligned so a warp reads 128 contiguous bytes per row.
// The two uint16 lanes are accumulated in two 32-bit registers (low halves add exactly
// modulo 2^16 inside a 32-bit add), so a row costs one load and two integer ops per thread.
#pragma once
#include "fhe_params.h"
#include "hd.h"

namespace b200 {

constexpr int KS_THREADS = SLOT_WORDS;       // one 32-bit word of the output slot per thread: 320 threads x 2 uint16
                                             // coefficients (128-bit flavour), 512 threads x 1 uint32 coefficient (80-bit)
constexpr uint32_t KS_NONE = 0xFFFFFFFFu;

struct KsJob {
    uint32_t u0, u1;   // indices into the lvl1 scratch buffer; u1 = KS_NONE unless MUX
    uint32_t out;      // arena slot
    uint32_t post;     // added to b after the switch, modulo 2^16 (mu_lvl0 for MUX)
};

// phase 1: 14-bit digit code of coefficient i (bits 31..18 of a_i + 2^17)
B200_HD uint16_t ks_code(const uint32_t* ubuf, const KsJob& job, int i)
{
    uint32_t u = ubuf[(size_t)job.u0 * U_STRIDE + i];
    if (job.u1 != KS_NONE) u += ubuf[(size_t)job.u1 * U_STRIDE + i];
    return (uint16_t)((u + (1u << (32 - 1 - KS_T * KS_BASEBIT))) >> (32 - KS_T * KS_BASEBIT));
}

// keyswitch.hpp:27-39: 32 -> 16-bit rounding of b when the lvl0 torus is narrower, plain copy when it is 32 bits wide
B200_HD uint32_t ks_b_rounded(const uint32_t* ubuf, const KsJob& job)
{
    uint32_t b = ubuf[(size_t)job.u0 * U_STRIDE + N1];
    if (job.u1 != KS_NONE) b += ubuf[(size_t)job.u1 * U_STRIDE + N1];
    if (T0_BITS == 32) return b;
    return ((b + (1u << (31 - (T0_BITS & 31)))) >> (32 - (T0_BITS & 31))) & T0_MASK;
}

// one selected key row word into the accumulators: two uint16 lanes (the low halves add exactly modulo 2^16
// inside a 32-bit add) or one uint32 coefficient
B200_HD void ks_add(uint32_t w, uint32_t& lo, uint32_t& hi)
{
    lo += w;
    if (T0_BITS == 16) hi += w >> 16;
}

constexpr int KS_GROUPS = T0_BITS == 16 ? 3 : 2;  // row groups per CTA: group y walks coefficients i = y, y+G, ...
constexpr int KS_SPLIT = 4;   // CTAs per key switch on the narrow-frontier path (256 coefficients each)
constexpr int KS_SPLIT_MAX_GATES = 148;  // frontiers up to this size use the split path

// phase 2: thread (k, y) accumulates word k of every selected row of its coefficient group.
// ksk_words: [1024][7][3][320] uint32 (row padded to 640 uint16).  Returns the two partial sums.
B200_HD void ks_accumulate_group(const uint32_t* ksk_words, const uint16_t* codes, int k, int y, int ngroups,
                                 uint32_t& lo_out, uint32_t& hi_out)
{
    uint32_t lo = 0, hi = 0;
    for (int i = y; i < N1; i += ngroups) {
        const uint32_t code = codes[i];
        B200_UNROLL
        for (int j = 0; j < KS_T; j++) {
            const uint32_t g = (code >> (2 * (KS_T - 1 - j))) & 3u;
            const uint32_t* row = ksk_words + (size_t)((i * KS_T + j) * 3 + (g ? g - 1 : 0)) * KS_THREADS;
            ks_add(g ? row[k] : 0u, lo, hi);
        }
    }
    lo_out = lo;
    hi_out = hi;
}
// same walk restricted to coefficients [i0, i1), i = i0 + y, i0 + y + ngroups, ...: used when ONE key
// switch is split over several CTAs (narrow frontiers, where half the SMs would otherwise idle).
// `codes` holds the digit codes of that range only (codes[i - i0]).
B200_HD void ks_accumulate_range(const uint32_t* ksk_words, const uint16_t* codes, int k, int y, int ngroups, int i0,
                                 int i1, uint32_t& lo_out, uint32_t& hi_out)
{
    uint32_t lo = 0, hi = 0;
    for (int i = i0 + y; i < i1; i += ngroups) {
        const uint32_t code = codes[i - i0];
        B200_UNROLL
        for (int j = 0; j < KS_T; j++) {
            const uint32_t g = (code >> (2 * (KS_T - 1 - j))) & 3u;
            const uint32_t* row = ksk_words + (size_t)((i * KS_T + j) * 3 + (g ? g - 1 : 0)) * KS_THREADS;
            ks_add(g ? row[k] : 0u, lo, hi);
        }
    }
    lo_out = lo;
    hi_out = hi;
}
// final word from the summed partials
B200_HD uint32_t ks_finish(uint32_t lo, uint32_t hi, uint32_t b_rounded, uint32_t post, int k)
{
    uint32_t r_lo = 0u - lo, r_hi = 0u - hi;
    if (T0_BITS == 32) return k == N0 ? r_lo + b_rounded + post : r_lo;
    if (k == N0 / 2) r_lo += b_rounded + post;  // coefficient 636 = b lives in the low half of word 318
    return (r_lo & 0xFFFFu) | (r_hi << 16);
}

B200_HD uint32_t ks_accumulate(const uint32_t* ksk_words, const uint16_t* codes, uint32_t b_rounded,
                               uint32_t post, int k)
{
    uint32_t lo = 0, hi = 0;
    for (int i = 0; i < N1; i++) {
        const uint32_t code = codes[i];
        B200_UNROLL
        for (int j = 0; j < KS_T; j++) {
            const uint32_t g = (code >> (2 * (KS_T - 1 - j))) & 3u;
            const uint32_t* row = ksk_words + (size_t)((i * KS_T + j) * 3 + (g ? g - 1 : 0)) * KS_THREADS;
            ks_add(g ? row[k] : 0u, lo, hi);
        }
    }
    uint32_t r_lo = 0u - lo, r_hi = 0u - hi;
    if (k == N0 / 2) r_lo += b_rounded + post;  // coefficient 636 = b lives in the low half of word 318
    return (r_lo & 0xFFFFu) | (r_hi << 16);
}

// ---- wide frontiers: eight gates per CTA, one warp per gate (ks8_kernel) -----------------------------------
// ks_kernel reads every selected key row from L2 once per gate: 8192 gates x 5376 rows x 1280 B = 56 GB per batch, which
// is what bounds it (9.4 TB/s of L2 reads).  Here a warp owns a gate and walks the coefficients in the same order as the
// other seven warps of its CTA (and, loosely, the other CTAs of its SM), a block barrier every KS8_SYNC coefficients
// keeping them together: a key row one warp pulled into L1 serves the others, L2 traffic drops by the number of gates
// in step on an SM, and what is left is the two integer adds per 32-bit word.
// Lane l owns the word pairs {2(l + 32c), 2(l + 32c) + 1}, c < KS8_PAIRS: 64-bit loads, 256 contiguous bytes per warp.
constexpr int KS8_GATES = 8;
constexpr int KS8_SYNC = 16;
constexpr int KS8_PAIRS = KS_THREADS / 64;   // 5 (128-bit flavour), 8 (80-bit)
static_assert(KS_THREADS % 64 == 0, "a key row is a whole number of 64-bit words per lane");
struct Ks8Pair {
    uint32_t x, y;
};
B200_HD Ks8Pair ks8_load_pair(const uint32_t* row, int w)
{
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(row + w));
    return Ks8Pair{v.x, v.y};
#else
    return Ks8Pair{row[w], row[w + 1]};
#endif
}
// coefficients [i0, i1) of one gate; codes = the gate's digit codes.  The rows of KS8_BATCH digits are requested together
// and added afterwards (a warp that waited for each row in turn would spend an L2 round trip per row); a zero digit
// requests nothing and adds nothing (the digit is warp-uniform).
constexpr int KS8_BATCH = 3;
B200_HD void ks8_accumulate(const uint32_t* ksk_words, const uint16_t* codes, int i0, int i1, int lane,
                            uint32_t (&lo)[2 * KS8_PAIRS], uint32_t (&hi)[2 * KS8_PAIRS])
{
    for (int i = i0; i < i1; i++) {
        const uint32_t code = codes[i];
        B200_UNROLL
        for (int jb = 0; jb < KS_T; jb += KS8_BATCH) {
            Ks8Pair v[KS8_BATCH][KS8_PAIRS];
            uint32_t g[KS8_BATCH];
            B200_UNROLL
            for (int b = 0; b < KS8_BATCH; b++) {
                const int j = jb + b;
                if (j >= KS_T) continue;
                g[b] = (code >> (2 * (KS_T - 1 - j))) & 3u;
                if (!g[b]) continue;
                const uint32_t* row = ksk_words + (size_t)((i * KS_T + j) * 3 + (g[b] - 1)) * KS_THREADS;
                B200_UNROLL
                for (int c = 0; c < KS8_PAIRS; c++) v[b][c] = ks8_load_pair(row, 2 * (lane + 32 * c));
            }
            B200_UNROLL
            for (int b = 0; b < KS8_BATCH; b++) {
                if (jb + b >= KS_T || !g[b]) continue;
                B200_UNROLL
                for (int c = 0; c < KS8_PAIRS; c++) {
                    ks_add(v[b][c].x, lo[2 * c], hi[2 * c]);
                    ks_add(v[b][c].y, lo[2 * c + 1], hi[2 * c + 1]);
                }
            }
        }
    }
}
B200_HD void ks8_store(uint32_t* out_words, const uint32_t (&lo)[2 * KS8_PAIRS], const uint32_t (&hi)[2 * KS8_PAIRS],
                       uint32_t b_rounded, uint32_t post, int lane)
{
    B200_UNROLL
    for (int c = 0; c < KS8_PAIRS; c++) {
        const int w = 2 * (lane + 32 * c);
        out_words[w] = ks_finish(lo[2 * c], hi[2 * c], b_rounded, post, w);
        out_words[w + 1] = ks_finish(lo[2 * c + 1], hi[2 * c + 1], b_rounded, post, w + 1);
    }
}

// ---- bootstrap-free ops: NOT / COPY / CONST (gate.hpp:32-57) and the DFF tick (iyokan.hpp:1395-1402)
struct UnaryJob {
    uint32_t src, dst;
    uint32_t op;  // OP_NOT, OP_COPY, OP_CONST0, OP_CONST1
};
B200_HD uint32_t unary_word(const UnaryJob& job, const uint32_t* arena_words, int k)
{
    const uint32_t w = (job.op == OP_NOT || job.op == OP_COPY) ? arena_words[(size_t)job.src * KS_THREADS + k] : 0u;
    if (job.op == OP_COPY) return w;
    if (job.op == OP_NOT) {
        if (T0_BITS == 32) return 0u - w;
        const uint32_t lo = (0u - w) & 0xFFFFu, hi = (0u - (w >> 16)) & 0xFFFFu;
        return lo | (hi << 16);
    }
    const uint32_t b = (job.op == OP_CONST1) ? MU0 : ((0u - MU0) & T0_MASK);
    return (k == (T0_BITS == 32 ? N0 : N0 / 2)) ? b : 0u;
}

}  // namespace b200
