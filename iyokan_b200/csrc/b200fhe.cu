// b200fhe: sm_100a kernels + C ABI (include/b200fhe.h) of the batched TFHE gate back-end.
//
// Kernels (all hand-written CUDA, integer arithmetic, no tensor cores — the path is modular
// integer math; see DESIGN.md for the roofline of each):
//   br7_kernel<8>    blind rotation, 8 jobs / 16 warps per CTA (br7_phases.h)   — the hot kernel
//   br3_kernel<G>    blind rotation, G jobs / 2G warps per CTA, x3 interleave (br_phases.h)
//   br4_kernel       one job per CTA, br6_kernel one job per 2-CTA cluster: latency shapes for narrow levels
//   ks_kernel        sample-extracted lvl1 TLWE(s) -> lvl0 TLWE (ks_phases.h)
//   unary_kernel     NOT / COPY / CONST and the DFF tick gather-copy
//   bk_prep_kernel   raw TRGSW bootstrapping key -> NTT-domain 3-limb form (once per key)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <dlfcn.h>
#include <string>
#include <vector>

#include "../../include/b200fhe.h"
#ifndef B200FHE_80BIT
#include "br4_phases.h"
#include "br6_phases.h"
#include "br7_phases.h"
#ifdef B200FHE_WITH_BR9
#include "br9_phases.h"  // 4-point cluster experiment: measured slower than br6_kernel (profiles/r02_br9.md), not in the default build
#endif
#ifdef B200FHE_WITH_BR8
#include "br8_phases.h"  // quad-cluster experiment: measured slower than br6_kernel (profiles/r02_br8.md), not in the default build
#endif
#endif
#include "br_phases.h"
#include "brg_phases.h"
#include "gate_jobs.h"
#include "ks_phases.h"

using namespace b200;

// Per-phase cycle counters of the rotation kernels, debug builds only (-DB200FHE_PHASE_TIMING; scripts/gpu_phase_timing.py):
// the first and the last warp of CTA 0 add the cycles between consecutive marks.
#ifdef B200FHE_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[2][32];
__device__ unsigned long long g_cta_ns[1024][4];  // per CTA: kernel entry, loop start, loop end, exit (globaltimer)
#define CTA_STAMP(k)                                                                    \
    if (threadIdx.x == 0 && blockIdx.x < 1024) {                                        \
        unsigned long long _g;                                                          \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_g));                          \
        g_cta_ns[blockIdx.x][k] = _g;                                                   \
    }
// counters stay in registers (32-bit clock differences) and are written once by PHASE_FLUSH: a global read-modify-write
// per mark would itself cost more than most phases
#define PHASE_DECL                                                                                   \
    const int _pobs = (blockIdx.x == 0 && (threadIdx.x & 31) == 0)                                    \
                          ? (threadIdx.x == 0 ? 0 : ((int)threadIdx.x == (int)blockDim.x - 32 ? 1 : -1)) \
                          : -1;                                                                       \
    unsigned _pacc[24];                                                                               \
    const long long _pc0 = clock64();                                                                 \
    unsigned long long _pg0;                                                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_pg0));                                          \
    _Pragma("unroll") for (int _k = 0; _k < 24; _k++) _pacc[_k] = 0;                                  \
    unsigned _pt = (unsigned)clock();
#define PHASE_MARK(k)                              \
    do {                                           \
        const unsigned _n = (unsigned)clock();     \
        _pacc[k] += _n - _pt;                      \
        _pt = _n;                                  \
    } while (0)
#define PHASE_FLUSH                                                                          \
    if (_pobs >= 0) {                                                                        \
        _Pragma("unroll") for (int _k = 0; _k < 24; _k++) g_phase_cycles[_pobs][_k] += _pacc[_k]; \
        unsigned long long _pg1;                                                             \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_pg1));                             \
        g_phase_cycles[_pobs][30] += (unsigned long long)(clock64() - _pc0); /* loop cycles */ \
        g_phase_cycles[_pobs][31] += _pg1 - _pg0;                            /* loop ns */     \
    }
#else
#define PHASE_DECL
#define PHASE_MARK(k)
#define PHASE_FLUSH
#define CTA_STAMP(k)
#endif

// =====================================================================================
// kernels
// =====================================================================================

// Generic shape (brg_phases.h): any gadget length / limb count / lvl0 torus width; the kernel of the 80-bit flavour,
// also selectable at 128 bits (variant 1) where the parity tests pin it against the oracle.
template <int G>
__global__ void __launch_bounds__(64 * G, 1)
brg_kernel(const BrJob* __restrict__ jobs, int njobs, const torus0_t* __restrict__ arena,
           const uint32_t* __restrict__ bk_ntt, const tw_t* __restrict__ tw2f_g, const tw_t* __restrict__ tw2i_g,
           uint32_t* __restrict__ ubuf, int n_iter)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    BrgSmem<G> sm;
    sm.carve(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = warp >> 1, q = warp & 1;
    for (int k = tid; k < TW2_LEN; k += 64 * G) {
        sm.tw2f[k] = tw2f_g[k];
        sm.tw2i[k] = tw2i_g[k];
    }
    int job = blockIdx.x * G + g;
    const bool valid = job < njobs;
    if (!valid) job = njobs - 1;  // duplicate work, keeps every barrier uniform
    const BrJob jb = jobs[job];
    uint32_t accr[32];
    brg_prologue<G>(sm, jb, arena, g, q, lane, accr);
    __syncthreads();
    for (int i = 0; i < n_iter; i++) {
        const uint32_t* bk_i = bk_ntt + (size_t)i * BK_COLS * ROWS * N1;
        {
            uint32_t dreg[32];
            brg_rotate_diff<G>(sm, i, g, q, lane, accr, dreg);
#pragma unroll 1
            for (int d = 0; d < GL; d++) {
                brg_fwd_a<G>(sm, g, q, lane, d, dreg);
                __syncwarp();
                brg_fwd_b<G>(sm, g, q, lane, d);
            }
        }
        __syncthreads();
        brg_pointwise<G>(sm, bk_i, tid);
        __syncthreads();
        {
            uint32_t sum[32];
#pragma unroll 1
            for (int l = 0; l < LIMBS; l++) {
                brg_inv_a<G>(sm, g, q, lane, l);
                __syncwarp();
                brg_inv_b<G>(sm, g, q, lane, l, sum);
            }
            brg_acc_update<G>(sm, g, q, lane, sum, accr);
        }
        __syncwarp();
    }
    if (valid) brg_epilogue<G>(sm, g, q, lane, ubuf + (size_t)job * U_STRIDE);
}

#ifndef B200FHE_80BIT
// 12-warp throughput shape: a warp owns one accumulator polynomial and runs its three transforms in lock step
// (ct_stage3 / gs_stage3): 3x the ILP per warp and one twiddle fetch for three butterflies
template <int G>
__global__ void __launch_bounds__(64 * G, 1)
br3_kernel(const BrJob* __restrict__ jobs, int njobs, const torus0_t* __restrict__ arena,
           const uint32_t* __restrict__ bk_ntt, const tw_t* __restrict__ tw2f_g, const tw_t* __restrict__ tw2i_g,
           uint32_t* __restrict__ ubuf, int n_iter)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    BrSmem<G> sm;
    sm.carve(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = warp >> 1, q = warp & 1;
    for (int k = tid; k < TW2_LEN; k += 64 * G) {
        sm.tw2f[k] = tw2f_g[k];
        sm.tw2i[k] = tw2i_g[k];
    }
    int job = blockIdx.x * G + g;
    const bool valid = job < njobs;
    if (!valid) job = njobs - 1;
    const BrJob jb = jobs[job];
    uint32_t accr[32];
    br_prologue<G>(sm, jb, arena, g, q, lane, accr);
    __syncthreads();
    for (int i = 0; i < n_iter; i++) {
        const uint32_t* bk_i = bk_ntt + (size_t)i * BK_COLS * ROWS * N1;
        uint32_t bk0[BK_COLS][ROWS];
        {
            uint32_t x0[32], x1[32], x2[32];
            br_fwd3_a<G>(sm, i, g, q, lane, accr, x0, x1, x2);
            __syncwarp();
            br_fwd3_b<G>(sm, g, q, lane, x0, x1, x2);
        }
        __syncwarp();
        br_fwd3_c<G>(sm, g, q, lane);
        pw_load(bk_i, tid, bk0);
        __syncthreads();
        br_pointwise<G>(sm, bk_i, tid, bk0);
        __syncthreads();
        br_inv3_a<G>(sm, g, q, lane);
        __syncwarp();
        br_inv3_b<G>(sm, g, q, lane, accr);
        __syncwarp();
        br_inv3_c<G>(sm, g, q, lane, accr);
        __syncwarp();
    }
    if (valid) br_epilogue<G>(sm, g, q, lane, ubuf + (size_t)job * U_STRIDE);
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 16-warp throughput shape (br7_phases.h): eight jobs per CTA on XOR-swizzled tiles, two transforms of a
// warp in lock step and the third alone so that 512 threads fit the register file (<= 128 each).
// The CTA is cut into G/J barrier groups of J jobs; group k starts k/(G/J) of a step late (skew_cycles apart).
template <int G, int J>
__global__ void __launch_bounds__(64 * G, 1)
br7_kernel(const BrJob* __restrict__ jobs, int njobs, const torus0_t* __restrict__ arena,
           const uint32_t* __restrict__ bk_ntt, const tw_t* __restrict__ tw2f_g, const tw_t* __restrict__ tw2i_g,
           const uint32_t* __restrict__ r4_g, uint32_t* __restrict__ ubuf, int n_iter, int skew_cycles)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Br7Smem<G> sm;
    sm.carve(smem_raw);
    CTA_STAMP(0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = warp >> 1, q = warp & 1;
    const int group = g / J, g0 = group * J, tig = tid - 64 * g0;
    for (int k = tid; k < TW2_LEN; k += 64 * G) {
        sm.tw2f[k] = tw2f_g[k];
        sm.tw2i[k] = tw2i_g[k];
    }
    for (int k = tid; k < R4_WORDS; k += 64 * G) sm.r4[k] = r4_g[k];
    int job = blockIdx.x * G + g;
    const bool valid = job < njobs;
    if (!valid) job = njobs - 1;  // duplicate work, keeps every barrier uniform
    const BrJob jb = jobs[job];
    uint32_t accr[32];
    br7_prologue<G>(sm, jb, arena, g, q, lane, accr);
    __syncthreads();
    if (J < G && group > 0 && skew_cycles > 0) {
        const long long t0 = clock64(), wait = (long long)group * skew_cycles;
        while (clock64() - t0 < wait) {}
    }
    PHASE_DECL
    CTA_STAMP(1);
    for (int i = 0; i < n_iter; i++) {
        const uint32_t* bk_i = bk_ntt + (size_t)i * BK_COLS * ROWS * N1;
        uint32_t bk0[BK_COLS][ROWS];
        {  // one rotated difference, kept in registers: digit 0 first (x1 passes leave room for it), then digits 1 and 2
            uint32_t dv[32];
            {
                uint32_t x0[32];
                br7_fwd0_a<G>(sm, i, g, q, lane, accr, dv, x0);
                __syncwarp();  // last read of the accumulator copy that shares tile 3q with digit 0
                br7_fwd0_b<G>(sm, g, q, lane, x0);
            }
            __syncwarp();
            PHASE_MARK(0);
            br7_fwd0_c<G>(sm, g, q, lane);
            PHASE_MARK(1);
            br7_fwd12_a<G>(sm, g, q, lane, dv);
            __syncwarp();
            PHASE_MARK(2);
        }
        br7_fwd12_c<G>(sm, g, q, lane);
        PHASE_MARK(3);
        pw_load(bk_i, tig, bk0);  // key words of the pointwise stage in flight across the barrier
        if (J == G) __syncthreads(); else named_barrier_sync(1 + group, 64 * J);
        PHASE_MARK(4);
        br7_pointwise<G, J>(sm, bk_i, g0, tig, bk0);
        PHASE_MARK(5);
        if (J == G) __syncthreads(); else named_barrier_sync(1 + group, 64 * J);
        PHASE_MARK(6);
        br7_inv01_a<G>(sm, g, q, lane);
        __syncwarp();
        PHASE_MARK(7);
        br7_inv01_b<G>(sm, g, q, lane, accr);
        PHASE_MARK(8);
        br7_inv2_a<G>(sm, g, q, lane);
        __syncwarp();
        PHASE_MARK(9);
        br7_inv2_b<G>(sm, g, q, lane, accr);
        __syncwarp();
        PHASE_MARK(10);
    }
    PHASE_FLUSH
    CTA_STAMP(2);
    if (valid) br7_epilogue<G>(sm, g, q, lane, ubuf + (size_t)job * U_STRIDE);
    CTA_STAMP(3);
}


// ---- latency shape: one job per CTA, 6 teams of 64 threads, key staged by bulk-async copies ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// one thread: arm the barrier with the byte count and issue the bulk copies of bk_ntt[i] (6 x 24 KB)
__device__ __forceinline__ void key_stage_issue(const Br4Smem& sm, const uint32_t* bk_i)
{
    constexpr uint32_t CHUNK = ROWS * N1 * 4;  // one output column: 24,576 B
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the buffer precede the async writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(sm.mbar)),
                 "r"((uint32_t)(BK_COLS * CHUNK))
                 : "memory");
#pragma unroll
    for (int c = 0; c < BK_COLS; c++)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sm.keyb + (size_t)c * ROWS * N1)),
                     "l"(bk_i + (size_t)c * ROWS * N1), "r"(CHUNK), "r"(smem_u32(sm.mbar))
                     : "memory");
}

__global__ void __launch_bounds__(BR4_THREADS, 1)
br4_kernel(const BrJob* __restrict__ jobs, int njobs, const torus0_t* __restrict__ arena,
           const uint32_t* __restrict__ bk_ntt, const BlockTw* __restrict__ tw_g, uint32_t* __restrict__ ubuf, int n_iter)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Br4Smem sm;
    sm.carve(smem_raw);
    const int tid = threadIdx.x, team = tid >> 6, t = tid & 63, q = team / GL, d = team % GL;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tw_g);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sm.tw);
        for (int k = tid; k < (int)(sizeof(BlockTw) / 4); k += BR4_THREADS) dst[k] = src[k];
    }
    if (tid == 0) mbar_init(sm.mbar, 1);
    const int job = blockIdx.x;
    const BrJob jb = jobs[job];
    br4_prologue(sm, jb, arena, tid);
    __syncthreads();
    if (tid == 0 && n_iter > 0) key_stage_issue(sm, bk_ntt);

    for (int i = 0; i < n_iter; i++) {
        br4_fwd_p1(sm, i, q, d, t);
        named_barrier_sync(1 + team, TEAM_THREADS);
        br4_fwd_p2(sm, q, d, t);
        named_barrier_sync(1 + team, TEAM_THREADS);
        br4_fwd_p3(sm, q, d, t);
        mbar_wait(sm.mbar, (uint32_t)(i & 1));  // key of step i has landed
        __syncthreads();
        br4_pointwise_item(sm, tid);
        br4_pointwise_item(sm, tid + BR4_THREADS);
        __syncthreads();
        if (tid == 0 && i + 1 < n_iter) key_stage_issue(sm, bk_ntt + (size_t)(i + 1) * BR4_KEY_WORDS);
        br4_inv_pA(sm, q, d, t);
        named_barrier_sync(1 + team, TEAM_THREADS);
        br4_inv_pB(sm, q, d, t);
        named_barrier_sync(1 + team, TEAM_THREADS);
        br4_inv_pC(sm, q, d, t);
        named_barrier_sync(7 + q, GL * TEAM_THREADS);  // the three limb teams of polynomial q
    }
    __syncthreads();
    br4_epilogue(sm, tid, ubuf + (size_t)job * U_STRIDE);
}


// ---- cluster helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// arrival that publishes no data (only "I have finished reading"): no fence, no store drain
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
template <class T>
__device__ __forceinline__ T* map_to_cta(T* p, uint32_t rank)  // generic address of p in CTA `rank` of the cluster
{
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((uint64_t)p), "r"(rank));
    return reinterpret_cast<T*>(out);
}

// ---- fine-grained cluster shape: one job per 2-CTA cluster, 3 teams of 128 threads per CTA (br6_phases.h) ----
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BR6_THREADS, 1)
br6_kernel(const BrJob* __restrict__ jobs, int njobs, const torus0_t* __restrict__ arena,
           const uint32_t* __restrict__ bk_ntt, const Block8Tw* __restrict__ tw_g, uint32_t* __restrict__ ubuf, int n_iter)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Br6Smem sm;
    sm.carve(smem_raw);
    const int tid = threadIdx.x, d = tid >> 7, t = tid & 127;
    const int q = (int)cluster_ctarank();
    const int job = blockIdx.x >> 1;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tw_g);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sm.tw);
        for (int k = tid; k < (int)(sizeof(Block8Tw) / 4); k += BR6_THREADS) dst[k] = src[k];
    }
    uint64_t* mbar_dig = sm.mbar + 1;  // counts the bytes of the three digit tiles the peer copies in per step
    if (tid == 0) {
        mbar_init(sm.mbar, 1);
        mbar_init(mbar_dig, 1);
    }
    const BrJob jb = jobs[job];
    br6_prologue(sm, jb, arena, q, tid);
    __syncthreads();
    const uint32_t* key0 = bk_ntt + (size_t)q * BR6_KEY_WORDS;  // columns 3q..3q+2 of step 0
    auto stage = [&](int i) {
        constexpr uint32_t BYTES = BR6_KEY_WORDS * 4;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(sm.mbar)), "r"(BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sm.keyb)),
                     "l"(key0 + (size_t)i * BR4_KEY_WORDS), "r"(BYTES), "r"(smem_u32(sm.mbar))
                     : "memory");
    };
    constexpr uint32_t TILE_BYTES = B8_WORDS * 4;
    const uint32_t my_tile = smem_u32(sm.in_tile(q * GL + d));
    uint32_t peer_tile, peer_bar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_tile) : "r"(my_tile), "r"((uint32_t)(q ^ 1)));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_bar) : "r"(smem_u32(mbar_dig)), "r"((uint32_t)(q ^ 1)));
    if (tid == 0 && n_iter > 0) stage(0);
    cluster_arrive();  // both CTAs have initialised their barriers before anyone copies into the other
    cluster_wait();
    cluster_arrive_relaxed();  // phase (B) of "step -1"

    PHASE_DECL
    for (int i = 0; i < n_iter; i++) {
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar_dig)),
                         "r"(GL * TILE_BYTES)
                         : "memory");
        br6_fwd_p1(sm, i, q, d, t);
        PHASE_MARK(0);
        named_barrier_sync(1 + d, TEAM8_THREADS);
        PHASE_MARK(1);
        br6_fwd_p2(sm, q, d, t);
        PHASE_MARK(2);
        named_barrier_sync(1 + d, TEAM8_THREADS);
        PHASE_MARK(3);
        br6_fwd_p3(sm, q, d, t);
        PHASE_MARK(4);
        named_barrier_sync(1 + d, TEAM8_THREADS);
        PHASE_MARK(5);
        br6_fwd_p4(sm, q, d, t);
        PHASE_MARK(6);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cluster_wait();    // (B) the peer's pointwise stage of the previous step no longer reads my copies
        PHASE_MARK(7);
        named_barrier_sync(1 + d, TEAM8_THREADS);
        PHASE_MARK(8);
        if (t == 0)
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             peer_tile),
                         "r"(my_tile), "r"(TILE_BYTES), "r"(peer_bar)
                         : "memory");
        mbar_wait(sm.mbar, (uint32_t)(i & 1));
        PHASE_MARK(9);
        uint64_t pacc[LIMBS][4];
        named_barrier_sync(7, BR6_THREADS);  // the local teams' tiles are complete
        PHASE_MARK(10);
        br6_pw_local(sm, q, tid, pacc);
        PHASE_MARK(11);
        mbar_wait(mbar_dig, (uint32_t)(i & 1));  // the peer's three tiles have landed
        PHASE_MARK(12);
        br6_pw_finish(sm, q, tid, pacc);
        PHASE_MARK(13);
        cluster_arrive_relaxed();  // (B) for the next step
        __syncthreads();
        PHASE_MARK(14);
        if (tid == 0 && i + 1 < n_iter) stage(i + 1);
        br6_inv_pA(sm, d, t);
        PHASE_MARK(15);
        named_barrier_sync(1 + d, TEAM8_THREADS);
        PHASE_MARK(16);
        br6_inv_pB(sm, d, t);
        PHASE_MARK(17);
        named_barrier_sync(1 + d, TEAM8_THREADS);
        PHASE_MARK(18);
        br6_inv_pC(sm, d, t);
        PHASE_MARK(19);
        named_barrier_sync(1 + d, TEAM8_THREADS);
        PHASE_MARK(20);
        br6_inv_pD(sm, d, t);
        PHASE_MARK(21);
        __syncthreads();
        PHASE_MARK(22);
    }
    PHASE_FLUSH
    cluster_wait();
    br6_epilogue(sm, q, tid, ubuf + (size_t)job * U_STRIDE);
}

#ifdef B200FHE_WITH_BR9
// 4-point cluster shape (br9_phases.h): the same cluster protocol as br6_kernel, transforms in five two-stage passes of
// 256 threads - 24 warps per SM
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BR9_THREADS, 1)
br9_kernel(const BrJob* __restrict__ jobs, int njobs, const torus0_t* __restrict__ arena,
           const uint32_t* __restrict__ bk_ntt, const Block4Tw* __restrict__ tw_g, uint32_t* __restrict__ ubuf, int n_iter)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Br9Smem sm;
    sm.carve(smem_raw);
    const int tid = threadIdx.x, d = tid >> 8, t = tid & 255;
    const int q = (int)cluster_ctarank();
    const int job = blockIdx.x >> 1;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tw_g);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sm.tw);
        for (int k = tid; k < (int)(sizeof(Block4Tw) / 4); k += BR9_THREADS) dst[k] = src[k];
    }
    uint64_t* mbar_dig = sm.mbar + 1;  // counts the bytes of the three digit tiles the peer copies in per step
    if (tid == 0) {
        mbar_init(sm.mbar, 1);
        mbar_init(mbar_dig, 1);
    }
    const BrJob jb = jobs[job];
    br9_prologue(sm, jb, arena, q, tid);
    __syncthreads();
    const uint32_t* key0 = bk_ntt + (size_t)q * BR6_KEY_WORDS;  // columns 3q..3q+2 of step 0
    auto stage = [&](int i) {
        constexpr uint32_t BYTES = BR6_KEY_WORDS * 4;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(sm.mbar)), "r"(BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sm.keyb)),
                     "l"(key0 + (size_t)i * BR4_KEY_WORDS), "r"(BYTES), "r"(smem_u32(sm.mbar))
                     : "memory");
    };
    constexpr uint32_t TILE_BYTES = B8_WORDS * 4;
    const uint32_t my_tile = smem_u32(sm.in_tile(q * GL + d));
    uint32_t peer_tile, peer_bar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_tile) : "r"(my_tile), "r"((uint32_t)(q ^ 1)));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_bar) : "r"(smem_u32(mbar_dig)), "r"((uint32_t)(q ^ 1)));
    if (tid == 0 && n_iter > 0) stage(0);
    cluster_arrive();  // both CTAs have initialised their barriers before anyone copies into the other
    cluster_wait();
    cluster_arrive_relaxed();  // phase (B) of "step -1"

    for (int i = 0; i < n_iter; i++) {
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar_dig)),
                         "r"(GL * TILE_BYTES)
                         : "memory");
        br9_fwd_p1(sm, i, q, d, t);
        named_barrier_sync(1 + d, TEAM4_THREADS);
        br9_fwd_p2(sm, q, d, t);
        named_barrier_sync(1 + d, TEAM4_THREADS);
        br9_fwd_p3(sm, q, d, t);
        named_barrier_sync(1 + d, TEAM4_THREADS);
        br9_fwd_p4(sm, q, d, t);
        named_barrier_sync(1 + d, TEAM4_THREADS);
        br9_fwd_p5(sm, q, d, t);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cluster_wait();    // (B) the peer's pointwise stage of the previous step no longer reads my copies
        named_barrier_sync(1 + d, TEAM4_THREADS);
        if (t == 0)
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             peer_tile),
                         "r"(my_tile), "r"(TILE_BYTES), "r"(peer_bar)
                         : "memory");
        mbar_wait(sm.mbar, (uint32_t)(i & 1));
        uint64_t pacc[LIMBS][4];
        named_barrier_sync(7, BR9_THREADS);  // the local teams' tiles are complete
        br6_pw_local(sm, q, tid, pacc);
        mbar_wait(mbar_dig, (uint32_t)(i & 1));  // the peer's three tiles have landed
        br6_pw_finish(sm, q, tid, pacc);
        cluster_arrive_relaxed();  // (B) for the next step
        __syncthreads();
        if (tid == 0 && i + 1 < n_iter) stage(i + 1);
        br9_inv_pA(sm, d, t);
        named_barrier_sync(1 + d, TEAM4_THREADS);
        br9_inv_pB(sm, d, t);
        named_barrier_sync(1 + d, TEAM4_THREADS);
        br9_inv_pC(sm, d, t);
        named_barrier_sync(1 + d, TEAM4_THREADS);
        br9_inv_pD(sm, d, t);
        named_barrier_sync(1 + d, TEAM4_THREADS);
        br9_inv_pE(sm, d, t);
        __syncthreads();
    }
    cluster_wait();
    br9_epilogue(sm, q, tid, ubuf + (size_t)job * U_STRIDE);
}
#endif  // B200FHE_WITH_BR9

#ifdef B200FHE_WITH_BR8
// ---- quad-cluster shape: one job per 4-CTA cluster, CTA (q, h) = polynomial q, transform half h (br8_phases.h) ----
__device__ __forceinline__ uint32_t map_shared_rank(uint32_t smem_addr, uint32_t rank)
{
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_addr), "r"(rank));
    return out;
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(BR8_THREADS, 1)
br8_kernel(const BrJob* __restrict__ jobs, int njobs, const torus0_t* __restrict__ arena,
           const uint32_t* __restrict__ bk_ntt, const tw_t* __restrict__ twf_g, const tw_t* __restrict__ twi_g,
           uint32_t* __restrict__ ubuf, int n_iter)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Br8Smem sm;
    sm.carve(smem_raw);
    const int tid = threadIdx.x, d = tid >> 6, t = tid & 63;
    const uint32_t rank = cluster_ctarank();
    const int q = (int)(rank >> 1), h = (int)(rank & 1);
    const int job = blockIdx.x >> 2;
    for (int k = tid; k < 1024; k += BR8_THREADS) {
        sm.twf[k] = twf_g[k];
        sm.twi[k] = twi_g[k];
    }
    if (tid == 0)
        for (int b = 0; b < 5; b++) mbar_init(sm.mbar + b, 1);
    const BrJob jb = jobs[job];
    br8_prologue(sm, jb, arena, q, tid);
    __syncthreads();
    // this CTA's quarter of a step's key: columns (q, l), rows r, positions [512h, 512h + 512): 18 chunks of 2 KB
    auto stage = [&](int i) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect(sm.mbar, BR8_KEY_WORDS * 4);
        const uint32_t* src = bk_ntt + (size_t)i * BR4_KEY_WORDS + (size_t)q * LIMBS * ROWS * N1 + 512 * h;
#pragma unroll
        for (int c = 0; c < LIMBS * ROWS; c++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(sm.keyb + c * 512)),
                         "l"(src + (size_t)c * N1), "r"(2048u), "r"(smem_u32(sm.mbar))
                         : "memory");
    };
    constexpr uint32_t HB = 512 * 4;  // payload of a half tile (the 64 padding words are never sent)
    const uint32_t rA = rank ^ 2u, rB = rank ^ 1u;  // partner for the digit tiles / for the halves of this polynomial
    Br8Peer peerA[2], peerB, peerC;
    peerA[0].bar = map_shared_rank(smem_u32(sm.mbar + 1), rA);
    peerA[1].bar = map_shared_rank(smem_u32(sm.mbar + 2), rA);
    peerA[0].dst = map_shared_rank(smem_u32(sm.peer + (size_t)d * H_WORDS), rA);
    peerA[1].dst = map_shared_rank(smem_u32(sm.peer + (size_t)(GL + d) * H_WORDS), rA);
    peerB.bar = map_shared_rank(smem_u32(sm.mbar + 3), rB);
    peerB.dst = map_shared_rank(smem_u32(sm.half2 + (size_t)d * H_WORDS), rB);
    peerC.bar = map_shared_rank(smem_u32(sm.mbar + 4), rB);
    peerC.dst = map_shared_rank(smem_u32(sm.accb + 512 * h), rB);
    if (tid == 0 && n_iter > 0) stage(0);
    cluster_arrive();  // every CTA has initialised its barriers before anyone stores into it
    cluster_wait();

    for (int i = 0; i < n_iter; i++) {
        const int par = i & 1;
        if (i > 0) mbar_wait(sm.mbar + 4, (uint32_t)((i - 1) & 1));  // the other half of the accumulator, step i-1
        if (tid == 0) {
            mbar_expect(sm.mbar + 1 + par, GL * HB);
            mbar_expect(sm.mbar + 3, LIMBS * HB);
            mbar_expect(sm.mbar + 4, 2048u);
        }
        uint32_t* mydig = sm.dig + (size_t)(par * GL + d) * H_WORDS;
        br8_fwd_p1(sm, i, h, d, t);
        named_barrier_sync(1 + d, BR8_TEAM);
        br8_fwd_p2(mydig, sm.twf, h, t);
        named_barrier_sync(1 + d, BR8_TEAM);
        br8_fwd_p3(mydig, sm.twf, h, t, peerA[par]);               // results also go to CTA (1-q, h) as they are produced
        mbar_wait(sm.mbar, (uint32_t)par);                          // key quarter of step i
        mbar_wait(sm.mbar + 1 + par, (uint32_t)((i >> 1) & 1));     // digit tiles of the other polynomial
        __syncthreads();                                            // the local teams' tiles are complete
        br8_pointwise(sm, q, par, tid);
        __syncthreads();
        if (tid == 0 && i + 1 < n_iter) stage(i + 1);
        uint32_t* myout = sm.outb + (size_t)d * H_WORDS;
        br8_inv_pA(myout, sm.twi, h, t);
        named_barrier_sync(1 + d, BR8_TEAM);
        br8_inv_pB(myout, sm.twi, h, t);
        named_barrier_sync(1 + d, BR8_TEAM);
        br8_inv_pC(myout, h, t, peerB);                             // results also go to CTA (q, 1-h)
        mbar_wait(sm.mbar + 3, (uint32_t)par);                      // the other half of the three limb columns
        br8_inv_join(sm, h, d, t);                                  // (own values: each thread re-reads what it wrote)
        __syncthreads();                                            // all three limbs are in the accumulator
        br8_send_acc(sm, h, tid, peerC);
    }
    if (n_iter > 0) mbar_wait(sm.mbar + 4, (uint32_t)((n_iter - 1) & 1));
    cluster_arrive();  // nobody exits while a peer's copy into it may be in flight
    cluster_wait();
    br8_epilogue(sm, q, h, tid, ubuf + (size_t)job * U_STRIDE);
}

#endif  // B200FHE_WITH_BR8
#endif  // !B200FHE_80BIT

__global__ void __launch_bounds__(KS_THREADS * KS_GROUPS)
ks_kernel(const KsJob* __restrict__ jobs, const uint32_t* __restrict__ ubuf,
          const uint32_t* __restrict__ ksk_words, torus0_t* __restrict__ arena)
{
    __shared__ uint16_t codes[N1];
    __shared__ uint32_t b_sh;
    __shared__ uint32_t part[KS_GROUPS - 1][2][KS_THREADS];
    const KsJob job = jobs[blockIdx.x];
    const int k = threadIdx.x, y = threadIdx.y, tid = y * KS_THREADS + k;
    for (int i = tid; i < N1; i += KS_THREADS * KS_GROUPS) codes[i] = ks_code(ubuf, job, i);
    if (tid == 0) b_sh = ks_b_rounded(ubuf, job);
    __syncthreads();
    uint32_t lo, hi;
    ks_accumulate_group(ksk_words, codes, k, y, KS_GROUPS, lo, hi);
    if (y > 0) {
        part[y - 1][0][k] = lo;
        part[y - 1][1][k] = hi;
    }
    __syncthreads();
    if (y == 0) {
#pragma unroll
        for (int g = 0; g < KS_GROUPS - 1; g++) {
            lo += part[g][0][k];
            hi += part[g][1][k];
        }
        reinterpret_cast<uint32_t*>(arena + (size_t)job.out * SLOT_STRIDE)[k] = ks_finish(lo, hi, b_sh, job.post, k);
    }
}

// Wide frontiers: eight gates per CTA, one warp per gate, key rows shared through L1 (ks_phases.h)
__global__ void __launch_bounds__(32 * KS8_GATES, 4)
ks8_kernel(const KsJob* __restrict__ jobs, int njobs, const uint32_t* __restrict__ ubuf,
           const uint32_t* __restrict__ ksk_words, torus0_t* __restrict__ arena)
{
    __shared__ uint16_t codes[KS8_GATES][N1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gate = blockIdx.x * KS8_GATES + warp;
    const bool valid = gate < njobs;
    const KsJob job = jobs[valid ? gate : njobs - 1];  // surplus warps repeat the last gate: barriers stay uniform
    for (int i = lane; i < N1; i += 32) codes[warp][i] = ks_code(ubuf, job, i);
    const uint32_t b = ks_b_rounded(ubuf, job);
    uint32_t lo[2 * KS8_PAIRS], hi[2 * KS8_PAIRS];
#pragma unroll
    for (int k = 0; k < 2 * KS8_PAIRS; k++) lo[k] = hi[k] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < N1; i0 += KS8_SYNC) {
        ks8_accumulate(ksk_words, codes[warp], i0, i0 + KS8_SYNC, lane, lo, hi);
        __syncthreads();
    }
    if (valid) ks8_store(reinterpret_cast<uint32_t*>(arena + (size_t)job.out * SLOT_STRIDE), lo, hi, b, job.post, lane);
}

// Narrow frontiers: one key switch split over KS_SPLIT CTAs (256 coefficients each) + a tiny combine
// kernel, so that a level of a few dozen gates uses all SMs for its key switch as well.
__global__ void __launch_bounds__(KS_THREADS * KS_GROUPS)
ks_split_kernel(const KsJob* __restrict__ jobs, const uint32_t* __restrict__ ubuf,
                const uint32_t* __restrict__ ksk_words, uint32_t* __restrict__ partial)
{
    constexpr int SPAN = N1 / KS_SPLIT;
    __shared__ uint16_t codes[SPAN];
    __shared__ uint32_t part[KS_GROUPS - 1][2][KS_THREADS];
    const int gate = blockIdx.x / KS_SPLIT, piece = blockIdx.x % KS_SPLIT, i0 = piece * SPAN;
    const KsJob job = jobs[gate];
    const int k = threadIdx.x, y = threadIdx.y, tid = y * KS_THREADS + k;
    for (int i = tid; i < SPAN; i += KS_THREADS * KS_GROUPS) codes[i] = ks_code(ubuf, job, i0 + i);
    __syncthreads();
    uint32_t lo, hi;
    ks_accumulate_range(ksk_words, codes, k, y, KS_GROUPS, i0, i0 + SPAN, lo, hi);
    if (y > 0) {
        part[y - 1][0][k] = lo;
        part[y - 1][1][k] = hi;
    }
    __syncthreads();
    if (y == 0) {
#pragma unroll
        for (int g = 0; g < KS_GROUPS - 1; g++) {
            lo += part[g][0][k];
            hi += part[g][1][k];
        }
        uint32_t* out = partial + (size_t)blockIdx.x * 2 * KS_THREADS;
        out[k] = lo;
        out[KS_THREADS + k] = hi;
    }
}
__global__ void __launch_bounds__(KS_THREADS)
ks_combine_kernel(const KsJob* __restrict__ jobs, const uint32_t* __restrict__ ubuf,
                  const uint32_t* __restrict__ partial, torus0_t* __restrict__ arena)
{
    const KsJob job = jobs[blockIdx.x];
    const int k = threadIdx.x;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int p = 0; p < KS_SPLIT; p++) {
        const uint32_t* in = partial + (size_t)(blockIdx.x * KS_SPLIT + p) * 2 * KS_THREADS;
        lo += in[k];
        hi += in[KS_THREADS + k];
    }
    reinterpret_cast<uint32_t*>(arena + (size_t)job.out * SLOT_STRIDE)[k] =
        ks_finish(lo, hi, ks_b_rounded(ubuf, job), job.post, k);
}

// Reads every source before any destination is written (two launches: gather, scatter) so that
// a DFF chain Q1 <- D1 = Q0 ticks correctly even when src and dst sets overlap.
__global__ void __launch_bounds__(KS_THREADS)
unary_gather_kernel(const UnaryJob* __restrict__ jobs, const uint32_t* __restrict__ arena_words,
                    uint32_t* __restrict__ stage_words)
{
    const UnaryJob job = jobs[blockIdx.x];
    stage_words[(size_t)blockIdx.x * KS_THREADS + threadIdx.x] = unary_word(job, arena_words, threadIdx.x);
}
__global__ void __launch_bounds__(KS_THREADS)
unary_scatter_kernel(const UnaryJob* __restrict__ jobs, const uint32_t* __restrict__ stage_words,
                     uint32_t* __restrict__ arena_words)
{
    const UnaryJob job = jobs[blockIdx.x];
    arena_words[(size_t)job.dst * KS_THREADS + threadIdx.x] = stage_words[(size_t)blockIdx.x * KS_THREADS + threadIdx.x];
}

constexpr int BKPREP_WARPS = 4;
__global__ void __launch_bounds__(32 * BKPREP_WARPS)
bk_prep_kernel(const uint32_t* __restrict__ bk_raw, uint32_t* __restrict__ bk_ntt, const tw_t* __restrict__ tw2f,
               tw_t scale, int ntasks)
{
    __shared__ __align__(16) uint32_t tiles[BKPREP_WARPS][TILE_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * BKPREP_WARPS + warp;
    if (task >= ntasks) return;  // whole warp exits together; only __syncwarp below
    const int l = task % LIMBS, q = (task / LIMBS) % 2, r = (task / (2 * LIMBS)) % ROWS, i = task / (2 * LIMBS * ROWS);
    const uint32_t* raw = bk_raw + ((size_t)(i * ROWS + r) * 2 + q) * N1;
    uint32_t* out = bk_ntt + ((size_t)(i * BK_COLS + q * LIMBS + l) * ROWS + r) * N1;
    brg_bk_prep_a(raw, l, lane, tiles[warp]);
    __syncwarp();
    bk_prep_b(tiles[warp], tw2f, scale, lane, out);
}

// test hook helpers: c [n][637] dense -> arena-like [n][640]
__global__ void pad_tlwe0_kernel(const torus0_t* __restrict__ dense, torus0_t* __restrict__ padded, size_t n)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * SLOT_STRIDE) return;
    const size_t s = idx / SLOT_STRIDE, k = idx % SLOT_STRIDE;
    padded[idx] = k < TLWE0_LEN ? dense[s * TLWE0_LEN + k] : (torus0_t)0;
}

// =====================================================================================
// host side
// =====================================================================================

static thread_local std::string g_err;
static int fail(const std::string& msg)
{
    g_err = msg;
    return 1;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" +  \
                        std::to_string(__LINE__) + ")");                                           \
    } while (0)

constexpr int BR_MAX_SEGMENTS = 5;  // launches one frontier's blind rotations are cut into, at most

struct b200fhe_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
#ifdef B200FHE_80BIT
    int G = 4;
    int variant = 1;
#else
    int G = 8;
    int variant = 7;  // pinned shape when autotune is off (b200fhe_set_kernel_variant / _jobs_per_cta)
#endif
    bool autotune = true;  // pick (variant, G) per batch size unless the caller pinned them
    int br7_group = 8;     // jobs per barrier group of br7_kernel (8 = CTA-wide pointwise stage, 4 or 2 = skewed groups)
    int br7_skew = 0;      // start delay between consecutive groups, SM cycles
    // frontiers of at least this many key switches take ks8_kernel (8 gates per CTA).  80-bit flavour: never - its 2048-byte
    // rows double the registers per lane and the shape measured slower than ks_kernel (11.5 against 9.3 ms per 8192 gates)
    int ks8_min = T0_BITS == 16 ? 1400 : 0x7fffffff;
    NttTables* tab = nullptr;
    tw_t* d_tw2f = nullptr;
    uint32_t* d_r4 = nullptr;       // digit x twiddle tables of the first two forward stages (br7_kernel)
    tw_t* d_tw2i = nullptr;
    tw_t* d_twfull = nullptr;       // [2][1024] psi_rev and inverse (br8_kernel)
#ifndef B200FHE_80BIT
    BlockTw* d_blocktw = nullptr;   // team-NTT twiddles (br4_kernel)
    Block8Tw* d_block8tw = nullptr; // 128-thread team NTT (br6_kernel)
#ifdef B200FHE_WITH_BR9
    Block4Tw* d_block4tw = nullptr; // 256-thread team NTT (br9_kernel)
#endif
#endif
    uint32_t* d_bk_ntt = nullptr;   // [636][6][6][1024]
    torus0_t* d_ksk = nullptr;      // [1024][t][3][SLOT_STRIDE]
    bool keys = false;
    torus0_t* d_arena = nullptr;
    bool arena_owned = false;
    size_t n_slots = 0;
    // job staging (pinned host + device), grown on demand
    size_t cap = 0;
    BrJob *h_br = nullptr, *d_br = nullptr;
    KsJob *h_ks = nullptr, *d_ks = nullptr;
    UnaryJob *h_un = nullptr, *d_un = nullptr;
    uint32_t* d_ubuf = nullptr;     // [2*cap][U_STRIDE]
    uint32_t* d_unstage = nullptr;  // [cap][320]
    uint32_t* d_kspart = nullptr;   // [KS_SPLIT_MAX_GATES][KS_SPLIT][2][320] partial sums of the split key switch
    cudaEvent_t ev_staged = nullptr, ev_t[3] = {nullptr, nullptr, nullptr};
    uint8_t* d_stage = nullptr;  // landing buffer of large uploads (packed rows), repacked into slots on the device
    size_t stage_bytes = 0;
    bool staged_pending = false, timed = false;
    // per-segment timing of the most recent launch plan (see plan_rotation)
    cudaEvent_t ev_seg[BR_MAX_SEGMENTS + 1] = {};
    int seg_n = 0, seg_variant[BR_MAX_SEGMENTS] = {}, seg_G[BR_MAX_SEGMENTS] = {}, seg_count[BR_MAX_SEGMENTS] = {};
    uint64_t launches = 0;
    // multi-GPU exchange (one process per GPU): NCCL communicator, created by b200fhe_comm_init
    void* comm = nullptr;
    int rank = 0, world = 1;
};

static int set_dev(b200fhe_ctx* c)
{
    CK(cudaSetDevice(c->device));
    return 0;
}

// NCCL, bound at run time (see the exchange section below)
struct NcclId { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load()
{
    if (g_nccl.lib) return 0;
    const char* names[] = {getenv("B200FHE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
        if (n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return fail(std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
    NcclApi a;
    a.lib = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy) return fail("libnccl lacks the expected symbols");
    g_nccl = a;
    return 0;
}
static int nccl_fail(const char* what, int rc)
{
    return fail(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error") + " (" +
                std::to_string(rc) + ")");
}

// ---- launch plan ------------------------------------------------------------------------------------
// A frontier of n rotation jobs is cut into launches of the shapes below so that the modelled time is
// minimal: whole waves of the throughput shape, the tail on the latency shapes.  The table holds, per
// shape, the jobs one wave of 148 SMs takes and the duration of a wave.  The defaults were measured on a
// B200 at 1965 MHz (profiles/r02_latency_table.json); b200fhe_calibrate() re-measures them on the device
// at hand (one short wave per shape), so a differently clocked or power-capped part plans with its own
// numbers.  The plan is an unbounded-knapsack DP over "one more wave of shape s" (cheap: 5 shapes).
#ifdef B200FHE_80BIT
constexpr int N_SHAPES = 1;
#else
constexpr int N_SHAPES = 5;
#endif
struct BrShape { int variant, G, wave_jobs; double wave_ms; };
struct BrSegment { int variant, G, count; };
struct PlanTable {
#ifdef B200FHE_80BIT
    BrShape shape[N_SHAPES] = {{1, 4, 592, 12.0}};  // generic shape only (brg_kernel<4>); calibrated at key load
#else
    BrShape shape[N_SHAPES] = {
        {7, 8, 1184, 18.98}, // 16-warp throughput shape: 62.4 k rotations/s
        {3, 6, 888, 14.70},  // 12-warp throughput shape: 60.4 k/s, fills the gap between one and two 16-warp waves
        {3, 4, 592, 10.25},
        {4, 1, 148, 3.02},   // one job per SM
        {6, 1, 74, 2.06},    // one job per 2-SM cluster: lowest latency
    };
#endif
    double launch_ms = 0.01;  // per extra launch: breaks ties in favour of fewer segments
    bool calibrated = false;
    // DP cache: best[m] = (ms, shape of the last wave) for m jobs, m < DP_MAX
    static constexpr int DP_MAX = 4 * 1184 + 1;  // >= a few waves of the widest shape
    double best_ms[DP_MAX];
    int8_t best_shape[DP_MAX];
    bool dp_valid = false;
    void build()
    {
        best_ms[0] = 0.0;
        best_shape[0] = -1;
        for (int m = 1; m < DP_MAX; m++) {
            double bm = 1e30;
            int bs = 0;
            for (int k = 0; k < N_SHAPES; k++) {
                const BrShape& sh = shape[k];
                const int rest = m > sh.wave_jobs ? m - sh.wave_jobs : 0;
                // a new launch is needed whenever the shape changes; approximating by "per wave of a shape other
                // than the previous one" keeps the DP one-dimensional
                const double t = sh.wave_ms + best_ms[rest] + ((rest > 0 && best_shape[rest] != k) ? launch_ms : 0.0);
                if (t < bm - 1e-12) bm = t, bs = k;
            }
            best_ms[m] = bm;
            best_shape[m] = (int8_t)bs;
        }
        dp_valid = true;
    }
    // counts[k] = jobs assigned to shape k; returns modelled ms
    double solve(int n, int (&counts)[N_SHAPES])
    {
        if (!dp_valid) build();
        for (int& c : counts) c = 0;
        double ms = 0.0;
        // far above the DP range the best-throughput shape takes whole waves
        int top = 0;
        for (int k = 1; k < N_SHAPES; k++)
            if (shape[k].wave_jobs / shape[k].wave_ms > shape[top].wave_jobs / shape[top].wave_ms) top = k;
        while (n >= DP_MAX) {
            counts[top] += shape[top].wave_jobs;
            n -= shape[top].wave_jobs;
            ms += shape[top].wave_ms;
        }
        ms += best_ms[n];
        while (n > 0) {
            const int k = best_shape[n];
            const int take = n < shape[k].wave_jobs ? n : shape[k].wave_jobs;
            counts[k] += take;
            n -= take;
        }
        return ms;
    }
};
static PlanTable g_plan;  // process-wide: the planning entry points of the ABI take no context

static int plan_rotation(const b200fhe_ctx* c, int njobs, BrSegment (&seg)[BR_MAX_SEGMENTS], double* model_ms = nullptr)
{
    if (!c->autotune) {
        seg[0] = BrSegment{c->variant, c->G, njobs};
        if (model_ms) *model_ms = 0.0;
        return 1;
    }
    int counts[N_SHAPES];
    const double ms = g_plan.solve(njobs, counts);
    if (model_ms) *model_ms = ms;
    int n = 0;
    for (int k = 0; k < N_SHAPES; k++)  // table order = widest shape first: full waves first, the tail last
        if (counts[k] > 0) seg[n++] = BrSegment{g_plan.shape[k].variant, g_plan.shape[k].G, counts[k]};
    return n;
}

template <int G>
static int brg_launch(b200fhe_ctx* c, int njobs, const torus0_t* arena, uint32_t* ubuf, const BrJob* d_jobs)
{
    const int grid = (njobs + G - 1) / G;
    brg_kernel<G><<<grid, 64 * G, BrgSmem<G>::BYTES, c->stream>>>(d_jobs, njobs, arena, c->d_bk_ntt, c->d_tw2f,
                                                                  c->d_tw2i, ubuf, N0);
    CK(cudaGetLastError());
    c->launches++;
    return 0;
}

#ifndef B200FHE_80BIT
template <int G>
static int br3_launch(b200fhe_ctx* c, int njobs, const torus0_t* arena, uint32_t* ubuf, const BrJob* d_jobs)
{
    const int grid = (njobs + G - 1) / G;
    br3_kernel<G><<<grid, 64 * G, BrSmem<G>::BYTES, c->stream>>>(d_jobs, njobs, arena, c->d_bk_ntt, c->d_tw2f,
                                                                  c->d_tw2i, ubuf, N0);
    CK(cudaGetLastError());
    c->launches++;
    return 0;
}

template <int G, int J>
static int br7_launch(b200fhe_ctx* c, int njobs, const torus0_t* arena, uint32_t* ubuf, const BrJob* d_jobs)
{
    const int grid = (njobs + G - 1) / G;
    br7_kernel<G, J><<<grid, 64 * G, Br7Smem<G>::BYTES, c->stream>>>(d_jobs, njobs, arena, c->d_bk_ntt, c->d_tw2f,
                                                                     c->d_tw2i, c->d_r4, ubuf, N0, c->br7_skew);
    CK(cudaGetLastError());
    c->launches++;
    return 0;
}

static int br4_launch(b200fhe_ctx* c, int njobs, const torus0_t* arena, uint32_t* ubuf, const BrJob* d_jobs)
{
    br4_kernel<<<njobs, BR4_THREADS, Br4Smem::BYTES, c->stream>>>(d_jobs, njobs, arena, c->d_bk_ntt, c->d_blocktw, ubuf, N0);
    CK(cudaGetLastError());
    c->launches++;
    return 0;
}

#ifdef B200FHE_WITH_BR8
static int br8_launch(b200fhe_ctx* c, int njobs, const torus0_t* arena, uint32_t* ubuf, const BrJob* d_jobs)
{
    br8_kernel<<<4 * njobs, BR8_THREADS, Br8Smem::BYTES, c->stream>>>(d_jobs, njobs, arena, c->d_bk_ntt, c->d_twfull,
                                                                       c->d_twfull + 1024, ubuf, N0);
    CK(cudaGetLastError());
    c->launches++;
    return 0;
}
#endif

#ifdef B200FHE_WITH_BR9
static int br9_launch(b200fhe_ctx* c, int njobs, const torus0_t* arena, uint32_t* ubuf, const BrJob* d_jobs)
{
    br9_kernel<<<2 * njobs, BR9_THREADS, Br9Smem::BYTES, c->stream>>>(d_jobs, njobs, arena, c->d_bk_ntt, c->d_block4tw, ubuf, N0);
    CK(cudaGetLastError());
    c->launches++;
    return 0;
}
#endif

static int br6_launch(b200fhe_ctx* c, int njobs, const torus0_t* arena, uint32_t* ubuf, const BrJob* d_jobs)
{
    br6_kernel<<<2 * njobs, BR6_THREADS, Br6Smem::BYTES, c->stream>>>(d_jobs, njobs, arena, c->d_bk_ntt, c->d_block8tw, ubuf, N0);
    CK(cudaGetLastError());
    c->launches++;
    return 0;
}

#endif

// dynamic shared memory opt-in of every blind-rotation shape, once per device (never inside a graph capture)
static int set_kernel_attrs()
{
    CK(cudaFuncSetAttribute(brg_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BrgSmem<4>::BYTES));
    CK(cudaFuncSetAttribute(brg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BrgSmem<2>::BYTES));
#ifndef B200FHE_80BIT
    CK(cudaFuncSetAttribute(br3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BrSmem<2>::BYTES));
    CK(cudaFuncSetAttribute(br3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BrSmem<4>::BYTES));
    CK(cudaFuncSetAttribute(br3_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BrSmem<6>::BYTES));
    CK(cudaFuncSetAttribute(br7_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Br7Smem<8>::BYTES));
    CK(cudaFuncSetAttribute(br7_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Br7Smem<8>::BYTES));
    CK(cudaFuncSetAttribute(br7_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Br7Smem<8>::BYTES));
    CK(cudaFuncSetAttribute(br4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Br4Smem::BYTES));
    CK(cudaFuncSetAttribute(br6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Br6Smem::BYTES));
#ifdef B200FHE_WITH_BR9
    CK(cudaFuncSetAttribute(br9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Br9Smem::BYTES));
#endif
#ifdef B200FHE_WITH_BR8
    CK(cudaFuncSetAttribute(br8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Br8Smem::BYTES));
#endif
#endif
    return 0;
}

static int br_dispatch_one(b200fhe_ctx* c, int variant, int G, int njobs, const torus0_t* arena, uint32_t* ubuf,
                           const BrJob* d_jobs)
{
    if (variant == 1) {
        switch (G) {
        case 2: return brg_launch<2>(c, njobs, arena, ubuf, d_jobs);
        case 4: return brg_launch<4>(c, njobs, arena, ubuf, d_jobs);
        default: return fail("variant 1 (generic shape) supports 2 or 4 jobs per CTA");
        }
    }
#ifndef B200FHE_80BIT
    if (variant == 7) {
        switch (c->br7_group) {
        case 2: return br7_launch<8, 2>(c, njobs, arena, ubuf, d_jobs);
        case 4: return br7_launch<8, 4>(c, njobs, arena, ubuf, d_jobs);
        default: return br7_launch<8, 8>(c, njobs, arena, ubuf, d_jobs);
        }
    }
#ifdef B200FHE_WITH_BR8
    if (variant == 8) return br8_launch(c, njobs, arena, ubuf, d_jobs);
#endif
    if (variant == 6) return br6_launch(c, njobs, arena, ubuf, d_jobs);
#ifdef B200FHE_WITH_BR9
    if (variant == 9) return br9_launch(c, njobs, arena, ubuf, d_jobs);
#endif
    if (variant == 4) return br4_launch(c, njobs, arena, ubuf, d_jobs);
    if (variant == 3) {
        switch (G) {
        case 2: return br3_launch<2>(c, njobs, arena, ubuf, d_jobs);
        case 4: return br3_launch<4>(c, njobs, arena, ubuf, d_jobs);
        case 6: return br3_launch<6>(c, njobs, arena, ubuf, d_jobs);
        default: return fail("variant 3 supports 2, 4 or 6 jobs per CTA");
        }
    }
    return fail("kernel variant must be 1, 3, 4, 6 or 7 (8 and 9 are experiments: build with -DB200FHE_WITH_BR8 / _BR9)");
#else
    return fail("the 80-bit flavour carries the generic shape only (variant 1)");
#endif
}

// Measures one wave of every shape on this device (dummy jobs on a zeroed slot: the kernels' time does not depend on
// the data) and replaces the table's defaults, so that a differently clocked or power-capped part plans - and, through
// b200fhe_plan_ms, schedules netlists - with its own numbers.  Called by b200fhe_load_keys; ~0.1 s.
static int calibrate(b200fhe_ctx* c)
{
    const char* off = getenv("B200FHE_NO_CALIBRATE");
    if (off && off[0] == '1') return 0;
    int maxw = 0;
    for (const BrShape& sh : g_plan.shape) maxw = std::max(maxw, sh.wave_jobs);
    torus0_t* d_slot = nullptr;
    uint32_t* d_u = nullptr;
    BrJob* d_jobs = nullptr;
    CK(cudaMalloc(&d_slot, SLOT_BYTES));
    CK(cudaMalloc(&d_u, (size_t)maxw * U_STRIDE * 4));
    CK(cudaMalloc(&d_jobs, (size_t)maxw * sizeof(BrJob)));
    CK(cudaMemsetAsync(d_slot, 0, SLOT_BYTES, c->stream));
    std::vector<BrJob> jobs(maxw, BrJob{{0u, 0u, 0u}, {1, 0, 0}, 0, 0u});
    CK(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(BrJob), cudaMemcpyHostToDevice, c->stream));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const uint64_t l0 = c->launches;
    for (int k = 0; k < N_SHAPES; k++) {
        BrShape& sh = g_plan.shape[k];
        float best = 1e30f;
        for (int rep = 0; rep < 2; rep++) {  // first run warms the instruction cache and L2
            CK(cudaEventRecord(e0, c->stream));
            if (br_dispatch_one(c, sh.variant, sh.G, sh.wave_jobs, d_slot, d_u, d_jobs)) return 1;
            CK(cudaEventRecord(e1, c->stream));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0) best = std::min(best, ms);
        }
        sh.wave_ms = best;
    }
    c->launches = l0;  // not the caller's work
    g_plan.calibrated = true;
    g_plan.dp_valid = false;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_slot);
    cudaFree(d_u);
    cudaFree(d_jobs);
    return 0;
}

// rotation job k writes ubuf[k]: segments are contiguous ranges of the job list
static int br_dispatch(b200fhe_ctx* c, int njobs, const torus0_t* arena, uint32_t* ubuf, const BrJob* d_jobs,
                       bool timing = true)
{
    BrSegment seg[BR_MAX_SEGMENTS];
    const int nseg = plan_rotation(c, njobs, seg);
    int first = 0;
    if (timing) CK(cudaEventRecord(c->ev_seg[0], c->stream));
    for (int k = 0; k < nseg; k++) {
        if (br_dispatch_one(c, seg[k].variant, seg[k].G, seg[k].count, arena, ubuf + (size_t)first * U_STRIDE,
                            d_jobs + first))
            return 1;
        if (timing) CK(cudaEventRecord(c->ev_seg[k + 1], c->stream));
        c->seg_variant[k] = seg[k].variant;
        c->seg_G[k] = (seg[k].variant == 4 || seg[k].variant == 6 || seg[k].variant == 8 || seg[k].variant == 9) ? 1 : seg[k].G;
        c->seg_count[k] = seg[k].count;
        first += seg[k].count;
    }
    c->seg_n = nseg;
    return 0;
}

// key switch of `nks` gates: one CTA per gate, or KS_SPLIT CTAs per gate + combine for narrow frontiers
static int ks_dispatch(b200fhe_ctx* c, size_t nks, const KsJob* d_jobs, const uint32_t* ubuf, torus0_t* arena)
{
    const uint32_t* ksk = reinterpret_cast<const uint32_t*>(c->d_ksk);
    if (nks <= (size_t)KS_SPLIT_MAX_GATES) {
        ks_split_kernel<<<(unsigned)(nks * KS_SPLIT), dim3(KS_THREADS, KS_GROUPS), 0, c->stream>>>(d_jobs, ubuf, ksk, c->d_kspart);
        ks_combine_kernel<<<(unsigned)nks, KS_THREADS, 0, c->stream>>>(d_jobs, ubuf, c->d_kspart, arena);
        c->launches += 2;
    } else if (nks >= (size_t)c->ks8_min) {
        ks8_kernel<<<(unsigned)((nks + KS8_GATES - 1) / KS8_GATES), 32 * KS8_GATES, 0, c->stream>>>(d_jobs, (int)nks, ubuf, ksk, arena);
        c->launches++;
    } else {
        ks_kernel<<<(unsigned)nks, dim3(KS_THREADS, KS_GROUPS), 0, c->stream>>>(d_jobs, ubuf, ksk, arena);
        c->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

static int ensure_cap(b200fhe_ctx* c, size_t n)
{
    if (n <= c->cap) return 0;
    CK(cudaStreamSynchronize(c->stream));
    size_t cap = c->cap ? c->cap : 1024;
    while (cap < n) cap *= 2;
    if (c->h_br) cudaFreeHost(c->h_br);
    if (c->h_ks) cudaFreeHost(c->h_ks);
    if (c->h_un) cudaFreeHost(c->h_un);
    if (c->d_br) cudaFree(c->d_br);
    if (c->d_ks) cudaFree(c->d_ks);
    if (c->d_un) cudaFree(c->d_un);
    if (c->d_ubuf) cudaFree(c->d_ubuf);
    if (c->d_unstage) cudaFree(c->d_unstage);
    c->cap = 0;
    CK(cudaHostAlloc(&c->h_br, 2 * cap * sizeof(BrJob), cudaHostAllocDefault));
    CK(cudaHostAlloc(&c->h_ks, cap * sizeof(KsJob), cudaHostAllocDefault));
    CK(cudaHostAlloc(&c->h_un, cap * sizeof(UnaryJob), cudaHostAllocDefault));
    CK(cudaMalloc(&c->d_br, 2 * cap * sizeof(BrJob)));
    CK(cudaMalloc(&c->d_ks, cap * sizeof(KsJob)));
    CK(cudaMalloc(&c->d_un, cap * sizeof(UnaryJob)));
    CK(cudaMalloc(&c->d_ubuf, 2 * cap * (size_t)U_STRIDE * 4));
    CK(cudaMalloc(&c->d_unstage, cap * (size_t)KS_THREADS * 4));
    c->cap = cap;
    return 0;
}

extern "C" {

const char* b200fhe_last_error(void) { return g_err.c_str(); }

int b200fhe_create(b200fhe_ctx** out, int device)
{
    if (!out) return fail("null out pointer");
    *out = nullptr;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("no such CUDA device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(std::string("b200fhe is built for sm_100a only; device is sm_") + std::to_string(prop.major) +
                    std::to_string(prop.minor));
    if (set_kernel_attrs()) return 1;
    b200fhe_ctx* c = new b200fhe_ctx();
    c->device = device;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_staged, cudaEventDisableTiming));
    for (auto& e : c->ev_t) CK(cudaEventCreate(&e));
    for (auto& e : c->ev_seg) CK(cudaEventCreate(&e));
    if (const char* e = getenv("B200FHE_BR7_GROUP")) c->br7_group = atoi(e);   // experiment knobs (profiles/r02_br7_groups.md)
    if (const char* e = getenv("B200FHE_BR7_SKEW")) c->br7_skew = atoi(e);
    if (const char* e = getenv("B200FHE_KS8_MIN")) c->ks8_min = atoi(e);
    c->tab = new NttTables();
    ntt_tables_init(*c->tab);
    // every transfer goes through the context's non-blocking stream: the legacy default stream is not
    // ordered with it
    CK(cudaMemcpyToSymbolAsync(c_twf_u, h_twf_u, sizeof(h_twf_u), 0, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyToSymbolAsync(c_twi_u, h_twi_u, sizeof(h_twi_u), 0, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMalloc(&c->d_tw2f, sizeof(c->tab->tw2f)));
    CK(cudaMalloc(&c->d_tw2i, sizeof(c->tab->tw2i)));
    CK(cudaMemcpyAsync(c->d_tw2f, c->tab->tw2f, sizeof(c->tab->tw2f), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_tw2i, c->tab->tw2i, sizeof(c->tab->tw2i), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMalloc(&c->d_r4, sizeof(c->tab->r4)));
    CK(cudaMemcpyAsync(c->d_r4, c->tab->r4, sizeof(c->tab->r4), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMalloc(&c->d_twfull, 2 * sizeof(c->tab->fwd)));
    CK(cudaMemcpyAsync(c->d_twfull, c->tab->fwd, sizeof(c->tab->fwd), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_twfull + 1024, c->tab->inv, sizeof(c->tab->inv), cudaMemcpyHostToDevice, c->stream));
#ifndef B200FHE_80BIT
    BlockTw* btw = new BlockTw();
    block_tw_init(*c->tab, *btw);
    CK(cudaMalloc(&c->d_blocktw, sizeof(BlockTw)));
    CK(cudaMemcpyAsync(c->d_blocktw, btw, sizeof(BlockTw), cudaMemcpyHostToDevice, c->stream));
    Block8Tw* b8tw = new Block8Tw();
    block8_tw_init(*c->tab, *b8tw);
    CK(cudaMalloc(&c->d_block8tw, sizeof(Block8Tw)));
    CK(cudaMemcpyAsync(c->d_block8tw, b8tw, sizeof(Block8Tw), cudaMemcpyHostToDevice, c->stream));
#endif
    CK(cudaMalloc(&c->d_kspart, (size_t)KS_SPLIT_MAX_GATES * KS_SPLIT * 2 * KS_THREADS * 4));
    CK(cudaStreamSynchronize(c->stream));
#ifndef B200FHE_80BIT
    delete btw;
    delete b8tw;
#ifdef B200FHE_WITH_BR9
    Block4Tw* b4tw = new Block4Tw();
    block4_tw_init(*c->tab, *b4tw);
    CK(cudaMalloc(&c->d_block4tw, sizeof(Block4Tw)));
    CK(cudaMemcpyAsync(c->d_block4tw, b4tw, sizeof(Block4Tw), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    delete b4tw;
#endif
#endif
    *out = c;
    return 0;
}

void b200fhe_destroy(b200fhe_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_tw2f);
    cudaFree(c->d_tw2i);
    cudaFree(c->d_r4);
    cudaFree(c->d_twfull);
#ifndef B200FHE_80BIT
    cudaFree(c->d_blocktw);
    cudaFree(c->d_block8tw);
#ifdef B200FHE_WITH_BR9
    cudaFree(c->d_block4tw);
#endif
#endif
    cudaFree(c->d_bk_ntt);
    cudaFree(c->d_ksk);
    if (c->arena_owned) cudaFree(c->d_arena);
    if (c->h_br) cudaFreeHost(c->h_br);
    if (c->h_ks) cudaFreeHost(c->h_ks);
    if (c->h_un) cudaFreeHost(c->h_un);
    cudaFree(c->d_br);
    cudaFree(c->d_ks);
    cudaFree(c->d_un);
    cudaFree(c->d_ubuf);
    cudaFree(c->d_unstage);
    cudaFree(c->d_kspart);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    cudaEventDestroy(c->ev_staged);
    cudaFree(c->d_stage);
    for (auto& e : c->ev_t) cudaEventDestroy(e);
    for (auto& e : c->ev_seg) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c->tab;
    delete c;
}

int b200fhe_set_jobs_per_cta(b200fhe_ctx* c, int g)
{
    if (!c) return fail("null context");
    if (g == 0) {  // back to the batch-size heuristic
        c->autotune = true;
#ifdef B200FHE_80BIT
        c->G = 4;
        c->variant = 1;
#else
        c->G = 8;
        c->variant = 7;
#endif
        return 0;
    }
    if (g != 2 && g != 4 && g != 6 && g != 8) return fail("jobs per CTA must be 2, 4 (variants 1, 3), 6 (variant 3) or 8 (variant 7)");
    c->G = g;
    c->autotune = false;
    return 0;
}

int b200fhe_set_kernel_variant(b200fhe_ctx* c, int variant)
{
    if (!c) return fail("null context");
    if (variant == 0) return b200fhe_set_jobs_per_cta(c, 0);
    if (variant != 1 && variant != 3 && variant != 4 && variant != 6 && variant != 7 && variant != 8 && variant != 9)
        return fail("kernel variant must be 0 (auto), 1, 3, 4, 6, 7, 8 or 9");
    c->variant = variant;
    c->autotune = false;
    return 0;
}

int b200fhe_load_keys(b200fhe_ctx* c, const uint32_t* bk_raw, const torus0_t* ksk)
{
    if (!c || !bk_raw || !ksk) return fail("null argument");
    if (set_dev(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    const size_t bk_raw_bytes = B200FHE_BK_WORDS * 4;
    const size_t bk_ntt_bytes = (size_t)N0 * BK_COLS * ROWS * N1 * 4;
    const size_t ksk_rows = (size_t)N1 * KS_T * 3;
    if (!c->d_bk_ntt) CK(cudaMalloc(&c->d_bk_ntt, bk_ntt_bytes));
    if (!c->d_ksk) CK(cudaMalloc(&c->d_ksk, ksk_rows * KSK_ROW * sizeof(torus0_t)));
    uint32_t* d_raw = nullptr;
    CK(cudaMalloc(&d_raw, bk_raw_bytes));
    CK(cudaMemcpyAsync(d_raw, bk_raw, bk_raw_bytes, cudaMemcpyHostToDevice, c->stream));
    const int ntasks = N0 * ROWS * 2 * LIMBS;
    bk_prep_kernel<<<(ntasks + BKPREP_WARPS - 1) / BKPREP_WARPS, 32 * BKPREP_WARPS, 0, c->stream>>>(
        d_raw, c->d_bk_ntt, c->d_tw2f, c->tab->bk_scale, ntasks);
    CK(cudaGetLastError());
    c->launches++;
    // key-switching key: pad every 637-element row to 640 (1280 B, 16-byte aligned rows)
    CK(cudaMemsetAsync(c->d_ksk, 0, ksk_rows * KSK_ROW * sizeof(torus0_t), c->stream));
    CK(cudaMemcpy2DAsync(c->d_ksk, KSK_ROW * sizeof(torus0_t), ksk, TLWE0_LEN * sizeof(torus0_t), TLWE0_LEN * sizeof(torus0_t), ksk_rows, cudaMemcpyHostToDevice,
                         c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(d_raw));
    c->keys = true;
    if (const char* e = getenv("B200FHE_L2_PERSIST"); e && e[0] == '1') {
        // experiment (profiles/r02_l2_persist.md): pin the NTT-domain bootstrapping key in L2 across waves
        int max_persist = 0, max_window = 0;
        CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device));
        CK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device));
        const size_t set_aside = std::min(bk_ntt_bytes, (size_t)max_persist);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside));
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = c->d_bk_ntt;
        attr.accessPolicyWindow.num_bytes = std::min(bk_ntt_bytes, (size_t)max_window);
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)set_aside / (double)attr.accessPolicyWindow.num_bytes);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        fprintf(stderr, "b200fhe: L2 persisting window %zu MB (set-aside %zu MB, hit ratio %.2f)\n",
                attr.accessPolicyWindow.num_bytes >> 20, set_aside >> 20, attr.accessPolicyWindow.hitRatio);
    }
    return calibrate(c);
}

int b200fhe_arena_alloc(b200fhe_ctx* c, size_t n_slots)
{
    if (!c || n_slots == 0) return fail("bad arena size");
    if (set_dev(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    if (c->arena_owned && c->d_arena) CK(cudaFree(c->d_arena));
    c->d_arena = nullptr;
    CK(cudaMalloc(&c->d_arena, n_slots * SLOT_BYTES));
    // on the context's own (non-blocking) stream: a legacy-stream memset would not be ordered with the
    // uploads that follow and could wipe them
    CK(cudaMemsetAsync(c->d_arena, 0, n_slots * SLOT_BYTES, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->arena_owned = true;
    c->n_slots = n_slots;
    return 0;
}

int b200fhe_arena_attach(b200fhe_ctx* c, void* dev_ptr, size_t n_slots)
{
    if (!c || !dev_ptr || n_slots == 0) return fail("bad arena");
    if (set_dev(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    if (c->arena_owned && c->d_arena) CK(cudaFree(c->d_arena));
    c->d_arena = reinterpret_cast<torus0_t*>(dev_ptr);
    c->arena_owned = false;
    c->n_slots = n_slots;
    return 0;
}

size_t b200fhe_arena_slots(const b200fhe_ctx* c) { return c ? c->n_slots : 0; }
void* b200fhe_arena_dev_ptr(const b200fhe_ctx* c) { return c ? c->d_arena : nullptr; }
void* b200fhe_stream(const b200fhe_ctx* c) { return c ? (void*)c->stream : nullptr; }
uint64_t b200fhe_launch_count(const b200fhe_ctx* c) { return c ? c->launches : 0; }

static int check_slots(b200fhe_ctx* c, const uint32_t* ids, size_t n)
{
    if (!c->d_arena) return fail("no arena allocated");
    for (size_t i = 0; i < n; i++)
        if (ids[i] >= c->n_slots) return fail("slot id out of range");
    return 0;
}

// copies runs of consecutive slot ids with one strided copy each
// packed rows [n][TLWE0_LEN] -> slots of SLOT_STRIDE elements (one thread per element)
__global__ void repack_rows_kernel(const torus0_t* __restrict__ packed, torus0_t* __restrict__ slots, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n * TLWE0_LEN) slots[(k / TLWE0_LEN) * SLOT_STRIDE + k % TLWE0_LEN] = packed[k];
}

int b200fhe_upload(b200fhe_ctx* c, const uint32_t* ids, const torus0_t* host, size_t n)
{
    if (!c || (n && (!ids || !host))) return fail("null argument");
    if (set_dev(c) || check_slots(c, ids, n)) return 1;
    constexpr size_t ROW = TLWE0_LEN * sizeof(torus0_t);
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && ids[j] == ids[j - 1] + 1) j++;
        const size_t rows = j - i;
        torus0_t* dst = c->d_arena + (size_t)ids[i] * SLOT_STRIDE;
        if (rows >= 256) {
            // a strided 2-D copy of 1274-byte rows runs at ~9 GB/s over PCIe; one contiguous copy (~50 GB/s) into a
            // landing buffer plus a device-side repack is 4x faster for whole batches
            if (c->stage_bytes < rows * ROW) {
                CK(cudaStreamSynchronize(c->stream));
                cudaFree(c->d_stage);
                c->d_stage = nullptr;
                c->stage_bytes = 0;
                CK(cudaMalloc(&c->d_stage, rows * ROW));
                c->stage_bytes = rows * ROW;
            }
            CK(cudaMemcpyAsync(c->d_stage, host + i * TLWE0_LEN, rows * ROW, cudaMemcpyHostToDevice, c->stream));
            const size_t elems = rows * TLWE0_LEN;
            repack_rows_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, c->stream>>>(
                reinterpret_cast<const torus0_t*>(c->d_stage), dst, rows);
            CK(cudaGetLastError());
        } else {
            CK(cudaMemcpy2DAsync(dst, SLOT_BYTES, host + i * TLWE0_LEN, ROW, ROW, rows, cudaMemcpyHostToDevice, c->stream));
        }
        i = j;
    }
    return 0;
}

int b200fhe_download(b200fhe_ctx* c, const uint32_t* ids, torus0_t* host, size_t n)
{
    if (!c || (n && (!ids || !host))) return fail("null argument");
    if (set_dev(c) || check_slots(c, ids, n)) return 1;
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && ids[j] == ids[j - 1] + 1) j++;
        CK(cudaMemcpy2DAsync(host + i * TLWE0_LEN, TLWE0_LEN * sizeof(torus0_t), c->d_arena + (size_t)ids[i] * SLOT_STRIDE,
                             SLOT_BYTES, TLWE0_LEN * sizeof(torus0_t), j - i, cudaMemcpyDeviceToHost, c->stream));
        i = j;
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int b200fhe_gate_batch(b200fhe_ctx* c, const uint8_t* opcode, const uint32_t* in0, const uint32_t* in1,
                       const uint32_t* in2, const uint32_t* out, size_t n)
{
    if (!c) return fail("null context");
    if (n == 0) return 0;
    if (!opcode || !out) return fail("null argument");
    if (!c->keys) return fail("keys not loaded");
    if (!c->d_arena) return fail("no arena allocated");
    if (set_dev(c) || ensure_cap(c, n)) return 1;
    if (c->staged_pending) {
        CK(cudaEventSynchronize(c->ev_staged));
        c->staged_pending = false;
    }
    BatchCounts cnt;
    if (const char* err = build_gate_jobs(opcode, in0, in1, in2, out, n, c->n_slots, c->h_br, c->h_ks, c->h_un, cnt))
        return fail(err);
    const size_t nbr = cnt.nbr, nks = cnt.nks, nun = cnt.nun;
    if (nbr) CK(cudaMemcpyAsync(c->d_br, c->h_br, nbr * sizeof(BrJob), cudaMemcpyHostToDevice, c->stream));
    if (nks) CK(cudaMemcpyAsync(c->d_ks, c->h_ks, nks * sizeof(KsJob), cudaMemcpyHostToDevice, c->stream));
    if (nun) CK(cudaMemcpyAsync(c->d_un, c->h_un, nun * sizeof(UnaryJob), cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->ev_staged, c->stream));
    c->staged_pending = true;

    // every kernel of a batch reads only slots written by earlier batches, so order is free;
    // the bootstrap-free ops go first, staged through a scratch buffer.
    if (nun) {
        uint32_t* aw = reinterpret_cast<uint32_t*>(c->d_arena);
        unary_gather_kernel<<<(unsigned)nun, KS_THREADS, 0, c->stream>>>(c->d_un, aw, c->d_unstage);
        unary_scatter_kernel<<<(unsigned)nun, KS_THREADS, 0, c->stream>>>(c->d_un, c->d_unstage, aw);
        CK(cudaGetLastError());
        c->launches += 2;
    }
    c->timed = false;
    if (nbr) {
        CK(cudaEventRecord(c->ev_t[0], c->stream));
        if (br_dispatch(c, (int)nbr, c->d_arena, c->d_ubuf, c->d_br)) return 1;
        CK(cudaEventRecord(c->ev_t[1], c->stream));
        if (ks_dispatch(c, nks, c->d_ks, c->d_ubuf, c->d_arena)) return 1;
        CK(cudaEventRecord(c->ev_t[2], c->stream));
        c->timed = true;
    }
    return 0;
}

int b200fhe_dff_tick(b200fhe_ctx* c, const uint32_t* src, const uint32_t* dst, size_t n)
{
    if (!c) return fail("null context");
    if (n == 0) return 0;
    if (!src || !dst) return fail("null argument");
    if (!c->d_arena) return fail("no arena allocated");
    if (set_dev(c) || ensure_cap(c, n) || check_slots(c, src, n) || check_slots(c, dst, n)) return 1;
    if (c->staged_pending) {
        CK(cudaEventSynchronize(c->ev_staged));
        c->staged_pending = false;
    }
    for (size_t i = 0; i < n; i++) c->h_un[i] = UnaryJob{src[i], dst[i], (uint32_t)OP_COPY};
    CK(cudaMemcpyAsync(c->d_un, c->h_un, n * sizeof(UnaryJob), cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->ev_staged, c->stream));
    c->staged_pending = true;
    uint32_t* aw = reinterpret_cast<uint32_t*>(c->d_arena);
    unary_gather_kernel<<<(unsigned)n, KS_THREADS, 0, c->stream>>>(c->d_un, aw, c->d_unstage);
    unary_scatter_kernel<<<(unsigned)n, KS_THREADS, 0, c->stream>>>(c->d_un, c->d_unstage, aw);
    CK(cudaGetLastError());
    c->launches += 2;
    return 0;
}

int b200fhe_sync(b200fhe_ctx* c)
{
    if (!c) return fail("null context");
    if (set_dev(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int b200fhe_query(b200fhe_ctx* c)
{
    if (!c) return -1;
    cudaError_t e = cudaStreamQuery(c->stream);
    if (e == cudaSuccess) return 0;
    if (e == cudaErrorNotReady) return 1;
    g_err = cudaGetErrorString(e);
    return -1;
}

int b200fhe_last_batch_ms(b200fhe_ctx* c, float* br_ms, float* ks_ms)
{
    if (!c) return fail("null context");
    if (!c->timed) return fail("no timed batch");
    CK(cudaEventSynchronize(c->ev_t[2]));
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, c->ev_t[0], c->ev_t[1]));
    CK(cudaEventElapsedTime(&b, c->ev_t[1], c->ev_t[2]));
    if (br_ms) *br_ms = a;
    if (ks_ms) *ks_ms = b;
    return 0;
}

int b200fhe_plan_rotation(int njobs, int* variant, int* jobs_per_cta, int* jobs, int max_segments)
{
    if (njobs <= 0) return 0;
    b200fhe_ctx tmp;  // default context state = heuristic on; no device is touched
    BrSegment seg[BR_MAX_SEGMENTS];
    const int nseg = plan_rotation(&tmp, njobs, seg);
    const int n = nseg < max_segments ? nseg : max_segments;
    for (int k = 0; k < n; k++) {
        if (variant) variant[k] = seg[k].variant;
        if (jobs_per_cta) jobs_per_cta[k] = (seg[k].variant == 4 || seg[k].variant == 6 || seg[k].variant == 8 || seg[k].variant == 9) ? 1 : seg[k].G;
        if (jobs) jobs[k] = seg[k].count;
    }
    return n;
}

int b200fhe_plan_table(int* variant, int* jobs_per_cta, int* wave_jobs, double* wave_ms, int max_shapes)
{
    const int n = N_SHAPES < max_shapes ? N_SHAPES : max_shapes;
    for (int k = 0; k < n; k++) {
        if (variant) variant[k] = g_plan.shape[k].variant;
        if (jobs_per_cta) jobs_per_cta[k] = g_plan.shape[k].G;
        if (wave_jobs) wave_jobs[k] = g_plan.shape[k].wave_jobs;
        if (wave_ms) wave_ms[k] = g_plan.shape[k].wave_ms;
    }
    return g_plan.calibrated ? n : -n;
}

double b200fhe_plan_ms(int njobs)
{
    if (njobs <= 0) return 0.0;
    b200fhe_ctx tmp;
    BrSegment seg[BR_MAX_SEGMENTS];
    double ms = 0.0;
    plan_rotation(&tmp, njobs, seg, &ms);
    return ms;
}

int b200fhe_last_batch_segments(b200fhe_ctx* c, int* variant, int* jobs_per_cta, int* jobs, float* ms, int max_segments)
{
    if (!c) {
        fail("null context");
        return -1;
    }
    if (!c->timed) {
        fail("no timed batch");
        return -1;
    }
    if (cudaEventSynchronize(c->ev_t[2]) != cudaSuccess) {
        fail("cudaEventSynchronize failed");
        return -1;
    }
    int n = c->seg_n < max_segments ? c->seg_n : max_segments;
    for (int k = 0; k < n; k++) {
        float t = 0;
        if (cudaEventElapsedTime(&t, c->ev_seg[k], c->ev_seg[k + 1]) != cudaSuccess) {
            fail("cudaEventElapsedTime failed");
            return -1;
        }
        if (variant) variant[k] = c->seg_variant[k];
        if (jobs_per_cta) jobs_per_cta[k] = c->seg_G[k];
        if (jobs) jobs[k] = c->seg_count[k];
        if (ms) ms[k] = t;
    }
    return n;
}

int b200fhe_gates_host(b200fhe_ctx* c, const uint8_t* opcode, const torus0_t* in0_host, const torus0_t* in1_host,
                       const torus0_t* in2_host, torus0_t* out_host, size_t n)
{
    if (!c) return fail("null context");
    if (n == 0) return 0;
    if (!opcode || !out_host) return fail("null argument");
    if (c->n_slots < 4 * n) return fail("arena too small for b200fhe_gates_host (needs 4*n slots)");
    if (set_dev(c)) return 1;
    // Copies and evaluation run back to back on the context's stream.  Overlapping them in chunks was measured and is
    // slower (profiles/r02_host_path.md): the key switch of a 2368-gate chunk streams the 27 MB key through L2 for a
    // quarter of the gates, which costs more than the 0.6 ms of copies it hides.
    std::vector<uint32_t> ids(4 * n);
    for (size_t i = 0; i < 4 * n; i++) ids[i] = (uint32_t)i;
    const torus0_t* ins[3] = {in0_host, in1_host, in2_host};
    for (int k = 0; k < 3; k++)
        if (ins[k] && b200fhe_upload(c, ids.data() + k * n, ins[k], n)) return 1;
    if (b200fhe_gate_batch(c, opcode, ids.data(), ids.data() + n, ids.data() + 2 * n, ids.data() + 3 * n, n)) return 1;
    return b200fhe_download(c, ids.data() + 3 * n, out_host, n);
}

int b200fhe_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return fail("null argument");
    CK(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return 0;
}
int b200fhe_host_free(void* ptr)
{
    CK(cudaFreeHost(ptr));
    return 0;
}

}  // extern "C"

// ---- multi-GPU exchange --------------------------------------------------------------------------------
// The netlist engine shards a dependency level over the ranks (one process per GPU, replicated arena layout)
// and replicates the level's outputs with ONE in-place all-gather over NVLink.  NCCL is bound at run time
// (dlopen of libnccl.so.2, the library torch.distributed already has in the process, or the system one), so
// single-GPU users of this library need no NCCL at all.  Reference hook this replaces: cufhe::SetGPUNum +
// round-robin streams through host memory (cuFHE include/cufhe_gpu.cuh:164-169, src/iyokan_cufhe.cpp:533).
static int exchange_issue(b200fhe_ctx* c, size_t first_slot, size_t slots_per_rank)
{
    if (c->world == 1 || slots_per_rank == 0) return 0;
    if (!c->comm) return fail("no communicator: call b200fhe_comm_init first");
    if (first_slot + slots_per_rank * (size_t)c->world > c->n_slots) return fail("exchange range exceeds the arena");
    uint8_t* base = reinterpret_cast<uint8_t*>(c->d_arena) + first_slot * (size_t)SLOT_BYTES;
    const size_t bytes = slots_per_rank * (size_t)SLOT_BYTES;
    const int rc = g_nccl.AllGather(base + (size_t)c->rank * bytes, base, bytes, /*ncclUint8*/ 1, c->comm, c->stream);
    if (rc) return nccl_fail("ncclAllGather", rc);
    return 0;
}

// ---- programs: a static schedule recorded once and replayed as one CUDA graph ---------------------------
// The netlist is the same every clock cycle, so the host builds the job lists of every dependency level
// ONCE, uploads them once, and captures all launches of a clock (unary ops, blind rotations, key switches,
// exchanges, the DFF tick) into one CUDA graph; a clock cycle is then a single cudaGraphLaunch.
struct ProgStep {
    int kind = 0;  // 0 = gate batch, 1 = tick (parallel copy), 2 = exchange
    size_t br_off = 0, nbr = 0, ks_off = 0, nks = 0, un_off = 0, nun = 0;
    size_t first_slot = 0, slots_per_rank = 0;
};
struct b200fhe_program {
    b200fhe_ctx* c = nullptr;
    std::vector<BrJob> h_br;
    std::vector<KsJob> h_ks;
    std::vector<UnaryJob> h_un;
    std::vector<ProgStep> steps;
    BrJob* d_br = nullptr;
    KsJob* d_ks = nullptr;
    UnaryJob* d_un = nullptr;
    uint32_t* d_ubuf = nullptr;
    uint32_t* d_unstage = nullptr;
    size_t max_nbr = 0, max_nun = 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool finalized = false;
    uint64_t launches_per_replay = 0, rotations = 0, exchanges = 0, exchanged_slots = 0;
    double model_ms = 0.0;
};

static int program_issue_step(b200fhe_program* p, const ProgStep& st)
{
    b200fhe_ctx* c = p->c;
    uint32_t* aw = reinterpret_cast<uint32_t*>(c->d_arena);
    if (st.kind == 2) return exchange_issue(c, st.first_slot, st.slots_per_rank);
    if (st.nun) {  // gather, then scatter: every source is read before any destination is written
        unary_gather_kernel<<<(unsigned)st.nun, KS_THREADS, 0, c->stream>>>(p->d_un + st.un_off, aw, p->d_unstage);
        unary_scatter_kernel<<<(unsigned)st.nun, KS_THREADS, 0, c->stream>>>(p->d_un + st.un_off, p->d_unstage, aw);
        CK(cudaGetLastError());
        c->launches += 2;
    }
    if (st.nbr) {
        if (br_dispatch(c, (int)st.nbr, c->d_arena, p->d_ubuf, p->d_br + st.br_off, false)) return 1;
        if (ks_dispatch(c, st.nks, p->d_ks + st.ks_off, p->d_ubuf, c->d_arena)) return 1;
    }
    return 0;
}
static int program_issue(b200fhe_program* p)
{
    for (const ProgStep& st : p->steps)
        if (program_issue_step(p, st)) return 1;
    return 0;
}

extern "C" {

int b200fhe_comm_unique_id(uint8_t* id128)
{
    if (!id128) return fail("null argument");
    if (nccl_load()) return 1;
    NcclId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) return nccl_fail("ncclGetUniqueId", rc);
    std::memcpy(id128, id.internal, 128);
    return 0;
}

int b200fhe_comm_init(b200fhe_ctx* c, int rank, int world, const uint8_t* id128)
{
    if (!c) return fail("null context");
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank / world size");
    if (c->comm) {
        g_nccl.CommDestroy(c->comm);
        c->comm = nullptr;
    }
    c->rank = rank;
    c->world = world;
    if (world == 1) return 0;
    if (!id128) return fail("null unique id");
    if (set_dev(c) || nccl_load()) return 1;
    NcclId id;
    std::memcpy(id.internal, id128, 128);
    const int rc = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (rc) return nccl_fail("ncclCommInitRank", rc);
    // every rank must derive the SAME schedule from b200fhe_plan_ms: agree on the launch-plan table (max over ranks)
    if (g_nccl.AllReduce) {
        double h[N_SHAPES], *d = nullptr;
        for (int k = 0; k < N_SHAPES; k++) h[k] = g_plan.shape[k].wave_ms;
        CK(cudaMalloc(&d, sizeof(h)));
        CK(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
        const int rr = g_nccl.AllReduce(d, d, N_SHAPES, /*ncclFloat64*/ 8, /*ncclMax*/ 2, c->comm, c->stream);
        if (rr) return nccl_fail("ncclAllReduce", rr);
        CK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(d));
        for (int k = 0; k < N_SHAPES; k++) g_plan.shape[k].wave_ms = h[k];
        g_plan.dp_valid = false;
    }
    return 0;
}

int b200fhe_comm_rank(const b200fhe_ctx* c) { return c ? c->rank : 0; }
int b200fhe_comm_world(const b200fhe_ctx* c) { return c ? c->world : 1; }

int b200fhe_exchange(b200fhe_ctx* c, size_t first_slot, size_t slots_per_rank)
{
    if (!c) return fail("null context");
    if (!c->d_arena) return fail("no arena allocated");
    if (set_dev(c)) return 1;
    return exchange_issue(c, first_slot, slots_per_rank);
}

int b200fhe_program_create(b200fhe_ctx* c, b200fhe_program** out)
{
    if (!c || !out) return fail("null argument");
    *out = new b200fhe_program();
    (*out)->c = c;
    return 0;
}

void b200fhe_program_destroy(b200fhe_program* p)
{
    if (!p) return;
    if (p->c) {
        cudaSetDevice(p->c->device);
        cudaStreamSynchronize(p->c->stream);
    }
    if (p->exec) cudaGraphExecDestroy(p->exec);
    if (p->graph) cudaGraphDestroy(p->graph);
    cudaFree(p->d_br);
    cudaFree(p->d_ks);
    cudaFree(p->d_un);
    cudaFree(p->d_ubuf);
    cudaFree(p->d_unstage);
    delete p;
}

int b200fhe_program_batch(b200fhe_program* p, const uint8_t* opcode, const uint32_t* in0, const uint32_t* in1,
                          const uint32_t* in2, const uint32_t* out, size_t n)
{
    if (!p) return fail("null program");
    if (p->finalized) return fail("program already finalized");
    if (n == 0) return 0;
    if (!opcode || !out) return fail("null argument");
    if (!p->c->d_arena) return fail("no arena allocated");
    std::vector<BrJob> br(2 * n);
    std::vector<KsJob> ks(n);
    std::vector<UnaryJob> un(n);
    BatchCounts cnt;
    if (const char* err = build_gate_jobs(opcode, in0, in1, in2, out, n, p->c->n_slots, br.data(), ks.data(), un.data(), cnt))
        return fail(err);
    ProgStep st;
    st.kind = 0;
    st.br_off = p->h_br.size();
    st.nbr = cnt.nbr;
    st.ks_off = p->h_ks.size();
    st.nks = cnt.nks;
    st.un_off = p->h_un.size();
    st.nun = cnt.nun;
    p->h_br.insert(p->h_br.end(), br.begin(), br.begin() + cnt.nbr);
    p->h_ks.insert(p->h_ks.end(), ks.begin(), ks.begin() + cnt.nks);
    p->h_un.insert(p->h_un.end(), un.begin(), un.begin() + cnt.nun);
    p->max_nbr = std::max(p->max_nbr, cnt.nbr);
    p->max_nun = std::max(p->max_nun, cnt.nun);
    p->rotations += cnt.nbr;
    if (cnt.nbr) p->model_ms += b200fhe_plan_ms((int)cnt.nbr);
    p->steps.push_back(st);
    return 0;
}

int b200fhe_program_tick(b200fhe_program* p, const uint32_t* src, const uint32_t* dst, size_t n)
{
    if (!p) return fail("null program");
    if (p->finalized) return fail("program already finalized");
    if (n == 0) return 0;
    if (!src || !dst) return fail("null argument");
    if (!p->c->d_arena) return fail("no arena allocated");
    if (check_slots(p->c, src, n) || check_slots(p->c, dst, n)) return 1;
    ProgStep st;
    st.kind = 1;
    st.un_off = p->h_un.size();
    st.nun = n;
    for (size_t i = 0; i < n; i++) p->h_un.push_back(UnaryJob{src[i], dst[i], (uint32_t)OP_COPY});
    p->max_nun = std::max(p->max_nun, n);
    p->steps.push_back(st);
    return 0;
}

int b200fhe_program_exchange(b200fhe_program* p, size_t first_slot, size_t slots_per_rank)
{
    if (!p) return fail("null program");
    if (p->finalized) return fail("program already finalized");
    if (slots_per_rank == 0 || p->c->world == 1) return 0;
    if (first_slot + slots_per_rank * (size_t)p->c->world > p->c->n_slots) return fail("exchange range exceeds the arena");
    ProgStep st;
    st.kind = 2;
    st.first_slot = first_slot;
    st.slots_per_rank = slots_per_rank;
    p->exchanges++;
    p->exchanged_slots += slots_per_rank * (size_t)p->c->world;
    p->steps.push_back(st);
    return 0;
}

int b200fhe_program_finalize(b200fhe_program* p)
{
    if (!p) return fail("null program");
    if (p->finalized) return 0;
    b200fhe_ctx* c = p->c;
    if (!c->keys && p->rotations) return fail("keys not loaded");
    if (set_dev(c)) return 1;
    auto up = [&](auto** d, const auto& h) -> int {
        if (h.empty()) return 0;
        CK(cudaMalloc(d, h.size() * sizeof(h[0])));
        CK(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(h[0]), cudaMemcpyHostToDevice, c->stream));
        return 0;
    };
    if (up(&p->d_br, p->h_br) || up(&p->d_ks, p->h_ks) || up(&p->d_un, p->h_un)) return 1;
    if (p->max_nbr) CK(cudaMalloc(&p->d_ubuf, p->max_nbr * (size_t)U_STRIDE * 4));
    if (p->max_nun) CK(cudaMalloc(&p->d_unstage, p->max_nun * (size_t)KS_THREADS * 4));
    CK(cudaStreamSynchronize(c->stream));
    p->finalized = true;
    // capture one replay; if the capture is refused (driver / NCCL build without graph support) the program
    // still runs, launch by launch, from the resident job lists
    const uint64_t l0 = c->launches;
    const char* no_graph = getenv("B200FHE_NO_GRAPH");
    bool captured = false;
    if (!(no_graph && no_graph[0] == '1') && !p->steps.empty()) {
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int rc = program_issue(p);
            cudaGraph_t g = nullptr;
            const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            if (rc == 0 && e == cudaSuccess && g && cudaGraphInstantiate(&p->exec, g, 0) == cudaSuccess) {
                p->graph = g;
                captured = true;
            } else {
                if (g) cudaGraphDestroy(g);
                p->exec = nullptr;
                cudaGetLastError();  // clear the sticky capture error
            }
        }
        p->launches_per_replay = c->launches - l0;
        c->launches = l0;
    }
    if (!captured) {  // count the launches of one eager replay without running it twice: issue once now
        p->launches_per_replay = 0;
    }
    return 0;
}

int b200fhe_program_launch(b200fhe_program* p)
{
    if (!p) return fail("null program");
    if (!p->finalized && b200fhe_program_finalize(p)) return 1;
    b200fhe_ctx* c = p->c;
    if (set_dev(c)) return 1;
    if (p->exec) {
        CK(cudaGraphLaunch(p->exec, c->stream));
        c->launches += p->launches_per_replay;
        return 0;
    }
    return program_issue(p);
}

int b200fhe_program_profile(b200fhe_program* p, float* step_ms, size_t cap, size_t* nsteps)
{
    if (!p) return fail("null program");
    if (!p->finalized && b200fhe_program_finalize(p)) return 1;
    b200fhe_ctx* c = p->c;
    if (set_dev(c)) return 1;
    const size_t n = p->steps.size();
    if (nsteps) *nsteps = n;
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    int rc = 0;
    CK(cudaEventRecord(ev[0], c->stream));
    for (size_t k = 0; k < n && !rc; k++) {
        rc = program_issue_step(p, p->steps[k]);
        if (!rc && cudaEventRecord(ev[k + 1], c->stream) != cudaSuccess) rc = fail("cudaEventRecord failed");
    }
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail("cudaStreamSynchronize failed");
    for (size_t k = 0; k < n && !rc; k++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) != cudaSuccess) rc = fail("cudaEventElapsedTime failed");
        if (step_ms && k < cap) step_ms[k] = ms;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return rc;
}

int b200fhe_program_info(const b200fhe_program* p, uint64_t* rotations, uint64_t* launches_per_replay, uint64_t* exchanges,
                         uint64_t* exchanged_slots, int* is_graph, double* model_ms)
{
    if (!p) return fail("null program");
    if (rotations) *rotations = p->rotations;
    if (launches_per_replay) *launches_per_replay = p->launches_per_replay;
    if (exchanges) *exchanges = p->exchanges;
    if (exchanged_slots) *exchanged_slots = p->exchanged_slots;
    if (is_graph) *is_graph = p->exec ? 1 : 0;
    if (model_ms) *model_ms = p->model_ms;
    return 0;
}

}  // extern "C"

extern "C" {

// ---- test hooks -------------------------------------------------------------------------

int b200fhe_test_bootstrap_lvl1(b200fhe_ctx* c, const torus0_t* c_host, uint32_t* tlwe1_host, size_t n)
{
    if (!c || !c_host || !tlwe1_host) return fail("null argument");
    if (!c->keys) return fail("keys not loaded");
    if (n == 0) return 0;
    if (set_dev(c)) return 1;
    torus0_t *d_dense = nullptr, *d_pad = nullptr;
    uint32_t* d_u = nullptr;
    BrJob* d_jobs = nullptr;
    std::vector<BrJob> jobs(n);
    for (size_t i = 0; i < n; i++) {
        jobs[i] = BrJob{{(uint32_t)i, 0u, 0u}, {1, 0, 0}, 0, 0u};
    }
    CK(cudaMalloc(&d_dense, n * TLWE0_LEN * sizeof(torus0_t)));
    CK(cudaMalloc(&d_pad, n * SLOT_BYTES));
    CK(cudaMalloc(&d_u, n * (size_t)U_STRIDE * 4));
    CK(cudaMalloc(&d_jobs, n * sizeof(BrJob)));
    CK(cudaMemcpyAsync(d_dense, c_host, n * TLWE0_LEN * sizeof(torus0_t), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_jobs, jobs.data(), n * sizeof(BrJob), cudaMemcpyHostToDevice, c->stream));
    const size_t tot = n * SLOT_STRIDE;
    pad_tlwe0_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(d_dense, d_pad, n);
    c->launches++;
    if (br_dispatch(c, (int)n, d_pad, d_u, d_jobs)) return 1;
    CK(cudaMemcpy2DAsync(tlwe1_host, TLWE1_LEN * 4, d_u, U_STRIDE * 4, TLWE1_LEN * 4, n, cudaMemcpyDeviceToHost,
                         c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_dense);
    cudaFree(d_pad);
    cudaFree(d_u);
    cudaFree(d_jobs);
    return 0;
}

int b200fhe_test_keyswitch(b200fhe_ctx* c, const uint32_t* tlwe1_host, torus0_t* tlwe0_host, size_t n)
{
    if (!c || !tlwe1_host || !tlwe0_host) return fail("null argument");
    if (!c->keys) return fail("keys not loaded");
    if (n == 0) return 0;
    if (set_dev(c)) return 1;
    uint32_t* d_u = nullptr;
    torus0_t* d_out = nullptr;
    KsJob* d_jobs = nullptr;
    std::vector<KsJob> jobs(n);
    for (size_t i = 0; i < n; i++) jobs[i] = KsJob{(uint32_t)i, KS_NONE, (uint32_t)i, 0u};
    CK(cudaMalloc(&d_u, n * (size_t)U_STRIDE * 4));
    CK(cudaMalloc(&d_out, n * SLOT_BYTES));
    CK(cudaMalloc(&d_jobs, n * sizeof(KsJob)));
    CK(cudaMemcpy2DAsync(d_u, U_STRIDE * 4, tlwe1_host, TLWE1_LEN * 4, TLWE1_LEN * 4, n, cudaMemcpyHostToDevice,
                         c->stream));
    CK(cudaMemcpyAsync(d_jobs, jobs.data(), n * sizeof(KsJob), cudaMemcpyHostToDevice, c->stream));
    if (ks_dispatch(c, n, d_jobs, d_u, d_out)) return 1;
    CK(cudaMemcpy2DAsync(tlwe0_host, TLWE0_LEN * sizeof(torus0_t), d_out, SLOT_BYTES, TLWE0_LEN * sizeof(torus0_t), n, cudaMemcpyDeviceToHost,
                         c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_u);
    cudaFree(d_out);
    cudaFree(d_jobs);
    return 0;
}

int b200fhe_test_read_bk_ntt(b200fhe_ctx* c, uint32_t* out_host, size_t first_i, size_t count_i)
{
    if (!c || !out_host) return fail("null argument");
    if (!c->keys) return fail("keys not loaded");
    if (first_i + count_i > (size_t)N0) return fail("range");
    if (set_dev(c)) return 1;
    const size_t per = (size_t)BK_COLS * ROWS * N1;
    CK(cudaMemcpyAsync(out_host, c->d_bk_ntt + first_i * per, count_i * per * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"

#ifdef B200FHE_PHASE_TIMING
// debug builds only: read and clear the per-phase cycle counters (two observer warps x 32 marks)
extern "C" int b200fhe_debug_cta_ns(unsigned long long* out)
{
    return cudaMemcpyFromSymbol(out, g_cta_ns, sizeof(unsigned long long) * 4096) == cudaSuccess ? 0 : -1;
}
extern "C" int b200fhe_debug_phase_cycles(unsigned long long* out, int clear)
{
    if (cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 64) != cudaSuccess) return -1;
    if (clear) {
        unsigned long long z[64] = {};
        if (cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif
