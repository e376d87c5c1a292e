// Host/device portability shims.  Every arithmetic / per-thread "phase" function of the
// gate-bootstrap kernels is written against these so that the SAME source compiles
//   * with nvcc into the sm_100a kernels (b200fhe.cu), and
//   * with plain g++ into the lock-step CPU simulator used by the "not gpu" tests
//     (tests/sim/br_sim.cpp), which executes the kernels thread by thread, phase by phase.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#define B200_UNROLL _Pragma("unroll")
#else
#define B200_HD inline __attribute__((always_inline))
#define B200_UNROLL
#endif

namespace b200 {

// twiddle in Shoup form: w in [0,p), ws = floor(w * 2^32 / p)
struct tw_t {
    uint32_t w, ws;
};

struct alignas(16) u32x4 {
    uint32_t x, y, z, w;
};

// High word of a 32x32 product.  Measured on B200 (scripts/microbench/pipes.cu): IMAD.HI.U32 and
// IMAD.WIDE.U32 both issue at half the rate of IMAD (0.97 vs 1.9 warp-inst/clk/SM), so asking for
// the full 64-bit product instead buys nothing and costs a register pair.
B200_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

}  // namespace b200
