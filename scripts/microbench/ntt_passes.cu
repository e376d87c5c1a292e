// How close can the real NTT passes (ntt_warp.h) get to the integer-multiply pipe bound?
// Runs each pass type back to back on registers (and, optionally, with the tile traffic of the
// kernel) and reports SM cycles per warp-pass against the 640-cycle IMAD bound (80 butterflies x 8).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../iyokan_b200/csrc/ntt_warp.h"
using namespace b200;

// 32 x 32 transpose (register index <-> lane) with warp shuffles only: five exchange rounds of 16 SHFL + 32 SEL.
// north_star names warp-shuffle butterflies; this is what replacing the warp-private tile by shuffles costs.
__device__ __forceinline__ void transpose_shfl(uint32_t (&x)[32], int lane)
{
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) {
        const bool up = (lane & k) != 0;
#pragma unroll
        for (int a = 0; a < 32; a++) {
            if (a & k) continue;
            const uint32_t send = up ? x[a] : x[a | k];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, k);
            if (up) x[a] = recv; else x[a | k] = recv;
        }
    }
}

template <int MODE>
__global__ void k(uint32_t* out, const tw_t* tw2f_g, const tw_t* tw2i_g, int iters)
{
    extern __shared__ __align__(16) uint8_t smem[];
    tw_t* tw2f = reinterpret_cast<tw_t*>(smem);
    tw_t* tw2i = tw2f + TW2_LEN;
    uint32_t* tiles = reinterpret_cast<uint32_t*>(tw2i + TW2_LEN);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < TW2_LEN; i += blockDim.x) { tw2f[i] = tw2f_g[i]; tw2i[i] = tw2i_g[i]; }
    __syncthreads();
    uint32_t* tile = tiles + warp * TILE_WORDS;
    uint32_t x[32];
#pragma unroll
    for (int a = 0; a < 32; a++) x[a] = (threadIdx.x * 32 + a) % P;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) fwd_pass1(x);
        if (MODE == 1) fwd_pass2(x, tw2f, lane);
        if (MODE == 2) inv_pass1(x, tw2i, lane);
        if (MODE == 3) inv_pass2(x);
        if (MODE == 4) {  // full forward transform with tile traffic, as in phase F
            fwd_pass1(x); tile_store_col(tile, x, lane); __syncwarp();
            tile_load_row(tile, x, lane); fwd_pass2(x, tw2f, lane); tile_store_row(tile, x, lane); __syncwarp();
            tile_load_col(tile, x, lane);
        }
        if (MODE == 6) {  // full forward transform, both transposes by shuffles instead of the tile
            fwd_pass1(x); transpose_shfl(x, lane);
            fwd_pass2(x, tw2f, lane); transpose_shfl(x, lane);
        }
        if (MODE == 5) {  // full inverse transform with tile traffic, as in phase I
            tile_load_row(tile, x, lane); inv_pass1(x, tw2i, lane); tile_store_row(tile, x, lane); __syncwarp();
            tile_load_col(tile, x, lane); inv_pass2(x); tile_store_col(tile, x, lane); __syncwarp();
        }
#pragma unroll
        for (int a = 0; a < 32; a++) x[a] = fix_lt8p_to_lt2p(x[a]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int a = 0; a < 32; a++) s += x[a];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int passes, const tw_t* f, const tw_t* i)
{
    uint32_t* d; cudaMalloc(&d, 148 * 1024 * 4);
    for (int warps : {4, 8, 12, 16, 24}) {
        size_t smem = 2 * TW2_LEN * sizeof(tw_t) + (size_t)warps * TILE_WORDS * 4;
        cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int iters = 2000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<MODE><<<148, warps * 32, smem>>>(d, f, i, 10);
        cudaEventRecord(e0);
        k<MODE><<<148, warps * 32, smem>>>(d, f, i, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        double cyc = ms * 1e-3 * clk * 1e3;
        double per = cyc / ((double)iters * warps * passes) ;   // SM cycles per warp-pass
        printf("%-34s warps/SM=%2d  %.1f SM-cycles per warp-pass (IMAD bound 160.0)  util %.2f  err=%s\n", name, warps, per, 160.0 / per,
               cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d);
}

int main()
{
    NttTables* t = new NttTables(); ntt_tables_init(*t);
    cudaMemcpyToSymbol(c_twf_u, h_twf_u, sizeof(h_twf_u)); cudaMemcpyToSymbol(c_twi_u, h_twi_u, sizeof(h_twi_u));
    tw_t *f, *i; cudaMalloc(&f, sizeof(t->tw2f)); cudaMalloc(&i, sizeof(t->tw2i));
    cudaMemcpy(f, t->tw2f, sizeof(t->tw2f), cudaMemcpyHostToDevice); cudaMemcpy(i, t->tw2i, sizeof(t->tw2i), cudaMemcpyHostToDevice);
    run<0>("fwd_pass1 (regs, const twiddles)", 1, f, i);
    run<1>("fwd_pass2 (regs, smem twiddles)", 1, f, i);
    run<2>("inv_pass1 (regs, smem twiddles)", 1, f, i);
    run<3>("inv_pass2 (regs, const twiddles)", 1, f, i);
    run<4>("forward NTT + tile traffic", 2, f, i);
    run<5>("inverse NTT + tile traffic", 2, f, i);
    run<6>("forward NTT + shuffle transposes", 2, f, i);
    return 0;
}
