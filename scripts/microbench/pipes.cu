// Instruction-throughput microbenchmark for the integer pipes the NTT kernel leans on (sm_100a).
// Reports warp-instructions per cycle per SM for each op (and two mixes) at several warp counts.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
#define CHAINS 8

template <int OP>
__global__ void k(unsigned* out, unsigned a0, unsigned b0)
{
    unsigned x[CHAINS], y[CHAINS];
    unsigned long long w[CHAINS];
    double f[CHAINS];
    const double fa = 1.0 + 1e-9 * a0, fb = 1e-7 * b0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { x[i] = a0 + threadIdx.x + i; y[i] = b0 * (i + 1); w[i] = x[i]; f[i] = (double)(threadIdx.x + i); }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(b0));
            if (OP == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            if (OP == 2) {  // full 64-bit product, operands change every iteration; + 1 IADD to fold
                unsigned lo, hi;
                asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0,%1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x[i]), "r"(y[i]));
                asm volatile("add.u32 %0, %1, %2;" : "=r"(x[i]) : "r"(lo), "r"(hi));
            }
            if (OP == 9) {  // mul.hi + IADD with the same dependency shape
                unsigned hi;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x[i]), "r"(y[i]));
                asm volatile("add.u32 %0, %1, %2;" : "=r"(x[i]) : "r"(hi), "r"(y[i]));
            }
            if (OP == 10) {  // mad.wide with a 64-bit accumulator and changing multiplicand
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b0));
            }
            if (OP == 11) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(fa), "d"(fb));  // FP64 pipe alone
            if (OP == 12) {  // does the FP64 pipe run beside the integer-multiply pipe?  one DFMA per IMAD.WIDE
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(fa), "d"(fb));
            }
            if (OP == 13) {  // the conversions an FP64 pointwise stage would need: u32 -> f64 and back
                double t;
                asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(t) : "r"(x[i]));
                asm volatile("cvt.rzi.u32.f64 %0, %1;" : "=r"(x[i]) : "d"(t));
            }
            if (OP == 14) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));  // IMAD.WIDE alone
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            if (OP == 4) { asm volatile("sub.u32 %0, %1, %2;" : "=r"(y[i]) : "r"(x[i]), "r"(b0)); asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i])); }
            if (OP == 5) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(b0)); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(b0)); }
            if (OP == 6) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(b0)); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a0)); }
            if (OP == 7) asm volatile("shr.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b0));
            if (OP == 8) {  // Shoup butterfly shape: hi, lo, mad, add, sub-add
                unsigned q, t;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(y[i]), "r"(b0));
                asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(y[i]), "r"(a0));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(b0));
                asm volatile("sub.u32 %0, %1, %2;" : "=r"(y[i]) : "r"(x[i]), "r"(t));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(t));
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += x[i] + y[i] + (unsigned)w[i] + (unsigned)(w[i] >> 32) + (unsigned)(long long)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int instr_per_chain)
{
    unsigned* d;
    cudaMalloc(&d, 148 * 1024 * 4 * 4);
    for (int warps : {4, 8, 16, 32}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<OP><<<148, warps * 32>>>(d, 3, 5);
        cudaEventRecord(e0);
        k<OP><<<148, warps * 32>>>(d, 3, 5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        double cycles = ms * 1e-3 * clk * 1e3;
        double winst = (double)warps * ITER * CHAINS * instr_per_chain;
        printf("%-28s warps/SM=%2d  %.3f warp-inst/clk/SM  (%.2f ms)\n", name, warps, winst / cycles, ms);
    }
    cudaFree(d);
}

int main()
{
    run<0>("IMAD (mad.lo)", 1);
    run<1>("IMAD.HI (mul.hi)", 1);
    run<2>("mul.wide + IADD", 2);
    run<9>("mul.hi + IADD", 2);
    run<10>("mad.wide(acc64) + IADD", 2);
    run<3>("IADD", 1);
    run<4>("sub+min (VIADDMNMX?)", 2);
    run<5>("IMAD + IADD", 2);
    run<6>("IMAD.HI + 2 IADD", 3);
    run<7>("SHF", 1);
    run<8>("Shoup butterfly (5 instr)", 5);
    run<14>("IMAD.WIDE (acc64) alone", 1);
    run<11>("DFMA alone", 1);
    run<12>("IMAD.WIDE + DFMA (1:1)", 2);
    run<13>("cvt u32->f64 + f64->u32", 2);
    return 0;
}
