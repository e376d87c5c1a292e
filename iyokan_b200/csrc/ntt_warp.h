// Warp-level 1024-point negacyclic NTT over Z_p (p = 536856577), 32 points per lane.
//
// Replaces, for the gate-bootstrap path, the reference's polynomial multipliers:
//   CPU : TwistIFFT/TwistFFT -> spqlios double-precision FFT (TFHEpp include/mulfft.hpp:69-134)
//   GPU : NTT1024/NTTInv1024 over 2^64-2^32+1, 128 threads x 8 points, 4 block-synchronised passes
//         (cuFHE include/ntt_gpu/ntt_1024_device.cuh:139-204)
// B200 design: ONE warp owns a whole polynomial.  The merged (twist-free) Cooley-Tukey /
// Gentleman-Sande negacyclic transform is split 32 x 32: five radix-2 stages run entirely in
// registers, one transpose goes through a warp-private padded shared-memory tile (only
// __syncwarp, never __syncthreads), five more stages run in registers.  Butterflies use
// Shoup/Harvey lazy arithmetic in [0, 8p): 3 integer multiplies + 2 adds; the range fixes on the
// sum path are ALU-only (add + unsigned min) because the integer-multiply pipe is the bottleneck.
//
// Index conventions (array position j = 32*a + b, a = j >> 5, b = j & 31):
//   forward : input natural order, output "bit-reversed" order (positions are only ever
//             consumed position-wise by the pointwise stage and by the inverse).
//             stage s (0..9): pairs (j, j + t), t = 512 >> s, twiddle psi_rev[2^s + (j >> (10-s))]
//             pass 1 = stages 0..4, lane = b, registers indexed by a  (twiddles lane-independent)
//             pass 2 = stages 5..9, lane = a, registers indexed by b  (twiddles lane-dependent)
//   inverse : exactly the reverse stage order with inverse twiddles; the 1/N factor is folded
//             into the bootstrapping-key precomputation.
// Tile layout in shared memory: element (a, b) at word a*36 + b  (row padding 32 -> 36 words keeps
// both the column-wise 32-bit accesses and the row-wise 128-bit accesses bank-conflict free).
#pragma once
#include "hd.h"
#include "modarith.h"

namespace b200 {

constexpr int TILE_ROW = 36;                 // padded row pitch in words
constexpr int TILE_WORDS = 32 * TILE_ROW;    // 1152 words = 4608 B per polynomial tile
constexpr int TW2_LEN = 31 * 32;             // lane-dependent twiddles per direction
// entries per table of the table-driven first two forward stages: one per digit value.  Only the 6-bit gadget of the
// 128-bit set uses them (five 64-entry tables); for wider gadgets the tables are stubs and fwd_start_r4_group must not be
// called (the 80-bit flavour compiles the specialised phase headers only for the simulator and runs brg_kernel alone)
constexpr bool R4_ENABLED = BGBIT <= 6;
constexpr int R4_DIGITS = R4_ENABLED ? (1 << BGBIT) : 4;
constexpr int R4_WORDS = 5 * R4_DIGITS;

B200_HD int tile_idx(int a, int b) { return a * TILE_ROW + b; }
B200_HD int tile_of_j(int j) { return (j >> 5) * TILE_ROW + (j & 31); }

// ---- twiddle storage -------------------------------------------------------------------
// psi_rev index idx < 32  : lane-independent, __constant__ on the device
// psi_rev index idx >= 32 : lane-dependent, table tw2[((2^(s-5) - 1) + h) * 32 + lane] = psi_rev[2^s + lane*2^(s-5) + h]
struct NttTables {
    tw_t fwd[1024];      // psi_rev[idx]      (idx 1..1023)
    tw_t inv[1024];      // psi_rev[idx]^-1
    tw_t tw2f[TW2_LEN];  // lane-dependent forward table
    tw_t tw2i[TW2_LEN];  // lane-dependent inverse table
    tw_t bk_scale;       // 2^32 * N^-1 mod p: folded into the NTT-domain bootstrapping key
    uint32_t r4[R4_WORDS];  // products of the 2^Bgbit possible digits with the twiddles of forward stages 0 and 1 (fwd_start_r4)
};

inline tw_t h_twf_u[32];
inline tw_t h_twi_u[32];
#if defined(__CUDACC__)
__constant__ tw_t c_twf_u[32];
__constant__ tw_t c_twi_u[32];
#endif

B200_HD tw_t twf_u(int idx)
{
#if defined(__CUDA_ARCH__)
    return c_twf_u[idx];
#else
    return h_twf_u[idx];
#endif
}
B200_HD tw_t twi_u(int idx)
{
#if defined(__CUDA_ARCH__)
    return c_twi_u[idx];
#else
    return h_twi_u[idx];
#endif
}

inline uint32_t bitrev10(uint32_t x)
{
    uint32_t r = 0;
    for (int i = 0; i < 10; i++) r |= ((x >> i) & 1u) << (9 - i);
    return r;
}

// Build all tables on the host (also fills h_twf_u / h_twi_u used by the CPU simulator).
inline void ntt_tables_init(NttTables& t)
{
    for (uint32_t idx = 0; idx < 1024; idx++) {
        const uint32_t w = mod_pow(PSI, bitrev10(idx));
        t.fwd[idx] = make_tw(w);
        t.inv[idx] = make_tw(mod_inv(w));
    }
    for (int idx = 0; idx < 32; idx++) {
        h_twf_u[idx] = t.fwd[idx];
        h_twi_u[idx] = t.inv[idx];
    }
    for (int ls = 0; ls < 5; ls++)          // global stage s = 5 + ls
        for (int h = 0; h < (1 << ls); h++)
            for (int lane = 0; lane < 32; lane++) {
                const int idx = (32 << ls) + (lane << ls) + h;
                const int pos = ((1 << ls) - 1 + h) * 32 + lane;
                t.tw2f[pos] = t.fwd[idx];
                t.tw2i[pos] = t.inv[idx];
            }
    const uint32_t two32 = (uint32_t)((1ull << 32) % P);
    t.bk_scale = make_tw(mod_mul(two32, mod_inv(1024)));
    // fwd_start_r4: digit field value i stands for the digit i - Bg/2; tables E, A, B, C, D of its products, exact in [0, p)
    const uint32_t w1 = t.fwd[1].w, w2 = t.fwd[2].w, w3 = t.fwd[3].w;
    for (uint32_t i = 0; i < (uint32_t)R4_DIGITS; i++) {
        const uint32_t half = R4_DIGITS / 2, d = i >= half ? i - half : P - (half - i);
        const uint32_t m31 = mod_mul(mod_mul(w3, w1), d);
        t.r4[0 * R4_DIGITS + i] = mod_mul(w1, d);
        t.r4[1 * R4_DIGITS + i] = mod_mul(w2, d);
        t.r4[2 * R4_DIGITS + i] = mod_mul(mod_mul(w2, w1), d);
        t.r4[3 * R4_DIGITS + i] = mod_mul(w3, d);
        t.r4[4 * R4_DIGITS + i] = m31 ? P - m31 : 0u;
    }
}

// ---- register stages -------------------------------------------------------------------
// One radix-2 stage over the 32 registers of a lane.  LS = local stage 0..4: pairs (i, i + half),
// half = 16 >> LS, arranged in 2^LS groups that share a twiddle.  tw(g) returns the group twiddle.
// FIX: 0 = none, 1 = sum-path operand < 8p -> < 4p, 2 = < 8p -> < 2p (both ALU-only)
template <int FIX>
B200_HD uint32_t apply_fix(uint32_t x)
{
    if (FIX == 1) return fix_lt8p_to_lt4p(x);
    if (FIX == 2) return fix_lt8p_to_lt2p(x);
    return x;
}

template <int LS, int FIX, class TwFn>
B200_HD void ct_stage(uint32_t (&x)[32], TwFn tw)
{
    constexpr int half = 16 >> LS;
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

template <int LS, int FIX, class TwFn>
B200_HD void gs_stage(uint32_t (&x)[32], TwFn tw)
{
    constexpr int half = 16 >> LS;
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

// Lazy-range schedule (all values < 8p < 2^32; Shoup products are always < 2p):
//   forward, input < p + 1024:  s0 <3p | s1 <5p | s2 <7p | fix1 s3 <6p | s4 <8p | fix1 s5 <6p | s6 <8p |
//                               fix1 s7 <6p | s8 <8p | fix2 s9 <4p                       => output < 4p
//   inverse, input < 4p: every stage forms S = U + V < 8p and folds it back below 4p (fix1);
//                        the product operand U - V + 4p stays in (0, 8p)                  => output < 4p
// Forward pass 1: global stages 0..4 (lane-independent twiddles from constant memory).
B200_HD void fwd_pass1(uint32_t (&x)[32])
{
    ct_stage<0, 0>(x, [](int g) { return twf_u(1 + g); });
    ct_stage<1, 0>(x, [](int g) { return twf_u(2 + g); });
    ct_stage<2, 0>(x, [](int g) { return twf_u(4 + g); });
    ct_stage<3, 1>(x, [](int g) { return twf_u(8 + g); });
    ct_stage<4, 0>(x, [](int g) { return twf_u(16 + g); });
}

// Forward pass 2: global stages 5..9, lane-dependent twiddles read from tw2 (shared memory).
B200_HD void fwd_pass2(uint32_t (&x)[32], const tw_t* tw2, int lane)
{
    ct_stage<0, 1>(x, [=](int g) { return tw2[(0 + g) * 32 + lane]; });
    ct_stage<1, 0>(x, [=](int g) { return tw2[(1 + g) * 32 + lane]; });
    ct_stage<2, 1>(x, [=](int g) { return tw2[(3 + g) * 32 + lane]; });
    ct_stage<3, 0>(x, [=](int g) { return tw2[(7 + g) * 32 + lane]; });
    ct_stage<4, 2>(x, [=](int g) { return tw2[(15 + g) * 32 + lane]; });
}

// Inverse pass 1: global stages 9..5 (lane-dependent).  Input < 4p, output < 4p.
B200_HD void inv_pass1(uint32_t (&x)[32], const tw_t* tw2, int lane)
{
    gs_stage<4, 1>(x, [=](int g) { return tw2[(15 + g) * 32 + lane]; });
    gs_stage<3, 1>(x, [=](int g) { return tw2[(7 + g) * 32 + lane]; });
    gs_stage<2, 1>(x, [=](int g) { return tw2[(3 + g) * 32 + lane]; });
    gs_stage<1, 1>(x, [=](int g) { return tw2[(1 + g) * 32 + lane]; });
    gs_stage<0, 1>(x, [=](int g) { return tw2[(0 + g) * 32 + lane]; });
}

// Inverse pass 2: global stages 4..0 (lane-independent).  Input < 4p, output < 4p.
B200_HD void inv_pass2(uint32_t (&x)[32])
{
    gs_stage<4, 1>(x, [](int g) { return twi_u(16 + g); });
    gs_stage<3, 1>(x, [](int g) { return twi_u(8 + g); });
    gs_stage<2, 1>(x, [](int g) { return twi_u(4 + g); });
    gs_stage<1, 1>(x, [](int g) { return twi_u(2 + g); });
    gs_stage<0, 1>(x, [](int g) { return twi_u(1 + g); });
}

// ---- three transforms interleaved in one lane -------------------------------------------
// A warp that owns the three digit (or limb) polynomials of one accumulator polynomial can run
// them in lock step: 48 independent butterflies per stage instead of 16 (more ILP per warp, so
// fewer warps saturate the multiply pipe) and ONE twiddle fetch per group instead of three.
template <int LS, int FIX, class TwFn>
B200_HD void ct_stage3(uint32_t (&x0)[32], uint32_t (&x1)[32], uint32_t (&x2)[32], TwFn tw)
{
    constexpr int half = 16 >> LS;
    B200_UNROLL
    for (int g = 0; g < (1 << LS); g++) {
        const tw_t w = tw(g);
        B200_UNROLL
        for (int k = 0; k < half; k++) {
            const int i0 = g * 2 * half + k, i1 = i0 + half;
            const uint32_t X0 = apply_fix<FIX>(x0[i0]), X1 = apply_fix<FIX>(x1[i0]), X2 = apply_fix<FIX>(x2[i0]);
            const uint32_t T0 = shoup_mul(x0[i1], w), T1 = shoup_mul(x1[i1], w), T2 = shoup_mul(x2[i1], w);
            x0[i0] = X0 + T0;
            x0[i1] = X0 - T0 + P2;
            x1[i0] = X1 + T1;
            x1[i1] = X1 - T1 + P2;
            x2[i0] = X2 + T2;
            x2[i1] = X2 - T2 + P2;
        }
    }
}
template <int LS, int FIX, class TwFn>
B200_HD void gs_stage3(uint32_t (&x0)[32], uint32_t (&x1)[32], uint32_t (&x2)[32], TwFn tw)
{
    constexpr int half = 16 >> LS;
    B200_UNROLL
    for (int g = 0; g < (1 << LS); g++) {
        const tw_t w = tw(g);
        B200_UNROLL
        for (int k = 0; k < half; k++) {
            const int i0 = g * 2 * half + k, i1 = i0 + half;
            const uint32_t U0 = x0[i0], V0 = x0[i1], U1 = x1[i0], V1 = x1[i1], U2 = x2[i0], V2 = x2[i1];
            x0[i0] = apply_fix<FIX>(U0 + V0);
            x1[i0] = apply_fix<FIX>(U1 + V1);
            x2[i0] = apply_fix<FIX>(U2 + V2);
            x0[i1] = shoup_mul(U0 - V0 + P4, w);
            x1[i1] = shoup_mul(U1 - V1 + P4, w);
            x2[i1] = shoup_mul(U2 - V2 + P4, w);
        }
    }
}
#define B200_X3 uint32_t (&x0)[32], uint32_t (&x1)[32], uint32_t (&x2)[32]
B200_HD void fwd_pass1_x3(B200_X3)
{
    ct_stage3<0, 0>(x0, x1, x2, [](int g) { return twf_u(1 + g); });
    ct_stage3<1, 0>(x0, x1, x2, [](int g) { return twf_u(2 + g); });
    ct_stage3<2, 0>(x0, x1, x2, [](int g) { return twf_u(4 + g); });
    ct_stage3<3, 1>(x0, x1, x2, [](int g) { return twf_u(8 + g); });
    ct_stage3<4, 0>(x0, x1, x2, [](int g) { return twf_u(16 + g); });
}
B200_HD void fwd_pass2_x3(B200_X3, const tw_t* tw2, int lane)
{
    ct_stage3<0, 1>(x0, x1, x2, [=](int g) { return tw2[(0 + g) * 32 + lane]; });
    ct_stage3<1, 0>(x0, x1, x2, [=](int g) { return tw2[(1 + g) * 32 + lane]; });
    ct_stage3<2, 1>(x0, x1, x2, [=](int g) { return tw2[(3 + g) * 32 + lane]; });
    ct_stage3<3, 0>(x0, x1, x2, [=](int g) { return tw2[(7 + g) * 32 + lane]; });
    ct_stage3<4, 2>(x0, x1, x2, [=](int g) { return tw2[(15 + g) * 32 + lane]; });
}
B200_HD void inv_pass1_x3(B200_X3, const tw_t* tw2, int lane)
{
    gs_stage3<4, 1>(x0, x1, x2, [=](int g) { return tw2[(15 + g) * 32 + lane]; });
    gs_stage3<3, 1>(x0, x1, x2, [=](int g) { return tw2[(7 + g) * 32 + lane]; });
    gs_stage3<2, 1>(x0, x1, x2, [=](int g) { return tw2[(3 + g) * 32 + lane]; });
    gs_stage3<1, 1>(x0, x1, x2, [=](int g) { return tw2[(1 + g) * 32 + lane]; });
    gs_stage3<0, 1>(x0, x1, x2, [=](int g) { return tw2[(0 + g) * 32 + lane]; });
}
B200_HD void inv_pass2_x3(B200_X3)
{
    gs_stage3<4, 1>(x0, x1, x2, [](int g) { return twi_u(16 + g); });
    gs_stage3<3, 1>(x0, x1, x2, [](int g) { return twi_u(8 + g); });
    gs_stage3<2, 1>(x0, x1, x2, [](int g) { return twi_u(4 + g); });
    gs_stage3<1, 1>(x0, x1, x2, [](int g) { return twi_u(2 + g); });
    gs_stage3<0, 1>(x0, x1, x2, [](int g) { return twi_u(1 + g); });
}

// ---- two transforms interleaved in one lane (br7_kernel: 16 warps per SM need <= 128 registers,
// so a warp runs two of its three transforms in lock step and the third one alone) -------------
template <int LS, int FIX, class TwFn>
B200_HD void ct_stage2(uint32_t (&x0)[32], uint32_t (&x1)[32], TwFn tw)
{
    constexpr int half = 16 >> LS;
    B200_UNROLL
    for (int g = 0; g < (1 << LS); g++) {
        const tw_t w = tw(g);
        B200_UNROLL
        for (int k = 0; k < half; k++) {
            const int i0 = g * 2 * half + k, i1 = i0 + half;
            const uint32_t X0 = apply_fix<FIX>(x0[i0]), X1 = apply_fix<FIX>(x1[i0]);
            const uint32_t T0 = shoup_mul(x0[i1], w), T1 = shoup_mul(x1[i1], w);
            x0[i0] = X0 + T0;
            x0[i1] = X0 - T0 + P2;
            x1[i0] = X1 + T1;
            x1[i1] = X1 - T1 + P2;
        }
    }
}
template <int LS, int FIX, class TwFn>
B200_HD void gs_stage2(uint32_t (&x0)[32], uint32_t (&x1)[32], TwFn tw)
{
    constexpr int half = 16 >> LS;
    B200_UNROLL
    for (int g = 0; g < (1 << LS); g++) {
        const tw_t w = tw(g);
        B200_UNROLL
        for (int k = 0; k < half; k++) {
            const int i0 = g * 2 * half + k, i1 = i0 + half;
            const uint32_t U0 = x0[i0], V0 = x0[i1], U1 = x1[i0], V1 = x1[i1];
            x0[i0] = apply_fix<FIX>(U0 + V0);
            x1[i0] = apply_fix<FIX>(U1 + V1);
            x0[i1] = shoup_mul(U0 - V0 + P4, w);
            x1[i1] = shoup_mul(U1 - V1 + P4, w);
        }
    }
}
#define B200_X2 uint32_t (&x0)[32], uint32_t (&x1)[32]
B200_HD void fwd_pass1_x2(B200_X2)
{
    ct_stage2<0, 0>(x0, x1, [](int g) { return twf_u(1 + g); });
    ct_stage2<1, 0>(x0, x1, [](int g) { return twf_u(2 + g); });
    ct_stage2<2, 0>(x0, x1, [](int g) { return twf_u(4 + g); });
    ct_stage2<3, 1>(x0, x1, [](int g) { return twf_u(8 + g); });
    ct_stage2<4, 0>(x0, x1, [](int g) { return twf_u(16 + g); });
}
B200_HD void fwd_pass2_x2(B200_X2, const tw_t* tw2, int lane)
{
    ct_stage2<0, 1>(x0, x1, [=](int g) { return tw2[(0 + g) * 32 + lane]; });
    ct_stage2<1, 0>(x0, x1, [=](int g) { return tw2[(1 + g) * 32 + lane]; });
    ct_stage2<2, 1>(x0, x1, [=](int g) { return tw2[(3 + g) * 32 + lane]; });
    ct_stage2<3, 0>(x0, x1, [=](int g) { return tw2[(7 + g) * 32 + lane]; });
    ct_stage2<4, 2>(x0, x1, [=](int g) { return tw2[(15 + g) * 32 + lane]; });
}
B200_HD void inv_pass1_x2(B200_X2, const tw_t* tw2, int lane)
{
    gs_stage2<4, 1>(x0, x1, [=](int g) { return tw2[(15 + g) * 32 + lane]; });
    gs_stage2<3, 1>(x0, x1, [=](int g) { return tw2[(7 + g) * 32 + lane]; });
    gs_stage2<2, 1>(x0, x1, [=](int g) { return tw2[(3 + g) * 32 + lane]; });
    gs_stage2<1, 1>(x0, x1, [=](int g) { return tw2[(1 + g) * 32 + lane]; });
    gs_stage2<0, 1>(x0, x1, [=](int g) { return tw2[(0 + g) * 32 + lane]; });
}
B200_HD void inv_pass2_x2(B200_X2)
{
    gs_stage2<4, 1>(x0, x1, [](int g) { return twi_u(16 + g); });
    gs_stage2<3, 1>(x0, x1, [](int g) { return twi_u(8 + g); });
    gs_stage2<2, 1>(x0, x1, [](int g) { return twi_u(4 + g); });
    gs_stage2<1, 1>(x0, x1, [](int g) { return twi_u(2 + g); });
    gs_stage2<0, 1>(x0, x1, [](int g) { return twi_u(1 + g); });
}

// ---- table-driven first two forward stages ------------------------------------------------
// The inputs of a forward transform are gadget digits: 2^Bgbit possible values.  Stages 0 and 1 only combine four of
// them (registers k, k+8, k+16, k+24) with the three twiddles w1, w2, w3 = psi_rev[1..3]:
//   z[k]    = d0 + w1 d2 + w2 d1 + w2 w1 d3        z[k+8]  = d0 + w1 d2 - (w2 d1 + w2 w1 d3)
//   z[k+16] = d0 - w1 d2 + w3 d1 - w3 w1 d3        z[k+24] = d0 - w1 d2 - (w3 d1 - w3 w1 d3)
// so the four Shoup multiplications of a group become five look-ups in tables of 2^Bgbit exact residues
// (NttTables::r4 = E | A | B | C | D; a shared-memory bank holds two entries of a 64-entry table, so a warp-wide look-up
// is at most two wavefronts).  Outputs are < 4p + Bg/2 < 5p: the bound stage 1 leaves in the lazy-range schedule above,
// and congruent mod p to what ct_stage<0>, ct_stage<1> produce.  SHIFT = bit position of the digit field in dv.
template <int SHIFT>
B200_HD void fwd_start_r4_group(const uint32_t* r4, uint32_t dv0, uint32_t dv1, uint32_t dv2, uint32_t dv3, uint32_t& z0,
                                uint32_t& z1, uint32_t& z2, uint32_t& z3)
{
    constexpr uint32_t M = R4_DIGITS - 1;
    const uint32_t i0 = (dv0 >> SHIFT) & M, i1 = (dv1 >> SHIFT) & M, i2 = (dv2 >> SHIFT) & M, i3 = (dv3 >> SHIFT) & M;
    const uint32_t X0 = i0 + (P - R4_DIGITS / 2);
    const uint32_t E = r4[i2];
    const uint32_t T1 = r4[R4_DIGITS + i1] + r4[2 * R4_DIGITS + i3];
    const uint32_t T3 = r4[3 * R4_DIGITS + i1] + r4[4 * R4_DIGITS + i3];
    const uint32_t u0 = X0 + E, u2 = X0 - E + P;
    z0 = u0 + T1;
    z1 = u0 - T1 + P2;
    z2 = u2 + T3;
    z3 = u2 - T3 + P2;
}
// stages 2..4 of pass 1 (after fwd_start_r4_group filled the registers)
B200_HD void fwd_pass1_tail(uint32_t (&x)[32])
{
    ct_stage<2, 0>(x, [](int g) { return twf_u(4 + g); });
    ct_stage<3, 1>(x, [](int g) { return twf_u(8 + g); });
    ct_stage<4, 0>(x, [](int g) { return twf_u(16 + g); });
}
B200_HD void fwd_pass1_tail_x2(B200_X2)
{
    ct_stage2<2, 0>(x0, x1, [](int g) { return twf_u(4 + g); });
    ct_stage2<3, 1>(x0, x1, [](int g) { return twf_u(8 + g); });
    ct_stage2<4, 0>(x0, x1, [](int g) { return twf_u(16 + g); });
}

// ---- XOR-swizzled tile (br7_kernel): 32 x 32 words with NO row padding.  Element (a, b) lives at
// word a*32 + (b ^ ((a & 7) << 2)): the 16-byte chunk index is XORed with the low row bits, so the
// row-wise 128-bit accesses of 8 consecutive rows hit 8 different bank groups, while the column-wise
// 32-bit accesses of a warp stay a permutation of one row (conflict free either way).
constexpr int STILE_WORDS = 1024;
B200_HD int stile_of_j(int j) { return j ^ (((j >> 5) & 7) << 2); }
B200_HD void stile_store_col(uint32_t* tile, const uint32_t (&x)[32], int lane)
{
    B200_UNROLL
    for (int a = 0; a < 32; a++) tile[a * 32 + (lane ^ ((a & 7) << 2))] = x[a];
}
B200_HD void stile_load_col(const uint32_t* tile, uint32_t (&x)[32], int lane)
{
    B200_UNROLL
    for (int a = 0; a < 32; a++) x[a] = tile[a * 32 + (lane ^ ((a & 7) << 2))];
}
B200_HD void stile_store_row(uint32_t* tile, const uint32_t (&x)[32], int lane)
{
    u32x4* row = reinterpret_cast<u32x4*>(tile + lane * 32);
    const int s = lane & 7;
    B200_UNROLL
    for (int k = 0; k < 8; k++) row[k ^ s] = u32x4{x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]};
}
B200_HD void stile_load_row(const uint32_t* tile, uint32_t (&x)[32], int lane)
{
    const u32x4* row = reinterpret_cast<const u32x4*>(tile + lane * 32);
    const int s = lane & 7;
    B200_UNROLL
    for (int k = 0; k < 8; k++) {
        const u32x4 v = row[k ^ s];
        x[4 * k] = v.x;
        x[4 * k + 1] = v.y;
        x[4 * k + 2] = v.z;
        x[4 * k + 3] = v.w;
    }
}

// ---- tile access -----------------------------------------------------------------------
// column access: register index = a, lane = b  (32-bit, conflict free)
B200_HD void tile_store_col(uint32_t* tile, const uint32_t (&x)[32], int lane)
{
    B200_UNROLL
    for (int a = 0; a < 32; a++) tile[tile_idx(a, lane)] = x[a];
}
B200_HD void tile_load_col(const uint32_t* tile, uint32_t (&x)[32], int lane)
{
    B200_UNROLL
    for (int a = 0; a < 32; a++) x[a] = tile[tile_idx(a, lane)];
}
// row access: register index = b, lane = a  (128-bit, conflict free thanks to the 36-word pitch)
B200_HD void tile_store_row(uint32_t* tile, const uint32_t (&x)[32], int lane)
{
    u32x4* row = reinterpret_cast<u32x4*>(tile + lane * TILE_ROW);
    B200_UNROLL
    for (int k = 0; k < 8; k++) row[k] = u32x4{x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]};
}
B200_HD void tile_load_row(const uint32_t* tile, uint32_t (&x)[32], int lane)
{
    const u32x4* row = reinterpret_cast<const u32x4*>(tile + lane * TILE_ROW);
    B200_UNROLL
    for (int k = 0; k < 8; k++) {
        const u32x4 v = row[k];
        x[4 * k] = v.x;
        x[4 * k + 1] = v.y;
        x[4 * k + 2] = v.z;
        x[4 * k + 3] = v.w;
    }
}

}  // namespace b200
