// Per-thread phases of br7_kernel: the 16-warp throughput shape of the blind rotation.
//
// Same arithmetic, tiles-in-place protocol and key layout as br3_kernel (br_phases.h; reference
// functions cited there: TFHEpp gatebootstrapping.hpp:19-71, detwfa.hpp:36-49, trgsw.hpp:62-131), but
// sized so that EIGHT jobs (16 warps) are resident per SM instead of six (12 warps): ncu showed the
// 12-warp shape losing a quarter of the integer-multiply pipe to fixed-latency waits that three warps
// per scheduler cannot cover (profiles/r01_br_kernel_auto.md).  Two changes make room:
//   * tiles are XOR-swizzled instead of row-padded (4096 B instead of 4608 B per polynomial):
//     8 jobs x 6 tiles = 192 KB + twiddles + mod-switched a_i = 222,720 B of the 227 KB per CTA;
//   * a warp runs two of its three transforms in lock step and the third alone (x2 + x1 instead of
//     x3), so accumulator (32) + data (64) registers fit the 128-register budget of 512 threads.
// Digit 0 shares its tile with the natural-order accumulator copy; the rotated difference is computed once, kept in
// registers across digit 0's (x1) passes and reused for digits 1 and 2; limb 2 is inverted last.  The first two forward
// stages are table look-ups on the 6-bit digits (ntt_warp.h fwd_start_r4_group).
//
// Step sequence (phases separated by __syncwarp unless noted):
//   F0a   rotated difference (kept) -> digit 0 -> pass 1            (only read of the accumulator copy)
//   F0b   column store into tile 3q
//   F0c   row load -> pass 2 -> row store
//   F12a  digits 1,2 from the kept difference -> pass 1 (x2) -> column stores into tiles 3q+1, 3q+2
//   F12c  row loads -> pass 2 (x2) -> row stores   ; __syncthreads ; pointwise ; __syncthreads
//   I01a  row loads of limbs 0,1 -> inverse pass 1 (x2) -> row stores
//   I01b  column loads -> inverse pass 2 (x2) -> lift, acc += v0 + (v1 << 11)
//   I2a   row load of limb 2 -> inverse pass 1 -> row store
//   I2b   column load -> inverse pass 2 -> lift, acc += v2 << 22 ; refresh the accumulator copy in tile 3q
#pragma once
#include "br_phases.h"

namespace b200 {

template <int G>
struct Br7Smem {
    static constexpr int DBUF_WORDS = G * ROWS * STILE_WORDS;
    static constexpr int ABAR_HALFS = G * SLOT_STRIDE;
    static constexpr size_t BYTES =
        (size_t)DBUF_WORDS * 4 + 2 * (size_t)TW2_LEN * sizeof(tw_t) + (size_t)R4_WORDS * 4 + (size_t)ABAR_HALFS * 2;
    uint32_t* dbuf;
    tw_t* tw2f;
    tw_t* tw2i;
    uint32_t* r4;    // digit x twiddle tables of the first two forward stages (ntt_warp.h fwd_start_r4_group)
    uint16_t* abar;
    B200_HD void carve(void* base)
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(base);
        dbuf = reinterpret_cast<uint32_t*>(p);
        p += (size_t)DBUF_WORDS * 4;
        tw2f = reinterpret_cast<tw_t*>(p);
        p += (size_t)TW2_LEN * sizeof(tw_t);
        tw2i = reinterpret_cast<tw_t*>(p);
        p += (size_t)TW2_LEN * sizeof(tw_t);
        r4 = reinterpret_cast<uint32_t*>(p);
        p += (size_t)R4_WORDS * 4;
        abar = reinterpret_cast<uint16_t*>(p);
    }
    B200_HD uint32_t* tile(int g, int row) const { return dbuf + (size_t)(g * ROWS + row) * STILE_WORDS; }
    B200_HD uint32_t* acc(int g, int q) const { return tile(g, q * GL); }  // natural order, whole tile
};

// prologue: identical to br_prologue (mod switch, test vector), on the swizzled carve-up
template <int G>
B200_HD void br7_prologue(const Br7Smem<G>& sm, const BrJob& job, const uint16_t* arena, int g, int q, int lane,
                          uint32_t (&accr)[32])
{
    for (int i = q * 32 + lane; i < N0; i += 64) {
        const uint32_t c = br_lincomb(job, arena, i);
        sm.abar[g * SLOT_STRIDE + i] = (uint16_t)((c + 16u) >> 5);
    }
    const uint32_t bbar = 2u * N1 - (br_lincomb(job, arena, N0) >> 5);
    uint32_t* acc = sm.acc(g, q);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        const int n = 32 * a + lane;
        uint32_t v = 0;
        if (q == 1) {
            const uint32_t m = ((uint32_t)n - bbar) & (2u * N1 - 1);
            v = (m & N1) ? (0u - MU1) : MU1;
        }
        accr[a] = v;
        acc[n] = v;
    }
}

// (X^abar - 1) * acc + decomposition offsets, coefficient 32a + lane (utils.hpp:130-144, trgsw.hpp:62-78)
B200_HD uint32_t br7_dv(const uint32_t* acc, uint32_t base, int a, uint32_t accr_a)
{
    const uint32_t m = (base + 32u * a) & (2u * N1 - 1);
    const uint32_t v = acc[m & (N1 - 1)];
    const uint32_t neg = 0u - ((m >> NBIT) & 1u);
    return ((v ^ neg) - neg) - accr_a + (DEC_OFFSET + DEC_ROUND);
}

// F12c
template <int G>
B200_HD void br7_fwd12_c(const Br7Smem<G>& sm, int g, int q, int lane)
{
    uint32_t x1[32], x2[32];
    uint32_t* t = sm.tile(g, q * GL + 1);
    stile_load_row(t, x1, lane);
    stile_load_row(t + STILE_WORDS, x2, lane);
    fwd_pass2_x2(x1, x2, sm.tw2f, lane);
    stile_store_row(t, x1, lane);
    stile_store_row(t + STILE_WORDS, x2, lane);
}
template <int G>
B200_HD void br7_fwd0_b(const Br7Smem<G>& sm, int g, int q, int lane, const uint32_t (&x0)[32])
{
    stile_store_col(sm.tile(g, q * GL), x0, lane);
}
template <int G>
B200_HD void br7_fwd0_c(const Br7Smem<G>& sm, int g, int q, int lane)
{
    uint32_t x0[32];
    uint32_t* t = sm.tile(g, q * GL);
    stile_load_row(t, x0, lane);
    fwd_pass2(x0, sm.tw2f, lane);
    stile_store_row(t, x0, lane);
}

// F0a: the rotated difference is computed ONCE and kept in registers (dv); digit 0 -> stages 0,1 by table look-up -> rest
// of pass 1.  Digit 0 goes first because its x1 passes leave room for the 32 registers of dv; x0 is returned so the caller
// can __syncwarp (last read of the accumulator copy that shares tile 3q with digit 0) before the store.
template <int G>
B200_HD void br7_fwd0_a(const Br7Smem<G>& sm, int i, int g, int q, int lane, const uint32_t (&accr)[32], uint32_t (&dv)[32],
                        uint32_t (&x0)[32])
{
    const uint32_t abar = sm.abar[g * SLOT_STRIDE + i];
    const uint32_t* acc = sm.acc(g, q);
    const uint32_t base = ((uint32_t)lane - abar) & (2u * N1 - 1);
    B200_UNROLL
    for (int a = 0; a < 32; a++) dv[a] = br7_dv(acc, base, a, accr[a]);
    B200_UNROLL
    for (int k = 0; k < 8; k++)
        fwd_start_r4_group<32 - BGBIT>(sm.r4, dv[k], dv[k + 8], dv[k + 16], dv[k + 24], x0[k], x0[k + 8], x0[k + 16], x0[k + 24]);
    fwd_pass1_tail(x0);
}
// F12a: digits 1 and 2 from the kept difference -> table look-up -> rest of pass 1 (x2) -> column stores
template <int G>
B200_HD void br7_fwd12_a(const Br7Smem<G>& sm, int g, int q, int lane, const uint32_t (&dv)[32])
{
    uint32_t x1[32], x2[32];
    B200_UNROLL
    for (int k = 0; k < 8; k++) {
        fwd_start_r4_group<32 - 2 * BGBIT>(sm.r4, dv[k], dv[k + 8], dv[k + 16], dv[k + 24], x1[k], x1[k + 8], x1[k + 16], x1[k + 24]);
        fwd_start_r4_group<32 - 3 * BGBIT>(sm.r4, dv[k], dv[k + 8], dv[k + 16], dv[k + 24], x2[k], x2[k + 8], x2[k + 16], x2[k + 24]);
    }
    fwd_pass1_tail_x2(x1, x2);
    stile_store_col(sm.tile(g, q * GL + 1), x1, lane);
    stile_store_col(sm.tile(g, q * GL + 2), x2, lane);
}

// Jobs g0 .. g0+J-1 of the CTA form a barrier group of 64*J threads that runs its own pointwise stage:
// thread `tig` of the group walks positions tig, tig + 64J, ... and reuses the 36 key words of a position
// for the J jobs of the group.  J = G is the CTA-wide stage of br3_kernel; smaller groups (each with its own
// named barrier, started a fraction of a step apart) keep the phases of the groups out of step, so the
// multiply pipe sees one group's transform passes while another group sits in shared-memory exchanges.
template <int G, int J>
B200_HD void br7_pw_compute(const Br7Smem<G>& sm, int g0, int j, const uint32_t (&bkv)[BK_COLS][ROWS])
{
    const int off = stile_of_j(j);
    B200_UNROLL
    for (int gg = 0; gg < J; gg++) {
        const int g = g0 + gg;
        uint32_t d[ROWS], o[BK_COLS];
        B200_UNROLL
        for (int r = 0; r < ROWS; r++) d[r] = sm.tile(g, r)[off];
        B200_UNROLL
        for (int c = 0; c < BK_COLS; c++) {
            uint64_t acc = 0;
            B200_UNROLL
            for (int r = 0; r < ROWS; r++) acc += (uint64_t)d[r] * bkv[c][r];
            o[c] = redc64(acc);  // < 4p
        }
        B200_UNROLL
        for (int c = 0; c < BK_COLS; c++) sm.tile(g, c)[off] = o[c];
    }
}
template <int G, int J>
B200_HD void br7_pointwise(const Br7Smem<G>& sm, const uint32_t* bk_i, int g0, int tig, uint32_t (&bk0)[BK_COLS][ROWS])
{
    constexpr int T = 64 * J;
    uint32_t bk1[BK_COLS][ROWS];
    for (int j = tig; j < N1; j += 2 * T) {
        const int j1 = j + T, j2 = j + 2 * T;
        if (j1 < N1) pw_load(bk_i, j1, bk1);
        br7_pw_compute<G, J>(sm, g0, j, bk0);
        if (j1 < N1) {
            if (j2 < N1) pw_load(bk_i, j2, bk0);
            br7_pw_compute<G, J>(sm, g0, j1, bk1);
        }
    }
}

// I01a
template <int G>
B200_HD void br7_inv01_a(const Br7Smem<G>& sm, int g, int q, int lane)
{
    uint32_t x0[32], x1[32];
    uint32_t* t = sm.tile(g, q * LIMBS);
    stile_load_row(t, x0, lane);
    stile_load_row(t + STILE_WORDS, x1, lane);
    inv_pass1_x2(x0, x1, sm.tw2i, lane);
    stile_store_row(t, x0, lane);
    stile_store_row(t + STILE_WORDS, x1, lane);
}
// I01b
template <int G>
B200_HD void br7_inv01_b(const Br7Smem<G>& sm, int g, int q, int lane, uint32_t (&accr)[32])
{
    uint32_t x0[32], x1[32];
    uint32_t* t = sm.tile(g, q * LIMBS);
    stile_load_col(t, x0, lane);
    stile_load_col(t + STILE_WORDS, x1, lane);
    inv_pass2_x2(x0, x1);
    B200_UNROLL
    for (int a = 0; a < 32; a++)
        accr[a] += (uint32_t)centered_lift(x0[a]) + ((uint32_t)centered_lift(x1[a]) << LIMB_BITS);
}
// I2a
template <int G>
B200_HD void br7_inv2_a(const Br7Smem<G>& sm, int g, int q, int lane)
{
    uint32_t x2[32];
    uint32_t* t = sm.tile(g, q * LIMBS + 2);
    stile_load_row(t, x2, lane);
    inv_pass1(x2, sm.tw2i, lane);
    stile_store_row(t, x2, lane);
}
// I2b: every lane finished its column loads of tile 3q before the __syncwarp that precedes this phase,
// so the accumulator copy can be refreshed here
template <int G>
B200_HD void br7_inv2_b(const Br7Smem<G>& sm, int g, int q, int lane, uint32_t (&accr)[32])
{
    uint32_t x2[32];
    stile_load_col(sm.tile(g, q * LIMBS + 2), x2, lane);
    inv_pass2(x2);
    uint32_t* acc = sm.acc(g, q);
    B200_UNROLL
    for (int a = 0; a < 32; a++) {
        accr[a] += (uint32_t)centered_lift(x2[a]) << (2 * LIMB_BITS);
        acc[32 * a + lane] = accr[a];
    }
}

template <int G>
B200_HD void br7_epilogue(const Br7Smem<G>& sm, int g, int q, int lane, uint32_t* u_out)
{
    const uint32_t* acc = sm.acc(g, q);
    if (q == 0) {
        for (int j = lane; j < N1; j += 32) u_out[j] = (j == 0) ? acc[0] : 0u - acc[N1 - j];
    } else if (lane == 0) {
        u_out[N1] = acc[0];
    }
}

}  // namespace b200
