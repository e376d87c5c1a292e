// Lock-step CPU simulator of the CUDA kernels (test infrastructure for the "not gpu" suite).
//
// It includes the very same per-thread phase functions the sm_100a kernels are built from
// (iyokan_b200/csrc/{ntt_warp,br_phases,ks_phases}.h) and runs them thread by thread, phase by
// phase, with a plain byte array standing in for shared memory.  Because no phase function
// communicates across threads except through that array between phases, this reproduces the
// kernel's arithmetic, index math and data layout exactly; what it cannot catch are missing
// barriers, which the GPU parity tests cover.
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef B200FHE_80BIT
#include "../../iyokan_b200/csrc/br4_phases.h"
#include "../../iyokan_b200/csrc/br6_phases.h"
#include "../../iyokan_b200/csrc/br7_phases.h"
#include "../../iyokan_b200/csrc/br8_phases.h"
#endif
#include "../../iyokan_b200/csrc/br_phases.h"
#include "../../iyokan_b200/csrc/brg_phases.h"
#include "../../iyokan_b200/csrc/gate_jobs.h"
#include "../../iyokan_b200/csrc/ks_phases.h"
#include "../../iyokan_b200/csrc/br9_phases.h"

using namespace b200;

static NttTables g_tab;
#ifndef B200FHE_80BIT
static BlockTw g_btw;
static Block8Tw g_b8tw;
static Block4Tw g_b4tw;
#endif
static bool g_init = false;

// Thread-order control: every phase function is claimed to be free of intra-phase cross-thread communication, so
// the result must not depend on the order in which the simulator runs the threads (or warps) of a phase.  Mode 0 =
// ascending, 1 = descending, 2 = a fixed permutation; tests run the kernels under all three and compare.
static int g_order = 0;
extern "C" void sim_set_thread_order(int mode) { g_order = mode; }
static inline int ord(int i, int n) { return g_order == 0 ? i : g_order == 1 ? n - 1 - i : (i * 7 + 3) % n; }

extern "C" void sim_init()
{
    if (!g_init) {
        ntt_tables_init(g_tab);
#ifndef B200FHE_80BIT
        block_tw_init(g_tab, g_btw);
        block8_tw_init(g_tab, g_b8tw);
        block4_tw_init(g_tab, g_b4tw);
#endif
        g_init = true;
    }
}

// forward NTT of 1024 residues (natural order in, tile order out: position j = 32a+b at out[j])
static void warp_forward(const uint32_t* in, uint32_t* out)
{
    std::vector<uint32_t> tile(TILE_WORDS);
    for (int lane = 0; lane < 32; lane++) {
        uint32_t x[32];
        for (int a = 0; a < 32; a++) x[a] = in[32 * a + lane];
        fwd_pass1(x);
        tile_store_col(tile.data(), x, lane);
    }
    for (int lane = 0; lane < 32; lane++) {
        uint32_t x[32];
        tile_load_row(tile.data(), x, lane);
        fwd_pass2(x, g_tab.tw2f, lane);
        for (int b = 0; b < 32; b++) out[32 * lane + b] = x[b];
    }
}
static void warp_inverse(const uint32_t* in, uint32_t* out)
{
    std::vector<uint32_t> tile(TILE_WORDS);
    for (int lane = 0; lane < 32; lane++) {
        uint32_t x[32];
        for (int b = 0; b < 32; b++) x[b] = in[32 * lane + b];
        inv_pass1(x, g_tab.tw2i, lane);
        tile_store_row(tile.data(), x, lane);
    }
    for (int lane = 0; lane < 32; lane++) {
        uint32_t x[32];
        tile_load_col(tile.data(), x, lane);
        inv_pass2(x);
        for (int a = 0; a < 32; a++) out[32 * a + lane] = x[a];
    }
}

// c = a (*) b in Z_p[X]/(X^1024+1) through the warp NTT; inputs in [0,p), output in [0,p)
extern "C" void sim_negacyclic_mul_modp(const uint32_t* a, const uint32_t* b, uint32_t* c)
{
    sim_init();
    std::vector<uint32_t> fa(N1), fb(N1), fc(N1);
    warp_forward(a, fa.data());
    warp_forward(b, fb.data());
    const uint32_t ninv = mod_inv(1024);
    for (int j = 0; j < N1; j++)
        fc[j] = mod_mul(mod_mul(reduce_full(fa[j]), reduce_full(fb[j])), ninv);
    warp_inverse(fc.data(), c);
    for (int j = 0; j < N1; j++) c[j] = reduce_full(c[j]);
}

// max value seen at the output of forward / inverse transforms for worst-case-ish inputs
extern "C" void sim_ntt_ranges(const uint32_t* in, uint32_t* fwd_max, uint32_t* inv_max)
{
    sim_init();
    std::vector<uint32_t> f(N1), g(N1);
    warp_forward(in, f.data());
    uint32_t m = 0;
    for (auto v : f) m = v > m ? v : m;
    *fwd_max = m;
    warp_inverse(in, g.data());
    m = 0;
    for (auto v : g) m = v > m ? v : m;
    *inv_max = m;
}

// bk_raw [n_i][6][2][1024] -> bk_ntt [n_i][BK_COLS][ROWS][1024]
extern "C" void sim_bk_prepare(const uint32_t* bk_raw, uint32_t* bk_ntt, int n_i)
{
    sim_init();
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n_i; i++) {
        std::vector<uint32_t> tile(TILE_WORDS);
        for (int r = 0; r < ROWS; r++)
            for (int q = 0; q < 2; q++)
                for (int l = 0; l < LIMBS; l++) {
                    const uint32_t* raw = bk_raw + ((size_t)(i * ROWS + r) * 2 + q) * N1;
                    uint32_t* out = bk_ntt + ((size_t)(i * BK_COLS + q * LIMBS + l) * ROWS + r) * N1;
                    for (int lane = 0; lane < 32; lane++) brg_bk_prep_a(raw, l, lane, tile.data());
                    for (int lane = 0; lane < 32; lane++) bk_prep_b(tile.data(), g_tab.tw2f, g_tab.bk_scale, lane, out);
                }
    }
}

// ---- variant 1: generic shape (brg_kernel), both flavours ----
template <int G>
static void sim_brg_cta(const BrJob* jobs, int njobs, int cta, const torus0_t* arena, const uint32_t* bk_ntt,
                        uint32_t* ubuf, int n_iter)
{
    constexpr int T = 64 * G, W = 2 * G;
    std::vector<uint8_t> smem(BrgSmem<G>::BYTES + 16);
    BrgSmem<G> sm;
    sm.carve(smem.data());
    std::memcpy(sm.tw2f, g_tab.tw2f, sizeof(g_tab.tw2f));
    std::memcpy(sm.tw2i, g_tab.tw2i, sizeof(g_tab.tw2i));
    struct Regs {
        uint32_t accr[32], dreg[32], sum[32];
    };
    std::vector<Regs> regs(T);
    auto jobof = [&](int g) {
        int j = cta * G + g;
        return j < njobs ? j : njobs - 1;
    };
    for (int w = 0; w < W; w++)
        for (int lane = 0; lane < 32; lane++)
            brg_prologue<G>(sm, jobs[jobof(w >> 1)], arena, w >> 1, w & 1, lane, regs[w * 32 + lane].accr);
    for (int i = 0; i < n_iter; i++) {
        for (int ww = 0; ww < W; ww++) {
            const int w = ord(ww, W), g = w >> 1, q = w & 1;
            for (int l = 0; l < 32; l++) {
                Regs& r = regs[w * 32 + ord(l, 32)];
                brg_rotate_diff<G>(sm, i, g, q, ord(l, 32), r.accr, r.dreg);
            }
            for (int d = 0; d < GL; d++) {
                for (int l = 0; l < 32; l++) brg_fwd_a<G>(sm, g, q, ord(l, 32), d, regs[w * 32 + ord(l, 32)].dreg);
                for (int l = 0; l < 32; l++) brg_fwd_b<G>(sm, g, q, ord(l, 32), d);
            }
        }
        const uint32_t* bk_i = bk_ntt + (size_t)i * BK_COLS * ROWS * N1;
        for (int t = 0; t < T; t++) brg_pointwise<G>(sm, bk_i, ord(t, T));
        for (int ww = 0; ww < W; ww++) {
            const int w = ord(ww, W), g = w >> 1, q = w & 1;
            for (int lm = 0; lm < LIMBS; lm++) {
                for (int l = 0; l < 32; l++) brg_inv_a<G>(sm, g, q, ord(l, 32), lm);
                for (int l = 0; l < 32; l++) brg_inv_b<G>(sm, g, q, ord(l, 32), lm, regs[w * 32 + ord(l, 32)].sum);
            }
            for (int l = 0; l < 32; l++) {
                Regs& r = regs[w * 32 + ord(l, 32)];
                brg_acc_update<G>(sm, g, q, ord(l, 32), r.sum, r.accr);
            }
        }
    }
    for (int w = 0; w < W; w++) {
        const int g = w >> 1, q = w & 1;
        if (cta * G + g >= njobs) continue;
        for (int lane = 0; lane < 32; lane++) brg_epilogue<G>(sm, g, q, lane, ubuf + (size_t)(cta * G + g) * U_STRIDE);
    }
}

extern "C" void sim_blind_rotate1(int G, const void* jobs_raw, int njobs, const torus0_t* arena, const uint32_t* bk_ntt,
                                  uint32_t* ubuf, int n_iter)
{
    sim_init();
    const BrJob* jobs = reinterpret_cast<const BrJob*>(jobs_raw);
    const int ncta = (njobs + G - 1) / G;
#pragma omp parallel for schedule(dynamic, 1)
    for (int cta = 0; cta < ncta; cta++) {
        if (G == 2) sim_brg_cta<2>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);
        else if (G == 4) sim_brg_cta<4>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);
        else std::abort();
    }
}

#ifndef B200FHE_80BIT
// ---- variant 3: interleaved transforms (br3_kernel) ----
template <int G>
static void sim_br3_cta(const BrJob* jobs, int njobs, int cta, const torus0_t* arena, const uint32_t* bk_ntt,
                        uint32_t* ubuf, int n_iter)
{
    constexpr int T = 64 * G, W = 2 * G;
    std::vector<uint8_t> smem(BrSmem<G>::BYTES + 16);
    BrSmem<G> sm;
    sm.carve(smem.data());
    std::memcpy(sm.tw2f, g_tab.tw2f, sizeof(g_tab.tw2f));
    std::memcpy(sm.tw2i, g_tab.tw2i, sizeof(g_tab.tw2i));
    struct Regs {
        uint32_t accr[32], x0[32], x1[32], x2[32];
    };
    std::vector<Regs> regs(T);
    auto jobof = [&](int g) {
        int j = cta * G + g;
        return j < njobs ? j : njobs - 1;
    };
    for (int w = 0; w < W; w++)
        for (int lane = 0; lane < 32; lane++)
            br_prologue<G>(sm, jobs[jobof(w >> 1)], arena, w >> 1, w & 1, lane, regs[w * 32 + lane].accr);
    for (int i = 0; i < n_iter; i++) {
        for (int ww = 0; ww < W; ww++) {
            const int w = ord(ww, W), g = w >> 1, q = w & 1;
            for (int l = 0; l < 32; l++) {
                const int lane = ord(l, 32);
                Regs& r = regs[w * 32 + lane];
                br_fwd3_a<G>(sm, i, g, q, lane, r.accr, r.x0, r.x1, r.x2);
            }
            for (int l = 0; l < 32; l++) {
                const int lane = ord(l, 32);
                Regs& r = regs[w * 32 + lane];
                br_fwd3_b<G>(sm, g, q, lane, r.x0, r.x1, r.x2);
            }
            for (int l = 0; l < 32; l++) br_fwd3_c<G>(sm, g, q, ord(l, 32));
        }
        const uint32_t* bk_i = bk_ntt + (size_t)i * BK_COLS * ROWS * N1;
        for (int t = 0; t < T; t++) {
            const int tid = ord(t, T);
            uint32_t bk0[BK_COLS][ROWS];
            pw_load(bk_i, tid, bk0);
            br_pointwise<G>(sm, bk_i, tid, bk0);
        }
        for (int ww = 0; ww < W; ww++) {
            const int w = ord(ww, W), g = w >> 1, q = w & 1;
            for (int l = 0; l < 32; l++) br_inv3_a<G>(sm, g, q, ord(l, 32));
            for (int l = 0; l < 32; l++) br_inv3_b<G>(sm, g, q, ord(l, 32), regs[w * 32 + ord(l, 32)].accr);
            for (int l = 0; l < 32; l++) br_inv3_c<G>(sm, g, q, ord(l, 32), regs[w * 32 + ord(l, 32)].accr);
        }
    }
    for (int w = 0; w < W; w++) {
        const int g = w >> 1, q = w & 1;
        if (cta * G + g >= njobs) continue;
        for (int lane = 0; lane < 32; lane++)
            br_epilogue<G>(sm, g, q, lane, ubuf + (size_t)(cta * G + g) * U_STRIDE);
    }
}

extern "C" void sim_blind_rotate3(int G, const void* jobs_raw, int njobs, const torus0_t* arena,
                                  const uint32_t* bk_ntt, uint32_t* ubuf, int n_iter)
{
    sim_init();
    const BrJob* jobs = reinterpret_cast<const BrJob*>(jobs_raw);
    const int ncta = (njobs + G - 1) / G;
#pragma omp parallel for schedule(dynamic, 1)
    for (int cta = 0; cta < ncta; cta++) {
        if (G == 2) sim_br3_cta<2>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);
        else if (G == 4) sim_br3_cta<4>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);
        else if (G == 6) sim_br3_cta<6>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);
        else std::abort();
    }
}

// ---- variant 7: 16-warp throughput shape on swizzled tiles (br7_kernel) ----

template <int G, int J>
static void sim_br7_cta(const BrJob* jobs, int njobs, int cta, const torus0_t* arena, const uint32_t* bk_ntt,
                        uint32_t* ubuf, int n_iter)
{
    constexpr int T = 64 * G, W = 2 * G;
    std::vector<uint8_t> smem(Br7Smem<G>::BYTES + 16);
    Br7Smem<G> sm;
    sm.carve(smem.data());
    std::memcpy(sm.tw2f, g_tab.tw2f, sizeof(g_tab.tw2f));
    std::memcpy(sm.tw2i, g_tab.tw2i, sizeof(g_tab.tw2i));
    std::memcpy(sm.r4, g_tab.r4, sizeof(g_tab.r4));
    struct Regs {
        uint32_t accr[32], x0[32], dv[32];
    };
    std::vector<Regs> regs(T);
    auto jobof = [&](int g) {
        int j = cta * G + g;
        return j < njobs ? j : njobs - 1;
    };
    for (int w = 0; w < W; w++)
        for (int lane = 0; lane < 32; lane++)
            br7_prologue<G>(sm, jobs[jobof(w >> 1)], arena, w >> 1, w & 1, lane, regs[w * 32 + lane].accr);
    for (int i = 0; i < n_iter; i++) {
        for (int ww = 0; ww < W; ww++) {
            const int w = ord(ww, W), g = w >> 1, q = w & 1;
            for (int l = 0; l < 32; l++) {
                Regs& r = regs[w * 32 + ord(l, 32)];
                br7_fwd0_a<G>(sm, i, g, q, ord(l, 32), r.accr, r.dv, r.x0);
            }
            for (int l = 0; l < 32; l++) br7_fwd0_b<G>(sm, g, q, ord(l, 32), regs[w * 32 + ord(l, 32)].x0);
            for (int l = 0; l < 32; l++) br7_fwd0_c<G>(sm, g, q, ord(l, 32));
            for (int l = 0; l < 32; l++) br7_fwd12_a<G>(sm, g, q, ord(l, 32), regs[w * 32 + ord(l, 32)].dv);
            for (int l = 0; l < 32; l++) br7_fwd12_c<G>(sm, g, q, ord(l, 32));
        }
        const uint32_t* bk_i = bk_ntt + (size_t)i * BK_COLS * ROWS * N1;
        for (int t = 0; t < T; t++) {  // barrier groups of J jobs, each with its own pointwise stage
            const int tid = ord(t, T), g0 = (tid / (64 * J)) * J, tig = tid - 64 * g0;
            uint32_t bk0[BK_COLS][ROWS];
            pw_load(bk_i, tig, bk0);
            br7_pointwise<G, J>(sm, bk_i, g0, tig, bk0);
        }
        for (int ww = 0; ww < W; ww++) {
            const int w = ord(ww, W), g = w >> 1, q = w & 1;
            for (int l = 0; l < 32; l++) br7_inv01_a<G>(sm, g, q, ord(l, 32));
            // I01b and I2a share a phase in the kernel (no __syncwarp between them): they touch different tiles
            for (int l = 0; l < 32; l++) {
                br7_inv01_b<G>(sm, g, q, ord(l, 32), regs[w * 32 + ord(l, 32)].accr);
                br7_inv2_a<G>(sm, g, q, ord(l, 32));
            }
            for (int l = 0; l < 32; l++) br7_inv2_b<G>(sm, g, q, ord(l, 32), regs[w * 32 + ord(l, 32)].accr);
        }
    }
    for (int w = 0; w < W; w++) {
        const int g = w >> 1, q = w & 1;
        if (cta * G + g >= njobs) continue;
        for (int lane = 0; lane < 32; lane++)
            br7_epilogue<G>(sm, g, q, lane, ubuf + (size_t)(cta * G + g) * U_STRIDE);
    }
}

extern "C" void sim_blind_rotate7(int G, const void* jobs_raw, int njobs, const torus0_t* arena,
                                  const uint32_t* bk_ntt, uint32_t* ubuf, int n_iter)
{
    sim_init();
    const BrJob* jobs = reinterpret_cast<const BrJob*>(jobs_raw);
    const int Gj = G > 8 ? 8 : G;
    const int ncta = (njobs + Gj - 1) / Gj;
#pragma omp parallel for schedule(dynamic, 1)
    for (int cta = 0; cta < ncta; cta++) {
        if (G == 2) sim_br7_cta<2, 2>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);
        else if (G == 8) sim_br7_cta<8, 8>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);
        else if (G == 84) sim_br7_cta<8, 4>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);  // 8 jobs, groups of 4
        else if (G == 82) sim_br7_cta<8, 2>(jobs, njobs, cta, arena, bk_ntt, ubuf, n_iter);  // 8 jobs, groups of 2
        else std::abort();
    }
}

// ---- variant 4: one job per CTA, 6 teams of 64 threads (br4_kernel) ----
// team transform on its own: natural order in, NTT positions out (must equal the warp transform bit for bit)
extern "C" void sim_block_forward(const uint32_t* in, uint32_t* out)
{
    sim_init();
    std::vector<uint32_t> tile(BT_WORDS);
    for (int t = 0; t < TEAM_THREADS; t++) {
        uint32_t x[16];
        for (int a = 0; a < 16; a++) x[a] = in[64 * a + t];
        blk_fwd_p1(x);
        blk_store_p1(tile.data(), x, t);
    }
    for (int t = 0; t < TEAM_THREADS; t++) {
        tw_t w[15];
        blk_load_tw2(g_btw.p2f, t, w);
        blk_fwd_p2(tile.data(), w, t);
    }
    for (int t = 0; t < TEAM_THREADS; t++) blk_fwd_p3(tile.data(), g_btw.p3f, t);
    for (int j = 0; j < N1; j++) out[j] = tile[bt_pad(j)];
}
extern "C" void sim_block_inverse(const uint32_t* in, uint32_t* out)
{
    sim_init();
    std::vector<uint32_t> tile(BT_WORDS);
    for (int j = 0; j < N1; j++) tile[bt_pad(j)] = in[j];
    for (int t = 0; t < TEAM_THREADS; t++) blk_inv_pA(tile.data(), g_btw.p3i, t);
    for (int t = 0; t < TEAM_THREADS; t++) {
        tw_t w[15];
        blk_load_tw2(g_btw.p2i, t, w);
        blk_inv_pB(tile.data(), w, t);
    }
    for (int t = 0; t < TEAM_THREADS; t++) {
        uint32_t x[16];
        blk_load_p1(tile.data(), x, t);
        blk_inv_pC(x);
        for (int a = 0; a < 16; a++) out[64 * a + t] = x[a];
    }
}
extern "C" void sim_warp_forward(const uint32_t* in, uint32_t* out)
{
    sim_init();
    warp_forward(in, out);
}
extern "C" void sim_warp_inverse(const uint32_t* in, uint32_t* out)
{
    sim_init();
    warp_inverse(in, out);
}

static void sim_br4_cta(const BrJob* jobs, int job, const torus0_t* arena, const uint32_t* bk_ntt, uint32_t* ubuf,
                        int n_iter)
{
    std::vector<uint8_t> smem(Br4Smem::BYTES + 128);
    Br4Smem sm;
    sm.carve(reinterpret_cast<void*>(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127));
    std::memcpy(sm.tw, &g_btw, sizeof(BlockTw));
    for (int tid = 0; tid < BR4_THREADS; tid++) br4_prologue(sm, jobs[job], arena, tid);
    auto each = [&](auto fn) {
        for (int k = 0; k < BR4_THREADS; k++) {
            const int tid = ord(k, BR4_THREADS), team = tid >> 6;
            fn(team / GL, team % GL, tid & 63);
        }
    };
    for (int i = 0; i < n_iter; i++) {
        std::memcpy(sm.keyb, bk_ntt + (size_t)i * BR4_KEY_WORDS, (size_t)BR4_KEY_WORDS * 4);  // the bulk-async copy
        each([&](int q, int d, int t) { br4_fwd_p1(sm, i, q, d, t); });
        each([&](int q, int d, int t) { br4_fwd_p2(sm, q, d, t); });
        each([&](int q, int d, int t) { br4_fwd_p3(sm, q, d, t); });
        for (int k = 0; k < BR4_THREADS; k++) {
            const int tid = ord(k, BR4_THREADS);
            br4_pointwise_item(sm, tid);
            br4_pointwise_item(sm, tid + BR4_THREADS);
        }
        each([&](int q, int d, int t) { br4_inv_pA(sm, q, d, t); });
        each([&](int q, int d, int t) { br4_inv_pB(sm, q, d, t); });
        each([&](int q, int d, int t) { br4_inv_pC(sm, q, d, t); });
    }
    for (int tid = 0; tid < BR4_THREADS; tid++) br4_epilogue(sm, tid, ubuf + (size_t)job * U_STRIDE);
}

extern "C" void sim_blind_rotate4(const void* jobs_raw, int njobs, const torus0_t* arena, const uint32_t* bk_ntt,
                                  uint32_t* ubuf, int n_iter)
{
    sim_init();
    const BrJob* jobs = reinterpret_cast<const BrJob*>(jobs_raw);
#pragma omp parallel for schedule(dynamic, 1)
    for (int job = 0; job < njobs; job++) sim_br4_cta(jobs, job, arena, bk_ntt, ubuf, n_iter);
}


// ---- variant 6: cluster shape with 128-thread x 8-point teams (br6_kernel) ----
extern "C" void sim_block8_forward(const uint32_t* in, uint32_t* out)
{
    sim_init();
    std::vector<uint32_t> tile(B8_WORDS);
    for (int t = 0; t < TEAM8_THREADS; t++) {
        uint32_t x[8];
        for (int a = 0; a < 8; a++) x[a] = in[128 * a + t];
        blk8_fwd_p1(x);
        blk8_store_p1(tile.data(), x, t);
    }
    for (int t = 0; t < TEAM8_THREADS; t++) blk8_fwd_p2(tile.data(), g_b8tw.q2f, t);
    for (int t = 0; t < TEAM8_THREADS; t++) blk8_fwd_p3(tile.data(), g_b8tw.q3f, t);
    for (int t = 0; t < TEAM8_THREADS; t++) blk8_fwd_p4(tile.data(), g_b8tw.q4f, t);
    for (int j = 0; j < N1; j++) out[j] = tile[b8_pad(j)];
}
extern "C" void sim_block8_inverse(const uint32_t* in, uint32_t* out)
{
    sim_init();
    std::vector<uint32_t> tile(B8_WORDS);
    for (int j = 0; j < N1; j++) tile[b8_pad(j)] = in[j];
    for (int t = 0; t < TEAM8_THREADS; t++) blk8_inv_pA(tile.data(), g_b8tw.q4i, t);
    for (int t = 0; t < TEAM8_THREADS; t++) blk8_inv_pB(tile.data(), g_b8tw.q3i, t);
    for (int t = 0; t < TEAM8_THREADS; t++) blk8_inv_pC(tile.data(), g_b8tw.q2i, t);
    for (int t = 0; t < TEAM8_THREADS; t++) {
        uint32_t x[8];
        blk8_load_p1(tile.data(), x, t);
        blk8_inv_pD(x);
        for (int a = 0; a < 8; a++) out[128 * a + t] = x[a];
    }
}

static void sim_br6_cluster(const BrJob* jobs, int job, const torus0_t* arena, const uint32_t* bk_ntt, uint32_t* ubuf,
                            int n_iter)
{
    std::vector<uint8_t> smem[2] = {std::vector<uint8_t>(Br6Smem::BYTES + 128), std::vector<uint8_t>(Br6Smem::BYTES + 128)};
    Br6Smem sm[2];
    for (int q = 0; q < 2; q++) {
        sm[q].carve(reinterpret_cast<void*>(((uintptr_t)smem[q].data() + 127) & ~(uintptr_t)127));
        std::memcpy(sm[q].tw, &g_b8tw, sizeof(Block8Tw));
        for (int tid = 0; tid < BR6_THREADS; tid++) br6_prologue(sm[q], jobs[job], arena, q, tid);
    }
    auto each = [&](auto fn) {
        for (int qq = 0; qq < 2; qq++)
            for (int k = 0; k < BR6_THREADS; k++) {
                const int q = ord(qq, 2), tid = ord(k, BR6_THREADS);
                fn(q, tid >> 7, tid & 127);
            }
    };
    std::vector<uint64_t> pacc((size_t)2 * BR6_THREADS * LIMBS * 4);
    auto acc_of = [&](int q, int tid) -> uint64_t(&)[LIMBS][4] {
        return *reinterpret_cast<uint64_t(*)[LIMBS][4]>(pacc.data() + ((size_t)q * BR6_THREADS + tid) * LIMBS * 4);
    };
    for (int i = 0; i < n_iter; i++) {
        for (int q = 0; q < 2; q++)
            std::memcpy(sm[q].keyb, bk_ntt + (size_t)i * BR4_KEY_WORDS + (size_t)q * BR6_KEY_WORDS, (size_t)BR6_KEY_WORDS * 4);
        each([&](int q, int d, int t) { br6_fwd_p1(sm[q], i, q, d, t); });
        each([&](int q, int d, int t) { br6_fwd_p2(sm[q], q, d, t); });
        each([&](int q, int d, int t) { br6_fwd_p3(sm[q], q, d, t); });
        each([&](int q, int d, int t) { br6_fwd_p4(sm[q], q, d, t); });
        for (int q = 0; q < 2; q++)
            for (int d = 0; d < GL; d++)
                std::memcpy(sm[q ^ 1].in_tile(q * GL + d), sm[q].in_tile(q * GL + d), (size_t)B8_WORDS * 4);
        for (int q = 0; q < 2; q++)
            for (int k = 0; k < BR6_THREADS; k++) br6_pw_local(sm[q], q, ord(k, BR6_THREADS), acc_of(q, ord(k, BR6_THREADS)));
        for (int q = 0; q < 2; q++)
            for (int k = 0; k < BR6_THREADS; k++) br6_pw_finish(sm[q], q, ord(k, BR6_THREADS), acc_of(q, ord(k, BR6_THREADS)));
        each([&](int q, int d, int t) { br6_inv_pA(sm[q], d, t); });
        each([&](int q, int d, int t) { br6_inv_pB(sm[q], d, t); });
        each([&](int q, int d, int t) { br6_inv_pC(sm[q], d, t); });
        each([&](int q, int d, int t) { br6_inv_pD(sm[q], d, t); });
    }
    for (int q = 0; q < 2; q++)
        for (int tid = 0; tid < BR6_THREADS; tid++) br6_epilogue(sm[q], q, tid, ubuf + (size_t)job * U_STRIDE);
}

extern "C" void sim_blind_rotate6(const void* jobs_raw, int njobs, const torus0_t* arena, const uint32_t* bk_ntt,
                                  uint32_t* ubuf, int n_iter)
{
    sim_init();
    const BrJob* jobs = reinterpret_cast<const BrJob*>(jobs_raw);
#pragma omp parallel for schedule(dynamic, 1)
    for (int job = 0; job < njobs; job++) sim_br6_cluster(jobs, job, arena, bk_ntt, ubuf, n_iter);
}



// ---- variant 9: cluster shape with 4-point threads (br9_kernel): br6's protocol, five two-stage passes of 256 threads ----
static void sim_br9_cluster(const BrJob* jobs, int job, const torus0_t* arena, const uint32_t* bk_ntt, uint32_t* ubuf,
                            int n_iter)
{
    std::vector<uint8_t> smem[2] = {std::vector<uint8_t>(Br9Smem::BYTES + 128), std::vector<uint8_t>(Br9Smem::BYTES + 128)};
    Br9Smem sm[2];
    for (int q = 0; q < 2; q++) {
        sm[q].carve(reinterpret_cast<void*>(((uintptr_t)smem[q].data() + 127) & ~(uintptr_t)127));
        std::memcpy(sm[q].tw, &g_b4tw, sizeof(Block4Tw));
        for (int tid = 0; tid < BR9_THREADS; tid++) br9_prologue(sm[q], jobs[job], arena, q, tid);
    }
    auto each = [&](auto fn) {
        for (int qq = 0; qq < 2; qq++)
            for (int k = 0; k < BR9_THREADS; k++) {
                const int q = ord(qq, 2), tid = ord(k, BR9_THREADS);
                fn(q, tid >> 8, tid & 255);
            }
    };
    std::vector<uint64_t> pacc((size_t)2 * BR9_THREADS * LIMBS * 4);
    auto acc_of = [&](int q, int tid) -> uint64_t(&)[LIMBS][4] {
        return *reinterpret_cast<uint64_t(*)[LIMBS][4]>(pacc.data() + ((size_t)q * BR9_THREADS + tid) * LIMBS * 4);
    };
    for (int i = 0; i < n_iter; i++) {
        for (int q = 0; q < 2; q++)
            std::memcpy(sm[q].keyb, bk_ntt + (size_t)i * BR4_KEY_WORDS + (size_t)q * BR6_KEY_WORDS, (size_t)BR6_KEY_WORDS * 4);
        each([&](int q, int d, int t) { br9_fwd_p1(sm[q], i, q, d, t); });
        each([&](int q, int d, int t) { br9_fwd_p2(sm[q], q, d, t); });
        each([&](int q, int d, int t) { br9_fwd_p3(sm[q], q, d, t); });
        each([&](int q, int d, int t) { br9_fwd_p4(sm[q], q, d, t); });
        each([&](int q, int d, int t) { br9_fwd_p5(sm[q], q, d, t); });
        for (int q = 0; q < 2; q++)
            for (int d = 0; d < GL; d++)
                std::memcpy(sm[q ^ 1].in_tile(q * GL + d), sm[q].in_tile(q * GL + d), (size_t)B8_WORDS * 4);
        for (int q = 0; q < 2; q++)
            for (int k = 0; k < BR9_THREADS; k++) br6_pw_local(sm[q], q, ord(k, BR9_THREADS), acc_of(q, ord(k, BR9_THREADS)));
        for (int q = 0; q < 2; q++)
            for (int k = 0; k < BR9_THREADS; k++) br6_pw_finish(sm[q], q, ord(k, BR9_THREADS), acc_of(q, ord(k, BR9_THREADS)));
        each([&](int q, int d, int t) { br9_inv_pA(sm[q], d, t); });
        each([&](int q, int d, int t) { br9_inv_pB(sm[q], d, t); });
        each([&](int q, int d, int t) { br9_inv_pC(sm[q], d, t); });
        each([&](int q, int d, int t) { br9_inv_pD(sm[q], d, t); });
        each([&](int q, int d, int t) { br9_inv_pE(sm[q], d, t); });
    }
    for (int q = 0; q < 2; q++)
        for (int tid = 0; tid < BR9_THREADS; tid++) br9_epilogue(sm[q], q, tid, ubuf + (size_t)job * U_STRIDE);
}

extern "C" void sim_blind_rotate9(const void* jobs_raw, int njobs, const torus0_t* arena, const uint32_t* bk_ntt,
                                  uint32_t* ubuf, int n_iter)
{
    sim_init();
    const BrJob* jobs = reinterpret_cast<const BrJob*>(jobs_raw);
#pragma omp parallel for schedule(dynamic, 1)
    for (int job = 0; job < njobs; job++) sim_br9_cluster(jobs, job, arena, bk_ntt, ubuf, n_iter);
}

// ---- variant 8: quad-cluster shape (br8_kernel); the four CTAs advance phase by phase, the three DSMEM exchanges
// become memcpys at the points where the kernel issues the bulk copies ----
static void sim_br8_cluster(const BrJob* jobs, int job, const torus0_t* arena, const uint32_t* bk_ntt, uint32_t* ubuf,
                            int n_iter)
{
    std::vector<uint8_t> smem[4];
    Br8Smem sm[4];
    for (int r = 0; r < 4; r++) {
        smem[r].assign(Br8Smem::BYTES + 128, 0);
        sm[r].carve(reinterpret_cast<void*>(((uintptr_t)smem[r].data() + 127) & ~(uintptr_t)127));
        std::memcpy(sm[r].twf, g_tab.fwd, sizeof(g_tab.fwd));
        std::memcpy(sm[r].twi, g_tab.inv, sizeof(g_tab.inv));
        for (int tid = 0; tid < BR8_THREADS; tid++) br8_prologue(sm[r], jobs[job], arena, r >> 1, tid);
    }
    auto each = [&](auto fn) {  // fn(rank, q, h, d, t)
        for (int rr = 0; rr < 4; rr++)
            for (int k = 0; k < BR8_THREADS; k++) {
                const int r = ord(rr, 4), tid = ord(k, BR8_THREADS);
                fn(r, r >> 1, r & 1, tid >> 6, tid & 63);
            }
    };
    for (int i = 0; i < n_iter; i++) {
        const int par = i & 1;
        for (int r = 0; r < 4; r++) {  // key quarter of CTA (q, h)
            const int q = r >> 1, h = r & 1;
            for (int c = 0; c < LIMBS * ROWS; c++)
                std::memcpy(sm[r].keyb + c * 512, bk_ntt + (size_t)i * BR4_KEY_WORDS + (size_t)(q * LIMBS * ROWS + c) * N1 + 512 * h, 2048);
        }
        each([&](int r, int, int h, int d, int t) { br8_fwd_p1(sm[r], i, h, d, t); });
        each([&](int r, int, int h, int d, int t) { br8_fwd_p2(sm[r].dig + (size_t)(par * GL + d) * H_WORDS, sm[r].twf, h, t); });
        each([&](int r, int, int h, int d, int t) { br8_fwd_p3(sm[r].dig + (size_t)(par * GL + d) * H_WORDS, sm[r].twf, h, t); });
        for (int r = 0; r < 4; r++)  // exchange A: digit half tiles -> CTA r ^ 2
            std::memcpy(sm[r ^ 2].peer + (size_t)par * GL * H_WORDS, sm[r].dig + (size_t)par * GL * H_WORDS, (size_t)GL * H_WORDS * 4);
        for (int rr = 0; rr < 4; rr++)
            for (int k = 0; k < BR8_THREADS; k++) br8_pointwise(sm[ord(rr, 4)], ord(rr, 4) >> 1, par, ord(k, BR8_THREADS));
        each([&](int r, int, int h, int d, int t) { br8_inv_pA(sm[r].outb + (size_t)d * H_WORDS, sm[r].twi, h, t); });
        each([&](int r, int, int h, int d, int t) { br8_inv_pB(sm[r].outb + (size_t)d * H_WORDS, sm[r].twi, h, t); });
        each([&](int r, int, int h, int d, int t) { br8_inv_pC(sm[r].outb + (size_t)d * H_WORDS, h, t); });
        for (int r = 0; r < 4; r++)  // exchange B: half results -> CTA r ^ 1
            std::memcpy(sm[r ^ 1].half2, sm[r].outb, (size_t)LIMBS * H_WORDS * 4);
        each([&](int r, int, int h, int d, int t) { br8_inv_join(sm[r], h, d, t); });
        for (int r = 0; r < 4; r++)  // exchange C: updated accumulator half -> CTA r ^ 1
            std::memcpy(sm[r ^ 1].accb + 512 * (r & 1), sm[r].accb + 512 * (r & 1), 2048);
    }
    for (int r = 0; r < 4; r++)
        for (int tid = 0; tid < BR8_THREADS; tid++) br8_epilogue(sm[r], r >> 1, r & 1, tid, ubuf + (size_t)job * U_STRIDE);
}

extern "C" void sim_blind_rotate8(const void* jobs_raw, int njobs, const torus0_t* arena, const uint32_t* bk_ntt,
                                  uint32_t* ubuf, int n_iter)
{
    sim_init();
    const BrJob* jobs = reinterpret_cast<const BrJob*>(jobs_raw);
#pragma omp parallel for schedule(dynamic, 1)
    for (int job = 0; job < njobs; job++) sim_br8_cluster(jobs, job, arena, bk_ntt, ubuf, n_iter);
}

#endif  // !B200FHE_80BIT

// ksk_dev: uint16 [1024][7][3][640]; jobs: packed KsJob (16 bytes each)
extern "C" void sim_keyswitch(const void* jobs_raw, int njobs, const uint32_t* ubuf, const torus0_t* ksk_dev,
                              torus0_t* arena)
{
    const KsJob* jobs = reinterpret_cast<const KsJob*>(jobs_raw);
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < njobs; n++) {
        uint16_t codes[N1];
        for (int i = 0; i < N1; i++) codes[i] = ks_code(ubuf, jobs[n], i);
        const uint32_t b = ks_b_rounded(ubuf, jobs[n]);
        uint32_t* out = reinterpret_cast<uint32_t*>(arena + (size_t)jobs[n].out * SLOT_STRIDE);
        for (int k = 0; k < KS_THREADS; k++) {  // grouped accumulation exactly as ks_kernel does it
            uint32_t lo = 0, hi = 0;
            for (int y = 0; y < KS_GROUPS; y++) {
                uint32_t l, h;
                ks_accumulate_group(reinterpret_cast<const uint32_t*>(ksk_dev), codes, k, y, KS_GROUPS, l, h);
                lo += l;
                hi += h;
            }
            out[k] = ks_finish(lo, hi, b, jobs[n].post, k);
        }
    }
}

// wide-frontier path: eight gates per CTA, one warp per gate (ks8_kernel)
extern "C" void sim_keyswitch8(const void* jobs_raw, int njobs, const uint32_t* ubuf, const torus0_t* ksk_dev,
                               torus0_t* arena)
{
    const KsJob* jobs = reinterpret_cast<const KsJob*>(jobs_raw);
    const int ncta = (njobs + KS8_GATES - 1) / KS8_GATES;
#pragma omp parallel for schedule(dynamic, 1)
    for (int cta = 0; cta < ncta; cta++) {
        std::vector<uint16_t> codes((size_t)KS8_GATES * N1);
        struct Regs {
            uint32_t lo[2 * KS8_PAIRS], hi[2 * KS8_PAIRS];
        };
        std::vector<Regs> regs(32 * KS8_GATES);
        std::memset(regs.data(), 0, regs.size() * sizeof(Regs));
        auto jobof = [&](int w) {
            const int g = cta * KS8_GATES + w;
            return jobs[g < njobs ? g : njobs - 1];
        };
        for (int w = 0; w < KS8_GATES; w++)
            for (int i = 0; i < N1; i++) codes[(size_t)w * N1 + i] = ks_code(ubuf, jobof(w), i);
        for (int i0 = 0; i0 < N1; i0 += KS8_SYNC)
            for (int ww = 0; ww < KS8_GATES; ww++) {
                const int w = ord(ww, KS8_GATES);
                for (int l = 0; l < 32; l++) {
                    Regs& r = regs[w * 32 + ord(l, 32)];
                    ks8_accumulate(reinterpret_cast<const uint32_t*>(ksk_dev), codes.data() + (size_t)w * N1, i0, i0 + KS8_SYNC,
                                   ord(l, 32), r.lo, r.hi);
                }
            }
        for (int w = 0; w < KS8_GATES; w++) {
            if (cta * KS8_GATES + w >= njobs) continue;
            const KsJob job = jobof(w);
            for (int l = 0; l < 32; l++) {
                Regs& r = regs[w * 32 + l];
                ks8_store(reinterpret_cast<uint32_t*>(arena + (size_t)job.out * SLOT_STRIDE), r.lo, r.hi, ks_b_rounded(ubuf, job),
                          job.post, l);
            }
        }
    }
}

// narrow-frontier path: KS_SPLIT CTAs per switch (ks_split_kernel) + ks_combine_kernel
extern "C" void sim_keyswitch_split(const void* jobs_raw, int njobs, const uint32_t* ubuf, const torus0_t* ksk_dev,
                                    torus0_t* arena)
{
    const KsJob* jobs = reinterpret_cast<const KsJob*>(jobs_raw);
    constexpr int SPAN = N1 / KS_SPLIT;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < njobs; n++) {
        std::vector<uint32_t> partial((size_t)KS_SPLIT * 2 * KS_THREADS);
        for (int piece = 0; piece < KS_SPLIT; piece++) {  // one CTA each
            uint16_t codes[SPAN];
            const int i0 = piece * SPAN;
            for (int i = 0; i < SPAN; i++) codes[i] = ks_code(ubuf, jobs[n], i0 + i);
            for (int k = 0; k < KS_THREADS; k++) {
                uint32_t lo = 0, hi = 0;
                for (int y = 0; y < KS_GROUPS; y++) {
                    uint32_t l, h;
                    ks_accumulate_range(reinterpret_cast<const uint32_t*>(ksk_dev), codes, k, y, KS_GROUPS, i0, i0 + SPAN, l, h);
                    lo += l;
                    hi += h;
                }
                partial[(size_t)piece * 2 * KS_THREADS + k] = lo;
                partial[(size_t)piece * 2 * KS_THREADS + KS_THREADS + k] = hi;
            }
        }
        uint32_t* out = reinterpret_cast<uint32_t*>(arena + (size_t)jobs[n].out * SLOT_STRIDE);
        for (int k = 0; k < KS_THREADS; k++) {  // combine kernel
            uint32_t lo = 0, hi = 0;
            for (int piece = 0; piece < KS_SPLIT; piece++) {
                lo += partial[(size_t)piece * 2 * KS_THREADS + k];
                hi += partial[(size_t)piece * 2 * KS_THREADS + KS_THREADS + k];
            }
            out[k] = ks_finish(lo, hi, ks_b_rounded(ubuf, jobs[n]), jobs[n].post, k);
        }
    }
}

extern "C" void sim_unary(const void* jobs_raw, int njobs, torus0_t* arena)
{
    const UnaryJob* jobs = reinterpret_cast<const UnaryJob*>(jobs_raw);
    for (int n = 0; n < njobs; n++) {
        uint32_t tmp[KS_THREADS];
        for (int k = 0; k < KS_THREADS; k++) tmp[k] = unary_word(jobs[n], reinterpret_cast<const uint32_t*>(arena), k);
        std::memcpy(arena + (size_t)jobs[n].dst * SLOT_STRIDE, tmp, sizeof(tmp));
    }
}

// Whole gate frontier exactly as b200fhe_gate_batch runs it: job building, unary gather/scatter,
// blind rotations, key switches.  Returns 0, or -1 with *err set.
extern "C" int sim_gate_batch(int G, const uint8_t* opcode, const uint32_t* in0, const uint32_t* in1,
                              const uint32_t* in2, const uint32_t* out, size_t n, torus0_t* arena, size_t n_slots,
                              const uint32_t* bk_ntt, const torus0_t* ksk_dev, const char** err)
{
    sim_init();
    std::vector<BrJob> br(2 * n + 1);
    std::vector<KsJob> ks(n + 1);
    std::vector<UnaryJob> un(n + 1);
    BatchCounts cnt;
    if (const char* e = build_gate_jobs(opcode, in0, in1, in2, out, n, n_slots, br.data(), ks.data(), un.data(), cnt)) {
        if (err) *err = e;
        return -1;
    }
    if (cnt.nun) {  // gather then scatter, like the two unary kernels
        std::vector<uint32_t> stage(cnt.nun * KS_THREADS);
        for (size_t j = 0; j < cnt.nun; j++)
            for (int k = 0; k < KS_THREADS; k++)
                stage[j * KS_THREADS + k] = unary_word(un[j], reinterpret_cast<const uint32_t*>(arena), k);
        for (size_t j = 0; j < cnt.nun; j++)
            std::memcpy(arena + (size_t)un[j].dst * SLOT_STRIDE, &stage[j * KS_THREADS], KS_THREADS * 4);
    }
    if (cnt.nbr) {
        std::vector<uint32_t> ubuf(cnt.nbr * (size_t)U_STRIDE);
#ifndef B200FHE_80BIT
        if (G == 8) sim_blind_rotate7(G, br.data(), (int)cnt.nbr, arena, bk_ntt, ubuf.data(), N0);
        else if (G > 0) sim_blind_rotate3(G, br.data(), (int)cnt.nbr, arena, bk_ntt, ubuf.data(), N0);
        else
#endif
            sim_blind_rotate1(G < 0 ? -G : 4, br.data(), (int)cnt.nbr, arena, bk_ntt, ubuf.data(), N0);
        sim_keyswitch(ks.data(), (int)cnt.nks, ubuf.data(), ksk_dev, arena);
    }
    return 0;
}

extern "C" int sim_sizeof_brjob() { return (int)sizeof(BrJob); }
extern "C" int sim_sizeof_ksjob() { return (int)sizeof(KsJob); }
extern "C" uint32_t sim_prime() { return P; }
extern "C" int sim_flavour_bits() { return T0_BITS == 32 ? 80 : 128; }
extern "C" int sim_n0() { return N0; }
