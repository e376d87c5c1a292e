/* b200fhe — C ABI of the B200-native TFHE gate-evaluation back-end for Iyokan.
 *
 * This is the drop-in boundary for Iyokan's gate workers: a host (Iyokan's Task/Worker
 * scheduler, or iyokan_b200's own levelised netlist engine) keeps every TLWE ciphertext in a
 * device-resident slot arena and submits the scheduler's ready-gate frontier as ONE batch.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   b200fhe_create / b200fhe_destroy      cufhe::SetGPUNum + Initialize/CleanUp bookkeeping
 *                                         (thirdparty/cuFHE/include/cufhe_gpu.cuh:53-73)
 *   b200fhe_load_keys                     cufhe::Initialize(ek): BootstrappingKeyToNTT +
 *                                         KeySwitchingKeyToDevice
 *                                         (thirdparty/cuFHE/src/cufhe_gates_gpu.cu:42-47,
 *                                          src/bootstrap_gpu.cu:89-115, src/keyswitch_gpu.cu:9-15);
 *                                         inputs are EvalKey::bklvl01 / iksklvl10 as laid out in memory
 *                                         (TFHEpp include/params.hpp:102-128, cloudkey.hpp:333-357)
 *   b200fhe_arena_alloc/upload/download   cufhe::Ctxt<P> host/device pairs + CtxtCopyH2D/D2H per gate
 *                                         (cufhe_gpu.cuh:112-131, cufhe_gates_gpu.cu:145-157)
 *   b200fhe_gate_batch                    TaskTFHEppGate{AND..MUX,NOT,CONST*}::startSync ->
 *                                         TFHEpp::Hom* (src/iyokan_tfhepp.hpp:109-144) and
 *                                         TaskCUFHEGate*::startAsyncImpl -> cufhe::And/.../Mux
 *                                         (src/iyokan_cufhe.hpp:207-262), one call per FRONTIER
 *   b200fhe_dff_tick                      TaskDFF::tick, output <- input(0) for every DFF/RAM cell
 *                                         (src/iyokan.hpp:1395-1402, driven by NetworkRunner::tick :2050-2054)
 *   b200fhe_query / b200fhe_sync          cufhe::StreamQuery / Synchronize (cufhe_gpu.cuh:60,201)
 *   b200fhe_last_error                    replaces CuSafeCall's exit(-1) (include/details/error_gpu.cuh:31-66):
 *                                         nothing in this library exits; the caller maps non-zero to error::die.
 *
 * Conventions: every function returns 0 on success, non-zero on failure (see b200fhe_last_error).
 * All pointers are plain host pointers unless the name says "dev".  The library is not
 * thread-safe per context; Iyokan calls it from its single scheduler thread
 * (src/iyokan_tfhepp.cpp:28-47).  One context = one GPU = (in multi-GPU runs) one process.
 *
 * Environment read by the library (all optional; none changes results, only how the work is launched):
 *   B200FHE_NO_CALIBRATE=1   keep the compiled-in launch-plan table instead of timing one wave per shape at key load
 *   B200FHE_NO_GRAPH=1       replay programs launch by launch instead of as one CUDA graph
 *   B200FHE_KS8_MIN=n        frontier width from which the key switch takes eight gates per CTA (default 1400)
 *   B200FHE_NCCL_LIB=path    NCCL library to dlopen for b200fhe_comm_* (default libnccl.so.2)
 *   B200FHE_BR7_GROUP / B200FHE_BR7_SKEW / B200FHE_L2_PERSIST   experiment knobs kept for the measurements under profiles/
 */
#ifndef B200FHE_H
#define B200FHE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Parameter set: a compile-time flavour, as in the reference (-DIYOKAN_80BIT_SECURITY=On, CMakeLists.txt:28-30).
 * Callers of libb200fhe80.so define B200FHE_80BIT before including this header. */
#ifdef B200FHE_80BIT
/* 80-bit parameter set (TFHEpp include/params/CGGI16.hpp): n = 500, 32-bit lvl0 torus, l = 2, Bg = 2^10, t = 8 */
typedef uint32_t b200fhe_torus0; /* lvl0param::T */
#define B200FHE_N0 500           /* lvl0 dimension; a TLWE lvl0 is 501 uint32 = 2004 bytes */
#define B200FHE_TLWE0_LEN 501
#define B200FHE_BK_WORDS (500ull * 4 * 2 * 1024)       /* uint32 [n][(k+1)l][k+1][N]  raw TRGSW */
#define B200FHE_KSK_ELEMS (1024ull * 8 * 3 * 501)      /* uint32 [N][t][3][n+1] */
#else
/* 128-bit parameter set (TFHEpp include/params/128bit.hpp) */
typedef uint16_t b200fhe_torus0;
#define B200FHE_N0 636           /* lvl0 dimension; a TLWE lvl0 is 637 uint16 = 1274 bytes */
#define B200FHE_TLWE0_LEN 637
#define B200FHE_BK_WORDS (636ull * 6 * 2 * 1024)       /* uint32 [n][(k+1)l][k+1][N]  raw TRGSW */
#define B200FHE_KSK_ELEMS (1024ull * 7 * 3 * 637)      /* uint16 [N][t][3][n+1] */
#define B200FHE_KSK_HALFS B200FHE_KSK_ELEMS
#endif
#define B200FHE_N1 1024          /* lvl1 ring degree */
#define B200FHE_TLWE1_LEN 1025

/* Gate opcodes.  ANDNOT = HomANDYN, ORNOT = HomORYN (src/iyokan_tfhepp.hpp:133,136).
 * MUX computes in2 ? in1 : in0 (HomMUX(out, in(2), in(1), in(0)), iyokan_tfhepp.hpp:140). */
enum b200fhe_op {
    B200FHE_AND = 0, B200FHE_NAND = 1, B200FHE_ANDNOT = 2, B200FHE_OR = 3, B200FHE_NOR = 4,
    B200FHE_ORNOT = 5, B200FHE_XOR = 6, B200FHE_XNOR = 7, B200FHE_MUX = 8, B200FHE_NOT = 9,
    B200FHE_COPY = 10, B200FHE_CONST0 = 11, B200FHE_CONST1 = 12, B200FHE_ANDNY = 13, B200FHE_ORNY = 14,
    B200FHE_NUM_OPS = 15
};

typedef struct b200fhe_ctx b200fhe_ctx;

/* lifetime ------------------------------------------------------------------------------- */
int b200fhe_create(b200fhe_ctx **out, int device);
void b200fhe_destroy(b200fhe_ctx *ctx);
const char *b200fhe_last_error(void);
/* tuning knob: rotation jobs per CTA (2, 4, 6 for variant 3; 8 for variant 7); 0 = default: chosen per batch
 * size (and kernel variant with it).  Pinning either knob switches the heuristic off. */
int b200fhe_set_jobs_per_cta(b200fhe_ctx *ctx, int g);
/* tuning knob: blind-rotation kernel shape; 0 = back to the launch plan.
 * 7 = eight jobs / 16 warps per CTA on swizzled tiles (the throughput shape the plan prefers),
 * 3 = G jobs / 2G warps per CTA, the three transforms of a warp interleaved (12 warps at G = 6),
 * 4 = one job per CTA, 12 warps (6 teams of 64 threads), key staged into shared memory by bulk-async
 *     copies: low latency for one dependency level (jobs-per-CTA is ignored),
 * 6 = one job per 2-CTA thread-block cluster (one accumulator polynomial per SM, digit tiles exchanged by
 *     bulk-async copies through distributed shared memory), 128-thread x 8-point teams: lowest latency, <= 74 jobs */
int b200fhe_set_kernel_variant(b200fhe_ctx *ctx, int variant);

/* keys: raw bootstrapping key + key-switching key in the reference's memory layout.
 * Copies to the device, converts the bootstrapping key to NTT form there. */
int b200fhe_load_keys(b200fhe_ctx *ctx, const uint32_t *bk_raw, const b200fhe_torus0 *ksk);

/* ciphertext arena ------------------------------------------------------------------------ */
int b200fhe_arena_alloc(b200fhe_ctx *ctx, size_t n_slots);
/* use caller-owned device memory (n_slots * 1280 bytes [2048 at 80 bits], zero-initialised) as the arena */
int b200fhe_arena_attach(b200fhe_ctx *ctx, void *dev_ptr, size_t n_slots);
size_t b200fhe_arena_slots(const b200fhe_ctx *ctx);
void *b200fhe_arena_dev_ptr(const b200fhe_ctx *ctx);
/* tlwe_host is [n][TLWE0_LEN] lvl0 torus words, densely packed (the reference's std::array<T, n+1>) */
int b200fhe_upload(b200fhe_ctx *ctx, const uint32_t *slot_ids, const b200fhe_torus0 *tlwe_host, size_t n);
int b200fhe_download(b200fhe_ctx *ctx, const uint32_t *slot_ids, b200fhe_torus0 *tlwe_host, size_t n);

/* evaluation (asynchronous on the context's stream) --------------------------------------- */
/* One frontier: gate i reads slots in0[i], in1[i], in2[i] (unused operands ignored, may be NULL
 * arrays when no gate needs them) and writes slot out[i].  Gates of one batch must be
 * independent of each other. */
int b200fhe_gate_batch(b200fhe_ctx *ctx, const uint8_t *opcode, const uint32_t *in0, const uint32_t *in1,
                       const uint32_t *in2, const uint32_t *out, size_t n);
/* dst[i] <- src[i] for all i, as one parallel step (all reads happen before all writes) */
int b200fhe_dff_tick(b200fhe_ctx *ctx, const uint32_t *src, const uint32_t *dst, size_t n);
int b200fhe_sync(b200fhe_ctx *ctx);
int b200fhe_query(b200fhe_ctx *ctx); /* 0 = idle, 1 = busy, <0 = error */

/* programs: a static schedule recorded once, replayed as ONE CUDA graph ------------------------------ */
/* A netlist is the same every clock cycle, so its host (include/b200net.h) records the frontiers of a clock
 * once - b200fhe_program_batch per step, b200fhe_program_exchange where a step was sharded over ranks,
 * b200fhe_program_tick for the DFF update - and finalises: the job lists are uploaded once and all launches
 * are captured into one CUDA graph.  b200fhe_program_launch then costs one cudaGraphLaunch per clock.
 * This replaces the reference's per-gate scheduling turn: Worker::update pops ONE ready node, starts it and
 * polls it (src/iyokan.hpp:851-874, run loop src/iyokan_tfhepp.cpp:28-47); cuFHE adds three PCIe copies and
 * a cudaStreamQuery per gate (src/iyokan_cufhe.hpp:217-241).  Arguments as b200fhe_gate_batch / _dff_tick. */
typedef struct b200fhe_program b200fhe_program;
int b200fhe_program_create(b200fhe_ctx *ctx, b200fhe_program **out);
void b200fhe_program_destroy(b200fhe_program *prog);
int b200fhe_program_batch(b200fhe_program *prog, const uint8_t *opcode, const uint32_t *in0, const uint32_t *in1,
                          const uint32_t *in2, const uint32_t *out, size_t n);
int b200fhe_program_tick(b200fhe_program *prog, const uint32_t *src, const uint32_t *dst, size_t n);
int b200fhe_program_exchange(b200fhe_program *prog, size_t first_slot, size_t slots_per_rank);
int b200fhe_program_finalize(b200fhe_program *prog);
int b200fhe_program_launch(b200fhe_program *prog); /* asynchronous on the context's stream */
/* Profiling replay: the same launches issued one recorded step at a time (no graph) with a CUDA event between the
 * steps; synchronous.  step_ms[k] = device time of the k-th recorded batch / exchange / tick (up to cap entries),
 * *nsteps = how many were recorded.  What the reference's ProgressGraphMaker samples per node (src/iyokan.hpp:128-278)
 * is available here per frontier. */
int b200fhe_program_profile(b200fhe_program *prog, float *step_ms, size_t cap, size_t *nsteps);
/* any pointer may be NULL.  is_graph = 0 when the capture was refused and the program replays launch by launch */
int b200fhe_program_info(const b200fhe_program *prog, uint64_t *rotations, uint64_t *launches_per_replay,
                         uint64_t *exchanges, uint64_t *exchanged_slots, int *is_graph, double *model_ms);

/* multi-GPU exchange: one process per GPU, identical arena layout on every rank ------------------------ */
/* Replaces cufhe::SetGPUNum + round-robin streams through host memory (cuFHE include/cufhe_gpu.cuh:164-169,
 * src/iyokan_cufhe.cpp:533).  Rank 0 creates a 128-byte id (ncclUniqueId) and hands it to the other ranks
 * by any host channel; every rank then calls b200fhe_comm_init.  b200fhe_exchange is ONE in-place all-gather
 * over NVLink (NCCL, bound at run time with dlopen): rank r contributes slots
 * [first_slot + r*slots_per_rank, first_slot + (r+1)*slots_per_rank), afterwards every rank holds all of them. */
int b200fhe_comm_unique_id(uint8_t *id128);
int b200fhe_comm_init(b200fhe_ctx *ctx, int rank, int world, const uint8_t *id128);
int b200fhe_comm_rank(const b200fhe_ctx *ctx);
int b200fhe_comm_world(const b200fhe_ctx *ctx);
int b200fhe_exchange(b200fhe_ctx *ctx, size_t first_slot, size_t slots_per_rank);

/* End-to-end convenience with HOST operands: uploads, evaluates, downloads (synchronous).
 * in*_host / out_host are [n][637] uint16; needs an arena of at least 4*n slots. */
int b200fhe_gates_host(b200fhe_ctx *ctx, const uint8_t *opcode, const b200fhe_torus0 *in0_host,
                       const b200fhe_torus0 *in1_host, const b200fhe_torus0 *in2_host, b200fhe_torus0 *out_host, size_t n);

/* pinned host memory helpers (so callers written in any language can stage without torch) */
int b200fhe_host_alloc(void **ptr, size_t bytes);
int b200fhe_host_free(void *ptr);

/* instrumentation ------------------------------------------------------------------------- */
/* kernels launched by this context so far (all kinds) */
uint64_t b200fhe_launch_count(const b200fhe_ctx *ctx);
/* device time (ms, CUDA events on the context's stream) spent in the blind-rotation kernel /
 * key-switch kernel by the most recent b200fhe_gate_batch; valid after b200fhe_sync */
int b200fhe_last_batch_ms(b200fhe_ctx *ctx, float *blind_rotate_ms, float *keyswitch_ms);
/* the launch plan of that batch: a frontier is cut into at most five blind-rotation launches (full
 * waves of the throughput kernels, the tail on the latency kernels).  Fills up to max_segments entries
 * (kernel variant, jobs per CTA, rotation jobs, device ms) and returns the number of segments, -1 on error. */
int b200fhe_last_batch_segments(b200fhe_ctx *ctx, int *variant, int *jobs_per_cta, int *jobs, float *ms,
                                int max_segments);
/* the plan the heuristic would choose for a frontier of `njobs` blind rotations (pure host logic, no device) */
int b200fhe_plan_rotation(int njobs, int *variant, int *jobs_per_cta, int *jobs, int max_segments);
/* modelled blind-rotation time (ms, B200 tables) of that plan: what a scheduler uses to decide whether
 * splitting a dependency level across GPUs pays for the exchange */
double b200fhe_plan_ms(int njobs);
/* the table behind the plan: per shape (kernel variant, jobs per CTA) the jobs one wave of 148 SMs takes and its
 * duration.  Compile-time defaults were measured on a B200 at 1965 MHz; b200fhe_load_keys re-measures them on the
 * device at hand, and b200fhe_comm_init makes all ranks agree (max) so that every rank derives the same schedule.
 * Returns the number of shapes, negated while the table still holds the defaults. */
int b200fhe_plan_table(int *variant, int *jobs_per_cta, int *wave_jobs, double *wave_ms, int max_shapes);
void *b200fhe_stream(const b200fhe_ctx *ctx);

/* test hooks: stage-level access used by the parity tests --------------------------------- */
/* blind rotation + sample extraction only: c [n][637] uint16 (already linearly combined)
 * -> lvl1 TLWE [n][1025] uint32 (GateBootstrappingTLWE2TLWE, gatebootstrapping.hpp:188-197) */
int b200fhe_test_bootstrap_lvl1(b200fhe_ctx *ctx, const b200fhe_torus0 *c_host, uint32_t *tlwe1_host, size_t n);
/* identity key switch only: [n][1025] uint32 -> [n][637] uint16 (keyswitch.hpp:11-52) */
int b200fhe_test_keyswitch(b200fhe_ctx *ctx, const uint32_t *tlwe1_host, b200fhe_torus0 *tlwe0_host, size_t n);
/* NTT-domain bootstrapping key as stored on the device: uint32 [636][6][6][1024] */
int b200fhe_test_read_bk_ntt(b200fhe_ctx *ctx, uint32_t *out_host, size_t first_i, size_t count_i);

#ifdef __cplusplus
}
#endif
#endif
