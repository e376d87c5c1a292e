/* b200net — host-side netlist engine above the b200fhe C ABI.
 *
 * Role in the reference: the part of Iyokan's scheduler that sits immediately before the hot path —
 * NetworkRunner::run / Worker::update / ReadyQueue / DepNode::propagate (src/iyokan.hpp:774-883,
 * 1982-2062; run loop src/iyokan_tfhepp.cpp:28-47) and NetworkRunner::tick (:2050-2054).  The
 * netlist is static across clock cycles, so instead of popping one ready node per Worker turn the
 * engine levelises the DAG once and replays it: ONE b200fhe_gate_batch per dependency level and ONE
 * b200fhe_dff_tick per clock.  It also carries a plaintext evaluator of the same
 * netlist, the role Iyokan's plain back-end plays in its tests (src/iyokan_plain.hpp:80-117).
 *
 * A netlist is given as flat arrays over node ids 0..n-1 (the blueprint / Yosys-JSON front end
 * stays outside: tests/tools/netlist_tools.py converts the reference's formats).
 */
#ifndef B200NET_H
#define B200NET_H
#include <stddef.h>
#include <stdint.h>

#include "b200fhe.h"

#ifdef __cplusplus
extern "C" {
#endif

/* node kinds: gate opcodes 0..14 of enum b200fhe_op, plus */
#define B200NET_INPUT 32  /* value supplied by the host every cycle (INPUT / ROM wires) */
#define B200NET_DFF 33    /* Q <- D at tick; in0 = D (TaskDFF, RAM cells) */
#define B200NET_OUTPUT 34 /* wire: alias of in0 (TaskWIRE OUTPUT) */

typedef struct b200net b200net;

/* Builds, validates (arity, ranges, combinational loops) and levelises.  in*[i] = -1 when unused. */
int b200net_create(b200net **out, size_t n_nodes, const uint8_t *kind, const int32_t *in0,
                   const int32_t *in1, const int32_t *in2);
void b200net_destroy(b200net *net);
const char *b200net_last_error(void);

/* structure queries */
size_t b200net_num_nodes(const b200net *net);
size_t b200net_num_levels(const b200net *net);              /* combinational depth in gate levels */
size_t b200net_level_width(const b200net *net, size_t level); /* gates (incl. NOT/COPY/CONST) in a level */
size_t b200net_level_bootstraps(const b200net *net, size_t level); /* blind rotations the level costs */
size_t b200net_bootstraps_per_cycle(const b200net *net);    /* 2-input gates = 1, MUX = 2, others 0 */
size_t b200net_num_dff(const b200net *net);
int32_t b200net_node_level(const b200net *net, size_t node); /* 0 for INPUT/DFF, k for gates */
/* arena slot that holds a node's ciphertext (OUTPUT nodes resolve to their driver) */
uint32_t b200net_slot_of(const b200net *net, size_t node);
size_t b200net_num_slots(const b200net *net);
/* slots of one level are contiguous: [base, base + width); padded_width(world) = world*ceil(width/world) */
uint32_t b200net_level_slot_base(const b200net *net, size_t level);

/* plaintext back-end: values[n_nodes] holds one bit per node.  Caller fills INPUT nodes (and DFF
 * nodes before the first cycle); eval fills every gate / OUTPUT node; tick does Q <- D. */
int b200net_plain_eval(const b200net *net, uint8_t *values);
int b200net_plain_tick(const b200net *net, uint8_t *values);

/* assigns arena slots for a given number of ranks (no GPU needed); implied by b200net_bind */
int b200net_layout(b200net *net, int world_size);
/* encrypted back-end on one GPU context (keys already loaded).  Lays out and allocates the arena. */
int b200net_bind(b200net *net, b200fhe_ctx *ctx, int world_size);
/* tlwe is [n][TLWE0_LEN] lvl0 torus words; nodes must be INPUT or DFF nodes for set, any node for get */
int b200net_set(b200net *net, const uint32_t *nodes, const b200fhe_torus0 *tlwe, size_t n);
/* resume from a snapshot (iyokan --resume, src/iyokan_tfhepp.cpp:603-617): restores the value of ANY node that
 * holds state - gate outputs included, since the next tick copies them into the DFFs */
int b200net_restore(b200net *net, const uint32_t *nodes, const b200fhe_torus0 *tlwe, size_t n);
int b200net_get(b200net *net, const uint32_t *nodes, b200fhe_torus0 *tlwe, size_t n);
int b200net_tick(b200net *net);                               /* all DFFs: Q <- D, one call */
int b200net_run(b200net *net);                                /* every level, whole width */
/* multi-GPU building block: evaluate only rank's contiguous share of one level */
int b200net_run_level_shard(b200net *net, size_t level, int rank, int world_size);

/* ---- static schedule + compiled replay (the path the front ends use) ----------------------------------- */
#define B200NET_PACK 1u /* pull gates with slack forward into under-full steps when the launch-plan model gains */
/* Schedules one clock cycle for `world` ranks (steps = frontiers; each replicated on every rank or sharded with
 * one all-gather), assigns arena slots, and compiles rank `rank`'s share into two b200fhe programs (clock, tick),
 * each ONE CUDA graph.  ctx must have keys loaded and, for world > 1, b200fhe_comm_init done.
 * Afterwards b200net_run / b200net_tick replay the graphs; set / get / restore work as after b200net_bind. */
int b200net_bind_rank(b200net *net, b200fhe_ctx *ctx, int rank, int world, unsigned flags);
/* the schedule alone (no GPU): what b200net_bind_rank would run */
int b200net_schedule(b200net *net, int world, unsigned flags);
size_t b200net_num_steps(const b200net *net);
size_t b200net_step_jobs(const b200net *net, size_t step, int *sharded); /* rotation jobs of a step, all ranks */
/* the gates `rank` evaluates in a step (replicated ones + its share); returns the count, fills up to cap node ids */
size_t b200net_step_gates(const b200net *net, size_t step, int rank, uint32_t *nodes, size_t cap);
/* the all-gather that follows a step: slots [first + r*per_rank, first + (r+1)*per_rank) come from rank r; per_rank = 0: none */
int b200net_step_exchange(const b200net *net, size_t step, uint32_t *first_slot, size_t *slots_per_rank);
int b200net_schedule_info(const b200net *net, size_t *steps, size_t *collectives, size_t *exchanged_slots,
                          double *model_ms, int *packed);
/* One evaluation of the clock like b200net_run, but step by step with device timers (b200fhe_program_profile):
 * step_ms[s] = time of schedule step s on this rank, its all-gather included (up to cap entries).  Synchronous.
 * Feeds the per-node time / graph dumps of the front end (the reference's ProgressGraphMaker, src/iyokan.hpp:128-278). */
int b200net_profile_run(b200net *net, float *step_ms, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
