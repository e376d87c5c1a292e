// b200net: levelised netlist engine above the b200fhe C ABI (see include/b200net.h).
#include "../../include/b200net.h"

#include <algorithm>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string& m)
{
    g_err = m;
    return 1;
}

int arity(uint8_t kind)
{
    if (kind <= B200FHE_XNOR || kind == B200FHE_ANDNY || kind == B200FHE_ORNY) return 2;
    if (kind == B200FHE_MUX) return 3;
    if (kind == B200FHE_NOT || kind == B200FHE_COPY || kind == B200NET_DFF || kind == B200NET_OUTPUT) return 1;
    if (kind == B200FHE_CONST0 || kind == B200FHE_CONST1 || kind == B200NET_INPUT) return 0;
    return -1;
}
bool is_gate(uint8_t kind) { return kind < B200FHE_NUM_OPS; }
int bootstraps(uint8_t kind)
{
    if (kind == B200FHE_MUX) return 2;
    return (is_gate(kind) && arity(kind) == 2) ? 1 : 0;
}

// plaintext truth tables, same opcode meaning as gate_jobs.h / TFHEpp gate.hpp
uint8_t plain_gate(uint8_t op, uint8_t a, uint8_t b, uint8_t s)
{
    switch (op) {
    case B200FHE_AND: return a & b;
    case B200FHE_NAND: return !(a & b);
    case B200FHE_ANDNOT: return a & !b;
    case B200FHE_OR: return a | b;
    case B200FHE_NOR: return !(a | b);
    case B200FHE_ORNOT: return a | !b;
    case B200FHE_XOR: return a ^ b;
    case B200FHE_XNOR: return !(a ^ b);
    case B200FHE_MUX: return s ? b : a;
    case B200FHE_NOT: return !a;
    case B200FHE_COPY: return a;
    case B200FHE_CONST0: return 0;
    case B200FHE_CONST1: return 1;
    case B200FHE_ANDNY: return (!a) & b;
    case B200FHE_ORNY: return (!a) | b;
    default: return 0;
    }
}

}  // namespace

struct b200net {
    size_t n = 0;
    std::vector<uint8_t> kind;
    std::vector<int32_t> in[3];
    std::vector<int32_t> level;                  // 0 = INPUT/DFF, k>=1 gates, OUTPUT = level of driver
    std::vector<std::vector<uint32_t>> levels;   // levels[k-1] = gate nodes of level k
    std::vector<uint32_t> dffs, outputs, inputs;
    std::vector<uint32_t> slot;                  // node -> arena slot
    std::vector<uint32_t> level_base;            // first slot of each gate level
    size_t n_slots = 0, scratch_base = 0;
    size_t n_boot = 0;
    int world = 1;
    b200fhe_ctx* ctx = nullptr;
    // per-level batch arrays, prepared once at bind time (the "replay" part)
    struct Batch {
        std::vector<uint8_t> op;
        std::vector<uint32_t> a, b, c, o;
    };
    std::vector<Batch> batches;
    std::vector<uint32_t> tick_src, tick_dst;
    // ---- compiled schedule (b200net_bind_rank): steps instead of ASAP levels, one CUDA graph per clock ----
    struct Step {
        std::vector<uint32_t> rep;   // gates every rank evaluates (no exchange)
        std::vector<uint32_t> shd;   // gates split over the ranks, outputs all-gathered after the step
        std::vector<size_t> cut;     // shd[cut[r] .. cut[r+1]) belongs to rank r (world + 1 entries)
        size_t chunk = 0;            // slots per rank of the sharded region
        uint32_t base_rep = 0, base_shd = 0;
        size_t jobs_rep = 0, jobs_shd = 0;
    };
    std::vector<Step> steps;
    std::vector<uint8_t> step_records;  // per schedule step: how many program steps (batch, exchange) it recorded
    bool compiled = false, packed = false;
    int rank = 0;
    double model_ms = 0.0;
    b200fhe_program *run_prog = nullptr, *tick_prog = nullptr;

    uint32_t resolve(uint32_t node) const
    {
        while (kind[node] == B200NET_OUTPUT) node = (uint32_t)in[0][node];
        return node;
    }
};

extern "C" {

const char* b200net_last_error(void) { return g_err.c_str(); }

int b200net_create(b200net** out, size_t n, const uint8_t* kind, const int32_t* in0, const int32_t* in1,
                   const int32_t* in2)
{
    if (!out || !kind || !in0 || !in1 || !in2) return fail("null argument");
    *out = nullptr;
    auto net = new b200net();
    net->n = n;
    net->kind.assign(kind, kind + n);
    net->in[0].assign(in0, in0 + n);
    net->in[1].assign(in1, in1 + n);
    net->in[2].assign(in2, in2 + n);
    auto bail = [&](const std::string& m) {
        delete net;
        return fail(m);
    };
    // validation (reference: TaskNetwork::checkValid, src/iyokan.hpp:1002-1015)
    for (size_t i = 0; i < n; i++) {
        const int ar = arity(kind[i]);
        if (ar < 0) return bail("node " + std::to_string(i) + ": unknown kind " + std::to_string(kind[i]));
        for (int k = 0; k < 3; k++) {
            const int32_t v = net->in[k][i];
            if (k < ar) {
                if (v < 0 || (size_t)v >= n) return bail("node " + std::to_string(i) + ": input out of range");
                if (net->kind[v] == B200NET_OUTPUT && false) return bail("unreachable");
            } else if (v != -1) {
                return bail("node " + std::to_string(i) + ": too many inputs for its kind");
            }
        }
        if (kind[i] == B200NET_INPUT) net->inputs.push_back((uint32_t)i);
        if (kind[i] == B200NET_DFF) net->dffs.push_back((uint32_t)i);
        if (kind[i] == B200NET_OUTPUT) net->outputs.push_back((uint32_t)i);
        net->n_boot += bootstraps(kind[i]);
    }
    // levelise with Kahn's algorithm over the combinational edges (DFF outputs are sources, DFF
    // D-inputs are sinks), detecting combinational loops
    std::vector<int32_t> pending(n, 0);
    std::vector<std::vector<uint32_t>> users(n);
    for (size_t i = 0; i < n; i++) {
        if (kind[i] == B200NET_DFF || kind[i] == B200NET_INPUT) continue;
        const int ar = arity(kind[i]);
        pending[i] = ar;
        for (int k = 0; k < ar; k++) users[net->in[k][i]].push_back((uint32_t)i);
    }
    net->level.assign(n, -1);
    std::vector<uint32_t> ready;
    for (size_t i = 0; i < n; i++)
        if (pending[i] == 0) {
            net->level[i] = (kind[i] == B200NET_DFF || kind[i] == B200NET_INPUT) ? 0 : 1;  // CONST gates: level 1
            ready.push_back((uint32_t)i);
        }
    size_t done = 0;
    while (!ready.empty()) {
        const uint32_t u = ready.back();
        ready.pop_back();
        done++;
        for (uint32_t v : users[u]) {
            const int32_t lu = net->level[u] + (net->kind[v] == B200NET_OUTPUT ? 0 : 1);
            net->level[v] = std::max(net->level[v], lu);
            if (--pending[v] == 0) ready.push_back(v);
        }
    }
    if (done != n) return bail("combinational loop in the netlist");
    int32_t depth = 0;
    for (size_t i = 0; i < n; i++)
        if (is_gate(kind[i])) depth = std::max(depth, net->level[i]);
    net->levels.resize(depth);
    for (size_t i = 0; i < n; i++)
        if (is_gate(kind[i])) net->levels[net->level[i] - 1].push_back((uint32_t)i);
    *out = net;
    return 0;
}

void b200net_destroy(b200net* net)
{
    if (!net) return;
    b200fhe_program_destroy(net->run_prog);
    b200fhe_program_destroy(net->tick_prog);
    delete net;
}
size_t b200net_num_nodes(const b200net* net) { return net->n; }
size_t b200net_num_levels(const b200net* net) { return net->levels.size(); }
size_t b200net_level_width(const b200net* net, size_t l) { return l < net->levels.size() ? net->levels[l].size() : 0; }
size_t b200net_level_bootstraps(const b200net* net, size_t l)
{
    if (l >= net->levels.size()) return 0;
    size_t n = 0;
    for (uint32_t node : net->levels[l]) n += bootstraps(net->kind[node]);
    return n;
}
size_t b200net_bootstraps_per_cycle(const b200net* net) { return net->n_boot; }
size_t b200net_num_dff(const b200net* net) { return net->dffs.size(); }
int32_t b200net_node_level(const b200net* net, size_t node) { return node < net->n ? net->level[node] : -1; }
uint32_t b200net_slot_of(const b200net* net, size_t node)
{
    return (node < net->n && !net->slot.empty()) ? net->slot[net->resolve((uint32_t)node)] : 0xFFFFFFFFu;
}
size_t b200net_num_slots(const b200net* net) { return net->n_slots; }
uint32_t b200net_level_slot_base(const b200net* net, size_t l)
{
    return l < net->level_base.size() ? net->level_base[l] : 0xFFFFFFFFu;
}

int b200net_plain_eval(const b200net* net, uint8_t* v)
{
    if (!net || !v) return fail("null argument");
    for (const auto& lv : net->levels)
        for (uint32_t g : lv) {
            const uint8_t a = net->in[0][g] >= 0 ? v[net->resolve(net->in[0][g])] : 0;
            const uint8_t b = net->in[1][g] >= 0 ? v[net->resolve(net->in[1][g])] : 0;
            const uint8_t s = net->in[2][g] >= 0 ? v[net->resolve(net->in[2][g])] : 0;
            v[g] = plain_gate(net->kind[g], a & 1, b & 1, s & 1) & 1;
        }
    for (uint32_t o : net->outputs) v[o] = v[net->resolve(o)];
    return 0;
}

int b200net_plain_tick(const b200net* net, uint8_t* v)
{
    if (!net || !v) return fail("null argument");
    std::vector<uint8_t> d(net->dffs.size());
    for (size_t i = 0; i < d.size(); i++) d[i] = v[net->resolve(net->in[0][net->dffs[i]])];
    for (size_t i = 0; i < d.size(); i++) v[net->dffs[i]] = d[i];
    return 0;
}

// Slot layout: [sources (INPUT, DFF)] [level 1, padded to world] [level 2, padded] ...
int b200net_layout(b200net* net, int world)
{
    if (!net) return fail("null argument");
    if (world < 1) return fail("world_size must be >= 1");
    net->world = world;
    net->slot.assign(net->n, 0xFFFFFFFFu);
    uint32_t next = 0;
    for (uint32_t i : net->inputs) net->slot[i] = next++;
    for (uint32_t i : net->dffs) net->slot[i] = next++;
    net->level_base.clear();
    for (const auto& lv : net->levels) {
        net->level_base.push_back(next);
        for (size_t k = 0; k < lv.size(); k++) net->slot[lv[k]] = next + (uint32_t)k;
        const size_t chunk = (lv.size() + world - 1) / world;
        next += (uint32_t)(chunk * world);
    }
    net->n_slots = next;
    net->batches.clear();
    for (const auto& lv : net->levels) {
        b200net::Batch bt;
        for (uint32_t g : lv) {
            bt.op.push_back(net->kind[g]);
            auto s = [&](int k) { return net->in[k][g] >= 0 ? net->slot[net->resolve(net->in[k][g])] : 0u; };
            bt.a.push_back(s(0));
            bt.b.push_back(s(1));
            bt.c.push_back(s(2));
            bt.o.push_back(net->slot[g]);
        }
        net->batches.push_back(std::move(bt));
    }
    net->tick_src.clear();
    net->tick_dst.clear();
    for (uint32_t d : net->dffs) {
        net->tick_src.push_back(net->slot[net->resolve(net->in[0][d])]);
        net->tick_dst.push_back(net->slot[d]);
    }
    return 0;
}

int b200net_bind(b200net* net, b200fhe_ctx* ctx, int world)
{
    if (!net || !ctx) return fail("null argument");
    if (b200net_layout(net, world)) return 1;
    net->ctx = ctx;
    net->compiled = false;
    if (b200fhe_arena_alloc(ctx, std::max<size_t>(net->n_slots, 1))) return fail(b200fhe_last_error());
    return 0;
}

int b200net_set(b200net* net, const uint32_t* nodes, const b200fhe_torus0* tlwe, size_t n)
{
    if (!net || !net->ctx) return fail("netlist is not bound to a context");
    std::vector<uint32_t> s(n);
    for (size_t i = 0; i < n; i++) {
        if (nodes[i] >= net->n) return fail("node out of range");
        const uint8_t k = net->kind[nodes[i]];
        if (k != B200NET_INPUT && k != B200NET_DFF) return fail("only INPUT and DFF nodes can be set");
        s[i] = net->slot[nodes[i]];
    }
    if (b200fhe_upload(net->ctx, s.data(), tlwe, n) || b200fhe_sync(net->ctx)) return fail(b200fhe_last_error());
    return 0;
}

int b200net_restore(b200net* net, const uint32_t* nodes, const b200fhe_torus0* tlwe, size_t n)
{
    if (!net || !net->ctx) return fail("netlist is not bound to a context");
    std::vector<uint32_t> s(n);
    for (size_t i = 0; i < n; i++) {
        if (nodes[i] >= net->n) return fail("node out of range");
        if (net->kind[nodes[i]] == B200NET_OUTPUT) return fail("OUTPUT wires alias their driver and hold no state");
        s[i] = net->slot[nodes[i]];
    }
    if (b200fhe_upload(net->ctx, s.data(), tlwe, n) || b200fhe_sync(net->ctx)) return fail(b200fhe_last_error());
    return 0;
}

int b200net_get(b200net* net, const uint32_t* nodes, b200fhe_torus0* tlwe, size_t n)
{
    if (!net || !net->ctx) return fail("netlist is not bound to a context");
    std::vector<uint32_t> s(n);
    for (size_t i = 0; i < n; i++) {
        if (nodes[i] >= net->n) return fail("node out of range");
        s[i] = net->slot[net->resolve(nodes[i])];
    }
    if (b200fhe_download(net->ctx, s.data(), tlwe, n)) return fail(b200fhe_last_error());
    return 0;
}

int b200net_tick(b200net* net)
{
    if (!net || !net->ctx) return fail("netlist is not bound to a context");
    if (net->compiled) {
        if (net->tick_src.empty()) return 0;
        if (b200fhe_program_launch(net->tick_prog)) return fail(b200fhe_last_error());
        return 0;
    }
    if (b200fhe_dff_tick(net->ctx, net->tick_src.data(), net->tick_dst.data(), net->tick_src.size()))
        return fail(b200fhe_last_error());
    return 0;
}

int b200net_run_level_shard(b200net* net, size_t level, int rank, int world)
{
    if (!net || !net->ctx) return fail("netlist is not bound to a context");
    if (net->compiled) return fail("netlist was compiled with b200net_bind_rank: use b200net_run");
    if (level >= net->batches.size()) return fail("level out of range");
    // world may be 1 (whole level, used for levels too narrow to shard) or the bound world size
    if ((world != net->world && world != 1) || rank < 0 || rank >= world) return fail("rank/world mismatch with bind");
    const auto& bt = net->batches[level];
    const size_t w = bt.op.size(), chunk = (w + world - 1) / world;
    const size_t lo = std::min(w, (size_t)rank * chunk), hi = std::min(w, lo + chunk);
    if (hi == lo) return 0;
    if (b200fhe_gate_batch(net->ctx, bt.op.data() + lo, bt.a.data() + lo, bt.b.data() + lo, bt.c.data() + lo,
                           bt.o.data() + lo, hi - lo))
        return fail(b200fhe_last_error());
    return 0;
}

int b200net_run(b200net* net)
{
    if (!net || !net->ctx) return fail("netlist is not bound to a context");
    if (net->compiled) {
        if (b200fhe_program_launch(net->run_prog)) return fail(b200fhe_last_error());
        return 0;
    }
    for (size_t l = 0; l < net->batches.size(); l++)
        if (b200net_run_level_shard(net, l, 0, 1)) return 1;
    return 0;
}

}  // extern "C"

// ---- static scheduling ---------------------------------------------------------------------------------
// The reference orders ready nodes by HEFT upward rank (graph::doRankuSort, src/iyokan.cpp:4-98) and lets
// N workers pop them one at a time.  Here the whole clock cycle is scheduled once: a sequence of STEPS,
// each one frontier of independent gates.  With packing off a step is an ASAP dependency level.  With
// packing on, gates that have slack (ALAP level later than ASAP level) are pulled forward into steps whose
// mandatory gates leave SMs idle, as long as the launch-plan model (b200fhe_plan_ms) says the step gets
// cheaper per job than the throughput shape would later charge for them; the schedule with the lower
// modelled clock time is kept.  On `world` ranks a step is either replicated (every rank evaluates all of
// it, no exchange) or sharded (contiguous shares balanced by rotation count, outputs all-gathered).
namespace {

struct Scheduler {
    const b200net& net;
    int world;
    double exch_ms;
    std::vector<int32_t> alap;
    std::vector<uint8_t> jobs;
    Scheduler(const b200net& n, int w, double x) : net(n), world(w), exch_ms(x) {}

    // Rotation jobs the busiest rank gets when `total` jobs are cut into `world` contiguous shares of whole gates
    // (a MUX is two jobs and is not split): every share is filled up to this load, an earlier share may fall one
    // short of it, so one spare job per earlier rank is budgeted.  The launch plan has cliffs (75 jobs cost one more
    // launch than 74), so the shares must never exceed what the cost model assumed.
    size_t share(size_t total) const { return world == 1 ? total : (total + 2 * (size_t)world - 2) / world; }

    // cost (ms) of one step with `total` rotation jobs: replicated, or sharded over the ranks
    double step_cost(size_t total, bool* shard) const
    {
        const double rep = b200fhe_plan_ms((int)total);
        if (world == 1 || total == 0) {
            if (shard) *shard = false;
            return rep;
        }
        const double shd = b200fhe_plan_ms((int)share(total)) + exch_ms;
        if (shard) *shard = shd < rep;
        return shd < rep ? shd : rep;
    }

    void compute_alap()
    {
        const size_t n = net.n;
        const int32_t D = (int32_t)net.levels.size();
        alap.assign(n, D);
        jobs.assign(n, 0);
        for (size_t i = 0; i < n; i++) jobs[i] = (uint8_t)bootstraps(net.kind[i]);
        // reverse ASAP order: users of a gate are on strictly later levels
        for (int32_t l = D - 1; l >= 0; l--)
            for (uint32_t g : net.levels[l]) {
                const int ar = arity(net.kind[g]);
                for (int k = 0; k < ar; k++) {
                    const uint32_t src = net.resolve((uint32_t)net.in[k][g]);
                    if (is_gate(net.kind[src])) alap[src] = std::min(alap[src], alap[g] - 1);
                }
            }
    }

    // returns modelled ms; fills steps (without slots)
    double build(bool pack, std::vector<b200net::Step>& steps) const
    {
        const size_t n = net.n;
        const int32_t D = (int32_t)net.levels.size();
        steps.clear();
        steps.resize(D);
        std::vector<int32_t> when(n, -1);  // step (1-based) a node's value is produced in; 0 = source
        for (size_t i = 0; i < n; i++)
            if (net.kind[i] == B200NET_INPUT || net.kind[i] == B200NET_DFF) when[i] = 0;
        // the throughput rate later steps would charge for a job that is not pulled forward
        const double wide = 4736.0;
        const double ms_per_job = step_cost((size_t)wide * world, nullptr) / (wide * world);
        std::vector<uint32_t> pending;  // gates not yet scheduled, kept sorted by (alap, asap)
        for (const auto& lv : net.levels) pending.insert(pending.end(), lv.begin(), lv.end());
        std::stable_sort(pending.begin(), pending.end(), [&](uint32_t a, uint32_t b) {
            return alap[a] != alap[b] ? alap[a] < alap[b] : net.level[a] < net.level[b];
        });
        auto ready = [&](uint32_t g, int32_t L) {
            const int ar = arity(net.kind[g]);
            for (int k = 0; k < ar; k++) {
                const int32_t w = when[net.resolve((uint32_t)net.in[k][g])];
                if (w < 0 || w >= L) return false;
            }
            return true;
        };
        double total_ms = 0.0;
        for (int32_t L = 1; L <= D; L++) {
            std::vector<uint32_t> take, cand;
            size_t mj = 0;
            for (uint32_t g : pending) {
                const bool mandatory = pack ? alap[g] <= L : net.level[g] == L;
                if (mandatory) {
                    take.push_back(g);
                    mj += jobs[g];
                } else if (pack && net.level[g] <= L && ready(g, L)) {
                    if (jobs[g] == 0) take.push_back(g);  // bootstrap-free gates cost nothing: as soon as possible
                    else cand.push_back(g);
                }
            }
            if (pack && !cand.empty()) {
                // candidate fill targets: the capacity points of the launch plan, per rank
                static const size_t points[] = {74, 148, 222, 296, 592, 888, 1184, 2368, 3552, 4736};
                std::vector<size_t> pre(cand.size() + 1, 0);
                for (size_t k = 0; k < cand.size(); k++) pre[k + 1] = pre[k] + jobs[cand[k]];
                double best = step_cost(mj, nullptr);
                size_t best_k = 0;
                for (size_t pt : points) {
                    // the largest total whose busiest share still fits the capacity point
                    const size_t target = world == 1 ? pt : pt * (size_t)world - 2 * ((size_t)world - 1);
                    if (target <= mj) continue;
                    // largest prefix of the candidates that fits the target
                    size_t k = std::upper_bound(pre.begin(), pre.end(), target - mj) - pre.begin() - 1;
                    if (k == 0) continue;
                    const double score = step_cost(mj + pre[k], nullptr) - ms_per_job * (double)pre[k];
                    if (score < best - 1e-9) best = score, best_k = k;
                    if (k == cand.size()) break;
                }
                take.insert(take.end(), cand.begin(), cand.begin() + best_k);
                mj += pre[best_k];
            }
            for (uint32_t g : take) when[g] = L;
            pending.erase(std::remove_if(pending.begin(), pending.end(), [&](uint32_t g) { return when[g] == L; }),
                          pending.end());
            // keep slot order deterministic: by node id inside a step
            std::sort(take.begin(), take.end());
            b200net::Step& st = steps[L - 1];
            bool shard = false;
            total_ms += step_cost(mj, &shard);
            if (shard) {
                st.shd = std::move(take);
                st.jobs_shd = mj;
            } else {
                st.rep = std::move(take);
                st.jobs_rep = mj;
            }
        }
        return total_ms;
    }
};

void assign_slots(b200net* net, int world)
{
    net->slot.assign(net->n, 0xFFFFFFFFu);
    uint32_t next = 0;
    for (uint32_t i : net->inputs) net->slot[i] = next++;
    for (uint32_t i : net->dffs) net->slot[i] = next++;
    for (auto& st : net->steps) {
        st.base_rep = next;
        for (uint32_t g : st.rep) net->slot[g] = next++;
        st.cut.assign(world + 1, 0);
        st.chunk = 0;
        st.base_shd = next;
        if (!st.shd.empty()) {
            // contiguous shares of whole gates, each filled up to the load the cost model assumed (Scheduler::share):
            // gates without a rotation (NOT / COPY / CONST) ride along for free
            size_t total = 0;
            for (uint32_t g : st.shd) total += bootstraps(net->kind[g]);
            const size_t load = (total + 2 * (size_t)world - 2) / world;
            size_t k = 0;
            for (int r = 0; r < world; r++) {
                st.cut[r] = k;
                size_t have = 0;
                while (k < st.shd.size() && (r == world - 1 || have + bootstraps(net->kind[st.shd[k]]) <= load))
                    have += bootstraps(net->kind[st.shd[k++]]);
            }
            st.cut[world] = st.shd.size();
            for (int q = 0; q < world; q++) st.chunk = std::max(st.chunk, st.cut[q + 1] - st.cut[q]);
            for (int q = 0; q < world; q++)
                for (size_t k = st.cut[q]; k < st.cut[q + 1]; k++)
                    net->slot[st.shd[k]] = st.base_shd + (uint32_t)(q * st.chunk + (k - st.cut[q]));
            next += (uint32_t)(st.chunk * world);
        }
    }
    net->n_slots = next;
    net->tick_src.clear();
    net->tick_dst.clear();
    for (uint32_t d : net->dffs) {
        net->tick_src.push_back(net->slot[net->resolve(net->in[0][d])]);
        net->tick_dst.push_back(net->slot[d]);
    }
}

}  // namespace

extern "C" {

int b200net_schedule(b200net* net, int world, unsigned flags)
{
    if (!net) return fail("null argument");
    if (world < 1) return fail("world_size must be >= 1");
    Scheduler sc(*net, world, 0.05);
    sc.compute_alap();
    std::vector<b200net::Step> asap, packed;
    const double ms_asap = sc.build(false, asap);
    net->packed = false;
    net->model_ms = ms_asap;
    net->steps = std::move(asap);
    if (flags & B200NET_PACK) {
        const double ms_packed = sc.build(true, packed);
        if (ms_packed < ms_asap) {
            net->steps = std::move(packed);
            net->packed = true;
            net->model_ms = ms_packed;
        }
    }
    // drop empty trailing / inner steps (packing can empty a level)
    net->steps.erase(std::remove_if(net->steps.begin(), net->steps.end(),
                                    [](const b200net::Step& s) { return s.rep.empty() && s.shd.empty(); }),
                     net->steps.end());
    net->world = world;
    assign_slots(net, world);
    net->batches.clear();  // the per-level arrays of b200net_layout do not describe this slot assignment
    net->level_base.clear();
    return 0;
}

int b200net_bind_rank(b200net* net, b200fhe_ctx* ctx, int rank, int world, unsigned flags)
{
    if (!net || !ctx) return fail("null argument");
    if (rank < 0 || rank >= world) return fail("rank out of range");
    if (world != b200fhe_comm_world(ctx) && world != 1) return fail("world size differs from the context's communicator");
    if (b200net_schedule(net, world, flags)) return 1;
    net->ctx = ctx;
    net->rank = rank;
    if (b200fhe_arena_alloc(ctx, std::max<size_t>(net->n_slots, 1))) return fail(b200fhe_last_error());
    b200fhe_program_destroy(net->run_prog);
    b200fhe_program_destroy(net->tick_prog);
    net->run_prog = net->tick_prog = nullptr;
    if (b200fhe_program_create(ctx, &net->run_prog) || b200fhe_program_create(ctx, &net->tick_prog))
        return fail(b200fhe_last_error());
    std::vector<uint8_t> op;
    std::vector<uint32_t> a, b, c, o;
    net->step_records.clear();
    for (const auto& st : net->steps) {
        op.clear(), a.clear(), b.clear(), c.clear(), o.clear();
        auto add = [&](uint32_t g) {
            op.push_back(net->kind[g]);
            auto s = [&](int k) { return net->in[k][g] >= 0 ? net->slot[net->resolve(net->in[k][g])] : 0u; };
            a.push_back(s(0));
            b.push_back(s(1));
            c.push_back(s(2));
            o.push_back(net->slot[g]);
        };
        for (uint32_t g : st.rep) add(g);
        if (!st.shd.empty())
            for (size_t k = st.cut[rank]; k < st.cut[rank + 1]; k++) add(st.shd[k]);
        uint8_t recorded = 0;
        if (!op.empty()) {
            if (b200fhe_program_batch(net->run_prog, op.data(), a.data(), b.data(), c.data(), o.data(), op.size()))
                return fail(b200fhe_last_error());
            recorded++;
        }
        if (!st.shd.empty() && world > 1) {
            if (b200fhe_program_exchange(net->run_prog, st.base_shd, st.chunk)) return fail(b200fhe_last_error());
            recorded++;
        }
        net->step_records.push_back(recorded);
    }
    if (!net->tick_src.empty() &&
        b200fhe_program_tick(net->tick_prog, net->tick_src.data(), net->tick_dst.data(), net->tick_src.size()))
        return fail(b200fhe_last_error());
    if (b200fhe_program_finalize(net->run_prog) || b200fhe_program_finalize(net->tick_prog))
        return fail(b200fhe_last_error());
    net->compiled = true;
    return 0;
}

int b200net_profile_run(b200net* net, float* step_ms, size_t cap)
{
    if (!net || !net->ctx) return fail("netlist is not bound to a context");
    if (!net->compiled) return fail("b200net_profile_run needs b200net_bind_rank");
    size_t nrec = 0;
    for (uint8_t r : net->step_records) nrec += r;
    std::vector<float> rec(nrec, 0.0f);
    size_t seen = 0;
    if (b200fhe_program_profile(net->run_prog, rec.data(), rec.size(), &seen)) return fail(b200fhe_last_error());
    if (seen != nrec) return fail("recorded steps of the program do not match the schedule");
    size_t k = 0;
    for (size_t s = 0; s < net->step_records.size(); s++) {
        float ms = 0;
        for (uint8_t r = 0; r < net->step_records[s]; r++) ms += rec[k++];
        if (step_ms && s < cap) step_ms[s] = ms;
    }
    return 0;
}

size_t b200net_num_steps(const b200net* net) { return net ? net->steps.size() : 0; }
size_t b200net_step_gates(const b200net* net, size_t step, int rank, uint32_t* nodes, size_t cap)
{
    if (!net || step >= net->steps.size() || rank < 0 || rank >= net->world) return 0;
    const auto& st = net->steps[step];
    size_t n = 0;
    auto put = [&](uint32_t g) {
        if (nodes && n < cap) nodes[n] = g;
        n++;
    };
    for (uint32_t g : st.rep) put(g);
    if (!st.shd.empty())
        for (size_t k = st.cut[rank]; k < st.cut[rank + 1]; k++) put(st.shd[k]);
    return n;
}
int b200net_step_exchange(const b200net* net, size_t step, uint32_t* first_slot, size_t* slots_per_rank)
{
    if (!net || step >= net->steps.size()) return fail("step out of range");
    const auto& st = net->steps[step];
    if (first_slot) *first_slot = st.base_shd;
    if (slots_per_rank) *slots_per_rank = (st.shd.empty() || net->world == 1) ? 0 : st.chunk;
    return 0;
}
size_t b200net_step_jobs(const b200net* net, size_t step, int* sharded)
{
    if (!net || step >= net->steps.size()) return 0;
    const auto& st = net->steps[step];
    if (sharded) *sharded = st.shd.empty() ? 0 : 1;
    return st.jobs_rep + st.jobs_shd;
}
int b200net_schedule_info(const b200net* net, size_t* steps, size_t* collectives, size_t* exchanged_slots,
                          double* model_ms, int* packed)
{
    if (!net) return fail("null argument");
    size_t nc = 0, ns = 0;
    for (const auto& st : net->steps)
        if (!st.shd.empty() && net->world > 1) nc++, ns += st.chunk * net->world;
    if (steps) *steps = net->steps.size();
    if (collectives) *collectives = nc;
    if (exchanged_slots) *exchanged_slots = ns;
    if (model_ms) *model_ms = net->model_ms;
    if (packed) *packed = net->packed ? 1 : 0;
    return 0;
}

}  // extern "C"
