// iyokan_b200.hpp — the B200 back-end as a plugin of Iyokan's KEPT scheduler.
//
// Compiled against the UNMODIFIED reference (src/iyokan.hpp): it supplies exactly what a back-end supplies there
//   struct XWorkerInfo, Task subclasses per gate kind, TaskXGateDFF / WIRE / Mem, XNetworkBuilder (the 12 name##Impl
//   factories, src/iyokan.hpp:1259-1282), XWorker (getWorkerInfo, :829-883), a NetworkRunner instantiation (:1982-2062),
//   processAllGates (the test hook, src/iyokan_tfhepp.cpp:577-601)
// in the shape of the cuFHE flavour (src/iyokan_cufhe.hpp:207-312), with two differences that make it a B200 back-end:
//   * ciphertexts never leave the GPU: a Task's value is a SLOT of the device-resident arena (B200Slot, 4 bytes), not a
//     TLWELvl0 in host memory copied around every gate (cufhe_gates_gpu.cu:145-157);
//   * a gate's startAsyncImpl does not launch anything: it RECORDS (opcode, input slots, output slot) into the frontier
//     of the process-wide B200Runtime.  The run loop lets every idle Worker pop one ready node (Worker::update, kept as
//     it is), then submits everything recorded in that sweep as ONE b200fhe_gate_batch; stream order makes the results
//     visible to the next batch, so the tasks are marked finished at submission and the stock propagate loop carries on.
//     cuFHE keeps 800 streams and polls cudaStreamQuery per gate instead (src/iyokan_cufhe.cpp:259, iyokan_cufhe.hpp:234-237).
// DFF ticks are recorded the same way and submitted as one b200fhe_dff_tick before the next sweep: all registers tick
// simultaneously (the reference copies them one by one in hash-map order, src/iyokan.hpp:982-986).
//
// tests/ref_link/b200_test0.cpp instantiates the reference's own templated unit tests (src/test0.cpp:43-455) with
// B200NetworkBuilder; `make -C oracle reflink` builds it, tests/test_gpu_ref_link.py runs it on the B200.
#ifndef IYOKAN_B200_HPP
#define IYOKAN_B200_HPP

#include <algorithm>
#include <memory>
#include <stdexcept>
#include <vector>

#include "iyokan.hpp"
#include "packet.hpp"
#include "tfhepp_cufhe_wrapper.hpp"

extern "C" {
#include "b200fhe.h"
}

static_assert(sizeof(TLWELvl0) == B200FHE_TLWE0_LEN * sizeof(uint16_t), "TLWELvl0 is the wire format of the C ABI");

// One per process: the GPU context, the slot allocator, the frontier under construction, the pending ticks.
class B200Runtime {
private:
    b200fhe_ctx* ctx_ = nullptr;
    size_t capacity_ = 0;
    uint32_t next_ = 0;
    std::vector<uint32_t> free_;
    // frontier of the current sweep
    std::vector<uint8_t> op_;
    std::vector<uint32_t> in0_, in1_, in2_, out_;
    std::vector<bool*> done_;
    std::vector<uint32_t> tickSrc_, tickDst_;
    size_t numBatches_ = 0, numGates_ = 0, maxBatch_ = 0;

    void ck(int rc, const char* what) const
    {
        if (rc) error::die("b200fhe: ", what, ": ", b200fhe_last_error());  // the library never exits (error.hpp:21-47)
    }

public:
    static B200Runtime& instance()
    {
        static B200Runtime rt;
        return rt;
    }

    // keys as EvalKey holds them (what cufhe::Initialize(ek) takes, cufhe_gates_gpu.cu:42-47)
    void init(const TFHEpp::EvalKey& ek, int device = 0, size_t numSlots = size_t(1) << 16)
    {
        if (ctx_) shutdown();
        if (!ek.bklvl01 || !ek.iksklvl10) error::die("EvalKey lacks bklvl01 / iksklvl10");
        ck(b200fhe_create(&ctx_, device), "b200fhe_create");
        ck(b200fhe_load_keys(ctx_, reinterpret_cast<const uint32_t*>(ek.bklvl01.get()),
                             reinterpret_cast<const uint16_t*>(ek.iksklvl10.get())),
           "b200fhe_load_keys");
        ck(b200fhe_arena_alloc(ctx_, numSlots), "b200fhe_arena_alloc");
        capacity_ = numSlots;
        next_ = 0;
        free_.clear();
    }

    void shutdown()
    {
        if (ctx_) b200fhe_destroy(ctx_);
        ctx_ = nullptr;
        capacity_ = 0;
    }

    bool ready() const { return ctx_ != nullptr; }
    b200fhe_ctx* ctx() const { return ctx_; }
    size_t numBatches() const { return numBatches_; }
    size_t numGates() const { return numGates_; }
    size_t maxBatch() const { return maxBatch_; }

    uint32_t allocSlot()
    {
        if (!ctx_) error::die("B200Runtime::init must run before any network is built (slots are allocated by the Tasks)");
        if (!free_.empty()) {
            const uint32_t s = free_.back();
            free_.pop_back();
            return s;
        }
        if (next_ >= capacity_) error::die("B200Runtime: slot arena exhausted (", capacity_, " slots)");
        return next_++;
    }
    void freeSlot(uint32_t s)
    {
        if (ctx_) free_.push_back(s);
    }

    // ---- what the Tasks call ----
    void record(uint8_t op, uint32_t a, uint32_t b, uint32_t c, uint32_t out, bool* done)
    {
        op_.push_back(op);
        in0_.push_back(a);
        in1_.push_back(b);
        in2_.push_back(c);
        out_.push_back(out);
        done_.push_back(done);
    }
    void recordTick(uint32_t src, uint32_t dst)
    {
        tickSrc_.push_back(src);
        tickDst_.push_back(dst);
    }

    // ---- what the run loop calls ----
    void flushTicks()
    {
        if (tickSrc_.empty()) return;
        ck(b200fhe_dff_tick(ctx_, tickSrc_.data(), tickDst_.data(), tickSrc_.size()), "b200fhe_dff_tick");
        tickSrc_.clear();
        tickDst_.clear();
    }
    void flush()
    {
        if (op_.empty()) return;
        ck(b200fhe_gate_batch(ctx_, op_.data(), in0_.data(), in1_.data(), in2_.data(), out_.data(), op_.size()),
           "b200fhe_gate_batch");
        for (bool* d : done_) *d = true;
        numBatches_++;
        numGates_ += op_.size();
        maxBatch_ = std::max(maxBatch_, op_.size());
        op_.clear(), in0_.clear(), in1_.clear(), in2_.clear(), out_.clear(), done_.clear();
    }

    // ---- host access to single slots (test glue, front end) ----
    void upload(uint32_t slot, const TLWELvl0& c)
    {
        flushTicks();
        ck(b200fhe_upload(ctx_, &slot, reinterpret_cast<const uint16_t*>(c.data()), 1), "b200fhe_upload");
        ck(b200fhe_sync(ctx_), "b200fhe_sync");
    }
    TLWELvl0 download(uint32_t slot)
    {
        flushTicks();
        TLWELvl0 c;
        ck(b200fhe_download(ctx_, &slot, reinterpret_cast<uint16_t*>(c.data()), 1), "b200fhe_download");
        return c;
    }
    void copy(uint32_t src, uint32_t dst)
    {
        flushTicks();
        ck(b200fhe_dff_tick(ctx_, &src, &dst, 1), "b200fhe_dff_tick");
    }
};

// A Task's value: one slot of the arena.  Default construction allocates (Task's constructor does
// std::make_shared<OutType>(), src/iyokan.hpp:391-396); assignment copies the CIPHERTEXT on the device, which keeps
// the reference's generic `output() = input(0)` (TaskMem::set, TaskDFF::tick) meaningful for this type.
struct B200Slot {
    uint32_t id;
    B200Slot() : id(B200Runtime::instance().allocSlot()) {}
    B200Slot(const B200Slot& o) : id(B200Runtime::instance().allocSlot()) { B200Runtime::instance().copy(o.id, id); }
    B200Slot& operator=(const B200Slot& o)
    {
        if (id != o.id) B200Runtime::instance().copy(o.id, id);
        return *this;
    }
    ~B200Slot() { B200Runtime::instance().freeSlot(id); }
};

struct B200WorkerInfo {
    B200Runtime* rt;
};

using TaskB200Gate = Task<B200Slot, B200Slot, B200WorkerInfo>;
using TaskB200GateMem = TaskMem<B200Slot, B200Slot, B200WorkerInfo>;

inline TLWELvl0 b200TrivialTLWE(bool bit)  // (0,...,0, +-mu): HomCONSTANTONE / ZERO, TFHEpp gate.hpp:32-44
{
    TLWELvl0 t{};
    t[Lvl0::n] = bit ? Lvl0::μ : static_cast<Lvl0::T>(-Lvl0::μ);
    return t;
}

// DFF / RAM cell: Q <- D at tick (TaskDFF::tick, src/iyokan.hpp:1395-1402), initial value trivial 0
// (TaskTFHEppGateDFF, src/iyokan_tfhepp.hpp:17-49)
class TaskB200GateDFF : public TaskDFF<B200Slot, B200Slot, B200WorkerInfo> {
private:
    Bit initialValue_;

public:
    TaskB200GateDFF() : initialValue_(0_b) { setInitialValue(); }
    TaskB200GateDFF(Bit initValue) : initialValue_(initValue) { setInitialValue(); }

    void setInitialValue() { B200Runtime::instance().upload(output().id, b200TrivialTLWE(initialValue_ == 1_b)); }

    void tick() override
    {
        TaskMem<B200Slot, B200Slot, B200WorkerInfo>::tick();
        B200Runtime::instance().recordTick(input(0).id, output().id);  // all ticks of a clock: one b200fhe_dff_tick
    }
};

// INPUT / OUTPUT / ROM wires (TaskTFHEppGateWIRE, src/iyokan_tfhepp.hpp:59-107): a wire with an input copies it
class TaskB200GateWIRE : public TaskB200GateMem {
private:
    bool done_ = false;

    void startAsyncImpl(B200WorkerInfo wi, ProgressGraphMaker* graph) override
    {
        if (graph) graph->startNode(this->depnode()->label());
        if (getInputSize() == 1) {
            done_ = false;
            wi.rt->record(B200FHE_COPY, input(0).id, 0, 0, output().id, &done_);
        }
        else {
            assert(getInputSize() == 0);
        }
    }

public:
    TaskB200GateWIRE() {}
    TaskB200GateWIRE(bool inputNeeded) : TaskB200GateMem(inputNeeded ? 1 : 0) {}
    bool hasFinished() const override { return getInputSize() == 0 || done_; }
    void tick() override
    {
        TaskB200GateMem::tick();
        done_ = false;
    }
};

// Gate tasks: DEFINE_TASK_GATE of the reference (src/iyokan_tfhepp.hpp:109-144) with the TFHEpp call replaced by a
// record into the frontier.  ANDNOT = HomANDYN, ORNOT = HomORYN; MUX = HomMUX(out, in(2), in(1), in(0)) = in2 ? in1 : in0.
#define DEFINE_TASK_B200_GATE(name, numInputs, opcode)                                                     \
    class TaskB200Gate##name : public TaskB200Gate {                                                       \
    private:                                                                                               \
        bool done_ = false;                                                                                \
        void startAsyncImpl(B200WorkerInfo wi) override                                                    \
        {                                                                                                  \
            done_ = false;                                                                                 \
            wi.rt->record(opcode, (numInputs) > 0 ? input(0).id : 0, (numInputs) > 1 ? input(1).id : 0,   \
                          (numInputs) > 2 ? input(2).id : 0, output().id, &done_);                         \
        }                                                                                                  \
                                                                                                           \
    public:                                                                                                \
        TaskB200Gate##name() : TaskB200Gate(numInputs) {}                                                  \
        bool hasFinished() const override { return done_; }                                                \
        void tick() override                                                                               \
        {                                                                                                  \
            TaskB200Gate::tick();                                                                          \
            done_ = false;                                                                                 \
        }                                                                                                  \
    };
DEFINE_TASK_B200_GATE(AND, 2, B200FHE_AND)
DEFINE_TASK_B200_GATE(NAND, 2, B200FHE_NAND)
DEFINE_TASK_B200_GATE(ANDNOT, 2, B200FHE_ANDNOT)
DEFINE_TASK_B200_GATE(OR, 2, B200FHE_OR)
DEFINE_TASK_B200_GATE(NOR, 2, B200FHE_NOR)
DEFINE_TASK_B200_GATE(ORNOT, 2, B200FHE_ORNOT)
DEFINE_TASK_B200_GATE(XOR, 2, B200FHE_XOR)
DEFINE_TASK_B200_GATE(XNOR, 2, B200FHE_XNOR)
DEFINE_TASK_B200_GATE(MUX, 3, B200FHE_MUX)
DEFINE_TASK_B200_GATE(NOT, 1, B200FHE_NOT)
DEFINE_TASK_B200_GATE(CONSTONE, 0, B200FHE_CONST1)
DEFINE_TASK_B200_GATE(CONSTZERO, 0, B200FHE_CONST0)
#undef DEFINE_TASK_B200_GATE

class B200NetworkBuilder
    : public NetworkBuilder<TaskB200Gate, TaskB200GateMem, TaskB200GateDFF, TaskB200GateWIRE, B200WorkerInfo> {
private:
#define DEFINE_GATE_IMPL(name) \
    std::shared_ptr<TaskB200Gate> name##Impl() override { return std::make_shared<TaskB200Gate##name>(); }
    DEFINE_GATE_IMPL(AND);
    DEFINE_GATE_IMPL(NAND);
    DEFINE_GATE_IMPL(ANDNOT);
    DEFINE_GATE_IMPL(OR);
    DEFINE_GATE_IMPL(NOR);
    DEFINE_GATE_IMPL(ORNOT);
    DEFINE_GATE_IMPL(XOR);
    DEFINE_GATE_IMPL(XNOR);
    DEFINE_GATE_IMPL(MUX);
    DEFINE_GATE_IMPL(NOT);
    DEFINE_GATE_IMPL(CONSTONE);
    DEFINE_GATE_IMPL(CONSTZERO);
#undef DEFINE_GATE_IMPL
};

using B200Network = B200NetworkBuilder::NetworkType;

class B200Worker : public Worker<B200WorkerInfo> {
private:
    B200WorkerInfo wi_;
    B200WorkerInfo getWorkerInfo() override { return wi_; }

public:
    B200Worker(ReadyQueue<B200WorkerInfo>& readyQueue, size_t& numFinishedTargets, B200WorkerInfo wi,
               std::shared_ptr<ProgressGraphMaker> graph)
        : Worker(readyQueue, numFinishedTargets, graph), wi_(wi)
    {
    }
};

// Frontier-draining run loop.  One sweep = every Worker gets one Worker::update() turn: idle workers pop a ready node
// each and record it, busy workers whose batch has been submitted propagate to their dependents.  After the sweep the
// recorded frontier goes to the GPU as ONE batch.  `numWorkers` bounds the batch size, not the parallelism.
class B200NetworkRunner {
private:
    NetworkRunner<B200WorkerInfo, B200Worker> runner_;
    std::shared_ptr<ProgressGraphMaker> graph_;
    B200WorkerInfo wi_;

public:
    B200NetworkRunner(int numWorkers, B200WorkerInfo wi, std::shared_ptr<ProgressGraphMaker> graph = nullptr)
        : graph_(graph), wi_(wi)
    {
        for (int i = 0; i < numWorkers; i++) runner_.addWorker(wi, graph_);
    }

    void addNetwork(std::shared_ptr<B200Network> net) { runner_.addNetwork(net); }

    void run()
    {
        if (graph_) graph_->reset();
        wi_.rt->flushTicks();
        runner_.prepareToRun();
        while (runner_.getNumFinishedTargets() < runner_.numNodes()) {
            assert(runner_.isRunning() && "Detected infinite loop");
            runner_.update();
            wi_.rt->flush();
        }
    }

    void tick() { runner_.tick(); }

    void setSDFFInitialValue() { runner_.setSDFFInitialValue<TaskB200GateDFF>(); }
};

// the test hook every back-end provides (src/iyokan_tfhepp.cpp:577-601, src/iyokan_cufhe.cpp:854-878)
inline void processAllGates(B200Network& net, int numWorkers, B200WorkerInfo wi,
                            std::shared_ptr<ProgressGraphMaker> graph = nullptr)
{
    ReadyQueue<B200WorkerInfo> readyQueue;
    wi.rt->flushTicks();
    net.pushReadyTasks(readyQueue);

    size_t numFinishedTargets = 0;
    std::vector<B200Worker> workers;
    workers.reserve(numWorkers);
    for (int i = 0; i < numWorkers; i++) workers.emplace_back(readyQueue, numFinishedTargets, wi, graph);

    while (numFinishedTargets < net.numNodes()) {
        assert(std::any_of(workers.begin(), workers.end(), [](auto&& w) { return w.isWorking(); }) || !readyQueue.empty());
        for (auto&& w : workers) w.update();
        wi.rt->flush();
    }
    assert(readyQueue.empty());
}

#endif
