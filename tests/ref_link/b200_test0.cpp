// TEST INFRASTRUCTURE: the reference's OWN templated unit tests (src/test0.cpp:43-455: testNOT, testMUX, testBinopGates,
// testFromJSONtest_{pass,and,and_4_2,mux,addr,register,counter}_4bit, testSequentialCircuit, testPrioritySetVisitor)
// instantiated with the B200 back-end plugin (iyokan_b200/host/iyokan_b200.hpp), exactly as src/test0.cpp:854-899
// instantiates them for the plain, TFHEpp and cuFHE back-ends.
//
// The reference's test0.cpp is #included as it lies in the reference tree (its main() renamed, nothing copied); this file
// adds the three glue overloads every back-end provides (src/test0.cpp:460-474, 581-598, 696-715) and a main().
// Scheduler, netlist readers, Task / DepNode / ReadyQueue / Worker: the reference's, unmodified.  Gate evaluation:
// b200fhe_gate_batch through the C ABI, one batch per ready frontier.  Inputs are FRESH encryptions under a key generated
// here with the reference's key generator (not trivial ciphertexts as in the cuFHE glue), outputs are decrypted with it.
//
// Built by `make -C oracle reflink`; run from a directory that holds test/iyokanl1-json/ (tests/golden/ref_assets) by
// tests/test_gpu_ref_link.py.  All checks are the reference's `assert`s: the build must not define NDEBUG.
#ifdef NDEBUG
#error "test0's checks are asserts: build without NDEBUG"
#endif

#define main test0_reference_main
#include "test0.cpp"
#undef main

#include "../../iyokan_b200/host/iyokan_b200.hpp"

namespace {
struct B200TestHelper {
    std::shared_ptr<SecretKey> sk = std::make_shared<SecretKey>();
    std::shared_ptr<EvalKey> ek = std::make_shared<EvalKey>();
    B200TestHelper()
    {
        ek->emplaceiksk<Lvl10>(*sk);  // the two members the gate path needs (src/iyokan-packet.cpp:144-160)
        ek->emplacebk<Lvl01>(*sk);
        B200Runtime::instance().init(*ek, 0, size_t(1) << 14);
    }
    static B200TestHelper& instance()
    {
        static B200TestHelper h;
        return h;
    }
};
}  // namespace

void processAllGates(B200Network& net, std::shared_ptr<ProgressGraphMaker> graph = nullptr)
{
    processAllGates(net, 256, B200WorkerInfo{&B200Runtime::instance()}, graph);
}

void setInput(std::shared_ptr<TaskB200GateMem> task, int val)
{
    auto& h = B200TestHelper::instance();
    B200Runtime::instance().upload(task->get().id, TFHEpp::bootsSymEncrypt<Lvl0>({static_cast<uint8_t>(val ? 1 : 0)}, *h.sk).at(0));
}

int getOutput(std::shared_ptr<TaskB200GateMem> task)
{
    auto& h = B200TestHelper::instance();
    return TFHEpp::bootsSymDecrypt<Lvl0>({B200Runtime::instance().download(task->get().id)}, *h.sk)[0];
}

#define RUN(test)                                    \
    do {                                             \
        test<B200NetworkBuilder>();                  \
        std::printf("ok  %s<B200NetworkBuilder>\n", #test); \
        std::fflush(stdout);                         \
    } while (0)

int main()
{
    AsyncThread::setNumThreads(2);
    B200TestHelper::instance();  // keys + GPU context before the first Task allocates a slot
    RUN(testNOT);
    RUN(testMUX);
    RUN(testBinopGates);
    RUN(testFromJSONtest_pass_4bit);
    RUN(testFromJSONtest_and_4bit);
    RUN(testFromJSONtest_and_4_2bit);
    RUN(testFromJSONtest_mux_4bit);
    RUN(testFromJSONtest_addr_4bit);
    RUN(testFromJSONtest_register_4bit);
    RUN(testSequentialCircuit);
    RUN(testFromJSONtest_counter_4bit);
    RUN(testPrioritySetVisitor);
    auto& rt = B200Runtime::instance();
    std::printf("test0 suite green on the B200 back-end: %zu gates in %zu batches (widest %zu)\n", rt.numGates(), rt.numBatches(),
                rt.maxBatch());
    rt.shutdown();
    return 0;
}
