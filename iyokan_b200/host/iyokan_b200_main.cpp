// iyokan-b200: `iyokan tfhe` with the gate workers replaced by the B200 back-end.
//
// This translation unit is compiled AGAINST THE UNMODIFIED REFERENCE (it includes its src/iyokan.hpp and
// src/packet.hpp and links its TFHEpp objects; `make -C oracle iyokan-b200`, in the container that holds the
// reference tree).  Everything that is not the hot path is the reference's own code, used as it is:
//   NetworkBlueprint                        TOML blueprint parser                      (src/iyokan.hpp:1671-1945)
//   readNetwork / YosysJSONReader / IyokanL1JSONReader                                 (:2064-2515)
//   makeROMWithMUX / makeRAMWithMUX         builtin MUX memories (incl. the embedded mux-ram netlists, :2517-2762)
//   TFHEPacket, readFromArchive / writeToArchive, TFHEpp::EvalKey   packet and key I/O (src/packet.hpp:208-340)
// What is replaced is the back-end: instead of instantiating Task<>/Worker<> templates per gate
// (src/iyokan_tfhepp.hpp:109-192) the readers are pointed at `FlatBuilder`, a NetworkBuilder that records the
// DAG as flat arrays; the sub-networks are merged along the blueprint's [connect] edges exactly as
// TFHEppFrontend's constructor does (src/iyokan_tfhepp.cpp:312-458) and handed to the level-replay engine
// (include/b200net.h -> include/b200fhe.h), which evaluates one frontier of independent gates per CUDA launch
// plan.  The cycle protocol is TFHEppFrontend::go (src/iyokan_tfhepp.cpp:465-566).
//
// Usage (option names of src/main.cpp:100-175):
//   iyokan-b200 tfhe --blueprint B.toml --evalkey EK -i req.enc -o res.enc -c N [--skip-reset] [--quiet]
//                    [--snapshot S] [--dump-prefix P --secret-key SK]
//                    [--dump-time-csv-prefix P] [--dump-graph-json-prefix P] [--dump-graph-dot-prefix P]
//   iyokan-b200 tfhe --resume S --evalkey EK -o res.enc -c N [...]
// Multi-GPU (what `--num-gpu N` asks of the reference, src/iyokan_cufhe.cpp:533): one process per GPU, launched N
// times with RANK / LOCAL_RANK / WORLD_SIZE in the environment (e.g. by torchrun or scripts/launch_ranks.sh); rank 0
// publishes the communicator id through the file named by B200FHE_ID_FILE (default /tmp/b200fhe_id_<MASTER_PORT>).
// Every rank evaluates its share of the static schedule (b200net_bind_rank); rank 0 writes the result packet.
// Snapshot / resume (src/main.cpp:116,160-166; TFHEppFrontend::serialize, iyokan_tfhepp.cpp:568-572) store what this
// front end needs to continue: blueprint path, cycle counter, the request packet and the ciphertext of every node.
// The per-cycle dumps (--dump-time-csv-prefix / --dump-graph-json-prefix / --dump-graph-dot-prefix, written by
// ProgressGraphMaker in the reference, src/iyokan.hpp:128-278, iyokan_tfhepp.cpp:538-555) keep the reference's file
// formats; a node's start / end are the measured window of the schedule step it was evaluated in (b200net_profile_run),
// since a frontier, not a gate, is the unit of execution here.  The plaintext mode lives in the Python front end.
#include <chrono>
#include <filesystem>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "iyokan.hpp"
#include "packet.hpp"

extern "C" {
#include "../../include/b200fhe.h"
#include "../../include/b200net.h"
}

// ---- a NetworkBuilder (duck-typed to what the reference's readers call) that records flat arrays ----
// (global namespace, like the reference's builders: makeROMWithMUX / makeRAMWithMUX find their helpers by ADL)
struct FlatNet;
class FlatBuilder {
    friend struct FlatNet;

protected:
    std::vector<uint8_t> kind_;
    std::vector<std::array<int32_t, 3>> in_;
    std::vector<uint8_t> nin_;
    std::map<TaskLabel, int> named_;  // ("input"|"output"|"rom"|"ram", port name, bit) -> node id

    int add(uint8_t kind)
    {
        kind_.push_back(kind);
        in_.push_back({-1, -1, -1});
        nin_.push_back(0);
        return (int)kind_.size() - 1;
    }
    int addNamed(uint8_t kind, const char* label, const std::string& port, int bit)
    {
        const int id = add(kind);
        if (!named_.emplace(TaskLabel{label, port, bit}, id).second)
            error::die("duplicate port ", label, " ", port, "[", bit, "]");
        return id;
    }
    // what RAMNetworkBuilder<NetworkBuilder>::RAM expects from its base (src/iyokan.hpp:1285-1300)
    struct IdHandle {
        int id;
        const IdHandle* operator->() const { return this; }
        const IdHandle* depnode() const { return this; }
        IdHandle label() const { return *this; }
    };
    IdHandle addNamedDFF(const std::string& label, const std::string& port, int bit)
    {
        return IdHandle{addNamed(B200NET_DFF, label.c_str(), port, bit)};
    }

public:
    using NetworkType = FlatNet;
    int AND() { return add(B200FHE_AND); }
    int NAND() { return add(B200FHE_NAND); }
    int ANDNOT() { return add(B200FHE_ANDNOT); }
    int OR() { return add(B200FHE_OR); }
    int NOR() { return add(B200FHE_NOR); }
    int ORNOT() { return add(B200FHE_ORNOT); }
    int XOR() { return add(B200FHE_XOR); }
    int XNOR() { return add(B200FHE_XNOR); }
    int MUX() { return add(B200FHE_MUX); }
    int NOT() { return add(B200FHE_NOT); }
    int CONSTONE() { return add(B200FHE_CONST1); }
    int CONSTZERO() { return add(B200FHE_CONST0); }
    int DFF() { return add(B200NET_DFF); }
    int ROM(const std::string& port, int bit) { return addNamed(B200NET_INPUT, "rom", port, bit); }
    int INPUT(const std::string& port, int bit) { return addNamed(B200NET_INPUT, "input", port, bit); }
    int OUTPUT(const std::string& port, int bit) { return addNamed(B200NET_OUTPUT, "output", port, bit); }
    void connect(int from, int to)
    {
        if (nin_.at(to) >= 3) error::die("node ", to, " has too many inputs");
        in_[to][nin_[to]++] = from;
    }
};

struct FlatNet {
    std::vector<uint8_t> kind;
    std::vector<std::array<int32_t, 3>> in;
    std::map<TaskLabel, int> named;
    FlatNet(FlatBuilder&& b) : kind(std::move(b.kind_)), in(std::move(b.in_)), named(std::move(b.named_)) {}
    void checkValid(error::Stack&) const {}  // arities are validated by b200net_create on the merged design
};

namespace {

// ---- TLWE helpers: trivial ciphertexts (0,...,0, +-mu), iyokan_tfhepp.hpp:23-27 ----
TLWELvl0 trivial(bool bit)
{
    TLWELvl0 t{};
    t[TFHEpp::lvl0param::n] = bit ? TFHEpp::lvl0param::μ : static_cast<TFHEpp::lvl0param::T>(-TFHEpp::lvl0param::μ);
    return t;
}
static_assert(sizeof(TLWELvl0) == B200FHE_TLWE0_LEN * sizeof(uint16_t), "TLWELvl0 is the wire format of the C ABI");

void ck(int rc, const char* what)
{
    if (rc) error::die("b200: ", what, ": ", b200net_last_error(), " / ", b200fhe_last_error());
}

struct Design {
    std::vector<uint8_t> kind;
    std::vector<int32_t> in0, in1, in2;
    std::map<std::string, std::pair<int, const FlatNet*>> nets;  // name -> (offset into the merged arrays, net)
    int node(const blueprint::Port& p) const
    {
        auto it = nets.find(p.nodeName);
        if (it == nets.end()) error::die("Invalid network name: ", p.nodeName);
        auto jt = it->second.second->named.find(p.portLabel);
        if (jt == it->second.second->named.end())
            error::die("Invalid port: ", p.nodeName, "/", p.portLabel.portName, "[", p.portLabel.portBit, "] (", p.portLabel.kind, ")");
        return it->second.first + jt->second;
    }
    int node(const std::string& net, const char* label, const std::string& port, int bit) const
    {
        return node(blueprint::Port{net, TaskLabel{label, port, bit}});
    }
};

// what --snapshot stores and --resume reads (cereal portable binary, like every other file of the tool chain)
struct B200Snapshot {
    std::string blueprintPath;
    int currentCycle = 0;
    TFHEPacket req;
    uint64_t numNodes = 0;
    std::vector<TLWELvl0> values;  // ciphertext of every node, in node order
    template <class Archive>
    void serialize(Archive& ar)
    {
        ar(blueprintPath, currentCycle, req, numNodes, values);
    }
};

const char* kindName(uint8_t k)
{
    switch (k) {
    case B200FHE_AND: return "AND";
    case B200FHE_NAND: return "NAND";
    case B200FHE_ANDNOT: return "ANDNOT";
    case B200FHE_OR: return "OR";
    case B200FHE_NOR: return "NOR";
    case B200FHE_ORNOT: return "ORNOT";
    case B200FHE_XOR: return "XOR";
    case B200FHE_XNOR: return "XNOR";
    case B200FHE_MUX: return "MUX";
    case B200FHE_NOT: return "NOT";
    case B200FHE_CONST1: return "CONSTONE";
    case B200FHE_CONST0: return "CONSTZERO";
    case B200NET_DFF: return "DFF";
    case B200NET_INPUT: return "INPUT";
    case B200NET_OUTPUT: return "WIRE";
    default: return "?";
    }
}

}  // namespace

int main(int argc, char** argv)
{
    std::string blueprintPath, evalkeyPath, inPath, outPath, snapshotPath, resumePath, dumpPrefix, secretKeyPath;
    std::string timeCsvPrefix, graphJsonPrefix, graphDotPrefix;
    int numCycles = -1;
    bool skipReset = false, quiet = false, statsJson = false;
    if (argc < 2 || std::string(argv[1]) != "tfhe")
        error::die("usage: iyokan-b200 tfhe (--blueprint B -i IN | --resume S) --evalkey EK -o OUT -c N");
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) error::die("missing value for ", a);
            return argv[++i];
        };
        if (a == "--blueprint") blueprintPath = next();
        else if (a == "--evalkey") evalkeyPath = next();
        else if (a == "-i" || a == "--in") inPath = next();
        else if (a == "-o" || a == "--out") outPath = next();
        else if (a == "-c") numCycles = std::stoi(next());
        else if (a == "--snapshot") snapshotPath = next();
        else if (a == "--resume") resumePath = next();
        else if (a == "--dump-prefix") dumpPrefix = next();
        else if (a == "--secret-key") secretKeyPath = next();
        else if (a == "--dump-time-csv-prefix") timeCsvPrefix = next();
        else if (a == "--dump-graph-json-prefix") graphJsonPrefix = next();
        else if (a == "--dump-graph-dot-prefix") graphDotPrefix = next();
        else if (a == "--skip-reset") skipReset = true;
        else if (a == "--quiet") quiet = true;
        else if (a == "--stats-json") statsJson = true;  // one JSON line on stdout: clock time, schedule, launches (bench.py)
        else if (a == "--verbose" || a == "--enable-gpu" || a == "--show-combinational-progress") {}
        else if (a == "--cpu" || a == "--gpu" || a == "--num-gpu" || a == "--gpu_num" || a == "--sched") next();  // CPU scheduler knobs
        else error::die("unknown option ", a);
    }
    if (quiet) spdlog::set_level(spdlog::level::err);
    if (evalkeyPath.empty() || outPath.empty()) error::die("--evalkey and -o are required");
    B200Snapshot snap;
    const bool resumed = !resumePath.empty();
    if (resumed) {
        if (!blueprintPath.empty() || !inPath.empty()) error::die("--resume excludes --blueprint and -i");
        readFromArchive(snap, resumePath);
        blueprintPath = snap.blueprintPath;
    } else if (blueprintPath.empty() || inPath.empty()) {
        error::die("--blueprint and -i are required for a new run");
    }
    if (!dumpPrefix.empty() && secretKeyPath.empty()) error::die("--dump-prefix needs --secret-key");
    const bool profiling = !timeCsvPrefix.empty() || !graphJsonPrefix.empty() || !graphDotPrefix.empty();

    // ---- the reference's loader: blueprint, netlists, builtin MUX memories ----
    const NetworkBlueprint bp{blueprintPath};
    if (bp.needsCircuitKey())  // type = "rom" / "ram": same ports, same function, evaluated here as MUX memories
        spdlog::warn("blueprint declares CMUX memories: evaluating them as MUX memories (no circuit bootstrapping on this back-end)");
    std::map<std::string, std::shared_ptr<FlatNet>> name2net;
    for (const auto& file : bp.files()) name2net.emplace(file.name, readNetwork<FlatBuilder>(file));
    for (const auto& rom : bp.builtinROMs()) name2net.emplace(rom.name, makeROMWithMUX<FlatBuilder>(rom.inAddrWidth, rom.outRdataWidth));
    for (const auto& ram : bp.builtinRAMs()) {
        if (ram.inWdataWidth != ram.outRdataWidth) error::die("Invalid RAM size; wdata and rdata widths differ");
        name2net.emplace(ram.name, makeRAMWithMUX<FlatBuilder>(ram.inAddrWidth, ram.outRdataWidth));
    }

    // ---- merge the sub-networks into one DAG (what connectTasks does across networks, iyokan_tfhepp.cpp:428-435) ----
    Design d;
    for (auto& [name, net] : name2net) {
        const int off = (int)d.kind.size();
        d.nets.emplace(name, std::make_pair(off, net.get()));
        for (size_t i = 0; i < net->kind.size(); i++) {
            d.kind.push_back(net->kind[i]);
            auto rel = [&](int v) { return v < 0 ? -1 : v + off; };
            d.in0.push_back(rel(net->in[i][0]));
            d.in1.push_back(rel(net->in[i][1]));
            d.in2.push_back(rel(net->in[i][2]));
        }
    }
    for (auto&& [key, port] : bp.atPorts()) d.node(port);  // existence check, as the reference does first
    for (const auto& [src, dst] : bp.edges()) {            // the consumer's INPUT wire becomes an alias of the producer
        const int s = d.node(src), t = d.node(dst);
        if (d.kind[t] != B200NET_INPUT) error::die("port connected twice: ", dst.nodeName, "/", dst.portLabel.portName);
        d.kind[t] = B200NET_OUTPUT;
        d.in0[t] = s;
    }

    // ---- keys and request, read by the reference's own (cereal) code ----
    const auto ek = readFromArchive<TFHEpp::EvalKey>(evalkeyPath);
    if (!ek.bklvl01 || !ek.iksklvl10) error::die("EvalKey lacks bklvl01 / iksklvl10 (run iyokan-packet genevalkey)");
    const TFHEPacket req = resumed ? snap.req : readFromArchive<TFHEPacket>(inPath);
    const int startCycle = resumed ? snap.currentCycle : 0;
    if (numCycles < 0) numCycles = req.numCycles.value_or(-1);
    if (numCycles < 0) error::die("the number of cycles is given neither by -c nor by the request packet");

    auto envInt = [](const char* name, int dflt) {
        const char* v = std::getenv(name);
        return v ? std::atoi(v) : dflt;
    };
    const int rank = envInt("RANK", 0), world = envInt("WORLD_SIZE", 1), localRank = envInt("LOCAL_RANK", rank);
    b200fhe_ctx* ctx = nullptr;
    if (b200fhe_create(&ctx, localRank)) error::die("b200fhe_create: ", b200fhe_last_error());
    if (b200fhe_load_keys(ctx, reinterpret_cast<const uint32_t*>(ek.bklvl01.get()),
                          reinterpret_cast<const uint16_t*>(ek.iksklvl10.get())))
        error::die("b200fhe_load_keys: ", b200fhe_last_error());
    if (world > 1) {  // communicator id: rank 0 writes it to a file, the others wait for it
        const char* idFile = std::getenv("B200FHE_ID_FILE");
        const std::string path = idFile ? idFile : "/tmp/b200fhe_id_" + std::string(std::getenv("MASTER_PORT") ? std::getenv("MASTER_PORT") : "0");
        uint8_t id[128];
        if (rank == 0) {
            if (b200fhe_comm_unique_id(id)) error::die("b200fhe_comm_unique_id: ", b200fhe_last_error());
            std::ofstream(path + ".tmp", std::ios::binary).write(reinterpret_cast<const char*>(id), 128);
            std::rename((path + ".tmp").c_str(), path.c_str());
        } else {
            for (int tries = 0;; tries++) {
                std::ifstream f(path, std::ios::binary);
                if (f && f.read(reinterpret_cast<char*>(id), 128)) break;
                if (tries > 6000) error::die("timed out waiting for the communicator id in ", path);
                std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
        }
        if (b200fhe_comm_init(ctx, rank, world, id)) error::die("b200fhe_comm_init: ", b200fhe_last_error());
        if (rank == 0) std::remove(path.c_str());  // every rank has joined once comm_init returns
    }
    b200net* net = nullptr;
    ck(b200net_create(&net, d.kind.size(), d.kind.data(), d.in0.data(), d.in1.data(), d.in2.data()), "b200net_create");
    ck(b200net_bind_rank(net, ctx, rank, world, B200NET_PACK), "b200net_bind_rank");
    {
        size_t steps = 0, coll = 0, slots = 0;
        double model = 0;
        int packed = 0;
        b200net_schedule_info(net, &steps, &coll, &slots, &model, &packed);
        spdlog::info("rank {}/{}: {} steps ({}), {} all-gathers and {} KiB exchanged per clock, modelled {:.1f} ms per clock", rank,
                     world, steps, packed ? "slack-packed" : "ASAP levels", coll, slots * 1280 / 1024, model);
    }
    spdlog::info("{} nodes, {} levels, {} bootstraps per cycle, {} DFF", b200net_num_nodes(net), b200net_num_levels(net),
                 b200net_bootstraps_per_cycle(net), b200net_num_dff(net));

    auto set = [&](const std::vector<uint32_t>& nodes, const std::vector<TLWELvl0>& vals) {
        if (nodes.empty()) return;
        ck(b200net_set(net, nodes.data(), reinterpret_cast<const uint16_t*>(vals.data()), nodes.size()), "b200net_set");
    };
    {  // DFF / RAM cells start at trivial 0
        std::vector<uint32_t> nodes;
        for (size_t i = 0; i < d.kind.size(); i++)
            if (d.kind[i] == B200NET_DFF) nodes.push_back((uint32_t)i);
        set(nodes, std::vector<TLWELvl0>(nodes.size(), trivial(false)));
    }
    // memories named in the request: MUX ROM contents now, MUX RAM contents on the first cycle
    auto memNodes = [&](const std::string& name, const char* label, const char* port, size_t n) {
        std::vector<uint32_t> nodes;
        for (size_t i = 0; i < n; i++) nodes.push_back((uint32_t)d.node(name, label, port, (int)i));
        return nodes;
    };
    for (const auto& rom : bp.builtinROMs()) {
        auto it = req.romInTLWE.find(rom.name);
        if (it == req.romInTLWE.end()) continue;
        const size_t n = (size_t(1) << rom.inAddrWidth) * rom.outRdataWidth;
        if (it->second.size() != n) error::die("Invalid request packet: wrong length of ROM ", rom.name);
        set(memNodes(rom.name, "rom", "romdata", n), it->second);
    }
    // external ports
    const bool hasReset = bp.at("reset").has_value();
    auto resetNode = [&] { return std::vector<uint32_t>{(uint32_t)d.node(*bp.at("reset"))}; };
    for (const auto& [name, stream] : req.bits) {
        if (name == "reset") error::die("@reset cannot be set by the request");
        if (!bp.atPortWidths().count(name)) error::die("Invalid request packet: unknown port @", name);
    }

    // ---- result packet (makeResPacket, iyokan_tfhepp.cpp:176-227) ----
    auto get = [&](const std::vector<uint32_t>& nodes) {
        std::vector<TLWELvl0> vals(nodes.size());
        if (!nodes.empty()) ck(b200net_get(net, nodes.data(), reinterpret_cast<uint16_t*>(vals.data()), nodes.size()), "b200net_get");
        return vals;
    };
    auto makeRes = [&](int cyclesDone) {
        TFHEPacket res{{}, {}, {}, {}, {}, cyclesDone};
        for (const auto& [name, width] : bp.atPortWidths()) {
            // one entry per port bit, written at its own index (makeResPacket resizes to atPortBit + 1,
            // iyokan_tfhepp.cpp:182-190); bits tied to ground or left unconnected stay trivial 0
            std::vector<uint32_t> nodes;
            std::vector<int> bits;
            for (int b = 0; b < width; b++)
                if (const auto port = bp.at(name, b); port && port->portLabel.kind == "output") {
                    nodes.push_back((uint32_t)d.node(*port));
                    bits.push_back(b);
                }
            if (bits.empty()) continue;
            std::vector<TLWELvl0> vals(bits.back() + 1, trivial(false));
            const auto got = get(nodes);
            for (size_t k = 0; k < bits.size(); k++) vals[bits[k]] = got[k];
            res.bits.emplace(name, std::move(vals));
        }
        for (const auto& ram : bp.builtinRAMs()) {
            const size_t n = (size_t(1) << ram.inAddrWidth) * ram.outRdataWidth;
            auto nodes = memNodes(ram.name, "ram", "ramdata", n);
            // A CMUX RAM is updated during the cycle (its image already holds the last cycle's write); the MUX RAM that
            // stands in for it keeps that value on the cells' D inputs until the next tick: report those.
            if (ram.type == blueprint::BuiltinRAM::TYPE::CMUX_MEMORY && cyclesDone > 0)
                for (auto& v : nodes) v = (uint32_t)d.in0[v];
            res.ramInTLWE.emplace(ram.name, get(nodes));
        }
        return res;
    };

    // ---- cycle protocol (TFHEppFrontend::go) ----
    std::vector<uint32_t> allNodes(d.kind.size());
    for (size_t i = 0; i < allNodes.size(); i++) allNodes[i] = (uint32_t)i;
    if (resumed) {
        if (snap.numNodes != d.kind.size() || snap.values.size() != d.kind.size())
            error::die("snapshot does not belong to this blueprint (", snap.numNodes, " nodes, the blueprint has ", d.kind.size(), ")");
        std::vector<uint32_t> nodes;  // wires alias their driver and hold no state of their own
        std::vector<TLWELvl0> vals;
        for (size_t i = 0; i < d.kind.size(); i++)
            if (d.kind[i] != B200NET_OUTPUT) nodes.push_back((uint32_t)i), vals.push_back(snap.values[i]);
        ck(b200net_restore(net, nodes.data(), reinterpret_cast<const uint16_t*>(vals.data()), nodes.size()), "b200net_restore");
    }
    std::optional<TFHEpp::SecretKey> secretKey;
    if (!dumpPrefix.empty()) secretKey = readFromArchive<TFHEpp::SecretKey>(secretKeyPath);
    // per-node times of a profiled cycle: the window of the schedule step that evaluated the node
    const size_t numSteps = b200net_num_steps(net);
    std::vector<int> stepOfNode(d.kind.size(), -1);
    if (profiling) {
        std::vector<uint32_t> buf(d.kind.size());
        for (size_t st = 0; st < numSteps; st++)
            for (int r = 0; r < world; r++) {
                const size_t n = b200net_step_gates(net, st, r, buf.data(), buf.size());
                for (size_t k = 0; k < n; k++) stepOfNode[buf[k]] = (int)st;
            }
    }
    auto dumpProfile = [&](int cycle, std::chrono::system_clock::time_point t0, const std::vector<float>& stepMs) {
        using namespace utility;
        using namespace std::chrono;
        std::vector<system_clock::time_point> begin(numSteps + 1, t0);
        for (size_t st = 0; st < numSteps; st++)
            begin[st + 1] = begin[st] + duration_cast<system_clock::duration>(duration<double, std::milli>(stepMs[st]));
        // nodes in the order they start: inputs, registers and wires at the beginning of the cycle, then step by step
        std::vector<uint32_t> order(allNodes);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return stepOfNode[x] < stepOfNode[y]; });
        std::vector<int> index(d.kind.size());
        for (size_t k = 0; k < order.size(); k++) index[order[k]] = (int)k;
        auto startOf = [&](uint32_t v) { return stepOfNode[v] < 0 ? t0 : begin[stepOfNode[v]]; };
        auto endOf = [&](uint32_t v) { return stepOfNode[v] < 0 ? t0 : begin[stepOfNode[v] + 1]; };
        auto desc = [&](uint32_t v) { return stepOfNode[v] < 0 ? std::string() : fok("step ", stepOfNode[v]); };
        if (!timeCsvPrefix.empty()) {  // ProgressGraphMaker::dumpTimeCSV
            auto os = openOfstream(fok(timeCsvPrefix, "-", cycle, ".csv"));
            for (uint32_t v : order)
                *os << "\"" << startOf(v) << "\",\"" << endOf(v) << "\",\"" << index[v] << "\",\"" << v << "\",\"" << kindName(d.kind[v])
                    << "\",\"" << desc(v) << "\"" << std::endl;
        }
        const std::vector<int32_t>* ins[3] = {&d.in0, &d.in1, &d.in2};
        if (!graphJsonPrefix.empty()) {  // ProgressGraphMaker::dumpJSON
            picojson::object nodes;
            for (uint32_t v : order) {
                picojson::object j;
                j.emplace("start", fok(startOf(v)));
                j.emplace("end", fok(endOf(v)));
                j.emplace("index", (double)index[v]);
                j.emplace("id", (double)v);
                j.emplace("kind", std::string(kindName(d.kind[v])));
                j.emplace("desc", desc(v));
                nodes.emplace(fok(v), j);
            }
            picojson::array edges;
            int ne = 0;
            for (uint32_t v : allNodes)
                for (auto in : ins)
                    if ((*in)[v] >= 0) {
                        picojson::object j;
                        j.emplace("index", (double)ne++);
                        j.emplace("from", (double)(*in)[v]);
                        j.emplace("to", (double)v);
                        edges.emplace_back(j);
                    }
            picojson::object root;
            root.emplace("nodes", nodes);
            root.emplace("edges", edges);
            *openOfstream(fok(graphJsonPrefix, "-", cycle, ".json")) << picojson::value(root);
        }
        if (!graphDotPrefix.empty()) {  // ProgressGraphMaker::dumpDOT
            auto os = openOfstream(fok(graphDotPrefix, "-", cycle, ".dot"));
            *os << "digraph progress_graph_maker {" << std::endl << "node [ shape = record ];" << std::endl;
            for (uint32_t v : order) {
                *os << "n" << v << " [label = \"{" << kindName(d.kind[v]);
                if (stepOfNode[v] >= 0) *os << "|" << desc(v);
                *os << "}\"];" << std::endl;
            }
            *os << std::endl;
            int ne = 0;
            for (uint32_t v : allNodes)
                for (auto in : ins)
                    if ((*in)[v] >= 0) *os << "n" << (*in)[v] << " -> n" << v << " [label = \"" << ne++ << "\"];" << std::endl;
            *os << "}" << std::endl;
        }
    };
    if (!resumed && hasReset && !skipReset) {
        set(resetNode(), {trivial(true)});
        ck(b200net_run(net), "b200net_run");
    }
    if (b200fhe_sync(ctx)) error::die("b200fhe_sync: ", b200fhe_last_error());
    // the clock cycles are timed on their own (the reference prints one "done. (N us)" per cycle, iyokan_tfhepp.cpp:557)
    const uint64_t launches0 = b200fhe_launch_count(ctx);
    const auto t0 = std::chrono::steady_clock::now();
    for (int c = 0; c < numCycles; c++) {
        const int cur = startCycle + c;  // cycles count on across --snapshot / --resume (currentCycle_)
        if (secretKey && rank == 0)      // dumpDecryptedPacket (iyokan_tfhepp.cpp:298-305): the state BEFORE cycle `cur`
            writeToArchive(utility::fok(dumpPrefix, "-", cur), makeRes(cur).decrypt(*secretKey));
        ck(b200net_tick(net), "b200net_tick");
        if (c == 0 && hasReset) set(resetNode(), {trivial(false)});
        if (cur == 0) {
            for (const auto& ram : bp.builtinRAMs()) {
                auto it = req.ramInTLWE.find(ram.name);
                if (it == req.ramInTLWE.end()) continue;
                const size_t n = (size_t(1) << ram.inAddrWidth) * ram.outRdataWidth;
                if (it->second.size() != n) error::die("Invalid request packet: wrong length of RAM ", ram.name);
                set(memNodes(ram.name, "ram", "ramdata", n), it->second);
            }
        }
        for (const auto& [name, stream] : req.bits) {  // setCircularInputs (iyokan_tfhepp.cpp:274-296)
            const int width = bp.atPortWidths().at(name);
            std::vector<uint32_t> nodes;
            std::vector<TLWELvl0> vals;
            for (int b = 0; b < width; b++) {
                const auto port = bp.at(name, b);
                if (!port || port->portLabel.kind != "input") continue;
                nodes.push_back((uint32_t)d.node(*port));
                vals.push_back(stream.at((size_t(width) * cur + b) % stream.size()));
            }
            set(nodes, vals);
        }
        if (profiling) {
            std::vector<float> stepMs(numSteps, 0.0f);
            if (b200fhe_sync(ctx)) error::die("b200fhe_sync: ", b200fhe_last_error());
            const auto tc = std::chrono::system_clock::now();
            ck(b200net_profile_run(net, stepMs.data(), stepMs.size()), "b200net_profile_run");
            if (rank == 0) dumpProfile(cur, tc, stepMs);
        } else {
            ck(b200net_run(net), "b200net_run");
        }
    }
    if (b200fhe_sync(ctx)) error::die("b200fhe_sync: ", b200fhe_last_error());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    spdlog::info("done. ({} us for {} cycle(s), {:.0f} bootstraps/s)", (long long)(secs * 1e6), numCycles,
                 secs > 0 ? b200net_bootstraps_per_cycle(net) * (double)numCycles / secs : 0.0);
    if (statsJson) {
        size_t steps = 0, coll = 0, slots = 0;
        double model = 0;
        int packed = 0;
        b200net_schedule_info(net, &steps, &coll, &slots, &model, &packed);
        std::printf("{\"rank\": %d, \"world\": %d, \"cycles\": %d, \"seconds\": %.6f, \"bootstraps_per_cycle\": %zu, \"nodes\": %zu, "
                    "\"levels\": %zu, \"steps\": %zu, \"packed\": %s, \"collectives_per_cycle\": %zu, \"exchanged_bytes_per_cycle\": %zu, "
                    "\"model_s_per_cycle\": %.6f, \"gpu_launches_per_cycle\": %.1f}\n",
                    rank, world, numCycles, secs, b200net_bootstraps_per_cycle(net), b200net_num_nodes(net), b200net_num_levels(net), steps,
                    packed ? "true" : "false", coll, slots * 1280, model / 1e3,
                    numCycles > 0 ? (double)(b200fhe_launch_count(ctx) - launches0) / numCycles : 0.0);
        std::fflush(stdout);
    }

    const TFHEPacket res = makeRes(startCycle + numCycles);
    if (rank == 0) writeToArchive(outPath, res);
    if (!snapshotPath.empty() && rank == 0) {
        B200Snapshot out;
        out.blueprintPath = std::filesystem::absolute(blueprintPath).string();
        out.currentCycle = startCycle + numCycles;
        out.req = req;
        out.numNodes = d.kind.size();
        out.values = get(allNodes);
        writeToArchive(snapshotPath, out);
    }
    b200net_destroy(net);
    b200fhe_destroy(ctx);
    return 0;
}
