// TEST INFRASTRUCTURE — not part of the product path.
//
// Thin command-line driver around the UNMODIFIED reference TFHEpp gate path
// (vendored under /root/reference/thirdparty/cuFHE/thirdparties/TFHEpp).
// It is compiled by oracle/Makefile straight from the reference's own
// sources where they lie; only this driver file is ours.  The binary lands
// in oracle/_ref/ (git-ignored) and is used
//   * by tests/ to pin the exact-integer oracle (oracle/tfhe_oracle.c) and the
//     CUDA path against TFHEpp's decrypted bits on the same keys and inputs,
//   * by bench.py --impl reference / cpu_baseline as the CPU baseline.
//
// Reference entry points exercised (the calls Iyokan's TFHEpp worker makes,
// src/iyokan_tfhepp.hpp:131-144):
//   TFHEpp::HomNAND/AND/OR/...<lvl01param, lvl1param::mu, lvl10param>
//       (TFHEpp include/gate.hpp:59-230)
//   TFHEpp::HomMUX<lvl0param>, HomNOT<lvl0param>, HomCONSTANTONE/ZERO
//       (gate.hpp:32-57, 231-262)
//   tlweSymEncrypt / tlweSymDecrypt<lvl0param> (tlwe.hpp:12-25, 77-87)
//   bkgen<lvl01param>, ikskgen<lvl10param>     (cloudkey.hpp:18-52, 188-203)
//
// File formats (all little-endian flat arrays, shared with oracle/ and tests/):
//   sk0.bin  u16[636]                lvl0 secret key (binary)
//   sk1.bin  u32[1024]               lvl1 secret key (ternary, as two's compl.)
//   bk.bin   u32[636][6][2][1024]    raw TRGSW bootstrapping key (bklvl01)
//   ksk.bin  u16[1024][7][3][637]    identity key-switching key (iksklvl10)
//   tlwe arrays: u16[count][637]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <tfhe++.hpp>

using namespace TFHEpp;
using TLWE0 = TLWE<lvl0param>;
using T0 = lvl0param::T;                       // uint16_t (128-bit set), uint32_t (80-bit set, -DUSE_80BIT_SECURITY)
constexpr size_t TLWE0_LEN = lvl0param::n + 1;

#ifdef USE_80BIT_SECURITY
static_assert(sizeof(TLWE0) == 501 * 4, "80-bit parameter build expected (params/CGGI16.hpp)");
static_assert(sizeof(BootstrappingKey<lvl01param>) == 500ull * 4 * 2 * 1024 * 4);
static_assert(sizeof(KeySwitchingKey<lvl10param>) == 1024ull * 8 * 3 * 501 * 4);
#else
static_assert(sizeof(TLWE0) == 637 * 2, "128-bit parameter build expected");
static_assert(sizeof(BootstrappingKey<lvl01param>) == 636ull * 6 * 2 * 1024 * 4);
static_assert(sizeof(KeySwitchingKey<lvl10param>) == 1024ull * 7 * 3 * 637 * 2);
#endif

// Opcode numbering shared with include/b200fhe.h.
enum Op : uint8_t {
    OP_AND = 0, OP_NAND, OP_ANDNOT, OP_OR, OP_NOR, OP_ORNOT, OP_XOR, OP_XNOR,
    OP_MUX, OP_NOT, OP_COPY, OP_CONST0, OP_CONST1, OP_ANDNY, OP_ORNY
};

static void die(const char* msg)
{
    std::fprintf(stderr, "ref_driver: %s\n", msg);
    std::exit(2);
}

template <class T>
static std::vector<T> slurp(const std::string& path)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) die(("cannot open " + path).c_str());
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<T> v(sz / sizeof(T));
    if (sz && std::fread(v.data(), 1, sz, f) != (size_t)sz) die("short read");
    std::fclose(f);
    return v;
}

static void spit(const std::string& path, const void* p, size_t bytes)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) die(("cannot write " + path).c_str());
    if (bytes && std::fwrite(p, 1, bytes, f) != bytes) die("short write");
    std::fclose(f);
}

static Key<lvl0param> load_sk0(const std::string& dir)
{
    auto v = slurp<T0>(dir + "/sk0.bin");
    if (v.size() != lvl0param::n) die("sk0.bin size");
    Key<lvl0param> k;
    std::memcpy(k.data(), v.data(), sizeof(k));
    return k;
}

static void load_evalkey(const std::string& dir, EvalKey& ek)
{
    ek.bklvl01 = std::make_unique_for_overwrite<BootstrappingKey<lvl01param>>();
    ek.iksklvl10 = std::make_unique_for_overwrite<KeySwitchingKey<lvl10param>>();
    {
        auto v = slurp<uint32_t>(dir + "/bk.bin");
        if (v.size() * 4 != sizeof(BootstrappingKey<lvl01param>)) die("bk.bin size");
        std::memcpy(ek.bklvl01.get(), v.data(), v.size() * 4);
    }
    {
        auto v = slurp<T0>(dir + "/ksk.bin");
        if (v.size() * sizeof(T0) != sizeof(KeySwitchingKey<lvl10param>)) die("ksk.bin size");
        std::memcpy(ek.iksklvl10.get(), v.data(), v.size() * sizeof(T0));
    }
    ek.emplacebk2bkfft<lvl01param>();  // FFT form derived from the same raw key
}

static int cmd_keygen(const std::string& dir)
{
    SecretKey sk;  // std::random_device-seeded Randen (utils.hpp:15-20)
    auto bk = std::make_unique_for_overwrite<BootstrappingKey<lvl01param>>();
    auto ksk = std::make_unique_for_overwrite<KeySwitchingKey<lvl10param>>();
    bkgen<lvl01param>(*bk, sk);
    ikskgen<lvl10param>(*ksk, sk);
    spit(dir + "/sk0.bin", sk.key.lvl0.data(), sizeof(sk.key.lvl0));
    spit(dir + "/sk1.bin", sk.key.lvl1.data(), sizeof(sk.key.lvl1));
    spit(dir + "/bk.bin", bk.get(), sizeof(*bk));
    spit(dir + "/ksk.bin", ksk.get(), sizeof(*ksk));
    return 0;
}

static int cmd_encrypt(const std::string& dir, const std::string& bits, const std::string& out)
{
    auto key = load_sk0(dir);
    auto p = slurp<uint8_t>(bits);
    std::vector<TLWE0> c(p.size());
    for (size_t i = 0; i < p.size(); i++)
        c[i] = tlweSymEncrypt<lvl0param>(p[i] ? lvl0param::μ : -lvl0param::μ, key);
    spit(out, c.data(), c.size() * sizeof(TLWE0));
    return 0;
}

static int cmd_decrypt(const std::string& dir, const std::string& in, const std::string& out)
{
    auto key = load_sk0(dir);
    auto raw = slurp<T0>(in);
    size_t n = raw.size() / 637;
    std::vector<uint8_t> p(n);
    for (size_t i = 0; i < n; i++) {
        TLWE0 c;
        std::memcpy(c.data(), raw.data() + i * TLWE0_LEN, sizeof(c));
        p[i] = tlweSymDecrypt<lvl0param>(c, key);
    }
    spit(out, p.data(), n);
    return 0;
}

#ifdef USE_80BIT_SECURITY
// The pinned TFHEpp's br->iks HomGate overload (gate.hpp:8-18) does not compile with a 32-bit lvl0 torus: it takes the
// offset as a signed template argument and gate.hpp:82,99,116,167,184 pass -mu, which narrows.
// The 80-bit driver therefore forms the linear combination itself - the three lines of HomGate, with the
// (casign, cbsign, offset) table of gate.hpp:59-230 - and calls the reference's own GateBootstrapping
// (gatebootstrapping.hpp:241-250: BlindRotate + SampleExtractIndex + IdentityKeySwitch), which is everything else.
static void hom_gate80(TLWE0& r, const TLWE0& a, const TLWE0& b, int sa, int sb, int off, const EvalKey& ek)
{
    TLWE0 c;
    for (size_t i = 0; i <= lvl0param::n; i++) c[i] = (T0)((uint32_t)sa * a[i] + (uint32_t)sb * b[i]);
    c[lvl0param::n] += (T0)((uint32_t)off * lvl0param::μ);
    GateBootstrapping<lvl01param, lvl1param::μ, lvl10param>(r, c, ek);
}
static void run_gate(uint8_t op, TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0& c,
                     const EvalKey& ek)
{
    switch (op) {
    case OP_NAND:   hom_gate80(r, a, b, -1, -1, +1, ek); break;  // gate.hpp:65
    case OP_NOR:    hom_gate80(r, a, b, -1, -1, -1, ek); break;  // :82
    case OP_XNOR:   hom_gate80(r, a, b, -2, -2, -2, ek); break;  // :99
    case OP_AND:    hom_gate80(r, a, b, +1, +1, -1, ek); break;  // :116
    case OP_OR:     hom_gate80(r, a, b, +1, +1, +1, ek); break;  // :133
    case OP_XOR:    hom_gate80(r, a, b, +2, +2, +2, ek); break;  // :150
    case OP_ANDNY:  hom_gate80(r, a, b, -1, +1, -1, ek); break;  // :167
    case OP_ANDNOT: hom_gate80(r, a, b, +1, -1, -1, ek); break;  // :184 HomANDYN
    case OP_ORNY:   hom_gate80(r, a, b, -1, +1, +1, ek); break;  // :201
    case OP_ORNOT:  hom_gate80(r, a, b, +1, -1, +1, ek); break;  // :218 HomORYN
    case OP_MUX: {  // HomMUX<lvl0param>, gate.hpp:231-262 (lvl0 branch), cs = c, c1 = b, c0 = a
        TLWE0 t1, t0;
        for (size_t i = 0; i <= lvl0param::n; i++) {
            t1[i] = c[i] + b[i];
            t0[i] = (T0)(0u - c[i]) + a[i];
        }
        t1[lvl0param::n] -= lvl0param::μ;
        t0[lvl0param::n] -= lvl0param::μ;
        TLWE<lvl1param> u1, u0;
        GateBootstrappingTLWE2TLWEFFT<lvl01param>(u1, t1, *ek.bkfftlvl01, μpolygen<lvl1param, lvl1param::μ>());
        GateBootstrappingTLWE2TLWEFFT<lvl01param>(u0, t0, *ek.bkfftlvl01, μpolygen<lvl1param, lvl1param::μ>());
        for (size_t i = 0; i <= lvl1param::n; i++) u1[i] += u0[i];
        IdentityKeySwitch<lvl10param>(r, u1, *ek.iksklvl10);
        r[lvl0param::n] += lvl0param::μ;
        break;
    }
    case OP_NOT:    for (size_t i = 0; i <= lvl0param::n; i++) r[i] = (T0)(0u - a[i]); break;  // HomNOT, gate.hpp:47-51
    case OP_COPY:   r = a; break;
    case OP_CONST0: r = {}; r[lvl0param::n] = (T0)(0u - lvl0param::μ); break;                  // gate.hpp:32-44
    case OP_CONST1: r = {}; r[lvl0param::n] = lvl0param::μ; break;
    default: die("bad opcode");
    }
}
#else
static void run_gate(uint8_t op, TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0& c,
                     const EvalKey& ek)
{
    constexpr auto mu = lvl1param::μ;
    using br = lvl01param;
    using ks = lvl10param;
    switch (op) {
    case OP_AND:    HomAND<br, mu, ks>(r, a, b, ek); break;
    case OP_NAND:   HomNAND<br, mu, ks>(r, a, b, ek); break;
    case OP_ANDNOT: HomANDYN<br, mu, ks>(r, a, b, ek); break;   // iyokan_tfhepp.hpp:133
    case OP_OR:     HomOR<br, mu, ks>(r, a, b, ek); break;
    case OP_NOR:    HomNOR<br, mu, ks>(r, a, b, ek); break;
    case OP_ORNOT:  HomORYN<br, mu, ks>(r, a, b, ek); break;    // iyokan_tfhepp.hpp:136
    case OP_XOR:    HomXOR<br, mu, ks>(r, a, b, ek); break;
    case OP_XNOR:   HomXNOR<br, mu, ks>(r, a, b, ek); break;
    case OP_ANDNY:  HomANDNY<br, mu, ks>(r, a, b, ek); break;
    case OP_ORNY:   HomORNY<br, mu, ks>(r, a, b, ek); break;
    // Iyokan: HomMUX(out, in(2)=S, in(1)=B, in(0)=A); here a=in0, b=in1, c=in2.
    case OP_MUX:    HomMUX<lvl0param>(r, c, b, a, ek); break;
    case OP_NOT:    HomNOT<lvl0param>(r, a); break;
    case OP_COPY:   HomCOPY<lvl0param>(r, a); break;
    case OP_CONST0: HomCONSTANTZERO<lvl0param>(r); break;
    case OP_CONST1: HomCONSTANTONE<lvl0param>(r); break;
    default: die("bad opcode");
    }
}

#endif

// gates DIR ops.bin in0.bin in1.bin in2.bin out.bin nthreads [repeat]
// in*.bin are u16[count][637]; pass "-" for an unused operand.
static int cmd_gates(int argc, char** argv)
{
    if (argc < 9) die("usage: gates DIR ops in0 in1 in2 out nthreads [repeat]");
    std::string dir = argv[2];
    auto ops = slurp<uint8_t>(argv[3]);
    size_t n = ops.size();
    auto load = [&](const char* p) {
        std::vector<TLWE0> v(n);
        if (std::strcmp(p, "-") != 0) {
            auto raw = slurp<T0>(p);
            if (raw.size() != n * TLWE0_LEN) die("operand size mismatch");
            std::memcpy(v.data(), raw.data(), n * sizeof(TLWE0));
        }
        return v;
    };
    auto a = load(argv[4]), b = load(argv[5]), c = load(argv[6]);
    std::vector<TLWE0> out(n);
    int nthreads = std::atoi(argv[8]);
    int repeat = argc > 9 ? std::atoi(argv[9]) : 1;
    if (nthreads <= 0) nthreads = std::thread::hardware_concurrency();
    EvalKey ek;
    load_evalkey(dir, ek);

    auto t0 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < repeat; rep++) {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([&, t] {
                for (size_t i = t; i < n; i += nthreads)
                    run_gate(ops[i], out[i], a[i], b[i], c[i], ek);
            });
        for (auto& x : th) x.join();
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    spit(argv[7], out.data(), n * sizeof(TLWE0));
    size_t boots = 0;
    for (auto op : ops) boots += (op == OP_MUX) ? 2 : (op <= OP_XNOR || op >= OP_ANDNY) ? 1 : 0;
    std::printf("{\"gates\": %zu, \"bootstraps\": %zu, \"repeat\": %d, \"threads\": %d, "
                "\"seconds\": %.6f, \"bootstraps_per_s\": %.3f}\n",
                n, boots, repeat, nthreads, sec, boots * (double)repeat / sec);
    return 0;
}


// ---- primitive-level commands used by tests/golden/make_golden.py ----
// blindrotate DIR in(u16[c][637]) out(u32[c][1025]): GateBootstrappingTLWE2TLWEFFT (gatebootstrapping.hpp:188-197)
static int cmd_blindrotate(const std::string& dir, const std::string& in, const std::string& out)
{
    EvalKey ek;
    load_evalkey(dir, ek);
    auto raw = slurp<T0>(in);
    size_t n = raw.size() / TLWE0_LEN;
    std::vector<TLWE<lvl1param>> res(n);
    for (size_t i = 0; i < n; i++) {
        TLWE0 c;
        std::memcpy(c.data(), raw.data() + i * TLWE0_LEN, sizeof(c));
        GateBootstrappingTLWE2TLWEFFT<lvl01param>(res[i], c, *ek.bkfftlvl01,
                                                  μpolygen<lvl1param, lvl1param::μ>());
    }
    spit(out, res.data(), n * sizeof(TLWE<lvl1param>));
    return 0;
}

// keyswitch DIR in(u32[c][1025]) out(u16[c][637]): IdentityKeySwitch<lvl10param> (keyswitch.hpp:11-52)
static int cmd_keyswitch(const std::string& dir, const std::string& in, const std::string& out)
{
    auto ksk = std::make_unique_for_overwrite<KeySwitchingKey<lvl10param>>();
    auto v = slurp<T0>(dir + "/ksk.bin");
    if (v.size() * sizeof(T0) != sizeof(*ksk)) die("ksk.bin size");
    std::memcpy(ksk.get(), v.data(), v.size() * sizeof(T0));
    auto raw = slurp<uint32_t>(in);
    size_t n = raw.size() / 1025;
    std::vector<TLWE0> res(n);
    for (size_t i = 0; i < n; i++) {
        TLWE<lvl1param> c;
        std::memcpy(c.data(), raw.data() + i * 1025, sizeof(c));
        IdentityKeySwitch<lvl10param>(res[i], c, *ksk);
    }
    spit(out, res.data(), n * sizeof(TLWE0));
    return 0;
}

// decompose in(u32[c][1024]) out(i32[c][3][1024]): Decomposition<lvl1param> (trgsw.hpp:62-78)
static int cmd_decompose(const std::string& in, const std::string& out)
{
    auto raw = slurp<uint32_t>(in);
    size_t n = raw.size() / 1024;
    std::vector<DecomposedPolynomial<lvl1param>> res(n);
    for (size_t i = 0; i < n; i++) {
        Polynomial<lvl1param> p;
        std::memcpy(p.data(), raw.data() + i * 1024, sizeof(p));
        Decomposition<lvl1param>(res[i], p);
    }
    spit(out, res.data(), n * sizeof(res[0]));
    return 0;
}

// mulxai in(u32[c][1024]) a(u32[c]) out(u32[c][2][1024]): [0]=PolynomialMulByXai, [1]=...MinusOne (utils.hpp:113-144)
static int cmd_mulxai(const std::string& in, const std::string& as, const std::string& out)
{
    auto raw = slurp<uint32_t>(in);
    auto a = slurp<uint32_t>(as);
    size_t n = a.size();
    std::vector<std::array<Polynomial<lvl1param>, 2>> res(n);
    for (size_t i = 0; i < n; i++) {
        Polynomial<lvl1param> p;
        std::memcpy(p.data(), raw.data() + i * 1024, sizeof(p));
        PolynomialMulByXai<lvl1param>(res[i][0], p, a[i]);
        PolynomialMulByXaiMinusOne<lvl1param>(res[i][1], p, a[i]);
    }
    spit(out, res.data(), n * sizeof(res[0]));
    return 0;
}

// cmuxstep trgsw(u32[6][2][1024]) acc(u32[2][1024]) abar out(u32[2][1024]):
// CMUXFFTwithPolynomialMulByXaiMinusOne<lvl01param> (detwfa.hpp:36-49)
static int cmd_cmuxstep(const std::string& tg, const std::string& accin, uint32_t abar,
                        const std::string& out)
{
    auto t = slurp<uint32_t>(tg);
    auto a = slurp<uint32_t>(accin);
    if (t.size() != 2 * lvl1param::l * 2 * 1024 || a.size() != 2 * 1024) die("cmuxstep sizes");
    TRGSW<lvl1param> trgsw;
    std::memcpy(trgsw.data(), t.data(), sizeof(trgsw));
    BootstrappingKeyElementFFT<lvl01param> el;
    el[0] = ApplyFFT2trgsw<lvl1param>(trgsw);
    TRLWE<lvl1param> acc;
    std::memcpy(acc.data(), a.data(), sizeof(acc));
    CMUXFFTwithPolynomialMulByXaiMinusOne<lvl01param>(acc, el, abar);
    spit(out, acc.data(), sizeof(acc));
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 2) die("usage: ref_driver keygen|encrypt|decrypt|gates|blindrotate|keyswitch|decompose|mulxai|cmuxstep ...");
    std::string cmd = argv[1];
    if (cmd == "keygen" && argc == 3) return cmd_keygen(argv[2]);
    if (cmd == "encrypt" && argc == 5) return cmd_encrypt(argv[2], argv[3], argv[4]);
    if (cmd == "decrypt" && argc == 5) return cmd_decrypt(argv[2], argv[3], argv[4]);
    if (cmd == "gates") return cmd_gates(argc, argv);
    if (cmd == "blindrotate" && argc == 5) return cmd_blindrotate(argv[2], argv[3], argv[4]);
    if (cmd == "keyswitch" && argc == 5) return cmd_keyswitch(argv[2], argv[3], argv[4]);
    if (cmd == "decompose" && argc == 4) return cmd_decompose(argv[2], argv[3]);
    if (cmd == "mulxai" && argc == 5) return cmd_mulxai(argv[2], argv[3], argv[4]);
    if (cmd == "cmuxstep" && argc == 6)
        return cmd_cmuxstep(argv[2], argv[3], (uint32_t)std::strtoul(argv[4], nullptr, 10), argv[5]);
    die("bad command line");
    return 2;
}
