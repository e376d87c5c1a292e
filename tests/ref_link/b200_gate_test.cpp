// TEST INFRASTRUCTURE: the reference-side binding of include/b200fhe.h, compiled against the UNMODIFIED reference
// headers (TFHEpp, the library Iyokan's gate workers call) and linked with libb200fhe.so.
//
// It is the B200 counterpart of the reference's own GPU gate test (cuFHE test/test_gate_gpu.cc: random bits,
// every gate type through the batched API, decrypt, compare, print ms per gate) and of what
// `TaskTFHEppGate*::startSync` does per gate (src/iyokan_tfhepp.hpp:109-144): the key objects, ciphertext types,
// encryption and decryption are the reference's; only the gate evaluation goes through the C ABI.  What it proves:
//   * b200fhe_load_keys accepts `ek.bklvl01` / `ek.iksklvl10` exactly as TFHEpp holds them in memory
//     (BootstrappingKey<lvl01param>, KeySwitchingKey<lvl10param>, include/params.hpp:102-128);
//   * `TFHEpp::TLWE<lvl0param>` arrays are the host ciphertext format of b200fhe_gates_host / upload / download;
//   * results decrypt (TFHEpp::tlweSymDecrypt) to the same bits as TFHEpp::Hom* on the same inputs and keys.
// Built by `make -C oracle reflink` in the container that has the reference tree; the binary travels to the GPU
// box under oracle/_ref/ and is run there by tests/test_gpu_ref_link.py.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <tfhe++.hpp>

extern "C" {
#include "../../include/b200fhe.h"
}

using namespace TFHEpp;
using TLWE0 = TLWE<lvl0param>;
static_assert(sizeof(TLWE0) == B200FHE_TLWE0_LEN * sizeof(uint16_t), "TLWE<lvl0param> is the wire format of the C ABI");
static_assert(sizeof(BootstrappingKey<lvl01param>) == B200FHE_BK_WORDS * sizeof(uint32_t), "raw bootstrapping key layout");

#define CK(call)                                                                     \
    do {                                                                             \
        if ((call) != 0) {                                                           \
            std::fprintf(stderr, "%s failed: %s\n", #call, b200fhe_last_error());    \
            return 2;                                                                \
        }                                                                            \
    } while (0)

struct GateCase {
    const char* name;
    uint8_t op;
    int arity;
    bool (*plain)(bool, bool, bool);
    void (*ref)(TLWE0&, const TLWE0&, const TLWE0&, const TLWE0&, const EvalKey&);
};

int main(int argc, char** argv)
{
    const size_t n = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 64;   // gates per type
    const size_t nref = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 4; // of which also run through TFHEpp itself
    SecretKey sk;
    EvalKey ek;
    ek.emplacebk<lvl01param>(sk);     // what `iyokan-packet genevalkey` does (src/iyokan-packet.cpp:144-160)
    ek.emplaceiksk<lvl10param>(sk);
    ek.emplacebk2bkfft<lvl01param>(); // only for the TFHEpp side of the comparison

    b200fhe_ctx* ctx = nullptr;
    CK(b200fhe_create(&ctx, 0));
    CK(b200fhe_load_keys(ctx, reinterpret_cast<const uint32_t*>(ek.bklvl01.get()),
                         reinterpret_cast<const uint16_t*>(ek.iksklvl10.get())));
    CK(b200fhe_arena_alloc(ctx, 4 * n));

    const GateCase cases[] = {
        {"NAND", B200FHE_NAND, 2, [](bool a, bool b, bool) { return !(a && b); },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0&, const EvalKey& k) { HomNAND<lvl01param, lvl1param::μ, lvl10param>(r, a, b, k); }},
        {"AND", B200FHE_AND, 2, [](bool a, bool b, bool) { return a && b; },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0&, const EvalKey& k) { HomAND<lvl01param, lvl1param::μ, lvl10param>(r, a, b, k); }},
        {"ANDNOT", B200FHE_ANDNOT, 2, [](bool a, bool b, bool) { return a && !b; },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0&, const EvalKey& k) { HomANDYN<lvl01param, lvl1param::μ, lvl10param>(r, a, b, k); }},
        {"OR", B200FHE_OR, 2, [](bool a, bool b, bool) { return a || b; },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0&, const EvalKey& k) { HomOR<lvl01param, lvl1param::μ, lvl10param>(r, a, b, k); }},
        {"NOR", B200FHE_NOR, 2, [](bool a, bool b, bool) { return !(a || b); },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0&, const EvalKey& k) { HomNOR<lvl01param, lvl1param::μ, lvl10param>(r, a, b, k); }},
        {"ORNOT", B200FHE_ORNOT, 2, [](bool a, bool b, bool) { return a || !b; },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0&, const EvalKey& k) { HomORYN<lvl01param, lvl1param::μ, lvl10param>(r, a, b, k); }},
        {"XOR", B200FHE_XOR, 2, [](bool a, bool b, bool) { return a != b; },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0&, const EvalKey& k) { HomXOR<lvl01param, lvl1param::μ, lvl10param>(r, a, b, k); }},
        {"XNOR", B200FHE_XNOR, 2, [](bool a, bool b, bool) { return a == b; },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0&, const EvalKey& k) { HomXNOR<lvl01param, lvl1param::μ, lvl10param>(r, a, b, k); }},
        // Iyokan's MUX task: HomMUX(out, in(2) = S, in(1) = B, in(0) = A), i.e. S ? B : A (src/iyokan_tfhepp.hpp:140)
        {"MUX", B200FHE_MUX, 3, [](bool a, bool b, bool s) { return s ? b : a; },
         [](TLWE0& r, const TLWE0& a, const TLWE0& b, const TLWE0& s, const EvalKey& k) { HomMUX<lvl0param>(r, s, b, a, k); }},
        {"NOT", B200FHE_NOT, 1, [](bool a, bool, bool) { return !a; },
         [](TLWE0& r, const TLWE0& a, const TLWE0&, const TLWE0&, const EvalKey&) { HomNOT<lvl0param>(r, a); }},
    };

    std::mt19937 rng(20261017);
    std::vector<TLWE0> ca(n), cb(n), cc(n), out(n);
    std::vector<uint8_t> pa(n), pb(n), pc(n), ops(n);
    int failures = 0;
    for (const GateCase& g : cases) {
        for (size_t i = 0; i < n; i++) {
            pa[i] = rng() & 1, pb[i] = rng() & 1, pc[i] = rng() & 1;
            ca[i] = tlweSymEncrypt<lvl0param>(pa[i] ? lvl0param::μ : -lvl0param::μ, sk.key.lvl0);
            cb[i] = tlweSymEncrypt<lvl0param>(pb[i] ? lvl0param::μ : -lvl0param::μ, sk.key.lvl0);
            cc[i] = tlweSymEncrypt<lvl0param>(pc[i] ? lvl0param::μ : -lvl0param::μ, sk.key.lvl0);
            ops[i] = g.op;
        }
        const auto t0 = std::chrono::steady_clock::now();
        CK(b200fhe_gates_host(ctx, ops.data(), reinterpret_cast<const uint16_t*>(ca.data()),
                              g.arity > 1 ? reinterpret_cast<const uint16_t*>(cb.data()) : nullptr,
                              g.arity > 2 ? reinterpret_cast<const uint16_t*>(cc.data()) : nullptr,
                              reinterpret_cast<uint16_t*>(out.data()), n));
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        size_t bad = 0, bad_ref = 0;
        for (size_t i = 0; i < n; i++) {
            const bool got = tlweSymDecrypt<lvl0param>(out[i], sk.key.lvl0);
            if (got != g.plain(pa[i], pb[i], pc[i])) bad++;
            if (i < nref) {  // the reference's own evaluation of the same gate on the same ciphertexts
                TLWE0 r;
                g.ref(r, ca[i], cb[i], cc[i], ek);
                if (tlweSymDecrypt<lvl0param>(r, sk.key.lvl0) != got) bad_ref++;
            }
        }
        std::printf("%-7s %zu gates  %.3f ms/gate (host buffers)  wrong bits %zu  differ from TFHEpp %zu/%zu\n", g.name, n,
                    ms / n, bad, bad_ref, nref);
        failures += (int)(bad + bad_ref);
    }
    b200fhe_destroy(ctx);
    std::printf(failures ? "FAIL\n" : "PASS\n");
    return failures ? 1 : 0;
}
