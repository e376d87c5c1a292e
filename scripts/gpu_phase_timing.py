"""Per-phase cycle breakdown of br6_kernel (latency shape) and br7_kernel<8,8> (throughput shape) on the GPU.

Needs the debug build: scripts/build_timing_lib.sh -> iyokan_b200/csrc/libb200fhe_timing.so, loaded through B200FHE_LIB.
Prints cycles per CMUX step and phase for the first and the last warp of CTA 0."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200FHE_LIB", os.path.join(ROOT, "iyokan_b200", "csrc", "libb200fhe_timing.so"))
import oracle as O
from iyokan_b200 import Context, OPS
from iyokan_b200 import lib as L

raw = ctypes.CDLL(os.environ["B200FHE_LIB"])
keys = O.cached_keys(20261017)
ctx = Context(0); ctx.load_keys(keys.bk, keys.ksk)
N = 2368
rng = np.random.default_rng(1)
pa, pb = rng.integers(0, 2, N, dtype=np.uint8), rng.integers(0, 2, N, dtype=np.uint8)
a, b = O.encrypt_bits(1, keys, pa), O.encrypt_bits(2, keys, pb)
ctx.arena_alloc(3 * N); ids = np.arange(3 * N, dtype=np.uint32)
ctx.upload(ids[:N], a); ctx.upload(ids[N:2 * N], b)
ops = np.full(N, OPS["NAND"], np.uint8)
NAMES = {
    6: ["fwd_p1", "bar", "fwd_p2", "bar", "fwd_p3", "bar", "fwd_p4", "fence+cluster_wait", "team bar", "copy issue+key wait",
        "cta bar", "pw_local", "wait peer tiles", "pw_finish", "arrive+cta bar", "inv_pA", "bar", "inv_pB", "bar", "inv_pC",
        "bar", "inv_pD", "cta bar"],
    7: ["fwd0_a+b (rotate, digit 0, pass 1)", "fwd0_c (pass 2)", "fwd12_a (digits 1-2, pass 1)", "fwd12_c (pass 2)",
        "key prefetch + cta bar", "pointwise", "cta bar", "inv01_a", "inv01_b", "inv2_a", "inv2_b (+acc)"],
}
out = {}
CASES = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("PT_CASES", "6:74,7:1184").split(",")]
TAG = os.environ.get("PT_TAG", "")
for variant, nb in CASES:
    ctx.set_kernel_variant(variant)
    for rep in range(2):
        ctx.gate_batch(ops[:nb], ids[:nb], ids[N:N + nb], None, ids[2 * N:2 * N + nb]); ctx.sync()
    buf = (ctypes.c_ulonglong * 64)()
    raw.b200fhe_debug_phase_cycles(buf, 1)
    ctx.gate_batch(ops[:nb], ids[:nb], ids[N:N + nb], None, ids[2 * N:2 * N + nb]); ctx.sync()
    ms = ctx.last_batch_ms()[0]
    raw.b200fhe_debug_phase_cycles(buf, 1)
    rawc = np.array(list(buf), dtype=np.float64).reshape(2, 32)
    print(f"{TAG} variant {variant}: loop of the first warp = {rawc[0, 30]:.0f} cycles in {rawc[0, 31]:.0f} ns "
          f"-> {rawc[0, 30] / max(rawc[0, 31], 1) * 1e3:.0f} MHz")
    if variant == 7:
        cb = (ctypes.c_ulonglong * 4096)()
        raw.b200fhe_debug_cta_ns(cb)
        ct = np.array(list(cb), dtype=np.float64).reshape(1024, 4)[: (nb + 7) // 8]
        t0 = ct[:, 0].min()
        for k, nm in enumerate(("kernel entry", "loop start", "loop end", "exit")):
            print(f"{TAG}   CTA {nm:12s}: min {(ct[:, k].min() - t0) / 1e6:8.3f} ms  max {(ct[:, k].max() - t0) / 1e6:8.3f} ms")
    cyc = rawc / 636.0
    cyc[:, 30:] = 0
    got = ctx.download(ids[2 * N:2 * N + nb])
    assert np.array_equal(O.decrypt_bits(keys, got), 1 - (pa[:nb] & pb[:nb])), "wrong bits"
    names = NAMES[variant]
    print(f"{TAG} variant {variant}, {nb} jobs: {ms:.3f} ms per launch (instrumented build); cycles per CMUX step")
    for k, nm in enumerate(names):
        print(f"  {k:2d} {nm:40s} first warp {cyc[0, k]:8.0f}   last warp {cyc[1, k]:8.0f}")
    print(f"     {'total':40s} first warp {cyc[0].sum():8.0f}   last warp {cyc[1].sum():8.0f}", flush=True)
    out[str(variant)] = {"jobs": nb, "ms": ms, "phases": names, "first_warp": cyc[0, :len(names)].tolist(),
                         "last_warp": cyc[1, :len(names)].tolist()}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"phase_timing{TAG}.json"), "w"), indent=1)
