"""br8_kernel (quad-cluster latency shape) on the GPU: parity vs the oracle, then launch time vs number of jobs."""
import os, sys
# needs the experiment build: bash scripts/build_experiment_lib.sh (the shipped library does not contain this kernel)
os.environ.setdefault("B200FHE_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "iyokan_b200", "csrc", "libb200fhe_exp.so"))
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O
from iyokan_b200 import Context, OPS
keys = O.cached_keys(20261017)
ctx = Context(0); ctx.load_keys(keys.bk, keys.ksk)
N = 160
rng = np.random.default_rng(1)
pa, pb = rng.integers(0, 2, N, dtype=np.uint8), rng.integers(0, 2, N, dtype=np.uint8)
a, b = O.encrypt_bits(1, keys, pa), O.encrypt_bits(2, keys, pb)
ctx.arena_alloc(3 * N); ids = np.arange(3 * N, dtype=np.uint32)
ctx.upload(ids[:N], a); ctx.upload(ids[N:2 * N], b)
ops = np.full(N, OPS["NAND"], np.uint8)
for variant in (8, 6):
    ctx.set_kernel_variant(variant)
    row = {}
    for nb in (1, 8, 16, 30, 32, 34, 36, 37, 40, 74, 148):
        for rep in range(3):
            ctx.gate_batch(ops[:nb], ids[:nb], ids[N:N + nb], None, ids[2 * N:2 * N + nb]); ctx.sync()
        row[nb] = round(ctx.last_batch_ms()[0], 3)
    got = ctx.download(ids[2 * N:2 * N + 148])
    ok_bits = bool(np.array_equal(O.decrypt_bits(keys, got), 1 - (pa[:148] & pb[:148])))
    exact = bool(np.array_equal(got[:4], O.gate_batch(keys, ops[:4], a[:4], b[:4])))
    print(f"variant {variant}: bits_ok={ok_bits} exact={exact} ms={row}", flush=True)
c = np.zeros((3, 637), np.uint16)
c[0, :8] = [0xFFFF, 0xFFF0, 0, 15, 0x8000, 0x7FF0, 0x8010, 0x0010]; c[0, 636] = 0x001F; c[1, 600:636] = 0x8000; c[1, 636] = 0xFFFF; c[2, 636] = 1 << 13
ctx.set_kernel_variant(8)
print("edges exact:", bool(np.array_equal(ctx.test_bootstrap_lvl1(c), O.bootstrap_to_lvl1(keys, c))), flush=True)
