"""Tiny run of every blind-rotation shape for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O
from iyokan_b200 import Context, OPS
keys = O.cached_keys(20261017)
rng = np.random.default_rng(3)
n = int(os.environ.get("NJOBS", "3"))
c = O.encrypt_bits(1, keys, rng.integers(0, 2, n, dtype=np.uint8))
want = O.bootstrap_to_lvl1(keys, c)
with Context(0) as ctx:
    ctx.load_keys(keys.bk, keys.ksk)
    for variant, g in ((3, 4), (3, 6), (4, 1), (5, 1), (6, 1)):
        ctx.set_kernel_variant(variant)
        ctx.set_jobs_per_cta(g)
        got = ctx.test_bootstrap_lvl1(c)
        print(variant, g, "exact" if np.array_equal(got, want) else "MISMATCH", flush=True)
    ctx.set_jobs_per_cta(0)
    ctx.arena_alloc(16)
    ops = np.array([OPS["NAND"], OPS["MUX"], OPS["NOT"]], np.uint8)
    out = ctx.gates_host(ops, c[:3], c[:3], c[:3])
    print("gates", out.shape, flush=True)
