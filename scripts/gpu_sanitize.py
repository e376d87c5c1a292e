"""Tiny run of every shipped kernel shape for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200FHE_NO_CALIBRATE", "1")   # the calibration waves would run for minutes under the sanitizer
os.environ.setdefault("B200FHE_KS8_MIN", "150")       # eight-gates-per-CTA key switch on a small frontier
import oracle as O
from iyokan_b200 import Context, OPS
keys = O.cached_keys(20261017)
rng = np.random.default_rng(3)
n = int(os.environ.get("NJOBS", "3"))
c = O.encrypt_bits(1, keys, rng.integers(0, 2, n, dtype=np.uint8))
want = O.bootstrap_to_lvl1(keys, c)
with Context(0) as ctx:
    ctx.load_keys(keys.bk, keys.ksk)
    for variant, g in ((7, 8), (3, 4), (3, 6), (4, 1), (6, 1), (1, 4)):
        ctx.set_kernel_variant(variant)
        if g > 1:
            ctx.set_jobs_per_cta(g)
        got = ctx.test_bootstrap_lvl1(c)
        print(variant, g, "exact" if np.array_equal(got, want) else "MISMATCH", flush=True)
    ctx.set_kernel_variant(0)
    ctx.arena_alloc(4 * 300)
    ops = np.array([OPS["NAND"], OPS["MUX"], OPS["NOT"]], np.uint8)
    out = ctx.gates_host(ops, c[:3], c[:3], c[:3])      # ks_split path, unary kernels
    print("gates", out.shape, flush=True)
    # key switch shapes on lvl1 samples: 3 (split), 149 (one CTA per gate), 153 (eight gates per CTA, ragged last CTA)
    u = np.zeros((153, 1025), np.uint32)
    u[:] = rng.integers(0, 2**32, (153, 1025), dtype=np.uint64).astype(np.uint32)
    ks = ctx.test_keyswitch(u)
    same = np.array_equal(ctx.test_keyswitch(u[:149]), ks[:149]) and np.array_equal(ctx.test_keyswitch(u[:3]), ks[:3])
    print("keyswitch shapes agree" if same else "keyswitch MISMATCH", flush=True)
    # staged upload (>= 256 contiguous slots) + repack kernel, then a download
    big = np.tile(c[:1], (300, 1))
    ids = np.arange(300, dtype=np.uint32)
    ctx.upload(ids, big)
    print("upload exact" if np.array_equal(ctx.download(ids), big) else "upload MISMATCH", flush=True)
