"""Key-switch time vs frontier width for ks_kernel (one CTA per gate) and ks8_kernel (eight gates per CTA, rows shared in L1)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import numpy as np
    import oracle as O
    from iyokan_b200 import Context, OPS
    keys = O.cached_keys(20261017)
    ctx = Context(0); ctx.load_keys(keys.bk, keys.ksk)
    N = 8192
    rng = np.random.default_rng(1)
    pa, pb = rng.integers(0, 2, N, dtype=np.uint8), rng.integers(0, 2, N, dtype=np.uint8)
    a, b = O.encrypt_bits(1, keys, pa), O.encrypt_bits(2, keys, pb)
    ctx.arena_alloc(3 * N); ids = np.arange(3 * N, dtype=np.uint32)
    ctx.upload(ids[:N], a); ctx.upload(ids[N:2 * N], b)
    ops = np.full(N, OPS["NAND"], np.uint8)
    row = {}
    for nb in (149, 296, 592, 1184, 2368, 4736, 8192):
        best = 1e9
        for rep in range(3):
            ctx.gate_batch(ops[:nb], ids[:nb], ids[N:N + nb], None, ids[2 * N:2 * N + nb]); ctx.sync()
            best = min(best, ctx.last_batch_ms()[1])
        row[nb] = round(best, 3)
    got = ctx.download(ids[2 * N:3 * N])
    ok = bool(np.array_equal(O.decrypt_bits(keys, got), 1 - (pa & pb)))
    pick = np.arange(0, N, 257)
    exact = bool(np.array_equal(got[pick], O.gate_batch(keys, ops[pick], a[pick], b[pick], nthreads=8)))
    print(json.dumps({"kernel": sys.argv[1], "ks_ms": row, "bits_ok": ok, "exact": exact}), flush=True)
else:
    for name, mn in (("ks8_kernel", "150"), ("ks_kernel", "100000000")):
        subprocess.run([sys.executable, __file__, name], env=dict(os.environ, B200FHE_KS8_MIN=mn))
