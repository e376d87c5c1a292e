"""b200fhe_gates_host: one 8192-gate HomNAND batch with pinned host buffers against the device-resident batch, and the copies alone."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if True:
    sys.path.insert(0, ROOT)
    import numpy as np, time
    import oracle as O
    from iyokan_b200 import Context, OPS
    from iyokan_b200.lib import PinnedBuffer
    keys = O.cached_keys(20261017)
    ctx = Context(0); ctx.load_keys(keys.bk, keys.ksk)
    n = 8192
    rng = np.random.default_rng(1)
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    ha, hb, ho = (PinnedBuffer((n, 637), np.uint16) for _ in range(3))
    ha.array[:] = O.encrypt_bits(1, keys, pa); hb.array[:] = O.encrypt_bits(2, keys, pb)
    ops = np.full(n, OPS["NAND"], np.uint8)
    ctx.arena_alloc(4 * n)
    ids = np.arange(4 * n, dtype=np.uint32)
    for _ in range(2):
        ctx.gates_host(ops, ha.array, hb.array, None, out=ho.array)
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); ctx.gates_host(ops, ha.array, hb.array, None, out=ho.array); t.append(time.perf_counter() - t0)
    ok = bool(np.array_equal(O.decrypt_bits(keys, ho.array), 1 - (pa & pb)))
    r = []
    for _ in range(5):
        ctx.sync(); t0 = time.perf_counter(); ctx.gate_batch(ops, ids[:n], ids[n:2 * n], None, ids[3 * n:]); ctx.sync(); r.append(time.perf_counter() - t0)
    u = []
    for _ in range(3):
        ctx.sync(); t0 = time.perf_counter(); ctx.upload(ids[:n], ha.array); ctx.upload(ids[n:2 * n], hb.array); ctx.sync(); u.append(time.perf_counter() - t0)
    d = []
    for _ in range(3):
        ctx.sync(); t0 = time.perf_counter(); ctx.download(ids[3 * n:], out=ho.array); d.append(time.perf_counter() - t0)
    print(json.dumps({"host_ms": round(min(t) * 1e3, 3), "resident_ms": round(min(r) * 1e3, 3),
                      "upload_ms": round(min(u) * 1e3, 3), "download_ms": round(min(d) * 1e3, 3), "ok": ok}), flush=True)
