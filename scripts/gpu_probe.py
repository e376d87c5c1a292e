"""First-contact GPU probe: stage-level parity + raw timings of the batched gate path."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from iyokan_b200 import Context, OPS  # noqa: E402

out = {}
keys = O.cached_keys(20261017)
rng = np.random.default_rng(3)

ctx = Context(0)
t = time.time()
ctx.load_keys(keys.bk, keys.ksk)
out["load_keys_s"] = time.time() - t

# --- stage parity: BK precompute vs CPU simulator
from iyokan_b200 import build as B
sim = ctypes.CDLL(str(B.build_sim()))
bk_sim = np.zeros((2, 6, 6, 1024), np.uint32)
sim.sim_bk_prepare(keys.bk.ctypes.data_as(ctypes.c_void_p), bk_sim.ctypes.data_as(ctypes.c_void_p), 2)
out["bk_ntt_equal_sim"] = bool(np.array_equal(ctx.test_read_bk_ntt(0, 2), bk_sim))

# --- stage parity: blind rotate and key switch vs oracle
n = 4
pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
ca, cb = O.encrypt_bits(1, keys, pa), O.encrypt_bits(2, keys, pb)
c = (-ca.astype(np.int32) - cb.astype(np.int32)).astype(np.uint16)
c[:, 636] += np.uint16(1 << 13)
u_ref = O.bootstrap_to_lvl1(keys, c)
for g in (1, 2, 3, 4, 6):
    ctx.set_jobs_per_cta(g)
    u = ctx.test_bootstrap_lvl1(c)
    out[f"br_exact_G{g}"] = bool(np.array_equal(u, u_ref))
    if not out[f"br_exact_G{g}"]:
        out[f"br_mismatch_G{g}"] = int((u != u_ref).sum())
out["ks_exact"] = bool(np.array_equal(ctx.test_keyswitch(u_ref), O.keyswitch(keys, u_ref)))
print(json.dumps(out), flush=True)

# --- timing: 8192 resident NAND gates
N = int(os.environ.get("PROBE_N", "8192"))
bits_a, bits_b = rng.integers(0, 2, N, dtype=np.uint8), rng.integers(0, 2, N, dtype=np.uint8)
A, Bc = O.encrypt_bits(5, keys, bits_a), O.encrypt_bits(6, keys, bits_b)
ctx.arena_alloc(3 * N)
ids = np.arange(3 * N, dtype=np.uint32)
ctx.upload(ids[:N], A)
ctx.upload(ids[N:2 * N], Bc)
ops = np.full(N, OPS["NAND"], np.uint8)
for g in (2, 3, 4, 6):
    ctx.set_jobs_per_cta(g)
    for rep in range(2):
        t = time.time()
        ctx.gate_batch(ops, ids[:N], ids[N:2 * N], None, ids[2 * N:])
        ctx.sync()
        wall = time.time() - t
        br, ks = ctx.last_batch_ms()
        out[f"G{g}_rep{rep}"] = {"wall_s": wall, "br_ms": br, "ks_ms": ks, "boot_per_s": N / wall}
        print(json.dumps({f"G{g}_rep{rep}": out[f"G{g}_rep{rep}"]}), flush=True)
res = ctx.download(ids[2 * N:])
bits = O.decrypt_bits(keys, res)
out["nand_bits_ok"] = bool(np.array_equal(bits, 1 - (bits_a & bits_b)))
# exact check on a sample against the oracle
k = 8
want = O.gate_batch(keys, ops[:k], A[:k], Bc[:k])
out["nand_exact_sample"] = bool(np.array_equal(res[:k], want))
# small batches (latency)
for nb in (1, 32, 148, 296, 592):
    ctx.set_jobs_per_cta(2)
    t = time.time()
    ctx.gate_batch(ops[:nb], ids[:nb], ids[N:N + nb], None, ids[2 * N:2 * N + nb])
    ctx.sync()
    br, ks = ctx.last_batch_ms()
    out[f"small_{nb}"] = {"wall_s": time.time() - t, "br_ms": br, "ks_ms": ks}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
ctx.close()
