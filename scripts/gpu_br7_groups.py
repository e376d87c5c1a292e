"""Ablation of br7_kernel's barrier groups: jobs per group J in {8, 4, 2} x start skew between groups.
Prints ms per launch for 1184 jobs (one wave) and 2368 jobs (two waves); feeds profiles/r02_br7_groups.md."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O
from iyokan_b200 import Context, OPS
keys = O.cached_keys(20261017)
N = 2368
rng = np.random.default_rng(1)
pa, pb = rng.integers(0, 2, N, dtype=np.uint8), rng.integers(0, 2, N, dtype=np.uint8)
a, b = O.encrypt_bits(1, keys, pa), O.encrypt_bits(2, keys, pb)
ops = np.full(N, OPS["NAND"], np.uint8)
ids = np.arange(3 * N, dtype=np.uint32)
out = {}
configs = [(8, 0)] + [(4, s) for s in (0, 15000, 29000)] + [(2, s) for s in (0, 7000, 14700)]
if os.environ.get("CONFIGS"):
    configs = [tuple(int(y) for y in x.split(":")) for x in os.environ["CONFIGS"].split(",")]
for J, skew in configs:
    os.environ["B200FHE_BR7_GROUP"], os.environ["B200FHE_BR7_SKEW"] = str(J), str(skew)
    with Context(0) as ctx:
        ctx.load_keys(keys.bk, keys.ksk)
        ctx.arena_alloc(3 * N)
        ctx.upload(ids[:N], a); ctx.upload(ids[N:2 * N], b)
        ctx.set_kernel_variant(7); ctx.set_jobs_per_cta(8)
        row = {}
        for nb in (1184, 2368):
            ts = []
            for rep in range(3):
                ctx.gate_batch(ops[:nb], ids[:nb], ids[N:N + nb], None, ids[2 * N:2 * N + nb]); ctx.sync()
                ts.append(ctx.last_batch_ms()[0])
            row[nb] = round(min(ts[1:]), 3)
        got = ctx.download(ids[2 * N:2 * N + 64])
        row["ok"] = bool(np.array_equal(O.decrypt_bits(keys, got), 1 - (pa[:64] & pb[:64])))
        out[f"J{J}_skew{skew}"] = row
        print(f"J={J} skew={skew}", row, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "br7_groups.json"), "w"), indent=1)
