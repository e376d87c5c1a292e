"""Blind-rotation kernel time vs batch size for each jobs-per-CTA setting (feeds the host heuristic)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O
from iyokan_b200 import Context, OPS
keys = O.cached_keys(20261017)
ctx = Context(0); ctx.load_keys(keys.bk, keys.ksk)
N = 2368
rng = np.random.default_rng(1)
a = O.encrypt_bits(1, keys, rng.integers(0, 2, N, dtype=np.uint8)); b = O.encrypt_bits(2, keys, rng.integers(0, 2, N, dtype=np.uint8))
ctx.arena_alloc(3 * N); ids = np.arange(3 * N, dtype=np.uint32)
ctx.upload(ids[:N], a); ctx.upload(ids[N:2 * N], b)
ops = np.full(N, OPS["NAND"], np.uint8)
table = {}
import itertools
sel = os.environ.get('VARIANTS', '6:1,4:1,3:4,3:6,7:8')
sizes = [int(x) for x in os.environ.get('SIZES', '1,74,148,296,444,592,888,1184,1776,2368').split(',')]
for variant, g in (tuple(int(y) for y in x.split(':')) for x in sel.split(',')):
    ctx.set_kernel_variant(variant)
    if g > 1:
        ctx.set_jobs_per_cta(g)
    row = {}
    for nb in sizes:
        for rep in range(2):
            ctx.gate_batch(ops[:nb], ids[:nb], ids[N:N + nb], None, ids[2 * N:2 * N + nb]); ctx.sync()
        row[nb] = round(ctx.last_batch_ms()[0], 3)
    table[f'v{variant}_G{g}'] = row
    print(f'v{variant}_G{g}', row, flush=True)
json.dump(table, open(os.path.join(ROOT, "gpurun_out", "latency_table.json"), "w"), indent=1)
