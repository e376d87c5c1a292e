"""BASELINE.json configs[4]: 80-bit vs 128-bit parameter sweep - Hom{NAND,MUX} throughput and the blind-rotation
kernel's key-stream GB/s (algorithmic bytes of SURVEY.md 8(d) / measured kernel time) at N = 1/2/4/8 GPUs.

The flavour is a property of the process (B200FHE_FLAVOUR = "" | "80", like the reference's compile-time
IYOKAN_80BIT_SECURITY); scripts/sweep_params.sh runs both.  Under torchrun every rank evaluates its own batch (the path
shards per gate; weak scaling) and rank 0 prints one JSON line per gate kind: bootstraps/s over all ranks (device-timed on
the library stream, max over ranks), decrypted outputs checked against the truth table and a sample bit-exactly against
the oracle, and the reference's CPU gate path (oracle/_ref/ref_driver[80]) timed on a bounded sample of the same inputs.

    python scripts/sweep_params.py [--batch 8192] [--steps 5]
    python -m torch.distributed.run --nproc-per-node N ... scripts/sweep_params.py
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402  (input generation, checker and CPU baseline only)
from iyokan_b200 import Context, OPS  # noqa: E402
from iyokan_b200 import lib as L  # noqa: E402


def cpu_reference(keys, ops, ca, cb, cc, threads):
    if not O.have_ref():
        return None
    d = Path(tempfile.mkdtemp(prefix="b200fhe_sweep_"))
    try:
        keys.save(d)
        for name, arr in (("ops", ops), ("a", ca), ("b", cb), ("c", cc)):
            arr.tofile(d / f"{name}.bin")
        out = O.ref("gates", d, d / "ops.bin", d / "a.bin", d / "b.bin", d / "c.bin", d / "o.bin", threads, 1)
        info = json.loads(out.strip().splitlines()[-1])
        info["out"] = np.fromfile(d / "o.bin", dtype=O.T0).reshape(ops.size, O.TLWE0)
        return info
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192, help="blind rotations per GPU and step (MUX counts two)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, local, world = (int(os.environ.get(k, "0" if k != "WORLD_SIZE" else "1")) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bits = 80 if O.FLAVOUR == "80" else 128
    keys = O.cached_keys(20261017)
    ctx = Context(local)
    ctx.load_keys(keys.bk, keys.ksk)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    bk_bytes = O.N0 * 2 * O.L * 2 * 1024 * 8                                   # key stream per rotation, SURVEY.md 8(d)
    tlwe_bytes = O.TLWE0 * np.dtype(O.T0).itemsize
    for kind in ("NAND", "MUX"):
        per_gate = 2 if kind == "MUX" else 1
        n = args.batch // per_gate
        rng = np.random.default_rng(500 + rank)
        pa, pb, pc = (rng.integers(0, 2, n, dtype=np.uint8) for _ in range(3))
        ca, cb, cc = (O.encrypt_bits(60 + 3 * rank + k, keys, x) for k, x in enumerate((pa, pb, pc)))
        ops = np.full(n, OPS[kind], np.uint8)
        ctx.arena_alloc(4 * n)
        ids = np.arange(4 * n, dtype=np.uint32)
        for k, c in enumerate((ca, cb, cc)):
            ctx.upload(ids[k * n:(k + 1) * n], c)

        def step():
            with torch.cuda.stream(stream):
                flush.zero_()
            ctx.gate_batch(ops, ids[:n], ids[n:2 * n], ids[2 * n:3 * n], ids[3 * n:])

        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        br_ms, ks_ms = ctx.last_batch_ms()
        plan = ctx.last_batch_segments()
        got = ctx.download(ids[3 * n:])
        want_bits = O.plain_gate_vec(ops, pa, pb, pc)
        ok = bool(np.array_equal(O.decrypt_bits(keys, got), want_bits))
        ok = ok and bool(np.array_equal(got[:3], O.gate_batch(keys, ops[:3], ca[:3], cb[:3], cc[:3])))
        if world > 1:
            t = torch.tensor([ms, float(ok)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:], op=dist.ReduceOp.MIN)
            ms, ok = float(t[0]), bool(t[1])
        if rank == 0:
            rot = n * per_gate
            line = {"config": "80-bit vs 128-bit sweep [BASELINE.json configs[4]]", "params_bits": bits, "gate": f"Hom{kind}",
                    "n_gpus": world, "gates_per_gpu": n, "bootstraps_per_s": world * rot * args.steps / (ms / 1e3),
                    "gates_per_s": world * n * args.steps / (ms / 1e3), "ms_per_step": ms / args.steps,
                    "blind_rotation_ms": br_ms, "keyswitch_ms": ks_ms,
                    "br_key_stream_GBps": rot * bk_bytes / (br_ms / 1e3) / 1e9,
                    "br_key_stream_frac_of_hbm_peak": rot * bk_bytes / (br_ms / 1e3) / 1e9 / 6553.9,
                    "key_bytes_per_rotation": bk_bytes, "tlwe_bytes": tlwe_bytes,
                    "launch_plan": [(s["variant"], s["jobs_per_cta"], s["jobs"]) for s in plan], "outputs_ok": ok,
                    "plan_table": L.plan_table()["shapes"]}
            if not args.no_cpu and world == 1:
                cores = os.cpu_count() or 1
                m = min(n, cores * 32)
                info = cpu_reference(keys, ops[:m], ca[:m], cb[:m], cc[:m], cores)
                if info:
                    line["cpu_reference"] = {"bootstraps_per_s": info["bootstraps_per_s"], "cores": cores, "gates": m,
                                             "bits_ok": bool(np.array_equal(O.decrypt_bits(keys, info["out"]), want_bits[:m])),
                                             "what": f"oracle/_ref/ref_driver{O.FLAVOUR}: TFHEpp gate path, {m} gates on {cores} threads"}
            print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
