#!/usr/bin/env python
"""bench.py — TFHE gate bootstraps / second on the batched HomNAND workload (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (TFHEpp)

One "step" = one pass of the hot path over one batch of 8192 independent HomNAND gates
(fresh lvl0 encryptions of seeded random bits, 128-bit parameters).  With N > 1 every rank
evaluates its own batch (the path shards per gate with no data-path collective; keys replicated),
so scaling is weak and `value` = N * 8192 * K / max-over-ranks time.

`value`  : ciphertexts resident in HBM when the timed region starts, device-timed (CUDA events on
           the library's stream), max over ranks.
`e2e`    : the same metric through the C ABI call with HOST buffers (b200fhe_gates_host): per step
           the two input arrays go host->device from pinned memory and the outputs come back.
`roofline`: blind-rotation kernel only (the dominant kernel), algorithmic bytes per SURVEY.md §8(d).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BATCH = 8192
WORKLOAD = ("batched HomNAND microbench: 8192 independent gate bootstraps per GPU, "
            "128-bit params (n=636, N=1024, l=3, Bgbit=6, t=7) [BASELINE.json configs[1]]")
BK_BYTES_PER_ROTATION = 636 * 2 * 3 * 2 * 1024 * 8      # 62,521,344
KS_BYTES_PER_SWITCH = 1024 * 7 * 637 * 2 * 3 // 4        # 6,849,024: one of three rows per (i, j), 3/4 of them non-zero
BYTES_PER_BOOTSTRAP = BK_BYTES_PER_ROTATION + KS_BYTES_PER_SWITCH + 3 * 1274   # 69,374,190
BR_KERNEL_BYTES_PER_JOB = BK_BYTES_PER_ROTATION + 2 * 1274 + 4100               # key stream + TLWE in + lvl1 out
KEY_SEED = 20261017
KERNEL_NAMES = {(7, 8): "br7_kernel<8>", (3, 6): "br3_kernel<6>", (3, 4): "br3_kernel<4>", (3, 2): "br3_kernel<2>",
                (4, 1): "br4_kernel", (6, 1): "br6_kernel"}


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        self.proc = subprocess.Popen([exe, f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # the load samples are the upper half of what we saw (idle samples bracket the region)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(keys, rank: int, n: int):
    import oracle as O

    rng = np.random.default_rng(1000 + rank)
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    return pa, pb, O.encrypt_bits(50 + 2 * rank, keys, pa), O.encrypt_bits(51 + 2 * rank, keys, pb)


def cpu_reference_run(keys, n_gates: int, threads: int, repeat: int = 1):
    """Time the UNMODIFIED reference (TFHEpp HomNAND, oracle/_ref/ref_driver) on the host cores."""
    import oracle as O

    if not O.have_ref():
        return None
    d = Path(tempfile.mkdtemp(prefix="b200fhe_ref_"))
    try:
        keys.save(d)
        pa, pb, ca, cb = make_inputs(keys, 0, n_gates)
        ca.tofile(d / "a.bin")
        cb.tofile(d / "b.bin")
        np.full(n_gates, O.OPS["NAND"], np.uint8).tofile(d / "ops.bin")
        out = O.ref("gates", d, d / "ops.bin", d / "a.bin", d / "b.bin", "-", d / "o.bin", threads, repeat)
        info = json.loads(out.strip().splitlines()[-1])
        got = O.decrypt_bits(keys, np.fromfile(d / "o.bin", dtype=np.uint16).reshape(n_gates, 637))
        info["bits_ok"] = bool(np.array_equal(got, 1 - (pa & pb)))
        return info
    finally:
        shutil.rmtree(d, ignore_errors=True)


def cpu_port_run(keys, n_gates: int, threads: int):
    """Fallback CPU baseline: the exact-integer oracle port (only when oracle/_ref is absent)."""
    import oracle as O

    pa, pb, ca, cb = make_inputs(keys, 0, n_gates)
    ops = np.full(n_gates, O.OPS["NAND"], np.uint8)
    t = time.time()
    out = O.gate_batch(keys, ops, ca, cb, nthreads=threads)
    sec = time.time() - t
    ok = bool(np.array_equal(O.decrypt_bits(keys, out), 1 - (pa & pb)))
    return {"gates": n_gates, "bootstraps": n_gates, "seconds": sec, "bootstraps_per_s": n_gates / sec,
            "threads": threads, "bits_ok": ok}


def run_reference(args, rank: int, world: int):
    """--impl reference: rank 0 alone times the reference CPU path on a bounded sample per step."""
    if rank != 0:
        return
    import oracle as O

    keys = O.cached_keys(KEY_SEED)
    cores = os.cpu_count() or 1
    sample = min(BATCH, max(cores * 64, 256))  # ~1 s of work per step on all host cores
    if O.have_ref():
        kind = "reference"
        cpu_reference_run(keys, cores, cores)  # warm-up (page in keys, FFT tables)
        t = time.time()
        info = cpu_reference_run(keys, sample, cores, repeat=args.steps)
        total_s = info["seconds"]
        ok = info["bits_ok"]
        del t
    else:
        kind = "port"
        t = time.time()
        ok = True
        for _ in range(args.steps):
            ok &= cpu_port_run(keys, sample, cores)["bits_ok"]
        total_s = time.time() - t
    value = sample * args.steps / total_s
    line = {
        "impl": "reference", "metric": "tfhe_gate_bootstraps_per_s", "value": value, "unit": "bootstraps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if kind == "reference" else "u32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "sample_per_step": sample,
                                         "parallelism": f"{cores} host threads, one gate per thread at a time"},
        "cpu_baseline": {"value": value, "unit": "bootstraps/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} HomNAND gates per step x {args.steps} steps, {cores} host threads"},
        "e2e": {"value": value, "unit": "bootstraps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "outputs_ok": bool(ok),
    }
    print(json.dumps(line), flush=True)


def BN_binary_available(tools) -> bool:
    return tools.O.IYOKAN_B200.exists() and os.environ.get("B200FHE_NETLIST_HOST", "") != "python"


def netlist_legs(args, ctx, stream, rank, world, barrier):
    """north_star's workload through the product path (see bench_netlist.py); every rank takes part."""
    import torch
    import torch.distributed as dist

    import bench_netlist as BN
    from iyokan_b200.packet import read_eval_key

    names = [x for x in args.netlist_cases.split(",") if x] or (["cahp-pearl-mux", "cahp-ruby-mux", "mux-ram-8-16-16"] if world == 1 else ["cahp-pearl-mux"])
    box = [tempfile.mkdtemp(prefix="b200fhe_net_") if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    work = Path(box[0])
    lines = []
    try:
        tools = BN.RefTools(work)
        if not tools.have_packet:
            if rank == 0:
                print("bench.py: oracle/_ref/iyokan-packet is not built; netlist leg skipped", file=sys.stderr)
            return []
        assets = work / "test"
        if rank == 0:
            BN._assets(assets)
            tools.genkeys()
        barrier()
        use_binary = BN_binary_available(tools)
        if not use_binary:
            bk, ksk = read_eval_key(tools.ek)
            ctx.load_keys(bk, ksk)

        def allreduce_max(x):
            if world == 1:
                return float(x)
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def stream_events(fn):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            ev0.record(stream)
            fn()
            ev1.record(stream)
            barrier()
            ms = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        for name in names:
            if use_binary:
                line = BN.run_case_binary(name, tools, assets, rank, int(os.environ.get("LOCAL_RANK", "0")), world,
                                          args.netlist_cycles, barrier, allreduce_max, cpu=not args.no_cpu_baseline)
            else:
                line = BN.run_case(name, tools, assets, ctx, rank, world, args.netlist_cycles, barrier, stream_events,
                                   cpu=not args.no_cpu_baseline)
            if line is not None:
                lines.append(line)
    finally:
        barrier()
        if rank == 0:
            shutil.rmtree(work, ignore_errors=True)
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--jobs-per-cta", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-netlist", action="store_true", help="skip the netlist leg (CAHP processor / mux-ram)")
    ap.add_argument("--netlist-cycles", type=int, default=10)
    ap.add_argument("--netlist-cases", default="", help="comma list; default: cahp-pearl-mux (+ cahp-ruby-mux, mux-ram-8-16-16 at N=1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import oracle as O  # checker + cpu_baseline leg only
    from iyokan_b200 import Context, OPS, PinnedBuffer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.batch
    keys = O.cached_keys(KEY_SEED)
    pa, pb, ca, cb = make_inputs(keys, rank, n)

    ctx = Context(local_rank)
    if args.variant:
        ctx.set_kernel_variant(args.variant)
    if args.jobs_per_cta:
        ctx.set_jobs_per_cta(args.jobs_per_cta)
    ctx.load_keys(keys.bk, keys.ksk)
    ctx.arena_alloc(4 * n)
    ids = np.arange(4 * n, dtype=np.uint32)
    ctx.upload(ids[:n], ca)
    ctx.upload(ids[n:2 * n], cb)
    ops = np.full(n, OPS["NAND"], np.uint8)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        with torch.cuda.stream(stream):
            flush.zero_()
        ctx.gate_batch(ops, ids[:n], ids[n:2 * n], None, ids[3 * n:])

    # host (pinned) buffers for the end-to-end arm
    h_a, h_b, h_o = (PinnedBuffer((n, 637), np.uint16) for _ in range(3))
    h_a.array[:] = ca
    h_b.array[:] = cb

    def step_e2e():
        with torch.cuda.stream(stream):
            flush.zero_()
        ctx.gates_host(ops, h_a.array, h_b.array, None, out=h_o.array)

    def timed(step_fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        br_ms = []
        barrier()
        ev0.record(stream)
        for _ in range(steps):
            step_fn()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, br_ms

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    l0 = ctx.launch_count
    ms_total, _ = timed(step_resident, args.steps)
    launches = ctx.launch_count - l0
    # per-launch duration of the dominant kernel, CUDA events on the launching stream
    br_list, ks_list, seg_runs = [], [], []
    for _ in range(min(args.steps, 5)):
        step_resident()
        ctx.sync()
        a, b = ctx.last_batch_ms()
        br_list.append(a)
        ks_list.append(b)
        seg_runs.append(ctx.last_batch_segments())
    clocks = sampler.stop() if rank == 0 else None

    out_res = ctx.download(ids[3 * n:])
    bits_ok = bool(np.array_equal(O.decrypt_bits(keys, out_res), 1 - (pa & pb)))
    k = min(n, 64)  # ciphertext-exact against the oracle on a strided sample of the batch (all host threads)
    pick = np.linspace(0, n - 1, k).astype(np.int64)
    exact_ok = bool(np.array_equal(out_res[pick], O.gate_batch(keys, ops[pick], ca[pick], cb[pick],
                                                                nthreads=os.cpu_count() or 1)))

    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    e2e_ok = bool(np.array_equal(h_o.array, out_res))

    if world > 1:
        flags = torch.tensor([int(bits_ok and exact_ok and e2e_ok)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        all_ok = bool(flags.item())
    else:
        all_ok = bits_ok and exact_ok and e2e_ok

    net_lines = []
    if not args.no_netlist and not args.variant and not args.jobs_per_cta and n == BATCH:
        net_lines = netlist_legs(args, ctx, stream, rank, world, barrier)

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        value = world * n * args.steps / (ms_total / 1e3)
        br_ms = float(np.mean(br_list))
        # launch plan of one step (same every step): the dominant kernel is the segment with most jobs
        segs = [{**seg_runs[0][k], "ms": float(np.mean([r[k]["ms"] for r in seg_runs]))} for k in range(len(seg_runs[0]))]
        for sg in segs:
            sg["kernel"] = KERNEL_NAMES.get((sg["variant"], sg["jobs_per_cta"]), f"variant {sg['variant']}")
        dom = max(segs, key=lambda sg: sg["jobs"])
        achieved = dom["jobs"] * BR_KERNEL_BYTES_PER_JOB / (dom["ms"] / 1e3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "br_kernel_traffic.json"
        if tp.exists():  # dram bytes of ONE launch of the dominant kernel, from the committed ncu --set full capture
            try:
                tj = json.load(open(tp))
                if tj.get("kernel") == dom["kernel"] and tj.get("jobs_in_captured_launch") == dom["jobs"]:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "tfhe_gate_bootstraps_per_s", "value": value, "unit": "bootstraps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": n, "parallelism": f"replicated keys, {world} independent batch shard(s)",
                       "l2": "256 MiB device memset between steps inside the timed region (L2 flush); "
                             "working set keys+ciphertexts ~176 MB > 126 MB L2"},
            "e2e": {"value": world * n * args.steps / (ms_e2e / 1e3), "unit": "bootstraps/s",
                    "h2d_bytes_per_step": 2 * n * 1274, "d2h_bytes_per_step": n * 1274},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": dom["kernel"] + " (blind rotation)", "peak_source": peak_src,
                         "kernel_ms_per_launch": dom["ms"], "jobs_per_launch": dom["jobs"],
                         "algorithmic_bytes_per_launch": dom["jobs"] * BR_KERNEL_BYTES_PER_JOB,
                         "launch_plan": segs, "blind_rotation_ms_per_step": br_ms,
                         "ks_kernel_ms_per_launch": float(np.mean(ks_list)),
                         "whole_step_frac": n * BR_KERNEL_BYTES_PER_JOB / (br_ms / 1e3) / 1e9 / peak,
                         "gate_model_frac": (value / world) * BYTES_PER_BOOTSTRAP / (peak * 1e9)},
            "outputs_ok": all_ok,
            "checks": {"decrypted_bits": f"all {n} gates per rank == NAND of the plaintexts",
                       "ciphertext_exact": f"{k} gates per rank (strided over the batch) bit-identical to the oracle",
                       "e2e_equals_resident": "host-buffer result == device-resident result, whole batch"},
        }
        # honest bound of the dominant kernel: the int32 multiply pipe (DRAM is idle: keys sit in L2).  Pipe cycles per
        # CMUX step and job from the SASS of br7_kernel<8,8> (scripts/sass_count.py, profiles/r02_sass_count.txt; per thread
        # and step: 1897 IMAD x 2 + 960 IMAD.HI x 4 + 576 IMAD.WIDE x 4 + 177 IMAD.IADD/SHL/MOV x 2 cycles = 10,292 per warp;
        # 16 warps on 4 schedulers, 8 jobs per SM)
        sm_hz = 1e6 * ((clocks or {}).get("sm_mhz") or 1965.0)
        pipe_cycles_per_rotation = 636 * 10292 * 4 / 8
        alu_peak = 148 * sm_hz / pipe_cycles_per_rotation
        line["roofline"]["alu"] = {"bound": "int32 multiply pipe (fmaheavy), 148 SMs at the sampled SM clock",
                                   "pipe_cycles_per_rotation": pipe_cycles_per_rotation, "peak": alu_peak,
                                   "achieved": dom["jobs"] / (dom["ms"] / 1e3), "unit": "rotations/s",
                                   "frac": dom["jobs"] / (dom["ms"] / 1e3) / alu_peak}
        from iyokan_b200.lib import plan_table

        line["roofline"]["plan_table"] = plan_table()   # ms per wave of each shape, measured at key load on this device
        if net_lines:
            line["netlist"] = net_lines[0]
            if len(net_lines) > 1:
                line["netlist_more"] = net_lines[1:]
            line["outputs_ok"] = bool(all_ok and all(x["outputs_ok"] for x in net_lines))
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = min(n, max(cores * 512, 256))  # ~10 s of host work: the whole batch on a 16-core box
            info = cpu_reference_run(keys, sample, cores)
            kind = "reference"
            if info is None:
                info, kind = cpu_port_run(keys, max(cores * 2, 32), cores), "port"
            line["cpu_baseline"] = {"value": info["bootstraps_per_s"], "unit": "bootstraps/s", "cores": cores,
                                    "kind": kind, "bits_ok": info["bits_ok"],
                                    "sample": f"{info['gates']} HomNAND gates of the same workload, {cores} host threads, "
                                              f"{info['seconds']:.2f} s"}
        print(json.dumps(line), flush=True)

    for hb in (h_a, h_b, h_o):
        hb.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
