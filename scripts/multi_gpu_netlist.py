"""torchrun entry: one netlist evaluated on N GPUs with level-sliced sharding + per-level NCCL all-gather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/multi_gpu_netlist.py --case cahp-pearl-mux --cycles 2
Prints one JSON line per case from rank 0: cycles/s, bootstraps/s, collectives per cycle, outputs verified
against the plaintext back-end.  (Also runs with N = 1 without torchrun.)"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402  (checker: encrypt inputs / decrypt outputs)
from iyokan_b200 import Context, netlist as N  # noqa: E402
from iyokan_b200.shard import GpuBackend, ShardedRunner  # noqa: E402

NL = ROOT / "tests" / "golden" / "netlists"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="cahp-pearl-mux")
    ap.add_argument("--cycles", type=int, default=2)
    ap.add_argument("--min-shard-width", type=int, default=-1, help="-1 = cost model (b200fhe_plan_ms)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    keys = O.cached_keys(424242)
    case = json.load(open(NL / "cases.json"))[args.case]
    nl = N.Netlist.load(NL / f"{args.case}.npz")
    eng = N.NetEngine(nl)
    ctx = Context(local)
    ctx.load_keys(keys.bk, keys.ksk)
    be = GpuBackend(nl, eng, ctx, world)
    run = ShardedRunner(nl, eng, be, rank, world, min_shard_width=None if args.min_shard_width < 0 else args.min_shard_width)
    enc = lambda bits: O.encrypt_bits(31, keys, bits)  # same seed on every rank -> identical ciphertexts  # noqa: E731
    req = case["request"]
    dffs = np.nonzero(nl.kind == N.DFF)[0]
    be.set_nodes(dffs, np.tile(N.trivial(0), (dffs.size, 1)))
    for mem, e in req["rom"].items():
        be.set_nodes(nl.mem[mem], enc(N.bits_of(e["bytes"], e["size"])[:len(nl.mem[mem])]))
    plain = N.PlainRunner(nl)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if "reset" in nl.in_ports:
        be.set_nodes(nl.in_ports["reset"], N.trivial(1)[None])
        run.run()
    times = []
    for c in range(args.cycles):
        run.tick()
        if c == 0:
            if "reset" in nl.in_ports:
                be.set_nodes(nl.in_ports["reset"], N.trivial(0)[None])
            for mem, e in req["ram"].items():
                be.set_nodes(nl.mem[mem], enc(N.bits_of(e["bytes"], e["size"])[:len(nl.mem[mem])]))
        for p, e in req["bits"].items():
            w = len(nl.in_ports[p])
            stream = N.bits_of(e["bytes"], e["size"])
            be.set_nodes(nl.in_ports[p], enc(np.array([stream[(w * c + b) % len(stream)] for b in range(w)], np.uint8)))
        barrier()
        t = time.time()
        c0 = run.collectives
        run.run()
        barrier()
        times.append(time.time() - t)
        ncoll = run.collectives - c0
    want = plain.run(args.cycles, inputs={p: N.bits_of(e["bytes"], e["size"]) for p, e in req["bits"].items()},
                     rams={p: N.bits_of(e["bytes"], e["size"]) for p, e in req["ram"].items()},
                     roms={p: N.bits_of(e["bytes"], e["size"]) for p, e in req["rom"].items()})
    ok = all(np.array_equal(O.decrypt_bits(keys, be.get_nodes(nodes)), want[p]) for p, nodes in nl.out_ports.items())
    if world > 1:
        t = torch.tensor([max(times[-1], 0.0), float(ok)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.MIN)
        last, ok = float(t[0]), bool(t[1])
    else:
        last = times[-1]
    if rank == 0:
        print(json.dumps({"case": args.case, "n_gpus": world, "cycles": args.cycles, "s_per_cycle": last,
                          "bootstraps_per_cycle": eng.bootstraps_per_cycle,
                          "bootstraps_per_s": eng.bootstraps_per_cycle / last, "levels": eng.num_levels,
                          "collectives_per_cycle": ncoll, "min_shard_width": args.min_shard_width,
                          "outputs_match_plain_backend": ok}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
