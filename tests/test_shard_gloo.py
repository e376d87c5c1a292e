"""N > 1 path on CPU: the engine's static schedule (b200net_schedule: replicated / sharded steps, slot ranges,
one all-gather per sharded step) replayed with the gloo backend, world_size 2, plaintext back-end standing in for
the GPU.  On the GPU the same schedule is compiled into one CUDA graph per rank (b200net_bind_rank)."""
import json
import os
import socket
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

NL = Path(__file__).resolve().parent / "golden" / "netlists"


def _worker(rank, world, port, name, flags, q):
    import torch.distributed as dist

    from iyokan_b200 import netlist as N
    from iyokan_b200.shard import PlainBackend, ScheduledRunner

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = json.load(open(NL / "cases.json"))[name]
    nl = N.Netlist.load(NL / f"{name}.npz")
    eng = N.NetEngine(nl)
    be = PlainBackend(nl, eng, world, flags)
    run = ScheduledRunner(nl, eng, be, rank, world)
    req = case["request"]
    for mem, e in req["rom"].items():
        be.set_nodes(nl.mem[mem], N.bits_of(e["bytes"], e["size"])[:len(nl.mem[mem])])
    if "reset" in nl.in_ports:
        be.set_nodes(nl.in_ports["reset"], [1])
        run.run()
    for c in range(case["cycles"]):
        run.tick()
        if c == 0:
            if "reset" in nl.in_ports:
                be.set_nodes(nl.in_ports["reset"], [0])
            for mem, e in req["ram"].items():
                be.set_nodes(nl.mem[mem], N.bits_of(e["bytes"], e["size"])[:len(nl.mem[mem])])
        for p, e in req["bits"].items():
            w = len(nl.in_ports[p])
            stream = N.bits_of(e["bytes"], e["size"])
            be.set_nodes(nl.in_ports[p], [stream[(w * c + b) % len(stream)] for b in range(w)])
        run.run()
    out = {p: N.bytes_of(be.get_nodes(nodes)) for p, nodes in nl.out_ports.items()}
    ram = {m: N.bytes_of(be.get_nodes(nl.mem[m])) for m in req["ram"]}
    q.put((rank, out, ram, run.collectives, run.exchanged_slots))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name,flags", [("mux-ram-8-16-16", 0), ("cahp-pearl-mux", 1), ("cahp-ruby-mux", 1)])
def test_two_rank_sharded_run_matches_reference_golden(name, flags):
    # flags: 0 = ASAP levels, 1 = slack-aware packing (B200NET_PACK)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, flags, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = json.load(open(NL / "cases.json"))[name]["expected"]
    for rank, out, ram, ncoll, nslots in results:
        for port_name, e in exp["bits"].items():
            assert out[port_name][:len(e["bytes"])] == e["bytes"], (rank, port_name)
        for m, e in exp["ram"].items():
            assert ram[m][:len(e["bytes"])] == e["bytes"], (rank, m)
        assert ncoll > 0 and nslots > 0
    # both ranks issued the same number of collectives (otherwise the run would have dead-locked)
    assert results[0][3] == results[1][3]
    assert np.isfinite(results[0][4])
