"""`python -m iyokan_b200 plain|tfhe ...`: the command line of `iyokan plain` / `iyokan tfhe` (src/main.cpp:12-277)
on the plaintext host engine and on the B200 back-end (CMUX memories are evaluated as MUX memories).

    python -m iyokan_b200 plain --blueprint B.toml -i req.plain -o res.plain [-c N] [--snapshot S]
    python -m iyokan_b200 tfhe  --blueprint B.toml --evalkey EK -i req.enc -o res.enc -c N [--snapshot S]
    python -m iyokan_b200 plain|tfhe --resume S -c N -o res            (tfhe also needs --evalkey)
    python -m iyokan_b200 packet toml2packet|packet2toml --in X --out Y (plain packets, as iyokan-packet does)
    python -m torch.distributed.run --nproc-per-node N -m iyokan_b200 tfhe ...   (N GPUs of one box, NCCL)

Same option names as the reference; options that only tune its CPU scheduler (--cpu, --sched, --gpu, --num-gpu,
--show-combinational-progress, --dump-time-csv-prefix, --dump-graph-*) are accepted and ignored; --dump-prefix
writes the result packet after every cycle (encrypted in tfhe mode).  Requests may be given as binary packets
(what `iyokan-packet toml2packet / enc` writes) or, in plain mode, directly as the TOML source.  Errors follow the
reference: message on stderr, exit status 1 (src/error.hpp:21-47)."""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

from .blueprint import read_blueprint
from .frontend import Frontend, FrontendError
from .packet import PacketError, PlainPacket, TFHEPacket, read_eval_key


def die(*msg):
    print("[error]", *msg, file=sys.stderr)
    raise SystemExit(1)


def _load_plain_request(path) -> PlainPacket:
    data = Path(path).read_bytes()
    if data[:1] == b"\x01":
        try:
            return PlainPacket.loads(data)
        except PacketError:
            pass
    try:
        return PlainPacket.from_toml(data.decode())
    except Exception as e:  # noqa: BLE001
        die(f"Invalid request packet {path}: neither a PlainPacket archive nor its TOML source ({e})")


def _common(sp):
    sp.add_argument("--blueprint")
    sp.add_argument("-i", "--in", dest="input")
    sp.add_argument("-o", "--out", dest="output")
    sp.add_argument("-c", dest="cycles", type=int)
    sp.add_argument("--snapshot")
    sp.add_argument("--resume")
    sp.add_argument("--skip-reset", action="store_true")
    sp.add_argument("--quiet", action="store_true")
    sp.add_argument("--verbose", action="store_true")
    sp.add_argument("--stdout-csv", action="store_true")
    sp.add_argument("--dump-prefix")
    for ignored in ("--cpu", "--sched", "--gpu", "--num-gpu", "--gpu_num", "--dump-time-csv-prefix",
                    "--dump-graph-json-prefix", "--dump-graph-dot-prefix", "--secret-key"):
        sp.add_argument(ignored, help=argparse.SUPPRESS)
    for ignored in ("--show-combinational-progress", "--enable-gpu", "--no-stdout-csv"):
        sp.add_argument(ignored, action="store_true", help=argparse.SUPPRESS)


def _init_distributed(mode):
    """One process per GPU under torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment)."""
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 0, 1
    import torch
    import torch.distributed as dist

    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if mode == "tfhe":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    return rank, local, world


def _run(args, mode):
    rank, local, world = _init_distributed(mode)
    log = (lambda *a: None) if args.quiet or rank else (lambda *a: print("[info]", *a, file=sys.stderr))
    if bool(args.blueprint) == bool(args.resume):
        die("exactly one of --blueprint (a new run) and --resume (from a snapshot) is required")
    if not args.output:
        die("-o/--out is required")
    ctx = None
    if mode == "tfhe":
        if not args.evalkey:
            die("--evalkey is required")
        from .lib import B200FheError, Context

        try:
            ctx = Context(local if world > 1 else args.gpu_index)
            bk, ksk = read_eval_key(args.evalkey)
            ctx.load_keys(bk, ksk)
        except (B200FheError, PacketError, OSError) as e:
            die(e)
    ok = False
    try:
        if args.resume:
            if not Frontend.is_snapshot(args.resume) or Frontend.snapshot_mode(args.resume) != mode:
                die("Invalid resume file:", args.resume)
            fe = Frontend.resume(args.resume, ctx, rank, world)
            cycles = args.cycles
            if cycles is None:
                die("-c is required with --resume")
        else:
            if not args.input:
                die("-i/--in is required")
            nl = read_blueprint(args.blueprint)
            fe = Frontend(nl, mode, ctx, rank, world)
            req = _load_plain_request(args.input) if mode == "plain" else TFHEPacket.load(args.input)
            fe.load_request(req)
            cycles = args.cycles if args.cycles is not None else req.num_cycles
            if cycles is None or cycles < 0:
                die("the number of cycles is given neither by -c nor by the request packet")
        log(f"{mode} on {world} process(es): {fe.nl.n} nodes, {fe.eng.num_levels} levels, {fe.eng.bootstraps_per_cycle} bootstraps/cycle, "
            f"{fe.eng.num_dff} DFF; running {cycles} cycle(s) from cycle {fe.cycle}")
        t0 = time.time()
        fe.run(cycles, skip_reset=args.skip_reset, dump_prefix=args.dump_prefix)
        log(f"done. ({int((time.time() - t0) * 1e6)} us)")
        res = fe.result()
        if rank == 0:
            res.save(args.output)
            if args.snapshot:
                fe.snapshot(args.snapshot)
        if rank == 0 and args.stdout_csv and mode == "plain":
            for name in sorted(res.bits):
                print(f"{fe.cycle},{name},{sum(int(b) << i for i, b in enumerate(res.bits[name]))}")
        ok = True
    except (FrontendError, PacketError, ValueError, KeyError, OSError, ZeroDivisionError) as e:
        die(e)
    finally:
        if world > 1:
            import torch.distributed as dist

            if dist.is_initialized():
                if ok:  # a rank that failed must not wait for the others (they may sit in a different collective)
                    dist.barrier()
                    dist.destroy_process_group()
                else:
                    os._exit(1)
        if ctx is not None:
            ctx.close()
    return 0


def _packet(args):
    try:
        if args.cmd == "toml2packet":
            PlainPacket.from_toml(Path(args.input).read_text()).save(args.output)
        else:
            text = PlainPacket.load(args.input).to_toml()
            if args.output:
                Path(args.output).write_text(text)
            else:
                print(text)
    except (PacketError, ValueError, OSError) as e:
        die(e)
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m iyokan_b200", description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="mode", required=True)
    _common(sub.add_parser("plain", help="plaintext back-end (the reference's `iyokan plain`)"))
    t = sub.add_parser("tfhe", help="encrypted back-end on the GPU (the reference's `iyokan tfhe`)")
    _common(t)
    t.add_argument("--evalkey")
    t.add_argument("--gpu-index", type=int, default=0)
    k = sub.add_parser("packet", help="plain packet <-> TOML (iyokan-packet toml2packet / packet2toml)")
    k.add_argument("cmd", choices=["toml2packet", "packet2toml"])
    k.add_argument("--in", dest="input", required=True)
    k.add_argument("--out", dest="output")
    args = ap.parse_args(argv)
    if args.mode == "packet":
        return _packet(args)
    return _run(args, args.mode)


if __name__ == "__main__":
    raise SystemExit(main())
