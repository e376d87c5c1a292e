"""bench.py's netlist leg: north_star's workload (VSP CAHP processor, mux-ram) through the product path, on the reference's
own blueprint / netlist / request files, keys and packets written by the reference's `iyokan-packet`, result decrypted by
it and compared with the reference's plaintext back-end; the reference's CPU back-end (`iyokan tfhe --cpu N`) timed in
the same run on the same keys and request.

Product path measured: `iyokan-b200` (iyokan_b200/host/iyokan_b200_main.cpp) = the reference's OWN blueprint / Yosys-JSON
loader, builtin MUX-memory generators and packet / key I/O (src/iyokan.hpp, src/packet.hpp, compiled unmodified) in front
of this repo's engine: b200net static schedule (b200net_bind_rank) -> one CUDA graph per clock on every rank (kernels +
NCCL all-gathers of the sharded steps) -> result packet.  One process per GPU: under torchrun every rank spawns the binary
with its RANK / LOCAL_RANK / WORLD_SIZE.  It therefore evaluates exactly the netlist the reference's CPU back-end evaluates
(same node and bootstrap counts).  When the binary is not built (no reference tree at build time) the leg falls back to the
Python front end of this repo (same engine, own loader) and says so in `host`.
`oracle/_ref/*` (the reference compiled here, test infrastructure) is used for key generation, encryption, decryption,
the plaintext expectation and the CPU baseline only - never on the measured path.
"""
from __future__ import annotations

import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent

CASES = {
    # name: (blueprint, request, golden cycles, bootstraps per clock are reported by the engine)
    "cahp-pearl-mux": ("config-toml/cahp-pearl-mux.toml", "in/test09.in"),
    "cahp-ruby-mux": ("config-toml/cahp-ruby-mux.toml", "in/test09.in"),
    "mux-ram-8-16-16": ("config-toml/mux-ram-8-16-16.toml", "in/test08.in"),
}


def _assets(dst: Path) -> Path:
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_ref_assets

    return make_ref_assets.materialise(dst)


class RefTools:
    """Keys / packets by the reference's own tools (oracle/_ref, built by oracle/Makefile from the reference sources)."""

    def __init__(self, work: Path):
        import oracle as O

        self.O, self.work = O, work
        self.sk, self.ek = work / "sk", work / "ek"
        self.have_packet = O.have_iyokan_packet()
        self.have_iyokan = O.IYOKAN_REF.exists()

    def genkeys(self):
        self.O.iyokan_packet("genkey", "--type", "tfhepp", "--out", self.sk)
        self.O.iyokan_packet("genevalkey", "--in", self.sk, "--out", self.ek)  # 2.2 GB, as the reference writes it

    def encrypt_request(self, toml_in: Path, tag: str):
        req, enc = self.work / f"{tag}.req", self.work / f"{tag}.req.enc"
        self.O.iyokan_packet("toml2packet", "--in", toml_in, "--out", req)
        self.O.iyokan_packet("enc", "--key", self.sk, "--in", req, "--out", enc)
        return req, enc

    def decrypt(self, enc: Path) -> Path:
        out = enc.with_suffix(".dec")
        self.O.iyokan_packet("dec", "--key", self.sk, "--in", enc, "--out", out)
        return out

    def plain(self, blueprint: Path, req: Path, cycles: int, tag: str) -> Path:
        out = self.work / f"{tag}.plain"
        r = self.O.iyokan_ref("plain", "--blueprint", blueprint, "-i", req, "-o", out, "-c", cycles, "--quiet")
        if r.returncode != 0:
            raise RuntimeError("iyokan plain failed: " + r.stderr[-500:])
        return out

    def cpu_tfhe(self, blueprint: Path, enc: Path, cores: int, tag: str, timeout=1500):
        """One clock of the reference's CPU back-end; returns its own per-cycle time ('done. (N us)', iyokan_tfhepp.cpp:557)."""
        out = self.work / f"{tag}.cpu.enc"
        t = time.time()
        r = self.O.iyokan_ref("tfhe", "--blueprint", blueprint, "--evalkey", self.ek, "-i", enc, "-o", out, "-c", 1, "--cpu", cores,
                              "--skip-reset", timeout=timeout)
        wall = time.time() - t
        if r.returncode != 0:
            raise RuntimeError("iyokan tfhe failed: " + r.stderr[-500:])
        m = re.findall(r"done\. \((\d+) us\)", r.stderr + r.stdout)
        return (int(m[-1]) / 1e6 if m else wall), wall, out


def _finish_line(line, name, tools, bp, req, enc, res_enc, cycles, nl_for_plain, cpu):
    """rank 0: decrypt + compare with the plaintext back-end, then time the reference's CPU back-end on the same files."""
    from iyokan_b200.frontend import Frontend
    from iyokan_b200.packet import PlainPacket

    got = PlainPacket.load(tools.decrypt(res_enc))
    if tools.have_iyokan:
        want = PlainPacket.load(tools.plain(bp, req, cycles, name))
        line["checker"] = "iyokan-packet dec == reference `iyokan plain` on the same blueprint and request"
    else:
        fp = Frontend(nl_for_plain(), "plain")
        fp.load_request(PlainPacket.load(req))
        fp.run(cycles)
        want = fp.result()
        line["checker"] = "iyokan-packet dec == this repo's plaintext evaluator (reference iyokan binary not built)"
    ok = got.num_cycles == want.num_cycles and set(got.bits) == set(want.bits)
    ok = ok and all(np.array_equal(got.bits[k], want.bits[k]) for k in want.bits)
    ok = ok and all(np.array_equal(got.ram[k], want.ram[k]) for k in want.ram)
    line["outputs_ok"] = bool(ok)
    boots, s_per_cycle = line["bootstraps_per_cycle"], line["s_per_cycle"]
    if cpu and tools.have_iyokan:
        cores = os.cpu_count() or 1
        sec, wall, _ = tools.cpu_tfhe(bp, enc, cores, name)
        line["cpu_baseline"] = {
            "value": boots / sec, "unit": "bootstraps/s", "s_per_cycle": sec, "cores": cores, "kind": "reference",
            "sample": f"`iyokan tfhe --cpu {cores} -c 1 --skip-reset` on the same blueprint, evalkey and encrypted request: "
                      f"1 clock, its own 'done. (N us)' line ({wall:.0f} s wall incl. loading the 2.2 GB key)"}
        line["speedup_vs_cpu"] = sec / s_per_cycle
    return line


def run_case_binary(name: str, tools: RefTools, assets: Path, rank: int, local_rank: int, world: int, cycles: int, barrier,
                    allreduce_max, cpu: bool):
    """The measured path: one `iyokan-b200` process per GPU (spawned by this rank), clock time = max over ranks."""
    bp_rel, req_rel = CASES[name]
    bp = assets / bp_rel
    if rank == 0:
        tools.encrypt_request(assets / req_rel, name)
    barrier()
    req, enc = tools.work / f"{name}.req", tools.work / f"{name}.req.enc"
    res_enc = tools.work / f"{name}.res.enc"
    env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(local_rank), WORLD_SIZE=str(world),
               B200FHE_ID_FILE=str(tools.work / f"{name}.ncclid"))
    r = subprocess.run([str(tools.O.IYOKAN_B200), "tfhe", "--blueprint", str(bp), "--evalkey", str(tools.ek), "-i", str(enc), "-o",
                        str(res_enc), "-c", str(cycles), "--stats-json"], env=env, capture_output=True, text=True, timeout=1800)
    if r.returncode != 0:
        raise RuntimeError(f"iyokan-b200 failed on rank {rank}: " + (r.stderr or r.stdout)[-800:])
    st = json.loads(r.stdout.strip().splitlines()[-1])
    secs = allreduce_max(st["seconds"])
    barrier()
    if rank != 0:
        return None
    s_per_cycle = secs / cycles
    line = {"case": name, "host": "iyokan-b200 (reference loader + packet I/O, this repo's engine through the C ABI), one process per GPU",
            "blueprint": "test/" + bp_rel, "request": "test/" + req_rel, "cycles": cycles, "n_gpus": world,
            "s_per_cycle": s_per_cycle, "bootstraps_per_cycle": st["bootstraps_per_cycle"],
            "bootstraps_per_s": st["bootstraps_per_cycle"] / s_per_cycle, "nodes": st["nodes"], "levels": st["levels"],
            "steps": st["steps"], "packed": st["packed"], "collectives_per_cycle": st["collectives_per_cycle"],
            "exchanged_bytes_per_cycle": st["exchanged_bytes_per_cycle"], "model_s_per_cycle": st["model_s_per_cycle"],
            "gpu_launches_per_cycle": st["gpu_launches_per_cycle"], "cuda_graph": os.environ.get("B200FHE_NO_GRAPH", "0") != "1",
            "timing": "host steady_clock around the clock loop + final stream sync inside each process (reset pass excluded), max over ranks"}

    def nl_for_plain():
        from iyokan_b200.blueprint import read_blueprint

        return read_blueprint(bp)

    return _finish_line(line, name, tools, bp, req, enc, res_enc, cycles, nl_for_plain, cpu)


def run_case(name: str, tools: RefTools, assets: Path, ctx, rank: int, world: int, cycles: int, barrier, stream_events,
             cpu: bool, group=None):
    """Fallback host (Python front end of this repo, same engine): evaluates one blueprint for `cycles` clocks on `world`
    ranks; returns the dict of the bench line (rank 0) or None."""
    from iyokan_b200.blueprint import read_blueprint
    from iyokan_b200.frontend import Frontend
    from iyokan_b200.packet import PlainPacket, TFHEPacket

    bp_rel, req_rel = CASES[name]
    bp = assets / bp_rel
    if rank == 0:
        tools.encrypt_request(assets / req_rel, name)
    barrier()
    req, enc = tools.work / f"{name}.req", tools.work / f"{name}.req.enc"
    nl = read_blueprint(bp)
    t0 = time.time()
    fe = Frontend(nl, "tfhe", ctx, rank, world, group)
    bind_s = time.time() - t0
    fe.load_request(TFHEPacket.load(enc))
    fe.run(0)            # reset pass (and first replay of the graph: warm-up)
    ctx.sync()
    barrier()
    l0 = ctx.launch_count
    ms = stream_events(lambda: fe.run(cycles))
    launches = ctx.launch_count - l0
    res = fe.result()
    info = fe.eng.schedule_info()
    boots = fe.eng.bootstraps_per_cycle
    line = None
    if rank == 0:
        res_enc = tools.work / f"{name}.res.enc"
        res.save(res_enc)
        s_per_cycle = ms / 1e3 / cycles
        line = {
            "case": name, "host": "python front end of this repo (iyokan-b200 binary not built)",
            "blueprint": "test/" + bp_rel, "request": "test/" + req_rel, "cycles": cycles, "n_gpus": world,
            "s_per_cycle": s_per_cycle, "bootstraps_per_cycle": boots, "bootstraps_per_s": boots / s_per_cycle,
            "levels": fe.eng.num_levels, "steps": info["steps"], "packed": info["packed"],
            "collectives_per_cycle": info["collectives"], "exchanged_bytes_per_cycle": info["exchanged_slots"] * 1280,
            "model_s_per_cycle": info["model_ms"] / 1e3, "gpu_launches_per_cycle": launches / max(cycles, 1),
            "cuda_graph": os.environ.get("B200FHE_NO_GRAPH", "0") != "1", "bind_s": bind_s,
            "timing": "CUDA events on the library stream around the clock loop (reset pass excluded), max over ranks",
        }
        line = _finish_line(line, name, tools, bp, req, enc, res_enc, cycles, lambda: nl, cpu)
    barrier()
    fe.eng.close()
    return line
