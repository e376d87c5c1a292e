"""Front end: the cycle protocol of `iyokan plain` / `iyokan tfhe` over a flat netlist, with snapshot / resume.

Mirrors TFHEppFrontend / PlainFrontend (src/iyokan_tfhepp.cpp:116-573, src/iyokan_plain.cpp:115-; SURVEY.md
Appendix B); memories are gate-level MUX ROM / RAM (a blueprint's CMUX memories are built the same way):
  * request packet -> ROM contents at construction, RAM contents and `@`-port bit streams from the first cycle on
    (setInitialRAM / setCircularInputs, iyokan_tfhepp.cpp:274-296): bit b of port p at cycle c =
    stream[(width * c + b) mod size];
  * optional reset pass (reset <- 1, one run), then per cycle: tick (every DFF / RAM cell Q <- D), reset <- 0 on
    the first cycle, inputs, run; no tick after the last run (iyokan_tfhepp.cpp:486-561);
  * result packet = OUTPUT ports + RAM cells as they are, numCycles = cycles executed so far (:176-227);
  * `--snapshot` / `--resume` (iyokan_tfhepp.cpp:568-572,603-617): the reference serialises its whole Task graph
    with cereal; here the state is the netlist plus one value (bit or TLWE) per node, the cycle counter and the
    request streams, stored as one .npz.  A run of N cycles equals a run of k cycles, snapshot, resume for N - k.
The two back-ends share every line of the protocol: `plain` evaluates bits on the host engine
(b200net_plain_eval, the reference's plain back-end), `tfhe` evaluates ciphertexts on the GPU through the
C ABI (b200net_run -> b200fhe_gate_batch).  Key material never enters this module.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from .lib import TLWE0_LEN
from .netlist import DFF, INPUT, NetEngine, Netlist, trivial
from .packet import PlainPacket, TFHEPacket

SNAPSHOT_MAGIC = "iyokan_b200-snapshot-1"


class FrontendError(ValueError):
    pass


class Frontend:
    def __init__(self, nl: Netlist, mode: str, ctx=None, rank: int = 0, world: int = 1, group=None):
        """rank / world / group: one process per GPU under torch.distributed (`torchrun -m iyokan_b200 tfhe ...`):
        every rank holds the whole state, a step of the static schedule is split between the ranks when that pays and
        its outputs are all-gathered (b200net_bind_rank; iyokan_b200/shard.py); rank 0 writes the result."""
        if mode not in ("plain", "tfhe"):
            raise FrontendError("mode must be 'plain' or 'tfhe'")
        if mode == "tfhe" and ctx is None:
            raise FrontendError("tfhe mode needs a b200fhe Context with keys loaded (there is no CPU fallback)")
        self.nl, self.mode, self.ctx = nl, mode, ctx
        self.rank, self.world = rank, world
        self.eng = NetEngine(nl)
        self.cycle = 0            # clock cycles executed so far
        self.reset_done = False
        self.streams: dict = {}   # port -> bits / TLWEs of the request packet
        self.rams: dict = {}
        self._all = np.arange(nl.n, dtype=np.uint32)
        self.be = self.runner = None
        if world > 1 and mode == "plain":
            from .shard import PlainBackend, ScheduledRunner

            self.be = PlainBackend(nl, self.eng, world)
            self.runner = ScheduledRunner(nl, self.eng, self.be, rank, world, group=group)
        elif mode == "plain":
            self.v = np.zeros(nl.n, np.uint8)
        else:  # one CUDA graph per clock; for world > 1 the all-gathers of the sharded steps are inside it
            from .shard import setup_comm

            setup_comm(ctx, rank, world, group)
            self.eng.bind_rank(ctx, rank, world)
        if mode == "tfhe":
            dffs = np.nonzero(nl.kind == DFF)[0]
            if dffs.size:  # DFF initial value: trivial 0 (iyokan_tfhepp.hpp:23-27)
                self.eng.set(dffs, np.tile(trivial(0), (dffs.size, 1)))

    # ---- state access, uniform over the two back-ends ----
    def _set(self, nodes, values):
        keep = [i for i, n in enumerate(nodes) if n >= 0 and i < len(values)]
        if not keep:
            return
        idx = np.array([nodes[i] for i in keep], np.uint32)
        if self.mode == "plain":
            vals = np.asarray(values, np.uint8)[keep]
            if self.be is not None:
                self.be.set_nodes(idx, vals)
            else:
                self.v[idx] = vals
        else:
            self.eng.set(idx, np.ascontiguousarray(np.asarray(values, np.uint16)[keep]))

    def _get(self, nodes):
        nodes = np.asarray(nodes, np.int64)
        idx = np.maximum(nodes, 0).astype(np.uint32)
        if self.mode == "plain":
            out = self.be.get_nodes(idx) if self.be is not None else self.v[idx].copy()
        else:
            out = self.eng.get(idx)
        out[nodes < 0] = 0 if self.mode == "plain" else trivial(0)  # port bits tied to ground (blueprint TOGND)
        return out

    def _const(self, bit):
        return np.array([bit], np.uint8) if self.mode == "plain" else trivial(bit)[None]

    def _eval(self):
        if self.runner is not None:
            self.runner.run()
        elif self.mode == "plain":
            self.eng.plain_eval(self.v)
        else:
            self.eng.run()

    def _tick(self):
        if self.runner is not None:
            self.runner.tick()
        elif self.mode == "plain":
            self.eng.plain_tick(self.v)
        else:
            self.eng.tick()

    # ---- request ----
    def load_request(self, req):
        want = PlainPacket if self.mode == "plain" else TFHEPacket
        if not isinstance(req, want):
            raise FrontendError(f"{self.mode} mode takes a {want.__name__}")
        bits = req.bits
        rams = req.ram if self.mode == "plain" else req.ram_in_tlwe
        roms = req.rom if self.mode == "plain" else req.rom_in_tlwe
        if self.mode == "tfhe":
            # `iyokan-packet enc` writes every memory twice: as TRLWEs (for CMUX memories) and as TLWEs (for MUX
            # memories, src/packet.hpp:208-220); this back-end runs mux-rom / mux-ram blueprints and uses the latter
            for name in list(req.ram) + list(req.rom):
                if name in self.nl.mem and name not in rams and name not in roms:
                    raise FrontendError(f"memory {name!r} is given only in CMUX (TRLWE) form")
        for name in bits:
            if len(bits[name]) == 0:
                raise FrontendError(f"request gives an empty bit stream for @{name}")
            if name == "reset":
                raise FrontendError("@reset cannot be set by the request (iyokan_tfhepp.cpp:284-285)")
            if name not in self.nl.in_ports:
                raise FrontendError(f"request drives unknown input port @{name}")
        for kind, mems in (("RAM", rams), ("ROM", roms)):
            for name, v in mems.items():
                if name not in self.nl.mem:
                    raise FrontendError(f"request initialises unknown memory {name!r}")
                if len(v) != len(self.nl.mem[name]):  # iyokan_tfhepp.cpp:242-259 ("wrong length of RAM")
                    raise FrontendError(f"Invalid request packet: wrong length of {kind} {name!r} "
                                        f"({len(v)} bits given, the memory holds {len(self.nl.mem[name])})")
        self.streams = {k: np.asarray(v) for k, v in bits.items()}
        self.rams = {k: np.asarray(v) for k, v in rams.items() if self.nl.kind[self.nl.mem[k][0]] == DFF}
        for name, v in roms.items():
            self._set(self.nl.mem[name], np.asarray(v))
        for name, v in rams.items():  # a "RAM" made of INPUT wires is a ROM in disguise
            if name not in self.rams:
                self._set(self.nl.mem[name], np.asarray(v))

    # ---- protocol ----
    def run(self, cycles: int, skip_reset: bool = False, dump_prefix: str | None = None):
        """dump_prefix: before every cycle c the state reached so far is written as a result packet to `<prefix>-<c>`
        (c = cycles executed: `-0` is the state after the reset pass; the final state is the -o packet), as
        iyokan --dump-prefix does (src/iyokan_tfhepp.cpp:520-533).  The reference decrypts the dumps with
        --secret-key; here they stay encrypted in tfhe mode - key material never enters the back-end - and
        `iyokan-packet dec` opens them."""
        if cycles < 0:
            raise FrontendError("number of cycles must be >= 0")
        has_reset = "reset" in self.nl.in_ports
        if self.cycle == 0 and not self.reset_done and has_reset and not skip_reset:
            self._set(self.nl.in_ports["reset"], self._const(1))
            self._eval()
        self.reset_done = True
        for _ in range(cycles):
            c = self.cycle
            if dump_prefix and self.rank == 0:
                self.result().save(f"{dump_prefix}-{c}")
            self._tick()
            if c == 0:
                if has_reset:
                    self._set(self.nl.in_ports["reset"], self._const(0))
                for name, v in self.rams.items():
                    self._set(self.nl.mem[name], v)
            for port, stream in self.streams.items():
                nodes = self.nl.in_ports[port]
                w = len(nodes)
                self._set(nodes, stream[[(w * c + b) % len(stream) for b in range(w)]])
            self._eval()
            self.cycle += 1
        if self.mode == "tfhe":
            self.ctx.sync()

    def result(self):
        """Result packet: OUTPUT ports, RAM cells (ROM is not returned), numCycles."""
        out = {p: self._get(nodes) for p, nodes in self.nl.out_ports.items()}
        ram = {}
        for m, nodes in self.nl.mem.items():
            if self.nl.kind[nodes[0]] != DFF:
                continue
            if m in self.nl.write_through and self.cycle > 0:
                # CMUX RAM semantics: the write of the last cycle is already part of the memory image; in the MUX
                # formulation that value sits on the cells' D inputs until the next tick
                nodes = [int(self.nl.in0[n]) for n in nodes]
            ram[m] = self._get(nodes)
        if self.mode == "plain":
            return PlainPacket(ram=ram, bits=out, num_cycles=self.cycle)
        return TFHEPacket(ram_in_tlwe=ram, bits=out, num_cycles=self.cycle)

    # ---- snapshot / resume ----
    def snapshot(self, path):
        meta = {"magic": SNAPSHOT_MAGIC, "mode": self.mode, "cycle": self.cycle, "reset_done": self.reset_done,
                "in_ports": self.nl.in_ports, "out_ports": self.nl.out_ports, "mem": self.nl.mem,
                "write_through": self.nl.write_through,
                "streams": sorted(self.streams), "rams": sorted(self.rams)}
        arrays = {"kind": self.nl.kind, "in0": self.nl.in0, "in1": self.nl.in1, "in2": self.nl.in2,
                  "state": self._get(self._all),
                  "meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
        for k, v in self.streams.items():
            arrays[f"stream__{k}"] = v
        for k, v in self.rams.items():
            arrays[f"ram__{k}"] = v
        with open(path, "wb") as f:  # np.savez would append ".npz" to a bare path
            np.savez_compressed(f, **arrays)

    @staticmethod
    def is_snapshot(path) -> bool:
        try:
            with np.load(path) as z:
                return json.loads(bytes(z["meta"]).decode()).get("magic") == SNAPSHOT_MAGIC
        except Exception:  # noqa: BLE001
            return False

    @staticmethod
    def snapshot_mode(path) -> str:
        with np.load(path) as z:
            return json.loads(bytes(z["meta"]).decode())["mode"]

    @staticmethod
    def resume(path, ctx=None, rank: int = 0, world: int = 1, group=None) -> "Frontend":
        if not Frontend.is_snapshot(path):
            raise FrontendError(f"Invalid resume file: {path}")
        z = np.load(path)
        meta = json.loads(bytes(z["meta"]).decode())
        nl = Netlist(z["kind"], z["in0"], z["in1"], z["in2"], meta["in_ports"], meta["out_ports"], meta["mem"],
                     meta.get("write_through", []))
        fe = Frontend(nl, meta["mode"], ctx, rank, world, group)
        fe.cycle, fe.reset_done = int(meta["cycle"]), bool(meta["reset_done"])
        fe.streams = {k: z[f"stream__{k}"] for k in meta["streams"]}
        fe.rams = {k: z[f"ram__{k}"] for k in meta["rams"]}
        state = z["state"]
        real = np.nonzero(nl.kind != 34)[0].astype(np.uint32)  # OUTPUT wires alias their drivers
        if fe.mode == "plain":
            if fe.be is not None:
                fe.be.set_nodes(real, state[real].astype(np.uint8))
            else:
                fe.v = state.astype(np.uint8).copy()
        else:
            if state.shape != (nl.n, TLWE0_LEN):
                raise FrontendError("snapshot state has the wrong shape")
            fe.eng.restore(real, np.ascontiguousarray(state[real]))
        return fe


__all__ = ["Frontend", "FrontendError", "INPUT"]
