"""Flat gate-level netlists and the levelised engine (ctypes binding of include/b200net.h).

A netlist is four flat arrays over node ids (kind, in0, in1, in2) plus named ports.  This mirrors what
Iyokan's NetworkBuilder produces after the blueprint's [connect] edges have merged every sub-network
into one DAG (src/iyokan_tfhepp.cpp:428-435).  The C++ engine
(iyokan_b200/host/b200net.cpp) validates, levelises and replays it: one `b200fhe_gate_batch` per level,
one `b200fhe_dff_tick` per clock.
"""
from __future__ import annotations

import ctypes
import json
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from .lib import FLAVOUR, MU0, OPS, T0, TLWE0_LEN, B200FheError, Context

HOST = Path(__file__).resolve().parent / "host"
NET_LIB = HOST / f"libb200net{FLAVOUR}.so"

INPUT, DFF, OUTPUT = 32, 33, 34


class NetBuilder:
    """Tiny builder with the same vocabulary as Iyokan's NetworkBuilder (AND(), MUX(), DFF(), INPUT(...))."""

    def __init__(self):
        self.kind, self.ins = [], []
        self.in_ports: dict[str, list[int]] = {}
        self.out_ports: dict[str, list[int]] = {}
        self.mem: dict[str, list[int]] = {}
        self.write_through: list[str] = []

    def _add(self, kind, *ins):
        self.kind.append(kind)
        self.ins.append(list(ins) + [-1] * (3 - len(ins)))
        return len(self.kind) - 1

    def gate(self, name, *ins):
        return self._add(OPS[name], *ins)

    def input(self, port, bit):
        n = self._add(INPUT)
        self.in_ports.setdefault(port, [])
        lst = self.in_ports[port]
        lst.extend([-1] * (bit + 1 - len(lst)))
        lst[bit] = n
        return n

    def output(self, port, bit, src):
        n = self._add(OUTPUT, src)
        lst = self.out_ports.setdefault(port, [])
        lst.extend([-1] * (bit + 1 - len(lst)))
        lst[bit] = n
        return n

    def dff(self, d=-1, mem=None, index=None):
        n = self._add(DFF, d)
        if mem is not None:
            lst = self.mem.setdefault(mem, [])
            lst.extend([-1] * (index + 1 - len(lst)))
            lst[index] = n
        return n

    def set_dff_input(self, dff, d):
        self.ins[dff][0] = d

    def alias(self, node, src):
        """Turn an INPUT node into a wire fed by `src` (what a blueprint [connect] edge does)."""
        self.kind[node] = OUTPUT
        self.ins[node] = [src, -1, -1]

    def build(self) -> "Netlist":
        ins = np.array(self.ins, dtype=np.int32).reshape(-1, 3)
        return Netlist(np.array(self.kind, np.uint8), ins[:, 0].copy(), ins[:, 1].copy(), ins[:, 2].copy(),
                       {k: list(v) for k, v in self.in_ports.items()}, {k: list(v) for k, v in self.out_ports.items()},
                       {k: list(v) for k, v in self.mem.items()}, list(self.write_through))


@dataclass
class Netlist:
    kind: np.ndarray
    in0: np.ndarray
    in1: np.ndarray
    in2: np.ndarray
    in_ports: dict = field(default_factory=dict)    # external input port -> INPUT node per bit
    out_ports: dict = field(default_factory=dict)   # external output port -> node per bit
    mem: dict = field(default_factory=dict)         # "ram"/"rom" name -> node per memory bit (DFF or INPUT)
    # RAMs declared as CMUX memories in the blueprint: the reference updates them DURING the cycle (the result packet
    # shows the last cycle's write), whereas a MUX RAM cell changes at the next tick; for these names the result is
    # read from the cells' next-state inputs (see Frontend.result)
    write_through: list = field(default_factory=list)

    @property
    def n(self):
        return int(self.kind.size)

    def save(self, path):
        meta = json.dumps({"in_ports": self.in_ports, "out_ports": self.out_ports, "mem": self.mem,
                           "write_through": self.write_through})
        np.savez_compressed(path, kind=self.kind, in0=self.in0, in1=self.in1, in2=self.in2,
                            meta=np.frombuffer(meta.encode(), dtype=np.uint8))

    @staticmethod
    def load(path) -> "Netlist":
        z = np.load(path)
        meta = json.loads(bytes(z["meta"]).decode())
        return Netlist(z["kind"], z["in0"], z["in1"], z["in2"], meta["in_ports"], meta["out_ports"], meta["mem"],
                       meta.get("write_through", []))


# ---- generators (self-contained circuits for tests; structure follows the reference's descriptions) ----

def mux_rom(b: NetBuilder, addr_nodes, width, name="rom"):
    """MUX-tree ROM: per output bit, 2^A leaves reduced by A levels of MUXes selected by addr bit i
    (circuit described in src/iyokan.hpp:2538-2593). ROM bit index = word * width + bit."""
    a = len(addr_nodes)
    outs = []
    for bit in range(width):
        work = [b.input(f"{name}/romdata", bit + w * width) for w in range(1 << a)]
        for i in range(a):
            work = [b.gate("MUX", work[j], work[j + 1], addr_nodes[i]) for j in range(0, len(work), 2)]
        outs.append(work[0])
    b.mem[name] = list(b.in_ports.pop(f"{name}/romdata"))
    return outs


def mux_ram(b: NetBuilder, addr_nodes, wren, wdata_nodes, name="ram"):
    """MUX RAM (circuit described in src/iyokan.hpp:2646-2762): a DMUX tree decodes wren by address
    (out0 = ANDNOT(in, sel), out1 = AND(in, sel)), each cell is a DFF holding MUX(hold, wdata, we),
    the read port is a MUX tree.  RAM bit index = addr * width + bit."""
    a, width = len(addr_nodes), len(wdata_nodes)
    we = [wren]
    for i in reversed(range(a)):   # most significant address bit first so that leaf order == address
        we = [g for w in we for g in (b.gate("ANDNOT", w, addr_nodes[i]), b.gate("AND", w, addr_nodes[i]))]
    outs = []
    for bit in range(width):
        cells = []
        for ad in range(1 << a):
            q = b.dff(mem=name, index=ad * width + bit)
            b.set_dff_input(q, b.gate("MUX", q, wdata_nodes[bit], we[ad]))
            cells.append(q)
        work = cells
        for i in range(a):
            work = [b.gate("MUX", work[j], work[j + 1], addr_nodes[i]) for j in range(0, len(work), 2)]
        outs.append(work[0])
    return outs


def ripple_adder(width: int) -> Netlist:
    b = NetBuilder()
    a = [b.input("a", i) for i in range(width)]
    c = [b.input("b", i) for i in range(width)]
    carry = None
    for i in range(width):
        x = b.gate("XOR", a[i], c[i])
        if carry is None:
            s, carry = x, b.gate("AND", a[i], c[i])
        else:
            s = b.gate("XOR", x, carry)
            carry = b.gate("OR", b.gate("AND", a[i], c[i]), b.gate("AND", x, carry))
        b.output("sum", i, s)
    b.output("sum", width, carry)
    return b.build()


def counter(width: int) -> Netlist:
    """Synchronous counter with reset, DFF state: q <- reset ? 0 : q + 1."""
    b = NetBuilder()
    rst = b.input("reset", 0)
    qs = [b.dff() for _ in range(width)]
    carry = None
    for i, q in enumerate(qs):
        nxt = b.gate("NOT", q) if carry is None else b.gate("XOR", q, carry)
        carry = q if carry is None else b.gate("AND", q, carry)
        b.set_dff_input(q, b.gate("ANDNOT", nxt, rst))
        b.output("out", i, q)
    return b.build()


# ---- ctypes binding of include/b200net.h ----

_net_lib = None


def load_net():
    global _net_lib
    if _net_lib is None:
        if not NET_LIB.exists():
            raise B200FheError(f"{NET_LIB} is missing: run __graft_entry__.build()")
        from .lib import load as load_fhe

        load_fhe()  # libb200net.so links against libb200fhe.so
        lib = ctypes.CDLL(str(NET_LIB))
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        sig = {
            "b200net_create": (ci, [ctypes.POINTER(vp), sz, vp, vp, vp, vp]),
            "b200net_destroy": (None, [vp]),
            "b200net_last_error": (ctypes.c_char_p, []),
            "b200net_num_nodes": (sz, [vp]), "b200net_num_levels": (sz, [vp]),
            "b200net_level_width": (sz, [vp, sz]), "b200net_level_bootstraps": (sz, [vp, sz]),
            "b200net_bootstraps_per_cycle": (sz, [vp]),
            "b200net_num_dff": (sz, [vp]), "b200net_node_level": (ctypes.c_int32, [vp, sz]),
            "b200net_slot_of": (ctypes.c_uint32, [vp, sz]), "b200net_num_slots": (sz, [vp]),
            "b200net_level_slot_base": (ctypes.c_uint32, [vp, sz]),
            "b200net_plain_eval": (ci, [vp, vp]), "b200net_plain_tick": (ci, [vp, vp]),
            "b200net_layout": (ci, [vp, ci]), "b200net_bind": (ci, [vp, vp, ci]), "b200net_set": (ci, [vp, vp, vp, sz]), "b200net_restore": (ci, [vp, vp, vp, sz]),
            "b200net_get": (ci, [vp, vp, vp, sz]), "b200net_tick": (ci, [vp]), "b200net_run": (ci, [vp]),
            "b200net_run_level_shard": (ci, [vp, sz, ci, ci]),
            "b200net_bind_rank": (ci, [vp, vp, ci, ci, ctypes.c_uint]), "b200net_schedule": (ci, [vp, ci, ctypes.c_uint]),
            "b200net_num_steps": (sz, [vp]), "b200net_step_jobs": (sz, [vp, sz, vp]),
            "b200net_schedule_info": (ci, [vp, vp, vp, vp, vp, vp]),
            "b200net_step_gates": (sz, [vp, sz, ci, vp, sz]), "b200net_step_exchange": (ci, [vp, sz, vp, vp]),
            "b200net_profile_run": (ci, [vp, vp, sz]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _net_lib = lib
    return _net_lib


NET_EXPORTS = [
    "b200net_create", "b200net_destroy", "b200net_last_error", "b200net_num_nodes", "b200net_num_levels",
    "b200net_level_width", "b200net_level_bootstraps", "b200net_bootstraps_per_cycle", "b200net_num_dff", "b200net_node_level",
    "b200net_slot_of", "b200net_num_slots", "b200net_level_slot_base", "b200net_plain_eval", "b200net_plain_tick",
    "b200net_layout", "b200net_bind", "b200net_set", "b200net_restore", "b200net_get", "b200net_tick", "b200net_run", "b200net_run_level_shard",
    "b200net_bind_rank", "b200net_schedule", "b200net_num_steps", "b200net_step_jobs", "b200net_schedule_info",
    "b200net_step_gates", "b200net_step_exchange", "b200net_profile_run",
]


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


class NetEngine:
    """Levelised netlist: structure queries, plaintext back-end, and the encrypted back-end on one GPU."""

    def __init__(self, nl: Netlist):
        self.lib = load_net()
        self.nl = nl
        self._k = np.ascontiguousarray(nl.kind, np.uint8)
        self._i = [np.ascontiguousarray(x, np.int32) for x in (nl.in0, nl.in1, nl.in2)]
        h = ctypes.c_void_p()
        if self.lib.b200net_create(ctypes.byref(h), nl.n, _p(self._k), _p(self._i[0]), _p(self._i[1]), _p(self._i[2])):
            raise B200FheError(self.lib.b200net_last_error().decode())
        self._h = h
        self.ctx = None
        self.world = 1

    def _ck(self, rc):
        if rc:
            raise B200FheError(self.lib.b200net_last_error().decode())

    def close(self):
        if self._h is not None:
            self.lib.b200net_destroy(self._h)
            self._h = None

    # structure
    @property
    def num_levels(self):
        return int(self.lib.b200net_num_levels(self._h))

    @property
    def level_widths(self):
        return [int(self.lib.b200net_level_width(self._h, l)) for l in range(self.num_levels)]

    @property
    def level_bootstraps(self):
        return [int(self.lib.b200net_level_bootstraps(self._h, l)) for l in range(self.num_levels)]

    @property
    def bootstraps_per_cycle(self):
        return int(self.lib.b200net_bootstraps_per_cycle(self._h))

    @property
    def num_dff(self):
        return int(self.lib.b200net_num_dff(self._h))

    @property
    def num_slots(self):
        return int(self.lib.b200net_num_slots(self._h))

    def slot_of(self, node):
        return int(self.lib.b200net_slot_of(self._h, node))

    def level_slot_base(self, level):
        return int(self.lib.b200net_level_slot_base(self._h, level))

    # plaintext back-end
    def plain_eval(self, values: np.ndarray):
        self._ck(self.lib.b200net_plain_eval(self._h, _p(values)))

    def plain_tick(self, values: np.ndarray):
        self._ck(self.lib.b200net_plain_tick(self._h, _p(values)))

    def layout(self, world_size: int = 1):
        self.world = world_size
        self._ck(self.lib.b200net_layout(self._h, world_size))

    # encrypted back-end
    def bind(self, ctx: Context, world_size: int = 1):
        self.ctx, self.world = ctx, world_size
        self._ck(self.lib.b200net_bind(self._h, ctx._h, world_size))

    PACK = 1

    def bind_rank(self, ctx: Context, rank: int = 0, world_size: int = 1, flags: int = 1):
        """Schedule for `world_size` ranks and compile this rank's share into CUDA graphs (clock, tick)."""
        self.ctx, self.world = ctx, world_size
        self._ck(self.lib.b200net_bind_rank(self._h, ctx._h, rank, world_size, flags))

    def schedule(self, world_size: int = 1, flags: int = 1):
        """The static schedule alone (no GPU): steps, sharding decisions, slots."""
        self.world = world_size
        self._ck(self.lib.b200net_schedule(self._h, world_size, flags))

    @property
    def step_jobs(self):
        out = []
        for k in range(int(self.lib.b200net_num_steps(self._h))):
            sh = ctypes.c_int()
            out.append((int(self.lib.b200net_step_jobs(self._h, k, ctypes.byref(sh))), bool(sh.value)))
        return out

    @property
    def num_steps(self):
        return int(self.lib.b200net_num_steps(self._h))

    def step_gates(self, step: int, rank: int) -> np.ndarray:
        n = int(self.lib.b200net_step_gates(self._h, step, rank, None, 0))
        out = np.zeros(max(n, 1), np.uint32)
        self.lib.b200net_step_gates(self._h, step, rank, _p(out), n)
        return out[:n]

    def step_exchange(self, step: int):
        first, per = ctypes.c_uint32(), ctypes.c_size_t()
        self._ck(self.lib.b200net_step_exchange(self._h, step, ctypes.byref(first), ctypes.byref(per)))
        return int(first.value), int(per.value)

    def schedule_info(self) -> dict:
        st, co, ex = (ctypes.c_size_t() for _ in range(3))
        ms, pk = ctypes.c_double(), ctypes.c_int()
        self._ck(self.lib.b200net_schedule_info(self._h, ctypes.byref(st), ctypes.byref(co), ctypes.byref(ex), ctypes.byref(ms),
                                                ctypes.byref(pk)))
        return {"steps": st.value, "collectives": co.value, "exchanged_slots": ex.value, "model_ms": ms.value,
                "packed": bool(pk.value)}

    def set(self, nodes, tlwe):
        nodes = np.ascontiguousarray(nodes, np.uint32)
        tlwe = np.ascontiguousarray(tlwe, T0)
        assert tlwe.size == nodes.size * TLWE0_LEN
        self._ck(self.lib.b200net_set(self._h, _p(nodes), _p(tlwe), nodes.size))

    def restore(self, nodes, tlwe):
        nodes = np.ascontiguousarray(nodes, np.uint32)
        tlwe = np.ascontiguousarray(tlwe, T0)
        assert tlwe.size == nodes.size * TLWE0_LEN
        self._ck(self.lib.b200net_restore(self._h, _p(nodes), _p(tlwe), nodes.size))

    def get(self, nodes) -> np.ndarray:
        nodes = np.ascontiguousarray(nodes, np.uint32)
        out = np.empty((nodes.size, TLWE0_LEN), T0)
        self._ck(self.lib.b200net_get(self._h, _p(nodes), _p(out), nodes.size))
        return out

    def tick(self):
        self._ck(self.lib.b200net_tick(self._h))

    def run(self):
        self._ck(self.lib.b200net_run(self._h))

    def profile_run(self) -> np.ndarray:
        """One evaluation of the clock, step by step with device timers: ms per schedule step on this rank."""
        ms = np.zeros(self.num_steps, np.float32)
        self._ck(self.lib.b200net_profile_run(self._h, _p(ms), ms.size))
        return ms

    def run_level_shard(self, level, rank, world):
        self._ck(self.lib.b200net_run_level_shard(self._h, level, rank, world))


def trivial(bit) -> np.ndarray:
    """Noiseless ciphertext (0,...,0, +-mu): what Iyokan feeds for reset/constants (iyokan_tfhepp.hpp:23-27)."""
    t = np.zeros(TLWE0_LEN, T0)
    t[-1] = MU0 if bit else (-MU0) & ((1 << (8 * np.dtype(T0).itemsize)) - 1)
    return t


def bits_of(byte_list, nbits):
    """Packet bit order: bit i = (bytes[i // 8] >> (i % 8)) & 1 (src/packet.hpp)."""
    out = np.zeros(nbits, np.uint8)
    for i in range(min(nbits, 8 * len(byte_list))):
        out[i] = (byte_list[i // 8] >> (i % 8)) & 1
    return out


def bytes_of(bits):
    bits = np.asarray(bits, np.uint8)
    out = [0] * ((bits.size + 7) // 8)
    for i, b in enumerate(bits):
        out[i // 8] |= int(b) << (i % 8)
    return out


class PlainRunner:
    """Cycle protocol of TFHEppFrontend::go / doPlain on the plaintext back-end.

    Test helper for netlists made with NetBuilder; the command-line protocol (TOGND bits, write-through RAM images,
    snapshots, length checks with the reference's messages) is `frontend.Frontend`."""

    def __init__(self, nl: Netlist, eng: NetEngine | None = None):
        self.nl, self.eng = nl, eng or NetEngine(nl)
        self.v = np.zeros(nl.n, np.uint8)

    def set_mem(self, name, bits):
        nodes = self.nl.mem[name]
        bits = np.asarray(bits, np.uint8)
        if bits.size != len(nodes):
            raise ValueError(f"memory {name}: {bits.size} bits given, {len(nodes)} cells")
        self.v[np.array(nodes)] = bits

    def get_mem(self, name):
        return self.v[np.array(self.nl.mem[name])].copy()

    def _set_port(self, port, bits):
        for node, b in zip(self.nl.in_ports[port], bits):
            if node >= 0:
                self.v[node] = b

    def run(self, cycles, inputs=None, rams=None, roms=None):
        """inputs: port -> bit stream (value of bit b at cycle c = stream[(width*c + b) % len])."""
        inputs = inputs or {}
        for name, bits in (roms or {}).items():
            self.set_mem(name, bits)
        has_reset = "reset" in self.nl.in_ports
        if has_reset:  # reset pass: reset <- 1, run once
            self._set_port("reset", [1])
            self.eng.plain_eval(self.v)
        for c in range(cycles):
            self.eng.plain_tick(self.v)
            if c == 0:
                if has_reset:
                    self._set_port("reset", [0])
                for name, bits in (rams or {}).items():
                    self.set_mem(name, bits)
            for port, stream in inputs.items():
                w = len(self.nl.in_ports[port])
                self._set_port(port, [stream[(w * c + b) % len(stream)] for b in range(w)])
            self.eng.plain_eval(self.v)
        # an unconnected (TOGND) bit of an output port reads 0
        return {p: np.array([self.v[n] if n >= 0 else 0 for n in nodes], np.uint8) for p, nodes in self.nl.out_ports.items()}


class EncryptedRunner:
    """Same cycle protocol on the GPU back-end (TFHEppFrontend::go, src/iyokan_tfhepp.cpp:465-566).  Like PlainRunner a
    helper for builder-made netlists and all-gate packets; write-through (CMUX) RAM images are `frontend.Frontend`'s job.

    Inputs are CIPHERTEXTS: `inputs[port]` is a TLWE stream [size][637] uint16 (what TFHEPacket.bits holds),
    `rams` / `roms` are [bits][637] arrays (TFHEPacket.ramInTLWE / romInTLWE).  For tests that start from
    plaintext, pass `encrypt(bits) -> [n][637]` and plain bit arrays instead; key material never enters the
    engine, as in Iyokan where iyokan-packet encrypts and decrypts."""

    def __init__(self, nl: Netlist, ctx: Context, encrypt=None, eng: NetEngine | None = None):
        self.nl, self.ctx, self.encrypt = nl, ctx, encrypt
        self.eng = eng or NetEngine(nl)
        self.eng.bind_rank(ctx, 0, 1)
        dffs = np.nonzero(nl.kind == DFF)[0]
        if dffs.size:  # DFF initial value: trivial 0 (iyokan_tfhepp.hpp:23-27)
            self.eng.set(dffs, np.tile(trivial(0), (dffs.size, 1)))

    def _ct(self, x):
        x = np.asarray(x)
        if x.dtype == T0 and x.ndim == 2 and x.shape[1] == TLWE0_LEN:
            return x
        if self.encrypt is None:
            raise ValueError("plaintext bits given but no encrypt callback")
        return self.encrypt(np.asarray(x, np.uint8))

    def _set_nodes(self, nodes, ct, what=None):
        if what is not None and len(ct) != len(nodes):
            raise ValueError(f"memory {what}: {len(ct)} ciphertexts given, {len(nodes)} cells")
        keep = [i for i, n in enumerate(nodes) if n >= 0 and i < len(ct)]
        if keep:
            self.eng.set(np.array([nodes[i] for i in keep], np.uint32), np.ascontiguousarray(ct[keep]))

    def run(self, cycles, inputs=None, rams=None, roms=None):
        inputs = {p: self._ct(v) for p, v in (inputs or {}).items()}
        for name, v in (roms or {}).items():
            self._set_nodes(self.nl.mem[name], self._ct(v), what=name)
        has_reset = "reset" in self.nl.in_ports
        if has_reset:
            # reset pass (iyokan_tfhepp.cpp:487-500): reset <- trivial 1, every other input still holds
            # the all-zero TLWE a default-constructed Task has (the arena is zero-initialised)
            self._set_nodes(self.nl.in_ports["reset"], trivial(1)[None])
            self.eng.run()
        for c in range(cycles):
            self.eng.tick()
            if c == 0:
                if has_reset:
                    self._set_nodes(self.nl.in_ports["reset"], trivial(0)[None])
                for name, v in (rams or {}).items():
                    self._set_nodes(self.nl.mem[name], self._ct(v), what=name)
            for port, stream in inputs.items():  # setCircularInputs, iyokan_tfhepp.cpp:274-296
                w = len(self.nl.in_ports[port])
                idx = [(w * c + b) % len(stream) for b in range(w)]
                self._set_nodes(self.nl.in_ports[port], stream[idx])
            self.eng.run()
        self.ctx.sync()
        return {p: self._get_nodes(nodes) for p, nodes in self.nl.out_ports.items()}

    def _get_nodes(self, nodes):
        """Ciphertexts of `nodes`; an unconnected (TOGND, -1) bit reads as the trivial 0 the reference leaves there."""
        out = np.tile(trivial(0), (len(nodes), 1))
        keep = [i for i, n in enumerate(nodes) if n >= 0]
        if keep:
            out[keep] = self.eng.get(np.array([nodes[i] for i in keep], np.uint32))
        return out

    def get_mem(self, name):
        return self._get_nodes(self.nl.mem[name])


def run_packet(nl: Netlist, ctx: Context, req, cycles: int | None = None):
    """`iyokan tfhe -i req -o res` for an all-gate blueprint: TFHEPacket in, TFHEPacket out
    (TFHEppFrontend::go + makeResPacket, src/iyokan_tfhepp.cpp:176-227,465-566)."""
    from .packet import TFHEPacket

    n = cycles if cycles is not None else req.num_cycles
    if n is None or n < 0:
        raise ValueError("number of cycles must be given (packet has none)")
    for name in list(req.bits):
        if name not in nl.in_ports:
            raise ValueError(f"request packet drives unknown input port @{name}")
    if nl.write_through:
        raise ValueError("blueprint declares CMUX RAMs (write-through images): run it through frontend.Frontend")
    runner = EncryptedRunner(nl, ctx)
    rams = {k: v for k, v in req.ram_in_tlwe.items() if k in nl.mem}
    roms = {k: v for k, v in req.rom_in_tlwe.items() if k in nl.mem}
    out = runner.run(n, inputs=req.bits, rams=rams, roms=roms)
    res = TFHEPacket(bits=out, num_cycles=n)
    for name, nodes in nl.mem.items():
        if nl.kind[nodes[0]] == DFF:  # RAM cells are part of the result packet, ROM is not
            res.ram_in_tlwe[name] = runner.get_mem(name)
    return res
