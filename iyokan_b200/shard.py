"""Multi-GPU evaluation of one netlist: level-sliced sharding + one all-gather per level.

The reference's multi-GPU mode replicates the keys and round-robins gates over devices, moving every
ciphertext through host memory around each gate (cufhe_gpu.cuh:164-169, cufhe_gates_gpu.cu:145-157);
it has no collective.  Here: one process per GPU, keys and the slot arena replicated, each dependency
level's gates are split into `world` contiguous shares (b200net_run_level_shard), and the TLWEs a level
produces are exchanged with ONE all-gather over NVLink (NCCL through torch.distributed) before the next
level starts.  The exchange has to be per level, not per clock: the cuts cross combinational logic (the
processor core, ROM and RAM are coupled inside a cycle).  Payload per level is
width x 1280 B, so the step is latency bound; levels narrower than `min_shard_width` are computed
redundantly on every rank instead (no exchange at all), which is what keeps the narrow levels of a
processor core from paying one collective each.

The same orchestration runs on CPU tensors with the plaintext back-end and the gloo backend; that is how
the N > 1 path is tested without GPUs (tests/test_shard_gloo.py).
"""
from __future__ import annotations

import numpy as np

from .netlist import DFF, INPUT, OUTPUT, NetEngine, Netlist

SLOT = 640  # uint16 per arena slot


class PlainBackend:
    """Bits instead of ciphertexts, a CPU uint8 tensor as the 'arena' (one byte per slot)."""

    elem = 1

    def __init__(self, nl: Netlist, eng: NetEngine, world: int):
        import torch

        self.nl, self.eng = nl, eng
        eng.layout(world)
        self._slot = np.array([eng.slot_of(i) for i in range(nl.n)], dtype=np.int64)
        self.arena = torch.zeros(eng.num_slots, dtype=torch.uint8)
        self._np = self.arena.numpy()
        lv = np.array([eng.lib.b200net_node_level(eng._h, i) for i in range(nl.n)])
        self._level_nodes = [np.nonzero((lv == k + 1) & (nl.kind < 15))[0] for k in range(eng.num_levels)]
        # the engine assigns slots inside a level in node order
        for k, nodes in enumerate(self._level_nodes):
            assert np.array_equal(self._slot[nodes], eng.level_slot_base(k) + np.arange(nodes.size))

    def set_nodes(self, nodes, bits):
        self._np[self._slot[np.asarray(nodes)]] = np.asarray(bits, np.uint8)

    def get_nodes(self, nodes):
        return self._np[self._slot[np.asarray(nodes)]].copy()

    def run_level_shard(self, level, lo, hi):
        g = self._level_nodes[level][lo:hi]
        k = self.nl.kind[g]

        def val(arr):
            idx = arr[g]
            return np.where(idx >= 0, self._np[self._slot[np.maximum(idx, 0)]], 0).astype(np.uint8)

        a, b, s = val(self.nl.in0), val(self.nl.in1), val(self.nl.in2)
        table = [a & b, 1 - (a & b), a & (1 - b), a | b, 1 - (a | b), a | (1 - b), a ^ b, 1 - (a ^ b),
                 np.where(s == 1, b, a), 1 - a, a, np.zeros_like(a), np.ones_like(a), (1 - a) & b, (1 - a) | b]
        out = np.zeros_like(a)
        for op in range(15):
            out = np.where(k == op, table[op], out)
        self._np[self._slot[g]] = out

    def tick(self):
        d = np.nonzero(self.nl.kind == DFF)[0]
        src = self._slot[self.nl.in0[d]]
        self._np[self._slot[d]] = self._np[src].copy()

    def sync(self):
        pass

    def stream_ctx(self):
        import contextlib

        return contextlib.nullcontext()


class GpuBackend:
    """Ciphertexts on one GPU: the engine's arena wrapped as a torch int16 tensor (NCCL has no uint16) for the all-gather."""

    elem = SLOT * 2  # bytes: the arena is exposed as uint8 (ProcessGroupNCCL rejects uint16 / int16)

    def __init__(self, nl: Netlist, eng: NetEngine, ctx, world: int):
        import torch

        self.nl, self.eng, self.ctx = nl, eng, ctx
        eng.bind(ctx, world)
        n = eng.num_slots * SLOT * 2

        class _Arena:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ctx.arena_dev_ptr, False), "version": 2}

        self.arena = torch.as_tensor(_Arena(), device=torch.device("cuda", ctx.device))
        self._stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", ctx.device))
        self._torch = torch

    def set_nodes(self, nodes, tlwe):
        self.eng.set(np.asarray(nodes, np.uint32), tlwe)

    def get_nodes(self, nodes):
        return self.eng.get(np.asarray(nodes, np.uint32))

    def run_level_shard(self, level, lo, hi, rank=None, world=None):
        raise NotImplementedError  # sharding is done inside the engine, see ShardedRunner._compute

    def tick(self):
        self.eng.tick()

    def sync(self):
        self.ctx.sync()

    def stream_ctx(self):
        return self._torch.cuda.stream(self._stream)


class ShardedRunner:
    """Runs the levels of a bound netlist on `world` ranks.  All ranks hold identical arenas between levels."""

    def __init__(self, nl: Netlist, eng: NetEngine, backend, rank: int, world: int, group=None,
                 min_shard_width: int | None = 0, collective_ms: float = 0.1):
        """min_shard_width: levels with fewer gates are computed redundantly on every rank.  None selects the
        cost model instead: a level is split only when the launch plan of one share plus the exchange
        (`collective_ms`) is faster than the plan of the whole level (b200fhe_plan_ms, measured B200 tables) -
        e.g. a 100-rotation level fits one 3.0 ms wave either way, but its 50-rotation halves fit the 2.4 ms
        cluster kernel."""
        self.nl, self.eng, self.be = nl, eng, backend
        self.rank, self.world, self.group = rank, world, group
        self.widths = eng.level_widths
        self.min_shard_width = min_shard_width
        self.exchanged_slots = 0
        self.collectives = 0
        if min_shard_width is None:
            from .lib import plan_ms

            jobs = eng.level_bootstraps
            self.replicate = [world == 1 or plan_ms(j) <= plan_ms(-(-j // world)) + collective_ms for j in jobs]
        else:
            self.replicate = [world == 1 or w < min_shard_width for w in self.widths]

    def _share(self, level):
        w = self.widths[level]
        chunk = (w + self.world - 1) // self.world
        lo = min(w, self.rank * chunk)
        return chunk, lo, min(w, lo + chunk)

    def _compute(self, level, everyone: bool):
        w = self.widths[level]
        if isinstance(self.be, PlainBackend):
            if everyone:
                self.be.run_level_shard(level, 0, w)
            else:
                _, lo, hi = self._share(level)
                self.be.run_level_shard(level, lo, hi)
        else:
            if everyone:  # every rank computes the whole level in one batch
                self.eng.run_level_shard(level, 0, 1)
            else:
                self.eng.run_level_shard(level, self.rank, self.world)

    def run(self):
        import torch.distributed as dist

        for level, w in enumerate(self.widths):
            replicate = self.replicate[level]
            self._compute(level, everyone=replicate)
            if replicate:
                continue
            chunk, _, _ = self._share(level)
            base = self.eng.level_slot_base(level) * self.be.elem
            n = chunk * self.be.elem
            out = self.be.arena[base: base + n * self.world]
            inp = out[self.rank * n: (self.rank + 1) * n]
            if isinstance(self.be, PlainBackend):
                inp = inp.clone()  # gloo: keep source and destination disjoint
            with self.be.stream_ctx():
                dist.all_gather_into_tensor(out, inp, group=self.group)
            self.exchanged_slots += chunk * self.world
            self.collectives += 1

    def tick(self):
        self.be.tick()
