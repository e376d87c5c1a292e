"""Multi-GPU evaluation of one netlist: one process per GPU, a static schedule shared by all ranks.

The reference's multi-GPU mode replicates the keys and round-robins gates over devices, moving every
ciphertext through host memory around each gate (cufhe_gpu.cuh:164-169, cufhe_gates_gpu.cu:145-157); it has
no collective.  Here the host engine (iyokan_b200/host/b200net.cpp, `b200net_schedule`) cuts a clock cycle
into steps once; every step is either replicated (all ranks evaluate all of it: narrow critical levels, no
exchange) or sharded (contiguous shares balanced by rotation count, then ONE in-place all-gather of the step's
slot range over NVLink).  On the GPU the whole clock - kernels and NCCL all-gathers - is a single CUDA graph
per rank, built by `b200net_bind_rank` / `b200fhe_program_*` behind the C ABI; Python only creates the
communicator (`setup_comm`).

The classes below replay the SAME schedule on CPU tensors with the plaintext back-end and any
torch.distributed backend (gloo in the tests): that is how the N > 1 schedule - dependencies, shares, slot
ranges, exchanges - is verified without GPUs (tests/test_shard_gloo.py).
"""
from __future__ import annotations

import numpy as np

from .netlist import DFF, NetEngine, Netlist


def setup_comm(ctx, rank: int, world: int, group=None):
    """Create the library's NCCL communicator: rank 0's 128-byte id travels over torch.distributed."""
    if world == 1:
        ctx.comm_init(0, 1, None)
        return
    import torch.distributed as dist

    from .lib import comm_unique_id

    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(rank, world, box[0])


class PlainBackend:
    """Bits instead of ciphertexts, a CPU uint8 tensor as the 'arena' (one byte per slot)."""

    def __init__(self, nl: Netlist, eng: NetEngine, world: int, flags: int = 1):
        import torch

        self.nl, self.eng = nl, eng
        eng.schedule(world, flags)
        self._slot = np.array([eng.slot_of(i) for i in range(nl.n)], dtype=np.int64)
        self.arena = torch.zeros(eng.num_slots, dtype=torch.uint8)
        self._np = self.arena.numpy()

    def set_nodes(self, nodes, bits):
        self._np[self._slot[np.asarray(nodes)]] = np.asarray(bits, np.uint8)

    def get_nodes(self, nodes):
        return self._np[self._slot[np.asarray(nodes)]].copy()

    def run_gates(self, g):
        """Evaluate the gates `g` (node ids) of one step: all read slots written by earlier steps."""
        if g.size == 0:
            return
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


class ScheduledRunner:
    """Replays the engine's static schedule on `world` ranks; all ranks hold identical arenas between steps."""

    def __init__(self, nl: Netlist, eng: NetEngine, backend: PlainBackend, rank: int, world: int, group=None):
        self.nl, self.eng, self.be = nl, eng, backend
        self.rank, self.world, self.group = rank, world, group
        self.steps = [(eng.step_gates(k, rank).astype(np.int64), eng.step_exchange(k)) for k in range(eng.num_steps)]
        self.exchanged_slots = 0
        self.collectives = 0

    def run(self):
        import torch.distributed as dist

        for gates, (first, per) in self.steps:
            self.be.run_gates(gates)
            if per == 0 or self.world == 1:
                continue
            out = self.be.arena[first: first + per * self.world]
            inp = out[self.rank * per: (self.rank + 1) * per].clone()  # gloo: keep source and destination disjoint
            dist.all_gather_into_tensor(out, inp, group=self.group)
            self.exchanged_slots += per * self.world
            self.collectives += 1

    def tick(self):
        self.be.tick()
