"""ctypes binding of the C ABI in ``include/b200fhe.h`` (built from ``csrc/b200fhe.cu``).

The library is built in-tree by ``__graft_entry__.build()`` / ``iyokan_b200.build.build()`` as
``iyokan_b200/csrc/libb200fhe.so``.  There is NO fallback: if the shared library is missing or
no sm_100a GPU is present, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

CSRC = Path(__file__).resolve().parent / "csrc"
# Parameter flavour: a property of the process, like the reference's compile-time IYOKAN_80BIT_SECURITY switch
# (CMakeLists.txt:28-30).  B200FHE_FLAVOUR="" loads libb200fhe.so (128-bit set), "80" loads libb200fhe80.so (80-bit set).
FLAVOUR = os.environ.get("B200FHE_FLAVOUR", "")
if FLAVOUR not in ("", "80"):
    raise ImportError(f"B200FHE_FLAVOUR must be '' (128-bit) or '80', not {FLAVOUR!r}")
LIB_PATH = Path(os.environ["B200FHE_LIB"]) if os.environ.get("B200FHE_LIB") else CSRC / f"libb200fhe{FLAVOUR}.so"  # override: debug builds

N1, TLWE1_LEN = 1024, 1025
if FLAVOUR == "80":   # TFHEpp include/params/CGGI16.hpp
    N0, GL, KS_T, LIMBS, T0, MU0, SLOT_BYTES = 500, 2, 8, 5, np.uint32, 1 << 29, 2048
else:                 # TFHEpp include/params/128bit.hpp
    N0, GL, KS_T, LIMBS, T0, MU0, SLOT_BYTES = 636, 3, 7, 3, np.uint16, 1 << 13, 1280
TLWE0_LEN = N0 + 1
BK_SHAPE = (N0, 2 * GL, 2, 1024)              # raw TRGSW bootstrapping key, uint32
KSK_SHAPE = (1024, KS_T, 3, TLWE0_LEN)        # identity key-switching key, lvl0 torus words
BK_NTT_SHAPE = (N0, 2 * LIMBS, 2 * GL, 1024)  # device form: [i][poly*LIMBS+limb][row][position]

OPS = {
    "AND": 0, "NAND": 1, "ANDNOT": 2, "OR": 3, "NOR": 4, "ORNOT": 5, "XOR": 6, "XNOR": 7,
    "MUX": 8, "NOT": 9, "COPY": 10, "CONST0": 11, "CONST1": 12, "ANDNY": 13, "ORNY": 14,
}
# gate bootstraps (blind rotations) per opcode: SURVEY.md §8(d)
BOOTSTRAPS = {op: (2 if name == "MUX" else 0 if name in ("NOT", "COPY", "CONST0", "CONST1") else 1)
              for name, op in OPS.items()}


class B200FheError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load libb200fhe.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200FheError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, u64, sz, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int
    sig = {
        "b200fhe_create": (ci, [ctypes.POINTER(vp), ci]),
        "b200fhe_destroy": (None, [vp]),
        "b200fhe_last_error": (ctypes.c_char_p, []),
        "b200fhe_set_jobs_per_cta": (ci, [vp, ci]),
        "b200fhe_set_kernel_variant": (ci, [vp, ci]),
        "b200fhe_load_keys": (ci, [vp, vp, vp]),
        "b200fhe_arena_alloc": (ci, [vp, sz]),
        "b200fhe_arena_attach": (ci, [vp, vp, sz]),
        "b200fhe_arena_slots": (sz, [vp]),
        "b200fhe_arena_dev_ptr": (vp, [vp]),
        "b200fhe_upload": (ci, [vp, vp, vp, sz]),
        "b200fhe_download": (ci, [vp, vp, vp, sz]),
        "b200fhe_gate_batch": (ci, [vp, vp, vp, vp, vp, vp, sz]),
        "b200fhe_dff_tick": (ci, [vp, vp, vp, sz]),
        "b200fhe_sync": (ci, [vp]),
        "b200fhe_query": (ci, [vp]),
        "b200fhe_gates_host": (ci, [vp, vp, vp, vp, vp, vp, sz]),
        "b200fhe_host_alloc": (ci, [ctypes.POINTER(vp), sz]),
        "b200fhe_host_free": (ci, [vp]),
        "b200fhe_launch_count": (u64, [vp]),
        "b200fhe_last_batch_ms": (ci, [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
        "b200fhe_last_batch_segments": (ci, [vp, vp, vp, vp, vp, ci]),
        "b200fhe_plan_rotation": (ci, [ci, vp, vp, vp, ci]),
        "b200fhe_plan_ms": (ctypes.c_double, [ci]),
        "b200fhe_plan_table": (ci, [vp, vp, vp, vp, ci]),
        "b200fhe_stream": (vp, [vp]),
        "b200fhe_program_create": (ci, [vp, ctypes.POINTER(vp)]),
        "b200fhe_program_destroy": (None, [vp]),
        "b200fhe_program_batch": (ci, [vp, vp, vp, vp, vp, vp, sz]),
        "b200fhe_program_tick": (ci, [vp, vp, vp, sz]),
        "b200fhe_program_exchange": (ci, [vp, sz, sz]),
        "b200fhe_program_finalize": (ci, [vp]),
        "b200fhe_program_launch": (ci, [vp]),
        "b200fhe_program_profile": (ci, [vp, vp, sz, vp]),
        "b200fhe_program_info": (ci, [vp, vp, vp, vp, vp, vp, vp]),
        "b200fhe_comm_unique_id": (ci, [vp]),
        "b200fhe_comm_init": (ci, [vp, ci, ci, vp]),
        "b200fhe_comm_rank": (ci, [vp]),
        "b200fhe_comm_world": (ci, [vp]),
        "b200fhe_exchange": (ci, [vp, sz, sz]),
        "b200fhe_test_bootstrap_lvl1": (ci, [vp, vp, vp, sz]),
        "b200fhe_test_keyswitch": (ci, [vp, vp, vp, sz]),
        "b200fhe_test_read_bk_ntt": (ci, [vp, vp, sz, sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


EXPORTS = [
    "b200fhe_create", "b200fhe_destroy", "b200fhe_last_error", "b200fhe_set_jobs_per_cta", "b200fhe_set_kernel_variant",
    "b200fhe_load_keys",
    "b200fhe_arena_alloc", "b200fhe_arena_attach", "b200fhe_arena_slots", "b200fhe_arena_dev_ptr",
    "b200fhe_upload", "b200fhe_download", "b200fhe_gate_batch", "b200fhe_dff_tick", "b200fhe_sync",
    "b200fhe_query", "b200fhe_gates_host", "b200fhe_host_alloc", "b200fhe_host_free", "b200fhe_launch_count",
    "b200fhe_last_batch_ms", "b200fhe_last_batch_segments", "b200fhe_plan_rotation", "b200fhe_plan_ms", "b200fhe_plan_table", "b200fhe_stream", "b200fhe_test_bootstrap_lvl1", "b200fhe_test_keyswitch",
    "b200fhe_test_read_bk_ntt",
    "b200fhe_program_create", "b200fhe_program_destroy", "b200fhe_program_batch", "b200fhe_program_tick",
    "b200fhe_program_exchange", "b200fhe_program_finalize", "b200fhe_program_launch", "b200fhe_program_profile", "b200fhe_program_info",
    "b200fhe_comm_unique_id", "b200fhe_comm_init", "b200fhe_comm_rank", "b200fhe_comm_world", "b200fhe_exchange",
]


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _u32(a, n=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if n is not None and a.size != n:
        raise ValueError("slot id array has the wrong length")
    return a


class PinnedBuffer:
    """Page-locked host array (cudaHostAlloc through the C ABI)."""

    def __init__(self, shape, dtype):
        lib = load()
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = ctypes.c_void_p()
        if lib.b200fhe_host_alloc(ctypes.byref(p), nbytes):
            raise B200FheError(lib.b200fhe_last_error().decode())
        self._p = p
        buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            load().b200fhe_host_free(self._p)
            self._p = None


def plan_rotation(njobs: int):
    """Launch plan of the batch-size heuristic for `njobs` blind rotations: [(variant, jobs_per_cta, jobs)]."""
    lib = load()
    v, g, n = (np.zeros(8, np.int32) for _ in range(3))
    k = lib.b200fhe_plan_rotation(int(njobs), _ptr(v), _ptr(g), _ptr(n), 8)
    return [(int(v[i]), int(g[i]), int(n[i])) for i in range(k)]


def plan_table() -> dict:
    """The launch-plan table: {"calibrated": bool, "shapes": [(variant, jobs_per_cta, wave_jobs, wave_ms)]}."""
    lib = load()
    v, g, w = (np.zeros(8, np.int32) for _ in range(3))
    ms = np.zeros(8, np.float64)
    k = lib.b200fhe_plan_table(_ptr(v), _ptr(g), _ptr(w), _ptr(ms), 8)
    return {"calibrated": k > 0, "shapes": [(int(v[i]), int(g[i]), int(w[i]), float(ms[i])) for i in range(abs(k))]}


def plan_ms(njobs: int) -> float:
    """Modelled blind-rotation time (ms) of a frontier of `njobs` rotations under the launch plan."""
    return float(load().b200fhe_plan_ms(int(njobs)))


def comm_unique_id() -> bytes:
    """128-byte communicator id (ncclUniqueId); rank 0 creates it and hands it to the other ranks."""
    lib = load()
    buf = (ctypes.c_uint8 * 128)()
    if lib.b200fhe_comm_unique_id(buf):
        raise B200FheError(lib.b200fhe_last_error().decode())
    return bytes(buf)


class Program:
    """A static schedule (frontiers, exchanges, tick) recorded once and replayed as one CUDA graph."""

    def __init__(self, ctx: "Context"):
        self._lib, self.ctx = ctx._lib, ctx
        h = ctypes.c_void_p()
        ctx._ck(self._lib.b200fhe_program_create(ctx._h, ctypes.byref(h)))
        self._h = h

    def batch(self, opcode, in0, in1, in2, out):
        op = np.ascontiguousarray(opcode, dtype=np.uint8)
        n = op.size
        a, b, c, o = _u32(in0, n), _u32(in1, n), _u32(in2, n), _u32(out, n)
        self.ctx._ck(self._lib.b200fhe_program_batch(self._h, _ptr(op), _ptr(a), _ptr(b), _ptr(c), _ptr(o), n))

    def tick(self, src, dst):
        s, d = _u32(src), _u32(dst)
        self.ctx._ck(self._lib.b200fhe_program_tick(self._h, _ptr(s), _ptr(d), s.size))

    def exchange(self, first_slot, slots_per_rank):
        self.ctx._ck(self._lib.b200fhe_program_exchange(self._h, first_slot, slots_per_rank))

    def finalize(self):
        self.ctx._ck(self._lib.b200fhe_program_finalize(self._h))

    def launch(self):
        self.ctx._ck(self._lib.b200fhe_program_launch(self._h))

    def info(self) -> dict:
        r, l, e, s = (ctypes.c_uint64() for _ in range(4))
        g, m = ctypes.c_int(), ctypes.c_double()
        self.ctx._ck(self._lib.b200fhe_program_info(self._h, ctypes.byref(r), ctypes.byref(l), ctypes.byref(e), ctypes.byref(s),
                                                    ctypes.byref(g), ctypes.byref(m)))
        return {"rotations": r.value, "launches_per_replay": l.value, "exchanges": e.value, "exchanged_slots": s.value,
                "is_graph": bool(g.value), "model_ms": m.value}

    def close(self):
        if self._h is not None:
            self._lib.b200fhe_program_destroy(self._h)
            self._h = None


class Context:
    """One GPU's evaluation context: keys, ciphertext arena, stream."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = ctypes.c_void_p()
        if self._lib.b200fhe_create(ctypes.byref(h), device):
            raise B200FheError(self._lib.b200fhe_last_error().decode())
        self._h = h
        self.device = device

    def _ck(self, rc):
        if rc:
            raise B200FheError(self._lib.b200fhe_last_error().decode())

    def close(self):
        if self._h is not None:
            self._lib.b200fhe_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_jobs_per_cta(self, g: int):
        self._ck(self._lib.b200fhe_set_jobs_per_cta(self._h, g))

    def set_kernel_variant(self, variant: int):
        self._ck(self._lib.b200fhe_set_kernel_variant(self._h, variant))

    def load_keys(self, bk_raw: np.ndarray, ksk: np.ndarray):
        bk_raw = np.ascontiguousarray(bk_raw, dtype=np.uint32)
        ksk = np.ascontiguousarray(ksk, dtype=T0)
        if bk_raw.size != int(np.prod(BK_SHAPE)) or ksk.size != int(np.prod(KSK_SHAPE)):
            raise ValueError("key arrays have the wrong size for this parameter set")
        self._ck(self._lib.b200fhe_load_keys(self._h, _ptr(bk_raw), _ptr(ksk)))

    def arena_alloc(self, n_slots: int):
        self._ck(self._lib.b200fhe_arena_alloc(self._h, n_slots))

    def arena_attach(self, dev_ptr: int, n_slots: int):
        self._ck(self._lib.b200fhe_arena_attach(self._h, ctypes.c_void_p(dev_ptr), n_slots))

    @property
    def arena_slots(self) -> int:
        return int(self._lib.b200fhe_arena_slots(self._h))

    @property
    def arena_dev_ptr(self) -> int:
        return int(self._lib.b200fhe_arena_dev_ptr(self._h) or 0)

    @property
    def stream(self) -> int:
        return int(self._lib.b200fhe_stream(self._h) or 0)

    def upload(self, slot_ids, tlwe: np.ndarray):
        ids = _u32(slot_ids)
        tlwe = np.ascontiguousarray(tlwe, dtype=T0)
        if tlwe.size != ids.size * TLWE0_LEN:
            raise ValueError(f"tlwe array must be [n][{TLWE0_LEN}] {np.dtype(T0).name}")
        self._ck(self._lib.b200fhe_upload(self._h, _ptr(ids), _ptr(tlwe), ids.size))
        self.sync()  # the source may be pageable numpy memory

    def download(self, slot_ids, out: np.ndarray | None = None) -> np.ndarray:
        ids = _u32(slot_ids)
        if out is None:
            out = np.empty((ids.size, TLWE0_LEN), T0)
        self._ck(self._lib.b200fhe_download(self._h, _ptr(ids), _ptr(out), ids.size))
        return out

    def gate_batch(self, opcode, in0, in1, in2, out):
        op = np.ascontiguousarray(opcode, dtype=np.uint8)
        n = op.size
        # keep the converted id arrays alive across the call (ctypes only sees raw addresses)
        a, b, c, o = _u32(in0, n), _u32(in1, n), _u32(in2, n), _u32(out, n)
        self._ck(self._lib.b200fhe_gate_batch(self._h, _ptr(op), _ptr(a), _ptr(b), _ptr(c), _ptr(o), n))

    def dff_tick(self, src, dst):
        s, d = _u32(src), _u32(dst, None)
        if s.size != d.size:
            raise ValueError("src/dst length mismatch")
        self._ck(self._lib.b200fhe_dff_tick(self._h, _ptr(s), _ptr(d), s.size))

    def sync(self):
        self._ck(self._lib.b200fhe_sync(self._h))

    # ---- multi-GPU exchange (NCCL bound inside the library) ----
    def comm_init(self, rank: int, world: int, unique_id: bytes | None):
        """Join the communicator of `world` ranks; `unique_id` = the 128 bytes rank 0 got from comm_unique_id()."""
        buf = None if unique_id is None else (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._lib.b200fhe_comm_init(self._h, rank, world, buf))

    def exchange(self, first_slot: int, slots_per_rank: int):
        self._ck(self._lib.b200fhe_exchange(self._h, first_slot, slots_per_rank))

    def program(self) -> "Program":
        return Program(self)

    def query(self) -> int:
        return int(self._lib.b200fhe_query(self._h))

    def gates_host(self, opcode, in0, in1, in2, out: np.ndarray | None = None) -> np.ndarray:
        op = np.ascontiguousarray(opcode, dtype=np.uint8)
        n = op.size
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=T0) for a in (in0, in1, in2)]
        if out is None:
            out = np.empty((n, TLWE0_LEN), T0)
        self._ck(self._lib.b200fhe_gates_host(self._h, _ptr(op), _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                                               _ptr(out), n))
        return out

    @property
    def launch_count(self) -> int:
        return int(self._lib.b200fhe_launch_count(self._h))

    def last_batch_ms(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        self._ck(self._lib.b200fhe_last_batch_ms(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def last_batch_segments(self):
        """Launch plan of the most recent gate_batch: list of dicts (variant, jobs_per_cta, jobs, ms)."""
        v, g, n = (np.zeros(8, np.int32) for _ in range(3))
        ms = np.zeros(8, np.float32)
        k = self._lib.b200fhe_last_batch_segments(self._h, _ptr(v), _ptr(g), _ptr(n), _ptr(ms), 8)
        if k < 0:
            raise B200FheError(self._lib.b200fhe_last_error().decode())
        return [{"variant": int(v[i]), "jobs_per_cta": int(g[i]), "jobs": int(n[i]), "ms": float(ms[i])} for i in range(k)]

    # ---- stage-level test hooks ----
    def test_bootstrap_lvl1(self, c: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(c, dtype=T0).reshape(-1, TLWE0_LEN)
        out = np.empty((c.shape[0], TLWE1_LEN), np.uint32)
        self._ck(self._lib.b200fhe_test_bootstrap_lvl1(self._h, _ptr(c), _ptr(out), c.shape[0]))
        return out

    def test_keyswitch(self, u: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.uint32).reshape(-1, TLWE1_LEN)
        out = np.empty((u.shape[0], TLWE0_LEN), T0)
        self._ck(self._lib.b200fhe_test_keyswitch(self._h, _ptr(u), _ptr(out), u.shape[0]))
        return out

    def test_read_bk_ntt(self, first_i: int, count_i: int) -> np.ndarray:
        out = np.empty((count_i,) + BK_NTT_SHAPE[1:], np.uint32)
        self._ck(self._lib.b200fhe_test_read_bk_ntt(self._h, _ptr(out), first_i, count_i))
        return out
