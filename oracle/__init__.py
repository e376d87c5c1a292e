"""TEST INFRASTRUCTURE — ctypes loader for the exact-integer oracle and the reference driver.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``iyokan_b200``)
never does: it fails loudly when its CUDA library is missing instead of falling
back to anything in here.

Parity status: pinned against the unmodified reference (see tfhe_oracle.h).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
# Parameter flavour, fixed per process like the reference's compile-time switch (IYOKAN_80BIT_SECURITY): the environment
# variable B200FHE_FLAVOUR ("" = 128-bit, "80" = 80-bit) is read once, at import, by this package and by iyokan_b200.lib.
FLAVOUR = os.environ.get("B200FHE_FLAVOUR", "")
if FLAVOUR not in ("", "80"):
    raise ImportError(f"B200FHE_FLAVOUR must be '' (128-bit) or '80', not {FLAVOUR!r}")
LIB_PATH = HERE / f"libtfhe_oracle{FLAVOUR}.so"
REF_DRIVER = HERE / "_ref" / f"ref_driver{FLAVOUR}"
IYOKAN_PACKET = HERE / "_ref" / "iyokan-packet"   # the reference's own packet tool, built unmodified
REF_LINK_TEST = HERE / "_ref" / "b200_gate_test"  # tests/ref_link/b200_gate_test.cpp: TFHEpp types over the C ABI
IYOKAN_REF = HERE / "_ref" / "iyokan"             # the reference's own iyokan (plain + tfhe on CPU), built unmodified
IYOKAN_B200 = HERE / "_ref" / "iyokan-b200"       # iyokan_b200/host/iyokan_b200_main.cpp: the reference's loader + our engine
B200_TEST0 = HERE / "_ref" / "b200_test0"         # the reference's src/test0.cpp templated tests on the B200 plugin (iyokan_b200.hpp)

if FLAVOUR == "80":   # TFHEpp include/params/CGGI16.hpp
    N0, N1, L, T = 500, 1024, 2, 8
    MU0, T0, S0 = 1 << 29, np.uint32, np.int32
else:                 # TFHEpp include/params/128bit.hpp
    N0, N1, L, T = 636, 1024, 3, 7
    MU0, T0, S0 = 1 << 13, np.uint16, np.int16
TLWE0, TLWE1, ROWS = N0 + 1, N1 + 1, 2 * L
MU1 = 1 << 29

OPS = {
    "AND": 0, "NAND": 1, "ANDNOT": 2, "OR": 3, "NOR": 4, "ORNOT": 5, "XOR": 6, "XNOR": 7,
    "MUX": 8, "NOT": 9, "COPY": 10, "CONST0": 11, "CONST1": 12, "ANDNY": 13, "ORNY": 14,
}


def plain_gate(op: int, a, b, c):
    """Plaintext truth table, Iyokan's plain back-end (src/iyokan_plain.hpp:80-117)."""
    a, b, c = (np.asarray(x, dtype=np.uint8) for x in (a, b, c))
    table = {
        0: a & b, 1: 1 - (a & b), 2: a & (1 - b), 3: a | b, 4: 1 - (a | b), 5: a | (1 - b),
        6: a ^ b, 7: 1 - (a ^ b), 8: np.where(c == 1, b, a), 9: 1 - a, 10: a,
        11: np.zeros_like(a), 12: np.ones_like(a), 13: (1 - a) & b, 14: (1 - a) | b,
    }
    return table[int(op)].astype(np.uint8)


def plain_gate_vec(ops, a, b, c) -> np.ndarray:
    ops = np.asarray(ops, dtype=np.uint8)
    out = np.empty(ops.size, np.uint8)
    for i, op in enumerate(ops):
        out[i] = plain_gate(op, a[i], b[i], c[i])
    return out


def _optional(cmd) -> None:
    """Build step whose product only enables extra tests: a failure is reported, not fatal."""
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        import sys

        print(f"oracle.build: optional step {' '.join(cmd[-2:])} failed:\n{r.stderr[-1500:]}", file=sys.stderr)


def build(force: bool = False) -> None:
    """Compile the C restatement (and the reference driver when the reference tree exists)."""
    libs = [HERE / "libtfhe_oracle.so", HERE / "libtfhe_oracle80.so"]
    if force or not all(x.exists() for x in libs) or min(x.stat().st_mtime for x in libs) < (HERE / "tfhe_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "oracle"], check=True, capture_output=True)
    if Path("/root/reference/thirdparty/cuFHE/thirdparties/TFHEpp/include").is_dir():
        if (force or not REF_DRIVER.exists() or not IYOKAN_PACKET.exists()
                or REF_DRIVER.stat().st_mtime < (HERE / "ref_driver.cpp").stat().st_mtime):
            subprocess.run(["make", "-C", str(HERE), "-j8", "ref"], check=True, capture_output=True)
        if force or not IYOKAN_REF.exists():  # optional: the differential tests skip without it
            _optional(["make", "-C", str(HERE), "-j4", "refbin"])
        # reference-side binding test (TFHEpp types + the product's C ABI); needs the CUDA library built first
        so = HERE.parent / "iyokan_b200" / "csrc" / "libb200fhe.so"
        src = HERE.parent / "tests" / "ref_link" / "b200_gate_test.cpp"
        host = HERE.parent / "iyokan_b200" / "host"
        deps = [so, src, host / "libb200net.so", host / "iyokan_b200_main.cpp", host / "iyokan_b200.hpp",
                HERE.parent / "tests" / "ref_link" / "b200_test0.cpp"]
        outs = [REF_LINK_TEST, IYOKAN_B200, B200_TEST0]
        if all(d.exists() for d in deps) and (force or not all(o.exists() for o in outs)
                                              or min(o.stat().st_mtime for o in outs) < max(d.stat().st_mtime for d in deps)):
            _optional(["make", "-C", str(HERE), "-j4", "reflink"])


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB_PATH))
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Keys:
    """Flat key arrays in the reference's memory layout (TFHEpp params.hpp:102-128)."""

    def __init__(self, sk0, sk1, bk, ksk):
        self.sk0, self.sk1, self.bk, self.ksk = sk0, sk1, bk, ksk

    def save(self, d):
        d = Path(d)
        d.mkdir(parents=True, exist_ok=True)
        self.sk0.tofile(d / "sk0.bin")
        self.sk1.astype(np.uint32).tofile(d / "sk1.bin")
        self.bk.tofile(d / "bk.bin")
        self.ksk.tofile(d / "ksk.bin")

    @staticmethod
    def load(d):
        d = Path(d)
        return Keys(
            np.fromfile(d / "sk0.bin", dtype=T0),
            np.fromfile(d / "sk1.bin", dtype=np.uint32).astype(np.int32),
            np.fromfile(d / "bk.bin", dtype=np.uint32).reshape(N0, ROWS, 2, N1),
            np.fromfile(d / "ksk.bin", dtype=T0).reshape(N1, T, 3, TLWE0),
        )


def keygen(seed: int) -> Keys:
    sk0 = np.empty(N0, T0)
    sk1 = np.empty(N1, np.int32)
    bk = np.empty((N0, ROWS, 2, N1), np.uint32)
    ksk = np.empty((N1, T, 3, TLWE0), T0)
    lib().orc_keygen(ctypes.c_uint64(seed), _p(sk0), _p(sk1), _p(bk), _p(ksk))
    return Keys(sk0, sk1, bk, ksk)


_KEY_CACHE: dict[int, Keys] = {}


def cached_keys(seed: int) -> Keys:
    if seed not in _KEY_CACHE:
        _KEY_CACHE[seed] = keygen(seed)
    return _KEY_CACHE[seed]


def encrypt_bits(seed: int, keys: Keys, bits) -> np.ndarray:
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.empty((bits.size, TLWE0), T0)
    lib().orc_encrypt_bits(ctypes.c_uint64(seed), _p(keys.sk0), _p(bits), ctypes.c_size_t(bits.size), _p(out))
    return out


def decrypt_bits(keys: Keys, c) -> np.ndarray:
    c = np.ascontiguousarray(c, dtype=T0).reshape(-1, TLWE0)
    bits = np.empty(c.shape[0], np.uint8)
    lib().orc_decrypt_bits(_p(keys.sk0), _p(c), ctypes.c_size_t(c.shape[0]), _p(bits))
    return bits


def phase(keys: Keys, c) -> np.ndarray:
    c = np.ascontiguousarray(c, dtype=T0).reshape(-1, TLWE0)
    ph = np.empty(c.shape[0], S0)
    lib().orc_phase(_p(keys.sk0), _p(c), ctypes.c_size_t(c.shape[0]), _p(ph))
    return ph


def phase1(keys: Keys, c) -> np.ndarray:
    c = np.ascontiguousarray(c, dtype=np.uint32).reshape(-1, TLWE1)
    ph = np.empty(c.shape[0], np.int32)
    lib().orc_phase1(_p(keys.sk1), _p(c), ctypes.c_size_t(c.shape[0]), _p(ph))
    return ph


def gate_batch(keys: Keys, ops, in0, in1=None, in2=None, nthreads: int = 0) -> np.ndarray:
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    n = ops.size
    arrs = []
    for a in (in0, in1, in2):
        if a is None:
            a = np.zeros((n, TLWE0), T0)
        arrs.append(np.ascontiguousarray(a, dtype=T0).reshape(n, TLWE0))
    out = np.empty((n, TLWE0), T0)
    lib().orc_gate_batch(_p(ops), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(out), ctypes.c_size_t(n),
                         _p(keys.bk), _p(keys.ksk), ctypes.c_int(nthreads))
    return out


def decompose(poly) -> np.ndarray:
    poly = np.ascontiguousarray(poly, dtype=np.uint32)
    out = np.empty((L, N1), np.int32)
    lib().orc_decompose(_p(poly), _p(out))
    return out


def mul_xai(poly, a: int, minus_one: bool) -> np.ndarray:
    poly = np.ascontiguousarray(poly, dtype=np.uint32)
    out = np.empty(N1, np.uint32)
    (lib().orc_mul_xai_minus_one if minus_one else lib().orc_mul_xai)(_p(poly), ctypes.c_uint32(a), _p(out))
    return out


def cmux_step(acc, trgsw, abar: int) -> np.ndarray:
    acc = np.array(acc, dtype=np.uint32).reshape(2, N1).copy()
    trgsw = np.ascontiguousarray(trgsw, dtype=np.uint32)
    lib().orc_cmux_step(_p(acc), _p(trgsw), ctypes.c_uint32(abar))
    return acc


def mod_switch(c):
    c = np.ascontiguousarray(c, dtype=T0)
    abar = np.empty(N0, np.uint32)
    bbar = ctypes.c_uint32(0)
    lib().orc_mod_switch(_p(c), _p(abar), ctypes.byref(bbar))
    return abar, bbar.value


def blind_rotate(keys: Keys, c) -> np.ndarray:
    c = np.ascontiguousarray(c, dtype=T0)
    acc = np.empty((2, N1), np.uint32)
    lib().orc_blind_rotate(_p(c), _p(keys.bk), _p(acc))
    return acc


def bootstrap_to_lvl1(keys: Keys, c) -> np.ndarray:
    c = np.ascontiguousarray(c, dtype=T0).reshape(-1, TLWE0)
    out = np.empty((c.shape[0], TLWE1), np.uint32)
    for i in range(c.shape[0]):
        lib().orc_bootstrap_to_lvl1(_p(c[i]), _p(keys.bk), _p(out[i]))
    return out


def keyswitch(keys: Keys, u) -> np.ndarray:
    u = np.ascontiguousarray(u, dtype=np.uint32).reshape(-1, TLWE1)
    out = np.empty((u.shape[0], TLWE0), T0)
    for i in range(u.shape[0]):
        lib().orc_keyswitch(_p(u[i]), _p(keys.ksk), _p(out[i]))
    return out


def num_bootstraps(ops) -> int:
    ops = np.asarray(ops, dtype=np.uint8)
    return int(((ops <= 7) | (ops >= 13)).sum() + 2 * (ops == 8).sum())


def iyokan_ref(*args, timeout=1800) -> subprocess.CompletedProcess:
    """Run the reference's own `iyokan` binary (oracle/_ref/iyokan)."""
    return subprocess.run([str(IYOKAN_REF), *map(str, args)], capture_output=True, text=True, timeout=timeout)


def have_ref() -> bool:
    return REF_DRIVER.exists() and os.access(REF_DRIVER, os.X_OK)


def ref(*args, check=True) -> str:
    """Run the unmodified-reference driver (oracle/_ref/ref_driver)."""
    r = subprocess.run([str(REF_DRIVER), *map(str, args)], check=check, capture_output=True, text=True)
    return r.stdout


def have_iyokan_packet() -> bool:
    return IYOKAN_PACKET.exists() and os.access(IYOKAN_PACKET, os.X_OK)


def iyokan_packet(*args) -> str:
    """Run the reference's iyokan-packet tool (oracle/_ref/iyokan-packet)."""
    r = subprocess.run([str(IYOKAN_PACKET), *map(str, args)], check=True, capture_output=True, text=True)
    return r.stdout
