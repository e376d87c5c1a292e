"""Generate tests/golden/tfhepp_golden.npz from the UNMODIFIED reference (oracle/_ref/ref_driver).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py                        # 128-bit parameters -> tfhepp_golden.npz
    B200FHE_FLAVOUR=80 python tests/golden/make_golden.py     # 80-bit parameters (ref_driver80) -> tfhepp_golden80.npz
The keys are produced by the oracle's deterministic integer-only key generator from KEY_SEED, so
the tests can regenerate exactly the same key material anywhere; only the reference's OUTPUTS on
those keys are stored.  A SHA-256 of the key arrays guards that assumption.
"""
import hashlib
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402

KEY_SEED = 424242


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    keys = O.keygen(KEY_SEED)
    rng = np.random.default_rng(2026)
    g = {"key_seed": np.uint64(KEY_SEED)}
    g["bk_sha256"] = np.frombuffer(hashlib.sha256(keys.bk.tobytes()).digest(), dtype=np.uint8)
    g["ksk_sha256"] = np.frombuffer(hashlib.sha256(keys.ksk.tobytes()).digest(), dtype=np.uint8)
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        keys.save(d)
        # Decomposition<lvl1param> (trgsw.hpp:62-78), exact
        p = rng.integers(0, 2**32, (4, 1024), dtype=np.uint32)
        p[0, :8] = [0, 1, 2**31, 2**32 - 1, 2**26, 2**26 - 1, 2**13, 2**13 - 1]
        p[0, 8:12] = [2**22, 2**22 - 1, 2**11, 2**11 - 1]
        p.tofile(d / "p.bin")
        O.ref("decompose", d / "p.bin", d / "dec.bin")
        g["decompose_in"], g["decompose_out"] = p, np.fromfile(d / "dec.bin", dtype=np.int32).reshape(4, O.L, 1024)
        # PolynomialMulByXai / MulByXaiMinusOne (utils.hpp:113-144), exact
        q = rng.integers(0, 2**32, (8, 1024), dtype=np.uint32)
        a = np.array([0, 1, 5, 1023, 1024, 1025, 2047, 2048], dtype=np.uint32)
        q.tofile(d / "q.bin")
        a.tofile(d / "a.bin")
        O.ref("mulxai", d / "q.bin", d / "a.bin", d / "mx.bin")
        g["mulxai_in"], g["mulxai_a"] = q, a
        g["mulxai_out"] = np.fromfile(d / "mx.bin", dtype=np.uint32).reshape(8, 2, 1024)
        # one CMUX step (detwfa.hpp:36-49) on a random accumulator: FFT result, compare within rounding
        trgsw = keys.bk[7].copy()
        acc = rng.integers(0, 2**32, (2, 1024), dtype=np.uint32)
        trgsw.tofile(d / "tg.bin")
        acc.tofile(d / "acc.bin")
        O.ref("cmuxstep", d / "tg.bin", d / "acc.bin", 777, d / "cm.bin")
        g["cmux_trgsw"], g["cmux_acc"], g["cmux_abar"] = trgsw, acc, np.uint32(777)
        g["cmux_out"] = np.fromfile(d / "cm.bin", dtype=np.uint32).reshape(2, 1024)
        # gates: every opcode x 4 seeded inputs through TFHEpp::Hom*
        names = list(O.OPS)
        ops = np.repeat(np.array([O.OPS[n] for n in names], dtype=np.uint8), 4)
        n = ops.size
        pa, pb, pc = (rng.integers(0, 2, n, dtype=np.uint8) for _ in range(3))
        ca, cb, cc = (O.encrypt_bits(s, keys, b) for s, b in ((101, pa), (102, pb), (103, pc)))
        for name, arr in (("ops", ops), ("ca", ca), ("cb", cb), ("cc", cc)):
            arr.tofile(d / f"{name}.bin")
        O.ref("gates", d, d / "ops.bin", d / "ca.bin", d / "cb.bin", d / "cc.bin", d / "go.bin", 8)
        g["gate_ops"], g["gate_pa"], g["gate_pb"], g["gate_pc"] = ops, pa, pb, pc
        g["gate_enc_seeds"] = np.array([101, 102, 103], dtype=np.uint64)
        g["gate_out_tfhepp"] = np.fromfile(d / "go.bin", dtype=O.T0).reshape(n, O.TLWE0)
        # blind rotation + extraction (gatebootstrapping.hpp:188-197): lvl1 TLWE of the reference
        c = (-ca[:4].astype(np.int64) - cb[:4].astype(np.int64)).astype(O.T0)
        c[:, O.N0] += O.T0(O.MU0)
        c.tofile(d / "rot.bin")
        O.ref("blindrotate", d, d / "rot.bin", d / "rot_out.bin")
        g["br_in"] = c
        g["br_out_tfhepp"] = np.fromfile(d / "rot_out.bin", dtype=np.uint32).reshape(4, 1025)
        # IdentityKeySwitch<lvl10param> (keyswitch.hpp:11-52) on the oracle's exact lvl1 samples: exact
        u = O.bootstrap_to_lvl1(keys, c)
        u.tofile(d / "ks_in.bin")
        O.ref("keyswitch", d, d / "ks_in.bin", d / "ks_out.bin")
        g["ks_in"], g["ks_out_tfhepp"] = u, np.fromfile(d / "ks_out.bin", dtype=O.T0).reshape(4, O.TLWE0)
    out = Path(__file__).with_name(f"tfhepp_golden{O.FLAVOUR}.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
