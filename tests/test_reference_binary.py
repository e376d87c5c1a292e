"""Differential tests against the reference's OWN binary (`iyokan`, compiled unmodified from its sources into
oracle/_ref/ by `make -C oracle refbin`): the same hand-written blueprints and request packets through `iyokan plain`
/ `iyokan tfhe` and through `python -m iyokan_b200 plain` / `tfhe` must give the same result packets."""
import os
import time
from pathlib import Path

import numpy as np
import pytest

import oracle as O
from iyokan_b200.cli import main
from iyokan_b200.packet import PlainPacket

FIX = Path(__file__).resolve().parent / "fixtures"
needs_ref = pytest.mark.skipif(not (O.IYOKAN_REF.exists() and O.have_iyokan_packet()), reason="oracle/_ref/iyokan not built")


def cli(*argv):
    try:
        return main([str(a) for a in argv])
    except SystemExit as e:
        return e.code


def lookup_request():
    return PlainPacket(rom={"rom": np.array([(w >> b) & 1 for w in (0x3, 0xC, 0x5, 0xA) for b in range(4)], np.uint8)},
                       ram={"ram": np.array([(w >> b) & 1 for w in (0xF, 0x0, 0x9, 0x6) for b in range(4)], np.uint8)},
                       bits={"raddr": np.array([1, 0, 0, 1, 1, 1, 0, 0], np.uint8), "waddr": np.array([0, 0, 1, 0, 0, 1], np.uint8),
                             "wren": np.array([1, 1, 0, 1], np.uint8)})


def same_packet(a: PlainPacket, b: PlainPacket):
    assert a.num_cycles == b.num_cycles
    for x, y in ((a.bits, b.bits), (a.ram, b.ram)):
        assert sorted(x) == sorted(y)
        for k in x:
            assert np.array_equal(x[k], y[k]), k


@needs_ref
@pytest.mark.parametrize("blueprint,cycles", [("counter-2bit.toml", 5), ("lookup.toml", 6), ("lookup-cmux.toml", 6)])
def test_plain_mode_equals_the_reference_binary(tmp_path, blueprint, cycles):
    req = lookup_request() if blueprint.startswith("lookup") else PlainPacket()
    (tmp_path / "req.toml").write_text(req.to_toml())
    O.iyokan_packet("toml2packet", "--in", tmp_path / "req.toml", "--out", tmp_path / "req")
    r = O.iyokan_ref("plain", "--blueprint", FIX / blueprint, "-i", tmp_path / "req", "-o", tmp_path / "ref.res", "-c", cycles,
                     "--dump-prefix", tmp_path / "ref.dump")
    assert r.returncode == 0, r.stdout + r.stderr
    assert cli("plain", "--blueprint", FIX / blueprint, "-i", tmp_path / "req", "-o", tmp_path / "our.res", "-c", cycles,
               "--dump-prefix", tmp_path / "our.dump", "--quiet") == 0
    same_packet(PlainPacket.load(tmp_path / "our.res"), PlainPacket.load(tmp_path / "ref.res"))
    for c in range(cycles):  # every intermediate cycle too: outputs and RAM image as the reference dumps them
        same_packet(PlainPacket.load(f"{tmp_path / 'our.dump'}-{c}"), PlainPacket.load(f"{tmp_path / 'ref.dump'}-{c}"))


@needs_ref
@pytest.mark.gpu
def test_tfhe_mode_equals_the_reference_binary(tmp_path):
    """Same keys, same encrypted request: `iyokan tfhe` on the host cores vs this back-end on the GPU."""
    sk, ek = tmp_path / "sk", tmp_path / "ek"
    O.iyokan_packet("genkey", "--type", "tfhepp", "--out", sk)
    O.iyokan_packet("genevalkey", "--in", sk, "--out", ek)
    (tmp_path / "req.toml").write_text(lookup_request().to_toml())
    O.iyokan_packet("toml2packet", "--in", tmp_path / "req.toml", "--out", tmp_path / "req")
    O.iyokan_packet("enc", "--key", sk, "--in", tmp_path / "req", "--out", tmp_path / "req.enc")
    bp, cycles = FIX / "lookup.toml", 5
    t0 = time.time()
    r = O.iyokan_ref("tfhe", "--blueprint", bp, "--evalkey", ek, "-i", tmp_path / "req.enc", "-o", tmp_path / "ref.enc", "-c", cycles,
                     "--cpu", os.cpu_count() or 1)
    t_ref = time.time() - t0
    assert r.returncode == 0, r.stdout + r.stderr
    t0 = time.time()
    assert cli("tfhe", "--blueprint", bp, "--evalkey", ek, "-i", tmp_path / "req.enc", "-o", tmp_path / "our.enc", "-c", cycles,
               "--quiet") == 0
    t_our = time.time() - t0
    for name in ("ref", "our"):
        O.iyokan_packet("dec", "--key", sk, "--in", tmp_path / f"{name}.enc", "--out", tmp_path / f"{name}.res")
    same_packet(PlainPacket.load(tmp_path / "our.res"), PlainPacket.load(tmp_path / "ref.res"))
    print(f"reference iyokan tfhe: {t_ref:.2f} s on {os.cpu_count()} cores; this back-end: {t_our:.2f} s (both incl. key loading)")
    ek.unlink()
