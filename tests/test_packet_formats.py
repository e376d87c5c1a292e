"""Iyokan wire formats (cereal PortableBinary) against files written by the reference's own iyokan-packet
(tests/golden/packets/: `genkey`, `toml2packet`, `enc` run in the build container)."""
from pathlib import Path

import numpy as np
import pytest

import oracle as O
from iyokan_b200 import packet as K

PK = Path(__file__).resolve().parent / "golden" / "packets"


def test_reads_reference_encrypted_packet_and_key():
    pkt = K.TFHEPacket.load(PK / "req.enc")
    assert set(pkt.bits) == {"io_inA", "io_inB"} and pkt.num_cycles == 1
    assert not pkt.ram and not pkt.rom and not pkt.ram_in_tlwe and not pkt.rom_in_tlwe
    sk0 = K.read_secret_key_lvl0(PK / "sk")
    assert sk0.shape == (636,) and set(np.unique(sk0)) <= {0, 1}
    keys = O.Keys(sk0, None, None, None)
    # the reference encrypted io_inA = 2, io_inB = 4 (bit i of a byte is element i, src/iyokan-packet.cpp:44-57)
    assert list(O.decrypt_bits(keys, pkt.bits["io_inA"])) == [0, 1, 0, 0]
    assert list(O.decrypt_bits(keys, pkt.bits["io_inB"])) == [0, 0, 1, 0]
    assert K.secret_key_params_bytes(PK / "sk") == 112


def test_writer_is_byte_identical_to_the_reference():
    raw = (PK / "req.enc").read_bytes()
    assert K.TFHEPacket.loads(raw).dumps() == raw
    rawp = (PK / "req").read_bytes()
    pp = K.PlainPacket.loads(rawp)
    assert list(pp.bits["io_inA"]) == [0, 1, 0, 0] and pp.num_cycles == 1
    assert pp.dumps() == rawp


def test_round_trip_with_memories_and_no_cycles():
    rng = np.random.default_rng(0)
    p = K.TFHEPacket(ram={"ram": rng.integers(0, 2**32, (3, 2, 1024), dtype=np.uint32)},
                     ram_in_tlwe={"ram": rng.integers(0, 2**16, (5, 637), dtype=np.uint16)},
                     rom_in_tlwe={"rom": rng.integers(0, 2**16, (2, 637), dtype=np.uint16)},
                     bits={"a": rng.integers(0, 2**16, (1, 637), dtype=np.uint16), "b": np.zeros((0, 637), np.uint16)})
    q = K.TFHEPacket.loads(p.dumps())
    assert q.num_cycles is None
    for a, b in ((p.ram, q.ram), (p.ram_in_tlwe, q.ram_in_tlwe), (p.rom_in_tlwe, q.rom_in_tlwe), (p.bits, q.bits)):
        assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)


def test_malformed_archives_are_rejected():
    raw = (PK / "req.enc").read_bytes()
    with pytest.raises(K.PacketError):
        K.TFHEPacket.loads(b"\x00" + raw[1:])      # big-endian flag
    with pytest.raises(K.PacketError):
        K.TFHEPacket.loads(raw[:-3])               # truncated
    with pytest.raises(K.PacketError):
        K.TFHEPacket.loads(raw + b"\x00")          # trailing garbage


@pytest.mark.skipif(not O.have_iyokan_packet(), reason="oracle/_ref/iyokan-packet not built")
def test_reference_tool_decrypts_our_packets(tmp_path):
    # a packet we write must be readable by the reference's `dec` + `packet2toml`
    sk0 = K.read_secret_key_lvl0(PK / "sk")
    keys = O.Keys(sk0, None, None, None)
    bits = np.array([1, 0, 1, 1, 0, 0, 1, 0], np.uint8)
    K.TFHEPacket(bits={"out": O.encrypt_bits(3, keys, bits)}, num_cycles=5).save(tmp_path / "res.enc")
    O.iyokan_packet("dec", "--key", PK / "sk", "--in", tmp_path / "res.enc", "--out", tmp_path / "res")
    toml = O.iyokan_packet("packet2toml", "--in", tmp_path / "res")
    assert "cycles" in toml and "5" in toml
    pp = K.PlainPacket.loads((tmp_path / "res").read_bytes())
    assert list(pp.bits["out"]) == list(bits) and pp.num_cycles == 5
