"""Iyokan wire formats: cereal PortableBinary packets and key files.

Reads and writes exactly what `iyokan-packet` produces, so keys and encrypted request packets made by the
reference's own tools can be fed to this back-end and its result packets decrypted by them:

  TFHEPacket  (src/packet.hpp:208-220)   ar(ram, ramInTLWE, rom, romInTLWE, bits, numCycles)
  PlainPacket (src/packet.hpp:193-206)   ar(ram, rom, bits, numCycles), Bit = 1 byte
  SecretKey   (TFHEpp include/key.hpp:27-36)        ar(key.lvl0, key.lvl1, key.lvl2, params)
  EvalKey     (TFHEpp include/cloudkey.hpp:362-368) ar(params, bklvl01, bklvl02, bkfftlvl01, ..., iksklvl10, ...)

cereal PortableBinary layout: one flag byte (1 = little endian), then fields in declaration order;
unordered_map = u64 count + (key, value) pairs, string = u64 length + bytes, vector = u64 count + elements,
std::array of arithmetic = raw bytes, optional = 1 byte (0 = engaged) + value, shared_ptr = u32 id
(msb set = first occurrence, object follows; 0 = null).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

TLWE0_BYTES = 637 * 2
TRLWE1_BYTES = 2 * 1024 * 4
SK_KEY_BYTES = 636 * 2 + 1024 * 4 + 2048 * 8

# serialised sizes of the EvalKey members that precede iksklvl10 (128-bit parameter set)
_BK01 = 636 * 6 * 2 * 1024 * 4
_BK02 = 636 * 8 * 2 * 2048 * 8
_BKFFT01 = 636 * 6 * 2 * 1024 * 8
_BKFFT02 = 636 * 8 * 2 * 2048 * 8
_BKNTT01, _BKNTT02 = _BKFFT01, _BKFFT02
_IKSK10 = 1024 * 7 * 3 * 637 * 2
_EVALKEY_ORDER = [("bklvl01", _BK01), ("bklvl02", _BK02), ("bkfftlvl01", _BKFFT01), ("bkfftlvl02", _BKFFT02),
                  ("bknttlvl01", _BKNTT01), ("bknttlvl02", _BKNTT02), ("iksklvl10", _IKSK10)]


class PacketError(ValueError):
    pass


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.o = memoryview(data), 0
        if self.u8() != 1:
            raise PacketError("not a little-endian cereal PortableBinary archive")

    def take(self, n):
        if self.o + n > len(self.d):
            raise PacketError("truncated archive")
        v = self.d[self.o:self.o + n]
        self.o += n
        return v

    def u8(self):
        return self.take(1)[0]

    def u32(self):
        return struct.unpack("<I", self.take(4))[0]

    def i32(self):
        return struct.unpack("<i", self.take(4))[0]

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def string(self):
        return bytes(self.take(self.u64())).decode()

    def map_of_vectors(self, elem_bytes, dtype, shape):
        out = {}
        for _ in range(self.u64()):
            name = self.string()
            n = self.u64()
            out[name] = np.frombuffer(self.take(n * elem_bytes), dtype=dtype).reshape((n,) + shape).copy()
        return out

    def optional_int(self):
        return None if self.u8() else self.i32()


class _Writer:
    def __init__(self):
        self.parts = [b"\x01"]

    def u64(self, v):
        self.parts.append(struct.pack("<Q", v))

    def string(self, s):
        b = s.encode()
        self.u64(len(b))
        self.parts.append(b)

    def map_of_vectors(self, m, dtype):
        self.u64(len(m))
        for name, arr in m.items():
            self.string(name)
            arr = np.ascontiguousarray(arr, dtype=dtype)
            self.u64(arr.shape[0])
            self.parts.append(arr.tobytes())

    def optional_int(self, v):
        if v is None:
            self.parts.append(b"\x01")
        else:
            self.parts.append(b"\x00" + struct.pack("<i", v))

    def bytes(self):
        return b"".join(self.parts)


@dataclass
class TFHEPacket:
    ram: dict = field(default_factory=dict)         # name -> [n][2][1024] uint32 (CMUX memory; passed through)
    ram_in_tlwe: dict = field(default_factory=dict)  # name -> [n][637] uint16
    rom: dict = field(default_factory=dict)
    rom_in_tlwe: dict = field(default_factory=dict)
    bits: dict = field(default_factory=dict)        # port -> [n][637] uint16
    num_cycles: int | None = None

    @staticmethod
    def loads(data: bytes) -> "TFHEPacket":
        r = _Reader(data)
        p = TFHEPacket()
        p.ram = r.map_of_vectors(TRLWE1_BYTES, np.uint32, (2, 1024))
        p.ram_in_tlwe = r.map_of_vectors(TLWE0_BYTES, np.uint16, (637,))
        p.rom = r.map_of_vectors(TRLWE1_BYTES, np.uint32, (2, 1024))
        p.rom_in_tlwe = r.map_of_vectors(TLWE0_BYTES, np.uint16, (637,))
        p.bits = r.map_of_vectors(TLWE0_BYTES, np.uint16, (637,))
        p.num_cycles = r.optional_int()
        if r.o != len(r.d):
            raise PacketError("trailing bytes in TFHEPacket")
        return p

    def dumps(self) -> bytes:
        w = _Writer()
        w.map_of_vectors(self.ram, np.uint32)
        w.map_of_vectors(self.ram_in_tlwe, np.uint16)
        w.map_of_vectors(self.rom, np.uint32)
        w.map_of_vectors(self.rom_in_tlwe, np.uint16)
        w.map_of_vectors(self.bits, np.uint16)
        w.optional_int(self.num_cycles)
        return w.bytes()

    @staticmethod
    def load(path) -> "TFHEPacket":
        return TFHEPacket.loads(Path(path).read_bytes())

    def save(self, path):
        Path(path).write_bytes(self.dumps())


@dataclass
class PlainPacket:
    ram: dict = field(default_factory=dict)   # name -> uint8 bits
    rom: dict = field(default_factory=dict)
    bits: dict = field(default_factory=dict)
    num_cycles: int | None = None

    @staticmethod
    def loads(data: bytes) -> "PlainPacket":
        r = _Reader(data)
        p = PlainPacket()
        p.ram = r.map_of_vectors(1, np.uint8, ())
        p.rom = r.map_of_vectors(1, np.uint8, ())
        p.bits = r.map_of_vectors(1, np.uint8, ())
        p.num_cycles = r.optional_int()
        return p

    def dumps(self) -> bytes:
        w = _Writer()
        for m in (self.ram, self.rom, self.bits):
            w.map_of_vectors(m, np.uint8)
        w.optional_int(self.num_cycles)
        return w.bytes()


    @staticmethod
    def load(path) -> "PlainPacket":
        return PlainPacket.loads(Path(path).read_bytes())

    def save(self, path):
        Path(path).write_bytes(self.dumps())

    # TOML form of `iyokan-packet toml2packet / packet2toml` (src/iyokan-packet.cpp:253-326; test/in/*.in):
    # cycles = N, [[bits]] / [[ram]] / [[rom]] entries with name, size (bits) and little-endian bytes
    @staticmethod
    def from_toml(text: str) -> "PlainPacket":
        import tomllib

        t = tomllib.loads(text)
        p = PlainPacket(num_cycles=t.get("cycles"))
        for kind, dst in (("bits", p.bits), ("ram", p.ram), ("rom", p.rom)):
            for e in t.get(kind, []):
                size, data = int(e["size"]), list(e["bytes"])
                bits = np.zeros(size, np.uint8)
                for i in range(min(size, 8 * len(data))):
                    bits[i] = (int(data[i // 8]) >> (i % 8)) & 1
                dst[e["name"]] = bits
        return p

    def to_toml(self) -> str:
        lines = [] if self.num_cycles is None else [f"cycles = {self.num_cycles}", ""]
        for kind, src in (("ram", self.ram), ("rom", self.rom), ("bits", self.bits)):
            for name in sorted(src):
                bits = np.asarray(src[name], np.uint8) & 1
                data = [0] * ((bits.size + 7) // 8)
                for i, b in enumerate(bits):
                    data[i // 8] |= int(b) << (i % 8)
                lines += [f"[[{kind}]]", f'name = "{name}"', f"size = {bits.size}", f"bytes = {data}", ""]
        return "\n".join(lines)


LWE_PARAMS_BYTES = 112  # serialised lweParams of the pinned TFHEpp (128-bit set): what precedes bklvl01 in an EvalKey


def read_secret_key_lvl0(path) -> np.ndarray:
    """lvl0 secret key (636 x uint16, binary) from an `iyokan-packet genkey` file."""
    data = Path(path).read_bytes()
    if data[0] != 1 or len(data) < 1 + SK_KEY_BYTES:
        raise PacketError("not a TFHEpp SecretKey archive")
    return np.frombuffer(data, dtype=np.uint16, count=636, offset=1).copy()


def secret_key_params_bytes(path) -> int:
    """Serialised size of lweParams, measured from a SecretKey file (fields precede it there, follow it nowhere)."""
    return Path(path).stat().st_size - 1 - SK_KEY_BYTES


def read_eval_key(path, params_bytes: int = LWE_PARAMS_BYTES):
    """(bklvl01 raw uint32 [636][6][2][1024], iksklvl10 uint16 [1024][7][3][637]) from `iyokan-packet genevalkey`.

    Streams only the two members the gate path needs out of the ~2.2 GB file."""
    path = Path(path)
    size = path.stat().st_size
    with open(path, "rb") as f:
        if f.read(1) != b"\x01":
            raise PacketError("not a little-endian cereal PortableBinary archive")
        off = 1 + params_bytes
        found, seen = {}, []
        for name, nbytes in _EVALKEY_ORDER:
            f.seek(off)
            raw = f.read(4)
            if len(raw) != 4:
                raise PacketError("truncated EvalKey")
            (pid,) = struct.unpack("<I", raw)
            off += 4
            if pid == 0:
                continue
            if not pid & 0x80000000 or (pid & 0x7FFFFFFF) != len(seen) + 1:
                raise PacketError(f"EvalKey layout not recognised at {name} (shared_ptr id {pid:#x}): "
                                  "was the file written by `iyokan-packet genevalkey` of the pinned TFHEpp?")
            seen.append(name)
            if off + nbytes > size:
                raise PacketError(f"truncated EvalKey while reading {name}")
            if name in ("bklvl01", "iksklvl10"):
                found[name] = off
            off += nbytes
    if "bklvl01" not in found or "iksklvl10" not in found:
        raise PacketError("EvalKey lacks bklvl01 / iksklvl10 (run iyokan-packet genevalkey)")
    bk = np.fromfile(path, dtype=np.uint32, count=_BK01 // 4, offset=found["bklvl01"]).reshape(636, 6, 2, 1024)
    ksk = np.fromfile(path, dtype=np.uint16, count=_IKSK10 // 2, offset=found["iksklvl10"]).reshape(1024, 7, 3, 637)
    return bk, ksk
