"""Encrypted netlist runs on the GPU through the host engine (b200net -> b200fhe C ABI), against the
reference's golden result packets and, for small circuits, bit-exactly against the oracle evaluated
level by level on the CPU."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle as O
from iyokan_b200 import netlist as N

pytestmark = pytest.mark.gpu
NL = Path(__file__).resolve().parent / "golden" / "netlists"
CASES = json.load(open(NL / "cases.json"))


class OracleNet:
    """CPU model of the engine: same slot arena semantics, gates evaluated by the exact-integer oracle."""

    def __init__(self, nl, keys):
        self.nl, self.keys = nl, keys
        self.eng = N.NetEngine(nl)
        self.ct = np.zeros((nl.n, 637), np.uint16)
        self.level = np.array([self.eng.lib.b200net_node_level(self.eng._h, i) for i in range(nl.n)])

    def resolve(self, n):
        while self.nl.kind[n] == N.OUTPUT:
            n = self.nl.in0[n]
        return n

    def run(self):
        for lv in range(1, self.level.max() + 1):
            g = np.nonzero((self.level == lv) & (self.nl.kind < 15))[0]
            if g.size == 0:
                continue
            ins = []
            for arr in (self.nl.in0, self.nl.in1, self.nl.in2):
                idx = np.array([self.resolve(a) if a >= 0 else 0 for a in arr[g]])
                ins.append(self.ct[idx])
            self.ct[g] = O.gate_batch(self.keys, self.nl.kind[g], *ins)

    def tick(self):
        d = np.nonzero(self.nl.kind == N.DFF)[0]
        self.ct[d] = self.ct[[self.resolve(self.nl.in0[i]) for i in d]]


def test_counter_bit_exact_vs_oracle_and_golden(gpu_ctx, keys):
    nl = N.Netlist.load(NL / "counter-4bit.npz")
    enc = lambda bits: O.encrypt_bits(99, keys, bits)  # noqa: E731
    r = N.EncryptedRunner(nl, gpu_ctx, enc)
    out = r.run(CASES["counter-4bit"]["cycles"])
    bits = O.decrypt_bits(keys, out["out"])
    assert N.bytes_of(bits) == CASES["counter-4bit"]["expected"]["bits"]["out"]["bytes"]   # out == 2 (test13.out)
    # oracle replay of the same protocol: reset pass, then 3 x (tick, run)
    o = OracleNet(nl, keys)
    dffs = np.nonzero(nl.kind == N.DFF)[0]
    o.ct[dffs] = N.trivial(0)
    rst = nl.in_ports["reset"][0]
    o.ct[rst] = N.trivial(1)
    o.run()
    for c in range(3):
        o.tick()
        if c == 0:
            o.ct[rst] = N.trivial(0)
        o.run()
    want = o.ct[[o.resolve(n) for n in nl.out_ports["out"]]]
    assert np.array_equal(out["out"], want)


def test_adder_fresh_inputs(gpu_ctx, keys):
    nl = N.ripple_adder(4)
    r = N.EncryptedRunner(nl, gpu_ctx, lambda bits: O.encrypt_bits(5, keys, bits))
    out = r.run(1, inputs={"a": N.bits_of([11], 4), "b": N.bits_of([6], 4)})
    assert N.bytes_of(O.decrypt_bits(keys, out["sum"]))[0] == 17


@pytest.mark.parametrize("name", ["addr-4bit", "div-8bit"])
def test_reference_cases_encrypted(gpu_ctx, keys, name):
    case = CASES[name]
    nl = N.Netlist.load(NL / f"{name}.npz")
    r = N.EncryptedRunner(nl, gpu_ctx, lambda bits: O.encrypt_bits(17, keys, bits))
    inputs = {p: N.bits_of(e["bytes"], e["size"]) for p, e in case["request"]["bits"].items()}
    out = r.run(case["cycles"], inputs=inputs)
    for port, e in case["expected"]["bits"].items():
        assert N.bytes_of(O.decrypt_bits(keys, out[port])[:e["size"]]) == e["bytes"], port


def test_cahp_pearl_processor_runs_the_program(gpu_ctx, keys):
    """VSP CAHP-pearl with MUX ROM/RAM (33k nodes, ~30.8k bootstraps/cycle): 3 cycles -> reg_x0 == 42."""
    case = CASES["cahp-pearl-mux"]
    nl = N.Netlist.load(NL / "cahp-pearl-mux.npz")
    r = N.EncryptedRunner(nl, gpu_ctx, lambda bits: O.encrypt_bits(23, keys, bits))
    req = case["request"]
    out = r.run(case["cycles"], rams={p: N.bits_of(e["bytes"], e["size"]) for p, e in req["ram"].items()},
                roms={p: N.bits_of(e["bytes"], e["size"]) for p, e in req["rom"].items()})
    for port, e in case["expected"]["bits"].items():
        assert N.bytes_of(O.decrypt_bits(keys, out[port])[:e["size"]]) == e["bytes"], port
    ram = O.decrypt_bits(keys, r.get_mem("ram"))
    e = case["expected"]["ram"]["ram"]
    assert N.bytes_of(ram[:e["size"]]) == e["bytes"]


def _run_case_encrypted(gpu_ctx, keys, name, seed):
    case = CASES[name]
    nl = N.Netlist.load(NL / f"{name}.npz")
    r = N.EncryptedRunner(nl, gpu_ctx, lambda bits: O.encrypt_bits(seed, keys, bits))
    req = case["request"]
    out = r.run(case["cycles"], inputs={p: N.bits_of(e["bytes"], e["size"]) for p, e in req["bits"].items()},
                rams={p: N.bits_of(e["bytes"], e["size"]) for p, e in req["ram"].items()},
                roms={p: N.bits_of(e["bytes"], e["size"]) for p, e in req["rom"].items()})
    for port, e in case["expected"]["bits"].items():
        assert N.bytes_of(O.decrypt_bits(keys, out[port])[:e["size"]]) == e["bytes"], (name, port)
    for mem, e in case["expected"]["ram"].items():
        got = O.decrypt_bits(keys, r.get_mem(mem))
        assert N.bytes_of(got[:e["size"]]) == e["bytes"], (name, mem)


def test_mux_ram_golden_eight_cycles(gpu_ctx, keys):
    """test/in/test08.in -> test/out/test08.out: 8 write/read cycles, rdata and the whole 4096-bit RAM image."""
    _run_case_encrypted(gpu_ctx, keys, "mux-ram-8-16-16", 29)


def test_cahp_ruby_processor_golden(gpu_ctx, keys):
    """VSP CAHP-ruby with MUX ROM/RAM, test09 program, 7 cycles (test/out/test09-ruby.out)."""
    _run_case_encrypted(gpu_ctx, keys, "cahp-ruby-mux", 31)


@pytest.mark.skipif(not O.have_iyokan_packet(), reason="oracle/_ref/iyokan-packet not built")
def test_drop_in_with_the_reference_packet_tools(tmp_path):
    """Keys and encrypted request from the reference's own `iyokan-packet` (genkey, genevalkey, enc); this
    back-end evaluates the netlist; the reference's `dec` + plain packets give the result (test.rb:287-315
    flow with `iyokan tfhe` replaced).  Cases: addr-4bit-04 and counter-4bit-13."""
    from iyokan_b200 import Context, packet as K

    sk, ek = tmp_path / "sk", tmp_path / "ek"
    O.iyokan_packet("genkey", "--type", "tfhepp", "--out", sk)
    O.iyokan_packet("genevalkey", "--in", sk, "--out", ek)            # ~2.2 GB, as the reference writes it
    bk, ksk = K.read_eval_key(ek, K.secret_key_params_bytes(sk))
    ek.unlink()
    with Context(0) as ctx:
        ctx.load_keys(bk, ksk)
        for name in ("addr-4bit", "counter-4bit"):
            case = CASES[name]
            lines = [f"cycles = {case['cycles']}"]
            for port, e in case["request"]["bits"].items():
                lines += ["[[bits]]", f'name = "{port}"', f"size = {e['size']}", f"bytes = {e['bytes']}"]
            (tmp_path / "req.in").write_text("\n".join(lines) + "\n")
            O.iyokan_packet("toml2packet", "--in", tmp_path / "req.in", "--out", tmp_path / "req")
            O.iyokan_packet("enc", "--key", sk, "--in", tmp_path / "req", "--out", tmp_path / "req.enc")
            res = N.run_packet(N.Netlist.load(NL / f"{name}.npz"), ctx, K.TFHEPacket.load(tmp_path / "req.enc"))
            res.save(tmp_path / "res.enc")
            O.iyokan_packet("dec", "--key", sk, "--in", tmp_path / "res.enc", "--out", tmp_path / "res")
            plain = K.PlainPacket.loads((tmp_path / "res").read_bytes())
            assert plain.num_cycles == case["cycles"]
            for port, e in case["expected"]["bits"].items():
                assert N.bytes_of(plain.bits[port][:e["size"]]) == e["bytes"], (name, port)


@pytest.mark.skipif(not O.have_iyokan_packet(), reason="oracle/_ref/iyokan-packet not built")
def test_cli_tfhe_end_to_end_with_reference_tools_and_resume(tmp_path):
    """`python -m iyokan_b200 tfhe` in place of `iyokan tfhe` (test.rb:287-315): blueprint -> netlist, EvalKey and
    encrypted request written by the reference's iyokan-packet, result decrypted by it; equals the plain run.
    Then snapshot after 2 cycles + resume for 2 more == one 4-cycle run, byte for byte."""
    from iyokan_b200.cli import main
    from iyokan_b200.packet import PlainPacket

    fix = Path(__file__).resolve().parent / "fixtures"
    sk, ek = tmp_path / "sk", tmp_path / "ek"
    O.iyokan_packet("genkey", "--type", "tfhepp", "--out", sk)
    O.iyokan_packet("genevalkey", "--in", sk, "--out", ek)
    rom_words = [0x3, 0xC, 0x5, 0xA]
    req = PlainPacket(rom={"rom": np.array([(w >> b) & 1 for w in rom_words for b in range(4)], np.uint8)},
                      ram={"ram": np.zeros(16, np.uint8)},
                      bits={"raddr": np.array([1, 0, 0, 1, 1, 1], np.uint8), "waddr": np.array([0, 0, 1, 0, 0, 1], np.uint8),
                            "wren": np.array([1, 1, 0], np.uint8)})
    (tmp_path / "req.toml").write_text(req.to_toml())
    O.iyokan_packet("toml2packet", "--in", tmp_path / "req.toml", "--out", tmp_path / "req")
    O.iyokan_packet("enc", "--key", sk, "--in", tmp_path / "req", "--out", tmp_path / "req.enc")

    def cli(*argv):
        try:
            return main([str(a) for a in argv])
        except SystemExit as e:
            return e.code

    bp = fix / "lookup.toml"
    assert cli("tfhe", "--blueprint", bp, "--evalkey", ek, "-i", tmp_path / "req.enc", "-o", tmp_path / "res.enc", "-c", 4,
               "--quiet") == 0
    O.iyokan_packet("dec", "--key", sk, "--in", tmp_path / "res.enc", "--out", tmp_path / "res")
    assert cli("plain", "--blueprint", bp, "-i", tmp_path / "req", "-o", tmp_path / "res.plain", "-c", 4, "--quiet") == 0
    got, want = PlainPacket.load(tmp_path / "res"), PlainPacket.load(tmp_path / "res.plain")
    assert got.num_cycles == want.num_cycles == 4
    assert np.array_equal(got.bits["rdata"], want.bits["rdata"]) and np.array_equal(got.ram["ram"], want.ram["ram"])
    # the same design declared with CMUX memories (type = "rom" / "ram") is evaluated as MUX memories; its RAM image
    # follows the CMUX convention (the last cycle's write is already visible)
    assert cli("tfhe", "--blueprint", fix / "lookup-cmux.toml", "--evalkey", ek, "-i", tmp_path / "req.enc", "-o",
               tmp_path / "res_cmux.enc", "-c", 4, "--quiet") == 0
    O.iyokan_packet("dec", "--key", sk, "--in", tmp_path / "res_cmux.enc", "--out", tmp_path / "res_cmux")
    assert cli("plain", "--blueprint", fix / "lookup-cmux.toml", "-i", tmp_path / "req", "-o", tmp_path / "res_cmux.plain", "-c", 4,
               "--quiet") == 0
    gc, wc = PlainPacket.load(tmp_path / "res_cmux"), PlainPacket.load(tmp_path / "res_cmux.plain")
    assert np.array_equal(gc.bits["rdata"], wc.bits["rdata"]) and np.array_equal(gc.ram["ram"], wc.ram["ram"])
    # snapshot / resume on the encrypted back-end
    assert cli("tfhe", "--blueprint", bp, "--evalkey", ek, "-i", tmp_path / "req.enc", "-o", tmp_path / "half.enc", "-c", 2,
               "--snapshot", tmp_path / "snap", "--quiet") == 0
    assert cli("tfhe", "--evalkey", ek, "--resume", tmp_path / "snap", "-o", tmp_path / "resumed.enc", "-c", 2, "--quiet") == 0
    assert (tmp_path / "resumed.enc").read_bytes() == (tmp_path / "res.enc").read_bytes()
    ek.unlink()


@pytest.mark.skipif(not O.have_iyokan_packet(), reason="oracle/_ref/iyokan-packet not built")
def test_cli_tfhe_two_gpus(tmp_path):
    """`torchrun --nproc-per-node 2 -m iyokan_b200 tfhe`: same result bits as the plain run (needs 2 GPUs)."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from iyokan_b200.cli import main
    from iyokan_b200.packet import PlainPacket

    root = Path(__file__).resolve().parents[1]
    bp = root / "tests" / "fixtures" / "lookup.toml"
    sk, ek = tmp_path / "sk", tmp_path / "ek"
    O.iyokan_packet("genkey", "--type", "tfhepp", "--out", sk)
    O.iyokan_packet("genevalkey", "--in", sk, "--out", ek)
    req = PlainPacket(rom={"rom": np.array([(w >> b) & 1 for w in (7, 8, 2, 13) for b in range(4)], np.uint8)},
                      bits={"raddr": np.array([1, 1, 0, 1], np.uint8), "waddr": np.array([0, 1, 1, 0], np.uint8),
                            "wren": np.array([1], np.uint8)})
    (tmp_path / "req.toml").write_text(req.to_toml())
    O.iyokan_packet("toml2packet", "--in", tmp_path / "req.toml", "--out", tmp_path / "req")
    O.iyokan_packet("enc", "--key", sk, "--in", tmp_path / "req", "--out", tmp_path / "req.enc")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", "-m", "iyokan_b200", "tfhe", "--blueprint", str(bp), "--evalkey", str(ek), "-i",
           str(tmp_path / "req.enc"), "-o", str(tmp_path / "res.enc"), "-c", "3", "--quiet"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    O.iyokan_packet("dec", "--key", sk, "--in", tmp_path / "res.enc", "--out", tmp_path / "res")
    try:
        main(["plain", "--blueprint", str(bp), "-i", str(tmp_path / "req"), "-o", str(tmp_path / "res.plain"), "-c", "3", "--quiet"])
    except SystemExit as e:
        assert e.code == 0
    got, want = PlainPacket.load(tmp_path / "res"), PlainPacket.load(tmp_path / "res.plain")
    assert np.array_equal(got.bits["rdata"], want.bits["rdata"]) and np.array_equal(got.ram["ram"], want.ram["ram"])
    ek.unlink()
