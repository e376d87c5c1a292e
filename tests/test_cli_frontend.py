"""`python -m iyokan_b200 plain ...`: blueprint loader, cycle protocol, packets and snapshot/resume on the
plaintext back-end (no GPU).  Hand-written fixtures run everywhere; the reference's own test.rb cases
(test/config-toml + test/in -> test/out) run where the reference tree is mounted: every blueprint of test.rb,
those declared with CMUX memories included (evaluated as MUX memories)."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from iyokan_b200 import blueprint as B
from iyokan_b200.cli import main
from iyokan_b200.frontend import Frontend, FrontendError
from iyokan_b200.packet import PlainPacket

ROOT = Path(__file__).resolve().parents[1]
FIX = ROOT / "tests" / "fixtures"
REF = Path("/root/reference")


def word(bits):
    return sum(int(b) << i for i, b in enumerate(bits))


def run_cli(*argv):
    try:
        return main([str(a) for a in argv])
    except SystemExit as e:
        return e.code


def test_counter_blueprint_yosys_reader_and_protocol(tmp_path):
    nl = B.read_blueprint(FIX / "upcount2.toml")
    assert sorted(nl.in_ports) == ["reset"] and sorted(nl.out_ports) == ["one", "out"]
    (tmp_path / "req.in").write_text("cycles = 3\n")
    for cycles, want in ((1, 0), (2, 1), (3, 2), (6, 1)):
        assert run_cli("plain", "--blueprint", FIX / "upcount2.toml", "-i", tmp_path / "req.in", "-o",
                       tmp_path / "res", "-c", cycles, "--quiet") == 0
        res = PlainPacket.load(tmp_path / "res")
        assert res.num_cycles == cycles and word(res.bits["out"]) == want and word(res.bits["one"]) == 1
    # cycles from the packet when -c is absent
    assert run_cli("plain", "--blueprint", FIX / "upcount2.toml", "-i", tmp_path / "req.in", "-o", tmp_path / "res",
                   "--quiet") == 0
    assert word(PlainPacket.load(tmp_path / "res").bits["out"]) == 2


def test_rom_ram_builtins_and_circular_inputs(tmp_path):
    rom_words = [0x3, 0xC, 0x5, 0xA]
    rom_bits = [(w >> b) & 1 for w in rom_words for b in range(4)]
    req = PlainPacket(rom={"rom": np.array(rom_bits, np.uint8)}, ram={"ram": np.zeros(16, np.uint8)},
                      bits={"raddr": np.array([1, 0, 0, 1, 1, 1], np.uint8),     # 1, 2, 3 (2 bits per cycle, LSB first)
                            "waddr": np.array([0, 0, 1, 0, 0, 1], np.uint8),     # 0, 1, 2
                            "wren": np.array([1, 1, 0], np.uint8)})              # third cycle does not write
    req.save(tmp_path / "req")
    assert run_cli("plain", "--blueprint", FIX / "lookup.toml", "-i", tmp_path / "req", "-o", tmp_path / "res", "-c", 4,
                   "--quiet") == 0
    res = PlainPacket.load(tmp_path / "res")
    ram = [word(res.ram["ram"][4 * a: 4 * a + 4]) for a in range(4)]
    # cycle 3 wraps around to the first stream entries: raddr 1, waddr 0, wren 1 again (written at the next tick,
    # which never comes: the result shows the RAM as the last tick left it)
    assert ram == [rom_words[1], rom_words[2], 0, 0]
    assert word(res.bits["rdata"]) == ram[0]
    assert "rom" not in res.ram and res.num_cycles == 4
    # declared with CMUX memories (type = "rom" / "ram") the design is built from the same MUX circuits, but its RAM
    # image follows the CMUX convention: a write is visible in the image of the cycle that performs it
    # (checked against the reference's own binary in tests/test_reference_binary.py)
    assert run_cli("plain", "--blueprint", FIX / "lookup-cmux.toml", "-i", tmp_path / "req", "-o", tmp_path / "res_cmux",
                   "-c", 3, "--quiet") == 0
    assert run_cli("plain", "--blueprint", FIX / "lookup.toml", "-i", tmp_path / "req", "-o", tmp_path / "res3", "-c", 3,
                   "--quiet") == 0
    cm, mx = PlainPacket.load(tmp_path / "res_cmux"), PlainPacket.load(tmp_path / "res3")
    assert np.array_equal(cm.bits["rdata"], mx.bits["rdata"])
    # cycle 2 does not write (wren = 0), so after 3 cycles both conventions show the same image ...
    assert np.array_equal(cm.ram["ram"], mx.ram["ram"])
    # ... while after 2 cycles the CMUX image already holds RAM[1] <- ROM[2], the MUX image only after the next tick
    for bp, want in (("lookup-cmux.toml", rom_words[2]), ("lookup.toml", 0)):
        assert run_cli("plain", "--blueprint", FIX / bp, "-i", tmp_path / "req", "-o", tmp_path / "r2", "-c", 2, "--quiet") == 0
        assert word(PlainPacket.load(tmp_path / "r2").ram["ram"][4:8]) == want
    # packet <-> TOML round trip (iyokan-packet packet2toml / toml2packet)
    assert run_cli("packet", "packet2toml", "--in", tmp_path / "res", "--out", tmp_path / "res.toml") == 0
    assert run_cli("packet", "toml2packet", "--in", tmp_path / "res.toml", "--out", tmp_path / "res2") == 0
    assert (tmp_path / "res2").read_bytes() == (tmp_path / "res").read_bytes()


def test_snapshot_resume_equals_one_run(tmp_path):
    bp, req = FIX / "upcount2.toml", tmp_path / "req.in"
    req.write_text("cycles = 1\n")
    assert run_cli("plain", "--blueprint", bp, "-i", req, "-o", tmp_path / "a", "-c", 5, "--quiet") == 0
    assert run_cli("plain", "--blueprint", bp, "-i", req, "-o", tmp_path / "b1", "-c", 2, "--snapshot", tmp_path / "snap",
                   "--quiet") == 0
    assert Frontend.is_snapshot(tmp_path / "snap")
    assert run_cli("plain", "--resume", tmp_path / "snap", "-o", tmp_path / "b2", "-c", 3, "--quiet") == 0
    assert (tmp_path / "a").read_bytes() == (tmp_path / "b2").read_bytes()
    assert PlainPacket.load(tmp_path / "b2").num_cycles == 5
    # a snapshot of the other mode, or a file that is no snapshot, is refused like the reference does
    assert run_cli("tfhe", "--evalkey", tmp_path / "nokey", "--resume", tmp_path / "snap", "-o", tmp_path / "x", "-c", 1) == 1
    assert run_cli("plain", "--resume", req, "-o", tmp_path / "x", "-c", 1) == 1


def test_dump_prefix_writes_one_packet_per_cycle(tmp_path):
    # test.rb "cahp-diamond-dump-prefix-00" in miniature: --dump-prefix P leaves P-0 ... P-(N-1) behind, P-c being the
    # state after c cycles (same numbering as the reference: tests/test_reference_binary.py compares them file by file)
    (tmp_path / "req.in").write_text("cycles = 4\n")
    assert run_cli("plain", "--blueprint", FIX / "upcount2.toml", "-i", tmp_path / "req.in", "-o", tmp_path / "res",
                   "--dump-prefix", tmp_path / "dump", "--quiet") == 0
    for c, want in ((1, 0), (2, 1), (3, 2)):
        p = PlainPacket.load(f"{tmp_path / 'dump'}-{c}")
        assert p.num_cycles == c and word(p.bits["out"]) == want
    assert PlainPacket.load(f"{tmp_path / 'dump'}-0").num_cycles == 0
    assert not Path(f"{tmp_path / 'dump'}-4").exists() and word(PlainPacket.load(tmp_path / "res").bits["out"]) == 3


def test_error_behaviour(tmp_path):
    bp = FIX / "upcount2.toml"
    (tmp_path / "bad.in").write_text('[[bits]]\nname = "nosuchport"\nsize = 1\nbytes = [1]\n')
    assert run_cli("plain", "--blueprint", bp, "-i", tmp_path / "bad.in", "-o", tmp_path / "r", "-c", 1) == 1
    (tmp_path / "rst.in").write_text('[[bits]]\nname = "reset"\nsize = 1\nbytes = [1]\n')
    assert run_cli("plain", "--blueprint", bp, "-i", tmp_path / "rst.in", "-o", tmp_path / "r", "-c", 1) == 1
    (tmp_path / "none.in").write_text("")
    assert run_cli("plain", "--blueprint", bp, "-i", tmp_path / "none.in", "-o", tmp_path / "r") == 1   # no cycle count
    (tmp_path / "odd.toml").write_text('[[builtin]]\ntype = "quantum-ram"\nname = "ram"\nin_addr_width = 8\n'
                                       'in_wdata_width = 8\nout_rdata_width = 8\n')
    assert run_cli("plain", "--blueprint", tmp_path / "odd.toml", "-i", tmp_path / "none.in", "-o", tmp_path / "r",
                   "-c", 1) == 1
    with pytest.raises(FrontendError):
        Frontend(B.read_blueprint(bp), "tfhe", None)   # encrypted mode never falls back to the CPU
    # a RAM / ROM image of the wrong length is refused, as the reference does ("wrong length of RAM",
    # iyokan_tfhepp.cpp:242-259): lookup.toml holds a 16-bit ROM and a 16-bit RAM
    (tmp_path / "short.in").write_text('cycles = 1\n[[ram]]\nname = "ram"\nsize = 8\nbytes = [255]\n')
    assert run_cli("plain", "--blueprint", FIX / "lookup.toml", "-i", tmp_path / "short.in", "-o", tmp_path / "r") == 1
    (tmp_path / "long.in").write_text('cycles = 1\n[[rom]]\nname = "rom"\nsize = 24\nbytes = [1, 2, 3]\n')
    assert run_cli("plain", "--blueprint", FIX / "lookup.toml", "-i", tmp_path / "long.in", "-o", tmp_path / "r") == 1
    (tmp_path / "fit.in").write_text('cycles = 1\n[[rom]]\nname = "rom"\nsize = 16\nbytes = [1, 2]\n')
    assert run_cli("plain", "--blueprint", FIX / "lookup.toml", "-i", tmp_path / "fit.in", "-o", tmp_path / "r") == 0


def test_module_entry_point(tmp_path):
    (tmp_path / "req.in").write_text("cycles = 2\n")
    r = subprocess.run([sys.executable, "-m", "iyokan_b200", "plain", "--blueprint", str(FIX / "upcount2.toml"), "-i",
                        str(tmp_path / "req.in"), "-o", str(tmp_path / "res"), "--stdout-csv"], cwd=ROOT,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "2,out,1" in r.stdout and "done." in r.stderr


def _torchrun(nproc, port, *argv):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), "-m", "iyokan_b200", *[str(a) for a in argv]]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_two_process_run_equals_single_process(tmp_path):
    """`torchrun -m iyokan_b200 plain` (gloo, world size 2): level-sliced sharding + all-gather behind the same
    command line; result and snapshot/resume identical to the single-process run."""
    import socket

    def port():
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        p = s.getsockname()[1]
        s.close()
        return p

    rom_words = [0x9, 0x6, 0xF, 0x1]
    req = PlainPacket(rom={"rom": np.array([(w >> b) & 1 for w in rom_words for b in range(4)], np.uint8)},
                      bits={"raddr": np.array([0, 0, 1, 0, 0, 1, 1, 1], np.uint8), "waddr": np.array([1, 1, 0, 1, 1, 0, 0, 0], np.uint8),
                            "wren": np.array([1], np.uint8)})
    req.save(tmp_path / "req")
    bp = FIX / "lookup.toml"
    assert run_cli("plain", "--blueprint", bp, "-i", tmp_path / "req", "-o", tmp_path / "one", "-c", 5, "--quiet") == 0
    r = _torchrun(2, port(), "plain", "--blueprint", bp, "-i", tmp_path / "req", "-o", tmp_path / "two", "-c", 5, "--quiet")
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "two").read_bytes() == (tmp_path / "one").read_bytes()
    r = _torchrun(2, port(), "plain", "--blueprint", bp, "-i", tmp_path / "req", "-o", tmp_path / "h", "-c", 2, "--snapshot",
                  tmp_path / "snap", "--quiet")
    assert r.returncode == 0, r.stderr[-2000:]
    r = _torchrun(2, port(), "plain", "--resume", tmp_path / "snap", "-o", tmp_path / "resumed", "-c", 3, "--quiet")
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "resumed").read_bytes() == (tmp_path / "one").read_bytes()


# ---- the reference's own end-to-end cases (test.rb:384-548), all-gate blueprints, plain mode ----
TEST_RB = [  # (blueprint, request, golden result, cycles)
    ("cahp-diamond-mux", "test00.in", "test00-diamond.out", 8), ("cahp-emerald-mux", "test00.in", "test00-emerald.out", 6),
    ("cahp-ruby-mux", "test09.in", "test09-ruby.out", 7), ("cahp-pearl-mux", "test09.in", "test09-pearl.out", 3),
    ("cahp-diamond-mux", "test01.in", "test01-diamond.out", 346), ("cahp-emerald-mux", "test01.in", "test01-emerald.out", 261),
    ("cahp-ruby-mux", "test10.in", "test10-ruby.out", 362), ("cahp-pearl-mux", "test10.in", "test10-pearl.out", 264),
    ("cahp-ruby-mux-1KiB", "test11.in", "test11.out", 7), ("const-4bit", "test22.in", "test22.out", 1),
    ("addr-4bit", "test04.in", "test04.out", 1), ("pass-addr-pass-4bit", "test04.in", "test04.out", 1),
    ("addr-register-4bit", "test16.in", "test16.out", 3), ("div-8bit", "test05.in", "test05.out", 1),
    ("mux-ram-addr8bit", "test06.in", "test06.out", 16), ("mux-ram-addr9bit", "test07.in", "test07.out", 16),
    ("mux-ram-8-16-16", "test08.in", "test08.out", 8), ("counter-4bit", "test13.in", "test13.out", 3),
    ("big-mult", "test21.in", "test21.out", 1),
    # blueprints declared with CMUX memories (type = "rom" / "ram"): evaluated as MUX memories, same golden packets
    ("cahp-diamond", "test00.in", "test00-diamond.out", 8), ("cahp-emerald", "test00.in", "test00-emerald.out", 6),
    ("cahp-ruby", "test09.in", "test09-ruby.out", 7), ("cahp-pearl", "test09.in", "test09-pearl.out", 3),
    ("cahp-diamond", "test01.in", "test01-diamond.out", 346), ("cahp-emerald", "test01.in", "test01-emerald.out", 261),
    ("cahp-ruby", "test10.in", "test10-ruby.out", 362), ("cahp-pearl", "test10.in", "test10-pearl.out", 264),
    ("cahp-ruby", "test14.in", "test14.out", 20), ("cahp-ruby-iyokanl1", "test09.in", "test09-ruby.out", 7),
    ("ram-addr8bit", "test06.in", "test06.out", 16), ("ram-addr9bit", "test07.in", "test07.out", 16),
    ("ram-8-16-16", "test08.in", "test08.out", 8), ("rom-7-32", "test12.in", "test12.out", 1), ("rom-4-8", "test15.in", "test15.out", 1),
    ("dff-reset", "test23.in", "test23.out", 1),
]


@pytest.mark.skipif(not (REF / "test" / "config-toml").exists(), reason="reference tree not mounted")
@pytest.mark.parametrize("bp,fin,fout,cycles", TEST_RB, ids=[f"{c[0]}-{c[1][4:6]}" for c in TEST_RB])
def test_reference_test_rb_cases_plain(tmp_path, bp, fin, fout, cycles):
    assert run_cli("plain", "--blueprint", REF / "test" / "config-toml" / f"{bp}.toml", "-i", REF / "test" / "in" / fin,
                   "-o", tmp_path / "res", "-c", cycles, "--quiet") == 0
    got = PlainPacket.load(tmp_path / "res")
    want = PlainPacket.from_toml((REF / "test" / "out" / fout).read_text())
    assert got.num_cycles == want.num_cycles == cycles
    assert sorted(got.bits) == sorted(want.bits)
    for name, bits in want.bits.items():
        assert np.array_equal(got.bits[name][:bits.size], bits), name
    for name, bits in want.ram.items():
        assert np.array_equal(got.ram[name][:bits.size], bits), name


def test_blueprint_loader_rejects_what_the_reference_rejects(tmp_path):
    """Error behaviour of the loader (NetworkBlueprint / YosysJSONReader die() cases, src/iyokan.hpp:1731-1895,2124)."""
    import json

    good = json.loads((FIX / "upcount2-netlist.json").read_text())

    def design(mutate):
        d = json.loads(json.dumps(good))
        mutate(d["modules"]["UpCount2"])
        (tmp_path / "d.json").write_text(json.dumps(d))
        (tmp_path / "d.toml").write_text('[[file]]\ntype = "yosys-json"\npath = "d.json"\nname = "core"\n[connect]\n'
                                         '"core/reset" = "@reset"\n"@out[0:1]" = "core/io_q[0:1]"\n')
        return tmp_path / "d.toml"

    with pytest.raises(ValueError, match="constant driver"):      # a cell input tied to a constant
        B.read_blueprint(design(lambda m: m["cells"]["inc0"]["connections"].update(A=["1"])))
    with pytest.raises(ValueError, match="unsupported cell type"):  # e.g. a latch or an $_SDFF_ cell
        B.read_blueprint(design(lambda m: m["cells"]["q0"].update(type="$_DLATCH_P_")))
    with pytest.raises(ValueError, match="no driver"):              # a signal nobody drives
        B.read_blueprint(design(lambda m: m["cells"]["inc1"]["connections"].update(B=[99])))
    (tmp_path / "both.toml").write_text('[connect]\n"@a" = "@b"\n')
    with pytest.raises(ValueError, match="Invalid connect"):
        B.read_blueprint(tmp_path / "both.toml")
    (tmp_path / "gnd.toml").write_text('[connect]\nTOGND = ["core/x[0:3]"]\n')
    with pytest.raises(ValueError, match="TOGND"):
        B.read_blueprint(tmp_path / "gnd.toml")
    (tmp_path / "nonet.toml").write_text('[connect]\n"@out" = "ghost/io_out"\n')
    with pytest.raises(ValueError, match="Invalid network name"):
        B.read_blueprint(tmp_path / "nonet.toml")
    with pytest.raises(ValueError, match="Invalid output port"):
        B.read_blueprint(design(lambda m: m["ports"].pop("io_q")))
    # and the command line turns every one of them into exit status 1 with a message, never a traceback
    (tmp_path / "req.in").write_text("cycles = 1\n")
    assert run_cli("plain", "--blueprint", tmp_path / "both.toml", "-i", tmp_path / "req.in", "-o", tmp_path / "r", "-c", 1) == 1
    assert run_cli("plain", "--blueprint", tmp_path / "nonet.toml", "-i", tmp_path / "req.in", "-o", tmp_path / "r", "-c", 1) == 1


def test_external_input_bit_fans_out_to_every_consumer(tmp_path):
    """One external input bit wired to several sub-networks drives ALL of them in this loader.  The reference keeps one
    connection per external bit in a map filled in the iteration order of an unordered TOML table
    (src/iyokan.hpp:1876-1884), so which consumer it drives is unspecified there; blueprints that rely on it are
    ambiguous in the reference and deterministic here."""
    import shutil

    shutil.copy(FIX / "upcount2-netlist.json", tmp_path / "n.json")
    (tmp_path / "two.toml").write_text(
        '[[file]]\nname = "a"\ntype = "yosys-json"\npath = "n.json"\n[[file]]\nname = "b"\ntype = "yosys-json"\npath = "n.json"\n'
        '[connect]\n"a/reset" = "@reset"\n"b/reset" = "@reset"\n"@outa[0:1]" = "a/io_q[0:1]"\n"@outb[0:1]" = "b/io_q[0:1]"\n')
    (tmp_path / "req.in").write_text("cycles = 3\n")
    assert run_cli("plain", "--blueprint", tmp_path / "two.toml", "-i", tmp_path / "req.in", "-o", tmp_path / "res", "-c", 3,
                   "--quiet") == 0
    res = PlainPacket.load(tmp_path / "res")
    assert word(res.bits["outa"]) == 2 and word(res.bits["outb"]) == 2
