"""The reference-side binding, for real: a C++ program written against the UNMODIFIED reference headers (TFHEpp key,
ciphertext, encryption and decryption types - what Iyokan's gate tasks use) evaluates every gate type through the
C ABI of libb200fhe.so and compares with TFHEpp::Hom* on the same ciphertexts (tests/ref_link/b200_gate_test.cpp,
built by `make -C oracle reflink` where the reference tree exists; the binary travels under oracle/_ref/)."""
import subprocess

import pytest

import oracle as O


def _run(*args):
    return subprocess.run([str(O.REF_LINK_TEST), *map(str, args)], capture_output=True, text=True, timeout=900)


@pytest.mark.gpu
@pytest.mark.skipif(not O.REF_LINK_TEST.exists(), reason="oracle/_ref/b200_gate_test not built")
def test_tfhepp_types_through_the_c_abi():
    r = _run(96, 3)   # 96 gates per type (cluster + one-job-per-SM kernels), 3 of each also through TFHEpp itself
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "PASS"
    assert len(lines) == 11 and all("wrong bits 0" in ln and "differ from TFHEpp 0/3" in ln for ln in lines[:-1])


@pytest.mark.skipif(not O.REF_LINK_TEST.exists(), reason="oracle/_ref/b200_gate_test not built")
def test_binding_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run(4, 1)
    assert r.returncode == 2 and "b200fhe_create" in r.stderr   # no silent CPU path behind the C ABI


@pytest.mark.gpu
@pytest.mark.skipif(not (O.IYOKAN_B200.exists() and O.have_iyokan_packet()), reason="oracle/_ref/iyokan-b200 not built")
@pytest.mark.parametrize("blueprint,cycles", [("lookup.toml", 4), ("lookup-cmux.toml", 4), ("upcount2.toml", 3)])
def test_iyokan_b200_binary_with_the_reference_loader(tmp_path, blueprint, cycles):
    """`iyokan-b200 tfhe`: the reference's own blueprint / Yosys / MUX-memory loader and packet code (compiled from
    its sources) in front of the B200 engine, against keys and packets made by the reference's iyokan-packet; the
    decrypted result equals the plaintext run of the same blueprint and request."""
    from pathlib import Path

    import numpy as np

    from iyokan_b200.cli import main
    from iyokan_b200.packet import PlainPacket

    bp = Path(__file__).resolve().parent / "fixtures" / blueprint
    sk, ek = tmp_path / "sk", tmp_path / "ek"
    O.iyokan_packet("genkey", "--type", "tfhepp", "--out", sk)
    O.iyokan_packet("genevalkey", "--in", sk, "--out", ek)
    if blueprint.startswith("lookup"):
        req = PlainPacket(rom={"rom": np.array([(w >> b) & 1 for w in (3, 12, 5, 10) for b in range(4)], np.uint8)},
                          ram={"ram": np.zeros(16, np.uint8)},
                          bits={"raddr": np.array([1, 0, 0, 1, 1, 1], np.uint8), "waddr": np.array([0, 0, 1, 0, 0, 1], np.uint8),
                                "wren": np.array([1, 1, 0], np.uint8)})
    else:
        req = PlainPacket()
    (tmp_path / "req.toml").write_text(req.to_toml())
    O.iyokan_packet("toml2packet", "--in", tmp_path / "req.toml", "--out", tmp_path / "req")
    O.iyokan_packet("enc", "--key", sk, "--in", tmp_path / "req", "--out", tmp_path / "req.enc")
    r = subprocess.run([str(O.IYOKAN_B200), "tfhe", "--blueprint", str(bp), "--evalkey", str(ek), "-i", str(tmp_path / "req.enc"),
                        "-o", str(tmp_path / "res.enc"), "-c", str(cycles)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bootstraps per cycle" in r.stdout + r.stderr
    O.iyokan_packet("dec", "--key", sk, "--in", tmp_path / "res.enc", "--out", tmp_path / "res")
    try:
        main(["plain", "--blueprint", str(bp), "-i", str(tmp_path / "req"), "-o", str(tmp_path / "res.plain"), "-c", str(cycles),
              "--quiet"])
    except SystemExit as e:
        assert e.code == 0
    got, want = PlainPacket.load(tmp_path / "res"), PlainPacket.load(tmp_path / "res.plain")
    assert got.num_cycles == want.num_cycles == cycles
    assert sorted(got.bits) == sorted(want.bits)
    for name in want.bits:
        assert np.array_equal(got.bits[name], want.bits[name]), name
    for name in want.ram:
        assert np.array_equal(got.ram[name], want.ram[name]), name
    ek.unlink()


@pytest.mark.gpu
@pytest.mark.skipif(not O.B200_TEST0.exists(), reason="oracle/_ref/b200_test0 not built")
def test_reference_test0_suite_on_the_b200_plugin(tmp_path):
    """The reference's own templated unit tests (src/test0.cpp:43-455) instantiated with B200NetworkBuilder: the plugin
    iyokan_b200/host/iyokan_b200.hpp compiled against the unmodified src/iyokan.hpp (Task / DepNode / ReadyQueue / Worker /
    NetworkBuilder / IyokanL1JSONReader are the reference's), gates evaluated through the C ABI, one batch per frontier."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    import make_ref_assets

    make_ref_assets.materialise(tmp_path / "test")   # the tests open test/iyokanl1-json/*.json relative to the cwd
    r = subprocess.run([str(O.B200_TEST0)], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for t in ("testNOT", "testMUX", "testBinopGates", "testFromJSONtest_pass_4bit", "testFromJSONtest_and_4bit",
              "testFromJSONtest_and_4_2bit", "testFromJSONtest_mux_4bit", "testFromJSONtest_addr_4bit",
              "testFromJSONtest_register_4bit", "testSequentialCircuit", "testFromJSONtest_counter_4bit",
              "testPrioritySetVisitor"):
        assert f"ok  {t}<B200NetworkBuilder>" in r.stdout, t
    assert "test0 suite green on the B200 back-end" in r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(not (O.IYOKAN_B200.exists() and O.have_iyokan_packet()), reason="oracle/_ref/iyokan-b200 not built")
def test_iyokan_b200_snapshot_resume_and_dumps(tmp_path):
    """`--snapshot` / `--resume` (2 + 2 cycles == 4 cycles, ciphertext for ciphertext), `--dump-prefix --secret-key`
    (decrypted packet before every cycle) and the ProgressGraphMaker-format dumps (`--dump-time-csv-prefix`,
    `--dump-graph-json-prefix`, `--dump-graph-dot-prefix`; src/iyokan_tfhepp.cpp:298-305,538-555, src/main.cpp:116-122)."""
    import csv
    import json
    from pathlib import Path

    import numpy as np

    from iyokan_b200.packet import PlainPacket, TFHEPacket

    bp = Path(__file__).resolve().parent / "fixtures" / "lookup.toml"
    sk, ek = tmp_path / "sk", tmp_path / "ek"
    O.iyokan_packet("genkey", "--type", "tfhepp", "--out", sk)
    O.iyokan_packet("genevalkey", "--in", sk, "--out", ek)
    req = PlainPacket(rom={"rom": np.array([(w >> b) & 1 for w in (3, 12, 5, 10) for b in range(4)], np.uint8)},
                      ram={"ram": np.zeros(16, np.uint8)},
                      bits={"raddr": np.array([1, 0, 0, 1, 1, 1], np.uint8), "waddr": np.array([0, 0, 1, 0, 0, 1], np.uint8),
                            "wren": np.array([1, 1, 0], np.uint8)})
    (tmp_path / "req.toml").write_text(req.to_toml())
    O.iyokan_packet("toml2packet", "--in", tmp_path / "req.toml", "--out", tmp_path / "req")
    O.iyokan_packet("enc", "--key", sk, "--in", tmp_path / "req", "--out", tmp_path / "req.enc")

    def run(*extra):
        r = subprocess.run([str(O.IYOKAN_B200), "tfhe", "--evalkey", str(ek), *map(str, extra)], capture_output=True, text=True,
                           timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr

    run("--blueprint", bp, "-i", tmp_path / "req.enc", "-o", tmp_path / "four.enc", "-c", 4, "--dump-prefix", tmp_path / "dump",
        "--secret-key", sk, "--dump-time-csv-prefix", tmp_path / "time", "--dump-graph-json-prefix", tmp_path / "graph",
        "--dump-graph-dot-prefix", tmp_path / "dot")
    run("--blueprint", bp, "-i", tmp_path / "req.enc", "-o", tmp_path / "two.enc", "-c", 2, "--snapshot", tmp_path / "snap")
    run("--resume", tmp_path / "snap", "-o", tmp_path / "two_two.enc", "-c", 2)
    a, b = TFHEPacket.load(tmp_path / "four.enc"), TFHEPacket.load(tmp_path / "two_two.enc")
    assert a.num_cycles == b.num_cycles == 4
    for name in a.bits:   # gate evaluation is deterministic: the resumed run reproduces the ciphertexts, not just the bits
        assert np.array_equal(a.bits[name], b.bits[name]), name
    assert np.array_equal(a.ram_in_tlwe["ram"], b.ram_in_tlwe["ram"])
    # decrypted dump before cycle 3 == a plain 3-cycle run's result
    O.iyokan_packet("dec", "--key", sk, "--in", tmp_path / "four.enc", "--out", tmp_path / "four")
    d3 = PlainPacket.load(tmp_path / "dump-3")
    from iyokan_b200.cli import main
    try:
        main(["plain", "--blueprint", str(bp), "-i", str(tmp_path / "req"), "-o", str(tmp_path / "three.plain"), "-c", "3", "--quiet"])
    except SystemExit as e:
        assert e.code == 0
    want3 = PlainPacket.load(tmp_path / "three.plain")
    for name in want3.bits:
        assert np.array_equal(d3.bits[name], want3.bits[name]), name
    # ProgressGraphMaker formats: "start","end","index","id","kind","desc" per node; nodes / edges; a digraph
    rows = list(csv.reader(open(tmp_path / "time-2.csv")))
    assert len(rows[0]) == 6 and {r[4] for r in rows} >= {"MUX", "DFF", "INPUT"}
    assert all(r[0] <= r[1] for r in rows)
    g = json.load(open(tmp_path / "graph-2.json"))
    assert len(g["nodes"]) == len(rows) and len(g["edges"]) > len(rows) // 2
    assert all(0 <= e["from"] < len(rows) and 0 <= e["to"] < len(rows) for e in g["edges"])
    dot = open(tmp_path / "dot-2.dot").read()
    assert dot.startswith("digraph progress_graph_maker {") and dot.count("->") == len(g["edges"])
