"""Host netlist engine (levelised replay) on the plaintext back-end, against the reference's golden
result packets (test/out/*.out, converted by tests/golden/make_netlists.py) and structural facts from
SURVEY.md §8(d)."""
import json
from pathlib import Path

import numpy as np
import pytest

from iyokan_b200 import netlist as N

NL = Path(__file__).resolve().parent / "golden" / "netlists"
CASES = json.load(open(NL / "cases.json"))


def run_case(name):
    case = CASES[name]
    nl = N.Netlist.load(NL / f"{name}.npz")
    runner = N.PlainRunner(nl)
    req = case["request"]
    inputs = {p: N.bits_of(e["bytes"], e["size"]) for p, e in req["bits"].items()}
    rams = {p: N.bits_of(e["bytes"], e["size"]) for p, e in req["ram"].items()}
    roms = {p: N.bits_of(e["bytes"], e["size"]) for p, e in req["rom"].items()}
    out = runner.run(case["cycles"], inputs=inputs, rams=rams, roms=roms)
    return nl, runner, out, case["expected"]


@pytest.mark.parametrize("name", ["counter-4bit", "addr-4bit", "div-8bit", "mux-ram-8-16-16", "cahp-pearl-mux",
                                  "cahp-ruby-mux"])
def test_plain_run_matches_reference_golden(name):
    nl, runner, out, exp = run_case(name)
    for port, e in exp["bits"].items():
        assert N.bytes_of(out[port][:e["size"]]) == e["bytes"], port
    for mem, e in exp["ram"].items():
        assert N.bytes_of(runner.get_mem(mem)[:e["size"]]) == e["bytes"], mem


def test_structure_matches_survey():
    eng = N.NetEngine(N.Netlist.load(NL / "counter-4bit.npz"))
    assert eng.bootstraps_per_cycle == 11 and eng.num_dff == 4 and eng.num_levels == 4
    assert eng.level_widths == [3, 3, 3, 2]
    eng = N.NetEngine(N.Netlist.load(NL / "mux-ram-8-16-16.npz"))
    assert eng.bootstraps_per_cycle == 18985 and eng.num_dff == 4096                     # config 3
    eng = N.NetEngine(N.Netlist.load(NL / "cahp-pearl-mux.npz"))
    assert eng.nl.n == 32998                                                             # SURVEY §8(a) callers row
    assert abs(eng.bootstraps_per_cycle - 30775) < 50                                    # BASELINE.md §2 probe


def test_generators_and_validation():
    nl = N.ripple_adder(4)
    r = N.PlainRunner(nl)
    for a, b in [(3, 9), (15, 15), (0, 0), (7, 8)]:
        out = r.run(1, inputs={"a": N.bits_of([a], 4), "b": N.bits_of([b], 4)})
        assert N.bytes_of(out["sum"])[0] == a + b
    c = N.PlainRunner(N.counter(4))
    assert N.bytes_of(c.run(6)["out"])[0] == 5     # q after reset pass + 6 ticks: 0,1,..,5
    # MUX RAM generator: write then read back
    b = N.NetBuilder()
    addr = [b.input("addr", i) for i in range(3)]
    wren = b.input("wren", 0)
    wd = [b.input("wdata", i) for i in range(4)]
    for i, o in enumerate(N.mux_ram(b, addr, wren, wd)):
        b.output("rdata", i, o)
    rr = N.PlainRunner(b.build())
    out = rr.run(2, inputs={"addr": N.bits_of([5], 3), "wren": [1], "wdata": N.bits_of([0xA], 4)})
    assert N.bytes_of(out["rdata"])[0] == 0xA and N.bytes_of(rr.get_mem("ram"))[5 * 4 // 8] != 0
    # validation: combinational loop and bad arity are rejected
    with pytest.raises(N.B200FheError, match="loop"):
        N.NetEngine(N.Netlist(np.array([0, 0], np.uint8), np.array([1, 0], np.int32), np.array([1, 0], np.int32),
                              np.array([-1, -1], np.int32)))
    with pytest.raises(N.B200FheError, match="range"):
        N.NetEngine(N.Netlist(np.array([9], np.uint8), np.array([5], np.int32), np.array([-1], np.int32),
                              np.array([-1], np.int32)))


def test_net_abi_exports():
    import ctypes
    import re

    root = Path(__file__).resolve().parents[1]
    text = re.sub(r"/\*.*?\*/", "", (root / "include" / "b200net.h").read_text(), flags=re.S)
    names = sorted(set(re.findall(r"\b(b200net_[a-z0-9_]+)\s*\(", text)))
    h = N.load_net()
    assert sorted(N.NET_EXPORTS) == names
    for n in names:
        assert hasattr(h, n)


def test_cost_model_reproduces_the_measured_clock_cycles():
    """scripts/model_netlist.py (launch-plan model + sharding policy, no GPU) against the clocks measured on 1, 2, 4 and 8
    B200s and committed under profiles/: within 12 %, and the number of collectives per clock exactly."""
    import importlib.util
    import json
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("model_netlist", root / "scripts" / "model_netlist.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    checked = 0
    for case in ("cahp-pearl-mux", "cahp-ruby-mux", "mux-ram-8-16-16"):
        nl = N.Netlist.load(NL / f"{case}.npz")
        for n in (1, 2, 4, 8):
            log = root / "profiles" / f"r01_net_{case}_n{n}.log"
            if not log.exists():
                continue
            meas = json.loads(log.read_text().strip().splitlines()[-1])
            m = mod.model(nl, n)
            assert abs(m["ms_per_cycle"] / 1e3 - meas["s_per_cycle"]) / meas["s_per_cycle"] < 0.12, (case, n, m, meas)
            assert m["collectives"] == meas["collectives_per_cycle"], (case, n)
            assert m["bootstraps_per_cycle"] == meas["bootstraps_per_cycle"]
            checked += 1
    assert checked >= 6


@pytest.mark.parametrize("name", ["cahp-pearl-mux", "mux-ram-8-16-16", "counter-4bit"])
@pytest.mark.parametrize("world", [1, 2, 8])
@pytest.mark.parametrize("flags", [0, 1])
def test_static_schedule_is_a_valid_partition(name, world, flags):
    """b200net_schedule: every gate is evaluated exactly once (replicated gates once per rank, sharded gates by exactly
    one rank), only after everything it reads, sharded outputs sit in the exchanged slot range of their step at the
    owner's offset, and packing never makes the modelled clock slower than ASAP levels."""
    nl = N.Netlist.load(NL / f"{name}.npz")
    eng = N.NetEngine(nl)
    eng.schedule(world, 0)
    asap_ms = eng.schedule_info()["model_ms"]
    eng.schedule(world, flags)
    info = eng.schedule_info()
    assert info["model_ms"] <= asap_ms + 1e-9
    gates = np.nonzero(nl.kind < 15)[0]
    when = np.full(nl.n, -1)          # step that produces a node's value (0 for sources)
    when[(nl.kind == N.INPUT) | (nl.kind == N.DFF)] = 0
    owners = np.zeros(nl.n, np.int64)
    nsharded = 0
    for k in range(eng.num_steps):
        first, per = eng.step_exchange(k)
        per_rank = [eng.step_gates(k, r) for r in range(world)]
        common = set(per_rank[0].tolist())
        for g in per_rank[1:]:
            common &= set(g.tolist())
        for r, g in enumerate(per_rank):
            for node in g:
                owners[node] += 1
                assert when[node] in (-1, k + 1)
                # inputs come from earlier steps
                for arr in (nl.in0, nl.in1, nl.in2):
                    src = arr[node]
                    if src < 0:
                        continue
                    while nl.kind[src] == N.OUTPUT:
                        src = nl.in0[src]
                    assert 0 <= when[src] <= k, (name, k, node)
                if node not in common or world == 1:
                    if world > 1:
                        assert per > 0 and first + r * per <= eng.slot_of(node) < first + (r + 1) * per
                        nsharded += 1
        for g in per_rank:
            when[g] = k + 1
    assert np.all(when[gates] > 0)
    shared = owners[gates]
    assert set(np.unique(shared)) <= {1, world}       # sharded: one owner; replicated: all ranks
    assert info["collectives"] == sum(1 for k in range(eng.num_steps) if eng.step_exchange(k)[1] > 0)
    if world > 1 and name != "counter-4bit":
        assert nsharded > 0 and info["collectives"] > 0
    eng.close()
