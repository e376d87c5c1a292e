"""Property tests (hypothesis) for the host logic around the hot path: packet codecs, the levelised engine against a
direct evaluator on random netlists, and the kernel simulator against the oracle on random gate batches."""
import ctypes

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle as O
from iyokan_b200 import netlist as N
from iyokan_b200.lib import OPS
from iyokan_b200.packet import PlainPacket, TFHEPacket

names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz_0123456789", min_size=1, max_size=12)
bit_arrays = st.lists(st.integers(0, 1), min_size=0, max_size=40).map(lambda v: np.array(v, np.uint8))


@given(ram=st.dictionaries(names, bit_arrays, max_size=3), rom=st.dictionaries(names, bit_arrays, max_size=3),
       bits=st.dictionaries(names, bit_arrays, max_size=4), cycles=st.one_of(st.none(), st.integers(-1, 10**6)))
@settings(max_examples=60, deadline=None)
def test_plain_packet_binary_and_toml_round_trip(ram, rom, bits, cycles):
    p = PlainPacket(ram=ram, rom=rom, bits=bits, num_cycles=cycles)
    q = PlainPacket.loads(p.dumps())
    assert q.num_cycles == cycles and q.dumps() == p.dumps()
    for a, b in ((p.ram, q.ram), (p.rom, q.rom), (p.bits, q.bits)):
        assert sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a)
    t = PlainPacket.from_toml(p.to_toml())
    assert t.num_cycles == cycles
    for a, b in ((p.ram, t.ram), (p.rom, t.rom), (p.bits, t.bits)):
        assert sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a)


@given(seed=st.integers(0, 2**32 - 1), n=st.integers(0, 3), cycles=st.one_of(st.none(), st.integers(0, 99)))
@settings(max_examples=25, deadline=None)
def test_tfhe_packet_round_trip(seed, n, cycles):
    rng = np.random.default_rng(seed)
    mk = lambda k: {f"p{i}": rng.integers(0, 65536, (int(rng.integers(0, 5)), 637), dtype=np.uint16) for i in range(k)}  # noqa: E731
    p = TFHEPacket(ram_in_tlwe=mk(n), rom_in_tlwe=mk(n), bits=mk(n + 1), num_cycles=cycles)
    data = p.dumps()
    q = TFHEPacket.loads(data)
    assert q.dumps() == data and q.num_cycles == cycles
    assert all(np.array_equal(p.bits[k], q.bits[k]) for k in p.bits)


def _random_netlist(rng, n_in, n_dff, n_gate):
    """Random sequential circuit: gates read earlier gates, inputs or DFFs; DFFs read any gate."""
    b = N.NetBuilder()
    ins = [b.input("x", i) for i in range(n_in)]
    dffs = [b.dff() for _ in range(n_dff)]
    pool = ins + dffs
    two = ["AND", "NAND", "ANDNOT", "OR", "NOR", "ORNOT", "XOR", "XNOR", "ANDNY", "ORNY"]
    for _ in range(n_gate):
        k = rng.integers(0, 14)
        pick = lambda: int(pool[rng.integers(0, len(pool))])  # noqa: E731
        if k < 10:
            g = b.gate(two[k], pick(), pick())
        elif k == 10:
            g = b.gate("MUX", pick(), pick(), pick())
        elif k == 11:
            g = b.gate("NOT", pick())
        else:
            g = b.gate("CONST1" if k == 12 else "CONST0")
        pool.append(g)
    for d in dffs:
        b.set_dff_input(d, int(pool[rng.integers(0, len(pool))]))
    for i in range(min(4, len(pool))):
        b.output("y", i, int(pool[-1 - i]))
    return b.build()


def _direct_eval(nl, v):
    """Reference semantics, node by node in creation order (gates only read earlier nodes or DFFs)."""
    T = {OPS[k]: k for k in OPS}
    for i in range(nl.n):
        k = int(nl.kind[i])
        if k >= 32:
            if k == N.OUTPUT:
                v[i] = v[nl.in0[i]]
            continue
        a = v[nl.in0[i]] if nl.in0[i] >= 0 else 0
        c = v[nl.in1[i]] if nl.in1[i] >= 0 else 0
        s = v[nl.in2[i]] if nl.in2[i] >= 0 else 0
        v[i] = {"AND": a & c, "NAND": 1 - (a & c), "ANDNOT": a & (1 - c), "OR": a | c, "NOR": 1 - (a | c),
                "ORNOT": a | (1 - c), "XOR": a ^ c, "XNOR": 1 - (a ^ c), "MUX": c if s else a, "NOT": 1 - a, "COPY": a,
                "CONST0": 0, "CONST1": 1, "ANDNY": (1 - a) & c, "ORNY": (1 - a) | c}[T[k]]


@given(seed=st.integers(0, 2**32 - 1), n_in=st.integers(1, 5), n_dff=st.integers(0, 6), n_gate=st.integers(1, 60))
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
def test_levelised_engine_equals_direct_evaluation(seed, n_in, n_dff, n_gate):
    rng = np.random.default_rng(seed)
    nl = _random_netlist(rng, n_in, n_dff, n_gate)
    eng = N.NetEngine(nl)
    assert sum(eng.level_widths) == int(np.sum(nl.kind < 15))
    assert eng.bootstraps_per_cycle == sum(eng.level_bootstraps)
    v = np.zeros(nl.n, np.uint8)
    w = v.copy()
    dffs = np.nonzero(nl.kind == N.DFF)[0]
    for _ in range(4):
        x = rng.integers(0, 2, n_in, dtype=np.uint8)
        v[nl.in_ports["x"]] = x
        w[nl.in_ports["x"]] = x
        eng.plain_eval(v)
        _direct_eval(nl, w)
        assert np.array_equal(v, w)
        eng.plain_tick(v)
        w[dffs] = w[nl.in0[dffs]].copy()
        assert np.array_equal(v, w)
    eng.close()


@given(seed=st.integers(0, 2**32 - 1))
@settings(max_examples=3, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
def test_simulated_kernels_equal_oracle_on_random_frontiers(sim, keys, bk_ntt_sim, ksk_dev, seed):
    """Random opcodes / operands through the CPU lock-step model of the CUDA kernels == exact-integer oracle."""
    rng = np.random.default_rng(seed)
    n = 3
    ops = rng.integers(0, 15, n).astype(np.uint8)
    cts = [O.encrypt_bits(int(rng.integers(1, 1000)), keys, rng.integers(0, 2, n, dtype=np.uint8)) for _ in range(3)]
    arena = np.zeros((4 * n, 640), np.uint16)
    for k in range(3):
        arena[k * n:(k + 1) * n, :637] = cts[k]
    ids = np.arange(4 * n, dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    err = ctypes.c_char_p()
    rc = sim.sim_gate_batch(4, p(ops), p(ids[:n]), p(ids[n:2 * n]), p(ids[2 * n:3 * n]), p(ids[3 * n:]), ctypes.c_size_t(n),
                            p(arena), ctypes.c_size_t(arena.shape[0]), p(bk_ntt_sim), p(ksk_dev), ctypes.byref(err))
    assert rc == 0, err.value
    assert np.array_equal(arena[3 * n:, :637], O.gate_batch(keys, ops, *cts))
