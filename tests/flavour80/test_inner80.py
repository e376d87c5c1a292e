"""80-bit parameter flavour (TFHEpp include/params/CGGI16.hpp: n = 500, 32-bit lvl0 torus, l = 2, Bg = 2^10, t = 8).

Runs in its own pytest process with B200FHE_FLAVOUR=80 (launched by tests/test_flavour80.py): oracle, simulator and
CUDA library are the -DORC_80BIT / -DB200FHE_80BIT builds of the same sources.  Pinning of the 80-bit oracle:
tests/golden/tfhepp_golden80.npz holds the outputs of the UNMODIFIED reference compiled with -DUSE_80BIT_SECURITY
(oracle/_ref/ref_driver80, generator tests/golden/make_golden.py).
"""
import ctypes
import hashlib

import numpy as np
import pytest

import oracle as O

assert O.FLAVOUR == "80", "run through tests/test_flavour80.py (B200FHE_FLAVOUR=80)"
BRJOB = np.dtype([("in", np.uint32, 3), ("sgn", np.int8, 3), ("pad", np.int8), ("off", np.uint32)])
KSJOB = np.dtype([("u0", np.uint32), ("u1", np.uint32), ("out", np.uint32), ("post", np.uint32)])


def p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def wrap32(d):
    return (d.astype(np.int64) + 2**31) % 2**32 - 2**31


# ---- the oracle against the reference's own 80-bit outputs ------------------------------------------------
def test_parameters():
    assert (O.N0, O.L, O.T, O.MU0, O.T0) == (500, 2, 8, 1 << 29, np.uint32)


def test_keygen_is_reproducible(golden, keys):
    assert hashlib.sha256(keys.bk.tobytes()).digest() == golden["bk_sha256"].tobytes()
    assert hashlib.sha256(keys.ksk.tobytes()).digest() == golden["ksk_sha256"].tobytes()


def test_decomposition_exact(golden):
    for poly, want in zip(golden["decompose_in"], golden["decompose_out"]):
        got = O.decompose(poly)
        assert np.array_equal(got, want) and got.min() >= -512 and got.max() <= 511


def test_cmux_step_within_fft_rounding(golden):
    # l = 2, Bg = 2^10: the reference's double-precision FFT is further from exact than at 128 bits (products up to 2^52)
    got = O.cmux_step(golden["cmux_acc"], golden["cmux_trgsw"], int(golden["cmux_abar"]))
    diff = np.abs(wrap32(got.astype(np.int64) - golden["cmux_out"].astype(np.int64)))
    assert diff.max() <= 64, diff.max()


def test_identity_keyswitch_exact(golden, keys):
    assert np.array_equal(O.keyswitch(keys, golden["ks_in"]), golden["ks_out_tfhepp"])


def test_blind_rotate_phase_matches_reference(golden, keys):
    mine = O.phase1(keys, O.bootstrap_to_lvl1(keys, golden["br_in"]))
    ref = O.phase1(keys, golden["br_out_tfhepp"])
    assert np.array_equal(np.sign(mine), np.sign(ref))
    for ph in (mine, ref):
        assert np.all(np.abs(np.abs(ph.astype(np.int64)) - O.MU1) < 2**26)


def test_every_gate_decrypts_like_the_reference(golden, keys):
    ops, pa, pb, pc = (golden[k] for k in ("gate_ops", "gate_pa", "gate_pb", "gate_pc"))
    ca, cb, cc = (O.encrypt_bits(int(sd), keys, bits) for sd, bits in zip(golden["gate_enc_seeds"], (pa, pb, pc)))
    mine = O.gate_batch(keys, ops, ca, cb, cc)
    want = O.plain_gate_vec(ops, pa, pb, pc)
    assert np.array_equal(O.decrypt_bits(keys, golden["gate_out_tfhepp"]), want)
    assert np.array_equal(O.decrypt_bits(keys, mine), want)
    free = np.isin(ops, [O.OPS[n] for n in ("NOT", "COPY", "CONST0", "CONST1")])
    assert np.array_equal(mine[free], golden["gate_out_tfhepp"][free])
    for c in (mine, golden["gate_out_tfhepp"]):
        ph = O.phase(keys, c).astype(np.int64)
        assert np.all(np.abs(np.abs(ph) - O.MU0) < O.MU0 // 2)


def test_mod_switch_wraps_in_the_lvl0_word(keys):
    # gatebootstrapping.hpp:58-65 with a 32-bit lvl0 torus: (a + 2^20) is formed in uint32, so a-bar never reaches 2N
    c = np.zeros(501, np.uint32)
    c[0], c[1], c[2], c[500] = 0xFFFFFFFF, 0xFFF00000, (1 << 20) - 1, (1 << 21) - 1
    abar, bbar = O.mod_switch(c)
    assert abar[0] == 0 and abar[1] == 0 and abar[2] == 0 and bbar == 2048
    c[500] = 0xFFFFFFFF
    assert O.mod_switch(c)[1] == 1


# ---- the kernels' phase functions in the lock-step simulator (libbr_sim80.so) --------------------------------
def test_sim_flavour(sim):
    assert sim.sim_flavour_bits() == 80 and sim.sim_n0() == 500


def test_bk_limbs_recombine(sim, keys, bk_ntt_sim):
    # five centred limbs of 7,7,6,6,6 bits recombine to the raw key coefficient modulo 2^32
    raw = keys.bk[3, 1, 0].astype(np.int64)
    v, limbs = raw.copy(), []
    for w in (7, 7, 6, 6):
        x = ((v & ((1 << w) - 1)) ^ (1 << (w - 1))) - (1 << (w - 1))
        limbs.append(x)
        v = (v - x) >> w
    limbs.append(np.where(v >= 2**5 + 1, v - 2**6, v))
    assert all(np.abs(x).max() <= 64 for x in limbs)
    acc = sum(x << s for x, s in zip(limbs, (0, 7, 14, 20, 26)))
    assert np.array_equal(acc % 2**32, raw % 2**32)
    assert bk_ntt_sim.shape == (500, 10, 4, 1024)


@pytest.mark.parametrize("G", [2, 4])
def test_sim_blind_rotate_bit_exact(sim, keys, bk_ntt_sim, golden, G):
    n = 3
    c = golden["br_in"][:n]
    arena = np.zeros((n, 512), np.uint32)
    arena[:, :501] = c
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, 0, 0)
        jobs[g]["sgn"] = (1, 0, 0)
    ubuf = np.zeros((n, 1028), np.uint32)
    sim.sim_blind_rotate1(G, p(jobs), n, p(arena), p(bk_ntt_sim), p(ubuf), 500)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


def test_sim_keyswitch_bit_exact(sim, keys, ksk_dev, golden):
    u = golden["ks_in"]
    n = u.shape[0]
    ubuf = np.zeros((n, 1028), np.uint32)
    ubuf[:, :1025] = u
    jobs = np.zeros(n, KSJOB)
    jobs["u0"], jobs["u1"], jobs["out"] = np.arange(n), 0xFFFFFFFF, np.arange(n)
    for fn in (sim.sim_keyswitch, sim.sim_keyswitch_split):
        arena = np.zeros((n, 512), np.uint32)
        fn(p(jobs), n, p(ubuf), p(ksk_dev), p(arena))
        assert np.array_equal(arena[:, :501], golden["ks_out_tfhepp"])     # == the reference's IdentityKeySwitch, bit for bit


def test_sim_full_frontier_all_opcodes(sim, keys, bk_ntt_sim, ksk_dev):
    names = ["NAND", "AND", "OR", "NOR", "XOR", "XNOR", "ANDNOT", "ORNOT", "ANDNY", "ORNY", "MUX", "NOT", "COPY", "CONST1", "CONST0"]
    ops = np.array([O.OPS[x] for x in names], np.uint8)
    n = ops.size
    rng = np.random.default_rng(80)
    pa, pb, pc = (rng.integers(0, 2, n, dtype=np.uint8) for _ in range(3))
    ca, cb, cc = (O.encrypt_bits(s, keys, b) for s, b in ((1, pa), (2, pb), (3, pc)))
    arena = np.zeros((4 * n, 512), np.uint32)
    arena[:n, :501], arena[n:2 * n, :501], arena[2 * n:3 * n, :501] = ca, cb, cc
    ids = np.arange(4 * n, dtype=np.uint32)
    err = ctypes.c_char_p()
    rc = sim.sim_gate_batch(-4, p(ops), p(ids[:n]), p(ids[n:2 * n]), p(ids[2 * n:3 * n]), p(ids[3 * n:]), ctypes.c_size_t(n),
                            p(arena), ctypes.c_size_t(4 * n), p(bk_ntt_sim), p(ksk_dev), ctypes.byref(err))
    assert rc == 0, err.value
    got = arena[3 * n:, :501]
    assert np.array_equal(got, O.gate_batch(keys, ops, ca, cb, cc))
    assert np.array_equal(O.decrypt_bits(keys, got), O.plain_gate_vec(ops, pa, pb, pc))


def test_library_exports_and_plan():
    # the 80-bit C-ABI library loads, exports every symbol and plans with its one (generic) shape; no GPU touched
    from iyokan_b200 import lib

    h = lib.load()
    assert lib.FLAVOUR == "80" and lib.LIB_PATH.name == "libb200fhe80.so"
    assert not [s for s in lib.EXPORTS if not hasattr(h, s)]
    assert lib.plan_rotation(1000) == [(1, 4, 1000)]
    from iyokan_b200 import netlist

    assert not [s for s in netlist.NET_EXPORTS if not hasattr(netlist.load_net(), s)]


# ---- GPU parity at 80 bits (through the C ABI of libb200fhe80.so) -----------------------------------------
@pytest.mark.gpu
def test_gpu_bk_ntt_matches_simulator(gpu_ctx, bk_ntt_sim):
    for first in (0, 250, 496):
        assert np.array_equal(gpu_ctx.test_read_bk_ntt(first, 4), bk_ntt_sim[first:first + 4])


@pytest.mark.gpu
def test_gpu_blind_rotate_and_keyswitch_bit_exact(gpu_ctx, keys, golden):
    c = golden["br_in"]
    u = gpu_ctx.test_bootstrap_lvl1(c)
    assert np.array_equal(u, O.bootstrap_to_lvl1(keys, c))
    edge = np.zeros((2, 501), np.uint32)
    edge[0, :6] = [0xFFFFFFFF, 0xFFF00000, 0, (1 << 20) - 1, 0x80000000, 0x7FF00000]
    edge[0, 500] = (1 << 21) - 1
    edge[1, 500] = 0xFFFFFFFF
    assert np.array_equal(gpu_ctx.test_bootstrap_lvl1(edge), O.bootstrap_to_lvl1(keys, edge))
    assert np.array_equal(gpu_ctx.test_keyswitch(golden["ks_in"]), golden["ks_out_tfhepp"])


@pytest.mark.gpu
def test_gpu_every_opcode_bit_exact(gpu_ctx, keys):
    names = ["NAND", "AND", "OR", "NOR", "XOR", "XNOR", "ANDNOT", "ORNOT", "ANDNY", "ORNY", "MUX", "NOT", "COPY", "CONST1", "CONST0"]
    ops = np.repeat(np.array([O.OPS[x] for x in names], np.uint8), 3)
    n = ops.size
    rng = np.random.default_rng(81)
    pa, pb, pc = (rng.integers(0, 2, n, dtype=np.uint8) for _ in range(3))
    ca, cb, cc = (O.encrypt_bits(s, keys, b) for s, b in ((4, pa), (5, pb), (6, pc)))
    gpu_ctx.arena_alloc(4 * n)
    got = gpu_ctx.gates_host(ops, ca, cb, cc)
    assert np.array_equal(got, O.gate_batch(keys, ops, ca, cb, cc))
    assert np.array_equal(O.decrypt_bits(keys, got), O.plain_gate_vec(ops, pa, pb, pc))


@pytest.mark.gpu
def test_gpu_wide_frontier_and_netlist(gpu_ctx, keys):
    # 1500 NAND + MUX gates (several waves of the generic shape), decrypted bits; then a 4-bit adder through b200net80
    from iyokan_b200 import netlist as N

    n = 1500
    rng = np.random.default_rng(82)
    ops = np.where(np.arange(n) % 3 == 0, O.OPS["MUX"], O.OPS["NAND"]).astype(np.uint8)
    pa, pb, pc = (rng.integers(0, 2, n, dtype=np.uint8) for _ in range(3))
    ca, cb, cc = (O.encrypt_bits(s, keys, b) for s, b in ((7, pa), (8, pb), (9, pc)))
    gpu_ctx.arena_alloc(4 * n)
    got = gpu_ctx.gates_host(ops, ca, cb, cc)
    assert np.array_equal(O.decrypt_bits(keys, got), O.plain_gate_vec(ops, pa, pb, pc))
    assert np.array_equal(got[:6], O.gate_batch(keys, ops[:6], ca[:6], cb[:6], cc[:6]))
    nl = N.ripple_adder(4)
    r = N.EncryptedRunner(nl, gpu_ctx, lambda bits: O.encrypt_bits(5, keys, bits))
    out = r.run(1, inputs={"a": N.bits_of([11], 4), "b": N.bits_of([6], 4)})
    assert N.bytes_of(O.decrypt_bits(keys, out["sum"]))[0] == 17
