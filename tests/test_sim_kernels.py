"""CPU lock-step simulation of the CUDA kernels (same source as the device code) vs the oracle."""
import ctypes

import numpy as np
import pytest

import oracle as O

BRJOB = np.dtype([("in", "<u4", 3), ("sgn", "i1", 3), ("pad", "i1"), ("off", "<u4")])
KSJOB = np.dtype([("u0", "<u4"), ("u1", "<u4"), ("out", "<u4"), ("post", "<u4")])


def p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def test_struct_layouts(sim):
    assert sim.sim_sizeof_brjob() == BRJOB.itemsize == 20
    assert sim.sim_sizeof_ksjob() == KSJOB.itemsize == 16


def test_warp_ntt_negacyclic_product(sim):
    P = sim.sim_prime()
    rng = np.random.default_rng(1)
    for trial in range(2):
        a = rng.integers(0, P, 1024, dtype=np.uint32)
        b = rng.integers(0, P, 1024, dtype=np.uint32)
        if trial == 1:  # extreme residues
            a[:] = P - 1
            b[:] = P - 1
        c = np.zeros(1024, np.uint32)
        sim.sim_negacyclic_mul_modp(p(a), p(b), p(c))
        A = a.astype(object)
        B = b.astype(object)
        full = np.convolve(A, B)
        ref = full[:1024].copy()
        ref[:1023] -= full[1024:]
        assert np.array_equal(np.array([int(v) % P for v in ref], dtype=np.uint64), c.astype(np.uint64))


def test_lazy_ranges_never_overflow(sim):
    # worst-case lazy inputs: forward takes digits+p (< p+64), inverse takes < 4p
    P = sim.sim_prime()
    fwd_in = np.full(1024, P + 31, np.uint32)
    inv_in = np.full(1024, 4 * P - 1, np.uint32)
    fmax, imax = ctypes.c_uint32(), ctypes.c_uint32()
    sim.sim_ntt_ranges(p(fwd_in), ctypes.byref(fmax), ctypes.byref(ctypes.c_uint32()))
    sim.sim_ntt_ranges(p(inv_in), ctypes.byref(ctypes.c_uint32()), ctypes.byref(imax))
    assert fmax.value < 4 * P
    assert imax.value < 4 * P


def test_bk_limbs_recombine(sim, keys, bk_ntt_sim):
    P = sim.sim_prime()
    assert bk_ntt_sim.max() < P
    # recombination identity on raw values: x0 + 2^11 x1 + 2^22 x2 == raw (mod 2^32)
    raw = keys.bk[3].astype(np.int64).ravel()
    v = raw.astype(np.uint32).view(np.int32).astype(np.int64)
    x0 = ((v & 2047) ^ 1024) - 1024
    v1 = (v - x0) >> 11
    x1 = ((v1 & 2047) ^ 1024) - 1024
    x2 = (v1 - x1) >> 11
    assert np.abs(x0).max() <= 1024 and np.abs(x1).max() <= 1024 and np.abs(x2).max() <= 512
    assert np.array_equal((x0 + (x1 << 11) + (x2 << 22)) % 2**32, raw % 2**32)


@pytest.mark.parametrize("G", [2, 4])
def test_blind_rotate_variant1_generic_bit_exact(sim, keys, bk_ntt_sim, G):
    # generic shape (brg_phases.h: loops over GL digits / LIMBS limbs, torus-width agnostic): the kernel of the
    # 80-bit flavour, pinned here at 128 bits against the same oracle as the specialised shapes
    rng = np.random.default_rng(10 + G)
    n = 3  # ragged last CTA
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    ca, cb = O.encrypt_bits(1, keys, pa), O.encrypt_bits(2, keys, pb)
    arena = np.zeros((2 * n, 640), np.uint16)
    arena[:n, :637], arena[n:, :637] = ca, cb
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, n + g, 0)
        jobs[g]["sgn"] = (-1, 1, 0)          # ANDNY: -a + b - mu
        jobs[g]["off"] = (-(1 << 13)) & 0xFFFF
    ubuf = np.zeros((n, 1028), np.uint32)
    sim.sim_blind_rotate1(G, p(jobs), n, p(arena), p(bk_ntt_sim), p(ubuf), 636)
    c = (-ca.astype(np.int32) + cb.astype(np.int32)).astype(np.uint16)
    c[:, 636] -= np.uint16(1 << 13)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


@pytest.mark.parametrize("G", [2, 4])
def test_blind_rotate_variant3_bit_exact(sim, keys, bk_ntt_sim, G):
    # interleaved transforms (ct_stage3 / gs_stage3): same result as the other two decompositions
    rng = np.random.default_rng(30 + G)
    n = 3
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    ca, cb = O.encrypt_bits(5, keys, pa), O.encrypt_bits(6, keys, pb)
    arena = np.zeros((2 * n, 640), np.uint16)
    arena[:n, :637], arena[n:, :637] = ca, cb
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, n + g, 0)
        jobs[g]["sgn"] = (1, 1, 0)           # OR: a + b + mu
        jobs[g]["off"] = 1 << 13
    ubuf = np.zeros((n, 1028), np.uint32)
    sim.sim_blind_rotate3(G, p(jobs), n, p(arena), p(bk_ntt_sim), p(ubuf), 636)
    c = (ca.astype(np.int32) + cb.astype(np.int32)).astype(np.uint16)
    c[:, 636] += np.uint16(1 << 13)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


@pytest.mark.parametrize("G", [2, 8, 84, 82])   # 84 / 82: eight jobs per CTA in barrier groups of 4 / 2
def test_blind_rotate_variant7_bit_exact(sim, keys, bk_ntt_sim, G):
    # 16-warp throughput shape (br7_phases.h): swizzled tiles, x1 + x2 interleave, one rotated difference kept in registers
    rng = np.random.default_rng(70 + G)
    n = 9 if G > 8 else 3  # ragged last CTA
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    ca, cb = O.encrypt_bits(17, keys, pa), O.encrypt_bits(18, keys, pb)
    arena = np.zeros((2 * n, 640), np.uint16)
    arena[:n, :637], arena[n:, :637] = ca, cb
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, n + g, 0)
        jobs[g]["sgn"] = (2, 2, 0)          # XOR: 2a + 2b + 2mu
        jobs[g]["off"] = 2 << 13
    ubuf = np.zeros((n, 1028), np.uint32)
    sim.sim_blind_rotate7(G, p(jobs), n, p(arena), p(bk_ntt_sim), p(ubuf), 636)
    c = (2 * ca.astype(np.int32) + 2 * cb.astype(np.int32)).astype(np.uint16)
    c[:, 636] += np.uint16(2 << 13)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


def test_team_ntt_equals_warp_ntt(sim):
    # the 64-thread x 16-point transform (ntt_block.h) reproduces the warp transform value for value,
    # lazy representation included, so the two kernels share one bootstrapping-key layout
    P = sim.sim_prime()
    rng = np.random.default_rng(7)
    for lo, hi in ((0, P + 64), (0, 4 * P)):
        x = rng.integers(lo, hi, 1024, dtype=np.uint32)
        a, b = np.zeros(1024, np.uint32), np.zeros(1024, np.uint32)
        sim.sim_warp_forward(p(x), p(a))
        sim.sim_block_forward(p(x), p(b))
        assert np.array_equal(a, b)
        sim.sim_warp_inverse(p(x), p(a))
        sim.sim_block_inverse(p(x), p(b))
        assert np.array_equal(a, b)
        # 128-thread x 8-point transform (ntt_block8.h)
        sim.sim_warp_forward(p(x), p(a))
        sim.sim_block8_forward(p(x), p(b))
        assert np.array_equal(a, b)
        sim.sim_warp_inverse(p(x), p(a))
        sim.sim_block8_inverse(p(x), p(b))
        assert np.array_equal(a, b)


def test_blind_rotate_variant4_bit_exact(sim, keys, bk_ntt_sim):
    # latency shape (br4_phases.h): one job per CTA, staged key, shared-memory accumulator
    rng = np.random.default_rng(44)
    n = 3
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    ca, cb = O.encrypt_bits(7, keys, pa), O.encrypt_bits(8, keys, pb)
    arena = np.zeros((2 * n, 640), np.uint16)
    arena[:n, :637], arena[n:, :637] = ca, cb
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, n + g, 0)
        jobs[g]["sgn"] = (-2, -2, 0)         # XNOR: -2a - 2b - 2mu
        jobs[g]["off"] = (-(2 << 13)) & 0xFFFF
    ubuf = np.zeros((n, 1028), np.uint32)
    sim.sim_blind_rotate4(p(jobs), n, p(arena), p(bk_ntt_sim), p(ubuf), 636)
    c = (-2 * ca.astype(np.int32) - 2 * cb.astype(np.int32)).astype(np.uint16)
    c[:, 636] -= np.uint16(2 << 13)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


def test_blind_rotate_variant6_bit_exact(sim, keys, bk_ntt_sim):
    # fine-grained cluster shape (br6_phases.h): 128-thread x 8-point teams
    rng = np.random.default_rng(66)
    n = 3
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    ca, cb = O.encrypt_bits(11, keys, pa), O.encrypt_bits(12, keys, pb)
    arena = np.zeros((2 * n, 640), np.uint16)
    arena[:n, :637], arena[n:, :637] = ca, cb
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, n + g, 0)
        jobs[g]["sgn"] = (1, -1, 0)          # ANDYN: a - b - mu
        jobs[g]["off"] = (-(1 << 13)) & 0xFFFF
    ubuf = np.zeros((n, 1028), np.uint32)
    sim.sim_blind_rotate6(p(jobs), n, p(arena), p(bk_ntt_sim), p(ubuf), 636)
    c = (ca.astype(np.int32) - cb.astype(np.int32)).astype(np.uint16)
    c[:, 636] -= np.uint16(1 << 13)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


def test_blind_rotate_variant9_bit_exact(sim, keys, bk_ntt_sim):
    # cluster shape with 4-point threads (br9_phases.h, ntt_block4.h): 256-thread teams, five two-stage passes
    rng = np.random.default_rng(99)
    n = 3
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    ca, cb = O.encrypt_bits(31, keys, pa), O.encrypt_bits(32, keys, pb)
    arena = np.zeros((2 * n, 640), np.uint16)
    arena[:n, :637], arena[n:, :637] = ca, cb
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, n + g, 0)
        jobs[g]["sgn"] = (-1, -1, 0)         # NOR: -a - b - mu
        jobs[g]["off"] = (-(1 << 13)) & 0xFFFF
    ubuf = np.zeros((n, 1028), np.uint32)
    sim.sim_blind_rotate9(p(jobs), n, p(arena), p(bk_ntt_sim), p(ubuf), 636)
    c = (-ca.astype(np.int32) - cb.astype(np.int32)).astype(np.uint16)
    c[:, 636] -= np.uint16(1 << 13)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


def test_blind_rotate_variant8_bit_exact(sim, keys, bk_ntt_sim):
    # quad-cluster shape (br8_phases.h): 4 CTAs per job, every transform cut into two 512-position halves
    rng = np.random.default_rng(88)
    n = 3
    pa, pb = rng.integers(0, 2, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)
    ca, cb = O.encrypt_bits(21, keys, pa), O.encrypt_bits(22, keys, pb)
    arena = np.zeros((2 * n, 640), np.uint16)
    arena[:n, :637], arena[n:, :637] = ca, cb
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, n + g, 0)
        jobs[g]["sgn"] = (-1, -1, 0)         # NOR: -a - b - mu
        jobs[g]["off"] = (-(1 << 13)) & 0xFFFF
    ubuf = np.zeros((n, 1028), np.uint32)
    sim.sim_blind_rotate8(p(jobs), n, p(arena), p(bk_ntt_sim), p(ubuf), 636)
    c = (-ca.astype(np.int32) - cb.astype(np.int32)).astype(np.uint16)
    c[:, 636] -= np.uint16(1 << 13)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


def test_phases_are_thread_order_independent(sim, keys, bk_ntt_sim):
    """The simulator's claim (and the kernels' barrier placement) rests on every phase function being free of
    intra-phase cross-thread communication.  If that holds, running the threads and warps of each phase in
    descending or permuted order must not change a single bit; a read-after-write or write-after-read hazard inside
    a phase (i.e. a missing barrier) would.  Checked for the six production shapes on 40 CMUX steps."""
    rng = np.random.default_rng(99)
    n, steps = 2, 40
    ca, cb = O.encrypt_bits(15, keys, rng.integers(0, 2, n, dtype=np.uint8)), O.encrypt_bits(16, keys, rng.integers(0, 2, n, dtype=np.uint8))
    arena = np.zeros((2 * n, 640), np.uint16)
    arena[:n, :637], arena[n:, :637] = ca, cb
    jobs = np.zeros(n, BRJOB)
    for g in range(n):
        jobs[g]["in"] = (g, n + g, 0)
        jobs[g]["sgn"] = (-1, -1, 0)
        jobs[g]["off"] = 1 << 13
    runs = {"br3": lambda u: sim.sim_blind_rotate3(2, p(jobs), n, p(arena), p(bk_ntt_sim), p(u), steps),
            "br7": lambda u: sim.sim_blind_rotate7(2, p(jobs), n, p(arena), p(bk_ntt_sim), p(u), steps),
            "brg": lambda u: sim.sim_blind_rotate1(2, p(jobs), n, p(arena), p(bk_ntt_sim), p(u), steps),
            "br4": lambda u: sim.sim_blind_rotate4(p(jobs), n, p(arena), p(bk_ntt_sim), p(u), steps),
            "br6": lambda u: sim.sim_blind_rotate6(p(jobs), n, p(arena), p(bk_ntt_sim), p(u), steps),
            "br8": lambda u: sim.sim_blind_rotate8(p(jobs), n, p(arena), p(bk_ntt_sim), p(u), steps)}
    try:
        ref = None
        for name, run in runs.items():
            outs = []
            for mode in (0, 1, 2):
                sim.sim_set_thread_order(mode)
                u = np.zeros((n, 1028), np.uint32)
                run(u)
                outs.append(u)
            assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2]), name
            ref = outs[0] if ref is None else ref
            assert np.array_equal(outs[0], ref), name   # and the shapes agree with each other after 40 steps
    finally:
        sim.sim_set_thread_order(0)


def test_blind_rotate_abar_edges(sim, keys, bk_ntt_sim):
    # a-bar = 2N (c_i = 0xFFFF), a-bar = 0, a-bar = N and b-bar = 2N / 1
    c = np.zeros((2, 637), np.uint16)
    c[0, :8] = [0xFFFF, 0xFFF0, 0, 15, 0x8000, 0x7FF0, 0x8010, 0x0010]
    c[0, 636] = 0x001F
    c[1, 600:636] = 0x8000
    c[1, 636] = 0xFFFF
    arena = np.zeros((2, 640), np.uint16)
    arena[:, :637] = c
    jobs = np.zeros(2, BRJOB)
    for g in range(2):
        jobs[g]["in"] = (g, 0, 0)
        jobs[g]["sgn"] = (1, 0, 0)
    ubuf = np.zeros((2, 1028), np.uint32)
    sim.sim_blind_rotate7(2, p(jobs), 2, p(arena), p(bk_ntt_sim), p(ubuf), 636)
    assert np.array_equal(ubuf[:, :1025], O.bootstrap_to_lvl1(keys, c))


def test_keyswitch_bit_exact(sim, keys, ksk_dev, golden):
    u = golden["ks_in"]
    n = u.shape[0]
    ubuf = np.zeros((n, 1028), np.uint32)
    ubuf[:, :1025] = u
    jobs = np.zeros(n, KSJOB)
    jobs["u0"], jobs["u1"], jobs["out"] = np.arange(n), 0xFFFFFFFF, np.arange(n)
    arena = np.full((n, 640), 0xABCD, np.uint16)
    sim.sim_keyswitch(p(jobs), n, p(ubuf), p(ksk_dev), p(arena))
    assert np.array_equal(arena[:, :637], golden["ks_out_tfhepp"])  # == the reference, bit for bit
    assert np.all(arena[:, 637:] == 0)
    # narrow-frontier path: four CTAs per switch + combine
    arena2 = np.full((n, 640), 0xABCD, np.uint16)
    sim.sim_keyswitch_split(p(jobs), n, p(ubuf), p(ksk_dev), p(arena2))
    assert np.array_equal(arena2, arena)
    # wide-frontier path: eight gates per CTA, one warp per gate (the last CTA is ragged unless n is a multiple of 8)
    arena3 = np.full((n, 640), 0xABCD, np.uint16)
    sim.sim_keyswitch8(p(jobs), n, p(ubuf), p(ksk_dev), p(arena3))
    assert np.array_equal(arena3, arena)
    # MUX tail: two summed lvl1 samples and + mu after the switch go through all paths alike
    mux = np.zeros(1, KSJOB)
    mux["u0"], mux["u1"], mux["out"], mux["post"] = 0, 1, 0, 1 << 13
    a1, a2, a3 = (np.zeros((1, 640), np.uint16) for _ in range(3))
    sim.sim_keyswitch(p(mux), 1, p(ubuf), p(ksk_dev), p(a1))
    sim.sim_keyswitch_split(p(mux), 1, p(ubuf), p(ksk_dev), p(a2))
    sim.sim_keyswitch8(p(mux), 1, p(ubuf), p(ksk_dev), p(a3))
    assert np.array_equal(a1, a2) and np.array_equal(a1, a3) and not np.array_equal(a1[0], arena[0])


def run_batch(sim, G, ops, arena, in0, in1, in2, out, bk_ntt_sim, ksk_dev):
    err = ctypes.c_char_p()
    ops = np.ascontiguousarray(ops, np.uint8)
    arrs = [None if a is None else np.ascontiguousarray(a, np.uint32) for a in (in0, in1, in2, out)]
    rc = sim.sim_gate_batch(G, p(ops), p(arrs[0]), p(arrs[1]), p(arrs[2]), p(arrs[3]), ctypes.c_size_t(ops.size),
                            p(arena), ctypes.c_size_t(arena.shape[0]), p(bk_ntt_sim), p(ksk_dev), ctypes.byref(err))
    return rc, err.value


def test_full_frontier_all_opcodes(sim, keys, bk_ntt_sim, ksk_dev):
    names = list(O.OPS)
    ops = np.array([O.OPS[n] for n in names], np.uint8)
    n = ops.size
    rng = np.random.default_rng(5)
    pa, pb, pc = (rng.integers(0, 2, n, dtype=np.uint8) for _ in range(3))
    ca, cb, cc = (O.encrypt_bits(s, keys, b) for s, b in ((21, pa), (22, pb), (23, pc)))
    arena = np.zeros((4 * n, 640), np.uint16)
    arena[:n, :637], arena[n:2 * n, :637], arena[2 * n:3 * n, :637] = ca, cb, cc
    ids = np.arange(4 * n, dtype=np.uint32)
    rc, err = run_batch(sim, 2, ops, arena, ids[:n], ids[n:2 * n], ids[2 * n:3 * n], ids[3 * n:], bk_ntt_sim, ksk_dev)
    assert rc == 0, err
    got = arena[3 * n:, :637]
    assert np.array_equal(got, O.gate_batch(keys, ops, ca, cb, cc))
    assert np.array_equal(O.decrypt_bits(keys, got.copy()), O.plain_gate_vec(ops, pa, pb, pc))
    assert np.all(arena[:, 637:] == 0)


def test_frontier_errors_and_dff_overlap(sim, keys, bk_ntt_sim, ksk_dev):
    arena = np.zeros((4, 640), np.uint16)
    ids = np.arange(4, dtype=np.uint32)
    rc, err = run_batch(sim, 2, [O.OPS["NAND"]], arena, ids[:1], None, None, ids[3:], bk_ntt_sim, ksk_dev)
    assert rc == -1 and b"input slot" in err
    rc, err = run_batch(sim, 2, [99], arena, ids[:1], ids[:1], None, ids[3:], bk_ntt_sim, ksk_dev)
    assert rc == -1 and b"opcode" in err
    rc, err = run_batch(sim, 2, [O.OPS["NOT"]], arena, ids[:1], None, None, np.array([7], np.uint32), bk_ntt_sim, ksk_dev)
    assert rc == -1 and b"output slot" in err
    # COPY chain with overlapping src/dst behaves like a simultaneous DFF tick
    arena[:, 0] = [10, 20, 30, 40]
    rc, err = run_batch(sim, 2, [O.OPS["COPY"]] * 3, arena, ids[:3], None, None, ids[1:], bk_ntt_sim, ksk_dev)
    assert rc == 0
    assert list(arena[:, 0]) == [10, 10, 20, 30]
