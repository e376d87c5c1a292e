"""The exact-integer oracle against vectors produced by the UNMODIFIED reference (TFHEpp).

Fixtures: tests/golden/tfhepp_golden.npz, generated in the build container by
tests/golden/make_golden.py through oracle/_ref/ref_driver.
"""
import hashlib

import numpy as np

import oracle as O


def wrap32(d):
    return (d.astype(np.int64) + 2**31) % 2**32 - 2**31


def test_keygen_is_reproducible(golden, keys):
    # the fixtures only make sense if the deterministic key generator reproduces the same keys here
    assert hashlib.sha256(keys.bk.tobytes()).digest() == golden["bk_sha256"].tobytes()
    assert hashlib.sha256(keys.ksk.tobytes()).digest() == golden["ksk_sha256"].tobytes()


def test_decomposition_exact(golden):
    # TFHEpp Decomposition<lvl1param>, trgsw.hpp:62-78
    for p, want in zip(golden["decompose_in"], golden["decompose_out"]):
        got = O.decompose(p)
        assert np.array_equal(got, want)
        assert got.min() >= -32 and got.max() <= 31


def test_mul_by_xai_exact(golden):
    # utils.hpp:113-144, including a = 0, N, 2N-1 and the a = 2N identity
    for p, a, want in zip(golden["mulxai_in"], golden["mulxai_a"], golden["mulxai_out"]):
        assert np.array_equal(O.mul_xai(p, int(a), False), want[0])
        assert np.array_equal(O.mul_xai(p, int(a), True), want[1])


def test_cmux_step_within_fft_rounding(golden):
    # detwfa.hpp:36-49: the reference's double-precision FFT equals the exact product up to rounding
    got = O.cmux_step(golden["cmux_acc"], golden["cmux_trgsw"], int(golden["cmux_abar"]))
    diff = np.abs(wrap32(got.astype(np.int64) - golden["cmux_out"].astype(np.int64)))
    assert diff.max() <= 4, diff.max()


def test_identity_keyswitch_exact(golden, keys):
    # keyswitch.hpp:11-52 is pure integer arithmetic in the reference: bit-exact
    assert np.array_equal(O.keyswitch(keys, golden["ks_in"]), golden["ks_out_tfhepp"])


def test_blind_rotate_phase_matches_reference(golden, keys):
    # After 636 CMUX steps the reference's FFT rounding changes individual digits, so ciphertexts
    # differ while both stay valid encryptions of the same message: compare decrypted phases.
    mine = O.phase1(keys, O.bootstrap_to_lvl1(keys, golden["br_in"]))
    ref = O.phase1(keys, golden["br_out_tfhepp"])
    assert np.array_equal(np.sign(mine), np.sign(ref))
    for ph in (mine, ref):
        assert np.all(np.abs(np.abs(ph.astype(np.int64)) - O.MU1) < 2**26)  # blind-rotation noise << mu/8


def test_every_gate_decrypts_like_the_reference(golden, keys):
    ops, pa, pb, pc = (golden[k] for k in ("gate_ops", "gate_pa", "gate_pb", "gate_pc"))
    s = golden["gate_enc_seeds"]
    ca, cb, cc = (O.encrypt_bits(int(sd), keys, p) for sd, p in zip(s, (pa, pb, pc)))
    mine = O.gate_batch(keys, ops, ca, cb, cc)
    want_bits = O.plain_gate_vec(ops, pa, pb, pc)
    assert np.array_equal(O.decrypt_bits(keys, golden["gate_out_tfhepp"]), want_bits)
    assert np.array_equal(O.decrypt_bits(keys, mine), want_bits)
    # bootstrap-free gates are integer-only in the reference too: identical ciphertexts
    free = np.isin(ops, [O.OPS[n] for n in ("NOT", "COPY", "CONST0", "CONST1")])
    assert np.array_equal(mine[free], golden["gate_out_tfhepp"][free])
    # noise margin: |phase| stays within mu/2 of +-mu for both
    for c in (mine, golden["gate_out_tfhepp"]):
        ph = O.phase(keys, c).astype(np.int32)
        assert np.all(np.abs(np.abs(ph) - O.MU0) < O.MU0 // 2)


def test_mod_switch_edges(keys):
    # gatebootstrapping.hpp:26-30,58-65: a-bar may reach 2N for the uint16 torus; b-bar is unrounded
    c = np.zeros(637, np.uint16)
    c[0], c[1], c[2], c[636] = 0xFFFF, 0xFFF0, 15, 0x001F
    abar, bbar = O.mod_switch(c)
    assert abar[0] == 2048 and abar[1] == 2048 and abar[2] == 0 and abar[3] == 0
    assert bbar == 2048
    c[636] = 0xFFFF
    assert O.mod_switch(c)[1] == 1


def test_trivial_inputs_skip_path(keys):
    # Iyokan's reset/const wires are trivial ciphertexts (a = 0): every a-bar is 0, the reference
    # skips all CMUXes (gatebootstrapping.hpp:66) and the exact model must agree bit for bit.
    t1 = np.zeros((1, 637), np.uint16)
    t1[0, 636] = O.MU0
    t0 = np.zeros((1, 637), np.uint16)
    t0[0, 636] = (-O.MU0) & 0xFFFF
    out = O.gate_batch(keys, [O.OPS["NAND"], O.OPS["NAND"]], np.vstack([t1, t1]), np.vstack([t1, t0]))
    assert list(O.decrypt_bits(keys, out)) == [0, 1]
