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
