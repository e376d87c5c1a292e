import ctypes
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden" / "tfhepp_golden.npz"
TEST_KEY_SEED = 424242  # same keys as the golden fixtures, so one key set serves every test


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def keys():
    import oracle as O

    return O.cached_keys(TEST_KEY_SEED)


@pytest.fixture(scope="session")
def sim():
    """ctypes handle of the lock-step CPU simulator of the CUDA kernels."""
    from iyokan_b200 import build as B

    lib = ctypes.CDLL(str(B.build_sim()))
    lib.sim_prime.restype = ctypes.c_uint32
    lib.sim_gate_batch.restype = ctypes.c_int
    return lib


@pytest.fixture(scope="session")
def bk_ntt_sim(sim, keys):
    out = np.zeros((636, 6, 6, 1024), np.uint32)
    sim.sim_bk_prepare(keys.bk.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), 636)
    return out


@pytest.fixture(scope="session")
def ksk_dev(keys):
    out = np.zeros((1024, 7, 3, 640), np.uint16)
    out[..., :637] = keys.ksk
    return out


@pytest.fixture(scope="session")
def gpu_ctx(keys):
    """One CUDA context with keys loaded for the whole GPU session (fails loudly without the .so)."""
    from iyokan_b200 import Context

    ctx = Context(0)
    ctx.load_keys(keys.bk, keys.ksk)
    yield ctx
    ctx.close()
