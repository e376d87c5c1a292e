import ctypes
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import os

# Parameter flavour of this pytest process ("" = 128-bit, "80" = 80-bit; read once at import by oracle and
# iyokan_b200.lib, like the reference's compile-time switch).  The 80-bit tests live in tests/flavour80/ and run in their
# own process (tests/test_flavour80.py launches it with B200FHE_FLAVOUR=80); every other module is 128-bit only.
FLAVOUR = os.environ.get("B200FHE_FLAVOUR", "")
collect_ignore = ["flavour80"] if FLAVOUR == "" else [p.name for p in Path(__file__).parent.glob("test_*.py")]
GOLDEN = ROOT / "tests" / "golden" / f"tfhepp_golden{FLAVOUR}.npz"
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

    lib = ctypes.CDLL(str(B.build_sim(flavour=FLAVOUR)))
    lib.sim_prime.restype = ctypes.c_uint32
    lib.sim_gate_batch.restype = ctypes.c_int
    return lib


@pytest.fixture(scope="session")
def bk_ntt_sim(sim, keys):
    import oracle as O

    limbs = 5 if FLAVOUR == "80" else 3
    out = np.zeros((O.N0, 2 * limbs, O.ROWS, 1024), np.uint32)
    sim.sim_bk_prepare(keys.bk.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), O.N0)
    return out


@pytest.fixture(scope="session")
def ksk_dev(keys):
    import oracle as O

    out = np.zeros((1024, O.T, 3, 512 if FLAVOUR == "80" else 640), O.T0)
    out[..., :O.TLWE0] = keys.ksk
    return out


@pytest.fixture(scope="session")
def gpu_ctx(keys):
    """One CUDA context with keys loaded for the whole GPU session (fails loudly without the .so)."""
    from iyokan_b200 import Context

    ctx = Context(0)
    ctx.load_keys(keys.bk, keys.ksk)
    yield ctx
    ctx.close()
