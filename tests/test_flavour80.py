"""The 80-bit parameter flavour (BASELINE.json configs[4]) is a second build of the same sources, selected per process
like the reference's -DIYOKAN_80BIT_SECURITY: its tests (tests/flavour80/) run in a child pytest with B200FHE_FLAVOUR=80."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _run(marker):
    env = dict(os.environ, B200FHE_FLAVOUR="80")
    return subprocess.run([sys.executable, "-m", "pytest", str(ROOT / "tests" / "flavour80"), "-x", "-q", "-m", marker,
                           "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)


def test_flavour80_cpu_suite():
    r = _run("not gpu")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


@pytest.mark.gpu
def test_flavour80_gpu_suite():
    r = _run("gpu")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
