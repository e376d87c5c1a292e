"""The C-ABI library loads without a GPU and exports every symbol include/b200fhe.h declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def declared_functions():
    text = (ROOT / "include" / "b200fhe.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fhe_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from iyokan_b200 import build as B
    from iyokan_b200 import lib

    so = B.build_cuda()
    h = ctypes.CDLL(str(so))
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/b200fhe.h but not exported by {so.name}"
    assert sorted(lib.EXPORTS) == names  # the ctypes binding covers the whole header


def test_no_gpu_means_loud_failure():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from iyokan_b200 import B200FheError, Context

    with pytest.raises(B200FheError):
        Context(0)


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under iyokan_b200/ may reference it
    for path in (ROOT / "iyokan_b200").rglob("*"):
        if path.suffix in (".py", ".h", ".cu", ".cpp"):
            txt = path.read_text()
            assert "import oracle" not in txt and "tfhe_oracle" not in txt, path


def test_launch_plan_covers_every_frontier_size():
    # host-only logic (no device): segments cover the frontier exactly, throughput waves come first,
    # narrow frontiers use the latency shapes, and the <= 74-job tail goes to the 2-SM cluster shape
    from iyokan_b200 import build as B
    from iyokan_b200.lib import plan_rotation

    B.build_cuda()
    from iyokan_b200.lib import plan_ms

    rank = {(7, 8): 0, (3, 6): 1, (3, 4): 2, (4, 1): 3, (6, 1): 4}
    prev = 0.0
    for n in list(range(1, 2500)) + [4115, 8192, 8961, 100000]:
        plan = plan_rotation(n)
        assert 1 <= len(plan) <= 5
        assert sum(c for _, _, c in plan) == n
        order = [(v, g) for v, g, _ in plan]
        assert order == sorted(order, key=lambda vg: rank[vg]) and len(set(order)) == len(order)
        for v, g, c in plan:
            if (v, g) == (6, 1):
                assert c <= 74
        ms = plan_ms(n)
        assert ms >= prev - 1e-9   # the modelled time never decreases with the frontier size
        prev = ms if n < 2500 else prev
    assert plan_rotation(30) == [(6, 1, 30)]
    assert plan_rotation(74) == [(6, 1, 74)]
    assert plan_rotation(100) == [(4, 1, 100)]
    assert plan_rotation(148 + 20) == [(4, 1, 148), (6, 1, 20)]
    assert plan_rotation(8192)[0][:2] == (7, 8)      # wide frontiers: the 16-warp throughput shape
    assert plan_rotation(0) == []


def test_headers_are_plain_c(tmp_path):
    # the drop-in boundary is a C ABI: both headers must compile as C99 (no C++-isms, no CUDA or torch types) and a C
    # caller must link against the libraries
    import shutil
    import subprocess

    from iyokan_b200 import build as B

    B.build_cuda()
    B.build_host()
    cc = "/usr/bin/gcc" if shutil.which("/usr/bin/gcc") else "gcc"
    src = tmp_path / "caller.c"
    src.write_text('''
#include "b200fhe.h"
#include "b200net.h"
#include <stdio.h>
int main(void) {
    b200fhe_ctx *ctx = 0;
    int variant[8], g[8], jobs[8];
    int n = b200fhe_plan_rotation(8192, variant, g, jobs, 8);
    printf("%d segments, %.2f ms modelled\\n", n, b200fhe_plan_ms(8192));
    if (b200fhe_create(&ctx, 0) != 0) { printf("no device: %s\\n", b200fhe_last_error()); return n >= 1 ? 0 : 1; }
    b200fhe_destroy(ctx);
    return 0;
}
''')
    exe = tmp_path / "caller"
    csrc, host = ROOT / "iyokan_b200" / "csrc", ROOT / "iyokan_b200" / "host"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
                        f"-L{csrc}", "-lb200fhe", f"-L{host}", "-lb200net", f"-Wl,-rpath,{csrc}", f"-Wl,-rpath,{host}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "segments" in r.stdout
