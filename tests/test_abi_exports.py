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
