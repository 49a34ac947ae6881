"""The C-ABI library loads without a GPU and exports every symbol include/dove_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "dove_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dove_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from dove_b200 import _lib, build
    build.build(verbose=False)
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dove_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(names)
    assert lib.dove_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    from dove_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.DoveError):
        _lib.init()
    lib = _lib.load()
    assert lib.dove_init(0) != 0 and b"no CUDA device" in lib.dove_last_error()
    # compute entry points refuse to run un-initialised
    assert lib.dove_post_scale_bf16(None, None, 0, None) == -3


def test_product_never_imports_oracle():
    for f in (ROOT / "dove_b200").glob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
