"""CPU: the C-ABI library builds/loads and exports exactly the symbols include/fcp_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def header_functions():
    text = (REPO / "include" / "fcp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcp_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from face_crop_plus_b200 import _abi
    assert header_functions() == sorted(_abi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from face_crop_plus_b200 import _abi, build
    build.build()
    lib = _abi.load_library()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.fcp_version()


def test_no_cpu_fallback():
    import torch
    from face_crop_plus_b200 import _abi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_abi.FcpError):
        _abi.Context(0)


def test_product_never_imports_oracle():
    for p in (REPO / "face_crop_plus_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
