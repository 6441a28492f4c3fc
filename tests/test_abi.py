"""The C-ABI library loads and exports every symbol include/pantas_aug.h declares (no GPU needed)."""
import os
import re

from pantas_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pantas_aug.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build_library()
    lib = _lib.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_strerror_and_version_work_without_a_gpu():
    lib = _lib.load_library()
    assert lib.pt_abi_version() == 1
    assert b"KeyError" in lib.pt_strerror(6)


def test_create_fails_loudly_without_a_device():
    import ctypes

    import torch

    if torch.cuda.is_available():
        return
    lib = _lib.load_library()
    ctx = ctypes.c_void_p()
    assert lib.pt_create(0, ctypes.byref(ctx)) < 0
    from pantas_b200.engine import AugmentEngine
    from pantas_b200.errors import NativeLibraryError
    import pytest

    with pytest.raises(NativeLibraryError):
        AugmentEngine(0)
