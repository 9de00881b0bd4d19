"""CPU-only: libhsg.so loads and exports every symbol include/hsg.h declares (no compute calls)."""
import ctypes
import os
import re

from higashi_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "hsg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hsg_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libhsg.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert _lib.load().hsg_abi_version() == 1


def test_weight_slot_enum_matches_header():
    text = open(os.path.join(ROOT, "include", "hsg.h")).read()
    body = text[text.index("enum hsg_weight"):text.index("HSG_W_COUNT")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"HSG_(W_[A-Z0-9_]+)", body)
    assert names == _lib.WEIGHT_SLOTS


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(_lib.HsgError):
        _lib.Context(0)
