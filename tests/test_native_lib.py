"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol include/xdfm.h declares
(no compute calls without a GPU)."""
import ctypes
import os

import pytest

from deepctr import _native


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as G
    G.build()
    L = ctypes.CDLL(_native.LIB_PATH)
    syms = _native.header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export: " + s
    assert set(_native._SIGS.keys()) == set(syms), set(_native._SIGS.keys()) ^ set(syms)


def test_version_and_error_string():
    L = _native.lib()
    assert L.xdfm_version() >= 100
    assert isinstance(L.xdfm_last_error(), bytes)


def test_argument_validation_without_gpu():
    L = _native.lib()
    rc = L.xdfm_split_input(None, 4, 3, _native.i32_array([0]), 1000, _native.i32_array([]), 0, None, None, None)
    assert rc != 0 and b"out of range" in L.xdfm_last_error()
    assert L.xdfm_embed_bwd_workspace_bytes(1000) > 0


def test_ops_raise_on_cpu_tensors():
    import torch
    from deepctr import ops
    with pytest.raises(RuntimeError):
        ops.split_input(torch.zeros(2, 3), [0], [1, 2])
