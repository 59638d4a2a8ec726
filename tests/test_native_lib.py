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


def test_bag_layout_validation_without_gpu():
    """xdfm_bag_pool_fwd / _bwd check the slot -> field layout on the host before anything is launched (B = 0: no launch)."""
    L = _native.lib()
    a = _native.i32_array

    def fwd(slot0, slen, mode, lencol, S, nlen=0, lens=None, argmax=None, den=None):
        return L.xdfm_bag_pool_fwd(None, None, lens, nlen, 0, S, 8, len(slot0), a(slot0), a(slen), a(mode), a(lencol), None, argmax, den, None)

    dummy = ctypes.c_void_p(16)          # never dereferenced with B = 0
    assert fwd([0, 1], [1, 3], [0, 1], [-1, -1], 4) == 0                                    # fixed field + 'sum' over 3 slots
    assert fwd([0, 1], [1, 3], [0, 2], [-1, -1], 4, den=dummy) == 0                         # 'mean' with its divisor buffer
    assert fwd([0, 1], [1, 3], [0, 2], [-1, -1], 4) != 0 and b"divisor" in L.xdfm_last_error()
    assert fwd([0, 1], [1, 3], [0, 3], [-1, -1], 4) != 0 and b"argmax" in L.xdfm_last_error()
    assert fwd([0, 2], [1, 3], [0, 1], [-1, -1], 5) != 0 and b"contiguous" in L.xdfm_last_error()  # gap between the fields
    assert fwd([0, 1], [1, 3], [0, 1], [-1, -1], 5) != 0 and b"cover" in L.xdfm_last_error()       # slots left over
    assert fwd([0, 1], [2, 3], [0, 1], [-1, -1], 5) != 0 and b"exactly one slot" in L.xdfm_last_error()
    assert fwd([0, 1], [1, 3], [0, 7], [-1, -1], 4) != 0 and b"unknown mode" in L.xdfm_last_error()
    assert fwd([0, 1], [1, 3], [0, 1], [-1, 0], 4) != 0 and b"length column" in L.xdfm_last_error()  # no lens tensor given
    assert fwd([0, 1], [1, 3], [0, 1], [-1, 0], 4, nlen=1, lens=dummy) == 0
    assert fwd([0], [65], [1], [-1], 65) != 0                                                # more than 64 slots
    rc = L.xdfm_bag_pool_bwd(None, None, None, 0, None, None, 0, 4, 8, 2, a([0, 1]), a([1, 3]), a([0, 1]), a([-1, -1]), None, None)
    assert rc == 0
