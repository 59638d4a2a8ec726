"""ctypes binding of libxdfm_sm100a.so (C ABI declared in include/xdfm.h).

There is no CPU fallback: if the library is missing or the device is not CUDA, the ops raise.
"""
import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("XDFM_LIB", os.path.join(_PKG_ROOT, "libxdfm_sm100a.so"))
HEADER_PATH = os.path.join(os.path.dirname(_PKG_ROOT), "include", "xdfm.h")

MAX_FIELDS = 64
MAX_DENSE = 256
ACT = {"none": 0, "linear": 0, None: 0, "relu": 1, "tanh": 2, "sigmoid": 3}
BAG = {"single": 0, "sum": 1, "mean": 2, "max": 3}
OPT = {"sgd": 0, "adam": 1, "adagrad": 2, "rmsprop": 3}


class OptCfg(ctypes.Structure):
    _fields_ = [("kind", c_int32), ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
                ("alpha", c_float), ("lr_decay", c_float), ("l2", c_float)]


_P = c_void_p
_SIGS = {
    "xdfm_last_error": (c_char_p, []),
    "xdfm_version": (c_int, []),
    "xdfm_device_cc": (c_int, []),
    "xdfm_launch_count": (ctypes.c_longlong, []),
    "xdfm_split_input": (c_int, [_P, c_int64, c_int, POINTER(c_int32), c_int, POINTER(c_int32), c_int, _P, _P, _P]),
    "xdfm_embed_gather": (c_int, [POINTER(_P), POINTER(_P), POINTER(c_int32), _P, c_int64, c_int, c_int, _P, _P, c_int, _P, _P, _P]),
    "xdfm_embed_bwd_workspace_bytes": (c_int64, [c_int64]),
    "xdfm_embed_bwd_segments": (c_int, [_P, c_int64, c_int, POINTER(c_int64), POINTER(c_int32), c_int64, _P, c_int64, _P, _P, _P, _P, _P]),
    "xdfm_embed_bwd_reduce": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "xdfm_embed_bwd_scatter_dense": (c_int, [POINTER(_P), POINTER(c_int64), c_int, c_int, _P, _P, _P, c_int64, _P]),
    "xdfm_cin_fwd_f32": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_int, c_int, _P]),
    "xdfm_cin_dy": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, c_int, _P, _P]),
    "xdfm_cin_bwd_f32_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int, c_int]),
    "xdfm_cin_bwd_f32": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int64, _P]),
    "xdfm_gemm_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "xdfm_gemm_f32": (c_int, [c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, _P, c_int64, _P]),
    "xdfm_act_bwd": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "xdfm_wcolsum_workspace_bytes": (c_int64, [c_int]),
    "xdfm_wcolsum": (c_int, [_P, c_int64, c_int, c_int, _P, _P, c_int, _P, c_int64, _P]),
    "xdfm_head_fwd": (c_int, [_P, _P, _P, c_int, _P, _P, c_int, _P, c_int64, c_int, _P, _P]),
    "xdfm_head_bwd": (c_int, [_P, _P, c_int64, c_int, _P, c_int, _P, c_int, _P, _P, _P, _P]),
    "xdfm_head_bwd_fused_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "xdfm_head_bwd_fused": (c_int, [_P, _P, c_int64, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, _P, _P, _P, c_int64, _P]),
    "xdfm_bce_sum": (c_int, [_P, _P, c_int64, c_float, _P, _P, _P, _P]),
    "xdfm_to_rows_bf16": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, _P]),
    "xdfm_cin_tc_set_cluster": (None, [c_int]),
    "xdfm_cin_tc_set_pair": (None, [c_int]),
    "xdfm_cin_tc_wprime_elems": (c_int64, [c_int, c_int, c_int, c_int]),
    "xdfm_cin_fwd_tc": (c_int, [_P, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_int,
                                c_int, _P]),
    "xdfm_cin_dy_rows": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, c_int64, c_int, c_int, _P, _P]),
    "xdfm_cin_dy_rows_cols": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, c_int64, c_int, c_int, _P, _P,
                                      _P]),
    "xdfm_cin_dy_db_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "xdfm_cin_dy_rows_cols_db": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, c_int64, c_int, c_int, _P, _P,
                                         _P, _P, c_int64, _P]),
    "xdfm_from_rows_f32": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, c_int, _P]),
    "xdfm_add_rows_f32": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, _P]),
    "xdfm_cin_bwd_dx_tc_wt_elems": (c_int64, [c_int, c_int, c_int, c_int]),
    "xdfm_cin_bwd_dx_tc": (c_int, [_P, _P, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "xdfm_cin_bwd_dx_tc_dy": (c_int, [_P, _P, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, c_int64, c_int, _P]),
    "xdfm_rows_to_cols_bf16": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P]),
    "xdfm_cin_bwd_dw_tc_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int, c_int]),
    "xdfm_cin_bwd_dw_tc": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, c_int64, _P]),
    "xdfm_tc_selftest_gemm": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "xdfm_tc_latency_probe": (c_int, [_P, c_int, _P]),
    "xdfm_ipc_alloc": (c_int, [c_int64, POINTER(_P)]),
    "xdfm_ipc_free": (c_int, [_P]),
    "xdfm_ipc_export": (c_int, [_P, _P]),
    "xdfm_ipc_open": (c_int, [_P, POINTER(_P)]),
    "xdfm_ipc_close": (c_int, [_P]),
    "xdfm_embed_gather_sharded": (c_int, [_P, _P, _P, POINTER(c_int32), _P, c_int64, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P]),
    "xdfm_shard_workspace_bytes": (c_int64, [c_int64]),
    "xdfm_shard_segments": (c_int, [_P, c_int64, c_int, c_int, c_uint32, _P, POINTER(c_int32), _P, c_int64, _P, _P, _P, _P, _P, _P]),
    "xdfm_shard_pull_segments": (c_int, [POINTER(_P), POINTER(_P), POINTER(_P), POINTER(_P), c_int, c_int, c_uint32, c_int, c_int64,
                                         _P, c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "xdfm_mhsa_fwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P]),
    "xdfm_mhsa_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P]),
    "xdfm_add_ln_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_float, c_int, _P, _P, _P, _P]),
    "xdfm_add_ln_bwd_blocks": (c_int, [c_int64]),
    "xdfm_add_ln_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P]),
    "xdfm_attn_pool_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "xdfm_attn_pool_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "xdfm_small_linear_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_int, _P, _P, _P, _P]),
    "xdfm_small_linear_set_staged": (None, [c_int]),
    "xdfm_mhsa_set_row_blocked": (None, [c_int]),
    "xdfm_mhsa_fwd_dropout": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, ctypes.c_float, _P, _P, _P, _P]),
    "xdfm_mhsa_bwd_dropout": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, ctypes.c_float, _P, _P, _P, _P, _P]),
    "xdfm_mhsa_dropout_mask": (c_int, [c_int64, c_int, c_int, ctypes.c_float, _P, _P, _P]),
    "xdfm_small_linear_bwd_dx": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "xdfm_small_linear_bwd_dw_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int]),
    "xdfm_small_linear_bwd_dw": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "xdfm_autodis_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "xdfm_autodis_param_count": (c_int64, [c_int, c_int]),
    "xdfm_autodis_bwd_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int]),
    "xdfm_autodis_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P]),
    "xdfm_bag_pool_fwd": (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32),
                                  POINTER(c_int32), _P, _P, _P, _P]),
    "xdfm_bag_pool_bwd": (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32),
                                  POINTER(c_int32), POINTER(c_int32), _P, _P]),
    "xdfm_sfg_row_weights": (c_int, [_P, c_int64, c_int, _P, _P]),
    "xdfm_masked_ce": (c_int, [_P, _P, c_int64, _P, c_int64, c_int, _P, _P, _P]),
    "xdfm_masked_mse": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, _P]),
    "xdfm_gemm_tc_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "xdfm_gemm_tc": (c_int, [c_int, c_int, c_int, _P, c_int64, _P, c_int64, _P, c_int, _P, c_int, _P, c_int64, _P]),
    "xdfm_cvt_bf16": (c_int, [_P, c_int, c_int, c_int64, c_int, _P, c_int64, _P]),
    "xdfm_cvt_bf16_both_workspace_bytes": (c_int64, [c_int, c_int]),
    "xdfm_cvt_bf16_both": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, _P, c_int64, _P, c_int64, _P, _P, c_int64, _P]),
    "xdfm_set_rows_opt_dense_version": (None, [c_int]),
    "xdfm_set_replay_packed": (c_int, [c_int]),
    "xdfm_cin_dx_set_debug": (None, [c_int]),
    "xdfm_cin_dw_set_jp": (None, [c_int]),
    "xdfm_cin_dw_set_pack": (None, [c_int]),
    "xdfm_cin_dx0_finish": (c_int, [_P, c_int, _P, c_int64, c_int64, c_int, c_int, c_int, _P, _P]),
    "xdfm_cin_dx_set_trace": (None, [_P]),
    "xdfm_opt_tick_hist": (c_int, [_P, POINTER(OptCfg), _P, c_int64, c_int64, _P]),
    "xdfm_rows_catchup": (c_int, [POINTER(OptCfg), _P, _P, c_int64, POINTER(_P), POINTER(_P), POINTER(_P), _P, POINTER(c_int64), c_int, c_int,
                                  _P, _P, c_int64, _P, _P]),
    "xdfm_rows_mark_current": (c_int, [_P, _P, _P, c_int64, _P, _P]),
    "xdfm_rows_flush": (c_int, [POINTER(OptCfg), _P, _P, c_int64, POINTER(_P), POINTER(_P), POINTER(_P), _P, POINTER(c_int64), c_int, c_int, _P,
                                _P]),
    "xdfm_embed_gather_sharded_lazy": (c_int, [_P, _P, POINTER(c_int32), _P, c_int64, c_int, c_int, c_int, POINTER(OptCfg), POINTER(OptCfg), _P, _P,
                                               c_int64, _P, _P, c_int, _P, _P, _P]),
    "xdfm_embed_fetch_unique_sharded": (c_int, [_P, c_int, c_uint32, c_int, _P, _P, _P, _P, c_int64, POINTER(OptCfg), POINTER(OptCfg), _P, _P,
                                                c_int64, _P, _P, _P, _P]),
    "xdfm_embed_expand_unique": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P, c_int, _P, _P, _P]),
    "xdfm_embed_expand_unique_lin_rows": (c_int, [_P, _P, c_int64, _P, _P]),
    "xdfm_opt_tick": (c_int, [_P, POINTER(OptCfg), _P]),
    "xdfm_flat_opt": (c_int, [POINTER(OptCfg), _P, c_int64, _P, _P, _P, _P, _P, c_float, _P, _P]),
    "xdfm_rows_opt": (c_int, [POINTER(OptCfg), _P, POINTER(_P), POINTER(_P), POINTER(_P), POINTER(c_int64), c_int, c_int, _P, _P, _P,
                              c_int64, c_float, _P, c_int, _P, _P]),
}

_lib = None


def header_symbols():
    """Every function name declared in include/xdfm.h."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xdfm_[a-z0-9_]+)\s*\(", src)))


def lib():
    """Load the shared library (once) and attach signatures.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libxdfm_sm100a.so not found at %s -- build it with `python xdeepfm-pytorch_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("XDFM_CIN_DW_JP") == "1":      # experiment switch (profiles/r01d_cin_dw_findings.md); default: automatic
        L.xdfm_cin_dw_set_jp(1)
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError("libxdfm: " + lib().xdfm_last_error().decode())


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be a contiguous CUDA tensor."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("xdeepfm-b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("xdeepfm-b200 ops need contiguous tensors")
    return c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else ptr(t).value
    return arr


def i32_array(vals):
    return (c_int32 * len(vals))(*[int(v) for v in vals])


def i64_array(vals):
    return (c_int64 * len(vals))(*[int(v) for v in vals])
