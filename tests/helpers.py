"""Shared fixture loading for the parity tests (oracle = checker only)."""
import json
import os

import numpy as np
import torch

from oracle.xdeepfm_oracle import ModelSpec, make_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FWD_BWD_CASES = ["xdeepfm_small", "xdeepfm_small_nosplit", "xdeepfm_small_linearact", "xdeepfm_small_nodense",
                 "xdeepfm_small_zipf", "attn_small", "attn_small_3heads", "attn_v2_small", "xdeepfm_cfg1", "xdeepfm_cfg2"]
PRO_CASES = ["pro_small", "pro_small_allrows_noattn", "pro_small_nodense", "pro_autodis_small", "pro_autodis_b5_nosfg"]
FIT_CASES = ["fit_small_adam", "fit_small_sgd", "fit_small_adagrad", "fit_small_rmsprop"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = json.loads(str(z["spec_json"]))
    d["cin_layer_size"] = tuple(d["cin_layer_size"])
    d["dnn_hidden_units"] = tuple(d["dnn_hidden_units"])
    if "sfg_hidden_units" in d:
        d["sfg_hidden_units"] = tuple(d["sfg_hidden_units"])
    spec = ModelSpec(**d)
    params = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    if not params:
        params = make_params(spec, seed=int(z["seed"]))
        chk = sum(v.double().sum().item() for v in params.values())
        assert abs(chk - float(z["param_checksum"])) < 1e-6 * max(1.0, abs(chk)), "seeded params differ from fixture"
    return spec, params, z


def golden_grads(z):
    return {k[len("grad::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad::")}


def golden_gradsamples(z):
    """{name: (stride, every stride-th element of the flattened reference gradient)} for tensors too large to store whole."""
    if "grad_sample_stride" not in z.files:
        return {}
    stride = int(z["grad_sample_stride"])
    return {k[len("gradsample::"):]: (stride, torch.from_numpy(z[k])) for k in z.files if k.startswith("gradsample::")}


def golden_gradnorms(z):
    return {k[len("gradnorm::"):]: float(z[k]) for k in z.files if k.startswith("gradnorm::")}


def build_product_model(spec, device="cuda:0", **extra):
    """The product model (xdeepfm-pytorch_b200/deepctr) for a ModelSpec."""
    from deepctr.inputs import DenseFeat, SparseFeat
    from deepctr import models as M
    cols = [SparseFeat(n, v, spec.embedding_dim) for n, v in zip(spec.sparse_names, spec.vocab_sizes)] + \
           [DenseFeat(n, 1) for n in spec.dense_names]
    common = dict(dnn_hidden_units=spec.dnn_hidden_units, cin_layer_size=spec.cin_layer_size,
                  cin_split_half=spec.cin_split_half, cin_activation=spec.cin_activation,
                  l2_reg_linear=spec.l2_reg_linear, l2_reg_embedding=spec.l2_reg_embedding,
                  l2_reg_dnn=spec.l2_reg_dnn, l2_reg_cin=spec.l2_reg_cin, device=device)
    common.update(extra)
    if spec.variant == "xdeepfm":
        return M.xDeepFM(cols, cols, **common)
    if spec.variant == "pro":
        from deepctr.xdeepfm_pro import xDeepFMPro
        kw = dict(use_sfg=spec.use_sfg, sfg_weight=spec.sfg_weight, sfg_hidden_units=spec.sfg_hidden_units, sfg_dropout=0.0,
                  sfg_positive_only=spec.sfg_positive_only, sfg_use_label_attention=spec.sfg_use_label_attention,
                  use_autodis=spec.use_autodis, autodis_buckets=spec.autodis_buckets, autodis_temperature=spec.autodis_temperature)
        kw.update(common)
        return xDeepFMPro(cols, cols, **kw)
    if spec.variant == "attn":
        return M.xDeepFMAttention(cols, cols, cin_num_heads=spec.num_heads, cin_use_layer_norm=spec.use_layer_norm,
                                  cin_use_residual=spec.use_residual, **common)
    return M.xDeepFMAttentionV2(cols, cols, cin_num_heads=spec.num_heads, cin_use_layer_norm=spec.use_layer_norm,
                                cin_use_residual=spec.use_residual, cin_num_attn_layers=spec.num_attn_layers, **common)


def assert_close(a, b, rtol, atol, what=""):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    if not bool((err <= tol).all()):
        i = int(torch.argmax(err - tol))
        raise AssertionError("%s: max |err| %.3e (ref %.3e) at flat index %d; rtol %.1e atol %.1e; frac bad %.4f" % (
            what, err.flatten()[i].item(), b.flatten()[i].item(), i, rtol, atol, float((err > tol).double().mean())))


def load_varlen_case(name):
    """Fixture of a model with multi-value features: (column descriptors, feature names, params, npz)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    desc = json.loads(str(z["columns_json"]))
    names = json.loads(str(z["feature_names"]))
    params = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    return desc, names, params, z


def product_columns(desc):
    from deepctr.inputs import DenseFeat, SparseFeat, VarLenSparseFeat
    cols = []
    for d in desc:
        if d["kind"] == "sparse":
            cols.append(SparseFeat(d["name"], d["vocab"], d["dim"]))
        elif d["kind"] == "dense":
            cols.append(DenseFeat(d["name"], d["dim"]))
        else:
            cols.append(VarLenSparseFeat(SparseFeat(d["name"], d["vocab"], d["dim"]), maxlen=d["maxlen"], combiner=d["combiner"],
                                         length_name=d["length_name"]))
    return cols


def build_varlen_product_model(desc, device="cuda:0", variant="xdeepfm"):
    from deepctr.models import xDeepFM, xDeepFMAttention, xDeepFMAttentionV2
    cols = product_columns(desc)
    common = dict(dnn_hidden_units=(32, 16), cin_layer_size=(16, 8), l2_reg_linear=1e-3, l2_reg_embedding=1e-3, l2_reg_dnn=1e-3,
                  l2_reg_cin=1e-3, device=device)
    if variant == "attn":
        return xDeepFMAttention(cols, cols, cin_num_heads=4, **common)
    if variant == "attn_v2":
        return xDeepFMAttentionV2(cols, cols, cin_num_heads=2, cin_num_attn_layers=2, **common)
    if variant == "pro_nosfg":
        from deepctr.xdeepfm_pro import xDeepFMPro
        return xDeepFMPro(cols, cols, use_sfg=False, use_autodis=False, **common)
    return xDeepFM(cols, cols, **common)
