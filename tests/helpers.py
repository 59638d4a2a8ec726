"""Shared fixture loading for the parity tests (oracle = checker only)."""
import json
import os

import numpy as np
import torch

from oracle.xdeepfm_oracle import ModelSpec, make_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FWD_BWD_CASES = ["xdeepfm_small", "xdeepfm_small_nosplit", "xdeepfm_small_linearact", "xdeepfm_small_nodense",
                 "xdeepfm_small_zipf", "attn_small", "attn_small_3heads", "attn_v2_small", "xdeepfm_cfg1"]
FIT_CASES = ["fit_small_adam", "fit_small_sgd", "fit_small_adagrad", "fit_small_rmsprop"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = json.loads(str(z["spec_json"]))
    d["cin_layer_size"] = tuple(d["cin_layer_size"])
    d["dnn_hidden_units"] = tuple(d["dnn_hidden_units"])
    spec = ModelSpec(**d)
    params = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    if not params:
        params = make_params(spec, seed=int(z["seed"]))
        chk = sum(v.double().sum().item() for v in params.values())
        assert abs(chk - float(z["param_checksum"])) < 1e-6 * max(1.0, abs(chk)), "seeded params differ from fixture"
    return spec, params, z


def golden_grads(z):
    return {k[len("grad::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad::")}


def golden_gradnorms(z):
    return {k[len("gradnorm::"):]: float(z[k]) for k in z.files if k.startswith("gradnorm::")}
