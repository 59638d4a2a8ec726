"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

The reference itself is the source of every number written here; the restatement in
`oracle/xdeepfm_oracle.py` is only used for its deterministic parameter/input builders
(`make_params`, `make_inputs`) so that tests can regenerate identical inputs on the GPU box.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from oracle.xdeepfm_oracle import ModelSpec, make_inputs, make_params  # noqa: E402

load_reference()
from deepctr.inputs import DenseFeat, SparseFeat  # noqa: E402
from deepctr.layers.interaction import CIN  # noqa: E402
from deepctr.models import xDeepFM, xDeepFMAttention, xDeepFMAttentionV2  # noqa: E402


def spec_to_json(spec: ModelSpec):
    d = dict(spec.__dict__)
    d["cin_layer_size"] = list(d["cin_layer_size"])
    d["dnn_hidden_units"] = list(d["dnn_hidden_units"])
    d["sfg_hidden_units"] = list(d["sfg_hidden_units"])
    return d


def build_reference_model(spec: ModelSpec, **extra):
    cols = [SparseFeat(n, v, spec.embedding_dim) for n, v in zip(spec.sparse_names, spec.vocab_sizes)] + \
           [DenseFeat(n, 1) for n in spec.dense_names]
    common = dict(dnn_hidden_units=spec.dnn_hidden_units, cin_layer_size=spec.cin_layer_size,
                  cin_split_half=spec.cin_split_half, cin_activation=spec.cin_activation,
                  l2_reg_linear=spec.l2_reg_linear, l2_reg_embedding=spec.l2_reg_embedding,
                  l2_reg_dnn=spec.l2_reg_dnn, l2_reg_cin=spec.l2_reg_cin, device="cpu")
    common.update(extra)
    if spec.variant == "xdeepfm":
        return xDeepFM(cols, cols, **common)
    if spec.variant == "pro":
        from deepctr.xdeepfm_pro import xDeepFMPro
        return xDeepFMPro(cols, cols, use_sfg=spec.use_sfg, sfg_weight=spec.sfg_weight, sfg_hidden_units=spec.sfg_hidden_units,
                          sfg_dropout=0.0, sfg_positive_only=spec.sfg_positive_only,
                          sfg_use_label_attention=spec.sfg_use_label_attention, use_autodis=spec.use_autodis,
                          autodis_buckets=spec.autodis_buckets, autodis_temperature=spec.autodis_temperature, **common)
    if spec.variant == "attn":
        return xDeepFMAttention(cols, cols, cin_num_heads=spec.num_heads, cin_use_layer_norm=spec.use_layer_norm,
                                cin_use_residual=spec.use_residual, **common)
    return xDeepFMAttentionV2(cols, cols, cin_num_heads=spec.num_heads, cin_use_layer_norm=spec.use_layer_norm,
                              cin_use_residual=spec.use_residual, cin_num_attn_layers=spec.num_attn_layers, **common)


def cin_kat():
    """RNG-free known-answer test of the reference CIN (recipe from SURVEY.md 8c)."""
    cin = CIN(3, (4, 2), "relu", True)
    K = [9, 6]
    with torch.no_grad():
        for l, conv in enumerate(cin.conv1ds):
            H = conv.weight.shape[0]
            W = torch.tensor([[(((h * K[l] + k) * 3) % 7 - 2) / 10 for k in range(K[l])] for h in range(H)])
            conv.weight.copy_(W.unsqueeze(-1))
            conv.bias.copy_(torch.tensor([(h + 1) / 10 for h in range(H)]))
    x = torch.tensor([((2 * i) % 5 - 1) / 2 for i in range(12)], dtype=torch.float32).reshape(2, 3, 2).requires_grad_(True)
    out = cin(x)
    out.sum().backward()
    kat = {
        "recipe": "W_l[h,k,0]=(((h*K_l+k)*3)%7-2)/10 (K_0=9,K_1=6); b_l[h]=(h+1)/10; x.flat[i]=((2*i)%5-1)/2; x:[2,3,2]; L=out.sum()",
        "x": x.detach().flatten().tolist(),
        "out": out.detach().tolist(),
        "dx": x.grad.flatten().tolist(),
        "dW0": cin.conv1ds[0].weight.grad.flatten().tolist(),
        "dW1": cin.conv1ds[1].weight.grad.flatten().tolist(),
        "db0": cin.conv1ds[0].bias.grad.tolist(),
        "db1": cin.conv1ds[1].bias.grad.tolist(),
    }
    with open(os.path.join(GOLD, "cin_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("cin_kat out", kat["out"])


def forward_backward_case(name, spec: ModelSpec, B, seed, store_params=True, grad_keys=None, zipf=False, grad_sample_stride=0):
    """y_pred / BCE-sum / total loss / parameter gradients of one reference train-step (no optimizer).
    `grad_sample_stride` > 0: tensors outside `grad_keys` also leave every stride-th element (flat order) as `gradsample::`."""
    params = make_params(spec, seed=seed)
    X, y = make_inputs(spec, B, seed=seed, zipf=zipf)
    model = build_reference_model(spec)
    missing = model.load_state_dict(params, strict=True)
    model.train()
    y_pred = model(X).squeeze()
    loss = torch.nn.functional.binary_cross_entropy(y_pred, y, reduction="sum")   # basemodel.py:254
    reg = model.get_regularization_loss()                                          # basemodel.py:255
    total = loss + reg + model.aux_loss
    model.zero_grad()
    total.backward()
    out = {"X": X.numpy(), "y": y.numpy(), "y_pred": y_pred.detach().numpy(),
           "loss": loss.detach().numpy(), "reg_loss": reg.detach().numpy(), "total": total.detach().numpy()}
    named = dict(model.named_parameters())
    for k, p in named.items():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        if grad_keys is None or k in grad_keys:
            out["grad::" + k] = g.numpy()
        elif grad_sample_stride > 0:
            out["gradsample::" + k] = g.flatten()[::grad_sample_stride].numpy().copy()
        out["gradnorm::" + k] = np.float64(g.double().norm().item())
    if grad_sample_stride > 0:
        out["grad_sample_stride"] = np.int64(grad_sample_stride)
    model.eval()
    with torch.no_grad():
        out["y_pred_eval"] = model(X).numpy()
    if store_params:
        for k, v in params.items():
            out["param::" + k] = v.numpy()
    else:
        out["param_checksum"] = np.float64(sum(v.double().sum().item() for v in params.values()))
    out["spec_json"] = np.array(json.dumps(spec_to_json(spec)))
    out["seed"] = np.int64(seed)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "loss", float(loss), "total", float(total), "y_pred[:3]", y_pred[:3].tolist())


def pro_case(name, spec: ModelSpec, B, seed):
    """One reference xDeepFM Pro train step (basemodel_sfg.py:316-349): y_pred, BCE-sum, sfg_loss, total, parameter gradients."""
    params = make_params(spec, seed=seed)
    X, y = make_inputs(spec, B, seed=seed)
    model = build_reference_model(spec)
    model.load_state_dict(params, strict=True)
    model.train()
    y_pred, info = model.forward_with_sfg(X, y)
    y_pred = y_pred.squeeze()
    loss = torch.nn.functional.binary_cross_entropy(y_pred, y, reduction="sum")
    reg = model.get_regularization_loss()
    sfg = info["sfg_loss"] if info is not None else torch.tensor(0.0)
    total = loss + reg + model.aux_loss + model.sfg_weight * sfg
    model.zero_grad()
    total.backward()
    out = {"X": X.numpy(), "y": y.numpy(), "y_pred": y_pred.detach().numpy(), "loss": loss.detach().numpy(),
           "reg_loss": reg.detach().numpy(), "sfg_loss": np.float64(float(sfg)), "total": total.detach().numpy()}
    for k, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out["grad::" + k] = g.numpy()
    model.eval()
    with torch.no_grad():
        out["y_pred_eval"] = model(X).numpy()
    for k, v in params.items():
        out["param::" + k] = v.numpy()
    out["spec_json"] = np.array(json.dumps(spec_to_json(spec)))
    out["seed"] = np.int64(seed)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "loss", float(loss), "sfg", float(sfg), "total", float(total))


def fit_case(name, spec: ModelSpec, N, batch_size, epochs, optimizer, seed, lr=None, store_params=True, final_keys=None):
    """Trajectory of the reference's own fit(): History['loss'] per epoch + final weights + predictions.
    `store_params=False` (large shapes): seeded parameters are regenerated by the test (checksum stored), final weights are kept in
    full for `final_keys` only and as (norm of the final tensor, norm of its movement) for every tensor."""
    params = make_params(spec, seed=seed)
    X, y = make_inputs(spec, N, seed=seed)
    model = build_reference_model(spec)
    model.load_state_dict(params, strict=True)
    model.compile(optimizer, "binary_crossentropy", metrics=["binary_crossentropy", "auc"])
    if lr is not None:
        for gparam in model.optim.param_groups:       # xdftrain.py:282-284
            gparam["lr"] = lr
    names = list(model.feature_index.keys())
    xdict = {n: X[:, i].numpy().copy() for i, n in enumerate(names)}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        hist = model.fit(xdict, y.numpy().reshape(-1, 1), batch_size=batch_size, epochs=epochs, verbose=0,
                         shuffle=False, validation_data=(dict(xdict), y.numpy().reshape(-1, 1)))
        pred = model.predict(dict(xdict), batch_size=batch_size)
    out = {"X": X.numpy(), "y": y.numpy(), "pred": pred,
           "history_sfg_loss": np.array(hist.history.get("sfg_loss", []), dtype=np.float64),
           "history_loss": np.array(hist.history["loss"], dtype=np.float64),
           "history_val_auc": np.array(hist.history["val_auc"], dtype=np.float64),
           "history_val_bce": np.array(hist.history["val_binary_crossentropy"], dtype=np.float64),
           "spec_json": np.array(json.dumps(spec_to_json(spec))), "seed": np.int64(seed),
           "batch_size": np.int64(batch_size), "epochs": np.int64(epochs), "optimizer": np.array(optimizer),
           "lr": np.float64(-1.0 if lr is None else lr)}
    if store_params:
        for k, v in params.items():
            out["param::" + k] = v.numpy()
        for k, v in model.state_dict().items():
            out["final::" + k] = v.numpy()
    else:
        out["param_checksum"] = np.float64(sum(v.double().sum().item() for v in params.values()))
        for k, v in model.state_dict().items():
            if final_keys is not None and k in final_keys:
                out["final::" + k] = v.numpy()
            out["finalnorm::" + k] = np.float64(v.double().norm().item())
            out["movednorm::" + k] = np.float64((v.double() - params[k].double()).norm().item())
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "history loss", hist.history["loss"], "val_auc", hist.history["val_auc"])


def small_spec(**kw):
    base = dict(sparse_names=["C%d" % i for i in range(1, 6)], vocab_sizes=[7, 13, 50, 3, 29], embedding_dim=8,
                dense_names=["I1", "I2", "I3"], cin_layer_size=(16, 8), dnn_hidden_units=(32, 16),
                l2_reg_linear=1e-3, l2_reg_embedding=1e-3, l2_reg_dnn=1e-3, l2_reg_cin=1e-3)
    base.update(kw)
    return ModelSpec(**base)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    cin_kat()
    forward_backward_case("xdeepfm_small", small_spec(), B=32, seed=1)
    forward_backward_case("xdeepfm_small_nosplit", small_spec(cin_split_half=False, cin_layer_size=(6, 5, 4)), B=17, seed=2)
    forward_backward_case("xdeepfm_small_linearact", small_spec(cin_activation="linear", cin_layer_size=(8, 6, 3)), B=9, seed=3)
    forward_backward_case("xdeepfm_small_nodense", small_spec(dense_names=[], embedding_dim=4), B=16, seed=4)
    forward_backward_case("xdeepfm_small_zipf", small_spec(vocab_sizes=[1000, 13, 5000, 3, 29]), B=64, seed=5, zipf=True)
    forward_backward_case("attn_small", small_spec(variant="attn", num_heads=4), B=12, seed=6)
    forward_backward_case("attn_small_3heads", small_spec(variant="attn", num_heads=3, embedding_dim=10), B=8, seed=7)
    forward_backward_case("attn_v2_small", small_spec(variant="attn_v2", num_heads=2, num_attn_layers=2), B=12, seed=8)
    # BASELINE.json config 1 shape (26 sparse + 13 dense, D=8, CIN (256,128), DNN (256,256)); vocab 100/field
    cfg1 = ModelSpec(sparse_names=["C%d" % i for i in range(1, 27)], vocab_sizes=[100] * 26, embedding_dim=8,
                     dense_names=["I%d" % i for i in range(1, 14)])
    keep = {"embedding_dict.C1.weight", "embedding_dict.C26.weight", "linear_model.embedding_dict.C3.weight",
            "linear_model.weight", "out.bias", "cin.conv1ds.0.bias", "cin.conv1ds.1.bias", "cin_linear.weight",
            "dnn.linears.1.bias", "dnn_linear.weight"}
    forward_backward_case("xdeepfm_cfg1", cfg1, B=64, seed=9, store_params=False, grad_keys=keep)
    fit_case("fit_small_adam", small_spec(), N=96, batch_size=32, epochs=3, optimizer="adam", seed=11, lr=1e-2)
    fit_case("fit_small_sgd", small_spec(), N=96, batch_size=32, epochs=2, optimizer="sgd", seed=12)
    fit_case("fit_small_adagrad", small_spec(), N=80, batch_size=32, epochs=2, optimizer="adagrad", seed=13)
    fit_case("fit_small_rmsprop", small_spec(), N=80, batch_size=32, epochs=2, optimizer="rmsprop", seed=14, lr=1e-3)


def cfg2_spec(vocab=200, **kw):
    """BASELINE.json configs[1] shape -- the configuration bench.py measures: 26 sparse + 13 dense, D = 16, CIN (200, 200, 200),
    DNN (400, 400); vocabularies capped (the tables are not what the tensor-core path changes)."""
    base = dict(sparse_names=["C%d" % i for i in range(1, 27)], vocab_sizes=[vocab] * 26, embedding_dim=16,
                dense_names=["I%d" % i for i in range(1, 14)], cin_layer_size=(200, 200, 200), dnn_hidden_units=(400, 400),
                l2_reg_linear=1e-5, l2_reg_embedding=1e-5)
    base.update(kw)
    return ModelSpec(**base)


def main_cfg2():
    """Fixtures at the benchmarked shape (`python -m oracle.make_golden cfg2`): one train step (all gradients: small tensors and CIN
    layer 0 in full, the rest as norms + every 61st element) and a 2-epoch Adam fit() trajectory."""
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    keep = {"embedding_dict.C1.weight", "embedding_dict.C26.weight", "linear_model.embedding_dict.C3.weight",
            "linear_model.weight", "out.bias", "cin.conv1ds.0.bias", "cin.conv1ds.1.bias", "cin.conv1ds.2.bias", "cin_linear.weight",
            "dnn.linears.0.bias", "dnn.linears.1.bias", "dnn_linear.weight", "cin.conv1ds.0.weight"}
    forward_backward_case("xdeepfm_cfg2", cfg2_spec(), B=128, seed=61, store_params=False, grad_keys=keep, zipf=True,
                          grad_sample_stride=61)
    fit_case("fit_cfg2_adam", cfg2_spec(vocab=50), N=768, batch_size=128, epochs=2, optimizer="adam", seed=62, lr=1e-3,
             store_params=False, final_keys={"out.bias", "cin_linear.weight", "dnn_linear.weight", "cin.conv1ds.2.bias",
                                             "linear_model.weight", "embedding_dict.C2.weight"})


def main_pro():
    """xDeepFM Pro fixtures (added after the first batch; `python -m oracle.make_golden pro` regenerates only these)."""
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    pro = dict(variant="pro", sfg_hidden_units=(16, 8))
    pro_case("pro_small", small_spec(**pro), B=40, seed=21)
    pro_case("pro_small_allrows_noattn", small_spec(sfg_positive_only=False, sfg_use_label_attention=False, sfg_weight=0.5, **pro),
             B=24, seed=22)
    pro_case("pro_small_nodense", small_spec(dense_names=[], embedding_dim=4, **pro), B=24, seed=23)
    fit_case("fit_pro_small_adam", small_spec(**pro), N=96, batch_size=32, epochs=2, optimizer="adam", seed=24, lr=1e-2)
    main_autodis()


def main_autodis():
    """xDeepFM Pro with the AutoDis dense-feature encoder (`python -m oracle.make_golden autodis` regenerates only these)."""
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    pro = dict(variant="pro", sfg_hidden_units=(16, 8), use_autodis=True)
    pro_case("pro_autodis_small", small_spec(autodis_buckets=16, **pro), B=40, seed=31)
    pro_case("pro_autodis_b5_nosfg", small_spec(autodis_buckets=5, autodis_temperature=0.5, use_sfg=False, embedding_dim=4, **pro),
             B=150, seed=32)
    fit_case("fit_pro_autodis_adam", small_spec(autodis_buckets=8, **pro), N=96, batch_size=32, epochs=2, optimizer="adam", seed=33,
             lr=1e-2)


# ---- multi-value (VarLenSparseFeat) features: SURVEY.md 8f-4 -----------------------------------------------------
VARLEN_COLUMNS = [
    {"kind": "sparse", "name": "C1", "vocab": 7, "dim": 8},
    {"kind": "dense", "name": "I1", "dim": 1},
    {"kind": "varlen", "name": "G", "vocab": 11, "dim": 8, "maxlen": 5, "combiner": "mean", "length_name": None},
    {"kind": "sparse", "name": "C2", "vocab": 29, "dim": 8},
    {"kind": "varlen", "name": "H", "vocab": 9, "dim": 8, "maxlen": 3, "combiner": "max", "length_name": None},
    {"kind": "dense", "name": "I2", "dim": 2},
    {"kind": "varlen", "name": "S", "vocab": 13, "dim": 8, "maxlen": 4, "combiner": "sum", "length_name": "S_len"},
    {"kind": "varlen", "name": "M", "vocab": 6, "dim": 8, "maxlen": 2, "combiner": "mean", "length_name": "M_len"},
]


def columns_from_desc(desc, SparseFeat, DenseFeat, VarLenSparseFeat):
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


def varlen_inputs(desc, feature_index, n, seed):
    """Flat float32 input matrix in feature_index order: ids uniform (0 = padding, ~40 % of the sequence positions; some sequences
    entirely padded), lengths uniform in [0, maxlen], dense U[0,1)."""
    g = torch.Generator().manual_seed(seed)
    width = max(b for _, b in feature_index.values())
    X = torch.zeros(n, width)
    by_name = {d["name"]: d for d in desc}
    for name, (a, b) in feature_index.items():
        d = by_name.get(name)
        if d is None:                                      # a length column
            owner = [q for q in desc if q.get("length_name") == name][0]
            X[:, a] = torch.randint(0, owner["maxlen"] + 1, (n,), generator=g).float()
        elif d["kind"] == "dense":
            X[:, a:b] = torch.rand(n, b - a, generator=g)
        elif d["kind"] == "sparse":
            X[:, a] = torch.randint(0, d["vocab"], (n,), generator=g).float()
        else:
            ids = torch.randint(1, d["vocab"], (n, b - a), generator=g)
            keep = torch.rand(n, b - a, generator=g) < 0.6
            keep[::7] = False                              # every 7th sample: an entirely padded sequence
            if d["combiner"] == "max":
                keep[:, 0] = True      # 'max' over an entirely padded sequence is -1e9 in the reference (sequence.py:69-72): not a model input
            X[:, a:b] = (ids * keep).float()
    y = (torch.rand(n, generator=g) < 0.3).float()
    return X, y


def varlen_params(model, seed):
    """Deterministic, non-degenerate parameters for the reference model's own state_dict layout."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in model.state_dict().items():
        scale = 0.3 if "embedding_dict" in k else 0.08
        out[k] = (torch.randn(v.shape, generator=g) * scale).to(v.dtype)
    return out


def seqpool_cases():
    """SequencePoolingLayer of the reference (sequence.py:9-79): every mode under both mask conventions, outputs and input gradients."""
    from deepctr.layers.sequence import SequencePoolingLayer
    g = torch.Generator().manual_seed(41)
    out = {}
    for E in (8, 3, 1):
        B, T = 13, 6
        x = torch.randn(B, T, E, generator=g)
        w = torch.randn(B, 1, E, generator=g)
        mask = torch.rand(B, T, generator=g) < 0.6
        mask[3] = False
        mask[5] = True
        length = torch.randint(0, T + 1, (B, 1), generator=g)
        length[0], length[1] = 0, T
        out["x_E%d" % E], out["w_E%d" % E], out["mask_E%d" % E], out["len_E%d" % E] = x.numpy(), w.numpy(), mask.numpy(), length.numpy()
        for mode in ("sum", "mean", "max"):
            for masking in (True, False):
                if mode == "max" and not masking:
                    continue        # the reference raises here (sequence.py:69: `1 - mask` on the bool mask of _sequence_mask)
                xi = x.clone().requires_grad_(True)
                layer = SequencePoolingLayer(mode=mode, supports_masking=masking)
                o = layer([xi, mask if masking else length])
                (o * w).sum().backward()
                key = "%s_%s_E%d" % (mode, "mask" if masking else "len", E)
                out["out_" + key], out["dx_" + key] = o.detach().numpy(), xi.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "seqpool_cases.npz"), **out)
    print("seqpool_cases", len(out), "arrays")


def varlen_model_case(name, B, seed, variant="xdeepfm"):
    """One reference train step (no optimizer) with multi-value features in the linear, CIN and DNN parts: xDeepFM, the attention
    variants (`variant` = attn / attn_v2) or xDeepFMPro without SFG (the reference's SFG decoder does not accept multi-value
    columns: its input width is sized from the SparseFeat columns only and the forward fails with a shape error)."""
    from deepctr.inputs import VarLenSparseFeat
    cols = columns_from_desc(VARLEN_COLUMNS, SparseFeat, DenseFeat, VarLenSparseFeat)
    common = dict(dnn_hidden_units=(32, 16), cin_layer_size=(16, 8), l2_reg_linear=1e-3, l2_reg_embedding=1e-3, l2_reg_dnn=1e-3,
                  l2_reg_cin=1e-3, device="cpu")
    if variant == "attn":
        model = xDeepFMAttention(cols, cols, cin_num_heads=4, **common)
    elif variant == "attn_v2":
        model = xDeepFMAttentionV2(cols, cols, cin_num_heads=2, cin_num_attn_layers=2, **common)
    elif variant == "pro_nosfg":
        from deepctr.xdeepfm_pro import xDeepFMPro
        model = xDeepFMPro(cols, cols, use_sfg=False, use_autodis=False, **common)
    else:
        model = xDeepFM(cols, cols, **common)
    params = varlen_params(model, seed)
    model.load_state_dict(params, strict=True)
    X, y = varlen_inputs(VARLEN_COLUMNS, model.feature_index, B, seed)
    model.train()
    y_pred = model(X).squeeze()
    loss = torch.nn.functional.binary_cross_entropy(y_pred, y, reduction="sum")
    reg = model.get_regularization_loss()
    total = loss + reg + model.aux_loss
    model.zero_grad()
    total.backward()
    out = {"X": X.numpy(), "y": y.numpy(), "y_pred": y_pred.detach().numpy(), "loss": loss.detach().numpy(),
           "reg_loss": reg.detach().numpy(), "total": total.detach().numpy(), "columns_json": np.array(json.dumps(VARLEN_COLUMNS)),
           "feature_names": np.array(json.dumps(list(model.feature_index.keys()))), "variant": np.array(variant)}
    for k, p in model.named_parameters():
        out["grad::" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    model.eval()
    with torch.no_grad():
        out["y_pred_eval"] = model(X).numpy()
        out["linear_logit"] = model.linear_model(X).numpy()
    for k, v in params.items():
        out["param::" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "loss", float(loss), "total", float(total), "y_pred[:3]", y_pred[:3].tolist())


def varlen_fit_case(name, N, batch_size, epochs, seed, lr):
    """Trajectory of the reference's own fit() / predict() on dict inputs with [N, maxlen] id matrices."""
    from deepctr.inputs import VarLenSparseFeat
    cols = columns_from_desc(VARLEN_COLUMNS, SparseFeat, DenseFeat, VarLenSparseFeat)
    model = xDeepFM(cols, cols, dnn_hidden_units=(32, 16), cin_layer_size=(16, 8), l2_reg_linear=1e-3, l2_reg_embedding=1e-3,
                    l2_reg_dnn=1e-3, l2_reg_cin=1e-3, device="cpu")
    params = varlen_params(model, seed)
    model.load_state_dict(params, strict=True)
    X, y = varlen_inputs(VARLEN_COLUMNS, model.feature_index, N, seed)
    model.compile("adam", "binary_crossentropy", metrics=["binary_crossentropy", "auc"])
    for gparam in model.optim.param_groups:
        gparam["lr"] = lr
    xdict = {n: X[:, a:b].numpy().copy() if b - a > 1 else X[:, a].numpy().copy() for n, (a, b) in model.feature_index.items()}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        hist = model.fit(dict(xdict), y.numpy().reshape(-1, 1), batch_size=batch_size, epochs=epochs, verbose=0, shuffle=False,
                         validation_data=(dict(xdict), y.numpy().reshape(-1, 1)))
        pred = model.predict(dict(xdict), batch_size=batch_size)
    out = {"X": X.numpy(), "y": y.numpy(), "pred": pred, "history_loss": np.array(hist.history["loss"], dtype=np.float64),
           "history_val_auc": np.array(hist.history["val_auc"], dtype=np.float64),
           "history_val_bce": np.array(hist.history["val_binary_crossentropy"], dtype=np.float64),
           "columns_json": np.array(json.dumps(VARLEN_COLUMNS)),
           "feature_names": np.array(json.dumps(list(model.feature_index.keys()))),
           "batch_size": np.int64(batch_size), "epochs": np.int64(epochs), "lr": np.float64(lr)}
    for k, v in params.items():
        out["param::" + k] = v.numpy()
    for k, v in model.state_dict().items():
        out["final::" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "history loss", hist.history["loss"], "val_auc", hist.history["val_auc"])


def main_varlen():
    """Multi-value feature fixtures (`python -m oracle.make_golden varlen` regenerates only these)."""
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    seqpool_cases()
    varlen_model_case("xdeepfm_varlen", B=48, seed=51)
    varlen_fit_case("fit_varlen_adam", N=96, batch_size=32, epochs=2, seed=52, lr=1e-2)
    main_varlen_variants()


def main_varlen_variants():
    """Multi-value features in the other model classes (`python -m oracle.make_golden varlen_variants`)."""
    varlen_model_case("attn_varlen", B=40, seed=53, variant="attn")
    varlen_model_case("attn_v2_varlen", B=40, seed=54, variant="attn_v2")
    varlen_model_case("pro_nosfg_varlen", B=40, seed=55, variant="pro_nosfg")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pro":
        main_pro()
    elif len(sys.argv) > 1 and sys.argv[1] == "autodis":
        main_autodis()
    elif len(sys.argv) > 1 and sys.argv[1] == "varlen":
        main_varlen()
    elif len(sys.argv) > 1 and sys.argv[1] == "varlen_variants":
        os.makedirs(GOLD, exist_ok=True)
        torch.set_num_threads(4)
        main_varlen_variants()
    elif len(sys.argv) > 1 and sys.argv[1] == "cfg2":
        main_cfg2()
    else:
        main()
        main_pro()
        main_varlen()
