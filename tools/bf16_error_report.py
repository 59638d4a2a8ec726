"""Per-tensor error of the bf16 (tcgen05) configuration against the reference fixtures: which part (CIN / DNN) contributes what.
    python tools/bf16_error_report.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "xdeepfm-pytorch_b200")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from tests.helpers import build_product_model, golden_grads, golden_gradsamples, load_case

DEV = "cuda:0"
cases = sys.argv[1:] or ["xdeepfm_small", "xdeepfm_small_zipf", "xdeepfm_cfg1", "xdeepfm_cfg2"]
for name in cases:
    spec, params, z = load_case(name)
    for cin_p, dnn_p in (("fp32", "fp32"), ("bf16", "fp32"), ("fp32", "bf16"), ("bf16", "bf16")):
        model = build_product_model(spec, DEV)
        model.load_state_dict(params, strict=True)
        model.cin.precision, model.dnn.precision = cin_p, dnn_p
        X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
        model.train()
        y_pred = model(X).squeeze()
        loss = F.binary_cross_entropy(y_pred, y, reduction="sum")
        total = loss + model.get_regularization_loss()
        model.zero_grad()
        total.backward()
        named = dict(model.named_parameters())
        yp = torch.from_numpy(z["y_pred"]).to(DEV)
        print("%s cin=%s dnn=%s: y_pred max abs err %.3e max rel %.3e; loss rel %.3e" % (
            name, cin_p, dnn_p, (y_pred - yp).abs().max().item(), ((y_pred - yp).abs() / yp).max().item(),
            abs(loss.item() - float(z["loss"])) / float(z["loss"])))
        rows = []
        for k, g in list(golden_grads(z).items()) + [(k + "[::s]", v[1]) for k, v in golden_gradsamples(z).items()]:
            kk = k.replace("[::s]", "")
            got = named[kk].grad.detach().double().cpu()
            if k.endswith("[::s]"):
                got = got.flatten()[::golden_gradsamples(z)[kk][0]]
            ref = g.double()
            err = (got - ref).abs()
            mx = ref.abs().max().item()
            big = ref.abs() >= 0.05 * mx
            rows.append((k, err.max().item() / max(mx, 1e-30), (err.norm() / max(ref.norm().item(), 1e-30)).item(),
                         (err[big] / ref.abs()[big]).max().item() if big.any() else 0.0))
        for k, a, b, c in rows:
            print("    %-44s max|err|/max|ref| %.2e   ||err||/||ref|| %.2e   max rel err (|ref|>=5%% max) %.2e" % (k, a, b, c))
