"""GPU bf16 configuration vs the bf16 emulation of the oracle (oracle/bf16_emulation.py): the per-element relative tolerance each
tensor would need (|err| / max(|ref|, 5 % of max |ref|)), and the same against the fp32 reference fixture.
    python tools/bf16_emulation_report.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "xdeepfm-pytorch_b200")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from oracle import bf16_emulation as E
from tests.helpers import build_product_model, golden_grads, load_case

DEV = "cuda:0"


def need(got, ref, floor=0.05):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    b = torch.clamp(ref.abs(), min=floor * max(ref.abs().max().item(), 1e-30))
    return ((got - ref).abs() / b).max().item(), ((got - ref).norm() / max(ref.norm().item(), 1e-30)).item()


for name in sys.argv[1:] or ["xdeepfm_small", "xdeepfm_small_nosplit", "xdeepfm_small_linearact", "xdeepfm_small_zipf", "xdeepfm_cfg1",
                             "xdeepfm_cfg2", "attn_small", "attn_v2_small"]:
    spec, params, z = load_case(name)
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    model.cin.precision = model.dnn.precision = "bf16"
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"])
    model.train()
    y_pred = model(X.to(DEV)).squeeze()
    loss = F.binary_cross_entropy(y_pred, y.to(DEV), reduction="sum")
    (loss + model.get_regularization_loss()).backward()
    yp_e, loss_e, _, grads_e = E.loss_and_grads_bf16(params, spec, X, y)
    print("%s: y_pred vs emulation max rel %.2e, loss rel %.2e" % (name, ((y_pred.cpu().double() - yp_e.squeeze()).abs() / yp_e.squeeze()).max().item(),
                                                                    abs(loss.item() - loss_e.item()) / loss_e.item()))
    worst = (0, 0, "")
    fixture = golden_grads(z)
    for k, p in model.named_parameters():
        a, b = need(p.grad, grads_e[k])
        line = "    %-44s vs emulation: per-element %.2e  norm %.2e" % (k, a, b)
        if k in fixture:
            c, d = need(p.grad, fixture[k])
            line += "   | vs fp32 reference: per-element %.2e  norm %.2e" % (c, d)
        print(line)
        worst = max(worst, (a, b, k))
    print("  worst vs emulation:", worst)
