"""CPU: the bf16 emulation of the oracle (oracle/bf16_emulation.py) against the fixtures of the unmodified fp32 reference.

Pins what bfloat16 operand rounding -- at exactly the points where the tcgen05 kernels round -- does to one reference train step,
without any GPU: the same distances tests/test_gpu_bf16_model.py measures for the CUDA path (which agrees with this emulation per
element), so they are a property of the precision BASELINE.json configs[1] names, not of the kernels."""
import pytest
import torch

from oracle import bf16_emulation as E
from oracle import xdeepfm_oracle as O
from tests.helpers import golden_grads, golden_gradsamples, load_case


@pytest.mark.parametrize("name,pred_rtol,norm_rtol,elem_atol", [
    # measured (worst tensor): norm-wise 4.0e-2 / 7.7e-2 / 3.1e-2 / 1.18e-1 / 3.3e-2 / 5.2e-3; per element of the max 9e-2 / 1.6e-1 /
    # 4.8e-2 / 1.85e-1 / 6.2e-2 / 7e-3.  The last case has a LINEAR CIN activation: without ReLU masks that can flip on
    # near-zero pre-activations the whole step stays within 0.5 % -- the large entries above are mask flips, not accumulated rounding
    ("xdeepfm_small", 5e-3, 1.5e-1, 2.5e-1), ("xdeepfm_small_zipf", 5e-3, 1.5e-1, 2.5e-1), ("xdeepfm_cfg1", 5e-3, 1.5e-1, 2.5e-1),
    ("xdeepfm_cfg2", 5e-3, 1.5e-1, 2.5e-1), ("attn_small", 2e-2, 1.5e-1, 2.5e-1), ("xdeepfm_small_linearact", 5e-3, 1e-2, 1.5e-2)])
def test_emulated_bf16_step_stays_within_the_stated_distance_of_the_fp32_reference(name, pred_rtol, norm_rtol, elem_atol):
    spec, params, z = load_case(name)
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"])
    y_pred, loss, _, grads = E.loss_and_grads_bf16(params, spec, X, y)
    ref = torch.from_numpy(z["y_pred"]).double()
    assert ((y_pred.squeeze(-1) - ref).abs() / ref).max().item() <= pred_rtol
    assert abs(loss.item() - float(z["loss"])) <= 5e-3 * float(z["loss"])
    pairs = [(k, grads[k], g.double()) for k, g in golden_grads(z).items()]
    pairs += [(k, grads[k].flatten()[::s], g.double()) for k, (s, g) in golden_gradsamples(z).items()]
    worst = 0.0
    for k, got, g in pairs:
        rel = ((got - g).norm() / max(g.norm().item(), 1e-30)).item()
        worst = max(worst, rel)
        assert rel <= norm_rtol, "%s: ||err|| / ||ref|| = %.3e" % (k, rel)
        assert (got - g).abs().max().item() <= elem_atol * g.abs().max().item() + 1e-30, k
    assert worst >= 1e-4, "bf16 rounding must be visible (is the emulation rounding at all?)"


def test_emulation_without_wide_layers_is_the_oracle():
    """Layers with K, N <= 32 stream in exact fp32 in the product (ops.small_linear_ok); the emulation's rounding hooks must be the
    only difference from the oracle: with rounding disabled it reproduces the oracle's gradients to float64 accuracy."""
    spec, params, z = load_case("xdeepfm_small")
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"])
    saved = E.bf16
    E.bf16 = lambda t: t
    try:
        y_e, loss_e, _, g_e = E.loss_and_grads_bf16(params, spec, X, y)
    finally:
        E.bf16 = saved
    p64 = {k: v.double() for k, v in params.items()}
    y_o, loss_o, _, g_o = O.loss_and_grads(p64, spec, X.double(), y.double())
    assert torch.allclose(y_e, y_o, rtol=1e-12, atol=1e-14)
    for k in g_o:
        assert torch.allclose(g_e[k], g_o[k], rtol=1e-9, atol=1e-12 * max(g_o[k].abs().max().item(), 1e-30)), k
