"""GPU parity of the CIN operator (forward + backward) against the CPU oracle and the reference's known answers."""
import json
import os

import pytest
import torch

from oracle import xdeepfm_oracle as O
from tests.helpers import GOLDEN, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run_product(x0, Ws, bs, split_half, act, pool, gout):
    from deepctr import ops
    m = x0.shape[1]
    cfg = ops.CINConfig(m, [W.shape[0] for W in Ws], split_half, act, pool=pool)
    x = x0.to(DEV).requires_grad_(True)
    wb = []
    for W, b in zip(Ws, bs):
        wb += [W.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)]
    out = ops.CINFunction.apply(cfg, x, *wb)
    out.backward(gout.to(DEV))
    return out.detach(), x.grad, [t.grad for t in wb]


def _product_masks(x0, Ws, bs, split_half, act):
    """ReLU active-sets chosen by the product kernels (layer outputs y > 0), via the C ABI."""
    from deepctr import _native as N
    from deepctr import ops
    B, m, D = x0.shape
    cfg = ops.CINConfig(m, [W.shape[0] for W in Ws], split_half, act, pool=True)
    x = x0.to(DEV).contiguous()
    out = torch.empty(B, cfg.fm, device=DEV)
    xk, stride, masks = x, m * D, []
    for k, H in enumerate(cfg.layer_size):
        y = torch.empty(B, H, D, device=DEV)
        Wd, bd = Ws[k].to(DEV).reshape(H, -1).contiguous(), bs[k].to(DEV)     # keep alive across the launch
        N.check(N.lib().xdfm_cin_fwd_f32(N.ptr(x), N.ptr(xk), stride, N.ptr(Wd), N.ptr(bd), B, m, cfg.Hp[k], H, D, cfg.act,
                                         N.ptr(y), cfg.direct_begin[k],
                                         N.ptr(out), None, cfg.fm, cfg.col_off[k], N.stream_ptr()))
        masks.append((y > 0).cpu())
        xk, stride = y, H * D
    return masks


def _run_oracle(x0, Ws, bs, split_half, act, pool, gout, masks=None):
    x = x0.double().requires_grad_(True)
    Wd = [W.double().requires_grad_(True) for W in Ws]
    bd = [b.double().requires_grad_(True) for b in bs]
    out = O.cin_forward(x, Wd, bd, split_half, act, pool=pool, relu_masks=masks)
    out.backward(gout.double())
    g = []
    for W, b in zip(Wd, bd):
        g += [W.grad, b.grad]
    return out.detach(), x.grad, g


def _rand_case(B, m, D, layers, split_half, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, m, D, generator=g) * 0.5
    Ws, bs = [], []
    prev = m
    for H in layers:
        K = prev * m
        Ws.append(torch.randn(H, K, 1, generator=g) / K ** 0.5)
        bs.append(torch.randn(H, generator=g) * 0.1)
        prev = H // 2 if split_half else H
    return x0, Ws, bs, g


CASES = [
    # B, m, D, layers, split_half, act, pool
    (5, 3, 2, (4, 2), True, "relu", True),
    (33, 5, 8, (16, 8), True, "relu", True),
    (17, 5, 8, (6, 5, 4), False, "relu", True),
    (9, 4, 10, (8, 6, 3), True, "linear", True),      # D not a power of two, odd last layer
    (12, 5, 8, (16, 8), True, "relu", False),         # un-pooled maps (attention variant input)
    (64, 26, 8, (256, 128), True, "relu", True),      # BASELINE config 1 shape
    (40, 26, 16, (200, 200, 200), True, "relu", True),  # BASELINE config 2 shape
    (16, 22, 32, (256, 256), True, "relu", True),     # config 4 shape (2 of 4 layers)
    (6, 26, 64, (256, 128), True, "relu", True),      # config 5 shape
]


@pytest.mark.parametrize("case", CASES, ids=[str(c[:5]) for c in CASES])
def test_cin_fp32_matches_oracle(case):
    B, m, D, layers, split_half, act, pool = case
    x0, Ws, bs, g = _rand_case(B, m, D, layers, split_half, seed=B + m + D)
    # Pre-activations that sit on the ReLU kink (|pre| ~ rounding noise) may get different gradient masks in fp32 and
    # fp64 -- a property of ReLU, not an error.  The oracle therefore differentiates with the active-set the kernel chose,
    # after checking that this active-set disagrees with the fp64 one only where |pre| is rounding noise.
    masks = None
    if act == "relu":
        masks = _product_masks(x0, Ws, bs, split_half, act)
        pres = O.cin_preactivations(x0.double(), [W.double() for W in Ws], [b.double() for b in bs], split_half, act)
        for mk, pre in zip(masks, pres):
            flipped = mk != (pre > 0)
            assert flipped.double().mean().item() < 1e-3 and (not flipped.any() or pre[flipped].abs().max().item() < 1e-4)
    fm = (sum(layers[:-1]) // 2 + layers[-1]) if split_half else sum(layers)
    gout = torch.randn((B, fm) if pool else (B, fm, D), generator=g)
    out, dx, dwb = _run_product(x0, Ws, bs, split_half, act, pool, gout)
    ro, rdx, rdwb = _run_oracle(x0, Ws, bs, split_half, act, pool, gout, masks)
    # fp32 CUDA-core path vs fp64 oracle: 1e-4 relative to the tensor's scale (K up to 3328 fp32 accumulations)
    assert_close(out, ro, 1e-4, 1e-5 * max(ro.abs().max().item(), 1.0), "cin out")
    assert_close(dx, rdx, 1e-4, 1e-4 * rdx.abs().max().item(), "cin dx0")
    for i, (a, b) in enumerate(zip(dwb, rdwb)):
        assert_close(a, b.reshape(a.shape), 1e-4, 1e-4 * b.abs().max().item(), "cin grad %d" % i)


def test_cin_known_answer_forward():
    """Reference-recorded KAT (tests/golden/cin_kat.json).  Only the forward is pinned here: one layer-0 pre-activation
    is exactly 0 in decimal arithmetic, so the ReLU gradient of the KAT depends on summation order (see test_oracle)."""
    kat = json.load(open(os.path.join(GOLDEN, "cin_kat.json")))
    K = [9, 6]
    Ws = [torch.tensor([[(((h * K[l] + k) * 3) % 7 - 2) / 10 for k in range(K[l])] for h in range(H)], dtype=torch.float32).unsqueeze(-1)
          for l, H in enumerate((4, 2))]
    bs = [torch.tensor([(h + 1) / 10 for h in range(H)], dtype=torch.float32) for H in (4, 2)]
    x = torch.tensor(kat["x"], dtype=torch.float32).reshape(2, 3, 2)
    out, dx, dwb = _run_product(x, Ws, bs, True, "relu", True, torch.ones(2, 4))
    assert_close(out, torch.tensor(kat["out"]), 1e-6, 1e-6, "KAT out")
    assert_close(dwb[3], torch.tensor(kat["db1"]), 1e-6, 1e-6, "KAT db1")
    assert_close(dwb[2].flatten(), torch.tensor(kat["dW1"]), 1e-5, 1e-6, "KAT dW1")


def test_cin_rejects_bad_rank():
    from deepctr.layers import CIN
    with pytest.raises(ValueError):
        CIN(3, (4, 2), device=DEV)(torch.zeros(2, 6, device=DEV))
    with pytest.raises(ValueError):
        CIN(3, (3, 2), split_half=True)
    with pytest.raises(ValueError):
        CIN(3, ())


TC_MODEL_CASES = [
    (33, 5, 8, (16, 8), True, "relu", True),
    (17, 5, 8, (6, 5, 4), False, "relu", True),
    (12, 5, 8, (16, 8), True, "relu", False),
    (64, 26, 8, (256, 128), True, "relu", True),
    (40, 26, 16, (200, 200, 200), True, "relu", True),
    (16, 22, 32, (256, 256), True, "relu", True),
    (9, 6, 16, (8, 6, 3), True, "linear", True),
]


@pytest.mark.parametrize("case", TC_MODEL_CASES, ids=[str(c[:5]) for c in TC_MODEL_CASES])
def test_cin_bf16_tensor_core_path_matches_oracle(case):
    """bf16 tensor-core CIN (forward + backward through autograd) vs the fp64 oracle on bf16-rounded x0 / W.
    Tolerance: bf16 products and bf16 layer outputs -> 2e-2 of the tensor's scale (stated bf16 tolerance)."""
    from deepctr import ops
    B, m, D, layers, split_half, act, pool = case
    x0, Ws, bs, g = _rand_case(B, m, D, layers, split_half, seed=B + m + D + 1)
    x0 = x0.to(torch.bfloat16).float()
    Ws = [W.to(torch.bfloat16).float() for W in Ws]
    fm = (sum(layers[:-1]) // 2 + layers[-1]) if split_half else sum(layers)
    gout = torch.randn((B, fm) if pool else (B, fm, D), generator=g)
    cfg = ops.CINConfig(m, [W.shape[0] for W in Ws], split_half, act, pool=pool, impl="bf16")
    x = x0.to(DEV).requires_grad_(True)
    wb = []
    for W, b in zip(Ws, bs):
        wb += [W.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)]
    out = ops.cin_apply(cfg, x, *wb)
    out.backward(gout.to(DEV))
    masks = None
    if act == "relu":
        # active-set chosen by the bf16 kernels (from the saved bf16 layer outputs)
        ctx_masks = []
        xk_ref = None
        pres = O.cin_preactivations(x0.double(), [W.double() for W in Ws], [b.double() for b in bs], split_half, act)
        # recompute the kernel's outputs layer by layer through the C ABI to get its masks
        from tests.test_gpu_tc import tc_layer, to_rows
        x0t = to_rows(x0.to(DEV), (m + 7) // 8 * 8)
        xkt = x0t
        for k, H in enumerate(cfg.layer_size):
            y, yt, _, _ = tc_layer(x0t, xkt, Ws[k].to(DEV).reshape(H, -1).contiguous(), bs[k].to(DEV), B, m, cfg.Hp[k], H, D, act,
                                   cfg.direct_begin[k], cfg.fm, cfg.col_off[k])
            ctx_masks.append((y.float() > 0).cpu())
            xkt = yt
        masks = ctx_masks
        for mk, pre in zip(masks, pres):
            flipped = mk != (pre > 0)
            assert flipped.double().mean().item() < 2e-2
    ro, rdx, rdwb = _run_oracle(x0, Ws, bs, split_half, act, pool, gout, masks)
    tol = 2e-2
    assert_close(out, ro, tol, tol * max(ro.abs().max().item(), 1.0), "cin out (bf16 path)")
    assert_close(x.grad, rdx, tol, tol * rdx.abs().max().item(), "cin dx0 (bf16 path)")
    for i, (a, b) in enumerate(zip([t.grad for t in wb], rdwb)):
        assert_close(a, b.reshape(a.shape), tol, tol * b.abs().max().item(), "cin grad %d (bf16 path)" % i)
