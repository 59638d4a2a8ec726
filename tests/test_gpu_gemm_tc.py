"""Tensor-core dense layers (csrc/gemm_tc.cu): bf16 operands / fp32 accumulate.  Reference = the same product computed in fp64 on
bf16-rounded operands (tolerance 1e-5 of the output scale: only the accumulation order differs), and the fp32 oracle layer at the
stated bf16 tolerance 2e-2."""
import pytest
import torch

from deepctr import ops
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bf(t):
    return t.to(torch.bfloat16).double()


@pytest.mark.parametrize("M,N,K,act", [(8192, 400, 429, "relu"), (300, 16, 40, None), (128, 256, 64, "tanh"), (1, 1, 1, None),
                                      (1000, 429, 400, None), (400, 429, 8192, None), (64, 600, 100, "sigmoid"), (129, 257, 65, "relu")])
def test_gemm_tc_matches_bf16_product(M, N, K, act):
    g = torch.Generator().manual_seed(M + N + K)
    A, Bm, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    ref = _bf(A) @ _bf(Bm).t() + bias.double()
    ref = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, None: lambda t: t}[act](ref)
    Ab, Bb = ops.cvt_bf16(A.to(DEV)), ops.cvt_bf16(Bm.to(DEV))
    assert Ab.shape == (M, (K + 7) // 8 * 8) and bool((Ab[:, K:] == 0).all())
    out = ops.gemm_tc(Ab, Bb, M, N, K, bias.to(DEV), ops.N.ACT[act])
    assert_close(out, ref, 1e-5, 1e-5 * max(ref.abs().max().item(), 1.0), "gemm_tc")


def test_cvt_bf16_transpose():
    g = torch.Generator().manual_seed(0)
    src = torch.randn(77, 45, generator=g)
    big = torch.randn(77, 60, generator=g)
    big[:, :45] = src
    t = ops.cvt_bf16(big.to(DEV)[:, :45], transpose=True)          # strided source (row pitch 60)
    assert t.shape == (45, 80)
    assert torch.equal(t[:, :77].cpu(), src.t().to(torch.bfloat16)) and bool((t[:, 77:] == 0).all())


@pytest.mark.parametrize("B,K,Nn,act", [(513, 429, 400, "relu"), (64, 40, 24, None)])
def test_linear_act_bf16_forward_backward(B, K, Nn, act):
    g = torch.Generator().manual_seed(B)
    x, W, b, dy = torch.randn(B, K, generator=g), torch.randn(Nn, K, generator=g) / K ** 0.5, torch.randn(Nn, generator=g), \
        torch.randn(B, Nn, generator=g)
    xg, Wg, bg = x.to(DEV).requires_grad_(True), W.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    out = ops.linear_act(xg, Wg, bg, act, precision="bf16")
    out.backward(dy.to(DEV))
    xd, Wd, bd = x.double().requires_grad_(True), W.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = xd @ Wd.t() + bd
    if act == "relu":
        # the gradient reference uses the kernel's own ReLU active set: pre-activations within bf16 rounding of zero may
        # legitimately land on either side, and a flipped unit changes whole rows of dx
        mask = (out.detach().cpu() > 0).double()
        assert ((ref.detach() > 0).double() != mask).double().mean().item() < 2e-2
        ref = ref * mask
    ref.backward(dy.double())
    tol = 2e-2          # stated bf16 tolerance (operands rounded to 8 mantissa bits), relative to each tensor's scale
    for got, want, what in ((out, ref, "y"), (xg.grad, xd.grad, "dx"), (Wg.grad, Wd.grad, "dW"), (bg.grad, bd.grad, "db")):
        assert_close(got, want, tol, tol * want.abs().max().item(), what)


@pytest.mark.parametrize("act", ["linear", "relu", "tanh", "sigmoid"])
@pytest.mark.parametrize("R,C", [(8192, 400), (100, 37), (64, 64), (1, 8), (333, 429), (4096, 1)])
def test_cvt_bf16_both_one_pass_for_every_operand(R, C, act):
    """xdfm_cvt_bf16_both: g = dy * act'(y) in one pass -> bf16 rows, bf16 transposed copy, fp32 column sums (the bias gradient);
    compared with the separate act_bwd / cvt_bf16 / wcolsum results it replaces."""
    from deepctr import ops, _native as Nv
    g = torch.Generator().manual_seed(R * 7 + C)
    dy = torch.randn(R, C, generator=g).to(DEV)
    y = torch.randn(R, C, generator=g).to(DEV)
    if act == "sigmoid":
        y = torch.sigmoid(y)
    a = Nv.ACT[act]
    rows, cols, colsum = ops.cvt_bf16_both(dy, True, True, y=None if a == 0 else y, act=a, want_colsum=True)
    torch.cuda.synchronize()
    if act == "linear":
        want = dy
    elif act == "relu":
        want = torch.where(y > 0, dy, torch.zeros_like(dy))
    elif act == "tanh":
        want = dy * (1 - y * y)
    else:
        want = dy * y * (1 - y)
    C8, R8 = (C + 7) // 8 * 8, (R + 7) // 8 * 8
    assert rows.shape == (R, C8) and cols.shape == (C, R8)
    if act in ("linear", "relu"):          # the same fp32 value rounded once: bit-exact
        assert torch.equal(rows[:, :C], want.to(torch.bfloat16))
    else:                                  # the kernel may contract dy * (1 - y * y) differently: one bf16 ulp
        assert_close(rows[:, :C].float(), want, 2.0 ** -7, 1e-30, "rows")
    assert torch.equal(cols[:, :R], rows[:, :C].t())          # both copies hold the same bf16 values
    assert bool((rows[:, C:] == 0).all()) and bool((cols[:, R:] == 0).all())
    assert_close(colsum, want.double().sum(0), 1e-5, 1e-5 * max(1.0, float(want.abs().sum(0).max())), "colsum")
    # no outputs requested individually
    r2, c2, s2 = ops.cvt_bf16_both(dy, True, False)
    assert c2 is None and s2 is None and torch.equal(r2[:, :C], dy.to(torch.bfloat16))


@pytest.mark.parametrize("B,fm,hd,binary", [(8192, 400, 400, 1), (37, 5, 0, 1), (100, 0, 33, 0), (16, 300, 7, 1), (1, 1, 1, 1)])
def test_head_backward_in_one_pass(B, fm, hd, binary):
    """xdfm_head_bwd_fused == xdfm_head_bwd + the three weighted column sums (autograd of xdeepfm.py:88-105), against float64."""
    from deepctr import _native as Nv
    L = Nv.lib()
    g = torch.Generator().manual_seed(B + fm + hd)
    dy = torch.randn(B, generator=g).to(DEV)
    y = torch.rand(B, generator=g).to(DEV)
    cin = torch.randn(B, fm, generator=g).to(DEV) if fm else None
    dnn = torch.randn(B, hd, generator=g).to(DEV) if hd else None
    wc = torch.randn(fm, generator=g).to(DEV) if fm else None
    wd = torch.randn(hd, generator=g).to(DEV) if hd else None
    dlogit = torch.full((B,), float("nan"), device=DEV)
    d_cin = torch.full((B, fm), float("nan"), device=DEV) if fm else None
    d_dnn = torch.full((B, hd), float("nan"), device=DEV) if hd else None
    d_w = torch.full((fm + hd + 1,), float("nan"), device=DEV)
    ws = torch.empty(L.xdfm_head_bwd_fused_workspace_bytes(B, fm, hd), dtype=torch.uint8, device=DEV)
    for _ in range(2):
        Nv.check(L.xdfm_head_bwd_fused(Nv.ptr(dy), Nv.ptr(y), B, binary, Nv.ptr(cin), Nv.ptr(wc), fm, Nv.ptr(dnn), Nv.ptr(wd), hd,
                                       Nv.ptr(dlogit), Nv.ptr(d_cin), Nv.ptr(d_dnn), Nv.ptr(d_w), Nv.ptr(ws), ws.numel(), Nv.stream_ptr()))
        torch.cuda.synchronize()
        first = d_w.clone() if _ == 0 else first
    assert torch.equal(first, d_w)                      # fixed summation order: bit-reproducible
    gl = dy.double() * (y.double() * (1 - y.double())) if binary else dy.double()
    assert_close(dlogit, gl, 1e-6, 1e-7, "dlogit")
    if fm:
        assert_close(d_cin, gl[:, None] * wc.double()[None, :], 1e-6, 1e-7, "d_cin")
        assert_close(d_w[:fm], (gl[:, None] * cin.double()).sum(0), 1e-5, 1e-5 * (B ** 0.5), "d_w_cin")
    if hd:
        assert_close(d_dnn, gl[:, None] * wd.double()[None, :], 1e-6, 1e-7, "d_dnn")
        assert_close(d_w[fm:fm + hd], (gl[:, None] * dnn.double()).sum(0), 1e-5, 1e-5 * (B ** 0.5), "d_w_dnn")
    assert_close(d_w[fm + hd:], gl.sum().reshape(1), 1e-5, 1e-5 * (B ** 0.5), "d_bias")


def test_parameter_gradients_written_in_place_equal_autograd_accumulation():
    """ops.direct_param_grads (fused training step): the first gradient of a parameter in a backward pass is written straight into
    its zeroed p.grad view, later uses of the same parameter are accumulated by autograd -- same bits as the plain path."""
    g = torch.Generator().manual_seed(3)
    W = torch.nn.Parameter(torch.randn(24, 40, generator=g).to(DEV))
    b = torch.nn.Parameter(torch.randn(24, generator=g).to(DEV))
    x1, x2 = torch.randn(64, 40, generator=g).to(DEV), torch.randn(32, 40, generator=g).to(DEV)
    W.grad, b.grad = torch.zeros_like(W), torch.zeros_like(b)
    ptrs = (W.grad.data_ptr(), b.grad.data_ptr())

    def run(direct):
        W.grad.zero_()
        b.grad.zero_()
        ops.direct_param_grads(direct)
        try:
            out = ops.LinearActTC.apply(x1, W, b, 1).sum() + ops.LinearActTC.apply(x2, W, b, 1).square().sum()
            out.backward()
        finally:
            ops.direct_param_grads(False)
        assert (W.grad.data_ptr(), b.grad.data_ptr()) == ptrs
        return W.grad.clone(), b.grad.clone()

    plain, direct = run(False), run(True)
    assert float(plain[0].abs().max()) > 0 and float(plain[1].abs().max()) > 0
    assert torch.equal(plain[0], direct[0]) and torch.equal(plain[1], direct[1])
