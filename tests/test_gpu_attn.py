"""Field self-attention block (csrc/attn.cu) against the oracle's restatement of deepctr/layers/cin_attention.py (fp64 on CPU).
Tolerances: fp32 kernels with exp2f/rsqrtf -> 2e-5 relative to each tensor's scale."""
import math

import pytest
import torch

from deepctr import ops
from oracle import xdeepfm_oracle as O
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, what, tol=2e-5):
    b = torch.as_tensor(b)
    assert_close(a, b, tol, tol * max(b.abs().max().item(), 1e-6), what)


@pytest.mark.parametrize("B,L,E,h", [(3, 256, 16, 4), (5, 24, 10, 2), (2, 40, 8, 4), (4, 17, 12, 1), (2, 300, 32, 2), (1, 8, 6, 6),
                                     (0, 16, 8, 2), (3, 333, 8, 2), (2, 5, 4, 1), (4, 1030, 16, 4)])
@pytest.mark.parametrize("row_blocked", [1, 0])
def test_mhsa_core_matches_softmax_attention(B, L, E, h, row_blocked):
    from deepctr import _native as Nv
    Nv.lib().xdfm_mhsa_set_row_blocked(row_blocked)
    try:
        _mhsa_core_case(B, L, E, h)
    finally:
        Nv.lib().xdfm_mhsa_set_row_blocked(1)


def _mhsa_core_case(B, L, E, h):
    g = torch.Generator().manual_seed(B * 1000 + L + E)
    q, k, v, do = (torch.randn(B, L, E, generator=g) for _ in range(4))
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    hd = E // h
    split = lambda t: t.view(B, L, h, hd).transpose(1, 2)
    p = torch.softmax(split(qd).matmul(split(kd).transpose(-2, -1)) / math.sqrt(hd), dim=-1)
    ref = p.matmul(split(vd)).transpose(1, 2).contiguous().view(B, L, E)
    ref.backward(do.double())
    qg, kg, vg = (t.to(DEV).requires_grad_(True) for t in (q, k, v))
    out = ops.MHSACore.apply(qg, kg, vg, h)
    out.backward(do.to(DEV))
    if B == 0:
        assert out.shape == (0, L, E)
        return
    _close(out, ref, "attention output")
    _close(qg.grad, qd.grad, "dq")
    _close(kg.grad, kd.grad, "dk")
    _close(vg.grad, vd.grad, "dv")


@pytest.mark.parametrize("rows_shape,E,residual,normalize", [((4, 33), 16, True, True), ((1000,), 10, False, True), ((7, 5), 8, True, False),
                                                               ((3, 300), 64, True, True), ((513,), 32, True, True)])
def test_add_layer_norm(rows_shape, E, residual, normalize):
    g = torch.Generator().manual_seed(E)
    a = torch.randn(*rows_shape, E, generator=g)
    r = torch.randn(*rows_shape, E, generator=g) if residual else None
    gamma, beta = 1 + 0.1 * torch.randn(E, generator=g), 0.1 * torch.randn(E, generator=g)
    dy = torch.randn(*rows_shape, E, generator=g)
    ad, gd, bd = a.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rd = r.double().requires_grad_(True) if residual else None
    x = ad + rd if residual else ad
    ref = torch.nn.functional.layer_norm(x, (E,), gd, bd, 1e-5) if normalize else x
    ref.backward(dy.double())
    ag, gg, bg = a.to(DEV).requires_grad_(True), gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    rg = r.to(DEV).requires_grad_(True) if residual else None
    out = ops.AddLayerNorm.apply(ag, rg, gg if normalize else None, bg if normalize else None, 1e-5, normalize)
    out.backward(dy.to(DEV))
    _close(out, ref, "layer norm output")
    _close(ag.grad, ad.grad, "d a")
    if residual:
        _close(rg.grad, rd.grad, "d residual")
    if normalize:
        _close(gg.grad, gd.grad, "d gamma", 1e-4)
        _close(bg.grad, bd.grad, "d beta", 1e-4)


@pytest.mark.parametrize("B,L,E", [(3, 256, 16), (5, 24, 10), (2, 300, 64), (1, 1, 8)])
def test_attention_pooling(B, L, E):
    g = torch.Generator().manual_seed(L)
    score, x, dout = 2 * torch.randn(B, L, 1, generator=g), torch.randn(B, L, E, generator=g), torch.randn(B, E, generator=g)
    sd, xd = score.double().requires_grad_(True), x.double().requires_grad_(True)
    ref = (torch.softmax(sd, dim=1) * xd).sum(dim=1)
    ref.backward(dout.double())
    sg, xg = score.to(DEV).requires_grad_(True), x.to(DEV).requires_grad_(True)
    out = ops.AttnPool.apply(sg, xg)
    out.backward(dout.to(DEV))
    _close(out, ref, "pooled")
    _close(sg.grad, sd.grad, "d score")
    _close(xg.grad, xd.grad, "d x")


@pytest.mark.parametrize("variant,heads,layers,ln,res", [("attn", 4, 1, True, True), ("attn", 3, 1, False, True), ("attn", 2, 1, True, False),
                                                        ("attn_v2", 2, 2, True, True), ("attn_v2", 4, 1, False, False)])
def test_cin_attention_tail_matches_oracle(variant, heads, layers, ln, res):
    """CINAttention / CINAttentionV2 modules (maps -> MHSA -> residual/LN -> pooling [-> projection]) with oracle weights."""
    from deepctr.layers.cin_attention import CINAttention, CINAttentionV2
    m, E, sizes, B = 5, 8, (16, 8), 9
    spec = O.ModelSpec(sparse_names=["C%d" % i for i in range(m)], vocab_sizes=[10] * m, embedding_dim=E, cin_layer_size=sizes,
                       variant=variant, num_heads=heads, use_layer_norm=ln, use_residual=res, num_attn_layers=layers)
    params = O.make_params(spec, seed=heads + layers)
    cin_sd = {k[len("cin."):]: v for k, v in params.items() if k.startswith("cin.")}
    if variant == "attn":
        mod = CINAttention(m, E, sizes, num_heads=heads, use_layer_norm=ln, use_residual=res, device=DEV)
    else:
        mod = CINAttentionV2(m, E, sizes, num_heads=heads, use_layer_norm=ln, use_residual=res, num_attn_layers=layers, device=DEV)
    mod.load_state_dict(cin_sd, strict=True)
    g = torch.Generator().manual_seed(3)
    x0 = 0.5 * torch.randn(B, m, E, generator=g)
    pd = {k: v.double().requires_grad_(True) for k, v in params.items()}
    Ws, bs = O.cin_params(pd, spec)
    xd = x0.double().requires_grad_(True)
    maps = O.cin_forward(xd, Ws, bs, True, "relu", pool=False)
    ref = O.cin_attention_tail(pd, spec, maps)
    gout = torch.randn(ref.shape, generator=g)
    ref.backward(gout.double())
    xg = x0.to(DEV).requires_grad_(True)
    out = mod(xg)
    out.backward(gout.to(DEV))
    _close(out, ref, "cin attention output", 5e-5)
    # fp32 kernels vs the fp64 oracle through ~10 chained ops with cancelling sums (sum_l dscore = 0): 3e-3 of each gradient's scale
    _close(xg.grad, xd.grad, "d x0", 3e-3)
    for name, p in mod.named_parameters():
        _close(p.grad, pd["cin." + name].grad, "grad " + name, 3e-3)


@pytest.mark.parametrize("R,K,N,nq,act,bias", [
    (4096 + 37, 16, 16, 3, None, False),      # the Q/K/V projections of BASELINE config 3 (E = 16), ragged last tile
    (1000, 16, 16, 1, "tanh", True),          # attention-pooling hidden layer
    (777, 16, 1, 1, None, False),             # attention-pooling score layer (N = 1)
    (300, 10, 10, 3, None, False),            # E = 10 (script default): not a multiple of 4 -> scalar load / store paths
    (129, 32, 32, 3, None, False), (5, 8, 6, 2, "relu", False), (1, 3, 5, 1, "sigmoid", True), (0, 16, 16, 3, None, False),
    (70000, 8, 8, 1, None, True)])
@pytest.mark.parametrize("staged", [0, 2, 4, 9])
def test_small_linear_matches_torch(R, K, N, nq, act, bias, staged):
    """Narrow-layer kernels (csrc/smalllin.cu) against fp64 torch: y, dx, every dW, db at 2e-5 of each tensor's scale; the
    two-stage reductions are bit-reproducible.  staged = rows per thread of the coalesced row kernel (0 = per-lane rows)."""
    from deepctr import _native as Nv0
    Nv0.lib().xdfm_small_linear_set_staged(staged)
    try:
        _small_linear_case(R, K, N, nq, act, bias)
    finally:
        Nv0.lib().xdfm_small_linear_set_staged(SL_STAGED_DEFAULT)


SL_STAGED_DEFAULT = 2


def _small_linear_case(R, K, N, nq, act, bias):
    g = torch.Generator().manual_seed(R + K * 7 + N)
    x = torch.randn(R, K, generator=g)
    Ws = [torch.randn(N, K, generator=g) / math.sqrt(K) for _ in range(nq)]
    b = 0.3 * torch.randn(N, generator=g) if bias else None
    dys = [torch.randn(R, N, generator=g) for _ in range(nq)]
    fn = {None: lambda t: t, "tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[act]
    xd = x.double().requires_grad_(True)
    Wd = [W.double().requires_grad_(True) for W in Ws]
    bd = b.double().requires_grad_(True) if bias else None
    refs = [fn(xd @ W.t() + (bd if bias else 0.0)) for W in Wd]
    torch.autograd.backward(refs, [d.double() for d in dys])

    def run():
        xg = x.to(DEV).requires_grad_(True)
        Wg = [W.to(DEV).requires_grad_(True) for W in Ws]
        bg = b.to(DEV).requires_grad_(True) if bias else None
        from deepctr import _native as Nv
        out = ops.SmallLinear.apply(xg, bg, Nv.ACT[act], *Wg)
        outs = list(out) if nq > 1 else [out]
        torch.autograd.backward(outs, [d.to(DEV) for d in dys])
        return outs, xg.grad, [W.grad for W in Wg], (bg.grad if bias else None)

    outs, dx, dWs, db = run()
    for q in range(nq):
        assert outs[q].shape == (R, N)
        if R == 0:
            assert float(dWs[q].abs().sum()) == 0.0
            continue
        _close(outs[q], refs[q].detach(), "y%d" % q)
        _close(dWs[q], Wd[q].grad, "dW%d" % q)
    if R == 0:
        return
    _close(dx, xd.grad, "dx")
    if bias:
        _close(db, bd.grad, "db")
    outs2, dx2, dWs2, db2 = run()
    assert all(torch.equal(a, b_) for a, b_ in zip(dWs, dWs2)) and torch.equal(dx, dx2)
    if bias:
        assert torch.equal(db, db2)


def test_small_linear_partial_outputs_and_limits():
    """Unused outputs of the fused Q/K/V pass get zero gradients; layers wider than 32 are refused by the C ABI (linear_act routes
    them to the GEMM kernels)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(50, 16, generator=g).to(DEV).requires_grad_(True)
    Ws = [torch.randn(16, 16, generator=g).to(DEV).requires_grad_(True) for _ in range(3)]
    q, k, v = ops.linear_multi(x, Ws)
    (q.sum() + 2 * v.sum()).backward()
    assert float(Ws[1].grad.abs().sum()) == 0.0 and float(Ws[0].grad.abs().sum()) > 0
    _close(Ws[2].grad, 2 * x.detach().sum(0, keepdim=True).expand(16, 16), "dWv")
    wide = torch.randn(48, 16, generator=g).to(DEV)
    with pytest.raises(RuntimeError):
        ops.SmallLinear.apply(x.detach(), None, 0, wide)
    y = ops.linear_act(x.detach(), wide)                       # falls through to the SGEMM path
    _close(y, x.detach().double().cpu() @ wide.double().cpu().t(), "wide layer")


# ---------------------------------------------------------------------------------------------------------------------------
# attention dropout (reference: nn.Dropout on the probabilities, cin_attention.py:54, 86)
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,L,E,heads,p", [(3, 37, 8, 4, 0.25), (2, 64, 16, 4, 0.5), (2, 21, 10, 2, 0.1), (1, 256, 16, 4, 0.3)])
def test_mhsa_core_with_dropout_matches_torch_under_the_same_mask(B, L, E, heads, p):
    """The fused core recomputes the keep mask from (seed, sample, head, query, key); xdfm_mhsa_dropout_mask materialises the same
    mask, with which plain torch reproduces forward and backward: o = (mask / (1 - p) * softmax(q k^T / sqrt(hd))) v."""
    from deepctr import _native as Nv
    from deepctr import ops
    g = torch.Generator().manual_seed(B * 1000 + L)
    q, k, v = (torch.randn(B, L, E, generator=g) for _ in range(3))
    dout = torch.randn(B, L, E, generator=g)
    ops.reset_dropout_state()
    st = ops.dropout_state(torch.device(DEV))
    seed = st.clone()
    mask = torch.empty((B, heads, L, L), dtype=torch.uint8, device=DEV)
    Nv.check(Nv.lib().xdfm_mhsa_dropout_mask(B, L, heads, p, Nv.ptr(seed), Nv.ptr(mask), Nv.stream_ptr()))
    keep = mask.double().cpu()
    frac = keep.mean().item()
    n = keep.numel()
    assert abs(frac - (1 - p)) < 5 * (p * (1 - p) / n) ** 0.5 + 1e-3, "kept fraction %.4f, expected %.4f" % (frac, 1 - p)
    qd, kd, vd = (t.to(DEV).requires_grad_(True) for t in (q, k, v))
    o = ops.MHSACore.apply(qd, kd, vd, heads, p)
    assert int(st.item()) != int(seed.item()), "the counter must advance"
    o.backward(dout.to(DEV))
    hd = E // heads
    qr, kr, vr = (t.double().requires_grad_(True) for t in (q, k, v))
    qh, kh, vh = (t.view(B, L, heads, hd).transpose(1, 2) for t in (qr, kr, vr))
    probs = torch.softmax(qh @ kh.transpose(-2, -1) / hd ** 0.5, dim=-1) * keep / (1 - p)
    ref = (probs @ vh).transpose(1, 2).reshape(B, L, E)
    ref.backward(dout.double())
    assert_close(o, ref, 2e-5, 2e-5 * ref.abs().max().item(), "o")
    for name, got, want in (("dq", qd.grad, qr.grad), ("dk", kd.grad, kr.grad), ("dv", vd.grad, vr.grad)):
        assert_close(got, want, 2e-4, 2e-5 * want.abs().max().item(), name)
    # a second call draws a different mask; eval-style p = 0 is the plain kernel
    o2 = ops.MHSACore.apply(qd, kd, vd, heads, p)
    assert not torch.equal(o2, o)


def test_attention_model_trains_with_attn_dropout_and_is_deterministic_in_eval():
    """xDeepFMAttention(cin_attn_dropout > 0): training steps run (the reference applies nn.Dropout there; this build raised before),
    the loss decreases, eval-mode predictions are dropout-free and reproducible, CUDA-graph replayed steps keep drawing new masks."""
    from deepctr import ops
    from deepctr.inputs import DenseFeat, SparseFeat
    from deepctr.models import xDeepFMAttention
    g = torch.Generator().manual_seed(3)
    vocab = [30, 11, 200, 7, 50]
    cols = [SparseFeat("C%d" % i, v, 8) for i, v in enumerate(vocab)] + [DenseFeat("I0", 1), DenseFeat("I1", 1)]
    model = xDeepFMAttention(cols, cols, dnn_hidden_units=(32, 16), cin_layer_size=(16, 8), cin_num_heads=4, cin_attn_dropout=0.2,
                             device=DEV, init_std=0.05)
    assert model.cin.mhsa.dropout_rate == 0.2
    model.compile("adam", "binary_crossentropy")
    B = 64
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in vocab], 1).to(torch.int32).to(DEV)
    dense = torch.rand(B, 2, generator=g).to(DEV)
    y = ((ids[:, 0] % 2 == 0) ^ (dense[:, 0] > 0.5)).float()
    model.train()
    losses, seeds = [], []
    for _ in range(12):
        accum = torch.zeros(1, dtype=torch.float64, device=DEV)
        model.train_step(ids, dense, y, accum)
        losses.append(accum.item())
        seeds.append(int(ops.dropout_state(torch.device(DEV)).item()))
    assert model._graphs, "the dropout step is CUDA-graph capturable (the seed lives on the device)"
    assert len(set(seeds)) == len(seeds), "every step (replayed ones included) advances the dropout counter"
    assert losses[-1] < losses[0]
    model.eval()
    with torch.no_grad():
        a, b = model.forward_ids(ids, dense), model.forward_ids(ids, dense)
    assert torch.equal(a, b)
