"""TEST INFRASTRUCTURE ONLY -- the oracle's xDeepFM train step with the ROUNDING POINTS of the bf16 tensor-core configuration.

`cin.precision = dnn.precision = 'bf16'` (what bench.py measures) keeps the reference's mathematics (oracle/xdeepfm_oracle.py,
which follows deepctr/layers/interaction.py:207-248, core.py:120-134, xdeepfm.py:79-107) and changes only where values are rounded
to bfloat16 before they enter a tensor-core GEMM.  This module restates those rounding points on top of the oracle's functions so
that a model-level test can separate two questions:

  (1) do the CUDA kernels compute exactly the arithmetic they claim?   GPU vs this emulation: tight, per element.
  (2) how far does that arithmetic sit from the fp32 reference?         emulation (and GPU) vs the reference fixtures: the
      intrinsic cost of 8-bit significands, amplified where a gradient is a sum of cancelling per-sample terms.

Rounding points (file:line of the product code they mirror):
  CIN layer forward   x0, X^{k-1} stored as bf16 rows (ops.py CINFunctionTC.forward: xdfm_to_rows_bf16; the epilogue writes the next
                      layer's input as bf16); the outer product Z = X^{k-1} * X^0 is rounded to bf16 when it is generated
                      (csrc/cin_tc.cu producers: __hmul2); W rounded to bf16 (cin_prep_w); accumulation, bias, activation and the
                      pooled / direct outputs in fp32 (from the TMEM accumulators).
  CIN layer backward  dY = act'(y) * (d direct + d next) rounded to bf16 (cin_dy_rows_cols); dW = dY^T Z with Z regenerated in bf16;
                      db from the bf16 dY; dZ = dY W in fp32, contracted with the bf16 X^0 / X^{k-1} in fp32 (cin_tc_bwd_dx.cu).
  DNN layer           x, W rounded to bf16 (ops.py LinearActTC: cvt_bf16), fp32 accumulate + bias + activation; backward: dY*act'
                      rounded to bf16 for dX and dW, db from the fp32 values.  Layers with K, N <= 32 run exact fp32 (SmallLinear).
Everything else (gather, first-order term, heads, sigmoid, BCE, L2) is fp32 in both configurations.
Accumulations here run in float64 (the kernels accumulate in fp32: the difference is ~1e-7, far below one bf16 ulp).
"""
import torch
import torch.nn.functional as F

from . import xdeepfm_oracle as O


def bf16(t):
    """Round to nearest-even bfloat16, keep the container dtype."""
    return t.to(torch.bfloat16).to(t.dtype)


class _CinLayer(torch.autograd.Function):
    """(y, yb) = layer(x0b, xkb, W, b): y = act(W' Z + b) unrounded (direct / pooled outputs), yb = bf16(y) (next layer's input)."""

    @staticmethod
    def forward(ctx, x0b, xkb, W, b, act):
        B, m, D = x0b.shape
        Hp = xkb.shape[1]
        Z = bf16(xkb[:, :, None, :] * x0b[:, None, :, :]).reshape(B, Hp * m, D)
        Wb = bf16(W.reshape(W.shape[0], -1))
        pre = torch.einsum("hk,bkd->bhd", Wb, Z) + b.view(1, -1, 1)
        y = torch.relu(pre) if act == "relu" else pre
        yb = bf16(y)
        ctx.save_for_backward(x0b, xkb, Wb, Z, yb)
        ctx.act, ctx.wshape = act, W.shape
        return y, yb

    @staticmethod
    def backward(ctx, dy, dyb_in):
        x0b, xkb, Wb, Z, yb = ctx.saved_tensors
        B, m, D = x0b.shape
        Hp = xkb.shape[1]
        g = dy + dyb_in
        if ctx.act == "relu":
            g = g * (yb > 0).to(g.dtype)
        gb = bf16(g)
        dW = torch.einsum("bhd,bkd->hk", gb, Z).reshape(ctx.wshape)
        db = gb.sum(dim=(0, 2))
        dZ = torch.einsum("bhd,hk->bkd", gb, Wb).reshape(B, Hp, m, D)
        dxk = (dZ * x0b[:, None, :, :]).sum(2)
        dx0 = (dZ * xkb[:, :, None, :]).sum(1)
        return dx0, dxk, dW, db, None


class _RoundSTE(torch.autograd.Function):
    """bf16 rounding with a straight-through gradient (the fp32 gradient of the rounded copy is handed to the fp32 source)."""

    @staticmethod
    def forward(ctx, x):
        return bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


def cin_forward_bf16(x0, weights, biases, split_half=True, activation="relu", pool=True):
    if activation not in ("relu", "linear", None):
        raise NotImplementedError("emulation covers relu / linear CIN activations")
    act = "relu" if activation == "relu" else "linear"
    x0b = _RoundSTE.apply(x0)
    hidden = x0b
    finals, n = [], len(weights)
    for k, (W, b) in enumerate(zip(weights, biases)):
        y, yb = _CinLayer.apply(x0b, hidden, W, b, act)
        H = y.shape[1]
        if split_half:
            if k != n - 1:
                hidden, direct = yb[:, : H // 2], y[:, H // 2:]
            else:
                direct, hidden = y, None
        else:
            direct, hidden = y, yb
        finals.append(direct)
    maps = torch.cat(finals, dim=1)
    return maps.sum(-1) if pool else maps


class _DenseLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, act):
        xb, Wb = bf16(x), bf16(W)
        pre = xb.matmul(Wb.t()) + b
        y = torch.relu(pre) if act == "relu" else pre
        ctx.save_for_backward(xb, Wb, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, Wb, y = ctx.saved_tensors
        g = dy * (y > 0).to(dy.dtype) if ctx.act == "relu" else dy
        gb = bf16(g)
        return gb.matmul(Wb), gb.t().matmul(xb), g.sum(0), None


def dnn_forward_bf16(x, weights, biases, activation="relu"):
    if activation not in ("relu", "linear", None):
        raise NotImplementedError("emulation covers relu / linear DNN activations")
    act = "relu" if activation == "relu" else "linear"
    for W, b in zip(weights, biases):
        if W.shape[0] <= 32 and W.shape[1] <= 32:       # ops.small_linear_ok: narrow layers stream in exact fp32
            pre = x.matmul(W.t()) + b
            x = torch.relu(pre) if act == "relu" else pre
        else:
            x = _DenseLayer.apply(x, W, b, act)
    return x


def logit_bf16(params, spec, X):
    """oracle.xdeepfm_logit with the CIN and the DNN replaced by their bf16-rounded forms (xDeepFM and the attention variants)."""
    if spec.variant == "pro":
        raise NotImplementedError("the emulation covers xDeepFM and the attention variants")
    ids, dense = O.split_input(spec, X)
    dt = params["out.bias"].dtype
    dense = dense.to(dt)
    emb = O.embedding_lookup(params, spec, ids)
    logit = O.linear_logit(params, spec, ids, dense)
    if len(spec.cin_layer_size) > 0:
        Ws, bs = O.cin_params(params, spec)
        if spec.variant == "xdeepfm":
            cin_out = cin_forward_bf16(emb, Ws, bs, spec.cin_split_half, spec.cin_activation, pool=True)
        else:
            cin_out = O.cin_attention_tail(params, spec, cin_forward_bf16(emb, Ws, bs, spec.cin_split_half, spec.cin_activation, pool=False))
        logit = logit + cin_out.matmul(params["cin_linear.weight"].t())
    if len(spec.dnn_hidden_units) > 0:
        dnn_in = torch.cat([emb.reshape(emb.shape[0], -1), dense], dim=-1) if spec.nd > 0 else emb.reshape(emb.shape[0], -1)
        Ws, bs = O.dnn_params(params, spec)
        dnn_out = dnn_forward_bf16(dnn_in, Ws, bs, spec.dnn_activation)
        logit = logit + dnn_out.matmul(params["dnn_linear.weight"].t())
    return logit


def loss_and_grads_bf16(params, spec, X, y):
    """(y_pred, BCE-sum, total, grads by name) of one train step in the emulated bf16 configuration; float64 containers."""
    leaves = {k: v.detach().double().clone().requires_grad_(True) for k, v in params.items()}
    Xd = X.double()
    y_pred = torch.sigmoid(logit_bf16(leaves, spec, Xd) + leaves["out.bias"])
    loss = F.binary_cross_entropy(y_pred.squeeze(-1), y.double().reshape(-1), reduction="sum")
    total = loss + O.reg_loss(leaves, spec).squeeze()
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return y_pred.detach(), loss.detach(), total.detach(), grads
