"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's xDeepFM hot path.

This is the parity oracle for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package never does.

What it restates (reference = Syclus123/xDeepFM-pytorch, vendored DeepCTR-Torch 0.2.9):

  split_input            deepctr/models/basemodel.py:368-370, 377-378  (ids = X[:, col].long(), dense = column views)
  embedding_lookup       deepctr/models/basemodel.py:354-380, deepctr/inputs.py:158-180
  linear_logit           deepctr/models/basemodel.py:63-92
  cin_forward            deepctr/layers/interaction.py:207-248
  dnn_forward            deepctr/layers/core.py:120-134, deepctr/inputs.py:126-138
  mhsa / attention_pool  deepctr/layers/cin_attention.py:63-97, 130-144
  cin_attention_tail     deepctr/layers/cin_attention.py:296-318 (v1), 452-466 (v2)
  xdeepfm_forward        deepctr/models/xdeepfm.py:79-107, deepctr/models/xdeepfm_attn.py:143-173, 269-301
  prediction             deepctr/layers/core.py:154-160
  reg_loss               deepctr/models/basemodel.py:412-428 (+ group registration :126-127, xdeepfm.py:57-60, 74-75)
  train_loss             deepctr/models/basemodel.py:254-257  (BCE on probabilities, reduction='sum')
  sfg_loss               deepctr/xdeepfm_pro/sfg_decoder.py:116-157, 198-204, 266-309; basemodel_sfg.py:420-476
  autodis_forward        deepctr/xdeepfm_pro/autodis.py:63-69, 100-127; xdeepfm_pro.py:130-153, 233-240

All arithmetic on this path lives in the third-party dependency torch (pinned torch==2.6.0 in the
reference's requirements.txt:60; this image has 2.11.0): the restatement is written with plain torch
CPU tensor ops (index, einsum, matmul, softmax, layer_norm) in float32 or float64, and gradients come
from torch autograd over those ops.  Parameters are passed as a dict keyed by the reference's
`state_dict()` names, so the same checkpoint drives reference, oracle and product.

PINNING: the reference ships no tests / golden vectors.  The oracle is pinned against (1) the RNG-free
CIN known-answer test recorded from the reference (`tests/golden/cin_kat.json`) and
(2) fixtures produced by running the unmodified reference in the build container
(`oracle/make_golden.py` -> `tests/golden/*.npz`).  See `tests/test_oracle.py`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class ModelSpec:
    """Shape/hyper-parameter description of one model instance (mirrors the ctor kwargs)."""
    sparse_names: List[str]
    vocab_sizes: List[int]
    embedding_dim: int
    dense_names: List[str] = field(default_factory=list)
    cin_layer_size: Tuple[int, ...] = (256, 128)
    cin_split_half: bool = True
    cin_activation: str = "relu"
    dnn_hidden_units: Tuple[int, ...] = (256, 256)
    dnn_activation: str = "relu"
    l2_reg_linear: float = 1e-5
    l2_reg_embedding: float = 1e-5
    l2_reg_dnn: float = 0.0
    l2_reg_cin: float = 0.0
    # attention variants (cin_attention.py)
    variant: str = "xdeepfm"        # xdeepfm | attn | attn_v2 | pro
    num_heads: int = 4
    use_layer_norm: bool = True
    use_residual: bool = True
    num_attn_layers: int = 1
    # xDeepFM Pro (xdeepfm_pro.py:57-87): SFG decoder; dropout is 0 in every parity case
    use_sfg: bool = True
    sfg_weight: float = 0.1
    sfg_hidden_units: Tuple[int, ...] = (128, 64)
    sfg_positive_only: bool = True
    sfg_use_label_attention: bool = True
    # AutoDis encoder of the dense features for the DNN branch (xdeepfm_pro.py:82-86, 130-153)
    use_autodis: bool = False
    autodis_buckets: int = 16
    autodis_temperature: float = 1.0
    # column order of the flat input matrix X: feature_index order (inputs.py:99-123).  The scripts
    # build sparse columns first, then dense (xdftrain.py:240-256).
    sparse_first: bool = True

    @property
    def m(self):
        return len(self.sparse_names)

    @property
    def nd(self):
        return len(self.dense_names)

    def sparse_cols(self):
        return list(range(0, self.m)) if self.sparse_first else list(range(self.nd, self.nd + self.m))

    def dense_cols(self):
        return list(range(self.m, self.m + self.nd)) if self.sparse_first else list(range(0, self.nd))

    @property
    def featuremap_num(self):
        ls = self.cin_layer_size
        return (sum(ls[:-1]) // 2 + ls[-1]) if self.cin_split_half else sum(ls)


def get_valid_num_heads(embed_dim, num_heads):
    """cin_attention.py:15-23 -- lower num_heads until it divides embed_dim."""
    if embed_dim % num_heads == 0:
        return num_heads
    for h in range(num_heads, 0, -1):
        if embed_dim % h == 0:
            return h
    return 1


def split_input(spec: ModelSpec, X: torch.Tensor):
    """ids = X[:, col].long() (truncation), dense = float columns (basemodel.py:368-370, 377-378)."""
    ids = X[:, spec.sparse_cols()].long()
    dense = X[:, spec.dense_cols()]
    return ids, dense


def embedding_lookup(params: Dict[str, torch.Tensor], spec: ModelSpec, ids: torch.Tensor, prefix="embedding_dict."):
    """m independent gathers weight_f[ids[:, f]] -> [B, m, D] (cat on dim=1, xdeepfm.py:86)."""
    rows = [params[prefix + name + ".weight"][ids[:, f]] for f, name in enumerate(spec.sparse_names)]
    return torch.stack(rows, dim=1)


def linear_logit(params, spec: ModelSpec, ids, dense):
    """First-order term: sum_f w_f[id_f] + dense @ weight (basemodel.py:63-92)."""
    B = ids.shape[0]
    dt = params["out.bias"].dtype
    logit = torch.zeros(B, 1, dtype=dt)
    if spec.m > 0:
        emb = embedding_lookup(params, spec, ids, prefix="linear_model.embedding_dict.")  # [B, m, 1]
        logit = logit + emb.sum(dim=1)
    if spec.nd > 0:
        logit = logit + dense.to(dt).matmul(params["linear_model.weight"])
    return logit


def _activation(name, x):
    if name is None or name == "linear":
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    raise NotImplementedError(name)


def cin_forward(x0: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                split_half=True, activation="relu", pool=True, relu_masks=None):
    """Compressed Interaction Network (interaction.py:207-248).

    x0 [B, m, D];  weights[k] is the Conv1d(k=1) weight [H_k, h_{k-1}*m, 1] (or [H_k, h_{k-1}*m]);
    outer-product row order j = h*m + mm (h over X^{k-1}, mm over X^0).
    Returns [B, featuremap_num] (pool=True, sum over D) or the un-pooled maps [B, fm, D].
    """
    if x0.dim() != 3:
        raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % x0.dim())
    B, m, D = x0.shape
    hidden = x0
    finals = []
    n = len(weights)
    for k, (W, b) in enumerate(zip(weights, biases)):
        W2 = W.reshape(W.shape[0], -1)
        z = torch.einsum("bhd,bmd->bhmd", hidden, x0).reshape(B, hidden.shape[1] * m, D)
        y = torch.einsum("hk,bkd->bhd", W2, z) + b.view(1, -1, 1)     # Conv1d(kernel_size=1) == GEMM + bias
        if relu_masks is not None and activation == "relu":
            # test hook: use a GIVEN active-set (e.g. the one the fp32/bf16 kernel chose) so that pre-activations on
            # the ReLU kink do not turn rounding noise into O(1) gradient differences
            y = y * relu_masks[k].to(y.dtype)
        else:
            y = _activation(activation, y)
        H = W2.shape[0]
        if split_half:
            if k != n - 1:
                hidden, direct = y[:, : H // 2], y[:, H // 2:]
            else:
                direct, hidden = y, None
        else:
            direct, hidden = y, y
        finals.append(direct)
    maps = torch.cat(finals, dim=1)
    return maps.sum(-1) if pool else maps


def cin_preactivations(x0, weights, biases, split_half=True, activation="relu"):
    """Pre-activation tensors [B, H_k, D] of every CIN layer (test helper: lets parity tests drop samples whose
    pre-activation sits on the ReLU kink, where fp32 and fp64 legitimately pick different gradient masks)."""
    B, m, D = x0.shape
    hidden, pres = x0, []
    n = len(weights)
    for k, (W, b) in enumerate(zip(weights, biases)):
        W2 = W.reshape(W.shape[0], -1)
        z = torch.einsum("bhd,bmd->bhmd", hidden, x0).reshape(B, hidden.shape[1] * m, D)
        pre = torch.einsum("hk,bkd->bhd", W2, z) + b.view(1, -1, 1)
        pres.append(pre)
        y = _activation(activation, pre)
        H = W2.shape[0]
        hidden = y[:, : H // 2] if (split_half and k != n - 1) else y
    return pres


def dnn_forward(x, weights, biases, activation="relu"):
    """Linear(+bias) -> activation per layer (core.py:120-134); dropout p=0, no BN."""
    for W, b in zip(weights, biases):
        x = _activation(activation, x.matmul(W.t()) + b)
    return x


def mhsa(x, Wq, Wk, Wv, Wo, num_heads):
    """MultiHeadSelfAttention.forward (cin_attention.py:63-97), dropout p=0."""
    B, L, E = x.shape
    hd = E // num_heads
    q = x.matmul(Wq.t()).view(B, L, num_heads, hd).transpose(1, 2)
    k = x.matmul(Wk.t()).view(B, L, num_heads, hd).transpose(1, 2)
    v = x.matmul(Wv.t()).view(B, L, num_heads, hd).transpose(1, 2)
    s = q.matmul(k.transpose(-2, -1)) / math.sqrt(hd)
    p = torch.softmax(s, dim=-1)
    o = p.matmul(v).transpose(1, 2).contiguous().view(B, L, E)
    return o.matmul(Wo.t())


def attention_pool(x, W0, b0, w2):
    """AttentionPooling.forward (cin_attention.py:130-144): Linear->Tanh->Linear(.,1), softmax over L."""
    s = torch.tanh(x.matmul(W0.t()) + b0).matmul(w2.t())      # [B, L, 1]
    a = torch.softmax(s, dim=1)
    return (a * x).sum(dim=1)


def cin_attention_tail(params, spec: ModelSpec, maps):
    """Everything after `torch.cat(final_result, dim=1)` in CINAttention / CINAttentionV2."""
    E = maps.shape[-1]
    h = get_valid_num_heads(E, spec.num_heads)
    if spec.variant == "attn":
        a = mhsa(maps, params["cin.mhsa.W_q.weight"], params["cin.mhsa.W_k.weight"],
                 params["cin.mhsa.W_v.weight"], params["cin.mhsa.W_o.weight"], h)
        if spec.use_residual:
            a = a + maps
        if spec.use_layer_norm:
            a = F.layer_norm(a, (E,), params["cin.layer_norm.weight"], params["cin.layer_norm.bias"])
        pooled = attention_pool(a, params["cin.attn_pooling.attention.0.weight"],
                                params["cin.attn_pooling.attention.0.bias"],
                                params["cin.attn_pooling.attention.2.weight"])
        return pooled.matmul(params["cin.output_proj.weight"].t())
    assert spec.variant == "attn_v2"
    r = maps
    for i in range(spec.num_attn_layers):
        p = "cin.mhsa_layers.%d." % i
        a = mhsa(r, params[p + "W_q.weight"], params[p + "W_k.weight"], params[p + "W_v.weight"],
                 params[p + "W_o.weight"], h)
        if spec.use_residual:
            a = a + r
        if spec.use_layer_norm:
            a = F.layer_norm(a, (E,), params["cin.layer_norms.%d.weight" % i], params["cin.layer_norms.%d.bias" % i])
        r = a
    return attention_pool(r, params["cin.attn_pooling.attention.0.weight"],
                          params["cin.attn_pooling.attention.0.bias"],
                          params["cin.attn_pooling.attention.2.weight"])


def autodis_forward(params, spec: ModelSpec, dense):
    """AutoDis (autodis.py:100-127): per dense feature Linear(1,nb) -> LeakyReLU(0.2) -> Linear(nb,nb), softmax(scores / temp_f),
    weighted sum of the feature's meta-embeddings; features concatenated -> [B, nd * D]."""
    P = "autodis_encoder.autodis."
    outs = []
    for f in range(spec.nd):
        v = dense[:, f:f + 1]
        h = F.leaky_relu(v.matmul(params[P + "bucket_projectors.%d.0.weight" % f].t()) + params[P + "bucket_projectors.%d.0.bias" % f], 0.2)
        scores = h.matmul(params[P + "bucket_projectors.%d.2.weight" % f].t()) + params[P + "bucket_projectors.%d.2.bias" % f]
        w = torch.softmax(scores / params[P + "feature_temperatures"][f], dim=-1)
        outs.append(w.matmul(params[P + "meta_embeddings"][f]))
    return torch.cat(outs, dim=-1)


def _uses_autodis(spec: ModelSpec):
    return spec.variant == "pro" and spec.use_autodis and spec.nd > 0


def cin_params(params, spec: ModelSpec):
    n = len(spec.cin_layer_size)
    return ([params["cin.conv1ds.%d.weight" % k] for k in range(n)],
            [params["cin.conv1ds.%d.bias" % k] for k in range(n)])


def dnn_params(params, spec: ModelSpec):
    n = len(spec.dnn_hidden_units)
    return ([params["dnn.linears.%d.weight" % k] for k in range(n)],
            [params["dnn.linears.%d.bias" % k] for k in range(n)])


def xdeepfm_logit(params, spec: ModelSpec, X, return_parts=False):
    """final_logit before PredictionLayer (xdeepfm.py:79-103; attn models identical up to the CIN tail)."""
    ids, dense = split_input(spec, X)
    dt = params["out.bias"].dtype
    dense = dense.to(dt)
    emb = embedding_lookup(params, spec, ids)                 # [B, m, D]
    lin = linear_logit(params, spec, ids, dense)
    logit = lin
    parts = {"emb": emb, "linear_logit": lin}
    if len(spec.cin_layer_size) > 0:
        Ws, bs = cin_params(params, spec)
        if spec.variant in ("xdeepfm", "pro"):
            cin_out = cin_forward(emb, Ws, bs, spec.cin_split_half, spec.cin_activation, pool=True)
        else:
            maps = cin_forward(emb, Ws, bs, spec.cin_split_half, spec.cin_activation, pool=False)
            parts["cin_maps"] = maps
            cin_out = cin_attention_tail(params, spec, maps)
        cin_logit = cin_out.matmul(params["cin_linear.weight"].t())
        parts["cin_out"], parts["cin_logit"] = cin_out, cin_logit
        logit = logit + cin_logit
    if len(spec.dnn_hidden_units) > 0:
        dnn_dense = autodis_forward(params, spec, dense) if _uses_autodis(spec) else dense      # xdeepfm_pro.py:233-242
        dnn_in = torch.cat([emb.reshape(emb.shape[0], -1), dnn_dense], dim=-1) if spec.nd > 0 else emb.reshape(emb.shape[0], -1)
        Ws, bs = dnn_params(params, spec)
        dnn_out = dnn_forward(dnn_in, Ws, bs, spec.dnn_activation)
        dnn_logit = dnn_out.matmul(params["dnn_linear.weight"].t())
        parts["dnn_out"], parts["dnn_logit"] = dnn_out, dnn_logit
        logit = logit + dnn_logit
    if return_parts:
        return logit, parts
    return logit


def xdeepfm_forward(params, spec: ModelSpec, X):
    """y_pred [B,1] = sigmoid(final_logit + bias) (core.py:154-160)."""
    return torch.sigmoid(xdeepfm_logit(params, spec, X) + params["out.bias"])


def reg_groups(params, spec: ModelSpec):
    """(names, l2) groups registered by the reference ctor."""
    groups = []
    groups.append(([k for k in params if k.startswith("embedding_dict.")], spec.l2_reg_embedding))       # basemodel.py:126
    groups.append(([k for k in params if k.startswith("linear_model.")], spec.l2_reg_linear))            # basemodel.py:127
    groups.append(([k for k in params if k.startswith("dnn.") and "weight" in k and "bn" not in k], spec.l2_reg_dnn))  # xdeepfm.py:57-58
    groups.append(([k for k in params if k == "dnn_linear.weight"], spec.l2_reg_dnn))                    # xdeepfm.py:60
    groups.append(([k for k in params if k.startswith("cin.") and "weight" in k[len("cin."):]], spec.l2_reg_cin))  # xdeepfm.py:74-75
    return groups


def reg_loss(params, spec: ModelSpec):
    """sum over registered tensors of sum(l2 * p^2) (basemodel.py:412-428)."""
    total = torch.zeros(1, dtype=params["out.bias"].dtype)
    for names, l2 in reg_groups(params, spec):
        if l2 > 0:
            for k in names:
                total = total + torch.sum(l2 * torch.square(params[k]))
    return total


def train_loss(params, spec: ModelSpec, X, y):
    """(loss, total_loss): BCE(y_pred, y, 'sum') and loss + reg_loss (basemodel.py:245-257)."""
    y_pred = xdeepfm_forward(params, spec, X).squeeze(-1)
    loss = F.binary_cross_entropy(y_pred, y.to(y_pred.dtype).reshape(-1), reduction="sum")
    return loss, loss + reg_loss(params, spec).squeeze()


def sfg_loss(params, spec: ModelSpec, X, y):
    """SFG reconstruction loss of xDeepFM Pro in training mode, dropout p = 0 (sfg_decoder.py:116-157 decoder forward,
    :198-204 label-aware attention, :266-309 loss; targets re-read from X: basemodel_sfg.py:446-466)."""
    ids, dense = split_input(spec, X)
    dt = params["out.bias"].dtype
    dense = dense.to(dt)
    emb = embedding_lookup(params, spec, ids)                                  # [B, m, D]
    x = torch.cat([emb.reshape(emb.shape[0], -1), dense], dim=-1)               # sparse embeddings in field order, then dense
    labels = y.reshape(-1)
    P = "sfg_decoder."
    if spec.sfg_use_label_attention:
        lab = params[P + "label_attention.label_embedding.weight"][labels.long()]
        h = torch.relu(torch.cat([x, lab], dim=-1).matmul(params[P + "label_attention.attention_net.0.weight"].t())
                       + params[P + "label_attention.attention_net.0.bias"])
        gate = torch.sigmoid(h.matmul(params[P + "label_attention.attention_net.2.weight"].t())
                             + params[P + "label_attention.attention_net.2.bias"])
        x = x * gate
    hid = x
    for i in range(len(spec.sfg_hidden_units)):
        hid = torch.relu(hid.matmul(params[P + "shared_layers.%d.weight" % (3 * i)].t()) + params[P + "shared_layers.%d.bias" % (3 * i)])
    if spec.sfg_positive_only:
        mask = (labels == 1).to(dt)
        num = mask.sum() + 1e-8
    else:
        mask = torch.ones_like(labels, dtype=dt)
        num = labels.shape[0]
    total_sparse = torch.zeros((), dtype=dt)
    for f, name in enumerate(spec.sparse_names):
        logits = hid.matmul(params[P + "sparse_heads.%s.weight" % name].t()) + params[P + "sparse_heads.%s.bias" % name]
        ce = F.cross_entropy(logits, ids[:, f], reduction="none")
        total_sparse = total_sparse + (ce * mask).sum() / num
    total_dense = torch.zeros((), dtype=dt)
    if spec.nd > 0:
        pred = hid.matmul(params[P + "dense_head.weight"].t()) + params[P + "dense_head.bias"]
        mse = ((pred - dense) ** 2).mean(dim=-1)
        total_dense = (mse * mask).sum() / num
    return total_sparse + total_dense


def pro_loss_and_grads(params, spec: ModelSpec, X, y):
    """xDeepFM Pro train step (basemodel_sfg.py:316-349): total = BCE_sum + reg + sfg_weight * sfg_loss."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    y_pred = xdeepfm_forward(leaves, spec, X)
    loss = F.binary_cross_entropy(y_pred.squeeze(-1), y.to(y_pred.dtype).reshape(-1), reduction="sum")
    sfg = sfg_loss(leaves, spec, X, y) if spec.use_sfg else torch.zeros((), dtype=y_pred.dtype)
    total = loss + reg_loss(leaves, spec).squeeze() + spec.sfg_weight * sfg
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return y_pred.detach(), loss.detach(), sfg.detach(), total.detach(), grads


def loss_and_grads(params, spec: ModelSpec, X, y):
    """Run the restated train-step maths; returns (y_pred, loss, total, grads-by-name)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    y_pred = xdeepfm_forward(leaves, spec, X)
    loss = F.binary_cross_entropy(y_pred.squeeze(-1), y.to(y_pred.dtype).reshape(-1), reduction="sum")
    total = loss + reg_loss(leaves, spec).squeeze()
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return y_pred.detach(), loss.detach(), total.detach(), grads


# ---------------------------------------------------------------------------------------------
# deterministic parameter / input builders shared by fixtures, tests, smoke() and bench.py
# ---------------------------------------------------------------------------------------------

def param_shapes(spec: ModelSpec):
    """state_dict names -> shapes."""
    D, m, nd = spec.embedding_dim, spec.m, spec.nd
    shapes = {}
    for name, V in zip(spec.sparse_names, spec.vocab_sizes):
        shapes["embedding_dict.%s.weight" % name] = (V, D)
    if nd > 0:
        shapes["linear_model.weight"] = (nd, 1)
    for name, V in zip(spec.sparse_names, spec.vocab_sizes):
        shapes["linear_model.embedding_dict.%s.weight" % name] = (V, 1)
    shapes["out.bias"] = (1,)
    if len(spec.dnn_hidden_units) > 0:
        dims = [m * D + (nd * D if _uses_autodis(spec) else nd)] + list(spec.dnn_hidden_units)
        for i in range(len(dims) - 1):
            shapes["dnn.linears.%d.weight" % i] = (dims[i + 1], dims[i])
            shapes["dnn.linears.%d.bias" % i] = (dims[i + 1],)
        shapes["dnn_linear.weight"] = (1, dims[-1])
    if len(spec.cin_layer_size) > 0:
        prev = m
        for k, H in enumerate(spec.cin_layer_size):
            shapes["cin.conv1ds.%d.weight" % k] = (H, prev * m, 1)
            shapes["cin.conv1ds.%d.bias" % k] = (H,)
            prev = H // 2 if spec.cin_split_half else H
        E = D
        if spec.variant == "attn":
            for w in "qkvo":
                shapes["cin.mhsa.W_%s.weight" % w] = (E, E)
            if spec.use_layer_norm:
                shapes["cin.layer_norm.weight"] = (E,)
                shapes["cin.layer_norm.bias"] = (E,)
        elif spec.variant == "attn_v2":
            for i in range(spec.num_attn_layers):
                for w in "qkvo":
                    shapes["cin.mhsa_layers.%d.W_%s.weight" % (i, w)] = (E, E)
            if spec.use_layer_norm:
                for i in range(spec.num_attn_layers):
                    shapes["cin.layer_norms.%d.weight" % i] = (E,)
                    shapes["cin.layer_norms.%d.bias" % i] = (E,)
        if spec.variant in ("attn", "attn_v2"):
            shapes["cin.attn_pooling.attention.0.weight"] = (E, E)
            shapes["cin.attn_pooling.attention.0.bias"] = (E,)
            shapes["cin.attn_pooling.attention.2.weight"] = (1, E)
        if spec.variant == "attn":
            shapes["cin.output_proj.weight"] = (spec.featuremap_num, E)
        shapes["cin_linear.weight"] = (1, E if spec.variant == "attn_v2" else spec.featuremap_num)
    if spec.variant == "pro" and spec.use_sfg:
        # sfg_decoder.py:42-83 (shared_layers = Linear, ReLU, Dropout per hidden unit -> indices 0, 3, ...)
        d_in = m * D + nd
        prev = d_in
        for i, h in enumerate(spec.sfg_hidden_units):
            shapes["sfg_decoder.shared_layers.%d.weight" % (3 * i)] = (h, prev)
            shapes["sfg_decoder.shared_layers.%d.bias" % (3 * i)] = (h,)
            prev = h
        for name, V in zip(spec.sparse_names, spec.vocab_sizes):
            shapes["sfg_decoder.sparse_heads.%s.weight" % name] = (V, prev)
            shapes["sfg_decoder.sparse_heads.%s.bias" % name] = (V,)
        if nd > 0:
            shapes["sfg_decoder.dense_head.weight"] = (nd, prev)
            shapes["sfg_decoder.dense_head.bias"] = (nd,)
        if spec.sfg_use_label_attention:
            h0 = spec.sfg_hidden_units[0] if spec.sfg_hidden_units else 64
            shapes["sfg_decoder.label_attention.label_embedding.weight"] = (2, h0)
            shapes["sfg_decoder.label_attention.attention_net.0.weight"] = (h0, d_in + h0)
            shapes["sfg_decoder.label_attention.attention_net.0.bias"] = (h0,)
            shapes["sfg_decoder.label_attention.attention_net.2.weight"] = (d_in, h0)
            shapes["sfg_decoder.label_attention.attention_net.2.bias"] = (d_in,)
    if _uses_autodis(spec):
        # autodis.py:56-74 (appended last so that the seeded values of every other fixture stay what they were)
        nb = spec.autodis_buckets
        P = "autodis_encoder.autodis."
        shapes[P + "meta_embeddings"] = (nd, nb, D)
        for f in range(nd):
            shapes[P + "bucket_projectors.%d.0.weight" % f] = (nb, 1)
            shapes[P + "bucket_projectors.%d.0.bias" % f] = (nb,)
            shapes[P + "bucket_projectors.%d.2.weight" % f] = (nb, nb)
            shapes[P + "bucket_projectors.%d.2.bias" % f] = (nb,)
        shapes[P + "feature_temperatures"] = (nd,)
    return shapes


def make_params(spec: ModelSpec, seed=0, scale=None, dtype=torch.float32):
    """Seeded parameters with O(0.1..1) magnitudes so that every branch contributes to the logit
    (the reference's own init_std=1e-4 makes CIN/DNN terms ~1e-8, useless for parity testing)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in param_shapes(spec).items():
        fan_in = shp[1] if len(shp) > 1 else shp[0]
        if k.startswith("embedding_dict."):
            std = 0.5
        elif k.startswith("linear_model."):
            std = 0.1
        elif k.endswith(".bias"):
            std = 0.1
        elif "layer_norm" in k and k.endswith("weight"):
            out[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)).to(dtype)
            continue
        elif k.endswith("feature_temperatures"):
            out[k] = (0.6 + 0.8 * torch.rand(shp, generator=g)).to(dtype)       # positive, per-feature distinct
            continue
        elif k.endswith("meta_embeddings"):
            std = 0.5
        else:
            std = 1.0 / math.sqrt(max(fan_in, 1))
        if scale is not None and k in scale:
            std = scale[k]
        out[k] = (std * torch.randn(shp, generator=g)).to(dtype)
    return out


def make_inputs(spec: ModelSpec, B, seed=0, zipf=False):
    """Seeded synthetic batch: X float32 [B, m+nd] (ids stored as floats, as the reference does),
    y float32 [B].  ids < 2^24 so the float32 round trip is exact."""
    g = torch.Generator().manual_seed(seed + 1000)
    cols = []
    for V in spec.vocab_sizes:
        if zipf:
            u = torch.rand(B, generator=g)
            ids = torch.clamp((V ** u).long() - 1, 0, V - 1)       # log-uniform ~ Zipf(1)
        else:
            ids = torch.randint(0, V, (B,), generator=g)
        cols.append(ids.to(torch.float32))
    sparse = torch.stack(cols, 1) if cols else torch.zeros(B, 0)
    dense = torch.rand(B, spec.nd, generator=g)
    X = torch.cat([sparse, dense], 1) if spec.sparse_first else torch.cat([dense, sparse], 1)
    y = (torch.rand(B, generator=g) < 0.25).to(torch.float32)
    return X, y


def sequence_pool(seq, aux, mode, supports_masking):
    """SequencePoolingLayer.forward (sequence.py:51-79): seq [B, T, E]; aux = mask [B, T] (supports_masking) or length [B, 1].
    'mean' divides by (length + 1e-8) with the length as given; 'max' reduces seq - (1 - mask) * 1e9.  (The reference itself
    raises for mode='max' with a length input -- `1 - mask` on a bool mask, sequence.py:69 -- the restatement defines it by the
    same formula.)"""
    B, T, E = seq.shape
    if supports_masking:
        mask = aux.reshape(B, T).to(seq.dtype)
        length = mask.sum(dim=-1, keepdim=True)
    else:
        length = aux.reshape(B, 1)
        mask = (torch.arange(T).view(1, T) < length).to(seq.dtype)
        length = length.to(seq.dtype)
    mask = mask.unsqueeze(2)
    if mode == "max":
        return (seq - (1 - mask) * 1e9).max(dim=1, keepdim=True)[0]
    hist = (seq * mask).sum(dim=1)
    if mode == "mean":
        hist = hist / (length + 1e-8)
    return hist.unsqueeze(1)


def bag_pool(emb, ids, lens, fields):
    """Slot tensor [B, S, D] -> field tensor [B, F, D] (the layout of include/xdfm.h xdfm_bag_pool_fwd, restating
    get_varlen_pooling_list, inputs.py:141-155): `fields` = [(n_slots, mode, lencol)], mode 'single' copies the slot."""
    out, s = [], 0
    for n, mode, lencol in fields:
        x = emb[:, s:s + n]
        if mode == "single":
            out.append(x)
        elif lencol < 0:
            out.append(sequence_pool(x, ids[:, s:s + n] != 0, mode, True))
        else:
            out.append(sequence_pool(x, lens[:, lencol:lencol + 1], mode, False))
        s += n
    return torch.cat(out, dim=1)


def varlen_feature_index(desc):
    """build_input_features (inputs.py:99-123) for column descriptors {kind: sparse|dense|varlen, name, ...}: name -> (start, end)."""
    index, cursor = {}, 0
    for d in desc:
        if d["name"] in index:
            continue
        width = {"sparse": 1, "dense": d.get("dim", 1), "varlen": d.get("maxlen", 1)}[d["kind"]]
        index[d["name"]] = (cursor, cursor + width)
        cursor += width
        if d["kind"] == "varlen" and d.get("length_name") is not None and d["length_name"] not in index:
            index[d["length_name"]] = (cursor, cursor + 1)
            cursor += 1
    return index


def _varlen_fields(params, desc, X, prefix):
    """sparse_embedding_list + varlen_sparse_embedding_list (basemodel.py:354-380; the linear model: basemodel.py:65-78):
    [B, 1, W] per SparseFeat, then per VarLenSparseFeat the pooled sequence (inputs.py:141-155, 212-225), concatenated on dim 1."""
    fi = varlen_feature_index(desc)
    out = []
    for d in desc:
        if d["kind"] == "sparse":
            a, _ = fi[d["name"]]
            out.append(params[prefix + d["name"] + ".weight"][X[:, a].long()].unsqueeze(1))
    for d in desc:
        if d["kind"] == "varlen":
            a, b = fi[d["name"]]
            ids = X[:, a:b].long()
            seq = params[prefix + d["name"] + ".weight"][ids]                      # [B, T, W]
            if d.get("length_name") is None:
                out.append(sequence_pool(seq, ids != 0, d["combiner"], True))
            else:
                la, lb = fi[d["length_name"]]
                out.append(sequence_pool(seq, X[:, la:lb].long(), d["combiner"], False))
    return torch.cat(out, dim=1)


def varlen_xdeepfm_forward(params, desc, X, cin_layers=2, dnn_layers=2, return_parts=False):
    """xDeepFM.forward (xdeepfm.py:79-107) for a model whose linear and deep parts use the columns `desc`, multi-value features
    included; relu CIN with split_half, relu DNN, binary task."""
    fi = varlen_feature_index(desc)
    dt = params["out.bias"].dtype
    dense = torch.cat([X[:, fi[d["name"]][0]:fi[d["name"]][1]] for d in desc if d["kind"] == "dense"], dim=-1).to(dt)
    lin = _varlen_fields(params, desc, X, "linear_model.embedding_dict.").sum(dim=1) + dense.matmul(params["linear_model.weight"])
    emb = _varlen_fields(params, desc, X, "embedding_dict.")
    cin_out = cin_forward(emb, [params["cin.conv1ds.%d.weight" % k] for k in range(cin_layers)],
                          [params["cin.conv1ds.%d.bias" % k] for k in range(cin_layers)], True, "relu", pool=True)
    dnn_out = dnn_forward(torch.cat([emb.reshape(emb.shape[0], -1), dense], dim=-1),
                          [params["dnn.linears.%d.weight" % k] for k in range(dnn_layers)],
                          [params["dnn.linears.%d.bias" % k] for k in range(dnn_layers)], "relu")
    logit = lin + cin_out.matmul(params["cin_linear.weight"].t()) + dnn_out.matmul(params["dnn_linear.weight"].t())
    y = torch.sigmoid(logit + params["out.bias"])
    return (y, {"linear_logit": lin, "emb": emb}) if return_parts else y
