"""CIN with multi-head self-attention pooling (reference: deepctr/layers/cin_attention.py:15-466) on the fused kernels.

Same module / parameter names as the reference (state_dict keys `mhsa.W_{q,k,v,o}`, `layer_norm`, `attn_pooling.attention.{0,2}`,
`output_proj`; V2: `mhsa_layers.<i>`, `layer_norms.<i>`).  The CIN layers emit the un-pooled direct-connect maps [B, L, E] straight
from the contraction epilogue; the attention core keeps the [L, L] scores on chip (ops.MHSACore)."""
import math
import os

import torch
import torch.nn as nn

from .. import ops
from .activation import activation_layer, activation_name
from .interaction import _build_cin_convs, _cin_wb


def _get_valid_num_heads(embed_dim, num_heads):
    """Largest head count <= num_heads that divides embed_dim (reference: cin_attention.py:15-23)."""
    if embed_dim % num_heads == 0:
        return num_heads
    for h in range(num_heads, 0, -1):
        if embed_dim % h == 0:
            return h
    return 1


class MultiHeadSelfAttention(nn.Module):
    """(batch, seq_len, embed_dim) -> same shape (reference: cin_attention.py:26-97).  Bias-free projections, xavier-uniform."""

    def __init__(self, embed_dim, num_heads=4, dropout=0.0, device="cpu"):
        super().__init__()
        num_heads = _get_valid_num_heads(embed_dim, num_heads)
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.scale = math.sqrt(self.head_dim)
        self.W_q = nn.Linear(embed_dim, embed_dim, bias=False)
        self.W_k = nn.Linear(embed_dim, embed_dim, bias=False)
        self.W_v = nn.Linear(embed_dim, embed_dim, bias=False)
        self.W_o = nn.Linear(embed_dim, embed_dim, bias=False)
        self.dropout = nn.Dropout(dropout)
        self.dropout_rate = dropout
        self.precision = "fp32"      # projections: 'fp32' SGEMM | 'bf16' tcgen05 GEMM (the attention core itself is always fp32)
        for mod in (self.W_q, self.W_k, self.W_v, self.W_o):
            nn.init.xavier_uniform_(mod.weight)
        self.to(device)

    def forward(self, x):
        q, k, v = ops.linear_multi(x, [self.W_q.weight, self.W_k.weight, self.W_v.weight], precision=self.precision)
        # dropout on the probabilities (cin_attention.py:54, 86) happens inside the fused core: the keep mask is a counter-based hash
        # the backward recomputes; active in training mode only, like nn.Dropout
        p = float(self.dropout_rate) if (self.training and self.dropout_rate > 0) else 0.0
        if p >= 1.0:
            return ops.linear_act(torch.zeros_like(x), self.W_o.weight, precision=self.precision) + 0.0 * (q + k + v)
        o = ops.MHSACore.apply(q, k, v, self.num_heads, p)
        return ops.linear_act(o, self.W_o.weight, precision=self.precision)


class AttentionPooling(nn.Module):
    """(batch, seq_len, embed_dim) -> (batch, embed_dim): Linear -> Tanh -> Linear(., 1), softmax over the sequence, weighted sum
    (reference: cin_attention.py:100-144)."""

    def __init__(self, embed_dim, hidden_dim=None, device="cpu"):
        super().__init__()
        hidden_dim = hidden_dim or embed_dim
        self.attention = nn.Sequential(nn.Linear(embed_dim, hidden_dim), nn.Tanh(), nn.Linear(hidden_dim, 1, bias=False))
        self.precision = "fp32"
        for mod in self.attention:
            if isinstance(mod, nn.Linear):
                nn.init.xavier_uniform_(mod.weight)
                if mod.bias is not None:
                    nn.init.zeros_(mod.bias)
        self.to(device)

    def forward(self, x):
        t = ops.linear_act(x, self.attention[0].weight, self.attention[0].bias, "tanh", precision=self.precision)
        score = ops.linear_act(t, self.attention[2].weight, precision=self.precision)            # [B, L, 1]
        return ops.AttnPool.apply(score, x)


class _CINAttentionBase(nn.Module):
    def _init_cin(self, field_size, embedding_size, layer_size, activation, split_half, use_layer_norm, use_residual, l2_reg, seed):
        self.layer_size, self.split_half, self.l2_reg, self.seed = layer_size, split_half, l2_reg, seed
        self.embedding_size, self.use_layer_norm, self.use_residual = embedding_size, use_layer_norm, use_residual
        act = activation_name(activation)
        if act is None or act == "tanh":
            raise NotImplementedError("CIN activation '%s' is not fused in this build (relu / linear / sigmoid)" % activation)
        self.activation = activation_layer(activation)
        _build_cin_convs(self, field_size, layer_size, split_half)
        self.featuremap_num = sum(layer_size[:-1]) // 2 + layer_size[-1] if split_half else sum(layer_size)
        self._cfg = ops.CINConfig(field_size, layer_size, split_half, act, pool=False,
                                  impl=os.environ.get("XDFM_CIN_PRECISION", "fp32"))

    @property
    def precision(self):
        return self._cfg.impl

    @precision.setter
    def precision(self, value):
        if value not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self._cfg.impl = value
        for mod in self.modules():
            if isinstance(mod, (MultiHeadSelfAttention, AttentionPooling)):
                mod.precision = value

    def _maps(self, inputs):
        if len(inputs.shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
        return ops.cin_apply(self._cfg, inputs, *_cin_wb(self))          # [B, featuremap_num, E]

    def _attend(self, result, mhsa, norm):
        attn = mhsa(result)
        if not self.use_residual and not self.use_layer_norm:
            return attn
        return ops.AddLayerNorm.apply(attn, result if self.use_residual else None, norm.weight if norm is not None else None,
                                      norm.bias if norm is not None else None, norm.eps if norm is not None else 0.0,
                                      norm is not None)


class CINAttention(_CINAttentionBase):
    """(batch, field_size, embedding_size) -> (batch, featuremap_num) (reference: cin_attention.py:147-318)."""

    def __init__(self, field_size, embedding_size, layer_size=(128, 128), activation="relu", split_half=True, num_heads=4,
                 attn_dropout=0.0, use_layer_norm=True, use_residual=True, l2_reg=1e-5, seed=1024, device="cpu"):
        super().__init__()
        self._init_cin(field_size, embedding_size, layer_size, activation, split_half, use_layer_norm, use_residual, l2_reg, seed)
        self.mhsa = MultiHeadSelfAttention(embedding_size, num_heads, attn_dropout, device=device)
        if use_layer_norm:
            self.layer_norm = nn.LayerNorm(embedding_size)
        self.attn_pooling = AttentionPooling(embedding_size, embedding_size, device=device)
        self.output_proj = nn.Linear(embedding_size, self.featuremap_num, bias=False)
        nn.init.xavier_uniform_(self.output_proj.weight)
        self.to(device)

    def forward(self, inputs):
        result = self._maps(inputs)
        attn = self._attend(result, self.mhsa, self.layer_norm if self.use_layer_norm else None)
        pooled = self.attn_pooling(attn)
        return ops.linear_act(pooled, self.output_proj.weight)


class CINAttentionV2(_CINAttentionBase):
    """(batch, field_size, embedding_size) -> (batch, embedding_size): stacked attention layers, no output projection
    (reference: cin_attention.py:321-466)."""

    def __init__(self, field_size, embedding_size, layer_size=(128, 128), activation="relu", split_half=True, num_heads=4,
                 attn_dropout=0.0, use_layer_norm=True, use_residual=True, num_attn_layers=1, l2_reg=1e-5, seed=1024, device="cpu"):
        super().__init__()
        self._init_cin(field_size, embedding_size, layer_size, activation, split_half, use_layer_norm, use_residual, l2_reg, seed)
        self.num_attn_layers = num_attn_layers
        self.mhsa_layers = nn.ModuleList()
        self.layer_norms = nn.ModuleList() if use_layer_norm else None
        for _ in range(num_attn_layers):
            self.mhsa_layers.append(MultiHeadSelfAttention(embedding_size, num_heads, attn_dropout, device=device))
            if use_layer_norm:
                self.layer_norms.append(nn.LayerNorm(embedding_size))
        self.attn_pooling = AttentionPooling(embedding_size, embedding_size, device=device)
        self.to(device)

    def forward(self, inputs):
        result = self._maps(inputs)
        for i in range(self.num_attn_layers):
            result = self._attend(result, self.mhsa_layers[i], self.layer_norms[i] if self.use_layer_norm else None)
        return self.attn_pooling(result)
