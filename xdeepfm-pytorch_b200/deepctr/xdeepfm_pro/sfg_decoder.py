"""SFG decoder, label-aware attention and reconstruction loss (reference: deepctr/xdeepfm_pro/sfg_decoder.py:19-311).

Same module / parameter names as the reference (`shared_layers.{0,3}`, `sparse_heads.<name>`, `dense_head`,
`label_attention.label_embedding`, `label_attention.attention_net.{0,2}`); every Linear runs through the fused GEMM + bias +
activation kernel, the per-field cross-entropy / MSE through one-pass masked-loss kernels (csrc/sfg.cu), and nothing is read
back to the host inside a step (the reference calls `.item()` m + 2 times per step: sfg_decoder.py:293, 305, 309)."""
import torch
import torch.nn as nn

from .. import ops


class LabelAwareAttention(nn.Module):
    """sigmoid gate over the decoder input conditioned on the label (reference: sfg_decoder.py:160-204)."""

    def __init__(self, input_dim, hidden_dim=64, device='cpu'):
        super().__init__()
        self.precision = "fp32"
        self.label_embedding = nn.Embedding(2, hidden_dim)
        self.attention_net = nn.Sequential(nn.Linear(input_dim + hidden_dim, hidden_dim), nn.ReLU(),
                                           nn.Linear(hidden_dim, input_dim), nn.Sigmoid())
        self.to(device)

    def forward(self, x, labels):
        if len(labels.shape) > 1:
            labels = labels.squeeze(-1)
        # 2-row lookup written as a blend (labels are 0/1): elementwise only, so the step stays CUDA-graph capturable and the
        # backward is a plain reduction instead of index_put
        w = self.label_embedding.weight
        lab = labels.to(w.dtype).reshape(-1, 1)
        label_emb = w[0].unsqueeze(0) + lab * (w[1] - w[0]).unsqueeze(0)
        combined = torch.cat([x, label_emb], dim=-1)
        h = ops.linear_act(combined, self.attention_net[0].weight, self.attention_net[0].bias, "relu", precision=self.precision)
        return ops.linear_act(h, self.attention_net[2].weight, self.attention_net[2].bias, "sigmoid", precision=self.precision)


class SFGDecoder(nn.Module):
    """MLP decoder reconstructing the original features from the embeddings (reference: sfg_decoder.py:19-157)."""

    def __init__(self, embedding_dim, sparse_feature_dims, dense_feature_names, hidden_units=(128, 64), dropout_rate=0.1,
                 use_label_aware_attention=True, device='cpu'):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.sparse_feature_dims = sparse_feature_dims
        self.dense_feature_names = dense_feature_names
        self.use_label_aware_attention = use_label_aware_attention
        self.device = device
        self.num_sparse_features = len(sparse_feature_dims)
        self.num_dense_features = len(dense_feature_names)
        self.dropout_rate = dropout_rate
        self._precision = "fp32"     # 'fp32' (SGEMM, reference precision) | 'bf16' (tcgen05 GEMMs for decoder layers and per-field heads)
        input_dim = self.num_sparse_features * embedding_dim + self.num_dense_features
        layers = []
        prev_dim = input_dim
        for hidden_dim in hidden_units:
            layers += [nn.Linear(prev_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout_rate)]
            prev_dim = hidden_dim
        self.shared_layers = nn.Sequential(*layers)
        self.sparse_heads = nn.ModuleDict()
        for feat_name, vocab_size in sparse_feature_dims.items():
            self.sparse_heads[feat_name] = nn.Linear(prev_dim, vocab_size)
        self.dense_head = nn.Linear(prev_dim, self.num_dense_features) if self.num_dense_features > 0 else None
        if use_label_aware_attention:
            self.label_attention = LabelAwareAttention(input_dim, hidden_units[0] if hidden_units else 64, device=device)
        self.to(device)

    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, value):
        if value not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self._precision = value
        if self.use_label_aware_attention:
            self.label_attention.precision = value

    def hidden(self, decoder_input, labels=None):
        """decoder_input [B, m*D + nd] -> last shared hidden state [B, h_last]."""
        if self.use_label_aware_attention and labels is not None:
            decoder_input = decoder_input * self.label_attention(decoder_input, labels)
        h = decoder_input
        for mod in self.shared_layers:
            if isinstance(mod, nn.Linear):
                h = ops.linear_act(h, mod.weight, mod.bias, "relu", precision=self._precision)
            elif isinstance(mod, nn.Dropout) and self.dropout_rate > 0:
                h = mod(h)
        return h

    def forward(self, sparse_embeddings, dense_values, labels=None):
        """Reference-shaped call: lists of [B,1,D] embeddings and [B,1] dense values -> (dict of logits, dense predictions)."""
        parts = [emb.reshape(emb.shape[0], -1) for emb in sparse_embeddings] + list(dense_values)
        decoder_input = torch.cat(parts, dim=-1) if len(parts) > 1 else parts[0]
        h = self.hidden(decoder_input, labels)
        sparse_logits = {name: ops.linear_act(h, head.weight, head.bias, precision=self._precision)
                         for name, head in self.sparse_heads.items()}
        if self.dense_head is not None:
            dense_preds = ops.linear_act(h, self.dense_head.weight, self.dense_head.bias, precision=self._precision)
        else:
            dense_preds = torch.zeros(h.shape[0], 0, device=h.device)
        return sparse_logits, dense_preds


class SFGLoss(nn.Module):
    """Cross-entropy (sparse) + MSE (dense) reconstruction loss, optionally on positive samples only
    (reference: sfg_decoder.py:207-311).  `loss_dict` values are device tensors (no host sync)."""

    def __init__(self, sparse_feature_names, dense_feature_names, sparse_weight=1.0, dense_weight=1.0, positive_only=True,
                 label_smooth=0.0, device='cpu'):
        super().__init__()
        self.sparse_feature_names = sparse_feature_names
        self.dense_feature_names = dense_feature_names
        self.sparse_weight, self.dense_weight = sparse_weight, dense_weight
        self.positive_only, self.label_smooth, self.device = positive_only, label_smooth, device

    def forward(self, sparse_logits, dense_preds, sparse_targets, dense_targets, labels):
        row_w = ops.sfg_row_weights(labels, self.positive_only)
        loss_dict = {}
        total_sparse = None
        for name in self.sparse_feature_names:
            if name in sparse_logits and name in sparse_targets:
                tgt = sparse_targets[name].reshape(-1, 1).to(torch.int32).contiguous()
                l = ops.MaskedCE.apply(sparse_logits[name], tgt, 0, row_w)
                total_sparse = l if total_sparse is None else total_sparse + l
                loss_dict['sfg_sparse_%s' % name] = l.detach()
        total = None if total_sparse is None else self.sparse_weight * total_sparse
        if len(self.dense_feature_names) > 0 and dense_preds.shape[1] > 0:
            d = ops.MaskedMSE.apply(dense_preds, dense_targets, row_w)
            loss_dict['sfg_dense'] = d.detach()
            total = self.dense_weight * d if total is None else total + self.dense_weight * d
        if total is None:
            total = torch.zeros(1, device=labels.device)
        loss_dict['sfg_total'] = total.detach()
        return total.reshape(()), loss_dict
