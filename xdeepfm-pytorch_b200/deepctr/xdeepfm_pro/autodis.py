"""AutoDis soft-bucket encoder for dense features (reference: deepctr/xdeepfm_pro/autodis.py:20-238) on one fused kernel.

Same classes, constructor signatures, parameter names and state_dict keys as the reference (`meta_embeddings`,
`bucket_projectors.<f>.{0,2}.{weight,bias}`, `feature_temperatures`); the arithmetic of AutoDisLayer.forward (13 x two Linears,
LeakyReLU, temperature softmax, matmul with the meta-embeddings, cat) is `ops.AutoDis` = csrc/autodis.cu.  CUDA only."""
import torch
import torch.nn as nn

from .. import ops


class AutoDisLayer(nn.Module):
    """autodis.py:20-149.  forward(dense_values) accepts the reference's list of [B, 1] tensors or one [B, num_features] tensor."""

    def __init__(self, num_features, num_buckets=16, embedding_dim=8, temperature=1.0, keep_raw=True, device='cpu'):
        super().__init__()
        self.num_features = num_features
        self.num_buckets = num_buckets
        self.embedding_dim = embedding_dim
        self.temperature = temperature
        self.keep_raw = keep_raw
        self.device = device
        if num_features > 0:
            # same construction order as the reference (autodis.py:56-74) so that a seeded init consumes the RNG identically
            self.meta_embeddings = nn.Parameter(torch.randn(num_features, num_buckets, embedding_dim) * 0.01)
            self.bucket_projectors = nn.ModuleList([
                nn.Sequential(nn.Linear(1, num_buckets), nn.LeakyReLU(0.2), nn.Linear(num_buckets, num_buckets))
                for _ in range(num_features)])
            self.feature_temperatures = nn.Parameter(torch.ones(num_features) * temperature)
        self.to(device)

    def _values(self, dense_values):
        if isinstance(dense_values, torch.Tensor):
            return dense_values
        cols = [v.unsqueeze(-1) if v.dim() == 1 else v for v in dense_values]
        for v in cols:
            if v.shape[-1] != 1:
                raise ValueError("AutoDis expects one scalar per dense feature (DenseFeat dimension 1); got %s" % (tuple(v.shape),))
        return torch.cat(cols, dim=-1)

    def _proj_params(self):
        out = []
        for seq in self.bucket_projectors:
            out += [seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias]
        return out

    def forward(self, dense_values):
        """-> (flat [B, num_features * E], list of [B, 1, E] views), as the reference (autodis.py:79-129)."""
        if self.num_features == 0 or (not isinstance(dense_values, torch.Tensor) and len(dense_values) == 0):
            batch_size = dense_values[0].shape[0] if len(dense_values) else 1
            return torch.zeros(batch_size, 0, device=self.device), []
        x = self._values(dense_values)
        if x.shape[1] != self.num_features:
            raise ValueError("AutoDis: %d dense values for %d features" % (x.shape[1], self.num_features))
        flat = ops.AutoDis.apply(x, self.meta_embeddings, self.feature_temperatures, *self._proj_params())
        per = flat.view(x.shape[0], self.num_features, self.embedding_dim)
        return flat, [per[:, f:f + 1, :] for f in range(self.num_features)]

    def get_bucket_indices(self, dense_values):
        """Dominant bucket per feature (autodis.py:131-149); analysis helper, plain torch."""
        x = self._values(dense_values)
        out = []
        with torch.no_grad():
            for f, seq in enumerate(self.bucket_projectors):
                out.append(seq(x[:, f:f + 1]).argmax(dim=-1))
        return out


class DenseFeatureEncoder(nn.Module):
    """autodis.py:152-238."""

    def __init__(self, dense_feature_names, embedding_dim=8, use_autodis=True, num_buckets=16, temperature=1.0, device='cpu'):
        super().__init__()
        self.dense_feature_names = dense_feature_names
        self.embedding_dim = embedding_dim
        self.use_autodis = use_autodis
        self.num_features = len(dense_feature_names)
        self.device = device
        if use_autodis and self.num_features > 0:
            self.autodis = AutoDisLayer(num_features=self.num_features, num_buckets=num_buckets, embedding_dim=embedding_dim,
                                        temperature=temperature, device=device)
        else:
            self.autodis = None
        self.to(device)

    def forward(self, dense_values):
        """-> (encoded [B, out_dim], per-feature list, raw values [B, num_features])."""
        empty = (not isinstance(dense_values, torch.Tensor)) and len(dense_values) == 0
        if self.num_features == 0 or empty:
            batch_size = 1 if empty else dense_values[0].shape[0]
            z = torch.zeros(batch_size, 0, device=self.device)
            return z, [], z.clone()
        raw = dense_values if isinstance(dense_values, torch.Tensor) else torch.cat(list(dense_values), dim=-1)
        if self.use_autodis and self.autodis is not None:
            flat, emb_list = self.autodis(raw)
            return flat, emb_list, raw
        return raw, [raw[:, f:f + 1].unsqueeze(-1) for f in range(raw.shape[1])], raw

    def get_output_dim(self):
        return self.num_features * self.embedding_dim if self.use_autodis else self.num_features
