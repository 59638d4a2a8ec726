"""SequencePoolingLayer (reference: deepctr/layers/sequence.py:9-79) on the bag-pooling kernel (csrc/bag.cu).

Inside the models the sequence positions are slots of the fused multi-table gather and the pooling is part of
`BaseModel.embed` / `linear_logit`; this module keeps the reference's stand-alone layer (same constructor, same
input convention) for user code that calls it directly.  The attention / RNN sequence layers of the reference
(DIN / DIEN families) are outside the xDeepFM path."""
import torch
import torch.nn as nn

from .. import ops


class SequencePoolingLayer(nn.Module):
    """Pooling (sum, mean, max) over a variable-length sequence of embeddings.

    Input: [seq_value [B, T, E], mask [B, T] (supports_masking=True) or seq_len [B, 1] (supports_masking=False)].
    Output: [B, 1, E].  'mean' divides by (length + 1e-8); 'max' reduces x - (1 - mask) * 1e9 (sequence.py:69-77)."""

    def __init__(self, mode="mean", supports_masking=False, device="cpu"):
        super().__init__()
        if mode not in ["sum", "mean", "max"]:
            raise ValueError("parameter mode should in [sum, mean, max]")
        self.supports_masking = supports_masking
        self.mode = mode
        self.device = device
        self._layouts = {}

    def forward(self, seq_value_len_list):
        seq, aux = seq_value_len_list
        ops.require_cuda(seq, "SequencePoolingLayer")
        B, T, _ = seq.shape
        lay = self._layouts.get(T)
        if lay is None:
            lay = self._layouts[T] = ops.BagLayout([(T, self.mode, -1 if self.supports_masking else 0)])
        if self.supports_masking:
            ids = (aux.reshape(B, T) != 0).to(torch.int32)          # the kernel's mask test is `!= 0`
            lens = None
        else:
            ids = torch.zeros((B, T), dtype=torch.int32, device=seq.device)
            lens = aux.reshape(B, 1).to(torch.int32)
        return ops.BagPool.apply(lay, seq, ids, lens)
