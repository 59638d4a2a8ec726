"""Compressed Interaction Network (reference: deepctr/layers/interaction.py:159-248) on the fused CIN kernels."""
import os

import torch.nn as nn

from .. import ops
from .activation import activation_layer, activation_name


def _build_cin_convs(module, field_size, layer_size, split_half):
    """conv1ds[k] = nn.Conv1d(h_{k-1}*m, H_k, 1) with bias and torch default init (interaction.py:189-201);
    the modules only hold parameters -- the contraction runs in ops.CINFunction."""
    if len(layer_size) == 0:
        raise ValueError("layer_size must be a list(tuple) of length greater than 1")
    module.field_nums = [field_size]
    module.conv1ds = nn.ModuleList()
    for i, size in enumerate(layer_size):
        module.conv1ds.append(nn.Conv1d(module.field_nums[-1] * module.field_nums[0], size, 1))
        if split_half:
            if i != len(layer_size) - 1 and size % 2 > 0:
                raise ValueError("layer_size must be even number except for the last layer when split_half=True")
            module.field_nums.append(size // 2)
        else:
            module.field_nums.append(size)


def _cin_wb(module):
    wb = []
    for conv in module.conv1ds:
        wb += [conv.weight, conv.bias]
    return wb


class CIN(nn.Module):
    """Input (batch, field_size, embedding_size) -> (batch, featuremap_num), featuremap_num =
    sum(layer_size[:-1]) // 2 + layer_size[-1] if split_half else sum(layer_size)."""

    def __init__(self, field_size, layer_size=(128, 128), activation="relu", split_half=True, l2_reg=1e-5, seed=1024,
                 device="cpu"):
        super().__init__()
        self.layer_size, self.split_half, self.l2_reg, self.seed = layer_size, split_half, l2_reg, seed
        act = activation_name(activation)
        if act is None or act == "tanh":
            raise NotImplementedError("CIN activation '%s' is not fused in this build (relu / linear / sigmoid)" % activation)
        self.activation = activation_layer(activation)
        _build_cin_convs(self, field_size, layer_size, split_half)
        self._cfg = ops.CINConfig(field_size, layer_size, split_half, act, pool=True,
                                  impl=os.environ.get("XDFM_CIN_PRECISION", "fp32"))
        self.to(device)

    @property
    def precision(self):
        """'fp32' (CUDA-core kernels, the reference's precision; default) or 'bf16' (tcgen05 tensor-core kernels, fp32 accumulate)."""
        return self._cfg.impl

    @precision.setter
    def precision(self, value):
        if value not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self._cfg.impl = value

    def forward(self, inputs):
        if len(inputs.shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
        return ops.cin_apply(self._cfg, inputs, *_cin_wb(self))
