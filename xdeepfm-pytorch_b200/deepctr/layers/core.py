"""DNN and PredictionLayer (reference: deepctr/layers/core.py:67-160) on the fused GEMM+bias+activation kernel."""
import os

import torch
import torch.nn as nn

from .. import ops
from .activation import activation_layer, activation_name


class DNN(nn.Module):
    """Multi-layer perceptron: Linear(+bias) -> [BatchNorm1d] -> activation -> Dropout per layer.

    Parameters live in `linears` (nn.Linear containers, same state_dict keys as the reference: `linears.<i>.weight/bias`);
    weights N(0, init_std), biases torch default (core.py:114-116).  Without BN each layer is ONE fused kernel.
    """

    def __init__(self, inputs_dim, hidden_units, activation="relu", l2_reg=0, dropout_rate=0, use_bn=False,
                 init_std=0.0001, dice_dim=3, seed=1024, device="cpu"):
        super().__init__()
        if len(hidden_units) == 0:
            raise ValueError("hidden_units is empty!!")
        self.dropout_rate, self.seed, self.l2_reg, self.use_bn = dropout_rate, seed, l2_reg, use_bn
        self.dropout = nn.Dropout(dropout_rate)
        dims = [inputs_dim] + list(hidden_units)
        self.linears = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])
        if use_bn:
            self.bn = nn.ModuleList([nn.BatchNorm1d(dims[i + 1]) for i in range(len(dims) - 1)])
        self._fused_act = activation_name(activation)
        self.precision = os.environ.get("XDFM_DNN_PRECISION", "fp32")   # 'fp32' (SGEMM, reference precision) | 'bf16' (tcgen05)
        self.activation_layers = nn.ModuleList(
            [activation_layer(activation, dims[i + 1], dice_dim) for i in range(len(dims) - 1)])
        for name, p in self.linears.named_parameters():
            if "weight" in name:
                nn.init.normal_(p, mean=0, std=init_std)
        self.to(device)

    def forward(self, inputs):
        x = inputs
        for i, lin in enumerate(self.linears):
            if self.use_bn or self._fused_act is None:
                x = ops.linear_act(x, lin.weight, lin.bias, None, precision=self.precision)
                if self.use_bn:
                    x = self.bn[i](x)
                x = self.activation_layers[i](x)
            else:
                x = ops.linear_act(x, lin.weight, lin.bias, self._fused_act, precision=self.precision)
            if self.dropout_rate > 0:
                x = self.dropout(x)
        return x


class PredictionLayer(nn.Module):
    """bias + sigmoid head (reference: core.py:144-160).  Inside the models the bias/sigmoid are fused into the
    LogitHead kernel; this module's forward is kept for stand-alone use."""

    def __init__(self, task="binary", use_bias=True, **kwargs):
        if task not in ["binary", "multiclass", "regression"]:
            raise ValueError("task must be binary,multiclass or regression")
        super().__init__()
        self.use_bias, self.task = use_bias, task
        if use_bias:
            self.bias = nn.Parameter(torch.zeros((1,)))

    def forward(self, X):
        return ops.LogitHead.apply(X, None, None, None, None, self.bias if self.use_bias else None, self.task == "binary")
