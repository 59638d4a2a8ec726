"""Activation factory (reference: deepctr/layers/activation.py:57-88).  Only the activations used by the xDeepFM
configs are fused into the CUDA epilogues (relu / linear / sigmoid / tanh); Dice and PReLU belong to other model
families and are rejected explicitly."""
import torch.nn as nn

FUSABLE = ("relu", "linear", "sigmoid", "tanh")


class Identity(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, inputs):
        return inputs


def activation_name(act):
    """Canonical lower-case name of a fusable activation, or None if `act` is not a fusable string."""
    if act is None:
        return "linear"
    if isinstance(act, str) and act.lower() in FUSABLE:
        return act.lower()
    return None


def activation_layer(act_name, hidden_size=None, dice_dim=2):
    if isinstance(act_name, str):
        name = act_name.lower()
        if name == "sigmoid":
            return nn.Sigmoid()
        if name == "linear":
            return Identity()
        if name == "relu":
            return nn.ReLU(inplace=True)
        if name == "tanh":
            return nn.Tanh()
        raise NotImplementedError("activation '%s' is outside the xDeepFM hot path of this build" % act_name)
    if isinstance(act_name, type) and issubclass(act_name, nn.Module):
        return act_name()
    raise NotImplementedError
