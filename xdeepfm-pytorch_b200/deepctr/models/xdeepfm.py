"""xDeepFM (reference: deepctr/models/xdeepfm.py:16-107): Linear + CIN + DNN -> sigmoid, on the fused B200 ops."""
import torch

from .. import ops
from ..layers import CIN
from .basemodel import BaseModel


class xDeepFM(BaseModel):
    """Same constructor as the reference (xdeepfm.py:42-45).

    :param linear_feature_columns / dnn_feature_columns: feature columns of the linear / deep part.
    :param dnn_hidden_units, cin_layer_size, cin_split_half, cin_activation: network shape.
    :param l2_reg_linear, l2_reg_embedding, l2_reg_dnn, l2_reg_cin: L2 strengths of the four parameter groups.
    :param init_std, seed, dnn_dropout, dnn_activation, dnn_use_bn, task, device, gpus: as in the reference.
    """

    def __init__(self, linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(256, 256),
                 cin_layer_size=(256, 128,), cin_split_half=True, cin_activation='relu', l2_reg_linear=0.00001,
                 l2_reg_embedding=0.00001, l2_reg_dnn=0, l2_reg_cin=0, init_std=0.0001, seed=1024, dnn_dropout=0,
                 dnn_activation='relu', dnn_use_bn=False, task='binary', device='cpu', gpus=None):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task, device=device, gpus=gpus)
        self._add_deep_tower(dnn_feature_columns, dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, init_std, device)
        self._add_cin(dnn_feature_columns, cin_layer_size, cin_split_half, l2_reg_cin, device,
                      lambda fields: CIN(fields, cin_layer_size, cin_activation, cin_split_half, l2_reg_cin, seed, device=device))
        self.to(device)

    def cin_output(self, emb):
        return self.cin(emb)

    def forward_ids(self, ids, dense):
        """ids int32 [B, m_all], dense float32 [B, nd_all] (device) -> y_pred [B, 1]."""
        lin = self.linear_logit(ids, dense)
        emb = self.embed(ids) if (self.use_cin or self.use_dnn) and self._emb_plan is not None else None
        cin_out = w_cin = dnn_out = w_dnn = None
        if self.use_cin:
            cin_out, w_cin = self.cin_output(emb), self.cin_linear.weight
        if self.use_dnn:
            parts = []
            if emb is not None:
                parts.append(emb.reshape(emb.shape[0], -1))
            dd = self.dnn_dense(dense)
            if dd.shape[1] > 0:
                parts.append(dd)
            dnn_in = parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)
            dnn_out, w_dnn = self.dnn(dnn_in), self.dnn_linear.weight
        return ops.LogitHead.apply(lin, cin_out, w_cin, dnn_out, w_dnn, self.out.bias if self.out.use_bias else None,
                                   self.task == "binary")
