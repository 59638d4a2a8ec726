"""xDeepFM (reference: deepctr/models/xdeepfm.py:16-107): Linear + CIN + DNN -> sigmoid, on the fused B200 ops."""
import torch
import torch.nn as nn

from .. import ops
from ..layers import CIN, DNN
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
        self.dnn_hidden_units = dnn_hidden_units
        self.use_dnn = len(dnn_feature_columns) > 0 and len(dnn_hidden_units) > 0
        if self.use_dnn:
            self.dnn = DNN(self.compute_input_dim(dnn_feature_columns), dnn_hidden_units, activation=dnn_activation,
                           l2_reg=l2_reg_dnn, dropout_rate=dnn_dropout, use_bn=dnn_use_bn, init_std=init_std, device=device)
            self.dnn_linear = nn.Linear(dnn_hidden_units[-1], 1, bias=False).to(device)
            self.add_regularization_weight(
                filter(lambda x: 'weight' in x[0] and 'bn' not in x[0], self.dnn.named_parameters()), l2=l2_reg_dnn)
            self.add_regularization_weight(self.dnn_linear.weight, l2=l2_reg_dnn)

        self.cin_layer_size = cin_layer_size
        self.use_cin = len(self.cin_layer_size) > 0 and len(dnn_feature_columns) > 0
        if self.use_cin:
            field_num = len(self.embedding_dict)
            if cin_split_half:
                self.featuremap_num = sum(cin_layer_size[:-1]) // 2 + cin_layer_size[-1]
            else:
                self.featuremap_num = sum(cin_layer_size)
            self.cin = CIN(field_num, cin_layer_size, cin_activation, cin_split_half, l2_reg_cin, seed, device=device)
            self.cin_linear = nn.Linear(self.featuremap_num, 1, bias=False).to(device)
            self.add_regularization_weight(filter(lambda x: 'weight' in x[0], self.cin.named_parameters()), l2=l2_reg_cin)
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
