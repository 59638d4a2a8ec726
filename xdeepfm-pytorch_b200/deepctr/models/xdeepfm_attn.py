"""xDeepFM with attention-pooled CIN (reference: deepctr/models/xdeepfm_attn.py:25-301) on the fused B200 ops."""
import torch.nn as nn

from ..inputs import SparseFeat, VarLenSparseFeat
from ..layers import DNN
from ..layers.cin_attention import CINAttention, CINAttentionV2
from .basemodel import BaseModel
from .xdeepfm import xDeepFM


class _xDeepFMAttnBase(BaseModel):
    forward_ids = xDeepFM.forward_ids
    cin_output = xDeepFM.cin_output

    def _build_dnn(self, dnn_feature_columns, dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, init_std, device):
        self.dnn_hidden_units = dnn_hidden_units
        self.use_dnn = len(dnn_feature_columns) > 0 and len(dnn_hidden_units) > 0
        if self.use_dnn:
            self.dnn = DNN(self.compute_input_dim(dnn_feature_columns), dnn_hidden_units, activation=dnn_activation,
                           l2_reg=l2_reg_dnn, dropout_rate=dnn_dropout, use_bn=dnn_use_bn, init_std=init_std, device=device)
            self.dnn_linear = nn.Linear(dnn_hidden_units[-1], 1, bias=False).to(device)
            self.add_regularization_weight(
                filter(lambda x: 'weight' in x[0] and 'bn' not in x[0], self.dnn.named_parameters()), l2=l2_reg_dnn)
            self.add_regularization_weight(self.dnn_linear.weight, l2=l2_reg_dnn)

    def _get_embedding_size(self, feature_columns):
        for feat in feature_columns:
            if isinstance(feat, SparseFeat):
                return feat.embedding_dim
            if isinstance(feat, VarLenSparseFeat):
                return feat.sparsefeat.embedding_dim
        return 4


class xDeepFMAttention(_xDeepFMAttnBase):
    """Same constructor as the reference (xdeepfm_attn.py:55-62); CIN output = attention-pooled maps projected to featuremap_num."""

    def __init__(self, linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(256, 256), cin_layer_size=(256, 128,),
                 cin_split_half=True, cin_activation='relu', cin_num_heads=4, cin_attn_dropout=0.0, cin_use_layer_norm=True,
                 cin_use_residual=True, l2_reg_linear=0.00001, l2_reg_embedding=0.00001, l2_reg_dnn=0, l2_reg_cin=0,
                 init_std=0.0001, seed=1024, dnn_dropout=0, dnn_activation='relu', dnn_use_bn=False, task='binary', device='cpu',
                 gpus=None):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task, device=device, gpus=gpus)
        self._build_dnn(dnn_feature_columns, dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, init_std, device)
        self.cin_layer_size = cin_layer_size
        self.use_cin = len(self.cin_layer_size) > 0 and len(dnn_feature_columns) > 0
        if self.use_cin:
            field_num = len(self.embedding_dict)
            embedding_size = self._get_embedding_size(dnn_feature_columns)
            if cin_split_half:
                self.featuremap_num = sum(cin_layer_size[:-1]) // 2 + cin_layer_size[-1]
            else:
                self.featuremap_num = sum(cin_layer_size)
            self.cin = CINAttention(field_size=field_num, embedding_size=embedding_size, layer_size=cin_layer_size,
                                    activation=cin_activation, split_half=cin_split_half, num_heads=cin_num_heads,
                                    attn_dropout=cin_attn_dropout, use_layer_norm=cin_use_layer_norm, use_residual=cin_use_residual,
                                    l2_reg=l2_reg_cin, seed=seed, device=device)
            self.cin_linear = nn.Linear(self.featuremap_num, 1, bias=False).to(device)
            self.add_regularization_weight(filter(lambda x: 'weight' in x[0], self.cin.named_parameters()), l2=l2_reg_cin)
        self.to(device)


class xDeepFMAttentionV2(_xDeepFMAttnBase):
    """Same constructor as the reference (xdeepfm_attn.py:185-193); CIN output = [B, embedding_size], cin_linear is Linear(E, 1)."""

    def __init__(self, linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(256, 256), cin_layer_size=(256, 128,),
                 cin_split_half=True, cin_activation='relu', cin_num_heads=4, cin_attn_dropout=0.0, cin_use_layer_norm=True,
                 cin_use_residual=True, cin_num_attn_layers=1, l2_reg_linear=0.00001, l2_reg_embedding=0.00001, l2_reg_dnn=0,
                 l2_reg_cin=0, init_std=0.0001, seed=1024, dnn_dropout=0, dnn_activation='relu', dnn_use_bn=False, task='binary',
                 device='cpu', gpus=None):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task, device=device, gpus=gpus)
        self._build_dnn(dnn_feature_columns, dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, init_std, device)
        self.cin_layer_size = cin_layer_size
        self.use_cin = len(self.cin_layer_size) > 0 and len(dnn_feature_columns) > 0
        if self.use_cin:
            field_num = len(self.embedding_dict)
            embedding_size = self._get_embedding_size(dnn_feature_columns)
            self.embedding_size_cin = embedding_size
            self.cin = CINAttentionV2(field_size=field_num, embedding_size=embedding_size, layer_size=cin_layer_size,
                                      activation=cin_activation, split_half=cin_split_half, num_heads=cin_num_heads,
                                      attn_dropout=cin_attn_dropout, use_layer_norm=cin_use_layer_norm,
                                      use_residual=cin_use_residual, num_attn_layers=cin_num_attn_layers, l2_reg=l2_reg_cin,
                                      seed=seed, device=device)
            self.cin_linear = nn.Linear(embedding_size, 1, bias=False).to(device)
            self.add_regularization_weight(filter(lambda x: 'weight' in x[0], self.cin.named_parameters()), l2=l2_reg_cin)
        self.to(device)
