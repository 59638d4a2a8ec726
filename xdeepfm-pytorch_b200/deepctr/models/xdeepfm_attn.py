"""xDeepFM with attention-pooled CIN (reference: deepctr/models/xdeepfm_attn.py:25-301) on the fused B200 ops."""
from ..inputs import SparseFeat, VarLenSparseFeat
from ..layers.cin_attention import CINAttention, CINAttentionV2
from .basemodel import BaseModel
from .xdeepfm import xDeepFM


class _xDeepFMAttnBase(BaseModel):
    forward_ids = xDeepFM.forward_ids
    cin_output = xDeepFM.cin_output

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
        self._add_deep_tower(dnn_feature_columns, dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, init_std, device)
        E = self._get_embedding_size(dnn_feature_columns)
        self._add_cin(dnn_feature_columns, cin_layer_size, cin_split_half, l2_reg_cin, device,
                      lambda fields: CINAttention(field_size=fields, embedding_size=E, layer_size=cin_layer_size, activation=cin_activation,
                                                  split_half=cin_split_half, num_heads=cin_num_heads, attn_dropout=cin_attn_dropout,
                                                  use_layer_norm=cin_use_layer_norm, use_residual=cin_use_residual, l2_reg=l2_reg_cin,
                                                  seed=seed, device=device))
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
        self._add_deep_tower(dnn_feature_columns, dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, init_std, device)
        E = self._get_embedding_size(dnn_feature_columns)
        self._add_cin(dnn_feature_columns, cin_layer_size, cin_split_half, l2_reg_cin, device,
                      lambda fields: CINAttentionV2(field_size=fields, embedding_size=E, layer_size=cin_layer_size, activation=cin_activation,
                                                    split_half=cin_split_half, num_heads=cin_num_heads, attn_dropout=cin_attn_dropout,
                                                    use_layer_norm=cin_use_layer_norm, use_residual=cin_use_residual,
                                                    num_attn_layers=cin_num_attn_layers, l2_reg=l2_reg_cin, seed=seed, device=device),
                      head_width=E)        # the V2 block emits one attention-pooled [B, E] vector
        if self.use_cin:
            self.embedding_size_cin = E
        self.to(device)
