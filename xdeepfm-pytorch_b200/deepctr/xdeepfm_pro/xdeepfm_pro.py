"""xDeepFMPro / xDeepFMProLight (reference: deepctr/xdeepfm_pro/xdeepfm_pro.py:19-394) on the fused B200 ops."""
import torch

from .. import ops
from ..layers import CIN
from .basemodel_sfg import BaseModelSFG


class xDeepFMPro(BaseModelSFG):
    """Same constructor as the reference (xdeepfm_pro.py:57-87)."""

    def __init__(self, linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(256, 256), cin_layer_size=(256, 128),
                 cin_split_half=True, cin_activation='relu', l2_reg_linear=0.00001, l2_reg_embedding=0.00001, l2_reg_dnn=0,
                 l2_reg_cin=0, init_std=0.0001, seed=1024, dnn_dropout=0, dnn_activation='relu', dnn_use_bn=False, task='binary',
                 device='cpu', gpus=None, use_sfg=True, sfg_weight=0.1, sfg_hidden_units=(128, 64), sfg_dropout=0.1,
                 sfg_positive_only=True, sfg_use_label_attention=True, use_autodis=False, autodis_buckets=16,
                 autodis_temperature=1.0):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task, device=device, gpus=gpus,
                         use_sfg=use_sfg, sfg_weight=sfg_weight, sfg_hidden_units=sfg_hidden_units, sfg_dropout=sfg_dropout,
                         sfg_positive_only=sfg_positive_only, sfg_use_label_attention=sfg_use_label_attention)
        self.use_autodis = use_autodis
        # AutoDis encoder of the dense features for the DNN branch (xdeepfm_pro.py:130-143); built before the DNN like the reference
        self.autodis_encoder = None
        dnn_input_dim = self.compute_input_dim(dnn_feature_columns)
        if use_autodis and len(self.dense_feature_columns) > 0:
            from .autodis import DenseFeatureEncoder
            if any(fc.dimension != 1 for fc in self.dense_feature_columns):
                raise ValueError("use_autodis=True needs DenseFeat(dimension=1) columns (the reference's Linear(1, buckets) projectors)")
            self.autodis_encoder = DenseFeatureEncoder(dense_feature_names=[fc.name for fc in self.dense_feature_columns],
                                                       embedding_dim=self.embedding_dim, use_autodis=True, num_buckets=autodis_buckets,
                                                       temperature=autodis_temperature, device=device)
            # raw dense columns are replaced by their AutoDis embeddings in the tower's input (xdeepfm_pro.py:150-153)
            dnn_input_dim += self.autodis_encoder.get_output_dim() - sum(fc.dimension for fc in self.dense_feature_columns)
        self._add_deep_tower(dnn_feature_columns, dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, init_std, device,
                             input_dim=dnn_input_dim)
        self._add_cin(dnn_feature_columns, cin_layer_size, cin_split_half, l2_reg_cin, device,
                      lambda fields: CIN(fields, cin_layer_size, cin_activation, cin_split_half, l2_reg_cin, seed, device=device))
        self._last_emb = None
        self.to(device)

    def forward_ids(self, ids, dense):
        """Same network as xDeepFM (xdeepfm_pro.py:219-262); the embeddings of the step are kept for the SFG decoder."""
        lin = self.linear_logit(ids, dense)
        emb = self.embed(ids) if self._emb_plan is not None else None
        self._last_emb = emb
        cin_out = w_cin = dnn_out = w_dnn = None
        if self.use_cin:
            cin_out, w_cin = self.cin(emb), self.cin_linear.weight
        if self.use_dnn:
            parts = []
            if emb is not None:
                parts.append(emb.reshape(emb.shape[0], -1))
            dd = self.dnn_dense(dense)
            if dd.shape[1] > 0:
                if self.autodis_encoder is not None:   # xdeepfm_pro.py:233-240
                    dd = self.autodis_encoder(dd)[0]
                parts.append(dd)
            dnn_in = parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)
            dnn_out, w_dnn = self.dnn(dnn_in), self.dnn_linear.weight
        return ops.LogitHead.apply(lin, cin_out, w_cin, dnn_out, w_dnn, self.out.bias if self.out.use_bias else None,
                                   self.task == "binary")

    def get_embedding_analysis(self, X):
        """Embedding statistics for collapse diagnosis (reference: xdeepfm_pro.py:281-327); analysis helper, plain torch."""
        with torch.no_grad():
            ids, _ = self.split_input(X.to(self.device))
            all_embeddings = self.embed(ids)
            flat = all_embeddings.reshape(all_embeddings.shape[0], -1)
            normalized = flat / (flat.norm(dim=1, keepdim=True) + 1e-8)
            cos = torch.mm(normalized, normalized.t())
            avg = (cos.sum() - cos.trace()) / (cos.numel() - cos.shape[0])
            return {'mean_embedding': all_embeddings.mean(dim=0), 'std_embedding': all_embeddings.std(dim=0),
                    'embedding_variance': all_embeddings.var(dim=0).mean(), 'avg_sample_cosine_similarity': avg,
                    'num_fields': all_embeddings.shape[1], 'embedding_dim': all_embeddings.shape[2]}


class xDeepFMProLight(xDeepFMPro):
    """Smaller defaults (reference: xdeepfm_pro.py:330-394)."""

    def __init__(self, linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(128, 64), cin_layer_size=(128, 64),
                 cin_split_half=True, cin_activation='relu', l2_reg_linear=0.00001, l2_reg_embedding=0.00001, l2_reg_dnn=0,
                 l2_reg_cin=0, init_std=0.0001, seed=1024, dnn_dropout=0, dnn_activation='relu', dnn_use_bn=False, task='binary',
                 device='cpu', gpus=None, use_sfg=True, sfg_weight=0.05, sfg_hidden_units=(64, 32), sfg_dropout=0.1,
                 sfg_positive_only=True, sfg_use_label_attention=True, use_autodis=False, autodis_buckets=8,
                 autodis_temperature=1.0):
        super().__init__(linear_feature_columns=linear_feature_columns, dnn_feature_columns=dnn_feature_columns,
                         dnn_hidden_units=dnn_hidden_units, cin_layer_size=cin_layer_size, cin_split_half=cin_split_half,
                         cin_activation=cin_activation, l2_reg_linear=l2_reg_linear, l2_reg_embedding=l2_reg_embedding,
                         l2_reg_dnn=l2_reg_dnn, l2_reg_cin=l2_reg_cin, init_std=init_std, seed=seed, dnn_dropout=dnn_dropout,
                         dnn_activation=dnn_activation, dnn_use_bn=dnn_use_bn, task=task, device=device, gpus=gpus, use_sfg=use_sfg,
                         sfg_weight=sfg_weight, sfg_hidden_units=sfg_hidden_units, sfg_dropout=sfg_dropout,
                         sfg_positive_only=sfg_positive_only, sfg_use_label_attention=sfg_use_label_attention,
                         use_autodis=use_autodis, autodis_buckets=autodis_buckets, autodis_temperature=autodis_temperature)
