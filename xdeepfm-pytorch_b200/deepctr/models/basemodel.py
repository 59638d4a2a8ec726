"""BaseModel: the reference's training/inference API (deepctr/models/basemodel.py:96-527) on the B200 kernels.

Same public surface -- ctor arguments, compile / fit / evaluate / predict, input_from_feature_columns,
compute_input_dim, add_regularization_weight, get_regularization_loss, add_auxiliary_loss, state_dict keys, History --
with a different engine underneath:

  * ids travel as an int32 [B, m] matrix next to the float32 dense [B, nd] matrix (the reference pushes ids through one
    float32 matrix and corrupts ids >= 2^24: basemodel.py:195-198, 242, 368-370); `forward(X_float)` is still accepted.
  * one fused multi-table gather instead of 2*m nn.Embedding calls; CIN / DNN / head are fused CUDA ops (ops.py).
  * fit() is sync-free inside an epoch: loss and regulariser are accumulated on the device, per-step train metrics are
    computed from a device-side prediction log at epoch end; batches are staged through pinned memory on a copy stream.
  * named optimizers map to fused kernels (optim.py) that keep the reference's dense-table semantics.
"""
from __future__ import print_function

import time

import numpy as np
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..callbacks import CallbackList, History
from ..inputs import (DenseFeat, SparseFeat, VarLenSparseFeat, build_input_features, create_embedding_matrix,
                      dense_columns, sparse_columns, table_rows, varlen_columns)
from ..layers import CIN, DNN, PredictionLayer
from ..layers.utils import slice_arrays
from ..optim import FusedOptimizer, TableSet


def _expand_cols(feature_index, columns):
    cols = []
    for fc in columns:
        a, b = feature_index[fc.name]
        cols.extend(range(a, b))
    return cols


def _slot_cols(feature_index, columns, varlen=()):
    """Input-matrix columns of the ids looked up together: one slot per SparseFeat, then `maxlen` slots per VarLenSparseFeat
    (the order of the reference's sparse_embedding_list + varlen_sparse_embedding_list, basemodel.py:368-380)."""
    return _expand_cols(feature_index, columns) + _expand_cols(feature_index, varlen)


def _make_plan(columns, embedding_dict, width, varlen=()):
    names = list(embedding_dict.keys())
    table_of = [names.index(fc.embedding_name) for fc in columns]
    for fc in varlen:
        table_of += [names.index(fc.embedding_name)] * fc.maxlen
    rows = [table_rows(embedding_dict[n]) for n in names]
    return ops.SparsePlan(table_of, rows, width)


def _length_names(varlen):
    out = []
    for fc in varlen:
        if fc.length_name is not None and fc.length_name not in out:
            out.append(fc.length_name)
    return out


def _make_bag(columns, varlen, len_names):
    """ops.BagLayout of a lookup with multi-value features (None without them): fixed fields first, then one pooled field
    per VarLenSparseFeat, masked by id != 0 or by its length column (inputs.py:141-155)."""
    if not varlen:
        return None
    fields = [(1, "single", -1) for _ in columns]
    for fc in varlen:
        fields.append((fc.maxlen, fc.combiner, -1 if fc.length_name is None else len_names.index(fc.length_name)))
    return ops.BagLayout(fields)


class Linear(nn.Module):
    """First-order term: sum_f w_f[id_f] + dense @ weight -> [B, 1] (reference: basemodel.py:34-92)."""

    def __init__(self, feature_columns, feature_index, init_std=0.0001, device="cpu"):
        super().__init__()
        self.feature_index = feature_index
        self.device = device
        self.sparse_feature_columns = sparse_columns(feature_columns)
        self.dense_feature_columns = dense_columns(feature_columns)
        self.varlen_sparse_feature_columns = varlen_columns(feature_columns)
        self.embedding_dict = create_embedding_matrix(feature_columns, init_std, linear=True, sparse=False, device=device)
        if len(self.dense_feature_columns) > 0:
            self.weight = nn.Parameter(torch.empty(sum(fc.dimension for fc in self.dense_feature_columns), 1, device=device))
            nn.init.normal_(self.weight, mean=0, std=init_std)
        self._sparse_cols = _slot_cols(feature_index, self.sparse_feature_columns, self.varlen_sparse_feature_columns)
        self._dense_cols = _expand_cols(feature_index, self.dense_feature_columns)
        self._plan = _make_plan(self.sparse_feature_columns, self.embedding_dict, 1, self.varlen_sparse_feature_columns)
        self._cache = ops.SegmentCache()
        self._bind_lengths(_length_names(self.varlen_sparse_feature_columns))

    def _bind_lengths(self, len_names):
        """`len_names`: the length columns (in the order of the `lens` tensor forward_ids receives) of the sequence features."""
        self._len_names = list(len_names)
        self._len_cols = [self.feature_index[n][0] for n in self._len_names]
        self._bag = _make_bag(self.sparse_feature_columns, self.varlen_sparse_feature_columns, self._len_names)

    def _tables(self):
        return [emb.weight for emb in self.embedding_dict.values()]

    def forward_ids(self, ids, dense, cache=None, lens=None):
        w = self.weight if len(self.dense_feature_columns) > 0 else None
        if self._bag is None:
            return ops.LinearTerm.apply(self._plan, cache or self._cache, ids, dense if w is not None else None, w, *self._tables())
        # multi-value features: first-order rows of every slot, pooled per field under the sequence mask, summed over the fields
        # (basemodel.py:63-92: cat(sparse + pooled sequence embeddings, -1).sum(-1) + dense @ weight)
        rows = ops.SparseGather.apply(self._plan, cache or self._cache, ids, *self._tables())
        logit = ops.BagPool.apply(self._bag, rows, ids, lens).sum(dim=1)
        if w is not None:
            logit = logit + ops.linear_act(dense, w.t().contiguous(), None, None)
        return logit

    def forward(self, X, sparse_feat_refine_weight=None):
        if sparse_feat_refine_weight is not None:
            raise NotImplementedError("sparse_feat_refine_weight (IFM/DIFM) is outside the xDeepFM hot path")
        ids, dense = ops.split_input(X, self._sparse_cols, self._dense_cols)
        lens = ops.split_input(X, self._len_cols, [])[0] if self._len_cols else None
        return self.forward_ids(ids, dense, lens=lens)


class BaseModel(nn.Module):
    def __init__(self, linear_feature_columns, dnn_feature_columns, l2_reg_linear=1e-5, l2_reg_embedding=1e-5,
                 init_std=0.0001, seed=1024, task="binary", device="cpu", gpus=None):
        super().__init__()
        torch.manual_seed(seed)
        self.dnn_feature_columns = dnn_feature_columns
        self.linear_feature_columns = linear_feature_columns
        self.device = device
        self.gpus = gpus
        if gpus and str(self.gpus[0]) not in self.device:
            raise ValueError("`gpus[0]` should be the same gpu with `device`")
        self.reg_loss = torch.zeros((1,), device=device)
        self.aux_loss = torch.zeros((1,), device=device)
        self.feature_index = build_input_features(list(linear_feature_columns) + list(dnn_feature_columns))
        self.embedding_dict = create_embedding_matrix(dnn_feature_columns, init_std, sparse=False, device=device)
        self.linear_model = Linear(linear_feature_columns, self.feature_index, device=device)

        self.regularization_weight = []
        self.add_regularization_weight(self.embedding_dict.parameters(), l2=l2_reg_embedding)
        self.add_regularization_weight(self.linear_model.parameters(), l2=l2_reg_linear)
        self._l2_reg_embedding, self._l2_reg_linear = l2_reg_embedding, l2_reg_linear

        self.out = PredictionLayer(task, )
        self.task = task

        # fused-lookup plans: the sparse / dense columns of the deep part and of the linear part
        self._dnn_sparse = sparse_columns(dnn_feature_columns)
        self._dnn_varlen = varlen_columns(dnn_feature_columns)
        self._dnn_dense = dense_columns(dnn_feature_columns)
        dims = set(fc.embedding_dim for fc in self._dnn_sparse + self._dnn_varlen)
        if len(dims) > 1:
            raise ValueError("embedding_dim of SparseFeat and VarlenSparseFeat must be same in this model!")
        self._emb_dim = dims.pop() if dims else 0
        self._emb_plan = _make_plan(self._dnn_sparse, self.embedding_dict, self._emb_dim, self._dnn_varlen) \
            if (self._dnn_sparse or self._dnn_varlen) else None
        self._seg_cache = ops.SegmentCache()
        # union column layout used by the int32/float32 feed
        all_cols = []
        seen = set()
        for fc in list(linear_feature_columns) + list(dnn_feature_columns):
            if fc.name not in seen:
                seen.add(fc.name)
                all_cols.append(fc)
        self._all_sparse = sparse_columns(all_cols)
        self._all_varlen = varlen_columns(all_cols)
        self._all_dense = dense_columns(all_cols)
        # id-type columns of the feed: one per SparseFeat, maxlen per VarLenSparseFeat, then the length columns
        self._len_names = _length_names(self._all_varlen)
        self._len_cols = [self.feature_index[n][0] for n in self._len_names]
        self._all_sparse_cols = _slot_cols(self.feature_index, self._all_sparse, self._all_varlen) + self._len_cols
        self._all_dense_cols = _expand_cols(self.feature_index, self._all_dense)
        n_ids = len(self._all_sparse_cols)
        slot_pos, cursor = {}, 0
        for fc in self._all_sparse + self._all_varlen:
            width = fc.maxlen if isinstance(fc, VarLenSparseFeat) else 1
            slot_pos[fc.name] = list(range(cursor, cursor + width))
            cursor += width
        self._len_sel = list(range(cursor, n_ids))

        def slots_of(sparse, varlen):
            return [p for fc in list(sparse) + list(varlen) for p in slot_pos[fc.name]]

        self._dnn_sparse_sel = self._selector(slots_of(self._dnn_sparse, self._dnn_varlen), n_ids)
        self._lin_sparse_sel = self._selector(slots_of(self.linear_model.sparse_feature_columns,
                                                       self.linear_model.varlen_sparse_feature_columns), n_ids)
        self._emb_bag = _make_bag(self._dnn_sparse, self._dnn_varlen, self._len_names)
        self.linear_model._bind_lengths(self._len_names)
        dcol_pos = {c: i for i, c in enumerate(self._all_dense_cols)}
        self._dnn_dense_sel = self._selector([dcol_pos[c] for c in _expand_cols(self.feature_index, self._dnn_dense)],
                                             len(self._all_dense_cols))
        self._lin_dense_sel = self._selector([dcol_pos[c] for c in self.linear_model._dense_cols], len(self._all_dense_cols))

        self.to(device)
        self._is_graph_network = True
        self._ckpt_saved_epoch = False
        self.history = History()
        self.stop_training = False
        self.sparse_embedding_update = False   # True: update only rows touched by the batch (NOT reference semantics)
        self._dist = None                      # deepctr.distributed.DistContext after distribute()
        self._optimizer_spec = None
        import os as _os
        self.use_cuda_graph = _os.environ.get("XDFM_CUDA_GRAPH", "1") != "0"
        self._graphs, self._graph_seen, self._graph_failed = {}, {}, False
        self._capturing_half = False
        self._capturing = False
        self._in_train_step = False

    # Not picklable / not meaningful in another process: captured CUDA graphs (static buffers in a private pool), cached segment
    # sorts, device index tensors.  `torch.save(model)` -- the reference's ModelCheckpoint default, callbacks.py:58-61 -- drops them;
    # they are re-created lazily.
    _TRANSIENT = ("_graphs", "_graph_seen", "_sel_cache", "_seg_cache", "_tob_accum", "_dist")

    def __getstate__(self):
        state = dict(self.__dict__)
        if state.get("_dist") is not None:
            raise RuntimeError("a distributed model cannot be pickled: save model.state_dict() (collective) instead")
        for k in self._TRANSIENT:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._graphs, self._graph_seen, self._graph_failed = {}, {}, False
        self._seg_cache = ops.SegmentCache()
        self._dist = None
        self._capturing = self._capturing_half = self._in_train_step = False

    # ------------------------------------------------------------------------------------------
    # sub-network builders shared by xDeepFM, the attention variants and xDeepFM Pro.  Modules are created in the reference's
    # order (deep tower, its head, CIN, its head) so that a seeded construction draws the same initial weights, and registered
    # under the reference's attribute names (state_dict keys).
    # ------------------------------------------------------------------------------------------
    def _add_deep_tower(self, columns, hidden_units, activation, l2, dropout, use_bn, init_std, device, input_dim=None):
        """`dnn` + `dnn_linear` over the flattened field embeddings and dense values (reference: xdeepfm.py:50-60).  L2 group:
        the tower's weight matrices (no biases, no BatchNorm parameters) and the head."""
        self.dnn_hidden_units = hidden_units
        self.use_dnn = len(columns) > 0 and len(hidden_units) > 0
        if not self.use_dnn:
            return
        width = self.compute_input_dim(columns) if input_dim is None else input_dim
        self.dnn = DNN(width, hidden_units, activation=activation, l2_reg=l2, dropout_rate=dropout, use_bn=use_bn,
                       init_std=init_std, device=device)
        self.dnn_linear = nn.Linear(hidden_units[-1], 1, bias=False).to(device)
        decayed = [(n, w) for n, w in self.dnn.named_parameters() if "weight" in n and "bn" not in n]
        self.add_regularization_weight(decayed, l2=l2)
        self.add_regularization_weight(self.dnn_linear.weight, l2=l2)

    def _add_cin(self, columns, layer_size, split_half, l2, device, make_cin, head_width=None):
        """`cin` + `cin_linear` (reference: xdeepfm.py:62-75).  `make_cin(field_num)` builds the interaction module;
        featuremap_num = what the pooled CIN emits: every layer's direct half, the whole last layer (or all maps without split_half).
        L2 group: every CIN parameter whose name contains 'weight'."""
        self.cin_layer_size = layer_size
        self.use_cin = len(layer_size) > 0 and len(columns) > 0
        if not self.use_cin:
            return
        maps = sum(layer_size)
        if split_half:
            maps = sum(layer_size[:-1]) // 2 + layer_size[-1]
        self.featuremap_num = maps
        self.cin = make_cin(len(self.embedding_dict))
        self.cin_linear = nn.Linear(maps if head_width is None else head_width, 1, bias=False).to(device)
        self.add_regularization_weight([(n, w) for n, w in self.cin.named_parameters() if "weight" in n], l2=l2)

    @staticmethod
    def _selector(idx, n):
        return None if idx == list(range(n)) else idx

    def _select(self, t, sel):
        """Columns `sel` of t (None = all).  The index tensor lives on the device and is built once: indexing with a Python list
        would copy it host -> device on every call, which a CUDA-graph capture of the step does not allow."""
        if sel is None:
            return t
        cache = self.__dict__.setdefault("_sel_cache", {})
        key = (tuple(sel), t.device)
        idx = cache.get(key)
        if idx is None:
            idx = cache[key] = torch.tensor(list(sel), dtype=torch.int64, device=t.device)
        # the deep and the first-order part usually select the same columns of the same batch: one tensor for both (the row-sharded
        # lookup recognises a batch by tensor identity and then fetches its distinct rows once).  The memo holds `t`, so its
        # address cannot be recycled while it is remembered; an in-place write bumps _version.
        slot = ("last", t.dtype)
        memo = cache.get(slot)
        if memo is not None and memo[0] is t and memo[1] == t._version and memo[2] == key[0]:
            return memo[3]
        out = torch.index_select(t, 1, idx)
        cache[slot] = (t, t._version, key[0], out)
        return out

    # ------------------------------------------------------------------------------------------
    # inputs
    # ------------------------------------------------------------------------------------------
    def split_input(self, X):
        """X float [B, n_columns] -> (ids int32 [B, m_all], dense float32 [B, nd_all]) on the device."""
        return ops.split_input(X, self._all_sparse_cols, self._all_dense_cols)

    def embed(self, ids_all):
        """Fused multi-table gather -> [B, m, D] for the deep part's sparse features."""
        ids = self._select(ids_all, self._dnn_sparse_sel)
        if self._dist is None and any(hasattr(e, "deferred_rows") for e in self.embedding_dict.values()):
            raise RuntimeError("this model was built under deepctr.inputs.deferred_tables(): call model.distribute() first")
        if self._dist is not None:
            from ..distributed import ShardedGather
            sh = self._dist.sharded
            emb = ShardedGather.apply(sh, ids, sh.anchor)
            if self._emb_bag is not None:   # multi-value features: one slot per sequence position, pooled per field
                emb = ops.BagPool.apply(self._emb_bag, emb, ids, self._lens(ids_all))
            return emb
        self._tables_current(self._emb_plan, ids)
        tables = [emb.weight for emb in self.embedding_dict.values()]
        emb = ops.SparseGather.apply(self._emb_plan, self._seg_cache, ids, *tables)
        if self._emb_bag is not None:       # multi-value features: [B, slots, D] -> [B, fields, D]
            emb = ops.BagPool.apply(self._emb_bag, emb, ids, self._lens(ids_all))
        return emb

    def _lens(self, ids_all):
        """int32 [B, n_length_columns] sequence lengths of the batch (None without length columns)."""
        return self._select(ids_all, self._len_sel) if self._len_sel else None

    def linear_logit(self, ids_all, dense_all):
        ids = self._select(ids_all, self._lin_sparse_sel)
        dense = self._select(dense_all, self._lin_dense_sel)
        if self._dist is not None:
            from ..distributed import ShardedLinearRows, ShardedLinearTerm
            sh = self._dist.sharded
            has_w = len(self.linear_model.dense_feature_columns) > 0
            if self.linear_model._bag is not None:
                # multi-value features: per-lookup rows, pooled per field under the sequence mask, summed over the fields
                rows = ShardedLinearRows.apply(sh, ids, sh.anchor)
                logit = ops.BagPool.apply(self.linear_model._bag, rows, ids, self._lens(ids_all)).sum(dim=1)
                if has_w:
                    logit = logit + ops.linear_act(dense, self.linear_model.weight.t().contiguous(), None, None)
                return logit
            return ShardedLinearTerm.apply(sh, ids, dense if has_w else None, self.linear_model.weight if has_w else None, sh.anchor)
        self._tables_current(self.linear_model._plan, ids)
        return self.linear_model.forward_ids(ids, dense, cache=self._seg_cache, lens=self._lens(ids_all))

    def _tables_current(self, plan, ids):
        """Lazy dense-table semantics (optim.FusedOptimizer): rows the optimizer has postponed are replayed before they are read --
        the looked-up rows in a training step, every row otherwise."""
        opt = getattr(self, "optim", None)
        if isinstance(opt, FusedOptimizer) and opt._dirty:
            # "inside a training step", not module.training: the reference's fit() leaves the model in eval mode after the first
            # validation pass (basemodel.py:202, :332), and a full flush per step would stream every table row every step
            if self._in_train_step:
                opt.catch_up(plan, self._seg_cache, ids)
            else:
                opt.flush()

    def dnn_dense(self, dense_all):
        return self._select(dense_all, self._dnn_dense_sel)

    def input_from_feature_columns(self, X, feature_columns, embedding_dict, support_dense=True):
        """Reference-shaped helper (basemodel.py:354-380): ([B,1,D] per sparse feature, [B,w] per dense feature)."""
        sp, de, vl = sparse_columns(feature_columns), dense_columns(feature_columns), varlen_columns(feature_columns)
        if not support_dense and len(de) > 0:
            raise ValueError("DenseFeat is not supported in dnn_feature_columns")
        ids, _ = ops.split_input(X, _slot_cols(self.feature_index, sp, vl), [])
        plan = _make_plan(sp, embedding_dict, (sp + vl)[0].embedding_dim, vl) if (sp or vl) else None
        emb_list = []
        if sp or vl:
            emb = ops.SparseGather.apply(plan, ops.SegmentCache(), ids, *[e.weight for e in embedding_dict.values()])
            if vl:
                names = _length_names(vl)
                lens = ops.split_input(X, [self.feature_index[n][0] for n in names], [])[0] if names else None
                emb = ops.BagPool.apply(_make_bag(sp, vl, names), emb, ids, lens)
            emb_list = list(torch.split(emb, 1, dim=1))
        dense_list = [X[:, self.feature_index[fc.name][0]:self.feature_index[fc.name][1]] for fc in de]
        return emb_list, dense_list

    def compute_input_dim(self, feature_columns, include_sparse=True, include_dense=True, feature_group=False):
        sp = [fc for fc in feature_columns if isinstance(fc, (SparseFeat, VarLenSparseFeat))] if len(feature_columns) else []
        de = dense_columns(feature_columns)
        dense_dim = sum(fc.dimension for fc in de)
        sparse_dim = len(sp) if feature_group else sum(fc.embedding_dim for fc in sp)
        return (sparse_dim if include_sparse else 0) + (dense_dim if include_dense else 0)

    @property
    def embedding_size(self):
        sp = [fc for fc in self.dnn_feature_columns if isinstance(fc, (SparseFeat, VarLenSparseFeat))]
        sizes = set(fc.embedding_dim for fc in sp)
        if len(sizes) > 1:
            raise ValueError("embedding_dim of SparseFeat and VarlenSparseFeat must be same in this model!")
        return list(sizes)[0]

    def forward(self, X):
        X = X.to(self.device) if not X.is_cuda else X
        ids, dense = self.split_input(X)
        return self.forward_ids(ids, dense)

    def forward_ids(self, ids, dense):
        raise NotImplementedError

    # ------------------------------------------------------------------------------------------
    # regularisation
    # ------------------------------------------------------------------------------------------
    def add_regularization_weight(self, weight_list, l1=0.0, l2=0.0):
        if isinstance(weight_list, torch.nn.parameter.Parameter):
            weight_list = [weight_list]
        else:
            weight_list = list(weight_list)
        self.regularization_weight.append((weight_list, l1, l2))

    def get_regularization_loss(self):
        """sum(l1*|p|) + sum(l2*p^2) over the registered tensors (reference: basemodel.py:412-428); autograd-visible.
        The fused fit path does not call this: the optimizer kernels add 2*l2*w and accumulate the loss value."""
        dev = next(self.parameters()).device
        total = torch.zeros((1,), device=dev)
        for weight_list, l1, l2 in self.regularization_weight:
            for w in weight_list:
                p = w[1] if isinstance(w, tuple) else w
                if l1 > 0:
                    total = total + torch.sum(l1 * torch.abs(p))
                if l2 > 0:
                    total = total + torch.sum(l2 * torch.square(p))
        return total

    def add_auxiliary_loss(self, aux_loss, alpha):
        self.aux_loss = aux_loss * alpha

    def _l2_map(self):
        m = {}
        has_l1 = False
        for weight_list, l1, l2 in self.regularization_weight:
            has_l1 = has_l1 or l1 > 0
            for w in weight_list:
                p = w[1] if isinstance(w, tuple) else w
                m[id(p)] = m.get(id(p), 0.0) + float(l2)
        return m, has_l1

    # ------------------------------------------------------------------------------------------
    # compile
    # ------------------------------------------------------------------------------------------
    def compile(self, optimizer, loss=None, metrics=None):
        self.metrics_names = ["loss"]
        self._optimizer_spec = optimizer
        self.optim = self._get_optim(optimizer)
        self.loss_func = self._get_loss_func(loss)
        self._loss_name = loss if isinstance(loss, str) else None
        self.metrics = self._get_metrics(metrics)

    def _table_sets(self):
        l2map, _ = self._l2_map()
        sets = []
        if self._dist is not None:
            return sets                 # tables are row-sharded: the optimizer updates them through self._dist.sharded
        emb_params = [e.weight for e in self.embedding_dict.values()]
        if self._emb_plan is not None:
            sets.append(TableSet(self._emb_plan, emb_params, l2map.get(id(emb_params[0]), 0.0)))
        lin_params = [e.weight for e in self.linear_model.embedding_dict.values()]
        if lin_params:
            sets.append(TableSet(self.linear_model._plan, lin_params, l2map.get(id(lin_params[0]), 0.0)))
        return sets

    def _get_optim(self, optimizer):
        if not isinstance(optimizer, str):
            return optimizer
        if optimizer not in ("sgd", "adam", "adagrad", "rmsprop"):
            raise NotImplementedError
        sets = self._table_sets()
        table_ids = set(id(p) for ts in sets for p in ts.params)
        l2map, _ = self._l2_map()
        l2_sharded = (0.0, 0.0)
        if self._dist is not None:
            emb_p = [e.weight for e in self.embedding_dict.values()]
            lin_p = [e.weight for e in self.linear_model.embedding_dict.values()]
            table_ids = set(id(p) for p in emb_p + lin_p)
            l2_sharded = (l2map.get(id(emb_p[0]), 0.0), l2map.get(id(lin_p[0]), 0.0))
        dense_named = [(n, p) for n, p in self.named_parameters() if id(p) not in table_ids]
        return FusedOptimizer(optimizer, dense_named, sets, l2map, dist_ctx=self._dist, l2_sharded=l2_sharded)

    # ------------------------------------------------------------------------------------------
    # multi-GPU (replaces nn.DataParallel, basemodel.py:206-209): one process per GPU, see deepctr/distributed.py
    # ------------------------------------------------------------------------------------------
    def distribute(self, group=None, max_batch=None):
        """Collective.  Row-shard the embedding / first-order tables over the ranks of `group` (NVLink peer memory) and make
        the dense parameters data-parallel (rank 0's values are broadcast).  `max_batch` = largest per-GPU batch."""
        from .. import distributed
        if self._dist is not None:
            raise RuntimeError("distribute() was already called on this model")
        distributed.attach(self, group, max_batch)
        if self._optimizer_spec is not None:
            self.optim = self._get_optim(self._optimizer_spec)     # re-bind the optimizer to the sharded tables
        return self

    def state_dict(self, *args, **kwargs):
        """Reference key layout.  With row-sharded tables this is a COLLECTIVE: every rank must call it; the
        full tables are re-assembled from the shards (save on rank 0 only)."""
        opt = getattr(self, "optim", None)
        if isinstance(opt, FusedOptimizer):
            opt.flush()                          # postponed row updates (lazy dense-table semantics) are applied first
        sd = super().state_dict(*args, **kwargs)
        if self._dist is not None:
            from ..distributed import gather_tables
            prefix = kwargs.get("prefix", args[1] if len(args) > 1 else "")
            for k, v in gather_tables(self).items():
                sd[prefix + k] = v
        return sd

    def load_state_dict(self, state_dict, strict=True, **kwargs):
        opt = getattr(self, "optim", None)
        if isinstance(opt, FusedOptimizer):
            opt.flush()
        if self._dist is None:
            return super().load_state_dict(state_dict, strict=strict, **kwargs)
        from ..distributed import scatter_tables
        used = set(scatter_tables(self, state_dict))
        rest = {k: v for k, v in state_dict.items() if k not in used}
        for k in used:                       # placeholders keep the module's own (empty) table parameters untouched
            rest[k] = super().state_dict()[k]
        return super().load_state_dict(rest, strict=strict, **kwargs)

    def _get_loss_func(self, loss):
        if isinstance(loss, str):
            return self._get_loss_func_single(loss)
        if isinstance(loss, list):
            return [self._get_loss_func_single(l) for l in loss]
        return loss

    def _get_loss_func_single(self, loss):
        if loss == "binary_crossentropy":
            return F.binary_cross_entropy
        if loss == "mse":
            return F.mse_loss
        if loss == "mae":
            return F.l1_loss
        raise NotImplementedError

    def _log_loss(self, y_true, y_pred, eps=1e-7, normalize=True, sample_weight=None, labels=None):
        from sklearn.metrics import log_loss
        return log_loss(y_true, y_pred, eps, normalize, sample_weight, labels)

    @staticmethod
    def _accuracy_score(y_true, y_pred):
        from sklearn.metrics import accuracy_score
        return accuracy_score(y_true, np.where(y_pred > 0.5, 1, 0))

    def _get_metrics(self, metrics, set_eps=False):
        from sklearn.metrics import log_loss, mean_squared_error, roc_auc_score
        out = {}
        for metric in metrics or []:
            if metric in ("binary_crossentropy", "logloss"):
                out[metric] = self._log_loss if set_eps else log_loss
            if metric == "auc":
                out[metric] = roc_auc_score
            if metric == "mse":
                out[metric] = mean_squared_error
            if metric in ("accuracy", "acc"):
                out[metric] = self._accuracy_score
            self.metrics_names.append(metric)
        return out

    def _in_multi_worker_mode(self):
        return None

    # ------------------------------------------------------------------------------------------
    # host-side data plumbing
    # ------------------------------------------------------------------------------------------
    def _as_list(self, x):
        if isinstance(x, dict):
            x = [x[name] for name in self.feature_index]
        x = list(x)
        for i in range(len(x)):
            if len(x[i].shape) == 1:
                x[i] = np.expand_dims(x[i], axis=1)
        return x

    def _host_arrays(self, x):
        """list of per-feature arrays (feature_index order) -> pinned (ids int32 [N, m_all], dense float32 [N, nd_all])."""
        x = self._as_list(x)
        names = list(self.feature_index.keys())
        by_name = dict(zip(names, x))
        n = x[0].shape[0] if x else 0
        ids = np.empty((n, len(self._all_sparse_cols)), dtype=np.int32)
        for j, fc in enumerate(self._all_sparse):
            col = np.asarray(by_name[fc.name]).reshape(n, -1)[:, 0]
            ids[:, j] = col.astype(np.int64) if col.dtype.kind in "fc" else col   # truncation == .long()
        j = len(self._all_sparse)
        for fc in self._all_varlen:         # [n, maxlen] id matrix per multi-value feature
            a = np.asarray(by_name[fc.name]).reshape(n, -1)
            if a.shape[1] != fc.maxlen:
                raise ValueError("feature '%s': expected %d positions per sample, got %d" % (fc.name, fc.maxlen, a.shape[1]))
            ids[:, j:j + fc.maxlen] = a.astype(np.int64) if a.dtype.kind in "fc" else a
            if n and (ids[:, j:j + fc.maxlen].min() < 0 or ids[:, j:j + fc.maxlen].max() >= fc.vocabulary_size):
                raise IndexError("feature '%s': id out of range [0, %d)" % (fc.name, fc.vocabulary_size))
            j += fc.maxlen
        for name in self._len_names:
            col = np.asarray(by_name[name]).reshape(n, -1)[:, 0]
            ids[:, j] = col.astype(np.int64) if col.dtype.kind in "fc" else col
            j += 1
        dense = np.empty((n, len(self._all_dense_cols)), dtype=np.float32)
        j = 0
        for fc in self._all_dense:
            a = np.asarray(by_name[fc.name]).reshape(n, -1)
            dense[:, j:j + a.shape[1]] = a
            j += a.shape[1]
        for fc, j in zip(self._all_sparse, range(ids.shape[1])):
            if n and (ids[:, j].min() < 0 or ids[:, j].max() >= fc.vocabulary_size):
                raise IndexError("feature '%s': id out of range [0, %d)" % (fc.name, fc.vocabulary_size))
        ids_t, dense_t = torch.from_numpy(ids), torch.from_numpy(dense)
        if torch.cuda.is_available():
            ids_t, dense_t = ids_t.pin_memory(), dense_t.pin_memory()
        return ids_t, dense_t

    # Device-staged epochs: when the whole input fits comfortably in HBM the feature arrays go to the device ONCE
    # per fit() / predict() call, column by column (no [N, m] host matrix, no pinned staging copy), ids are range-checked there, and
    # every batch is a device-side gather (shuffle) or slice.  Larger inputs keep the pinned double-buffered host feeder.
    STAGE_FRACTION = 0.25          # of the currently free device memory

    def _can_stage(self, n):
        if os.environ.get("XDFM_DEVICE_STAGING", "1") == "0" or torch.device(self.device).type != "cuda":
            return False
        need = n * (len(self._all_sparse_cols) * 4 + len(self._all_dense_cols) * 4 + 8) * 2      # arrays + one conversion temporary
        free, _ = torch.cuda.mem_get_info(torch.device(self.device))
        return need <= self.STAGE_FRACTION * free

    def _device_arrays(self, x):
        """list / dict of per-feature arrays -> device (ids int32 [N, m_all], dense float32 [N, nd_all]); same column layout, id
        truncation (== .long()) and IndexError behaviour as _host_arrays."""
        x = self._as_list(x)
        names = list(self.feature_index.keys())
        by_name = dict(zip(names, x))
        n = x[0].shape[0] if x else 0
        dev = torch.device(self.device)
        ids = torch.empty((n, len(self._all_sparse_cols)), dtype=torch.int32, device=dev)
        dense = torch.empty((n, len(self._all_dense_cols)), dtype=torch.float32, device=dev)
        checks = []            # (feature name, vocabulary size, device [2] = (min, max))

        def put_ids(j, arr, width, fc):
            a = np.asarray(arr).reshape(n, -1)
            if width == 1:
                a = a[:, :1]
            elif a.shape[1] != width:
                raise ValueError("feature '%s': expected %d positions per sample, got %d" % (fc.name, width, a.shape[1]))
            t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            if t.is_floating_point():
                t = t.long()                                   # truncation == .long() (basemodel.py:368-370)
            if n and fc is not None:
                checks.append((fc.name, fc.vocabulary_size, torch.stack([t.min(), t.max()]).long()))
            ids[:, j:j + width] = t
            return j + width

        j = 0
        for fc in self._all_sparse:
            j = put_ids(j, by_name[fc.name], 1, fc)
        for fc in self._all_varlen:
            j = put_ids(j, by_name[fc.name], fc.maxlen, fc)
        for name in self._len_names:
            j = put_ids(j, by_name[name], 1, None)
        j = 0
        for fc in self._all_dense:
            a = np.asarray(by_name[fc.name]).reshape(n, -1)
            dense[:, j:j + a.shape[1]] = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            j += a.shape[1]
        if checks:
            mm = torch.stack([c[2] for c in checks]).cpu()     # one sync for all features
            for (name, vocab, _), (lo, hi) in zip(checks, mm.tolist()):
                if lo < 0 or hi >= vocab:
                    raise IndexError("feature '%s': id out of range [0, %d)" % (name, vocab))
        return ids, dense

    def _device_batches(self, ids, dense, y, batch_size, order=None):
        """Batches of device-staged arrays: slices, or gathers through the (device copy of the) epoch permutation."""
        n = ids.shape[0] if order is None else order.shape[0]
        steps = (n - 1) // batch_size + 1 if n > 0 else 0
        for i in range(steps):
            lo, hi = i * batch_size, min(n, (i + 1) * batch_size)
            if order is None:
                yield ids[lo:hi], dense[lo:hi], None if y is None else y[lo:hi]
            else:
                idx = order[lo:hi]
                yield ids.index_select(0, idx), dense.index_select(0, idx), None if y is None else y.index_select(0, idx)

    def _dist_local_order(self, order, sample_num, batch_size):
        """Row indices this rank trains on, in step order: its slice of every global batch (identical permutation on all ranks)."""
        from ..distributed import rank_slice
        ctx = self._dist
        if order is not None and ctx.world > 1:
            o = order.to(torch.device(self.device))
            ctx.broadcast(o, 0)
            order = o.cpu()
        gbs = batch_size * ctx.world
        parts = []
        for lo in range(0, sample_num, gbs):
            a, b = rank_slice(lo, min(sample_num, lo + gbs), ctx.rank, ctx.world)
            parts.append(torch.arange(a, b) if order is None else order[a:b])
        return torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64)

    def _batches(self, ids, dense, y, batch_size, order=None):
        """Yield device batches (ids, dense, y or None); H2D copies run one batch ahead on a side stream."""
        dev = torch.device(self.device)
        n = ids.shape[0] if order is None else order.shape[0]
        steps = (n - 1) // batch_size + 1 if n > 0 else 0
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        stage = [None, None]

        def launch(i):
            lo, hi = i * batch_size, min(n, (i + 1) * batch_size)
            slot = i & 1
            if order is None:
                h = (ids[lo:hi], dense[lo:hi], None if y is None else y[lo:hi])
            else:
                idx = order[lo:hi]
                if stage[slot] is None:
                    stage[slot] = (torch.empty((batch_size, ids.shape[1]), dtype=ids.dtype).pin_memory(),
                                   torch.empty((batch_size, dense.shape[1]), dtype=dense.dtype).pin_memory(),
                                   None if y is None else torch.empty((batch_size,) + tuple(y.shape[1:]), dtype=y.dtype).pin_memory(),
                                   [None])
                si, sd, sy, ev = stage[slot]
                if ev[0] is not None:
                    ev[0].synchronize()
                k = hi - lo
                torch.index_select(ids, 0, idx, out=si[:k])
                torch.index_select(dense, 0, idx, out=sd[:k])
                if y is not None:
                    torch.index_select(y, 0, idx, out=sy[:k])
                h = (si[:k], sd[:k], None if y is None else sy[:k])
            with torch.cuda.stream(copy_stream):
                d = tuple(None if t is None else t.to(dev, non_blocking=True) for t in h)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            if order is not None:
                stage[slot][3][0] = ev
            return d, ev

        nxt = launch(0) if steps > 0 else None
        for i in range(steps):
            cur, ev = nxt
            nxt = launch(i + 1) if i + 1 < steps else None
            main.wait_event(ev)
            for t in cur:
                if t is not None:
                    t.record_stream(main)
            yield cur

    # ------------------------------------------------------------------------------------------
    # fit / evaluate / predict
    # ------------------------------------------------------------------------------------------
    def _fused_ok(self):
        _, has_l1 = self._l2_map()
        aux_zero = (not torch.is_tensor(self.aux_loss)) or (not self.aux_loss.requires_grad and float(self.aux_loss.abs().sum()) == 0.0)
        return isinstance(self.optim, FusedOptimizer) and not has_l1 and aux_zero and not isinstance(self.loss_func, list)

    def train_step(self, ids, dense, y, loss_accum, pred_log=None, pred_off=0, host_labels=None):
        """One fused training step on device tensors: forward, loss, backward, optimizer (+L2).  No host sync.
        `host_labels` (optional): the same labels as a host tensor / array, for models that size work from them (xDeepFM Pro).

        The step issues ~120 kernel launches from Python (3-3.5 ms of host time at BASELINE config 2, about the GPU time of the
        step itself), so after two eager steps with the same shapes and hyper-parameters it is captured into a CUDA graph and
        replayed: inputs are copied into the graph's static buffers, the loss comes back through a static accumulator.
        `model.use_cuda_graph = False` (or XDFM_CUDA_GRAPH=0) keeps every step eager."""
        if host_labels is not None:
            self._host_label_hint(host_labels)
        if self.use_cuda_graph and ids.shape[0] > 0:
            out = self._train_step_graphed(ids, dense, y, loss_accum, pred_log, pred_off)
            if out is not None:
                return out
        return self._train_step_eager(ids, dense, y, loss_accum, pred_log, pred_off)

    # ---- CUDA-graph replay of the fused step ------------------------------------------------------------
    def _graph_key(self, ids, dense, y):
        opt = self.optim
        hp = tuple((k, v) for k, v in sorted(opt.param_groups[0].items()) if isinstance(v, (int, float, tuple)))
        prec = tuple(getattr(getattr(self, n, None), "precision", None) for n in ("cin", "dnn"))
        flat = opt._flat["w"].data_ptr() if opt._flat is not None else 0
        return (tuple(ids.shape), tuple(dense.shape), tuple(y.shape), id(opt), hp, prec, opt._hist_base, opt.lazy_tables,
                opt.sparse_embedding_update, self.training, getattr(self, "sfg_weight", None), flat, ops.workspace_generation())

    def _train_step_graphed(self, ids, dense, y, loss_accum, pred_log, pred_off):
        opt = self.optim
        if not isinstance(opt, FusedOptimizer) or self._graph_failed or ops.TIMERS is not None:
            return None
        if opt._lazy_active() and opt.steps + 2 - opt._hist_base >= opt._hist_cap:
            return None                      # the history window is about to be rebased: take the eager path for that step
        key = self._graph_key(ids, dense, y)
        st = self._graphs.get(key)
        if st is None:
            seen = self._graph_seen.get(key, 0)
            self._graph_seen[key] = seen + 1
            if seen < 2:
                return None                  # eager warm-up (allocators, workspaces, NCCL communicators, lazy-table state)
            try:
                st = self._capture_step(ids, dense, y)
            except Exception as e:           # pragma: no cover - depends on driver / torch build
                import warnings
                self._recover_after_failed_capture(ids.device)
                if os.environ.get("XDFM_GRAPH_DEBUG"):
                    raise
                warnings.warn("CUDA-graph capture of the training step failed (%s); staying on eager launches" % (e,))
                self._graph_failed = True
                return None
            if key != self._graph_key(ids, dense, y):
                # the capture itself grew a cached workspace (first use of a larger shape): its graph holds addresses that are
                # still valid (the NEW buffers), but every older graph is stale -- drop them, keep this one under the new key
                self._graphs.clear()
                key = self._graph_key(ids, dense, y)
            if len(self._graphs) >= 4:       # a handful of batch shapes at most (full batch, last partial batch)
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = st
        st["ids"].copy_(ids, non_blocking=True)
        st["dense"].copy_(dense, non_blocking=True)
        st["y"].copy_(y.reshape(st["y"].shape), non_blocking=True)
        st["graph"].replay()
        if st.get("graph_apply") is not None:
            # multi-GPU: the two NCCL collectives stay outside the graphs (captured collectives can hang at process-group teardown)
            opt.step_exchange()
            st["graph_apply"].replay()
            opt.step_barrier()
        opt.steps += 1
        if opt._lazy_active():
            opt._dirty = True
            opt.maybe_flush()
        loss_accum += st["loss"]
        if pred_log is not None:
            pred_log[pred_off:pred_off + ids.shape[0]] = st["y_pred"].detach().reshape(-1)
        return st["y_pred"]

    @staticmethod
    def _recover_after_failed_capture(device):
        """A capture that dies in capture_end() leaves torch's CUDA generator flagged as 'capturing' (every later randn / dropout on
        the device raises 'Offset increment outside graph capture').  A trivial successful capture runs the generator's capture
        prologue / epilogue pair again and clears the flag."""
        try:
            torch.cuda.synchronize(device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                torch.zeros(1, device=device)
            del g
            torch.cuda.synchronize(device)
        except Exception:                    # pragma: no cover
            pass

    def _capture_step(self, ids, dense, y):
        opt = self.optim
        st = {"ids": ids.clone(), "dense": dense.clone(), "y": y.clone(),
              "loss": torch.zeros(1, dtype=torch.float64, device=ids.device)}
        steps0, dirty0 = opt.steps, opt._dirty
        opt._dirty = True                    # the captured step always runs the (idempotent) catch-up of the looked-up rows
        torch.cuda.synchronize(ids.device)
        graph = torch.cuda.CUDAGraph()
        self._capturing = True
        try:
            return self._capture_graphs(opt, st, graph, steps0, dirty0)
        finally:
            self._capturing = False

    def _capture_graphs(self, opt, st, graph, steps0, dirty0):
        if self._dist is None:
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                st["loss"].zero_()
                st["y_pred"] = self._train_step_eager(st["ids"], st["dense"], st["y"], st["loss"])
            st["graph_apply"] = None
        else:
            # graph 1: forward, loss, backward, batch side of the sharded scatter-add; graph 2: dense + shard optimizer kernels
            self._capturing_half = True
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    st["loss"].zero_()
                    st["y_pred"] = self._train_step_eager(st["ids"], st["dense"], st["y"], st["loss"])
            finally:
                self._capturing_half = False
            graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph2, pool=graph.pool(), capture_error_mode="thread_local"):
                opt.step_apply(apply_l2=True)
            st["graph_apply"] = graph2
        opt.steps, opt._dirty = steps0, dirty0          # capturing does not execute the step
        st["graph"] = graph
        return st

    def _train_step_eager(self, ids, dense, y, loss_accum, pred_log=None, pred_off=0):
        opt = self.optim
        opt.prepare()       # flat parameter / gradient views must exist BEFORE backward accumulates into them
        opt.zero_grad()
        self._seg_cache.clear()     # the key holds a device address: never let a recycled allocation hit a previous batch's segments
        if self._dist is not None:
            self._dist.ensure_capacity(ids.shape[0] * self._dist.sharded.m)
            self._dist.sharded.stash = {}
            if ids.shape[0] == 0:           # this rank has no rows in the last partial batch: take part in the collectives only
                self._empty_step_collectives()
                opt.step(apply_l2=True)
                return None
        for ts in opt.table_sets:
            ts.plan.sparse_grad = True      # backward leaves (unique rows, segment sums) for the fused optimizer
        self._in_train_step = True
        ops.direct_param_grads(True)        # p.grad views were just zeroed: the hot layers write their gradients in place
        try:
            return self._train_step_inner(opt, ids, dense, y, loss_accum, pred_log, pred_off)
        finally:
            ops.direct_param_grads(False)
            self._in_train_step = False
            for ts in opt.table_sets:
                ts.plan.sparse_grad = False

    def _train_step_inner(self, opt, ids, dense, y, loss_accum, pred_log, pred_off):
        y_pred = self.forward_ids(ids, dense)
        yv = y.reshape(-1)
        if self._loss_name == "binary_crossentropy":
            _, dy = ops.bce_sum(y_pred, yv, loss_accum)
            y_pred.backward(dy.view_as(y_pred))
        else:
            loss = self.loss_func(y_pred.reshape(-1), yv, reduction="sum")
            loss_accum += loss.detach().double()
            loss.backward()
        self._optimizer_phases(opt)
        y_pred = y_pred.detach()             # the caller must not keep the step's autograd graph alive (see BaseModelSFG)
        if pred_log is not None:
            pred_log[pred_off:pred_off + yv.shape[0]] = y_pred.reshape(-1)
        return y_pred

    def _empty_step_collectives(self):
        """Hook: collectives a subclass issues inside its step body (a rank without rows must still take part)."""

    def _epoch_begin(self):
        """Hook for subclasses with extra device-side accumulators (xDeepFM Pro)."""

    def _epoch_extra(self, sample_num):
        """Hook: (extra sum added to the epoch's total loss, extra History entries)."""
        return 0.0, {}

    def _optimizer_phases(self, opt):
        """opt.step(apply_l2=True), or only its first phase while the first half of a multi-GPU step is being captured."""
        opt.step_local()
        if self._capturing_half:
            return
        opt.step_exchange()
        opt.step_apply(apply_l2=True)
        opt.step_barrier()
        if not self._capturing:
            opt.maybe_flush()

    def train_on_batch(self, ids, dense, y):
        """Public single-step API: HOST tensors (ids int32 [B, m_all], dense float32 [B, nd_all], y float32 [B]; pinned
        memory makes the copies asynchronous) -> python float BCE-sum of the batch.  H2D copy, fused step, D2H of the loss."""
        dev = torch.device(self.device)
        ids_d = ids.to(dev, non_blocking=True)
        dense_d = dense.to(dev, non_blocking=True)
        y_d = y.to(dev, non_blocking=True)
        if not hasattr(self, "_tob_accum") or self._tob_accum.device != dev:
            self._tob_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        self._tob_accum.zero_()
        self._host_label_hint(y)
        self.train_step(ids_d, dense_d, y_d, self._tob_accum)
        return float(self._tob_accum.item())

    def _host_label_hint(self, y_host):
        """Hook: the labels of the batch about to be stepped are still on the host (xDeepFM Pro sizes its positive-rows-only SFG
        pass from their count without a device sync).  Consumed by the next train_step."""

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, initial_epoch=0, validation_split=0.,
            validation_data=None, shuffle=True, callbacks=None):
        x = self._as_list(x)
        do_validation = False
        val_x, val_y = [], []
        if validation_data:
            do_validation = True
            if len(validation_data) == 2:
                val_x, val_y = validation_data
            elif len(validation_data) == 3:
                val_x, val_y, _ = validation_data
            else:
                raise ValueError("When passing a `validation_data` argument, it must contain either 2 items (x_val, y_val), "
                                 "or 3 items (x_val, y_val, val_sample_weights)")
            val_x = self._as_list(val_x)
        elif validation_split and 0. < validation_split < 1.:
            do_validation = True
            split_at = int(x[0].shape[0] * (1. - validation_split))
            x, val_x = slice_arrays(x, 0, split_at), slice_arrays(x, split_at)
            y, val_y = slice_arrays(y, 0, split_at), slice_arrays(y, split_at)
            if not isinstance(x, list):
                x, val_x = [x], [val_x]
        if batch_size is None:
            batch_size = 256
        if self.gpus and len(self.gpus) > 1:
            raise NotImplementedError(
                "single-process DataParallel (`gpus=[...]`) is replaced by one process per GPU: launch with torchrun and "
                "use deepctr.distributed (see INTEGRATION.md)")
        if torch.device(self.device).type != "cuda":
            raise RuntimeError("fit(): the xdeepfm-b200 path needs device='cuda:N' (sm_100a); no CPU fallback")

        staged = self._dist is None and self._can_stage(x[0].shape[0] if x else 0)
        ids, dense = self._device_arrays(x) if staged else self._host_arrays(x)
        y_t = torch.from_numpy(np.ascontiguousarray(np.asarray(y, dtype=np.float32)).reshape(ids.shape[0], -1))
        if y_t.shape[1] == 1:
            y_t = y_t.reshape(-1)
        y_dev = y_t.to(torch.device(self.device)) if staged else None      # labels stay on the host too (xDeepFM Pro counts positives there)
        if not staged:
            y_t = y_t.pin_memory()
        sample_num = ids.shape[0]
        steps_per_epoch = (sample_num - 1) // batch_size + 1
        ctx = self._dist
        if ctx is not None:
            # one process per GPU: `batch_size` is per GPU (as with the reference's `gpus=`, basemodel.py:209); every rank holds
            # the same arrays and trains on its slice of each global batch of batch_size * world samples
            if not self._fused_ok():
                raise NotImplementedError("distributed fit() needs a named optimizer ('sgd'/'adam'/'adagrad'/'rmsprop'), a single "
                                          "loss and no l1 / auxiliary loss")
            steps_per_epoch = (sample_num - 1) // (batch_size * ctx.world) + 1
            ctx.ensure_capacity(batch_size * ctx.sharded.m)
        dev = torch.device(self.device)
        self.train()
        fused = self._fused_ok()

        callbacks = CallbackList((callbacks or []) + [self.history])
        callbacks.set_model(self)
        callbacks.on_train_begin()
        callbacks.set_model(self)
        self.stop_training = False

        print("Train on {0} samples, validate on {1} samples, {2} steps per epoch".format(sample_num, len(val_y), steps_per_epoch))
        loss_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        total_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        for epoch in range(initial_epoch, epochs):
            callbacks.on_epoch_begin(epoch)
            epoch_logs = {}
            start_time = time.time()
            loss_accum.zero_()
            total_accum.zero_()
            self._epoch_begin()
            if fused:
                self.optim.prepare()
                self.optim.reg_accum.zero_()
            order = torch.randperm(sample_num) if shuffle else None
            local_steps = steps_per_epoch
            if ctx is not None:
                order = self._dist_local_order(order, sample_num, batch_size)
                local_steps = (order.shape[0] - 1) // batch_size + 1 if order.shape[0] > 0 else 0
            n_local = sample_num if order is None else order.shape[0]
            want_metrics = verbose > 0 and len(self.metrics) > 0
            pred_log = torch.empty(n_local, dtype=torch.float32, device=dev) if want_metrics else None
            off = 0
            # NB: like the reference (basemodel.py:202, :332 / basemodel_sfg.py:275, :488) the model is put in train mode ONCE
            # before the epoch loop; the validation predict() leaves it in eval mode, so from the second epoch on dropout -- and
            # the SFG term of xDeepFM Pro, which is gated on self.training (xdeepfm_pro.py:265) -- are off when validation data
            # is given.  Kept for result parity (tests/golden/fit_pro_small_adam.npz: reference History['sfg_loss'] = [0.41, 0.0]).
            feed = self._device_batches(ids, dense, y_dev, batch_size, None if order is None else order.to(dev)) if staged else \
                self._batches(ids, dense, y_t, batch_size, order)
            for ids_b, dense_b, y_b in feed:
                nb = ids_b.shape[0]
                if fused:
                    if ctx is None:
                        self._host_label_hint(y_t[off:off + nb] if order is None else y_t[order[off:off + nb]])
                    self.train_step(ids_b, dense_b, y_b, loss_accum, pred_log, off)
                else:
                    # generic path (user-supplied optimizer / loss list / l1 / aux loss): mirrors basemodel.py:245-262
                    y_pred = self.forward_ids(ids_b, dense_b).squeeze()
                    self.optim.zero_grad()
                    if isinstance(self.loss_func, list):
                        loss = sum(self.loss_func[i](y_pred[:, i], y_b[:, i], reduction="sum") for i in range(len(self.loss_func)))
                    else:
                        loss = self.loss_func(y_pred, y_b.squeeze(), reduction="sum")
                    total = loss + self.get_regularization_loss() + self.aux_loss
                    loss_accum += loss.detach().double()
                    total_accum += total.detach().double().reshape(-1)
                    total.backward()
                    self.optim.step()
                    if pred_log is not None:
                        pred_log[off:off + nb] = y_pred.detach().reshape(-1)
                off += nb
            if ctx is not None:
                for _ in range(steps_per_epoch - local_steps):      # ranks without rows in the last partial batch
                    self.train_step(ids[:0].to(dev), dense[:0].to(dev), y_t[:0].to(dev), loss_accum)
                ctx.all_reduce_sum(loss_accum)
            # ---- one host sync per epoch
            if fused:
                total_loss_epoch = float(loss_accum.item()) + self.optim.pop_reg_loss()
            else:
                total_loss_epoch = float(total_accum.item())
            extra_total, extra_logs = self._epoch_extra(sample_num)
            epoch_logs["loss"] = (total_loss_epoch + extra_total) / sample_num
            epoch_logs.update(extra_logs)
            if want_metrics:
                # per-step metrics of the epoch from the device-resident prediction log: float64 on the device (metrics_device.py);
                # metric functions it does not know, or steps sklearn would reject, go through the host exactly like the reference
                from ..metrics_device import step_metrics
                y_epoch = y_t if order is None else y_t[order]
                dev_vals = step_metrics(self, self.metrics, pred_log, y_epoch.to(dev), batch_size) if n_local > 0 else None
                if dev_vals is None:
                    pred_host = pred_log.cpu().numpy().astype("float64")
                    y_host = y_epoch.numpy()
                for name, fn in self.metrics.items():
                    if dev_vals is not None:
                        vals = dev_vals[name].cpu().numpy().tolist()
                    else:
                        vals = []
                        for s in range(local_steps):
                            lo, hi = s * batch_size, min(n_local, (s + 1) * batch_size)
                            vals.append(fn(y_host[lo:hi], pred_host[lo:hi]))
                    if ctx is not None:
                        # per-step metrics of the LOCAL sub-batches, averaged over all ranks' steps
                        t = torch.tensor([float(np.sum(vals)), float(len(vals))], dtype=torch.float64, device=dev)
                        ctx.all_reduce_sum(t)
                        epoch_logs[name] = float(t[0].item() / max(t[1].item(), 1.0))
                    else:
                        epoch_logs[name] = np.sum(vals) / steps_per_epoch
            if do_validation:
                for name, result in self.evaluate(val_x, val_y, batch_size).items():
                    epoch_logs["val_" + name] = result
            if verbose > 0:
                epoch_time = int(time.time() - start_time)
                print("Epoch {0}/{1}".format(epoch + 1, epochs))
                msg = "{0}s - loss: {1: .4f}".format(epoch_time, epoch_logs["loss"])
                for name in self.metrics:
                    msg += " - " + name + ": {0: .4f}".format(epoch_logs[name])
                if do_validation:
                    for name in self.metrics:
                        msg += " - val_" + name + ": {0: .4f}".format(epoch_logs["val_" + name])
                print(msg)
            callbacks.on_epoch_end(epoch, epoch_logs)
            if self.stop_training:
                break
        callbacks.on_train_end()
        return self.history

    def evaluate(self, x, y, batch_size=256):
        pred = self.predict(x, batch_size)
        return {name: fn(y, pred) for name, fn in self.metrics.items()}

    def predict(self, x, batch_size=256):
        """Batched no-grad forward; float64 [N, 1] numpy like the reference (basemodel.py:325-352), one D2H at the end."""
        if torch.device(self.device).type != "cuda":
            raise RuntimeError("predict(): the xdeepfm-b200 path needs device='cuda:N' (sm_100a); no CPU fallback")
        self.eval()
        x = self._as_list(x)
        staged = self._can_stage(x[0].shape[0] if x else 0)
        ids, dense = self._device_arrays(x) if staged else self._host_arrays(x)
        n = ids.shape[0]
        out = torch.empty((n, 1), dtype=torch.float32, device=torch.device(self.device))
        off = 0
        with torch.no_grad():
            feed = self._device_batches(ids, dense, None, batch_size) if staged else self._batches(ids, dense, None, batch_size, None)
            for ids_b, dense_b, _ in feed:
                yb = self.forward_ids(ids_b, dense_b)
                out[off:off + yb.shape[0]] = yb.reshape(-1, 1)
                off += yb.shape[0]
        return out.cpu().numpy().astype("float64")
