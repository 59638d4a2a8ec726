"""Feature-column API (drop-in for the reference's deepctr/inputs.py:20-245).

Same names, fields, defaults and column-layout rule as the reference so that `xdftrain*.py` and user code keep
working; the lookups themselves are done by the fused CUDA gather in `ops.py`, not by per-field nn.Embedding calls.
"""
from collections import OrderedDict, namedtuple

import torch
import torch.nn as nn

DEFAULT_GROUP_NAME = "default_group"


class SparseFeat(namedtuple("SparseFeat", ["name", "vocabulary_size", "embedding_dim", "use_hash", "dtype",
                                           "embedding_name", "group_name"])):
    """Categorical feature column (reference: inputs.py:20-38)."""
    __slots__ = ()

    def __new__(cls, name, vocabulary_size, embedding_dim=4, use_hash=False, dtype="int32", embedding_name=None,
                group_name=DEFAULT_GROUP_NAME):
        if embedding_dim == "auto":
            embedding_dim = 6 * int(pow(vocabulary_size, 0.25))
        if use_hash:
            print("Notice! Feature Hashing on the fly currently is not supported in torch version,"
                  "you can use tensorflow version!")
        return super().__new__(cls, name, vocabulary_size, embedding_dim, use_hash, dtype,
                               name if embedding_name is None else embedding_name, group_name)

    def __hash__(self):
        return hash(self.name)


class VarLenSparseFeat(namedtuple("VarLenSparseFeat", ["sparsefeat", "maxlen", "combiner", "length_name"])):
    """Variable-length (multi-value) categorical column (reference: inputs.py:41-77): `maxlen` id columns, pooled with
    `combiner` ('sum' | 'mean' | 'max') under the id != 0 mask, or under positions < `length_name` column when given.
    Looked up as `maxlen` slots of the fused gather and reduced by the bag-pooling kernel (csrc/bag.cu)."""
    __slots__ = ()

    def __new__(cls, sparsefeat, maxlen, combiner="mean", length_name=None):
        return super().__new__(cls, sparsefeat, maxlen, combiner, length_name)

    name = property(lambda self: self.sparsefeat.name)
    vocabulary_size = property(lambda self: self.sparsefeat.vocabulary_size)
    embedding_dim = property(lambda self: self.sparsefeat.embedding_dim)
    use_hash = property(lambda self: self.sparsefeat.use_hash)
    dtype = property(lambda self: self.sparsefeat.dtype)
    embedding_name = property(lambda self: self.sparsefeat.embedding_name)
    group_name = property(lambda self: self.sparsefeat.group_name)

    def __hash__(self):
        return hash(self.name)


class DenseFeat(namedtuple("DenseFeat", ["name", "dimension", "dtype"])):
    """Numeric feature column (reference: inputs.py:80-87)."""
    __slots__ = ()

    def __new__(cls, name, dimension=1, dtype="float32"):
        return super().__new__(cls, name, dimension, dtype)

    def __hash__(self):
        return hash(self.name)


def build_input_features(feature_columns):
    """OrderedDict name -> (start, end) column range of the flat input matrix, first occurrence wins
    (reference: inputs.py:99-123)."""
    index = OrderedDict()
    cursor = 0
    for fc in feature_columns:
        if fc.name in index:
            continue
        if isinstance(fc, SparseFeat):
            width = 1
        elif isinstance(fc, DenseFeat):
            width = fc.dimension
        elif isinstance(fc, VarLenSparseFeat):
            width = fc.maxlen
        else:
            raise TypeError("Invalid feature column type,got", type(fc))
        index[fc.name] = (cursor, cursor + width)
        cursor += width
        if isinstance(fc, VarLenSparseFeat) and fc.length_name is not None and fc.length_name not in index:
            index[fc.length_name] = (cursor, cursor + 1)
            cursor += 1
    return index


def get_feature_names(feature_columns):
    return list(build_input_features(feature_columns).keys())


def sparse_columns(feature_columns):
    return [fc for fc in feature_columns if isinstance(fc, SparseFeat)] if feature_columns else []


def dense_columns(feature_columns):
    return [fc for fc in feature_columns if isinstance(fc, DenseFeat)] if feature_columns else []


def varlen_columns(feature_columns):
    return [fc for fc in feature_columns if isinstance(fc, VarLenSparseFeat)] if feature_columns else []


_DEFERRED_TABLES = [False]


class deferred_tables:
    """`with deferred_tables(): model = xDeepFM(...)` builds the model WITHOUT materialising its embedding tables (1-row placeholders
    that remember their vocabulary size): for vocabularies that do not fit one GPU (BASELINE config 5: tables up to 50 M rows x 64).
    `model.distribute()` then initialises every rank's row shard in place, N(0, init_std) like create_embedding_matrix."""

    def __enter__(self):
        self._prev = _DEFERRED_TABLES[0]
        _DEFERRED_TABLES[0] = True
        return self

    def __exit__(self, *exc):
        _DEFERRED_TABLES[0] = self._prev
        return False


def table_rows(emb):
    """Rows of an embedding table module (the vocabulary size, also for deferred placeholders)."""
    return int(getattr(emb, "deferred_rows", emb.weight.shape[0]))


def create_embedding_matrix(feature_columns, init_std=0.0001, linear=False, sparse=False, device="cpu"):
    """nn.ModuleDict {embedding_name: nn.Embedding(vocab, D or 1)} initialised N(0, init_std)
    (reference: inputs.py:158-180).  The modules are parameter containers: the CUDA gather reads `.weight` directly."""
    tables = nn.ModuleDict()
    for fc in sparse_columns(feature_columns) + varlen_columns(feature_columns):
        if fc.embedding_name not in tables:
            if _DEFERRED_TABLES[0]:
                emb = nn.Embedding(1, 1 if linear else fc.embedding_dim, sparse=sparse)
                emb.deferred_rows = int(fc.vocabulary_size)
                emb.init_std = init_std
                tables[fc.embedding_name] = emb
            else:
                tables[fc.embedding_name] = nn.Embedding(fc.vocabulary_size, 1 if linear else fc.embedding_dim, sparse=sparse)
    for emb in tables.values():
        nn.init.normal_(emb.weight, mean=0, std=init_std)
    return tables.to(device)


def combined_dnn_input(sparse_embedding_list, dense_value_list):
    """Flatten + concat: all sparse embeddings (field order) then the dense values (reference: inputs.py:126-138)."""
    parts = []
    if len(sparse_embedding_list) > 0:
        parts.append(torch.flatten(torch.cat(sparse_embedding_list, dim=-1), start_dim=1))
    if len(dense_value_list) > 0:
        parts.append(torch.flatten(torch.cat(dense_value_list, dim=-1), start_dim=1))
    if not parts:
        raise NotImplementedError
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)


def varlen_embedding_lookup(X, embedding_dict, sequence_input_dict, varlen_sparse_feature_columns):
    """{feature name: [B, maxlen, D]} sequence embeddings (reference: inputs.py:212-225), one fused gather per feature."""
    from . import ops
    out = {}
    for fc in varlen_sparse_feature_columns:
        a, b = sequence_input_dict[fc.name]
        ids, _ = ops.split_input(X, list(range(a, b)), [])
        w = embedding_dict[fc.embedding_name].weight
        plan = ops.SparsePlan([0] * (b - a), [w.shape[0]], w.shape[1])
        out[fc.name] = ops.SparseGather.apply(plan, ops.SegmentCache(), ids, w)
    return out


def get_varlen_pooling_list(embedding_dict, features, feature_index, varlen_sparse_feature_columns, device):
    """[B, 1, D] per sequence feature (reference: inputs.py:141-155).  `embedding_dict` = the result of varlen_embedding_lookup."""
    from .layers.sequence import SequencePoolingLayer
    out = []
    for fc in varlen_sparse_feature_columns:
        seq_emb = embedding_dict[fc.name]
        if fc.length_name is None:
            a, b = feature_index[fc.name]
            mask = features[:, a:b].long() != 0
            out.append(SequencePoolingLayer(mode=fc.combiner, supports_masking=True, device=device)([seq_emb, mask]))
        else:
            a, b = feature_index[fc.length_name]
            length = features[:, a:b].long()
            out.append(SequencePoolingLayer(mode=fc.combiner, supports_masking=False, device=device)([seq_emb, length]))
    return out
