"""Drop-in surface of the `deepctr` package (no GPU): the imports the reference's CLI scripts make (xdftrain.py:22-24,
xdftrain_attn.py:26-28, xdftrain_pro.py:28-30), constructor signatures, state_dict keys and the Python-level error behaviour."""
import inspect

import pytest
import torch

from oracle import xdeepfm_oracle as O
from tests.helpers import build_product_model


def test_reference_script_imports_resolve():
    from deepctr.callbacks import EarlyStopping, History, ModelCheckpoint  # noqa: F401
    from deepctr.inputs import DenseFeat, SparseFeat, VarLenSparseFeat, build_input_features, combined_dnn_input, get_feature_names  # noqa: F401
    from deepctr.inputs import create_embedding_matrix, get_varlen_pooling_list, varlen_embedding_lookup  # noqa: F401
    from deepctr.layers import CIN, DNN, PredictionLayer, SequencePoolingLayer  # noqa: F401
    from deepctr.layers.sequence import SequencePoolingLayer as _SPL  # noqa: F401
    from deepctr.layers.cin_attention import AttentionPooling, CINAttention, CINAttentionV2, MultiHeadSelfAttention  # noqa: F401
    from deepctr.models import xDeepFM, xDeepFMAttention, xDeepFMAttentionV2  # noqa: F401
    from deepctr.xdeepfm_pro import (AutoDisLayer, BaseModelSFG, DenseFeatureEncoder, LabelAwareAttention, SFGDecoder, SFGLoss,  # noqa: F401
                                     xDeepFMPro, xDeepFMProLight)


def test_constructor_signatures_match_the_reference():
    from deepctr.models import xDeepFM, xDeepFMAttention, xDeepFMAttentionV2
    from deepctr.xdeepfm_pro import xDeepFMPro, xDeepFMProLight
    base = ["linear_feature_columns", "dnn_feature_columns", "dnn_hidden_units", "cin_layer_size", "cin_split_half", "cin_activation"]
    tail = ["l2_reg_linear", "l2_reg_embedding", "l2_reg_dnn", "l2_reg_cin", "init_std", "seed", "dnn_dropout", "dnn_activation",
            "dnn_use_bn", "task", "device", "gpus"]

    def names(cls):
        return [p for p in inspect.signature(cls.__init__).parameters if p != "self"]
    assert names(xDeepFM) == base + tail                                                                      # xdeepfm.py:42-45
    attn = ["cin_num_heads", "cin_attn_dropout", "cin_use_layer_norm", "cin_use_residual"]
    assert names(xDeepFMAttention) == base + attn + tail                                                      # xdeepfm_attn.py:55-62
    assert names(xDeepFMAttentionV2) == base + attn + ["cin_num_attn_layers"] + tail                          # xdeepfm_attn.py:185-193
    pro = ["use_sfg", "sfg_weight", "sfg_hidden_units", "sfg_dropout", "sfg_positive_only", "sfg_use_label_attention", "use_autodis",
           "autodis_buckets", "autodis_temperature"]
    assert names(xDeepFMPro) == base + tail + pro                                                             # xdeepfm_pro.py:57-87
    assert names(xDeepFMProLight) == base + tail + pro
    d = inspect.signature(xDeepFM.__init__).parameters
    assert d["dnn_hidden_units"].default == (256, 256) and d["cin_layer_size"].default == (256, 128) and d["l2_reg_linear"].default == 1e-5
    assert inspect.signature(xDeepFMProLight.__init__).parameters["sfg_weight"].default == 0.05


@pytest.mark.parametrize("variant", ["xdeepfm", "attn", "attn_v2", "pro", "pro+autodis"])
def test_state_dict_keys_and_shapes_match_the_reference_layout(variant):
    autodis = variant.endswith("+autodis")
    spec = O.ModelSpec(sparse_names=["C1", "C2", "C3"], vocab_sizes=[11, 5, 40], embedding_dim=8, dense_names=["I1", "I2"],
                       cin_layer_size=(8, 4), dnn_hidden_units=(8, 6), variant=variant.split("+")[0], num_heads=2, num_attn_layers=2,
                       sfg_hidden_units=(8, 4), use_autodis=autodis, autodis_buckets=6)
    model = build_product_model(spec, "cpu")
    want = O.param_shapes(spec)                                   # SURVEY.md 8a-K, probed on the reference
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}


def test_python_level_errors_mirror_the_reference():
    from deepctr.inputs import DenseFeat, SparseFeat
    from deepctr.layers import CIN
    from deepctr.models import xDeepFM
    with pytest.raises(ValueError):
        CIN(3, ())                                                # interaction.py:178-180
    with pytest.raises(ValueError):
        CIN(3, (3, 2), split_half=True)                           # interaction.py:195-197
    cols = [SparseFeat("C1", 10, 4), DenseFeat("I1", 1)]
    with pytest.raises(ValueError):
        xDeepFM(cols, cols, device="cuda:0", gpus=[1])            # basemodel.py:107-109
    m = xDeepFM(cols, cols, device="cpu")
    with pytest.raises(NotImplementedError):
        m.compile("lbfgs", "binary_crossentropy")                 # basemodel.py:458
    with pytest.raises(NotImplementedError):
        m.compile("adam", "hinge")                                # basemodel.py:480
    assert SparseFeat("a", 10000, "auto").embedding_dim == 6 * int(pow(10000, 0.25))     # inputs.py:29-30
    with pytest.raises(RuntimeError):
        m.fit({"C1": torch.zeros(4).numpy(), "I1": torch.zeros(4).numpy()}, torch.zeros(4, 1).numpy(), batch_size=2, verbose=0)   # no CPU fallback


def test_varlen_feature_column_surface_matches_the_reference():
    """VarLenSparseFeat (inputs.py:41-77): field order, defaults, forwarded properties, hashing; column layout with a length
    column (inputs.py:99-123); tables created for multi-value features too (inputs.py:158-180)."""
    from deepctr.inputs import DenseFeat, SparseFeat, VarLenSparseFeat, build_input_features, create_embedding_matrix, get_feature_names
    sf = SparseFeat("hist", 30, 8, embedding_name="item")
    v = VarLenSparseFeat(sf, maxlen=5)
    assert VarLenSparseFeat._fields == ("sparsefeat", "maxlen", "combiner", "length_name")
    assert (v.combiner, v.length_name, v.maxlen) == ("mean", None, 5)
    assert (v.name, v.vocabulary_size, v.embedding_dim, v.use_hash, v.dtype, v.embedding_name, v.group_name) == \
        ("hist", 30, 8, False, "int32", "item", "default_group")
    assert hash(v) == hash("hist")
    cols = [SparseFeat("item", 30, 8), VarLenSparseFeat(sf, 5, "sum", length_name="hist_len"), DenseFeat("price", 2),
            VarLenSparseFeat(SparseFeat("tags", 9, 8), 3, "max")]
    fi = build_input_features(cols)
    assert list(fi.items()) == [("item", (0, 1)), ("hist", (1, 6)), ("hist_len", (6, 7)), ("price", (7, 9)), ("tags", (9, 12))]
    assert get_feature_names(cols) == ["item", "hist", "hist_len", "price", "tags"]
    tables = create_embedding_matrix(cols, init_std=1e-4)
    assert {k: tuple(t.weight.shape) for k, t in tables.items()} == {"item": (30, 8), "tags": (9, 8)}        # shared table 'item'
    lin = create_embedding_matrix(cols, linear=True)
    assert {k: tuple(t.weight.shape) for k, t in lin.items()} == {"item": (30, 1), "tags": (9, 1)}
    with pytest.raises(TypeError):
        build_input_features([type("Other", (), {"name": "x"})()])               # inputs.py:121-122


def test_column_selection_is_shared_between_the_deep_and_the_first_order_part():
    """Host logic: BaseModel._select hands both parts ONE tensor for the same columns of the same batch (the row-sharded lookup
    recognises a batch by tensor identity and fetches its distinct rows once), and forgets it when the batch is rewritten in place
    or another batch arrives."""
    spec = O.ModelSpec(sparse_names=["C1", "C2", "C3"], vocab_sizes=[5, 7, 9], embedding_dim=4, dense_names=["I1"],
                       cin_layer_size=(8, 4), dnn_hidden_units=(8,))
    model = build_product_model(spec, "cpu")
    ids = torch.arange(12, dtype=torch.int32).reshape(4, 3)
    assert model._select(ids, None) is ids
    a = model._select(ids, [0, 2])
    assert model._select(ids, [0, 2]) is a and a.tolist() == [[0, 2], [3, 5], [6, 8], [9, 11]]
    assert model._select(ids, [1]) is not a
    b = model._select(ids, [0, 2])                       # a different selection in between: recomputed, same values
    assert torch.equal(a, b)
    ids.add_(1)                                          # in-place write bumps the version
    c = model._select(ids, [0, 2])
    assert c is not b and c.tolist() == [[1, 3], [4, 6], [7, 9], [10, 12]]
    other = ids.clone()
    assert model._select(other, [0, 2]) is not c
    dense = torch.rand(4, 2)
    d = model._select(dense, [1])
    assert model._select(ids, [0, 2]).dtype == torch.int32 and model._select(dense, [1]) is d      # one memo per dtype
