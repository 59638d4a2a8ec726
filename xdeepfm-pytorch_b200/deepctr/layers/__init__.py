from .activation import activation_layer, Identity
from .core import DNN, PredictionLayer
from .interaction import CIN
from .utils import concat_fun, slice_arrays
from .cin_attention import AttentionPooling, CINAttention, CINAttentionV2, MultiHeadSelfAttention
from .sequence import SequencePoolingLayer
