"""xDeepFM Pro -- xDeepFM with a Supervised-Feature-Generation auxiliary decoder (reference: deepctr/xdeepfm_pro/__init__.py:27-41),
on the fused B200 ops."""
from .autodis import AutoDisLayer, DenseFeatureEncoder
from .basemodel_sfg import BaseModelSFG
from .sfg_decoder import LabelAwareAttention, SFGDecoder, SFGLoss
from .xdeepfm_pro import xDeepFMPro, xDeepFMProLight

__all__ = ['xDeepFMPro', 'xDeepFMProLight', 'SFGDecoder', 'SFGLoss', 'LabelAwareAttention', 'AutoDisLayer',
           'DenseFeatureEncoder', 'BaseModelSFG']
