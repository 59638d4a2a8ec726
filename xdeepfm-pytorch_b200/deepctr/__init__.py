"""deepctr -- B200-native drop-in for the xDeepFM hot path of Syclus123/xDeepFM-pytorch's vendored deepctr package.

Same import paths as the reference (`deepctr.inputs`, `deepctr.models`, `deepctr.layers`, `deepctr.callbacks`,
`deepctr.xdeepfm_pro`); no tensorflow import, no PyPI version-check thread (reference: deepctr/__init__.py:5-6).
"""
from . import inputs  # noqa: F401
from . import layers  # noqa: F401
from . import models  # noqa: F401
from . import xdeepfm_pro  # noqa: F401

__version__ = "0.2.9+b200"
