"""AutoDis soft-bucket encoder for dense features (reference: deepctr/xdeepfm_pro/autodis.py:20-238).

Default-off in every reference configuration (`use_autodis=False`, xdeepfm_pro.py:85) and listed as a "next" item of the
hot-path scope: the classes exist for import compatibility and refuse to run instead of silently falling back
to stock PyTorch kernels."""
import torch.nn as nn

_MSG = ("AutoDis (use_autodis=True) is not part of the B200 hot path of this build yet (SURVEY.md 8f-4); the reference's "
        "default is use_autodis=False")


class AutoDisLayer(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(_MSG)


class DenseFeatureEncoder(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(_MSG)
