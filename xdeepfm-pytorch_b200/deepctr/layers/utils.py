"""Small host helpers kept from the reference surface (deepctr/layers/utils.py:12-70)."""
import numpy as np
import torch


def concat_fun(inputs, axis=-1):
    return inputs[0] if len(inputs) == 1 else torch.cat(inputs, dim=axis)


def slice_arrays(arrays, start=None, stop=None):
    """arrays[start:stop] for one array or a list of arrays; `start` may also be a list/array of indices."""
    if arrays is None:
        return [None]
    if isinstance(arrays, np.ndarray):
        arrays = [arrays]
    if isinstance(start, list) and stop is not None:
        raise ValueError("The stop argument has to be None if the value of start is a list.")
    by_index = hasattr(start, "__len__")
    if by_index and hasattr(start, "shape"):
        start = start.tolist()
    if isinstance(arrays, list):
        if by_index:
            return [None if x is None else x[start] for x in arrays]
        if len(arrays) == 1:
            return arrays[0][start:stop]
        return [None if x is None else x[start:stop] for x in arrays]
    if by_index:
        return arrays[start]
    if hasattr(start, "__getitem__"):
        return arrays[start:stop]
    return [None]
