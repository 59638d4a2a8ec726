"""Host helpers of the reference surface (deepctr/layers/utils.py:12-70): `concat_fun`, `slice_arrays`."""
import numpy as np
import torch


def concat_fun(inputs, axis=-1):
    """One tensor: returned as is; several: concatenated along `axis`."""
    if len(inputs) == 1:
        return inputs[0]
    return torch.cat(inputs, dim=axis)


def _take(a, sel, stop):
    """a[sel:stop] for an integer `sel`, a[sel] (fancy indexing) when `sel` is an index list; None stays None."""
    if a is None:
        return None
    return a[sel] if isinstance(sel, list) else a[sel:stop]


def slice_arrays(arrays, start=None, stop=None):
    """Row slice of one array or of every array of a list (fit()'s validation_split and batch slicing).

    `start` is an int (with `stop`) or a list / ndarray of row indices (then `stop` must be None).  A one-element list
    gives back the sliced array itself, as the reference does."""
    if arrays is None:
        return [None]
    if isinstance(start, np.ndarray):
        start = start.tolist()
    if isinstance(start, list) and stop is not None:
        raise ValueError("The stop argument has to be None if the value of start is a list.")
    if isinstance(arrays, np.ndarray):
        arrays = [arrays]
    if not isinstance(arrays, list):
        if isinstance(start, list) or hasattr(start, "__getitem__"):
            return _take(arrays, start, stop)
        return [None]
    if len(arrays) == 1 and not isinstance(start, list):
        return _take(arrays[0], start, stop)
    return [_take(a, start, stop) for a in arrays]
