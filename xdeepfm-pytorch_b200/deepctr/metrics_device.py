"""Per-step training metrics on the device.

The reference evaluates every metric on the HOST after every step (`basemodel.py:264-269`: `metric_fun(y.cpu().numpy(),
y_pred.cpu().numpy().astype("float64"))`, sklearn) and logs their mean over the steps of the epoch.  `fit()` here keeps the
predictions of an epoch on the device; this module turns that log into the same per-step values in float64 with a handful of
batched torch calls -- no per-step host round trip, no `steps x sklearn.roc_auc_score` at the end of the epoch.

Definitions (equal to sklearn's up to float64 rounding):
  auc                  Mann-Whitney U with mid-ranks for tied predictions / (n_pos * n_neg)  == trapezoidal area under the ROC
                       curve (`sklearn.metrics.roc_auc_score`)
  binary_crossentropy  mean(-(y log p + (1 - y) log(1 - p))), p and 1 - p clipped to [eps, 1 - eps], eps = float64 epsilon
                       (`sklearn.metrics.log_loss`, sklearn >= 1.5)
  mse                  mean((y - p)^2)
  accuracy             mean((p > 0.5) == y)
A step whose labels hold a single class makes sklearn's auc / log_loss raise; `step_metrics` then returns None and the caller
uses the host path, which raises exactly as the reference does.  Metric functions this module does not know are left to the host.
"""
import numpy as np
import torch

_EPS64 = float(np.finfo(np.float64).eps)


def _auc_rows(p, y):
    """p, y float64 [S, n] -> float64 [S]."""
    n = p.shape[1]
    sp, idx = torch.sort(p, dim=1)
    ys = torch.gather(y, 1, idx)
    pos = torch.arange(n, device=p.device, dtype=torch.int64).expand_as(sp)
    first_flag = torch.ones_like(sp, dtype=torch.bool)
    first_flag[:, 1:] = sp[:, 1:] != sp[:, :-1]
    last_flag = torch.ones_like(first_flag)
    last_flag[:, :-1] = first_flag[:, 1:]
    zero = torch.zeros((), dtype=torch.int64, device=p.device)
    first = torch.cummax(torch.where(first_flag, pos, zero), dim=1).values            # start of each element's tie group
    last = torch.cummin(torch.where(last_flag, pos, zero + (n - 1)).flip(1), dim=1).values.flip(1)
    rank = (first + last).double() * 0.5 + 1.0                                        # mid-rank, 1-based
    npos = ys.sum(1)
    nneg = float(n) - npos
    return ((rank * ys).sum(1) - npos * (npos + 1.0) * 0.5) / (npos * nneg)


def _logloss_rows(p, y):
    q = (1.0 - p).clamp(_EPS64, 1.0 - _EPS64)
    pc = p.clamp(_EPS64, 1.0 - _EPS64)
    return -(torch.xlogy(y, pc) + torch.xlogy(1.0 - y, q)).mean(1)


def _mse_rows(p, y):
    return ((y - p) ** 2).mean(1)


def _acc_rows(p, y):
    return ((p > 0.5).double() == y).double().mean(1)


def _kind(name, fn, model):
    """Which device metric computes what `fn` computes (None = unknown function: leave it to the host)."""
    from sklearn.metrics import log_loss, mean_squared_error, roc_auc_score
    if fn is roc_auc_score:
        return "auc"
    if fn is log_loss:
        return "logloss"
    if fn is mean_squared_error:
        return "mse"
    if getattr(fn, "__func__", fn) is getattr(type(model)._accuracy_score, "__func__", type(model)._accuracy_score):
        return "acc"
    return None


_ROWS = {"auc": _auc_rows, "logloss": _logloss_rows, "mse": _mse_rows, "acc": _acc_rows}


def step_metrics(model, metrics, pred_log, y_log, batch_size):
    """metrics: {name: host function}; pred_log / y_log: device tensors [n] of one epoch in step order.
    -> {name: float64 device tensor [steps]} (per-step values, as the reference appends them), or None if any metric is unknown
    to this module or a step cannot be evaluated the way sklearn would (single-class labels under auc / log_loss)."""
    kinds = {name: _kind(name, fn, model) for name, fn in metrics.items()}
    n = int(pred_log.shape[0])
    if n == 0 or any(k is None for k in kinds.values()) or y_log.dim() != 1 or y_log.shape[0] != n:
        return None
    p = pred_log.detach().double()
    y = y_log.detach().double()
    full = n // batch_size
    parts = []
    if full > 0:
        parts.append((p[:full * batch_size].view(full, batch_size), y[:full * batch_size].view(full, batch_size)))
    if n - full * batch_size > 0:
        parts.append((p[full * batch_size:].view(1, -1), y[full * batch_size:].view(1, -1)))
    if any(k in ("auc", "logloss") for k in kinds.values()):
        is_binary = bool(((y == 0) | (y == 1)).all().item())
        one_class = any(bool(((yy.sum(1) == 0) | (yy.sum(1) == yy.shape[1])).any().item()) for _, yy in parts)
        if not is_binary or one_class:
            return None
    return {name: torch.cat([_ROWS[k](pp, yy) for pp, yy in parts]) for name, k in kinds.items()}
