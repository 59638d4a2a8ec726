"""Per-step training metrics computed from the epoch's prediction log (deepctr/metrics_device.py) against sklearn, which is what
the reference calls after every step (basemodel.py:264-269).  The module is plain torch: checked here on CPU tensors; the GPU
test runs fit(verbose=1) with and without it."""
import numpy as np
import pytest
import torch
from sklearn.metrics import accuracy_score, log_loss, mean_squared_error, roc_auc_score

from deepctr import metrics_device as MD


class _M:
    @staticmethod
    def _accuracy_score(y_true, y_pred):
        return accuracy_score(y_true, np.where(y_pred > 0.5, 1, 0))


def _case(n, seed, ties):
    g = np.random.default_rng(seed)
    p = g.random(n).astype(np.float32)
    if ties:
        p = np.round(p, 1).astype(np.float32)              # many tied predictions, exact 0.0 and 1.0 included
    y = (g.random(n) < 0.3).astype(np.float32)
    return p, y


@pytest.mark.parametrize("n,bs,ties", [(1000, 128, False), (1000, 128, True), (256, 256, True), (77, 256, False), (640, 512, True)])
def test_step_metrics_equal_sklearn_per_step(n, bs, ties):
    p, y = _case(n, n + bs, ties)
    metrics = {"auc": roc_auc_score, "binary_crossentropy": log_loss, "mse": mean_squared_error, "acc": _M._accuracy_score}
    got = MD.step_metrics(_M(), metrics, torch.from_numpy(p), torch.from_numpy(y), bs)
    assert got is not None
    steps = (n - 1) // bs + 1
    for name, fn in metrics.items():
        want = [fn(y[s * bs:(s + 1) * bs], p[s * bs:(s + 1) * bs].astype("float64")) for s in range(steps)]
        assert got[name].dtype == torch.float64 and got[name].shape == (steps,)
        assert np.allclose(got[name].numpy(), want, rtol=1e-12, atol=1e-15), (name, got[name].numpy(), want)


def test_step_metrics_leave_what_sklearn_would_reject_or_what_they_do_not_know_to_the_host():
    p, y = _case(300, 3, False)
    y[128:256] = 1.0                                        # second step: one class only -> roc_auc_score raises on the host path
    assert MD.step_metrics(_M(), {"auc": roc_auc_score}, torch.from_numpy(p), torch.from_numpy(y), 128) is None
    assert MD.step_metrics(_M(), {"mse": mean_squared_error}, torch.from_numpy(p), torch.from_numpy(y), 128) is not None
    assert MD.step_metrics(_M(), {"mine": lambda a, b: 0.0}, torch.from_numpy(p), torch.from_numpy(y), 128) is None
    assert MD.step_metrics(_M(), {"mse": mean_squared_error}, torch.from_numpy(p), torch.from_numpy(np.stack([y, y], 1)), 128) is None
