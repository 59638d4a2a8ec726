"""Script-level drop-in proof: the reference's caller scripts run against THIS repo's
`deepctr` package.

1. `test_xdftrain_call_sequence`: replays, in our own words, exactly the calls xdftrain.py makes (xdftrain.py:259-285 build_model with
   the lr override through `model.optim.param_groups`, :417-452 a duck-typed TensorBoard-style callback with the
   `_implements_*_batch_hooks` methods + ModelCheckpoint(save_best_only, save_weights_only) + EarlyStopping, fit(shuffle=True,
   validation_data), :455-458 `load_state_dict(torch.load(ckpt, map_location="cpu"))`, predict, :495 save of the weights).
2. `test_unmodified_xdftrain_script_runs_on_this_package`: runs the reference's UNMODIFIED xdftrain.py (byte-for-byte copy made by
   oracle/build_ref.py into oracle/_ref/, shipped to the GPU box) in a subprocess whose `deepctr` is this repo's package, on a
   synthetic Criteo-format file, and checks the artefacts it writes.  Skipped when oracle/_ref is absent (fresh clone without
   /root/reference)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "xdeepfm-pytorch_b200")
REF_SCRIPT = os.path.join(ROOT, "oracle", "_ref", "xdftrain.py")


class DuckTypedLogger:
    """Keras-protocol callback that does NOT derive from any callback base class (the shape of xdftrain.py:31-97)."""

    def __init__(self):
        self.model = self.params = None
        self.epochs, self.began, self.ended = [], 0, 0

    def set_model(self, model):
        self.model = model

    def set_params(self, params):
        self.params = params

    def _implements_train_batch_hooks(self):
        return False

    def _implements_test_batch_hooks(self):
        return False

    def _implements_predict_batch_hooks(self):
        return False

    def on_batch_begin(self, batch, logs=None):
        pass

    def on_batch_end(self, batch, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        self.epochs.append((epoch, dict(logs or {})))

    def on_train_begin(self, logs=None):
        self.began += 1

    def on_train_end(self, logs=None):
        self.ended += 1


def _criteo_like(n, seed):
    g = np.random.default_rng(seed)
    vocab = [40, 12, 300, 150, 9, 5, 200, 30, 3, 120, 80, 260, 60, 7, 140, 220, 6, 90, 50, 4, 250, 8, 10, 180, 20, 110]
    sparse = np.stack([g.integers(0, v, n) for v in vocab], 1)
    dense = g.random((n, 13)).astype("float32")
    logit = 0.8 * np.sin(sparse[:, 2] * 0.7) + 1.2 * (sparse[:, 4] % 3 == 0) + 2.0 * dense[:, 0] - 1.8
    y = (g.random(n) < 1 / (1 + np.exp(-logit))).astype("float32")
    return vocab, sparse, dense, y


def test_xdftrain_call_sequence(tmp_path):
    from sklearn.metrics import log_loss, roc_auc_score
    from deepctr.callbacks import EarlyStopping, ModelCheckpoint
    from deepctr.inputs import DenseFeat, SparseFeat, get_feature_names
    from deepctr.models import xDeepFM
    vocab, sparse, dense, y = _criteo_like(6000, 5)
    sparse_features = ["C%d" % i for i in range(1, 27)]
    dense_features = ["I%d" % i for i in range(1, 14)]
    cols = [SparseFeat(f, vocabulary_size=int(sparse[:, i].max()) + 1, embedding_dim=10) for i, f in enumerate(sparse_features)] + \
           [DenseFeat(f, 1) for f in dense_features]
    feature_names = get_feature_names(cols + cols)
    data = {f: sparse[:, i].astype("int64") for i, f in enumerate(sparse_features)}
    data.update({f: dense[:, i] for i, f in enumerate(dense_features)})
    tr, ev = slice(0, 5000), slice(5000, 6000)
    train_x = {n: data[n][tr] for n in feature_names}
    eval_x = {n: data[n][ev] for n in feature_names}
    y_train, y_eval = y[tr].reshape(-1, 1), y[ev].reshape(-1, 1)
    # build_model (xdftrain.py:259-285)
    model = xDeepFM(linear_feature_columns=cols, dnn_feature_columns=cols, task="binary", l2_reg_embedding=1e-5, l2_reg_dnn=1e-5,
                    dnn_dropout=0.0, device=DEV)
    model.compile(optimizer="adam", loss="binary_crossentropy", metrics=["binary_crossentropy", "auc"])
    for param_group in model.optim.param_groups:
        param_group["lr"] = 0.004
    ckpt = str(tmp_path / "xdeepfm_best.pth")
    logger = DuckTypedLogger()
    callbacks = [logger, ModelCheckpoint(filepath=ckpt, monitor="val_auc", save_best_only=True, save_weights_only=True, mode="max", verbose=1)]
    callbacks.insert(1, EarlyStopping(monitor="val_auc", patience=50, mode="max", verbose=1))
    history = model.fit(train_x, y_train, batch_size=512, epochs=4, verbose=2, validation_data=(eval_x, y_eval), shuffle=True,
                        callbacks=callbacks)
    assert float(model.optim.param_groups[0]["lr"]) == 0.004
    assert logger.began == 1 and logger.ended == 1 and [e for e, _ in logger.epochs] == [0, 1, 2, 3]
    assert logger.model is model
    for key in ("loss", "binary_crossentropy", "auc", "val_binary_crossentropy", "val_auc"):
        assert len(history.history[key]) == 4, key
        assert set(logger.epochs[-1][1]) >= {key}
    assert history.history["loss"][-1] < history.history["loss"][0]
    assert os.path.exists(ckpt)
    best = int(np.argmax(history.history["val_auc"]))
    model.load_state_dict(torch.load(ckpt, map_location="cpu"))
    eval_pred = model.predict(eval_x, batch_size=8192)
    assert eval_pred.shape == (1000, 1) and eval_pred.dtype == np.float64
    auc = roc_auc_score(y_eval, eval_pred)
    assert abs(auc - history.history["val_auc"][best]) < 1e-6, "reloaded best checkpoint must reproduce the best epoch's val_auc"
    assert np.isfinite(log_loss(y_eval, eval_pred)) and auc > 0.6
    torch.save(model.state_dict(), str(tmp_path / "xdeepfm_weights.pth"))
    json.dumps(history.history)                     # xdftrain.py:497-498 dumps it as is: plain python floats only


@pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="oracle/_ref absent (run `python -m oracle.build_ref` where /root/reference exists)")
def test_unmodified_xdftrain_script_runs_on_this_package(tmp_path):
    vocab, sparse, dense, y = _criteo_like(4000, 9)
    header = ["label"] + ["I%d" % i for i in range(1, 14)] + ["C%d" % i for i in range(1, 27)]

    def write(path, lo, hi, labelled=True):
        with open(path, "w") as f:
            if labelled:
                f.write("\t".join(header) + "\n")
            for r in range(lo, hi):
                row = (["%d" % y[r]] if labelled else []) + ["%.4f" % v for v in dense[r]] + ["c%x" % v for v in sparse[r]]
                f.write("\t".join(row) + "\n")

    train, evalf, test = str(tmp_path / "train.txt"), str(tmp_path / "eval.txt"), str(tmp_path / "test.txt")
    write(train, 0, 3000)
    write(evalf, 3000, 3600)
    write(test, 3600, 4000, labelled=False)
    out_dir = str(tmp_path / "out")
    env = dict(os.environ, PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""), PYTHONDONTWRITEBYTECODE="1")
    # -c so that sys.path[0] is not the script's directory (oracle/_ref holds the reference's own deepctr next to the script)
    runner = ("import runpy, sys; sys.argv = %r; import deepctr, os; "
              "assert os.path.abspath(deepctr.__file__).startswith(%r), deepctr.__file__; "
              "runpy.run_path(%r, run_name='__main__')") % (
        ["xdftrain.py", "--data_path", train, "--eval_path", evalf, "--test_path", test, "--out_dir", out_dir, "--mode", "eval",
         "--device", DEV, "--epochs", "3", "--batch_size", "256", "--embedding_dim", "8", "--verbose", "2", "--use_early_stopping"],
        PKG, REF_SCRIPT)
    r = subprocess.run([sys.executable, "-c", runner], env=env, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    hist = json.load(open(os.path.join(out_dir, "history.json")))
    assert len(hist["loss"]) == 3 and len(hist["val_auc"]) == 3 and hist["loss"][-1] < hist["loss"][0]
    sd = torch.load(os.path.join(out_dir, "xdeepfm_weights.pth"), map_location="cpu")
    assert "cin.conv1ds.0.weight" in sd and "embedding_dict.C1.weight" in sd and "linear_model.embedding_dict.C26.weight" in sd
    assert os.path.exists(os.path.join(out_dir, "xdeepfm_best.pth"))
    preds = open(os.path.join(out_dir, "test_predictions.csv")).read().strip().splitlines()
    assert len(preds) == 401 and all(0.0 <= float(v) <= 1.0 for v in preds[1:])
    assert "[Eval] eval AUC" in r.stdout
