"""TEST INFRASTRUCTURE ONLY -- makes the unmodified reference importable in THIS container.

The reference (`/root/reference/deepctr`) imports tensorflow's Keras callbacks at module import
(`deepctr/models/basemodel.py:22-25`, `deepctr/callbacks.py:2-4`) and starts a PyPI version-check
thread (`deepctr/__init__.py:6`, `deepctr/utils.py:19-44`).  Neither exists / works offline, so this
loader injects minimal stand-ins *before* importing the reference.  Nothing of the reference is
modified or copied; the stand-ins only provide the Keras callback protocol the reference calls.

Used by `oracle/make_golden.py` (fixture generation) and by `bench.py --impl reference`.  `/root/reference` does not exist on
the GPU box; there the loader finds the byte-for-byte copy `oracle/_ref/` made by `oracle/build_ref.py` (git-ignored, shipped
with the snapshot).  The `-m gpu` parity tests and `smoke()` never call `load_reference()` (the product package is also named
`deepctr`; the reference is only ever imported in its own process).

Never import this from the product package.
"""
import importlib
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root():
    """/root/reference in the build container; on the GPU box the byte-for-byte copy made by oracle/build_ref.py (oracle/_ref)."""
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "deepctr")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = os.environ.get("XDFM_REFERENCE_ROOT") or _default_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deepctr"))


class _Callback:
    def __init__(self):
        self.model = None
        self.params = None

    def set_model(self, model):
        self.model = model

    def set_params(self, params):
        self.params = params

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass


class _History(_Callback):
    def on_train_begin(self, logs=None):
        self.epoch = []
        self.history = {}

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)


class _CallbackList:
    def __init__(self, callbacks=None):
        self.callbacks = list(callbacks or [])

    def set_model(self, model):
        self.model = model
        for c in self.callbacks:
            c.set_model(model)

    def set_params(self, params):
        for c in self.callbacks:
            c.set_params(params)

    def on_train_begin(self, logs=None):
        for c in self.callbacks:
            c.on_train_begin(logs)

    def on_train_end(self, logs=None):
        for c in self.callbacks:
            c.on_train_end(logs)

    def on_epoch_begin(self, epoch, logs=None):
        for c in self.callbacks:
            c.on_epoch_begin(epoch, logs)

    def on_epoch_end(self, epoch, logs=None):
        for c in self.callbacks:
            c.on_epoch_end(epoch, logs)


class _EarlyStopping(_Callback):
    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", **_):
        super().__init__()
        import numpy as np
        self.monitor, self.patience, self.verbose = monitor, patience, verbose
        self.min_delta = abs(min_delta)
        if mode == "max" or (mode == "auto" and ("acc" in monitor or "auc" in monitor)):
            self.monitor_op, self.best = np.greater, -np.inf
        else:
            self.monitor_op, self.best = np.less, np.inf
            self.min_delta *= -1
        self.wait = 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.monitor_op(cur - self.min_delta, self.best):
            self.best, self.wait = cur, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.model.stop_training = True


class _ModelCheckpoint(_Callback):
    # attributes read by the reference's subclass (`deepctr/callbacks.py:41-73`)
    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False,
                 save_weights_only=False, mode="auto", period=1, **_):
        super().__init__()
        import numpy as np
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.save_weights_only, self.period = save_best_only, save_weights_only, period
        self.epochs_since_last_save = 0
        if mode == "max" or (mode == "auto" and ("acc" in monitor or "auc" in monitor or monitor.startswith("fmeasure"))):
            self.monitor_op, self.best = np.greater, -np.inf
        else:
            self.monitor_op, self.best = np.less, np.inf


def _stub_module(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)  # torch._dynamo calls find_spec("tensorflow")
    m.__path__ = []
    return m


def install_stubs():
    """Install the tensorflow callback stand-ins and neutralise the PyPI thread."""
    try:
        import torch.utils.tensorboard  # noqa: F401  must be imported BEFORE the fake tensorflow exists
    except Exception:
        pass
    if "tensorflow" not in sys.modules:
        names = ["tensorflow", "tensorflow.python", "tensorflow.python.keras",
                 "tensorflow.python.keras.callbacks"]
        mods = {n: _stub_module(n) for n in names}
        cb = mods["tensorflow.python.keras.callbacks"]
        cb.CallbackList, cb.History = _CallbackList, _History
        cb.EarlyStopping, cb.ModelCheckpoint, cb.Callback = _EarlyStopping, _ModelCheckpoint, _Callback
        mods["tensorflow"].python = mods["tensorflow.python"]
        mods["tensorflow.python"].keras = mods["tensorflow.python.keras"]
        mods["tensorflow.python.keras"].callbacks = cb
        sys.modules.update(mods)
    try:
        import requests

        def _no_network(*a, **k):
            raise RuntimeError("offline")
        requests.get = _no_network
    except Exception:
        pass


def load_reference():
    """Return the reference's `deepctr` package (imported from REFERENCE_ROOT)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    install_stubs()
    if "deepctr" in sys.modules:
        mod = sys.modules["deepctr"]
        if not os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)):
            raise RuntimeError("a different `deepctr` is already imported: %s" % mod.__file__)
        return mod
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            mod = importlib.import_module("deepctr")
            importlib.import_module("deepctr.models")
            importlib.import_module("deepctr.xdeepfm_pro")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return mod
