"""Keras-protocol callbacks without tensorflow (the reference re-exports TF-Keras classes: deepctr/callbacks.py:1-73).

Same names and ctor arguments: EarlyStopping, ModelCheckpoint, History; plus the CallbackList used by fit().
Duck-typed callbacks (e.g. the TensorBoardCallback of xdftrain.py:31-97) are accepted as they are.
"""
import numpy as np
import torch


class Callback:
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

    def on_batch_begin(self, batch, logs=None):
        pass

    def on_batch_end(self, batch, logs=None):
        pass


class CallbackList:
    def __init__(self, callbacks=None):
        self.callbacks = list(callbacks or [])
        self.model = None

    def _each(self, method, *args):
        for cb in self.callbacks:
            fn = getattr(cb, method, None)
            if fn is not None:
                fn(*args)

    def set_model(self, model):
        self.model = model
        self._each("set_model", model)

    def set_params(self, params):
        self._each("set_params", params)

    def on_train_begin(self, logs=None):
        self._each("on_train_begin", logs)

    def on_train_end(self, logs=None):
        self._each("on_train_end", logs)

    def on_epoch_begin(self, epoch, logs=None):
        self._each("on_epoch_begin", epoch, logs)

    def on_epoch_end(self, epoch, logs=None):
        self._each("on_epoch_end", epoch, logs)


class History(Callback):
    def on_train_begin(self, logs=None):
        self.epoch = []
        self.history = {}

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)


def _monitor_direction(monitor, mode):
    if mode == "min":
        return np.less, np.inf
    if mode == "max":
        return np.greater, -np.inf
    if "acc" in monitor or "auc" in monitor or monitor.startswith("fmeasure"):
        return np.greater, -np.inf
    return np.less, np.inf


class EarlyStopping(Callback):
    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", baseline=None,
                 restore_best_weights=False):
        super().__init__()
        self.monitor, self.patience, self.verbose, self.baseline = monitor, patience, verbose, baseline
        self.restore_best_weights = restore_best_weights
        self.monitor_op, _ = _monitor_direction(monitor, mode)
        self.min_delta = abs(min_delta) * (1 if self.monitor_op is np.greater else -1)
        self.wait = self.stopped_epoch = 0
        self.best_weights = None

    def on_train_begin(self, logs=None):
        self.wait = self.stopped_epoch = 0
        self.best = self.baseline if self.baseline is not None else (np.inf if self.monitor_op is np.less else -np.inf)

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            return
        if self.monitor_op(current - self.min_delta, self.best):
            self.best, self.wait = current, 0
            if self.restore_best_weights:
                self.best_weights = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
                if self.restore_best_weights and self.best_weights is not None:
                    self.model.load_state_dict(self.best_weights)

    def on_train_end(self, logs=None):
        if self.stopped_epoch > 0 and self.verbose > 0:
            print("Epoch %05d: early stopping" % (self.stopped_epoch + 1))


class ModelCheckpoint(Callback):
    """Save `state_dict()` (save_weights_only) or the pickled model after an epoch (reference: callbacks.py:41-73)."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", period=1):
        super().__init__()
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.save_weights_only, self.period = save_best_only, save_weights_only, period
        self.epochs_since_last_save = 0
        self.monitor_op, self.best = _monitor_direction(monitor, mode)

    def _save(self, path):
        dist_ctx = getattr(self.model, "_dist", None)
        if dist_ctx is not None:
            # row-sharded tables: state_dict() is a collective that re-assembles the reference key layout; rank 0 writes
            sd = self.model.state_dict()
            if dist_ctx.rank == 0:
                torch.save(sd, path)
            return
        torch.save(self.model.state_dict() if self.save_weights_only else self.model, path)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if not self.save_best_only:
            if self.verbose > 0:
                print("Epoch %05d: saving model to %s" % (epoch + 1, path))
            self._save(path)
            return
        current = logs.get(self.monitor)
        if current is None:
            print("Can save best model only with %s available, skipping." % self.monitor)
        elif self.monitor_op(current, self.best):
            if self.verbose > 0:
                print("Epoch %05d: %s improved from %0.5f to %0.5f, saving model to %s"
                      % (epoch + 1, self.monitor, self.best, current, path))
            self.best = current
            self._save(path)
        elif self.verbose > 0:
            print("Epoch %05d: %s did not improve from %0.5f" % (epoch + 1, self.monitor, self.best))
