"""Keras-`Model`-shaped facade over the sm_100a engine: exactly the methods the reference calls on the
object returned by create_model (models.py:178,342,541,783 predict; :1218 load_weights; :1213 compile;
:146-155 fit_generator; get/set_weights for identical-weights parity)."""
from __future__ import annotations

import os

import numpy as np
import torch

from .engine import Engine


class Layer:
    def __init__(self, model, name, ksize, cin, cout):
        self._model, self.name = model, name
        self.kernel_size, self.filters, self.input_channels = (ksize, ksize), cout, cin
        self.trainable = True

    def get_weights(self):
        w, b = self._model.engine.master[self.name]
        return [w.cpu().numpy(), b.cpu().numpy()]

    def set_weights(self, ws):
        d = self._model.engine.get_weights_dict()
        d[self.name] = (ws[0], ws[1])
        self._model.engine.set_weights_dict(d)

    def count_params(self):
        k = self.kernel_size[0]
        return k * k * self.input_channels * self.filters + self.filters


class Model:
    def __init__(self, input_shape, engine=None, **engine_kw):
        self.input_shape = (None,) + tuple(input_shape)
        h, w, c = input_shape
        self.engine = engine if engine is not None else Engine(**engine_kw)
        sc = getattr(self.engine, "scale", 4)              # Engine (DifvdsrDouble) and Difvdsr4: x4; Difvdsr: x1
        self.output_shape = (None, sc * h, sc * w, 3)
        # the weighted layers in Keras `model.layers` order (depth-sorted: per 5/3 block a3, c5, b5, d3), which is
        # also the order of get_weights() / set_weights() lists; weightless layers have no object here
        by_name = {s[0]: s for s in self.engine.specs}
        self.layers = [Layer(self, *by_name[n]) for n, w in self.keras_layers() if w]
        self.optimizer = None
        self.loss = None
        self.metrics = []
        self.stop_training = False
        self._trainer = None

    # ------------------------------------------------------------------ inference
    def predict(self, x, batch_size=32, verbose=0):
        """float NHWC in [0,1] -> float32 NHWC, x4 (Keras Model.predict; batch_size only bounds memory here)."""
        x = np.asarray(x)
        if x.ndim != 4 or x.shape[-1] != 3:
            raise ValueError("Error when checking input: expected 4-D NHWC input with 3 channels, got %s" % (x.shape,))
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.engine.device)
        out = self.engine.forward_device(xd)
        return out.cpu().numpy()

    # ------------------------------------------------------------------ weights
    def get_weights(self):
        out = []
        for l in self.layers:
            out.extend(l.get_weights())
        return out

    def set_weights(self, ws):
        if len(ws) != 2 * len(self.layers):
            raise ValueError("You called `set_weights(weights)` with a weight list of length %d, but the model "
                             "was expecting %d weights." % (len(ws), 2 * len(self.layers)))
        d = {l.name: (ws[2 * i], ws[2 * i + 1]) for i, l in enumerate(self.layers)}
        self.engine.set_weights_dict(d)

    def count_params(self):
        return sum(l.count_params() for l in self.layers)

    def save_weights(self, path, overwrite=True):
        """Keras `save_weights`: an HDF5 file in the Keras 2 layout (sr100.h5lite writes it; no h5py needed), or an
        .npz with keys '<layer>/kernel:0', '<layer>/bias:0' when the path ends in .npz."""
        if not overwrite and os.path.exists(path):
            raise OSError("%s exists and overwrite=False" % path)
        wd = self.engine.get_weights_dict()
        if path.endswith(".npz"):
            d = {}
            for name, (w, b) in wd.items():
                d[name + "/kernel:0"] = w
                d[name + "/bias:0"] = b
            np.savez(path, **d)
            return path
        from . import h5lite
        h5lite.save_keras_weights(path, wd, layers=self.keras_layers())
        return path

    def keras_layers(self):
        """[(name, has_weights)] in Keras `model.layers` order (depth-sorted, sr100.keras_graph): what save_weights
        writes as `layer_names` and what load_weights zips positionally.  DifvdsrDouble has parallel branches (per
        5/3 block: a3, c5, b5, d3); the chain-shaped Difvdsr4 / Difvdsr graphs are in creation order."""
        if isinstance(self.engine, Engine):
            from .keras_graph import difvdsr_double_layers
            return difvdsr_double_layers()
        return [(n, True) for n, _, _, _ in self.engine.specs]

    def load_weights(self, path, by_name=False):
        """Keras `load_weights` of an HDF5 weight file (`save_weights` or `model.save` layout, models.py:1217-1218),
        or of the .npz exchange format.  HDF5 files load the way Keras loads them: POSITIONALLY -- the file's weighted
        layers in `layer_names` order against this model's weighted layers in `model.layers` order, names ignored
        (so a file whose auto-names are offset, e.g. written after a second create_model without clear_session,
        loads exactly as in Keras); by_name=True matches group names instead."""
        if not os.path.exists(path):
            raise OSError("Unable to open file (unable to open file: name = '%s', errno = 2, error message = "
                          "'No such file or directory')" % path)
        names = [n for n, _, _, _ in self.engine.specs]
        with open(path, "rb") as f:
            magic = f.read(4)
        if magic[:2] == b"PK":                           # npz (zip container)
            z = np.load(path)
            d = {n: (z[n + "/kernel:0"], z[n + "/bias:0"]) for n in names}
        else:
            from . import h5lite
            if by_name:
                d = h5lite.load_keras_weights(path, names)
            else:
                mine = [n for n, w in self.keras_layers() if w]
                theirs = h5lite.load_keras_weights_positional(path)
                if len(theirs) != len(mine):
                    raise ValueError("You are trying to load a weight file containing %d layers into a model with "
                                     "%d layers." % (len(theirs), len(mine)))
                d = {}
                for n, (fname, arrs) in zip(mine, theirs):
                    if len(arrs) != 2:
                        raise ValueError("Layer #%s (named \"%s\" in the current model) was found to correspond to "
                                         "layer %s in the save file. However the new layer %s expects 2 weights, but "
                                         "the saved weights have %d elements." % (mine.index(n), n, fname, n, len(arrs)))
                    k, b = (arrs[0], arrs[1]) if arrs[0].ndim > arrs[1].ndim else (arrs[1], arrs[0])
                    d[n] = (k, b)
        for (n, k, cin, cout) in self.engine.specs:
            if tuple(d[n][0].shape) != (k, k, cin, cout) or tuple(d[n][1].shape) != (cout,):
                raise ValueError("Layer %s: weight file holds kernel %s / bias %s, the model expects %s / %s"
                                 % (n, tuple(d[n][0].shape), tuple(d[n][1].shape), (k, k, cin, cout), (cout,)))
        self.engine.set_weights_dict(d)

    # ------------------------------------------------------------------ training facade
    def compile(self, optimizer=None, loss=None, metrics=None):
        self.optimizer, self.loss, self.metrics = optimizer, loss, list(metrics or [])
        if loss not in (None, "mse", "mean_squared_error"):
            raise ValueError("only loss='mse' is implemented (models.py:1213)")

    def _get_trainer(self):
        if self._trainer is None:
            opt = self.optimizer
            kw = dict(lr=getattr(opt, "lr", 1e-4), beta_1=getattr(opt, "beta_1", 0.9),
                      beta_2=getattr(opt, "beta_2", 0.999), epsilon=getattr(opt, "epsilon", 1e-7))
            if isinstance(self.engine, Engine):
                from .train import Trainer
                self._trainer = Trainer(self.engine, **kw)
            else:                                   # Difvdsr4 / Difvdsr (models.py:1079-1080, 1332-1333)
                from .planetrain import PlaneTrainer
                self._trainer = PlaneTrainer(self.engine, **kw)
        return self._trainer

    def train_on_batch(self, x, y):
        return self._get_trainer().train_on_batch(x, y)

    def fit_generator(self, generator, steps_per_epoch, epochs=1, callbacks=None, validation_data=None,
                      validation_steps=None, verbose=1, **_):
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin({})
        history = {}
        for epoch in range(epochs):
            losses = []
            for _ in range(int(steps_per_epoch)):
                x, y = next(generator)
                losses.append(self.train_on_batch(x, y))
            logs = {"loss": float(np.mean(losses)) if losses else float("nan")}
            if validation_data is not None and validation_steps:
                vl = []
                for _ in range(int(validation_steps)):
                    x, y = next(validation_data)
                    vl.append(self._get_trainer().evaluate(x, y))
                logs["val_loss"] = float(np.mean([v[0] for v in vl]))
                logs["val_acc"] = float(np.mean([v[1] for v in vl]))
            for k, v in logs.items():
                history.setdefault(k, []).append(v)
            if verbose:
                print("Epoch %d/%d - %s" % (epoch + 1, epochs, " - ".join("%s: %.6f" % kv for kv in logs.items())))
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        return history

    def summary(self):
        print("_" * 65)
        print("%-28s%-26s%s" % ("Layer (type)", "Output Shape", "Param #"))
        print("=" * 65)
        for l in self.layers:
            print("%-28s%-26s%d" % (l.name + " (Conv2D)", "(None, ?, ?, %d)" % l.filters, l.count_params()))
        print("=" * 65)
        print("Total params: {:,}".format(self.count_params()))
