"""Drop-in for ``pix2pose_model/ae_model.py``'s inference builders.

``aemodel_unet_resnet50(p)`` (ae_model.py:175) and ``aemodel_unet_prob(p)`` (ae_model.py:70) return
an object with the Keras surface ``recognition.py`` uses: ``load_weights(path)`` (:23,26) and
``predict(x) -> [decode, prob]`` (:84,129).  The forward runs on hand-written sm_100a kernels
(csrc/conv_tc.cuh) behind the C ABI of include/pix2pose_b200.h; there is no CPU path.
"""
import ctypes

import numpy as np

from . import _lib, weights as W


class Engine:
    """Per-device activation workspace + execution plan for one backbone (shared by all objects)."""

    _cache = {}

    def __init__(self, backbone, capacity=64, precision="fp16x3"):
        self.backbone = backbone
        self.capacity = int(capacity)
        self.precision = precision
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().p2p_engine_create(backbone.encode(), self.capacity, _lib.precision_code(precision),
                                                ctypes.byref(h)))
        self.handle = h

    @classmethod
    def shared(cls, backbone, capacity=64, precision="fp16x3"):
        key = (backbone, precision)
        e = cls._cache.get(key)
        if e is None or e.capacity < capacity:
            e = cls(backbone, capacity, precision)
            cls._cache[key] = e
        return e

    @property
    def launch_count(self):
        return int(_lib.lib().p2p_engine_launch_count(self.handle))

    def read_tensor(self, name, n):
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        L = _lib.lib()
        _lib.check(L.p2p_engine_read_tensor(self.handle, name.encode(), n, None, ctypes.byref(h), ctypes.byref(w),
                                            ctypes.byref(c)))
        out = np.empty((n, h.value, w.value, c.value), np.float32)
        _lib.check(L.p2p_engine_read_tensor(self.handle, name.encode(), n, _lib.fptr(out), None, None, None))
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().p2p_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class GeneratorModel:
    """Keras-``Model`` duck type of the generator (``generator_train`` in recognition.py:22-26)."""

    def __init__(self, backbone, capacity=64, precision="fp16x3", engine=None):
        self.backbone = backbone
        self.engine = engine if engine is not None else Engine.shared(backbone, capacity, precision)
        self._model = None
        self.weights = None

    # -- keras surface ---------------------------------------------------------------------------
    def load_weights(self, path_or_dict):
        """A Keras weight file (``inference.hdf5`` / ``pix2pose.NN-x.hdf5``; recognition.py:23-26) read by the built-in
        HDF5 reader (``weights.load_keras_hdf5``), an ``.npz`` written by ``weights.save_npz``, or a dict of
        Keras-layout arrays."""
        if isinstance(path_or_dict, dict):
            w = path_or_dict
            W.check_shapes(w, self.backbone)
        else:
            path = str(path_or_dict)
            if path.endswith((".hdf5", ".h5")):
                w = W.load_keras_hdf5(path, self.backbone)
            else:
                w = W.load_npz(path, self.backbone)
        blob = np.concatenate([np.asarray(w[n], np.float32).ravel() for n in W.param_names(self.backbone)])
        L = _lib.lib()
        if blob.size != L.p2p_param_count(self.backbone.encode()):
            raise ValueError("weight blob size %d != library count %d" % (blob.size, L.p2p_param_count(self.backbone.encode())))
        h = ctypes.c_void_p()
        _lib.check(L.p2p_model_create(self.engine.handle, _lib.fptr(blob), blob.size, ctypes.byref(h)))
        self._release()
        self._model = h
        self.weights = w

    def save_weights(self, path):
        """Keras ``Model.save_weights`` counterpart (tools/4_convert_weights_inference.py:52): ``.hdf5`` / ``.h5`` in the
        Keras layout, anything else as ``.npz``."""
        if self.weights is None:
            raise RuntimeError("save_weights() before load_weights()")
        if str(path).endswith((".hdf5", ".h5")):
            W.save_keras_hdf5(str(path), self.weights, self.backbone)
        else:
            W.save_npz(str(path), self.weights, self.backbone)

    def predict(self, x, batch_size=None, verbose=0):
        if self._model is None:
            raise RuntimeError("predict() before load_weights()")
        x = _lib.as_f32(x)
        if x.ndim != 4 or x.shape[1:] != (128, 128, 3):
            raise ValueError("expected input (N,128,128,3), got %s" % (x.shape,))
        n = x.shape[0]
        dec = np.empty((n, 128, 128, 3), np.float32)
        prob = np.empty((n, 128, 128, 1), np.float32)
        _lib.check(_lib.lib().p2p_predict(self.engine.handle, self._model, _lib.fptr(x), n, _lib.fptr(dec), _lib.fptr(prob)))
        return [dec, prob]

    def time_forward(self, x, warmup=3, iters=10):
        """Device-resident forward time in ms (CUDA events); x: (n<=capacity,128,128,3)."""
        x = _lib.as_f32(x)
        ms = ctypes.c_float()
        _lib.check(_lib.lib().p2p_time_forward(self.engine.handle, self._model, _lib.fptr(x), x.shape[0], warmup, iters,
                                               ctypes.byref(ms)))
        return ms.value

    def _release(self):
        if self._model is not None:
            _lib.lib().p2p_model_destroy(self._model)
            self._model = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


def aemodel_unet_resnet50(p=0.5, **kw):
    return GeneratorModel("resnet50", **kw)


def aemodel_unet_prob(p=0.5, **kw):
    return GeneratorModel("paper", **kw)
