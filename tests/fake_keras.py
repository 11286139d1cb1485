"""TEST INFRASTRUCTURE: a minimal stand-in for the slice of the Keras 2.2 functional API that the reference's
model-building code uses (pix2pose_model/ae_model.py:70-240, resnet50_mod.py:40-262), so that THAT CODE can be executed
here -- keras / tensorflow are not installable -- and the oracle's layer wiring checked against it.

``install()`` puts fake ``keras.*`` / ``tensorflow`` modules into ``sys.modules``; importing
``pix2pose_model.ae_model`` from /root/reference then works and ``aemodel_unet_resnet50()`` / ``aemodel_unet_prob()``
build a deferred graph: every layer call records a node, ``Model.predict`` evaluates the needed nodes with the SAME torch
primitives as oracle/net_oracle.py (``NetOracle.conv / convT / bn / dense / lrelu``), weights looked up by Keras layer
name.  What this pins is the reference's TOPOLOGY (which tensor feeds which layer, strides, paddings, channel slices,
concatenation order, layer names and construction order); the arithmetic of each primitive stays the oracle's
restatement of Keras semantics (unpinned)."""
import re
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

_counters = {}
_created = []          # every layer, in construction order


def reset():
    _counters.clear()
    del _created[:]


def _snake(name):
    s = re.sub("(.)([A-Z][a-z]+)", r"\1_\2", name)
    return re.sub("([a-z0-9])([A-Z])", r"\1_\2", s).lower()


class Node:
    def __init__(self, layer, inputs, index=0):
        self.layer, self.inputs, self.index = layer, inputs, index

    def __getitem__(self, key):                       # Lambda(lambda x: x[:, :, :, :32])
        return Node(_Slice(key), [self])


class Layer:
    def __init__(self, *args, name=None, **kw):
        cls = type(self).__name__
        if name is None:
            k = _snake(cls)
            _counters[k] = _counters.get(k, 0) + 1
            name = "%s_%d" % (k, _counters[k])
        self.name, self.args, self.kw = name, args, kw
        self.output = None
        _created.append(self)

    def __call__(self, x):
        self.output = Node(self, x if isinstance(x, (list, tuple)) else [x])
        return self.output

    def compute(self, ops, xs, wname):
        raise NotImplementedError(type(self).__name__)


class _Slice(Layer):
    def __init__(self, key):
        self.key, self.name = key, "slice"

    def compute(self, ops, xs, wname):
        n, h, w, c = self.key                          # NHWC slicing in the reference, NCHW here
        return xs[0][n, c, h, w].contiguous()


class Input(Layer):
    def __init__(self, shape=None, tensor=None, name=None, **kw):
        Layer.__init__(self, name=name)
        self.output = Node(self, [])


def _input(shape=None, tensor=None, **kw):            # keras.layers.Input returns the tensor, not the layer
    return Input(shape=shape, tensor=tensor, **kw).output


class Conv2D(Layer):
    def compute(self, ops, xs, wname):
        strides = self.kw.get("strides", (1, 1))
        return ops.conv(xs[0], wname, stride=strides[0], padding=self.kw.get("padding", "valid"))


class Conv2DTranspose(Layer):
    def compute(self, ops, xs, wname):
        assert self.kw.get("strides") == (2, 2) and self.kw.get("padding") == "same" and tuple(self.kw.get("kernel_size")) == (5, 5)
        return ops.convT(xs[0], wname)


class BatchNormalization(Layer):
    def compute(self, ops, xs, wname):
        assert self.kw.get("axis", 3) == 3
        return ops.bn(xs[0], wname)


class Dense(Layer):
    def compute(self, ops, xs, wname):
        return ops.dense(xs[0], wname)


class Activation(Layer):
    def compute(self, ops, xs, wname):
        return {"relu": F.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[self.args[0]](xs[0])


class LeakyReLU(Layer):
    def compute(self, ops, xs, wname):
        return ops.lrelu(xs[0])


class ZeroPadding2D(Layer):
    def compute(self, ops, xs, wname):
        p = self.kw.get("padding", self.args[0] if self.args else (1, 1))
        return F.pad(xs[0], (p[1], p[1], p[0], p[0]))


class MaxPooling2D(Layer):
    def compute(self, ops, xs, wname):
        from oracle.net_oracle import _same_pad
        k, s = self.args[0][0], self.kw["strides"][0]
        x = xs[0]
        if self.kw.get("padding", "valid") == "same":
            pt, pb, _ = _same_pad(x.shape[2], k, s)
            pl, pr, _ = _same_pad(x.shape[3], k, s)
            x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
        return F.max_pool2d(x, k, s)


class Concatenate(Layer):
    def compute(self, ops, xs, wname):
        return torch.cat(xs, 1)


class Add(Layer):
    def compute(self, ops, xs, wname):
        return xs[0] + xs[1]


class Lambda(Layer):
    def __call__(self, x):
        self.output = self.args[0](x)                 # the reference's lambdas only slice channels -> Node.__getitem__
        return self.output


class Flatten(Layer):
    def compute(self, ops, xs, wname):
        return xs[0].permute(0, 2, 3, 1).reshape(xs[0].shape[0], -1)       # Keras flattens NHWC


class Reshape(Layer):
    def compute(self, ops, xs, wname):
        h, w, c = self.args[0]
        return xs[0].reshape(-1, h, w, c if c > 0 else xs[0].shape[1] // (h * w)).permute(0, 3, 1, 2).contiguous()


class _Unused(Layer):
    """Layers the inference graph never evaluates (Dropout, AveragePooling2D of the unused ResNet stages, ...)."""

    def compute(self, ops, xs, wname):
        raise RuntimeError("layer %s is not part of the inference path" % self.name)


class Model(Layer):
    def __init__(self, inputs=None, outputs=None, name=None, **kw):
        Layer.__init__(self, name=name)
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        self.input = self.inputs[0]
        self.single_output = not isinstance(outputs, (list, tuple))
        self.weight_names = {}

    def __call__(self, x):
        outs = [Node(self, [x], i) for i in range(len(self.outputs))]
        return outs[0] if self.single_output else outs

    def get_layer(self, name):
        for l in _created:
            if l.name == name:
                return l
        raise ValueError("no layer " + name)

    def load_weights(self, *a, **kw):
        pass

    # ---- evaluation
    def _eval(self, node, ops, names, memo, bound):
        if id(node) in memo:
            return memo[id(node)]
        if id(node) in bound:
            v = bound[id(node)]
        elif isinstance(node.layer, Model):
            sub = node.layer
            x = self._eval(node.inputs[0], ops, names, memo, bound)
            inner = {id(sub.inputs[0]): x}
            v = sub._eval(sub.outputs[node.index], ops, names, {}, inner)
        else:
            xs = [self._eval(i, ops, names, memo, bound) for i in node.inputs]
            v = node.layer.compute(ops, xs, names.get(node.layer.name, node.layer.name))
        memo[id(node)] = v
        return v

    def run(self, ops, x_nchw, names):
        memo, bound = {}, {id(self.inputs[0]): x_nchw}
        return [self._eval(o, ops, names, memo, bound) for o in self.outputs]


def weighted_layers(model=None):
    """Layers that own weights, in construction order (= the order Keras' topological sort gives these chains); with
    ``model`` only those its outputs depend on (the reference builds all five ResNet stages and uses three)."""
    keep = None
    if model is not None:
        keep, stack = set(), list(model.outputs)
        seen = set()
        while stack:
            n = stack.pop()
            if id(n) in seen:
                continue
            seen.add(id(n))
            keep.add(id(n.layer))
            stack.extend(n.inputs)
            if isinstance(n.layer, Model):
                stack.append(n.layer.outputs[n.index])
    return [l for l in _created if isinstance(l, (Conv2D, Conv2DTranspose, BatchNormalization, Dense)) and (keep is None or id(l) in keep)]


class _Fake(types.ModuleType):
    def __getattr__(self, name):                      # anything the reference imports but never calls on this path
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Unused,), {})


def install():
    reset()
    mods = {}

    def mod(name, **attrs):
        m = _Fake(name)
        m.__dict__.update(attrs)
        mods[name] = m
        return m

    layers = dict(Input=_input, Conv2D=Conv2D, Conv2DTranspose=Conv2DTranspose, BatchNormalization=BatchNormalization, Dense=Dense,
                  Activation=Activation, LeakyReLU=LeakyReLU, ZeroPadding2D=ZeroPadding2D, MaxPooling2D=MaxPooling2D,
                  Concatenate=Concatenate, Lambda=Lambda, Flatten=Flatten, Reshape=Reshape, Layer=Layer,
                  add=lambda xs: Add()(xs))
    backend = mod("keras.backend", image_data_format=lambda: "channels_last", is_keras_tensor=lambda t: True)
    lay = mod("keras.layers", **layers)
    mod("keras.layers.normalization", BatchNormalization=BatchNormalization)
    mod("keras.layers.advanced_activations", LeakyReLU=LeakyReLU)
    models = mod("keras.models", Model=Model)
    mod("keras.regularizers", l2=lambda *a, **k: None)
    mod("keras.initializers", glorot_normal=lambda *a, **k: None)
    for n in ("keras.losses", "keras.optimizers", "keras.callbacks", "keras.engine", "keras.utils", "keras.utils.layer_utils",
              "keras.applications"):
        mod(n)
    mod("keras.engine.topology", get_source_inputs=lambda t: t)
    mod("keras.utils.data_utils", get_file=lambda *a, **k: "unused")
    mod("keras.applications.imagenet_utils", decode_predictions=None, preprocess_input=None,
        _obtain_input_shape=lambda input_shape, **k: input_shape)
    mod("keras", backend=backend, layers=lay, models=models, losses=mods["keras.losses"], optimizers=mods["keras.optimizers"],
        utils=mods["keras.utils"])
    mod("tensorflow")
    sys.modules.update(mods)
    return mods


def uninstall():
    for n in [k for k in sys.modules if k == "tensorflow" or k == "keras" or k.startswith("keras.") or k.startswith("pix2pose_model")]:
        mod = sys.modules[n]
        if isinstance(mod, _Fake) or n.startswith("pix2pose_model"):
            del sys.modules[n]
