"""Weight tables for the two Pix2Pose generator backbones: canonical layer order, Keras
layouts, synthetic initialisation, ``.npz`` I/O.

The layer ORDER below is the Keras topological order of
``pix2pose_model/ae_model.py:175-240`` (resnet50) / ``:70-150`` (paper) -- the order in
which ``Model.load_weights`` walks an ``inference*.hdf5`` (recognition.py:23-26; order-based,
not name-based).  Names are ours; layouts are Keras':

* Conv2D            ``kernel (kh,kw,Cin,Cout)``, ``bias (Cout,)``
* Conv2DTranspose   ``kernel (kh,kw,Cout,Cin)``, ``bias (Cout,)``
* Dense             ``kernel (in,out)``, ``bias (out,)``
* BatchNormalization ``gamma, beta, moving_mean, moving_variance`` each ``(C,)``
"""
import re

import numpy as np

CONV, CONVT, DENSE, BN = "conv", "convT", "dense", "bn"


def _resnet_block(stage, block, cin, f1, f2, f3, shortcut):
    base, bnb = "res%d%s_branch" % (stage, block), "bn%d%s_branch" % (stage, block)
    L = [(base + "2a", CONV, (1, 1, cin, f1)), (bnb + "2a", BN, (f1,)),
         (base + "2b", CONV, (3, 3, f1, f2)), (bnb + "2b", BN, (f2,)),
         (base + "2c", CONV, (1, 1, f2, f3)), (bnb + "2c", BN, (f3,))]
    if shortcut:
        L += [(base + "1", CONV, (1, 1, cin, f3)), (bnb + "1", BN, (f3,))]
    return L


def _decoder(skip3, skip2, skip1):
    return [
        ("dense_1", DENSE, (8 * 8 * 512, 256)), ("dense_2", DENSE, (256, 8 * 8 * 256)),
        ("convT1", CONVT, (5, 5, 256, 256)), ("bn_convT1", BN, (256,)),
        ("deconv1", CONV, (5, 5, 256 + skip3, 256)), ("bn_deconv1", BN, (256,)),
        ("convT2", CONVT, (5, 5, 128, 256)), ("bn_convT2", BN, (128,)),
        ("deconv2", CONV, (5, 5, 128 + skip2, 256)), ("bn_deconv2", BN, (256,)),
        ("convT3", CONVT, (5, 5, 64, 256)), ("bn_convT3", BN, (64,)),
        ("deconv3", CONV, (5, 5, 64 + skip1, 128)), ("bn_deconv3", BN, (128,)),
        ("convT_xyz", CONVT, (5, 5, 3, 128)), ("convT_prob", CONVT, (5, 5, 1, 128)),
    ]


def layer_table(backbone):
    """Ordered ``[(name, kind, shape)]`` for one backbone."""
    if backbone == "resnet50":
        L = [("conv1", CONV, (7, 7, 3, 64)), ("bn_conv1", BN, (64,))]
        L += _resnet_block(2, "a", 64, 64, 64, 256, True)
        L += _resnet_block(2, "b", 256, 64, 64, 256, False)
        L += _resnet_block(2, "c", 256, 64, 64, 256, False)
        L += _resnet_block(3, "a", 256, 128, 128, 512, True)
        for b in "bcd":
            L += _resnet_block(3, b, 512, 128, 128, 512, False)
        L += [("conv4_1", CONV, (5, 5, 512, 256)), ("bn_conv4_1", BN, (256,)),
              ("conv4_2", CONV, (5, 5, 512, 256)), ("bn_conv4_2", BN, (256,))]
        L += _decoder(128, 128, 32)
    elif backbone == "paper":
        L = []
        for lvl, cin, cout in ((1, 3, 64), (2, 128, 128), (3, 256, 128), (4, 256, 256)):
            for br in (1, 2):
                n = "conv%d_%d" % (lvl, br)
                L += [(n, CONV, (5, 5, cin, cout)), ("bn_" + n, BN, (cout,))]
        L += _decoder(128, 128, 64)
    else:
        raise ValueError("backbone must be 'resnet50' or 'paper', got %r" % (backbone,))
    return L


def param_names(backbone):
    out = []
    for name, kind, _ in layer_table(backbone):
        if kind == BN:
            out += [name + s for s in ("/gamma", "/beta", "/moving_mean", "/moving_variance")]
        else:
            out += [name + "/kernel", name + "/bias"]
    return out


def count_params(backbone):
    n = 0
    for _, kind, shp in layer_table(backbone):
        n += 4 * shp[0] if kind == BN else int(np.prod(shp)) + (shp[2] if kind == CONVT else shp[-1])
    return n


def synthetic_weights(backbone, seed=1, prob_bias=-1.5):
    """Deterministic random-init weights: He-normal kernels, BN
    gamma~U(.5,1.5) beta~N(0,.1) mean~N(0,.1) var~U(.5,1.5).  ``prob_bias`` shifts the error
    head so a useful share of pixels passes ``prob < th_outlier`` (config 3 note)."""
    rng = np.random.RandomState(seed)
    w = {}
    for name, kind, shp in layer_table(backbone):
        if kind == BN:
            c = shp[0]
            w[name + "/gamma"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
            w[name + "/beta"] = (0.1 * rng.randn(c)).astype(np.float32)
            w[name + "/moving_mean"] = (0.1 * rng.randn(c)).astype(np.float32)
            w[name + "/moving_variance"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
            continue
        if kind == CONV:
            fan_in, nb = shp[0] * shp[1] * shp[2], shp[3]
            gain = 2.0
        elif kind == CONVT:
            fan_in, nb = shp[0] * shp[1] * shp[3] / 4.0, shp[2]   # ~25/4 taps reach an output pixel
            gain = 2.0
        else:
            fan_in, nb = shp[0], shp[1]
            gain = 1.0                                            # no activation after the Dense pair
        if name.endswith("branch2c"):
            gain = 0.25                                           # keep the residual sum O(1)
        elif name.endswith("branch1"):
            gain = 1.0
        elif name in ("convT_xyz", "convT_prob"):
            gain = 0.5                                            # un-saturated tanh / sigmoid heads
        w[name + "/kernel"] = (np.sqrt(gain / fan_in) * rng.randn(*shp)).astype(np.float32)
        w[name + "/bias"] = (0.05 * rng.randn(nb)).astype(np.float32)
    w["convT_prob/bias"] = w["convT_prob/bias"] + np.float32(prob_bias)
    return w


def save_npz(path, weights, backbone):
    names = param_names(backbone)
    missing = [n for n in names if n not in weights]
    if missing:
        raise KeyError("weights missing %d tensors, first: %s" % (len(missing), missing[0]))
    np.savez(path, __backbone__=np.array(backbone), **{n: weights[n] for n in names})


def load_npz(path, backbone=None):
    with np.load(path, allow_pickle=False) as z:
        stored = str(z["__backbone__"]) if "__backbone__" in z.files else None
        if backbone is not None and stored is not None and stored != backbone:
            raise ValueError("weight file %s is for backbone %r, model is %r" % (path, stored, backbone))
        bb = backbone or stored
        if bb is None:
            raise ValueError("backbone unknown for %s" % path)
        w = {n: np.asarray(z[n], dtype=np.float32) for n in param_names(bb)}
    check_shapes(w, bb)
    return w


def check_shapes(weights, backbone):
    for name, kind, shp in layer_table(backbone):
        if kind == BN:
            for s in ("/gamma", "/beta", "/moving_mean", "/moving_variance"):
                if tuple(weights[name + s].shape) != tuple(shp):
                    raise ValueError("%s%s has shape %s, expected %s" % (name, s, weights[name + s].shape, shp))
        else:
            if tuple(weights[name + "/kernel"].shape) != tuple(shp):
                raise ValueError("%s/kernel has shape %s, expected %s" % (name, weights[name + "/kernel"].shape, shp))
            nb = shp[2] if kind == CONVT else shp[-1]
            if tuple(weights[name + "/bias"].shape) != (nb,):
                raise ValueError("%s/bias has shape %s, expected (%d,)" % (name, weights[name + "/bias"].shape, nb))


def _keras_file_layers(f):
    """Flat, ordered ``[(layer_name, {suffix: array})]`` of a Keras weight file: ``layer_names`` / ``weight_names``
    attributes in file order (keras/engine/saving.py load_weights_from_hdf5_group), a full-model file's
    ``model_weights`` group unwrapped, nested models (the ResNet part of ae_model.py:178-184 is a ``Model`` used as a
    layer) split into their own layers by the ``layer/weight:0`` prefix of every weight name."""
    g = f["model_weights"] if "layer_names" not in f.attrs and "model_weights" in f else f

    def names(attrs, key):
        if key in attrs:
            vals = list(np.asarray(attrs[key]).ravel())
        else:                       # Keras splits attributes larger than 64 KB into key0, key1, ...
            vals, i = [], 0
            while "%s%d" % (key, i) in attrs:
                vals += list(np.asarray(attrs["%s%d" % (key, i)]).ravel())
                i += 1
        return [v.decode("utf-8") if isinstance(v, bytes) else str(v) for v in vals]

    out, index = [], {}
    for lname in names(g.attrs, "layer_names"):
        grp = g[lname]
        for wname in names(grp.attrs, "weight_names"):
            arr = np.asarray(grp[wname], np.float32)
            parts = wname.split("/")
            owner, suffix = parts[-2] if len(parts) >= 2 else lname, parts[-1].split(":")[0]
            if owner not in index:
                index[owner] = {}
                out.append((owner, index[owner]))
            index[owner][suffix] = arr
    return out


# layer names given in the reference's model code: resnet50_mod.py:59-114, 200-202; ae_model.py:74-103, 118-138, 190-228
_EXPLICIT = re.compile(r"^(conv1|bn_conv1|res\d[a-z]_branch\w+|bn\d[a-z]_branch\w+|conv[1-4]_[12]|deconv[1-3])$")


def keras_layers_to_weights(file_layers, backbone):
    """Maps the ordered layers of a Keras file onto ``param_names(backbone)``.  Layers the reference names explicitly
    (``conv1``, ``bn_conv1``, ``res2a_branch2a`` ..., ``conv4_1`` ...; resnet50_mod.py:60-118, ae_model.py:74-106,
    190-195) are matched by NAME -- inside the nested ResNet model Keras orders ``branch2c`` / ``branch1`` by graph
    depth, not by our table.  The auto-named rest (``batch_normalization_k``, ``dense_k``, ``conv2d_transpose_k``,
    ``conv2d_k``) is matched in file order per layer kind, which is how ``Model.load_weights`` pairs them (topological
    order is the construction order for these chains)."""
    table = layer_table(backbone)
    by_name = {n: w for n, w in file_layers}
    used = set()
    out = {}

    def kind_of(name, w):
        if "gamma" in w or "moving_mean" in w:
            return BN
        k = w.get("kernel")
        if k is None:
            return None
        if k.ndim == 2:
            return DENSE
        return CONVT if "transpose" in name else CONV

    def put(name, kind, w, src):
        keys = ("gamma", "beta", "moving_mean", "moving_variance") if kind == BN else ("kernel", "bias")
        for k in keys:
            if k not in w:
                raise ValueError("layer %s (file layer %s) has no %r" % (name, src, k))
            out[name + "/" + k] = np.asarray(w[k], np.float32)

    for name, kind, _ in table:                      # pass 1: the names the reference sets explicitly, and only those --
        # our table's own names for auto-named Keras layers ('dense_1', 'dense_2', ...) collide with Keras' counters: a file
        # saved by a process that had built another Dense before holds 'dense_2' / 'dense_3', and matching 'dense_2' by name
        # would pair the wrong kernels where Keras' order-based load_weights pairs the right ones
        if not _EXPLICIT.match(name):
            continue
        if name in by_name and kind_of(name, by_name[name]) in (kind, CONV if kind == CONVT else kind):
            put(name, kind, by_name[name], name)
            used.add(name)
    rest = {BN: [], DENSE: [], CONV: [], CONVT: []}
    for n, w in file_layers:                          # pass 2: the auto-named remainder, per kind, in file order
        if n in used:
            continue
        k = kind_of(n, w)
        if k is not None:
            rest[k].append((n, w))
    for name, kind, shp in table:
        if name + ("/gamma" if kind == BN else "/kernel") in out:
            continue
        cand = rest[kind]
        if kind == CONVT and not cand:                # files whose transposed convs carry no 'transpose' in the name
            cand = rest[CONV]
        if not cand:
            raise ValueError("Keras file has no %s layer left for %s" % (kind, name))
        n, w = cand.pop(0)
        put(name, kind, w, n)
    check_shapes(out, backbone)
    return out


def load_keras_hdf5(path, backbone):
    """Import of a Keras ``inference*.hdf5`` / ``pix2pose.NN-x.hdf5`` weight file (recognition.py:23-26;
    tools/4_convert_weights_inference.py:50-52) through the built-in HDF5 reader
    (``hdf5_lite``; h5py is used instead when it is installed)."""
    try:
        import h5py
        opener = lambda p: h5py.File(p, "r")   # noqa: E731
    except ImportError:
        from . import hdf5_lite
        opener = hdf5_lite.File
    with opener(path) as f:
        layers = _keras_file_layers(f)
    return keras_layers_to_weights(layers, backbone)


def save_keras_hdf5(path, weights, backbone):
    """Writes ``weights`` in the layout ``Model.save_weights`` produces (one group per layer, ``layer_names`` /
    ``weight_names`` attributes, datasets ``<layer>/<layer>/<weight>:0``), with our table names as layer names.
    Used by the tests and as an export path; read back by ``load_keras_hdf5`` and by Keras' ``load_weights(by_name=False)``
    for a model built in the same order."""
    from . import hdf5_lite
    w = hdf5_lite.Writer()
    lnames = []
    for name, kind, _ in layer_table(backbone):
        keys = ("gamma", "beta", "moving_mean", "moving_variance") if kind == BN else ("kernel", "bias")
        w.create_group(name)
        wn = []
        for k in keys:
            w.create_dataset("%s/%s/%s:0" % (name, name, k), np.asarray(weights[name + "/" + k], np.float32))
            wn.append(("%s/%s:0" % (name, k)).encode())
        w.set_attr(name, "weight_names", np.array(wn))
        lnames.append(name.encode())
    w.set_attr("", "layer_names", np.array(lnames))
    w.set_attr("", "backend", np.array(b"tensorflow"))
    w.set_attr("", "keras_version", np.array(b"2.2.1"))
    w.save(path)
