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


def load_keras_hdf5(path, backbone):
    """Order-based import of a Keras ``inference*.hdf5`` (recognition.py:23-26).
    Needs h5py, which is not part of this image -- convert offline and ship the ``.npz``."""
    try:
        import h5py  # noqa: F401
    except ImportError as e:
        raise ImportError("reading Keras .hdf5 weights needs h5py; convert to .npz offline "
                          "(pix2pose_b200.weights.save_npz)") from e
    raise NotImplementedError("Keras HDF5 import is SURVEY §8(f) item 1 (next)")
