"""Generates tests/golden/topology_golden.npz by EXECUTING the reference's model-building code
(/root/reference/pix2pose_model/ae_model.py ``aemodel_unet_resnet50`` :175-240, ``aemodel_unet_prob`` :70-150, and
resnet50_mod.py :40-262) on top of tests/fake_keras.py, a stand-in for the Keras functional API whose layers evaluate with
the oracle's torch primitives.  The stored outputs therefore come from the reference's own wiring (inputs of every layer,
strides, paddings, channel slices, concatenation order, names); tests/test_oracle_net.py checks that the hand-written
forward of oracle/net_oracle.py reproduces them bit for bit, and that the construction order of the weight-carrying
layers is the order pix2pose_b200.weights.keras_layers_to_weights assumes.  Run in the build container:
  python tests/golden/make_topology_golden.py"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.net_oracle import NetOracle          # noqa: E402
from pix2pose_b200 import weights as W           # noqa: E402
from tests import fake_keras as FK               # noqa: E402

REF = "/root/reference/pix2pose_model"


def load_reference_ae_model():
    FK.install()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "pix2pose_model" or k.startswith("pix2pose_model.")}
    pkg = types.ModuleType("pix2pose_model")
    pkg.__path__ = [REF]                           # so that `import pix2pose_model.resnet50_mod` finds the reference's file
    sys.modules["pix2pose_model"] = pkg
    spec = importlib.util.spec_from_file_location("pix2pose_model.ae_model", os.path.join(REF, "ae_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, saved


def keras_names_to_table(backbone, model):
    """Keras layer name -> our table name: explicit names are shared, auto-named layers pair up per kind in order."""
    table = W.layer_table(backbone)
    ours = {n for n, _, _ in table}
    kind_of = {FK.BatchNormalization: W.BN, FK.Dense: W.DENSE, FK.Conv2D: W.CONV, FK.Conv2DTranspose: W.CONVT}
    rest = {k: [n for n, kk, _ in table if kk == k] for k in (W.BN, W.DENSE, W.CONV, W.CONVT)}
    layers = FK.weighted_layers(model)
    used = {l.name for l in layers if l.name in ours}
    for k in rest:
        rest[k] = [n for n in rest[k] if n not in used]
    names = {}
    for l in layers:
        names[l.name] = l.name if l.name in ours else rest[kind_of[type(l)]].pop(0)
    return names, [l.name for l in layers]


def main():
    out = {}
    x = np.random.RandomState(0).uniform(-1, 1, (2, 128, 128, 3)).astype(np.float32)
    for backbone, builder in (("resnet50", "aemodel_unet_resnet50"), ("paper", "aemodel_unet_prob")):
        ae, saved = load_reference_ae_model()
        model = getattr(ae, builder)(p=1.0)
        names, order = keras_names_to_table(backbone, model)
        w = W.synthetic_weights(backbone, 1)
        ops = NetOracle(w, backbone)
        with torch.no_grad():
            dec, prob = model.run(ops, torch.from_numpy(x).permute(0, 3, 1, 2).contiguous(), names)
        dec = dec.permute(0, 2, 3, 1).contiguous().numpy()
        prob = prob.permute(0, 2, 3, 1).contiguous().numpy()
        out[backbone + "_decode_s4"], out[backbone + "_prob_s4"] = dec[:, ::4, ::4], prob[:, ::4, ::4]
        out[backbone + "_keras_order"] = np.array(order)
        out[backbone + "_table_order"] = np.array([names[n] for n in order])
        FK.uninstall()
        sys.modules.update(saved)
        print(backbone, "decode", dec.shape, "weighted layers", len(order), order[:6], "...")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "topology_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
