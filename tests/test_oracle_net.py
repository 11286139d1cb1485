"""CPU: the torch restatement of the generator against algebraic identities, the shape table of
SURVEY.md §8a and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.net_oracle import NetOracle, _same_pad
from pix2pose_b200 import weights as W

GOLD = os.path.join(os.path.dirname(__file__), "golden", "net_golden.npz")


def test_same_padding_is_asymmetric_for_k5_s2():
    assert _same_pad(16, 5, 2) == (1, 2, 8)
    assert _same_pad(128, 5, 2) == (1, 2, 64)
    assert _same_pad(64, 3, 2) == (0, 1, 32)     # maxpool 3x3 s2 'same'
    assert _same_pad(32, 5, 1) == (2, 2, 32)


def test_conv_transpose_is_gradient_of_same_conv():
    """Keras Conv2DTranspose(k5,s2,'same') == autograd of the SAME-padded forward conv."""
    torch.manual_seed(0)
    w = {"c/kernel": np.random.RandomState(0).randn(5, 5, 4, 6).astype(np.float32), "c/bias": np.zeros(4, np.float32)}
    net = NetOracle(w, "paper", torch.float64)
    x = torch.randn(2, 6, 8, 8, dtype=torch.float64)
    y = net.convT(x, "c")                                            # (2,4,16,16)
    # forward conv with kernel (kh,kw,Cin=4,Cout=6) on a 16x16 image, SAME s2 -> 8x8; its VJP w.r.t. input
    inp = torch.randn(2, 4, 16, 16, dtype=torch.float64, requires_grad=True)
    k = torch.from_numpy(w["c/kernel"]).double().permute(3, 2, 0, 1)  # (Cout=6, Cin=4, kh, kw)
    out = F.conv2d(F.pad(inp, (1, 2, 1, 2)), k, stride=2)
    (g,) = torch.autograd.grad(out, inp, grad_outputs=x)
    assert torch.allclose(g, y, atol=1e-10)


def test_shapes_and_flatten_order():
    for bb, nparam in (("resnet50", 27904452), ("paper", 25740356)):
        assert W.count_params(bb) == nparam
        w = W.synthetic_weights(bb, 1)
        net = NetOracle(w, bb)
        net.taps = {}
        d, p = net.forward(np.zeros((1, 128, 128, 3), np.float32))
        assert d.shape == (1, 128, 128, 3) and p.shape == (1, 128, 128, 1)
        assert net.taps["f4"].shape == (1, 8, 8, 512) and net.taps["d0"].shape == (1, 8, 8, 256)
        assert net.taps["d1_uni"].shape == (1, 16, 16, 256) and net.taps["d3_uni"].shape == (1, 64, 64, 128)
        assert np.abs(d).max() <= 1 and p.min() >= 0 and p.max() <= 1
    assert net.taps["f1"].shape == (1, 64, 64, 128)


def test_golden_vectors():
    g = np.load(GOLD)
    x = np.random.RandomState(0).uniform(-1, 1, (2, 128, 128, 3)).astype(np.float32)
    for bb in ("resnet50", "paper"):
        d, p = NetOracle(W.synthetic_weights(bb, 1), bb).forward(x)
        assert np.abs(d[:, ::8, ::8, :] - g[bb + "_decode_s8"]).max() < 2e-5
        assert np.abs(p[:, ::8, ::8, :] - g[bb + "_prob_s8"]).max() < 2e-5


def test_fp64_mode_agrees():
    w = W.synthetic_weights("paper", 1)
    x = np.random.RandomState(1).uniform(-1, 1, (1, 128, 128, 3)).astype(np.float32)
    d32, _ = NetOracle(w, "paper").forward(x)
    d64, _ = NetOracle(w, "paper", torch.float64).forward(x)
    assert np.abs(d32 - d64).max() < 1e-4


# ---- pinned by the reference's own model-building code (tests/golden/make_topology_golden.py) ---------------------------
def _topology_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "topology_golden.npz"))


@pytest.mark.parametrize("backbone", ["resnet50", "paper"])
def test_forward_equals_the_graph_built_by_the_reference_code(backbone):
    """tests/golden/topology_golden.npz: ae_model.py / resnet50_mod.py of the reference EXECUTED on a fake Keras whose layers
    use this oracle's primitives.  Equal bits => the hand-written forward wires every layer exactly like the reference
    (inputs, strides, paddings, channel slices f1[:32] / f2[:128] / f3[:128], concatenation order, head order)."""
    from oracle.net_oracle import NetOracle
    from pix2pose_b200 import weights as W
    g = _topology_golden()
    x = np.random.RandomState(0).uniform(-1, 1, (2, 128, 128, 3)).astype(np.float32)
    d, p = NetOracle(W.synthetic_weights(backbone, 1), backbone).forward(x)
    assert np.array_equal(d[:, ::4, ::4], g[backbone + "_decode_s4"])
    assert np.array_equal(p[:, ::4, ::4], g[backbone + "_prob_s4"])


@pytest.mark.parametrize("backbone", ["resnet50", "paper"])
def test_keras_import_follows_the_reference_construction_order(backbone):
    """The weight-carrying layers in the order the reference code constructs them (Keras names: explicit ones and the
    auto-numbered batch_normalization_k / dense_k / conv2d_transpose_k) mapped by weights.keras_layers_to_weights give back
    every tensor in its place -- the importer's name / per-kind-order rule matches the reference's construction order."""
    from pix2pose_b200 import weights as W
    g = _topology_golden()
    w = W.synthetic_weights(backbone, 4)
    kinds = {n: k for n, k, _ in W.layer_table(backbone)}
    file_layers = []
    for kname, tname in zip(g[backbone + "_keras_order"], g[backbone + "_table_order"]):
        kname, tname = str(kname), str(tname)
        keys = ("gamma", "beta", "moving_mean", "moving_variance") if kinds[tname] == W.BN else ("kernel", "bias")
        file_layers.append((kname, {k: w[tname + "/" + k] for k in keys}))
    assert len(file_layers) == len(kinds)
    got = W.keras_layers_to_weights(file_layers, backbone)
    assert all(np.array_equal(got[k], w[k]) for k in W.param_names(backbone))
