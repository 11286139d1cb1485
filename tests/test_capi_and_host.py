"""CPU: the C-ABI library loads and exports every symbol include/pix2pose_b200.h declares; host-side
tables agree between Python and C++; the product fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from pix2pose_b200 import _lib, weights as W
from pix2pose_b200.recognition import _get_boxes, _get_boxes_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "pix2pose_b200.h")).read()
    names = re.findall(r"P2P_API\s+[\w\s\*]+?\b(p2p_\w+)\s*\(", hdr)
    assert len(names) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert set(names) >= set(_lib.lib()._sig_names)          # everything Python binds is declared


def test_param_tables_agree_with_library():
    L = _lib.lib()
    for bb, flops in (("resnet50", 10.701e9), ("paper", 12.582e9)):
        w = W.synthetic_weights(bb, 1)
        assert L.p2p_param_count(bb.encode()) == W.count_params(bb) == sum(int(np.prod(w[n].shape)) for n in W.param_names(bb))
        assert abs(L.p2p_flops_per_crop(bb.encode()) - flops) / flops < 1e-3
    assert L.p2p_param_count(b"vgg") == 0 and b"backbone" in L.p2p_last_error()


def test_weight_npz_roundtrip(tmp_path):
    w = W.synthetic_weights("paper", 3)
    W.save_npz(tmp_path / "w.npz", w, "paper")
    w2 = W.load_npz(tmp_path / "w.npz", "paper")
    assert all(np.array_equal(w[k], w2[k]) for k in W.param_names("paper"))
    with pytest.raises(ValueError):
        W.load_npz(tmp_path / "w.npz", "resnet50")
    bad = dict(w); bad["dense_1/kernel"] = bad["dense_1/kernel"][:10]
    with pytest.raises(ValueError):
        W.check_shapes(bad, "paper")


def test_get_boxes_matches_oracle():
    from oracle.recognition_oracle import get_boxes
    rng = np.random.RandomState(0)
    for _ in range(500):
        v0, u0 = rng.randint(-60, 460), rng.randint(-60, 620)
        bb = [v0, u0, v0 + rng.randint(1, 300), u0 + rng.randint(1, 300)]
        assert tuple(_get_boxes(1.5, bb, 480, 640)) == tuple(get_boxes(1.5, bb, 480, 640))
        assert tuple(_get_boxes_batch(1.5, [bb], 480, 640)[0]) == tuple(get_boxes(1.5, bb, 480, 640))
        fb = np.array(bb) * 1.37
        ct = np.array([rng.randint(0, 480), rng.randint(0, 640)])
        assert tuple(_get_boxes(1.5, fb, 480, 640, ct=ct, max_w=128)) == tuple(get_boxes(1.5, fb, 480, 640, ct=ct, max_w=128))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_cpu_fallback():
    from pix2pose_b200 import ae_model
    with pytest.raises(_lib.P2PError) as e:
        ae_model.GeneratorModel("paper", capacity=2)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    from pix2pose_b200.pnp import solve_pnp_ransac
    with pytest.raises(_lib.P2PError):
        solve_pnp_ransac(np.zeros((10, 3)), np.zeros((10, 2)), np.eye(3))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pix2pose_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/" not in src or f.endswith((".cu", ".cuh")) and "oracle/resize_oracle.py" in src, f
