"""CPU: the resize restatement (skimage semantics pinned in oracle/resize_oracle.py)."""
import os

import numpy as np
from scipy import ndimage

from oracle.resize_oracle import resize

GOLD = os.path.join(os.path.dirname(__file__), "golden", "resize_golden.npz")


def _coords(n_in, n_out):
    return (np.arange(n_out) + 0.5) * n_in / n_out - 0.5


def test_identity_at_scale_one():
    a = np.random.RandomState(0).rand(128, 128, 3)
    assert np.array_equal(resize(a, (128, 128), mode="reflect"), a)
    assert np.array_equal(resize(a, (128, 128), mode="constant", cval=0.5), a)


def test_matches_ndimage_mirror_and_grid_constant():
    a = np.random.RandomState(0).rand(128, 128)
    for oh, ow in [(200, 200), (77, 91), (128, 300), (301, 129)]:
        R, C = np.meshgrid(_coords(128, oh), _coords(128, ow), indexing="ij")
        ref = ndimage.map_coordinates(a, [R, C], order=1, mode="mirror")
        assert np.abs(resize(a, (oh, ow), mode="reflect", clip=False) - ref).max() < 1e-12
        ref = ndimage.map_coordinates(a, [R, C], order=1, mode="grid-constant", cval=0.5)
        assert np.abs(resize(a, (oh, ow), mode="constant", cval=0.5, clip=False) - ref).max() < 1e-12


def test_clip_keeps_exact_cval_only():
    a = np.full((8, 8), 0.25)
    out = resize(a, (16, 16), mode="constant", cval=1.0)     # border blends towards 1 -> clipped back to 0.25
    assert np.allclose(out, 0.25)
    m = np.ones((8, 8), bool)                                 # all-true mask: clip range [1,1], cval 0 kept only if exact
    assert (resize(m, (16, 16), mode="constant", cval=0) > 0.9).all()


def test_bool_input_and_golden():
    g = np.load(GOLD)
    a = np.random.RandomState(3).rand(128, 128)
    assert np.allclose(resize(a, (200, 173), mode="reflect")[::10, ::10], g["up_reflect"], atol=1e-14)
    assert np.allclose(resize(a, (200, 173), mode="constant", cval=1.0)[::10, ::10], g["up_const"], atol=1e-14)
    assert np.allclose(resize(a, (77, 91), mode="constant", cval=0.5)[::7, ::7], g["down_const"], atol=1e-14)
    assert np.array_equal((resize(a > 0.5, (150, 150), mode="constant", cval=0) > 0.9)[::5, ::5], g["mask_up"])
