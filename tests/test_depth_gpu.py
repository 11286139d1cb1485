"""GPU parity of pix2pose_b200.depth (getXYZ / get_normal) against the numpy / scipy / cv2 restatement
of pix2pose_util/common_util.py:13-90.  Tolerances: getXYZ and the refine=False normals are the same double arithmetic
in the same order -> bit-exact; with refine=True the Gaussian smoothing sums in scipy's order but its weights come from
two exp() implementations -> 1e-9 abs on unit normals."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
K = dict(fx=572.4114, fy=573.57043, cx=325.2611, cy=242.04899)


def _depth(seed, holes):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:480, 0:640]
    d = 700.0 + 40.0 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + 0.05 * (xx - 320) + rng.normal(0, 0.3, (480, 640))
    if holes:
        d[rng.rand(480, 640) < 0.02] = 0.0
        d[100:120, 200:230] = 0.0
        d[5, 7] = np.nan
    return d


@pytest.mark.parametrize("bbox", [np.array([0]), np.array([60, 80, 300, 420])])
def test_getxyz_bit_exact(bbox):
    from oracle import depth_oracle as O
    from pix2pose_b200 import depth as D
    d = _depth(0, False)
    want = O.getXYZ(d, K["fx"], K["fy"], K["cx"], K["cy"], bbox)
    got = D.getXYZ(d, K["fx"], K["fy"], K["cx"], K["cy"], bbox)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("bbox", [np.array([0]), np.array([60, 80, 300, 420])])
def test_normals_without_refinement_bit_exact(bbox):
    from oracle import depth_oracle as O
    from pix2pose_b200 import depth as D
    d = _depth(1, False)
    want = O.get_normal(d, bbox=bbox, refine=False, **K)
    got = D.get_normal(d, bbox=bbox, refine=False, **K)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert np.allclose(np.linalg.norm(got, axis=2), 1.0, atol=1e-12)


@pytest.mark.parametrize("bbox", [np.array([0]), np.array([90, 150, 250, 400])])
def test_normals_with_refinement(bbox):
    from oracle import depth_oracle as O
    from pix2pose_b200 import depth as D
    d = _depth(2, True)
    want = O.get_normal(d, bbox=bbox, refine=True, **K)
    got = D.get_normal(d, bbox=bbox, refine=True, **K)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-9


def test_flat_plane_normal_and_degenerate_inputs():
    from pix2pose_b200 import depth as D
    n = D.get_normal(np.full((64, 96), 500.0), bbox=np.array([0]), refine=False, **K)
    assert np.abs(n[..., 2]).min() > 0.999                      # fronto-parallel plane: normals along +-z
    z = D.get_normal(np.zeros((32, 32)), bbox=np.array([0]), refine=False, **K)
    assert np.array_equal(z, np.zeros((32, 32, 3)))             # zero cross product -> norm set to 1 -> zeros
    assert D.getXYZ(np.zeros((8, 8)), 1.0, 1.0, 0.0, 0.0, np.array([2, 2, 2, 5])).shape == (0, 3, 3)


def test_against_golden_vectors_of_the_reference_functions():
    """tests/golden/depth_golden.npz = outputs of pix2pose_util/common_util.py itself (tests/golden/make_depth_golden.py)."""
    import os
    from pix2pose_b200 import depth as D
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "depth_golden.npz"))
    fx, fy, cx, cy = g["K"]
    box = g["bbox"]
    kw = dict(fx=fx, fy=fy, cx=cx, cy=cy)
    assert np.array_equal(D.getXYZ(g["depth_plain"], fx, fy, cx, cy), g["xyz_full"])
    assert np.array_equal(D.getXYZ(g["depth_plain"], fx, fy, cx, cy, box), g["xyz_box"])
    assert np.array_equal(D.get_normal(g["depth_plain"], bbox=np.array([0]), refine=False, **kw), g["normal_full"])
    assert np.array_equal(D.get_normal(g["depth_plain"], bbox=box, refine=False, **kw), g["normal_box"])
    assert np.abs(D.get_normal(g["depth_holes"].copy(), bbox=np.array([0]), refine=True, **kw) - g["normal_refined_full"]).max() <= 1e-9
    assert np.abs(D.get_normal(g["depth_holes"].copy(), bbox=box, refine=True, **kw) - g["normal_refined_box"]).max() <= 1e-9
