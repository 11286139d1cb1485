"""CPU checks of oracle/depth_oracle.py (restatement of pix2pose_util/common_util.py:13-90): analytic cases."""
import os

import numpy as np

from oracle import depth_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "depth_golden.npz")


def test_oracle_equals_the_reference_functions_on_the_golden_vectors():
    """tests/golden/depth_golden.npz holds outputs of the REFERENCE's own getXYZ / get_normal
    (pix2pose_util/common_util.py, run by tests/golden/make_depth_golden.py in the build container): this pins the
    restatement -- same numpy / scipy / cv2 calls, so the match is exact."""
    g = np.load(GOLD)
    fx, fy, cx, cy = g["K"]
    box = g["bbox"]
    assert np.array_equal(O.getXYZ(g["depth_plain"], fx, fy, cx, cy), g["xyz_full"])
    assert np.array_equal(O.getXYZ(g["depth_plain"], fx, fy, cx, cy, box), g["xyz_box"])
    kw = dict(fx=fx, fy=fy, cx=cx, cy=cy)
    assert np.array_equal(O.get_normal(g["depth_plain"], bbox=np.array([0]), refine=False, **kw), g["normal_full"])
    assert np.array_equal(O.get_normal(g["depth_plain"], bbox=box, refine=False, **kw), g["normal_box"])
    assert np.allclose(O.get_normal(g["depth_holes"].copy(), bbox=np.array([0]), refine=True, **kw), g["normal_refined_full"], atol=1e-12)
    assert np.allclose(O.get_normal(g["depth_holes"].copy(), bbox=box, refine=True, **kw), g["normal_refined_box"], atol=1e-12)


def test_getxyz_backprojects_pixels():
    d = np.full((10, 12), 2.0)
    xyz = O.getXYZ(d, 4.0, 5.0, 6.0, 5.0)
    assert xyz.shape == (10, 12, 3) and np.all(xyz[..., 2] == 2.0)
    assert xyz[5, 6, 0] == 0.0 and xyz[5, 6, 1] == 0.0
    assert xyz[5, 10, 0] == (10 - 6) * 2.0 / 4.0 and xyz[9, 6, 1] == (9 - 5) * 2.0 / 5.0
    box = np.array([2, 3, 7, 9])
    assert np.array_equal(O.getXYZ(d, 4.0, 5.0, 6.0, 5.0, box), xyz[2:7, 3:9])
    # the pixel tables are int16: a fractional principal point truncates toward zero (common_util.py:16-18)
    assert O.getXYZ(d, 1.0, 1.0, 5.5, 0.0)[0, 5, 0] == 0.0 and O.getXYZ(d, 1.0, 1.0, 5.5, 0.0)[0, 4, 0] == -2.0


def test_normals_of_planes():
    n = O.get_normal(np.full((40, 50), 800.0), fx=500.0, fy=500.0, cx=25.0, cy=20.0, refine=False)
    assert np.allclose(np.abs(n[..., 2]), 1.0, atol=1e-12) and np.allclose(n[..., :2], 0.0, atol=1e-12)
    yy, xx = np.mgrid[0:40, 0:50]
    tilted = 800.0 + 2.0 * (xx - 25)
    nt = O.get_normal(tilted, fx=500.0, fy=500.0, cx=25.0, cy=20.0, refine=False)
    assert np.allclose(np.linalg.norm(nt, axis=2), 1.0) and np.abs(nt[20, 25, 0]) > 0.5      # tilted about the vertical axis
    holes = tilted.copy()
    holes[10:14, 10:14] = 0.0
    nr = O.get_normal(holes, fx=500.0, fy=500.0, cx=25.0, cy=20.0, refine=True, bbox=np.array([5, 5, 30, 40]))
    assert nr.shape == (25, 35, 3) and np.isfinite(nr).all()
