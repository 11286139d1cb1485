"""CPU: the restated est_pose / pnp_ransac control flow (oracle/recognition_oracle.py)."""
import numpy as np

from oracle.recognition_oracle import Pix2PoseOracle, get_boxes
from tests.planted import K_LM, OBJ, planted_case, rodrigues


def test_get_boxes_config1_roi():
    # 86x86 roi -> side 2*int(1.5*86/2) = 128, fully inside a 480x640 frame
    b = get_boxes(1.5, [197, 277, 283, 363], 480, 640)
    assert b == (176, 304, 256, 384, 176, 304, 256, 384, 0, 128, 0, 128)


def test_get_boxes_clipping_and_paste_offsets():
    b = get_boxes(1.5, [-20, -10, 90, 120], 480, 640)
    v1o, v2o, u1o, u2o, v1, v2, u1, u2, vv1, vv2, uu1, uu2 = b
    assert v1 == 0 and u1 == 0 and vv1 == -v1o and uu1 == -u1o
    assert (vv2 - vv1, uu2 - uu1) == (v2 - v1, u2 - u1)          # paste region matches the clipped crop
    b = get_boxes(1.5, [400, 560, 500, 660], 480, 640)
    assert b[5] == 480 and b[7] == 640 and (b[9] - b[8], b[11] - b[10]) == (b[5] - b[4], b[7] - b[6])
    # explicit centre and width cap (stage 2, recognition.py:110)
    b = get_boxes(1.5, np.array([10.5, 20.25, 90.0, 100.0]), 480, 640, ct=np.array([200, 300]), max_w=64)
    assert b[1] - b[0] == 64 and b[0] == 200 - 32


def test_small_roi_returns_sentinels():
    ora = Pix2PoseOracle(None, K_LM, 640, 480, OBJ)
    out = ora.est_pose(np.zeros((480, 640, 3), np.uint8), np.array([200, 300, 202, 302]))
    assert out[1] == -1 and out[2] == -1 and out[4] == -1 and out[0].shape == (1,)


def test_planted_pose_is_recovered():
    """2_1 normalise -> pnp_ransac de-normalise round trip + real cv2.solvePnPRansac: the planted R|t
    comes back through both stages (SURVEY §8c oracle self-validation)."""
    frame = np.random.RandomState(0).randint(0, 256, (480, 640, 3)).astype(np.uint8)
    R = rodrigues([0.4, -0.3, 0.2])
    t = np.array([15.0, -10.0, 700.0])
    # the ellipsoid projects to roughly 80x100 px around (cx + 15*fx/700, cy - 10*fy/700)
    roi = [182, 290, 286, 386]
    res, s1, s2, ora = planted_case(Pix2PoseOracle, frame, roi, R, t, th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15)
    assert not isinstance(res[1], int), "pose expected"
    ang = np.degrees(np.arccos(np.clip((np.trace(R.T @ res[2]) - 1) / 2, -1, 1)))
    assert ang < 2.0, ang                                    # uint8 XYZ quantisation + 10 % outliers
    assert np.linalg.norm(res[3] - t) / np.linalg.norm(t) < 0.02
    assert 0.3 < res[4] <= 1.5
    assert res[1].dtype == bool and res[1].shape == (480, 640) and res[0].dtype == np.uint8
