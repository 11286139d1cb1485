"""GPU probe: CPU-oracle pipeline vs GPU pipeline on real network outputs, candidate by candidate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.net_oracle import NetOracle
from oracle.recognition_oracle import Pix2PoseOracle
from pix2pose_b200 import weights as W
from pix2pose_b200.recognition import pix2pose, _Pose
from tests.planted import K_LM, OBJ

TH = dict(th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15)
w = W.synthetic_weights("resnet50", 1)
r = pix2pose(w, K_LM, 640, 480, OBJ, backbone="resnet50", capacity=16, max_dets=16, **TH)
ora = Pix2PoseOracle(NetOracle(w, "resnet50"), K_LM, 640, 480, OBJ, **TH)
frame = np.random.RandomState(0).randint(0, 256, (480, 640, 3)).astype(np.uint8)
frame[150:330, 230:410] = (frame[150:330, 230:410] // 4 + 100).astype(np.uint8)
for roi in ([197, 277, 283, 363], [100, 200, 260, 330]):
    ora.trace = {}
    want = ora.est_pose(frame, np.array(roi))
    got = r.est_pose(frame, np.array(roi))
    print("roi", roi, "bbox_t", list(want[5]), list(got[5]))
    print(" x1 equal", np.array_equal(r.debug_fetch(3, 0), ora.trace["x1"].astype(np.float32)))
    d1 = np.abs(r.debug_fetch(1, 0) - ora.trace["decode1"][0]).max()
    print(" decode1 maxdiff %.2e" % d1)
    for k in range(len(ora.trace["x2"])):
        x2g = r.debug_fetch(4, k)
        print("  cand", k, "x2 equal", np.array_equal(x2g, ora.trace["x2"][k].astype(np.float32)), "x2 maxdiff %.3f" % np.abs(x2g - ora.trace["x2"][k]).max(),
              "decode2 maxdiff %.2e" % np.abs(r.debug_fetch(2, k) - ora.trace["decode2"][k]).max(), "box", ora.trace["boxes2"][k][:4])
    for c in ora.trace["cands"]:
        p = _Pose(); p.best_cand = c["cid"]; p.best_box[:] = [int(v) for v in c["box"]]
        xyz, mask, _ = r._fetch_crop(0, p)
        d = np.abs(xyz.astype(int) - c["xyz_u8"].astype(int))
        print("  cand", c["cid"], "u8 diff>0 %.4f diff>1 %.4f" % ((d > 0).mean(), (d > 1).mean()), "n_inl", c["n_inliers"], "dist %.3f" % c["dist"])
    if not isinstance(want[1], int):
        d = np.abs(want[0].astype(int) - got[0].astype(int)) if want[0].shape == got[0].shape else None
        print(" returned crop diff>1", None if d is None else float((d > 1).mean()), "R equal", np.abs(want[2] - got[2]).max(), "frac", want[4], got[4])
