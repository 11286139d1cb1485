"""GPU bring-up of the est_pose device pipeline against the CPU oracle (same network outputs)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle.recognition_oracle import Pix2PoseOracle
from pix2pose_b200 import weights as W
from pix2pose_b200.recognition import pix2pose

K = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])
OBJ = np.array([50., 40., 60., 0., 0., 0.])


def main():
    bb = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
    w = W.synthetic_weights(bb, 1)
    rec = pix2pose(w, K, 640, 480, OBJ, th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15, backbone=bb)
    ora = Pix2PoseOracle(rec.generator_train, K, 640, 480, OBJ, th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15)
    rng = np.random.RandomState(0)
    frame = rng.randint(0, 256, (480, 640, 3)).astype(np.uint8)
    # smooth-ish content so that masks are not pure noise
    frame[150:330, 230:410] = (frame[150:330, 230:410] // 4 + 100).astype(np.uint8)
    rois = [[197, 277, 283, 363], [100, 200, 260, 330], [-20, -10, 90, 120], [400, 560, 500, 660], [200, 300, 203, 302],
            [50, 60, 300, 420]]
    for roi in rois:
        ora.trace = {}
        t0 = time.time()
        o = ora.est_pose(frame, np.array(roi))
        t1 = time.time()
        g = rec.est_pose(frame, np.array(roi))
        t2 = time.time()
        print("roi", roi, "oracle %.0f ms  gpu %.0f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
        if "x1" in ora.trace:
            x1 = rec.debug_fetch(3, 0)
            print("   x1 max diff %.3e" % np.abs(x1 - ora.trace["x1"]).max())
        if "x2" in ora.trace:
            for k in range(len(ora.trace["x2"])):
                x2 = rec.debug_fetch(4, k)
                print("   x2[%d] max diff %.3e  (n diff>1e-6: %d)" % (k, np.abs(x2 - ora.trace["x2"][k]).max(),
                                                                      int((np.abs(x2 - ora.trace["x2"][k]) > 1e-6).sum())))
        for c in ora.trace.get("cands", []):
            print("   oracle cand %d: n_inl %d n_non_gray %d dist %.3f box %s" % (c["cid"], c["n_inliers"], c["n_non_gray"], c["dist"], list(c["box"])[:8]))
        print("   bbox_t oracle %s gpu %s" % (list(o[5]), list(g[5])))
        if isinstance(o[1], int) or isinstance(g[1], int):
            print("   sentinel: oracle %s gpu %s ; img shapes %s %s ; img diff %s" % (
                o[1], g[1] if isinstance(g[1], int) else "array", np.shape(o[0]), np.shape(g[0]),
                np.abs(np.asarray(o[0], np.float64) - np.asarray(g[0], np.float64)).max() if np.shape(o[0]) == np.shape(g[0]) else "n/a"))
            continue
        print("   img_pred equal: %s (max diff %d)  mask equal: %s (%d vs %d px)" % (
            np.array_equal(o[0], g[0]), int(np.abs(o[0].astype(int) - g[0].astype(int)).max()) if o[0].shape == g[0].shape else -1,
            np.array_equal(o[1], g[1]), int(o[1].sum()), int(g[1].sum())))
        print("   frac_inlier oracle %.5f gpu %.5f | dR max %.3e  dt max %.3e" % (o[4], g[4], np.abs(o[2] - g[2]).max(), np.abs(o[3] - g[3]).max()))


if __name__ == "__main__":
    main()
