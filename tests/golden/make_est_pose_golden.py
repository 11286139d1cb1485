"""Generates tests/golden/est_pose_golden.npz by running the REFERENCE's own ``pix2pose.est_pose``
(pix2pose_model/recognition.py:70-193, with its ``get_boxes`` :28-69 and ``pnp_ransac`` :195-224) on planted-pose cases.

recognition.py cannot be imported (keras / tensorflow / skimage at module level), so the three method definitions are
compiled out of the file in memory (``ast``; nothing of the reference is copied into this repository) and called on a
stand-in object.  Two names the methods use come from outside:
  * ``resize``  -> oracle/resize_oracle.py (scikit-image is absent; this is the one UNPINNED semantic);
  * ``self.generator_train.predict`` -> tests/planted.py PlantedGenerator (analytic XYZ maps of an ellipsoid under a known
    pose), the same stub the oracle is given in the test, so the comparison isolates est_pose's own control flow:
    crops, masks, refined boxes, uint8 quantisation, candidate selection, the Q1-Q3 quirks.
Run in the build container:  python tests/golden/make_est_pose_golden.py"""
import ast
import os
import sys
import types

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
np.int = int                                           # alias numpy < 1.24 provided (recognition.py:79, :127, :191)
from oracle.recognition_oracle import Pix2PoseOracle   # noqa: E402
from oracle.resize_oracle import resize                # noqa: E402
from tests.planted import K_LM, OBJ, PlantedGenerator, planted_case, rodrigues   # noqa: E402

SRC = "/root/reference/pix2pose_model/recognition.py"
TH = dict(th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15)
CASES = [(0, [182, 290, 286, 386], [0.4, -0.3, 0.2], [15.0, -10.0, 700.0]),
         (1, [150, 250, 290, 370], [-0.2, 0.5, 1.0], [-20.0, 10.0, 800.0]),
         (2, [-20, -10, 120, 140], [0.1, 0.1, -0.4], [-260.0, -240.0, 900.0]),     # roi clipped at the top-left corner
         (3, [200, 300, 210, 306], [0.0, 0.0, 0.0], [0.0, 0.0, 700.0])]            # tiny roi


def reference_class():
    tree = ast.parse(open(SRC).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "pix2pose"][0]
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("get_boxes", "pnp_ransac", "est_pose")]
    ns = {"np": np, "cv2": cv2, "resize": resize}
    exec(compile(ast.Module(body=fns, type_ignores=[]), SRC, "exec"), ns)
    return type("RefPix2Pose", (), {k: ns[k] for k in ("get_boxes", "pnp_ransac", "est_pose")})


def frame():
    f = np.random.RandomState(0).randint(0, 256, (480, 640, 3)).astype(np.uint8)
    f[150:330, 230:410] = (f[150:330, 230:410] // 4 + 100).astype(np.uint8)
    return f


def main():
    Ref = reference_class()
    f = frame()
    out = {}
    for cid, roi, rv, t in CASES:
        R = rodrigues(rv)
        _, s1, s2, _ = planted_case(Pix2PoseOracle, f, roi, R, np.array(t), seed=cid, **TH)     # the planted network outputs
        ref = Ref()
        ref.camK, ref.res_x, ref.res_y = K_LM, 640, 480
        ref.th_o, ref.th_i, ref.box_size = TH["th_outlier"], TH["th_inlier"], 1.5
        ref.obj_scale, ref.obj_ct = OBJ[:3], OBJ[3:]
        ref.generator_train = PlantedGenerator(s1, s2)
        img_pred, mask_pred, rot, tra, frac, bbox_t = ref.est_pose(f, np.array(roi))
        ok = not (np.isscalar(mask_pred) and mask_pred == -1)
        out["c%d_roi" % cid], out["c%d_pose_true" % cid] = np.array(roi), np.concatenate([rv, t])
        # (the planted maps are not stored: the tests regenerate them with the same seeded planted_case call)
        out["c%d_ok" % cid] = np.array(ok)
        out["c%d_img_pred" % cid] = np.asarray(img_pred)
        out["c%d_bbox_t" % cid] = np.asarray(bbox_t)
        if ok:
            out["c%d_mask" % cid], out["c%d_R" % cid], out["c%d_t" % cid], out["c%d_frac" % cid] = np.packbits(mask_pred), rot, np.asarray(tra, float), np.array(frac)
        print("case", cid, "ok" if ok else "sentinel", "frac_inlier", frac, "bbox_t", list(np.asarray(bbox_t)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "est_pose_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
