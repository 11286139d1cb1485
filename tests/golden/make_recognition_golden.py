"""Generates tests/golden/recognition_golden.npz by running the REFERENCE's own methods
``pix2pose.get_boxes`` (pix2pose_model/recognition.py:28-69) and ``pix2pose.pnp_ransac`` (:195-224) on seeded inputs.

recognition.py cannot be imported here (it imports keras / tensorflow / skimage at module level), but these two methods
only use numpy and cv2: the script parses the file, compiles JUST those two function definitions out of it in memory
(``ast``; nothing of the reference is written into this repository) and calls them on a stand-in object that carries the
attributes they read (box_size, obj_scale, obj_ct, camK, th_i).  Run in the build container, where /root/reference
exists:  python tests/golden/make_recognition_golden.py"""
import ast
import os
import sys
import types

import cv2
import numpy as np

SRC = "/root/reference/pix2pose_model/recognition.py"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.planted import K_LM, OBJ, rodrigues   # noqa: E402


def reference_methods():
    tree = ast.parse(open(SRC).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "pix2pose"][0]
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("get_boxes", "pnp_ransac")]
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {"np": np, "cv2": cv2}
    exec(compile(mod, SRC, "exec"), ns)
    return ns["get_boxes"], ns["pnp_ransac"]


def planted_xyz_frame(seed, rvec, t, outlier_frac, H=480, W=640, box=(150, 330, 230, 410)):
    """A uint8 'XYZ image' of an ellipsoid under pose (R|t): the frame est_pose pastes its stage-2 output into (:152-154)."""
    rng = np.random.RandomState(seed)
    R = rodrigues(np.asarray(rvec, float))
    v1, v2, u1, u2 = box
    frame = np.zeros((H, W, 3), np.uint8)
    prob = np.ones((H, W))
    non_zero = np.zeros((v2 - v1, u2 - u1), bool)
    fx, fy, cx, cy = K_LM[0, 0], K_LM[1, 1], K_LM[0, 2], K_LM[1, 2]
    Rt = R.T
    for v in range(v1, v2):
        for u in range(u1, u2):
            d = Rt @ np.array([(u - cx) / fx, (v - cy) / fy, 1.0])           # ray in the object frame
            o = -Rt @ np.asarray(t, float)
            A = np.sum((d / OBJ[:3]) ** 2); B = 2 * np.sum(d * o / OBJ[:3] ** 2); C = np.sum((o / OBJ[:3]) ** 2) - 1
            disc = B * B - 4 * A * C
            if disc <= 0:
                continue
            p = o + (-B - np.sqrt(disc)) / (2 * A) * d                          # first hit with the ellipsoid (object units)
            q = np.clip((p / OBJ[:3] + 1) / 2 * 255, 0, 255)
            if rng.rand() < outlier_frac:
                q = rng.uniform(0, 255, 3)
            frame[v, u] = q.astype(np.uint8)
            non_zero[v - v1, u - u1] = True
            prob[v, u] = rng.uniform(0.0, 0.3)
    return frame, prob[v1:v2, u1:u2], non_zero, box


def main():
    get_boxes, pnp_ransac = reference_methods()
    out = {}
    # ---- get_boxes: rois inside, clipped on every side, non-square, with / without centre and max_w
    self_ = types.SimpleNamespace(box_size=1.5)
    cases, res = [], []
    rng = np.random.RandomState(3)
    for i in range(40):
        cy, cx = rng.randint(-20, 500), rng.randint(-20, 660)
        h, w = rng.randint(10, 300), rng.randint(10, 300)
        bbox = np.array([cy - h // 2, cx - w // 2, cy + h // 2, cx + w // 2])
        if i % 3 == 0:
            ct, mw = np.array([rng.randint(0, 480), rng.randint(0, 640)]), float(rng.randint(40, 400))
        else:
            ct, mw = np.array([-1]), 9999
        r = get_boxes(self_, bbox, 480, 640, ct=ct, max_w=mw)
        cases.append(np.concatenate([bbox, [ct[0], ct[-1] if len(ct) > 1 else -1, mw]]))
        res.append(np.array(r, float))
    out["boxes_in"], out["boxes_out"] = np.array(cases, float), np.array(res)
    # ---- pnp_ransac on planted poses
    obj = types.SimpleNamespace(obj_scale=OBJ[:3], obj_ct=OBJ[3:], camK=K_LM, th_i=0.15)
    for i, (seed, rv, t, of) in enumerate([(0, [0.4, -0.3, 0.2], [15.0, -10.0, 700.0], 0.0),
                                           (1, [-0.2, 0.5, 1.0], [-30.0, 20.0, 900.0], 0.3),
                                           (2, [0.1, 0.1, -0.4], [0.0, 0.0, 600.0], 0.1)]):
        frame, prob, nz, box = planted_xyz_frame(seed, rv, t, of)
        R, tr, mask, n_inl = pnp_ransac(obj, frame, prob, nz, box[0], box[1], box[2], box[3])
        out["pnp%d_frame_crop" % i] = frame[box[0]:box[1], box[2]:box[3]]
        out["pnp%d_prob" % i], out["pnp%d_nonzero" % i], out["pnp%d_box" % i] = prob, nz, np.array(box)
        out["pnp%d_R" % i], out["pnp%d_t" % i], out["pnp%d_mask" % i], out["pnp%d_ninl" % i] = R, np.asarray(tr, float), np.asarray(mask), n_inl
        out["pnp%d_pose_true" % i] = np.concatenate([rv, t])
    # too few correspondences -> the :214-215 sentinel
    few = np.zeros((480, 640, 3), np.uint8)
    R, tr, mask, n_inl = pnp_ransac(obj, few, np.ones((20, 20)), np.zeros((20, 20), bool), 100, 120, 100, 120)
    out["pnp_few_R"], out["pnp_few_t"], out["pnp_few_ninl"] = R, np.asarray(tr, float), n_inl
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "recognition_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "pnp inliers:", [int(out["pnp%d_ninl" % i]) for i in range(3)], "few:", n_inl)


if __name__ == "__main__":
    main()
