"""CPU: the per-image recognition loop (pix2pose_b200/evaluation.py) against a literal restatement of
tools/5_evaluation_bop_basic.py:286-349 driven by the same fake ``est_pose``."""
import numpy as np

from pix2pose_b200 import evaluation as E


class FakeRec:
    """Deterministic stand-in exposing the reference surface; results depend only on (object, roi)."""

    def __init__(self, oid):
        self.oid, self.camK = oid, None

    def est_pose(self, image, roi):
        h = int(roi[0] * 7 + roi[1] * 13 + self.oid * 101)
        if h % 5 == 0:
            return np.zeros(1), -1, -1, -1, -1, np.array(roi)
        mask = np.zeros(image.shape[:2], bool)
        mask[max(roi[0], 0):roi[2], max(roi[1], 0):roi[3]] = True
        R = np.eye(3) * (1 + (h % 7) * 0.01)
        return np.zeros((2, 2, 3), np.uint8), mask, R, np.array([h % 11, h % 13, 500.0 + h % 17]), 0.1 + (h % 9) / 10.0, np.array(roi)


def reference_loop(recs, image, rois, obj_orders, obj_ids, scores, masks, targets, inst_counts, cam_K, cand_factor, score_type, task_type):
    """Literal per-ROI loop of the reference driver (:286-349), sequential est_pose calls."""
    inst_count_est = np.zeros(len(inst_counts)); inst_count_pred = np.zeros(len(inst_counts))
    rs, ro, rR, rt = [], [], [], []
    for r_id, roi in enumerate(rois):
        if roi[0] == -1 and roi[1] == -1:
            continue
        obj_id = obj_ids[r_id]
        if obj_id not in targets:
            continue
        g = targets.index(obj_id)
        if inst_count_pred[g] > inst_counts[g] * cand_factor:
            continue
        inst_count_pred[g] += 1
        recs[obj_orders[r_id]].camK = cam_K.reshape(3, 3)
        img_pred, mask_pred, R, t, frac, bbox_t = recs[obj_orders[r_id]].est_pose(image, np.asarray(roi).astype(int))
        if isinstance(frac, int) and frac == -1:
            continue
        if score_type == 2:
            m = masks[:, :, r_id]
            union = np.sum(np.logical_or(m, mask_pred))
            iou = 0 if union <= 0 else np.sum(np.logical_and(m, mask_pred)) / union
            score = scores[r_id] * frac * iou * union
        else:
            score = scores[r_id]
        rs.append(score); ro.append(obj_id); rR.append(R); rt.append(t)
    if not rs:
        return []
    rs = np.array(rs); rs = rs / np.max(rs); order = np.argsort(1 - rs)
    out, total, n_inst = [], 0, np.sum(inst_counts)
    for i in order:
        g = targets.index(ro[i]); inst_count_est[g] += 1
        if task_type == "2" and inst_count_est[g] > inst_counts[g]:
            continue
        out.append((ro[i], rs[i], rR[i].flatten(), rt[i].flatten())); total += 1
        if task_type == "2" and total > n_inst:
            break
    return out


def detections():
    """Seeded detector output of one image (shared with tests/golden/make_evaluation_golden.py)."""
    rng = np.random.RandomState(0)
    image = rng.randint(0, 256, (120, 160, 3)).astype(np.uint8)
    targets, inst_counts = [3, 7, 9], [1, 2, 1]
    n = 14
    rois = [[rng.randint(0, 60), rng.randint(0, 80), rng.randint(61, 120), rng.randint(81, 160)] for _ in range(n)]
    rois[4] = [-1, -1, 5, 5]
    obj_ids = [int(rng.choice([3, 7, 9, 11])) for _ in range(n)]
    obj_orders = [targets.index(o) if o in targets else 0 for o in obj_ids]
    scores = rng.uniform(0.5, 1.0, n)
    masks = rng.rand(120, 160, n) > 0.5
    K = np.array([572.4, 0, 80, 0, 573.5, 60, 0, 0, 1.0])
    return image, targets, inst_counts, rois, obj_ids, obj_orders, scores, masks, K


def test_recognize_image_equals_the_reference_drivers_own_loop():
    """tests/golden/evaluation_golden.npz: lines 281-349 of tools/5_evaluation_bop_basic.py exec'd on these detections
    (tests/golden/make_evaluation_golden.py).  Same rows, same order, same scores, for both score types and both tasks."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "evaluation_golden.npz"))
    image, targets, inst_counts, rois, obj_ids, obj_orders, scores, masks, K = detections()
    for score_type in (1, 2):
        for task_type in ("1", "2"):
            recs = [FakeRec(t) for t in targets]
            got = E.recognize_image(recs, image, rois, obj_orders, obj_ids, scores, masks, targets, inst_counts, K, scene_id=5, im_id=9,
                                    cand_factor=2, score_type=score_type, task_type=task_type, backend=E.est_pose_loop)
            key = "s%d_t%s" % (score_type, task_type)
            assert [r["obj_id"] for r in got] == list(g[key + "_obj"]) and len(got) > 0
            assert np.array_equal(np.array([r["score"] for r in got]), g[key + "_score"])
            assert np.array_equal(np.array([r["R"] for r in got]), g[key + "_R"])
            assert np.array_equal(np.array([r["t"] for r in got]), g[key + "_t"])


def test_recognize_image_matches_reference_loop(tmp_path):
    image, targets, inst_counts, rois, obj_ids, obj_orders, scores, masks, K = detections()
    recs = [FakeRec(3), FakeRec(7), FakeRec(9)]
    for score_type in (1, 2):
        for task_type in ("1", "2"):
            want = reference_loop(recs, image, rois, obj_orders, obj_ids, scores, masks, targets, inst_counts, K, 2, score_type, task_type)
            got = E.recognize_image(recs, image, rois, obj_orders, obj_ids, scores, masks, targets, inst_counts, K, scene_id=5, im_id=9,
                                    cand_factor=2, score_type=score_type, task_type=task_type, backend=E.est_pose_loop)
            assert len(got) == len(want) and len(got) > 0
            for g, w in zip(got, want):
                assert g["obj_id"] == w[0] and np.isclose(g["score"], w[1]) and np.array_equal(g["R"], w[2]) and np.array_equal(g["t"], w[3])
                assert g["scene_id"] == 5 and g["im_id"] == 9
    E.save_bop_results(tmp_path / "r.csv", got)
    lines = open(tmp_path / "r.csv").read().split("\n")
    assert lines[0] == "scene_id,im_id,obj_id,score,R,t,time" and len(lines) == len(got) + 1
    assert len(lines[1].split(",")[4].split(" ")) == 9 and len(lines[1].split(",")[5].split(" ")) == 3
