"""Per-image recognition loop of ``tools/5_evaluation_bop_basic.py:281-349``: detections of one image
-> poses -> detection-score x inlier-fraction x mask-IoU scoring (score_type 2) -> normalise, sort, ViVo filter -> BOP
result rows.  The reference calls ``est_pose`` once per ROI; here the ROIs that survive the (result-independent)
candidate filter are grouped per object and run as one device batch each.  ``save_bop_results`` writes the BOP CSV
(bop_toolkit ``inout.save_bop_results`` format), so no bop_toolkit is needed for the output side."""
import time

import numpy as np


def select_rois(rois, obj_ids, obj_id_targets, inst_counts, cand_factor):
    """tools/5_evaluation_bop_basic.py:289-299: which detections get a pose estimate (depends only on the detector output)."""
    keep = []
    inst_count_pred = np.zeros(len(inst_counts), int)
    for r_id, roi in enumerate(rois):
        if roi[0] == -1 and roi[1] == -1:
            continue
        obj_id = obj_ids[r_id]
        if obj_id not in obj_id_targets:
            continue
        g = list(obj_id_targets).index(obj_id)
        if inst_count_pred[g] > inst_counts[g] * cand_factor:
            continue
        inst_count_pred[g] += 1
        keep.append(r_id)
    return keep


def est_pose_loop(recognizers, image, rois, obj_orders, r_ids):
    """Reference-style backend: one ``est_pose`` call per ROI (works with any object exposing the reference surface)."""
    return {r: recognizers[obj_orders[r]].est_pose(image, np.asarray(rois[r]).astype(int)) for r in r_ids}


class DeviceMaskIoU:
    """Stands in for ``mask_pred`` in the batched backend when the detector masks are known: |mask_pred & m| and
    |mask_pred | m| were counted on the device, so no mask crosses the bus."""

    def __init__(self, inter, union):
        self.inter, self.union = int(inter), int(union)


def est_pose_batched(recognizers, image, rois, obj_orders, r_ids, masks=None):
    """B200 backend: one ``est_pose_batch`` per object; returns the same 6-tuples as ``est_pose``.  With ``masks``
    ((H,W,n_rois) detector masks, the reference's layout) the mask entry is a ``DeviceMaskIoU``."""
    out = {}
    by_obj = {}
    for r in r_ids:
        by_obj.setdefault(obj_orders[r], []).append(r)
    H, W = image.shape[0], image.shape[1]
    for order, rs in by_obj.items():
        rec = recognizers[order]
        res = rec.est_pose_batch(image, [np.asarray(rois[r]).astype(int) for r in rs])
        iou = None
        if masks is not None and masks.shape[:2] == (H, W):
            iou = res.mask_iou(np.ascontiguousarray(np.moveaxis(np.asarray(masks)[:, :, rs], 2, 0)))
        for i, r in enumerate(rs):
            if res.status[i] != 1:
                out[r] = (None, -1, -1, -1, -1, res.bbox_t[i])
                continue
            if iou is not None:
                out[r] = (None, DeviceMaskIoU(iou[0][i], iou[1][i]), res.R[i], res.t[i], res.frac_inlier[i], res.bbox_t[i])
                continue
            xyz, mask, bx = res.crop(i)
            full = np.zeros((H, W), bool)
            full[bx[4]:bx[5], bx[6]:bx[7]] = mask
            out[r] = (xyz, full, res.R[i], res.t[i], res.frac_inlier[i], res.bbox_t[i])
    return out


def recognize_image(recognizers, image, rois, obj_orders, obj_ids, scores, masks, obj_id_targets, inst_counts, cam_K,
                    scene_id=0, im_id=0, cand_factor=2, score_type=2, task_type="2", detect_type="rcnn", backend=None, t_start=None):
    """Lines 281-349 of the reference driver for ONE image.  Returns the list of BOP result dicts it would append."""
    t1 = time.time() if t_start is None else t_start
    backend = backend or (est_pose_batched if hasattr(recognizers[0], "est_pose_batch") else est_pose_loop)
    for rec in recognizers:
        rec.camK = np.asarray(cam_K).reshape(3, 3)                               # :302
    keep = select_rois(rois, obj_ids, obj_id_targets, inst_counts, cand_factor)
    use_dev_iou = backend is est_pose_batched and score_type == 2 and detect_type == "rcnn" and masks is not None
    poses = backend(recognizers, image, rois, obj_orders, keep, masks=masks) if use_dev_iou else backend(recognizers, image, rois, obj_orders, keep)
    result_score, result_objid, result_R, result_t = [], [], [], []
    for r_id in keep:
        img_pred, mask_pred, rot_pred, tra_pred, frac_inlier, bbox_t = poses[r_id]
        if np.isscalar(frac_inlier) and frac_inlier == -1:                        # :305
            continue
        if score_type == 2 and detect_type == "rcnn":                            # :307-316
            m = masks[:, :, r_id]
            if m.shape[:2] != image.shape[:2]:
                raise ValueError("detector mask must have the image size (the reference resizes with skimage; not reproduced)")
            if isinstance(mask_pred, DeviceMaskIoU):
                union, inter = mask_pred.union, mask_pred.inter
            else:
                union, inter = np.sum(np.logical_or(m, mask_pred)), np.sum(np.logical_and(m, mask_pred))
            mask_iou = 0 if union <= 0 else inter / union
            score = scores[r_id] * frac_inlier * mask_iou * union
        else:
            score = scores[r_id]
        result_score.append(score); result_objid.append(obj_ids[r_id]); result_R.append(rot_pred); result_t.append(tra_pred)
    if len(result_score) == 0:
        return []
    result_score = np.array(result_score)
    result_score = result_score / np.max(result_score)                          # :327
    sorted_id = np.argsort(1 - result_score)                                    # :328
    time_spend = time.time() - t1
    n_inst = np.sum(inst_counts)
    inst_count_est = np.zeros(len(inst_counts))
    total_inst, results = 0, []
    for result_id in sorted_id:                                                 # :335-349
        obj_id = result_objid[result_id]
        g = list(obj_id_targets).index(obj_id)
        inst_count_est[g] += 1
        if task_type == "2" and inst_count_est[g] > inst_counts[g]:
            continue
        results.append({"scene_id": scene_id, "im_id": im_id, "obj_id": obj_id, "score": result_score[result_id],
                        "R": np.asarray(result_R[result_id]).flatten(), "t": np.asarray(result_t[result_id]).flatten(),
                        "time": time_spend})
        total_inst += 1
        if task_type == "2" and total_inst > n_inst:
            break
    return results


def save_bop_results(path, results, version="bop19"):
    """BOP result CSV: scene_id,im_id,obj_id,score,R (9 values, row-major, space separated),t (3 values, mm),time."""
    lines = ["scene_id,im_id,obj_id,score,R,t,time"]
    for r in results:
        lines.append("{},{},{},{},{},{},{}".format(
            r["scene_id"], r["im_id"], r["obj_id"], r["score"],
            " ".join(map(str, np.asarray(r["R"]).flatten().tolist())),
            " ".join(map(str, np.asarray(r["t"]).flatten().tolist())), r.get("time", -1)))
    with open(path, "w") as f:
        f.write("\n".join(lines))
