"""Multi-object detection streams: one weight set per object id resident on the
GPU, one shared engine + pipeline, detections grouped by object so that each object's detections run as one
batch.  Mirrors the per-object dispatch of tools/5_evaluation_bop_basic.py:206-225 (one ``recog.pix2pose`` per
model id) and :289-323 (``obj_pix2pose[model_ids_list.index(obj_id)].est_pose(image, roi)`` per ROI)."""
import numpy as np

from .recognition import pix2pose


class MultiObjectRecognizer:
    def __init__(self, weights, camK, res_x, res_y, obj_params, th_outlier=(0.15, 0.25, 0.35), th_inlier=0.15,
                 backbone="resnet50", precision="fp16x3", capacity=64, max_dets=64):
        """weights / obj_params: dict obj_id -> weight file (or dict of arrays) / 6-vector (bop_io.get_model_params)."""
        self.obj_ids = sorted(weights)
        self.models = {}
        for oid in self.obj_ids:
            th = th_outlier[oid] if isinstance(th_outlier, dict) else list(th_outlier)
            self.models[oid] = pix2pose(weights[oid], camK, res_x, res_y, obj_params[oid], th_outlier=th, th_inlier=th_inlier,
                                        backbone=backbone, precision=precision, capacity=capacity, max_dets=max_dets)

    def set_camK(self, camK):
        for m in self.models.values():
            m.camK = camK                          # tools/5_evaluation_bop_basic.py:302

    def est_pose_stream(self, frames, rois, obj_ids, frame_ids=None):
        """Poses for a stream of detections in detector order.  Returns (records (n,16), status (n,)): records as
        ``PoseBatchResult.records`` with column 15 = stream index."""
        rois = np.asarray(rois).reshape(-1, 4)
        obj_ids = np.asarray(obj_ids)
        n = len(rois)
        frame_ids = np.zeros(n, np.int64) if frame_ids is None else np.asarray(frame_ids)
        rec = np.zeros((n, 16))
        rec[:, 14] = -3                             # unknown object id
        uploaded = {}                               # pipeline -> device frames: one H2D copy per (shared) pipeline
        for oid in np.unique(obj_ids):
            if oid not in self.models:
                continue
            idx = np.nonzero(obj_ids == oid)[0]
            m = self.models[oid]
            pipe = m._pipeline(n)
            if id(pipe) not in uploaded:
                uploaded[id(pipe)] = m.upload_frames(frames, n)
            res = m.est_pose_batch(None, rois[idx], frame_ids[idx], frames_dev=uploaded[id(pipe)])
            r = res.records()
            r[:, 15] = idx
            rec[idx] = r
        rec[:, 15] = np.arange(n)
        return rec, rec[:, 14].astype(np.int32)
