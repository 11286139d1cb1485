"""Multi-object detection streams: one weight set per object id resident on the GPU, one
shared engine + device pipeline.  Mirrors the per-object dispatch of tools/5_evaluation_bop_basic.py:206-225 (one
``recog.pix2pose`` per model id) and :289-323 (``obj_pix2pose[model_ids_list.index(obj_id)].est_pose(image, roi)`` per
ROI) -- but the whole stream goes through ONE device run (``p2p_pipeline_run_multi``): detections are grouped by object,
only the generator forwards run per group (each with its object's weights); crops, masks, correspondences, EPnP-RANSAC
and candidate selection run over all objects' detections at once, and each detection carries its own object's
``obj_param`` and thresholds (cfg_tless_paper.json:12 gives every T-LESS object its own ``outlier_th``)."""
import ctypes

import numpy as np

from . import _lib
from .recognition import DET_DTYPE, POSE_DTYPE, _Det, _Pose, pix2pose


class MultiObjectRecognizer:
    def __init__(self, weights, camK, res_x, res_y, obj_params, th_outlier=(0.15, 0.25, 0.35), th_inlier=0.15,
                 backbone="resnet50", precision="fp16x3", capacity=64, max_dets=64):
        """weights / obj_params: dict obj_id -> weight file (or dict of arrays) / 6-vector (bop_io.get_model_params).
        th_outlier: one list for all objects, or dict obj_id -> list (all of the same length)."""
        self.obj_ids = sorted(weights)
        self.models = {}
        for oid in self.obj_ids:
            th = th_outlier[oid] if isinstance(th_outlier, dict) else list(th_outlier)
            self.models[oid] = pix2pose(weights[oid], camK, res_x, res_y, obj_params[oid], th_outlier=th, th_inlier=th_inlier,
                                        backbone=backbone, precision=precision, capacity=capacity, max_dets=max_dets)
        if len({len(np.asarray(m.th_o).reshape(-1)) for m in self.models.values()}) != 1:
            raise ValueError("all objects must use the same number of outlier thresholds")
        if len({id(m.generator_train.engine) for m in self.models.values()}) != 1:
            raise ValueError("all objects must share one engine (same backbone, precision and capacity)")

    def set_camK(self, camK):
        for m in self.models.values():
            m.camK = camK                          # tools/5_evaluation_bop_basic.py:302

    def upload_frames(self, frames, n_dets=1):
        return next(iter(self.models.values())).upload_frames(frames, n_dets)

    def est_pose_stream(self, frames, rois, obj_ids, frame_ids=None, frames_dev=None):
        """Poses for a stream of detections in detector order.  Returns (records (n,16), status (n,)): records as
        ``PoseBatchResult.records`` with column 15 = stream index; detections of unknown objects get status -3."""
        rois = np.asarray(rois).reshape(-1, 4)
        obj_ids = np.asarray(obj_ids)
        n = len(rois)
        frame_ids = np.zeros(n, np.int64) if frame_ids is None else np.asarray(frame_ids)
        rec = np.zeros((n, 16))
        rec[:, 14] = -3                             # unknown object id
        rec[:, 15] = np.arange(n)
        groups = [(oid, np.nonzero(obj_ids == oid)[0]) for oid in np.unique(obj_ids) if oid in self.models]
        n_known = int(sum(len(idx) for _, idx in groups))
        if n_known == 0:
            return rec, rec[:, 14].astype(np.int32)
        first = self.models[groups[0][0]]
        if frames_dev is None:
            frames_dev = first.upload_frames(frames, n_known)
        dev, F, H, W, pipe0 = frames_dev
        pipe = first._pipeline(n_known)
        if pipe0.value != pipe.value:
            raise RuntimeError("frames_dev belongs to a pipeline that was rebuilt; upload the frames again")
        dets = (_Det * n_known)()
        da = np.frombuffer(dets, dtype=DET_DTYPE, count=n_known)
        order, counts, handles, at = [], [], [], 0
        for oid, idx in groups:
            m = self.models[oid]
            part = m._make_dets((H, W), rois[idx], frame_ids[idx], None)
            da[at:at + len(idx)] = np.frombuffer(part, dtype=DET_DTYPE, count=len(idx))
            at += len(idx)
            order.append(idx); counts.append(len(idx)); handles.append(m.generator_train._model)
        order = np.concatenate(order)
        poses = (_Pose * n_known)()
        ent = first._entry()
        ent["run"] += 1
        _lib.check(_lib.lib().p2p_pipeline_run_multi(pipe, (ctypes.c_void_p * len(handles))(*[h.value for h in handles]),
                                                     (ctypes.c_int * len(counts))(*counts), len(counts), dev, 0, F, H, W, dets,
                                                     n_known, 5.0, 100, 0.99, poses))
        a = np.frombuffer(poses, dtype=POSE_DTYPE, count=n_known)
        rec[order, :9] = a["R"]
        rec[order, 9:12] = a["t"]
        rec[order, 12], rec[order, 13], rec[order, 14] = a["n_inliers"], a["frac_inlier"], a["status"]
        return rec, rec[:, 14].astype(np.int32)


class AsyncBatcher:
    """Throughput loops: batch k+1 is prepared on the host (boxes, detection records) and queued on the device while batch k
    still computes.  Two device pipelines are used alternately (one run in flight per pipeline, ``p2p_pipeline_set_async`` /
    ``p2p_pipeline_wait``); both share the engine's stream, so the device simply runs the batches back to back and never
    waits for Python.  ``owner``: a ``pix2pose`` (one object) or a ``MultiObjectRecognizer`` (``submit`` then takes the
    detections' object ids).  The synchronous ``est_pose_batch`` / ``est_pose_stream`` path is untouched."""

    def __init__(self, owner, max_dets):
        self.multi = owner if isinstance(owner, MultiObjectRecognizer) else None
        self.first = next(iter(owner.models.values())) if self.multi else owner
        eng = self.first.generator_train.engine
        self.n_th = len(np.asarray(self.first.th_o).reshape(-1))
        self.pipes = []
        for _ in range(2):
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().p2p_pipeline_create(eng.handle, int(max_dets), self.n_th, ctypes.byref(h)))
            _lib.check(_lib.lib().p2p_pipeline_set_async(h, 1))
            self.pipes.append(h)
        self.turn = 0
        self.inflight = [None, None]

    def close(self):
        for t in (0, 1):
            if self.inflight[t] is not None:
                self.result(self.inflight[t])
        for h in self.pipes:
            _lib.lib().p2p_pipeline_destroy(h)
        self.pipes = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return sum(int(_lib.lib().p2p_pipeline_launch_count(h)) for h in self.pipes)

    def upload_frames(self, frames):
        """Queues the H2D copy of `frames` for the batch the NEXT ``submit`` runs (same pipeline, its copy stream)."""
        frames = np.ascontiguousarray(np.asarray(frames, np.uint8))
        if frames.ndim == 3:
            frames = frames[None]
        pipe = self.pipes[self.turn]
        dev = ctypes.c_void_p()
        _lib.check(_lib.lib().p2p_pipeline_upload_frames(pipe, frames.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), frames.shape[0],
                                                         frames.shape[1], frames.shape[2], ctypes.byref(dev)))
        return (dev, frames.shape[0], frames.shape[1], frames.shape[2], self.turn)

    def submit(self, frames_dev, rois, frame_ids, obj_ids=None):
        """Queues one batch; returns a ticket for ``result``.  `frames_dev`: handle of ``upload_frames`` (of this batcher, or of
        the owner for frames that stay resident)."""
        dev, F, H, W = frames_dev[:4]
        t = self.turn
        if self.inflight[t] is not None:
            raise RuntimeError("both pipelines are busy: collect a result first")
        rois = np.asarray(rois).reshape(-1, 4)
        frame_ids = np.asarray(frame_ids)
        n = len(rois)
        if self.multi is None:
            groups = [(self.first, np.arange(n))]
        else:
            obj_ids = np.asarray(obj_ids)
            groups = [(self.multi.models[o], np.nonzero(obj_ids == o)[0]) for o in np.unique(obj_ids) if o in self.multi.models]
        n_known = int(sum(len(idx) for _, idx in groups))
        dets = (_Det * max(n_known, 1))()
        da = np.frombuffer(dets, dtype=DET_DTYPE, count=max(n_known, 1))
        order, counts, handles, at = [], [], [], 0
        for m, idx in groups:
            part = m._make_dets((H, W), rois[idx], frame_ids[idx], None)
            da[at:at + len(idx)] = np.frombuffer(part, dtype=DET_DTYPE, count=len(idx))
            at += len(idx)
            order.append(idx); counts.append(len(idx)); handles.append(m.generator_train._model)
        poses = (_Pose * max(n_known, 1))()
        pipe = self.pipes[t]
        _lib.check(_lib.lib().p2p_pipeline_set_box_size(pipe, float(self.first.box_size)))
        if n_known:
            _lib.check(_lib.lib().p2p_pipeline_run_multi(pipe, (ctypes.c_void_p * len(handles))(*[h.value for h in handles]),
                                                         (ctypes.c_int * len(counts))(*counts), len(counts), dev, 0, F, H, W, dets,
                                                         n_known, 5.0, 100, 0.99, poses))
        ticket = (t, poses, n, n_known, np.concatenate(order) if order else np.zeros(0, np.int64), dets)
        self.inflight[t] = ticket
        self.turn ^= 1
        return ticket

    def result(self, ticket):
        """Blocks until the ticket's batch is done: (records (n,16) in submission order, status (n,))."""
        t, poses, n, n_known, order, _dets = ticket
        if self.inflight[t] is not ticket:
            raise RuntimeError("ticket already collected")
        if n_known:
            _lib.check(_lib.lib().p2p_pipeline_wait(self.pipes[t], poses))
        self.inflight[t] = None
        rec = np.zeros((n, 16))
        rec[:, 14] = -3
        rec[:, 15] = np.arange(n)
        if n_known:
            a = np.frombuffer(poses, dtype=POSE_DTYPE, count=n_known)
            rec[order, :9] = a["R"]
            rec[order, 9:12] = a["t"]
            rec[order, 12], rec[order, 13], rec[order, 14] = a["n_inliers"], a["frac_inlier"], a["status"]
        self.last_n_cand = np.frombuffer(poses, dtype=POSE_DTYPE, count=max(n_known, 1))["n_cand"][:n_known].copy()
        return rec, rec[:, 14].astype(np.int32)

    def last_forward_ms(self, ticket_slot):
        ms = ctypes.c_double()
        _lib.check(_lib.lib().p2p_pipeline_forward_ms(self.pipes[ticket_slot], ctypes.byref(ms)))
        return ms.value
