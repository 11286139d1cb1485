"""Drop-in for ``pix2pose_model/recognition.py`` of kirumang/Pix2Pose on B200.

``pix2pose(weight_fn, camK, res_x, res_y, obj_param, ...)`` keeps the reference's constructor,
attributes (``camK`` is re-assigned by callers before every call,
tools/5_evaluation_bop_basic.py:302), ``get_boxes``, ``est_pose`` and ``pnp_ransac`` signatures and
return conventions (sentinel ``-1`` returns, never exceptions: recognition.py:79,127,191,215,219).
All arithmetic of the hot path runs in CUDA behind include/pix2pose_b200.h:

* ``generator_train.predict``  -> tcgen05 implicit-GEMM generator (csrc/conv_tc.cuh)
* crop / resize / masks / uint8 XYZ / correspondences -> csrc/pipeline.cu
* ``cv2.solvePnPRansac`` + ``cv2.Rodrigues`` -> csrc/pnp_ransac.cu

``est_pose_batch`` is the throughput entry point (many detections, one device pipeline run);
``est_pose`` is a batch of one.  There is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib, ae_model
from .pnp import solve_pnp_ransac


class _Det(ctypes.Structure):
    _fields_ = [("frame", ctypes.c_int), ("skip", ctypes.c_int), ("bbox", ctypes.c_int * 4), ("box1", ctypes.c_int * 12),
                ("fu", ctypes.c_double), ("fv", ctypes.c_double), ("uc", ctypes.c_double), ("vc", ctypes.c_double),
                ("scale", ctypes.c_double * 3), ("ct", ctypes.c_double * 3),
                ("pool_off", ctypes.c_longlong), ("cap_px", ctypes.c_int), ("seg", ctypes.c_int),
                ("th_o", ctypes.c_double * 8), ("th_i", ctypes.c_double)]


class _Pose(ctypes.Structure):
    _fields_ = [("R", ctypes.c_double * 9), ("t", ctypes.c_double * 3), ("frac_inlier", ctypes.c_double),
                ("status", ctypes.c_int), ("n_inliers", ctypes.c_int), ("best_cand", ctypes.c_int), ("n_cand", ctypes.c_int),
                ("bbox_t", ctypes.c_int * 4), ("best_box", ctypes.c_int * 12), ("n_init", ctypes.c_int),
                ("mask_all_true", ctypes.c_int), ("cand_base", ctypes.c_int), ("pad", ctypes.c_int)]


def _get_boxes(box_size, bbox, v_max, u_max, ct=np.array([-1]), max_w=9999):
    """recognition.py:28-69, verbatim arithmetic (Python floats, ``int()`` truncation)."""
    if ct[0] == -1:
        bbox_ct_v = int((bbox[0] + bbox[2]) / 2)
        bbox_ct_u = int((bbox[1] + bbox[3]) / 2)
    else:
        bbox_ct_v, bbox_ct_u = ct[0], ct[1]
    width, height = bbox[3] - bbox[1], bbox[2] - bbox[0]
    w = min(max_w, max(width * box_size, height * box_size))
    half = int(w / 2)
    lo_v, hi_v, lo_u, hi_u = bbox_ct_v - half, bbox_ct_v + half, bbox_ct_u - half, bbox_ct_u + half
    cv1, cv2_, cu1, cu2 = lo_v, hi_v, lo_u, hi_u
    pad_top = pad_left = cut_bottom = cut_right = 0
    if lo_v < 0:
        pad_top, cv1 = abs(lo_v), 0
    if hi_v > v_max:
        cut_bottom, cv2_ = -abs(hi_v - v_max), v_max
    if lo_u < 0:
        pad_left, cu1 = abs(lo_u), 0
    if hi_u > u_max:
        cut_right, cu2 = -abs(hi_u - u_max), u_max
    return (lo_v, hi_v, lo_u, hi_u, cv1, cv2_, cu1, cu2, pad_top, cut_bottom + (hi_v - lo_v), pad_left,
            cut_right + (hi_u - lo_u))


DET_DTYPE = np.dtype([("frame", "<i4"), ("skip", "<i4"), ("bbox", "<i4", 4), ("box1", "<i4", 12), ("fu", "<f8"), ("fv", "<f8"),
                      ("uc", "<f8"), ("vc", "<f8"), ("scale", "<f8", 3), ("ct", "<f8", 3), ("pool_off", "<i8"), ("cap_px", "<i4"),
                      ("seg", "<i4"), ("th_o", "<f8", 8), ("th_i", "<f8")], align=True)
POSE_DTYPE = np.dtype([("R", "<f8", 9), ("t", "<f8", 3), ("frac_inlier", "<f8"), ("status", "<i4"), ("n_inliers", "<i4"),
                       ("best_cand", "<i4"), ("n_cand", "<i4"), ("bbox_t", "<i4", 4), ("best_box", "<i4", 12), ("n_init", "<i4"),
                       ("mask_all_true", "<i4"), ("cand_base", "<i4"), ("pad", "<i4")], align=True)
assert DET_DTYPE.itemsize == ctypes.sizeof(_Det) and POSE_DTYPE.itemsize == ctypes.sizeof(_Pose)


def _get_boxes_batch(box_size, bboxes, v_max, u_max):
    """Vectorised ``get_boxes`` (recognition.py:28-69) for integer rois without explicit centre: (n,12) int64.
    Same arithmetic as the scalar version (float division, truncation toward zero)."""
    b = np.asarray(bboxes, np.int64).reshape(-1, 4)
    ct_v = np.trunc((b[:, 0] + b[:, 2]) / 2).astype(np.int64)
    ct_u = np.trunc((b[:, 1] + b[:, 3]) / 2).astype(np.int64)
    w = np.minimum(9999, np.maximum((b[:, 3] - b[:, 1]) * box_size, (b[:, 2] - b[:, 0]) * box_size))
    half = np.trunc(w / 2).astype(np.int64)
    lo_v, hi_v, lo_u, hi_u = ct_v - half, ct_v + half, ct_u - half, ct_u + half
    out = np.empty((b.shape[0], 12), np.int64)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = lo_v, hi_v, lo_u, hi_u
    out[:, 4], out[:, 5] = np.maximum(lo_v, 0), np.where(hi_v > v_max, v_max, hi_v)
    out[:, 6], out[:, 7] = np.maximum(lo_u, 0), np.where(hi_u > u_max, u_max, hi_u)
    out[:, 8] = np.where(lo_v < 0, -lo_v, 0)
    out[:, 9] = np.where(hi_v > v_max, -(hi_v - v_max), 0) + (hi_v - lo_v)
    out[:, 10] = np.where(lo_u < 0, -lo_u, 0)
    out[:, 11] = np.where(hi_u > u_max, -(hi_u - u_max), 0) + (hi_u - lo_u)
    return out


class PoseBatchResult:
    """Result of ``est_pose_batch``: struct-of-arrays view plus lazy crop access."""

    def __init__(self, owner, poses, n, run_id):
        self._owner, self._poses, self.n, self._run_id = owner, poses, n, run_id
        a = np.frombuffer(poses, dtype=POSE_DTYPE, count=max(n, 1))[:n]
        self.status = a["status"].copy()
        self.R = a["R"].reshape(n, 3, 3).copy()
        self.t = a["t"].copy()
        self.n_inliers = a["n_inliers"].copy()
        self.frac_inlier = a["frac_inlier"].copy()
        self.bbox_t = a["bbox_t"].astype(np.int64)
        self.n_cand = a["n_cand"].copy()

    def records(self):
        """(n, 16) float64 pose records for the multi-GPU gather: R9, t3, n_inliers, frac, status, index."""
        rec = np.zeros((self.n, 16))
        rec[:, :9] = self.R.reshape(self.n, 9)
        rec[:, 9:12] = self.t
        rec[:, 12], rec[:, 13], rec[:, 14] = self.n_inliers, self.frac_inlier, self.status
        rec[:, 15] = np.arange(self.n)
        return rec

    def crop(self, d):
        """(img_pred uint8 (h,w,3), valid_mask bool (h,w), box) of detection d's winning candidate.  Reads the pipeline's
        candidate pools, so it must be called before the next run of the (shared) pipeline -- a stale result raises."""
        return self._owner._fetch_crop(d, self._poses[d], self._run_id)

    def mask_iou(self, masks):
        """Device-side mask IoU ingredients (tools/5_evaluation_bop_basic.py:307-316): ``masks`` (n,H,W) bool/uint8 detector
        masks of this batch's detections -> (intersection (n,), union (n,)) with the full-frame ``mask_pred`` est_pose
        returns.  Must be called before the next run of the (shared) pipeline."""
        m = np.ascontiguousarray(np.asarray(masks).astype(np.uint8))
        if m.ndim != 3 or m.shape[0] != self.n:
            raise ValueError("masks must be (n,H,W) with n = %d detections, got %s" % (self.n, m.shape))
        out = np.zeros((self.n, 2), np.int64)
        if self.n:
            _lib.check(_lib.lib().p2p_pipeline_mask_iou(self._owner._live_pipe(self._run_id), m.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), self.n,
                                                        m.shape[1], m.shape[2], out.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))))
        return out[:, 0].copy(), out[:, 1].copy()


class pix2pose():
    def __init__(self, weight_fn, camK, res_x, res_y, obj_param, th_ransac=3.0, th_outlier=[0.1, 0.2, 0.3], th_inlier=0.1,
                 box_size=1.5, dist_coeff=None, backbone="paper", precision="fp16x3", capacity=64, max_dets=64, **kwargs):
        # recognition.py:11-26
        self.camK = camK
        self.res_x = res_x
        self.res_y = res_y
        self.th_ransac = th_ransac      # stored, never used (reprojectionError is hard-coded to 5, recognition.py:217)
        self.th_o = th_outlier
        self.th_i = th_inlier
        self.obj_scale = obj_param[:3]  # x,y,z
        self.obj_ct = obj_param[3:]     # x,y,z
        self.box_size = box_size
        self.dist_coeff = dist_coeff    # stored, never used (distCoeffs=None, recognition.py:216)
        if backbone == 'paper':
            self.generator_train = ae_model.aemodel_unet_prob(p=1.0, precision=precision, capacity=capacity)
        elif backbone == 'resnet50':
            self.generator_train = ae_model.aemodel_unet_resnet50(p=1.0, precision=precision, capacity=capacity)
        else:
            raise ValueError("backbone must be 'paper' or 'resnet50'")
        self.generator_train.load_weights(weight_fn)
        self._max_dets = int(max_dets)
        self._last_run = None

    # ------------------------------------------------------------------------------------------------
    def get_boxes(self, bbox, v_max, u_max, ct=np.array([-1]), max_w=9999):
        return _get_boxes(self.box_size, bbox, v_max, u_max, ct, max_w)

    # The device pipeline (stage buffers, candidate pools, PnP scratch) does not depend on the object: all
    # pix2pose instances that share an engine and a threshold count share ONE pipeline, like the reference's
    # per-object models share one TF session (tools/5_evaluation_bop_basic.py:112-114).
    _shared_pipes = {}     # (engine id, n thresholds) -> {"h": handle, "cap": max_dets, "eng": engine, "run": run counter, "old": [...]}

    def _entry(self):
        eng = self.generator_train.engine
        return pix2pose._shared_pipes.get((id(eng), len(self.th_o)))

    def _pipeline(self, n):
        """The shared device pipeline, grown if `n` detections do not fit.  A pipeline that is outgrown is NOT destroyed while
        the process lives (results and other objects may still hold buffers in it: it is parked in the entry's "old" list and
        its results are invalidated through the run counter); every access goes through the shared entry, never through a
        cached handle."""
        eng = self.generator_train.engine
        key = (id(eng), len(self.th_o))
        ent = pix2pose._shared_pipes.get(key)
        want = max(self._max_dets, n)
        if ent is None or ent["cap"] < n:
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().p2p_pipeline_create(eng.handle, want, len(self.th_o), ctypes.byref(h)))
            old = [] if ent is None else ent["old"] + [ent["h"]]
            ent = {"h": h, "cap": want, "eng": eng, "run": 0 if ent is None else ent["run"] + 1, "old": old}
            pix2pose._shared_pipes[key] = ent
        _lib.check(_lib.lib().p2p_pipeline_set_box_size(ent["h"], float(self.box_size)))   # pipelines are shared by objects
        return ent["h"]

    def _live_pipe(self, run_id=None):
        """Handle of the shared pipeline for reading back the buffers of run `run_id` (default: this object's last run);
        raises when another run (of any object sharing the pipeline) has overwritten them since."""
        ent = self._entry()
        if ent is None:
            raise RuntimeError("no pipeline run yet")
        want = self._last_run if run_id is None else run_id
        if want is None or ent["run"] != want:
            raise RuntimeError("stale result: the shared device pipeline has run again (run %s, result of run %s); fetch crops / "
                               "masks before the next est_pose call" % (ent["run"], want))
        return ent["h"]

    def last_forward_ms(self):
        """Device milliseconds the generator forwards of the shared pipeline's last run took (bench.py's roofline numerator)."""
        ms = ctypes.c_double()
        _lib.check(_lib.lib().p2p_pipeline_forward_ms(self._entry()["h"], ctypes.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        n = self.generator_train.engine.launch_count
        ent = self._entry()
        if ent is not None:
            n += int(_lib.lib().p2p_pipeline_launch_count(ent["h"]))
        return n

    def _make_dets(self, shape_hw, bboxes, frame_ids, camKs):
        H, W = shape_hw
        n = len(bboxes)
        dets = (_Det * max(n, 1))()
        if n == 0:
            return dets
        a = np.frombuffer(dets, dtype=DET_DTYPE, count=n)
        bb = np.array([[int(v) for v in b] for b in bboxes], np.int64).reshape(n, 4)
        box1 = _get_boxes_batch(self.box_size, bb, H, W)
        a["frame"] = np.asarray(frame_ids, np.int64)
        a["bbox"] = bb
        a["box1"] = box1
        a["skip"] = ((box1[:, 1] - box1[:, 0] < 5) | (box1[:, 3] - box1[:, 2] < 5) | (box1[:, 5] - box1[:, 4] < 5) |
                     (box1[:, 7] - box1[:, 6] < 5))                           # recognition.py:78
        if camKs is None:
            K = np.asarray(self.camK, np.float64).reshape(3, 3)
            a["fu"], a["fv"], a["uc"], a["vc"] = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        else:
            Ks = np.asarray(camKs, np.float64).reshape(n, 3, 3)
            a["fu"], a["fv"], a["uc"], a["vc"] = Ks[:, 0, 0], Ks[:, 1, 1], Ks[:, 0, 2], Ks[:, 1, 2]
        a["scale"] = np.asarray(self.obj_scale, np.float64)
        a["ct"] = np.asarray(self.obj_ct, np.float64)
        th = np.zeros(8)
        th[:len(self.th_o)] = np.asarray(self.th_o, np.float64).reshape(-1)   # cfg_tless_paper.json nests them: [[0.1]]
        a["th_o"] = th
        a["th_i"] = float(self.th_i)
        return dets

    def est_pose_batch(self, frames, bboxes, frame_ids=None, camKs=None, frames_dev=None):
        """Poses for many detections of this object in one device pipeline run.

        frames: (F,H,W,3) uint8 (or a single (H,W,3) image); bboxes: (n,4) rois [v0,u0,v1,u1];
        frame_ids: (n,) index of each roi's frame (default 0); camKs: optional per-detection intrinsics.
        Returns a ``PoseBatchResult``; ``status``: 1 pose found, 0 the reference would return its -1
        sentinels, -2 the crop was smaller than 5 px (recognition.py:78)."""
        n = len(bboxes)
        f32 = False
        if frames_dev is not None:
            dev, F, H, W, pipe0 = frames_dev
        else:
            frames = np.asarray(frames)
            if frames.ndim == 3:
                frames = frames[None]
            if frames.dtype != np.uint8:
                # the ICP driver passes a float32 image (0..255, invalid-depth pixels scaled by 0.1: non-integers;
                # 5_evaluation_bop_icp3d.py:369-370); numpy promotes `rgb - [128,128,128]` to float64 on it, so the crops are
                # cut from the float values (float64 input would be rounded to float32 here: the reference never passes it)
                frames, f32 = frames.astype(np.float32), True
            frames = np.ascontiguousarray(frames)
            F, H, W = frames.shape[0], frames.shape[1], frames.shape[2]
        if frame_ids is None:
            frame_ids = np.zeros(n, np.int64)
        dets = self._make_dets((H, W), bboxes, frame_ids, camKs)
        poses = (_Pose * max(n, 1))()
        th = np.ascontiguousarray(np.asarray(self.th_o, np.float64).reshape(-1))
        pipe = self._pipeline(n)
        ent = self._entry()
        thp = th.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        ent["run"] += 1
        self._last_run = ent["run"]
        if frames_dev is not None:
            if pipe0.value != pipe.value:
                raise RuntimeError("frames_dev belongs to a pipeline that was rebuilt; upload the frames again")
            _lib.check(_lib.lib().p2p_pipeline_run_device(pipe, self.generator_train._model, dev, F, H, W, dets, n, thp,
                                                          float(self.th_i), 5.0, 100, 0.99, poses))
        elif f32:
            _lib.check(_lib.lib().p2p_pipeline_run_f32(pipe, self.generator_train._model, _lib.fptr(frames), F, H, W, dets, n,
                                                       thp, float(self.th_i), 5.0, 100, 0.99, poses))
        else:
            _lib.check(_lib.lib().p2p_pipeline_run(
                pipe, self.generator_train._model, frames.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), F, H, W, dets, n,
                thp, float(self.th_i), 5.0, 100, 0.99, poses))
        return PoseBatchResult(self, poses, n, self._last_run)

    def upload_frames(self, frames, n_dets=1):
        """Copies (F,H,W,3) uint8 frames into the pipeline's device buffer; returns an opaque handle for
        ``est_pose_batch(..., frames_dev=handle)`` (keeps detector output on the device)."""
        frames = np.ascontiguousarray(np.asarray(frames, np.uint8))
        if frames.ndim == 3:
            frames = frames[None]
        pipe = self._pipeline(n_dets)
        dev = ctypes.c_void_p()
        _lib.check(_lib.lib().p2p_pipeline_upload_frames(pipe, frames.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                                         frames.shape[0], frames.shape[1], frames.shape[2], ctypes.byref(dev)))
        return (dev, frames.shape[0], frames.shape[1], frames.shape[2], pipe)

    def _fetch_crop(self, d, pose, run_id=None):
        bx = list(pose.best_box)
        h, w = max(bx[5] - bx[4], 0), max(bx[7] - bx[6], 0)
        xyz = np.zeros((h, w, 3), np.uint8)
        mask = np.zeros((h, w), np.uint8)
        if h * w > 0:
            _lib.check(_lib.lib().p2p_pipeline_fetch_crop(self._live_pipe(run_id), d, ctypes.byref(pose),
                                                          xyz.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                                          mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
        if pose.mask_all_true:
            mask[:] = 1
        return xyz, mask.astype(bool), bx

    def debug_fetch(self, what, index):
        """Parity hook: float buffers of the last run (1 dec1, 2 dec2, 3 x1, 4 x2, 5 prob1, 6 prob2)."""
        out = np.zeros((128, 128, 3) if what <= 4 else (128, 128), np.float32)
        _lib.check(_lib.lib().p2p_pipeline_fetch_buffer(self._live_pipe(), what, index, _lib.fptr(out)))
        return out

    def debug_override(self, stage, decode, prob, n_dets=1):
        """Parity hook: the next run uses these arrays instead of the network outputs of `stage`."""
        decode, prob = _lib.as_f32(decode), _lib.as_f32(prob)
        _lib.check(_lib.lib().p2p_pipeline_debug_override(self._pipeline(n_dets), stage, _lib.fptr(decode), _lib.fptr(prob), decode.shape[0]))

    def debug_select(self, Rt, n_inliers, status):
        """Parity hook: repeats the candidate selection of this object's last single-detection run with the given
        per-candidate PnP results (``Rt`` (k,12): R row-major then t) and returns the resulting pose record."""
        Rt = np.ascontiguousarray(np.asarray(Rt, np.float64).reshape(-1, 12))
        ni = np.ascontiguousarray(np.asarray(n_inliers, np.int32))
        st = np.ascontiguousarray(np.asarray(status, np.int32))
        pose = (_Pose * 1)()
        _lib.check(_lib.lib().p2p_pipeline_debug_select(self._live_pipe(), Rt.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                        ni.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                                        st.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(ni), 1, pose))
        return pose[0]

    def _fetch_pred(self, stage, index, zero_gray):
        dec = np.zeros((128, 128, 3), np.float32)
        _lib.check(_lib.lib().p2p_pipeline_fetch_decode(self._live_pipe(), stage, index, _lib.fptr(dec)))
        if zero_gray:
            dec[np.linalg.norm(dec, axis=2) < 0.3] = 0           # recognition.py:137-139
        return np.clip((dec + 1) / 2, 0, 1)                       # :85-87 / :141-143

    def est_pose(self, rgb, bbox, gt_trans=np.eye((4)), z_iter=False):
        """recognition.py:70-193 -> (img_pred, mask_pred, rot_pred, tra_pred, frac_inlier, bbox_t)."""
        res = self.est_pose_batch(rgb, [bbox])
        pose = res._poses[0]
        bbox_t = np.array(list(pose.bbox_t), int)
        if pose.status == -2:
            return np.zeros((1)), -1, -1, -1, -1, bbox_t                      # :79
        if pose.status != 1:
            if pose.n_cand <= 0:
                return self._fetch_pred(1, 0, False), -1, -1, -1, -1, bbox_t   # :125-127
            return self._fetch_pred(2, pose.cand_base + pose.n_cand - 1, True), -1, -1, -1, -1, bbox_t   # :189-191
        xyz, mask, bx = self._fetch_crop(0, pose)
        valid_mask_full = np.zeros((rgb.shape[0], rgb.shape[1]), bool)
        valid_mask_full[bx[4]:bx[5], bx[6]:bx[7]] = mask
        return (xyz, valid_mask_full, np.array(list(pose.R)).reshape(3, 3), np.array(list(pose.t)), pose.frac_inlier,
                bbox_t)                                                        # :193

    def pnp_ransac(self, rgb_aug_test, img_prob_ori, non_zero, v1, v2, u1, u2):
        """recognition.py:195-224 with the PnP itself on the GPU (host-side correspondence build:
        this standalone entry point takes host arrays, as the reference's does)."""
        xyz = rgb_aug_test[v1:v2, u1:u2].astype(np.float64) / 255 * 2 - 1
        for a in range(3):
            xyz[:, :, a] = xyz[:, :, a] * self.obj_scale[a] + self.obj_ct[a]
        valid_mask = np.logical_and(non_zero, img_prob_ori < self.th_i)
        vs, us = np.where(valid_mask == 1)
        n_pts = len(vs)
        if n_pts < 6:
            return np.eye(3), np.array([0, 0, 0]), valid_mask, -1
        img_pts = np.stack((us + u1, vs + v1), axis=1).astype(np.float64)
        ret, rvec, tvec, inliers, R, _ = solve_pnp_ransac(xyz[vs, us], img_pts, self.camK, 5.0, 100, 0.99)
        if inliers is None:
            return np.eye(3), np.array([0, 0, 0]), -1, -1
        return R, tvec[:, 0], valid_mask, len(inliers)
