"""GPU EPnP-RANSAC behind the signature of ``cv2.solvePnPRansac`` as recognition.py:216 calls it."""
import ctypes

import numpy as np

from . import _lib


def solve_pnp_ransac(obj_pts, img_pts, camK, reprojectionError=5.0, iterationsCount=100, confidence=0.99):
    """Returns ``(ret, rvec(3,1), tvec(3,1), inliers(m,1) int32 or None, R(3,3), iters_run)``.

    Replaces ``cv2.solvePnPRansac(obj, img, camK, None, flags=cv2.SOLVEPNP_EPNP, reprojectionError=5,
    iterationsCount=100)`` + ``cv2.Rodrigues`` (pix2pose_model/recognition.py:216-223)."""
    obj = np.ascontiguousarray(np.asarray(obj_pts, np.float64).reshape(-1, 3))
    img = np.ascontiguousarray(np.asarray(img_pts, np.float64).reshape(-1, 2))
    n = obj.shape[0]
    if img.shape[0] != n:
        raise ValueError("obj/img point counts differ: %d vs %d" % (n, img.shape[0]))
    K = np.ascontiguousarray(np.asarray(camK, np.float64).reshape(3, 3))
    rvec, tvec, R = np.zeros(3), np.zeros(3), np.zeros(9)
    mask = np.zeros(max(n, 1), np.uint8)
    ninl, iters_run = ctypes.c_int(), ctypes.c_int()
    dp = ctypes.POINTER(ctypes.c_double)
    _lib.check(_lib.lib().p2p_pnp_ransac(obj.ctypes.data_as(dp), img.ctypes.data_as(dp), n, K.ctypes.data_as(dp),
                                         float(reprojectionError), int(iterationsCount), float(confidence),
                                         rvec.ctypes.data_as(dp), tvec.ctypes.data_as(dp), R.ctypes.data_as(dp),
                                         ctypes.byref(ninl), mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                         ctypes.byref(iters_run)))
    if ninl.value < 0:
        return False, rvec.reshape(3, 1), tvec.reshape(3, 1), None, np.eye(3), iters_run.value
    inliers = np.nonzero(mask[:n])[0].astype(np.int32).reshape(-1, 1)
    return True, rvec.reshape(3, 1), tvec.reshape(3, 1), inliers, R.reshape(3, 3), iters_run.value


_RESULT_DTYPE = np.dtype([("rvec", "<f8", 3), ("tvec", "<f8", 3), ("R", "<f8", (3, 3)), ("n_inliers", "<i4"), ("best_iter", "<i4"),
                          ("iters_run", "<i4"), ("status", "<i4"), ("n_mask", "<i4"), ("pad", "<i4")])   # p2p_pnp_result_t


def solve_pnp_ransac_batch(obj_list, img_list, camK, reprojectionError=5.0, iterationsCount=100, confidence=0.99,
                           return_time=False):
    """``cv2.solvePnPRansac`` for a batch of independent problems in one device run (``p2p_pnp_ransac_batch``).

    ``obj_list[i]`` (n_i,3) / ``img_list[i]`` (n_i,2); ``camK`` one 3x3 or one per problem.  Returns a list of the tuples
    ``solve_pnp_ransac`` returns (and the device time in ms when ``return_time``)."""
    n_prob = len(obj_list)
    if len(img_list) != n_prob:
        raise ValueError("obj_list / img_list lengths differ")
    objs = [np.asarray(o, np.float64).reshape(-1, 3) for o in obj_list]
    imgs = [np.asarray(i, np.float64).reshape(-1, 2) for i in img_list]
    counts = np.array([o.shape[0] for o in objs], np.int32)
    if any(i.shape[0] != c for i, c in zip(imgs, counts)):
        raise ValueError("obj/img point counts differ")
    obj = np.ascontiguousarray(np.concatenate(objs) if n_prob else np.zeros((0, 3)))
    img = np.ascontiguousarray(np.concatenate(imgs) if n_prob else np.zeros((0, 2)))
    K = np.ascontiguousarray(np.asarray(camK, np.float64))
    per = 1 if K.size == 9 * n_prob and K.size != 9 else 0
    if K.size != (9 * n_prob if per else 9):
        raise ValueError("camK must be one 3x3 matrix or one per problem")
    res = np.zeros(n_prob, _RESULT_DTYPE)
    mask = np.zeros(max(int(counts.sum()), 1), np.uint8)
    ms = ctypes.c_float()
    dp = ctypes.POINTER(ctypes.c_double)
    _lib.check(_lib.lib().p2p_pnp_ransac_batch(obj.ctypes.data_as(dp), img.ctypes.data_as(dp),
                                               counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n_prob, K.ctypes.data_as(dp), per,
                                               float(reprojectionError), int(iterationsCount), float(confidence),
                                               res.ctypes.data_as(ctypes.c_void_p), mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                               ctypes.byref(ms)))
    out, at = [], 0
    for i in range(n_prob):
        r, n = res[i], int(counts[i])
        if r["n_inliers"] < 0:
            out.append((False, r["rvec"].reshape(3, 1).copy(), r["tvec"].reshape(3, 1).copy(), None, np.eye(3), int(r["iters_run"])))
        else:
            inl = np.nonzero(mask[at:at + n])[0].astype(np.int32).reshape(-1, 1)
            out.append((True, r["rvec"].reshape(3, 1).copy(), r["tvec"].reshape(3, 1).copy(), inl, r["R"].copy(), int(r["iters_run"])))
        at += n
    return (out, ms.value) if return_time else out
