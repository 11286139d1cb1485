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
