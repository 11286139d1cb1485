"""Depth -> ICP inputs on the GPU: drop-in for ``pix2pose_util/common_util.py`` ``getXYZ`` / ``get_normal``
(callers tools/5_evaluation_bop_icp3d.py:78-79, ros_kinetic/ros_pix2pose.py:180-181, 296-297).
Same signatures and return shapes.  The per-pixel arithmetic (int16 pixel tables, point back-projection, Gaussian
smoothing, second-order gradient, normalised cross product) runs in CUDA kernels in double (csrc/depth.cu); the one step
kept on the host is OpenCV's Navier-Stokes inpainting of the depth holes (``cv2.inpaint``, common_util.py:43-47), a
sequential fast-marching algorithm that is not part of this path's kernels.  No CPU fallback for the rest."""
import ctypes

import numpy as np

from . import _lib

_DP = ctypes.POINTER(ctypes.c_double)


def _bbox_arg(bbox, n_expected):
    bbox = np.asarray(bbox)
    if bbox.shape[0] == n_expected:
        b = np.ascontiguousarray(bbox.astype(np.int32))
        return b, b.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), (int(b[2] - b[0]), int(b[3] - b[1]))
    return None, None, None


def getXYZ(depth, fx, fy, cx, cy, bbox=np.array([0])):
    """common_util.py:13-30: (h,w,3) float64 camera-frame points, in the unit of ``depth``."""
    d = np.ascontiguousarray(np.asarray(depth, np.float64))
    H, W = d.shape
    keep, bptr, hw = _bbox_arg(bbox, 4) if np.asarray(bbox).shape[0] != 1 else (None, None, None)
    h, w = hw if hw else (H, W)
    out = np.zeros((h, w, 3))
    _lib.check(_lib.lib().p2p_depth_xyz(d.ctypes.data_as(_DP), H, W, float(fx), float(fy), float(cx), float(cy), bptr, out.ctypes.data_as(_DP)))
    return out


def get_normal(depth_refine, fx=-1, fy=-1, cx=-1, cy=-1, bbox=np.array([0]), refine=True):
    """common_util.py:32-90: (h,w,3) float64 unit normals of the (optionally hole-filled and smoothed) depth map."""
    d = np.asarray(depth_refine)
    sigma = 0.0
    if refine:
        import cv2
        d = np.nan_to_num(d)                                   # :43
        mask = np.zeros_like(d).astype(np.uint8)
        mask[d == 0] = 1
        d = cv2.inpaint(d.astype(np.float32), mask, 2, cv2.INPAINT_NS)   # :44-47 (host: see module docstring)
        sigma = 2.0                                            # :48 runs on the device
    d = np.ascontiguousarray(d.astype(np.float64))
    H, W = d.shape
    keep, bptr, hw = _bbox_arg(bbox, 4)
    h, w = hw if hw else (H, W)
    out = np.zeros((h, w, 3))
    _lib.check(_lib.lib().p2p_depth_normals(d.ctypes.data_as(_DP), H, W, float(fx), float(fy), float(cx), float(cy), bptr, sigma,
                                            out.ctypes.data_as(_DP)))
    return out
