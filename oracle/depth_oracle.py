"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the depth helpers of the reference's ICP path,
``pix2pose_util/common_util.py:13-90`` (``getXYZ``, ``get_normal``), with numpy / scipy / cv2 exactly as the reference
calls them (only ``np.float`` -> ``float``, removed in numpy >= 1.24).  Checker for ``pix2pose_b200/depth.py``;
never imported by the product.  PINNED: unlike the network half, this part of the reference imports here (numpy, cv2,
scipy only), so tests/golden/depth_golden.npz holds outputs of the reference's own functions
(tests/golden/make_depth_golden.py) and tests/test_oracle_depth.py checks this restatement against them exactly."""
import cv2
import numpy as np
from scipy import ndimage


def getXYZ(depth, fx, fy, cx, cy, bbox=np.array([0])):
    """common_util.py:13-30."""
    uv_table = np.zeros((depth.shape[0], depth.shape[1], 2), dtype=np.int16)
    column = np.arange(0, depth.shape[0])
    uv_table[:, :, 1] = np.arange(0, depth.shape[1]) - cx
    uv_table[:, :, 0] = column[:, np.newaxis] - cy
    if bbox.shape[0] == 1:
        xyz = np.zeros((depth.shape[0], depth.shape[1], 3))
        xyz[:, :, 0] = uv_table[:, :, 1] * depth * 1 / fx
        xyz[:, :, 1] = uv_table[:, :, 0] * depth * 1 / fy
        xyz[:, :, 2] = depth
    else:
        xyz = np.zeros((bbox[2] - bbox[0], bbox[3] - bbox[1], 3))
        xyz[:, :, 0] = uv_table[bbox[0]:bbox[2], bbox[1]:bbox[3], 1] * depth[bbox[0]:bbox[2], bbox[1]:bbox[3]] * 1 / fx
        xyz[:, :, 1] = uv_table[bbox[0]:bbox[2], bbox[1]:bbox[3], 0] * depth[bbox[0]:bbox[2], bbox[1]:bbox[3]] * 1 / fy
        xyz[:, :, 2] = depth[bbox[0]:bbox[2], bbox[1]:bbox[3]]
    return xyz


def refine_depth(depth_refine):
    """common_util.py:42-48: nan_to_num, Navier-Stokes inpainting of the zero pixels, Gaussian smoothing (sigma 2)."""
    depth_refine = np.nan_to_num(depth_refine)
    mask = np.zeros_like(depth_refine).astype(np.uint8)
    mask[depth_refine == 0] = 1
    depth_refine = depth_refine.astype(np.float32)
    depth_refine = cv2.inpaint(depth_refine, mask, 2, cv2.INPAINT_NS)
    depth_refine = depth_refine.astype(float)
    return ndimage.gaussian_filter(depth_refine, 2)


def get_normal(depth_refine, fx=-1, fy=-1, cx=-1, cy=-1, bbox=np.array([0]), refine=True):
    """common_util.py:32-90."""
    res_y, res_x = depth_refine.shape[0], depth_refine.shape[1]
    constant_x, constant_y = 1 / fx, 1 / fy
    if refine:
        depth_refine = refine_depth(depth_refine)
    uv_table = np.zeros((res_y, res_x, 2), dtype=np.int16)
    column = np.arange(0, res_y)
    uv_table[:, :, 1] = np.arange(0, res_x) - cx
    uv_table[:, :, 0] = column[:, np.newaxis] - cy
    if bbox.shape[0] == 4:
        uv_table = uv_table[bbox[0]:bbox[2], bbox[1]:bbox[3]]
        shape = (bbox[2] - bbox[0], bbox[3] - bbox[1], 3)
        depth_refine = depth_refine[bbox[0]:bbox[2], bbox[1]:bbox[3]]
    else:
        shape = (res_y, res_x, 3)
    v_x, v_y = np.zeros(shape), np.zeros(shape)
    uv_table_sign = np.copy(uv_table)
    dig = np.gradient(depth_refine, 2, edge_order=2)
    v_y[:, :, 0] = uv_table_sign[:, :, 1] * constant_x * dig[0]
    v_y[:, :, 1] = depth_refine * constant_y + (uv_table_sign[:, :, 0] * constant_y) * dig[0]
    v_y[:, :, 2] = dig[0]
    v_x[:, :, 0] = depth_refine * constant_x + uv_table_sign[:, :, 1] * constant_x * dig[1]
    v_x[:, :, 1] = uv_table_sign[:, :, 0] * constant_y * dig[1]
    v_x[:, :, 2] = dig[1]
    cross = np.cross(v_x.reshape(-1, 3), v_y.reshape(-1, 3))
    norm = np.expand_dims(np.linalg.norm(cross, axis=1), axis=1)
    norm[norm == 0] = 1
    cross = cross / norm
    cross = cross.reshape(shape)
    return np.nan_to_num(cross)
