"""CPU restatement of ``pix2pose_model/recognition.py`` (class ``pix2pose``), the reference's
per-detection host logic.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Every stage cites the reference lines it follows.  Deliberate substitutions, because the
reference cannot be imported here:

* ``generator_train``   -> any object with ``predict`` (tests pass ``oracle.net_oracle.NetOracle``)
* ``skimage resize``    -> ``oracle.resize_oracle.resize`` (parity unpinned, see there)
* ``np.int``            -> ``int`` (numpy >= 1.24 removed the alias)
* ``cv2.solvePnPRansac``/``cv2.Rodrigues`` are the real OpenCV calls (4.13.0 here, 3.4.2.17 pinned).

Reference quirks kept on purpose (they are observable behaviour):
Q1 the refined-box list grows before the size check that may skip the candidate (:111 vs :117-119),
   so after such a skip boxes and network outputs are paired off by one;
Q2 the returned ``bbox_t`` is the clipped box of the LAST candidate looped over, not of the winner
   (loop variables reused at :133, returned at :191/:193);
Q3 the stage-2 centre adds a 128-space offset to a frame-space centre without rescaling (:108-109).
"""
import cv2
import numpy as np

from .resize_oracle import resize


def get_boxes(box_size, bbox, v_max, u_max, ct=None, max_w=9999):
    """recognition.py:28-69 -> (v1_ori,v2_ori,u1_ori,u2_ori, v1,v2,u1,u2, vv1,vv2,uu1,uu2)."""
    if ct is None or ct[0] == -1:
        ct_v = int((bbox[0] + bbox[2]) / 2)
        ct_u = int((bbox[1] + bbox[3]) / 2)
    else:
        ct_v, ct_u = ct[0], ct[1]
    width = bbox[3] - bbox[1]
    height = bbox[2] - bbox[0]
    w = min(max_w, max(width * box_size, height * box_size))
    half = int(w / 2)
    v1o, v2o, u1o, u2o = ct_v - half, ct_v + half, ct_u - half, ct_u + half
    v1, v2, u1, u2 = v1o, v2o, u1o, u2o
    sv_min = su_min = sv_max = su_max = 0
    if v1o < 0:
        sv_min, v1 = abs(v1o), 0
    if v2o > v_max:
        sv_max, v2 = -abs(v2o - v_max), v_max
    if u1o < 0:
        su_min, u1 = abs(u1o), 0
    if u2o > u_max:
        su_max, u2 = -abs(u2o - u_max), u_max
    return (v1o, v2o, u1o, u2o, v1, v2, u1, u2, sv_min, sv_max + (v2o - v1o), su_min, su_max + (u2o - u1o))


class Pix2PoseOracle:
    def __init__(self, generator, camK, res_x, res_y, obj_param, th_ransac=3.0, th_outlier=(0.1, 0.2, 0.3),
                 th_inlier=0.1, box_size=1.5, dist_coeff=None):
        # recognition.py:10-26 (th_ransac and dist_coeff are stored and never used)
        self.generator_train = generator
        self.camK = np.asarray(camK, np.float64)
        self.res_x, self.res_y = res_x, res_y
        self.th_ransac = th_ransac
        self.th_o = list(th_outlier)
        self.th_i = th_inlier
        self.obj_scale = np.asarray(obj_param[:3], np.float64)
        self.obj_ct = np.asarray(obj_param[3:], np.float64)
        self.box_size = box_size
        self.dist_coeff = dist_coeff
        self.trace = None  # optional dict filled with intermediates for stage-wise parity tests

    def get_boxes(self, bbox, v_max, u_max, ct=None, max_w=9999):
        return get_boxes(self.box_size, bbox, v_max, u_max, ct, max_w)

    # ---------------------------------------------------------------------------------------------
    def _padded_crop(self, rgb, box, bg_full=None):
        """recognition.py:75-82 / :113-121: normalised crop pasted into a zero square, or None when
        one of the size guards (:78, :116-118) trips."""
        v1o, v2o, u1o, u2o, v1, v2, u1, u2, vv1, vv2, uu1, uu2 = box
        side_v, side_u = v2o - v1o, u2o - u1o
        patch = (rgb[v1:v2, u1:u2].astype(np.float64) - 128.0) / 128.0
        if bg_full is not None:
            patch[bg_full[v1:v2, u1:u2]] = 0
        if side_v < 5 or side_u < 5 or patch.shape[0] < 5 or patch.shape[1] < 5:
            return None
        base = np.zeros((max(side_v, 0), max(side_u, 0), 3))
        if bg_full is not None and (base[vv1:vv2, uu1:uu2].shape[0] == 0 or base[vv1:vv2, uu1:uu2].shape[1] == 0):
            return None
        base[vv1:vv2, uu1:uu2] = patch
        return resize(base, (128, 128), order=1, mode="reflect")

    def est_pose(self, rgb, bbox):
        """recognition.py:70-193."""
        H, W = rgb.shape[0], rgb.shape[1]
        box1 = self.get_boxes(bbox, H, W)                                         # :71
        v1o, v2o, u1o, u2o, v1, v2, u1, u2, vv1, vv2, uu1, uu2 = box1
        cx_o, cy_o = (bbox[3] + bbox[1]) / 2, (bbox[2] + bbox[0]) / 2              # :72-73
        w_stage_1 = v2o - v1o                                                     # :74
        x1 = self._padded_crop(rgb, box1)                                         # :75-82
        if x1 is None:
            return np.zeros((1)), -1, -1, -1, -1, np.array([v1, v2, u1, u2], int)  # :79
        decode, prob = self.generator_train.predict(np.expand_dims(x1, 0))        # :84
        img_pred = np.clip((decode[0] + 1) / 2, 0, 1)                             # :85-87
        non_gray = np.linalg.norm(decode[0], axis=2) > 0.3                        # :89
        n_init_mask = np.sum(non_gray)                                            # :90
        if self.trace is not None:
            self.trace.update(x1=x1, decode1=decode.copy(), prob1=prob.copy(), box1=box1)
        inputs, boxes = [], []
        for th_o in self.th_o:                                                    # :93
            non_gray_prob = np.logical_and(non_gray, prob[0, :, :, 0] < th_o)     # :94-95
            if np.sum(non_gray_prob) < 10:                                        # :96
                continue
            vs, us = np.where(non_gray)                                           # :98
            if len(vs) == 0:
                continue
            sv, su = (v2o - v1o) / 128, (u2o - u1o) / 128
            bb = np.array([vs.min() * sv, us.min() * su, vs.max() * sv, us.max() * su])       # :101-102
            m_ori = resize(non_gray_prob, (v2o - v1o, u2o - u1o), order=1, mode="constant", cval=0) > 0.9  # :103
            m_ori = m_ori[vv1:vv2, uu1:uu2]                                       # :104
            bg_full = np.ones((H, W), bool)                                       # :105
            bg_full[v1:v2, u1:u2] = np.invert(m_ori)                              # :106
            cx_m = int((np.mean(us) - (127 / 2)) + cx_o)                          # :108  (Q3)
            cy_m = int((np.mean(vs) - (127 / 2)) + cy_o)                          # :109
            box2 = self.get_boxes(bb, H, W, ct=np.array([cy_m, cx_m]), max_w=w_stage_1)  # :110
            boxes.append(box2)                                                    # :111  (Q1)
            x2 = self._padded_crop(rgb, box2, bg_full)                            # :113-121
            if x2 is None:
                continue
            inputs.append(x2)
        if len(inputs) <= 0:                                                      # :125-127
            return img_pred, -1, -1, -1, -1, np.array([v1, v2, u1, u2], int)
        decode, prob = self.generator_train.predict(np.array(inputs))             # :129
        decode = np.array(decode, copy=True)
        if self.trace is not None:
            self.trace.update(x2=np.array(inputs), decode2=decode.copy(), prob2=prob.copy(), boxes2=list(boxes),
                              cands=[])
        max_inlier, min_dist = -1, 9999999
        rot_pred = tra_pred = valid_mask_full = img_pred_f = None
        for cid in range(len(inputs)):                                            # :132
            v1o, v2o, u1o, u2o, v1, v2, u1, u2, vv1, vv2, uu1, uu2 = boxes[cid]   # :133 (Q1, Q2)
            sh = (v2o - v1o, u2o - u1o)
            prob_ori = resize(prob[cid, :, :, 0], sh, order=1, mode="constant", cval=1)[vv1:vv2, uu1:uu2]  # :134-135
            gray = np.linalg.norm(decode[cid], axis=2) < 0.3                      # :137
            non_gray2 = np.invert(gray)
            decode[cid, gray, :] = 0                                              # :139
            img_pred = np.clip((decode[cid] + 1) / 2, 0, 1)                       # :141-143
            img_pred_ori = resize(img_pred, sh, order=1, mode="constant", cval=0.5) * 255      # :144
            ng_ori = (resize(non_gray2.astype(float), sh, order=1, mode="constant", cval=0) > 0.9)[vv1:vv2, uu1:uu2]  # :146-147
            n_non_gray = np.sum(ng_ori)
            if n_non_gray < 10:                                                   # :149
                continue
            img_pred_ori = img_pred_ori[vv1:vv2, uu1:uu2]                         # :151
            frame = np.zeros((H, W, 3), np.uint8)                                 # :152
            frame[v1:v2, u1:u2] = [128, 128, 128]
            frame[v1:v2, u1:u2] = img_pred_ori                                    # :154  float -> uint8 truncation (F8)
            R_c, t_c, valid_mask, n_inl = self.pnp_ransac(frame, prob_ori, ng_ori, v1, v2, u1, u2)  # :156
            full = np.zeros((H, W), bool)                                         # :159-161
            full[v1:v2, u1:u2] = ng_ori
            vv, uu = np.where(full)
            ct_pt = np.array([np.mean(vv), np.mean(uu)])                          # :162
            if t_c[2] == 0:                                                       # :163
                dist = 99999
            else:
                pu = self.camK[0, 0] * t_c[0] / t_c[2] + self.camK[0, 2]
                pv = self.camK[1, 1] * t_c[1] / t_c[2] + self.camK[1, 2]
                dist = ((pv - ct_pt[0]) ** 2 + (pu - ct_pt[1]) ** 2) / (n_inl + 1e-6)          # :168
            if self.trace is not None:
                self.trace["cands"].append(dict(cid=cid, n_inliers=n_inl, R=R_c, t=t_c, dist=dist, n_non_gray=int(n_non_gray),
                                                xyz_u8=frame[v1:v2, u1:u2].copy(), valid_mask=valid_mask, box=boxes[cid]))
            if dist < min_dist:                                                   # :170-178
                rot_pred, tra_pred, max_inlier, min_dist = R_c, t_c, n_inl, dist
                valid_mask_full = np.zeros((H, W), bool)
                valid_mask_full[v1:v2, u1:u2] = valid_mask
                img_pred_f = img_pred_ori
        if max_inlier == -1:                                                      # :189-191
            return img_pred, -1, -1, -1, -1, np.array([v1, v2, u1, u2], int)
        return (img_pred_f.astype(np.uint8), valid_mask_full, rot_pred, tra_pred, max_inlier / n_init_mask,
                np.array([v1, v2, u1, u2], int))                                  # :193

    def build_correspondences(self, frame_u8, prob_ori, non_zero, v1, v2, u1, u2):
        """recognition.py:196-213 -> (obj_pts (n,3), img_pts (n,1,2), valid_mask)."""
        xyz = frame_u8[v1:v2, u1:u2].astype(np.float64) / 255 * 2 - 1
        xyz = xyz * self.obj_scale[None, None, :] + self.obj_ct[None, None, :]
        valid_mask = np.logical_and(non_zero, prob_ori < self.th_i)
        vs, us = np.where(valid_mask == 1)
        obj = xyz[vs, us]
        img = np.stack((us + u1, vs + v1), axis=1).astype(np.float64).reshape(-1, 1, 2)
        return obj, np.ascontiguousarray(img), valid_mask

    def pnp_ransac(self, frame_u8, prob_ori, non_zero, v1, v2, u1, u2):
        """recognition.py:195-224."""
        obj, img, valid_mask = self.build_correspondences(frame_u8, prob_ori, non_zero, v1, v2, u1, u2)
        if obj.shape[0] < 6:                                                      # :214-215
            return np.eye(3), np.array([0, 0, 0]), valid_mask, -1
        ret, rvec, tvec, inliers = cv2.solvePnPRansac(obj, img, self.camK, None, flags=cv2.SOLVEPNP_EPNP,
                                                      reprojectionError=5, iterationsCount=100)   # :216-217
        if inliers is None:                                                       # :218-219
            return np.eye(3), np.array([0, 0, 0]), -1, -1
        R = np.eye(3)
        cv2.Rodrigues(rvec, R)                                                    # :223
        return R, tvec[:, 0], valid_mask, len(inliers)
