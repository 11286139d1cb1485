"""Planted-pose fixtures: the network output is replaced by the analytic
XYZ map of an ellipsoid (semi-axes = obj_scale) under a known R|t, plus noise and outliers, so that
the two-stage pipeline has a meaningful pose to recover.  Test infrastructure."""
import numpy as np

K_LM = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])
OBJ = np.array([50., 40., 60., 0., 0., 0.])


def rodrigues(rv):
    import cv2
    return cv2.Rodrigues(np.asarray(rv, np.float64))[0]


def render_maps(box, R, t, K=K_LM, scale=OBJ[:3], rng=None, noise=0.01, outlier_frac=0.1):
    """decode (128,128,3) in [-1,1] and prob (128,128,1) for the square crop `box` (get_boxes tuple)."""
    v1o, v2o, u1o, u2o = box[0], box[1], box[2], box[3]
    sv, su = (v2o - v1o) / 128.0, (u2o - u1o) / 128.0
    ii, jj = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    v = v1o + (ii + 0.5) * sv - 0.5
    u = u1o + (jj + 0.5) * su - 0.5
    rays = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], -1)   # camera frame
    # ellipsoid in model frame: |x/scale| = 1 ; camera ray p = s*ray ; model x = R^T (p - t)
    d = rays @ R                      # R^T ray
    o = -(R.T @ t)                    # R^T (0 - t)
    dn, on = d / scale, o / scale
    a = (dn * dn).sum(-1)
    b = 2 * (dn * on).sum(-1)
    c = (on * on).sum() - 1
    disc = b * b - 4 * a * c
    hit = disc > 0
    s = np.where(hit, (-b - np.sqrt(np.where(hit, disc, 0))) / (2 * a), 0)
    xm = (s[..., None] * d + o) / scale          # normalised model coords on the surface, norm 1
    dec = np.where(hit[..., None], xm, 0.0)
    prob = np.where(hit, 0.03, 0.9)
    if rng is not None:
        dec = dec + noise * rng.randn(128, 128, 3) * hit[..., None]
        out = (rng.rand(128, 128) < outlier_frac) & hit
        dec[out] = rng.uniform(-1, 1, (int(out.sum()), 3))
        prob = prob + 0.01 * rng.rand(128, 128)
    return np.clip(dec, -1, 1).astype(np.float32), prob[..., None].astype(np.float32)


class PlantedGenerator:
    """``predict`` stand-in: call 1 returns the stage-1 maps, call 2 the stage-2 maps rendered for the
    refined boxes recorded by `boxes_for_stage2` (filled by the test after a dry run)."""

    def __init__(self, stage1, stage2=None):
        self.stage1, self.stage2, self.calls = stage1, stage2, 0

    def predict(self, x):
        self.calls += 1
        n = np.asarray(x).shape[0]
        dec, prob = self.stage1 if self.calls % 2 == 1 else self.stage2
        assert dec.shape[0] >= n, "planted stage-%d maps: have %d need %d" % (2 - self.calls % 2, dec.shape[0], n)
        return [dec[:n].copy(), prob[:n].copy()]


def planted_case(oracle_cls, frame, roi, R, t, seed=0, **kw):
    """Runs the oracle twice (dry run to learn the refined boxes, then with planted stage-2 maps).
    Returns (oracle_result, stage1_maps, stage2_maps, oracle)."""
    from oracle.recognition_oracle import get_boxes
    rng = np.random.RandomState(seed)
    H, W = frame.shape[:2]
    box1 = get_boxes(1.5, roi, H, W)
    d1, p1 = render_maps(box1, R, t, rng=rng)
    s1 = (d1[None], p1[None])
    dry = oracle_cls(PlantedGenerator(s1, (np.zeros((8, 128, 128, 3), np.float32), np.ones((8, 128, 128, 1), np.float32))),
                     K_LM, W, H, OBJ, **kw)
    dry.trace = {}
    dry.est_pose(frame, np.array(roi))
    boxes2 = dry.trace.get("boxes2", [])
    maps = [render_maps(b, R, t, rng=rng) for b in boxes2] or [(np.zeros((128, 128, 3), np.float32), np.ones((128, 128, 1), np.float32))]
    s2 = (np.stack([m[0] for m in maps]), np.stack([m[1] for m in maps]))
    ora = oracle_cls(PlantedGenerator(s1, s2), K_LM, W, H, OBJ, **kw)
    ora.trace = {}
    res = ora.est_pose(frame, np.array(roi))
    return res, s1, s2, ora
