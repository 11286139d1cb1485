"""GPU parity of the est_pose device pipeline (csrc/pipeline.cu) against the CPU restatement of
recognition.py, through the drop-in class pix2pose_b200.recognition.pix2pose."""
import ctypes

import numpy as np
import pytest

from pix2pose_b200 import weights as W
from tests.planted import K_LM, OBJ, planted_case, rodrigues

pytestmark = pytest.mark.gpu
TH = dict(th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15)       # cfg/cfg_bop2020_rgb.json:8-9


@pytest.fixture(scope="module")
def rec():
    from pix2pose_b200.recognition import pix2pose
    return pix2pose(W.synthetic_weights("resnet50", 1), K_LM, 640, 480, OBJ, backbone="resnet50", capacity=16, max_dets=16, **TH)


@pytest.fixture(scope="module")
def frame():
    f = np.random.RandomState(0).randint(0, 256, (480, 640, 3)).astype(np.uint8)
    f[150:330, 230:410] = (f[150:330, 230:410] // 4 + 100).astype(np.uint8)
    return f


ROIS = [[197, 277, 283, 363], [100, 200, 260, 330], [-20, -10, 90, 120], [400, 560, 500, 660], [50, 60, 300, 420]]


def _cand_crop(rec, k, box, cap_dummy=None):
    """uint8 XYZ crop + valid mask of candidate k of detection 0 (parity hook)."""
    from pix2pose_b200.recognition import _Pose
    p = _Pose()
    p.best_cand = k
    p.best_box[:] = [int(v) for v in box]
    return rec._fetch_crop(0, p)


@pytest.mark.parametrize("roi", ROIS)
def test_stagewise_bit_parity_with_same_network_outputs(rec, frame, roi):
    """Oracle and device pipeline are fed the SAME generator (the GPU one): the crops, masks, uint8 XYZ
    maps and valid masks of every candidate are integer / byte work and must agree bit for bit."""
    from oracle.recognition_oracle import Pix2PoseOracle
    ora = Pix2PoseOracle(rec.generator_train, K_LM, 640, 480, OBJ, **TH)
    ora.trace = {}
    want = ora.est_pose(frame, np.array(roi))
    got = rec.est_pose(frame, np.array(roi))
    assert list(got[5]) == list(want[5])
    assert np.array_equal(rec.debug_fetch(3, 0), ora.trace["x1"].astype(np.float32))
    for k in range(len(ora.trace["x2"])):
        assert np.array_equal(rec.debug_fetch(4, k), ora.trace["x2"][k].astype(np.float32)), k
    assert len(ora.trace["cands"]) > 0
    for c in ora.trace["cands"]:
        xyz, mask, _ = _cand_crop(rec, c["cid"], c["box"])
        assert np.array_equal(xyz, c["xyz_u8"]), c["cid"]
        assert np.array_equal(mask, np.asarray(c["valid_mask"], bool)), c["cid"]
    assert isinstance(got[1], int) == isinstance(want[1], int)
    if not isinstance(want[1], int):
        assert got[1].dtype == bool and got[1].shape == frame.shape[:2] and got[0].dtype == np.uint8
        assert got[2].shape == (3, 3) and got[3].shape == (3,)


# The configurations the reference ships: thresholds, backbone, frame size and frame dtype.
K_TLESS = np.array([[1075.65, 0, 360.0], [0, 1073.90, 270.0], [0, 0, 1]])
SHIPPED = {
    "cfg_bop2020": dict(th_outlier=[0.2, 0.3, 0.35], th_inlier=0.2, backbone="resnet50", hw=(480, 640), K=K_LM),          # cfg_bop2020.json:2,8-9
    "cfg_bop2019_paper": dict(th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15, backbone="paper", hw=(480, 640), K=K_LM),     # cfg_bop2019.json:2,8-9
    "cfg_tless_paper": dict(th_outlier=[[0.3]], th_inlier=0.1, backbone="paper", hw=(540, 720), K=K_TLESS),               # cfg_tless_paper.json:12-13 via 5_evaluation_bop_basic.py:217
    "ros_config": dict(th_outlier=[0.1, 0.2, 0.3, 0.4], th_inlier=0.15, backbone="resnet50", hw=(480, 640), K=K_LM),       # ros_kinetic/ros_config.json:7-9
    "icp3d_float32": dict(th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15, backbone="resnet50", hw=(480, 640), K=K_LM, f32=True),  # 5_evaluation_bop_icp3d.py:369-370
}


@pytest.mark.parametrize("name", sorted(SHIPPED))
def test_stagewise_bit_parity_in_shipped_configurations(name):
    """Same bit-parity bar as above for every configuration the reference ships: one threshold (nested list, as the T-LESS
    config hands it over), four thresholds, the 2020 thresholds, 720x540 frames, the paper backbone, and the float32 frame
    of the ICP driver (invalid-depth pixels scaled by 0.1, i.e. non-integer pixel values)."""
    from oracle.recognition_oracle import Pix2PoseOracle
    from pix2pose_b200.recognition import pix2pose
    c = SHIPPED[name]
    H, W = c["hw"]
    r = pix2pose(W_synth(c["backbone"]), c["K"], W, H, OBJ, backbone=c["backbone"], capacity=16, max_dets=16,
                 th_outlier=c["th_outlier"], th_inlier=c["th_inlier"])
    rng = np.random.RandomState(3)
    f = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    f[H // 3:2 * H // 3, W // 3:2 * W // 3] = (f[H // 3:2 * H // 3, W // 3:2 * W // 3] // 4 + 100).astype(np.uint8)
    if c.get("f32"):
        f = f.astype(np.float32)
        bad = rng.rand(H, W) < 0.3
        f[bad] = 0.1 * f[bad]                                                    # 5_evaluation_bop_icp3d.py:370
    ora = Pix2PoseOracle(r.generator_train, c["K"], W, H, OBJ, th_outlier=c["th_outlier"], th_inlier=c["th_inlier"])
    n_cands = 0
    for roi in ([H // 2 - 43, W // 2 - 43, H // 2 + 43, W // 2 + 43], [-20, -10, 90, 120], [H - 180, W - 230, H - 10, W + 20]):
        ora.trace = {}
        want = ora.est_pose(f, np.array(roi))
        got = r.est_pose(f, np.array(roi))
        assert list(got[5]) == list(want[5]), (name, roi)
        assert np.array_equal(r.debug_fetch(3, 0), ora.trace["x1"].astype(np.float32)), (name, roi)
        for k in range(len(ora.trace["x2"])):
            assert np.array_equal(r.debug_fetch(4, k), ora.trace["x2"][k].astype(np.float32)), (name, roi, k)
        for cd in ora.trace["cands"]:
            xyz, mask, _ = _cand_crop(r, cd["cid"], cd["box"])
            assert np.array_equal(xyz, cd["xyz_u8"]), (name, roi, cd["cid"])
            if not isinstance(cd["valid_mask"], int):         # -1: PnP found no consensus (recognition.py:219), no mask to compare
                assert np.array_equal(mask, np.asarray(cd["valid_mask"], bool)), (name, roi, cd["cid"])
            n_cands += 1
        assert isinstance(got[1], int) == isinstance(want[1], int)
        if not isinstance(want[1], int):
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[4] == want[4]
            assert np.abs(got[2] - want[2]).max() <= 1e-9 and np.abs(got[3] - want[3]).max() <= 1e-6 * np.abs(want[3]).max()
    assert n_cands > 0, name


def W_synth(backbone):
    return W.synthetic_weights(backbone, 1)


def test_refined_box_centre_minus_one_quirk(rec, frame):
    """recognition.py:29 treats ``ct[0] == -1`` as "no centre given"; the stage-2 call (:110) passes the computed centre
    ``[cy_m, cx_m]``, so a mask centroid that lands on row -1 (ROIs at the top edge) silently falls back to the centre of
    the 128-space bbox.  Planted stage-1 maps put the centroid exactly there; the device path must build the same refined
    box as the oracle (= the reference's behaviour)."""
    from oracle.recognition_oracle import Pix2PoseOracle
    from tests.planted import PlantedGenerator
    roi = np.array([-40, 260, 40, 340])                     # cy_o = 0
    dec = np.zeros((1, 128, 128, 3), np.float32)
    dec[0, 30:95, 20:108] = 0.5                            # rows 30..94: mean 62 -> cy_m = int(62 - 63.5 + 0) = int(-1.5) = -1
    prob = np.full((1, 128, 128, 1), 0.05, np.float32)
    s2 = (np.tile(dec, (3, 1, 1, 1)), np.tile(prob, (3, 1, 1, 1)))
    ora = Pix2PoseOracle(PlantedGenerator((dec, prob), s2), K_LM, 640, 480, OBJ, **TH)
    ora.trace = {}
    want = ora.est_pose(frame, roi)
    assert len(ora.trace["boxes2"]) == 3
    rec.debug_override(1, dec, prob)
    rec.debug_override(2, s2[0], s2[1])
    got = rec.est_pose(frame, roi)
    assert list(got[5]) == list(want[5])
    for k in range(len(ora.trace["x2"])):
        assert np.array_equal(rec.debug_fetch(4, k), ora.trace["x2"][k].astype(np.float32)), k
    # and the quirk really was exercised: with the centre honoured the box would sit one row higher
    from oracle.recognition_oracle import get_boxes
    b = ora.trace["boxes2"][0]
    assert (b[0] + b[1]) // 2 != -1


def test_non_default_box_size_matches_oracle(frame):
    """box_size is a constructor argument of the reference (recognition.py:10, :19); the device-side refined-box
    arithmetic follows it (p2p_pipeline_set_box_size)."""
    from oracle.recognition_oracle import Pix2PoseOracle
    from pix2pose_b200.recognition import pix2pose
    r = pix2pose(W.synthetic_weights("resnet50", 1), K_LM, 640, 480, OBJ, backbone="resnet50", capacity=16, max_dets=16, box_size=1.2, **TH)
    ora = Pix2PoseOracle(r.generator_train, K_LM, 640, 480, OBJ, box_size=1.2, **TH)
    ora.trace = {}
    roi = np.array([150, 250, 290, 370])
    want = ora.est_pose(frame, roi)
    got = r.est_pose(frame, roi)
    assert list(got[5]) == list(want[5])
    assert np.array_equal(r.debug_fetch(3, 0), ora.trace["x1"].astype(np.float32))
    for k in range(len(ora.trace["x2"])):
        assert np.array_equal(r.debug_fetch(4, k), ora.trace["x2"][k].astype(np.float32)), k
    for c in ora.trace["cands"]:
        xyz, mask, _ = _cand_crop(r, c["cid"], c["box"])
        assert np.array_equal(xyz, c["xyz_u8"]) and np.array_equal(mask, np.asarray(c["valid_mask"], bool)), c["cid"]
    # and the default is restored for other objects sharing the pipeline
    r2 = pix2pose(W.synthetic_weights("resnet50", 1), K_LM, 640, 480, OBJ, backbone="resnet50", capacity=16, max_dets=16, **TH)
    ora2 = Pix2PoseOracle(r2.generator_train, K_LM, 640, 480, OBJ, **TH)
    assert list(r2.est_pose(frame, roi)[5]) == list(ora2.est_pose(frame, roi)[5])


def test_sentinel_returns(rec, frame):
    out = rec.est_pose(frame, np.array([200, 300, 203, 302]))          # crop < 5 px (recognition.py:78-79)
    assert out[0].shape == (1,) and out[1] == -1 and out[2] == -1 and out[3] == -1 and out[4] == -1
    assert list(out[5]) == [199, 203, 299, 303]
    zero = rec.est_pose(np.zeros((480, 640, 3), np.uint8), np.array([0, 0, 128, 128]))   # the reference's warm-up call
    assert len(zero) == 6


@pytest.mark.parametrize("seed,roi,rv,t", [(0, [182, 290, 286, 386], [0.4, -0.3, 0.2], [15.0, -10.0, 700.0]),
                                           (1, [120, 200, 330, 420], [-0.7, 0.5, 1.1], [-20.0, -15.0, 420.0]),
                                           (2, [300, 420, 470, 630], [0.1, 0.9, -0.4], [180.0, 140.0, 800.0])])
def test_planted_pose_end_to_end(rec, frame, seed, roi, rv, t):
    """Planted XYZ maps (ellipsoid under a known pose, 10 % outliers) replace the network outputs on both
    sides.  All byte/integer stages must agree bit for bit, the same candidate must win, and the final R|t must equal the
    oracle's (real cv2.solvePnPRansac) to 1e-5 deg / 1e-9 relative (SURVEY section 8c allows 0.1 deg / 1e-3); both sit
    within 2.5 deg / 3 % of the planted truth (uint8 XYZ quantisation, 10 % outliers, clipped crops)."""
    from oracle.recognition_oracle import Pix2PoseOracle
    R, t = rodrigues(rv), np.array(t)
    want, s1, s2, ora = planted_case(Pix2PoseOracle, frame, roi, R, t, seed=seed, **TH)
    rec.debug_override(1, s1[0], s1[1])
    rec.debug_override(2, s2[0], s2[1])
    got = rec.est_pose(frame, np.array(roi))
    assert not isinstance(want[1], int) and not isinstance(got[1], int)
    assert list(got[5]) == list(want[5])
    for c in ora.trace["cands"]:
        xyz, mask, _ = _cand_crop(rec, c["cid"], c["box"])
        assert np.array_equal(xyz, c["xyz_u8"]) and np.array_equal(mask, np.asarray(c["valid_mask"], bool))
    ang = np.degrees(np.arccos(np.clip((np.trace(want[2].T @ got[2]) - 1) / 2, -1, 1)))
    assert ang <= 1e-5 and np.linalg.norm(got[3] - want[3]) / np.linalg.norm(want[3]) <= 1e-9
    assert got[4] == want[4]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    ang_true = np.degrees(np.arccos(np.clip((np.trace(R.T @ got[2]) - 1) / 2, -1, 1)))
    assert ang_true < 2.5 and np.linalg.norm(got[3] - t) / np.linalg.norm(t) < 0.03


def test_batch_equals_singles_at_config3_size(rec):
    """BASELINE config 3 shape (256 detections on 16 frames): the batched pipeline returns, for every
    detection, exactly what a batch of one returns (detections are independent)."""
    from pix2pose_b200.recognition import pix2pose
    big = pix2pose(rec.generator_train.weights, K_LM, 640, 480, OBJ, backbone="resnet50", capacity=64, max_dets=256, **TH)
    rng = np.random.RandomState(7)
    frames = rng.randint(0, 256, (16, 480, 640, 3)).astype(np.uint8)
    rois, fids = [], []
    for f in range(16):
        for _ in range(16):
            cy, cx, h, w = rng.randint(60, 420), rng.randint(60, 580), rng.randint(40, 120), rng.randint(40, 120)
            rois.append([cy - h // 2, cx - w // 2, cy + h // 2, cx + w // 2]); fids.append(f)
    res = big.est_pose_batch(frames, rois, fids)
    assert res.n == 256 and set(np.unique(res.status)) <= {-2, 0, 1}
    assert (res.status == 1).sum() >= 128
    for i in (0, 17, 100, 255):
        one = big.est_pose_batch(frames[fids[i]], [rois[i]])
        assert one.status[0] == res.status[i] and np.array_equal(one.bbox_t[0], res.bbox_t[i])
        assert np.array_equal(one.R[0], res.R[i]) and np.array_equal(one.t[0], res.t[i])
        assert one.n_inliers[0] == res.n_inliers[i]
    rec16 = res.records()
    assert rec16.shape == (256, 16) and np.array_equal(rec16[:, 15], np.arange(256))


def test_multi_object_stream_matches_per_object_calls(frame):
    """SURVEY §8d config 4 in miniature: two objects (two weight sets, one engine + pipeline), interleaved
    detection stream; every detection gets exactly the pose its own object's recogniser computes alone."""
    from pix2pose_b200.stream import MultiObjectRecognizer
    objs = {3: np.array([50., 40., 60., 0., 0., 0.]), 7: np.array([30., 30., 80., 1., -2., 3.])}
    wts = {3: W.synthetic_weights("paper", 1), 7: W.synthetic_weights("paper", 2)}
    multi = MultiObjectRecognizer(wts, K_LM, 640, 480, objs, backbone="paper", capacity=16, max_dets=16)
    rng = np.random.RandomState(3)
    frames = np.stack([frame, frame[::-1].copy()])
    rois, oids, fids = [], [], []
    for i in range(10):
        cy, cx, h, w = rng.randint(80, 400), rng.randint(80, 560), rng.randint(50, 120), rng.randint(50, 120)
        rois.append([cy - h // 2, cx - w // 2, cy + h // 2, cx + w // 2]); oids.append(3 if i % 2 == 0 else 7); fids.append(i % 2)
    rec, status = multi.est_pose_stream(frames, rois, oids, fids)
    assert rec.shape == (10, 16) and np.array_equal(rec[:, 15], np.arange(10))
    for i in range(10):
        one = multi.models[oids[i]].est_pose_batch(frames[fids[i]], [rois[i]])
        assert one.status[0] == status[i]
        assert np.array_equal(one.R[0].ravel(), rec[i, :9]) and np.array_equal(one.t[0], rec[i, 9:12])
    assert len({tuple(rec[0, :9]), tuple(rec[1, :9])}) == 2


def test_evaluation_loop_batched_equals_per_roi(frame):
    """the per-image loop of tools/5_evaluation_bop_basic.py:281-349 with the per-object device batches
    gives exactly the rows the reference-style one-est_pose-per-ROI loop gives (same recognisers, same detections)."""
    from pix2pose_b200 import evaluation as E
    from pix2pose_b200.recognition import pix2pose
    targets, inst_counts = [3, 7], [2, 1]
    objs = {3: np.array([50., 40., 60., 0., 0., 0.]), 7: np.array([30., 30., 80., 1., -2., 3.])}
    recs = [pix2pose(W.synthetic_weights("paper", s), K_LM, 640, 480, objs[o], backbone="paper", capacity=16, max_dets=16,
                     th_outlier=[0.15, 0.25, 0.35], th_inlier=0.15) for o, s in ((3, 1), (7, 2))]
    rng = np.random.RandomState(5)
    rois, oids = [], []
    for i in range(9):
        cy, cx, h, w = rng.randint(80, 400), rng.randint(80, 560), rng.randint(50, 120), rng.randint(50, 120)
        rois.append([cy - h // 2, cx - w // 2, cy + h // 2, cx + w // 2]); oids.append([3, 7, 11][i % 3])
    rois[2] = [-1, -1, 4, 4]
    orders = [targets.index(o) if o in targets else 0 for o in oids]
    scores = rng.uniform(0.5, 1, 9)
    masks = np.zeros((480, 640, 9), bool)
    for i, r in enumerate(rois):
        masks[max(r[0], 0):r[2], max(r[1], 0):r[3], i] = True
    a = E.recognize_image(recs, frame, rois, orders, oids, scores, masks, targets, inst_counts, K_LM, backend=E.est_pose_batched)
    b = E.recognize_image(recs, frame, rois, orders, oids, scores, masks, targets, inst_counts, K_LM, backend=E.est_pose_loop)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x["obj_id"] == y["obj_id"] and x["score"] == y["score"] and np.array_equal(x["R"], y["R"]) and np.array_equal(x["t"], y["t"])


def test_device_mask_iou_matches_numpy(rec, frame):
    """|mask_pred & m| and |mask_pred | m| counted on the device equal numpy on the full-frame masks
    (tools/5_evaluation_bop_basic.py:307-316)."""
    rng = np.random.RandomState(11)
    rois = [[197, 277, 283, 363], [100, 200, 260, 330], [-20, -10, 90, 120], [50, 60, 300, 420]]
    res = rec.est_pose_batch(frame, [np.array(r) for r in rois])
    masks = rng.rand(len(rois), 480, 640) > 0.5
    for i, r in enumerate(rois):
        masks[i, :max(r[0], 0)] = False
        masks[i, :, r[3]:] = False
    inter, union = res.mask_iou(masks)
    for i in range(len(rois)):
        full = np.zeros((480, 640), bool)
        if res.status[i] == 1:
            _, m, bx = res.crop(i)
            full[bx[4]:bx[5], bx[6]:bx[7]] = m
        assert inter[i] == np.sum(full & masks[i]) and union[i] == np.sum(full | masks[i]), i
    assert (res.status == 1).any()


def test_tless_like_stream_per_object_thresholds():
    """SURVEY section 8d config 5 in miniature: 720x540 frames, one outlier threshold PER OBJECT handed over as a nested list
    (cfg_tless_paper.json:12, 5_evaluation_bop_basic.py:217), all objects' detections in one device run: every detection
    gets exactly what its own object's recogniser returns for it alone."""
    from pix2pose_b200.stream import MultiObjectRecognizer
    ths = {1: [[0.15]], 4: [[0.3]], 5: [[0.2]]}
    objs = {1: np.array([50., 40., 60., 0., 0., 0.]), 4: np.array([30., 30., 80., 1., -2., 3.]), 5: np.array([45., 45., 20., 0., 5., 0.])}
    wts = {o: W.synthetic_weights("resnet50", o) for o in ths}
    multi = MultiObjectRecognizer(wts, K_TLESS, 720, 540, objs, th_outlier=ths, th_inlier=0.15, backbone="resnet50", capacity=16, max_dets=32)
    rng = np.random.RandomState(8)
    frames = rng.randint(0, 256, (2, 540, 720, 3)).astype(np.uint8)
    rois, oids, fids = [], [], []
    for i in range(14):
        cy, cx, h, w = rng.randint(80, 460), rng.randint(80, 640), rng.randint(50, 140), rng.randint(50, 140)
        rois.append([cy - h // 2, cx - w // 2, cy + h // 2, cx + w // 2]); oids.append([1, 4, 5, 9][i % 4]); fids.append(i % 2)
    rec, status = multi.est_pose_stream(frames, rois, oids, fids)
    assert rec.shape == (14, 16) and np.array_equal(rec[:, 15], np.arange(14))
    assert all(status[i] == -3 for i in range(14) if oids[i] == 9)
    for i in range(14):
        if oids[i] == 9:
            continue
        one = multi.models[oids[i]].est_pose_batch(frames[fids[i]], [rois[i]])
        assert one.status[0] == status[i], i
        assert np.array_equal(one.R[0].ravel(), rec[i, :9]) and np.array_equal(one.t[0], rec[i, 9:12]), i
        assert one.n_inliers[0] == rec[i, 12] and one.frac_inlier[0] == rec[i, 13]
    assert (status == 1).sum() >= 1, status


def test_stale_result_access_raises(rec, frame):
    """The device pipeline is shared by all objects: reading a result's crops after another run would return that run's
    pool data, so it must raise instead."""
    res = rec.est_pose_batch(frame, [np.array(ROIS[0])])
    assert res.status[0] == 1
    res.crop(0)
    rec.est_pose_batch(frame, [np.array(ROIS[1])])
    with pytest.raises(RuntimeError):
        res.crop(0)
    with pytest.raises(RuntimeError):
        res.mask_iou(np.zeros((1, 480, 640), bool))


def test_graph_replay_equals_plain_launches(rec, frame):
    """The captured CUDA graph of a run and the kernel-by-kernel path give identical records, call after call."""
    import os
    rois = [np.array(r) for r in ROIS]
    a = rec.est_pose_batch(frame, rois)
    b = rec.est_pose_batch(frame, rois)          # replay of the graph captured by the first call
    assert np.array_equal(a.R, b.R) and np.array_equal(a.t, b.t) and np.array_equal(a.status, b.status)
    assert np.array_equal(a.n_inliers, b.n_inliers) and np.array_equal(a.bbox_t, b.bbox_t)
    l0 = rec.launch_count
    rec.est_pose_batch(frame, rois)
    assert rec.launch_count - l0 >= 80           # replays still count their kernels (bench.py's gpu_launches): 2 forwards of 35 + 11


@pytest.mark.parametrize("seed,roi,rv,t", [(0, [182, 290, 286, 386], [0.4, -0.3, 0.2], [15.0, -10.0, 700.0]),
                                           (2, [300, 420, 470, 630], [0.1, 0.9, -0.4], [180.0, 140.0, 800.0])])
def test_candidate_selection_fed_with_cv2_results(rec, frame, seed, roi, rv, t):
    """A18 in isolation (recognition.py:158-178, :189-193): the device selection is re-run on the candidates of a planted
    run with the PnP results of the ORACLE (cv2's own R|t and inlier count per candidate, bit for bit) and must then return
    exactly the oracle's choice: the same R|t bits, inlier fraction and bbox_t -- including a variant where PnP failed for
    the would-be winner (t = 0 -> dist 99999, recognition.py:163-165) and one where it returned inliers = None (:219)."""
    from oracle.recognition_oracle import Pix2PoseOracle
    R, t = rodrigues(rv), np.array(t)
    want, s1, s2, ora = planted_case(Pix2PoseOracle, frame, roi, R, t, seed=seed, **TH)
    rec.debug_override(1, s1[0], s1[1])
    rec.debug_override(2, s2[0], s2[1])
    rec.est_pose(frame, np.array(roi))
    cands = ora.trace["cands"]
    assert len(cands) >= 2
    K = K_LM

    def reference_choice(cs):               # recognition.py:158-178 on the per-candidate values
        best, min_dist, max_inl = None, 9999999, -1
        for c in cs:
            if c["t"][2] == 0:
                dist = 99999
            else:
                dist = c["dist"] if c.get("dist_override") is None else c["dist_override"]
            if dist < min_dist:
                best, min_dist, max_inl = c, dist, c["n_inliers"]
        return best, max_inl

    for variant in ("as_is", "winner_pnp_failed", "winner_no_inliers"):
        cs = [dict(c) for c in cands]
        win0, _ = reference_choice(cs)
        if variant == "winner_pnp_failed":          # pnp_ransac's `n_pts < 6` return: eye(3), zeros, -1 inliers (:214-215)
            win0["R"], win0["t"], win0["n_inliers"] = np.eye(3), np.zeros(3), -1
        elif variant == "winner_no_inliers":        # inliers is None: eye(3), zeros, valid_mask -1, -1 (:218-219)
            win0["R"], win0["t"], win0["n_inliers"], win0["none"] = np.eye(3), np.zeros(3), -1, True
        Rt = np.array([np.concatenate([np.asarray(c["R"], float).ravel(), np.asarray(c["t"], float).ravel()]) for c in cs])
        ninl = [int(c["n_inliers"]) for c in cs]
        # status 1 = pose with inliers; 0 = the reference's "inliers is None" return; a failed pnp (-1 inliers, t = 0) also arrives as 0
        status = [1 if c["n_inliers"] >= 0 else 0 for c in cs]
        pose = rec.debug_select(Rt, ninl, status)
        best, max_inl = reference_choice(cs)
        if max_inl == -1:
            assert pose.status != 1
            continue
        assert pose.status == 1 and pose.best_cand == best["cid"], variant
        assert np.array_equal(np.array(list(pose.R)).reshape(3, 3), np.asarray(best["R"], float)), variant
        assert np.array_equal(np.array(list(pose.t)), np.asarray(best["t"], float).ravel()), variant
        assert pose.n_inliers == max_inl
        assert list(pose.bbox_t) == list(want[5])


def test_end_to_end_cpu_oracle_vs_gpu_on_real_network_outputs(frame, capsys):
    """Oracle generator (torch-CPU fp32) -> oracle pipeline against GPU generator -> GPU pipeline, on REAL network outputs
    (VERDICT r1, weak #3).  The two generators agree to ~2e-4 (tests/test_net_gpu.py), so a pixel whose ||decode|| or prob
    sits within that distance of a threshold (0.3, th_o, th_i) may flip, and `value * 255` may truncate to the neighbouring
    uint8 level.  The test bounds how far that propagates -- same pose / sentinel decision for every detection, a handful of
    flipped stage-1 mask pixels, and for the detections that return the same pose, uint8 XYZ crops that differ by at most
    one level in at most 1 % of the pixels -- and prints the measured fractions (profiles/r02_e2e_parity.md)."""
    import json
    from oracle.net_oracle import NetOracle
    from oracle.recognition_oracle import Pix2PoseOracle
    from pix2pose_b200.recognition import pix2pose
    w = W.synthetic_weights("resnet50", 1)
    r = pix2pose(w, K_LM, 640, 480, OBJ, backbone="resnet50", capacity=16, max_dets=16, **TH)
    ora = Pix2PoseOracle(NetOracle(w, "resnet50"), K_LM, 640, 480, OBJ, **TH)
    rng = np.random.RandomState(21)
    rois = [ROIS[0], ROIS[1], ROIS[4]]
    for _ in range(9):
        cy, cx, h, ww = rng.randint(100, 380), rng.randint(100, 540), rng.randint(60, 130), rng.randint(60, 130)
        rois.append([cy - h // 2, cx - ww // 2, cy + h // 2, cx + ww // 2])
    st = dict(n=0, same_outcome=0, same_bbox_t=0, same_geometry=0, same_pose=0, crops_identical=0, px=0, px_diff=0, px_diff_gt1=0, mask_px_diff=0,
              stage1_px=0, stage1_mask_flips=0)
    ang, dt = [], []
    for roi in rois:
        ora.trace = {}
        want = ora.est_pose(frame, np.array(roi))
        got = r.est_pose(frame, np.array(roi))
        st["n"] += 1
        same = isinstance(want[1], int) == isinstance(got[1], int)
        st["same_outcome"] += same
        st["same_bbox_t"] += list(want[5]) == list(got[5])
        # stage 1: the crops are bit-identical by construction (same frame, same box); count mask pixels that flip
        d1, p1 = r.debug_fetch(1, 0), r.debug_fetch(5, 0)
        ng_g = np.linalg.norm(d1, axis=2) > 0.3
        ng_o = np.linalg.norm(ora.trace["decode1"][0], axis=2) > 0.3
        st["stage1_px"] += ng_g.size * (1 + len(TH["th_outlier"]))
        st["stage1_mask_flips"] += int((ng_g != ng_o).sum())
        for th in TH["th_outlier"]:
            st["stage1_mask_flips"] += int(((p1 < th) != (ora.trace["prob1"][0, :, :, 0] < th)).sum())
        if not same or isinstance(want[1], int):
            continue
        ang.append(float(np.degrees(np.arccos(np.clip((np.trace(want[2].T @ got[2]) - 1) / 2, -1, 1)))))
        dt.append(float(np.linalg.norm(want[3] - got[3]) / np.linalg.norm(want[3])))
        if list(want[5]) == list(got[5]) and want[0].shape == got[0].shape:
            st["same_geometry"] += 1
        if ang[-1] <= 1e-5 and dt[-1] <= 1e-9 and want[0].shape == got[0].shape:      # same winner, same consensus set
            st["same_pose"] += 1
            d = np.abs(want[0].astype(int) - got[0].astype(int))
            st["crops_identical"] += int(d.max() == 0)
            st["px"] += d.size
            st["px_diff"] += int((d > 0).sum())
            st["px_diff_gt1"] += int((d > 1).sum())
            st["mask_px_diff"] += int((want[1] != got[1]).sum())
    st.update(pose_pairs=len(ang), ang_median=float(np.median(ang)) if ang else None, ang_max=max(ang, default=None),
              dt_median=float(np.median(dt)) if dt else None)
    with capsys.disabled():
        print("\nE2E_PARITY " + json.dumps(st))
    assert st["same_outcome"] == st["n"]                              # pose / sentinel decision agrees everywhere
    assert st["stage1_mask_flips"] <= 2e-3 * st["stage1_px"]          # ~2e-4 network error x density of values near a threshold
    # measured (profiles/r02_e2e_parity.md): 11 of 12 detections return the SAME pose to 1e-14 (same winning candidate, same
    # RANSAC consensus set) and uint8 XYZ crops that differ in ~0.1 % of the pixels, by one level; the twelfth picks another
    # candidate (on noise-like maps PnP poses are near-random and a single flipped correspondence changes the consensus)
    assert st["same_pose"] >= 0.75 * st["pose_pairs"]
    assert st["px_diff_gt1"] == 0 and st["px_diff"] <= 0.01 * st["px"]


def test_async_batcher_equals_synchronous_calls(rec, frame):
    """Two batches in flight on two alternating device pipelines (stream.AsyncBatcher) return exactly what the synchronous
    est_pose_batch returns for the same detections, in submission order, also when the batches differ."""
    from pix2pose_b200.stream import AsyncBatcher
    rois_a = np.array(ROIS)
    rois_b = np.array(ROIS[::-1][:3])
    want_a = rec.est_pose_batch(frame, rois_a).records()
    want_b = rec.est_pose_batch(frame, rois_b).records()
    b = AsyncBatcher(rec, 16)
    fdev = rec.upload_frames(frame, 16)
    t1 = b.submit(fdev, rois_a, np.zeros(len(rois_a), int))
    t2 = b.submit(fdev, rois_b, np.zeros(len(rois_b), int))
    with pytest.raises(RuntimeError):
        b.submit(fdev, rois_a, np.zeros(len(rois_a), int))           # both pipelines busy
    r1, s1 = b.result(t1)
    t3 = b.submit(b.upload_frames(frame), rois_a, np.zeros(len(rois_a), int))   # per-batch upload on the batcher's own copy stream
    r2, s2 = b.result(t2)
    r3, s3 = b.result(t3)
    for got, want in ((r1, want_a), (r2, want_b), (r3, want_a)):
        assert np.array_equal(got[:, :15], want[:, :15])
    assert b.last_forward_ms(t3[0]) > 0
    b.close()
