"""SURVEY.md §8d config 4 / 5 style stream: K objects (K weight sets resident, one shared engine + pipeline), detections
interleaved round-robin as a detector would emit them, grouped per object on the fly.
  python scripts/bench_stream.py [n_objects=8] [n_detections=512] [backbone=resnet50] [precision=fp16x3]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pix2pose_b200 import weights as W
from pix2pose_b200.stream import MultiObjectRecognizer

K = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])


def main():
    n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_det = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    bb = sys.argv[3] if len(sys.argv) > 3 else "resnet50"
    prec = sys.argv[4] if len(sys.argv) > 4 else "fp16x3"
    rng = np.random.RandomState(0)
    ids = list(range(1, n_obj + 1))
    t = time.time()
    wts = {i: W.synthetic_weights(bb, i) for i in ids}
    objs = {i: np.array([40. + i, 35. + i, 50. + i, 0., 0., 0.]) for i in ids}
    per_obj = (n_det + n_obj - 1) // n_obj
    multi = MultiObjectRecognizer(wts, K, 640, 480, objs, backbone=bb, precision=prec, capacity=max(64, 3 * per_obj), max_dets=per_obj)
    print("%d objects loaded in %.1f s" % (n_obj, time.time() - t), flush=True)
    n_frames = max(1, n_det // 16)
    frames = rng.randint(0, 256, (n_frames, 480, 640, 3)).astype(np.uint8)
    rois, oids, fids = [], [], []
    for d in range(n_det):
        cy, cx, h, w = rng.randint(80, 400), rng.randint(80, 560), rng.randint(50, 130), rng.randint(50, 130)
        rois.append([cy - h // 2, cx - w // 2, cy + h // 2, cx + w // 2]); oids.append(ids[d % n_obj]); fids.append(d % n_frames)
    multi.est_pose_stream(frames, rois, oids, fids)
    reps = 5
    t = time.perf_counter()
    for _ in range(reps):
        rec, status = multi.est_pose_stream(frames, rois, oids, fids)
    dt = (time.perf_counter() - t) / reps
    print("stream of %d detections over %d objects (%s, %s): %.1f ms -> %.0f detections/s (host frames in, records out); poses found %.2f"
          % (n_det, n_obj, bb, prec, dt * 1e3, n_det / dt, float((status == 1).mean())))


if __name__ == "__main__":
    main()
