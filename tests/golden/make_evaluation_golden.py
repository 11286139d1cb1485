"""Generates tests/golden/evaluation_golden.npz by executing the reference driver's own per-image loop
(/root/reference/tools/5_evaluation_bop_basic.py:281-349: ROI filter, est_pose calls, score_type-2 scoring, normalise /
sort, ViVo filter, result rows) on seeded detections.  The script file cannot be run (it loads datasets, Mask R-CNN and
Keras at module level), so exactly those source lines are read, de-indented, wrapped in a one-pass loop (they contain
``continue`` / ``break``) and exec'd in a namespace holding the variables the surrounding script would have set; the
recognisers are the deterministic FakeRec of tests/test_evaluation_host.py.  Nothing of the reference is copied into the
repository.  Run in the build container:  python tests/golden/make_evaluation_golden.py"""
import os
import sys
import textwrap
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
np.int, np.float = int, float                      # aliases numpy < 1.24 provided (:304, :310)
from tests.test_evaluation_host import FakeRec, detections   # noqa: E402

SRC = "/root/reference/tools/5_evaluation_bop_basic.py"


def reference_loop_source():
    lines = open(SRC).read().split("\n")[280:349]          # file lines 281..349
    body = textwrap.dedent("\n".join(lines))
    return "for __once in (0,):\n" + textwrap.indent(body, "    ")


def main():
    code = compile(reference_loop_source(), SRC, "exec")
    image, targets, inst_counts, rois, obj_ids, obj_orders, scores, masks, K = detections()
    out = {}
    for score_type in (1, 2):
        for task_type in ("1", "2"):
            ns = dict(np=np, time=time, t1=time.time(), rois=np.array(rois), obj_ids=obj_ids, obj_id_targets=targets,
                      inst_count_pred=np.zeros(len(inst_counts)), inst_count_est=np.zeros(len(inst_counts)), inst_counts=inst_counts,
                      cand_factor=2, obj_orders=obj_orders, obj_pix2pose=[FakeRec(t) for t in targets], cam_K=K, image_t=image,
                      score_type=score_type, detect_type="rcnn", masks=masks, scores=scores, task_type=task_type, scene_id=5, im_id=9,
                      result_dataset=[], resize=None)
            exec(code, ns)
            rows = ns["result_dataset"]
            key = "s%d_t%s" % (score_type, task_type)
            out[key + "_obj"] = np.array([r["obj_id"] for r in rows])
            out[key + "_score"] = np.array([r["score"] for r in rows])
            out[key + "_R"] = np.array([r["R"] for r in rows])
            out[key + "_t"] = np.array([r["t"] for r in rows])
            print(key, len(rows), "rows")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "evaluation_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
