"""Device time of the batched EPnP-RANSAC (p2p_pnp_ransac_batch) on 768 problems -- the stage-2 candidates of a 256-detection
bench step -- for well-posed batches (the replayed loop ends within the first wave of 32 hypotheses), hard ones (all 100
iterations, what the bench fixture with its random network outputs produces) and a mix.  Run on the GPU box:
  python scripts/bench_pnp_batch.py [n_problems] [points_per_problem]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pix2pose_b200.pnp import solve_pnp_ransac_batch   # noqa: E402
from tests.planted import K_LM                         # noqa: E402
from tests.test_pnp_gpu import _planted                # noqa: E402

n_prob = int(sys.argv[1]) if len(sys.argv) > 1 else 768
n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
rng = np.random.RandomState(0)
for name, fracs in (("30 % outliers", [0.3]), ("60 % outliers", [0.6]), ("85 % outliers", [0.85]), ("half 30 % / half 85 %", [0.3, 0.85])):
    objs, imgs = [], []
    for i in range(n_prob):
        pw, uv = _planted(rng, n_pts, fracs[i % len(fracs)], noise=1.0)
        objs.append(pw)
        imgs.append(uv)
    best, iters = 1e9, None
    for rep in range(4):
        t0 = time.perf_counter()
        res, ms = solve_pnp_ransac_batch(objs, imgs, K_LM, return_time=True)
        wall = time.perf_counter() - t0
        if rep:
            best = min(best, ms)
        iters = np.array([r[5] for r in res])
    found = sum(r[0] for r in res)
    print("%-24s %d problems x %d points: device %.3f ms (%.0f problems/s), host call %.0f ms; iterations run: median %d, max %d; poses %d"
          % (name, n_prob, n_pts, best, n_prob / best * 1e3, wall * 1e3, np.median(iters), iters.max(), found), flush=True)
