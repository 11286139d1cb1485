"""Which half of epnp_refit_kernel sets its time?  P2P_PROF_PNP=1 prints the per-stage device times of three batches:
consensus sets of <= 32 points (exact serial refit, 32 problems per block), of ~60 points and of ~3500 points (one block each)."""
import os
import sys

import numpy as np

os.environ["P2P_PROF_PNP"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pix2pose_b200.pnp import solve_pnp_ransac_batch   # noqa: E402
from tests.planted import K_LM                         # noqa: E402
from tests.test_pnp_gpu import _planted                # noqa: E402

rng = np.random.RandomState(0)
for name, n_pts in (("small path only (24 points)", 24), ("large path, 60 points", 60), ("large path, 5000 points", 5000),
                    ("mixed: 1 small problem among 767 large", -1)):
    objs, imgs = [], []
    for i in range(768):
        n = n_pts if n_pts > 0 else (24 if i == 0 else 60)
        pw, uv = _planted(rng, n, 0.0, noise=0.5)
        objs.append(pw)
        imgs.append(uv)
    print("==", name, flush=True)
    for rep in range(3):
        solve_pnp_ransac_batch(objs, imgs, K_LM)
