"""GPU probe: EPnP refit paths (small exact / large block-parallel) against cv2 on clean and noisy all-inlier sets."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from pix2pose_b200.pnp import solve_pnp_ransac
from tests.planted import K_LM
from tests.test_pnp_gpu import _planted

rng = np.random.RandomState(0)
for n in (6, 8, 20, 32, 33, 40, 81, 200, 2000, 16384):
    for of in (0.0, 0.6):
        pw, uv = _planted(rng, n, of, noise=1.0)
        ret, rv, tv, inl = cv2.solvePnPRansac(pw, uv.reshape(-1, 1, 2), K_LM, None, flags=cv2.SOLVEPNP_EPNP, reprojectionError=5,
                                              iterationsCount=100)
        g = solve_pnp_ransac(pw, uv, K_LM, 5.0, 100, 0.99)
        if inl is None or g[3] is None:
            print(n, of, "none", inl is None, g[3] is None)
            continue
        print("n %6d of %.1f inl %5d/%5d same %s  drv %.3e dt %.3e it %d" % (n, of, len(inl), len(g[3]), np.array_equal(inl, g[3]),
              np.abs(g[1] - rv).max(), np.abs(g[2] - tv).max(), g[5]), flush=True)
