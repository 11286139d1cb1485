"""GPU bring-up of the EPnP-RANSAC kernels against cv2.solvePnPRansac on planted poses."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np

from pix2pose_b200.pnp import solve_pnp_ransac

K = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])


def planted(rng, n, outlier_frac, noise):
    rv = rng.randn(3); rv *= rng.uniform(0.1, np.pi) / np.linalg.norm(rv)
    R, _ = cv2.Rodrigues(rv)
    t = np.array([rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(500, 1200)])
    pw = rng.uniform(-1, 1, (n, 3)) * np.array([50, 40, 60])
    pc = pw @ R.T + t
    uv = pc[:, :2] / pc[:, 2:] * np.array([K[0, 0], K[1, 1]]) + np.array([K[0, 2], K[1, 2]])
    uv += rng.randn(n, 2) * noise
    no = int(n * outlier_frac)
    idx = rng.choice(n, no, replace=False)
    uv[idx] += rng.uniform(-60, 60, (no, 2))
    return pw, uv, R, t


def main():
    rng = np.random.RandomState(0)
    worst_deg, worst_t = 0, 0
    for trial in range(40):
        n = int(rng.choice([6, 12, 50, 200, 2000, 8000, 16384]))
        of = float(rng.choice([0.0, 0.2, 0.4]))
        pw, uv, R, t = planted(rng, n, of, 1.0)
        t0 = time.time()
        ret, rv, tv, inl = cv2.solvePnPRansac(pw, uv.reshape(-1, 1, 2), K, None, flags=cv2.SOLVEPNP_EPNP,
                                              reprojectionError=5, iterationsCount=100)
        t1 = time.time()
        g_ret, g_rv, g_tv, g_inl, g_R, g_it = solve_pnp_ransac(pw, uv, K, 5.0, 100, 0.99)
        t2 = time.time()
        if not ret or inl is None:
            print("trial %d n=%d: cv2 failed ret=%s; gpu ret=%s" % (trial, n, ret, g_ret))
            continue
        Rcv, _ = cv2.Rodrigues(rv)
        dR = np.degrees(np.arccos(np.clip((np.trace(Rcv.T @ g_R) - 1) / 2, -1, 1)))
        dt = np.linalg.norm(g_tv[:, 0] - tv[:, 0]) / np.linalg.norm(tv)
        same = len(np.intersect1d(inl[:, 0], g_inl[:, 0])) if g_inl is not None else -1
        worst_deg, worst_t = max(worst_deg, dR), max(worst_t, dt)
        print("trial %2d n=%5d out=%.1f | inl cv %5d gpu %5d common %5d iters %3d | dR %.2e deg  dt %.2e | cv %.1f ms gpu %.1f ms" % (
            trial, n, of, len(inl), len(g_inl) if g_inl is not None else -1, same, g_it, dR, dt, (t1 - t0) * 1e3, (t2 - t1) * 1e3),
            flush=True)
    print("worst dR %.3e deg, worst dt %.3e" % (worst_deg, worst_t))


if __name__ == "__main__":
    main()
