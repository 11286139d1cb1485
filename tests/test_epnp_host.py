"""CPU: the __host__ __device__ EPnP / RANSAC building blocks (csrc/epnp_core.cuh, compiled for the host by
tests/host_harness) against the real OpenCV.  This is how the device math is checked without a GPU."""
import ctypes

import cv2
import numpy as np

from tests.planted import K_LM

DP = ctypes.POINTER(ctypes.c_double)
CAM = np.array([K_LM[0, 0], K_LM[1, 1], K_LM[0, 2], K_LM[1, 2]])


def P(a):
    return a.ctypes.data_as(DP)


def _case(rng, n, noise):
    rv = rng.randn(3); rv *= rng.uniform(0, np.pi) / np.linalg.norm(rv)
    R = cv2.Rodrigues(rv)[0]
    t = np.array([rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(400, 1200)])
    pw = rng.uniform(-1, 1, (n, 3)) * np.array([50, 40, 60])
    pc = pw @ R.T + t
    uv = pc[:, :2] / pc[:, 2:] * CAM[:2] + CAM[2:] + rng.randn(n, 2) * noise
    return np.ascontiguousarray(pw), np.ascontiguousarray(uv)


def test_linear_algebra_bitwise_equals_cv2(host_harness):
    """The restated OpenCV routines EPnP goes through (csrc/epnp_core.cuh: JacobiSVDImpl_ with OpenCV's own hypot, SVBkSb,
    MulTransposedR) against the cv2 functions themselves: EVERY output bit equal.  Includes the rank-10 12x12 case of the
    5-point solver, whose two null-space rows of U are pure rounding residue."""
    rng = np.random.RandomState(5)
    for _ in range(40):
        M = rng.randn(10, 12) * rng.uniform(1, 100)
        A = np.ascontiguousarray(M.T @ M)
        w, u, vt = cv2.SVDecomp(A.copy(), flags=cv2.SVD_FULL_UV)
        W, Ut, Vt = np.zeros(12), np.zeros((12, 12)), np.zeros((12, 12))
        host_harness.host_cv_svd12(P(A), P(W), P(Ut), P(Vt))
        assert np.array_equal(w[:, 0], W) and np.array_equal(u.T, Ut) and np.array_equal(vt, Vt)
    for _ in range(100):
        A = np.ascontiguousarray(rng.randn(3, 3) * rng.uniform(0.1, 50))
        if rng.rand() < 0.2:
            A[:, 2] = 0                                   # exactly rank deficient: the fixed-seed RNG completion path
        w, u, vt = cv2.SVDecomp(A.copy())
        W, Ut, Vt = np.zeros(3), np.zeros((3, 3)), np.zeros((3, 3))
        host_harness.host_cv_svd3(P(A), P(W), P(Ut), P(Vt))
        assert np.array_equal(w[:, 0], W) and np.array_equal(u.T, Ut) and np.array_equal(vt, Vt)
        inv = np.zeros((3, 3))
        host_harness.host_cv_invert3(P(A), P(inv))
        assert np.array_equal(cv2.invert(A, flags=cv2.DECOMP_SVD)[1], inv)
    for n in (3, 4, 5):
        for k in range(60):
            A = np.ascontiguousarray(rng.randn(6, n) if k % 3 else rng.randn(6, 2) @ rng.randn(2, n))   # full rank / rank 2
            b = np.ascontiguousarray(rng.randn(6))
            x = np.zeros(n)
            host_harness.host_cv_solve6(P(A), P(b), n, P(x))
            assert np.array_equal(cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_SVD)[1][:, 0], x)
    for n in (5, 9):
        for _ in range(20):
            al, us = np.ascontiguousarray(rng.randn(n, 4)), np.ascontiguousarray(rng.uniform(0, 640, (n, 2)))
            M = np.zeros((2 * n, 12))                     # epnp::fill_M
            for i in range(n):
                for j in range(4):
                    M[2 * i, 3 * j], M[2 * i, 3 * j + 2] = al[i, j] * CAM[0], al[i, j] * (CAM[2] - us[i, 0])
                    M[2 * i + 1, 3 * j + 1], M[2 * i + 1, 3 * j + 2] = al[i, j] * CAM[1], al[i, j] * (CAM[3] - us[i, 1])
            got = np.zeros((12, 12))
            host_harness.host_mtm(P(al), P(us), n, P(CAM), P(got))
            assert np.array_equal(cv2.mulTransposed(M, True), got)


def test_epnp_bitwise_equals_cv2_solvepnp(host_harness):
    """The whole EPnP (csrc/epnp_core.cuh solve_small) returns the SAME translation bits as cv2.solvePnP(EPNP) -- also for
    exactly 5 points, where the solution is a chaotic function of every rounding upstream of the 12x12 SVD -- and the same
    rotation up to libm's acos/sin/cos in Rodrigues (1e-13)."""
    rng = np.random.RandomState(1)
    for n in (5, 6, 7, 10, 100, 400):
        for _ in range(40):
            pw, uv = _case(rng, n, float(rng.choice([0.0, 0.5, 2.0, 10.0])))
            pw = pw.astype(np.float32).astype(np.float64)           # what solvePnPRansac hands to the minimal solver
            uv = uv.astype(np.float32).astype(np.float64)
            ok, rv, tv = cv2.solvePnP(pw.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), K_LM, None, flags=cv2.SOLVEPNP_EPNP)
            # solvePnP: undistortPoints -> normalised coordinates (double), epnp::init_points maps them back to pixels
            us = np.ascontiguousarray(((uv - CAM[2:]) * (1.0 / CAM[:2])) * CAM[:2] + CAM[2:])
            R9, t3 = np.zeros(9), np.zeros(3)
            host_harness.host_epnp_small(P(pw), P(us), n, P(CAM), P(R9), P(t3))
            assert np.array_equal(t3, tv[:, 0]), (n, t3, tv[:, 0])
            assert np.abs(R9.reshape(3, 3) - cv2.Rodrigues(rv)[0]).max() < 1e-13


def test_rodrigues_roundtrip(host_harness):
    rng = np.random.RandomState(2)
    for _ in range(50):
        rv = rng.randn(3); rv *= rng.uniform(0, np.pi) / np.linalg.norm(rv)
        R = np.ascontiguousarray(cv2.Rodrigues(rv)[0])
        r3, Ro = np.zeros(3), np.zeros(9)
        host_harness.host_rodrigues_roundtrip(P(R), P(r3), P(Ro))
        assert np.abs(r3 - cv2.Rodrigues(R)[0][:, 0]).max() < 1e-10
        assert np.abs(Ro.reshape(3, 3) - R).max() < 1e-12


def test_update_num_iters(host_harness):
    host_harness.host_update_iters.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
    assert host_harness.host_update_iters(0.99, 0.0, 5, 100) == 0
    assert host_harness.host_update_iters(0.99, 1.0, 5, 100) == 100
    assert host_harness.host_update_iters(0.99, 0.2, 5, 100) == 12       # log(.01)/log(1-.8^5) = 11.6
    assert host_harness.host_update_iters(0.99, 0.4, 5, 100) == 57


def test_ransac_emulation_matches_cv2_solvepnpransac(host_harness):
    """The full algorithm of csrc/pnp_ransac.cu (RNG replay, 5-point EPnP, float32 scoring, sequential accept/terminate
    replay, refit) emulated on the host from the same functions: inlier index sets, rvec and tvec IDENTICAL (bit for bit)
    to cv2.solvePnPRansac in every planted case."""
    import resource
    resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
    rng = np.random.RandomState(0)
    total = 0
    for trial in range(60):
        n = int(rng.choice([6, 12, 50, 200, 2000, 8000]))
        of = float(rng.choice([0, 0.2, 0.4, 0.6]))
        pw, uv = _case(rng, n, float(rng.choice([0.5, 1.0, 3.0])))
        no = int(n * of)
        idx = rng.choice(n, no, replace=False)
        uv[idx] += rng.uniform(-60, 60, (no, 2))
        ret, rvc, tvc, inl = cv2.solvePnPRansac(pw, uv.reshape(-1, 1, 2), K_LM, None, flags=cv2.SOLVEPNP_EPNP,
                                                reprojectionError=5, iterationsCount=100)
        rvec, tvec, mask, it = np.zeros(3), np.zeros(3), np.zeros(n, np.uint8), ctypes.c_int()
        r = host_harness.host_ransac(P(pw), P(uv), n, P(CAM), ctypes.c_float(5.0), 100, ctypes.c_double(0.99), P(rvec), P(tvec),
                                     mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), ctypes.byref(it))
        assert (inl is None) == (r < 0)
        if inl is None:
            continue
        total += 1
        assert np.array_equal(np.nonzero(mask)[0], inl[:, 0]), (trial, n, of)
        assert np.array_equal(rvec, rvc[:, 0]) and np.array_equal(tvec, tvc[:, 0]), (trial, n, of)
    assert total >= 40
