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


def test_epnp_matches_cv2_solvepnp(host_harness):
    """n >= 6: identical to cv2.solvePnP(EPNP) to 1e-8 (incl. OpenCV's JacobiSVD sign convention for the
    control points); n = 5 noise-free: identical; n = 5 noisy: equal up to the null-space basis noise."""
    rng = np.random.RandomState(1)
    for n in (6, 7, 10, 100, 400):
        for _ in range(25):
            pw, uv = _case(rng, n, float(rng.choice([0.5, 2.0, 10.0])))
            ok, rv, tv = cv2.solvePnP(pw.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), K_LM, None, flags=cv2.SOLVEPNP_EPNP)
            R9, t3 = np.zeros(9), np.zeros(3)
            host_harness.host_epnp_small(P(pw), P(uv), n, P(CAM), P(R9), P(t3))
            assert np.abs(R9.reshape(3, 3) - cv2.Rodrigues(rv)[0]).max() < 1e-8
            assert np.abs(t3 - tv[:, 0]).max() < 1e-6
    for _ in range(25):
        pw, uv = _case(rng, 5, 0.0)
        ok, rv, tv = cv2.solvePnP(pw.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), K_LM, None, flags=cv2.SOLVEPNP_EPNP)
        R9, t3 = np.zeros(9), np.zeros(3)
        host_harness.host_epnp_small(P(pw), P(uv), 5, P(CAM), P(R9), P(t3))
        assert np.abs(R9.reshape(3, 3) - cv2.Rodrigues(rv)[0]).max() < 1e-8


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
    """The full algorithm of csrc/pnp_ransac.cu (RNG replay, 5-point EPnP, float32 scoring, sequential
    accept/terminate replay, refit) emulated on the host from the same functions: identical inlier sets to
    cv2.solvePnPRansac in the large majority of planted cases, tolerance otherwise."""
    import resource
    resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
    rng = np.random.RandomState(0)
    exact = total = 0
    for trial in range(24):
        n = int(rng.choice([6, 12, 50, 200, 2000, 8000]))
        of = float(rng.choice([0, 0.2, 0.4]))
        pw, uv = _case(rng, n, 1.0)
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
        same = np.array_equal(np.nonzero(mask)[0], inl[:, 0])
        exact += same
        if same:   # same consensus set -> same refit (up to the null-space noise of very small sets)
            assert np.abs(rvec - rvc[:, 0]).max() < 1e-5 and np.linalg.norm(tvec - tvc[:, 0]) / np.linalg.norm(tvc) < 1e-5
        elif n >= 100:
            assert abs(int(mask.sum()) - len(inl)) <= 0.05 * n and np.abs(rvec - rvc[:, 0]).max() < 2e-2
    assert exact >= 0.6 * total, (exact, total)
