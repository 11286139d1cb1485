"""Generates tests/golden/depth_golden.npz by running the REFERENCE's own functions
(/root/reference/pix2pose_util/common_util.py: getXYZ :13-30, get_normal :32-90) on small seeded depth maps.
Run in the build container, where /root/reference exists:  python tests/golden/make_depth_golden.py
The module imports only numpy / cv2 / scipy, so it runs here unmodified; ``np.float`` / ``np.int`` (removed in
numpy >= 1.24, used at common_util.py:9-11, :47) are aliased to the builtins for the duration of the call."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
np.float, np.int = float, int                      # noqa: the aliases numpy < 1.24 provided
from pix2pose_util import common_util as ref       # noqa: E402

K = dict(fx=572.4114, fy=573.57043, cx=23.3, cy=19.6)


def depth_case(seed, holes):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:40, 0:48]
    d = 700.0 + 40.0 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + 0.5 * (xx - 24) + rng.normal(0, 0.3, (40, 48))
    if holes:
        d[rng.rand(40, 48) < 0.03] = 0.0
        d[10:14, 20:25] = 0.0
        d[3, 4] = np.nan
    return d


def main():
    out = {}
    box = np.array([6, 8, 30, 40])
    d0, d1 = depth_case(0, False), depth_case(1, True)
    out["depth_plain"], out["depth_holes"], out["bbox"] = d0, d1, box
    out["K"] = np.array([K["fx"], K["fy"], K["cx"], K["cy"]])
    out["xyz_full"] = ref.getXYZ(d0, K["fx"], K["fy"], K["cx"], K["cy"])
    out["xyz_box"] = ref.getXYZ(d0, K["fx"], K["fy"], K["cx"], K["cy"], box)
    out["normal_full"] = ref.get_normal(d0, bbox=np.array([0]), refine=False, **K)
    out["normal_box"] = ref.get_normal(d0, bbox=box, refine=False, **K)
    out["normal_refined_full"] = ref.get_normal(d1.copy(), bbox=np.array([0]), refine=True, **K)
    out["normal_refined_box"] = ref.get_normal(d1.copy(), bbox=box, refine=True, **K)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "depth_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
