"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference has no tests or golden vectors and cannot be
imported here, so these fixtures pin the *oracle* (torch-CPU fp32 restatement,
seeded synthetic weights/inputs) -- they guard against regressions of the oracle and give the GPU
tests a file-based expectation that does not need torch at test time."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.net_oracle import NetOracle  # noqa: E402
from oracle.resize_oracle import resize  # noqa: E402
from pix2pose_b200 import weights as W  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    x = np.random.RandomState(0).uniform(-1, 1, (2, 128, 128, 3)).astype(np.float32)
    for bb in ("resnet50", "paper"):
        net = NetOracle(W.synthetic_weights(bb, 1), bb)
        net.taps = {}
        d, p = net.forward(x)
        out[bb + "_decode_s8"] = d[:, ::8, ::8, :]          # every 8th pixel
        out[bb + "_prob_s8"] = p[:, ::8, ::8, :]
        out[bb + "_f4_mean"] = np.array([net.taps["f4"].mean(), net.taps["f4"].std()])
        out[bb + "_d3uni_s16"] = net.taps["d3_uni"][:, ::16, ::16, ::16]
    np.savez_compressed(os.path.join(HERE, "net_golden.npz"), **out)
    a = np.random.RandomState(3).rand(128, 128)
    np.savez_compressed(os.path.join(HERE, "resize_golden.npz"),
                        up_reflect=resize(a, (200, 173), mode="reflect")[::10, ::10],
                        up_const=resize(a, (200, 173), mode="constant", cval=1.0)[::10, ::10],
                        down_const=resize(a, (77, 91), mode="constant", cval=0.5)[::7, ::7],
                        mask_up=(resize(a > 0.5, (150, 150), mode="constant", cval=0) > 0.9)[::5, ::5])
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
