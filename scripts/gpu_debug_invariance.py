"""Batch-invariance / determinism probe: every intermediate tensor of crop 0, forward of (x0) vs forward of (x0, x1),
and the same forward twice.  python scripts/gpu_debug_invariance.py [backbone]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pix2pose_b200 import ae_model, weights as W

bb = sys.argv[1] if len(sys.argv) > 1 else "paper"
names = {"paper": ["f1", "f2", "f3", "f4", "enc", "d0", "d1", "d1_uni", "d2", "d2_uni", "d3", "d3_uni"],
         "resnet50": ["f1", "pool1", "act2a", "act2b", "act2c", "act3a", "act3d", "f4", "enc", "d0", "d1", "d1_uni", "d2", "d2_uni", "d3", "d3_uni"]}[bb]
m = ae_model.GeneratorModel(bb, capacity=2, precision="fp16x3")
m.load_weights(W.synthetic_weights(bb, 1))
x = np.random.RandomState(0).uniform(-1, 1, (2, 128, 128, 3)).astype(np.float32)


def run(xx):
    d, p = m.predict(xx)
    return {n: m.engine.read_tensor(n, 1).copy() for n in names}, d[0].copy()


t2, d2 = run(x)
t2b, d2b = run(x)
t1, d1 = run(x[:1])
for n in names:
    print("%-8s n=2 twice: %-5s   n=1 vs n=2: %-5s  max|diff| %.3e" % (n, np.array_equal(t2[n], t2b[n]), np.array_equal(t1[n], t2[n]),
                                                                  np.abs(t1[n] - t2[n]).max()))
print("decode   n=2 twice: %s   n=1 vs n=2: %s  max|diff| %.3e" % (np.array_equal(d2, d2b), np.array_equal(d1, d2), np.abs(d1 - d2).max()))
