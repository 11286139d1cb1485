"""Determinism stress: the same batch through the generator `reps` times (and through alternating batch sizes), every
output compared bit for bit.  python scripts/gpu_stress_determinism.py [backbone] [n] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pix2pose_b200 import ae_model, weights as W

bb = sys.argv[1] if len(sys.argv) > 1 else "paper"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 2
m = ae_model.GeneratorModel(bb, capacity=cap, precision="fp16x3")
m.load_weights(W.synthetic_weights(bb, 1))
x = np.random.RandomState(0).uniform(-1, 1, (n, 128, 128, 3)).astype(np.float32)
d0, p0 = m.predict(x)
bad = 0
for r in range(reps):
    d, p = m.predict(x)
    if not (np.array_equal(d, d0) and np.array_equal(p, p0)):
        bad += 1
        w = np.argwhere(d != d0)
        print("rep %d: %d decode values differ, crops %s, max|diff| %.3e" % (r, len(w), sorted(set(w[:, 0].tolist())), np.abs(d - d0).max()))
    for i in (0, n - 1):
        d1, p1 = m.predict(x[i:i + 1])
        if not (np.array_equal(d1[0], d0[i]) and np.array_equal(p1[0], p0[i])):
            bad += 1
            print("rep %d: single crop %d differs from its batch result, max|diff| %.3e (%d values)" % (
                r, i, np.abs(d1[0] - d0[i]).max(), int((d1[0] != d0[i]).sum())))
print("%s n=%d cap=%d reps=%d: %d mismatches" % (bb, n, cap, reps, bad))
