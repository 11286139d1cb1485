"""Replays the order of tests/test_net_gpu.py in one process, then probes every tensor of the batch-invariance case."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pix2pose_b200 import ae_model, weights as W

inputs = np.random.RandomState(0).uniform(-1, 1, (5, 128, 128, 3)).astype(np.float32)
pre = sys.argv[1] if len(sys.argv) > 1 else "all"
if pre in ("all", "x3"):
    for bb in ("resnet50", "paper"):
        m = ae_model.GeneratorModel(bb, capacity=8, precision="fp16x3")
        m.load_weights(W.synthetic_weights(bb, 1))
        m.predict(inputs)
        del m
if pre in ("all", "fp16"):
    for bb in ("resnet50", "paper"):
        m = ae_model.GeneratorModel(bb, capacity=4, precision="fp16")
        m.load_weights(W.synthetic_weights(bb, 1))
        m.predict(inputs[:2])
        del m
names = ["f1", "f2", "f3", "f4", "enc", "d0", "d1", "d1_uni", "d2", "d2_uni", "d3", "d3_uni"]
m = ae_model.GeneratorModel("paper", capacity=2, precision="fp16x3")
m.load_weights(W.synthetic_weights("paper", 1))
m.predict(np.zeros((0, 128, 128, 3), np.float32))
d5, p5 = m.predict(inputs)
ta = {n: m.engine.read_tensor(n, 1).copy() for n in names}   # crop 4 (last chunk, n = 1)
for i in (0, 4):
    d1, p1 = m.predict(inputs[i:i + 1])
    print("crop %d: decode equal %s  max|diff| %.3e" % (i, np.array_equal(d1[0], d5[i]), np.abs(d1[0] - d5[i]).max()))
    if i == 4:
        for n in names:
            t = m.engine.read_tensor(n, 1)
            print("   %-8s equal %-5s max|diff| %.3e" % (n, np.array_equal(t, ta[n]), np.abs(t - ta[n]).max()))
d2, _ = m.predict(inputs[:2])
tb = {n: m.engine.read_tensor(n, 1).copy() for n in names}
d1, _ = m.predict(inputs[:1])
print("crop 0 of (x0,x1) vs (x0): decode equal %s" % np.array_equal(d1[0], d2[0]))
for n in names:
    t = m.engine.read_tensor(n, 1)
    print("   %-8s equal %-5s max|diff| %.3e" % (n, np.array_equal(t, tb[n]), np.abs(t - tb[n]).max()))
