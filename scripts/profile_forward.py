"""Short command for ncu: three 64-crop resnet50 forwards (2 warm-up + 1) through the C ABI."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pix2pose_b200 import ae_model, weights as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
m = ae_model.GeneratorModel("resnet50", capacity=n, precision=prec)
m.load_weights(W.synthetic_weights("resnet50", 1))
x = np.random.RandomState(0).uniform(-1, 1, (n, 128, 128, 3)).astype(np.float32)
for _ in range(3):
    m.predict(x)
print("device ms per forward:", m.time_forward(x, 2, 5))
