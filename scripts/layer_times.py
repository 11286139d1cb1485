"""Per-launch CUDA-event times of one forward (device resident), plus the plain forward time.
  P2P_PROF_LAYERS=1 python scripts/layer_times.py [n=256] [precision=fp16x3] [backbone=resnet50]
Prints (stderr, from the engine) `launch:microseconds` for every launch, then the event-timed forward."""
import ctypes
import os
import sys

os.environ.setdefault("P2P_PROF_LAYERS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pix2pose_b200 import _lib, ae_model, weights as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
bb = sys.argv[3] if len(sys.argv) > 3 else "resnet50"
m = ae_model.GeneratorModel(bb, capacity=n, precision=prec)
m.load_weights(W.synthetic_weights(bb, 1))
x = np.random.RandomState(0).uniform(-1, 1, (n, 128, 128, 3)).astype(np.float32)
ms = (ctypes.c_double * 2)()
cnt = (ctypes.c_int * 2)()
for _ in range(2):
    _lib.check(_lib.lib().p2p_engine_profile_forward(m.engine.handle, m._model, _lib.fptr(x), n, ms, cnt))
sys.stderr.flush()
print("conv launches %.3f ms (%d), other %.3f ms (%d)" % (ms[0], cnt[0], ms[1], cnt[1]))
print("device ms per forward: %.3f" % m.time_forward(x, 3, 10))
