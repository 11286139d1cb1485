"""GPU bring-up: layer-by-layer comparison of the CUDA generator against the torch-CPU oracle.
Run on the GPU box:  python scripts/gpu_debug_net.py [backbone] [precision] [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle.net_oracle import NetOracle
from pix2pose_b200 import ae_model, weights as W


def main():
    bbs = [sys.argv[1]] if len(sys.argv) > 1 and sys.argv[1] != "all" else ["paper", "resnet50"]
    precs = [sys.argv[2]] if len(sys.argv) > 2 and sys.argv[2] != "all" else ["fp16x3", "fp16"]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    for bb in bbs:
        w = W.synthetic_weights(bb, 1)
        x = np.random.RandomState(0).uniform(-1, 1, (n, 128, 128, 3)).astype(np.float32)
        net = NetOracle(w, bb)
        net.taps = {}
        d0, p0 = net.forward(x)
        for prec in precs:
            t = time.time()
            m = ae_model.GeneratorModel(bb, capacity=max(n, 4), precision=prec)
            m.load_weights(w)
            print("[%s/%s] engine+model built in %.2fs" % (bb, prec, time.time() - t), flush=True)
            t = time.time()
            d, p = m.predict(x)
            print("[%s/%s] predict %.3fs" % (bb, prec, time.time() - t), flush=True)
            for name, ref in net.taps.items():
                try:
                    got = m.engine.read_tensor(name, n)
                except Exception as e:  # tensor not materialised by the engine
                    print("   %-8s (skipped: %s)" % (name, e))
                    continue
                err = np.abs(got - ref)
                print("   %-8s shape %-18s ref|max| %8.3f  max_err %.3e  mean_err %.3e  nan %d" % (
                    name, ref.shape[1:], np.abs(ref).max(), err.max(), err.mean(), int(np.isnan(got).sum())), flush=True)
            print("[%s/%s] decode max_err %.3e  prob max_err %.3e  (nan %d)" % (
                bb, prec, np.abs(d - d0).max(), np.abs(p - p0).max(), int(np.isnan(d).sum())), flush=True)
            del m


if __name__ == "__main__":
    main()


def timing(bb="resnet50", prec="fp16x3", n=64, reps=5):
    w = W.synthetic_weights(bb, 1)
    x = np.random.RandomState(0).uniform(-1, 1, (n, 128, 128, 3)).astype(np.float32)
    m = ae_model.GeneratorModel(bb, capacity=n, precision=prec)
    m.load_weights(w)
    m.predict(x)
    t = time.time()
    for _ in range(reps):
        m.predict(x)
    dt = (time.time() - t) / reps
    ms = m.time_forward(x, 3, 10)
    print("[timing %s/%s] n=%d  %.2f ms/batch (host in/out)  %.0f crops/s | device %.3f ms  %.0f crops/s" % (
        bb, prec, n, dt * 1e3, n / dt, ms, n / ms * 1e3), flush=True)


if __name__ == "__main__" and os.environ.get("P2P_TIMING"):
    for bb in ("resnet50", "paper"):
        for prec in ("fp16x3", "fp16"):
            timing(bb, prec, 64)
            timing(bb, prec, 256)
