"""CPU: the host logic of bench.py (argument handling, timing loop, JSON line, N = 1 / N > 1 differences) with the device
layer replaced by stand-ins -- no GPU needed.  Guards the contract: exactly ONE JSON line on stdout carrying metric / value /
e2e / roofline / clocks / gpu_launches, cpu_baseline only at N = 1."""
import ctypes
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
import ctypes, json, os, sys, types
import numpy as np
sys.path.insert(0, %(root)r)
world = int(os.environ.get("FAKE_WORLD", "1"))

class FakeLib:
    def p2p_engine_event_record(self, eng, i): return 0
    def p2p_engine_event_elapsed(self, eng, a, b, ms):
        ms._obj.value = 40.0 * 2
        return 0
    def p2p_engine_profile_forward(self, eng, model, x, n, msk, cnt):
        msk[0], msk[1], cnt[0], cnt[1] = 7.0, 0.1, 32, 3
        return 0
    def p2p_flops_per_crop(self, bb): return 10.7e9
    def p2p_engine_prof_begin(self, eng): return 0
    def p2p_engine_prof_end(self, eng, msk, cnt):
        msk[0], msk[1], cnt[0], cnt[1] = 56.0, 0.8, 256, 24      # two profiled steps
        return 0

import pix2pose_b200._lib as _lib
_lib.lib = lambda: FakeLib()
_lib.check = lambda rc: None
_lib.fptr = lambda a: None
_lib.pinned_array = lambda shape, dtype: np.zeros(shape, dtype)

import pix2pose_b200.dist as D
D.init = lambda: (0, 0, world)
D.barrier = lambda: None
D.max_over_ranks = lambda v: v
D.gather_records = lambda rec, idx, n: rec
D.shutdown = lambda: None
class FakeGather:
    def __init__(self, cap, n_total): self.n, self.last = n_total, None
    def submit(self, rec, idx): self.last = np.zeros((self.n, 16))
    def result(self): return self.last
D.AsyncGather = FakeGather

class FakeRes:
    def __init__(self, n):
        self.n_cand, self.status = np.full(n, 3), np.ones(n, int)
    def records(self): return np.zeros((len(self.status), 16))

class FakeRec:
    launch_count = 0
    def __init__(self, *a, **k):
        self.generator_train = types.SimpleNamespace(engine=types.SimpleNamespace(handle=None), _model=None)
    def upload_frames(self, frames, n): return object()
    def last_forward_ms(self): return 28.0
    def est_pose_batch(self, frames, rois, fids, frames_dev=None):
        FakeRec.launch_count += 150
        return FakeRes(len(rois))

import pix2pose_b200.recognition as R
R.pix2pose = FakeRec
import pix2pose_b200.stream as S
class FakeBatcher:
    launch_count = 0
    def __init__(self, owner, n): self.n, self.k = n, 0
    def upload_frames(self, frames): return (None, 1, 480, 640, 0)
    def submit(self, fdev, rois, fids, oids=None):
        FakeBatcher.launch_count += 150
        self.k += 1
        return (self.k %% 2, len(rois))
    def result(self, t):
        self.last_n_cand = np.full(t[1], 3)
        return np.zeros((t[1], 16)), np.ones(t[1], int)
    def last_forward_ms(self, slot): return 28.0
    def close(self): pass
S.AsyncBatcher = FakeBatcher
import pix2pose_b200.weights as W
W.synthetic_weights = lambda *a, **k: {}

import bench
bench.ClockSampler.run = lambda self: None
class FakeOra:
    def est_pose(self, frame, roi): return (None, np.zeros(1), None, None, 0.5, None)
bench.make_cpu_port = lambda cfg=3: FakeOra()
sys.argv = ["bench.py", "--gpus", str(world), "--steps", "2", "--warmup", "3", "--cpu-sample", "4"]
print("library banner that must not reach stdout")
os.write(1, b"raw fd-1 write before main\n")
bench.main()
'''


@pytest.mark.parametrize("world", [1, 2])
def test_bench_emits_exactly_one_json_line_with_the_contract_keys(world, tmp_path):
    script = tmp_path / "drive.py"
    script.write_text(DRIVER % {"root": ROOT})
    env = dict(os.environ, FAKE_WORLD=str(world), OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.split("\n") if l.strip()]
    json_lines = [l for l in lines if l.startswith("{")]
    assert len(json_lines) == 1, p.stdout
    assert lines[-1] == json_lines[0]                       # whatever ran before main() may print; after quiet_stdout() only the JSON line
    d = json.loads(json_lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == world and d["steps"] == 2 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["ms_per_step"] == pytest.approx(40.0) and d["value"] == pytest.approx(256 * world / 0.040)
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "step_frac"} and "workload" in d["config"]
    assert d["config"]["workload"].startswith("config3")          # the default stays the configuration the metric is quoted on
    assert d["roofline"]["achieved"] == pytest.approx(10.7e9 * 1024 / 28e-3 / 1e12)
    assert d["roofline"]["step_frac"] < d["roofline"]["frac"]
    assert d["gpu_launches"] == 300
    if world == 1:
        assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["cores"] >= 1
    else:
        assert "cpu_baseline" not in d
