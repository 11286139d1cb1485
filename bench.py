#!/usr/bin/env python
"""bench.py -- throughput of the Pix2Pose per-detection hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W             # this repo (CUDA), N = 1 or under torchrun
  python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port), rank 0 only

Workload (the configuration the metric "crops/sec end-to-end incl. PnP-RANSAC"
is quoted on): per GPU 256 detections of one object on 16 synthetic 480x640 frames (16 ROIs each),
resnet50 backbone, seeded synthetic weights, th_outlier [0.15,0.25,0.35], th_inlier 0.15, full
two-stage est_pose (1 + <=3 network forwards and <=3 PnP-RANSACs per detection).  One "step" = one pass
of that path over the 256 detections.  ``value`` = detections ("crops" handed to est_pose) per second
with the frames already in HBM; ``e2e`` = the same through recognition.pix2pose.est_pose_batch with
host frames (H2D of the frames and D2H of the pose records inside the timed region).  Weak scaling:
every rank owns 256 detections; one NCCL all_gather of 16-double pose records per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

K_LM = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])
OBJ = np.array([50., 40., 60., 0., 0., 0.])
TH_O, TH_I = [0.15, 0.25, 0.35], 0.15
N_FRAMES, ROIS_PER_FRAME, H, W = 16, 16, 480, 640
BACKBONE = "resnet50"
METRIC = "crops/sec (128x128) end-to-end incl. PnP-RANSAC"


def workload(rank=0):
    rng = np.random.RandomState(100 + rank)
    frames = rng.randint(0, 256, (N_FRAMES, H, W, 3)).astype(np.uint8)
    rois, fids = [], []
    for f in range(N_FRAMES):
        for _ in range(ROIS_PER_FRAME):
            cy, cx = rng.randint(80, H - 80), rng.randint(80, W - 80)
            h, w = rng.randint(50, 130), rng.randint(50, 130)
            rois.append([cy - h // 2, cx - w // 2, cy + h // 2, cx + w // 2])
            fids.append(f)
    return frames, np.array(rois), np.array(fids)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1])); mx.append(float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def conv_traffic(capacity):
    """DRAM bytes (read + write) of the conv_tc launches of ONE forward of `capacity` crops, from the committed ncu capture
    (profiles/conv_traffic.json, written from scripts/ncu_step_metrics.py output); None when no capture matches."""
    p = os.path.join(ROOT, "profiles", "conv_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        if int(d.get("capacity", -1)) == int(capacity):
            return float(d["dram_bytes_per_forward"])
    except (OSError, ValueError, KeyError):
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1391.7), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


# ---------------------------------------------------------------------------------------------------
def make_cpu_port():
    """The reference's CPU path restated (oracle/): torch-CPU generator + numpy resize + real cv2 PnP."""
    from oracle.net_oracle import NetOracle
    from oracle.recognition_oracle import Pix2PoseOracle
    from pix2pose_b200 import weights as Wt
    net = NetOracle(Wt.synthetic_weights(BACKBONE, 1), BACKBONE)
    return Pix2PoseOracle(net, K_LM, W, H, OBJ, th_outlier=TH_O, th_inlier=TH_I)


def cpu_sample(ora, frames, rois, fids, idx):
    """Times est_pose of the CPU port over detections `idx`; returns (detections/s, seconds, poses found)."""
    t = time.perf_counter()
    ok = 0
    for i in idx:
        out = ora.est_pose(frames[fids[i]], rois[i])
        ok += not isinstance(out[1], int)
    dt = time.perf_counter() - t
    return len(idx) / dt, dt, ok


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and may use the whole host."""
    import cv2
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    cv2.setNumThreads(n)


def cpu_threads():
    import cv2
    import torch
    return torch.get_num_threads(), cv2.getNumThreads()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    frames, rois, fids = workload(0)
    use_all_host_threads()
    ora = make_cpu_port()
    per_step = 6
    for s in range(max(args.warmup, 1)):
        cpu_sample(ora, frames, rois, fids, [s % len(rois)])
    t = time.perf_counter()
    n = 0
    for s in range(args.steps):
        b = (s * per_step) % (len(rois) - per_step)
        cpu_sample(ora, frames, rois, fids, range(b, b + per_step))
        n += per_step
    dt = time.perf_counter() - t
    cores, cvt = cpu_threads()
    val = n / dt
    emit(({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config3: two-stage est_pose, resnet50, 480x640 frames", "sample": "%d detections per step" % per_step},
        "cpu_baseline": {"value": val, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": "%d detections per step x %d steps of the bench workload; torch-CPU fp32 generator (%d threads; stand-in "
                                   "for Keras/TF-CPU, which cannot be installed here) + numpy resize + real cv2.solvePnPRansac (%d threads)"
                                   % (per_step, args.steps, cores, cvt)},
        "e2e": {"value": val, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes
    from pix2pose_b200 import _lib, dist as D, weights as Wt
    from pix2pose_b200.recognition import _Det, _Pose, pix2pose
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line (NCCL prints its version banner there)
    rank, local_rank, world = D.init()
    frames, rois, fids = workload(rank)
    n_det = len(rois)
    w = Wt.synthetic_weights(BACKBONE, 1)
    rec = pix2pose(w, K_LM, W, H, OBJ, th_outlier=TH_O, th_inlier=TH_I, backbone=BACKBONE, precision=args.precision,
                   capacity=args.capacity, max_dets=n_det)
    L = _lib.lib()
    eng = rec.generator_train.engine.handle
    n_total = n_det * world

    def gather(res):
        return D.gather_records(res.records(), rank * n_det + np.arange(n_det), n_total)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        D.barrier()
        _lib.check(L.p2p_engine_event_record(eng, 0))
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        _lib.check(L.p2p_engine_event_record(eng, 1))
        ms = ctypes.c_float()
        _lib.check(L.p2p_engine_event_elapsed(eng, 0, 1, ctypes.byref(ms)))     # device time on the launching stream
        wall = (time.perf_counter() - t0) * 1e3
        D.barrier()
        return D.max_over_ranks(ms.value), D.max_over_ranks(wall)

    # ---- value: frames resident in HBM
    fdev = rec.upload_frames(frames, n_det)
    state = {}

    def step_dev():
        state["res"] = rec.est_pose_batch(None, rois, fids, frames_dev=fdev)
        state["all"] = gather(state["res"])

    if args.profile:                      # short run for `ncu` launch lists: 1 warm-up + 1 step, no JSON
        step_dev()
        step_dev()
        return
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = rec.launch_count
    ms_dev, wall_dev = timed(step_dev, args.steps, max(args.warmup, 3))
    launches = (rec.launch_count - l0) // (args.steps + max(args.warmup, 3)) * args.steps
    clk = clocks.stop()
    res = state["res"]
    n_cand_mean = float(np.mean(res.n_cand))
    ok_frac = float(np.mean(res.status == 1))

    # ---- e2e: host frames (pinned) through the public API, H2D + D2H inside
    pinned = _lib.pinned_array(frames.shape, np.uint8)
    pinned[...] = frames

    def step_e2e():
        state["res"] = rec.est_pose_batch(pinned, rois, fids)
        state["all"] = gather(state["res"])

    ms_e2e, _ = timed(step_e2e, args.steps, 3)
    h2d = int(frames.nbytes + n_det * ctypes.sizeof(_Det) + 64)
    d2h = int(n_det * ctypes.sizeof(_Pose))

    if rank != 0:
        D.shutdown()
        return
    # ---- roofline of the dominant kernel class (tcgen05 implicit-GEMM conv), measured live
    x = np.random.RandomState(0).uniform(-1, 1, (args.capacity, 128, 128, 3)).astype(np.float32)
    msk = (ctypes.c_double * 2)()
    cnt = (ctypes.c_int * 2)()
    _lib.check(L.p2p_engine_profile_forward(eng, rec.generator_train._model, _lib.fptr(x), args.capacity, msk, cnt))
    flops_crop = L.p2p_flops_per_crop(BACKBONE.encode())
    achieved = flops_crop * args.capacity / (msk[0] * 1e-3) / 1e12
    peak, peak_src = measured_peaks()
    # ---- CPU baseline: bounded sample of the same workload
    cpu_baseline = None                                            # reported on rank 0 at N = 1 only
    if world == 1:
        use_all_host_threads()
        ora = make_cpu_port()
        cpu_sample(ora, frames, rois, fids, [0])                  # warm-up (thread pools)
        cpu_val, cpu_dt, cpu_ok = cpu_sample(ora, frames, rois, fids, range(args.cpu_sample))
        cores, cvthreads = cpu_threads()
        cpu_baseline = {"value": cpu_val, "unit": "crops/s", "cores": cores, "kind": "port",
                        "sample": "first %d detections of the same workload (%.1f s): torch-CPU fp32 generator (%d threads; stand-in for "
                                  "Keras/TF-CPU) + numpy resize + real cv2.solvePnPRansac (%d threads)" % (args.cpu_sample, cpu_dt, cores, cvthreads)}
    per_step_ms = ms_dev / args.steps
    value = n_total / (per_step_ms * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 hi/lo operand pairs, f32 accumulate (fp16x3)" if args.precision == "fp16x3" else "f16, f32 accumulate",
        "data": "synthetic",
        "config": {"workload": "config3: 256 detections/GPU on 16 synthetic 480x640 frames, 1 object, two-stage est_pose incl. EPnP-RANSAC",
                   "backbone": BACKBONE, "global_detections": n_total, "stage2_candidates_per_detection": n_cand_mean,
                   "network_crops_per_s": value * (1 + n_cand_mean), "pose_found_fraction": ok_frac,
                   "parallelism": "dp%d (detections sharded, 1 all_gather of pose records per step)" % world,
                   "precision": args.precision, "engine_capacity": args.capacity,
                   "l2": "per-step working set (>= 4 GB of activations + 112 MB weights) exceeds the 126 MB L2; no flush needed",
                   "wall_ms_per_step": wall_dev / args.steps},
        "clocks": clk,
        "e2e": {"value": n_total / (ms_e2e / args.steps * 1e-3), "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": conv_traffic(args.capacity), "peak_source": peak_src,
                     "kernel": "conv_tc_{pair,slab,persistent}_kernel (all %d tcgen05 conv launches of one %d-crop forward: %.3f ms; other kernels %.3f ms)" % (
                         cnt[0], args.capacity, msk[0], msk[1]),
                     "note": "algorithmic FLOPs (10.70 GFLOP/crop); fp16x3 issues 3 MMAs per k-step, so tensor-pipe work is 3x this"},
    }
    if cpu_baseline is not None:
        out["cpu_baseline"] = cpu_baseline
    emit(out)
    D.shutdown()


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (the NCCL version banner, torchrun notices) write to file
    descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON line goes to the
    saved original descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp16"])
    ap.add_argument("--capacity", type=int, default=256)
    ap.add_argument("--cpu-sample", type=int, default=96)    # ~15 s of CPU work on 16 cores
    ap.add_argument("--profile", action="store_true", help="one warm-up + one step only (for ncu launch lists)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
