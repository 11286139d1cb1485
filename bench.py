#!/usr/bin/env python
"""bench.py -- throughput of the Pix2Pose per-detection hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--config C]     # this repo (CUDA), N = 1 or under torchrun
  python bench.py --impl reference --steps K --warmup W [--config C]   # the reference's CPU path (oracle port), rank 0 only

Workloads = BASELINE.json configs; ``--config 3`` is the default and the configuration the metric
"crops/sec end-to-end incl. PnP-RANSAC" is quoted on:

  2  batch of 64 synthetic crops, one object, encoder-decoder only (no PnP); unit = network crops/s
  3  per GPU 256 detections of one object on 16 synthetic 480x640 frames (16 ROIs each), resnet50, th_outlier
     [0.15,0.25,0.35], th_inlier 0.15, full two-stage est_pose (1 + <=3 forwards and <=3 PnP-RANSACs per detection); weak
  4  LM-O: 8 objects (8 resident weight sets), 512 detections in detector (round-robin) order on 32 frames of 640x480,
     sharded over the GPUs by object group (strong scaling: the 512 are fixed), one device run per rank
  5  T-LESS: 30 objects (30 resident weight sets), 720x540 frames, 250 detections per GPU (weak: 250 g on g GPUs)

One "step" = one pass of the path over the rank's detections.  ``value`` = detections ("crops" handed to est_pose; config 2:
network crops) per second with the frames already in HBM; ``e2e`` = the same through the public Python API with pinned
host frames (the H2D copy of every step's frames and the D2H of its pose records inside the timed region).  Multi-GPU:
detections sharded, no data-path collective, one NCCL all_gather of 16-double pose records per step, consumed one step
later (pix2pose_b200.dist.AsyncGather).
"""
import argparse
import hashlib
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
K_TLESS = np.array([[1075.65, 0, 360.0], [0, 1073.90, 270.0], [0, 0, 1]])
OBJ = np.array([50., 40., 60., 0., 0., 0.])
TH_O, TH_I = [0.15, 0.25, 0.35], 0.15
BACKBONE = "resnet50"
METRIC = "crops/sec (128x128) end-to-end incl. PnP-RANSAC"
LMO_IDS = [1, 5, 6, 8, 9, 10, 11, 12]

CONFIGS = {
    2: dict(name="config2: batch of 64 synthetic crops, 1 object, encoder-decoder only (no PnP)", kind="net", batch=64),
    3: dict(name="config3: 256 detections/GPU on 16 synthetic 480x640 frames, 1 object, two-stage est_pose incl. EPnP-RANSAC",
            kind="single", H=480, W=640, K=K_LM, frames=16, per_gpu=256, scaling="weak"),
    4: dict(name="config4: LM-O stream, 8 objects (8 resident weight sets), 512 detections on 32 synthetic 640x480 frames in "
                 "detector order, sharded by object group, two-stage est_pose incl. EPnP-RANSAC",
            kind="multi", H=480, W=640, K=K_LM, frames=32, objects=LMO_IDS, total=512, scaling="strong"),
    5: dict(name="config5: T-LESS stream, 30 objects (30 resident weight sets), 250 detections/GPU on synthetic 720x540 frames, "
                 "ICP off, two-stage est_pose incl. EPnP-RANSAC",
            kind="multi", H=540, W=720, K=K_TLESS, frames=25, objects=list(range(1, 31)), per_gpu=250, scaling="weak"),
}


def obj_param(oid):
    """Synthetic norm_factor entry of object `oid` (x/y/z scale and centre in mm, tools/bop_io.py:33-42)."""
    r = np.random.RandomState(1000 + oid)
    return np.concatenate([r.uniform(30, 80, 3), r.uniform(-5, 5, 3)])


def make_rois(rng, n_frames, n, H, W):
    rois, fids = [], []
    for i in range(n):
        cy, cx = rng.randint(80, H - 80), rng.randint(80, W - 80)
        h, w = rng.randint(50, 130), rng.randint(50, 130)
        rois.append([cy - h // 2, cx - w // 2, cy + h // 2, cx + w // 2])
        fids.append(i * n_frames // n)
    return np.array(rois), np.array(fids)


def workload(cfg_id, rank=0, world=1):
    """(frames, rois, frame ids, object ids or None, global indices of this rank's detections, global detection count)."""
    c = CONFIGS[cfg_id]
    H, W = c["H"], c["W"]
    if c["kind"] == "single":
        rng = np.random.RandomState(100 + rank)
        frames = rng.randint(0, 256, (c["frames"], H, W, 3)).astype(np.uint8)
        rois, fids = make_rois(rng, c["frames"], c["per_gpu"], H, W)
        n = c["per_gpu"]
        return frames, rois, fids, None, rank * n + np.arange(n), n * world
    from pix2pose_b200.dist import shard_by_object
    total = c["total"] if c["scaling"] == "strong" else c["per_gpu"] * world
    rng = np.random.RandomState(200 + cfg_id)
    n_frames = c["frames"] * (world if c["scaling"] == "weak" else 1)
    rois, fids = make_rois(rng, n_frames, total, H, W)
    oids = np.array([c["objects"][i % len(c["objects"])] for i in range(total)])        # the order a detector emits: interleaved
    mine = shard_by_object(oids, rank, world)
    used = np.unique(fids[mine])                                                          # a rank only needs the frames its ROIs sit on
    remap = {f: i for i, f in enumerate(used)}
    frames = np.random.RandomState(300 + rank).randint(0, 256, (len(used), H, W, 3)).astype(np.uint8)
    return frames, rois[mine], np.array([remap[f] for f in fids[mine]]), oids[mine], mine, total


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


def conv_sources_sha():
    """Fingerprint of the sources the generator kernels are built from; profiles/conv_traffic.json records the one its ncu
    capture was taken with."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "pix2pose_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith(("conv_tc", "engine", "net_kernels", "sm100_ptx")) and f.endswith((".cu", ".cuh")):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def conv_traffic(capacity):
    """(DRAM bytes read + written by the conv_tc launches of ONE forward of `capacity` crops, note) from the committed
    `ncu --set full` capture (profiles/conv_traffic.json, written by scripts/ncu_step_metrics.py).  The number is refused
    (None) when the capture was taken at another batch size or with other kernel sources than the ones built now."""
    p = os.path.join(ROOT, "profiles", "conv_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
    except (OSError, ValueError):
        return None, "no ncu capture committed"
    if int(d.get("capacity", -1)) != int(capacity):
        return None, "ncu capture is for capacity %s" % d.get("capacity")
    if d.get("csrc_sha") != conv_sources_sha():
        return None, "stale: ncu capture %s predates the current kernel sources" % d.get("source", "?")
    return float(d["dram_bytes_per_forward"]), "ncu --set full, %s" % d.get("source", "?")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1391.7), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernels timed inside a long step)"
    return 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


# ---------------------------------------------------------------------------------------------------
class CpuPort:
    """The reference's CPU path restated (oracle/): torch-CPU generator + numpy resize + real cv2 PnP, one per object like
    tools/5_evaluation_bop_basic.py:206-225 keeps one Keras model per object (built lazily: only sampled objects)."""

    def __init__(self, cfg_id):
        self.c = CONFIGS[cfg_id]
        self.by_obj = {}

    def get(self, oid):
        if oid not in self.by_obj:
            from oracle.net_oracle import NetOracle
            from oracle.recognition_oracle import Pix2PoseOracle
            from pix2pose_b200 import weights as Wt
            single = self.c["kind"] == "single"
            net = NetOracle(Wt.synthetic_weights(BACKBONE, 1 if single else int(oid)), BACKBONE)
            self.by_obj[oid] = Pix2PoseOracle(net, self.c["K"], self.c["W"], self.c["H"], OBJ if single else obj_param(int(oid)),
                                              th_outlier=TH_O, th_inlier=TH_I)
        return self.by_obj[oid]

    def est_pose(self, frame, roi, oid=0):
        return self.get(oid).est_pose(frame, roi)


def make_cpu_port(cfg_id=3):
    return CpuPort(cfg_id)


def cpu_sample(ora, frames, rois, fids, idx, oids=None):
    """Times est_pose of the CPU port over detections `idx`; returns (detections/s, seconds, poses found)."""
    for i in idx:
        if oids is not None and hasattr(ora, "get"):
            ora.get(oids[i])                                      # model construction is not part of the per-detection path
    t = time.perf_counter()
    ok = 0
    for i in idx:
        out = ora.est_pose(frames[fids[i]], rois[i], oids[i]) if oids is not None else ora.est_pose(frames[fids[i]], rois[i])
        ok += not isinstance(out[1], int)
    dt = time.perf_counter() - t
    return len(idx) / dt, dt, ok


def cpu_net_sample(n_crops):
    """Config 2 on the CPU: the torch-CPU generator on `n_crops` crops, batch 64 (crops/s, seconds)."""
    from oracle.net_oracle import NetOracle
    from pix2pose_b200 import weights as Wt
    net = NetOracle(Wt.synthetic_weights(BACKBONE, 1), BACKBONE)
    x = np.random.RandomState(0).uniform(-1, 1, (n_crops, 128, 128, 3)).astype(np.float32)
    net.forward(x[:2])
    t = time.perf_counter()
    net.forward(x)
    dt = time.perf_counter() - t
    return n_crops / dt, dt


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


CPU_NOTE = ("torch-CPU fp32 generator (%d threads; stand-in for Keras/TF-CPU, which cannot be installed here) + numpy resize + real "
            "cv2.solvePnPRansac (%d threads)")


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    use_all_host_threads()
    cores, cvt = 1, 1
    c = CONFIGS[args.config]
    if c["kind"] == "net":
        cores, cvt = cpu_threads()
        per_step = 16
        cpu_net_sample(2)
        t = time.perf_counter()
        for _ in range(args.steps):
            cpu_net_sample(per_step)
        dt = time.perf_counter() - t
        n, sample = per_step * args.steps, "%d crops per step (generator only)" % per_step
    else:
        frames, rois, fids, oids, _, _ = workload(args.config, 0, 1)
        ora = make_cpu_port(args.config)
        per_step = 6
        for s in range(max(args.warmup, 1)):
            cpu_sample(ora, frames, rois, fids, [s % len(rois)], oids)
        idxs = [range((s * per_step) % (len(rois) - per_step), (s * per_step) % (len(rois) - per_step) + per_step) for s in range(args.steps)]
        if oids is not None:
            for r in idxs:
                for i in r:
                    ora.get(oids[i])
        t = time.perf_counter()
        n = 0
        for r in idxs:
            cpu_sample(ora, frames, rois, fids, r, oids)
            n += per_step
        dt = time.perf_counter() - t
        cores, cvt = cpu_threads()
        sample = "%d detections per step" % per_step
    val = n / dt
    emit(({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": c.get("scaling", "weak"),
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["name"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": "%s x %d steps of the bench workload; " % (sample, args.steps) + CPU_NOTE % (cores, cvt)},
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
    c = CONFIGS[args.config]
    L = _lib.lib()
    peak, peak_src = measured_peaks()
    flops_crop = L.p2p_flops_per_crop(BACKBONE.encode())
    dtype = "f16 hi/lo operand pairs, f32 accumulate (fp16x3)" if args.precision == "fp16x3" else "f16, f32 accumulate"

    def timed(fn, steps, warmup, eng):
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

    def profiled(fn, eng, steps=2):
        """conv / other generator-kernel milliseconds per step, every launch bracketed by events, on the real step."""
        msk, cnt = (ctypes.c_double * 2)(), (ctypes.c_int * 2)()
        fn()
        _lib.check(L.p2p_engine_prof_begin(eng))
        for _ in range(steps):
            fn()
        _lib.check(L.p2p_engine_prof_end(eng, msk, cnt))
        return msk[0] / steps, msk[1] / steps, cnt[0] // steps, cnt[1] // steps

    # ================================================================== config 2: generator only
    if c["kind"] == "net":
        from pix2pose_b200.ae_model import aemodel_unet_resnet50
        B = c["batch"]
        gen = aemodel_unet_resnet50(p=1.0, precision=args.precision, capacity=B)
        gen.load_weights(Wt.synthetic_weights(BACKBONE, 1))
        eng = gen.engine.handle
        x = np.random.RandomState(0).uniform(-1, 1, (B, 128, 128, 3)).astype(np.float32)
        xp = _lib.pinned_array(x.shape, np.float32); xp[...] = x
        dec = _lib.pinned_array((B, 128, 128, 3), np.float32); prob = _lib.pinned_array((B, 128, 128, 1), np.float32)
        iters = 50                                              # >= 50 forwards per measurement
        clocks = ClockSampler(local_rank); clocks.start()
        l0 = gen.engine.launch_count
        ms_fwd = []
        for _ in range(max(args.warmup, 3)):
            gen.time_forward(x, warmup=1, iters=5)
        D.barrier()
        for _ in range(args.steps):
            ms_fwd.append(gen.time_forward(x, warmup=0, iters=iters))     # device-resident input, CUDA events around `iters` forwards
        launches_per_fwd = (gen.engine.launch_count - l0) // (max(args.warmup, 3) * 6 + args.steps * iters)
        clk = clocks.stop()
        ms = D.max_over_ranks(float(np.mean(ms_fwd)))

        def step_e2e():
            _lib.check(L.p2p_predict(eng, gen._model, _lib.fptr(xp), B, _lib.fptr(dec), _lib.fptr(prob)))

        ms_e2e, _ = timed(step_e2e, args.steps * 5, 3, eng)
        msk, cnt = (ctypes.c_double * 2)(), (ctypes.c_int * 2)()
        _lib.check(L.p2p_engine_profile_forward(eng, gen._model, _lib.fptr(x), B, msk, cnt))
        if rank != 0:
            D.shutdown()
            return
        achieved = flops_crop * B / (msk[0] * 1e-3) / 1e12
        traffic, tnote = conv_traffic(B)
        out = {
            "metric": METRIC, "value": B * world / (ms * 1e-3), "unit": "crops/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": c["name"], "backbone": BACKBONE, "batch": B, "forwards_per_step": iters, "precision": args.precision,
                       "unit_note": "network crops (one 128x128x3 forward each); no PnP in this configuration",
                       "l2": "a forward streams ~1.3 GB of activations + 112 MB of weights through the 126 MB L2; no flush needed"},
            "clocks": clk,
            "e2e": {"value": B * world / (ms_e2e / (args.steps * 5) * 1e-3), "unit": "crops/s", "h2d_bytes_per_step": int(x.nbytes),
                    "d2h_bytes_per_step": int(dec.nbytes + prob.nbytes)},
            "gpu_launches": int(launches_per_fwd * iters * args.steps),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "step_frac": flops_crop * B / (ms * 1e-3) / 1e12 / peak, "traffic": traffic, "traffic_note": tnote,
                         "peak_source": peak_src,
                         "kernel": "conv_tc_{pair,slab,persistent}_kernel (%d tcgen05 conv launches of one %d-crop forward: %.3f ms; "
                                   "other kernels %.3f ms)" % (cnt[0], B, msk[0], msk[1])},
        }
        if world == 1:
            use_all_host_threads()
            v, dt = cpu_net_sample(32)
            cores, _ = cpu_threads()
            out["cpu_baseline"] = {"value": v, "unit": "crops/s", "cores": cores, "kind": "port",
                                   "sample": "32 crops through the torch-CPU fp32 generator (%.1f s, %d threads; stand-in for Keras/TF-CPU)" % (dt, cores)}
        emit(out)
        D.shutdown()
        return

    # ================================================================== configs 3 / 4 / 5: two-stage est_pose
    frames, rois, fids, oids, gidx, n_total = workload(args.config, rank, world)
    n_det = len(rois)
    H, W = c["H"], c["W"]
    if c["kind"] == "single":
        rec = pix2pose(Wt.synthetic_weights(BACKBONE, 1), c["K"], W, H, OBJ, th_outlier=TH_O, th_inlier=TH_I, backbone=BACKBONE,
                       precision=args.precision, capacity=args.capacity, max_dets=n_det)
        runner, n_objects = rec, 1
    else:
        from pix2pose_b200.stream import MultiObjectRecognizer
        objs = sorted(set(int(o) for o in (c["objects"] if not args.local_objects else np.unique(oids))))
        multi = MultiObjectRecognizer({o: Wt.synthetic_weights(BACKBONE, o) for o in objs}, c["K"], W, H, {o: obj_param(o) for o in objs},
                                      th_outlier=TH_O, th_inlier=TH_I, backbone=BACKBONE, precision=args.precision,
                                      capacity=args.capacity, max_dets=n_det)
        rec = next(iter(multi.models.values()))
        runner, n_objects = multi, len(objs)
    eng = rec.generator_train.engine.handle
    gather = D.AsyncGather((n_total + world - 1) // world + 1, n_total)
    state = {}

    def run_batch(frames_host, frames_dev):
        if c["kind"] == "single":
            res = rec.est_pose_batch(frames_host, rois, fids, frames_dev=frames_dev)
            state["n_cand"], state["status"] = res.n_cand, res.status
            return res.records()
        r, status = runner.est_pose_stream(frames_host, rois, oids, fids, frames_dev=frames_dev)
        state["status"] = status
        return r

    def step_dev():
        r = run_batch(None, state["fdev"])
        state.setdefault("fwd_ms", []).append(rec.last_forward_ms())          # event nodes inside the run just finished
        state["all"] = gather.result()                  # last step's gathered records (None on the first)
        gather.submit(r, gidx)

    state["fdev"] = runner.upload_frames(frames, n_det)
    if args.profile:                      # short run for `ncu` launch lists: 1 warm-up + 1 step, no JSON
        step_dev()
        step_dev()
        return
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = rec.launch_count
    warm = max(args.warmup, 3)
    ms_dev, wall_dev = timed(step_dev, args.steps, warm, eng)
    fwd_ms = D.max_over_ranks(float(np.mean(state["fwd_ms"][-args.steps:])))      # generator time per step INSIDE the timed region
    launches = (rec.launch_count - l0) // (args.steps + warm) * args.steps
    clk = clocks.stop()
    status = state["status"]
    ok_frac = float(np.mean(status == 1))
    # stage-2 candidates per detection (network crops per detection = 1 + this)
    if c["kind"] == "single":
        n_cand_mean = float(np.mean(state["n_cand"]))
    else:
        n_cand_mean = float(len(TH_O))                 # random weights: every threshold yields a candidate (checked on config 3)

    # ---- e2e: pinned host frames through the public API; every step's frames cross the bus inside the timed region
    # (on the copy stream, one step ahead of the run that consumes them) and its pose records come back
    pinned = _lib.pinned_array(frames.shape, np.uint8)
    pinned[...] = frames
    state["next"] = runner.upload_frames(pinned, n_det)

    def step_e2e():
        cur = state["next"]
        state["next"] = runner.upload_frames(pinned, n_det)      # H2D of the next step's frames overlaps this step's kernels
        r = run_batch(None, cur)
        state["all"] = gather.result()
        gather.submit(r, gidx)

    ms_e2e, _ = timed(step_e2e, args.steps, 5, eng)     # 5 warm-ups: two frame buffers alternate, each configuration runs plain once, then is captured
    h2d = int(frames.nbytes + n_det * ctypes.sizeof(_Det) + 64)
    d2h = int(n_det * ctypes.sizeof(_Pose))
    # ---- roofline of the dominant kernel class (tcgen05 implicit-GEMM conv), measured live on the real step
    conv_ms, other_ms, n_conv, n_other = profiled(lambda: run_batch(None, state["fdev"]), eng)
    all_rec = gather.result()
    if rank != 0:
        D.shutdown()
        return
    crops_per_step = n_det * (1 + n_cand_mean)
    achieved = flops_crop * crops_per_step / (fwd_ms * 1e-3) / 1e12
    per_step_ms = ms_dev / args.steps
    value = n_total / (per_step_ms * 1e-3)
    traffic, tnote = conv_traffic(args.capacity)
    # ---- CPU baseline: bounded sample of the same workload (rank 0 at N = 1 only)
    cpu_baseline = None
    if world == 1:
        use_all_host_threads()
        ora = make_cpu_port(args.config)
        cpu_sample(ora, frames, rois, fids, [0], oids)                  # warm-up (thread pools)
        cpu_val, cpu_dt, cpu_ok = cpu_sample(ora, frames, rois, fids, range(args.cpu_sample), oids)
        cores, cvthreads = cpu_threads()
        cpu_baseline = {"value": cpu_val, "unit": "crops/s", "cores": cores, "kind": "port",
                        "sample": "first %d detections of the same workload (%.1f s): " % (args.cpu_sample, cpu_dt) + CPU_NOTE % (cores, cvthreads)}
    out = {
        "metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": c["name"], "backbone": BACKBONE, "global_detections": int(n_total), "detections_this_rank": int(n_det),
                   "objects_resident": n_objects, "object_groups_this_rank": 1 if oids is None else int(len(np.unique(oids))),
                   "stage2_candidates_per_detection": n_cand_mean,
                   "network_crops_per_s": value * (1 + n_cand_mean), "pose_found_fraction": ok_frac,
                   "parallelism": "dp%d (detections sharded%s, no data-path collective; one all_gather of 16-double pose records per step, "
                                  "consumed one step later)" % (world, "" if oids is None else " by object group"),
                   "precision": args.precision, "engine_capacity": args.capacity,
                   "l2": "per-step working set (>= 4 GB of activations + 112 MB weights per object) exceeds the 126 MB L2; no flush needed",
                   "wall_ms_per_step": wall_dev / args.steps,
                   "gathered_records_ok": bool(all_rec is not None and all_rec.shape == (n_total, 16))},
        "clocks": clk,
        "e2e": {"value": n_total / (ms_e2e / args.steps * 1e-3), "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "step_frac": flops_crop * crops_per_step / (per_step_ms * 1e-3) / 1e12 / peak,
                     "traffic": traffic, "traffic_note": tnote, "peak_source": peak_src,
                     "kernel": "conv_tc_{pair,slab,persistent}_kernel = the generator forwards of a step (%d network crops): %.3f ms per step, from "
                               "CUDA event nodes around them inside the timed steps themselves (part of the captured graph).  Launch by "
                               "launch (separate profiled pass, events around every kernel, gaps included): %d tcgen05 conv launches %.3f ms, "
                               "%d other generator launches (stem input conversion, max-pool, split-K reduce) %.3f ms" % (
                                   int(crops_per_step), fwd_ms, n_conv, conv_ms, n_other, other_ms),
                     "note": "algorithmic FLOPs (10.70 GFLOP/crop); fp16x3 issues 3 MMAs per k-step, so tensor-pipe work is 3x this and frac is "
                             "bounded by 1/3; step_frac = the same FLOPs over the whole step (crops, masks, PnP, launch gaps included)"},
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
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json configuration (default 3: the one the metric is quoted on)")
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp16"])
    ap.add_argument("--capacity", type=int, default=256)
    ap.add_argument("--cpu-sample", type=int, default=96)    # ~15 s of CPU work on 16 cores
    ap.add_argument("--local-objects", action="store_true", help="configs 4/5: keep only the weight sets of this rank's object groups resident")
    ap.add_argument("--profile", action="store_true", help="one warm-up + one step only (for ncu launch lists)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
