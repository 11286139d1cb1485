"""Multi-GPU plumbing: one process per GPU (torchrun), detections sharded by index, no data-path
collective -- each detection is independent (tools/5_evaluation_bop_basic.py:289-323 has no cross-ROI
state) -- and ONE gather of fixed-size pose records per batch (16 float64 per detection:
R9, t3, n_inliers, frac_inlier, status, global index).  ``torch.distributed`` is plumbing only
(NCCL on GPUs, gloo in the CPU tests)."""
import os

import numpy as np

RECORD = 16


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_indices(n, rank, world):
    """Indices of the detections rank `rank` owns: i % world == rank (round-robin keeps per-object
    sub-batches balanced when the stream interleaves objects)."""
    return np.arange(rank, n, world, dtype=np.int64)


def shard_by_object(obj_ids, rank, world):
    """Indices of the detections rank `rank` owns when a multi-object stream is sharded: the stream is ordered by object id
    (stable) and cut into `world` contiguous ranges of near-equal length, so that a rank sees as few objects as possible
    and each object's detections form one large sub-batch there (one generator forward per object and rank; SURVEY
    section 8e: "within a rank, group by object id so one weight set serves a whole sub-batch")."""
    order = np.argsort(np.asarray(obj_ids), kind="stable")
    n = len(order)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return np.sort(order[lo:hi])


def init(backend=None):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif world == 1 and torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if torch.cuda.is_available():
        # the C library links its own CUDA runtime: point it at this rank's GPU as well
        from . import _lib
        _lib.check(_lib.lib().p2p_set_device(local_rank))
    return rank, local_rank, world


def shutdown():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def gather_records(local_records, global_index, n_total, device=None):
    """All ranks contribute (n_local, 16) records whose last column is overwritten with `global_index`;
    every rank gets the (n_total, 16) array ordered by global detection index.  One all_gather of a
    padded fixed-size tensor (<= 128 B per detection)."""
    import torch
    import torch.distributed as dist
    rec = np.array(local_records, np.float64).reshape(-1, RECORD)
    rec[:, RECORD - 1] = np.asarray(global_index, np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = np.zeros((n_total, RECORD))
        out[rec[:, RECORD - 1].astype(np.int64)] = rec
        return out
    world = dist.get_world_size()
    cap = (n_total + world - 1) // world
    pad = np.full((cap, RECORD), -1.0)
    pad[: rec.shape[0]] = rec
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(pad).to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    allr = torch.cat(parts).cpu().numpy()
    allr = allr[allr[:, RECORD - 1] >= 0]
    out = np.zeros((n_total, RECORD))
    out[allr[:, RECORD - 1].astype(np.int64)] = allr
    return out


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class AsyncGather:
    """The per-step result gather taken off the critical path: ``submit`` stages the local (n_local, 16) records in pinned
    memory and enqueues H2D -> all_gather -> D2H on a side stream without blocking the host; ``result`` (called one step
    later) waits for that step's event only.  No numpy -> pageable -> .cpu() hop and no implicit barrier inside the step.
    Falls back to the blocking ``gather_records`` without NCCL (gloo CPU tests, world 1)."""

    def __init__(self, n_local_max, n_total):
        import torch
        import torch.distributed as dist
        self.n_total = n_total
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.nccl = self.world > 1 and dist.get_backend() == "nccl"
        self.pending = None
        if self.nccl:
            dev = torch.device("cuda", torch.cuda.current_device())
            self.cap = n_local_max
            self.host_in = torch.full((self.cap, RECORD), -1.0, dtype=torch.float64).pin_memory()
            self.dev_in = torch.empty((self.cap, RECORD), dtype=torch.float64, device=dev)
            self.dev_out = torch.empty((self.world * self.cap, RECORD), dtype=torch.float64, device=dev)
            self.host_out = torch.empty((self.world * self.cap, RECORD), dtype=torch.float64).pin_memory()
            self.stream = torch.cuda.Stream(device=dev)
            self.done = torch.cuda.Event()

    def submit(self, local_records, global_index):
        import torch
        import torch.distributed as dist
        rec = np.array(local_records, np.float64).reshape(-1, RECORD)
        rec[:, RECORD - 1] = np.asarray(global_index, np.float64)
        if not self.nccl:
            self.pending = gather_records(rec, global_index, self.n_total)
            return
        if self.pending is not None:
            self.done.synchronize()          # the staging buffers are reused: the previous gather must have left them
        self.host_in.fill_(-1.0)
        self.host_in[: rec.shape[0]] = torch.from_numpy(rec)
        with torch.cuda.stream(self.stream):
            self.dev_in.copy_(self.host_in, non_blocking=True)
            dist.all_gather_into_tensor(self.dev_out, self.dev_in)
            self.host_out.copy_(self.dev_out, non_blocking=True)
            self.done.record(self.stream)
        self.pending = True

    def result(self):
        """(n_total, 16) records of the last submitted step, ordered by global detection index."""
        if self.pending is None:
            return None
        if not self.nccl:
            return self.pending
        self.done.synchronize()
        allr = self.host_out.numpy()
        allr = allr[allr[:, RECORD - 1] >= 0]
        out = np.zeros((self.n_total, RECORD))
        out[allr[:, RECORD - 1].astype(np.int64)] = allr
        return out
