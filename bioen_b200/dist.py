"""Multi-GPU plumbing: the structure axis N is sharded over the ranks of a torch.distributed job (one process
per GPU).  torch.distributed is used ONLY to bootstrap (broadcast the NCCL unique id, scatter/gather host
vectors); the per-evaluation exchanges are issued by libbioen_b200.so itself on its compute stream
(csrc/comm.cuh: a peer-memory kernel over NVLink, NCCL when peers cannot be mapped): M+3 doubles per log-weights
evaluation, 1-2 doubles per L-BFGS dot product.
"""
import ctypes

import numpy as np

from . import _lib
from .problem import FORCES, LOGW, Problem


def bind_to_gpu_numa(device):
    """Pin this process to the CPUs NVML reports as local to GPU `device` (its NUMA node) -- call it before allocating
    pinned host buffers, so that they land in memory next to the GPU's PCIe root.  With one process per GPU on a
    two-socket box this is what keeps the per-evaluation host<->device vector traffic of 8 ranks from crossing the
    socket interconnect.  Returns the CPU list, or None when NVML (or the affinity call) is not available."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def shard_bounds(n_total, rank, world):
    """Contiguous column range [lo, hi) of `rank`; the remainder is spread over the first ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_total, world):
    return [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]


def broadcast_bytes(payload, nbytes, src=0, group=None, device=None):
    """Broadcast a fixed-size byte string from `src` through torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank(group) == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def allgather_vector(local, n_total, group=None, device=None):
    """Concatenate the rank-local slices of an N-vector (host NumPy in, host NumPy out)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=torch.float64, device=device)
    pad[:local.size] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(pad.device)
    out = [torch.zeros(mx, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)])


def connect(problem, n_total, group=None, device=None):
    """Create the library's NCCL communicator for `problem` (this rank's column block)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    raw = b"\0" * 128
    if rank == 0:
        cbuf = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().bioen_b200_nccl_unique_id(cbuf), "nccl_unique_id")
        raw = cbuf.raw
    uid = broadcast_bytes(raw, 128, 0, group, device)
    problem.comm_init(uid, rank, world, n_total)
    return problem


class ShardedProblem:
    """A BioEn problem whose N columns are spread over the ranks of the default process group.

    Every rank passes the FULL host arrays (or its own slice with `presliced=True`); N-vectors returned by
    the methods are full-length (gathered), so the object can be used like `Problem` on every rank."""

    def __init__(self, yTilde, device, group=None, presliced=False, n_total=None, structure_major_only=False):
        """structure_major_only: every rank keeps only the structure-major copy of its block (forces method on the
        fused kernels, half the device memory; see Problem)."""
        import torch
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.tdev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else None
        yT = np.asarray(yTilde, dtype=np.float64)
        self.n_total = int(n_total if presliced else yT.shape[1])
        self.lo, self.hi = shard_bounds(self.n_total, self.rank, self.world)
        local = yT if presliced else np.ascontiguousarray(yT[:, self.lo:self.hi])
        self.m = local.shape[0]
        self.p = Problem(local, device=device, **({"structure_major_only": True} if structure_major_only else {}))
        if self.world > 1:
            connect(self.p, self.n_total, group, self.tdev)

    def _slice(self, v):
        v = _lib.vec(v)
        return v[self.lo:self.hi] if v.size == self.n_total else v

    def _gather(self, v):
        return allgather_vector(v, self.n_total, self.group, self.tdev) if self.world > 1 else v

    def set_logw(self, G, YTilde, theta):
        self.p.set_logw(self._slice(G), YTilde, theta)

    def set_forces(self, w0, YTilde, theta):
        self.p.set_forces(self._slice(w0), YTilde, theta)

    # The methods below take the same arguments as bioen_b200.Problem's (the optional `method` must be the one
    # that was set), so that optimize.{log_weights,forces}.find_optimum(..., problem=ShardedProblem(...)) runs the
    # reference API on all GPUs of the job.
    @property
    def method(self):
        return self.p.method

    @property
    def n(self):
        return self.n_total

    @property
    def device(self):
        return self.p.device

    def _check(self, method):
        if method is not None and method != self.p.method:
            raise ValueError("ShardedProblem: call set_logw / set_forces for this method first")
        return self.p.method == LOGW

    def set_theta(self, theta):
        self.p.set_theta(theta)

    def objective_and_gradient(self, x, method=None):
        if self._check(method):
            f, g = self.p.objective_and_gradient(self._slice(x))
            return f, self._gather(g)
        return self.p.objective_and_gradient(x)

    def objective(self, x, method=None):
        return self.p.objective(self._slice(x) if self._check(method) else x)

    def gradient(self, x, method=None):
        return self.objective_and_gradient(x, method)[1]

    def weights(self, x, method=None):
        w, s = self.p.weights(self._slice(x) if self._check(method) else x)
        return self._gather(w), s

    def average(self, w):
        """y.w over all ranks for full-length (or rank-local) weights; the partial sums are combined inside the
        library, every rank gets the same M-vector."""
        return self.p.average(self._slice(w))

    def affine_rows(self, scale, offset):
        """Row-affine transform of every rank's column block (bioen_b200.nuisance commits refits with it)."""
        self.p.affine_rows(scale, offset)

    def opt_lbfgs(self, x0, method=None, **cfg):
        if self._check(method):
            x, fmin, code, info = self.p.opt_lbfgs(self._slice(x0), **cfg)
            return self._gather(x), fmin, code, info
        return self.p.opt_lbfgs(x0, **cfg)

    def opt_gsl(self, x0, method=None, **cfg):
        if self._check(method):
            x, fmin, code, info = self.p.opt_gsl(self._slice(x0), **cfg)
            return self._gather(x), fmin, code, info
        return self.p.opt_gsl(x0, **cfg)

    def like(self, y):
        """A problem of the same kind (same devices, same sharding) holding another m' x N matrix, e.g. the
        un-normalised observables y for the post-processing y.wopt."""
        return ShardedProblem(y, self.p.device, group=self.group)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def theta_scan(self, thetas, x0=None, method=None, **cfg):
        """Batched theta scan on the sharded problem (same signature as Problem.theta_scan); X planes are gathered
        to full length for log-weights."""
        if self._check(method):
            if x0 is not None:
                x0 = np.asarray(x0, dtype=np.float64)
                x0 = x0[..., self.lo:self.hi] if x0.shape[-1] == self.n_total else x0
            X, fmin, codes, info = self.p.theta_scan(thetas, x0=x0, method=LOGW, **cfg)
            X = np.stack([self._gather(X[q]) for q in range(X.shape[0])])
            return X, fmin, codes, info
        return self.p.theta_scan(thetas, x0=x0, method=FORCES, **cfg)

    def close(self):
        cached = getattr(self, "_y_cache", None)
        if cached is not None:
            self._y_cache = None
            cached[1].close()
        self.p.close()


class LocalGroup:
    """The N-sharded problem inside ONE process: `world` bioen_b200.Problem objects (one host thread each) on the
    given devices -- which may all be the same GPU -- joined through bioen_b200_comm_init_local, exchanging over peer
    memory without NCCL or CUDA IPC.  `call(fn)` runs fn(rank, problem, lo, hi) on every rank concurrently and
    returns the list of results; every collective operation of the library (evaluations, minimisers) must be issued
    on all ranks this way.  Used by the tests to run the sharded kernels and the exchange protocol on a one-GPU box,
    and usable as a single-process multi-GPU driver."""

    _next_group = [1]

    def __init__(self, yTilde, world, devices=None, problem_cls=None):
        from concurrent.futures import ThreadPoolExecutor
        yT = np.asarray(yTilde, dtype=np.float64)
        self.world = int(world)
        self.devices = list(devices) if devices is not None else [0] * self.world
        self.m, self.n_total = yT.shape
        self.bounds = [shard_bounds(self.n_total, r, self.world) for r in range(self.world)]
        self.group = LocalGroup._next_group[0]
        LocalGroup._next_group[0] += 1
        self.pool = ThreadPoolExecutor(max_workers=self.world)
        cls = problem_cls or Problem
        self.problems = [None] * self.world

        def make(r):
            lo, hi = self.bounds[r]
            p = cls(np.ascontiguousarray(yT[:, lo:hi]), device=self.devices[r])
            self.problems[r] = p
            if self.world > 1:
                p.comm_init_local(self.group, r, self.world, self.n_total)
            return True
        list(self.pool.map(make, range(self.world)))

    def call(self, fn):
        futs = [self.pool.submit(fn, r, self.problems[r], *self.bounds[r]) for r in range(self.world)]
        return [f.result() for f in futs]

    def gather(self, parts):
        return np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in parts])

    def close(self):
        if self.pool is not None:
            self.call(lambda r, p, lo, hi: p.close())
            self.pool.shutdown()
            self.pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


__all__ = ["shard_bounds", "shard_sizes", "broadcast_bytes", "allgather_vector", "connect", "ShardedProblem", "LocalGroup", "bind_to_gpu_numa",
           "LOGW", "FORCES"]
