"""CPU, world_size 2 over gloo: host-side logic of the N-sharded path (bioen_b200/dist.py) -- shard bounds,
byte broadcast used for the NCCL id, vector gather -- and the communication SCHEDULE of a sharded evaluation
(local (max, sum-exp) pairs -> all-gather -> local partial avg -> ONE all-reduce of M+3 doubles -> local
gradient slice), emulated with NumPy per rank and checked against the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bioen_b200 import dist as D
        from oracle import oracle as O
        M, N, theta = 13, 1001, 4.0
        P = O.synthetic_problem(M, N, seed=5)
        rng = np.random.default_rng(11)
        G = 0.2 * rng.standard_normal(N)
        g = G + 0.3 * rng.standard_normal(N)
        lo, hi = D.shard_bounds(N, rank, world)
        # 1. byte broadcast (the NCCL unique id travels this way)
        payload = bytes(range(128)) if rank == 0 else b"\0" * 128
        assert D.broadcast_bytes(payload, 128) == bytes(range(128))
        # 2. sharded evaluation schedule, NumPy per rank + the same collectives the library issues
        yl, gl, Gl = P["yTilde"][:, lo:hi], g[lo:hi], G[lo:hi]
        mx = gl.max()
        pair = torch.tensor([mx, np.exp(gl - mx).sum()], dtype=torch.float64)
        pairs = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(pairs, pair)                                   # comm 1: 2 doubles per rank
        Mx = max(float(p[0]) for p in pairs)
        S = sum(float(p[1]) * np.exp(float(p[0]) - Mx) for p in pairs)
        wl = np.exp(gl - Mx) / S
        buf = torch.from_numpy(np.concatenate([yl @ wl, [np.dot(gl - Gl, wl), np.dot(gl, wl), np.dot(Gl, wl)]]))
        dist.all_reduce(buf)                                           # comm 2: M + 3 doubles
        buf = buf.numpy()
        avg, t0, gbar, Gbar = buf[:M], buf[M], buf[M + 1], buf[M + 2]
        r = avg - P["YTilde"].ravel()
        # log s0 from G the same way
        pg = torch.tensor([Gl.max(), np.exp(Gl - Gl.max()).sum()], dtype=torch.float64)
        pgs = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(pgs, pg)
        M0 = max(float(p[0]) for p in pgs)
        logs0 = M0 + np.log(sum(float(p[1]) * np.exp(float(p[0]) - M0) for p in pgs))
        f = theta * (t0 - (Mx + np.log(S)) + logs0) + 0.5 * np.dot(r, r)
        grad_l = wl * theta * (gl - gbar - Gl + Gbar) + wl * (yl.T @ r - np.dot(r, avg))
        grad = D.allgather_vector(grad_l, N)                           # host-side gather of the result
        fo, go = O.logw_fg(g, G, P["yTilde"], P["YTilde"], theta)
        assert abs(f - fo) < 1e-12 * abs(fo)
        assert np.max(np.abs(grad - go)) < 1e-12 * np.max(np.abs(go))
        out[rank] = (lo, hi)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_partition():
    from bioen_b200.dist import shard_bounds, shard_sizes
    for n, w in ((10, 3), (1000000, 8), (7, 8), (128, 1), (1001, 2)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 3, 3)


def test_sharded_schedule_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: (0, 501), 1: (501, 1001)}


class _StandIn:
    """CPU stand-in with bioen_b200.Problem's interface: this rank's columns, oracle arithmetic on the gathered
    problem.  It lets the world-size-2 test below drive dist.ShardedProblem and optimize.*.find_optimum -- the
    slicing, gathering and call signatures -- without a GPU.  (The GPU library under ShardedProblem is covered by
    tests/mgpu_check.py.)"""

    def __init__(self, local, device=0):
        from bioen_b200 import dist as D
        self.local = np.asarray(local, dtype=np.float64)
        self.m, self.n = self.local.shape
        self.device, self.method = device, None
        world = dist.get_world_size()
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([self.n]))
        self.sizes = [int(s) for s in sizes]
        self.n_total = sum(self.sizes)
        self.lo = sum(self.sizes[:dist.get_rank()])
        self.full = np.stack([D.allgather_vector(self.local[i], self.n_total) for i in range(self.m)])

    def comm_init(self, *a):
        pass

    def _full(self, v):
        from bioen_b200 import dist as D
        return D.allgather_vector(np.asarray(v, dtype=np.float64).ravel(), self.n_total)

    def _mine(self, v):
        return np.ascontiguousarray(v[self.lo:self.lo + self.n])

    def set_logw(self, G, Y, theta):
        self.G, self.Y, self.theta, self.method = self._full(G), np.asarray(Y, dtype=np.float64).ravel(), theta, 0

    def set_forces(self, w0, Y, theta):
        self.w0, self.Y, self.theta, self.method = self._full(w0), np.asarray(Y, dtype=np.float64).ravel(), theta, 1

    def _fg(self, x):
        from oracle import oracle as O
        if self.method == 0:
            return O.logw_fg_np(x, self.G, self.full, self.Y, self.theta)
        return O.forces_fg_np(x, self.w0, self.full, self.Y, self.theta)

    def objective(self, x, method=None):
        return self._fg(self._full(x) if self.method == 0 else np.asarray(x, dtype=np.float64).ravel())[0]

    def objective_and_gradient(self, x, method=None):
        if self.method == 0:
            f, g = self._fg(self._full(x))
            return f, self._mine(g)
        return self._fg(np.asarray(x, dtype=np.float64).ravel())

    def weights(self, x, method=None):
        from oracle import oracle as O
        if self.method == 0:
            w, s = O.logw_weights_np(self._full(x))
            return self._mine(w), s
        return self._mine(O.forces_weights_np(np.asarray(x, dtype=np.float64).ravel(), self.w0, self.full)), None

    def average(self, w):
        part = torch.from_numpy(self.local @ np.asarray(w, dtype=np.float64).ravel())
        dist.all_reduce(part)
        return part.numpy()

    def opt_lbfgs(self, x0, method=None, verbose=0, **cfg):
        from oracle import oracle as O
        start = self._full(x0) if self.method == 0 else np.asarray(x0, dtype=np.float64).ravel()
        r = O.lbfgs(self._fg, start, **cfg)
        x = self._mine(r["x"]) if self.method == 0 else r["x"]
        return x, r["fx"], r["code"], dict(iterations=r["iterations"], evaluations=r["evaluations"])

    def set_theta(self, theta):
        self.theta = theta

    def theta_scan(self, thetas, x0=None, method=None, verbose=0, **cfg):
        """Problem.theta_scan's signature; one oracle L-BFGS per theta from the shared start."""
        from oracle import oracle as O
        assert method == self.method
        X, fm, codes, its = [], [], [], []
        keep = self.theta
        for th in np.asarray(thetas, dtype=np.float64).ravel():
            self.theta = float(th)
            start = self._full(x0) if self.method == 0 else np.asarray(x0, dtype=np.float64).ravel()
            r = O.lbfgs(self._fg, start, **cfg)
            X.append(self._mine(r["x"]) if self.method == 0 else r["x"])
            fm.append(r["fx"])
            codes.append(r["code"])
            its.append(r["iterations"])
        self.theta = keep
        return np.stack(X), np.array(fm), np.array(codes), dict(iterations=np.array(its), evaluations=np.array(its),
                                                                 rounds=0, gemm_launches=0, seconds=0.0)

    def close(self):
        pass


def _worker_series(rank, world, port, out):
    """optimize.*.find_optimum_series(..., problem=ShardedProblem): the batched branch calls
    problem.theta_scan(chunk, x0=..., method=..., **cfg) and holds the post-processing matrix y as problem.like(y)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bioen_b200 import dist as D
        from bioen_b200 import optimize
        from oracle import oracle as O
        D.Problem = _StandIn
        D.connect = lambda p, n_total, group=None, device=None: p
        M, N = 7, 151
        P = O.synthetic_problem(M, N, seed=8)
        y = 0.5 * P["yTilde"] + 0.25
        thetas = [30.0, 3.0]
        cfg = optimize.minimize.Parameters("lbfgs")
        cfg["verbose"] = False
        kw = optimize.log_weights._lbfgs_kwargs(cfg)
        sp = D.ShardedProblem(P["yTilde"], device=0)
        res = optimize.log_weights.find_optimum_series(P["GInit"], P["G"], y, P["yTilde"], P["YTilde"], thetas, cfg,
                                                       problem=sp)
        assert len(res) == 2
        for (wopt, yopt, gopt, f0, f1), th in zip(res, thetas):
            r = O.lbfgs(lambda v: O.logw_fg_np(v, P["G"].ravel(), P["yTilde"], P["YTilde"].ravel(), th),
                        P["GInit"].ravel(), **kw)
            assert gopt.shape == (N,) and np.array_equal(gopt, r["x"]) and f1 == r["fx"] and f1 < f0
            assert wopt.shape == (N, 1) and np.allclose(yopt, y @ wopt.ravel(), rtol=1e-12)
        resf = optimize.forces.find_optimum_series(P["forces_init"], P["w0"], y, P["yTilde"], P["YTilde"], thetas, cfg,
                                                   problem=sp)
        for tup, th in zip(resf, thetas):
            r = O.lbfgs(lambda v: O.forces_fg_np(v, P["w0"].ravel(), P["yTilde"], P["YTilde"].ravel(), th),
                        P["forces_init"].ravel(), **kw)
            assert np.array_equal(tup[2], r["x"]) and tup[4] == r["fx"] and tup[0].shape == (N, 1)
            assert np.allclose(tup[1], y @ tup[0].ravel(), rtol=1e-12)
        with pytest.raises(ValueError):
            sp.theta_scan(thetas, method=D.LOGW)       # forces is the method that is set
        sp.close()
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_find_optimum_series_on_sharded_problem_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_series, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def _worker_find_optimum(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bioen_b200 import dist as D
        from bioen_b200 import optimize
        from oracle import oracle as O
        D.Problem = _StandIn                       # the GPU library is not available here
        D.connect = lambda p, n_total, group=None, device=None: p
        M, N, theta = 9, 203, 2.0
        P = O.synthetic_problem(M, N, seed=3)
        y = 0.5 * P["yTilde"] + 0.25               # "un-normalised observables": a different M x N matrix
        cfg = optimize.minimize.Parameters("lbfgs")
        cfg["verbose"] = False
        # log-weights through the reference API, N split over the two ranks
        sp = D.ShardedProblem(P["yTilde"], device=0)
        assert (sp.lo, sp.hi) == D.shard_bounds(N, rank, world) and sp.n == N
        wopt, yopt, gopt, f0, f1 = optimize.log_weights.find_optimum(P["GInit"], P["G"], y, P["yTilde"], P["YTilde"],
                                                                     theta, cfg, problem=sp)
        r = O.lbfgs(lambda v: O.logw_fg_np(v, P["G"].ravel(), P["yTilde"], P["YTilde"].ravel(), theta),
                    P["GInit"].ravel(), **optimize.log_weights._lbfgs_kwargs(cfg))
        assert wopt.shape == (N, 1) and gopt.shape == (N,) and yopt.shape == (M,)
        assert np.array_equal(gopt, r["x"]) and f1 == r["fx"] and f1 < f0
        w = np.exp(r["x"] - r["x"].max())
        w /= w.sum()
        assert np.allclose(wopt.ravel(), w, rtol=1e-13, atol=0) and np.allclose(yopt, y @ w, rtol=1e-12)
        # forces: replicated M-vector, weights gathered
        res = optimize.forces.find_optimum(P["forces_init"], P["w0"], y, P["yTilde"], P["YTilde"], theta, cfg,
                                           problem=sp)
        wopt, yopt, fopt, f0, f1, chi2, S = res
        r = O.lbfgs(lambda v: O.forces_fg_np(v, P["w0"].ravel(), P["yTilde"], P["YTilde"].ravel(), theta),
                    P["forces_init"].ravel(), **optimize.log_weights._lbfgs_kwargs(cfg))
        assert np.array_equal(fopt, r["x"]) and f1 == r["fx"]
        assert wopt.shape == (N, 1) and abs(wopt.sum() - 1.0) < 1e-12 and np.allclose(yopt, y @ wopt.ravel(), rtol=1e-12)
        assert abs(theta * S + chi2 - f1) < 1e-10 * abs(f1)
        # signature checks of the wrapper itself
        with pytest.raises(ValueError):
            sp.objective(P["GInit"], D.LOGW)       # forces is the method that is set
        sp.close()
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_find_optimum_on_sharded_problem_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_find_optimum, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
