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
