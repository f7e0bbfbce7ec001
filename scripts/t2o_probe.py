"""Diagnostic (torchrun or single): repeated device L-BFGS runs on a generated problem, per-evaluation overhead."""
import os, sys, time, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bioen_b200
from bioen_b200 import _lib
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M = 1000; NT = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
N = NT // world
rng = np.random.default_rng(12345); ytrue = rng.standard_normal(M); yobs = ytrue + 0.5 * rng.standard_normal(M)
p = bioen_b200.Problem(shape=(M, N), device=local)
if world > 1:
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = ctypes.create_string_buffer(128); _lib.load().bioen_b200_nccl_unique_id(raw)
        idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    p.comm_init(bytes(idbuf.cpu().numpy().tobytes()), rank, world, NT)
p.generate(12345, rank * N, ytrue / 0.5, 2.0)
p.set_logw(np.zeros(N), yobs / 0.5, 10.0)
for rep in range(4):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, fmin, code, info = p.opt_lbfgs(np.zeros(N), max_iterations=60)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    for _ in range(20): p.objective_and_gradient(x)
    de = (time.perf_counter() - t1) / 20
    if rank == 0:
        print("world %d N/gpu %d: lbfgs %.3f s, %d it, %d evals -> %.3f ms/eval ; host-vector eval %.3f ms" % (world, N, dt, info["iterations"], info["evaluations"], 1e3 * dt / info["evaluations"], 1e3 * de), flush=True)
p.close()
if world > 1: dist.destroy_process_group()
