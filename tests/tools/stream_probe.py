"""Diagnostic (GPU box): host -> device paths at config-3 size: whole-matrix upload, streamed y.w."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bioen_b200
M, N = 1000, 1000000
y = np.empty((M, N)); y[:] = np.arange(N) % 7 * 0.25
w = np.full(N, 1.0 / N)
for rep in range(2):
    t0 = time.perf_counter()
    p = bioen_b200.Problem(y)
    t1 = time.perf_counter()
    a = p.average(w)
    t2 = time.perf_counter()
    b = p.average_streamed(y, w)
    t3 = time.perf_counter()
    c = p.average_streamed(y, w, chunk_bytes=256 << 20)
    t4 = time.perf_counter()
    p.close()
    print("upload %.3f s (%.1f GB/s)  resident average %.4f s  streamed 1 GB chunks %.3f s  256 MB chunks %.3f s  max diff %.2e"
          % (t1 - t0, 8 / (t1 - t0), t2 - t1, t3 - t2, t4 - t3, max(np.max(np.abs(a - b)), np.max(np.abs(a - c)))), flush=True)
