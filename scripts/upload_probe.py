"""Diagnostic (GPU box): host -> device upload rate of an 8 GB pageable matrix, plain vs staged."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bioen_b200
M, N = 1000, 1_000_000
y = np.ones((M, N))
y[::7] = 2.0
for mode, thr in (("plain", 1), ("staged", 2), ("staged", 4), ("staged", 6), ("staged", 12), ("staged", 6)):
    os.environ["BIOEN_B200_UPLOAD"] = mode
    os.environ["BIOEN_B200_UPLOAD_THREADS"] = str(thr)
    t0 = time.perf_counter()
    p = bioen_b200.Problem(y)
    dt = time.perf_counter() - t0
    chk = p.download(7, 1, 12345, 3)
    p.close()
    print("%-6s threads %2d  %.3f s  %.1f GB/s  check %s" % (mode, thr, dt, M * N * 8 / dt / 1e9, chk.ravel()))
print("cores", os.cpu_count())
