"""Diagnostic (GPU box): L-BFGS at config 2 with plain launches / CUDA-graph replay (BIOEN_B200_GRAPHS=1), with and
without the two tiny memcpy nodes inside the graph (BIOEN_B200_GRAPHS_NOMEMCPY=1)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bioen_b200
M, N = 500, 100000
rng = np.random.default_rng(12345)
a = rng.standard_normal(M); YT = a + rng.standard_normal(M)
with bioen_b200.Problem(shape=(M, N)) as p:
    p.generate(12345, 0, a, 2.0)
    p.set_option(5, int(os.environ.get("PROBE_PERSISTENT", "0")))
    p.set_logw(np.zeros(N), YT, 10.0)
    for rep in range(2):
        t0 = time.perf_counter()
        x, f, code, info = p.opt_lbfgs(np.zeros(N), max_iterations=150)
        dt = time.perf_counter() - t0
        print("graphs=%s nomemcpy=%s persistent=%s: %.4f s, %d evals, %.3f ms per evaluation+update, f=%.8f" % (
            os.environ.get("BIOEN_B200_GRAPHS", "0"), os.environ.get("BIOEN_B200_GRAPHS_NOMEMCPY", "0"),
            os.environ.get("PROBE_PERSISTENT", "0"), dt, info["evaluations"], 1e3 * dt / info["evaluations"], f), flush=True)
