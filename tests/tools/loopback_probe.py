"""Debug aid: time every collective operation of an in-process group on one GPU (a timeout shows up as ~5 s)."""
import os
import sys
import time

import numpy as np

os.environ.setdefault("BIOEN_B200_P2P_TIMEOUT_S", "5")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bioen_b200 import dist as D  # noqa: E402
from oracle import oracle as O  # noqa: E402

M, N, theta, world = 37, 5001, 1.0, int(sys.argv[1]) if len(sys.argv) > 1 else 2
P = O.synthetic_problem(M, N, seed=12345)
rng = np.random.default_rng(3)
G = 0.1 * rng.standard_normal(N)
g1 = G + 0.1 * rng.standard_normal(N)


def timed(name, grp, fn):
    t0 = time.perf_counter()
    out = grp.call(fn)
    print("%-40s %.3f s" % (name, time.perf_counter() - t0), flush=True)
    return out


with D.LocalGroup(P["yTilde"], world) as grp:
    timed("set_logw", grp, lambda r, p, lo, hi: p.set_logw(G[lo:hi], P["YTilde"], theta))
    for fused in (1, 0, 1, 0):
        timed("set_option fused=%d" % fused, grp, lambda r, p, lo, hi: p.set_option(4, fused))
        for k in range(3):
            res = timed("  f+g", grp, lambda r, p, lo, hi: p.objective_and_gradient(g1[lo:hi]))
            print("     f =", [f for f, _ in res])
        res = timed("  f only", grp, lambda r, p, lo, hi: p.objective(g1[lo:hi]))
        res = timed("  weights", grp, lambda r, p, lo, hi: p.weights(g1[lo:hi])[0].sum())
        res = timed("  lbfgs 10 it", grp, lambda r, p, lo, hi: p.opt_lbfgs(G[lo:hi], max_iterations=10)[1])
        print("     fmin =", res)
