"""Small driver for compute-sanitizer (memcheck): every kernel family once, at ragged shapes.
    compute-sanitizer --tool memcheck python tests/tools/sanitize_smoke.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bioen_b200
from bioen_b200.problem import FORCES, LOGW
rng = np.random.default_rng(0)
for (M, N) in ((33, 1001), (300, 777), (1030, 515)):
    y = rng.standard_normal((M, N))
    Y = rng.standard_normal(M)
    w0 = rng.random(N) + 0.1
    w0 /= w0.sum()
    G = np.log(w0)
    with bioen_b200.Problem(y) as p:
        p.set_logw(G, Y, 3.0)
        p.objective_and_gradient(G + 0.1)
        p.opt_lbfgs(G, max_iterations=4)
        p.opt_gsl(G, max_iterations=3)
        p.theta_scan([10.0, 1.0, 5.0], x0=G, max_iterations=3)
        p.average(w0)
        for fused in (1, 0):
            p.set_option(1, fused)
            p.set_forces(w0, Y, 3.0)
            p.objective_and_gradient(np.full(M, 1e-3))
        p.set_option(1, 1)
        p.set_forces(w0, Y, 3.0)
        p.opt_lbfgs(np.zeros(M), max_iterations=4)
        p.theta_scan([10.0, 1.0], method=FORCES, max_iterations=3)
        w, _ = p.weights(np.zeros(M))
        p.forces_from_weights(w)
    with bioen_b200.Problem(shape=(M, N)) as p:
        p.generate(1, 0, np.zeros(M), 1.0)
        p.download(0, 2, 0, 5)
print("sanitize_smoke done")

# ---- round 2: persistent kernel on/off, column pass with the gradient epilogue (wide, odd N), coefficient-space
# L-BFGS update, row-affine transform, fp32 storage, zero-copy gradient into pinned memory, in-process group
import ctypes as C
from bioen_b200 import _lib, dist as D
for (M, N) in ((3, 76001), (40, 80003)):
    y = rng.standard_normal((M, N))
    Y = rng.standard_normal(M)
    G = 0.1 * rng.standard_normal(N)
    with bioen_b200.Problem(y) as p:
        for persistent in (0, 1):
            p.set_option(5, persistent)
            p.set_logw(G, Y, 3.0)
            p.objective_and_gradient(G + 0.1)
            p.objective(G)
            p.gradient(G)
            p.set_option(6, 1)
            p.opt_lbfgs(G, max_iterations=8)
            p.set_option(6, 0)
            p.set_option(1, 0)
            p.set_forces(np.full(N, 1.0 / N), Y, 3.0)
            p.objective_and_gradient(np.full(M, 1e-3))
        p.set_option(5, 0)
        p.set_logw(G, Y, 3.0)
        gpin = bioen_b200.pinned_empty(N)
        f = C.c_double()
        x = np.ascontiguousarray(G + 0.05)
        _lib.check(_lib.load().bioen_b200_eval(p._ctx, 0, _lib.ptr(x), C.byref(f), _lib.ptr(gpin)), "eval")
        p.affine_rows(np.full(M, 1.5), np.full(M, 0.25))
        p.set_option(7, 1)
        p.set_logw(G, Y, 3.0)
        p.objective_and_gradient(G + 0.1)
        p.set_option(5, 1)
        p.objective_and_gradient(G + 0.1)
os.environ["BIOEN_B200_P2P_TIMEOUT_S"] = "60"
y = rng.standard_normal((300, 3001))
Y = rng.standard_normal(300)
w0 = np.full(3001, 1.0 / 3001)
with D.LocalGroup(y, 2) as grp:
    grp.call(lambda r, p, lo, hi: p.set_logw(np.zeros(hi - lo), Y, 2.0))
    grp.call(lambda r, p, lo, hi: p.objective_and_gradient(np.zeros(hi - lo)))
    grp.call(lambda r, p, lo, hi: p.opt_lbfgs(np.zeros(hi - lo), max_iterations=4))
    grp.call(lambda r, p, lo, hi: p.set_forces(w0[lo:hi], Y, 2.0))
    grp.call(lambda r, p, lo, hi: p.objective_and_gradient(np.full(300, 1e-3)))
print("sanitize_smoke round 2 done")
