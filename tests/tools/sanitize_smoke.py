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
