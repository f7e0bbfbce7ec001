import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bioen_b200
from bioen_b200.problem import FORCES
from oracle import oracle as O
M, N = 28, 5001
P = O.synthetic_problem(M, N, seed=1)
rng = np.random.default_rng(0)
for name, w0 in (("uniform", np.full(N, 1.0 / N)), ("random", (lambda a: a / a.sum())(rng.random(N) + 0.2))):
    with bioen_b200.Problem(P["yTilde"]) as p:
        p.set_forces(w0, P["YTilde"], 1.0)
        for th in (300.0, 10.0):
            for mi in (1, 2, 5, 0):
                X, fmin, codes, info = p.theta_scan([th], method=FORCES, max_iterations=mi if mi else 5000)
                p.set_theta(th)
                x1, f1, c1, i1 = p.opt_lbfgs(np.zeros(M), max_iterations=mi if mi else 5000)
                print(name, th, "max_it", mi, "scan", codes[0], fmin[0], info["iterations"][0], info["evaluations"][0], "| single", c1, f1, i1["iterations"], i1["evaluations"], "| dx", np.max(np.abs(X[0] - x1)))
