"""Diagnostic (GPU box): fixed overheads at the ala5 size (M=28, N=50001)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bioen_b200
from bioen_b200 import optimize
from bioen_b200.optimize.ext import c_bioen
from oracle import oracle as O
M, N = 28, 50001
P = O.synthetic_problem(M, N, seed=12345)
cfg = optimize.minimize.Parameters("lbfgs"); cfg["verbose"] = False
def t(label, fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    print("%-46s first %.4f  best %.4f s" % (label, ts[0], min(ts)), flush=True)
    return r
def create():
    p = bioen_b200.Problem(P["yTilde"]); p.close()
t("Problem create+upload+destroy", create)
t("c_bioen.bioen_opt_lbfgs_forces (stateless)", lambda: c_bioen.bioen_opt_lbfgs_forces(np.zeros(M), P["w0"], P["yTilde"], P["YTilde"], 10.0, cfg))
t("c_bioen.bioen_opt_lbfgs_logw (stateless)", lambda: c_bioen.bioen_opt_lbfgs_logw(np.zeros(N), P["G"], P["yTilde"], P["YTilde"], 10.0, cfg))
t("forces.find_optimum", lambda: optimize.forces.find_optimum(P["forces_init"], P["w0"], P["y"], P["yTilde"], P["YTilde"], 10.0, cfg))
p = bioen_b200.Problem(P["yTilde"])
t("resident set_forces", lambda: p.set_forces(P["w0"], P["YTilde"], 10.0))
r = t("resident opt_lbfgs forces", lambda: p.opt_lbfgs(np.zeros(M)))
print(r[2], r[3])
t("resident objective_and_gradient forces", lambda: p.objective_and_gradient(np.zeros(M)), 20)
t("resident set_logw", lambda: p.set_logw(P["G"], P["YTilde"], 10.0))
r = t("resident opt_lbfgs logw", lambda: p.opt_lbfgs(np.zeros(N)))
print(r[2], r[3])
t("resident objective_and_gradient logw", lambda: p.objective_and_gradient(np.zeros(N)), 20)
from oracle import ref
if ref.available():
    ref.set_fast_openmp_flag(1)
    t("reference CPU opt_lbfgs_forces", lambda: ref.opt_lbfgs_forces(np.zeros(M), P["w0"], P["yTilde"], P["YTilde"], 10.0))
    t("reference CPU opt_lbfgs_logw", lambda: ref.opt_lbfgs_logw(np.zeros(N), P["G"], P["yTilde"], P["YTilde"], 10.0))
