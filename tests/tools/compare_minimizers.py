"""Diagnostic (GPU box): device minimisers vs the reference's stored end points and the oracle's restatement."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import FORCES_FIXTURES, LOGW_FIXTURES, load_golden  # noqa: E402
import bioen_b200  # noqa: E402
from bioen_b200.optimize.ext import c_bioen  # noqa: E402
from oracle import oracle as O  # noqa: E402

rel = lambda a, b: abs(a - b) / max(abs(b), 1e-300)
for name in LOGW_FIXTURES + FORCES_FIXTURES:
    d = load_golden(name)
    with bioen_b200.Problem(d["yTilde"]) as p:
        if d["kind"] == "logw":
            p.set_logw(d["G"], d["YTilde"], d["theta"]); x0 = d["GInit"].ravel()
            fg = lambda v: O.logw_fg(v, d["G"], d["yTilde"], d["YTilde"], d["theta"])
        else:
            p.set_forces(d["w0"], d["YTilde"], d["theta"]); x0 = d["forces_init"].ravel()
            fg = lambda v: O.forces_fg(v, d["w0"], d["yTilde"], d["YTilde"], d["theta"])
        for ls in range(4):
            x, fmin, code, info = p.opt_lbfgs(x0, linesearch=ls)
            r = O.lbfgs(fg, x0, linesearch=ls)
            print("%-36s lbfgs%d code gpu %5d ref %5d orc %5d | it gpu %4d orc %4d ev gpu %4d orc %4d | fmin gpu %.12g ref %.12g  rel(ref) %.2e rel(orc) %.2e"
                  % (name, ls, code, d["lbfgs%d_code" % ls], r["code"], info["iterations"], r["iterations"],
                     info["evaluations"], r["evaluations"], fmin, d["lbfgs%d_fmin" % ls],
                     rel(fmin, d["lbfgs%d_fmin" % ls]), rel(fmin, r["fx"])))
        for alg in ["conjugate_fr", "conjugate_pr", "bfgs2", "bfgs", "steepest_descent"]:
            x, fmin, code, info = p.opt_gsl(x0, algorithm=c_bioen.get_gsl_method(alg))
            r = O.gsl_minimize(fg, x0, algorithm=alg)
            print("%-36s gsl %-16s code gpu %3d ref %3d orc %3d | it gpu %4d orc %4d | fmin gpu %.12g ref %.12g rel(ref) %.2e rel(orc) %.2e"
                  % (name, alg, code, d["gsl_%s_code" % alg], r["code"], info["iterations"], r["iterations"], fmin,
                     d["gsl_%s_fmin" % alg], rel(fmin, d["gsl_%s_fmin" % alg]), rel(fmin, r["fx"])))
