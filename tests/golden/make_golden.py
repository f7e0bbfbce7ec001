"""Regenerates tests/golden/*.npz.  Runs ONLY where /root/reference exists (the build container).

For every fixture of the reference's own test-suite (/root/reference/test/optimize/data/*.pkl, the pickled
twins of the .h5 files its tests load -- h5py is not installed here) it stores
  * the inputs (GInit/G/y/yTilde/YTilde/w0/theta  or  forces_init/w0/y/yTilde/YTilde/theta),
  * the reference's golden scalar from the matching `.ref` file (the value its tests compare against at 10 %),
  * outputs of the UNMODIFIED reference C code (oracle/_ref/libbioen_ref.so, built by oracle/Makefile from
    the sources under /root/reference): objective + gradient at the start point and at a seeded perturbed
    point, and the end point (x, fmin, return code) of every minimiser the reference offers in C
    (liblbfgs with its 4 line searches, the 5 GSL multimin algorithms) with BioEn's default parameters.
A seeded synthetic "generic data" problem (SURVEY.md section 8d recipe) evaluated by the reference is added
as `synthetic_*.npz`.

Usage:  python tests/golden/make_golden.py
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402
from oracle import ref  # noqa: E402

DATA = "/root/reference/test/optimize/data/"
LOGW = ["data_16x15", "data_deer_test_logw_M808xN10", "data_potra_part_2_logw_M205xN10",
        "data_potra_part_1_logw_M808xN80", "data_potra_part_2_logw_M808xN10"]
FORCES = ["data_deer_test_forces_M808xN10", "data_forces_M64xN64"]
GSL_ALGS = ["conjugate_fr", "conjugate_pr", "bfgs2", "bfgs", "steepest_descent"]


def load(name):
    with open(DATA + name + ".pkl", "rb") as fh:
        return [np.asarray(a, dtype=np.float64) if hasattr(a, "shape") else float(a)
                for a in pickle.load(fh, encoding="latin1")]


def golden_scalar(name):
    with open(DATA + name + ".ref", "rb") as fh:
        return float(pickle.load(fh, encoding="latin1"))


def minimisers(kind, x0, a, yT, YT, theta):
    """End points of every C minimiser in the reference's reproducible mode (fast_openmp=0), plus -- as
    `*_fmin_alt` / `*_code_alt` -- the same run in its fast mode (fast_openmp=1, 3 threads, transposed cache):
    same arithmetic, different summation order.  |fmin - fmin_alt| is the reference's OWN end-point noise for
    that problem/minimiser; the GPU tests use it to widen the 1e-8 bar where the trajectory is chaotic."""
    out = {}
    lb = ref.opt_lbfgs_logw if kind == "logw" else ref.opt_lbfgs_forces
    gs = ref.opt_gsl_logw if kind == "logw" else ref.opt_gsl_forces
    for ls in range(4):
        x, fmin, code = lb(x0, a, yT, YT, theta, linesearch=ls)
        out["lbfgs%d_x" % ls], out["lbfgs%d_fmin" % ls], out["lbfgs%d_code" % ls] = x, fmin, code
    for alg in GSL_ALGS:
        x, fmin, code = gs(x0, a, yT, YT, theta, algorithm=alg)
        out["gsl_%s_x" % alg], out["gsl_%s_fmin" % alg], out["gsl_%s_code" % alg] = x, fmin, code
    ref.set_fast_openmp_flag(1)
    ref.set_num_threads(3)
    for ls in range(4):
        x, fmin, code = lb(x0, a, yT, YT, theta, linesearch=ls, caching=True)
        out["lbfgs%d_fmin_alt" % ls], out["lbfgs%d_code_alt" % ls] = fmin, code
    for alg in GSL_ALGS:
        x, fmin, code = gs(x0, a, yT, YT, theta, algorithm=alg, caching=True)
        out["gsl_%s_fmin_alt" % alg], out["gsl_%s_code_alt" % alg] = fmin, code
    ref.set_fast_openmp_flag(0)
    ref.set_num_threads(8)
    return out


def main():
    ref.set_fast_openmp_flag(0)          # the reference's reproducible mode
    rng = np.random.default_rng(20261018)
    for name in LOGW:
        GInit, G, y, yT, YT, w0, theta = load(name)
        g1 = GInit.ravel() + 0.1 * rng.standard_normal(GInit.size)
        d = dict(kind="logw", GInit=GInit, G=G, y=y, yTilde=yT, YTilde=YT, w0=w0, theta=theta,
                 ref_scalar=golden_scalar(name), probe=g1,
                 f_init=ref.logw_objective(GInit, G, yT, YT, theta),
                 grad_init=ref.logw_gradient(GInit, G, yT, YT, theta),
                 f_probe=ref.logw_objective(g1, G, yT, YT, theta),
                 grad_probe=ref.logw_gradient(g1, G, yT, YT, theta),
                 w_probe=ref.logw_weights(g1)[0])
        d.update(minimisers("logw", GInit, G, yT, YT, theta))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, d["f_init"], d["lbfgs2_fmin"], d["lbfgs2_code"])
    for name in FORCES:
        fi, w0, y, yT, YT, theta = load(name)
        f1 = fi.ravel() + 1e-3 * rng.standard_normal(fi.size)
        d = dict(kind="forces", forces_init=fi, w0=w0, y=y, yTilde=yT, YTilde=YT, theta=theta,
                 ref_scalar=golden_scalar(name), probe=f1,
                 f_init=ref.forces_objective(fi, w0, yT, YT, theta),
                 grad_init=ref.forces_gradient(fi, w0, yT, YT, theta),
                 f_probe=ref.forces_objective(f1, w0, yT, YT, theta),
                 grad_probe=ref.forces_gradient(f1, w0, yT, YT, theta),
                 w_probe=ref.forces_weights(f1, w0, yT))
        d.update(minimisers("forces", fi, w0, yT, YT, theta))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, d["f_init"], d["lbfgs2_fmin"], d["lbfgs2_code"])
    # synthetic generic-data problem: inputs are regenerated from the seed at test time, only the reference's
    # outputs are stored.
    for (M, N, theta) in [(100, 20000, 10.0), (37, 5001, 1.0)]:
        P = O.synthetic_problem(M, N, seed=12345)
        g1 = 0.1 * np.random.default_rng(1).standard_normal(N)
        f1 = 1e-3 * np.random.default_rng(2).standard_normal(M)
        d = dict(M=M, N=N, theta=theta, seed=12345,
                 logw_f=ref.logw_objective(g1, P["G"], P["yTilde"], P["YTilde"], theta),
                 logw_grad=ref.logw_gradient(g1, P["G"], P["yTilde"], P["YTilde"], theta),
                 forces_f=ref.forces_objective(f1, P["w0"], P["yTilde"], P["YTilde"], theta),
                 forces_grad=ref.forces_gradient(f1, P["w0"], P["yTilde"], P["YTilde"], theta))
        for ls in (0, 2):
            x, fmin, code = ref.opt_lbfgs_logw(P["GInit"], P["G"], P["yTilde"], P["YTilde"], theta, linesearch=ls)
            d["logw_lbfgs%d_x" % ls], d["logw_lbfgs%d_fmin" % ls], d["logw_lbfgs%d_code" % ls] = x, fmin, code
            x, fmin, code = ref.opt_lbfgs_forces(P["forces_init"], P["w0"], P["yTilde"], P["YTilde"], theta,
                                                 linesearch=ls)
            d["forces_lbfgs%d_x" % ls], d["forces_lbfgs%d_fmin" % ls], d["forces_lbfgs%d_code" % ls] = x, fmin, code
        x, fmin, code = ref.opt_gsl_logw(P["GInit"], P["G"], P["yTilde"], P["YTilde"], theta)
        d["logw_gsl_bfgs2_x"], d["logw_gsl_bfgs2_fmin"], d["logw_gsl_bfgs2_code"] = x, fmin, code
        x, fmin, code = ref.opt_gsl_forces(P["forces_init"], P["w0"], P["yTilde"], P["YTilde"], theta)
        d["forces_gsl_bfgs2_x"], d["forces_gsl_bfgs2_fmin"], d["forces_gsl_bfgs2_code"] = x, fmin, code
        np.savez_compressed(os.path.join(HERE, "synthetic_M%dxN%d.npz" % (M, N)), **d)
        print("synthetic", M, N, d["logw_f"], d["logw_lbfgs2_fmin"], d["logw_lbfgs2_code"],
              d["forces_lbfgs2_fmin"], d["forces_lbfgs2_code"])


if __name__ == "__main__":
    main()
