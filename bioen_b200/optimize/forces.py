"""Forces method: public API mirror of bioen/optimize/forces.py, GPU backed.

w_j ~ w0_j * exp(+ sum_i f_i yTilde_ij)  (sign as in c_bioen_kernels_forces.c:144,158 and forces.py:107-110).
"""
import time

import numpy as np

from . import common
from .ext import c_bioen
from .log_weights import _lbfgs_kwargs, _minimize_on_device, _run_scipy
from .. import _lib
from ..problem import FORCES, LBFGS_OK, Problem, lbfgs_strerror


# ---- synthetic "generic data" generators (forces.py:19-68) -------------------------------------------------
def gen_synthetic_data(M, N, YTrue, sig_exp):
    """Observations YObs ~ N(YTrue, sig_exp) and YTilde = YObs / sig_exp."""
    YObs = np.array(np.random.normal(YTrue, sig_exp))
    return YObs, YObs / sig_exp


def gen_sythetic_ensemble(M, N, YTrue, sig_exp, sig_sim):
    """y (M x N) ~ N(YTrue_i, sig_sim) and yTilde = y / sig_exp_i."""
    y = np.random.normal(np.asarray(YTrue, dtype=np.float64)[:, None], sig_sim, size=(len(YTrue), N))
    yTilde = y / np.asarray(sig_exp, dtype=np.float64)[:, None]
    return y, yTilde


def init_forces(M, val=0):
    """(M,1) array filled with val (forces.py:71-88)."""
    forces = np.zeros((M, 1))
    forces[:, 0] = val
    return forces


# ---- legacy host helpers -----------------------------------------------------------------------------------
def get_weights_from_forces(w0, y, forces):
    """(n,1) weights, NumPy (forces.py:91-113)."""
    f = _lib.vec(forces)
    x = f @ np.asarray(y, dtype=np.float64)
    e = _lib.vec(w0) * np.exp(x - x.max())
    return (e / e.sum()).reshape(-1, 1)


def _kl_and_chi2(w, w0, yTilde, YTilde):
    w, w0 = _lib.vec(w), _lib.vec(w0)
    ind = w > 0
    S = float(np.log(w[ind] / w0[ind]) @ w[ind])
    return S, common.chiSqrTerm(w, yTilde, YTilde)


def bioen_chi2_s_forces(forces, w0, yTilde, YTilde):
    """(S, chiSqr) with S = sum_j w_j log(w_j / w0_j) over w_j > 0 (forces.py:116-136)."""
    w = get_weights_from_forces(w0, yTilde, forces)
    return _kl_and_chi2(w, w0, yTilde, YTilde)


def check_params_forces(forcesInit, w0, y, yTilde, YTilde):
    """Shapes: forcesInit (m,1); w0 (n,1); y, yTilde (m,n); YTilde (1,m); ValueError otherwise
    (forces.py:139-183)."""
    m, n = yTilde.shape
    error = False
    for name, arr, expected in (("forcesInit", forcesInit, (m, 1)), ("w0", w0, (n, 1)), ("y", y, (m, n)),
                                ("YTilde", YTilde, (1, m))):
        if arr.shape != expected:
            print("Unexpected shape for variable: {}\nExpected: {}\nCurrent:  {}".format(name, expected, arr.shape))
            error = True
    if error:
        raise ValueError("arguments dimensionality for the 'forces' method are wrong")


def bioen_log_posterior(forces, w0, y, yTilde, YTilde, theta, use_c=True, caching=False):
    """Selector (forces.py:187-213)."""
    if use_c:
        return c_bioen.bioen_log_posterior_forces(forces, w0, yTilde, YTilde, theta)
    return bioen_log_posterior_base(forces, w0, yTilde, YTilde, theta)


def grad_bioen_log_posterior(forces, w0, y, yTilde, YTilde, theta, use_c=True, caching=False):
    """Selector (forces.py:216-243)."""
    if use_c:
        return c_bioen.grad_bioen_log_posterior_forces(forces, w0, yTilde, YTilde, theta)
    return grad_bioen_log_posterior_base(forces, w0, yTilde, YTilde, theta)


def bioen_log_posterior_base(forces, w0, yTilde, YTilde, theta, use_c=True):
    """Legacy NumPy objective (forces.py:246-289)."""
    S, chi2 = bioen_chi2_s_forces(forces, w0, yTilde, YTilde)
    return theta * S + chi2


def grad_bioen_log_posterior_base(forces, w0, yTilde, YTilde, theta, use_c=True):
    """Legacy NumPy gradient (forces.py:292-333)."""
    yT = np.asarray(yTilde, dtype=np.float64)
    w = get_weights_from_forces(w0, yT, forces).ravel()
    w0v = _lib.vec(w0)
    avg = yT @ w
    B = yT.T @ (avg - _lib.vec(YTilde))
    ratio = np.where(w > 0, w / w0v, 1.0)
    E = ((np.log(ratio) + 1.0) * theta + B) * w
    return (yT - avg[:, None]) @ E


def find_optimum(forcesInit, w0, y, yTilde, YTilde, theta, cfg, problem=None):
    """Minimise the BioEn log-posterior over the generalised forces (forces.py:336-548).

    Returns (wopt (n,1), yopt (m,), forces_opt (m,), fmin_initial, fmin_final, chiSqr, S).  `problem`
    (optional, not in the reference) is a bioen_b200.Problem that already holds yTilde.
    """
    check_params_forces(forcesInit, w0, y, yTilde, YTilde)
    caching = cfg["cache_ytilde_transposed"]
    if caching == "auto":
        caching = common.set_caching_heuristics(yTilde.shape[0], yTilde.shape[1])
    cfg["cache_ytilde_transposed"] = caching

    f0 = _lib.vec(forcesInit).copy()
    minimizer = cfg["minimizer"].upper()
    legacy = minimizer == "SCIPY" and cfg["use_c_functions"] is False
    own = problem is None
    if own:
        problem = Problem(yTilde)
    try:
        problem.set_forces(w0, YTilde, theta)
        fmin_initial = problem.objective(f0, FORCES)
        if cfg["verbose"]:
            print("fmin_initial", fmin_initial)
        start = time.time()
        if legacy:
            common.print_highlighted("FORCES -- Library scipy/PY", cfg["verbose"])
            res = _run_scipy(cfg, lambda x: bioen_log_posterior_base(x, w0, yTilde, YTilde, theta),
                             lambda x: grad_bioen_log_posterior_base(x, w0, yTilde, YTilde, theta), f0, "py")
            forces_opt, fmin_final = np.asarray(res[0]), float(res[1])
        else:
            forces_opt, fmin_final = _minimize_on_device(problem, FORCES, f0, cfg, "FORCES")
        if cfg["verbose"]:
            print("time elapsed ", time.time() - start)
        w, _ = problem.weights(forces_opt, FORCES)
        wopt = w.reshape(-1, 1)
        yavg = problem.average(w)
        yopt = yavg if y is yTilde else common.average_like(problem, w, y, keep=not own)
        # S and chi^2 at the optimum (forces.py:546): KL over w_j > 0, chi^2 from the resident yTilde
        ind = w > 0
        S = float(np.log(w[ind] / _lib.vec(w0)[ind]) @ w[ind])
        r = yavg - _lib.vec(YTilde)
        chiSqr = 0.5 * float(r @ r)
    finally:
        if own:
            problem.close()
    if cfg["verbose"]:
        print("========================")
        print("fmin_initial  = ", fmin_initial)
        print("fmin_final    = ", fmin_final)
        print("========================")
    return wopt, yopt, forces_opt, fmin_initial, fmin_final, chiSqr, S


def find_optimum_series(forcesInit, w0, y, yTilde, YTilde, thetas, cfg, batched=True, problem=None, strict=True):
    """The theta series (L-curve) of the reference's callers (bioen/analyze/procedure.py:62-83; the ala5 notebook
    runs exactly this with the forces method and liblbfgs) on ONE resident copy of yTilde.  Not in the reference API.

    batched=True (minimizer 'lbfgs'): up to 32 theta values are minimised together from forcesInit (lockstep
    device L-BFGS, four tensor-core GEMMs per batched evaluation).  batched=False: one find_optimum per theta,
    warm-started from the previous optimum.  Returns a list of find_optimum 7-tuples in input order.  strict=False:
    a theta whose minimisation ends with a liblbfgs error code yields None instead of raising RuntimeError.
    """
    check_params_forces(forcesInit, w0, y, yTilde, YTilde)
    thetas = [float(t) for t in np.asarray(thetas, dtype=np.float64).ravel()]
    own = problem is None
    if own:
        problem = Problem(yTilde)
    out = []
    try:
        if batched and cfg["minimizer"].upper() in ("LIBLBFGS", "LBFGS"):
            problem.set_forces(w0, YTilde, thetas[0] if thetas else 0.0)
            f0 = _lib.vec(forcesInit)
            w0v, Yv = _lib.vec(w0), _lib.vec(YTilde)
            yprob = problem if y is yTilde else problem.like(y)
            try:
                for lo in range(0, len(thetas), 32):
                    chunk = thetas[lo:lo + 32]
                    X, fmin, codes, _ = problem.theta_scan(chunk, x0=f0, method=FORCES, verbose=cfg["verbose"],
                                                           **_lbfgs_kwargs(cfg))
                    for q, th in enumerate(chunk):
                        if codes[q] not in LBFGS_OK:
                            if not strict:
                                out.append(None)
                                continue
                            raise RuntimeError("{}, liblbfgs return code: {}:{}".format(
                                "bioen_opt_lbfgs_forces", codes[q], lbfgs_strerror(codes[q])))
                        problem.set_theta(th)
                        fmin_initial = problem.objective(f0, FORCES)
                        w, _ = problem.weights(X[q], FORCES)
                        yavg = problem.average(w)
                        ind = w > 0
                        S = float(np.log(w[ind] / w0v[ind]) @ w[ind])
                        r = yavg - Yv
                        out.append((w.reshape(-1, 1), yavg if y is yTilde else yprob.average(w), X[q].copy(),
                                    fmin_initial, float(fmin[q]), 0.5 * float(r @ r), S))
            finally:
                if yprob is not problem:
                    yprob.close()
        else:
            f = forcesInit
            for th in thetas:
                try:
                    res = find_optimum(f, w0, y, yTilde, YTilde, th, cfg, problem=problem)
                except RuntimeError:
                    if strict:
                        raise
                    out.append(None)      # as a caller of the reference would: skip, keep the last good start
                    continue
                out.append(res)
                f = res[2].reshape(-1, 1)
    finally:
        if own:
            problem.close()
    return out
