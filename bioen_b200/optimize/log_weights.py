"""Log-weights method: public API mirror of bioen/optimize/log_weights.py, GPU backed.

`find_optimum` has the reference's signature, cfg handling, return tuple and error behaviour
(log_weights.py:409-621); underneath, yTilde is uploaded to the GPU once and everything -- the initial
objective, the whole minimisation (device-resident L-BFGS / GSL-style minimisers, or SciPy calling GPU
objective/gradient closures), the optimal weights and the averages -- runs on it.
"""
import time

import numpy as np
import scipy.optimize as sopt

from . import common
from .ext import c_bioen
from .. import _lib
from ..problem import GSL_OK, LBFGS_OK, LOGW, Problem, gsl_strerror, lbfgs_strerror


# ---- legacy host helpers (part of the reference's public surface; NumPy, small inputs) ------------------------
def getWeights(g):
    """w = exp(g)/sum exp(g), s = sum exp(g); g is (n,1) (log_weights.py:93-110, un-stabilised like the C)."""
    tmp = np.exp(g)
    s = tmp.sum()
    return np.array(tmp / s), s


def getGs(w):
    """Log-weights with the last one pinned to 0 (log_weights.py:113-127); w is (n,1)."""
    g = np.log(w)
    g -= g[-1, 0]
    return g


def init_log_weights(w0):
    """log_weights.py:71-90"""
    G = getGs(w0)
    GInit = getGs(np.array(w0))
    g = GInit.copy()
    gPrime = np.asarray(g[:-1].T)[0]
    return gPrime, g, G, GInit


def getWOpt(G, gPrimeOpt):
    """(n,1) weights for the optimal log-weights (log_weights.py:130-161)."""
    wopt, _ = getWeights(np.asarray(gPrimeOpt, dtype=np.float64).reshape(-1, 1))
    return wopt


def bioen_log_prior(w, s, g, G, theta):
    """theta * (g.w - G.w - log s + log s0) (log_weights.py:18-68)."""
    w, g, G = _lib.vec(w), _lib.vec(g), _lib.vec(G)
    s0 = np.exp(G).sum()
    return float(theta * (g @ w - G @ w - np.log(s) + np.log(s0)))


def bioen_log_posterior_base(gPrime, g, G, yTilde, YTilde, theta):
    """Legacy NumPy objective (log_weights.py:289-329).  Like the reference it stores gPrime into g[:, 0]."""
    g[:, 0] = np.asarray(gPrime, dtype=np.float64).ravel()
    w, s = getWeights(np.asarray(g))
    return bioen_log_prior(w, s, g, G, theta) + common.chiSqrTerm(w, yTilde, YTilde)


def grad_bioen_log_posterior_base(gPrime, g, G, yTilde, YTilde, theta):
    """Legacy NumPy gradient, vectorised (log_weights.py:332-406).  The prior term follows the C kernels
    (c_bioen_kernels_logw.c:214-217: g - <g> - G + <G>); the reference's Python twin has `-(G + <G>)` at
    log_weights.py:389, which is the same whenever <G> = 0 (all reference fixtures)."""
    gp = np.asarray(gPrime, dtype=np.float64).ravel()
    Gv = _lib.vec(G)
    yT = np.asarray(yTilde, dtype=np.float64)
    w, _ = getWeights(gp)
    avg = yT @ w
    r = avg - _lib.vec(YTilde)
    back = yT.T @ r - r @ avg
    return w * theta * (gp - gp @ w - Gv + Gv @ w) + w * back


def grad_chiSqrTerm(gPrime, g, G, yTilde, YTilde, theta):
    """Gradient of the chi^2 term w.r.t. the first n-1 log-weights, last pinned to 0 (log_weights.py:164-188)."""
    g[:-1, 0] = np.asarray(gPrime, dtype=np.float64).ravel()
    g[-1, 0] = 0
    w, _ = getWeights(np.asarray(g))
    w = w.ravel()
    yT = np.asarray(yTilde, dtype=np.float64)
    avg = yT @ w
    r = avg - _lib.vec(YTilde)
    return (w * (yT.T @ r - r @ avg))[:-1]


def check_params_logweights(GInit, G, y, yTilde, YTilde):
    """Shapes: GInit, G (n,1); y, yTilde (m,n); YTilde (1,m); ValueError otherwise (log_weights.py:191-233)."""
    m, n = yTilde.shape
    error = False
    for name, arr, expected in (("GInit", GInit, (n, 1)), ("G", G, (n, 1)), ("y", y, (m, n)),
                                ("YTilde", YTilde, (1, m))):
        if arr.shape != expected:
            print("Unexpected shape for variable: {}\nExpected: {}\nCurrent:  {}".format(name, expected, arr.shape))
            error = True
    if error:
        raise ValueError("arguments dimensionality for the 'log_weights' method are wrong")


# ---- selectors (log_weights.py:237-286) -------------------------------------------------------------------
def bioen_log_posterior(gPrime, g, G, yTilde, YTilde, theta, use_c=True, caching=False):
    if use_c:
        return c_bioen.bioen_log_posterior_logw(gPrime, g, G, yTilde, YTilde, theta, caching=caching)
    return bioen_log_posterior_base(gPrime, g, G, yTilde, YTilde, theta)


def grad_bioen_log_posterior(gPrime, g, G, yTilde, YTilde, theta, use_c=True, caching=False):
    if use_c:
        return c_bioen.grad_bioen_log_posterior_logw(gPrime, g, G, yTilde, YTilde, theta, caching=caching)
    return grad_bioen_log_posterior_base(gPrime, g, G, yTilde, YTilde, theta)


# ---- SciPy dispatch shared by both methods -----------------------------------------------------------------
def _run_scipy(cfg, f, fprime, x0, label):
    alg = cfg["algorithm"].lower()
    p = cfg["params"]
    if alg in ("lbfgs", "fmin_l_bfgs_b"):
        common.print_highlighted("method L-BFGS", cfg["verbose"])
        kw = dict(fprime=fprime, epsilon=p["epsilon"], pgtol=p["pgtol"], maxiter=p["max_iterations"])
        try:
            return sopt.fmin_l_bfgs_b(f, x0, disp=cfg["verbose"], **kw)      # as the reference calls it
        except TypeError:                                                    # SciPy >= 1.18 dropped `disp`
            return sopt.fmin_l_bfgs_b(f, x0, **kw)
    if alg in ("bfgs", "fmin_bfgs"):
        common.print_highlighted("method BFGS", cfg["verbose"])
        return sopt.fmin_bfgs(f, x0, fprime=fprime, epsilon=p["epsilon"], gtol=p["gtol"],
                              maxiter=p["max_iterations"], disp=cfg["verbose"], full_output=True)
    if alg in ("cg", "fmin_cg"):
        common.print_highlighted("method CG", cfg["verbose"])
        return sopt.fmin_cg(f, x0, fprime=fprime, epsilon=p["epsilon"], gtol=p["gtol"],
                            maxiter=p["max_iterations"], disp=cfg["verbose"], full_output=True)
    raise RuntimeError("Method '" + cfg["algorithm"] + "' not recognized for scipy/" + label +
                       " library (valid values =  'lbfgs', 'bfgs', 'cg' ) ")


def _lbfgs_kwargs(cfg):
    return {k: cfg["params"][k] for k in ("linesearch", "max_iterations", "delta", "epsilon", "ftol", "gtol",
                                          "wolfe", "past", "max_linesearch")}


def _gsl_kwargs(cfg):
    return dict(algorithm=c_bioen.get_gsl_method(cfg["algorithm"]), step_size=cfg["params"]["step_size"],
                tol=cfg["params"]["tol"], max_iterations=cfg["params"]["max_iterations"])


def _minimize_on_device(problem, method, x0, cfg, tag):
    """lbfgs / gsl / scipy dispatch on a resident problem; returns (xopt, fmin_final).  Error convention of
    c_bioen.pyx:432-438 / 516-520: RuntimeError naming the '... return code', result discarded."""
    minimizer = cfg["minimizer"].upper()
    if minimizer in ("LIBLBFGS", "LBFGS"):
        common.print_highlighted(tag + " -- device L-BFGS (liblbfgs semantics)", cfg["verbose"])
        x, fmin, code, _ = problem.opt_lbfgs(x0, method, verbose=cfg["verbose"], **_lbfgs_kwargs(cfg))
        if code not in LBFGS_OK:
            raise RuntimeError("{}, liblbfgs return code: {}:{}".format(
                "bioen_opt_lbfgs_" + tag.lower(), code, lbfgs_strerror(code)))
        return x, fmin
    if minimizer == "GSL":
        common.print_highlighted(tag + " -- device GSL multimin", cfg["verbose"])
        x, fmin, code, _ = problem.opt_gsl(x0, method, verbose=cfg["verbose"], **_gsl_kwargs(cfg))
        if code not in GSL_OK:
            raise RuntimeError("{}, GSL return code: {}:{}".format(
                "bioen_opt_bfgs_" + tag.lower(), code, gsl_strerror(code)))
        if code == c_bioen.gsl_continue:
            print(c_bioen.gsl_continue_msg)
        return x, fmin
    if minimizer == "SCIPY":
        common.print_highlighted(tag + " -- Library scipy/GPU", cfg["verbose"])
        res = _run_scipy(cfg, lambda x: problem.objective(x, method), lambda x: problem.gradient(x, method),
                         _lib.vec(x0), "c")
        return np.asarray(res[0]), float(res[1])
    raise RuntimeError("Library " + cfg["minimizer"] +
                       " not recognized (valid values =  'LIBLBFGS', 'GSL', 'scipy', 'scipy' ) ")


def find_optimum(GInit, G, y, yTilde, YTilde, theta, cfg, problem=None):
    """Minimise the BioEn log-posterior over the log-weights (log_weights.py:409-621).

    Returns (wopt (n,1), yopt (m,), gopt (n,), fmin_initial, fmin_final).  `problem` (optional, not in the
    reference) is a bioen_b200.Problem that already holds yTilde, e.g. to reuse one upload over a theta series.
    """
    check_params_logweights(GInit, G, y, yTilde, YTilde)
    caching = cfg["cache_ytilde_transposed"]
    if caching == "auto":
        caching = common.set_caching_heuristics(yTilde.shape[0], yTilde.shape[1])
    cfg["cache_ytilde_transposed"] = caching

    gPrime = _lib.vec(GInit).copy()
    minimizer = cfg["minimizer"].upper()
    legacy = minimizer == "SCIPY" and cfg["use_c_functions"] is False
    own = problem is None
    if own:
        problem = Problem(yTilde)
    try:
        problem.set_logw(G, YTilde, theta)
        fmin_initial = problem.objective(gPrime, LOGW)
        if cfg["verbose"]:
            print("fmin_initial", fmin_initial)
        start = time.time()
        if legacy:
            common.print_highlighted("LOGW -- Library scipy/PY", cfg["verbose"])
            g = np.array(GInit, dtype=np.float64).reshape(-1, 1)
            res = _run_scipy(cfg, lambda x: bioen_log_posterior_base(x, g, G, yTilde, YTilde, theta),
                             lambda x: grad_bioen_log_posterior_base(x, g, G, yTilde, YTilde, theta), gPrime, "py")
            gopt, fmin_final = np.asarray(res[0]), float(res[1])
        else:
            gopt, fmin_final = _minimize_on_device(problem, LOGW, gPrime, cfg, "LOGW")
        if cfg["verbose"]:
            print("time elapsed ", time.time() - start)
        w, _ = problem.weights(gopt, LOGW)
        wopt = w.reshape(-1, 1)
        yopt = problem.average(w) if y is yTilde else common.average_like(problem, w, y, keep=not own)
    finally:
        if own:
            problem.close()
    if cfg["verbose"]:
        print("========================")
        print("fmin_initial  = ", fmin_initial)
        print("fmin_final    = ", fmin_final)
        print("========================")
    return wopt, yopt, gopt, fmin_initial, fmin_final


def find_optimum_series(GInit, G, y, yTilde, YTilde, thetas, cfg, batched=True, problem=None, strict=True):
    """The theta series (L-curve) of the reference's callers -- the loop of bioen/analyze/procedure.py:62-83 and
    of the ala5 notebook's run_theta_series -- on ONE resident copy of yTilde.  Not part of the reference API.

    batched=True (minimizer 'lbfgs' only): up to 32 theta values are minimised together from GInit by lockstep
    device L-BFGS machines whose evaluations are fp64 tensor-core skinny GEMMs (yTilde streamed once per pass
    for all of them).  batched=False: one find_optimum per theta, warm-started from the previous optimum like
    the reference's callers do.  Returns a list of find_optimum 5-tuples, one per theta, in input order.  strict=False: a theta whose
    minimisation ends with a liblbfgs error code (the reference raises RuntimeError and discards the result; at
    large theta the line search can fail at rounding level, code -998) yields None instead of raising.
    """
    check_params_logweights(GInit, G, y, yTilde, YTilde)
    thetas = [float(t) for t in np.asarray(thetas, dtype=np.float64).ravel()]
    own = problem is None
    if own:
        problem = Problem(yTilde)
    out = []
    try:
        if batched and cfg["minimizer"].upper() in ("LIBLBFGS", "LBFGS"):
            problem.set_logw(G, YTilde, thetas[0] if thetas else 0.0)
            g0 = _lib.vec(GInit)
            yprob = problem if y is yTilde else problem.like(y)
            try:
                for lo in range(0, len(thetas), 32):
                    chunk = thetas[lo:lo + 32]
                    X, fmin, codes, _ = problem.theta_scan(chunk, x0=g0, method=LOGW, verbose=cfg["verbose"], **_lbfgs_kwargs(cfg))
                    for q, th in enumerate(chunk):
                        if codes[q] not in LBFGS_OK:
                            if not strict:
                                out.append(None)
                                continue
                            raise RuntimeError("{}, liblbfgs return code: {}:{}".format(
                                "bioen_opt_lbfgs_logw", codes[q], lbfgs_strerror(codes[q])))
                        problem.set_theta(th)
                        fmin_initial = problem.objective(g0, LOGW)
                        w, _ = problem.weights(X[q], LOGW)
                        out.append((w.reshape(-1, 1), yprob.average(w), X[q].copy(), fmin_initial, float(fmin[q])))
            finally:
                if yprob is not problem:
                    yprob.close()
        else:
            g = GInit
            for th in thetas:
                try:
                    res = find_optimum(g, G, y, yTilde, YTilde, th, cfg, problem=problem)
                except RuntimeError:
                    if strict:
                        raise
                    out.append(None)      # as a caller of the reference would: skip, keep the last good start
                    continue
                out.append(res)
                g = res[2].reshape(-1, 1)
    finally:
        if own:
            problem.close()
    return out
