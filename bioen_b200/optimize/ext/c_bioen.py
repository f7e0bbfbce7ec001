"""GPU-backed mirror of the reference's Cython module `bioen.optimize.ext.c_bioen`
(bioen/optimize/ext/c_bioen.pyx): the same 14 Python-callable functions, argument order, return values and
error behaviour, bound with ctypes to the reference-compatible symbols of libbioen_b200.so
(include/bioen_b200.h part 1).  Every call takes host NumPy arrays, exactly like the reference.

Unlike the pyx (c_bioen.pyx has no dtype / contiguity check and silently computes garbage for other layouts)
inputs are converted to C-contiguous float64 first.
"""
import ctypes as C
import sys
import time

import numpy as np

from ... import _lib

gsl_success = [0, -2, 27]          # c_bioen.pyx:112-116 (GSL_SUCCESS, GSL_CONTINUE, GSL_ENOPROG)
gsl_continue = -2
gsl_continue_msg = "GSL: the iteration has not converged yet; the current best point is returned"
lbfgs_success = [0, 1, 2]          # c_bioen.pyx:120


def _L():
    return _lib.load()


def set_fast_openmp_flag(flag):                      # c_bioen.pyx:176
    _L()._set_fast_openmp_flag(int(flag))


def get_fast_openmp_flag():                          # c_bioen.pyx:180
    return _L()._get_fast_openmp_flag()


def omp_set_num_threads(i):                          # c_bioen.pyx:184
    _L()._omp_set_num_threads(int(i))


_P = "gsl_multimin_fdfminimizer_"
_GSL_IDS = {"conjugate_fr": 0, _P + "conjugate_fr": 0, "conjugate_pr": 1, _P + "conjugate_pr": 1,
            "bfgs2": 2, _P + "vector_bfgs2": 2, "bfgs": 3, _P + "vector_bfgs": 3,
            "steepest_descent": 4, _P + "steepest_descent": 4}


def get_gsl_method(algorithm):                       # c_bioen.pyx:188-213
    if algorithm not in _GSL_IDS:
        raise RuntimeError("{}, GSL return code: {}:{}".format(
            "get_gsl_method", -1, " The algorithm " + str(algorithm) + " is not available."))
    return _GSL_IDS[algorithm]


def library_gsl():                                   # c_bioen.pyx:216
    return bool(_L()._library_gsl())


def library_lbfgs():                                 # c_bioen.pyx:230
    return bool(_L()._library_lbfgs())


def bioen_log_posterior_logw(gPrime, g, G, yTilde, YTilde, theta, caching=False):
    """c_bioen.pyx:246-292.  Faithful to the reference, the SECOND argument `g` is what the C objective
    receives as reference log-weights (c_bioen.pyx:278-279); `G` is unused here."""
    _lib.clear_pending()
    yT = _lib.mat(yTilde)
    m, n = yT.shape
    gp, gref, Y = _lib.vec(gPrime), _lib.vec(g), _lib.vec(YTilde)
    val = _L()._bioen_log_posterior_logw(_lib.ptr(gp), _lib.ptr(gref), _lib.ptr(yT), _lib.ptr(Y), None, None,
                                         float(theta), 0, None, None, None, m, n, 0.0)
    _lib.check_pending("bioen_log_posterior_logw")
    return val


def grad_bioen_log_posterior_logw(gPrime, g, G, yTilde, YTilde, theta, caching=False, print_timing=False):
    """c_bioen.pyx:295-359"""
    _lib.clear_pending()
    yT = _lib.mat(yTilde)
    m, n = yT.shape
    gp, Gv, Y = _lib.vec(gPrime), _lib.vec(G), _lib.vec(YTilde)
    gradient = np.empty(n, dtype=np.float64)
    t0 = time.time()
    _L()._grad_bioen_log_posterior_logw(_lib.ptr(gp), _lib.ptr(Gv), _lib.ptr(yT), _lib.ptr(Y), None,
                                        _lib.ptr(gradient), float(theta), 1 if caching else 0, None, None, None,
                                        m, n, 0.0)
    _lib.check_pending("grad_bioen_log_posterior_logw")
    if print_timing:
        print("_grad_bioen_log_posterior_logw: {}".format(time.time() - t0))
    return gradient


def _params(m, n, yT, Y, theta, caching, result, g=None, G=None, forces=None, w0=None):
    p = _lib.params_t()
    p.forces = _lib.ptr(forces) if forces is not None else None
    p.w0 = _lib.ptr(w0) if w0 is not None else None
    p.g = _lib.ptr(g) if g is not None else None
    p.G = _lib.ptr(G) if G is not None else None
    p.yTilde = _lib.ptr(yT)
    p.YTilde = _lib.ptr(Y)
    p.w = None
    p.result = _lib.ptr(result)
    p.theta = float(theta)
    p.yTildeT = None
    p.caching = 1 if caching else 0
    p.tmp_n = None
    p.tmp_m = None
    p.m = m
    p.n = n
    return p


def _gsl_cfg(params):
    c = _lib.gsl_config_params()
    c.algorithm = get_gsl_method(params["algorithm"])
    c.tol = params["params"]["tol"]
    c.step_size = params["params"]["step_size"]
    c.max_iterations = params["params"]["max_iterations"]
    return c


def _lbfgs_cfg(params):
    c = _lib.lbfgs_config_params()
    for k in ("linesearch", "max_iterations", "delta", "epsilon", "ftol", "gtol", "wolfe", "past",
              "max_linesearch"):
        setattr(c, k, params["params"][k])
    return c


def _visual(params):
    return _lib.visual_params(int(bool(params["debug"])), int(bool(params["verbose"])))


def _finish_gsl(name, errno, result, fmin):
    _lib.check_pending(name)
    if errno in gsl_success:
        if errno == gsl_continue:
            print(gsl_continue_msg)
        return result, fmin
    raise RuntimeError("{}, GSL return code: {}:{}".format(name, errno, _L().bioen_gsl_error(errno).decode()))


def _finish_lbfgs(name, errno, result, fmin):
    _lib.check_pending(name)
    if errno in lbfgs_success:
        return result, fmin
    raise RuntimeError("{}, liblbfgs return code: {}:{}".format(name, errno, _L().lbfgs_strerror(errno).decode()))


def bioen_opt_bfgs_logw(g, G, yTilde, YTilde, theta, params):
    """c_bioen.pyx:362-438 -> (xfinal[n], fmin)"""
    _lib.clear_pending()
    yT = _lib.mat(yTilde)
    m, n = yT.shape
    gv, Gv, Y = _lib.vec(g), _lib.vec(G), _lib.vec(YTilde)
    result = np.empty(n, dtype=np.float64)
    cfg = _gsl_cfg(params)
    errno = C.c_int(0)
    fmin = _L()._opt_bfgs_logw(_params(m, n, yT, Y, theta, params["cache_ytilde_transposed"], result, g=gv, G=Gv),
                               cfg, _visual(params), C.byref(errno))
    return _finish_gsl("bioen_opt_bfgs_logw", errno.value, result, fmin)


def bioen_opt_lbfgs_logw(g, G, yTilde, YTilde, theta, params):
    """c_bioen.pyx:441-520 -> (xfinal[n], fmin)"""
    _lib.clear_pending()
    yT = _lib.mat(yTilde)
    m, n = yT.shape
    gv, Gv, Y = _lib.vec(g), _lib.vec(G), _lib.vec(YTilde)
    result = np.empty(n, dtype=np.float64)
    cfg = _lbfgs_cfg(params)
    errno = C.c_int(0)
    fmin = _L()._opt_lbfgs_logw(_params(m, n, yT, Y, theta, params["cache_ytilde_transposed"], result, g=gv, G=Gv),
                                cfg, _visual(params), C.byref(errno))
    return _finish_lbfgs("bioen_opt_lbfgs_logw", errno.value, result, fmin)


def _forces_given_weights(forces, w0, yTilde, YTilde, theta, want_gradient):
    """The pyx computes the weights from the forces and then calls the given-weights C entry point with them
    (c_bioen.pyx:560-579, 620-641).  Here both steps run on ONE resident copy of yTilde (one upload, tile kernels
    only: no structure-major copy is made for a single evaluation)."""
    from ...problem import FORCES, Problem
    with Problem(yTilde) as p:
        p.set_option(1, 0)
        p.set_forces(w0, YTilde, theta)
        w, _ = p.weights(forces, FORCES)
        return p.forces_from_weights(w, gradient=want_gradient)


def bioen_log_posterior_forces(forces, w0, yTilde, YTilde, theta, caching=False):
    """c_bioen.pyx:523-581: weights from the forces, then the objective for those weights"""
    return _forces_given_weights(forces, w0, yTilde, YTilde, theta, False)


def grad_bioen_log_posterior_forces(forces, w0, yTilde, YTilde, theta, caching=False):
    """c_bioen.pyx:584-643"""
    return _forces_given_weights(forces, w0, yTilde, YTilde, theta, True)[1]


def bioen_opt_bfgs_forces(forces, w0, yTilde, YTilde, theta, params):
    """c_bioen.pyx:646-716 -> (xfinal[m], fmin)"""
    _lib.clear_pending()
    yT = _lib.mat(yTilde)
    m, n = yT.shape
    f, w0v, Y = _lib.vec(forces), _lib.vec(w0), _lib.vec(YTilde)
    result = np.empty(m, dtype=np.float64)
    cfg = _gsl_cfg(params)
    errno = C.c_int(0)
    fmin = _L()._opt_bfgs_forces(_params(m, n, yT, Y, theta, params["cache_ytilde_transposed"], result, forces=f,
                                         w0=w0v), cfg, _visual(params), C.byref(errno))
    return _finish_gsl("bioen_opt_bfgs_forces", errno.value, result, fmin)


def bioen_opt_lbfgs_forces(forces, w0, yTilde, YTilde, theta, params):
    """c_bioen.pyx:719-792 -> (xfinal[m], fmin)"""
    _lib.clear_pending()
    yT = _lib.mat(yTilde)
    m, n = yT.shape
    f, w0v, Y = _lib.vec(forces), _lib.vec(w0), _lib.vec(YTilde)
    result = np.empty(m, dtype=np.float64)
    cfg = _lbfgs_cfg(params)
    errno = C.c_int(0)
    fmin = _L()._opt_lbfgs_forces(_params(m, n, yT, Y, theta, params["cache_ytilde_transposed"], result, forces=f,
                                          w0=w0v), cfg, _visual(params), C.byref(errno))
    return _finish_lbfgs("bioen_opt_lbfgs_forces", errno.value, result, fmin)
