"""ctypes binding to oracle/_ref/libbioen_ref.so -- TEST INFRASTRUCTURE ONLY.

`libbioen_ref.so` is the UNMODIFIED reference implementation of the hot path (BioEn's
bioen/optimize/ext/c_bioen_kernels_{logw,forces}.c + c_bioen_common.c, linked with the vendored
liblbfgs 1.10 and the multimin subset of the vendored GSL 2.5), compiled in place from /root/reference by
oracle/Makefile.  This module plays the role of the reference's Cython layer
(bioen/optimize/ext/c_bioen.pyx) for that library so tests and the CPU-baseline leg of bench.py can call
the reference kernels and minimiser drivers directly.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product (bioen_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbioen_ref.so")

_dp = C.POINTER(C.c_double)


class params_t(C.Structure):  # c_bioen_common.h:44-60
    _fields_ = [("forces", _dp), ("w0", _dp), ("g", _dp), ("G", _dp), ("yTilde", _dp), ("YTilde", _dp),
                ("w", _dp), ("result", _dp), ("theta", C.c_double), ("yTildeT", _dp), ("caching", C.c_int),
                ("tmp_n", _dp), ("tmp_m", _dp), ("m", C.c_int), ("n", C.c_int)]


class gsl_config_params(C.Structure):  # c_bioen_common.h:62-67
    _fields_ = [("step_size", C.c_double), ("tol", C.c_double), ("max_iterations", C.c_int),
                ("algorithm", C.c_int)]


class lbfgs_config_params(C.Structure):  # c_bioen_common.h:69-79
    _fields_ = [("linesearch", C.c_int), ("max_iterations", C.c_int), ("delta", C.c_double),
                ("epsilon", C.c_double), ("ftol", C.c_double), ("gtol", C.c_double), ("wolfe", C.c_double),
                ("past", C.c_int), ("max_linesearch", C.c_int)]


class visual_params(C.Structure):  # c_bioen_common.h:89-92
    _fields_ = [("debug", C.c_size_t), ("verbose", C.c_size_t)]


GSL_ALGORITHMS = {"conjugate_fr": 0, "conjugate_pr": 1, "bfgs2": 2, "bfgs": 3, "steepest_descent": 4}
LBFGS_DEFAULTS = dict(linesearch=2, max_iterations=5000, delta=1e-6, epsilon=1e-6, ftol=1e-5, gtol=0.9,
                      wolfe=0.9, past=10, max_linesearch=100)  # config/bioen_optimize.yaml:33-46
GSL_DEFAULTS = dict(step_size=0.01, tol=0.001, max_iterations=5000)  # config/bioen_optimize.yaml:20-31


def available():
    return os.path.isfile(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libbioen_ref.so missing: run `make -C oracle ref` where "
                               "/root/reference exists")
        L = C.CDLL(LIB_PATH)
        L._get_weights.restype = C.c_double
        L._get_weights.argtypes = [_dp, _dp, C.c_size_t]
        L._bioen_log_posterior_logw.restype = C.c_double
        L._bioen_log_posterior_logw.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp,
                                                _dp, C.c_int, C.c_int, C.c_double]
        L._grad_bioen_log_posterior_logw.restype = None
        L._grad_bioen_log_posterior_logw.argtypes = L._bioen_log_posterior_logw.argtypes
        L._get_weights_from_forces.restype = None
        L._get_weights_from_forces.argtypes = [_dp, _dp, _dp, _dp, C.c_int, _dp, _dp, C.c_size_t, C.c_size_t]
        L._bioen_log_posterior_forces.restype = C.c_double
        L._bioen_log_posterior_forces.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp, _dp,
                                                  C.c_int, C.c_int]
        L._grad_bioen_log_posterior_forces.restype = None
        L._grad_bioen_log_posterior_forces.argtypes = L._bioen_log_posterior_forces.argtypes
        for nm in ("_opt_bfgs_logw", "_opt_bfgs_forces"):
            f = getattr(L, nm)
            f.restype = C.c_double
            f.argtypes = [params_t, gsl_config_params, visual_params, C.POINTER(C.c_int)]
        for nm in ("_opt_lbfgs_logw", "_opt_lbfgs_forces"):
            f = getattr(L, nm)
            f.restype = C.c_double
            f.argtypes = [params_t, lbfgs_config_params, visual_params, C.POINTER(C.c_int)]
        L.lbfgs_strerror.restype = C.c_char_p
        L.lbfgs_strerror.argtypes = [C.c_int]
        L.bioen_gsl_error.restype = C.c_char_p
        L.bioen_gsl_error.argtypes = [C.c_int]
        L._set_fast_openmp_flag.argtypes = [C.c_int]
        L._omp_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _vec(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _mat(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def set_fast_openmp_flag(flag):
    lib()._set_fast_openmp_flag(int(flag))


def set_num_threads(n):
    lib()._omp_set_num_threads(int(n))


# ---------------------------------------------------------------- log-weights: evaluation
def logw_weights(g):
    g = _vec(g)
    w = np.empty_like(g)
    s = lib()._get_weights(_p(g), _p(w), g.size)
    return w, s


def logw_objective(g, G, yTilde, YTilde, theta):
    """c_bioen.pyx:246-292 (with the real G passed where the pyx passes its 2nd positional argument)."""
    g, G, Y = _vec(g), _vec(G), _vec(YTilde)
    yT = _mat(yTilde)
    m, n = yT.shape
    w = np.empty(n)
    tn, tm = np.empty(n), np.empty(m)
    s = lib()._get_weights(_p(g), _p(w), n)
    return lib()._bioen_log_posterior_logw(_p(g), _p(G), _p(yT), _p(Y), _p(w), None, float(theta), 0, None,
                                           _p(tn), _p(tm), m, n, s)


def logw_gradient(g, G, yTilde, YTilde, theta, caching=False):
    """c_bioen.pyx:295-359."""
    g, G, Y = _vec(g), _vec(G), _vec(YTilde)
    yT = _mat(yTilde)
    m, n = yT.shape
    w, grad = np.empty(n), np.empty(n)
    tn, tm = np.empty(n), np.empty(m)
    yTT = np.ascontiguousarray(yT.T) if caching else np.empty(1)
    s = lib()._get_weights(_p(g), _p(w), n)
    lib()._grad_bioen_log_posterior_logw(_p(g), _p(G), _p(yT), _p(Y), _p(w), _p(grad), float(theta),
                                         1 if caching else 0, _p(yTT), _p(tn), _p(tm), m, n, s)
    return grad


class LogwEvaluator:
    """Pre-allocated f+g evaluation = interface_lbfgs_logw (c_bioen_kernels_logw.c:525-561); for timing."""

    def __init__(self, G, yTilde, YTilde, theta, caching=False):
        self.G, self.Y, self.yT = _vec(G), _vec(YTilde), _mat(yTilde)
        self.m, self.n = self.yT.shape
        self.theta = float(theta)
        self.caching = 1 if caching else 0
        self.yTT = np.ascontiguousarray(self.yT.T) if caching else np.empty(1)
        self.w, self.grad = np.empty(self.n), np.empty(self.n)
        self.tn, self.tm = np.empty(self.n), np.empty(self.m)

    def __call__(self, g):
        L = lib()
        s = L._get_weights(_p(g), _p(self.w), self.n)
        f = L._bioen_log_posterior_logw(_p(g), _p(self.G), _p(self.yT), _p(self.Y), _p(self.w), None,
                                        self.theta, self.caching, _p(self.yTT), _p(self.tn), _p(self.tm),
                                        self.m, self.n, s)
        L._grad_bioen_log_posterior_logw(_p(g), _p(self.G), _p(self.yT), _p(self.Y), _p(self.w),
                                         _p(self.grad), self.theta, self.caching, _p(self.yTT), _p(self.tn),
                                         _p(self.tm), self.m, self.n, -1.0)
        return f, self.grad


# ---------------------------------------------------------------- forces: evaluation
def forces_weights(forces, w0, yTilde, caching=False):
    f, w0 = _vec(forces), _vec(w0)
    yT = _mat(yTilde)
    m, n = yT.shape
    w, tn = np.empty(n), np.empty(n)
    yTT = np.ascontiguousarray(yT.T) if caching else np.empty(1)
    lib()._get_weights_from_forces(_p(w0), _p(yT), _p(f), _p(w), 1 if caching else 0, _p(yTT), _p(tn), m, n)
    return w


def forces_objective(forces, w0, yTilde, YTilde, theta):
    """c_bioen.pyx:523-581."""
    f, w0, Y = _vec(forces), _vec(w0), _vec(YTilde)
    yT = _mat(yTilde)
    m, n = yT.shape
    w, tn, tm = np.empty(n), np.empty(n), np.empty(m)
    L = lib()
    L._get_weights_from_forces(_p(w0), _p(yT), _p(f), _p(w), 0, None, _p(tn), m, n)
    return L._bioen_log_posterior_forces(_p(w0), _p(yT), _p(Y), _p(w), None, float(theta), 0, None, _p(tn),
                                         _p(tm), m, n)


def forces_gradient(forces, w0, yTilde, YTilde, theta):
    """c_bioen.pyx:584-643."""
    f, w0, Y = _vec(forces), _vec(w0), _vec(YTilde)
    yT = _mat(yTilde)
    m, n = yT.shape
    w, tn, tm, grad = np.empty(n), np.empty(n), np.empty(m), np.empty(m)
    L = lib()
    L._get_weights_from_forces(_p(w0), _p(yT), _p(f), _p(w), 0, None, _p(tn), m, n)
    L._grad_bioen_log_posterior_forces(_p(w0), _p(yT), _p(Y), _p(w), _p(grad), float(theta), 0, None, _p(tn),
                                       _p(tm), m, n)
    return grad


class ForcesEvaluator:
    """Pre-allocated f+g evaluation = interface_lbfgs_forces (c_bioen_kernels_forces.c:43-76)."""

    def __init__(self, w0, yTilde, YTilde, theta, caching=False):
        self.w0, self.Y, self.yT = _vec(w0), _vec(YTilde), _mat(yTilde)
        self.m, self.n = self.yT.shape
        self.theta = float(theta)
        self.caching = 1 if caching else 0
        self.yTT = np.ascontiguousarray(self.yT.T) if caching else np.empty(1)
        self.w, self.tn = np.empty(self.n), np.empty(self.n)
        self.tm, self.grad = np.empty(self.m), np.empty(self.m)

    def __call__(self, f):
        L = lib()
        a = (_p(self.w0), _p(self.yT))
        L._get_weights_from_forces(*a, _p(f), _p(self.w), self.caching, _p(self.yTT), _p(self.tn), self.m,
                                   self.n)
        val = L._bioen_log_posterior_forces(*a, _p(self.Y), _p(self.w), None, self.theta, self.caching,
                                            _p(self.yTT), _p(self.tn), _p(self.tm), self.m, self.n)
        L._grad_bioen_log_posterior_forces(*a, _p(self.Y), _p(self.w), _p(self.grad), self.theta,
                                           self.caching, _p(self.yTT), _p(self.tn), _p(self.tm), self.m,
                                           self.n)
        return val, self.grad


# ---------------------------------------------------------------- minimiser drivers
def _fill(cls, defaults, kw):
    d = dict(defaults)
    d.update(kw)
    return cls(**d)


def _params(m, n, yT, Y, theta, caching, *, g=None, G=None, forces=None, w0=None, result=None):
    keep = dict(w=np.empty(n), tn=np.empty(n), tm=np.empty(m),
                yTT=np.ascontiguousarray(yT.T) if caching else np.empty(1))
    p = params_t()
    if forces is not None:
        p.forces, p.w0 = _p(forces), _p(w0)
    if g is not None:
        p.g, p.G = _p(g), _p(G)
    p.yTilde, p.YTilde, p.w, p.result = _p(yT), _p(Y), _p(keep["w"]), _p(result)
    p.theta, p.yTildeT, p.caching = float(theta), _p(keep["yTT"]), 1 if caching else 0
    p.tmp_n, p.tmp_m, p.m, p.n = _p(keep["tn"]), _p(keep["tm"]), m, n
    return p, keep


def opt_lbfgs_logw(g0, G, yTilde, YTilde, theta, caching=False, verbose=0, **cfg):
    """c_bioen.pyx:441-520 -> _opt_lbfgs_logw (c_bioen_kernels_logw.c:581-669). Returns (x, fmin, code)."""
    g0, G, Y, yT = _vec(g0), _vec(G), _vec(YTilde), _mat(yTilde)
    m, n = yT.shape
    res = np.empty(n)
    p, keep = _params(m, n, yT, Y, theta, caching, g=g0, G=G, result=res)
    err = C.c_int(0)
    fmin = lib()._opt_lbfgs_logw(p, _fill(lbfgs_config_params, LBFGS_DEFAULTS, cfg),
                                 visual_params(0, int(verbose)), C.byref(err))
    return res, fmin, err.value


def opt_lbfgs_forces(f0, w0, yTilde, YTilde, theta, caching=False, verbose=0, **cfg):
    """c_bioen.pyx:719-792 -> _opt_lbfgs_forces (c_bioen_kernels_forces.c:574-662)."""
    f0, w0, Y, yT = _vec(f0), _vec(w0), _vec(YTilde), _mat(yTilde)
    m, n = yT.shape
    res = np.empty(m)
    p, keep = _params(m, n, yT, Y, theta, caching, forces=f0, w0=w0, result=res)
    err = C.c_int(0)
    fmin = lib()._opt_lbfgs_forces(p, _fill(lbfgs_config_params, LBFGS_DEFAULTS, cfg),
                                   visual_params(0, int(verbose)), C.byref(err))
    return res, fmin, err.value


def opt_gsl_logw(g0, G, yTilde, YTilde, theta, algorithm="bfgs2", caching=False, verbose=0, **cfg):
    """c_bioen.pyx:362-438 -> _opt_bfgs_logw (c_bioen_kernels_logw.c:367-509)."""
    g0, G, Y, yT = _vec(g0), _vec(G), _vec(YTilde), _mat(yTilde)
    m, n = yT.shape
    res = np.empty(n)
    p, keep = _params(m, n, yT, Y, theta, caching, g=g0, G=G, result=res)
    c = _fill(gsl_config_params, dict(GSL_DEFAULTS, algorithm=GSL_ALGORITHMS[algorithm]), cfg)
    err = C.c_int(0)
    fmin = lib()._opt_bfgs_logw(p, c, visual_params(0, int(verbose)), C.byref(err))
    return res, fmin, err.value


def opt_gsl_forces(f0, w0, yTilde, YTilde, theta, algorithm="bfgs2", caching=False, verbose=0, **cfg):
    """c_bioen.pyx:646-716 -> _opt_bfgs_forces (c_bioen_kernels_forces.c:431-570)."""
    f0, w0, Y, yT = _vec(f0), _vec(w0), _vec(YTilde), _mat(yTilde)
    m, n = yT.shape
    res = np.empty(m)
    p, keep = _params(m, n, yT, Y, theta, caching, forces=f0, w0=w0, result=res)
    c = _fill(gsl_config_params, dict(GSL_DEFAULTS, algorithm=GSL_ALGORITHMS[algorithm]), cfg)
    err = C.c_int(0)
    fmin = lib()._opt_bfgs_forces(p, c, visual_params(0, int(verbose)), C.byref(err))
    return res, fmin, err.value


def lbfgs_strerror(code):
    return lib().lbfgs_strerror(int(code)).decode()


def gsl_strerror(code):
    return lib().bioen_gsl_error(int(code)).decode()
