"""CPU: the product's liblbfgs line-search state machine (csrc/lbfgs.cuh LineSearchState -- the code the device
L-BFGS and the theta scan run on the host side) against the oracle's restatement of lbfgs.c:645-1001, on 1-D
functions, through the host-only C hook bioen_b200_selftest_linesearch.  No GPU involved."""
import ctypes as C
import math

import numpy as np
import pytest

from bioen_b200 import _lib

PHI = {
    "quadratic": (lambda t: (t - 1.3) ** 2 + 0.5, lambda t: 2 * (t - 1.3)),
    "quartic": (lambda t: (t - 2.0) ** 4 - 3 * t + 7.0, lambda t: 4 * (t - 2.0) ** 3 - 3),
    "exp": (lambda t: math.exp(0.7 * t) - 2.5 * t, lambda t: 0.7 * math.exp(0.7 * t) - 2.5),
    "more_thuente_1": (lambda t: -t / (t * t + 2.0), lambda t: (t * t - 2.0) / (t * t + 2.0) ** 2),
    "wiggly": (lambda t: 0.1 * math.sin(9 * t) + (t - 0.8) ** 2, lambda t: 0.9 * math.cos(9 * t) + 2 * (t - 0.8)),
}


def _oracle_search(oracle, name, ls, stp0, over):
    phi, dphi = PHI[name]
    p = dict(oracle.LBFGS_BIOEN_DEFAULTS)
    p.update(over)
    p["linesearch"] = ls

    def evaluate(x, g):
        g[0] = dphi(float(x[0]))
        return phi(float(x[0]))

    x, g, xp, s = np.zeros(1), np.array([dphi(0.0)]), np.zeros(1), np.ones(1)
    fn = oracle._ls_morethuente if ls == 0 else oracle._ls_backtracking
    code, f, stp = fn(evaluate, x, phi(0.0), g, s, stp0, xp, p)
    return code, f, stp


def _product_search(name, ls, stp0, over):
    phi, dphi = PHI[name]
    lib = _lib.load()
    cfg = _lib.lbfgs_config_params(linesearch=ls, max_iterations=0, delta=1e-6, epsilon=1e-6,
                                   ftol=over.get("ftol", 1e-5), gtol=over.get("gtol", 0.9),
                                   wolfe=over.get("wolfe", 0.9), past=10,
                                   max_linesearch=over.get("max_linesearch", 100))
    CB = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double))

    def cb(t, f, dg):
        f[0] = phi(t)
        dg[0] = dphi(t)

    stp, f, n = C.c_double(), C.c_double(), C.c_int()
    code = lib.bioen_b200_selftest_linesearch(cfg, phi(0.0), dphi(0.0), stp0, CB(cb), C.byref(stp), C.byref(f),
                                              C.byref(n))
    return code, f.value, stp.value, n.value


@pytest.mark.parametrize("name", sorted(PHI))
@pytest.mark.parametrize("ls", [0, 1, 2, 3])
@pytest.mark.parametrize("stp0", [1e-3, 1.0, 25.0])
def test_linesearch_matches_liblbfgs_restatement(oracle, name, ls, stp0):
    for over in ({}, {"ftol": 1e-4, "gtol": 0.1, "wolfe": 0.5}, {"max_linesearch": 3}):
        try:
            co, fo, so = _oracle_search(oracle, name, ls, stp0, over)
        except ValueError:
            continue      # the Python restatement's math.sqrt raises where C yields NaN: nothing to compare against
        cp, fp, sp, n = _product_search(name, ls, stp0, over)
        assert cp == co, (name, ls, stp0, over, cp, co)
        assert fp == fo and sp == so, (name, ls, stp0, over, (fp, sp), (fo, so))
        if cp > 0:
            assert n == cp


def test_linesearch_rejects_ascent_direction_and_bad_step():
    lib = _lib.load()
    cfg = _lib.lbfgs_config_params(linesearch=2, max_iterations=0, delta=0, epsilon=0, ftol=1e-5, gtol=0.9,
                                   wolfe=0.9, past=0, max_linesearch=10)
    CB = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double))
    cb = CB(lambda t, f, dg: None)
    assert lib.bioen_b200_selftest_linesearch(cfg, 1.0, +0.5, 1.0, cb, None, None, None) == -994   # INCREASEGRADIENT
    assert lib.bioen_b200_selftest_linesearch(cfg, 1.0, -0.5, 0.0, cb, None, None, None) == -995   # INVALIDPARAMETERS


def test_fletcher_interpolation_matches_gsl_restatement(oracle):
    """gsl_min.cuh fletcher::interpolate (the bracketing / sectioning step of vector_bfgs2's line search) against
    the oracle's restatement of linear_minimize.c, including the NaN-derivative (quadratic) branch."""
    lib = _lib.load()
    rng = np.random.default_rng(11)
    for _ in range(400):
        a, b = sorted(rng.uniform(-2, 5, 2))
        fa, fb = rng.uniform(-3, 3, 2)
        fpa = rng.uniform(-4, 0.5)
        fpb = rng.uniform(-4, 4) if rng.random() < 0.7 else float("nan")
        lo, hi = sorted(rng.uniform(a - 1, b + 2, 2))
        for order in (2, 3):
            want = oracle._interpolate(a, fa, fpa, b, fb, fpb, lo, hi, order)
            got = lib.bioen_b200_selftest_interpolate(a, fa, fpa, b, fb, fpb, lo, hi, order)
            assert got == want or (math.isnan(got) and math.isnan(want)), (a, fa, fpa, b, fb, fpb, lo, hi, order)
