"""GPU: the reference's UNMODIFIED Cython boundary (bioen/optimize/ext/c_bioen.pyx, cythonized by `make -C oracle
dropin` where /root/reference is mounted) linked against libbioen_b200.so instead of the reference's OpenMP C code
-- the drop-in of INTEGRATION.md section A.  Every Python-callable function of that module is driven here and
checked against the reference's stored outputs (tests/golden)."""
import glob
import os
import sys

import numpy as np
import pytest

from conftest import FORCES_FIXTURES, LOGW_FIXTURES, ROOT, grad_err, load_golden, rel

DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
HAVE = bool(glob.glob(os.path.join(DROPIN, "c_bioen*.so")))
needs_dropin = pytest.mark.skipif(not HAVE, reason="oracle/_ref/dropin not built (make -C oracle dropin)")


def _module():
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import c_bioen          # the reference's extension module name
    return c_bioen


@needs_dropin
def test_reference_cython_module_imports_and_reports_libraries():
    m = _module()
    for name in ("set_fast_openmp_flag", "get_fast_openmp_flag", "omp_set_num_threads", "get_gsl_method", "library_gsl",
                 "library_lbfgs", "bioen_log_posterior_logw", "grad_bioen_log_posterior_logw", "bioen_opt_bfgs_logw",
                 "bioen_opt_lbfgs_logw", "bioen_log_posterior_forces", "grad_bioen_log_posterior_forces",
                 "bioen_opt_bfgs_forces", "bioen_opt_lbfgs_forces"):
        assert callable(getattr(m, name)), name
    assert m.library_gsl() is True and m.library_lbfgs() is True
    m.set_fast_openmp_flag(1)
    assert m.get_fast_openmp_flag() == 1
    m.set_fast_openmp_flag(0)
    assert m.get_gsl_method("gsl_multimin_fdfminimizer_vector_bfgs2") == 2
    with pytest.raises(RuntimeError, match="return code"):
        m.get_gsl_method("TEST_INVALID")


def _cfg(minimizer):
    from bioen_b200 import optimize
    cfg = optimize.minimize.Parameters(minimizer)       # same dict layout as the reference's Parameters()
    cfg["verbose"] = False
    cfg["cache_ytilde_transposed"] = True                # exercise the yTildeT argument path of the pyx
    return cfg


@pytest.mark.gpu
@needs_dropin
@pytest.mark.parametrize("name", LOGW_FIXTURES)
def test_dropin_logw(name):
    m = _module()
    d = load_golden(name)
    G, yT, YT = np.ascontiguousarray(d["G"].ravel()), np.ascontiguousarray(d["yTilde"]), np.ascontiguousarray(d["YTilde"].ravel())
    for x, fk, gk in ((d["GInit"], "f_init", "grad_init"), (d["probe"], "f_probe", "grad_probe")):
        x = np.ascontiguousarray(np.asarray(x).ravel())
        f = m.bioen_log_posterior_logw(x, G, G, yT, YT, d["theta"])
        g = m.grad_bioen_log_posterior_logw(x, G, G, yT, YT, d["theta"], True)
        assert rel(f, d[fk]) < 1e-11 and grad_err(g, d[gk]) < 1e-11
    x0 = np.ascontiguousarray(d["GInit"].ravel())
    xo, fmin = m.bioen_opt_lbfgs_logw(x0, G, yT, YT, d["theta"], _cfg("lbfgs"))
    assert rel(fmin, d["lbfgs2_fmin"]) < 1e-8 and xo.shape == x0.shape
    xo, fmin = m.bioen_opt_bfgs_logw(x0, G, yT, YT, d["theta"], _cfg("gsl"))
    assert rel(fmin, d["gsl_bfgs2_fmin"]) < 1e-6


@pytest.mark.gpu
@needs_dropin
@pytest.mark.parametrize("name", FORCES_FIXTURES)
def test_dropin_forces(name):
    m = _module()
    d = load_golden(name)
    w0, yT, YT = np.ascontiguousarray(d["w0"].ravel()), np.ascontiguousarray(d["yTilde"]), np.ascontiguousarray(d["YTilde"].ravel())
    for x, fk, gk in ((d["forces_init"], "f_init", "grad_init"), (d["probe"], "f_probe", "grad_probe")):
        x = np.ascontiguousarray(np.asarray(x).ravel())
        f = m.bioen_log_posterior_forces(x, w0, yT, YT, d["theta"])
        g = m.grad_bioen_log_posterior_forces(x, w0, yT, YT, d["theta"], True)
        assert rel(f, d[fk]) < 1e-11 and grad_err(g, d[gk]) < 1e-11
    x0 = np.ascontiguousarray(d["forces_init"].ravel())
    xo, fmin = m.bioen_opt_lbfgs_forces(x0, w0, yT, YT, d["theta"], _cfg("lbfgs"))
    assert rel(fmin, d["lbfgs2_fmin"]) < 1e-8 and xo.shape == x0.shape
    xo, fmin = m.bioen_opt_bfgs_forces(x0, w0, yT, YT, d["theta"], _cfg("gsl"))
    assert rel(fmin, d["gsl_bfgs2_fmin"]) < 1e-6


@pytest.mark.gpu
@needs_dropin
def test_dropin_error_convention():
    """test_error_opt_logw.py / test_error_opt_forces.py of the reference, through its own Cython code."""
    m = _module()
    d = load_golden("data_16x15")
    G, yT, YT = np.ascontiguousarray(d["G"].ravel()), np.ascontiguousarray(d["yTilde"]), np.ascontiguousarray(d["YTilde"].ravel())
    x0 = np.ascontiguousarray(d["GInit"].ravel())
    cfg = _cfg("lbfgs")
    cfg["params"]["delta"] = -1
    with pytest.raises(RuntimeError, match="liblbfgs return code: -1015"):
        m.bioen_opt_lbfgs_logw(x0, G, yT, YT, d["theta"], cfg)
    cfg = _cfg("gsl")
    cfg["algorithm"] = "TEST_INVALID"
    with pytest.raises(RuntimeError, match="GSL return code"):
        m.bioen_opt_bfgs_logw(x0, G, yT, YT, d["theta"], cfg)
