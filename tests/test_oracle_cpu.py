"""CPU: the oracle (oracle/oracle.py + oracle/bioen_oracle.c) is pinned against the golden vectors produced
by the UNMODIFIED reference (tests/golden/make_golden.py) and, where oracle/_ref is built, against the
reference library itself."""
import numpy as np
import pytest

from conftest import FORCES_FIXTURES, LOGW_FIXTURES, grad_err, load_golden, rel

F_TOL = 5e-14      # the reference's own C-vs-NumPy tolerance (test_func_gradient_logw.py:35-57)
G_TOL = 1e-11     # north_star per-evaluation tolerance, norm-relative (forces gradients cancel: ~2e-12 seen)


@pytest.mark.parametrize("name", LOGW_FIXTURES)
def test_logw_eval_golden(oracle, name):
    d = load_golden(name)
    for x, fk, gk in ((d["GInit"], "f_init", "grad_init"), (d["probe"], "f_probe", "grad_probe")):
        f, g = oracle.logw_fg(x, d["G"], d["yTilde"], d["YTilde"], d["theta"])
        f2, g2 = oracle.logw_fg_np(x, d["G"], d["yTilde"], d["YTilde"], d["theta"])
        assert rel(f, d[fk]) < F_TOL and rel(f2, d[fk]) < F_TOL
        assert grad_err(g, d[gk]) < G_TOL and grad_err(g2, d[gk]) < G_TOL
    w, _ = oracle.logw_weights(d["probe"])
    assert np.max(np.abs(w - d["w_probe"])) < 1e-15


@pytest.mark.parametrize("name", FORCES_FIXTURES)
def test_forces_eval_golden(oracle, name):
    d = load_golden(name)
    for x, fk, gk in ((d["forces_init"], "f_init", "grad_init"), (d["probe"], "f_probe", "grad_probe")):
        f, g = oracle.forces_fg(x, d["w0"], d["yTilde"], d["YTilde"], d["theta"])
        f2, g2 = oracle.forces_fg_np(x, d["w0"], d["yTilde"], d["YTilde"], d["theta"])
        assert rel(f, d[fk]) < F_TOL and rel(f2, d[fk]) < F_TOL
        assert grad_err(g, d[gk]) < G_TOL and grad_err(g2, d[gk]) < G_TOL
    w = oracle.forces_weights(d["probe"], d["w0"], d["yTilde"])
    assert np.max(np.abs(w - d["w_probe"])) < 1e-15


@pytest.mark.parametrize("name", ["data_16x15", "data_potra_part_2_logw_M205xN10"])
@pytest.mark.parametrize("ls", [0, 1, 2, 3])
def test_lbfgs_restatement_logw(oracle, name, ls):
    d = load_golden(name)
    fg = lambda x: oracle.logw_fg(x, d["G"], d["yTilde"], d["YTilde"], d["theta"])
    r = oracle.lbfgs(fg, d["GInit"], linesearch=ls)
    assert r["code"] == d["lbfgs%d_code" % ls]
    assert rel(r["fx"], d["lbfgs%d_fmin" % ls]) < 1e-9
    if r["code"] in (0, 1, 2):
        assert np.max(np.abs(r["x"] - d["lbfgs%d_x" % ls])) < 1e-5


@pytest.mark.parametrize("name", FORCES_FIXTURES)
@pytest.mark.parametrize("ls", [0, 2])
def test_lbfgs_restatement_forces(oracle, name, ls):
    d = load_golden(name)
    fg = lambda x: oracle.forces_fg(x, d["w0"], d["yTilde"], d["YTilde"], d["theta"])
    r = oracle.lbfgs(fg, d["forces_init"], linesearch=ls)
    assert r["code"] == d["lbfgs%d_code" % ls]
    assert rel(r["fx"], d["lbfgs%d_fmin" % ls]) < 1e-8


@pytest.mark.parametrize("alg", ["conjugate_fr", "conjugate_pr", "bfgs2", "bfgs", "steepest_descent"])
def test_gsl_restatement_logw(oracle, alg):
    d = load_golden("data_16x15")
    fg = lambda x: oracle.logw_fg(x, d["G"], d["yTilde"], d["YTilde"], d["theta"])
    r = oracle.gsl_minimize(fg, d["GInit"], algorithm=alg)
    assert r["code"] == d["gsl_%s_code" % alg]
    assert rel(r["fx"], d["gsl_%s_fmin" % alg]) < 1e-9


@pytest.mark.parametrize("alg", ["bfgs2", "conjugate_fr"])
def test_gsl_restatement_forces(oracle, alg):
    d = load_golden("data_forces_M64xN64")
    fg = lambda x: oracle.forces_fg(x, d["w0"], d["yTilde"], d["YTilde"], d["theta"])
    r = oracle.gsl_minimize(fg, d["forces_init"], algorithm=alg)
    assert r["code"] == d["gsl_%s_code" % alg]
    assert rel(r["fx"], d["gsl_%s_fmin" % alg]) < 1e-8


def test_synthetic_golden(oracle):
    d = load_golden("synthetic_M37xN5001")
    P = oracle.synthetic_problem(int(d["M"]), int(d["N"]), seed=int(d["seed"]))
    g1 = 0.1 * np.random.default_rng(1).standard_normal(P["N"])
    f1 = 1e-3 * np.random.default_rng(2).standard_normal(P["M"])
    f, g = oracle.logw_fg(g1, P["G"], P["yTilde"], P["YTilde"], d["theta"])
    assert rel(f, d["logw_f"]) < F_TOL and grad_err(g, d["logw_grad"]) < G_TOL
    f, g = oracle.forces_fg(f1, P["w0"], P["yTilde"], P["YTilde"], d["theta"])
    assert rel(f, d["forces_f"]) < F_TOL and grad_err(g, d["forces_grad"]) < G_TOL


def test_against_reference_library_if_built(oracle):
    """Only where oracle/_ref/libbioen_ref.so exists (this container; it also travels to the GPU box)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    ref.set_fast_openmp_flag(0)
    P = oracle.synthetic_problem(23, 777, seed=7)
    g1 = 0.3 * np.random.default_rng(3).standard_normal(777)
    f = oracle.logw_fg(g1, P["G"], P["yTilde"], P["YTilde"], 3.0)
    assert rel(f[0], ref.logw_objective(g1, P["G"], P["yTilde"], P["YTilde"], 3.0)) < F_TOL
    assert grad_err(f[1], ref.logw_gradient(g1, P["G"], P["yTilde"], P["YTilde"], 3.0)) < G_TOL
    f1 = 1e-2 * np.random.default_rng(4).standard_normal(23)
    f = oracle.forces_fg(f1, P["w0"], P["yTilde"], P["YTilde"], 3.0)
    assert rel(f[0], ref.forces_objective(f1, P["w0"], P["yTilde"], P["YTilde"], 3.0)) < F_TOL
    assert grad_err(f[1], ref.forces_gradient(f1, P["w0"], P["yTilde"], P["YTilde"], 3.0)) < G_TOL


def test_host_generator_matches_the_numpy_restatement_of_the_device_generator():
    """oracle/hostgen.c (the matrix of bench.py's reference arm) against tests/util_rng.py (the NumPy restatement of
    the device's k_generate, which tests/test_gpu_fullsize.py checks against the device itself): same counter-based
    law, entries equal to a few ulp (libm cos/log vs NumPy's)."""
    import ctypes as C
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import oracle as O
    from util_rng import generic_ytilde_block
    lib = O.hostgen()
    dp = C.POINTER(C.c_double)
    m, n, col0 = 37, 5001, 123_456_789
    a = np.random.default_rng(0).standard_normal(m)
    Y = np.empty((m, n))
    lib.hostgen_generic_ytilde(Y.ctypes.data_as(dp), n, m, n, 12345, col0, a.ctypes.data_as(dp), 2.0)
    R = generic_ytilde_block(12345, a, 2.0, 0, m, col0, n)
    assert np.max(np.abs(Y - R)) < 1e-13
    assert abs(Y.std() - np.sqrt(4.0 + a.var())) < 0.05          # N(a_i, 2^2) entries
    # a sub-block generated with an offset is the same block of the big matrix
    Y2 = np.empty((m, 100))
    lib.hostgen_generic_ytilde(Y2.ctypes.data_as(dp), 100, m, 100, 12345, col0 + 700, a.ctypes.data_as(dp), 2.0)
    assert np.array_equal(Y2, Y[:, 700:800])
