"""CPU: host-side mirror of bioen.optimize -- cfg dicts, helpers, shape checks, legacy NumPy twins."""
import numpy as np
import pytest

from conftest import FORCES_FIXTURES, LOGW_FIXTURES, grad_err, load_golden, rel

from bioen_b200 import optimize


def test_parameters_defaults_match_reference_template():
    # values of bioen/optimize/config/bioen_optimize.yaml
    c = optimize.minimize.Parameters("lbfgs")
    assert c["minimizer"] == "lbfgs" and c["algorithm"] == "" and c["use_c_functions"] is True
    assert c["params"] == dict(linesearch=2, max_iterations=5000, delta=1e-6, epsilon=1e-6, ftol=1e-5, gtol=0.9,
                               wolfe=0.9, past=10, max_linesearch=100)
    assert c["cache_ytilde_transposed"] == "auto" and c["n_threads"] == -1
    assert c["debug"] is True and c["verbose"] is True
    c = optimize.minimize.Parameters("gsl")
    assert c["algorithm"] == "gsl_multimin_fdfminimizer_vector_bfgs2"
    assert c["params"] == dict(step_size=0.01, tol=0.001, max_iterations=5000)
    c = optimize.minimize.Parameters("scipy")
    assert c["algorithm"] == "fmin_bfgs" and c["use_c_functions"] is True
    assert c["params"] == dict(gtol=0.001, pgtol=0.001, epsilon=0.1, max_iterations=5000)


def test_parameter_overrides():
    c = optimize.minimize.Parameters("lbfgs", "lbfgs:epsilon=1e-5,lbfgs:linesearch=0,general:verbose=false")
    assert c["params"]["epsilon"] == 1e-5 and c["params"]["linesearch"] == 0 and c["verbose"] is False
    assert optimize.util.ntype("12") == 12 and optimize.util.ntype("1.5") == 1.5
    assert optimize.util.ntype("Yes") is True and optimize.util.ntype("abc") == "abc"


def test_util_relative_differences():
    u = optimize.util
    assert u.compute_relative_difference_for_values(1.1, 1.0) == pytest.approx(0.1)
    assert u.compute_relative_difference_for_values(0.3, 0.0) == 0.3
    d, idx = u.compute_relative_difference_for_arrays(np.array([1.0, 2.2, 0.5]), np.array([1.0, 2.0, 0.0]))
    assert d == pytest.approx(0.1) and idx == 1
    assert u.compute_relative_difference_for_arrays(np.ones(3), np.zeros(3)) == (0.0, 0)
    assert u.library_gsl() and u.library_lbfgs()


def test_gsl_method_ids_and_error():
    from bioen_b200.optimize.ext import c_bioen
    assert c_bioen.get_gsl_method("conjugate_fr") == 0
    assert c_bioen.get_gsl_method("gsl_multimin_fdfminimizer_conjugate_pr") == 1
    assert c_bioen.get_gsl_method("bfgs2") == 2
    assert c_bioen.get_gsl_method("gsl_multimin_fdfminimizer_vector_bfgs") == 3
    assert c_bioen.get_gsl_method("steepest_descent") == 4
    with pytest.raises(RuntimeError, match="return code"):
        c_bioen.get_gsl_method("TEST_INVALID")


def test_shape_checks():
    n, m = 7, 3
    ok = dict(GInit=np.zeros((n, 1)), G=np.zeros((n, 1)), y=np.zeros((m, n)), yTilde=np.zeros((m, n)),
              YTilde=np.zeros((1, m)))
    optimize.log_weights.check_params_logweights(**ok)
    for key, bad in (("GInit", (n,)), ("G", (1, n)), ("y", (n, m)), ("YTilde", (m, 1))):
        args = dict(ok)
        args[key] = np.zeros(bad)
        with pytest.raises(ValueError):
            optimize.log_weights.check_params_logweights(**args)
    okf = dict(forcesInit=np.zeros((m, 1)), w0=np.zeros((n, 1)), y=np.zeros((m, n)), yTilde=np.zeros((m, n)),
               YTilde=np.zeros((1, m)))
    optimize.forces.check_params_forces(**okf)
    for key, bad in (("forcesInit", (1, m)), ("w0", (n,)), ("YTilde", (m,))):
        args = dict(okf)
        args[key] = np.zeros(bad)
        with pytest.raises(ValueError):
            optimize.forces.check_params_forces(**args)
    cfg = optimize.minimize.Parameters("lbfgs")
    with pytest.raises(ValueError):
        optimize.log_weights.find_optimum(np.zeros(n), ok["G"], ok["y"], ok["yTilde"], ok["YTilde"], 1.0, cfg)


def test_caching_heuristics_and_helpers():
    assert optimize.common.set_caching_heuristics(1000, 10 ** 6) is True
    assert optimize.common.set_caching_heuristics(5000, 10 ** 7) is False
    w0 = np.full((5, 1), 0.2)
    g = optimize.log_weights.getGs(w0)
    assert g.shape == (5, 1) and np.all(g == 0)
    w, s = optimize.log_weights.getWeights(g)
    assert np.allclose(w, 0.2) and s == 5.0
    assert optimize.forces.init_forces(4).shape == (4, 1)
    assert np.allclose(optimize.forces.get_weights_from_forces(w0, np.ones((3, 5)), np.zeros((3, 1))), 0.2)


@pytest.mark.parametrize("name", LOGW_FIXTURES[:3])
def test_legacy_logw_twins_agree_with_reference_values(name):
    d = load_golden(name)
    g = d["GInit"].copy()
    f = optimize.log_weights.bioen_log_posterior_base(d["probe"], g, d["G"], d["yTilde"], d["YTilde"], d["theta"])
    gr = optimize.log_weights.grad_bioen_log_posterior_base(d["probe"], g, d["G"], d["yTilde"], d["YTilde"],
                                                            d["theta"])
    assert rel(f, d["f_probe"]) < 5e-14
    assert grad_err(gr, d["grad_probe"]) < 1e-11


@pytest.mark.parametrize("name", FORCES_FIXTURES)
def test_legacy_forces_twins_agree_with_reference_values(name):
    d = load_golden(name)
    f = optimize.forces.bioen_log_posterior_base(d["probe"], d["w0"], d["yTilde"], d["YTilde"], d["theta"])
    gr = optimize.forces.grad_bioen_log_posterior_base(d["probe"], d["w0"], d["yTilde"], d["YTilde"], d["theta"])
    assert rel(f, d["f_probe"]) < 5e-14
    assert grad_err(gr, d["grad_probe"]) < 1e-11
