"""GPU: opt-in fp32 STORAGE of yTilde (BIOEN_B200_OPT_FP32_STORAGE).  The matrix entries are rounded
to fp32 once; every product and sum stays fp64.  This is NOT parity-preserving and never the default: the tests pin
(a) that the kernels compute exactly what fp64 kernels compute on the fp32-rounded matrix (1e-11, i.e. the only
difference to the default path is the rounding of the entries), and (b) how far that rounding moves the results."""
import numpy as np
import pytest

from conftest import grad_err, rel

pytestmark = pytest.mark.gpu
OPT_FP32, OPT_PERSISTENT = 7, 5


# (8, 80001) is wide enough for the stand-alone column pass to carry the gradient epilogue (stream_colgrad_kernel<float>)
@pytest.mark.parametrize("M,N", [(3, 5), (33, 129), (64, 4096), (28, 50001), (257, 3001), (300, 40000), (8, 80001)])
@pytest.mark.parametrize("persistent", [0, 1])
def test_fp32_storage_computes_fp64_on_the_rounded_matrix(oracle, M, N, persistent):
    import bioen_b200
    P = oracle.synthetic_problem(M, N, seed=100 + M + N)
    rng = np.random.default_rng(M * 7 + N)
    G = 0.2 * rng.standard_normal(N)
    g1 = G + 0.1 * rng.standard_normal(N)
    w0 = rng.random(N) + 0.1
    w0 /= w0.sum()
    f1 = 1e-3 * rng.standard_normal(M)
    theta = 3.7
    y32 = P["yTilde"].astype(np.float32).astype(np.float64)
    with bioen_b200.Problem(P["yTilde"]) as p:
        p.set_option(OPT_PERSISTENT, persistent)
        p.set_option(OPT_FP32, 1)
        assert p.query(4) == 4
        p.set_logw(G, P["YTilde"], theta)
        f, g = p.objective_and_gradient(g1)
        fo, go = oracle.logw_fg(g1, G, y32, P["YTilde"], theta)
        assert rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11
        fe, ge = oracle.logw_fg(g1, G, P["yTilde"], P["YTilde"], theta)
        assert rel(f, fe) < 1e-5 and grad_err(g, ge) < 1e-4             # what the storage rounding costs
        p.set_forces(w0, P["YTilde"], theta)
        assert p.query(0) == 0                                          # tile passes: no structure-major fp64 copy
        f, g = p.objective_and_gradient(f1)
        fo, go = oracle.forces_fg(f1, w0, y32, P["YTilde"], theta)
        assert rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11
        assert grad_err(p.average(w0), y32 @ w0) < 1e-13
        with pytest.raises(RuntimeError):
            p.download()
        with pytest.raises(RuntimeError):
            p.set_option(OPT_FP32, 0)


def test_fp32_storage_minimisation_and_measured_error(oracle):
    import bioen_b200
    M, N, theta = 100, 20000, 10.0
    P = oracle.synthetic_problem(M, N, seed=12345)
    with bioen_b200.Problem(P["yTilde"]) as p:
        p.set_logw(P["G"], P["YTilde"], theta)
        x64, f64, c64, _ = p.opt_lbfgs(P["GInit"])
        w64, _ = p.weights(x64)
        p.set_option(OPT_FP32, 1)
        p.set_logw(P["G"], P["YTilde"], theta)
        x32, f32, c32, _ = p.opt_lbfgs(P["GInit"])
        w32, _ = p.weights(x32)
        assert c64 in (0, 1) and c32 in (0, 1)
        print("fp32 storage: optimum objective moves by %.2e relative, weights by %.2e max-abs"
              % (rel(f32, f64), np.max(np.abs(w32 - w64))))
        assert rel(f32, f64) < 1e-5 and np.max(np.abs(w32 - w64)) < 1e-4
        # a new fp64 upload returns the context to full precision
        p.upload_rows(0, P["yTilde"])
        assert p.query(4) == 8
        p.set_logw(P["G"], P["YTilde"], theta)
        fo, go = oracle.logw_fg(P["GInit"].ravel(), P["G"], P["yTilde"], P["YTilde"], theta)
        f, g = p.objective_and_gradient(P["GInit"])
        assert rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11
