"""GPU: per-evaluation parity of the CUDA path with the CPU oracle and the reference's golden values.

Tolerances (north_star): objective |df|/|f| < 1e-11, gradient max|dg|/max|g| < 1e-11, fp64.
Every call goes through the C ABI: part 1 (reference-compatible host-pointer symbols, via the c_bioen mirror)
and part 2 (resident handle, bioen_b200.Problem).
"""
import numpy as np
import pytest

from conftest import FORCES_FIXTURES, LOGW_FIXTURES, grad_err, load_golden, rel

pytestmark = pytest.mark.gpu

TOL = 1e-11


def _problem(yT):
    import bioen_b200
    return bioen_b200.Problem(yT)


@pytest.mark.parametrize("name", LOGW_FIXTURES)
def test_logw_golden_part1(name):
    from bioen_b200.optimize.ext import c_bioen
    d = load_golden(name)
    for x, fk, gk in ((d["GInit"], "f_init", "grad_init"), (d["probe"], "f_probe", "grad_probe")):
        f = c_bioen.bioen_log_posterior_logw(x.ravel(), d["G"], d["G"], d["yTilde"], d["YTilde"], d["theta"])
        g = c_bioen.grad_bioen_log_posterior_logw(x.ravel(), d["G"], d["G"], d["yTilde"], d["YTilde"], d["theta"])
        assert rel(f, d[fk]) < TOL
        assert grad_err(g, d[gk]) < TOL


@pytest.mark.parametrize("name", FORCES_FIXTURES)
def test_forces_golden_part1(name):
    from bioen_b200.optimize.ext import c_bioen
    d = load_golden(name)
    for x, fk, gk in ((d["forces_init"], "f_init", "grad_init"), (d["probe"], "f_probe", "grad_probe")):
        f = c_bioen.bioen_log_posterior_forces(x.ravel(), d["w0"], d["yTilde"], d["YTilde"], d["theta"])
        g = c_bioen.grad_bioen_log_posterior_forces(x.ravel(), d["w0"], d["yTilde"], d["YTilde"], d["theta"])
        assert rel(f, d[fk]) < TOL
        assert grad_err(g, d[gk]) < TOL


@pytest.mark.parametrize("name", LOGW_FIXTURES + FORCES_FIXTURES)
def test_golden_resident_handle(name):
    d = load_golden(name)
    with _problem(d["yTilde"]) as p:
        if d["kind"] == "logw":
            p.set_logw(d["G"], d["YTilde"], d["theta"])
        else:
            p.set_forces(d["w0"], d["YTilde"], d["theta"])
        f, g = p.objective_and_gradient(d["probe"])
        assert rel(f, d["f_probe"]) < TOL and grad_err(g, d["grad_probe"]) < TOL
        assert rel(p.objective(d["probe"]), d["f_probe"]) < TOL          # objective-only variant
        w, _ = p.weights(d["probe"])
        assert np.max(np.abs(w - d["w_probe"])) < 1e-14
        avg = p.average(w)
        assert grad_err(avg, d["yTilde"] @ d["w_probe"]) < 1e-13


def test_weights_part1(oracle):
    import ctypes as C
    from bioen_b200 import _lib
    rng = np.random.default_rng(5)
    g = rng.standard_normal(1000)
    w = np.empty_like(g)
    s = _lib.load()._get_weights(_lib.ptr(g), _lib.ptr(w), 1000)
    _lib.check_pending("_get_weights")
    w_ref, s_ref = oracle.logw_weights(g)
    assert rel(s, s_ref) < 1e-13 and np.max(np.abs(w - w_ref)) < 1e-15


# ragged / tiny / tile-boundary shapes: tiles are 32 rows x 128 columns, rows are padded to 16 doubles
SHAPES = [(1, 1), (1, 2), (2, 1), (3, 5), (31, 127), (32, 128), (33, 129), (64, 4096), (5, 20000),
          (100, 1000), (257, 3001), (500, 2049), (1000, 777)]


@pytest.mark.parametrize("M,N", SHAPES)
def test_synthetic_shapes(oracle, M, N):
    P = oracle.synthetic_problem(M, N, seed=100 + M + N)
    rng = np.random.default_rng(M * 7 + N)
    G = 0.2 * rng.standard_normal(N)
    g1 = G + 0.1 * rng.standard_normal(N)
    w0 = rng.random(N) + 0.1
    w0 /= w0.sum()
    f1 = 1e-3 * rng.standard_normal(M)
    theta = 3.7
    with _problem(P["yTilde"]) as p:
        p.set_logw(G, P["YTilde"], theta)
        f, g = p.objective_and_gradient(g1)
        fo, go = oracle.logw_fg(g1, G, P["yTilde"], P["YTilde"], theta)
        assert rel(f, fo) < TOL and grad_err(g, go) < TOL
        assert rel(p.objective(g1), fo) < TOL
        p.set_forces(w0, P["YTilde"], theta)
        f, g = p.objective_and_gradient(f1)
        fo, go = oracle.forces_fg(f1, w0, P["yTilde"], P["YTilde"], theta)
        assert rel(f, fo) < TOL and grad_err(g, go) < TOL
        assert rel(p.objective(f1), fo) < TOL
        w, _ = p.weights(f1)
        assert np.max(np.abs(w - oracle.forces_weights(f1, w0, P["yTilde"]))) < 1e-14
        # reference semantics: objective / gradient for GIVEN weights
        f2, g2 = p.forces_from_weights(w)
        assert rel(f2, fo) < TOL and grad_err(g2, go) < TOL


def test_golden_synthetic_reference_values(oracle):
    for name in ("synthetic_M37xN5001", "synthetic_M100xN20000"):
        d = load_golden(name)
        P = oracle.synthetic_problem(int(d["M"]), int(d["N"]), seed=int(d["seed"]))
        g1 = 0.1 * np.random.default_rng(1).standard_normal(P["N"])
        f1 = 1e-3 * np.random.default_rng(2).standard_normal(P["M"])
        with _problem(P["yTilde"]) as p:
            p.set_logw(P["G"], P["YTilde"], d["theta"])
            f, g = p.objective_and_gradient(g1)
            assert rel(f, d["logw_f"]) < TOL and grad_err(g, d["logw_grad"]) < TOL
            p.set_forces(P["w0"], P["YTilde"], d["theta"])
            f, g = p.objective_and_gradient(f1)
            assert rel(f, d["forces_f"]) < TOL and grad_err(g, d["forces_grad"]) < TOL


def test_large_log_weights_are_stabilised(oracle):
    """g up to +-600: exp(g) still fits fp64, the reference's un-stabilised sum works, so must we."""
    P = oracle.synthetic_problem(20, 3000, seed=3)
    rng = np.random.default_rng(9)
    g1 = rng.uniform(-600, 600, 3000)
    with _problem(P["yTilde"]) as p:
        p.set_logw(P["G"], P["YTilde"], 2.0)
        f, g = p.objective_and_gradient(g1)
    fo, go = oracle.logw_fg(g1, P["G"], P["yTilde"], P["YTilde"], 2.0)
    assert rel(f, fo) < TOL and grad_err(g, go) < TOL


def test_zero_weights_guard_forces(oracle):
    """w0_j = 0 entries take the DBL_MIN guard of c_bioen_kernels_forces.c:252-256."""
    P = oracle.synthetic_problem(11, 500, seed=4)
    w0 = np.full(500, 1.0 / 400)
    w0[::5] = 0.0
    f1 = 1e-2 * np.random.default_rng(1).standard_normal(11)
    with _problem(P["yTilde"]) as p:
        p.set_forces(w0, P["YTilde"], 5.0)
        f, g = p.objective_and_gradient(f1)
    fo, go = oracle.forces_fg(f1, w0, P["yTilde"], P["YTilde"], 5.0)
    assert rel(f, fo) < TOL and grad_err(g, go) < TOL


def test_run_to_run_bit_reproducible(oracle):
    """GPU counterpart of the reference's fast_openmp=0 mode (test_logw_reproducibility.py): identical bits."""
    P = oracle.synthetic_problem(300, 50000, seed=8)
    g1 = 0.1 * np.random.default_rng(2).standard_normal(50000)
    with _problem(P["yTilde"]) as p:
        p.set_logw(P["G"], P["YTilde"], 10.0)
        f0, g0 = p.objective_and_gradient(g1)
        for _ in range(5):
            f, g = p.objective_and_gradient(g1)
            assert f == f0 and np.array_equal(g, g0)


def test_device_generator_is_the_documented_hash():
    """bench.py generates yTilde on the device; util_rng restates the generator in NumPy."""
    from util_rng import generic_ytilde_block
    import bioen_b200
    M, N = 13, 1000
    a = np.linspace(-1, 1, M)
    with bioen_b200.Problem(shape=(M, N)) as p:
        p.generate(12345, 7000, a, 2.0)
        Y = p.download()
        blk = p.download(3, 4, 100, 50)
    ref = generic_ytilde_block(12345, a, 2.0, 0, M, 7000, N)
    assert np.max(np.abs(Y - ref)) < 1e-12
    assert np.array_equal(blk, Y[3:7, 100:150])
    assert abs(np.mean((Y - a[:, None]) / 2.0)) < 0.05 and abs(np.std((Y - a[:, None]) / 2.0) - 1) < 0.05


# fused two-pass forces kernels (structure-major copy, csrc/fused_pass.cuh): eligible for 256 <= M <= 8192;
# slab width C = 8 (M <= 768) ... 1 (M > 3072), register tiling KI = 1 ... 16
FUSED_SHAPES = [(256, 1000), (300, 9), (511, 4097), (1000, 20000), (1500, 3000), (3073, 700), (4100, 300),
                (8192, 70), (8193, 70)]


@pytest.mark.parametrize("M,N", FUSED_SHAPES)
def test_forces_fused_vs_unfused_vs_oracle(oracle, M, N):
    import bioen_b200
    P = oracle.synthetic_problem(M, N, seed=M + N)
    rng = np.random.default_rng(M)
    w0 = rng.random(N) + 0.05
    w0 /= w0.sum()
    f1 = (2e-2 / np.sqrt(M)) * rng.standard_normal(M)
    theta = 2.5
    fo, go = oracle.forces_fg(f1, w0, P["yTilde"], P["YTilde"], theta)
    res = {}
    for fused in (1, 0):
        with bioen_b200.Problem(P["yTilde"]) as p:
            p.set_option(1, fused)
            p.set_forces(w0, P["YTilde"], theta)
            f, g = p.objective_and_gradient(f1)
            assert rel(f, fo) < TOL and grad_err(g, go) < TOL, (fused, rel(f, fo), grad_err(g, go))
            assert rel(p.objective(f1), fo) < TOL
            res[fused] = (f, g, p.kernels_launched())
            # a second, different point through the same context (accumulators / barriers are reusable)
            f2, g2 = p.objective_and_gradient(-0.5 * f1)
            fo2, go2 = oracle.forces_fg(-0.5 * f1, w0, P["yTilde"], P["YTilde"], theta)
            assert rel(f2, fo2) < TOL and grad_err(g2, go2) < TOL
    assert rel(res[1][0], res[0][0]) < 1e-13


def test_forces_fused_minimiser_and_reproducibility(oracle):
    import bioen_b200
    P = oracle.synthetic_problem(400, 30000, seed=77)
    with bioen_b200.Problem(P["yTilde"]) as p:
        p.set_forces(P["w0"], P["YTilde"], 10.0)
        x, fmin, code, info = p.opt_lbfgs(P["forces_init"])
        r = oracle.lbfgs(lambda v: oracle.forces_fg(v, P["w0"], P["yTilde"], P["YTilde"], 10.0), P["forces_init"])
        assert code == r["code"] and rel(fmin, r["fx"]) < 1e-8, (code, r["code"], fmin, r["fx"], info)
        f0, g0 = p.objective_and_gradient(x)
        for _ in range(3):
            f, g = p.objective_and_gradient(x)
            assert f == f0 and np.array_equal(g, g0)


def test_chunked_upload_equals_full_upload(oracle):
    import bioen_b200
    P = oracle.synthetic_problem(45, 3003, seed=6)
    g1 = 0.2 * np.random.default_rng(1).standard_normal(3003)
    with bioen_b200.Problem(P["yTilde"]) as p:
        p.set_logw(P["G"], P["YTilde"], 4.0)
        want = p.objective_and_gradient(g1)
    with bioen_b200.Problem(shape=(45, 3003)) as p:
        for r0 in range(0, 45, 16):
            p.upload_rows(r0, P["yTilde"][r0:r0 + 16])
        assert np.array_equal(p.download(), P["yTilde"])
        p.set_logw(P["G"], P["YTilde"], 4.0)
        got = p.objective_and_gradient(g1)
        with pytest.raises(RuntimeError, match="out of range"):
            p.upload_rows(40, P["yTilde"][:16])
    assert got[0] == want[0] and np.array_equal(got[1], want[1])


def test_pinned_host_arrays(oracle):
    import gc
    import bioen_b200
    P = oracle.synthetic_problem(20, 1500, seed=2)
    yp = bioen_b200.pinned_empty((20, 1500))
    yp[:] = P["yTilde"]
    g1 = 0.1 * np.random.default_rng(0).standard_normal(1500)
    with bioen_b200.Problem(yp) as p:
        p.set_logw(P["G"], P["YTilde"], 2.0)
        f, g = p.objective_and_gradient(g1)
    fo, go = oracle.logw_fg(g1, P["G"], P["yTilde"], P["YTilde"], 2.0)
    assert rel(f, fo) < TOL and grad_err(g, go) < TOL
    view = yp[3:5]
    del yp
    gc.collect()
    assert np.array_equal(view, P["yTilde"][3:5])      # the pinned block lives as long as any view of it


def test_problem_use_after_close_raises(oracle):
    import bioen_b200
    P = oracle.synthetic_problem(5, 40, seed=1)
    p = bioen_b200.Problem(P["yTilde"])
    p.set_logw(P["G"], P["YTilde"], 1.0)
    p.close()
    p.close()                                   # idempotent
    with pytest.raises(RuntimeError, match="after close"):
        p.objective(np.zeros(40))
    with pytest.raises(ValueError):
        bioen_b200.Problem(np.zeros(7))         # not a matrix
    with pytest.raises(RuntimeError):
        bioen_b200.Problem(shape=(0, 5))        # empty problems are refused by the library


def test_reported_error_does_not_leak_into_the_next_stateless_call(oracle):
    """An error that was raised through a status code must not make the next reference-style (part 1) call fail."""
    import bioen_b200
    from bioen_b200.optimize.ext import c_bioen
    P = oracle.synthetic_problem(6, 50, seed=2)
    with pytest.raises(RuntimeError):
        bioen_b200.Problem(shape=(0, 3))
    with bioen_b200.Problem(P["yTilde"]) as p:
        with pytest.raises(RuntimeError):
            p.upload_rows(5, P["yTilde"][:4])
    f = c_bioen.bioen_log_posterior_logw(np.zeros(50), np.zeros(50), np.zeros(50), P["yTilde"], P["YTilde"], 2.0)
    fo, _ = oracle.logw_fg(np.zeros(50), np.zeros(50), P["yTilde"], P["YTilde"], 2.0)
    assert rel(f, fo) < TOL


@pytest.mark.parametrize("M,N", [(37, 5001), (300, 3000)])
def test_gradient_after_objective_runs_the_gradient_half_only(oracle, M, N):
    """SciPy-style callers ask f(x) and then fprime(x): Problem.gradient continues the pending objective-only
    evaluation (bioen_b200_grad_continue).  Bit-identical to a full evaluation; anything in between voids it."""
    P = oracle.synthetic_problem(M, N, seed=7)
    rng = np.random.default_rng(1)
    with _problem(P["yTilde"]) as p:
        for setter, x in ((lambda: p.set_logw(P["G"], P["YTilde"], 4.0), np.ravel(P["G"]) + 0.1 * rng.standard_normal(N)),
                          (lambda: p.set_forces(P["w0"], P["YTilde"], 4.0), 1e-3 * rng.standard_normal(M))):
            setter()
            f_full, g_full = p.objective_and_gradient(x)
            k0 = p.kernels_launched()
            f = p.objective(x)
            k1 = p.kernels_launched()
            g = p.gradient(x)                       # continuation
            k2 = p.kernels_launched()
            assert f == f_full and np.array_equal(g, g_full)
            g2 = p.gradient(x)                      # nothing pending any more: full evaluation
            k3 = p.kernels_launched()
            assert np.array_equal(g2, g_full)
            if p.query(3):    # persistent kernel: objective half, gradient half and a full evaluation are one launch each
                assert (k1 - k0, k2 - k1, k3 - k2) == (1, 1, 1)
            else:
                assert (k2 - k1) < (k3 - k2) and (k1 - k0) + (k2 - k1) == (k3 - k2)
            p.objective(x)
            p.weights(x)                            # moves the device state: the probe is void
            assert np.array_equal(p.gradient(x), g_full)
            p.objective(x)
            y = x.copy()
            y[0] += 1e-3                            # another point: full evaluation there
            assert np.array_equal(p.gradient(y), p.objective_and_gradient(y)[1])
            assert p._lib.bioen_b200_grad_continue(p._ctx, p.method, None) == 2
