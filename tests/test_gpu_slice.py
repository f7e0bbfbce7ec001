"""GPU: the shared-memory slice kernel (csrc/slice_eval.cuh) -- one cooperative launch per evaluation, the columns of
yTilde dealt to the CTAs' shared memory, one grid barrier.  It is the default path for every problem small enough
(all fixtures of the reference's test-suite, the ala5 example), so the golden-vector tests of test_gpu_eval.py and
the minimiser tests already run on it; this file pins that fact and covers what is specific to the kernel: the
geometry of the column deal (narrow / wide slices, odd tails, a single CTA), CTAs with very different local maxima,
zero prior weights, the gradient-only continuation, eligibility and the switches."""
import numpy as np
import pytest

from conftest import FORCES_FIXTURES, LOGW_FIXTURES, grad_err, load_golden, rel

pytestmark = pytest.mark.gpu
TOL = 1e-11
OPT_PERSISTENT, OPT_SLICE = 5, 8


@pytest.mark.parametrize("name", LOGW_FIXTURES + FORCES_FIXTURES)
def test_reference_fixtures_run_on_the_slice_kernel(name):
    """Every fixture of the reference's test-suite is slice-eligible and evaluated by the slice kernel by default:
    stored outputs of the reference's own C at both stored points (1e-11)."""
    import bioen_b200
    d = load_golden(name)
    with bioen_b200.Problem(d["yTilde"]) as p:
        assert p.query(7) == 1
        if d["kind"] == "logw":
            p.set_logw(d["G"], d["YTilde"], d["theta"])
            pts = ((d["GInit"], "f_init", "grad_init"), (d["probe"], "f_probe", "grad_probe"))
        else:
            p.set_forces(d["w0"], d["YTilde"], d["theta"])
            pts = ((d["forces_init"], "f_init", "grad_init"), (d["probe"], "f_probe", "grad_probe"))
        assert p.pass_kernel_name() == "slice_eval_kernel"
        n0 = p.query(6)
        for x, fk, gk in pts:
            f, g = p.objective_and_gradient(np.asarray(x).ravel())
            assert rel(f, d[fk]) < TOL and grad_err(g, d[gk]) < TOL
            f1 = p.objective(np.asarray(x).ravel())
            g1 = p.gradient(np.asarray(x).ravel())          # continuation of the objective-only launch
            assert f1 == f and np.array_equal(g1, g)
        assert p.query(6) - n0 == 6 and p.kernels_launched() >= 6


# (M, N): one CTA; 2 CTAs with an odd tail; tall and narrow (808 rows x 8 columns per CTA, 32 row groups in the column
# reduce); exactly 148 x 16 columns; one column more; wide slices (338, 514 columns); slices near the shared-memory
# limit (1000 x 22, 27 x 1000-column equivalents)
SHAPES = [(1, 1), (7, 3), (17, 8), (17, 9), (808, 10), (808, 100), (205, 10), (1000, 15), (3, 296), (5, 2368), (7, 2369),
          (64, 300), (100, 1000), (1000, 777), (28, 50001), (40, 76000), (9, 150001), (2, 400000)]


@pytest.mark.parametrize("M,N", SHAPES)
def test_slice_shapes_against_the_oracle(oracle, M, N):
    import bioen_b200
    P = oracle.synthetic_problem(M, N, seed=7 * M + N)
    rng = np.random.default_rng(M + 13 * N)
    G = 0.2 * rng.standard_normal(N)
    g1 = G + 0.1 * rng.standard_normal(N)
    w0 = rng.random(N) + 0.1
    w0 /= w0.sum()
    f1 = 1e-3 * rng.standard_normal(M)
    theta = 2.5
    with bioen_b200.Problem(P["yTilde"]) as p:
        assert p.query(7) == 1, "expected to be slice-eligible"
        p.set_logw(G, P["YTilde"], theta)
        n0 = p.query(6)
        f, g = p.objective_and_gradient(g1)
        fo, go = oracle.logw_fg(g1, G, P["yTilde"], P["YTilde"], theta)
        assert rel(f, fo) < TOL and grad_err(g, go) < TOL
        w, _ = p.weights(g1)
        fonly = p.objective(g1)
        assert fonly == f and np.array_equal(p.gradient(g1), g)
        assert np.max(np.abs(p.debug_read(1, N) - w)) < 1e-15          # weights left on the device by the kernel
        f2, g2 = p.objective_and_gradient(g1)
        assert f2 == f and np.array_equal(g2, g)                       # run-to-run bits
        p.set_forces(w0, P["YTilde"], theta)
        f, g = p.objective_and_gradient(f1)
        fo, go = oracle.forces_fg(f1, w0, P["yTilde"], P["YTilde"], theta)
        assert rel(f, fo) < TOL and grad_err(g, go) < TOL
        assert np.max(np.abs(p.debug_read(1, N) - oracle.forces_weights(f1, w0, P["yTilde"]))) < 1e-14
        fonly = p.objective(f1)
        assert fonly == f and np.array_equal(p.gradient(f1), g)
        f2, g2 = p.objective_and_gradient(f1)
        assert f2 == f and np.array_equal(g2, g)
        assert p.query(6) - n0 == 8


def test_slice_local_maxima_far_apart(oracle):
    """Log-weights that differ by hundreds between the column ranges of different CTAs (every CTA exponentiates
    against its OWN maximum; the combination rescales): no overflow, no loss against the oracle's global softmax."""
    import bioen_b200
    M, N, theta = 20, 6000, 1.3
    P = oracle.synthetic_problem(M, N, seed=3)
    rng = np.random.default_rng(4)
    G = rng.standard_normal(N)
    g1 = G + rng.standard_normal(N)
    g1[:1500] += 600.0
    g1[1500:3000] -= 600.0
    g1[3000:3100] += 595.0
    with bioen_b200.Problem(P["yTilde"]) as p:
        assert p.query(7) == 1
        p.set_logw(G, P["YTilde"], theta)
        f, g = p.objective_and_gradient(g1)
        fo, go = oracle.logw_fg(g1, G, P["yTilde"], P["YTilde"], theta)
        assert np.isfinite(f) and rel(f, fo) < TOL and grad_err(g, go) < TOL
        # forces that spread x_j = sum_i f_i y_ij over e^+-8 (different maxima in different CTAs)
        f1 = 0.3 * rng.standard_normal(M)
        w0 = rng.random(N) + 0.1
        w0 /= w0.sum()
        p.set_forces(w0, P["YTilde"], theta)
        f, g = p.objective_and_gradient(f1)
        fo, go = oracle.forces_fg(f1, w0, P["yTilde"], P["YTilde"], theta)
        assert np.isfinite(f) and rel(f, fo) < TOL and grad_err(g, go) < TOL


def test_slice_zero_prior_weights(oracle):
    """w0_j = 0 for a block of structures (a whole CTA's columns among them): the reference's DBL_MIN guard
    (c_bioen_kernels_forces.c:156-171) -- log-ratio 0, no NaN from 0 * log 0."""
    import bioen_b200
    M, N, theta = 12, 3000, 4.0
    P = oracle.synthetic_problem(M, N, seed=9)
    rng = np.random.default_rng(10)
    w0 = rng.random(N) + 0.1
    w0[100:400] = 0.0
    w0[::7] = 0.0
    w0 /= w0.sum()
    f1 = 1e-2 * rng.standard_normal(M)
    with bioen_b200.Problem(P["yTilde"]) as p:
        assert p.query(7) == 1
        p.set_forces(w0, P["YTilde"], theta)
        f, g = p.objective_and_gradient(f1)
        fo, go = oracle.forces_fg(f1, w0, P["yTilde"], P["YTilde"], theta)
        assert np.isfinite(f) and np.all(np.isfinite(g))
        assert rel(f, fo) < TOL and grad_err(g, go) < TOL
        w = p.debug_read(1, N)
        assert np.all(w[100:400] == 0.0)
        assert np.all(p.debug_read(4, N)[100:400] == 0.0)               # guarded log-ratio


def test_slice_eligibility_and_switches(oracle):
    import bioen_b200
    rng = np.random.default_rng(0)
    # too large for shared memory: 300 x 272 columns per CTA
    with bioen_b200.Problem(shape=(300, 40000)) as p:
        p.generate(1, 0, np.zeros(300), 1.0)
        assert p.query(7) == 0 and p.query(3) == 1
    y = rng.standard_normal((10, 500))
    with bioen_b200.Problem(y) as p:
        assert p.query(7) == 1 and p.query(3) == 1
        p.set_option(OPT_SLICE, 0)
        assert p.query(7) == 0 and p.query(3) == 1                      # the persistent kernel takes over
        p.set_option(OPT_SLICE, 1)
        p.set_option(OPT_PERSISTENT, 0)
        assert p.query(7) == 0 and p.query(3) == 0                      # stand-alone kernels
        p.set_option(OPT_PERSISTENT, -1)
        assert p.query(7) == 1
        p.set_option(7, 1)                                              # fp32 storage: tile kernels only
        assert p.query(7) == 0


def test_slice_minimisers_on_the_ala5_shape(oracle):
    """Device L-BFGS on the slice kernel at the ala5 shape: trial points are formed inside the kernel (xp + stp d),
    the lazy-gradient line search uses the objective-only and gradient-only launches; end point against the oracle's
    liblbfgs restatement."""
    import bioen_b200
    M, N, theta = 28, 50001, 10.0
    P = oracle.synthetic_problem(M, N, seed=12345)
    with bioen_b200.Problem(P["yTilde"]) as p:
        assert p.query(7) == 1
        p.set_forces(P["w0"], P["YTilde"], theta)
        n0 = p.query(6)
        x, fmin, code, info = p.opt_lbfgs(P["forces_init"])
        assert p.query(6) - n0 >= info["evaluations"]
        ro = oracle.lbfgs(lambda v: oracle.forces_fg(v, P["w0"], P["yTilde"], P["YTilde"], theta), P["forces_init"])
        noise = {-998, -1001, -1000, -999, -996}
        assert code == ro["code"] or code in noise or ro["code"] in noise
        assert rel(fmin, ro["fx"]) < (1e-8 if ro["iterations"] < 150 else 1e-4)
        p.set_logw(P["G"], P["YTilde"], theta)
        x, fmin, code, info = p.opt_lbfgs(P["GInit"], max_iterations=40)
        ro = oracle.lbfgs(lambda v: oracle.logw_fg(v, P["G"], P["yTilde"], P["YTilde"], theta), P["GInit"],
                          max_iterations=40)
        assert code == ro["code"] and rel(fmin, ro["fx"]) < 1e-8
