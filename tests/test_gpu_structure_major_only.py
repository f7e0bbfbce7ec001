"""GPU: BIOEN_B200_OPT_STRUCTURE_MAJOR_ONLY -- the context holds only the structure-major copy of yTilde that the fused
two-pass forces kernels read (half the device memory of the default forces set-up, which keeps both layouts; what lets
BASELINE config 5 run the forces method on 4 GPUs).  Same kernels, same data: evaluations and minimisers must agree
with the default set-up bit for bit, and with the oracle to 1e-11."""
import numpy as np
import pytest

from conftest import grad_err, rel

pytestmark = pytest.mark.gpu
TOL = 1e-11
OPT_YT_ONLY = 11


@pytest.mark.parametrize("M,N", [(256, 1000), (301, 3001), (1000, 777), (2049, 515), (4100, 300)])
def test_structure_major_only_matches_default_and_oracle(oracle, M, N):
    import bioen_b200
    P = oracle.synthetic_problem(M, N, seed=M + N)
    rng = np.random.default_rng(M * 3 + N)
    w0 = rng.random(N) + 0.1
    w0 /= w0.sum()
    f1 = 1e-3 * rng.standard_normal(M)
    theta = 4.2
    fo, go = oracle.forces_fg(f1, w0, P["yTilde"], P["YTilde"], theta)
    with bioen_b200.Problem(P["yTilde"]) as p:            # default: both layouts resident
        p.set_option(5, 0)                                 # (small shapes: keep them on the fused kernels)
        p.set_forces(w0, P["YTilde"], theta)
        assert p.query(0) == 1
        fd, gd = p.objective_and_gradient(f1)
        xd = p.opt_lbfgs(np.zeros(M), max_iterations=25)
        both = p.query(8)
    ldt = (M + 1) // 2 * 2
    for how in ("before upload", "after upload", "row chunks"):
        if how == "before upload":
            p = bioen_b200.Problem(P["yTilde"], structure_major_only=True)
        elif how == "after upload":
            p = bioen_b200.Problem(P["yTilde"])
            p.set_option(OPT_YT_ONLY, 1)
        else:
            p = bioen_b200.Problem(shape=(M, N), structure_major_only=True)
            for r0 in range(0, M, 97):
                p.upload_rows(r0, P["yTilde"][r0:r0 + 97])
        with p:
            assert p.query(9) == 1 and p.query(8) == N * ldt * 8 and p.query(8) < both
            p.set_forces(w0, P["YTilde"], theta)
            assert p.query(0) == 1 and p.query(3) == 0 and p.query(7) == 0
            f, g = p.objective_and_gradient(f1)
            assert f == fd and np.array_equal(g, gd), how          # same kernels on the same data
            assert rel(f, fo) < TOL and grad_err(g, go) < TOL
            assert p.objective(f1) == f and np.array_equal(p.gradient(f1), g)
            x = p.opt_lbfgs(np.zeros(M), max_iterations=25)
            assert x[2] == xd[2] and x[1] == xd[1] and np.array_equal(x[0], xd[0]), how
            w, _ = p.weights(f1)
            assert np.max(np.abs(w - oracle.forces_weights(f1, w0, P["yTilde"]))) < 1e-14
            v = rng.standard_normal(N)                             # averages of an arbitrary vector
            assert grad_err(p.average(v), P["yTilde"] @ v) < 1e-13
            blk = p.download(3, 40, 5, 200)
            assert np.array_equal(blk, P["yTilde"][3:43, 5:205])
            assert np.array_equal(p.download(), P["yTilde"])


def test_structure_major_only_generator_and_errors(oracle):
    import bioen_b200
    M, N = 300, 4099
    a = np.random.default_rng(1).standard_normal(M)
    with bioen_b200.Problem(shape=(M, N)) as p:
        p.generate(99, 1000, a, 2.0)
        ref = p.download()
    with bioen_b200.Problem(shape=(M, N), structure_major_only=True) as p:
        p.generate(99, 1000, a, 2.0)
        assert np.array_equal(p.download(), ref)                   # the generator writes the same matrix
        w0 = np.full(N, 1.0 / N)
        with pytest.raises(RuntimeError, match="row-major"):
            p.set_logw(np.zeros(N), a, 1.0)
        with pytest.raises(RuntimeError, match="row-major"):
            p.affine_rows(np.ones(M), np.zeros(M))
        p.set_forces(w0, a, 1.0)
        with pytest.raises(RuntimeError, match="row-major"):
            p.theta_scan([1.0, 2.0])
        with pytest.raises(RuntimeError, match="row-major"):
            p.forces_from_weights(w0)
        with pytest.raises(RuntimeError):
            p.set_option(OPT_YT_ONLY, 0)
        f, g = p.objective_and_gradient(np.zeros(M))
        fo, go = oracle.forces_fg(np.zeros(M), w0, ref, a[None, :], 1.0)
        assert rel(f, fo) < TOL and grad_err(g, go) < TOL
    with bioen_b200.Problem(shape=(100, 500)) as p:               # M < 256: no fused kernels, no such mode
        with pytest.raises(RuntimeError, match="256"):
            p.set_option(OPT_YT_ONLY, 1)


@pytest.mark.parametrize("minimizer,algorithm", [("lbfgs", ""), ("gsl", "bfgs2"), ("scipy", "lbfgs")])
def test_find_optimum_on_a_structure_major_only_problem(oracle, minimizer, algorithm):
    """The reference API (optimize.forces.find_optimum) on a resident problem that holds only the structure-major
    copy: every step of it -- initial objective, minimiser, weights, averages -- runs on the fused kernels and returns
    what the default set-up returns."""
    import bioen_b200
    from bioen_b200 import optimize
    M, N, theta = 300, 2500, 5.0
    P = oracle.synthetic_problem(M, N, seed=77)
    cfg = optimize.minimize.Parameters(minimizer)
    cfg["verbose"] = False
    if algorithm:
        cfg["algorithm"] = algorithm
    if minimizer == "gsl":
        cfg["params"]["max_iterations"] = 40
    res = {}
    for only in (False, True):
        with bioen_b200.Problem(P["yTilde"], structure_major_only=only) as p:
            p.set_option(5, 0)       # the default set-up on the fused kernels as well (not the slice kernel)
            res[only] = optimize.forces.find_optimum(P["forces_init"], P["w0"], P["yTilde"], P["yTilde"], P["YTilde"],
                                                     theta, cfg, problem=p)
    a, b = res[False], res[True]
    assert b[3] == a[3] and b[4] == a[4]                              # fmin_initial, fmin_final
    assert np.array_equal(b[2], a[2])                                 # forces at the optimum
    assert np.max(np.abs(b[0] - a[0])) < 1e-15                        # weights
    assert np.max(np.abs(b[1] - a[1])) < 1e-12 * np.max(np.abs(a[1]))  # y . wopt (fused-kernel average vs row pass)
    assert abs(b[5] - a[5]) < 1e-11 * abs(a[5]) and abs(b[6] - a[6]) < 1e-11 * max(abs(a[6]), 1e-12)
