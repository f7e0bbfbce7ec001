"""GPU, BASELINE.json's sizes.  Config 3 (N = 1e6 structures x M = 1e3 observables, 8 GB fp64, generated on the
device) and config 2 (1e5 x 500): per-evaluation parity against the reference's own C kernels on the matrix
downloaded from the device, converged optima against the reference's liblbfgs driver, plus size-independent
properties and the independent device paths (tile kernels, fused team kernels, tensor-core GEMMs) played against each
other and against NumPy on downloaded sub-blocks."""
import numpy as np
import pytest

from conftest import grad_err, rel

pytestmark = pytest.mark.gpu

M, N = 1000, 1_000_000
SEED, THETA = 12345, 10.0


@pytest.fixture(scope="module")
def big():
    import bioen_b200
    rng = np.random.default_rng(SEED)
    ytrue = rng.standard_normal(M)
    yobs = ytrue + 0.5 * rng.standard_normal(M)
    p = bioen_b200.Problem(shape=(M, N))
    p.generate(SEED, 0, ytrue / 0.5, 2.0)
    r = np.random.default_rng(7)
    w0 = r.random(N) + 0.5
    w0 /= w0.sum()
    ctx = dict(p=p, YT=yobs / 0.5, w0=w0, G=np.log(w0) - np.log(w0[-1]), g=np.log(w0) + 0.1 * r.standard_normal(N),
               f=(2e-2 / np.sqrt(M)) * r.standard_normal(M))
    yield ctx
    p.close()


def test_logw_gauge_invariance_and_directional_derivative(big):
    p = big["p"]
    p.set_logw(big["G"], big["YT"], THETA)
    f0, g0 = p.objective_and_gradient(big["g"])
    # L(g + c) = L(g)  =>  sum_j grad_j = 0   (log_weights.py:113-127: only the last log-weight is pinned on input)
    assert abs(g0.sum()) < 1e-11 * np.abs(g0).sum()
    assert rel(p.objective(big["g"] + 3.25), f0) < 1e-12
    d = np.random.default_rng(3).standard_normal(N)
    eps = 1e-4
    num = (p.objective(big["g"] + eps * d) - p.objective(big["g"] - eps * d)) / (2 * eps)
    assert rel(num, float(g0 @ d)) < 1e-6
    # run-to-run bit reproducibility at full size (the reference's fast_openmp = 0 contract)
    f1, g1 = p.objective_and_gradient(big["g"])
    assert f1 == f0 and np.array_equal(g1, g0)


def test_forces_equals_logw_at_the_same_weights(big):
    """theta*KL(w||w0) + chi2(w) through the fused two-pass forces kernels (structure-major copy) must equal the
    log-weights objective at g = log w through the tile kernels (observable-major matrix); and the chain rule
    dL/df_i = sum_j y_ij dL/dg_j links the two gradients."""
    p = big["p"]
    p.set_forces(big["w0"], big["YT"], THETA)
    ff, gf = p.objective_and_gradient(big["f"])
    w, _ = p.weights(big["f"])
    assert abs(w.sum() - 1.0) < 1e-12
    p.set_logw(np.log(big["w0"]), big["YT"], THETA)
    fl, gl = p.objective_and_gradient(np.log(w))
    assert rel(fl, ff) < 1e-11
    chain = p.average(gl)                     # Y . grad_g
    assert grad_err(chain, gf) < 1e-9


def test_forces_fused_equals_unfused_full_size(big):
    p = big["p"]
    res = []
    for fused in (0, 1):
        p.set_option(1, fused)
        p.set_forces(big["w0"], big["YT"], THETA)
        res.append(p.objective_and_gradient(big["f"]))
    assert rel(res[0][0], res[1][0]) < 1e-12
    assert grad_err(res[0][1], res[1][1]) < 1e-11


def test_passes_against_numpy_on_downloaded_blocks(big):
    """Both directions of the full-size matrix product checked exactly against NumPy on real device data."""
    p = big["p"]
    c0, nc = 517_000, 4096
    blk = p.download(0, M, c0, nc)                                   # 1000 x 4096 block of the resident matrix
    # row pass: a weight vector supported on the block only
    wblk = np.random.default_rng(1).random(nc)
    w = np.zeros(N)
    w[c0:c0 + nc] = wblk
    assert grad_err(p.average(w), blk @ wblk) < 1e-13
    # column pass: log(w_j / w_k) = x_j - x_k with x = Y^T f, for uniform prior weights
    p.set_forces(np.full(N, 1.0 / N), big["YT"], THETA)
    wf, _ = p.weights(big["f"])
    x = big["f"] @ blk
    lw = np.log(wf[c0:c0 + nc])
    assert np.max(np.abs((lw - lw[0]) - (x - x[0]))) < 1e-11
    # linearity of the row pass
    v1, v2 = np.random.default_rng(2).random((2, N))
    assert grad_err(p.average(2.0 * v1 - 3.0 * v2), 2.0 * p.average(v1) - 3.0 * p.average(v2)) < 1e-12


def test_theta_scan_planes_equal_single_runs_full_size(big):
    p = big["p"]
    p.set_logw(big["G"], big["YT"], THETA)
    thetas = np.array([100.0, 10.0, 1.0, 30.0, 3.0])
    X, fmin, codes, info = p.theta_scan(thetas, x0=big["G"], max_iterations=4)
    for q in (0, 1, 4):
        p.set_theta(thetas[q])
        x1, f1, c1, i1 = p.opt_lbfgs(big["G"], max_iterations=4)
        assert codes[q] == c1 == -997
        assert rel(fmin[q], f1) < 1e-10
        assert np.max(np.abs(X[q] - x1)) < 1e-9 * max(1.0, np.max(np.abs(x1)))


def test_lbfgs_converges_full_size(big):
    p = big["p"]
    p.set_logw(big["G"], big["YT"], THETA)
    f0 = p.objective(big["G"])
    x, fmin, code, info = p.opt_lbfgs(big["G"])
    assert code in (0, 1) and fmin < f0 and info["evaluations"] >= info["iterations"]
    assert rel(p.objective(x), fmin) < 5e-13
    w, _ = p.weights(x)
    assert abs(w.sum() - 1.0) < 1e-12 and w.min() >= 0.0
    p.set_forces(big["w0"], big["YT"], THETA)
    xf, ffin, cf, inf_ = p.opt_lbfgs(np.zeros(M))
    assert cf in (0, 1)
    # the two methods minimise the same functional over (nearly) the same set of weights
    assert abs(ffin - fmin) / fmin < 1e-2


# ---------------------------------------------------------------------------------------------------------------
# The reference's own C code on the SAME matrix (downloaded from the device), at the BASELINE sizes: config 3
# (1000 x 1e6, 8 GB) per evaluation, config 2 (500 x 1e5, 0.4 GB) per evaluation and at the converged optimum.
# north_star: objective and gradient within 1e-11 relative; converged weights within 1e-6 max-abs and the same final
# objective to 1e-8 relative.  The checker is oracle/_ref (the unmodified reference sources built by oracle/Makefile,
# shipped prebuilt), or -- where that library is absent -- the plain-C restatement oracle/bioen_oracle.c.
# ---------------------------------------------------------------------------------------------------------------
def _cpu_checker():
    import os
    from oracle import oracle as O
    from oracle import ref
    if ref.available():
        try:
            ref.set_num_threads(len(os.sched_getaffinity(0)))
        except Exception:
            pass
        ref.set_fast_openmp_flag(0)        # the reference's reproducible mode (sequential reductions)

        def logw(g, G, yT, Y, theta):
            f, gr = ref.LogwEvaluator(G, yT, Y, theta)(np.ascontiguousarray(g))
            return f, gr.copy()

        def forces(f, w0, yT, Y, theta):
            v, gr = ref.ForcesEvaluator(w0, yT, Y, theta)(np.ascontiguousarray(f))
            return v, gr.copy()
        return "reference", logw, forces
    return "port", O.logw_fg, O.forces_fg


def _need_host_ram(nbytes):
    try:
        import psutil
        if psutil.virtual_memory().available < nbytes:
            pytest.skip("needs %.0f GB of host RAM" % (nbytes / 1e9))
    except ImportError:
        pass


def test_cfg3_per_evaluation_against_the_reference_c(big):
    """1000 x 1e6: 7813 column blocks, 2*M*N*8 > 2^33 bytes per evaluation -- f and grad of both methods against the
    reference's C kernels on the identical matrix."""
    _need_host_ram(int(1.4 * M * N * 8))
    p = big["p"]
    kind, logw, forces = _cpu_checker()
    yT = p.download()
    p.set_logw(big["G"], big["YT"], THETA)
    f, g = p.objective_and_gradient(big["g"])
    fo, go = logw(big["g"], big["G"], yT, big["YT"], THETA)
    assert rel(f, fo) < 1e-11, (kind, f, fo)
    assert grad_err(g, go) < 1e-11, (kind, grad_err(g, go))
    p.set_forces(big["w0"], big["YT"], THETA)
    f, g = p.objective_and_gradient(big["f"])
    fo, go = forces(big["f"], big["w0"], yT, big["YT"], THETA)
    assert rel(f, fo) < 1e-11, (kind, f, fo)
    assert grad_err(g, go) < 1e-11, (kind, grad_err(g, go))


@pytest.fixture(scope="module")
def cfg2():
    """BASELINE config 2: N = 1e5 structures x M = 500 observables, generic-data recipe of SURVEY 8d."""
    import bioen_b200
    m, n = 500, 100_000
    rng = np.random.default_rng(SEED)
    ytrue = rng.standard_normal(m)
    yobs = ytrue + 0.5 * rng.standard_normal(m)
    p = bioen_b200.Problem(shape=(m, n))
    p.generate(SEED, 0, ytrue / 0.5, 2.0)
    ctx = dict(p=p, m=m, n=n, YT=yobs / 0.5, yT=p.download(), w0=np.full(n, 1.0 / n), G=np.zeros(n))
    yield ctx
    p.close()


def test_cfg2_per_evaluation_against_the_reference_c(cfg2):
    p, m, n = cfg2["p"], cfg2["m"], cfg2["n"]
    kind, logw, forces = _cpu_checker()
    r = np.random.default_rng(11)
    g1 = 0.1 * r.standard_normal(n)                  # SURVEY 8d parity point
    f1 = 1e-3 * r.standard_normal(m)
    for theta in (10.0, 0.1):
        p.set_logw(cfg2["G"], cfg2["YT"], theta)
        f, g = p.objective_and_gradient(g1)
        fo, go = logw(g1, cfg2["G"], cfg2["yT"], cfg2["YT"], theta)
        assert rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11, (kind, theta, rel(f, fo), grad_err(g, go))
        p.set_forces(cfg2["w0"], cfg2["YT"], theta)
        f, g = p.objective_and_gradient(f1)
        fo, go = forces(f1, cfg2["w0"], cfg2["yT"], cfg2["YT"], theta)
        assert rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11, (kind, theta, rel(f, fo), grad_err(g, go))


def test_cfg2_converged_optimum_against_the_reference_liblbfgs(cfg2):
    """Both methods minimised with BioEn's default liblbfgs settings by the device L-BFGS and by the reference's
    _opt_lbfgs_* drivers (the reference's own liblbfgs 1.10) on the identical matrix.

    north_star's bar -- final objective within 1e-8 relative, weights within 1e-6 max-abs -- is applied
      * to the forces method at its converged optimum (a well-conditioned M-dimensional problem: the reference's own
        two reduction modes end 3e-13 / 2e-10 apart there), and
      * to the log-weights method along the first 40 iterations (same code -997, same point): the minimiser logic at
        this size, before rounding differences have grown (the reference's two reduction modes on this matrix are
        7e-16 / 1e-17 apart after 3 iterations, 9e-13 / 4e-11 after 20, 5e-11 / 1e-9 after 40, 4e-7 / 5e-5 after 60:
        a factor ~10 every 5 iterations).
    The converged log-weights end point is NOT defined to that accuracy by liblbfgs's stop rule on an N = 1e5
    problem: measured here with the reference alone, fast_openmp = 0 vs 1 on this matrix end 2.8e-7 apart in f and
    9.9e-5 in max|dw| with BioEn's defaults (LBFGS_STOP after ~400 evaluations), and still 9e-8 / 3.6e-6 apart after
    2 247 vs 2 424 iterations with epsilon = 1e-7, delta = 1e-9.  So the converged log-weights
    comparison uses the reference's own mode-to-mode distance, measured in the test, as its yardstick (factor 3)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libbioen_ref.so (the reference's liblbfgs driver) is not present")
    import os
    p, m, n = cfg2["p"], cfg2["m"], cfg2["n"]
    ref.set_num_threads(len(os.sched_getaffinity(0)))
    ref.set_fast_openmp_flag(0)
    theta = THETA
    # ---- forces: converged, strict
    p.set_forces(cfg2["w0"], cfg2["YT"], theta)
    x, fmin, code, info = p.opt_lbfgs(np.zeros(m))
    xr, fr, cr = ref.opt_lbfgs_forces(np.zeros(m), cfg2["w0"], cfg2["yT"], cfg2["YT"], theta)
    assert code == cr and code in (0, 1), (code, cr)
    assert rel(fmin, fr) < 1e-8, ("forces fmin", fmin, fr, info)
    w, _ = p.weights(x)
    wr = ref.forces_weights(xr, cfg2["w0"], cfg2["yT"])
    assert np.max(np.abs(w - wr)) < 1e-6, ("forces weights", np.max(np.abs(w - wr)))
    assert rel(ref.forces_objective(x, cfg2["w0"], cfg2["yT"], cfg2["YT"], theta), fmin) < 1e-11
    fmin_forces = fmin
    # ---- log-weights: 40 iterations, strict
    p.set_logw(cfg2["G"], cfg2["YT"], theta)
    x, fmin, code, info = p.opt_lbfgs(np.zeros(n), max_iterations=40)
    xr, fr, cr = ref.opt_lbfgs_logw(np.zeros(n), cfg2["G"], cfg2["yT"], cfg2["YT"], theta, max_iterations=40)
    assert code == cr == -997, (code, cr)
    assert rel(fmin, fr) < 1e-8, ("logw 40 iterations fmin", fmin, fr)
    w, _ = p.weights(x)
    wr, _ = ref.logw_weights(xr)
    assert np.max(np.abs(w - wr)) < 1e-6, ("logw 40 iterations weights", np.max(np.abs(w - wr)))
    # ---- log-weights: BioEn defaults to the stop; yardstick = the reference against itself
    x, fmin, code, info = p.opt_lbfgs(np.zeros(n))
    xr0, fr0, cr0 = ref.opt_lbfgs_logw(np.zeros(n), cfg2["G"], cfg2["yT"], cfg2["YT"], theta)
    ref.set_fast_openmp_flag(1)
    xr1, fr1, cr1 = ref.opt_lbfgs_logw(np.zeros(n), cfg2["G"], cfg2["yT"], cfg2["YT"], theta)
    ref.set_fast_openmp_flag(0)
    assert code in (0, 1) and cr0 in (0, 1) and cr1 in (0, 1), (code, cr0, cr1)
    w, _ = p.weights(x)
    w0_, _ = ref.logw_weights(xr0)
    w1_, _ = ref.logw_weights(xr1)
    noise_f, noise_w = rel(fr1, fr0), float(np.max(np.abs(w1_ - w0_)))
    print("cfg2 logw defaults: reference mode-to-mode noise f %.2e w %.2e; device vs reference f %.2e w %.2e"
          % (noise_f, noise_w, rel(fmin, fr0), np.max(np.abs(w - w0_))))
    # liblbfgs stops here on `delta`: (f[k-10] - f[k]) / f[k] < 1e-6 -- wherever a trajectory happens to satisfy that
    # first, still ~2e-5 above the optimum (192.17358, which the forces method reaches).  End points of different
    # trajectories therefore scatter by a few 1e-6 in f (measured: the reference's own modes 2.8e-7 ... 8.1e-7 apart,
    # device vs reference 1e-7 ... 3.2e-6, depending on the evaluation path).  Bound: 10 x delta, and never tighter
    # than 3 x the reference's own scatter in this run.
    assert rel(fmin, fr0) <= max(1e-5, 3 * noise_f), ("logw fmin", fmin, fr0, fr1)
    assert np.max(np.abs(w - w0_)) <= max(5e-4, 3 * noise_w), ("logw weights", np.max(np.abs(w - w0_)), noise_w)
    # all three are equally close to the optimum that the well-conditioned forces problem pins down
    for fv in (fmin, fr0, fr1):
        assert 0.0 <= (fv - fmin_forces) / fmin_forces < 1e-4, (fv, fmin_forces)
    # whatever the trajectory, the objective the device reports at its end point is the reference's objective there
    assert rel(ref.logw_objective(x, cfg2["G"], cfg2["yT"], cfg2["YT"], theta), fmin) < 1e-11


def test_cfg5_shard_device_generated_matrix_and_sub_block_parity():
    """BASELINE config 5 (N = 1e7 x M = 5000, 400 GB) exists only sharded and only on the devices: each GPU generates
    its 5000 x 1.25e6 block (50 GB) with the counter-based generator.  Checked here on one such shard: (a) the generated
    entries equal the NumPy restatement of the generator (tests/util_rng.py) on a column block and on a row block that
    spans the whole shard, at the shard's global column offset; (b) both methods evaluated on a 5000 x 32768 column
    sub-block agree with the reference's C kernels to 1e-11; (c) the full shard evaluates reproducibly and
    gauge-invariantly (2 * M * N * 8 = 100 GB per evaluation)."""
    import bioen_b200
    import torch
    from util_rng import generic_ytilde_block
    m, n, shard = 5000, 1_250_000, 3                  # the 4th of 8 shards
    free_b, _ = torch.cuda.mem_get_info(0)
    if free_b < 60e9:
        pytest.skip("needs 60 GB of free HBM")
    _need_host_ram(int(6e9))
    rng = np.random.default_rng(SEED)
    ytrue = rng.standard_normal(m)
    YT = (ytrue + 0.5 * rng.standard_normal(m)) / 0.5
    col0 = shard * n
    with bioen_b200.Problem(shape=(m, n)) as p:
        p.generate(SEED, col0, ytrue / 0.5, 2.0)
        # (a) generator: column block (all rows) and row block (all columns of the shard)
        c0, nc = 777_000, 2048
        blk = p.download(0, m, c0, nc)
        exp = generic_ytilde_block(SEED, ytrue / 0.5, 2.0, 0, m, col0 + c0, nc)
        assert np.max(np.abs(blk - exp)) < 1e-13 * np.max(np.abs(exp))
        r0, nr = 4321, 8
        rows = p.download(r0, nr, 0, n)
        exp = generic_ytilde_block(SEED, ytrue / 0.5, 2.0, r0, nr, col0, n)
        assert np.max(np.abs(rows - exp)) < 1e-13 * np.max(np.abs(exp))
        # (b) sub-block parity against the reference C
        kind, logw, forces = _cpu_checker()
        nsub = 32768
        sub = p.download(0, m, c0, nsub)
        r = np.random.default_rng(5)
        g1 = 0.1 * r.standard_normal(nsub)
        f1 = (2e-2 / np.sqrt(m)) * r.standard_normal(m)
        w0 = np.full(nsub, 1.0 / nsub)
        with bioen_b200.Problem(sub) as q:
            q.set_logw(np.zeros(nsub), YT, THETA)
            f, g = q.objective_and_gradient(g1)
            fo, go = logw(g1, np.zeros(nsub), sub, YT, THETA)
            assert rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11, (kind, rel(f, fo), grad_err(g, go))
            q.set_forces(w0, YT, THETA)
            f, g = q.objective_and_gradient(f1)
            fo, go = forces(f1, w0, sub, YT, THETA)
            assert rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11, (kind, rel(f, fo), grad_err(g, go))
        # (c) the whole shard
        gfull = 0.1 * r.standard_normal(n)
        p.set_logw(np.zeros(n), YT, THETA)
        fa, ga = p.objective_and_gradient(gfull)
        fb, gb = p.objective_and_gradient(gfull)
        assert fa == fb and np.array_equal(ga, gb)
        assert abs(ga.sum()) < 1e-11 * np.abs(ga).sum()
        assert rel(p.objective(gfull + 1.5), fa) < 1e-12
        # the row pass on the whole shard against NumPy for weights supported on the downloaded block
        w = np.zeros(n)
        wb = r.random(nc)
        w[c0:c0 + nc] = wb
        assert grad_err(p.average(w), blk @ wb) < 1e-13
