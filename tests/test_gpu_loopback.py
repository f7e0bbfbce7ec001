"""GPU (one device is enough): the N-sharded path -- rank-local kernels, the peer-memory exchange protocol fused into
the producing kernels, the sharded L-BFGS -- run as an in-process group of 2 / 3 / 4 ranks that all live on cuda:0
(bioen_b200_comm_init_local), against the unsharded oracle and against the unsharded device path.  On a multi-GPU box
tests/test_gpu_multi.py runs the same path with one process per GPU over NCCL-bootstrapped CUDA IPC."""
import os

import numpy as np
import pytest

from conftest import grad_err, rel

pytestmark = pytest.mark.gpu

os.environ.setdefault("BIOEN_B200_P2P_TIMEOUT_S", "30")   # a rank that never arrives poisons the result after 30 s


def _problem(oracle, M, N, seed=12345):
    P = oracle.synthetic_problem(M, N, seed=seed)
    rng = np.random.default_rng(3)
    G = 0.1 * rng.standard_normal(N)
    g1 = G + 0.1 * rng.standard_normal(N)
    w0 = rng.random(N) + 0.1
    w0 /= w0.sum()
    f1 = 1e-3 * rng.standard_normal(M)
    return P, G, g1, w0, f1


@pytest.mark.parametrize("world,M,N,theta", [(2, 37, 5001, 1.0), (3, 100, 20000, 10.0), (4, 300, 4003, 3.0),
                                             (2, 1, 257, 2.0), (2, 1000, 2600, 5.0),
                                             (2, 8, 160001, 2.0)])   # wide shards: gradient formed in the column pass
def test_sharded_evaluation_on_one_gpu(oracle, world, M, N, theta):
    from bioen_b200 import dist as D
    P, G, g1, w0, f1 = _problem(oracle, M, N)
    with D.LocalGroup(P["yTilde"], world) as grp:
        assert grp.call(lambda r, p, lo, hi: p.comm_mode()) == ["p2p"] * world
        # log-weights: fused exchange (2 per evaluation), every rank gets the same f, gradient slices concatenate
        grp.call(lambda r, p, lo, hi: p.set_logw(G[lo:hi], P["YTilde"], theta))
        assert grp.call(lambda r, p, lo, hi: p.exchanges_per_eval()) == [2] * world
        res = grp.call(lambda r, p, lo, hi: p.objective_and_gradient(g1[lo:hi]))
        fo, go = oracle.logw_fg(g1, G, P["yTilde"], P["YTilde"], theta)
        assert all(f == res[0][0] for f, _ in res)
        assert rel(res[0][0], fo) < 1e-11
        assert grad_err(grp.gather([g for _, g in res]), go) < 1e-11
        # objective only, then the gradient of the same point (the split the minimisers use)
        fs = grp.call(lambda r, p, lo, hi: p.objective(g1[lo:hi]))
        assert all(rel(f, fo) < 1e-11 for f in fs)
        # the three-exchange path (what NCCL carries) gives the same numbers
        grp.call(lambda r, p, lo, hi: p.set_option(4, 0))
        assert grp.call(lambda r, p, lo, hi: p.exchanges_per_eval()) == [3] * world
        res3 = grp.call(lambda r, p, lo, hi: p.objective_and_gradient(g1[lo:hi]))
        assert rel(res3[0][0], res[0][0]) < 1e-13
        assert grad_err(grp.gather([g for _, g in res3]), grp.gather([g for _, g in res])) < 1e-12
        grp.call(lambda r, p, lo, hi: p.set_option(4, 1))
        # weights are normalised over all ranks
        ws = grp.call(lambda r, p, lo, hi: p.weights(g1[lo:hi])[0])
        wo, _ = oracle.logw_weights(g1)
        assert np.max(np.abs(grp.gather(ws) - wo)) < 1e-15
        # forces: replicated M-vector
        grp.call(lambda r, p, lo, hi: p.set_forces(w0[lo:hi], P["YTilde"], theta))
        res = grp.call(lambda r, p, lo, hi: p.objective_and_gradient(f1))
        fo, go = oracle.forces_fg(f1, w0, P["yTilde"], P["YTilde"], theta)
        assert all(rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11 for f, g in res)
        assert all(np.array_equal(g, res[0][1]) for _, g in res)
        # the forces exchanges as separate launches (what NCCL carries) give the same numbers
        grp.call(lambda r, p, lo, hi: p.set_option(4, 0))
        res3 = grp.call(lambda r, p, lo, hi: p.objective_and_gradient(f1))
        grp.call(lambda r, p, lo, hi: p.set_option(4, 1))
        assert rel(res3[0][0], res[0][0]) < 1e-13 and grad_err(res3[0][1], res[0][1]) < 1e-12
        fs = grp.call(lambda r, p, lo, hi: p.objective(f1))
        assert all(rel(f, fo) < 1e-11 for f in fs)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_lbfgs_on_one_gpu(oracle, world):
    """The sharded device L-BFGS (dot products summed over the ranks inside the vector kernels) follows the oracle's
    liblbfgs restatement: same code, same end point."""
    from bioen_b200 import dist as D
    M, N, theta = 37, 5001, 1.0
    P, G, g1, w0, f1 = _problem(oracle, M, N)
    with D.LocalGroup(P["yTilde"], world) as grp:
        grp.call(lambda r, p, lo, hi: p.set_logw(P["G"].ravel()[lo:hi], P["YTilde"], theta))
        for ls in (2, 0):
            res = grp.call(lambda r, p, lo, hi: p.opt_lbfgs(P["GInit"].ravel()[lo:hi], linesearch=ls))
            ro = oracle.lbfgs(lambda v: oracle.logw_fg(v, P["G"], P["yTilde"], P["YTilde"], theta), P["GInit"],
                              linesearch=ls)
            assert all(code == ro["code"] for _, _, code, _ in res)
            assert all(fm == res[0][1] for _, fm, _, _ in res)           # bit-identical on all ranks
            tol = 1e-8 if ro["iterations"] < 150 else 1e-4
            assert rel(res[0][1], ro["fx"]) < tol, (ls, res[0][1], ro["fx"], ro["iterations"])
        # the same minimisation with separate exchanges: same trajectory up to rounding
        grp.call(lambda r, p, lo, hi: p.set_option(4, 0))
        res3 = grp.call(lambda r, p, lo, hi: p.opt_lbfgs(P["GInit"].ravel()[lo:hi], max_iterations=30))
        grp.call(lambda r, p, lo, hi: p.set_option(4, 1))
        res2 = grp.call(lambda r, p, lo, hi: p.opt_lbfgs(P["GInit"].ravel()[lo:hi], max_iterations=30))
        assert res3[0][2] == res2[0][2] and rel(res3[0][1], res2[0][1]) < 1e-9
        # coefficient-space direction update (opt-in): one exchange of 39 doubles per iteration instead of 13 exchanges
        grp.call(lambda r, p, lo, hi: p.set_option(6, 1))
        resg = grp.call(lambda r, p, lo, hi: p.opt_lbfgs(P["GInit"].ravel()[lo:hi], max_iterations=30))
        grp.call(lambda r, p, lo, hi: p.set_option(6, 0))
        assert resg[0][2] == res2[0][2] and rel(resg[0][1], res2[0][1]) < 1e-9
        assert all(o[1] == resg[0][1] for o in resg)
        # forces
        grp.call(lambda r, p, lo, hi: p.set_forces(P["w0"].ravel()[lo:hi], P["YTilde"], theta))
        res = grp.call(lambda r, p, lo, hi: p.opt_lbfgs(P["forces_init"].ravel()))
        ro = oracle.lbfgs(lambda v: oracle.forces_fg(v, P["w0"], P["yTilde"], P["YTilde"], theta), P["forces_init"])
        assert all(code == ro["code"] for _, _, code, _ in res)
        assert rel(res[0][1], ro["fx"]) < (1e-7 if ro["iterations"] < 150 else 1e-4)


def test_sharded_matches_unsharded_at_config2_size():
    """500 x 1e5 split over 4 ranks on one GPU vs the single-context device path (same kernels, no exchange)."""
    import bioen_b200
    from bioen_b200 import dist as D
    M, N, theta = 500, 100_000, 10.0
    rng = np.random.default_rng(12345)
    ytrue = rng.standard_normal(M)
    YT = (ytrue + 0.5 * rng.standard_normal(M)) / 0.5
    with bioen_b200.Problem(shape=(M, N)) as whole:
        whole.generate(12345, 0, ytrue / 0.5, 2.0)
        yT = whole.download()
        g1 = 0.1 * rng.standard_normal(N)
        whole.set_logw(np.zeros(N), YT, theta)
        f0, g0 = whole.objective_and_gradient(g1)
        with D.LocalGroup(yT, 4) as grp:
            grp.call(lambda r, p, lo, hi: p.set_logw(np.zeros(hi - lo), YT, theta))
            res = grp.call(lambda r, p, lo, hi: p.objective_and_gradient(g1[lo:hi]))
            assert rel(res[0][0], f0) < 1e-12
            assert grad_err(grp.gather([g for _, g in res]), g0) < 1e-12
            out = grp.call(lambda r, p, lo, hi: p.opt_lbfgs(np.zeros(hi - lo), max_iterations=25))
        x1, f1, c1, _ = whole.opt_lbfgs(np.zeros(N), max_iterations=25)
        assert out[0][2] == c1 == -997
        assert rel(out[0][1], f1) < 1e-9
        assert np.max(np.abs(grp.gather([o[0] for o in out]) - x1)) < 1e-8 * max(1.0, np.max(np.abs(x1)))
