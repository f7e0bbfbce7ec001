"""GPU: nuisance-parameter refits on the resident matrix (bioen_b200/nuisance.py + bioen_b200_affine_rows) against
the reference's own refit code run on its DEER and scattering test data (fixtures: tests/golden/nuisance_*.npz,
generator tests/golden/make_nuisance_golden.py; reference: bioen/analyze/observables/observables.py:110-229)."""
import numpy as np
import pytest

from conftest import grad_err, load_golden, rel

pytestmark = pytest.mark.gpu


def _blocks(kind, d, value=None):
    from bioen_b200 import nuisance as NU
    rows = d["raw"].shape[0]
    v = float(d["m0"] if kind == "deer" else d["c0"]) if value is None else value
    return NU, [NU.Block(kind, 0, rows, err=d["err"], exp_fit=d["exp_fit"], value=v, name=kind)]


@pytest.mark.parametrize("kind", ["deer", "scattering"])
def test_refit_on_device_matches_the_reference(kind):
    import bioen_b200
    d = load_golden("nuisance_" + kind)
    for k, w in enumerate(d["weights"]):
        NU, blocks = _blocks(kind, d)
        base = NU.base_matrix(d["raw"], d["err"])
        with bioen_b200.Problem(NU.proc_sim(base, blocks)) as p:
            assert np.max(np.abs(p.download() - d["sim0"]) / np.abs(d["sim0"])) < 1e-14
            nr = NU.NuisanceRefit(p, blocks)
            for q, v in enumerate(d["probes"]):                         # moddepth_fit / coeff_fit: one row pass each
                assert rel(nr.chi2(blocks[0], v, w), d["chi2"][k, q]) < 1e-11
            k0 = p.kernels_launched()
            values = nr.update(w)                                       # update_sim: ONE row pass + ONE affine pass
            assert p.kernels_launched() - k0 <= 3
            assert rel(values[kind], d["fitted"][k]) < 1e-6
            got = p.download()
            exact = NU.proc_sim(base, blocks)                           # the blocks now carry the fitted value
            assert np.max(np.abs(got - exact) / np.abs(exact)) < 1e-13
            assert np.max(np.abs(got - d["sim_fitted"][k]) / np.abs(d["sim_fitted"][k])) < 1e-5
            # the problem keeps working on the transformed matrix: both methods against a fresh upload of it
            n = got.shape[1]
            Y = d["YTilde"].ravel()
            g1 = 0.3 * np.random.default_rng(k).standard_normal(n)
            f1 = 1e-3 * np.random.default_rng(k + 5).standard_normal(got.shape[0])
            w0 = np.full(n, 1.0 / n)
            with bioen_b200.Problem(got) as fresh:
                for setter, x in ((lambda q: q.set_logw(np.zeros(n), Y, 2.0), g1),
                                  (lambda q: q.set_forces(w0, Y, 2.0), f1)):
                    setter(p)
                    setter(fresh)
                    fa, ga = p.objective_and_gradient(x)
                    fb, gb = fresh.objective_and_gradient(x)
                    assert rel(fa, fb) < 1e-13 and grad_err(ga, gb) < 1e-12


def test_residual_helper_and_iterated_refits():
    """chi2_residuals is the row pass seen through a row-affine transform; repeated refit/commit cycles (the reference
    iterates weights <-> nuisance parameters, procedure.py:62-83) do not drift away from a fresh rebuild."""
    import bioen_b200
    d = load_golden("nuisance_deer")
    NU, blocks = _blocks("deer", d, value=0.2)
    base = NU.base_matrix(d["raw"], d["err"])
    rng = np.random.default_rng(3)
    with bioen_b200.Problem(NU.proc_sim(base, blocks)) as p:
        w = d["weights"][1]
        s, o = blocks[0].scale_offset(0.7)
        r = p.chi2_residuals(w, scale=s / 0.2, offset=o - (1.0 - 0.2) / d["err"] * (s / 0.2), YTilde=d["exp_fit"] / d["err"])
        assert rel(0.5 * float(r @ r), blocks[0].chi2(0.7, base @ w, w.sum())) < 1e-11
        nr = NU.NuisanceRefit(p, blocks)
        for it in range(12):
            w = rng.random(w.size) + 0.05
            w /= w.sum()
            nr.update(w)
        exact = NU.proc_sim(base, blocks)
        assert np.max(np.abs(p.download() - exact) / np.abs(exact)) < 1e-13      # ~1 ulp per committed refit at most
        with pytest.raises(RuntimeError):
            import torch
            t = torch.zeros((4, 16), dtype=torch.float64, device="cuda")
            q = bioen_b200.Problem(shape=(4, 16))
            q.adopt(t.data_ptr(), 16)
            q.affine_rows(np.ones(4), np.zeros(4))       # caller-owned matrix: refused


@pytest.mark.parametrize("kind", ["deer", "scattering"])
def test_weights_nuisance_iteration_matches_the_reference_workflow(kind):
    """The loop of bioen/analyze/procedure.py:40-83 -- optimise the weights, refit the nuisance parameter against
    them, rebuild the matrix, repeat; two theta values x four iterations -- recorded with the COMPLETE reference stack
    (its Python, its Cython module, its OpenMP kernels, liblbfgs; make_nuisance_golden.py) and repeated here on ONE
    resident matrix: device L-BFGS through the public find_optimum(problem=...), refit + in-place row-affine commit."""
    import bioen_b200
    from bioen_b200 import optimize
    from bioen_b200 import nuisance as NU
    d = load_golden("nuisance_" + kind)
    n = d["raw"].shape[1]
    start = float(d["m0"]) if kind == "deer" else NU.INITIAL
    NUm, blocks = _blocks(kind, d, value=start)
    base = NU.base_matrix(d["raw"], d["err"])
    cfg = optimize.minimize.Parameters("lbfgs")
    cfg["verbose"] = False
    cfg["params"]["epsilon"] = 1e-7
    cfg["params"]["delta"] = 1e-10
    cfg["params"]["max_iterations"] = 20000
    w0 = np.full((n, 1), 1.0 / n)
    log_w0 = optimize.log_weights.getGs(w0)
    log_wopt = optimize.log_weights.getGs(w0.copy())
    YT = d["YTilde"].reshape(1, -1)
    with bioen_b200.Problem(NU.proc_sim(base, blocks)) as P:
        refit = NU.NuisanceRefit(P, blocks)
        values = [refit.update(w0.ravel())[kind]]                         # update_sim_init
        k = 0
        for theta in d["wf_thetas"]:
            for it in range(int(d["wf_iterations"])):
                yT = P.download()                                         # only for the shape checks of the public API
                res = optimize.log_weights.find_optimum(log_wopt, log_w0, yT, yT, YT, float(theta), cfg, problem=P)
                wopt = res[0].ravel()
                assert rel(res[4], d["wf_fmin"][k]) < 1e-7, (theta, it, res[4], d["wf_fmin"][k])
                assert np.max(np.abs(wopt - d["wf_weights"][k])) < 1e-5, (theta, it)
                wmd = wopt.copy()
                wmd[wmd == 0.0] = 1e-150
                values.append(refit.update(wmd)[kind])
                k += 1
    assert np.max(np.abs(np.array(values) / d["wf_values"] - 1.0)) < 1e-5, (values, d["wf_values"])
