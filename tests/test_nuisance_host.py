"""CPU: the host logic of the nuisance-parameter refits (bioen_b200/nuisance.py) against values produced by the
reference's own code (bioen/analyze/observables/observables.py:110-229, fixtures made by
tests/golden/make_nuisance_golden.py on the reference's DEER and scattering test data).  The device is replaced by a
NumPy stand-in with Problem's two methods the refit uses; tests/test_gpu_nuisance.py runs the same on the GPU."""
import numpy as np
import pytest

from conftest import load_golden, rel


class HostProblem:
    def __init__(self, yTilde):
        self.y = np.array(yTilde, dtype=np.float64)
        self.m, self.n = self.y.shape
        self.transforms = 0

    def average(self, w):
        return self.y @ np.asarray(w, dtype=np.float64).ravel()

    def affine_rows(self, scale, offset):
        self.y = scale[:, None] * self.y + offset[:, None]
        self.transforms += 1


def _blocks(kind, d):
    from bioen_b200 import nuisance as NU
    rows = d["raw"].shape[0]
    value = d["m0"] if kind == "deer" else d["c0"]
    return NU, [NU.Block(kind, 0, rows, err=d["err"], exp_fit=d["exp_fit"], value=float(value), name=kind)]


@pytest.mark.parametrize("kind", ["deer", "scattering"])
def test_matrix_probes_and_refit_match_the_reference(kind):
    d = load_golden("nuisance_" + kind)
    NU, blocks = _blocks(kind, d)
    base = NU.base_matrix(d["raw"], d["err"])
    y0 = NU.proc_sim(base, blocks)
    assert np.max(np.abs(y0 - d["sim0"]) / np.abs(d["sim0"])) < 1e-14          # get_proc_sim
    for k, w in enumerate(d["weights"]):
        blocks[0].value = float(d["m0"] if kind == "deer" else d["c0"])
        prob = HostProblem(NU.proc_sim(base, blocks))
        nr = NU.NuisanceRefit(prob, blocks)
        for q, v in enumerate(d["probes"]):                                     # moddepth_fit / coeff_fit
            assert rel(nr.chi2(blocks[0], v, w), d["chi2"][k, q]) < 1e-11
        values = nr.update(w)                                                   # update_sim
        assert rel(values[kind], d["fitted"][k]) < 1e-6, (values, d["fitted"][k])
        assert prob.transforms == 1
        # the committed matrix is the reference's rebuilt one (up to the difference of the fitted values)
        blocks_ref = _blocks(kind, d)[1]
        blocks_ref[0].value = float(d["fitted"][k])
        assert np.max(np.abs(prob.y - d["sim_fitted"][k]) / np.abs(d["sim_fitted"][k])) < 1e-5
        assert np.max(np.abs(NU.proc_sim(base, blocks_ref) - d["sim_fitted"][k]) / np.abs(d["sim_fitted"][k])) < 1e-13
        # a second refit against the same weights is a fixed point; nothing is transformed
        again = nr.fit(w)
        assert rel(again[kind], values[kind]) < 1e-6


@pytest.mark.parametrize("kind", ["deer", "scattering"])
def test_initial_optimization_start_values(kind):
    d = load_golden("nuisance_" + kind)
    NU, blocks = _blocks(kind, d)
    blocks[0].value = NU.INITIAL
    base = NU.base_matrix(d["raw"], d["err"])
    prob = HostProblem(NU.proc_sim(base, blocks))
    nr = NU.NuisanceRefit(prob, blocks)
    assert blocks[0].value == NU.INITIAL_VALUES[kind]
    values = nr.update(d["weights"][1])                                         # update_sim_init
    assert rel(values[kind], d["fitted_from_initial"]) < 1e-6


def test_mixed_blocks_leave_fixed_rows_alone():
    from bioen_b200 import nuisance as NU
    dd, ds = load_golden("nuisance_deer"), load_golden("nuisance_scattering")
    n = 5
    raw = np.vstack([dd["raw"][:, :n], ds["raw"], dd["raw"][:7, :n] * 3.0])
    err = np.concatenate([dd["err"], ds["err"], np.ones(7)])
    r1, r2 = dd["raw"].shape[0], dd["raw"].shape[0] + ds["raw"].shape[0]
    blocks = [NU.Block("deer", 0, r1, err=dd["err"], exp_fit=dd["exp_fit"], value=0.3, name="label"),
              NU.Block("scattering", r1, r2, err=ds["err"], exp_fit=ds["exp_fit"], value=2e-6, name="saxs"),
              NU.Block("fixed", r2, r2 + 7)]
    base = NU.base_matrix(raw, err)
    prob = HostProblem(NU.proc_sim(base, blocks))
    fixed_before = prob.y[r2:].copy()
    nr = NU.NuisanceRefit(prob, blocks)
    w = np.full(n, 1.0 / n)
    values = nr.update(w)
    assert set(values) == {"label", "saxs"} and np.array_equal(prob.y[r2:], fixed_before)
    ref = NU.proc_sim(base, blocks)            # blocks now carry the fitted values
    assert np.max(np.abs(prob.y - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-12
    with pytest.raises(ValueError):
        nr.commit({"label": 0.0})
    with pytest.raises(ValueError):
        NU.Block("deer", 0, 3, err=np.ones(2), exp_fit=np.ones(3), value=0.1)
