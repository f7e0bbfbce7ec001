"""Generate tests/golden/nuisance_{deer,scattering}.npz with the REFERENCE's own nuisance-parameter code.

Run in the build container only (needs /root/reference):  python tests/golden/make_nuisance_golden.py

The reference refits the DEER modulation depth / the scattering scale between two weight optimisations
(bioen/analyze/observables/observables.py:146-216, called from bioen/analyze/procedure.py:82-83).  This script
imports that module unchanged (h5py, which the image lacks and fileio imports unconditionally, is stubbed: no HDF5 is
touched), fills an Observables object with the reference's own test data (test/deer/data, test/scattering/data;
parsed the way deer.py:89-155 / scattering.py do) and records what its methods return: the processed matrices
(get_proc_sim, get_proc_exp), chi^2 probes (moddepth_fit, coeff_fit) and whole refits (update_sim, update_sim_init).
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, REF)
from bioen.analyze.observables import observables as RO  # noqa: E402


class Bag:
    pass


def make_obs(experiment, payload, models):
    obs = RO.Observables.__new__(RO.Observables)
    obs.experiments = [experiment]
    obs.models_list = np.asarray(models, dtype=float)
    obs.observables = {experiment: payload}
    obs.nrestraints = payload.nrestraints
    obs.exp = RO.get_proc_exp(obs)
    return obs


def deer():
    path = os.path.join(REF, "test", "deer", "data")
    models = np.loadtxt(os.path.join(path, "models-deer.dat"))[:10]
    ln = "319-259"
    d = Bag()
    d.labels = [[319, 259]]
    exp = np.genfromtxt(os.path.join(path, "exp-319-259-deer.dat"), comments="#")
    d.exp_tmp = {ln: exp}
    d.exp_err_tmp = {ln: np.array([0.2] * len(exp))}                      # exp-error.dat: "319-259 0.2"
    d.moddepth = {ln: 0.41928739}                                          # moddepth-deer.dat
    d.nrestraints = len(exp)
    d.sim_tmp = {m: {ln: np.genfromtxt(os.path.join(path, "conf%d-319-259-deer.dat" % int(m)))[:, 1]} for m in models}
    obs = make_obs("deer", d, models)
    n = len(models)
    rng = np.random.default_rng(7)
    ws = [np.full((n, 1), 1.0 / n)]
    w = rng.random((n, 1)) + 0.05
    ws.append(w / w.sum())
    out = dict(raw=np.stack([d.sim_tmp[m][ln] for m in models], axis=1), err=d.exp_err_tmp[ln],
               exp_fit=exp[:, 1], exp_opt=exp[:, 2], m0=d.moddepth[ln], YTilde=np.asarray(obs.exp),
               weights=np.stack([np.asarray(w).ravel() for w in ws]))
    sim, _ = obs.get_proc_sim()
    out["sim0"] = np.asarray(sim)
    probes = np.array([0.05, 0.2, 0.41928739, 0.6, 0.95])
    out["probes"] = probes
    out["chi2"] = np.array([[obs.moddepth_fit(m, ln, np.matrix(w)) for m in probes] for w in ws])
    fitted, sims = [], []
    for w in ws:
        d.moddepth = {ln: 0.41928739}
        sim, _ = obs.update_sim(np.matrix(w))
        fitted.append(float(np.ravel(d.moddepth[ln])[0]))
        sims.append(np.asarray(sim))
    out["fitted"] = np.array(fitted)
    out["sim_fitted"] = np.stack(sims)
    # "initial-optimization": 0.15 start (observables.py:219-229)
    d.moddepth = {ln: "initial-optimization"}
    obs.update_sim_init(np.matrix(ws[1]))
    out["fitted_from_initial"] = float(np.ravel(d.moddepth[ln])[0])
    np.savez_compressed(os.path.join(HERE, "nuisance_deer.npz"), **out)
    print("deer: rows", len(exp), "models", n, "fitted", out["fitted"], out["fitted_from_initial"])


def scattering():
    path = os.path.join(REF, "test", "scattering", "data")
    models = np.arange(5.0)
    s = Bag()
    exp = np.genfromtxt(os.path.join(path, "lyz-exp.dat"), comments="#")
    s.exp_tmp = exp
    err = exp[:, 2].copy()
    err[err == 0.0] = 0.01
    s.exp_err_tmp = err
    s.nrestraints = len(exp)
    s.scaling_factor = 2.0e-6
    s.sim_tmp = {m: np.genfromtxt(os.path.join(path, "lyz%d-sim-saxs.dat" % int(m)))[:, 1] for m in models}
    obs = make_obs("scattering", s, models)
    n = len(models)
    rng = np.random.default_rng(11)
    ws = [np.full((n, 1), 1.0 / n)]
    w = rng.random((n, 1)) + 0.05
    ws.append(w / w.sum())
    out = dict(raw=np.stack([s.sim_tmp[m] for m in models], axis=1), err=err, exp_fit=exp[:, 1], exp_opt=exp[:, 1],
               c0=s.scaling_factor, YTilde=np.asarray(obs.exp), weights=np.stack([np.asarray(w).ravel() for w in ws]))
    sim, _ = obs.get_proc_sim()
    out["sim0"] = np.asarray(sim)
    probes = np.array([1.0e-6, 2.0e-6, 2.03e-6, 2.5e-6, 2.0e-4])
    out["probes"] = probes
    out["chi2"] = np.array([[obs.coeff_fit(c, np.matrix(w)) for c in probes] for w in ws])
    fitted, sims = [], []
    for w in ws:
        s.scaling_factor = 2.0e-6
        sim, _ = obs.update_sim(np.matrix(w))
        fitted.append(float(np.ravel(s.scaling_factor)[0]))
        sims.append(np.asarray(sim))
    out["fitted"] = np.array(fitted)
    out["sim_fitted"] = np.stack(sims)
    s.scaling_factor = "initial-optimization"                # 0.0002 start (observables.py:226-228)
    obs.update_sim_init(np.matrix(ws[1]))
    out["fitted_from_initial"] = float(np.ravel(s.scaling_factor)[0])
    np.savez_compressed(os.path.join(HERE, "nuisance_scattering.npz"), **out)
    print("scattering: rows", len(exp), "models", n, "fitted", out["fitted"], out["fitted_from_initial"])


if __name__ == "__main__":
    deer()
    scattering()
