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
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, REF)


def _load_reference_stack():
    """The reference package with ITS compiled extension: `make -C oracle refext` links the reference's Cython module
    against the reference's own C library; it is injected under the name the package imports."""
    import glob
    import importlib
    import importlib.util
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "refext"])
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cpu", "c_bioen*.so"))[0]
    import bioen.optimize.ext as E
    spec = importlib.util.spec_from_file_location("bioen.optimize.ext.c_bioen", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["bioen.optimize.ext.c_bioen"] = mod
    E.c_bioen = mod
    import bioen.optimize
    importlib.reload(bioen.optimize)
    return bioen.optimize


OPT = _load_reference_stack()
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


def workflow(obs, payload_get, payload_set, start_value, thetas, iterations, n):
    """The reference's iteration of bioen/analyze/procedure.py:40-83 (log-weights method, liblbfgs through the
    reference's own C stack, tight stop so that the optimum -- not the stop rule -- defines the result)."""
    optimize = OPT
    params = optimize.minimize.Parameters("lbfgs")
    params["verbose"] = False
    params["params"]["epsilon"] = 1e-7
    params["params"]["delta"] = 1e-10
    params["params"]["max_iterations"] = 20000
    w0 = np.matrix(np.full((n, 1), 1.0 / n))
    winit = w0.copy()
    wopt = winit.copy()
    log_w0 = optimize.log_weights.getGs(w0)
    exp = obs.exp.copy()
    payload_set(start_value)
    sim, sim_init = obs.update_sim_init(wopt)
    log_wopt = optimize.log_weights.getGs(winit)
    values, weights, fmins = [float(np.ravel(payload_get())[0])], [], []
    for theta in thetas:
        for i in range(iterations):
            out = optimize.log_weights.find_optimum(log_wopt, log_w0, sim_init, sim, exp, theta, params)
            wopt = out[0]
            wopt_md = np.matrix(wopt.copy())
            wopt_md[wopt_md == 0.0] = 1e-150
            sim, sim_init = obs.update_sim(wopt_md)
            values.append(float(np.ravel(payload_get())[0]))
            weights.append(np.asarray(wopt).ravel().copy())
            fmins.append(float(out[4]))
    return dict(wf_values=np.array(values), wf_weights=np.stack(weights), wf_fmin=np.array(fmins),
                wf_thetas=np.array(thetas, dtype=float), wf_iterations=iterations)


def add_workflows():
    """Append the weights <-> nuisance iteration to both fixtures."""
    for kind in ("deer", "scattering"):
        fn = os.path.join(HERE, "nuisance_%s.npz" % kind)
        d = dict(np.load(fn))
        n = d["raw"].shape[1]
        models = np.arange(float(n))
        bag = Bag()
        bag.nrestraints = d["raw"].shape[0]
        if kind == "deer":
            ln = "319-259"
            bag.labels = [[319, 259]]
            exp = np.zeros((bag.nrestraints, 3))
            exp[:, 1], exp[:, 2] = d["exp_fit"], d["exp_opt"]
            bag.exp_tmp = {ln: exp}
            bag.exp_err_tmp = {ln: d["err"].copy()}
            bag.moddepth = {ln: float(d["m0"])}
            bag.sim_tmp = {m: {ln: d["raw"][:, j].copy()} for j, m in enumerate(models)}
            get = lambda: bag.moddepth[ln]
            setv = lambda v: bag.moddepth.__setitem__(ln, v)
            start = float(d["m0"])
        else:
            exp = np.zeros((bag.nrestraints, 3))
            exp[:, 1], exp[:, 2] = d["exp_fit"], d["err"]
            bag.exp_tmp = exp
            bag.exp_err_tmp = d["err"].copy()
            bag.scaling_factor = float(d["c0"])
            bag.sim_tmp = {m: d["raw"][:, j].copy() for j, m in enumerate(models)}
            get = lambda: bag.scaling_factor
            setv = lambda v: setattr(bag, "scaling_factor", v)
            start = "initial-optimization"
        obs = make_obs(kind, bag, models)
        d.update(workflow(obs, get, setv, start, [10.0, 1.0], 4, n))
        np.savez_compressed(fn, **d)
        print(kind, "workflow values", d["wf_values"], "fmin", d["wf_fmin"])


if __name__ == "__main__":
    deer()
    scattering()
    add_workflows()
