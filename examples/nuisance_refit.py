#!/usr/bin/env python
"""The reference's iterative procedure for data with a nuisance parameter (bioen/analyze/procedure.py:40-83: optimise
the weights, refit the DEER modulation depth / the scattering scale against them, rebuild the matrix, repeat) on ONE
resident copy of the matrix.

    python examples/nuisance_refit.py [deer|scattering]

Data: the reference's own DEER (462 time points x 10 conformers) and lysozyme SAXS (197 q-points x 5 models) test
sets as stored in tests/golden/nuisance_*.npz.  Per iteration the GPU does one device L-BFGS, ONE row pass for the
whole least-squares refit (every trial value of the parameter is then O(rows) on the host) and one in-place
row-affine pass that commits the refitted value -- no rebuild of the matrix on the host, no second upload.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bioen_b200  # noqa: E402
from bioen_b200 import nuisance as NU, optimize  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "deer"
    d = np.load(os.path.join(ROOT, "tests", "golden", "nuisance_%s.npz" % kind))
    rows, n = d["raw"].shape
    start = float(d["m0"]) if kind == "deer" else NU.INITIAL
    blocks = [NU.Block(kind, 0, rows, err=d["err"], exp_fit=d["exp_fit"], value=start, name=kind)]
    base = NU.base_matrix(d["raw"], d["err"])                    # s_ij / err_i: the parameter-free matrix
    cfg = optimize.minimize.Parameters("lbfgs")
    cfg["verbose"] = False
    w0 = np.full((n, 1), 1.0 / n)
    G = optimize.log_weights.getGs(w0)
    YTilde = d["YTilde"].reshape(1, -1)
    with bioen_b200.Problem(NU.proc_sim(base, blocks)) as P:     # = Observables.get_proc_sim(), uploaded once
        refit = NU.NuisanceRefit(P, blocks)
        print("start value        %-22s -> refit against w0: %.8g" % (start, refit.update(w0.ravel())[kind]))
        for theta in (10.0, 1.0):
            for it in range(4):
                yT = P.download()
                wopt, yopt, gopt, f0, f1 = optimize.log_weights.find_optimum(G, G, yT, yT, YTilde, theta, cfg, problem=P)
                value = refit.update(wopt.ravel())[kind]
                print("theta %5g  iteration %d  objective %.8f  parameter %.8g" % (theta, it, f1, value))
    ref = d["wf_values"]
    print("reference stack, same loop: parameter %.8g ... %.8g" % (ref[1], ref[-1]))


if __name__ == "__main__":
    main()
