"""compute-sanitizer driver for the slice kernel (memcheck / racecheck): all three launch modes, both methods, at
shapes that exercise every thread layout (one CTA, narrow / wide slices, odd tails).
    compute-sanitizer --tool racecheck python tests/tools/sanitize_slice.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bioen_b200
rng = np.random.default_rng(0)
for (M, N) in ((7, 3), (17, 9), (808, 10), (100, 1000), (28, 5001), (5, 45001)):
    y = rng.standard_normal((M, N))
    Y = rng.standard_normal(M)
    w0 = rng.random(N) + 0.1
    w0 /= w0.sum()
    G = np.log(w0)
    with bioen_b200.Problem(y) as p:
        assert p.query(7) == 1
        p.set_logw(G, Y, 3.0)
        f, g = p.objective_and_gradient(G + 0.1)
        assert p.objective(G + 0.1) == f and np.array_equal(p.gradient(G + 0.1), g)
        p.opt_lbfgs(G, max_iterations=3)
        p.set_forces(w0, Y, 3.0)
        f, g = p.objective_and_gradient(np.full(M, 1e-3))
        assert p.objective(np.full(M, 1e-3)) == f and np.array_equal(p.gradient(np.full(M, 1e-3)), g)
        p.opt_lbfgs(np.zeros(M), max_iterations=3)
        assert p.query(6) > 0
print("sanitize_slice done")
