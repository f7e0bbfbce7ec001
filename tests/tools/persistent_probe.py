"""Diagnostic (GPU box): device-timed evaluations at small sizes, persistent kernel on / off.  Run under ncu
(--metrics gpu__time_duration.sum) to separate in-kernel time from launch overhead."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bioen_b200  # noqa: E402

dev = torch.device("cuda", 0)
steps = int(os.environ.get("STEPS", "20"))
for (M, N) in ((28, 50001), (500, 100000)):
    rng = np.random.default_rng(1)
    a = rng.standard_normal(M)
    YT = a + rng.standard_normal(M)
    with bioen_b200.Problem(shape=(M, N)) as p:
        p.generate(12345, 0, a, 2.0)
        for mode in (1, 0):
            p.set_option(5, mode)
            p.set_option(1, 0 if M < 256 else 1)
            for meth, name in ((0, "logw"), (1, "forces")):
                if meth == 0:
                    p.set_logw(np.zeros(N), YT, 10.0)
                else:
                    p.set_forces(np.full(N, 1.0 / N), YT, 10.0)
                n = N if meth == 0 else M
                x = torch.from_numpy((0.1 if meth == 0 else 1e-3) * rng.standard_normal(n)).to(dev)
                g = torch.zeros_like(x)
                ms, pass_ms, launches = p.time_evals(x.data_ptr(), g.data_ptr(), 3, steps, meth)
                print("M=%d N=%d %-6s persistent=%d: %.2f us per f+g evaluation, %d launches" %
                      (M, N, name, mode, 1e3 * ms / steps, launches // steps), flush=True)
