import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
LOGW_FIXTURES = ["data_16x15", "data_deer_test_logw_M808xN10", "data_potra_part_2_logw_M205xN10",
                 "data_potra_part_1_logw_M808xN80", "data_potra_part_2_logw_M808xN10"]
FORCES_FIXTURES = ["data_deer_test_forces_M808xN10", "data_forces_M64xN64"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (d[k].item() if d[k].shape == () else d[k]) for k in d.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()  # builds liboracle.so with gcc if missing
    return O


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def grad_err(a, b):
    """max|a-b| / max|b| -- the norm-relative gradient metric ('parity metric')."""
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def select_eval_path(p, mode):
    """How a Problem launches its evaluations: 0 stand-alone kernels, 1 persistent kernel (csrc/persistent_eval.cuh),
    2 shared-memory slice kernel (csrc/slice_eval.cuh; the default for every problem small enough -- all fixtures
    of the reference's test-suite are).  Skips the test when the slice kernel cannot take the problem."""
    p.set_option(5, 1 if mode else 0)
    p.set_option(8, 1 if mode == 2 else 0)
    if mode == 2 and p.query(7) != 1:
        pytest.skip("problem not eligible for the slice kernel")
