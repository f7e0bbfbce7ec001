"""GPU: the example scripts run end to end."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("method", ["forces", "log_weights"])
def test_theta_series_example(method):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "theta_series.py"), "--structures", "6001",
                          "--observables", "28", "--thetas", "12", "--method", method],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "batched find_optimum_series" in res.stdout
