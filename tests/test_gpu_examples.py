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


def test_c_abi_demo_runs(tmp_path):
    import shutil
    from bioen_b200 import _lib
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    _lib.load()
    exe = tmp_path / "c_abi_demo"
    libdir = os.path.dirname(_lib.library_path())
    subprocess.run([gcc, "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"),
                    "-o", str(exe), "-L", libdir, "-lbioen_b200", "-lm", "-Wl,-rpath," + libdir], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "c_abi_demo ok" in res.stdout, res.stdout + res.stderr


@pytest.mark.parametrize("kind", ["deer", "scattering"])
def test_nuisance_refit_example(kind):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "nuisance_refit.py"), kind],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "reference stack, same loop" in res.stdout and res.stdout.count("iteration") == 8
