"""CPU: libbioen_b200.so builds for sm_100a, loads, exports every symbol include/bioen_b200.h declares, and the
host-only entry points behave; compute entry points fail LOUDLY without a GPU (no fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT

import bioen_b200
from bioen_b200 import _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "bioen_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return {n for n in names if not n.startswith("defined")}


def test_header_and_binding_agree():
    hdr = header_functions()
    assert hdr == set(_lib.SIGNATURES), (hdr ^ set(_lib.SIGNATURES))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in header_functions():
        assert hasattr(lib, name), name


def test_reference_extern_block_is_covered():
    # the C symbols c_bioen.pyx binds (bioen/optimize/ext/c_bioen.pyx:10-129)
    needed = ["_get_weights", "_bioen_log_posterior_logw", "_grad_bioen_log_posterior_logw", "_opt_bfgs_logw",
              "_opt_lbfgs_logw", "_get_weights_from_forces", "_bioen_log_posterior_forces",
              "_grad_bioen_log_posterior_forces", "_opt_bfgs_forces", "_opt_lbfgs_forces", "_library_gsl",
              "_library_lbfgs", "_omp_set_num_threads", "_set_fast_openmp_flag", "_get_fast_openmp_flag",
              "bioen_gsl_error", "lbfgs_strerror"]
    lib = _lib.load()
    for n in needed:
        assert hasattr(lib, n)


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib._library_gsl() == 1 and lib._library_lbfgs() == 1
    lib._set_fast_openmp_flag(1)
    assert lib._get_fast_openmp_flag() == 1
    lib._set_fast_openmp_flag(0)
    assert lib._get_fast_openmp_flag() == 0
    lib._omp_set_num_threads(3)
    assert lib.lbfgs_strerror(0) == b"Convergence reached."
    assert lib.lbfgs_strerror(1) == b"LBFGS_STOP"
    assert lib.lbfgs_strerror(-997).startswith(b"The algorithm routine reaches the maximum number")
    assert lib.lbfgs_strerror(-1015) == b"Invalid parameter lbfgs_parameter_t::delta specified."
    assert lib.lbfgs_strerror(5) == b"(unknown)"
    assert lib.bioen_gsl_error(27) == b"iteration is not making progress towards solution"
    assert lib.bioen_gsl_error(-2) == b"the iteration has not converged yet"


def test_error_strings_match_reference_library_if_built():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    lib = _lib.load()
    for code in list(range(-1024, -993)) + [0, 1, 2, 7]:
        assert lib.lbfgs_strerror(code).decode() == ref.lbfgs_strerror(code), code
    for code in range(-2, 33):
        assert lib.bioen_gsl_error(code).decode() == ref.gsl_strerror(code), code


def test_struct_layouts_match_reference_header():
    # c_bioen_common.h:44-92 on LP64: params_t = 8 ptrs, double, ptr, int(+pad), 2 ptrs, 2 ints
    import ctypes as C
    assert C.sizeof(_lib.params_t) == 8 * 8 + 8 + 8 + 8 + 16 + 8
    assert C.sizeof(_lib.gsl_config_params) == 24
    assert C.sizeof(_lib.lbfgs_config_params) == 56
    assert C.sizeof(_lib.visual_params) == 16
    assert _lib.params_t.theta.offset == 64 and _lib.params_t.m.offset == 104 and _lib.params_t.n.offset == 108


@pytest.mark.skipif(bioen_b200.device_count() > 0, reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_gpu():
    y = np.ones((3, 5))
    with pytest.raises(RuntimeError, match="bioen_b200"):
        bioen_b200.Problem(y)
    from bioen_b200.optimize.ext import c_bioen
    with pytest.raises(RuntimeError):
        c_bioen.bioen_log_posterior_logw(np.zeros(5), np.zeros(5), np.zeros(5), y, np.zeros(3), 1.0)
    with pytest.raises(RuntimeError):
        c_bioen.grad_bioen_log_posterior_forces(np.zeros(3), np.full(5, 0.2), y, np.zeros(3), 1.0)


def test_plain_c_program_compiles_and_links_against_the_header(tmp_path):
    """include/bioen_b200.h is valid C (not only C++) and every symbol the demo uses resolves in the library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    _lib.load()
    exe = tmp_path / "c_abi_demo"
    libdir = os.path.dirname(_lib.library_path())
    res = subprocess.run([gcc, "-O1", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", str(exe), "-L", libdir, "-lbioen_b200",
                          "-lm", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
