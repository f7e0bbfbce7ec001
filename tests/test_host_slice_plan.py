"""CPU: launch geometry of the shared-memory slice kernel (csrc/slice_eval.cuh, slice_plan) through the host-only
hook bioen_b200_selftest_slice_plan.  The kernel relies on these invariants: the CTAs' column ranges tile [0, N) and
every CTA owns at least one column; column counts and strides are even (16-byte cp.async rows) and the stride is
= 2 mod 4 (bank conflicts); the slice and its vectors fit the shared memory asked for; the grid fits the device and
the tables; the thread layouts cover the slice.  No GPU involved."""
import ctypes as C

import pytest

from bioen_b200 import _lib

SMS, SMEM = 148, 232448 - 7168 - 256
THREADS = 512


def plan(m, n, sms=SMS, smem=SMEM):
    out = (C.c_longlong * 8)()
    ok = _lib.load().bioen_b200_selftest_slice_plan(m, n, sms, smem, out)
    return ok, dict(zip(("nc", "ncs", "ms", "grid", "cx_log2", "l_log2", "lt_log2", "smem"), list(out)))


SHAPES = [(1, 1), (7, 3), (17, 8), (17, 9), (808, 10), (808, 100), (205, 10), (1000, 15), (3, 296), (5, 2368),
          (7, 2369), (64, 300), (100, 1000), (1000, 777), (28, 50001), (40, 76000), (9, 150001), (2, 400000),
          (64, 4096), (5, 20000), (257, 3001), (500, 2049), (2000, 8), (16, 15), (64, 64), (1, 500000)]


@pytest.mark.parametrize("m,n", SHAPES)
def test_eligible_shapes(m, n):
    ok, p = plan(m, n)
    assert ok == 1, (m, n, p)
    nc, ncs, grid = p["nc"], p["ncs"], p["grid"]
    assert nc >= 8 and nc % 2 == 0 and ncs % 2 == 0 and ncs % 4 == 2 and ncs >= nc
    assert 1 <= grid <= SMS and grid <= 160
    assert (grid - 1) * nc < n <= grid * nc                  # the ranges tile [0, N); the last CTA owns >= 1 column
    mp = (m + 1) // 2 * 2
    assert p["smem"] == (m * ncs + 5 * nc + 3 * mp) * 8 <= SMEM
    assert p["ms"] >= 8 + m and p["ms"] % 16 == 0 and grid * p["ms"] * 8 <= 1 << 20
    cx, L, LT = 1 << p["cx_log2"], 1 << p["l_log2"], 1 << p["lt_log2"]
    assert cx <= THREADS and (cx >= nc or cx == THREADS)     # one column per thread unless the slice is wider than the CTA
    assert 1 <= L <= 32 and L <= max(1, cx) and 1 <= LT <= 32
    assert L == 1 or L * m <= THREADS or L * m < 2 * THREADS  # about one sweep of the CTA covers all rows
    assert LT == 1 or LT <= max(1, 2 * grid)


@pytest.mark.parametrize("m,n", [(300, 40000), (64, 100003), (1000, 100000), (1000, 1000000), (5000, 800), (2800, 8), (1, 1000000)])
def test_ineligible_shapes(m, n):
    ok, p = plan(m, n)
    assert ok == 0


def test_fewer_sms_and_less_shared_memory():
    ok, p = plan(28, 50001, sms=132)
    assert ok == 1 and p["grid"] <= 132 and p["grid"] * p["nc"] >= 50001
    ok, p = plan(28, 50001, smem=48 * 1024)                 # 28 x 338 doubles do not fit 48 KB
    assert ok == 0
    ok, p = plan(28, 5001, smem=48 * 1024)
    assert ok == 1 and p["smem"] <= 48 * 1024
