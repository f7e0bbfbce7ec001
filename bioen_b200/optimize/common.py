"""Shared helpers of the optimize API (mirror of bioen/optimize/common.py)."""
import numpy as np

from .. import _lib
from ..problem import Problem


def _is_matrix(a):
    return isinstance(a, np.matrix)


def chiSqrTerm(w, yTilde, YTilde):
    """0.5 * || yTilde.w - YTilde ||^2 (legacy host helper, bioen/optimize/common.py:5-39)."""
    v = np.asarray(yTilde, dtype=np.float64) @ _lib.vec(w) - _lib.vec(YTilde)
    return 0.5 * float(v @ v)


def getAve(w, y):
    """Ensemble averages y.w as a flat (m,) array (legacy host helper, bioen/optimize/common.py:42-60)."""
    return np.asarray(y, dtype=np.float64) @ _lib.vec(w)


def device_average(w, y, problem=None):
    """y.w computed on the GPU; `problem` is reused when it already holds `y`."""
    if problem is not None:
        return problem.average(w)
    with Problem(y) as p:
        return p.average(w)


def average_like(problem, w, y, keep=False):
    """y.w on the device(s) `problem` lives on (a Problem, or a dist.ShardedProblem over all GPUs of the job), for a
    host matrix y that is not the resident yTilde (the reference's post-processing, log_weights.py:612-613).

    keep=False (a one-off find_optimum): y is STREAMED through a small device buffer in row chunks
    (Problem.average_streamed) -- no second resident M x N matrix.  keep=True (the caller owns `problem` and will
    call again, e.g. a theta series): a resident problem for y is made once and cached on `problem` until it is
    closed or a different y arrives."""
    if keep:
        cached = getattr(problem, "_y_cache", None)
        if cached is None or cached[0] is not y:
            if cached is not None:
                cached[1].close()
            problem._y_cache = (y, problem.like(y))
        return problem._y_cache[1].average(w)
    if hasattr(problem, "average_streamed"):
        return problem.average_streamed(y, w)
    with problem.like(y) as q:
        return q.average(w)


def print_highlighted(text, verbose=True):
    """bioen/optimize/common.py:63-80"""
    if verbose:
        n = len(text)
        print("-" * n)
        print(text)
        print("-" * n)


def set_caching_heuristics(m, n):
    """bioen/optimize/common.py:83-106: True iff m*n*8 bytes <= 8 GiB.  The value is carried in the cfg dict for
    compatibility; the GPU kernels never need the transposed copy."""
    return not (m * n * 8 > 8 * 2 ** 30)
