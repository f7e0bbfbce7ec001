"""Nuisance-parameter refits on the resident matrix.

Between two weight optimisations the reference refits the DEER modulation depths and the scattering scale factor
against the current weights and rebuilds the simulated-data matrix on the host, entry by entry in a Python double
loop (bioen/analyze/observables/observables.py: get_proc_sim 110-143, moddepth_fit 146-171, coeff_fit 174-188,
update_sim 191-216, update_sim_init 219-229; driven by bioen/analyze/procedure.py:62-83).  Both parameters act
AFFINELY on the rows of yTilde:

    DEER label (rows of one spin-label pair)   yTilde_ij = (1 - m + m s_ij) / err_i  =  (m) * b_ij + (1 - m) / err_i
    scattering                                 yTilde_ij = c s_ij / err_i            =  (c) * b_ij
    generic / CD                               yTilde_ij = s_ij / err_i              =      b_ij          (no parameter)

with the parameter-free base b_ij = s_ij / err_i.  So with yTilde resident in HBM

  * the chi^2 of ANY trial value of a parameter needs the base averages  (b . w)_i  only: ONE row pass on the device
    (Problem.average) serves the whole least-squares fit of all blocks, every trial after that is O(rows) on the host;
  * committing the refitted values is the in-place row-affine transform Problem.affine_rows (one read + one write
    pass on the device) instead of a rebuild on the host and a new upload.

The fit itself is the reference's: scipy.optimize.leastsq on the scalar chi^2 of the block (observables.py:207,
212), same start values, same "initial-optimization" defaults (0.15 / 0.0002).
"""
import numpy as np
from scipy.optimize import leastsq

from . import _lib

DEER, SCATTERING, FIXED = "deer", "scattering", "fixed"
INITIAL = "initial-optimization"
INITIAL_VALUES = {DEER: 0.15, SCATTERING: 0.0002}          # observables.py:224, 228


class Block:
    """Rows [start, stop) of yTilde that share one nuisance parameter.

    kind      'deer' (modulation depth m), 'scattering' (scale factor c) or 'fixed'
    err       (rows,) experimental errors
    exp_fit   (rows,) experimental values the PARAMETER is fitted against (DEER: the background-corrected raw trace,
              exp_tmp[:, 1], observables.py:161; scattering: I(q), :179) -- not necessarily the values the weights are
              optimised against (DEER: the polynomial fit, exp_tmp[:, 2], :340)
    value     current parameter (float), or 'initial-optimization'
    """

    def __init__(self, kind, start, stop, err=None, exp_fit=None, value=None, name=None):
        if kind not in (DEER, SCATTERING, FIXED):
            raise ValueError("unknown block kind %r" % (kind,))
        self.kind, self.start, self.stop = kind, int(start), int(stop)
        n = self.stop - self.start
        self.err = np.ones(n) if err is None else _lib.vec(err)
        self.exp_fit = None if exp_fit is None else _lib.vec(exp_fit)
        if kind != FIXED and (self.exp_fit is None or self.exp_fit.size != n or self.err.size != n):
            raise ValueError("err and exp_fit must have one entry per row of the block")
        self.value = value
        self.name = name or "%s[%d:%d]" % (kind, self.start, self.stop)

    # row-affine map base -> yTilde for a parameter value:  yTilde = scale * base + offset
    def scale_offset(self, value):
        n = self.stop - self.start
        if self.kind == DEER:
            return np.full(n, float(value)), (1.0 - float(value)) / self.err
        if self.kind == SCATTERING:
            return np.full(n, float(value)), np.zeros(n)
        return np.ones(n), np.zeros(n)

    def chi2(self, value, base_avg, sum_w):
        """0.5 * || yTilde(value) . w - exp_fit / err ||^2 over the rows of the block -- what moddepth_fit / coeff_fit
        return (observables.py:171, 188), from the base averages (b . w)_i."""
        s, o = self.scale_offset(np.ravel(value)[0])
        r = s * base_avg + o * sum_w - self.exp_fit / self.err
        return 0.5 * float(r @ r)


def base_matrix(raw_rows, err):
    """b = s / err[:, None] for raw simulated data s (rows x models): the parameter-free matrix to keep resident."""
    return np.asarray(raw_rows, dtype=np.float64) / _lib.vec(err)[:, None]


def proc_sim(base, blocks):
    """yTilde for the blocks' current values from the base matrix on the HOST (NumPy twin of get_proc_sim,
    observables.py:110-143; used to build the first upload and by the tests)."""
    out = np.array(base, dtype=np.float64, copy=True)
    for b in blocks:
        if b.kind == FIXED:
            continue
        s, o = b.scale_offset(_start_value(b))
        out[b.start:b.stop] = s[:, None] * out[b.start:b.stop] + o[:, None]
    return out


def _start_value(block):
    if isinstance(block.value, str):
        if block.value != INITIAL:
            raise ValueError("block value must be a number or %r" % INITIAL)
        return INITIAL_VALUES[block.kind]
    return float(block.value)


class NuisanceRefit:
    """Refit loop state for one resident problem.

    `problem` (bioen_b200.Problem or dist.ShardedProblem) must hold yTilde built with the blocks' CURRENT values
    (e.g. uploaded from proc_sim(base, blocks)); rows outside all blocks are left alone.  `update(w)` mirrors
    Observables.update_sim (observables.py:191-216): refit every block against the weights w, then commit.
    """

    def __init__(self, problem, blocks):
        self.problem = problem
        self.blocks = list(blocks)
        self.m = problem.m
        for b in self.blocks:
            if not (0 <= b.start <= b.stop <= self.m):
                raise ValueError("block %s outside the matrix" % b.name)
            if b.kind != FIXED:
                b.value = _start_value(b)
        self._scale = np.ones(self.m)       # resident matrix = _scale * base + _offset (row-wise)
        self._offset = np.zeros(self.m)
        for b in self.blocks:
            self._scale[b.start:b.stop], self._offset[b.start:b.stop] = b.scale_offset(b.value)

    def base_averages(self, w):
        """(b . w)_i and sum(w): one row pass over the resident matrix, mapped back through the current transform."""
        w = _lib.vec(w)
        avg = self.problem.average(w)
        sw = float(w.sum())
        return (avg - self._offset * sw) / self._scale, sw

    def chi2(self, block, value, w=None, base_avg=None, sum_w=None):
        """chi^2 of one block for a trial value (= Observables.moddepth_fit / coeff_fit)."""
        if base_avg is None:
            base_avg, sum_w = self.base_averages(w)
        return block.chi2(value, base_avg[block.start:block.stop], sum_w)

    def fit(self, w):
        """Least-squares refit of every block against w (observables.py:203-213).  Returns {name: value}; nothing
        is committed."""
        base_avg, sw = self.base_averages(w)
        out = {}
        for b in self.blocks:
            if b.kind == FIXED:
                continue
            sl = base_avg[b.start:b.stop]
            opt, _ = leastsq(lambda v: b.chi2(v, sl, sw), b.value)
            out[b.name] = float(np.ravel(opt)[0])
        return out

    def commit(self, values):
        """Make the resident matrix the one of the given values: one in-place row-affine pass on the device."""
        new_s, new_o = self._scale.copy(), self._offset.copy()
        for b in self.blocks:
            if b.kind == FIXED or b.name not in values:
                continue
            v = float(values[b.name])
            if v == 0.0:
                raise ValueError("a zero %s parameter makes the block's rows independent of the simulated data; the "
                                 "resident matrix cannot be transformed back from that" % b.kind)
            new_s[b.start:b.stop], new_o[b.start:b.stop] = b.scale_offset(v)
            b.value = v
        a = new_s / self._scale                      # new = a * old + c
        c = new_o - self._offset * a
        if np.any(a != 1.0) or np.any(c != 0.0):
            self.problem.affine_rows(a, c)
        self._scale, self._offset = new_s, new_o

    def update(self, w):
        """Observables.update_sim: refit, then rebuild (here: transform in place).  Returns {name: value}."""
        values = self.fit(w)
        self.commit(values)
        return values
