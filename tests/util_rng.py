"""NumPy restatement of the on-device synthetic-data generator (bioen_b200/csrc/bioen_b200.cu, k_generate):
counter-based, so any block of the (possibly sharded, possibly 400 GB) matrix can be reproduced on the host."""
import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def generic_ytilde_block(seed, ytrue_over_sigma, inv_sigma, row0, nrows, col0, ncols):
    """yTilde[i, j] = a_i + inv_sigma * z(seed, i, j) for global rows/columns of the block."""
    with np.errstate(over="ignore"):
        i = np.arange(row0, row0 + nrows, dtype=np.uint64)[:, None]
        j = np.arange(col0, col0 + ncols, dtype=np.uint64)[None, :]
        ctr = (i << np.uint64(40)) + j
        h1 = _mix64(np.uint64(seed) + _G * (ctr + np.uint64(1)))
        h2 = _mix64(h1 + _G)
    u1 = ((h1 >> np.uint64(11)).astype(np.float64) + 1.0) * 2.0 ** -53
    u2 = (h2 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    a = np.asarray(ytrue_over_sigma, dtype=np.float64)[row0:row0 + nrows, None]
    return a + inv_sigma * z
