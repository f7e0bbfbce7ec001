"""One process, one generated matrix: the tile orders of k_read_tiles back to back (see read_order_probe.py)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bioen_b200  # noqa: E402
from bioen_b200 import _lib  # noqa: E402

M, N = 1000, 1000000
p = bioen_b200.Problem(shape=(M, N))
_lib.check(_lib.load().bioen_b200_alloc_ytilde(p._h), "alloc")
p.generate(12345, 0, np.zeros(M), 2.0)
g = ctypes.c_double()
os.environ["BIOEN_B200_READ_VARIANT"] = "4,2"
for rep in range(2):
    for order in ("-1", "2", "4", "0"):
        os.environ["BIOEN_B200_READ_ORDER"] = order
        _lib.check(_lib.load().bioen_b200_read_stream_peak(p._h, 10, ctypes.byref(g)), "read")
        print("order %2s  %.0f GB/s" % (order, g.value), flush=True)
