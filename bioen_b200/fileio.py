"""Input formats of the hot path: the reference's pickle / HDF5 problem files and
matrices that are streamed from disk straight into HBM.

Part 1 mirrors `bioen/fileio.py` (load 18-44, dump 47-68, convert_to_hdf5 71-93, load_pickle / load_hdf5 101-131,
dump_pickle / dump_hdf5 134-175): same function names, the suffix decides the format, HDF5 needs h5py (optional here:
the image this was built in has none, so the HDF5 branches raise ImportError with a clear text instead of failing at
import time like the reference does, fileio.py:10).

Part 2 is what the GPU adds: `load_problem` names the arrays of the reference's optimisation test files
(test/optimize/data/*.pkl, lists in the order of test_fileio_logw.py:37 / test_fileio_forces.py), and
`upload_streamed` copies a matrix that lives on disk (a .npy memory map, an HDF5 dataset, anything sliceable by rows)
into a resident bioen_b200.Problem in row chunks through two pinned staging buffers, the next chunk being read while
the previous one crosses PCIe -- the host never holds more than two chunks, so yTilde may be larger than host RAM.
"""
import os
import pickle
import string
import threading

import numpy as np

LOGW_KEYS = ["GInit", "G", "y", "yTilde", "YTilde", "w0", "theta"]          # test_fileio_logw.py:37
FORCES_KEYS = ["forces_init", "w0", "y", "yTilde", "YTilde", "theta"]       # test_fileio_forces.py


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:
        raise ImportError("bioen_b200.fileio: HDF5 files need the h5py package (not installed); "
                          "pickle (.pkl) and NumPy (.npy) inputs work without it") from e


# ---------------------------------------------------------------------------------------------------------------
# part 1 -- the reference's interface
# ---------------------------------------------------------------------------------------------------------------
def load(filename, hdf5_deep_mode=False, hdf5_keys=[]):
    """Load a pickle or HDF5 file, by suffix (bioen/fileio.py:18-44)."""
    extension = os.path.splitext(filename)[1]
    if extension == ".pkl":
        return load_pickle(filename)
    if extension == ".h5":
        return load_hdf5(filename, hdf5_deep_mode, hdf5_keys)
    raise ValueError("filename extension not recognized (only '.h5' or '.pkl')")


def dump(filename, data, hdf5_keys=[]):
    """Store a list (pickle, HDF5) or a dict with string keys (HDF5), by suffix (bioen/fileio.py:47-68)."""
    extension = os.path.splitext(filename)[1]
    if extension == ".pkl":
        dump_pickle(filename, data)
    elif extension == ".h5":
        dump_hdf5(filename, data, hdf5_keys)
    else:
        raise ValueError("filename extension not recognized (only '.h5' or '.pkl')")


def convert_to_hdf5(filename_pickle, filename_h5, hdf5_keys=[]):
    """bioen/fileio.py:71-93"""
    x = load_pickle(filename_pickle)
    assert isinstance(x, (list, tuple))
    dump_hdf5(filename_h5, x, hdf5_keys)


def load_pickle(file_name):
    """Python-3 pickles as they are; the reference's legacy files were written by Python 2 (NumPy arrays and
    np.matrix objects inside), which need the latin1 decoding."""
    with open(file_name, "rb") as fp:
        try:
            return pickle.load(fp)
        except UnicodeDecodeError:
            fp.seek(0)
            return pickle.load(fp, encoding="latin1")


def dump_pickle(file_name, data):
    with open(file_name, "wb") as fp:
        pickle.dump(data, fp)


def load_hdf5(file_name, hdf5_deep_mode=False, hdf5_keys=[]):
    """bioen/fileio.py:107-131: a dict of everything (deep mode), or the named / all top-level datasets as a list."""
    h5py = _h5py()

    def rec(group):
        out = {}
        for key, value in sorted(group.items()):
            out[key] = value[()] if isinstance(value, h5py.Dataset) else rec(value)
        return out

    with h5py.File(file_name, "r") as f:
        if hdf5_deep_mode:
            return rec(f)
        if hdf5_keys:
            return [f[k][()] for k in hdf5_keys if isinstance(f[k], h5py.Dataset)]
        return [v[()] for _, v in sorted(f.items()) if isinstance(v, h5py.Dataset)]


def _label(i):
    n = len(string.ascii_uppercase)
    return "{}{}".format(string.ascii_uppercase[i // n], string.ascii_uppercase[i % n])


def dump_hdf5(file_name, data, data_labels=[]):
    """bioen/fileio.py:139-175: list elements get the given labels, or sortable artificial ones ("AA", "AB", ...)."""
    h5py = _h5py()

    def rec(group, d):
        for key, value in d.items():
            if isinstance(value, dict):
                rec(group.create_group(key), value)
            else:
                group.create_dataset(key, data=value)

    with h5py.File(file_name, "w") as f:
        if isinstance(data, (list, tuple)):
            labels = list(data_labels) if len(data_labels) == len(data) else [_label(i) for i in range(len(data))]
            for lab, value in zip(labels, data):
                f.create_dataset(lab, data=value)
        elif isinstance(data, dict):
            rec(f, data)
        else:
            raise TypeError("data type unsupported")


# ---------------------------------------------------------------------------------------------------------------
# part 2 -- problems and streamed matrices
# ---------------------------------------------------------------------------------------------------------------
def load_problem(filename, kind=None):
    """A reference optimisation problem file as a dict of plain float64 ndarrays (np.matrix flattened to ndarray,
    shapes as the reference passes them: GInit, G, w0, forces_init (n|m, 1); y, yTilde (m, n); YTilde (1, m)).
    kind: 'logw' | 'forces' | None (guessed from the number of entries: 7 vs 6)."""
    data = load(filename, hdf5_keys=[]) if not filename.endswith(".h5") else None
    if data is None:
        h5py = _h5py()
        with h5py.File(filename, "r") as f:
            names = set(f.keys())
        keys = LOGW_KEYS if (kind == "logw" or (kind is None and "GInit" in names)) else FORCES_KEYS
        data = load_hdf5(filename, hdf5_keys=keys)
    else:
        if kind is None:
            kind = "logw" if len(data) == len(LOGW_KEYS) else "forces"
        keys = LOGW_KEYS if kind == "logw" else FORCES_KEYS
    if len(data) != len(keys):
        raise ValueError("%s holds %d entries, expected %d (%s)" % (filename, len(data), len(keys), ", ".join(keys)))
    out = {}
    for k, v in zip(keys, data):
        out[k] = float(np.asarray(v).ravel()[0]) if k == "theta" else np.array(v, dtype=np.float64)
    out["kind"] = "logw" if keys is LOGW_KEYS else "forces"
    return out


def open_matrix(source, dataset=None):
    """A row-sliceable (m, n) float64 view of a matrix on disk: '.npy' -> memory map; '.h5' -> dataset `dataset`
    (default 'yTilde'); arrays and anything with .shape and row slicing are passed through."""
    if isinstance(source, str):
        ext = os.path.splitext(source)[1]
        if ext == ".npy":
            return np.load(source, mmap_mode="r")
        if ext == ".h5":
            return _h5py().File(source, "r")[dataset or "yTilde"]
        raise ValueError("open_matrix: only '.npy' and '.h5' paths")
    return source


def upload_streamed(source, device=0, chunk_bytes=256 << 20, problem=None, dataset=None):
    """yTilde from disk into HBM in row chunks.  Returns the resident bioen_b200.Problem.

    Two page-locked staging buffers of `chunk_bytes`; a reader thread fills one (disk -> pinned host memory, with the
    conversion to C-contiguous float64) while the other is copied to the device (bioen_b200_upload_rows), so the disk
    read and the PCIe transfer overlap and the host holds at most two chunks of the matrix at any time."""
    from .problem import Problem, pinned_empty
    src = open_matrix(source, dataset)
    m, n = int(src.shape[0]), int(src.shape[1])
    own = problem is None
    if own:
        problem = Problem(shape=(m, n), device=device)
    elif (problem.m, problem.n) != (m, n):
        raise ValueError("problem shape %s does not match the matrix %s" % ((problem.m, problem.n), (m, n)))
    rows = max(1, min(m, int(chunk_bytes) // (8 * n)))
    bufs = [pinned_empty((rows, n)), pinned_empty((rows, n))]
    chunks = [(r0, min(rows, m - r0)) for r0 in range(0, m, rows)]
    filled = [threading.Event(), threading.Event()]
    free = [threading.Event(), threading.Event()]
    for e in free:
        e.set()
    err = []

    def reader():
        try:
            for k, (r0, nr) in enumerate(chunks):
                b = k & 1
                free[b].wait()
                free[b].clear()
                bufs[b][:nr] = src[r0:r0 + nr]          # disk -> pinned memory (dtype / layout conversion included)
                filled[b].set()
        except Exception as e:                           # surfaced by the uploading thread
            err.append(e)
            for ev in filled:
                ev.set()

    t = threading.Thread(target=reader, daemon=True)
    t.start()
    try:
        for k, (r0, nr) in enumerate(chunks):
            b = k & 1
            filled[b].wait()
            filled[b].clear()
            if err:
                raise err[0]
            problem.upload_rows(r0, bufs[b][:nr])
            free[b].set()
    except Exception:
        if own:
            problem.close()
        raise
    finally:
        for e in free:
            e.set()
        t.join()
    return problem


def problem_from_file(filename, device=0, **kw):
    """Resident Problem for the yTilde of a reference problem file (.pkl / .h5) or of a bare matrix file (.npy, or
    .h5 with dataset=...); returns (problem, data) where data is load_problem()'s dict (None for bare matrices)."""
    from .problem import Problem
    ext = os.path.splitext(filename)[1]
    if ext == ".npy" or (ext == ".h5" and kw.get("dataset")):
        return upload_streamed(filename, device=device, **kw), None
    data = load_problem(filename)
    return Problem(data["yTilde"], device=device), data
