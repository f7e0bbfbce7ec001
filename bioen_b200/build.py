"""Build libbioen_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

The library is compiled for sm_100a ONLY (-gencode arch=compute_100a,code=sm_100a); there is no other code
path.  nvcc cross-compiles without a GPU, so this also runs on CPU-only build boxes.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "bioen_b200.cu")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libbioen_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-diag-suppress", "550",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("bioen_b200: nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    d = os.path.join(HERE, "csrc")
    out = [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "bioen_b200.h"))
    return out


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources() if os.path.exists(s))


def build_library(force=False, verbose=False):
    """Compile bioen_b200/lib/libbioen_b200.so if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("bioen_b200: nvcc failed\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
