"""Build libldpc_b200.so in-tree with nvcc for sm_100a.

    python -m ldpc_decoders_b200.build [--force]

The shared library is a plain C-ABI object (include/ldpc_b200.h): it links only the CUDA
runtime (statically) and is loaded with ctypes — no torch extension machinery involved.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libldpc_b200.so")
SOURCES = ["ldpc_b200.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "ldpc_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def nvcc_path():
    for cand in (os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc"), "nvcc"):
        try:
            subprocess.run([cand, "--version"], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            return cand
        except (OSError, subprocess.CalledProcessError):
            continue
    raise RuntimeError("nvcc not found (needed to build libldpc_b200.so for sm_100a)")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources; returns its path."""
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
