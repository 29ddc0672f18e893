"""Builds libqtesla_b200.so (sm_100a) in-tree with nvcc. No JIT cache: the .so travels with the repo."""
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB = os.path.join(PKG_DIR, "libqtesla_b200.so")
SOURCES = [os.path.join(CSRC, "qt_capi.cu"), os.path.join(CSRC, "qt_reference_api.cpp")]
HEADERS = [os.path.join(CSRC, f) for f in
           ("qt_params.h", "qt_tables.h", "qt_tile.cuh", "qt_kernels.cuh", "qt_nussbaumer.cuh")] + [
    os.path.join(os.path.dirname(PKG_DIR), "include", "qtesla_b200.h"),
    os.path.join(os.path.dirname(PKG_DIR), "include", "qtesla_b200_reference_api.h")]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS if os.path.exists(f))


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
