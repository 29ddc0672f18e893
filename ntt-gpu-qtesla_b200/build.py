"""Builds libqtesla_b200.so (sm_100a) in-tree with nvcc.  No JIT cache: the built .so sits next to this file
(git-ignored, so the history stays source-only; it is NOT gpurun-ignored, so it travels to the GPU box)."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB = os.path.join(PKG_DIR, "libqtesla_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
# one translation unit holds every kernel (the __constant__ twiddle bank is defined once); the host-only
# units compile beside it
SOURCES = [os.path.join(CSRC, f) for f in ("qt_capi.cu", "qt_host.cu", "qt_reference_api.cpp")]
HEADERS = [os.path.join(CSRC, f) for f in
           ("qt_params.h", "qt_tables.h", "qt_tile.cuh", "qt_kernels.cuh", "qt_nussbaumer.cuh")] + [
    os.path.join(INCLUDE, "qtesla_b200.h"), os.path.join(INCLUDE, "qtesla_b200_reference_api.h")]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-diag-suppress=128", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC"]


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")


def _newer_than(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in deps if os.path.exists(f))


def stale():
    return _newer_than(LIB, SOURCES + HEADERS)


def build(force=False, verbose=False):
    """QT_NVCC_EXTRA="-DQT_...=..." QT_BUILD_TAG=name builds an A/B variant into build_ab/name/libqtesla_b200.so
    (select it at run time with QT_LIB_PATH); without a tag the product library next to this file is (re)built."""
    global OBJ_DIR, LIB
    tag = os.environ.get("QT_BUILD_TAG")
    extra = os.environ.get("QT_NVCC_EXTRA", "").split()  # A/B builds: -DQT_...=...
    if tag:
        OBJ_DIR = os.path.join(os.path.dirname(PKG_DIR), "build_ab", tag)
        LIB = os.path.join(OBJ_DIR, "libqtesla_b200.so")
        force = True
    elif extra:
        raise SystemExit("QT_NVCC_EXTRA without QT_BUILD_TAG would overwrite the product library with an A/B variant")
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = _obj(src)
        if force or extra or _newer_than(obj, [src] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.run([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-o", LIB] + objs, check=True)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
