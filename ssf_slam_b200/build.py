"""Builds the C-ABI shared library ``ssf_slam_b200/libssf_b200.so`` from ``csrc/*.cu`` with nvcc for
sm_100a (cross-compiles without a GPU).  The library is kept in-tree so it travels with the repo.
Each translation unit is compiled to an object under ``build/`` (in parallel, only when stale) and linked."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libssf_b200.so")
DEV_LIB = os.path.join(HERE, "libssf_b200_dev.so")   # developer entry points (csrc/dev/), not part of the product ABI
OBJ_DIR = os.path.join(HERE, "build")
NVCC_FLAGS = (["-DSSF_CV_TRACE"] if os.environ.get("SSF_CV_TRACE") else []) + ([] if os.environ.get("SSF_SPLIT_RNA") else ["-DSSF_SPLIT_TRUNC"]) + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", os.path.join(os.path.dirname(HERE), "include")]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def dev_sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "dev", "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "csrc", "dev", "*.h")) + glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h")))


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in deps)


def needs_build():
    return _stale(LIB, sources() + headers()) or _stale(DEV_LIB, dev_sources() + headers())


def _nvcc():
    return os.environ.get("NVCC") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = headers()
    todo = [s for s in sources() + dev_sources() if force or _stale(_obj(s), [s] + hdrs)]
    extra = ["-Xptxas", "-v"] if verbose else []
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        list(ex.map(lambda s: _run([_nvcc()] + NVCC_FLAGS + extra + ["-c", s, "-o", _obj(s)], verbose), todo))
    _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [_obj(s) for s in sources()], verbose)
    capi = [_obj(s) for s in sources() if os.path.basename(s) == "capi.cu"]   # error / launch-count plumbing (own copy)
    _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", DEV_LIB] + [_obj(s) for s in dev_sources()] + capi, verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
