"""Build libmvsdet_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m mvsdet_b200.build [--force] [--verbose]

The library has no torch dependency: plain nvcc, one object per .cu compiled in
parallel, linked into ``mvsdet_b200/lib/libmvsdet_b200.so``.  nvcc
cross-compiles without a GPU, so this runs in the CPU-only build container; the
.so travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmvsdet_b200.so")
OBJDIR = os.path.join(HERE, "_obj")
SOURCES = ("capi.cu", "pack.cu", "scene_setup.cu", "plane_sweep_fwd.cu", "plane_sweep_bwd.cu",
           "plane_sweep_bwd_run.cu", "group_corr.cu", "depth_topk.cu", "backproject.cu", "voxel_p2p.cu")
HEADERS = (os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "plane_sweep.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "mvsdet_b200.h"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-Xfatbin", "-compress-all"]
# experiment builds only (e.g. MVSD_EXTRA_NVCC_FLAGS="-DMVSD_KRUN=16"); part of the object digest
NVCC_FLAGS += os.environ.get("MVSD_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; the mvsdet_b200 library cannot be built")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _sources():
    paths = [os.path.join(CSRC, s) for s in SOURCES]
    missing = [p for p in paths if not os.path.isfile(p)]
    if missing:
        raise RuntimeError(f"mvsdet_b200 build: listed sources are missing: {missing}")
    return paths


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([src, *HEADERS])
        if (not force and os.path.isfile(obj) and os.path.isfile(stamp)
                and open(stamp).read() == dig):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as fh:
            fh.write(dig)
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    logs = "".join(log for _, log in results)
    if logs:
        with open(os.path.join(OBJDIR, "ptxas.log"), "w") as fh:
            fh.write(logs)
        if verbose:
            print(logs)
    newest_obj = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < newest_obj:
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
