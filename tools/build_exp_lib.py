#!/usr/bin/env python
"""Build an EXPERIMENT copy of the library: one source recompiled with extra flags, the other
objects reused; written to mvsdet_b200/lib/exp_<tag>.so (select it with MVSDET_B200_LIB=...).

    python tools/build_exp_lib.py g1mb4 plane_sweep_bwd_run.cu -DMVSD_EXP_BWD_G1 -DMVSD_EXP_BWD_MINB=4
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvsdet_b200 import build as B  # noqa: E402


def main():
    tag, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
    B.build()
    obj = f"/tmp/exp_{tag}.o"
    r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj],
                       capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    objs = [os.path.join(B.OBJDIR, s[:-3] + ".o") for s in B.SOURCES if s != src] + [obj]
    out = os.path.join(B.LIBDIR, f"exp_{tag}.so")
    subprocess.run([B._nvcc(), "-shared", "-o", out, *objs, "-lcudart"], check=True)
    print(out)


if __name__ == "__main__":
    main()
