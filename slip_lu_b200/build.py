"""Build recipe of libslip_lu_b200.so: nvcc for the CUDA layer (sm_100a), gcc for the C host layer.

Run as ``python -m slip_lu_b200.build`` or through ``__graft_entry__.build()``.  The library is built
in-tree (slip_lu_b200/libslip_lu_b200.so) so that it travels with the source tree.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INC = os.path.join(ROOT, "include")
OUT = os.path.join(HERE, "libslip_lu_b200.so")
BUILD = os.path.join(HERE, "csrc", "_build")


def _have_system_gmp_header() -> bool:
    r = subprocess.run(["gcc", "-include", "gmp.h", "-include", "mpfr.h", "-E", "-x", "c", "/dev/null"],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return r.returncode == 0


def _gmp_link_flags():
    if os.path.exists("/usr/lib/x86_64-linux-gnu/libgmp.so"):
        return ["-lgmp", "-lmpfr"]
    return ["-l:libgmp.so.10", "-l:libmpfr.so.6"]


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(BUILD, exist_ok=True)
    headers = glob.glob(os.path.join(INC, "*.h")) + glob.glob(os.path.join(HERE, "csrc", "host", "*.h"))
    inc = ["-I" + INC, "-I" + os.path.join(HERE, "csrc", "host")]
    if not _have_system_gmp_header():
        inc.append("-I" + os.path.join(INC, "gmp_abi"))
    objs = []
    cu = os.path.join(HERE, "csrc", "cuda", "slipcu.cu")
    cu_o = os.path.join(BUILD, "slipcu.o")
    if force or _stale(cu_o, [cu] + headers):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-I" + INC, "-c", cu, "-o", cu_o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    objs.append(cu_o)
    for src in sorted(glob.glob(os.path.join(HERE, "csrc", "host", "*.c"))):
        o = os.path.join(BUILD, os.path.basename(src)[:-2] + ".o")
        if force or _stale(o, [src] + headers):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-fopenmp", "-Wall", "-Wno-implicit-fallthrough",
                                   "-std=gnu11"] + inc + ["-c", src, "-o", o])
        objs.append(o)
    if force or _stale(OUT, objs):
        cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
        subprocess.check_call(["g++", "-shared", "-o", OUT] + objs +
                              ["-L" + cuda_lib, "-lcudart_static", "-lrt", "-ldl", "-lpthread", "-fopenmp", "-lm"] +
                              _gmp_link_flags())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
