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
# SuiteSparse COLAMD/AMD (third-party prerequisite of SLIP_LU_analyze, like GMP): a system
# libcolamd/libamd is used when present; this image has none, so the packages are compiled from a
# SuiteSparse source tree (the copy the reference distribution vendors) into _deps/.
SUITESPARSE_SRC = os.environ.get("SLIP_B200_SUITESPARSE_SRC", "/root/reference")
DEPS = os.path.join(HERE, "_deps")
ORDERING_SO = os.path.join(DEPS, "libsuitesparse_ordering.so")


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


def build_ordering(force: bool = False) -> str | None:
    """libsuitesparse_ordering.so = SuiteSparse COLAMD + AMD + SuiteSparse_config, unmodified, compiled
    from SUITESPARSE_SRC where it lies (nothing is copied into the repository)."""
    col = os.path.join(SUITESPARSE_SRC, "COLAMD", "Source", "colamd.c")
    if not os.path.exists(col):
        return ORDERING_SO if os.path.exists(ORDERING_SO) else None
    src = [col, os.path.join(SUITESPARSE_SRC, "SuiteSparse_config", "SuiteSparse_config.c")]
    src += sorted(glob.glob(os.path.join(SUITESPARSE_SRC, "AMD", "Source", "*.c")))
    if force or _stale(ORDERING_SO, src):
        os.makedirs(DEPS, exist_ok=True)
        inc = ["-I" + os.path.join(SUITESPARSE_SRC, d) for d in ("SuiteSparse_config", "COLAMD/Include", "AMD/Include")]
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-w", "-shared"] + inc + ["-o", ORDERING_SO] + src + ["-lm"])
    return ORDERING_SO


def build(force: bool = False, verbose: bool = False) -> str:
    build_ordering(force)
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
