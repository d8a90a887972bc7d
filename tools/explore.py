"""Dev tool: time SLIP_solve_mpq / SLIP_LU_factorize at growing n with the per-kernel counters."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
import slip_lu_b200  # noqa: E402
from slip_lu_b200 import capi, synth  # noqa: E402


class Counters(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("trisolve_launches", C.c_uint64), ("trisolve_ms", C.c_double),
                ("trisolve_bytes", C.c_double), ("trisolve_modmul", C.c_double), ("recon_ms", C.c_double),
                ("recon_mac", C.c_double), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
                ("device_ms", C.c_double), ("other_ms", C.c_double), ("trisolve_union_ms", C.c_double)]


def main():
    lib = slip_lu_b200.lib()
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [200, 400, 800]
    prof = "--prof" in sys.argv
    host_factors = "--factors" in sys.argv
    rep = 2 if "--repeat" in sys.argv else 1
    for n in [m for m in sizes for _ in range(rep)]:
        n, cp, ri, vals, b = synth.random_sparse(n, 10, 32, seed=1)
        o = lib.default_options()
        A = lib.sparse_from_csc(n, cp, ri, vals)
        B = lib.dense_from_rows(b)
        t = time.time(); S = lib.analyze(A, o); ta = time.time() - t
        lib.dll.slipcu_reset_counters()
        lib.dll.slipcu_set_profiling(1 if prof else 0)
        t = time.time()
        if host_factors:
            L, U, rhos, pinv = lib.factorize(A, S, o)
            x = lib.lu_solve(B, rhos, L, U, pinv)
            nnz = (L.contents.nz, U.contents.nz)
        else:
            x = lib.solve_mpq(A, S, B, o)
            nnz = None
        dt = time.time() - t
        c = Counters(); lib.dll.slipcu_get_counters(C.byref(c))
        ok = lib.dll.SLIP_check_solution(A, x, B) if not host_factors else "n/a"
        print(f"n={n} analyze {ta:.3f}s factor+solve {dt:.3f}s check={ok} nnz={nnz} launches={c.launches} "
              f"tri_ms={c.trisolve_ms:.1f} tri_GB={c.trisolve_bytes/1e9:.2f} "
              f"tri_GBps={(c.trisolve_bytes/1e9)/(c.trisolve_ms/1e3) if c.trisolve_ms else 0:.0f} "
              f"recon_ms={c.recon_ms:.1f} recon_Gmac={c.recon_mac/1e9:.1f} other_ms={c.other_ms:.1f}", flush=True)
        if host_factors:
            lib.free_sparse(L); lib.free_sparse(U)


if __name__ == "__main__":
    main()
