"""Dev tool (GPU box): run the product on packed reference systems (tests/golden/mats/refmats.npz),
compare with the reference's recorded digests, print times and the per-kernel counters.

    python tools/refmats_gpu.py [--factors] [--repeat N] NAME ...      (NAME: NSR8K, prob159, basislib/gen2, all)
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
import slip_lu_b200  # noqa: E402
from slip_lu_b200 import refmats  # noqa: E402
from oracle import binding as ob  # noqa: E402


class Counters(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("trisolve_launches", C.c_uint64), ("trisolve_ms", C.c_double),
                ("trisolve_bytes", C.c_double), ("trisolve_modmul", C.c_double), ("recon_ms", C.c_double),
                ("recon_mac", C.c_double), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
                ("device_ms", C.c_double), ("other_ms", C.c_double), ("trisolve_union_ms", C.c_double)]


def main():
    lib = slip_lu_b200.lib()
    lib.dll.SLIP_B200_last_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
    lib.dll.SLIP_B200_last_pinv.argtypes = [C.POINTER(C.c_int32), C.c_int]
    args = sys.argv[1:]
    factors = "--factors" in args
    rep = int(args[args.index("--repeat") + 1]) if "--repeat" in args else 1
    names = [a for a in args if not a.startswith("--") and not a.isdigit()]
    recs = refmats.records()
    if names == ["all"]:
        names = list(recs)
    for name in names:
        n, I, J, X, b = refmats.system(name)
        rec = recs[name]
        A = lib.sparse_from_triplets(n, I, J, X)
        B = lib.dense_from_rows(b)
        o = lib.default_options()
        nrhs = len(b[0])
        for it in range(rep):
            S = lib.analyze(A, o)
            lib.dll.slipcu_reset_counters()
            t = time.perf_counter()
            if factors:
                L, U, rhos, pinv = lib.factorize(A, S, o)
                tf = time.perf_counter() - t
                x = lib.lu_solve(B, rhos, L, U, pinv)
                dt = time.perf_counter() - t
                d = rec["digests"]
                ok = dict(L=str(ob.digest_slip_sparse(L)) == d["L"], U=str(ob.digest_slip_sparse(U)) == d["U"],
                          rhos=str(ob.digest_mpz_array(rhos, n)) == d["rhos"],
                          pinv=refmats.digest_ints(list(pinv)) == d["pinv"],
                          x=str(refmats.digest_mpq_mat(lib, x, n, nrhs)) == d["x_lu_solve"])
                extra = f"factorize {tf:.3f}s "
                lib.free_sparse(L); lib.free_sparse(U); lib.free_mpz_array(rhos, n)
            else:
                x = lib.solve_mpq(A, S, B, o)
                dt = time.perf_counter() - t
                pv = (C.c_int32 * n)()
                lib.dll.SLIP_B200_last_pinv(pv, n)
                ok = dict(x=str(refmats.digest_mpq_mat(lib, x, n, nrhs)) == rec["digests"]["x_solve_mpq"],
                          pinv=refmats.digest_ints(list(pv)) == rec["digests"]["pinv"])
                extra = ""
            c = Counters(); lib.dll.slipcu_get_counters(C.byref(c))
            st = (C.c_double * 13)(); lib.dll.SLIP_B200_last_stats(st, 13)
            ref_s = rec["ref_seconds"]["factorize"] + rec["ref_seconds"]["lu_solve"] if factors else rec["ref_seconds"]["solve_mpq"]
            print(f"{name} n={n} {'factorize+lu_solve' if factors else 'solve_mpq'} {extra}total {dt:.3f}s "
                  f"(reference, build container: {ref_s:.3f}s -> {ref_s / dt:.1f}x) parity={ok} "
                  f"channels={int(st[3])}/{int(st[10])} restarts={int(st[11])} verified={int(st[12])} "
                  f"launches={c.launches} tri_GB={c.trisolve_bytes / 1e9:.2f} dev_ms={c.device_ms:.1f}", flush=True)
            lib.free_mpq_mat(x, n, nrhs); lib.free_analysis(S)
        lib.free_dense(B); lib.free_sparse(A)


if __name__ == "__main__":
    main()
