"""Packs the reference distribution's own test systems (ExampleMats/NSR8K, prob159 and the BasisLIB
LP bases with 300 <= n < 700) into tests/golden/mats/refmats.npz and records what the UNMODIFIED
reference (oracle/_ref/libslip_ref.so) computes for them in tests/golden/refmats.json.

Run in the build container (needs /root/reference):   python tests/golden/make_refmats.py
These are the matrices the reference's demos read (Demo/SLIPLU.c, Demo/example2.c); bench.py's
head-to-head legs and the GPU parity tests load them from the packed file, because /root/reference
does not exist on the GPU box.  Only digests of the factors are stored (NSR8K has 3.2 M entries).
"""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slip_lu_b200 import capi, refmats, synth  # noqa: E402
from oracle import binding as ob               # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REFMATS = "/root/reference/SLIP_LU/ExampleMats"
LIMIT_S = 120


def reference_record(ref, name, n, I, J, X, b, family):
    o = ref.default_options()
    A = ref.sparse_from_triplets(n, I, J, X)
    B = ref.dense_from_rows(b)
    t = time.perf_counter(); S = ref.analyze(A, o); ta = time.perf_counter() - t
    t = time.perf_counter(); L, U, rhos, pinv = ref.factorize(A, S, o); tf = time.perf_counter() - t
    t = time.perf_counter(); x = ref.lu_solve(B, rhos, L, U, pinv); ts = time.perf_counter() - t
    B2 = ref.dense_from_rows(b)
    t = time.perf_counter(); xs = ref.solve_mpq(A, S, B2, o); tq = time.perf_counter() - t
    Lc, Uc = L.contents, U.contents
    mb = 0
    for M in (Lc, Uc):
        for k in range(M.nz):
            s = abs(M.x[k]._mp_size)
            if s:
                mb = max(mb, 64 * (s - 1) + int(M.x[k]._mp_d[s - 1]).bit_length())
    Lp = [Lc.p[k] for k in range(n + 1)]
    upd = sum(Lp[Uc.i[m] + 1] - Lp[Uc.i[m]] - 1 for m in range(Uc.nz))
    nrhs = len(b[0])
    rec = dict(
        name=name, family=family, n=n, nnz=len(X), nrhs=nrhs, nnz_L=Lc.nz, nnz_U=Uc.nz,
        det_bits=abs(capi.mpz_to_int(rhos[n - 1])).bit_length(), max_entry_bits=mb, ref_updates=upd,
        hadamard_bits=refmats.hadamard_bits(n, J, X),
        ref_seconds=dict(analyze=ta, factorize=tf, lu_solve=ts, solve_mpq=tq),
        digests=dict(L=str(ob.digest_slip_sparse(L)), U=str(ob.digest_slip_sparse(U)),
                     rhos=str(ob.digest_mpz_array(rhos, n)),
                     pinv=refmats.digest_ints(list(pinv)),
                     q=refmats.digest_ints([S.contents.q[k] for k in range(n)]),
                     x_lu_solve=str(refmats.digest_mpq_mat(ref, x, n, nrhs)),
                     x_solve_mpq=str(refmats.digest_mpq_mat(ref, xs, n, nrhs))))
    ref.free_mpq_mat(x, n, nrhs); ref.free_mpq_mat(xs, n, nrhs)
    ref.free_sparse(L); ref.free_sparse(U); ref.free_mpz_array(rhos, n); ref.free_analysis(S)
    ref.free_dense(B); ref.free_dense(B2); ref.free_sparse(A)
    print(f"{name}: n={n} nnzL={rec['nnz_L']} nnzU={rec['nnz_U']} det_bits={rec['det_bits']} "
          f"hadamard={rec['hadamard_bits']:.0f} factor={tf:.3f}s solve_mpq={tq:.3f}s", flush=True)
    return rec


def one(name, mat, rhs, family):
    ref = capi.SlipLib(ob.REF_SO)
    n, I, J, X = synth.read_triplet_file(mat)
    rec = reference_record(ref, name, n, I, J, X, synth.read_dense_file(rhs), family)
    print(json.dumps(rec))


def main():
    ob.build()
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        return one(*sys.argv[2:6])
    todo = [("NSR8K", os.path.join(REFMATS, "NSR8K_mat.txt"), os.path.join(REFMATS, "NSR8K_v.txt"), "ExampleMats"),
            ("prob159", os.path.join(REFMATS, "prob159_mat.txt"), os.path.join(REFMATS, "prob159_v.txt"), "ExampleMats")]
    bl = os.path.join(REFMATS, "BasisLIB_ALL", "RHS")
    for f in sorted(os.listdir(bl)):
        if not f.endswith(".mat"):
            continue
        with open(os.path.join(bl, f)) as fh:
            n = int(fh.readline().split()[0])
        if 300 <= n < 700 and os.path.exists(os.path.join(bl, f + ".rhs")):
            todo.append(("basislib/" + f[:-4], os.path.join(bl, f), os.path.join(bl, f + ".rhs"), "BasisLIB"))
    arrays, records = {}, []
    for name, mat, rhs, family in todo:
        # one child process per system: a few BasisLIB bases (gen1, gen2, gen4) keep the reference
        # busy for many minutes; those are left out (limit: LIMIT_S seconds of reference time)
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name, mat, rhs, family],
                                 capture_output=True, text=True, timeout=LIMIT_S)
        except subprocess.TimeoutExpired:
            print(f"{name}: reference needs more than {LIMIT_S} s; skipped", flush=True)
            continue
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if out.returncode != 0 or not line:
            print(f"{name}: reference failed; skipped\n{out.stderr[-300:]}", flush=True)
            continue
        rec = json.loads(line[-1])
        print(f"{name}: n={rec['n']} nnzL={rec['nnz_L']} nnzU={rec['nnz_U']} det_bits={rec['det_bits']} "
              f"factor={rec['ref_seconds']['factorize']:.3f}s", flush=True)
        records.append(rec)
        n, I, J, X = synth.read_triplet_file(mat)
        refmats.pack(arrays, name, n, I, J, X, synth.read_dense_file(rhs))
    os.makedirs(os.path.join(HERE, "mats"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "mats", "refmats.npz"), **arrays)
    with open(os.path.join(HERE, "refmats.json"), "w") as f:
        json.dump(dict(note="outputs of the unmodified reference (default options: COLAMD, SLIP_TOL_SMALLEST, tol 1) "
                            "on its own ExampleMats / BasisLIB systems; ref_seconds measured in the build container",
                       records=records), f, indent=0)
    print("packed", len(records), "systems,", os.path.getsize(os.path.join(HERE, "mats", "refmats.npz")), "bytes")


if __name__ == "__main__":
    main()
