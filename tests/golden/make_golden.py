"""Generates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libslip_ref.so).

Run in the build container (needs /root/reference for the ExampleMats inputs and the _ref build):
    python tests/golden/make_golden.py
Each fixture holds the input system, the options, and the reference's outputs: the column order
q, L, U, rhos, pinv (SLIP_LU_factorize), x from SLIP_LU_solve (factor order) and x from
SLIP_solve_mpq (original order, scaled).  Integers are stored as decimal strings (JSON numbers
cannot be trusted beyond 2^53 by all readers).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slip_lu_b200 import capi, synth  # noqa: E402
from oracle import binding as ob      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REFMATS = "/root/reference/SLIP_LU/ExampleMats"


def S(v):
    return str(v)


def digest_pairs(rows):
    """Order-dependent digest of a matrix of (num, den) pairs; same function in tests/cases.py."""
    import hashlib
    h = hashlib.sha256()
    for row in rows:
        for a, d in row:
            h.update(f"{a}/{d};".encode())
    return int.from_bytes(h.digest()[:8], "little")


def fixture(ref, name, n, cp, ri, vals, b, pivot, order, tol=None, note=""):
    o = ref.default_options(pivot=pivot, order=order, tol=tol)
    A = ref.sparse_from_csc(n, cp, ri, vals)
    B = ref.dense_from_rows(b)
    Sx = ref.analyze(A, o)
    q = [Sx.contents.q[k] for k in range(n)]
    L, U, rhos, pinv = ref.factorize(A, Sx, o)
    x = ref.lu_solve(B, rhos, L, U, pinv)
    nrhs = len(b[0])
    B2 = ref.dense_from_rows(b)
    xs = ref.solve_mpq(A, Sx, B2, o)
    Lp, Li, Lx = ref.sparse_to_py(L)
    Up, Ui, Ux = ref.sparse_to_py(U)
    rh = ref.mpz_array_to_py(rhos, n)
    x1 = ref.mpq_mat_to_py(x, n, nrhs)
    x2 = ref.mpq_mat_to_py(xs, n, nrhs)
    doc = dict(
        name=name, note=note, n=n, colptr=cp, rowidx=ri, values=[S(v) for v in vals],
        b=[[S(v) for v in row] for row in b],
        options=dict(pivot=pivot, order=order, tol=o.contents.tol),
        q=q, pinv=list(pinv), det=S(rh[-1]),
        L=dict(p=Lp, i=Li), U=dict(p=Up, i=Ui),
        digests=dict(L=S(ob.digest_slip_sparse(L)), U=S(ob.digest_slip_sparse(U)),
                     rhos=S(ob.digest_mpz_array(rhos, n)),
                     x_lu_solve=S(digest_pairs(x1)), x_solve_mpq=S(digest_pairs(x2))),
    )
    digits = sum(len(S(v)) for v in Lx) + sum(len(S(v)) for v in Ux)
    doc["explicit"] = digits < 150000
    if doc["explicit"]:
        # small cases carry every number; large ones are pinned by the digests above
        doc["rhos"] = [S(v) for v in rh]
        doc["L"]["x"] = [S(v) for v in Lx]
        doc["U"]["x"] = [S(v) for v in Ux]
        doc["x_lu_solve"] = [[[S(a), S(d)] for a, d in row] for row in x1]
        doc["x_solve_mpq"] = [[[S(a), S(d)] for a, d in row] for row in x2]
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print(name, "n", n, "nnzL", Lp[-1], "nnzU", Up[-1], "det bits", abs(rh[-1]).bit_length(), "explicit", doc["explicit"])


def doubles_fixture(ref):
    """Integer systems the reference's double builders produce (slip_expand_double_array)."""
    import ctypes as C
    (sysd, dvals) = synth.decimal_scaled(30, 4, 6, seed=3, nrhs=2)
    n, cp, ri, vals, b = sysd
    o = ref.default_options()
    A = ref.dll.SLIP_create_sparse()
    p = (C.c_int32 * (n + 1))(*cp); i = (C.c_int32 * len(ri))(*ri); x = (C.c_double * len(dvals))(*dvals)
    rc = ref.dll.SLIP_build_sparse_ccf_double(A, p, i, x, n, len(dvals), o)
    assert rc == 0
    _, _, ax = ref.sparse_to_py(A)
    num, den = capi.mpq_to_pair(A.contents.scale)
    tricky = [0.1, -0.25, 3.0, 1e-3, 123456.789, -7.5e-9, 0.0, 2.0 ** -20]
    A2 = ref.dll.SLIP_create_sparse()
    p2 = (C.c_int32 * 9)(*range(9)); i2 = (C.c_int32 * 8)(*range(8)); x2 = (C.c_double * 8)(*tricky)
    assert ref.dll.SLIP_build_sparse_ccf_double(A2, p2, i2, x2, 8, 8, o) == 0
    _, _, ax2 = ref.sparse_to_py(A2)
    num2, den2 = capi.mpq_to_pair(A2.contents.scale)
    single = [-0.75]
    A3 = ref.dll.SLIP_create_sparse()
    p3 = (C.c_int32 * 2)(0, 1); i3 = (C.c_int32 * 1)(0); x3 = (C.c_double * 1)(*single)
    assert ref.dll.SLIP_build_sparse_ccf_double(A3, p3, i3, x3, 1, 1, o) == 0
    _, _, ax3 = ref.sparse_to_py(A3)
    num3, den3 = capi.mpq_to_pair(A3.contents.scale)
    doc = dict(name="double_builders",
               decimal=dict(n=n, colptr=cp, rowidx=ri, doubles=dvals, ints=[S(v) for v in ax], scale=[S(num), S(den)]),
               tricky=dict(doubles=tricky, ints=[S(v) for v in ax2], scale=[S(num2), S(den2)]),
               single=dict(doubles=single, ints=[S(v) for v in ax3], scale=[S(num3), S(den3)]))
    with open(os.path.join(HERE, "double_builders.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("double_builders ok")


def mpfr_fixture(ref):
    """What the reference's mpfr builders and SLIP_solve_mpfr produce (slip_expand_mpfr_array/_mat,
    SLIP_get_mpfr_soln), at two precisions."""
    import ctypes as C
    doc = dict(name="mpfr_builders", cases=[])
    literals = ["0.1", "-0.25", "3", "1e-3", "123456.789", "-7.5e-9", "0", "1.52587890625e-05"]
    for prec in (128, 53):
        o = ref.default_options()
        o.contents.prec = prec
        nz = len(literals)
        x = ref.dll.SLIP_create_mpfr_array(nz, o)
        for k, t in enumerate(literals):
            capi.mpfr_set_decimal(x[k], t)
        A = ref.dll.SLIP_create_sparse()
        p = (C.c_int32 * (nz + 1))(*range(nz + 1)); i = (C.c_int32 * nz)(*range(nz))
        assert ref.dll.SLIP_build_sparse_ccf_mpfr(A, p, i, x, nz, nz, o) == 0
        _, _, ax = ref.sparse_to_py(A)
        num, den = capi.mpq_to_pair(A.contents.scale)
        # the same values as a 2 x 4 dense block, and one lone negative entry (sign of the "gcd")
        M = ref.dll.SLIP_create_mpfr_mat(2, 4, o)
        for r in range(2):
            for c in range(4):
                capi.mpfr_set_decimal(M[r][c], literals[4 * r + c])
        D = ref.dll.SLIP_create_dense()
        assert ref.dll.SLIP_build_dense_mpfr(D, M, 2, 4, o) == 0
        dx = [[S(capi.mpz_to_int(D.contents.x[r][c])) for c in range(4)] for r in range(2)]
        dnum, dden = capi.mpq_to_pair(D.contents.scale)
        x1 = ref.dll.SLIP_create_mpfr_array(1, o)
        capi.mpfr_set_decimal(x1[0], "-0.75")
        A1 = ref.dll.SLIP_create_sparse()
        assert ref.dll.SLIP_build_sparse_trip_mpfr(A1, (C.c_int32 * 1)(0), (C.c_int32 * 1)(0), x1, 1, 1, o) == 0
        _, _, ax1 = ref.sparse_to_py(A1)
        n1, d1 = capi.mpq_to_pair(A1.contents.scale)
        # SLIP_solve_mpfr on a small integer system
        n, cp, ri, vals, b = synth.random_sparse(24, 4, 20, seed=11, nrhs=2)
        As = ref.sparse_from_csc(n, cp, ri, vals); Bs = ref.dense_from_rows(b)
        o.contents.order = capi.SLIP_NO_ORDERING
        Sy = ref.analyze(As, o)
        X = ref.dll.SLIP_create_mpfr_mat(n, 2, o)
        assert ref.dll.SLIP_solve_mpfr(X, As, Sy, Bs, o) == 0
        sol = [[[S(v) for v in capi.mpfr_to_pair(X[r][c])] for c in range(2)] for r in range(n)]
        doc["cases"].append(dict(prec=prec, literals=literals, ints=[S(v) for v in ax], scale=[S(num), S(den)],
                                 dense_ints=dx, dense_scale=[S(dnum), S(dden)],
                                 single_ints=[S(v) for v in ax1], single_scale=[S(n1), S(d1)],
                                 solve=dict(n=n, nnz_per_col=4, bits=20, seed=11, nrhs=2, x=sol)))
    with open(os.path.join(HERE, "mpfr_builders.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("mpfr_builders ok")


def main():
    ob.build()
    ref = capi.SlipLib(ob.REF_SO)
    # config[0]: the reference demo system (example2.c: 10teams matrix + rhs)
    n, I, J, V = synth.read_triplet_file(os.path.join(REFMATS, "10teams_mat.txt"))
    b = synth.read_dense_file(os.path.join(REFMATS, "10teams_v.txt"))
    cp, ri, vals = synth.triplets_to_csc(n, I, J, V)
    fixture(ref, "10teams_default", n, cp, ri, vals, b, capi.SLIP_TOL_SMALLEST, capi.SLIP_COLAMD,
            note="BASELINE configs[0]: Demo/example2 input, default options")
    fixture(ref, "10teams_amd_largest", n, cp, ri, vals, b, capi.SLIP_LARGEST, capi.SLIP_AMD)
    n, I, J, V = synth.read_triplet_file(os.path.join(REFMATS, "test_mat.txt"))
    b = synth.read_dense_file(os.path.join(REFMATS, "test_rhs.txt"))
    cp, ri, vals = synth.triplets_to_csc(n, I, J, V)
    for piv in range(6):
        fixture(ref, f"testmat_pivot{piv}", n, cp, ri, vals, b, piv, capi.SLIP_COLAMD,
                tol=0.3 if piv in (3, 4) else None, note="ExampleMats/test_mat.txt + test_rhs.txt")
    # synthetic: the BASELINE config families at CPU-feasible sizes
    fixture(ref, "rand120_colamd", *synth.random_sparse(120, 8, 32, seed=3, nrhs=3),
            capi.SLIP_TOL_SMALLEST, capi.SLIP_COLAMD, note="configs[1] family (random sparse, 32-bit, COLAMD)")
    fixture(ref, "lap144_colamd", *synth.laplacian_2d(12, 64, seed=2, nrhs=2),
            capi.SLIP_TOL_SMALLEST, capi.SLIP_COLAMD, note="configs[2] family (2D Laplacian pattern, 64-bit)")
    fixture(ref, "lap100_amd_tol", *synth.laplacian_2d(10, 24, seed=4, nrhs=1, rhs_bits=16),
            capi.SLIP_TOL_LARGEST, capi.SLIP_AMD, tol=0.05)
    fixture(ref, "lp300_colamd", *synth.lp_basis(300, seed=5, nrhs=2),
            capi.SLIP_TOL_SMALLEST, capi.SLIP_COLAMD, note="configs[4] family (LP basis style)")
    (sysd, _dv) = synth.decimal_scaled(80, 5, 6, seed=8, nrhs=4)
    fixture(ref, "decimal80_multirhs", *sysd, capi.SLIP_TOL_SMALLEST, capi.SLIP_COLAMD,
            note="configs[3] family (scaled decimals, several right-hand sides)")
    doubles_fixture(ref)
    mpfr_fixture(ref)


if __name__ == "__main__":
    main()
