"""Parity on the reference distribution's own systems (ExampleMats/NSR8K, prob159, BasisLIB LP bases)
and on synthetic systems of the BASELINE families at sizes beyond the explicit fixtures.

What the unmodified reference computed for them is recorded in tests/golden/refmats.json and
tests/golden/synth_records.json (digests of L, U, rhos, pinv, q and x; made by
tests/golden/make_refmats.py / make_synth_records.py from oracle/_ref/libslip_ref.so).

CPU tests: the oracle restatement reproduces those records (pins the oracle at size).
GPU tests: the product reproduces them through SLIP_solve_mpq (x and the row permutation it chose)
and through SLIP_LU_factorize + SLIP_LU_solve (L, U, rhos, pinv, x), bit for bit.
"""
import ctypes as C

import pytest

from slip_lu_b200 import refmats, synth

RECS = refmats.records()


def have(name):
    return name in RECS


# systems the CPU oracle finishes in seconds
ORACLE_CASES = ["prob159", "basislib/25fv47", "basislib/bas1lp", "basislib/neos7", "basislib/qiu",
                "basislib/rosen2", "basislib/route", "basislib/stair"]
# through SLIP_solve_mpq on the GPU: every recorded system (61 of the reference's own + 4 synthetic at size)
SOLVE_CASES = sorted(RECS)
# through SLIP_LU_factorize + SLIP_LU_solve on the GPU (L and U come to the host as mpz_t)
FACTOR_CASES = ["NSR8K", "prob159", "synth/rand240", "synth/rand600", "synth/lap24",
                "basislib/rat7a", "basislib/newman2", "basislib/aa01", "basislib/cr42", "basislib/complex",
                "basislib/t0331-4l", "basislib/model4"]


def csc_of(name):
    n, I, J, X, b = refmats.system(name)
    cp, ri, vals = synth.triplets_to_csc(n, I, J, X)
    return n, cp, ri, vals, b


@pytest.mark.parametrize("name", ORACLE_CASES)
def test_oracle_reproduces_reference_records(product, oracle, name):
    """oracle/ref_oracle.c == the unmodified reference on the reference's own matrices (column order
    from the product's SLIP_LU_analyze, which must be the reference's COLAMD order)."""
    from conftest import have_ordering_library
    if not have(name):
        pytest.skip("record missing")
    if not have_ordering_library():
        pytest.skip("no COLAMD library")
    rec = RECS[name]
    n, I, J, X, b = refmats.system(name)
    A = product.sparse_from_triplets(n, I, J, X)
    o = product.default_options()
    S = product.analyze(A, o)
    q = [S.contents.q[k] for k in range(n)]
    assert refmats.digest_ints(q) == rec["digests"]["q"], "column order differs from the reference's"
    cp = [A.contents.p[k] for k in range(n + 1)]
    ri = [A.contents.i[k] for k in range(cp[n])]
    from slip_lu_b200.capi import mpz_to_int
    vals = [mpz_to_int(A.contents.x[k]) for k in range(cp[n])]
    f = oracle.factorize(n, cp, ri, vals, q)
    dL, dU, dr = f.digests()
    assert str(dL) == rec["digests"]["L"] and str(dU) == rec["digests"]["U"] and str(dr) == rec["digests"]["rhos"]
    assert refmats.digest_ints(f.pinv_py()) == rec["digests"]["pinv"]
    assert abs(f.rhos_py()[-1]).bit_length() == rec["det_bits"]
    product.free_analysis(S); product.free_sparse(A); product.free_options(o)


def test_records_cover_the_bench_workloads():
    for name in ("NSR8K", "prob159", "synth/rand240", "synth/lap24"):
        assert have(name), f"{name}: no reference record (tests/golden/make_refmats.py / make_synth_records.py)"
        for key in ("L", "U", "rhos", "pinv", "q", "x_lu_solve", "x_solve_mpq"):
            assert RECS[name]["digests"][key]


@pytest.mark.gpu
@pytest.mark.parametrize("name", SOLVE_CASES)
def test_solve_mpq_matches_reference_record(gpu, name):
    """x of SLIP_solve_mpq and the row permutation the GPU path chose (the approximate pivot search
    and the bound-mode restarts included) against the REFERENCE's x and pinv."""
    if not have(name):
        pytest.skip("record missing")
    lib = gpu
    lib.dll.SLIP_B200_last_pinv.argtypes = [C.POINTER(C.c_int32), C.c_int]
    rec = RECS[name]
    n, I, J, X, b = refmats.system(name)
    A = lib.sparse_from_triplets(n, I, J, X)
    B = lib.dense_from_rows(b)
    o = lib.default_options()
    S = lib.analyze(A, o)
    assert refmats.digest_ints([S.contents.q[k] for k in range(n)]) == rec["digests"]["q"]
    x = lib.solve_mpq(A, S, B, o)
    pv = (C.c_int32 * n)()
    assert lib.dll.SLIP_B200_last_pinv(pv, n) == n
    assert refmats.digest_ints(list(pv)) == rec["digests"]["pinv"], "row permutation differs from the reference's"
    assert str(refmats.digest_mpq_mat(lib, x, n, len(b[0]))) == rec["digests"]["x_solve_mpq"]
    lib.free_mpq_mat(x, n, len(b[0])); lib.free_analysis(S); lib.free_dense(B); lib.free_sparse(A); lib.free_options(o)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FACTOR_CASES)
def test_factorize_matches_reference_record(gpu, oracle, name):
    """L, U, rhos, pinv of SLIP_LU_factorize and x of SLIP_LU_solve against the reference's digests."""
    if not have(name):
        pytest.skip("record missing")
    lib = gpu
    rec = RECS[name]
    n, I, J, X, b = refmats.system(name)
    A = lib.sparse_from_triplets(n, I, J, X)
    B = lib.dense_from_rows(b)
    o = lib.default_options()
    S = lib.analyze(A, o)
    L, U, rhos, pinv = lib.factorize(A, S, o)
    d = rec["digests"]
    assert L.contents.nz == rec["nnz_L"] and U.contents.nz == rec["nnz_U"]
    assert refmats.digest_ints(list(pinv)) == d["pinv"]
    assert str(oracle.digest_mpz_array(rhos, n)) == d["rhos"]
    assert str(oracle.digest_slip_sparse(L)) == d["L"]
    assert str(oracle.digest_slip_sparse(U)) == d["U"]
    x = lib.lu_solve(B, rhos, L, U, pinv)
    assert str(refmats.digest_mpq_mat(lib, x, n, len(b[0]))) == d["x_lu_solve"]
    lib.free_mpq_mat(x, n, len(b[0]))
    lib.free_sparse(L); lib.free_sparse(U); lib.free_mpz_array(rhos, n)
    lib.free_analysis(S); lib.free_dense(B); lib.free_sparse(A); lib.free_options(o)
