"""Shared test cases and comparison helpers."""
from __future__ import annotations

import json
import os

from slip_lu_b200 import capi, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def colamd_like_order(n, colptr, rowidx):
    """A deterministic column order that needs no ordering library: by column count, ties by index."""
    return sorted(range(n), key=lambda j: (colptr[j + 1] - colptr[j], j))


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def small_cases():
    """(name, n, colptr, rowidx, values, b) for sizes the CPU oracle finishes in well under a second."""
    out = []
    out.append(("rand24", *synth.random_sparse(24, 4, 16, seed=7, nrhs=2)))
    out.append(("rand60", *synth.random_sparse(60, 6, 32, seed=11, nrhs=3)))
    out.append(("rand90_64bit", *synth.random_sparse(90, 5, 64, seed=5, nrhs=1, rhs_bits=64)))
    out.append(("lap49", *synth.laplacian_2d(7, 64, seed=2, nrhs=2)))
    out.append(("lap100_small", *synth.laplacian_2d(10, 8, seed=3, nrhs=1, rhs_bits=8)))
    out.append(("lp200", *synth.lp_basis(200, seed=5, nrhs=2)))
    out.append(("dense12", *synth.random_sparse(12, 12, 20, seed=9, nrhs=4)))
    out.append(("n1", 1, [0, 1], [0], [-7], [[21]]))
    return out


def run_library(lib: capi.SlipLib, n, cp, ri, vals, b, q, pivot=capi.SLIP_TOL_SMALLEST, tol=None):
    """factorize + LU_solve through the C interface; returns python data."""
    o = lib.default_options(pivot=pivot, order=capi.SLIP_NO_ORDERING, tol=tol)
    A = lib.sparse_from_csc(n, cp, ri, vals)
    B = lib.dense_from_rows(b)
    S = lib.analyze(A, o, q=q)
    L, U, rhos, pinv = lib.factorize(A, S, o)
    x = lib.lu_solve(B, rhos, L, U, pinv)
    res = dict(L=lib.sparse_to_py(L), U=lib.sparse_to_py(U), rhos=lib.mpz_array_to_py(rhos, n),
               pinv=list(pinv), x=lib.mpq_mat_to_py(x, n, len(b[0])))
    lib.free_mpq_mat(x, n, len(b[0]))
    lib.free_sparse(L); lib.free_sparse(U); lib.free_mpz_array(rhos, n)
    lib.free_analysis(S); lib.free_dense(B); lib.free_sparse(A); lib.free_options(o)
    return res


def run_oracle(ob, n, cp, ri, vals, b, q, pivot=capi.SLIP_TOL_SMALLEST, tol=1.0):
    f = ob.factorize(n, cp, ri, vals, q, pivot, tol)
    return dict(L=f.L_py(), U=f.U_py(), rhos=f.rhos_py(), pinv=f.pinv_py(), x=ob.solve(f, b))


def assert_same_factorization(got, want, label=""):
    assert got["pinv"] == want["pinv"], f"{label}: row permutation differs"
    assert got["rhos"] == want["rhos"], f"{label}: pivots differ at " \
        f"{next(i for i, (a, b) in enumerate(zip(got['rhos'], want['rhos'])) if a != b)}"
    for nm in ("L", "U"):
        gp, gi, gx = got[nm]
        wp, wi, wx = want[nm]
        assert gp == wp, f"{label}: {nm} column pointers differ"
        assert gi == wi, f"{label}: {nm} row indices differ"
        if gx != wx:
            bad = next(i for i, (a, b) in enumerate(zip(gx, wx)) if a != b)
            col = max(k for k in range(len(gp)) if gp[k] <= bad)
            raise AssertionError(f"{label}: {nm} value {bad} (column {col}) differs: {gx[bad]} vs {wx[bad]}")
    assert got["x"] == want["x"], f"{label}: solution differs"


def digest_pairs(rows):
    """Order-dependent digest of a matrix of (num, den) pairs (same as tests/golden/make_golden.py)."""
    import hashlib
    h = hashlib.sha256()
    for row in rows:
        for a, d in row:
            h.update(f"{a}/{d};".encode())
    return int.from_bytes(h.digest()[:8], "little")


GOLDEN_FACTOR_CASES = ["10teams_default", "10teams_amd_largest", "rand120_colamd", "lap144_colamd",
                       "lap100_amd_tol", "lp300_colamd", "decimal80_multirhs"] + \
                      [f"testmat_pivot{p}" for p in range(6)]


def golden_system(g):
    vals = [int(v) for v in g["values"]]
    b = [[int(v) for v in row] for row in g["b"]]
    return g["n"], g["colptr"], g["rowidx"], vals, b


def check_against_golden(g, got, ob=None, raw=None):
    """got: dict from run_library/run_oracle (python ints).  raw: optional (L, U, rhos) C objects
    for digest comparison without python conversion."""
    name = g["name"]
    assert got["pinv"] == g["pinv"], f"{name}: pinv"
    assert got["rhos"][-1] == int(g["det"]), f"{name}: determinant"
    assert got["L"][0] == g["L"]["p"] and got["L"][1] == g["L"]["i"], f"{name}: L pattern"
    assert got["U"][0] == g["U"]["p"] and got["U"][1] == g["U"]["i"], f"{name}: U pattern"
    assert digest_pairs(got["x"]) == int(g["digests"]["x_lu_solve"]), f"{name}: x digest"
    if g["explicit"]:
        assert got["rhos"] == [int(v) for v in g["rhos"]], f"{name}: rhos"
        assert got["L"][2] == [int(v) for v in g["L"]["x"]], f"{name}: L values"
        assert got["U"][2] == [int(v) for v in g["U"]["x"]], f"{name}: U values"
        assert got["x"] == [[(int(a), int(d)) for a, d in row] for row in g["x_lu_solve"]], f"{name}: x"
    if ob is not None:
        # digest of (p, i, limbs) exactly as the reference laid them out
        import ctypes as C
        from slip_lu_b200.capi import MpzStruct
        for nm in ("L", "U"):
            p, i, x = got[nm]
            arr = ob.MpzArray(x)
            d = ob.dll().ro_digest_csc(g["n"], (C.c_int * len(p))(*p), (C.c_int * max(len(i), 1))(*i), arr.arr)
            assert d == int(g["digests"][nm]), f"{name}: {nm} digest"
        arr = ob.MpzArray(got["rhos"])
        assert ob.dll().ro_digest_mpz(arr.arr, g["n"]) == int(g["digests"]["rhos"]), f"{name}: rhos digest"
