"""CPU: the C-ABI library builds, loads, and exports every symbol the headers declare; the host
logic that needs no GPU behaves like the reference; the compute entry points fail loudly
without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from slip_lu_b200 import capi, synth
import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(SLIP_[A-Za-z0-9_]+|slipcu_[a-z0-9_]+)\s*\(", txt))
                  - {"SLIP_FREE", "SLIP_free (p) ; (p) = NULL ; }"})


def test_exports_every_declared_symbol(product):
    out = subprocess.check_output(["nm", "-D", "--defined-only", product.path], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for h in ("SLIP_LU.h", "slip_b200_device.h") for s in _declared(h)
               if s not in exported and s not in ("SLIP_FREE",)]
    assert not missing, f"declared but not exported: {missing}"


def test_no_oracle_in_product(product):
    """The product library must not depend on anything under oracle/."""
    out = subprocess.check_output(["ldd", product.path], text=True)
    assert "oracle" not in out and "libslip_ref" not in out and "libref_oracle" not in out
    needed = subprocess.check_output(["readelf", "-d", product.path], text=True)
    assert "libgmp" in needed


def test_struct_layout_matches_interface(product):
    assert C.sizeof(capi.SLIP_options) == 40
    assert C.sizeof(capi.SLIP_sparse) == 72
    assert C.sizeof(capi.SLIP_dense) == 48
    o = product.default_options()
    assert (o.contents.pivot, o.contents.order, o.contents.tol, o.contents.print_level,
            o.contents.prec, o.contents.SLIP_MPFR_ROUND) == (3, 1, 1.0, 0, 128, 0)
    product.free_options(o)


def test_builders_roundtrip(product):
    n, cp, ri, vals, b = synth.random_sparse(30, 4, 90, seed=4, nrhs=2, rhs_bits=70)
    A = product.sparse_from_csc(n, cp, ri, vals)
    assert product.sparse_to_py(A) == (cp, ri, vals)
    assert capi.mpq_to_pair(A.contents.scale) == (1, 1)
    I = [i for i in ri]
    J = [j for j in range(n) for _ in range(cp[j], cp[j + 1])]
    T = product.sparse_from_triplets(n, I, J, vals)
    assert product.sparse_to_py(T) == (cp, ri, vals)
    B = product.dense_from_rows(b)
    assert [[capi.mpz_to_int(B.contents.x[r][c]) for c in range(2)] for r in range(n)] == b
    assert product.dll.SLIP_spok(A, product.default_options()) == 0
    product.free_sparse(A); product.free_sparse(T); product.free_dense(B)


def test_double_builder_matches_reference_golden(product):
    g = cases.load_golden("double_builders")
    o = product.default_options()
    for key in ("decimal", "tricky", "single"):
        d = g[key]
        vals = d["doubles"]
        nz = len(vals)
        if key == "decimal":
            n, cp, ri = d["n"], d["colptr"], d["rowidx"]
        else:
            n, cp, ri = nz, list(range(nz + 1)), list(range(nz))
        A = product.dll.SLIP_create_sparse()
        rc = product.dll.SLIP_build_sparse_ccf_double(A, (C.c_int32 * (n + 1))(*cp), (C.c_int32 * nz)(*ri),
                                                      (C.c_double * nz)(*vals), n, nz, o)
        assert rc == 0
        assert product.sparse_to_py(A)[2] == [int(v) for v in d["ints"]], key
        assert capi.mpq_to_pair(A.contents.scale) == (int(d["scale"][0]), int(d["scale"][1])), key
        product.free_sparse(A)


def test_mpfr_builders_match_reference_golden(product):
    """SLIP_build_sparse_{ccf,trip}_mpfr / SLIP_build_dense_mpfr (slip_expand_mpfr_array.c,
    slip_expand_mpfr_mat.c): same integers and scale as the reference, at two precisions."""
    g = cases.load_golden("mpfr_builders")
    for case in g["cases"]:
        o = product.default_options()
        o.contents.prec = case["prec"]
        lit = case["literals"]
        nz = len(lit)
        x = product.dll.SLIP_create_mpfr_array(nz, o)
        for k, t in enumerate(lit):
            capi.mpfr_set_decimal(x[k], t)
        A = product.dll.SLIP_create_sparse()
        rc = product.dll.SLIP_build_sparse_ccf_mpfr(A, (C.c_int32 * (nz + 1))(*range(nz + 1)),
                                                    (C.c_int32 * nz)(*range(nz)), x, nz, nz, o)
        assert rc == 0
        assert product.sparse_to_py(A)[2] == [int(v) for v in case["ints"]], case["prec"]
        assert capi.mpq_to_pair(A.contents.scale) == tuple(int(v) for v in case["scale"])
        product.free_sparse(A)
        M = product.dll.SLIP_create_mpfr_mat(2, 4, o)
        for r in range(2):
            for c in range(4):
                capi.mpfr_set_decimal(M[r][c], lit[4 * r + c])
        D = product.dll.SLIP_create_dense()
        assert product.dll.SLIP_build_dense_mpfr(D, M, 2, 4, o) == 0
        assert [[capi.mpz_to_int(D.contents.x[r][c]) for c in range(4)] for r in range(2)] == \
            [[int(v) for v in row] for row in case["dense_ints"]]
        assert capi.mpq_to_pair(D.contents.scale) == tuple(int(v) for v in case["dense_scale"])
        product.free_dense(D)
        x1 = product.dll.SLIP_create_mpfr_array(1, o)
        capi.mpfr_set_decimal(x1[0], "-0.75")
        A1 = product.dll.SLIP_create_sparse()
        assert product.dll.SLIP_build_sparse_trip_mpfr(A1, (C.c_int32 * 1)(0), (C.c_int32 * 1)(0), x1, 1, 1, o) == 0
        assert product.sparse_to_py(A1)[2] == [int(v) for v in case["single_ints"]]
        assert capi.mpq_to_pair(A1.contents.scale) == tuple(int(v) for v in case["single_scale"])
        product.free_sparse(A1)
        Mp = C.pointer(M); product.dll.SLIP_delete_mpfr_mat(Mp, 2, 4)
        xp = C.pointer(x); product.dll.SLIP_delete_mpfr_array(xp, nz)
        xp1 = C.pointer(x1); product.dll.SLIP_delete_mpfr_array(xp1, 1)
    # argument checks of the reference
    o = product.default_options()
    A = product.dll.SLIP_create_sparse()
    assert product.dll.SLIP_build_sparse_ccf_mpfr(A, None, None, None, 3, 3, o) == capi.SLIP_INCORRECT_INPUT
    product.free_sparse(A)


def test_get_mpfr_soln_rounds_like_mpfr_set_q(product):
    o = product.default_options()
    o.contents.prec = 80
    xq = product.dll.SLIP_create_mpq_mat(2, 1)
    capi.int_to_mpz(xq[0][0]._mp_num, 1); capi.int_to_mpz(xq[0][0]._mp_den, 3)
    capi.int_to_mpz(xq[1][0]._mp_num, -22); capi.int_to_mpz(xq[1][0]._mp_den, 7)
    X = product.dll.SLIP_create_mpfr_mat(2, 1, o)
    assert product.dll.SLIP_get_mpfr_soln(X, xq, 2, 1, o) == 0
    from fractions import Fraction
    for r, want in ((0, Fraction(1, 3)), (1, Fraction(-22, 7))):
        m, e = capi.mpfr_to_pair(X[r][0])
        got = Fraction(m) * (Fraction(2) ** e)
        assert abs(got - want) <= abs(want) * Fraction(1, 2 ** 79) and abs(m).bit_length() <= 80
    assert product.dll.SLIP_get_mpfr_soln(None, xq, 2, 1, o) == capi.SLIP_INCORRECT_INPUT


def test_analyze_matches_golden_orderings(product):
    """SLIP_LU_analyze (COLAMD / AMD through the SuiteSparse library found at run time, and the
    nnz guesses) reproduces the reference's q."""
    from conftest import have_ordering_library
    if not have_ordering_library():
        pytest.skip("no SuiteSparse ordering library available here")
    for name in ("10teams_default", "10teams_amd_largest", "rand120_colamd", "lap100_amd_tol"):
        g = cases.load_golden(name)
        n, cp, ri, vals, _ = cases.golden_system(g)
        A = product.sparse_from_csc(n, cp, ri, vals)
        o = product.default_options(order=g["options"]["order"])
        S = product.analyze(A, o)
        assert [S.contents.q[k] for k in range(n)] == g["q"], name
        product.free_analysis(S); product.free_sparse(A)
    A = product.sparse_from_csc(3, [0, 1, 2, 3], [0, 1, 2], [1, 2, 3])
    S = product.analyze(A, product.default_options(order=capi.SLIP_NO_ORDERING))
    assert [S.contents.q[k] for k in range(4)] == [0, 1, 2, 3] and S.contents.lnz == 5  # 10*nz clamped to ceil(n*n/2)


def test_permute_and_scale(product):
    n, nrhs = 4, 2
    x = product.dll.SLIP_create_mpq_mat(n, nrhs)
    for r in range(n):
        for c in range(nrhs):
            capi.int_to_mpz(x[r][c]._mp_num, 10 * r + c + 1)
    S = product.dll.SLIP_create_LU_analysis(n + 1)
    for k, v in enumerate([2, 0, 3, 1]):
        S.contents.q[k] = v
    assert product.dll.SLIP_permute_x(x, n, nrhs, S) == 0
    got = product.mpq_mat_to_py(x, n, nrhs)
    # x2[q[i]] = x[i]
    assert [row[0][0] for row in got] == [11, 31, 1, 21]


def test_compute_fails_loudly_without_gpu(product):
    if product.dll.SLIP_B200_device_count() > 0:
        pytest.skip("a GPU is present")
    n, cp, ri, vals, b = synth.random_sparse(10, 3, 16, seed=1)
    A = product.sparse_from_csc(n, cp, ri, vals)
    o = product.default_options(order=capi.SLIP_NO_ORDERING)
    S = product.analyze(A, o)
    with pytest.raises(capi.SlipError):
        product.factorize(A, S, o)
    B = product.dense_from_rows(b)
    with pytest.raises(capi.SlipError):
        product.solve_mpq(A, S, B, o)
    product.dll.SLIP_B200_last_error.restype = C.c_char_p
    assert b"no CUDA device" in product.dll.SLIP_B200_last_error()


def test_bad_arguments(product):
    o = product.default_options()
    assert product.dll.SLIP_LU_factorize(None, None, None, None, None, None, o) == capi.SLIP_INCORRECT_INPUT
    assert product.dll.SLIP_LU_solve(None, None, None, None, None, None) == capi.SLIP_INCORRECT_INPUT
    assert product.dll.SLIP_LU_analyze(None, None, o) == capi.SLIP_INCORRECT_INPUT


def test_header_is_layout_compatible_with_the_reference_header(tmp_path):
    """Compiles one probe against include/SLIP_LU.h and one against the reference's own header
    (where the reference tree is present) and compares sizeof / offsetof of every public struct and
    the enum values: a program built against either header can be linked with either library."""
    ref_inc = "/root/reference/SLIP_LU/Include"
    if not os.path.exists(os.path.join(ref_inc, "SLIP_LU.h")):
        pytest.skip("reference tree not present on this machine")
    probe = r'''
#include <stddef.h>
#include <stdio.h>
#include "SLIP_LU.h"
#define F(s, m) printf (#s "." #m " %zu %zu\n", offsetof (s, m), sizeof (((s *) 0)->m))
int main (void)
{
    printf ("sizeof %zu %zu %zu %zu\n", sizeof (SLIP_options), sizeof (SLIP_sparse), sizeof (SLIP_dense), sizeof (SLIP_LU_analysis)) ;
    F (SLIP_options, pivot) ; F (SLIP_options, order) ; F (SLIP_options, tol) ; F (SLIP_options, print_level) ;
    F (SLIP_options, prec) ; F (SLIP_options, SLIP_MPFR_ROUND) ;
    F (SLIP_sparse, m) ; F (SLIP_sparse, n) ; F (SLIP_sparse, nzmax) ; F (SLIP_sparse, nz) ; F (SLIP_sparse, p) ;
    F (SLIP_sparse, i) ; F (SLIP_sparse, x) ; F (SLIP_sparse, scale) ;
    F (SLIP_dense, m) ; F (SLIP_dense, n) ; F (SLIP_dense, x) ; F (SLIP_dense, scale) ;
    F (SLIP_LU_analysis, q) ; F (SLIP_LU_analysis, lnz) ; F (SLIP_LU_analysis, unz) ;
    printf ("info %d %d %d %d %d\n", SLIP_OK, SLIP_OUT_OF_MEMORY, SLIP_SINGULAR, SLIP_INCORRECT_INPUT, SLIP_INCORRECT) ;
    printf ("pivot %d %d %d %d %d %d\n", SLIP_SMALLEST, SLIP_DIAGONAL, SLIP_FIRST_NONZERO, SLIP_TOL_SMALLEST, SLIP_TOL_LARGEST, SLIP_LARGEST) ;
    printf ("order %d %d %d\n", SLIP_NO_ORDERING, SLIP_COLAMD, SLIP_AMD) ;
    return 0 ;
}
'''
    src = tmp_path / "probe.c"
    src.write_text(probe)
    have_gmp_h = subprocess.run(["gcc", "-include", "gmp.h", "-include", "mpfr.h", "-E", "-x", "c", "/dev/null"],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0
    abi = [] if have_gmp_h else ["-I" + os.path.join(ROOT, "include", "gmp_abi")]
    outs = []
    for name, inc in (("ours", [os.path.join(ROOT, "include")]),
                      ("ref", [ref_inc, "/root/reference/SuiteSparse_config", "/root/reference/COLAMD/Include",
                               "/root/reference/AMD/Include"])):
        exe = tmp_path / f"probe_{name}"
        subprocess.check_call(["gcc", "-w", "-o", str(exe), str(src)] + ["-I" + d for d in inc] + abi)
        outs.append(subprocess.check_output([str(exe)], text=True))
    assert outs[0] == outs[1]


def test_reference_demos_compile_and_link_unchanged(product, tmp_path):
    """The reference's own demo programs (Demo/example*.c, SLIPLU.c with demos.c) are compiled with
    -Wall -Werror against include/SLIP_LU.h and linked against libslip_lu_b200.so without a change:
    every type, prototype and symbol they use is there (they are not run here: that needs a GPU;
    the 10teams demo system is a golden fixture of the GPU tests)."""
    demo = "/root/reference/SLIP_LU/Demo"
    if not os.path.exists(os.path.join(demo, "demos.c")):
        pytest.skip("reference tree not present on this machine")
    have_gmp_h = subprocess.run(["gcc", "-include", "gmp.h", "-include", "mpfr.h", "-E", "-x", "c", "/dev/null"],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + demo]
    libs = ["-lgmp", "-lmpfr"] if have_gmp_h else ["-l:libgmp.so.10", "-l:libmpfr.so.6"]
    if not have_gmp_h:
        inc.append("-I" + os.path.join(ROOT, "include", "gmp_abi"))
    libdir = os.path.dirname(product.path)
    for prog in ("example", "example2", "example3", "example4", "example5", "SLIPLU"):
        subprocess.check_call(["gcc", "-Wall", "-Werror", "-Wno-unused", "-O1"] + inc +
                              ["-o", str(tmp_path / prog), os.path.join(demo, prog + ".c"), os.path.join(demo, "demos.c"),
                               "-L" + libdir, "-lslip_lu_b200", "-Wl,-rpath," + libdir] + libs + ["-lm"])
