"""ctypes binding of the SLIP_LU C interface (include/SLIP_LU.h).

This is the Python host-side mirror of the reference's public C API for the
factor/solve path (reference: SLIP_LU/Include/SLIP_LU.h:160-993).  It binds *a*
shared library that exports that interface: the product library
``slip_lu_b200/libslip_lu_b200.so`` (see :func:`slip_lu_b200.lib`) or, in tests
and in bench.py's reference arm only, the unmodified reference built at
``oracle/_ref/libslip_ref.so``.  The struct layouts below are the C ABI of the
header and are identical for both libraries, which is what makes the product a
drop-in.

Nothing in here computes: every numeric operation is a call into the bound C
library.  Python ints are converted to/from GMP ``mpz_t`` limbs at the boundary.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import os
from typing import Iterable, List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------
# GMP ABI structs (GMP manual, "Integer Internals"/"Rational Internals")
# ----------------------------------------------------------------------------


class MpzStruct(C.Structure):
    _fields_ = [("_mp_alloc", C.c_int), ("_mp_size", C.c_int),
                ("_mp_d", C.POINTER(C.c_uint64))]


class MpqStruct(C.Structure):
    _fields_ = [("_mp_num", MpzStruct), ("_mp_den", MpzStruct)]


class MpfrStruct(C.Structure):
    # __mpfr_struct of MPFR 4 (x86-64): precision, sign, exponent, limb pointer
    _fields_ = [("_mpfr_prec", C.c_long), ("_mpfr_sign", C.c_int), ("_mpfr_exp", C.c_long),
                ("_mpfr_d", C.c_void_p)]


class SLIP_options(C.Structure):
    # reference: SLIP_LU/Include/SLIP_LU.h:212-223
    _fields_ = [("pivot", C.c_int), ("order", C.c_int), ("tol", C.c_double),
                ("print_level", C.c_int32), ("prec", C.c_uint64),
                ("SLIP_MPFR_ROUND", C.c_int)]


class SLIP_sparse(C.Structure):
    # reference: SLIP_LU/Include/SLIP_LU.h:246-256
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("nzmax", C.c_int32),
                ("nz", C.c_int32), ("p", C.POINTER(C.c_int32)),
                ("i", C.POINTER(C.c_int32)), ("x", C.POINTER(MpzStruct)),
                ("scale", MpqStruct)]


class SLIP_dense(C.Structure):
    # reference: SLIP_LU/Include/SLIP_LU.h:277-284
    _fields_ = [("m", C.c_int32), ("n", C.c_int32),
                ("x", C.POINTER(C.POINTER(MpzStruct))), ("scale", MpqStruct)]


class SLIP_LU_analysis(C.Structure):
    # reference: SLIP_LU/Include/SLIP_LU.h:303-310
    _fields_ = [("q", C.POINTER(C.c_int32)), ("lnz", C.c_int32),
                ("unz", C.c_int32)]


# error codes, pivot and ordering enums (SLIP_LU.h:160-201)
SLIP_OK, SLIP_OUT_OF_MEMORY, SLIP_SINGULAR, SLIP_INCORRECT_INPUT, SLIP_INCORRECT = 0, -1, -2, -3, -4
SLIP_SMALLEST, SLIP_DIAGONAL, SLIP_FIRST_NONZERO, SLIP_TOL_SMALLEST, SLIP_TOL_LARGEST, SLIP_LARGEST = range(6)
SLIP_NO_ORDERING, SLIP_COLAMD, SLIP_AMD = range(3)


class SlipError(RuntimeError):
    def __init__(self, code: int, where: str):
        names = {0: "SLIP_OK", -1: "SLIP_OUT_OF_MEMORY", -2: "SLIP_SINGULAR",
                 -3: "SLIP_INCORRECT_INPUT", -4: "SLIP_INCORRECT"}
        super().__init__(f"{where} returned {names.get(code, code)}")
        self.code = code


def _find_gmp() -> C.CDLL:
    for name in ("libgmp.so.10", "libgmp.so", ctypes.util.find_library("gmp")):
        if not name:
            continue
        try:
            return C.CDLL(name, mode=C.RTLD_GLOBAL)
        except OSError:
            continue
    raise OSError("libgmp not found")


_gmp = None


class _Gmp:
    """libgmp entry points under their documented names (mpz_init, ...)."""

    def __init__(self, dll: C.CDLL):
        self.dll = dll
        Z, Q = C.POINTER(MpzStruct), C.POINTER(MpqStruct)
        table = {
            "mpz_init": ("__gmpz_init", None, [Z]),
            "mpz_clear": ("__gmpz_clear", None, [Z]),
            "mpz_neg": ("__gmpz_neg", None, [Z, Z]),
            "mpz_import": ("__gmpz_import", None, [Z, C.c_size_t, C.c_int, C.c_size_t, C.c_int,
                                                   C.c_size_t, C.c_void_p]),
            "mpq_init": ("__gmpq_init", None, [Q]),
            "mpq_clear": ("__gmpq_clear", None, [Q]),
            "mpq_canonicalize": ("__gmpq_canonicalize", None, [Q]),
        }
        for name, (sym, res, args) in table.items():
            f = getattr(dll, sym)
            f.restype, f.argtypes = res, args
            setattr(self, name, f)


def gmp() -> _Gmp:
    global _gmp
    if _gmp is None:
        _gmp = _Gmp(_find_gmp())
    return _gmp


_mpfr = None


def mpfr() -> C.CDLL:
    """libmpfr with the few entry points the tests need (values in and out of mpfr_t)."""
    global _mpfr
    if _mpfr is None:
        gmp()
        for name in ("libmpfr.so.6", "libmpfr.so", ctypes.util.find_library("mpfr")):
            if not name:
                continue
            try:
                _mpfr = C.CDLL(name, mode=C.RTLD_GLOBAL)
                break
            except OSError:
                continue
        if _mpfr is None:
            raise OSError("libmpfr not found")
        F, Z = C.POINTER(MpfrStruct), C.POINTER(MpzStruct)
        _mpfr.mpfr_set_str.restype, _mpfr.mpfr_set_str.argtypes = C.c_int, [F, C.c_char_p, C.c_int, C.c_int]
        _mpfr.mpfr_get_z_2exp.restype, _mpfr.mpfr_get_z_2exp.argtypes = C.c_long, [Z, F]
    return _mpfr


def mpfr_set_decimal(x, text: str, rnd: int = 0) -> None:
    """x (an initialised mpfr_t element) = the decimal string, rounded to x's precision."""
    xp = x if isinstance(x, C.POINTER(MpfrStruct)) else C.pointer(x)
    if mpfr().mpfr_set_str(xp, text.encode(), 10, rnd) != 0:
        raise ValueError(f"bad mpfr literal {text!r}")


def mpfr_to_pair(x) -> Tuple[int, int]:
    """Exact value of a finite mpfr as (mantissa, exponent): mantissa * 2**exponent; zero is (0, 0)."""
    xp = x if isinstance(x, C.POINTER(MpfrStruct)) else C.pointer(x)
    g = gmp()
    z = MpzStruct()
    g.mpz_init(C.byref(z))
    e = mpfr().mpfr_get_z_2exp(C.byref(z), xp)
    m = mpz_to_int(z)
    g.mpz_clear(C.byref(z))
    while m and m % 2 == 0:          # normalise so that equal values compare equal
        m //= 2; e += 1
    return (m, e) if m else (0, 0)


def mpz_to_int(z: MpzStruct) -> int:
    """Read a GMP integer's limbs directly (no library call)."""
    sz = z._mp_size
    if sz == 0:
        return 0
    nl = -sz if sz < 0 else sz
    v = int.from_bytes(C.string_at(z._mp_d, nl * 8), "little")
    return -v if sz < 0 else v


def int_to_mpz(z, v: int) -> None:
    """Assign Python int v to an *initialised* mpz (pointer or struct)."""
    g = gmp()
    zp = z if isinstance(z, C.POINTER(MpzStruct)) else C.pointer(z)
    a = -v if v < 0 else v
    nbytes = max(1, (a.bit_length() + 7) // 8)
    buf = a.to_bytes(nbytes, "little")
    g.mpz_import(zp, nbytes, -1, 1, 0, 0, buf)
    if v < 0:
        g.mpz_neg(zp, zp)


def mpq_to_pair(qv: MpqStruct) -> Tuple[int, int]:
    return mpz_to_int(qv._mp_num), mpz_to_int(qv._mp_den)


class SlipLib:
    """A loaded library exporting the SLIP_LU C interface."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        gmp()  # make sure libgmp is resident with RTLD_GLOBAL
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        P = C.POINTER
        sig = {
            "SLIP_initialize": (None, []),
            "SLIP_finalize": (None, []),
            "SLIP_create_default_options": (P(SLIP_options), []),
            "SLIP_create_sparse": (P(SLIP_sparse), []),
            "SLIP_delete_sparse": (None, [P(P(SLIP_sparse))]),
            "SLIP_create_dense": (P(SLIP_dense), []),
            "SLIP_delete_dense": (None, [P(P(SLIP_dense))]),
            "SLIP_create_LU_analysis": (P(SLIP_LU_analysis), [C.c_int32]),
            "SLIP_delete_LU_analysis": (None, [P(P(SLIP_LU_analysis))]),
            "SLIP_free": (None, [C.c_void_p]),
            "SLIP_malloc": (C.c_void_p, [C.c_size_t]),
            "SLIP_create_mpz_array": (P(MpzStruct), [C.c_int32]),
            "SLIP_delete_mpz_array": (None, [P(P(MpzStruct)), C.c_int32]),
            "SLIP_create_mpz_mat": (P(P(MpzStruct)), [C.c_int32, C.c_int32]),
            "SLIP_delete_mpz_mat": (None, [P(P(P(MpzStruct))), C.c_int32, C.c_int32]),
            "SLIP_create_mpq_mat": (P(P(MpqStruct)), [C.c_int32, C.c_int32]),
            "SLIP_delete_mpq_mat": (None, [P(P(P(MpqStruct))), C.c_int32, C.c_int32]),
            "SLIP_build_sparse_ccf_mpz": (C.c_int, [P(SLIP_sparse), P(C.c_int32), P(C.c_int32),
                                                   P(MpzStruct), C.c_int32, C.c_int32]),
            "SLIP_build_sparse_trip_mpz": (C.c_int, [P(SLIP_sparse), P(C.c_int32), P(C.c_int32),
                                                    P(MpzStruct), C.c_int32, C.c_int32]),
            "SLIP_build_sparse_ccf_int": (C.c_int, [P(SLIP_sparse), P(C.c_int32), P(C.c_int32),
                                                   P(C.c_int32), C.c_int32, C.c_int32]),
            "SLIP_build_sparse_ccf_double": (C.c_int, [P(SLIP_sparse), P(C.c_int32), P(C.c_int32),
                                                      P(C.c_double), C.c_int32, C.c_int32,
                                                      P(SLIP_options)]),
            "SLIP_build_dense_mpz": (C.c_int, [P(SLIP_dense), P(P(MpzStruct)), C.c_int32, C.c_int32]),
            "SLIP_LU_analyze": (C.c_int, [P(SLIP_LU_analysis), P(SLIP_sparse), P(SLIP_options)]),
            "SLIP_LU_factorize": (C.c_int, [P(SLIP_sparse), P(SLIP_sparse), P(SLIP_sparse),
                                           P(SLIP_LU_analysis), P(MpzStruct), P(C.c_int32),
                                           P(SLIP_options)]),
            "SLIP_LU_solve": (C.c_int, [P(P(MpqStruct)), P(SLIP_dense), P(MpzStruct),
                                       P(SLIP_sparse), P(SLIP_sparse), P(C.c_int32)]),
            "SLIP_permute_x": (C.c_int, [P(P(MpqStruct)), C.c_int32, C.c_int32, P(SLIP_LU_analysis)]),
            "SLIP_scale_x": (C.c_int, [P(P(MpqStruct)), P(SLIP_sparse), P(SLIP_dense)]),
            "SLIP_solve_mpq": (C.c_int, [P(P(MpqStruct)), P(SLIP_sparse), P(SLIP_LU_analysis),
                                        P(SLIP_dense), P(SLIP_options)]),
            "SLIP_solve_double": (C.c_int, [P(P(C.c_double)), P(SLIP_sparse), P(SLIP_LU_analysis),
                                           P(SLIP_dense), P(SLIP_options)]),
            "SLIP_check_solution": (C.c_int, [P(SLIP_sparse), P(P(MpqStruct)), P(SLIP_dense)]),
            "SLIP_create_mpfr_array": (P(MpfrStruct), [C.c_int32, P(SLIP_options)]),
            "SLIP_delete_mpfr_array": (None, [P(P(MpfrStruct)), C.c_int32]),
            "SLIP_create_mpfr_mat": (P(P(MpfrStruct)), [C.c_int32, C.c_int32, P(SLIP_options)]),
            "SLIP_delete_mpfr_mat": (None, [P(P(P(MpfrStruct))), C.c_int32, C.c_int32]),
            "SLIP_build_sparse_ccf_mpfr": (C.c_int, [P(SLIP_sparse), P(C.c_int32), P(C.c_int32),
                                                    P(MpfrStruct), C.c_int32, C.c_int32, P(SLIP_options)]),
            "SLIP_build_sparse_trip_mpfr": (C.c_int, [P(SLIP_sparse), P(C.c_int32), P(C.c_int32),
                                                     P(MpfrStruct), C.c_int32, C.c_int32, P(SLIP_options)]),
            "SLIP_build_dense_mpfr": (C.c_int, [P(SLIP_dense), P(P(MpfrStruct)), C.c_int32, C.c_int32,
                                               P(SLIP_options)]),
            "SLIP_solve_mpfr": (C.c_int, [P(P(MpfrStruct)), P(SLIP_sparse), P(SLIP_LU_analysis),
                                         P(SLIP_dense), P(SLIP_options)]),
            "SLIP_get_mpfr_soln": (C.c_int, [P(P(MpfrStruct)), P(P(MpqStruct)), C.c_int32, C.c_int32,
                                            P(SLIP_options)]),
            "SLIP_create_double_mat": (P(P(C.c_double)), [C.c_int32, C.c_int32]),
            "SLIP_delete_double_mat": (None, [P(P(P(C.c_double))), C.c_int32, C.c_int32]),
        }
        self.optional_missing: List[str] = []
        for name, (res, args) in sig.items():
            try:
                f = getattr(d, name)
            except AttributeError:
                self.optional_missing.append(name)
                continue
            f.restype = res
            f.argtypes = args
        d.SLIP_initialize()

    # ------------------------------------------------------------------ builders
    def default_options(self, pivot: Optional[int] = None, order: Optional[int] = None,
                        tol: Optional[float] = None):
        o = self.dll.SLIP_create_default_options()
        if not o:
            raise MemoryError("SLIP_create_default_options")
        if pivot is not None:
            o.contents.pivot = pivot
        if order is not None:
            o.contents.order = order
        if tol is not None:
            o.contents.tol = tol
        return o

    def _mpz_array(self, values: Sequence[int]):
        arr = self.dll.SLIP_create_mpz_array(len(values))
        if not arr:
            raise MemoryError("SLIP_create_mpz_array")
        for k, v in enumerate(values):
            int_to_mpz(arr[k], int(v))
        return arr

    def sparse_from_csc(self, n: int, colptr: Sequence[int], rowidx: Sequence[int],
                        values: Sequence[int]):
        """SLIP_build_sparse_ccf_mpz on Python ints."""
        nz = len(values)
        A = self.dll.SLIP_create_sparse()
        xa = self._mpz_array(values)
        p = (C.c_int32 * (n + 1))(*colptr)
        i = (C.c_int32 * nz)(*rowidx)
        rc = self.dll.SLIP_build_sparse_ccf_mpz(A, p, i, xa, n, nz)
        xp = C.pointer(xa) if False else None  # noqa: F841 (kept for clarity)
        tmp = C.cast(xa, C.POINTER(MpzStruct))
        self.dll.SLIP_delete_mpz_array(C.byref(tmp), nz)
        if rc != SLIP_OK:
            self.dll.SLIP_delete_sparse(C.byref(A))
            raise SlipError(rc, "SLIP_build_sparse_ccf_mpz")
        return A

    def sparse_from_triplets(self, n: int, I: Sequence[int], J: Sequence[int], values: Sequence[int]):
        nz = len(values)
        A = self.dll.SLIP_create_sparse()
        xa = self._mpz_array(values)
        ia = (C.c_int32 * nz)(*I)
        ja = (C.c_int32 * nz)(*J)
        rc = self.dll.SLIP_build_sparse_trip_mpz(A, ia, ja, xa, n, nz)
        tmp = C.cast(xa, C.POINTER(MpzStruct))
        self.dll.SLIP_delete_mpz_array(C.byref(tmp), nz)
        if rc != SLIP_OK:
            self.dll.SLIP_delete_sparse(C.byref(A))
            raise SlipError(rc, "SLIP_build_sparse_trip_mpz")
        return A

    def dense_from_rows(self, rows: Sequence[Sequence[int]]):
        m = len(rows)
        nrhs = len(rows[0])
        b = self.dll.SLIP_create_dense()
        mat = self.dll.SLIP_create_mpz_mat(m, nrhs)
        for r in range(m):
            for c in range(nrhs):
                int_to_mpz(mat[r][c], int(rows[r][c]))
        rc = self.dll.SLIP_build_dense_mpz(b, mat, m, nrhs)
        tmp = C.cast(mat, C.POINTER(C.POINTER(MpzStruct)))
        self.dll.SLIP_delete_mpz_mat(C.byref(tmp), m, nrhs)
        if rc != SLIP_OK:
            raise SlipError(rc, "SLIP_build_dense_mpz")
        return b

    # ------------------------------------------------------------------ path calls
    def analyze(self, A, opts, q: Optional[Sequence[int]] = None):
        n = A.contents.n
        S = self.dll.SLIP_create_LU_analysis(n + 1)
        if q is None:
            rc = self.dll.SLIP_LU_analyze(S, A, opts)
            if rc != SLIP_OK:
                raise SlipError(rc, "SLIP_LU_analyze")
        else:
            # caller-supplied column permutation (S->q is an input of SLIP_LU_factorize)
            for k in range(n):
                S.contents.q[k] = q[k]
            S.contents.q[n] = n
            S.contents.lnz = S.contents.unz = max(10 * A.contents.nz, n)
        return S

    def factorize(self, A, S, opts):
        """SLIP_LU_factorize -> (L, U, rhos, pinv) as raw C objects."""
        n = A.contents.n
        L = self.dll.SLIP_create_sparse()
        U = self.dll.SLIP_create_sparse()
        rhos = self.dll.SLIP_create_mpz_array(n)
        pinv = (C.c_int32 * n)()
        rc = self.dll.SLIP_LU_factorize(L, U, A, S, rhos, pinv, opts)
        if rc != SLIP_OK:
            self.dll.SLIP_delete_sparse(C.byref(L))
            self.dll.SLIP_delete_sparse(C.byref(U))
            tmp = C.cast(rhos, C.POINTER(MpzStruct))
            self.dll.SLIP_delete_mpz_array(C.byref(tmp), n)
            raise SlipError(rc, "SLIP_LU_factorize")
        return L, U, rhos, pinv

    def lu_solve(self, b, rhos, L, U, pinv):
        """SLIP_LU_solve -> mpq_t** (n x numRHS), in the permuted (LU) column order."""
        n, nrhs = b.contents.m, b.contents.n
        x = self.dll.SLIP_create_mpq_mat(n, nrhs)
        rc = self.dll.SLIP_LU_solve(x, b, rhos, L, U, pinv)
        if rc != SLIP_OK:
            raise SlipError(rc, "SLIP_LU_solve")
        return x

    def solve_mpq(self, A, S, b, opts):
        """SLIP_solve_mpq: factor + solve + permute + scale -> mpq_t**."""
        n, nrhs = A.contents.n, b.contents.n
        x = self.dll.SLIP_create_mpq_mat(n, nrhs)
        rc = self.dll.SLIP_solve_mpq(x, A, S, b, opts)
        if rc != SLIP_OK:
            tmp = C.cast(x, C.POINTER(C.POINTER(MpqStruct)))
            self.dll.SLIP_delete_mpq_mat(C.byref(tmp), n, nrhs)
            raise SlipError(rc, "SLIP_solve_mpq")
        return x

    # ------------------------------------------------------------------ readers
    @staticmethod
    def sparse_to_py(M) -> Tuple[List[int], List[int], List[int]]:
        m = M.contents
        p = [m.p[k] for k in range(m.n + 1)]
        nz = p[m.n]
        return p, [m.i[k] for k in range(nz)], [mpz_to_int(m.x[k]) for k in range(nz)]

    @staticmethod
    def mpz_array_to_py(arr, n: int) -> List[int]:
        return [mpz_to_int(arr[k]) for k in range(n)]

    @staticmethod
    def mpq_mat_to_py(x, n: int, nrhs: int) -> List[List[Tuple[int, int]]]:
        return [[mpq_to_pair(x[r][c]) for c in range(nrhs)] for r in range(n)]

    # ------------------------------------------------------------------ deleters
    def free_sparse(self, M):
        self.dll.SLIP_delete_sparse(C.byref(M))

    def free_dense(self, b):
        self.dll.SLIP_delete_dense(C.byref(b))

    def free_analysis(self, S):
        self.dll.SLIP_delete_LU_analysis(C.byref(S))

    def free_mpz_array(self, arr, n: int):
        tmp = C.cast(arr, C.POINTER(MpzStruct))
        self.dll.SLIP_delete_mpz_array(C.byref(tmp), n)

    def free_mpq_mat(self, x, n: int, nrhs: int):
        tmp = C.cast(x, C.POINTER(C.POINTER(MpqStruct)))
        self.dll.SLIP_delete_mpq_mat(C.byref(tmp), n, nrhs)

    def free_options(self, o):
        self.dll.SLIP_free(C.cast(o, C.c_void_p))
