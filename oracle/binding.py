"""ctypes binding of oracle/libref_oracle.so (the CPU restatement) -- TEST INFRASTRUCTURE.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence, Tuple

from slip_lu_b200.capi import MpqStruct, MpzStruct, gmp, int_to_mpz, mpq_to_pair, mpz_to_int

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libref_oracle.so")
DEMO_EXE = os.path.join(_HERE, "_ref", "demo_example2")
REF_SO = os.path.join(_HERE, "_ref", "libslip_ref.so")


class RoCsc(C.Structure):
    _fields_ = [("n", C.c_int), ("nz", C.c_int), ("p", C.POINTER(C.c_int)),
                ("i", C.POINTER(C.c_int)), ("x", C.POINTER(MpzStruct))]


def build(force: bool = False) -> None:
    """Compile the restatement (and, where /root/reference exists, oracle/_ref)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "ref_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if os.path.isdir("/root/reference/SLIP_LU/Source") and (force or not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    # the reference's demo program linked against the product library (run by a GPU test)
    product = os.path.join(os.path.dirname(_HERE), "slip_lu_b200", "libslip_lu_b200.so")
    if os.path.isfile("/root/reference/SLIP_LU/Demo/example2.c") and os.path.exists(product) and \
            (force or not os.path.exists(DEMO_EXE) or os.path.getmtime(DEMO_EXE) < os.path.getmtime(product)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "demo"])


_dll = None


def dll() -> C.CDLL:
    global _dll
    if _dll is None:
        build()
        gmp()
        _dll = C.CDLL(ORACLE_SO)
        P = C.POINTER
        _dll.ro_factorize.restype = C.c_int
        _dll.ro_factorize.argtypes = [C.c_int, P(C.c_int), P(C.c_int), P(MpzStruct), P(C.c_int),
                                      C.c_int, C.c_double, P(RoCsc), P(RoCsc), P(MpzStruct), P(C.c_int)]
        _dll.ro_solve.restype = C.c_int
        _dll.ro_solve.argtypes = [C.c_int, C.c_int, P(RoCsc), P(RoCsc), P(MpzStruct), P(C.c_int),
                                  P(MpzStruct), P(MpqStruct)]
        _dll.ro_free_csc.argtypes = [P(RoCsc)]
        _dll.ro_set_column_limit.argtypes = [C.c_int]
        _dll.ro_digest_mpz.restype = C.c_uint64
        _dll.ro_digest_mpz.argtypes = [P(MpzStruct), C.c_int]
        _dll.ro_digest_csc.restype = C.c_uint64
        _dll.ro_digest_csc.argtypes = [C.c_int, P(C.c_int), P(C.c_int), P(MpzStruct)]
        _dll.ro_digest_mpq.restype = C.c_uint64
        _dll.ro_digest_mpq.argtypes = [P(MpqStruct), C.c_int]
    return _dll


class MpzArray:
    """An owned, initialised array of mpz_t."""

    def __init__(self, values: Sequence[int] = (), count: int = 0):
        n = len(values) if values else count
        self.n = n
        self.arr = (MpzStruct * max(n, 1))()
        g = gmp(); dll()
        for k in range(n):
            g.mpz_init(C.byref(self.arr[k]))
        for k, v in enumerate(values):
            int_to_mpz(self.arr[k], int(v))

    def to_py(self) -> List[int]:
        return [mpz_to_int(self.arr[k]) for k in range(self.n)]

    def __del__(self):
        try:
            g = gmp()
            for k in range(self.n):
                g.mpz_clear(C.byref(self.arr[k]))
        except Exception:
            pass


class OracleFactors:
    def __init__(self, n, L, U, rhos, pinv):
        self.n, self.L, self.U, self.rhos, self.pinv = n, L, U, rhos, pinv

    @staticmethod
    def _csc_py(M: RoCsc):
        p = [M.p[k] for k in range(M.n + 1)]
        return p, [M.i[k] for k in range(M.nz)], [mpz_to_int(M.x[k]) for k in range(M.nz)]

    def L_py(self):
        return self._csc_py(self.L)

    def U_py(self):
        return self._csc_py(self.U)

    def rhos_py(self):
        return self.rhos.to_py()

    def pinv_py(self):
        return list(self.pinv)

    def digests(self) -> Tuple[int, int, int]:
        d = dll()
        return (d.ro_digest_csc(self.n, self.L.p, self.L.i, self.L.x),
                d.ro_digest_csc(self.n, self.U.p, self.U.i, self.U.x),
                d.ro_digest_mpz(self.rhos.arr, self.n))

    def __del__(self):
        try:
            dll().ro_free_csc(C.byref(self.L)); dll().ro_free_csc(C.byref(self.U))
        except Exception:
            pass


def factorize(n: int, colptr, rowidx, values, q, pivot: int = 3, tol: float = 1.0,
              max_cols: int = 0) -> OracleFactors:
    """max_cols > 0: only the first max_cols columns (bench.py's bounded sample of a large job)."""
    d = dll()
    d.ro_set_column_limit(max_cols)
    Ax = MpzArray(values)
    Ap = (C.c_int * (n + 1))(*colptr)
    Ai = (C.c_int * len(rowidx))(*rowidx)
    qa = (C.c_int * n)(*list(q)[:n])
    L, U = RoCsc(), RoCsc()
    rhos = MpzArray(count=n)
    pinv = (C.c_int * n)()
    rc = d.ro_factorize(n, Ap, Ai, Ax.arr, qa, pivot, tol, C.byref(L), C.byref(U), rhos.arr, pinv)
    if rc != 0:
        raise RuntimeError(f"ro_factorize returned {rc}")
    return OracleFactors(n, L, U, rhos, pinv)


def solve(f: OracleFactors, b_rows) -> List[List[Tuple[int, int]]]:
    """x = (L D^-1 U)^-1 P b as canonical (num, den) pairs, in factor column order."""
    d = dll(); g = gmp()
    n, nrhs = f.n, len(b_rows[0])
    b = MpzArray([b_rows[r][c] for r in range(n) for c in range(nrhs)])
    xq = (MpqStruct * (n * nrhs))()
    for k in range(n * nrhs):
        g.mpq_init(C.byref(xq[k]))
    rc = d.ro_solve(n, nrhs, C.byref(f.L), C.byref(f.U), f.rhos.arr, f.pinv, b.arr, xq)
    if rc != 0:
        raise RuntimeError(f"ro_solve returned {rc}")
    out = [[mpq_to_pair(xq[r * nrhs + c]) for c in range(nrhs)] for r in range(n)]
    for k in range(n * nrhs):
        g.mpq_clear(C.byref(xq[k]))
    return out


def digest_slip_sparse(M) -> int:
    """Digest of a SLIP_sparse* produced by any library bound through capi.SlipLib."""
    m = M.contents
    return dll().ro_digest_csc(m.n, C.cast(m.p, C.POINTER(C.c_int)), C.cast(m.i, C.POINTER(C.c_int)), m.x)


def digest_mpz_array(arr, n: int) -> int:
    return dll().ro_digest_mpz(C.cast(arr, C.POINTER(MpzStruct)), n)
