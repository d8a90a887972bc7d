"""Deterministic synthetic integer systems for the BASELINE.json configurations.

All generators return ``(n, colptr, rowidx, values, b)`` with Python-int values
(CSC, rows sorted inside each column) and ``b`` as a list of rows (n x nrhs).
They are seeded with ``random.Random`` so the same input is produced here and
on the GPU box; none of them reads /root/reference.
"""
from __future__ import annotations

import random
from typing import List, Tuple

Csc = Tuple[int, List[int], List[int], List[int], List[List[int]]]


def _val(rng: random.Random, bits: int) -> int:
    v = rng.randrange(1, 1 << (bits - 1)) if bits > 1 else 1
    return v if rng.random() < 0.5 else -v


def _rhs(rng: random.Random, n: int, nrhs: int, bits: int) -> List[List[int]]:
    return [[_val(rng, bits) for _ in range(nrhs)] for _ in range(n)]


def random_sparse(n: int, nnz_per_col: int = 10, bits: int = 32, seed: int = 0,
                  nrhs: int = 1, rhs_bits: int = 32) -> Csc:
    """BASELINE configs[1]: random sparse integer matrix, diagonal always present
    (structurally nonsingular), `nnz_per_col` entries per column, `bits`-bit entries."""
    rng = random.Random(seed)
    k = min(nnz_per_col, n)
    colptr, rowidx, values = [0], [], []
    for j in range(n):
        rows = {j}
        while len(rows) < k:
            rows.add(rng.randrange(n))
        for i in sorted(rows):
            rowidx.append(i)
            values.append(_val(rng, bits))
        colptr.append(len(rowidx))
    return n, colptr, rowidx, values, _rhs(rng, n, nrhs, rhs_bits)


def laplacian_2d(m: int, bits: int = 64, seed: int = 0, nrhs: int = 1, rhs_bits: int = 64) -> Csc:
    """BASELINE configs[2]: 5-point 2D-Laplacian *pattern* on an m x m grid (n = m*m)
    with random `bits`-bit integer entries."""
    rng = random.Random(seed)
    n = m * m
    colptr, rowidx, values = [0], [], []
    for y in range(m):
        for x in range(m):
            rows = []
            for dx, dy in ((0, -1), (-1, 0), (0, 0), (1, 0), (0, 1)):
                xx, yy = x + dx, y + dy
                if 0 <= xx < m and 0 <= yy < m:
                    rows.append(yy * m + xx)
            for i in sorted(rows):
                rowidx.append(i)
                values.append(_val(rng, bits))
            colptr.append(len(rowidx))
    return n, colptr, rowidx, values, _rhs(rng, n, nrhs, rhs_bits)


def lp_basis(n: int, seed: int = 0, nrhs: int = 1) -> Csc:
    """BASELINE configs[4]: LP-basis style matrix: mostly slack (unit) columns, the rest
    short structural columns with small integer coefficients (like BasisLIB bases)."""
    rng = random.Random(seed)
    colptr, rowidx, values = [0], [], []
    for j in range(n):
        if rng.random() < 0.55:
            rows = {j: rng.choice((1, -1))}
        else:
            rows = {j: rng.choice((1, -1, 2, -2, 3, 5, -7))}
            for _ in range(rng.randrange(1, 6)):
                rows[rng.randrange(n)] = rng.choice((1, -1, 1, -1, 2, -2, 3, -4, 6, 12, -25))
        for i in sorted(rows):
            rowidx.append(i)
            values.append(rows[i])
        colptr.append(len(rowidx))
    return n, colptr, rowidx, values, _rhs(rng, n, nrhs, 8)


def decimal_scaled(n: int, nnz_per_col: int = 6, digits: int = 6, seed: int = 0,
                   nrhs: int = 1) -> Tuple[Csc, List[float]]:
    """BASELINE configs[3]: a matrix given as doubles with `digits` decimal digits.  Returns
    the integer system obtained by scaling by 10**digits (what the double builders of the
    interface produce up to the common gcd) together with the double values."""
    rng = random.Random(seed)
    k = min(nnz_per_col, n)
    colptr, rowidx, values, dvals = [0], [], [], []
    for j in range(n):
        rows = {j}
        while len(rows) < k:
            rows.add(rng.randrange(n))
        for i in sorted(rows):
            iv = rng.randrange(1, 10 ** digits) * rng.choice((1, -1))
            rowidx.append(i)
            values.append(iv)
            dvals.append(iv / 10 ** digits)
        colptr.append(len(rowidx))
    return (n, colptr, rowidx, values, _rhs(rng, n, nrhs, 20)), dvals


def dense_head(n: int, head: int = 5, bits: int = 12, seed: int = 0, nrhs: int = 1) -> Csc:
    """A few long columns at little total cost: column 0 is dense, columns 1..head-1 have an entry in
    row 0 (so they fill in completely and are eliminated with columns of n, n-1, ... rows), the rest
    is diagonal.  Used to put elimination steps of a chosen length through the chunked kernels."""
    rng = random.Random(seed)
    I, J, V = [], [], []
    for r in range(n):
        I.append(r); J.append(0); V.append(_val(rng, bits))
    for k in range(1, n):
        if k < head:
            I.append(0); J.append(k); V.append(_val(rng, bits))
        I.append(k); J.append(k); V.append(_val(rng, bits))
    cp, ri, vals = triplets_to_csc(n, I, J, V)
    return n, cp, ri, vals, _rhs(rng, n, nrhs, bits)


def read_triplet_file(path: str):
    """Reader for the ExampleMats text format ("m n nz" then "i j value" lines, 0- or
    1-based decided from the first triplet like the reference demo reader,
    reference: SLIP_LU/Demo/demos.c:245-340).  Returns (n, I, J, values)."""
    with open(path) as f:
        toks = f.read().split()
    m, n, nz = int(toks[0]), int(toks[1]), int(toks[2])
    I, J, V = [], [], []
    t = 3
    for _ in range(nz):
        I.append(int(toks[t])); J.append(int(toks[t + 1])); V.append(int(toks[t + 2]))
        t += 3
    dec = 0 if min(I[0], J[0]) == 0 else 1
    return n, [i - dec for i in I], [j - dec for j in J], V


def read_dense_file(path: str) -> List[List[int]]:
    with open(path) as f:
        toks = f.read().split()
    m, n = int(toks[0]), int(toks[1])
    vals = [int(t) for t in toks[2:2 + m * n]]
    return [vals[r * n:(r + 1) * n] for r in range(m)]


def triplets_to_csc(n: int, I, J, V):
    """Column-bucket the triplets in input order (same order slip_trip_to_mat produces,
    reference: SLIP_LU/Source/slip_trip_to_mat.c)."""
    cnt = [0] * (n + 1)
    for j in J:
        cnt[j + 1] += 1
    for j in range(n):
        cnt[j + 1] += cnt[j]
    nxt = cnt[:-1].copy()
    ri = [0] * len(V)
    vv = [0] * len(V)
    for i, j, v in zip(I, J, V):
        p = nxt[j]; nxt[j] += 1
        ri[p] = i; vv[p] = v
    return cnt, ri, vv
