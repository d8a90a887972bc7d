"""Loader for packed integer test systems (tests/golden/mats/refmats.npz).

The file holds the reference distribution's own example systems (ExampleMats/NSR8K, prob159 and a
slice of the BasisLIB LP bases: the inputs of its Demo/SLIPLU.c and Demo/example2.c) in triplet
form, packed by tests/golden/make_refmats.py so that bench.py's head-to-head legs and the GPU parity
tests can read them where /root/reference does not exist.  Nothing here computes.
"""
from __future__ import annotations

import math
import os
import sys
from typing import Dict, List, Tuple

if hasattr(sys, "set_int_max_str_digits"):
    sys.set_int_max_str_digits(0)      # the digests below hash decimal strings of integers of tens of kbit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACKED = os.path.join(ROOT, "tests", "golden", "mats", "refmats.npz")
RECORDS = os.path.join(ROOT, "tests", "golden", "refmats.json")

_I64 = 1 << 63


def pack(arrays: Dict, name: str, n: int, I, J, X, b) -> None:
    """Adds one system to the dict that np.savez_compressed writes (values as int64 where they fit,
    decimal strings otherwise)."""
    import numpy as np
    key = name.replace("/", "__")
    arrays[key + ".n"] = np.array([n, len(b[0])], dtype=np.int64)
    arrays[key + ".I"] = np.array(I, dtype=np.int32)
    arrays[key + ".J"] = np.array(J, dtype=np.int32)
    flat_b = [v for row in b for v in row]
    for tag, vals in (("X", X), ("b", flat_b)):
        if all(-_I64 <= v < _I64 for v in vals):
            arrays[key + "." + tag] = np.array(vals, dtype=np.int64)
        else:
            arrays[key + "." + tag + "s"] = np.array([str(v) for v in vals])


_cache = None


def _file():
    global _cache
    if _cache is None:
        import numpy as np
        if not os.path.exists(PACKED):
            raise FileNotFoundError(f"{PACKED} is missing (python tests/golden/make_refmats.py in the build container)")
        _cache = np.load(PACKED)
    return _cache


def names() -> List[str]:
    return sorted(k[:-2].replace("__", "/") for k in _file().files if k.endswith(".n"))


def load(name: str) -> Tuple[int, List[int], List[int], List[int], List[List[int]]]:
    """(n, I, J, X, b) with Python ints; triplets in the order of the original file."""
    f = _file()
    key = name.replace("/", "__")
    n, nrhs = (int(v) for v in f[key + ".n"])
    I = f[key + ".I"].tolist()
    J = f[key + ".J"].tolist()
    X = f[key + ".X"].tolist() if key + ".X" in f.files else [int(s) for s in f[key + ".Xs"].tolist()]
    fb = f[key + ".b"].tolist() if key + ".b" in f.files else [int(s) for s in f[key + ".bs"].tolist()]
    return n, I, J, X, [fb[r * nrhs:(r + 1) * nrhs] for r in range(n)]


def records() -> Dict[str, dict]:
    """What the unmodified reference computed for each packed system and each synthetic system below
    (digests, sizes, its seconds in the build container)."""
    import json
    out = {}
    for path in (RECORDS, SYNTH_RECORDS):
        if os.path.exists(path):
            with open(path) as fh:
                out.update({r["name"]: r for r in json.load(fh)["records"]})
    return out


# Synthetic systems of the BASELINE config families that are pinned by reference digests
# (tests/golden/make_synth_records.py); regenerated from their seeds, nothing stored.
SYNTH_RECORDS = os.path.join(ROOT, "tests", "golden", "synth_records.json")
BENCH_SEED = 20261018
SYNTH = {
    "synth/rand240": dict(family="configs[1] generator (random sparse, 10 nnz/col, 32-bit)", gen="random", n=240, seed=BENCH_SEED),
    "synth/rand600": dict(family="configs[1] generator (random sparse, 10 nnz/col, 32-bit)", gen="random", n=600, seed=BENCH_SEED),
    "synth/lap24": dict(family="configs[2] generator (2D Laplacian pattern, 64-bit)", gen="laplacian", m=24, seed=7),
    "synth/lap32": dict(family="configs[2] generator (2D Laplacian pattern, 64-bit)", gen="laplacian", m=32, seed=7),
    "synth/lap40": dict(family="configs[2] generator (2D Laplacian pattern, 64-bit)", gen="laplacian", m=40, seed=7),
}


def synth_system(name: str):
    """(n, I, J, X, b) of a synthetic system, as triplets in CSC order."""
    from . import synth
    d = SYNTH[name]
    if d["gen"] == "random":
        n, cp, ri, vals, b = synth.random_sparse(d["n"], 10, 32, seed=d["seed"], nrhs=1)
    else:
        n, cp, ri, vals, b = synth.laplacian_2d(d["m"], 64, seed=d["seed"], nrhs=1)
    J = [j for j in range(n) for _ in range(cp[j], cp[j + 1])]
    return n, list(ri), J, list(vals), b


def system(name: str):
    """A packed reference system or a synthetic one, by name."""
    return synth_system(name) if name in SYNTH else load(name)


def hadamard_bits(n: int, J, X) -> float:
    """log2 of the column-norm Hadamard bound of det A."""
    col = [0] * n
    for j, x in zip(J, X):
        col[j] += x * x
    return sum(0.5 * math.log2(c) for c in col if c > 0)


def digest_ints(v) -> str:
    import hashlib
    h = hashlib.sha256()
    h.update(",".join(str(int(t)) for t in v).encode())
    return h.hexdigest()[:16]


def digest_mpq_mat(lib, x, n: int, nrhs: int) -> int:
    """Order-dependent digest of an mpq_t** result (same function as tests/cases.digest_pairs)."""
    import hashlib
    from .capi import mpq_to_pair
    h = hashlib.sha256()
    for r in range(n):
        for c in range(nrhs):
            a, d = mpq_to_pair(x[r][c])
            h.update(f"{a}/{d};".encode())
    return int.from_bytes(h.digest()[:8], "little")
