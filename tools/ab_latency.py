"""Dev tool (GPU box): the latency-bound workloads under the current environment, one line each.

    [ENV=...] python tools/ab_latency.py
"""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
import slip_lu_b200  # noqa: E402
from slip_lu_b200 import capi, refmats, synth  # noqa: E402

lib = slip_lu_b200.lib()
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("SLIP_B200_")) or "default"


def solve(A, B, o):
    S = lib.analyze(A, o)
    try:
        x = lib.solve_mpq(A, S, B, o)
    except capi.SlipError:
        lib.free_analysis(S)
        return
    lib.free_mpq_mat(x, A.contents.n, B.contents.n)
    lib.free_analysis(S)


out = []
for name in ("prob159", "NSR8K", "basislib/aa01"):
    n, I, J, X, b = refmats.system(name)
    A = lib.sparse_from_triplets(n, I, J, X); B = lib.dense_from_rows(b); o = lib.default_options()
    best = 1e9
    for it in range(4):
        t = time.perf_counter(); solve(A, B, o); dt = time.perf_counter() - t
        if it:
            best = min(best, dt)
    out.append(f"{name} {best * 1e3:.1f} ms")
systems = []
for g in range(128):
    n, cp, ri, vals, b = synth.lp_basis(500, seed=5000 + g, nrhs=1)
    systems.append((lib.sparse_from_csc(n, cp, ri, vals), lib.dense_from_rows(b), lib.default_options()))
solve(*systems[0])
t = time.perf_counter()
NT = int(os.environ.get("AB_THREADS", "8"))
with ThreadPoolExecutor(max_workers=NT) as pool:
    list(pool.map(lambda s: solve(*s), systems))
out.append(f"128 x lp500 ({NT} threads) {(time.perf_counter() - t) * 1e3:.0f} ms")
n, cp, ri, vals, b = synth.lp_basis(10000, seed=4, nrhs=1)
A = lib.sparse_from_csc(n, cp, ri, vals); B = lib.dense_from_rows(b); o = lib.default_options()
best = 1e9
for it in range(3):
    t = time.perf_counter(); solve(A, B, o); dt = time.perf_counter() - t
    if it:
        best = min(best, dt)
out.append(f"lp10000 1 rhs {best * 1e3:.0f} ms")
print(f"[{tag}] " + " | ".join(out), flush=True)
