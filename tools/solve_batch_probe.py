"""Dev tool (GPU box): one SLIP_solve_mpq with several right-hand sides; prints the number of
k_trisolve launches (the last `batches` of them are the forward substitutions of the batches) and
the algorithmic bytes of L, to set beside an ncu capture of one of those launches:

    python tools/solve_batch_probe.py 1200 16            # prints launches=T batches=B L_bytes=...
    ncu --set full --clock-control none -k regex:k_trisolve -s <T-B> -c 1 -o out python tools/solve_batch_probe.py 1200 16
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SLIP_B200_LOOKAHEAD", "0")
import __graft_entry__ as entry  # noqa: E402

entry.build()
import slip_lu_b200  # noqa: E402
from slip_lu_b200 import synth  # noqa: E402
import bench  # noqa: E402

lib = slip_lu_b200.lib()
lib.dll.SLIP_B200_last_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
n, nrhs = int(sys.argv[1]), int(sys.argv[2])
n, cp, ri, vals, b = synth.random_sparse(n, 10, 32, seed=3, nrhs=nrhs)
A = lib.sparse_from_csc(n, cp, ri, vals); B = lib.dense_from_rows(b); o = lib.default_options()
S = lib.analyze(A, o)
lib.dll.slipcu_reset_counters()
x = lib.solve_mpq(A, S, B, o)
c = bench.counters(lib)
st = bench.last_stats(lib)
batch = min(nrhs, (nrhs + 3) // 4) if nrhs >= 8 else nrhs
batches = (nrhs + batch - 1) // batch
print(f"trisolve_launches={c.trisolve_launches} batches={batches} rhs_per_batch={batch} channels={int(st[3])} nnz_L={int(st[1])} "
      f"L_bytes={st[1] * st[3] * 4:.0f} exact={lib.dll.SLIP_check_solution(A, x, B) == 0}", flush=True)
