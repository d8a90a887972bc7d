"""Dev tool (GPU box): where the channel-count restarts happen (SLIP_B200_TIMING=1 prints them)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SLIP_B200_TIMING"] = "1"
import __graft_entry__ as entry
entry.build()
import slip_lu_b200
from slip_lu_b200 import capi, refmats, synth
lib = slip_lu_b200.lib()
def run(name, A, B):
    o = lib.default_options()
    for it in range(2):
        S = lib.analyze(A, o)
        t = time.perf_counter(); x = lib.solve_mpq(A, S, B, o); dt = time.perf_counter() - t
        lib.free_mpq_mat(x, A.contents.n, B.contents.n); lib.free_analysis(S)
    print(f"== {name}: {dt:.3f}s", file=sys.stderr, flush=True)
for name in sys.argv[1:]:
    if name.startswith("lp"):
        n, cp, ri, vals, b = synth.lp_basis(int(name[2:]), seed=4 if name == "lp10000" else 5001, nrhs=1)
        run(name, lib.sparse_from_csc(n, cp, ri, vals), lib.dense_from_rows(b))
    else:
        n, I, J, X, b = refmats.system(name)
        run(name, lib.sparse_from_triplets(n, I, J, X), lib.dense_from_rows(b))
