"""Dev tool: larger instances of the other BASELINE config families, verified exactly (A x = b)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
import slip_lu_b200
from slip_lu_b200 import capi, synth
lib = slip_lu_b200.lib()

def run(name, sysm, order=capi.SLIP_COLAMD):
    n, cp, ri, vals, b = sysm
    o = lib.default_options(order=order)
    A = lib.sparse_from_csc(n, cp, ri, vals); B = lib.dense_from_rows(b)
    t = time.time(); S = lib.analyze(A, o); x = lib.solve_mpq(A, S, B, o); dt = time.time() - t
    ok = lib.dll.SLIP_check_solution(A, x, B)
    import ctypes as C
    st = (C.c_double * 10)(); lib.dll.SLIP_B200_last_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
    lib.dll.SLIP_B200_last_stats(st, 10)
    print(f"{name}: n={n} nrhs={len(b[0])} time {dt:.2f}s exact={ok == 0} nnzL={int(st[1])} nnzU={int(st[2])} channels={int(st[3])}", flush=True)

which = sys.argv[1:] or ["lp", "lap", "batch"]
if "lp" in which:
    for seed in range(3, 12):      # random LP bases of this size are sometimes exactly singular
        try:
            run(f"configs[3]-style LP basis (seed {seed}), 256 rhs", synth.lp_basis(10000, seed=seed, nrhs=256))
            break
        except capi.SlipError as e:
            print("seed", seed, e, flush=True)
if "lap" in which:
    run("configs[2]-style Laplacian 50x50, 64-bit", synth.laplacian_2d(50, 64, seed=4, nrhs=1))
if "batch" in which:
    t = time.time()
    for s in range(32):
        n, cp, ri, vals, b = synth.lp_basis(500, seed=100 + s, nrhs=1)
        o = lib.default_options()
        A = lib.sparse_from_csc(n, cp, ri, vals); B = lib.dense_from_rows(b)
        S = lib.analyze(A, o)
        try:
            x = lib.solve_mpq(A, S, B, o)
        except capi.SlipError:
            continue                      # exactly singular random basis
        assert lib.dll.SLIP_check_solution(A, x, B) == 0
    print(f"configs[4]-style: 32 systems n=500 in {time.time() - t:.2f}s ({(time.time() - t) / 32 * 1e3:.1f} ms each)", flush=True)
if "threads" in which:
    from slip_lu_b200.sharding import solve_batch_sharded
    systems = []
    seed = 100
    while len(systems) < 64:
        sm = synth.lp_basis(500, seed=seed, nrhs=1); seed += 1
        try:
            solve_batch_sharded(lib, [sm], 1, 0)
            systems.append(sm)
        except capi.SlipError:
            pass
    for th in (1, 4, 8, 16):
        t = time.time()
        solve_batch_sharded(lib, systems, 1, 0, threads=th)
        print(f"64 systems n=500, {th} host threads: {(time.time() - t) / 64 * 1e3:.1f} ms per system", flush=True)
