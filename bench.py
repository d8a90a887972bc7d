#!/usr/bin/env python
"""bench.py -- the hot path of SLIP LU (exact sparse factor + solve) on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 2000]
                    [--extras all|none|h2h,laplacian,lu,sharded,intmul]

Headline workload (BASELINE.json configs[1]): synthetic random sparse integer matrix, n = 2000,
10 nonzeros per column, 32-bit entries, COLAMD column order, 1 right-hand side.  One *step* is one
complete exact solve of that system through the reference-facing C interface:
SLIP_LU_analyze + SLIP_solve_mpq (the call sequence of the reference's Demo/example2.c), host mpz_t
inputs in, canonical host mpq_t solution out.

metric  limb_mul_ops_per_s: schoolbook-EQUIVALENT 32-bit limb multiplications of the REF
        elimination per second.  The work W of a system is a property of the input, computed by
        the same formula for both arms: every REF entry update (one per entry of L below the pivot,
        per elimination step j of a column) counts 3 * w_j^2 limb products (two w_j-limb products
        and one exact division), w_j = ceil(bits_j / 32), bits_j = Hadamard prefix bound of pivot j.
        It is a work model, not arithmetic the GPU performs (the GPU works on residues, O(w) per
        update): read it beside factor_solve_seconds and the roofline, never alone.
value   W / device time (CUDA events, A resident in HBM -> solution numerators reconstructed in
        HBM), all ranks.   e2e: W / wall time of the C-interface call with host buffers.

Extra keys (what the other BASELINE configs and the judge's questions need, measured in the same run):
  head_to_head   systems BOTH arms run in full at identical configuration: the reference's own
                 ExampleMats (NSR8K n = 5387, prob159), the n = 240 sample of configs[1] that the
                 reference arm times, a 24 x 24 Laplacian (configs[2] family).  GPU seconds through
                 SLIP_LU_analyze + SLIP_solve_mpq, CPU seconds of the unmodified reference on the
                 box's host (live, same call sequence), ratio, channels carried vs needed, parity
                 against the reference's recorded digests (x and the row permutation).
  laplacian      configs[2] family at the largest grid the default run affords, with its own
                 k_trisolve roofline (n = 20 000 with 64-bit entries does not fit any machine:
                 pivots of ~1.3 Mbit).
  lu_path        SLIP_LU_factorize + SLIP_LU_solve at the headline size (L and U returned as mpz_t).
  sharded        the two workloads that shard (configs[3], configs[4]), STRONG scaling: a fixed
                 job split over the N ranks through slip_lu_b200.sharding, no data-path collective.
  roofline.int_mul   k_trisolve's multiply rate against the device's measured IMAD peaks.

--impl reference times the UNMODIFIED reference (oracle/_ref/libslip_ref.so, built from
/root/reference by oracle/Makefile; the oracle port if that build is absent) on a bounded SAMPLE of
the workload: the same generator at n = --ref-n (default 240), because the n = 2000 system would
take the CPU days.  Its config says so (same_config: false); the measured same-configuration
CPU/GPU comparisons are the head_to_head entries of the b200 arm.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018


class Counters(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("trisolve_launches", C.c_uint64), ("trisolve_ms", C.c_double),
                ("trisolve_bytes", C.c_double), ("trisolve_modmul", C.c_double), ("recon_ms", C.c_double),
                ("recon_mac", C.c_double), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
                ("device_ms", C.c_double), ("other_ms", C.c_double), ("trisolve_union_ms", C.c_double)]


def workload(n, seed, rhs_seed=None):
    """configs[1]: random sparse integer matrix; rhs_seed gives a rank its own right-hand side for the
    same matrix, so that every GPU of a weak-scaling run has the same amount of work."""
    import random
    from slip_lu_b200 import synth
    n, cp, ri, vals, b = synth.random_sparse(n, 10, 32, seed=seed, nrhs=1)
    if rhs_seed:
        rng = random.Random(rhs_seed)
        b = [[rng.getrandbits(32) - (1 << 31) or 1] for _ in range(n)]
    return n, cp, ri, vals, b


def work_model(n, cp, vals, q, Lp, Up, Ui):
    """Schoolbook-equivalent limb multiplications of a factorization with factors (L, U) in final
    row numbering; same formula as slip_factorize.c (work_limbmul)."""
    colbits = []
    for j in range(n):
        ss = sum(v * v for v in vals[cp[j]:cp[j + 1]])
        colbits.append(0.5 * math.log2(ss) if ss > 0 else 0.0)
    cum, acc = [], 0.0
    for k in range(n):
        acc += colbits[q[k]]
        cum.append(acc)
    updates = 0.0
    limbmul = 0.0
    for k in range(n):
        for m in range(Up[k], Up[k + 1] - 1):          # entries above the diagonal
            j = Ui[m]
            ln = (Lp[j + 1] - Lp[j]) - 1
            w = math.ceil(cum[j] / 32.0)
            updates += ln
            limbmul += 3.0 * ln * w * w
    return updates, limbmul


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=1000):
        self.rows, self.proc, self.gpu, self.period_ms = [], None, gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = max([int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()] or [0])
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def headline_config(n, sample_n=None):
    base = (f"synthetic random sparse integer matrix n={n}, 10 nnz/col, 32-bit entries, COLAMD order, 1 RHS, "
            "SLIP_TOL_SMALLEST pivoting (defaults)")
    if sample_n is None:
        return {"workload": "BASELINE configs[1]: " + base,
                "per_gpu": "one independent system per GPU (the same matrix, a different right-hand side per rank: equal work per GPU), no data-path collective",
                "l2": "factor data streamed per column (>> 126 MB L2; ~10 GB of L residues at n=2000)"}
    return {"workload": f"bounded SAMPLE of BASELINE configs[1] (n={n}): the same generator at n={sample_n} -- "
                        f"synthetic random sparse integer matrix n={sample_n}, 10 nnz/col, 32-bit entries, COLAMD order, 1 RHS, "
                        "SLIP_TOL_SMALLEST pivoting (defaults)",
            "same_config": False,
            "sample_of": f"configs[1] n={n}",
            "why": f"the unmodified reference needs days of CPU for n={n} (pivots of ~64 kbit, 8e8 REF updates); "
                   "the metric is a schoolbook work model whose rate grows with operand size, so this arm's value is "
                   "NOT comparable with the b200 arm's n=2000 value; measured same-configuration comparisons: "
                   "the b200 arm's head_to_head entries"}


# ----------------------------------------------------------------------------------------------
# reference arm
# ----------------------------------------------------------------------------------------------
def reference_lib():
    from slip_lu_b200 import capi
    from oracle import binding as ob
    return capi.SlipLib(ob.REF_SO) if os.path.exists(ob.REF_SO) else None


def run_reference(args, rank):
    """The reference arm (and the cpu_baseline of the GPU arm): host cores only."""
    from oracle import binding as ob
    n, cp, ri, vals, b = workload(args.ref_n, args.seed)
    ref = reference_lib()
    kind = "reference" if ref is not None else "port"
    times = []
    W = None
    if kind == "reference":
        o = ref.default_options()
        A = ref.sparse_from_csc(n, cp, ri, vals)
        B = ref.dense_from_rows(b)
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            S = ref.analyze(A, o)
            L, U, rhos, pinv = ref.factorize(A, S, o)           # = SLIP_solve_mpq's own sequence,
            x = ref.lu_solve(B, rhos, L, U, pinv)              #   kept apart to read L and U for W
            ref.dll.SLIP_permute_x(x, n, 1, S)
            ref.dll.SLIP_scale_x(x, A, B)
            dt = time.perf_counter() - t0
            if W is None:
                q = [S.contents.q[k] for k in range(n)]
                Lp = [L.contents.p[k] for k in range(n + 1)]
                Up = [U.contents.p[k] for k in range(n + 1)]
                Ui = [U.contents.i[k] for k in range(Up[n])]
                W = work_model(n, cp, vals, q, Lp, Up, Ui)
            if it >= args.warmup:
                times.append(dt)
            ref.free_mpq_mat(x, n, 1); ref.free_sparse(L); ref.free_sparse(U)
            ref.free_mpz_array(rhos, n); ref.free_analysis(S)
    else:
        q = sorted(range(n), key=lambda j: (cp[j + 1] - cp[j], j))
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            f = ob.factorize(n, cp, ri, vals, q)
            ob.solve(f, b)
            dt = time.perf_counter() - t0
            if W is None:
                Lp, _, _ = f.L_py(); Up, Ui, _ = f.U_py()
                W = work_model(n, cp, vals, q, Lp, Up, Ui)
            if it >= args.warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    return {"value": W[1] / sec, "unit": "limb-mul/s", "cores": 1, "kind": kind, "seconds": sec,
            "updates": W[0], "limb_mul": W[1],
            "sample": f"same generator at n={args.ref_n} (10 nnz/col, 32-bit, COLAMD, 1 RHS): the full "
                      f"SLIP_LU_analyze + factorize + solve of the {'unmodified reference' if kind == 'reference' else 'oracle port'}, "
                      f"single thread (the reference is single-threaded); n={args.n} would take the CPU days"}


# ----------------------------------------------------------------------------------------------
# helpers of the b200 arm
# ----------------------------------------------------------------------------------------------
def counters(lib):
    c = Counters()
    lib.dll.slipcu_get_counters(C.byref(c))
    return c


def last_stats(lib):
    st = (C.c_double * 13)()
    lib.dll.SLIP_B200_last_stats(st, 13)
    return list(st)


def roofline_of(c, peak, peak_src, t_dev=None):
    # launches on the lookahead streams overlap: the kernel's time is the union of its launch intervals
    tri_ms = c.trisolve_union_ms if c.trisolve_union_ms > 0 else c.trisolve_ms
    achieved = (c.trisolve_bytes / 1e9) / (tri_ms / 1e3) if tri_ms > 0 else 0.0
    traffic, ratio, cap_name = None, None, None
    for name in ("r02_trisolve_ncu_full_sorted.json", "r02_trisolve_ncu_full.json"):
        try:   # DRAM bytes / algorithmic bytes of one `ncu --set full` capture of this kernel
            cap = json.load(open(os.path.join(ROOT, "profiles", name)))
            ratio = cap["k_trisolve"][0]["traffic_over_algorithmic"] if "k_trisolve" in cap else cap["traffic_over_algorithmic"]
            traffic = ratio * c.trisolve_bytes / max(1, c.trisolve_launches)
            cap_name = name
            break
        except Exception:
            continue
    return {"bound": "hbm", "kernel": "k_trisolve", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if peak else None, "peak_source": peak_src,
            "launches": int(c.trisolve_launches), "traffic": traffic,
            "traffic_source": ("algorithmic bytes x the DRAM/algorithmic ratio (%.4f) of profiles/%s (one ncu --set full capture of this kernel)" % (ratio, cap_name)) if cap_name else None,
            "algorithmic_bytes_per_launch": c.trisolve_bytes / max(1, c.trisolve_launches),
            "modmul_per_s": c.trisolve_modmul / (tri_ms / 1e3) if tri_ms > 0 else None,
            "kernel_ms": tri_ms, "kernel_ms_sum_of_launches": c.trisolve_ms,
            "timing": "CUDA events around every launch on its own stream; achieved = all algorithmic bytes / length of the union of the launch intervals (launches of the lookahead streams overlap)",
            "kernel_share_of_step": (tri_ms / 1e3) / t_dev if t_dev else None}


def exact_check_integers(lib, system, x):
    """A x = b, exactly, over the integers: with D = lcm of the denominators of x and N = D x,
    sum_j A_ij N_j == D b_i for every row.  Same statement as SLIP_check_solution, which does it in
    rational arithmetic (a gcd per operation: minutes at the sizes of the Laplacian block)."""
    from slip_lu_b200 import capi
    n, I, J, X, b = system
    nrhs = len(b[0])
    for c in range(nrhs):
        pairs = [capi.mpq_to_pair(x[r][c]) for r in range(n)]
        D = max((d for _, d in pairs), default=1)
        for _, d in pairs:
            if D % d:
                D = D * d // math.gcd(D, d)
        N = [a * (D // d) for a, d in pairs]
        acc = [0] * n
        for i, j, v in zip(I, J, X):
            acc[i] += v * N[j]
        if any(acc[i] != D * b[i][c] for i in range(n)):
            return False
    return True


def gpu_solve_system(lib, name, system, rec, repeat=3, profile=False):
    """analyze + solve_mpq of a named system (triplets); best wall time of `repeat` runs after one
    warm-up (the 10-100 ms systems vary by tens of per cent from run to run on a shared box), parity of x and of the row permutation against the reference's recorded digests."""
    from slip_lu_b200 import refmats
    n, I, J, X, b = system
    A = lib.sparse_from_triplets(n, I, J, X)
    B = lib.dense_from_rows(b)
    o = lib.default_options()
    nrhs = len(b[0])
    best, out = None, {}
    for it in range(repeat + 1):
        lib.dll.slipcu_reset_counters()
        if profile:
            lib.dll.slipcu_set_profiling(1)
        t0 = time.perf_counter()
        S = lib.analyze(A, o)
        x = lib.solve_mpq(A, S, B, o)
        dt = time.perf_counter() - t0
        if profile:
            lib.dll.slipcu_set_profiling(0)
        if it > 0 and (best is None or dt < best):
            best = dt
            c = counters(lib)
            st = last_stats(lib)
            out = {"gpu_e2e_s": dt, "gpu_device_s": c.device_ms / 1e3, "launches": int(c.launches),
                   "channels": int(st[3]), "channels_hadamard": int(st[10]), "restarts": int(st[11]),
                   "nnz_L": int(st[1]), "nnz_U": int(st[2]), "_counters": c}
            if rec is not None:
                pv = (C.c_int32 * n)()
                lib.dll.SLIP_B200_last_pinv(pv, n)
                out["parity"] = {"x": str(refmats.digest_mpq_mat(lib, x, n, nrhs)) == rec["digests"]["x_solve_mpq"],
                                 "pinv": refmats.digest_ints(list(pv)) == rec["digests"]["pinv"]}
            else:
                out["parity"] = {"A x = b exact (integer arithmetic: sum_j A_ij N_j == D b_i, D = lcm of the denominators)":
                                 exact_check_integers(lib, system, x)}
        lib.free_mpq_mat(x, n, nrhs); lib.free_analysis(S)
    lib.free_dense(B); lib.free_sparse(A); lib.free_options(o)
    return out


def cpu_solve_system(ref, system):
    """The unmodified reference on the same system: SLIP_LU_analyze + SLIP_solve_mpq, one run."""
    n, I, J, X, b = system
    A = ref.sparse_from_triplets(n, I, J, X)
    B = ref.dense_from_rows(b)
    o = ref.default_options()
    t0 = time.perf_counter()
    S = ref.analyze(A, o)
    x = ref.solve_mpq(A, S, B, o)
    dt = time.perf_counter() - t0
    ref.free_mpq_mat(x, n, len(b[0])); ref.free_analysis(S); ref.free_dense(B); ref.free_sparse(A); ref.free_options(o)
    return dt


def cpu_leg_subprocess(argv, timeout=900):
    """Every reference (CPU) leg runs in its OWN process: the reference's SLIP_initialize installs its
    allocation-tracking hooks into the process-wide libgmp (mp_set_memory_functions), and those hooks
    are not thread-safe, while the product's host layer calls GMP from OpenMP threads -- the two
    libraries must not share a process once the product is computing."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"error": (r.stderr or "no output")[-300:]}
    return json.loads(lines[-1])


H2H = ["synth/rand240", "NSR8K", "prob159", "synth/lap24"]
# GPU leg live, CPU seconds as recorded with the reference's digests in the build container (minutes
# of CPU each: not repeated in every bench run); the ratio of these rows is labelled accordingly
H2H_RECORDED = ["synth/rand600", "synth/lap32", "synth/lap40", "basislib/gen2", "basislib/rat7a"]


def head_to_head(lib, with_cpu):
    from slip_lu_b200 import refmats
    recs = refmats.records()
    out = []
    for name in H2H + H2H_RECORDED:
        try:
            system = refmats.system(name)
        except Exception as e:                                   # packed file missing
            out.append({"workload": name, "error": str(e)})
            continue
        rec = recs.get(name)
        g = gpu_solve_system(lib, name, system, rec)
        g.pop("_counters", None)
        row = {"workload": name, "n": system[0], "call": "SLIP_LU_analyze + SLIP_solve_mpq, default options"}
        if rec:
            row.update({"family": rec["family"], "det_bits": rec["det_bits"], "hadamard_bits": round(rec["hadamard_bits"]),
                        "channels_needed": math.ceil((rec["det_bits"] + 2) / 30.99),
                        "cpu_s_build_container": rec["ref_seconds"]["solve_mpq"]})
        row.update(g)
        if name in H2H_RECORDED:
            if rec:
                row["ratio_vs_build_container_cpu"] = rec["ref_seconds"]["solve_mpq"] / row["gpu_e2e_s"]
                row["cpu_kind"] = "unmodified reference, 1 thread, BUILD CONTAINER (recorded with its digests; not measured on this box)"
        elif with_cpu:
            leg = cpu_leg_subprocess(["--cpu-leg", name])
            if "cpu_s" in leg:
                row["cpu_s"] = leg["cpu_s"]
                row["cpu_kind"] = leg["kind"]
                row["ratio"] = row["cpu_s"] / row["gpu_e2e_s"]
            else:
                row["cpu_error"] = leg.get("error")
        out.append(row)
    return out


def laplacian_block(lib, m, peak, peak_src):
    from slip_lu_b200 import synth
    n, cp, ri, vals, b = synth.laplacian_2d(m, 64, seed=7, nrhs=1)
    J = [j for j in range(n) for _ in range(cp[j], cp[j + 1])]
    g = gpu_solve_system(lib, f"lap{m}", (n, list(ri), J, list(vals), b), None, repeat=1, profile=True)
    c = g.pop("_counters")
    return {"workload": f"BASELINE configs[2] family: 2D-Laplacian pattern on a {m} x {m} grid (n={n}), 64-bit entries, COLAMD, 1 RHS",
            "why_not_n_20000": "REF pivots grow by ~66 bits per column: ~1.3 Mbit at n=20 000, several 100 GB of factors (DESIGN.md section 7)",
            "factor_solve_seconds": g["gpu_e2e_s"], "device_seconds": g["gpu_device_s"], "channels": g["channels"],
            "nnz_L": g["nnz_L"], "nnz_U": g["nnz_U"], "exact": g["parity"], "launches": g["launches"],
            "roofline": roofline_of(c, peak, peak_src, g["gpu_device_s"])}


def lu_path_block(lib, A, B, o, n, system):
    """SLIP_LU_factorize (L, U, rhos as host mpz_t) + SLIP_LU_solve at the headline size."""
    S = lib.analyze(A, o)
    t0 = time.perf_counter()
    L, U, rhos, pinv = lib.factorize(A, S, o)
    t1 = time.perf_counter()
    x = lib.lu_solve(B, rhos, L, U, pinv)
    t2 = time.perf_counter()
    lib.dll.SLIP_permute_x(x, n, 1, S)
    lib.dll.SLIP_scale_x(x, A, B)
    ok = exact_check_integers(lib, system, x)
    out = {"call": "SLIP_LU_analyze + SLIP_LU_factorize + SLIP_LU_solve (L, U, rhos returned as host mpz_t)",
           "factorize_with_LU_seconds": t2 - t0, "factorize_seconds": t1 - t0, "lu_solve_seconds": t2 - t1,
           "nnz_L": int(L.contents.nz), "nnz_U": int(U.contents.nz), "exact": ok,
           "note": "one cold run; every entry of L and U is reconstructed positionally (k_garner_flow + k_limbs) and copied to the host"}
    lib.free_mpq_mat(x, n, 1); lib.free_sparse(L); lib.free_sparse(U); lib.free_mpz_array(rhos, n); lib.free_analysis(S)
    return out


# ----------------------------------------------------------------------------------------------
# sharded workloads (strong scaling): configs[4] batch of independent systems, configs[3] multi-RHS
# ----------------------------------------------------------------------------------------------
def sharded_block(lib, args, rank, world, barrier, reduce_max, reduce_sum):
    import numpy as np
    from slip_lu_b200 import capi, synth
    from slip_lu_b200.sharding import shard_range
    from concurrent.futures import ThreadPoolExecutor
    out = {"scaling": "strong", "ranks": world}
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    # (under torchrun every rank is already pinned to its own share of the cores, see main)
    threads = max(1, min(8, cores if world > 1 else cores))

    # (i) configs[4]: batch of independent LP-basis systems
    nsys, nb = args.batch_systems, args.batch_n
    lo, hi = shard_range(nsys, world, rank)
    built = []
    for g in range(lo, hi):
        n, cp, ri, vals, b = synth.lp_basis(nb, seed=5000 + g, nrhs=1)
        built.append((lib.sparse_from_csc(n, cp, ri, vals), lib.dense_from_rows(b)))

    def one(ab):
        A, B = ab
        o = lib.default_options()
        try:
            S = lib.analyze(A, o)
            try:
                x = lib.solve_mpq(A, S, B, o)
            except capi.SlipError as e:
                lib.free_analysis(S)
                return ("singular" if e.code == capi.SLIP_SINGULAR else f"error {e.code}", None)
            lib.free_analysis(S)
            return ("ok", x)
        finally:
            lib.free_options(o)

    if built:
        one(built[0])                                            # warm: tables, pool
    barrier()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as pool:
        res = list(pool.map(one, built))
    dt = time.perf_counter() - t0
    t_batch = reduce_max(dt)
    ok = sing = bad = 0
    for (A, B), (status, x) in zip(built, res):
        if status == "ok":
            if lib.dll.SLIP_check_solution(A, x, B) == 0:
                ok += 1
            else:
                bad += 1
            lib.free_mpq_mat(x, nb, 1)
        elif status == "singular":
            sing += 1
        else:
            bad += 1
        lib.free_dense(B); lib.free_sparse(A)
    tot = reduce_sum([ok, sing, bad])
    out["batch"] = {"workload": f"BASELINE configs[4]: {nsys} independent LP-basis style systems n={nb} (synth.lp_basis seeds 5000..), 1 RHS each, "
                                "SLIP_LU_analyze + SLIP_solve_mpq per system",
                    "seconds": t_batch, "systems_per_s": nsys / t_batch if t_batch > 0 else None,
                    "host_threads_per_rank": threads,
                    "verified_exact": int(tot[0]), "singular_by_construction": int(tot[1]), "failed": int(tot[2]),
                    "verification": "SLIP_check_solution (A x = b in rational arithmetic) on every solved system, outside the timed region"}

    # (ii) configs[3]: one matrix, many right-hand sides, columns of b sharded over the ranks
    n3, nrhs = args.mrhs_n, args.mrhs_rhs
    system = None
    for seed in range(3, 40):                                    # some random LP bases are exactly singular
        n, cp, ri, vals, _ = synth.lp_basis(n3, seed=seed, nrhs=1)
        A = lib.sparse_from_csc(n, cp, ri, vals)
        o = lib.default_options()
        S = lib.analyze(A, o)
        B1 = lib.dense_from_rows([[1 + (r % 7)] for r in range(n)])
        try:
            x = lib.solve_mpq(A, S, B1, o)
            lib.free_mpq_mat(x, n, 1)
            system = (A, S, o, seed)
            lib.free_dense(B1)
            break
        except capi.SlipError:
            lib.free_dense(B1); lib.free_analysis(S); lib.free_sparse(A); lib.free_options(o)
    if system is None:
        out["multi_rhs"] = {"error": "no nonsingular instance found"}
        return out
    A, S, o, seed = system
    lo, hi = shard_range(nrhs, world, rank)
    mine = hi - lo
    rng = np.random.default_rng(1234)
    ball = rng.integers(-(1 << 20), 1 << 20, size=(n3, nrhs), dtype=np.int32)
    ball[ball == 0] = 1
    t_local = t_fac = 0.0
    checked = good = 0
    if mine > 0:
        bs = np.ascontiguousarray(ball[:, lo:hi])
        rows = (C.POINTER(C.c_int32) * n3)(*[C.cast(bs[r].ctypes.data, C.POINTER(C.c_int32)) for r in range(n3)])
        lib.dll.SLIP_build_dense_int.restype = C.c_int
        lib.dll.SLIP_build_dense_int.argtypes = [C.POINTER(capi.SLIP_dense), C.POINTER(C.POINTER(C.c_int32)), C.c_int32, C.c_int32]
        B = lib.dll.SLIP_create_dense()
        assert lib.dll.SLIP_build_dense_int(B, rows, n3, mine) == 0
    barrier()
    if mine > 0:
        t0 = time.perf_counter()
        x = lib.solve_mpq(A, S, B, o)
        t_local = time.perf_counter() - t0
        t_fac = last_stats(lib)[9]
        # exact check of a sample of this rank's columns (rational A x = b costs ~1 s per column here)
        for cidx in sorted({0, mine - 1}):
            xcol = (C.POINTER(capi.MpqStruct) * n3)(*[C.cast(C.byref(x[r][cidx]), C.POINTER(capi.MpqStruct)) for r in range(n3)])
            bcol = lib.dense_from_rows([[int(bs[r][cidx])] for r in range(n3)])
            checked += 1
            good += int(lib.dll.SLIP_check_solution(A, C.cast(xcol, C.POINTER(C.POINTER(capi.MpqStruct))), bcol) == 0)
            lib.free_dense(bcol)
        lib.free_mpq_mat(x, n3, mine)
        lib.free_dense(B)
    t_all = reduce_max(t_local)
    t_fac_all = reduce_max(t_fac)
    tot = reduce_sum([checked, good])
    st = last_stats(lib)
    out["multi_rhs"] = {"workload": f"BASELINE configs[3]: LP-basis style matrix n={n3} (synth.lp_basis seed {seed}), {nrhs} right-hand sides "
                                    f"(21-bit integers), columns of b sharded over the ranks; every rank calls SLIP_LU_analyze once and "
                                    "SLIP_solve_mpq on its columns",
                        "seconds": t_all, "rhs_per_s": nrhs / t_all if t_all > 0 else None,
                        "replicated_factorization_seconds": t_fac_all,
                        "amdahl": "every rank factors the same matrix (the column chain does not shard): that term does not shrink with N",
                        "channels": int(st[3]), "nnz_L": int(st[1]), "nnz_U": int(st[2]),
                        "columns_checked_exactly": int(tot[0]), "columns_exact": int(tot[1]),
                        "verification": "SLIP_check_solution on the first and last column of every rank; bound-mode results are also verified inside the library (A N = det b)"}
    lib.free_analysis(S); lib.free_sparse(A); lib.free_options(o)
    if world == 1 and not args.no_cpu_baseline:
        leg = cpu_leg_subprocess(["--cpu-leg", f"sharded:{seed}", "--batch-n", str(nb), "--mrhs-n", str(n3), "--mrhs-rhs", str(nrhs)])
        if "batch_sample_s" in leg:
            cpu_batch = leg["batch_sample_s"] * nsys / leg["batch_sample_systems"]
            cpu_mrhs = leg["mrhs_factorize_s"] + leg["mrhs_solve_4_rhs_s"] * nrhs / 4.0
            out["batch"]["cpu"] = {"kind": leg["kind"], "sample": "the first 64 of the systems, one after the other",
                                   "sample_seconds": leg["batch_sample_s"], "whole_job_seconds_extrapolated": cpu_batch,
                                   "ratio_vs_gpu": cpu_batch / out["batch"]["seconds"]}
            out["multi_rhs"]["cpu"] = {"kind": leg["kind"], "sample": "SLIP_LU_analyze + SLIP_LU_factorize in full, SLIP_LU_solve on 4 of the right-hand sides",
                                       "factorize_seconds": leg["mrhs_factorize_s"], "solve_4_rhs_seconds": leg["mrhs_solve_4_rhs_s"],
                                       "whole_job_seconds_extrapolated": cpu_mrhs, "ratio_vs_gpu": cpu_mrhs / out["multi_rhs"]["seconds"]}
        else:
            out["cpu_error"] = leg.get("error")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--ref-n", type=int, default=240)
    ap.add_argument("--seed", type=int, default=SEED)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-leg", default=None, help="(internal) time the unmodified reference on one named head-to-head system and print {cpu_s}")
    ap.add_argument("--extras", default="all", help="all | none | comma list of h2h,laplacian,lu,sharded,intmul")
    ap.add_argument("--lap-grid", type=int, default=64, help="grid side of the configs[2]-family block")
    ap.add_argument("--batch-systems", type=int, default=512)
    ap.add_argument("--batch-n", type=int, default=500)
    ap.add_argument("--mrhs-n", type=int, default=10000)
    ap.add_argument("--mrhs-rhs", type=int, default=256)
    ap.add_argument("--clock-period-ms", type=int, default=1000, help="nvidia-smi sampling period during the timed region")
    ap.add_argument("--no-kernel-timers", action="store_true", help="(diagnostic) no CUDA-event brackets around the kernels: roofline is then not measured")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    extras = set(["h2h", "laplacian", "lu", "sharded", "intmul"]) if args.extras == "all" else \
        (set() if args.extras == "none" else set(args.extras.split(",")))

    # stdout carries exactly one JSON line: anything libraries print to fd 1 on the way (NCCL prints
    # its version there) is sent to stderr, the result goes to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cores = list(range(os.cpu_count() or 1))
    if world > 1:
        # torchrun pins OMP_NUM_THREADS=1 for multi-rank jobs; the host side of the interface (limb
        # export, canonical rationals) is OpenMP-parallel: every rank gets its own share of the
        # cores, and stays on it (ranks that roam over each other's cores cost 20 % at N = 8)
        per = max(1, len(cores) // world)
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        try:
            os.sched_setaffinity(0, mine)
        except (AttributeError, OSError):
            pass
        os.environ["OMP_NUM_THREADS"] = str(len(mine))
        os.environ.setdefault("OMP_PROC_BIND", "false")

    if args.cpu_leg:
        from slip_lu_b200 import refmats
        import __graft_entry__ as entry
        entry.build()
        ref = reference_lib()
        if ref is None:
            emit({"error": "oracle/_ref/libslip_ref.so is not built"})
            return
        if args.cpu_leg.startswith("sharded:"):
            # bounded samples of the two sharded jobs on the host: 64 of the configs[4] systems, and the
            # configs[3] factorization with 4 of its right-hand sides (seed of the matrix after the colon)
            from slip_lu_b200 import capi, synth
            t_batch = 0.0
            done = 0
            for g in range(64):
                n, cp, ri, vals, b = synth.lp_basis(args.batch_n, seed=5000 + g, nrhs=1)
                A = ref.sparse_from_csc(n, cp, ri, vals); B = ref.dense_from_rows(b); o = ref.default_options()
                t0 = time.perf_counter()                       # the interface calls only, as in the GPU leg
                S = ref.analyze(A, o)
                try:
                    x = ref.solve_mpq(A, S, B, o)
                    t_batch += time.perf_counter() - t0
                    ref.free_mpq_mat(x, n, 1)
                    done += 1
                except capi.SlipError:
                    t_batch += time.perf_counter() - t0
                ref.free_analysis(S); ref.free_dense(B); ref.free_sparse(A); ref.free_options(o)
            seed = int(args.cpu_leg.split(":")[1])
            n, cp, ri, vals, _ = synth.lp_basis(args.mrhs_n, seed=seed, nrhs=1)
            import numpy as np
            ball = np.random.default_rng(1234).integers(-(1 << 20), 1 << 20, size=(n, args.mrhs_rhs), dtype=np.int32)
            ball[ball == 0] = 1
            A = ref.sparse_from_csc(n, cp, ri, vals); o = ref.default_options()
            B = ref.dense_from_rows([[int(v) for v in ball[r, :4]] for r in range(n)])
            t0 = time.perf_counter()
            S = ref.analyze(A, o)
            L, U, rhos, pinv = ref.factorize(A, S, o)
            t_fac = time.perf_counter() - t0
            t0 = time.perf_counter()
            x = ref.lu_solve(B, rhos, L, U, pinv)
            t_sol = time.perf_counter() - t0
            emit({"kind": "unmodified reference (oracle/_ref), 1 thread, this box, own process",
                  "batch_sample_systems": 64, "batch_sample_s": t_batch, "batch_sample_solved": done,
                  "mrhs_factorize_s": t_fac, "mrhs_solve_4_rhs_s": t_sol})
            return
        emit({"cpu_s": cpu_solve_system(ref, refmats.system(args.cpu_leg)), "workload": args.cpu_leg,
              "kind": "unmodified reference (oracle/_ref), 1 thread, this box, own process"})
        return

    if args.impl == "reference":
        if rank != 0:
            return
        import __graft_entry__ as entry
        entry.build()
        r = run_reference(args, rank)
        emit({
            "impl": "reference", "metric": "limb_mul_ops_per_s", "value": r["value"], "unit": "limb-mul/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "GMP mpz (64-bit limbs)", "data": "synthetic",
            "config": headline_config(args.n, args.ref_n), "same_config": False,
            "cpu_baseline": {"value": r["value"], "unit": "limb-mul/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "factor_solve_seconds": r["seconds"], "ref_updates": r["updates"],
            "e2e": {"value": r["value"], "unit": "limb-mul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if local_rank == 0:
        entry.build()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    if local_rank != 0:
        entry.build()                       # no-op: everything is already built
    import slip_lu_b200
    lib = slip_lu_b200.lib()
    if lib.dll.SLIP_B200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU path")
    lib.dll.SLIP_B200_set_device(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib.dll.SLIP_B200_last_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
    lib.dll.SLIP_B200_last_pinv.argtypes = [C.POINTER(C.c_int32), C.c_int]

    n, cp, ri, vals, b = workload(args.n, args.seed, rhs_seed=rank)
    o = lib.default_options()
    A = lib.sparse_from_csc(n, cp, ri, vals)          # host mpz_t matrices: the interface's inputs
    B = lib.dense_from_rows(b)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(vs):
        t = torch.tensor([float(v) for v in vs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)      # the final gather of a sharded job
        return [float(v) for v in t.tolist()]

    def step():
        S = lib.analyze(A, o)                 # COLAMD on the host
        x = lib.solve_mpq(A, S, B, o)         # H2D of A and b, GPU factor + solve, D2H, host mpq_t result
        return S, x

    x_last = None
    for _ in range(args.warmup):
        S, x = step()
        lib.free_mpq_mat(x, n, 1); lib.free_analysis(S)

    sampler = ClockSampler(local_rank, args.clock_period_ms)
    lib.dll.slipcu_reset_counters()
    lib.dll.slipcu_set_profiling(0 if args.no_kernel_timers else 1)   # CUDA-event bracket around every k_trisolve launch
    sampler.start()                           # (process start-up of the sampler stays outside the clock)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if x_last is not None:
            lib.free_mpq_mat(x_last[1], n, 1); lib.free_analysis(x_last[0])
        x_last = step()
    torch.cuda.synchronize(dev)
    t_local = time.perf_counter() - t0
    clocks = sampler.stop()
    lib.dll.slipcu_set_profiling(0)
    c = counters(lib)
    st = last_stats(lib)
    updates, limbmul = st[4], st[5]

    # exactness of the timed result (outside the timed region): A x == b in rational arithmetic
    ok = lib.dll.SLIP_check_solution(A, x_last[1], B)
    if ok != 0:
        raise SystemExit("bench.py: the solution of the timed step does not satisfy A x = b exactly")
    lib.free_mpq_mat(x_last[1], n, 1); lib.free_analysis(x_last[0])

    t_wall = reduce_max(t_local)
    t_dev = reduce_max(c.device_ms / 1e3)
    W_total, U_total = reduce_sum([limbmul * args.steps, updates * args.steps])

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"

    out = {
        "metric": "limb_mul_ops_per_s", "value": W_total / t_dev, "unit": "limb-mul/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 residues (31-bit prime channels) + 32-bit limbs",
        "data": "synthetic", "config": headline_config(args.n), "clocks": clocks,
        "metric_note": "schoolbook-EQUIVALENT limb multiplications of the input (work model, same formula in both arms), not multiplies executed: read with factor_solve_seconds and the roofline",
        "e2e": {"value": W_total / t_wall, "unit": "limb-mul/s",
                "h2d_bytes_per_step": c.h2d_bytes / args.steps, "d2h_bytes_per_step": c.d2h_bytes / args.steps},
        "gpu_launches": int(c.launches),
        "factor_solve_seconds": t_wall / args.steps,
        "device_seconds_per_step": t_dev / args.steps,
        "ref_updates_per_s": U_total / t_wall,
        "problem": {"n": n, "nnz_L": st[1], "nnz_U": st[2], "channels": st[3], "channels_hadamard": st[10],
                    "bound_mode_restarts": st[11], "ref_updates": updates,
                    "limb_mul_equiv": limbmul, "exact_check": "A x = b verified in rational arithmetic"},
        "roofline": roofline_of(c, peak, peak_src, t_dev),
        "reconstruction": {"kernel": "k_fraccrt (approximate magnitudes for the pivot search) + k_garner_flow / k_limbs (exact solution numerators)",
                           "ms": c.recon_ms / args.steps,
                           "mac_per_s": c.recon_mac / (c.recon_ms / 1e3) if c.recon_ms > 0 else None},
    }

    if "intmul" in extras and rank == 0:
        pk = (C.c_double * 10)()
        if lib.dll.slipcu_measure_int_peaks(pk) == 0 and pk[9] > 0:
            mm = out["roofline"]["modmul_per_s"] or 0.0
            sms = 148
            mhz = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965)
            int_peak = pk[9] * sms * mhz * 1e6      # modular multiply-subtracts/s at the clock of the timed region
            out["roofline"]["int_mul"] = {
                "what": "k_trisolve's modular multiply-subtracts per second against the rate of the same operation on registers only",
                "unit": "modular multiply-subtracts/s (each: a Shoup product by the step's fixed multiplier, IMAD.HI + 2 IMAD, + 3 integer ALU operations)",
                "achieved": mm, "peak": int_peak, "frac": mm / int_peak if int_peak else None,
                "peak_per_sm_cycle": {"shoup_multiply_subtract": pk[9], "montgomery_multiply_subtract": pk[8],
                                      "imad_wide": pk[5], "imad": pk[6], "imad_hi": pk[7]},
                "peak_as_run_per_s": {"shoup_multiply_subtract": pk[4], "montgomery_multiply_subtract": pk[3],
                                      "imad_wide": pk[0], "imad": pk[1], "imad_hi": pk[2]},
                "sm_mhz_used": mhz, "imad_per_s": 3.0 * mm,
                "peak_source": "measured live (slipcu_measure_int_peaks): register-resident chains, 8 per thread, 8 CTAs x 256 threads per SM, "
                               "operations per SM cycle from clock64 inside the kernel x 148 SMs x the SM clock sampled during the timed region "
                               "(the microbenchmark itself runs power-capped at a lower clock: peak_as_run_per_s)"}

    block_s = {}

    def timed_block(name, fn):
        t = time.perf_counter()
        r = fn()
        block_s[name] = time.perf_counter() - t
        return r

    if world == 1 and not args.no_cpu_baseline:
        r = timed_block("cpu_baseline", lambda: cpu_leg_subprocess(
            ["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-n", str(args.ref_n), "--n", str(args.n),
             "--seed", str(args.seed)]))
        if "cpu_baseline" in r:
            out["cpu_baseline"] = dict(r["cpu_baseline"], seconds=r["factor_solve_seconds"], same_config=False,
                                       note="bounded sample (n=%d), not the headline configuration: see head_to_head for measured same-configuration ratios" % args.ref_n)
        else:
            out["cpu_baseline"] = {"error": r.get("error")}
    if world == 1 and "h2h" in extras:
        out["head_to_head"] = timed_block("head_to_head", lambda: head_to_head(lib, with_cpu=not args.no_cpu_baseline))
    if world == 1 and "laplacian" in extras:
        out["laplacian"] = timed_block("laplacian", lambda: laplacian_block(lib, args.lap_grid, peak, peak_src))
    if world == 1 and "lu" in extras:
        Jc = [j for j in range(n) for _ in range(cp[j], cp[j + 1])]
        out["lu_path"] = timed_block("lu_path", lambda: lu_path_block(lib, A, B, o, n, (n, list(ri), Jc, list(vals), b)))
    if "sharded" in extras:
        out["sharded"] = timed_block("sharded", lambda: sharded_block(lib, args, rank, world, barrier, reduce_max, reduce_sum))
    out["extras_wall_seconds"] = block_s
    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
