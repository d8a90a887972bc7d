#!/usr/bin/env python
"""bench.py -- the hot path of SLIP LU (exact sparse factor + solve) on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 2000]

Workload (BASELINE.json configs[1]): synthetic random sparse integer matrix, n = 2000, 10 nonzeros
per column, 32-bit entries, COLAMD column order, 1 right-hand side.  One *step* is one complete
exact solve of that system through the reference-facing C interface:
SLIP_LU_analyze + SLIP_solve_mpq (the call sequence of the reference's Demo/example2.c), host mpz_t
inputs in, canonical host mpq_t solution out.

metric  limb_mul_ops_per_s: schoolbook-equivalent 32-bit limb multiplications of the REF
        elimination per second.  The work W of a system is a property of the input, computed by
        the same formula for both arms: every REF entry update (one per entry of L below the pivot,
        per elimination step j of a column) counts 3 * w_j^2 limb products (two w_j-limb products
        and one exact division), w_j = ceil(bits_j / 32), bits_j = Hadamard prefix bound of pivot j.
value   W / device time (CUDA events, A resident in HBM -> solution numerators reconstructed in
        HBM), all ranks.   e2e: W / wall time of the C-interface call with host buffers.
Extra keys: factor_solve_seconds (the other half of BASELINE.json's metric), roofline of the
dominant kernel (k_trisolve, HBM-bound), cpu_baseline (the unmodified reference on host cores).

--impl reference times the UNMODIFIED reference (oracle/_ref/libslip_ref.so, built from
/root/reference by oracle/Makefile; the oracle port if that build is absent) on a bounded sample:
the same generator at n = --ref-n (default 240), because the n = 2000 system would take the CPU
days.  Same metric, same work formula.
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


class Counters(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("trisolve_launches", C.c_uint64), ("trisolve_ms", C.c_double),
                ("trisolve_bytes", C.c_double), ("trisolve_modmul", C.c_double), ("recon_ms", C.c_double),
                ("recon_mac", C.c_double), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
                ("device_ms", C.c_double), ("other_ms", C.c_double)]


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


def run_reference(args, rank):
    """The reference arm (and the cpu_baseline of the GPU arm): host cores only."""
    from slip_lu_b200 import capi
    from oracle import binding as ob
    n, cp, ri, vals, b = workload(args.ref_n, args.seed)
    kind = "reference" if os.path.exists(ob.REF_SO) else "port"
    times = []
    W = None
    if kind == "reference":
        ref = capi.SlipLib(ob.REF_SO)
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
                      "single thread (the reference is single-threaded); n=2000 would take the CPU days"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--ref-n", type=int, default=240)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period-ms", type=int, default=1000, help="nvidia-smi sampling period during the timed region")
    ap.add_argument("--no-kernel-timers", action="store_true", help="(diagnostic) no CUDA-event brackets around the kernels: roofline is then not measured")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

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
    if world > 1:
        # torchrun pins OMP_NUM_THREADS=1 for multi-rank jobs; the host side of the interface (limb
        # export, canonical rationals) is OpenMP-parallel: give every rank its share of the cores
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = str(max(1, cores // world))

    config = {"workload": f"BASELINE configs[1]: synthetic random sparse integer matrix n={args.n}, 10 nnz/col, "
                          "32-bit entries, COLAMD order, 1 RHS, SLIP_TOL_SMALLEST pivoting (defaults)",
              "per_gpu": "one independent system per GPU (the same matrix, a different right-hand side per rank: equal work per GPU), no data-path collective",
              "l2": "factor data streamed per column (>> 126 MB L2; ~10 GB of L residues at n=2000)"}

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
            "vs_baseline": None, "dtype": "GMP mpz (64-bit limbs)", "data": "synthetic", "config": config,
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

    n, cp, ri, vals, b = workload(args.n, args.seed, rhs_seed=rank)
    o = lib.default_options()
    A = lib.sparse_from_csc(n, cp, ri, vals)          # host mpz_t matrices: the interface's inputs
    B = lib.dense_from_rows(b)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

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
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if x_last is not None:
            lib.free_mpq_mat(x_last[1], n, 1); lib.free_analysis(x_last[0])
        x_last = step()
    torch.cuda.synchronize(dev)
    t_local = time.perf_counter() - t0
    clocks = sampler.stop()
    lib.dll.slipcu_set_profiling(0)
    c = Counters()
    lib.dll.slipcu_get_counters(C.byref(c))
    st = (C.c_double * 10)()
    lib.dll.SLIP_B200_last_stats(st, 10)
    updates, limbmul = st[4], st[5]

    # exactness of the timed result (outside the timed region): A x == b in rational arithmetic
    ok = lib.dll.SLIP_check_solution(A, x_last[1], B)
    if ok != 0:
        raise SystemExit("bench.py: the solution of the timed step does not satisfy A x = b exactly")

    t_wall = torch.tensor([t_local], dtype=torch.float64, device=dev)
    t_dev = torch.tensor([c.device_ms / 1e3], dtype=torch.float64, device=dev)
    w_all = torch.tensor([limbmul * args.steps, updates * args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_wall, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(w_all, op=dist.ReduceOp.SUM)       # the final gather of the sharded job
    t_wall, t_dev = float(t_wall.item()), float(t_dev.item())
    W_total, U_total = float(w_all[0].item()), float(w_all[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    achieved = (c.trisolve_bytes / 1e9) / (c.trisolve_ms / 1e3) if c.trisolve_ms > 0 else 0.0
    # DRAM traffic per launch: the ratio measured by one `ncu --set full` capture of this kernel
    # (profiles/r01_trisolve_ncu_full.json) applied to the live average algorithmic bytes per launch
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r01_trisolve_ncu_full.json")))
        traffic = cap["traffic_over_algorithmic"] * c.trisolve_bytes / max(1, c.trisolve_launches)
    except Exception:
        pass
    out = {
        "metric": "limb_mul_ops_per_s", "value": W_total / t_dev, "unit": "limb-mul/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 residues (31-bit prime channels) + 32-bit limbs",
        "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": W_total / t_wall, "unit": "limb-mul/s",
                "h2d_bytes_per_step": c.h2d_bytes / args.steps, "d2h_bytes_per_step": c.d2h_bytes / args.steps},
        "gpu_launches": int(c.launches),
        "factor_solve_seconds": t_wall / args.steps,
        "device_seconds_per_step": t_dev / args.steps,
        "ref_updates_per_s": U_total / t_wall,
        "problem": {"n": n, "nnz_L": st[1], "nnz_U": st[2], "channels": st[3], "ref_updates": updates,
                    "limb_mul_equiv": limbmul, "exact_check": "A x = b verified in rational arithmetic"},
        "roofline": {"bound": "hbm", "kernel": "k_trisolve", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak if peak else None, "peak_source": peak_src,
                     "launches": int(c.trisolve_launches), "traffic": traffic,
                     "algorithmic_bytes_per_launch": c.trisolve_bytes / max(1, c.trisolve_launches),
                     "modmul_per_s": c.trisolve_modmul / (c.trisolve_ms / 1e3) if c.trisolve_ms > 0 else None,
                     "kernel_share_of_step": (c.trisolve_ms / 1e3) / (t_dev) if t_dev > 0 else None},
        "reconstruction": {"kernel": "k_fraccrt (approximate magnitudes for the pivot search) + k_garner_flow / k_limbs (exact solution numerators)",
                           "ms": c.recon_ms / args.steps,
                           "mac_per_s": c.recon_mac / (c.recon_ms / 1e3) if c.recon_ms > 0 else None},
    }
    if world == 1 and not args.no_cpu_baseline:
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup = 1, 0
        r = run_reference(sub, 0)
        out["cpu_baseline"] = {"value": r["value"], "unit": "limb-mul/s", "cores": r["cores"],
                               "kind": r["kind"], "sample": r["sample"], "seconds": r["seconds"]}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
