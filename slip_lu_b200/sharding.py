"""Partitioning of independent work across GPUs (one process per GPU, no data-path collective).

The column chain of a single factorization does not shard.  What does: independent systems
(BASELINE configs[4], LP-basis batches) and the right-hand sides of one system (configs[3]).
Both are split into contiguous, balanced shards; every rank works alone on its shard and the only
communication is the final gather of results (or of their digests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced shard [lo, hi) of `total` items for `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rhs_columns(b_rows: Sequence[Sequence[int]], world: int, rank: int) -> Tuple[List[List[int]], Tuple[int, int]]:
    """Columns [lo, hi) of a dense right-hand side matrix given as rows."""
    nrhs = len(b_rows[0])
    lo, hi = shard_range(nrhs, world, rank)
    return [list(row[lo:hi]) for row in b_rows], (lo, hi)


def gather_objects(obj, world: int, rank: int):
    """Final gather on rank 0 through torch.distributed (any backend); identity for world == 1."""
    if world == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


def _solve_one(lib, system, options):
    n, cp, ri, vals, b = system
    o = options() if options else lib.default_options()
    A = lib.sparse_from_csc(n, cp, ri, vals)
    B = lib.dense_from_rows(b)
    S = lib.analyze(A, o)
    try:
        x = lib.solve_mpq(A, S, B, o)
        out = lib.mpq_mat_to_py(x, n, len(b[0]))
        lib.free_mpq_mat(x, n, len(b[0]))
    finally:
        lib.free_analysis(S); lib.free_dense(B); lib.free_sparse(A); lib.free_options(o)
    return out


def solve_batch_sharded(lib, systems, world: int, rank: int, options=None, threads: int = 1):
    """Solve the shard of `systems` (each (n, colptr, rowidx, values, b_rows)) owned by this rank with
    SLIP_LU_analyze + SLIP_solve_mpq; returns [(global_index, solution)] for the shard.

    Small systems leave most of a B200 idle (a column is a handful of short launches), so with
    threads > 1 several systems of the shard are in flight at once: every factorization session has
    its own CUDA stream, and the ctypes calls release the GIL."""
    lo, hi = shard_range(len(systems), world, rank)
    if threads <= 1:
        return [(g, _solve_one(lib, systems[g], options)) for g in range(lo, hi)]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads) as pool:
        futs = [(g, pool.submit(_solve_one, lib, systems[g], options)) for g in range(lo, hi)]
        return [(g, f.result()) for g, f in futs]
