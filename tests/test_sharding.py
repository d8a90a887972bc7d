"""CPU: the N>1 partitioning logic, world_size 2 over gloo (no GPU involved: each rank runs the
CPU oracle on its shard, the final gather and the merged result are what is under test)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from slip_lu_b200.sharding import shard_range
    for total in (0, 1, 7, 256, 512, 513):
        for world in (1, 2, 4, 8):
            parts = [shard_range(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from slip_lu_b200 import synth
    from slip_lu_b200.sharding import shard_range, shard_rhs_columns, gather_objects
    from oracle import binding as ob
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # (1) a batch of independent systems (configs[4] style), sharded by system
    systems = [synth.lp_basis(40, seed=s, nrhs=1) for s in range(5)]
    lo, hi = shard_range(len(systems), world, rank)
    mine = []
    for g in range(lo, hi):
        n, cp, ri, vals, b = systems[g]
        qq = list(range(n))
        f = ob.factorize(n, cp, ri, vals, qq)
        mine.append((g, ob.solve(f, b)))
    batch = gather_objects(mine, world, rank)
    # (2) one system, many right-hand sides (configs[3] style), sharded by column
    n, cp, ri, vals, b = synth.random_sparse(20, 4, 16, seed=9, nrhs=5)
    bs, (c0, c1) = shard_rhs_columns(b, world, rank)
    f = ob.factorize(n, cp, ri, vals, list(range(n)))
    cols = gather_objects((c0, c1, ob.solve(f, bs) if c1 > c0 else []), world, rank)
    if rank == 0:
        q.put((batch, cols))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    import torch.multiprocessing as mp
    from slip_lu_b200 import synth
    from oracle import binding as ob
    ob.build()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    batch, cols = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # batch: every system solved exactly once, same answers as a single process
    merged = dict(pair for part in batch for pair in part)
    assert sorted(merged) == list(range(5))
    for g in range(5):
        n, cp, ri, vals, b = synth.lp_basis(40, seed=g, nrhs=1)
        f = ob.factorize(n, cp, ri, vals, list(range(n)))
        assert merged[g] == ob.solve(f, b)
    # right-hand sides: concatenating the column shards gives the unsharded solution
    n, cp, ri, vals, b = synth.random_sparse(20, 4, 16, seed=9, nrhs=5)
    f = ob.factorize(n, cp, ri, vals, list(range(n)))
    want = ob.solve(f, b)
    got = [[] for _ in range(n)]
    for c0, c1, part in sorted(cols):
        for r in range(n):
            got[r].extend(part[r] if part else [])
    assert got == want
