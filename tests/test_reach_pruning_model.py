"""Host-side model of the symbolic side of slip_factorize.c: the pattern of column k is the set of
rows reachable from A(:,q[k]) through the finished columns of L (slip_reach.c / slip_dfs.c), and
the finished columns are pruned symmetrically (Eisenstat-Liu, prune_columns): once row `prow`
became the pivot of column k, every column j of the U part of column k whose L part contains prow
keeps only its pivotal rows visible to later searches.  The reach SETS must be the same with and
without pruning, for any pivot choices.  Pure Python, no GPU."""
import random


def _reach(col_rows, pinv, L, visible, klim):
    seen, out, stack = set(), [], []
    for r0 in col_rows:
        if r0 in seen:
            continue
        seen.add(r0); stack.append(r0)
        while stack:
            r = stack.pop()
            out.append(r)
            pos = pinv[r]
            if pos < klim:
                rows = L[pos] if visible is None else L[pos][:visible[pos]]
                for rr in rows:
                    if rr not in seen:
                        seen.add(rr); stack.append(rr)
    return set(out)


def _simulate(n, density, seed):
    rng = random.Random(seed)
    cols = []
    for j in range(n):
        rows = {j} | {rng.randrange(n) for _ in range(rng.randrange(1, density + 1))}
        cols.append(sorted(rows))
    pinv = list(range(n)); row_at = list(range(n))
    L_full, L_pruned, vis, pruned = [], [], [], []
    work_full = work_pruned = 0
    for k in range(n):
        pat_full = _reach(cols[k], pinv, L_full, None, k)
        pat_pruned = _reach(cols[k], pinv, L_pruned, vis, k)
        assert pat_full == pat_pruned, f"column {k}: pruned reach differs"
        work_full += sum(len(L_full[pinv[r]]) for r in pat_full if pinv[r] < k)
        work_pruned += sum(vis[pinv[r]] for r in pat_pruned if pinv[r] < k)
        cand = sorted(r for r in pat_full if pinv[r] >= k)
        assert cand, "structurally singular case: choose another seed"
        prow = rng.choice(cand)                               # any pivot rule
        oldpos, displaced = pinv[prow], row_at[k]
        row_at[k], row_at[oldpos] = prow, displaced
        pinv[prow], pinv[displaced] = k, oldpos
        upos = sorted(pinv[r] for r in pat_full if pinv[r] < k)
        L_full.append(list(cand)); L_pruned.append(list(cand)); vis.append(len(cand)); pruned.append(False)
        for j in upos:                                        # prune_columns
            if pruned[j]:
                continue
            rj = L_pruned[j]
            if prow not in rj[:vis[j]]:
                continue
            head, tail = 0, vis[j]
            while head < tail:
                if pinv[rj[head]] <= k:
                    head += 1
                else:
                    tail -= 1
                    rj[head], rj[tail] = rj[tail], rj[head]
            vis[j] = tail; pruned[j] = True
    return work_full, work_pruned


def test_pruned_reach_equals_full_reach():
    total_full = total_pruned = 0
    for seed in range(12):
        wf, wp = _simulate(60 + 7 * seed, 2 + seed % 4, seed)
        total_full += wf; total_pruned += wp
    assert total_pruned < total_full          # and it does save search work
