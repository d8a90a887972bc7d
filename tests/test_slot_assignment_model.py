"""Host-side model of the row assignment of k_slots (slipcu.cu; DESIGN.md section 4, "Last step of
round 2") and of the Shoup companion of k_trisolve.  Pure Python, no GPU.

k_slots: the rows of a pipeline chunk are handed to the threads of k_trisolve so that the G = 32/CH
row groups of one quarter-warp access fall into distinct shared-memory bank groups: warp a owns the
rows r = a (mod G) of the chunk (always 128) and orders them, stable, by (slot - a) mod G, rows
without a target last; the row of class a at position i of that order goes to row group
g = (i mod 32) * G + a, entry q = i // 32.  The model repeats the kernel's arithmetic (ballot ranks
included) and checks what k_trisolve and the bound CTA rely on:

* every row with a target appears exactly once, with its own slot, among the entries of the
  active row groups, and every entry of an active row group is written;
* the L rows of one access are one of each class (never a bank conflict on the stage buffer);
* on scattered patterns the targets of an access conflict far less than in the plain order.

Shoup companion: floor(ny * 2^32 / p) == (ny * 2^32 mod p) * (-1/p mod 2^32) mod 2^32, the identity
behind the single IMAD of k_trisolve."""
import random

import pytest


def popc(v):
    return bin(v).count("1")


def k_slots_chunk(CH, nrows, slot_of_row, pivrow, cnt, sort):
    """One CTA of k_slots: returns {list index: (slot, chunk row)} and the number of active row groups."""
    G, RG = 32 // CH, 1024 // CH
    R = 4 * RG
    up = (nrows + G - 1) // G * G
    active = RG if nrows == R else min(up, RG)                 # tri_active_groups
    npass = (((nrows + G - 1) // G) + 31) >> 5
    out = {}
    for a in range(G):                                         # warp a = class a
        slot = [[0] * 32 for _ in range(4)]
        key = [[0] * 32 for _ in range(4)]
        idx = [[0] * 32 for _ in range(4)]
        for p in range(4):
            for lane in range(32):
                r = a + G * (p * 32 + lane)
                target = r < nrows and r != pivrow
                slot[p][lane] = slot_of_row[r] if target else cnt
                key[p][lane] = (((slot[p][lane] - a) & (G - 1)) if sort else 0) if target else G
                idx[p][lane] = p * 32 + lane
        if sort or npass < 4:
            run = 0
            for k in range(G + 1):
                if not sort and 0 < k < G:
                    continue
                for p in range(4):
                    if p >= npass:
                        continue
                    ballot = sum(1 << lane for lane in range(32) if key[p][lane] == k)
                    for lane in range(32):
                        if key[p][lane] == k:
                            idx[p][lane] = run + popc(ballot & ((1 << lane) - 1))
                    run += popc(ballot)
        for p in range(4):
            for lane in range(32):
                g, q = (idx[p][lane] & 31) * G + a, idx[p][lane] >> 5
                r = a + G * (p * 32 + lane)
                if g < active:
                    assert 4 * g + q not in out, "two rows at one list position"
                    out[4 * g + q] = (slot[p][lane], r)
    return out, active


def run_chunk(CH, nrows, sort, seed, dense=False):
    rng = random.Random(seed)
    G = 32 // CH
    cnt = nrows if dense else rng.randint(nrows, 3 * nrows + 5)
    targets = sorted(rng.sample(range(cnt), nrows))
    piv = rng.randrange(nrows) if rng.random() < 0.7 else -1
    out, active = k_slots_chunk(CH, nrows, targets, piv, cnt, sort)
    assert len(out) == 4 * active                              # every entry of an active group is written
    seen = set()
    for slot, r in out.values():
        assert 0 <= r < 4096 // CH and r < (1 << 10) and slot <= cnt < (1 << 22)
        if slot != cnt:
            assert r < nrows and r != piv and targets[r] == slot and r not in seen
            seen.add(r)
    assert len(seen) == nrows - (1 if piv >= 0 else 0)         # every row with a target, once
    w_wave = w_acc = l_wave = 0
    for q in range(4):
        for g0 in range(0, active, G):
            ws, ls = {}, {}
            for g in range(g0, g0 + G):
                slot, r = out[4 * g + q]
                if slot != cnt:
                    ws[slot % G] = ws.get(slot % G, 0) + 1
                ls[r % G] = ls.get(r % G, 0) + 1
            assert max(ls.values()) == 1                       # L rows of an access: one of each class
            if ws:
                w_wave += max(ws.values()); w_acc += 1
    return w_wave, w_acc


@pytest.mark.parametrize("CH", [4, 8, 16, 32])
def test_every_row_is_assigned_once(CH):
    RG = 1024 // CH
    R = 4 * RG
    for nrows in (1, 2, 3, 5, RG - 1, RG, RG + 1, R // 2 + 3, 3 * R // 4, 3 * R // 4 + 1, R - 1, R):
        for sort in (0, 1):
            for seed in range(3):
                run_chunk(CH, nrows, sort, seed)


@pytest.mark.parametrize("CH", [4, 8])
def test_sorted_order_removes_most_conflicts(CH):
    R = 4096 // CH
    plain = [run_chunk(CH, R, 0, 100 + s) for s in range(12)]
    by_group = [run_chunk(CH, R, 1, 100 + s) for s in range(12)]
    rate = lambda rows: sum(w for w, _ in rows) / sum(a for _, a in rows)
    extra_plain, extra_sorted = rate(plain) - 1.0, rate(by_group) - 1.0
    assert extra_plain > 0.5                                   # scattered targets: ~2 wavefronts per access
    assert extra_sorted < 0.6 * extra_plain                    # most of the excess is gone
    # consecutive slots (a dense trailing block) are conflict-free in both orders
    for sort in (0, 1):
        w, a = run_chunk(CH, R, sort, 5, dense=True)
        assert w - a <= a // 50


def test_shoup_companion_is_one_product():
    from test_fraccrt_model import _primes_below_2_31
    rng = random.Random(3)
    R = 1 << 32
    for p in _primes_below_2_31(40):
        ninv = (-pow(p, -1, R)) % R                            # the Montgomery constant -1/p mod 2^32
        for ny in [0, 1, p - 1] + [rng.randrange(p) for _ in range(300)]:
            nym = ny * R % p                                   # Montgomery form of the multiplier
            assert (ny * R) // p == (nym * ninv) % R
