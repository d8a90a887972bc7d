"""GPU: channel sizing below the a-priori (Hadamard) bound.

Sessions may start with a fraction of the channels the Hadamard bound asks for.  SLIP_LU_factorize
sessions PROVE every column's size on the device (bound mode); SLIP_solve_* sessions MEASURE the
candidates and verify the final numerators exactly (A N = det b).  Both restart with more channels
when a column does not fit.  These tests force every branch of that machinery and compare with the
oracle (bit-exact L, U, rhos, pinv, x).
"""
import ctypes as C

import pytest

from slip_lu_b200 import capi, synth
import cases

pytestmark = pytest.mark.gpu


def stats(lib):
    lib.dll.SLIP_B200_last_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
    st = (C.c_double * 13)()
    lib.dll.SLIP_B200_last_stats(st, 13)
    return dict(channels=int(st[3]), hadamard=int(st[10]), restarts=int(st[11]), verified=int(st[12]))


def solve_mpq_py(lib, n, cp, ri, vals, b, q):
    o = lib.default_options(order=capi.SLIP_NO_ORDERING)
    A = lib.sparse_from_csc(n, cp, ri, vals)
    B = lib.dense_from_rows(b)
    S = lib.analyze(A, o, q=q)
    x = lib.solve_mpq(A, S, B, o)
    lib.dll.SLIP_B200_last_pinv.argtypes = [C.POINTER(C.c_int32), C.c_int]
    pv = (C.c_int32 * n)()
    lib.dll.SLIP_B200_last_pinv(pv, n)
    assert lib.dll.SLIP_check_solution(A, x, B) == 0
    got = lib.mpq_mat_to_py(x, n, len(b[0]))
    # SLIP_solve_mpq returns x in the original column order: undo Q to compare with SLIP_LU_solve's order
    xf = [got[q[i]] for i in range(n)]
    lib.free_mpq_mat(x, n, len(b[0])); lib.free_analysis(S); lib.free_dense(B); lib.free_sparse(A); lib.free_options(o)
    return xf, list(pv)


@pytest.mark.parametrize("mode", ["measured", "proven"])
def test_solve_restarts_when_the_start_is_too_small(gpu, oracle, monkeypatch, mode):
    """Wide random entries: the sizes follow the Hadamard bound, 32 channels cannot hold them, the
    session must notice (measured size / proven bound), restart, and still reproduce the oracle's
    row permutation and x."""
    monkeypatch.setenv("SLIP_B200_START_CHANNELS", "32")
    if mode == "proven":
        monkeypatch.setenv("SLIP_B200_BOUND", "proven")
    n, cp, ri, vals, b = synth.random_sparse(120, 6, 32, seed=77, nrhs=2)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    x, pinv = solve_mpq_py(gpu, n, cp, ri, vals, b, q)
    st = stats(gpu)
    assert st["restarts"] >= 1, st
    assert pinv == want["pinv"]
    assert x == want["x"]


def test_small_determinants_run_on_few_channels(gpu, oracle):
    """LP-basis style matrix: determinant far below its Hadamard bound.  The solve session must stay on
    a fraction of the Hadamard channel count, without restart, with the oracle's pinv and x."""
    n, cp, ri, vals, b = synth.lp_basis(400, seed=21, nrhs=2)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    x, pinv = solve_mpq_py(gpu, n, cp, ri, vals, b, q)
    st = stats(gpu)
    assert pinv == want["pinv"] and x == want["x"]
    det_bits = abs(want["rhos"][-1]).bit_length()
    assert st["channels"] * 31 < max(4 * det_bits, 40 * 31) or st["channels"] <= st["hadamard"] // 4, (st, det_bits)


def test_factorize_restarts_in_bound_mode(gpu, oracle, monkeypatch):
    """SLIP_LU_factorize (L and U returned): proven bound mode, forced restart, bit-exact factors."""
    monkeypatch.setenv("SLIP_B200_START_CHANNELS", "32")
    n, cp, ri, vals, b = synth.random_sparse(100, 6, 32, seed=78, nrhs=2)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    got = cases.run_library(gpu, n, cp, ri, vals, b, q)
    assert stats(gpu)["restarts"] >= 1
    cases.assert_same_factorization(got, want, "bound mode restart")


def test_large_rhs_is_verified_and_widened(gpu, oracle, monkeypatch):
    """Factors that fit few channels, a right-hand side that does not: det*x is reconstructed over the
    channels carried, fails the exact check A N = det b, and the solve starts over with more."""
    monkeypatch.setenv("SLIP_B200_START_CHANNELS", "32")
    import random
    want = None
    for seed in range(9, 30):                 # random LP bases are sometimes exactly singular
        n, cp, ri, vals, b = synth.lp_basis(150, seed=seed, nrhs=1)
        rng = random.Random(5)
        b = [[rng.getrandbits(3000) - (1 << 2999)] for _ in range(n)]
        q = cases.colamd_like_order(n, cp, ri)
        try:
            want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
            break
        except RuntimeError:
            continue
    assert want is not None
    x, pinv = solve_mpq_py(gpu, n, cp, ri, vals, b, q)
    assert pinv == want["pinv"] and x == want["x"]


def test_adaptive_off_gives_the_same_result(gpu, oracle, monkeypatch):
    monkeypatch.setenv("SLIP_B200_ADAPTIVE", "0")
    n, cp, ri, vals, b = synth.lp_basis(200, seed=5, nrhs=2)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    x, pinv = solve_mpq_py(gpu, n, cp, ri, vals, b, q)
    st = stats(gpu)
    assert st["restarts"] == 0 and st["channels"] >= st["hadamard"]
    assert pinv == want["pinv"] and x == want["x"]


def test_finalize_releases_cached_memory_and_the_library_keeps_working(gpu, oracle):
    """SLIP_finalize hands the cached device blocks and pinned buffers back to the driver (and drops
    resident factors); the next call allocates afresh and gives the same bits."""
    n, cp, ri, vals, b = synth.random_sparse(60, 5, 24, seed=91, nrhs=2)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    x1, p1 = solve_mpq_py(gpu, n, cp, ri, vals, b, q)
    gpu.dll.SLIP_finalize()
    gpu.dll.SLIP_initialize()
    x2, p2 = solve_mpq_py(gpu, n, cp, ri, vals, b, q)
    got = cases.run_library(gpu, n, cp, ri, vals, b, q)
    assert x1 == want["x"] and x2 == want["x"] and p1 == p2 == want["pinv"]
    cases.assert_same_factorization(got, want, "after SLIP_finalize")
