"""CPU: the oracle (oracle/ref_oracle.c) against the golden fixtures generated from the reference,
and against the reference library itself where oracle/_ref exists."""
import pytest

from slip_lu_b200 import capi
import cases


@pytest.mark.parametrize("name", cases.GOLDEN_FACTOR_CASES)
def test_oracle_matches_golden(oracle, name):
    g = cases.load_golden(name)
    n, cp, ri, vals, b = cases.golden_system(g)
    got = cases.run_oracle(oracle, n, cp, ri, vals, b, g["q"], g["options"]["pivot"], g["options"]["tol"])
    cases.check_against_golden(g, got, ob=oracle)


@pytest.mark.parametrize("pivot", range(6))
def test_oracle_matches_reference_live(oracle, reference, pivot):
    """Fresh random inputs every pivot rule: oracle == unmodified reference, entry for entry."""
    from slip_lu_b200 import synth
    tol = {3: 0.2, 4: 0.01}.get(pivot)
    for seed in (1, 2):
        n, cp, ri, vals, b = synth.random_sparse(40 + 10 * seed, 5, 24, seed=100 * pivot + seed, nrhs=2)
        o = reference.default_options(pivot=pivot, order=capi.SLIP_COLAMD, tol=tol)
        A = reference.sparse_from_csc(n, cp, ri, vals)
        S = reference.analyze(A, o)
        q = [S.contents.q[k] for k in range(n)]
        want = cases.run_library(reference, n, cp, ri, vals, b, q, pivot, tol)
        got = cases.run_oracle(oracle, n, cp, ri, vals, b, q, pivot, o.contents.tol)
        cases.assert_same_factorization(got, want, f"pivot {pivot} seed {seed}")


def test_oracle_singular(oracle):
    # two identical columns
    n, cp, ri, vals = 3, [0, 2, 4, 5], [0, 1, 0, 1, 2], [1, 2, 1, 2, 5]
    with pytest.raises(RuntimeError):
        oracle.factorize(n, cp, ri, vals, [0, 1, 2], 3, 1.0)


def _edge_systems():
    """Degenerate shapes: 1 x 1, diagonal, a permutation matrix, dense, explicit zeros among the entries,
    a zero right-hand side next to a wide one."""
    big = (1 << 70) + 3
    yield "1x1", 1, [0, 1], [0], [-7], [[5, 0]]
    yield "diag", 4, [0, 1, 2, 3, 4], [0, 1, 2, 3], [3, -2, big, 1], [[1, 0], [2, 0], [3, 0], [-4, 0]]
    yield "perm", 4, [0, 1, 2, 3, 4], [2, 0, 3, 1], [1, 1, 1, 1], [[1, big], [2, -big], [3, 7], [4, 0]]
    n = 6
    dense_vals = [((i * 7 + j * 13) % 11) - 5 + (3 if i == j else 0) for j in range(n) for i in range(n)]
    yield "dense", n, [n * j for j in range(n + 1)], [i for _ in range(n) for i in range(n)], dense_vals, \
        [[i + 1, -i] for i in range(n)]
    # explicit zeros stored in the pattern (the reference keeps them as entries)
    yield "zeros", 4, [0, 2, 5, 7, 9], [0, 1, 0, 1, 2, 2, 3, 0, 3], [2, 0, 1, 3, 0, 5, 1, 0, 4], \
        [[1, 0], [0, 0], [-3, 0], [2, 0]]


@pytest.mark.parametrize("case", list(_edge_systems()), ids=lambda c: c[0])
def test_oracle_matches_reference_on_degenerate_shapes(oracle, reference, case):
    name, n, cp, ri, vals, b = case
    q = list(range(n))
    for pivot in (capi.SLIP_SMALLEST, capi.SLIP_DIAGONAL, capi.SLIP_TOL_SMALLEST, capi.SLIP_LARGEST):
        try:
            want = cases.run_library(reference, n, cp, ri, vals, b, q, pivot, None)
        except capi.SlipError as e:
            assert e.code == capi.SLIP_SINGULAR
            with pytest.raises(RuntimeError):
                cases.run_oracle(oracle, n, cp, ri, vals, b, q, pivot, 1.0)
            continue
        o = reference.default_options(pivot=pivot, order=capi.SLIP_NO_ORDERING)
        got = cases.run_oracle(oracle, n, cp, ri, vals, b, q, pivot, o.contents.tol)
        reference.free_options(o)
        cases.assert_same_factorization(got, want, f"{name} pivot {pivot}")
        assert got["x"] == want["x"], f"{name} pivot {pivot}: x differs"
