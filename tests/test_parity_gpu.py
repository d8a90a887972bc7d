"""GPU: the CUDA path, through the C ABI, against the oracle and the golden fixtures."""
import ctypes as C
import os
import time

import pytest

from slip_lu_b200 import capi, synth
import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", cases.small_cases(), ids=lambda c: c[0])
def test_factor_and_solve_match_oracle(gpu, oracle, case):
    name, n, cp, ri, vals, b = case
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    got = cases.run_library(gpu, n, cp, ri, vals, b, q)
    cases.assert_same_factorization(got, want, name)


@pytest.mark.parametrize("pivot,tol", [(0, None), (1, None), (2, None), (3, 1.0), (3, 0.3), (3, 1e-4),
                                        (4, 0.5), (4, 0.01), (5, None)])
def test_every_pivot_rule(gpu, oracle, pivot, tol):
    for seed in (1, 2, 3):
        n, cp, ri, vals, b = synth.random_sparse(50, 5, 20, seed=10 * pivot + seed, nrhs=2)
        q = cases.colamd_like_order(n, cp, ri)
        t = 1.0 if tol is None else tol
        want = cases.run_oracle(oracle, n, cp, ri, vals, b, q, pivot, t)
        got = cases.run_library(gpu, n, cp, ri, vals, b, q, pivot, tol)
        cases.assert_same_factorization(got, want, f"pivot {pivot} tol {tol} seed {seed}")


@pytest.mark.parametrize("name", cases.GOLDEN_FACTOR_CASES)
def test_matches_reference_golden(gpu, oracle, name):
    """Bit-exact L, U, rhos, pinv and x against fixtures produced by the unmodified reference."""
    g = cases.load_golden(name)
    n, cp, ri, vals, b = cases.golden_system(g)
    got = cases.run_library(gpu, n, cp, ri, vals, b, g["q"], g["options"]["pivot"], g["options"]["tol"])
    cases.check_against_golden(g, got, ob=oracle)


@pytest.mark.parametrize("name", ["10teams_default", "testmat_pivot3", "decimal80_multirhs", "lp300_colamd"])
def test_solve_mpq_end_to_end(gpu, name):
    """SLIP_LU_analyze + SLIP_solve_mpq (the reference demo's call sequence, Demo/example2.c)."""
    g = cases.load_golden(name)
    n, cp, ri, vals, b = cases.golden_system(g)
    o = gpu.default_options(pivot=g["options"]["pivot"], order=g["options"]["order"], tol=g["options"]["tol"])
    A = gpu.sparse_from_csc(n, cp, ri, vals)
    B = gpu.dense_from_rows(b)
    from conftest import have_ordering_library
    if have_ordering_library():
        S = gpu.analyze(A, o)
        assert [S.contents.q[k] for k in range(n)] == g["q"]
    else:
        S = gpu.analyze(A, o, q=g["q"])
    x = gpu.solve_mpq(A, S, B, o)
    got = gpu.mpq_mat_to_py(x, n, len(b[0]))
    assert cases.digest_pairs(got) == int(g["digests"]["x_solve_mpq"])
    if g["explicit"]:
        assert got == [[(int(a), int(d)) for a, d in row] for row in g["x_solve_mpq"]]
    assert gpu.dll.SLIP_check_solution(A, x, B) == 0      # exact residual, rational arithmetic


def test_singular_and_edge_inputs(gpu):
    o = gpu.default_options(order=capi.SLIP_NO_ORDERING)
    # duplicate columns -> numerically singular
    A = gpu.sparse_from_csc(3, [0, 2, 4, 5], [0, 1, 0, 1, 2], [1, 2, 1, 2, 5])
    S = gpu.analyze(A, o)
    with pytest.raises(capi.SlipError) as e:
        gpu.factorize(A, S, o)
    assert e.value.code == capi.SLIP_SINGULAR
    # the zero sits in a column with a single candidate, which the host does not wait for: the
    # device reports it with the next column; same answer through SLIP_solve_mpq and with the
    # no-wait path switched off
    B = gpu.dense_from_rows([[1], [2], [3]])
    for single in ("1", "0"):
        os.environ["SLIP_B200_SINGLE"] = single
        try:
            with pytest.raises(capi.SlipError) as e:
                gpu.solve_mpq(A, S, B, o)
            assert e.value.code == capi.SLIP_SINGULAR
            with pytest.raises(capi.SlipError) as e:
                gpu.factorize(A, S, o)
            assert e.value.code == capi.SLIP_SINGULAR
        finally:
            del os.environ["SLIP_B200_SINGLE"]
    # structurally singular: empty row
    A = gpu.sparse_from_csc(3, [0, 1, 2, 3], [0, 0, 2], [1, 2, 3])
    S = gpu.analyze(A, o)
    with pytest.raises(capi.SlipError) as e:
        gpu.factorize(A, S, o)
    assert e.value.code == capi.SLIP_SINGULAR
    # explicit zeros in the input and a zero right-hand side
    n, cp, ri, vals = 3, [0, 2, 4, 6], [0, 1, 1, 2, 0, 2], [3, 0, 5, 0, 0, -2]
    got = cases.run_library(gpu, n, cp, ri, vals, [[0], [0], [0]], [0, 1, 2])
    assert got["rhos"] == [3, 15, -30] and all(v == (0, 1) for row in got["x"] for v in row)


def test_standalone_solve_uploads_factors(gpu, oracle):
    """SLIP_LU_solve on factors that are not resident on the GPU (here: built by the CPU oracle and
    handed over as plain mpz_t matrices) takes the upload path and still matches."""
    n, cp, ri, vals, b = synth.random_sparse(40, 5, 28, seed=21, nrhs=3, rhs_bits=200)
    q = cases.colamd_like_order(n, cp, ri)
    f = oracle.factorize(n, cp, ri, vals, q)
    want = oracle.solve(f, b)
    Lp, Li, Lx = f.L_py(); Up, Ui, Ux = f.U_py()
    L = gpu.sparse_from_csc(n, Lp, Li, Lx)
    U = gpu.sparse_from_csc(n, Up, Ui, Ux)
    rhos = gpu._mpz_array(f.rhos_py())
    pinv = (C.c_int32 * n)(*f.pinv_py())
    B = gpu.dense_from_rows(b)
    x = gpu.lu_solve(B, rhos, L, U, pinv)
    assert gpu.mpq_mat_to_py(x, n, 3) == want


def test_large_rhs_widens_channels(gpu, oracle):
    """A right-hand side far larger than A's entries: the resident factors are too narrow and the
    solve re-encodes them with more channels."""
    n, cp, ri, vals, b = synth.random_sparse(30, 4, 16, seed=33, nrhs=1, rhs_bits=2000)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    got = cases.run_library(gpu, n, cp, ri, vals, b, q)
    cases.assert_same_factorization(got, want, "large rhs")


def test_bench_scale_properties(gpu):
    """BASELINE configs[1] family at a size the CPU oracle cannot reach in test time: the factors
    are checked through size-independent properties computed exactly on the host:
      * A x = b exactly (SLIP_check_solution, rational arithmetic) for the solution of SLIP_solve_mpq
      * det from SLIP_LU_factorize equals the last pivot and divides det * x consistently
      * L and U are triangular in the pivot order with U's diagonal equal to rhos
      * factorize+LU_solve and solve_mpq agree (two different device paths)."""
    n, cp, ri, vals, b = synth.random_sparse(400, 6, 32, seed=77, nrhs=2)
    q = cases.colamd_like_order(n, cp, ri)
    o = gpu.default_options(order=capi.SLIP_NO_ORDERING)
    A = gpu.sparse_from_csc(n, cp, ri, vals)
    B = gpu.dense_from_rows(b)
    S = gpu.analyze(A, o, q=q)
    t0 = time.time()
    L, U, rhos, pinv = gpu.factorize(A, S, o)
    x1 = gpu.lu_solve(B, rhos, L, U, pinv)
    t1 = time.time()
    Up, Ui, Ux = gpu.sparse_to_py(U)
    Lp, Li, _ = gpu.sparse_to_py(L)
    rh = gpu.mpz_array_to_py(rhos, n)
    for k in range(n):
        assert Ui[Up[k + 1] - 1] == k and Ux[Up[k + 1] - 1] == rh[k]
        assert all(i <= k for i in Ui[Up[k]:Up[k + 1]]) and all(i >= k for i in Li[Lp[k]:Lp[k + 1]])
    assert sorted(pinv) == list(range(n)) and all(v != 0 for v in rh)
    assert gpu.dll.SLIP_permute_x(x1, n, 2, S) == 0
    x2 = gpu.solve_mpq(A, S, B, o)
    assert gpu.mpq_mat_to_py(x1, n, 2) == gpu.mpq_mat_to_py(x2, n, 2)
    assert gpu.dll.SLIP_check_solution(A, x2, B) == 0
    print(f"n=400 factor+solve (with host factors) {t1 - t0:.2f}s, nnz(L)={Lp[-1]}, det bits={abs(rh[-1]).bit_length()}")


def test_many_right_hand_sides(gpu, oracle):
    """configs[3] family: one factorization, many right-hand sides (batched by the solve kernels)."""
    n, cp, ri, vals, b = synth.random_sparse(48, 5, 24, seed=41, nrhs=37, rhs_bits=40)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    got = cases.run_library(gpu, n, cp, ri, vals, b, q)
    cases.assert_same_factorization(got, want, "37 rhs")
    # the same through SLIP_solve_mpq with a tiny batch budget (forces several batches)
    os.environ["SLIP_B200_SOLVE_BATCH_MB"] = "1"
    try:
        o = gpu.default_options(order=capi.SLIP_NO_ORDERING)
        A = gpu.sparse_from_csc(n, cp, ri, vals); B = gpu.dense_from_rows(b)
        S = gpu.analyze(A, o, q=q)
        x = gpu.solve_mpq(A, S, B, o)
        got2 = gpu.mpq_mat_to_py(x, n, 37)
    finally:
        del os.environ["SLIP_B200_SOLVE_BATCH_MB"]
    want2 = [None] * n
    for i in range(n):
        want2[q[i]] = want["x"][i]
    assert got2 == want2


def test_batch_of_independent_systems(gpu, oracle):
    """configs[4] family: a batch of LP-basis style systems, solved through the sharding helper
    (world size 1 here; the partitioning itself is covered on CPU with gloo)."""
    from slip_lu_b200.sharding import solve_batch_sharded
    systems = [synth.lp_basis(120, seed=s, nrhs=1) for s in range(12)]
    got = solve_batch_sharded(gpu, systems, 1, 0,
                              options=lambda: gpu.default_options(order=capi.SLIP_NO_ORDERING))
    # several systems in flight at once (one CUDA stream per session, host threads)
    got_mt = solve_batch_sharded(gpu, systems, 1, 0, threads=4,
                                 options=lambda: gpu.default_options(order=capi.SLIP_NO_ORDERING))
    assert got_mt == got
    for g, x in got:
        n, cp, ri, vals, b = systems[g]
        f = oracle.factorize(n, cp, ri, vals, list(range(n)))
        assert x == oracle.solve(f, b), f"system {g}"


def test_laplacian_pattern_64bit(gpu, oracle):
    """configs[2] family at a size the oracle still reaches: 16x16 grid, 64-bit entries."""
    n, cp, ri, vals, b = synth.laplacian_2d(16, 64, seed=6, nrhs=1)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    got = cases.run_library(gpu, n, cp, ri, vals, b, q)
    cases.assert_same_factorization(got, want, "lap256")
    assert abs(got["rhos"][-1]).bit_length() > 15000


def test_double_input_end_to_end(gpu, oracle):
    """double input -> integer system (SLIP_build_*_double) -> SLIP_solve_double, against the
    oracle run on the same integer system and scaled back."""
    from fractions import Fraction
    (sysd, dvals) = synth.decimal_scaled(40, 4, 5, seed=13, nrhs=2)
    n, cp, ri, _, b = sysd
    o = gpu.default_options(order=capi.SLIP_NO_ORDERING)
    A = gpu.dll.SLIP_create_sparse()
    nz = len(dvals)
    assert gpu.dll.SLIP_build_sparse_ccf_double(A, (C.c_int32 * (n + 1))(*cp), (C.c_int32 * nz)(*ri),
                                                (C.c_double * nz)(*dvals), n, nz, o) == 0
    _, _, ivals = gpu.sparse_to_py(A)
    sa = Fraction(*capi.mpq_to_pair(A.contents.scale))
    B = gpu.dense_from_rows(b)
    S = gpu.analyze(A, o)
    xd = gpu.dll.SLIP_create_double_mat(n, 2)
    assert gpu.dll.SLIP_solve_double(xd, A, S, B, o) == 0
    f = oracle.factorize(n, cp, ri, ivals, list(range(n)))
    want = oracle.solve(f, b)
    for i in range(n):
        for c in range(2):
            exact = Fraction(*want[i][c]) * sa           # x of the original (double) system
            # mpq_get_d truncates where float() rounds: allow one unit in the last place
            assert abs(xd[i][c] - float(exact)) <= abs(float(exact)) * 2.0 ** -51, (i, c)


def test_channel_prime_dividing_a_pivot_is_retired(gpu, oracle):
    """The first channel prime is 2^31-1.  A matrix whose first pivot IS that prime makes the pivot
    vanish in channel 0; the factorization must notice, retire the prime and still be exact."""
    p0 = 2 ** 31 - 1
    n, cp, ri, vals = 3, [0, 2, 4, 6], [0, 1, 1, 2, 0, 2], [p0, 3, 5, 7, 11, 13]
    b = [[1], [2], [3]]
    q = [0, 1, 2]
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q, capi.SLIP_LARGEST, 1.0)
    assert want["rhos"][0] == p0
    got = cases.run_library(gpu, n, cp, ri, vals, b, q, capi.SLIP_LARGEST)
    cases.assert_same_factorization(got, want, "pivot equal to a channel prime")
    # and the library keeps working afterwards (tables rebuilt without the retired prime)
    n, cp, ri, vals, b = synth.random_sparse(30, 4, 20, seed=77, nrhs=1)
    q = cases.colamd_like_order(n, cp, ri)
    cases.assert_same_factorization(cases.run_library(gpu, n, cp, ri, vals, b, q),
                                    cases.run_oracle(oracle, n, cp, ri, vals, b, q), "after retirement")


@pytest.mark.parametrize("env", [{"SLIP_B200_CH": "4"}, {"SLIP_B200_CH": "8"}, {"SLIP_B200_CH": "4", "SLIP_B200_X_GLOBAL": "1"},
                                 {"SLIP_B200_CH": "4", "SLIP_B200_CPT": "2"}, {"SLIP_B200_CANON_GMP": "1"},
                                 {"SLIP_B200_LOOKAHEAD": "0"}, {"SLIP_B200_LOOKAHEAD": "3", "SLIP_B200_LOOK_MIN": "0"},
                                 {"SLIP_B200_SINGLE": "0"}, {"SLIP_B200_LOOKAHEAD": "6", "SLIP_B200_LOOK_STEPS": "1"},
                                 {"SLIP_B200_BACKSUB": "0"}, {"SLIP_B200_BACKSUB": "0", "SLIP_B200_CH": "4"},
                                 {"SLIP_B200_CH": "16"}, {"SLIP_B200_CH": "32"}, {"SLIP_B200_X_GLOBAL": "1"},
                                 {"SLIP_B200_CH": "32", "SLIP_B200_X_GLOBAL": "1"}, {"SLIP_B200_GARNER": "1"},
                                 {"SLIP_B200_GARNER": "0"}, {"SLIP_B200_CPT": "2"}, {"SLIP_B200_CPT": "4"},
                                 {"SLIP_B200_GARNER_E": "1"}, {"SLIP_B200_GARNER_E": "3"}, {"SLIP_B200_GARNER_E": "5"},
                                 {"SLIP_B200_GARNER_E": "6"}, {"SLIP_B200_OVERLAP": "0"}, {"SLIP_B200_FRAC": "0"}, {"SLIP_B200_FRAC_MARGIN": "0"}, {"SLIP_B200_PRUNE": "0"}, {"SLIP_B200_SPEC": "1"},
                                 {"SLIP_B200_CPT": "2", "SLIP_B200_CH": "32", "SLIP_B200_X_GLOBAL": "1"}])
def test_kernel_variants_agree(gpu, oracle, env):
    """The kernel configurations that large problems select automatically (wider channel blocks,
    work vector in global memory when the pattern outgrows shared memory, the other Garner
    kernels) are forced here on a small case and must give the same bits."""
    n, cp, ri, vals, b = synth.random_sparse(72, 6, 40, seed=17, nrhs=2)
    q = cases.colamd_like_order(n, cp, ri)
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        got = cases.run_library(gpu, n, cp, ri, vals, b, q)
        o = gpu.default_options(order=capi.SLIP_NO_ORDERING)
        A = gpu.sparse_from_csc(n, cp, ri, vals); B = gpu.dense_from_rows(b)
        S = gpu.analyze(A, o, q=q)
        x = gpu.solve_mpq(A, S, B, o)
        assert gpu.dll.SLIP_check_solution(A, x, B) == 0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    cases.assert_same_factorization(got, want, str(env))


_MULTI_CHUNK = {}


@pytest.mark.parametrize("cpt", ["4", "2"])
def test_steps_spanning_several_chunks(gpu, oracle, cpt):
    """With 32-channel blocks a pipeline chunk holds 128 rows, so the dense trailing columns of this
    case are eliminated in several chunks per step (first/last-chunk flags, descriptor ring wrap)."""
    n, cp, ri, vals, b = synth.random_sparse(300, 9, 24, seed=5, nrhs=1)
    q = cases.colamd_like_order(n, cp, ri)
    if "want" not in _MULTI_CHUNK:
        _MULTI_CHUNK["want"] = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    want = _MULTI_CHUNK["want"]
    env = {"SLIP_B200_CH": "32", "SLIP_B200_CPT": cpt}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        got = cases.run_library(gpu, n, cp, ri, vals, b, q)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    cases.assert_same_factorization(got, want, f"multi-chunk cpt={cpt}")
    lp = want["L"][0]
    assert max(lp[k + 1] - lp[k] for k in range(n)) > 128, "case too sparse to span two chunks"


def test_solve_mpfr_matches_reference_golden(gpu):
    """SLIP_solve_mpfr (SLIP_solve_mpfr.c): the exact solution rounded to option->prec bits, value for
    value what the reference returned for the same system (fixture mpfr_builders.json)."""
    g = cases.load_golden("mpfr_builders")
    for case in g["cases"]:
        sv = case["solve"]
        n, cp, ri, vals, b = synth.random_sparse(sv["n"], sv["nnz_per_col"], sv["bits"], seed=sv["seed"], nrhs=sv["nrhs"])
        o = gpu.default_options(order=capi.SLIP_NO_ORDERING)
        o.contents.prec = case["prec"]
        A = gpu.sparse_from_csc(n, cp, ri, vals); B = gpu.dense_from_rows(b)
        S = gpu.analyze(A, o)
        X = gpu.dll.SLIP_create_mpfr_mat(n, sv["nrhs"], o)
        assert gpu.dll.SLIP_solve_mpfr(X, A, S, B, o) == 0
        got = [[list(capi.mpfr_to_pair(X[r][c])) for c in range(sv["nrhs"])] for r in range(n)]
        want = [[[int(v) for v in pair] for pair in row] for row in sv["x"]]
        assert got == want, case["prec"]


@pytest.mark.parametrize("n", [513, 1030])
def test_long_steps_cross_chunk_boundaries(gpu, oracle, n):
    """Elimination steps of n, n-1, ... rows with the default 8-channel blocks (512-row chunks): one
    full chunk plus a one-row tail at n = 513, two full chunks plus a short tail at n = 1030."""
    n, cp, ri, vals, b = synth.dense_head(n, head=5, bits=12, seed=n)
    q = list(range(n))
    want = cases.run_oracle(oracle, n, cp, ri, vals, b, q)
    got = cases.run_library(gpu, n, cp, ri, vals, b, q)
    cases.assert_same_factorization(got, want, f"dense head n={n}")


@pytest.mark.parametrize("env", [{"SLIP_B200_FRAC_VERIFY": "1"},
                                 {"SLIP_B200_FRAC_VERIFY": "1", "SLIP_B200_FRAC_MARGIN": "0"}])
@pytest.mark.parametrize("pivot", [capi.SLIP_SMALLEST, capi.SLIP_TOL_SMALLEST, capi.SLIP_LARGEST, capi.SLIP_TOL_LARGEST])
@pytest.mark.parametrize("case", ["random30bit", "laplacian_small_ints"])
def test_approximate_pivot_search_is_exact(gpu, pivot, env, case):
    """SLIP_solve_mpq searches the pivots on approximate magnitudes (fractional CRT) and accepts a
    choice only when it is proven.  With SLIP_B200_FRAC_VERIFY the library re-runs every accepted
    choice through the exact reconstruction + scan and fails on any difference (slot, diagonal
    flags, sign).  With no margin of words the columns first fail for lack of precision and are
    repeated with more words; the small-integer Laplacian is full of ties, which go to the exact
    scan.  The solution must satisfy A x = b exactly on every route."""
    if case == "random30bit":
        n, cp, ri, vals, b = synth.random_sparse(90, 6, 30, seed=41, nrhs=2)
    else:
        n, cp, ri, vals, b = synth.laplacian_2d(12, 8, seed=6, nrhs=2, rhs_bits=8)
    q = cases.colamd_like_order(n, cp, ri)
    tol = 0.3 if pivot in (capi.SLIP_TOL_SMALLEST, capi.SLIP_TOL_LARGEST) else None
    # the approximate search is normally reserved for sessions with hundreds of channels: force it here
    env = dict(env, SLIP_B200_FRAC_MIN_S="16", SLIP_B200_ADAPTIVE="0")
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        o = gpu.default_options(pivot=pivot, order=capi.SLIP_NO_ORDERING, tol=tol)
        A = gpu.sparse_from_csc(n, cp, ri, vals); B = gpu.dense_from_rows(b)
        S = gpu.analyze(A, o, q=q)
        x = gpu.solve_mpq(A, S, B, o)
        assert gpu.dll.SLIP_check_solution(A, x, B) == 0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_reference_demo_program_runs_unchanged(gpu, tmp_path):
    """Demo/example2.c of the reference (BASELINE configs[0]: read a triplet matrix and a dense
    right-hand side, SLIP_LU_analyze, SLIP_solve_mpq), compiled unmodified and linked against
    libslip_lu_b200.so by `make -C oracle demo`, run on the 10teams system written out from the
    golden fixture in the demo's own file formats."""
    import subprocess
    from oracle import binding as ob
    if not os.path.exists(ob.DEMO_EXE):
        pytest.skip("oracle/_ref/demo_example2 was not built (needs the reference tree at build time)")
    g = cases.load_golden("10teams_default")
    n, cp, ri, vals, b = cases.golden_system(g)
    mat, rhs = tmp_path / "mat.txt", tmp_path / "rhs.txt"
    with open(mat, "w") as f:
        f.write(f"{n} {n} {cp[n]}\n")
        for j in range(n):
            for a in range(cp[j], cp[j + 1]):
                f.write(f"{ri[a] + 1} {j + 1} {vals[a]}\n")          # 1-based triplets
    with open(rhs, "w") as f:
        f.write(f"{n} {len(b[0])}\n")
        for row in b:
            f.write(" ".join(str(v) for v in row) + "\n")
    out = subprocess.run([ob.DEMO_EXE, str(mat), str(rhs)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "all tests passed" in out.stdout, out.stdout + out.stderr
    assert "SLIP LU Factor & Solve time" in out.stdout
