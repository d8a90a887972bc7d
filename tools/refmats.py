"""Dev tool: run a library exporting the SLIP_LU interface on matrix files in the reference's triplet
text format (ExampleMats/*_mat.txt, BasisLIB_ALL/RHS/*.mat) and print times and size statistics.

    python tools/refmats.py [--lib ref|b200] [--factors] MAT [RHS] ...
"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slip_lu_b200 import capi  # noqa: E402


def read_triplets(path):
    with open(path) as f:
        tok = f.read().split()
    m, n, nz = int(tok[0]), int(tok[1]), int(tok[2])
    I = [int(t) for t in tok[3:3 + 3 * nz:3]]
    J = [int(t) for t in tok[4:4 + 3 * nz:3]]
    X = [int(t) for t in tok[5:5 + 3 * nz:3]]
    dec = 0 if min(I[0], J[0]) == 0 else 1      # the demo reader's rule (Demo/demos.c:291-300)
    return n, [i - dec for i in I], [j - dec for j in J], X


def read_rhs(path, n):
    with open(path) as f:
        tok = f.read().split()
    m, k = int(tok[0]), int(tok[1])
    vals = [int(t) for t in tok[2:2 + m * k]]
    return [vals[r * k:(r + 1) * k] for r in range(m)]


def hadamard_bits(n, J, X):
    col = [0] * n
    for j, x in zip(J, X):
        col[j] += x * x
    return sum(0.5 * math.log2(c) for c in col if c > 0)


def main():
    args = sys.argv[1:]
    which = "ref"
    factors = False
    files = []
    while args:
        a = args.pop(0)
        if a == "--lib":
            which = args.pop(0)
        elif a == "--factors":
            factors = True
        else:
            files.append(a)
    if which == "ref":
        from oracle import binding as ob
        lib = capi.SlipLib(ob.REF_SO)
    else:
        import __graft_entry__ as entry
        entry.build()
        import slip_lu_b200
        lib = slip_lu_b200.lib()
    for mat in files:
        rhs = None
        for cand in (mat.replace("_mat.txt", "_v.txt"), mat + ".rhs"):
            if cand != mat and os.path.exists(cand):
                rhs = cand
        n, I, J, X = read_triplets(mat)
        b = read_rhs(rhs, n) if rhs else [[1] for _ in range(n)]
        A = lib.sparse_from_triplets(n, I, J, X)
        B = lib.dense_from_rows(b)
        o = lib.default_options()
        t = time.time(); S = lib.analyze(A, o); ta = time.time() - t
        info = f"{os.path.basename(mat)} n={n} nnz={len(X)} hadamard_bits={hadamard_bits(n, J, X):.0f} analyze={ta:.3f}s"
        if factors or which == "ref":
            t = time.time(); L, U, rhos, pinv = lib.factorize(A, S, o); tf = time.time() - t
            t = time.time(); x = lib.lu_solve(B, rhos, L, U, pinv); ts = time.time() - t
            Lc, Uc = L.contents, U.contents
            mb = 0
            for M in (Lc, Uc):
                for k in range(M.nz):
                    s = abs(M.x[k]._mp_size)
                    if s:
                        mb = max(mb, 64 * (s - 1) + int(M.x[k]._mp_d[s - 1]).bit_length())
            det = abs(capi.mpz_to_int(rhos[n - 1])).bit_length()
            Lp = [Lc.p[k] for k in range(n + 1)]
            upd = 0
            for m_ in range(Uc.nz):
                j = Uc.i[m_]
                upd += Lp[j + 1] - Lp[j] - 1
            xb = 0
            for r in range(n):
                nu, de = capi.mpq_to_pair(x[r][0])
                xb = max(xb, abs(nu).bit_length(), de.bit_length())
            info += (f" factor={tf:.3f}s solve={ts:.3f}s nnzL={Lc.nz} nnzU={Uc.nz} max_entry_bits={mb} det_bits={det} "
                     f"x_bits={xb} updates~{upd}")
        else:
            t = time.time(); x = lib.solve_mpq(A, S, B, o); tf = time.time() - t
            info += f" solve_mpq={tf:.3f}s check={lib.dll.SLIP_check_solution(A, x, B)}"
        print(info, flush=True)


if __name__ == "__main__":
    main()
