/* oracle/ref_oracle.c -- CPU restatement of the SLIP LU factor/solve path.  TEST INFRASTRUCTURE.
 *
 * This file is the checker for slip_lu_b200's CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (slip_lu_b200/) never links, loads or calls it.
 *
 * It restates, over GMP integers, the algorithm of the reference's path:
 *   - left-looking REF LU, one sparse REF triangular solve per column
 *       (reference: SLIP_LU/Source/SLIP_LU_factorize.c:193-270,
 *                   SLIP_LU/Source/slip_REF_triangular_solve.c:84-262)
 *   - symbolic reach over the pattern of L and the ordering of the pattern by the
 *     current row permutation (slip_reach.c:18-52, slip_dfs.c:19-77, slip_sort_xi.c:24-47)
 *   - the six pivoting rules (slip_get_pivot.c:46-175, slip_get_smallest_pivot.c,
 *     slip_get_largest_pivot.c, slip_get_nonzero_pivot.c)
 *   - REF forward substitution, scaling by det, back substitution and division by det
 *       (SLIP_LU_solve.c:74-91, slip_forward_sub.c:64-155, slip_back_sub.c:30-56,
 *        slip_array_mul.c, slip_array_div.c)
 *
 * PARITY PIN: the reference ships no golden output vectors for this path (its demos only
 * print timings and run SLIP_check_solution), so this oracle is pinned against outputs of
 * the reference itself: oracle/_ref/libslip_ref.so (the unmodified reference compiled by
 * oracle/Makefile) in tests/test_oracle.py, against the committed fixtures
 * tests/golden/*.json generated from that library by tests/golden/make_golden.py, and at size
 * against the reference's digests of L, U, rhos and pinv on its own ExampleMats / BasisLIB
 * matrices (tests/golden/refmats.json, tests/test_refmats.py).
 *
 * The REF update is written once in its closed form
 *     x_i <- ( rho_j * rho_{j-1}/rho_{h_i} * x_i  -  l_ij * x_j ) / rho_{j-1},   rho_{-1} = 1
 * where h_i is the last elimination step applied to x_i ("history"); the reference spells
 * the same arithmetic as separate branches on x_i == 0 / h_i (slip_REF_triangular_solve.c:
 * 150-232).  All divisions are exact.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <gmp.h>

#define RO_OK 0
#define RO_OUT_OF_MEMORY (-1)
#define RO_SINGULAR (-2)
#define RO_INCORRECT_INPUT (-3)

typedef struct
{
    int n;       /* columns */
    int nz;      /* entries */
    int *p;      /* n+1 column pointers */
    int *i;      /* row indices */
    mpz_t *x;    /* values */
} ro_csc;

static int csc_alloc (ro_csc *M, int n, int cap)
{
    M->n = n; M->nz = 0;
    M->p = (int *) calloc ((size_t) n + 1, sizeof (int));
    M->i = (int *) malloc ((size_t) cap * sizeof (int));
    M->x = (mpz_t *) malloc ((size_t) cap * sizeof (mpz_t));
    return (M->p && M->i && M->x) ? RO_OK : RO_OUT_OF_MEMORY;
}

static int csc_reserve (ro_csc *M, int *cap, int need)
{
    if (need <= *cap) return RO_OK;
    int ncap = *cap;
    while (ncap < need) ncap = 2 * ncap + 16;
    int *ni = (int *) realloc (M->i, (size_t) ncap * sizeof (int));
    if (!ni) return RO_OUT_OF_MEMORY;
    M->i = ni;
    mpz_t *nx = (mpz_t *) realloc (M->x, (size_t) ncap * sizeof (mpz_t));
    if (!nx) return RO_OUT_OF_MEMORY;
    M->x = nx;
    *cap = ncap;
    return RO_OK;
}

void ro_free_csc (ro_csc *M)
{
    if (!M) return;
    if (M->x) { for (int k = 0; k < M->nz; k++) mpz_clear (M->x[k]); free (M->x); }
    free (M->i); free (M->p);
    M->x = NULL; M->i = NULL; M->p = NULL; M->nz = 0;
}

/* ---------------------------------------------------------------------------------------
 * symbolic: rows reachable from the rows of A(:,col) in the graph of L.  A pivotal row r
 * (position pos = pinv[r] < k) has edges to every row stored in column pos of L.
 * Returns the count; pattern[] is then sorted by current position pinv[].
 * ------------------------------------------------------------------------------------- */
static int cmp_int (const void *a, const void *b)
{
    int x = *(const int *) a, y = *(const int *) b;
    return (x > y) - (x < y);
}

static int reach_sorted (int n, int k, const int *Ap, const int *Ai, int col,
                         const ro_csc *L, const int *pinv, const int *row_at,
                         int *mark, int stamp, int *stack, int *pattern)
{
    int cnt = 0;
    for (int a = Ap[col]; a < Ap[col + 1]; a++)
    {
        int r0 = Ai[a];
        if (mark[r0] == stamp) continue;
        int sp = 0;
        stack[sp++] = r0; mark[r0] = stamp;
        while (sp > 0)
        {
            int r = stack[--sp];
            pattern[cnt++] = pinv[r];           /* store positions, translate back below */
            int pos = pinv[r];
            if (pos < k)
            {
                for (int m = L->p[pos]; m < L->p[pos + 1]; m++)
                {
                    int rr = L->i[m];
                    if (mark[rr] != stamp) { mark[rr] = stamp; stack[sp++] = rr; }
                }
            }
        }
    }
    qsort (pattern, (size_t) cnt, sizeof (int), cmp_int);
    for (int t = 0; t < cnt; t++) pattern[t] = row_at[pattern[t]];
    (void) n;
    return cnt;
}

/* x_i <- x_i * rho[to] / rho[from]   (rho[-1] == 1): bring x_i from level from+1 to level to+1 */
static void lift (mpz_t xi, mpz_t *rho, int from, int to)
{
    if (from >= to) return;
    mpz_mul (xi, xi, rho[to]);
    if (from >= 0) mpz_divexact (xi, xi, rho[from]);
}

/* ---------------------------------------------------------------------------------------
 * pivot choice among the not-yet-pivotal rows of the pattern, scanning in pattern order.
 * ------------------------------------------------------------------------------------- */
static int pick_extreme (int want_small, const int *pattern, int cnt, mpz_t *x, const int *pivotal)
{
    int best = -1;
    for (int t = 0; t < cnt; t++)
    {
        int r = pattern[t];
        if (pivotal[r] || mpz_sgn (x[r]) == 0) continue;
        if (best < 0) { best = r; continue; }
        int c = mpz_cmpabs (x[r], x[best]);
        if (want_small ? (c < 0) : (c > 0)) best = r;     /* strict: first of equals wins */
    }
    return best;
}

static int pick_first_nonzero (const int *pattern, int cnt, mpz_t *x, const int *pivotal)
{
    for (int t = 0; t < cnt; t++)
    {
        int r = pattern[t];
        if (!pivotal[r] && mpz_sgn (x[r]) != 0) return r;
    }
    return -1;
}

static int choose_pivot (int scheme, double tol, int col, const int *pattern, int cnt,
                         mpz_t *x, const int *pivotal)
{
    int diag_ok = (!pivotal[col] && mpz_sgn (x[col]) != 0);
    int piv;
    switch (scheme)
    {
        case 0: return pick_extreme (1, pattern, cnt, x, pivotal);
        case 1: return diag_ok ? col : pick_extreme (1, pattern, cnt, x, pivotal);
        case 2: return pick_first_nonzero (pattern, cnt, x, pivotal);
        case 3:
        case 4:
        {
            piv = pick_extreme (scheme == 3, pattern, cnt, x, pivotal);
            if (piv < 0 || !diag_ok) return piv;
            /* scheme 3: |smallest| / |diag| >= tol.
             * scheme 4: the reference forms diag / largest and takes mpq_abs of the
             * non-canonical fraction (slip_get_pivot.c:131-137), which leaves the sign of the
             * denominator in place: the operand handed to mpq_cmp is |diag| / largest with a
             * SIGNED denominator.  That is reproduced literally here (same GMP call on the
             * same operand) because it decides the pivot whenever the largest entry is
             * negative. */
            mpq_t ratio, t;
            mpq_init (ratio); mpq_init (t);
            if (scheme == 3)
            {
                mpz_abs (mpq_numref (ratio), x[piv]);
                mpz_abs (mpq_denref (ratio), x[col]);
            }
            else
            {
                mpz_abs (mpq_numref (ratio), x[col]);
                mpz_set (mpq_denref (ratio), x[piv]);
            }
            mpq_set_d (t, tol);
            int ge = (mpq_cmp (ratio, t) >= 0);
            mpq_clear (ratio); mpq_clear (t);
            return ge ? col : piv;
        }
        default: return pick_extreme (0, pattern, cnt, x, pivotal);
    }
}

/* ---------------------------------------------------------------------------------------
 * factorization  P A Q = L D^-1 U
 *   Ap/Ai/Ax : CSC input (Ax: array of n..nz mpz_t), q: column order (length n)
 *   L, U     : outputs (allocated here); row indices are final positions (pinv applied)
 *   rhos     : n initialised mpz_t, receives the pivots; pinv: n ints
 * ------------------------------------------------------------------------------------- */
/* bench.py's bounded sample of a workload the CPU cannot finish: only the first ro_column_limit
 * columns are factorized (a left-looking factorization never looks at later columns, so they are
 * exactly the first columns of the complete job); 0 = all.  L and U then hold those columns. */
static int ro_column_limit = 0;
void ro_set_column_limit (int m) { ro_column_limit = m > 0 ? m : 0; }

int ro_factorize (int n, const int *Ap, const int *Ai, mpz_t *Ax, const int *q,
                  int scheme, double tol, ro_csc *L, ro_csc *U, mpz_t *rhos, int *pinv)
{
    if (n <= 0 || !Ap || !Ai || !Ax || !q || !L || !U || !rhos || !pinv) return RO_INCORRECT_INPUT;
    int status = RO_OK;
    int lcap = 16 * n + 64, ucap = 16 * n + 64;
    memset (L, 0, sizeof (*L)); memset (U, 0, sizeof (*U));
    if (csc_alloc (L, n, lcap) || csc_alloc (U, n, ucap)) return RO_OUT_OF_MEMORY;

    int *row_at = (int *) malloc ((size_t) n * sizeof (int));    /* inverse of pinv */
    int *pivotal = (int *) calloc ((size_t) n, sizeof (int));
    int *hist = (int *) malloc ((size_t) n * sizeof (int));
    int *mark = (int *) calloc ((size_t) n, sizeof (int));
    int *stack = (int *) malloc ((size_t) n * sizeof (int));
    int *pattern = (int *) malloc ((size_t) n * sizeof (int));
    mpz_t *x = (mpz_t *) malloc ((size_t) n * sizeof (mpz_t));
    if (!row_at || !pivotal || !hist || !mark || !stack || !pattern || !x) return RO_OUT_OF_MEMORY;
    for (int r = 0; r < n; r++) { mpz_init (x[r]); pinv[r] = r; row_at[r] = r; }

    const int kend = (ro_column_limit > 0 && ro_column_limit < n) ? ro_column_limit : n;
    for (int k = 0; k < kend && status == RO_OK; k++)
    {
        int col = q[k];
        L->p[k] = L->nz; U->p[k] = U->nz;
        if (csc_reserve (L, &lcap, L->nz + n) || csc_reserve (U, &ucap, U->nz + n))
        { status = RO_OUT_OF_MEMORY; break; }

        int cnt = reach_sorted (n, k, Ap, Ai, col, L, pinv, row_at, mark, k + 1, stack, pattern);

        /* numeric: x = A(:,col) on the pattern */
        for (int t = 0; t < cnt; t++) { mpz_set_ui (x[pattern[t]], 0); hist[pattern[t]] = -1; }
        mpz_set_ui (x[col], 0);   /* the diagonal is queried by the pivot rules even if absent */
        for (int a = Ap[col]; a < Ap[col + 1]; a++) mpz_set (x[Ai[a]], Ax[a]);

        for (int t = 0; t < cnt; t++)
        {
            int r = pattern[t];
            int j = pinv[r];
            if (j >= k)
            {   /* row not yet pivotal: only bring it to level k */
                lift (x[r], rhos, hist[r], k - 1);
                continue;
            }
            /* row r was pivot j: x[r] becomes U(j,k); eliminate with column j of L */
            lift (x[r], rhos, hist[r], j - 1);
            if (mpz_sgn (x[r]) == 0) continue;          /* nothing to eliminate (values unchanged) */
            for (int m = L->p[j]; m < L->p[j + 1]; m++)
            {
                int i = L->i[m];
                if (pinv[i] <= j || mpz_sgn (L->x[m]) == 0) continue;
                if (mpz_sgn (x[i]) != 0)
                {
                    lift (x[i], rhos, hist[i], j - 1);
                    mpz_mul (x[i], x[i], rhos[j]);
                }
                mpz_submul (x[i], L->x[m], x[r]);
                if (j >= 1) mpz_divexact (x[i], x[i], rhos[j - 1]);
                hist[i] = j;
            }
        }

        int piv = choose_pivot (scheme, tol, col, pattern, cnt, x, pivotal);
        if (piv < 0) { status = RO_SINGULAR; break; }
        {   /* move row piv to position k */
            int oldpos = pinv[piv], displaced = row_at[k];
            row_at[k] = piv; row_at[oldpos] = displaced;
            pinv[piv] = k; pinv[displaced] = oldpos;
            pivotal[piv] = 1;
            mpz_set (rhos[k], x[piv]);
        }
        for (int t = 0; t < cnt; t++)
        {
            int r = pattern[t], pos = pinv[r];
            if (pos <= k) { U->i[U->nz] = r; mpz_init_set (U->x[U->nz], x[r]); U->nz++; }
            if (pos >= k) { L->i[L->nz] = r; mpz_init_set (L->x[L->nz], x[r]); L->nz++; }
        }
    }
    if (status == RO_OK)
    {
        for (int k = kend; k <= n; k++) { L->p[k] = L->nz; U->p[k] = U->nz; }
        for (int m = 0; m < L->nz; m++) L->i[m] = pinv[L->i[m]];
        for (int m = 0; m < U->nz; m++) U->i[m] = pinv[U->i[m]];
    }
    else { ro_free_csc (L); ro_free_csc (U); }
    for (int r = 0; r < n; r++) mpz_clear (x[r]);
    free (x); free (pattern); free (stack); free (mark); free (hist); free (pivotal); free (row_at);
    return status;
}

/* ---------------------------------------------------------------------------------------
 * solve  L D^-1 U y = P b ;  xq = y  (n x nrhs rationals, still in factor column order)
 *   b  : row-major n*nrhs mpz_t (not modified);  xq : row-major n*nrhs initialised mpq_t
 * ------------------------------------------------------------------------------------- */
int ro_solve (int n, int nrhs, const ro_csc *L, const ro_csc *U, mpz_t *rhos, const int *pinv,
              mpz_t *b, mpq_t *xq)
{
    if (n <= 0 || nrhs <= 0 || !L || !U || !rhos || !pinv || !b || !xq) return RO_INCORRECT_INPUT;
    mpz_t *y = (mpz_t *) malloc ((size_t) n * sizeof (mpz_t));
    int *hist = (int *) malloc ((size_t) n * sizeof (int));
    if (!y || !hist) return RO_OUT_OF_MEMORY;
    for (int r = 0; r < n; r++) mpz_init (y[r]);
    mpq_t det; mpq_init (det); mpq_set_z (det, rhos[n - 1]);

    for (int c = 0; c < nrhs; c++)
    {
        for (int r = 0; r < n; r++) { mpz_set (y[pinv[r]], b[(size_t) r * nrhs + c]); hist[r] = -1; }
        /* forward: REF elimination of the dense right-hand side */
        for (int j = 0; j < n; j++)
        {
            if (mpz_sgn (y[j]) == 0) continue;
            lift (y[j], rhos, hist[j], j - 1);
            for (int m = L->p[j]; m < L->p[j + 1]; m++)
            {
                int i = L->i[m];
                if (i <= j || mpz_sgn (L->x[m]) == 0) continue;
                if (mpz_sgn (y[i]) != 0)
                {
                    lift (y[i], rhos, hist[i], j - 1);
                    mpz_mul (y[i], y[i], rhos[j]);
                }
                mpz_submul (y[i], L->x[m], y[j]);
                if (j >= 1) mpz_divexact (y[i], y[i], rhos[j - 1]);
                hist[i] = j;
            }
        }
        /* scale by det, then back substitution with U (diagonal is the last entry of a column) */
        for (int r = 0; r < n; r++) mpz_mul (y[r], y[r], rhos[n - 1]);
        for (int j = n - 1; j >= 0; j--)
        {
            if (mpz_sgn (y[j]) == 0) continue;
            int last = U->p[j + 1] - 1;
            mpz_divexact (y[j], y[j], U->x[last]);
            for (int m = U->p[j]; m < last; m++)
                if (mpz_sgn (U->x[m]) != 0) mpz_submul (y[U->i[m]], U->x[m], y[j]);
        }
        for (int r = 0; r < n; r++)
        {
            mpq_ptr o = xq[(size_t) r * nrhs + c];
            mpq_set_z (o, y[r]);
            mpq_div (o, o, det);
        }
    }
    mpq_clear (det);
    for (int r = 0; r < n; r++) mpz_clear (y[r]);
    free (y); free (hist);
    return RO_OK;
}

/* ---------------------------------------------------------------------------------------
 * digests: 64-bit FNV-1a over structure and limbs, for comparing large factors without
 * materialising them as Python integers.  Works on any (p, i, mpz_t*) triple, i.e. on the
 * SLIP_sparse of either library and on ro_csc.
 * ------------------------------------------------------------------------------------- */
static uint64_t fnv (uint64_t h, const void *data, size_t len)
{
    const unsigned char *d = (const unsigned char *) data;
    for (size_t k = 0; k < len; k++) { h ^= d[k]; h *= 1099511628211ULL; }
    return h;
}

uint64_t ro_digest_mpz (mpz_t *x, int count)
{
    uint64_t h = 1469598103934665603ULL;
    for (int k = 0; k < count; k++)
    {
        int sz = x[k]->_mp_size;
        h = fnv (h, &sz, sizeof (sz));
        h = fnv (h, x[k]->_mp_d, (size_t) (sz < 0 ? -sz : sz) * sizeof (mp_limb_t));
    }
    return h;
}

uint64_t ro_digest_csc (int n, const int *p, const int *i, mpz_t *x)
{
    uint64_t h = 1469598103934665603ULL;
    h = fnv (h, p, ((size_t) n + 1) * sizeof (int));
    h = fnv (h, i, (size_t) p[n] * sizeof (int));
    uint64_t hx = ro_digest_mpz (x, p[n]);
    return fnv (h, &hx, sizeof (hx));
}

uint64_t ro_digest_mpq (mpq_t *x, int count)
{
    uint64_t h = 1469598103934665603ULL;
    for (int k = 0; k < count; k++)
    {
        uint64_t a = ro_digest_mpz ((mpz_t *) mpq_numref (x[k]), 1);
        uint64_t b = ro_digest_mpz ((mpz_t *) mpq_denref (x[k]), 1);
        h = fnv (h, &a, sizeof (a)); h = fnv (h, &b, sizeof (b));
    }
    return h;
}
