/* slip_solve.c -- SLIP_LU_solve and the SLIP_solve_* drivers.
 *
 * Mirrors SLIP_LU/Source/SLIP_LU_solve.c:44-94 (permute b, forward substitution, scale by det,
 * back substitution, divide by det), SLIP_solve_mpq.c / SLIP_solve_double.c, SLIP_permute_x.c,
 * SLIP_scale_x.c, SLIP_check_solution.c, SLIP_get_double_soln.c.  The substitutions run on the
 * GPU against the resident factors; the host turns the integer numerators det*x_i into canonical
 * rationals (the mpq_div of slip_array_div.c) and applies the permutation and the scale. */
#include <time.h>
#include "slip_internal.h"

static double now_s (void)
{
    struct timespec t ;
    clock_gettime (CLOCK_MONOTONIC, &t) ;
    return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec ;
}

/* ---- numerators det*x_i streamed from the device: x[i][c] = num / det, canonical ---- */
typedef struct
{
    mpq_t **x ;
    mpz_srcptr det ;
    /* verification of a bound-mode result (NULL: the channel count is proven a priori) */
    const SLIP_sparse *A ; const int32_t *q ; const SLIP_dense *b ;
    int failed ;
    /* |det| = Ds * Dr: Ds collects the prime factors below SMALL_PRIME_BOUND, Dr is the rest */
    int split_done ;
    mpz_t D, Ds, Dr ;
} rational_ctx ;

/* ---- canonical form of N_t / det for a whole column with ONE big gcd ----
 * mpq_canonicalize costs a gcd of two numbers of the size of det per entry (the mpq_div of
 * slip_array_div.c does the same in the reference).  All entries share the denominator, which allows
 * this instead: split |det| = Ds * Dr (Ds: every prime factor < 2^16, found by trial division once
 * per solve).  Any prime common to some N_t and Dr divides P = prod_t N_t mod Dr, so if
 * gcd (P, Dr) = 1 -- one gcd per column, a modular product per entry -- no N_t shares a factor with
 * Dr and gcd (N_t, det) = gcd (N_t mod Ds, Ds), a small-operand gcd.  Otherwise (rare) the entries
 * that share a factor with gcd (P, Dr) get a full gcd.  The result is the canonical form GMP
 * produces (positive denominator, no common factor): bit-identical to the reference's x. */
#define SMALL_PRIME_BOUND 65536

static void det_split (rational_ctx *ctx)
{
    mpz_init (ctx->D) ; mpz_init_set_ui (ctx->Ds, 1) ; mpz_init (ctx->Dr) ;
    mpz_abs (ctx->D, ctx->det) ;
    mpz_set (ctx->Dr, ctx->D) ;
    static uint8_t *sieve = NULL ;
    #pragma omp critical (slip_b200_sieve)
    if (!sieve)
    {
        uint8_t *sv = (uint8_t *) calloc (SMALL_PRIME_BOUND, 1) ;
        if (sv)
        {
            for (uint32_t i = 2 ; i * i < SMALL_PRIME_BOUND ; i++)
                if (!sv [i]) for (uint32_t j = i * i ; j < SMALL_PRIME_BOUND ; j += i) sv [j] = 1 ;
            sieve = sv ;
        }
    }
    if (sieve && mpz_cmp_ui (ctx->Dr, 1) > 0)
        for (uint32_t p = 2 ; p < SMALL_PRIME_BOUND ; p++)
        {
            if (sieve [p]) continue ;
            while (mpz_divisible_ui_p (ctx->Dr, p))
            {
                mpz_divexact_ui (ctx->Dr, ctx->Dr, p) ;
                mpz_mul_ui (ctx->Ds, ctx->Ds, p) ;
            }
        }
    ctx->split_done = 1 ;
}

static void canonical_column (rational_ctx *ctx, int c, int cnt) ;
static void rational_ctx_clear (rational_ctx *ctx) ;

/* test hook (not part of the public interface; tests/test_canonical.py): x[t][0] holds numerators
 * N_t with any denominator; on return x[t][0] = N_t / det in canonical form */
int slip_b200_canonical_column_selftest (mpq_t **x, int32_t n, const mpz_t det)
{
    if (!x || n <= 0 || mpz_sgn (det) == 0) return -1 ;
    rational_ctx ctx ;
    memset (&ctx, 0, sizeof (ctx)) ;
    ctx.x = x ; ctx.det = det ;
    canonical_column (&ctx, 0, n) ;
    rational_ctx_clear (&ctx) ;
    return 0 ;
}

static void rational_ctx_clear (rational_ctx *ctx)
{
    if (ctx->split_done) { mpz_clear (ctx->D) ; mpz_clear (ctx->Ds) ; mpz_clear (ctx->Dr) ; ctx->split_done = 0 ; }
}

/* x[t][c] <- N_t / det in canonical form, N_t already in the numerators */
static void canonical_column (rational_ctx *ctx, int c, int cnt)
{
    if (!ctx->split_done) det_split (ctx) ;
    const int det_neg = mpz_sgn (ctx->det) < 0 ;
    const int rough = mpz_cmp_ui (ctx->Dr, 1) > 0 ;
    mpz_t G ;                    /* gcd (prod N_t mod Dr, Dr) */
    mpz_init_set_ui (G, 1) ;
    if (rough)
    {
        mpz_t P ;
        mpz_init_set_ui (P, 1) ;
        #pragma omp parallel if (cnt > 16)
        {
            mpz_t part, tmp ;
            mpz_init_set_ui (part, 1) ; mpz_init (tmp) ;
            #pragma omp for schedule(static) nowait
            for (int t = 0 ; t < cnt ; t++)
            {
                mpz_srcptr N = mpq_numref (ctx->x [t][c]) ;
                if (mpz_sgn (N) == 0) continue ;
                mpz_mul (tmp, part, N) ;
                mpz_tdiv_r (part, tmp, ctx->Dr) ;
            }
            #pragma omp critical (slip_b200_prod)
            { mpz_mul (tmp, P, part) ; mpz_tdiv_r (P, tmp, ctx->Dr) ; }
            mpz_clear (part) ; mpz_clear (tmp) ;
        }
        mpz_gcd (G, P, ctx->Dr) ;
        mpz_clear (P) ;
    }
    const int g_trivial = mpz_cmp_ui (G, 1) == 0 ;
    const int ds_small = mpz_fits_ulong_p (ctx->Ds) ;
    const unsigned long ds_ul = ds_small ? mpz_get_ui (ctx->Ds) : 0 ;
    #pragma omp parallel if (cnt > 16)
    {
        mpz_t g, t2 ;
        mpz_init (g) ; mpz_init (t2) ;
        #pragma omp for schedule(dynamic, 16)
        for (int t = 0 ; t < cnt ; t++)
        {
            mpq_ptr q = ctx->x [t][c] ;
            mpz_ptr N = mpq_numref (q), Dn = mpq_denref (q) ;
            if (mpz_sgn (N) == 0) { mpz_set_ui (Dn, 1) ; continue ; }
            /* smooth part */
            if (ds_small)
            {
                if (ds_ul > 1) mpz_set_ui (g, mpz_gcd_ui (NULL, N, ds_ul)) ; else mpz_set_ui (g, 1) ;
            }
            else mpz_gcd (g, N, ctx->Ds) ;
            /* rough part: only entries that share a factor with G can share one with Dr */
            if (!g_trivial)
            {
                mpz_gcd (t2, N, G) ;
                if (mpz_cmp_ui (t2, 1) != 0) { mpz_gcd (t2, N, ctx->Dr) ; mpz_mul (g, g, t2) ; }
            }
            if (mpz_cmp_ui (g, 1) == 0) mpz_set (Dn, ctx->D) ;
            else { mpz_divexact (N, N, g) ; mpz_divexact (Dn, ctx->D, g) ; }
            if (det_neg) mpz_neg (N, N) ;
        }
        mpz_clear (g) ; mpz_clear (t2) ;
    }
    mpz_clear (G) ;
}


/* Exact check of one right-hand side: sum_i A(:,q[i]) * N_i == det * b(:,c) over the integers.
 * x = N/det is then THE solution of A x = b (A is nonsingular: every pivot was nonzero), whatever
 * the channel count it was reconstructed from -- a wrong reconstruction cannot pass. */
static int numerators_verify (const rational_ctx *ctx, int c, int n)
{
    const SLIP_sparse *A = ctx->A ;
    int ok = 1 ;
    mpz_t *acc = SLIP_create_mpz_array (n) ;
    mpz_t t ;
    if (!acc) return 0 ;
    mpz_init (t) ;
    for (int32_t i = 0 ; i < n ; i++)
    {
        mpz_srcptr Ni = mpq_numref (ctx->x [i][c]) ;
        if (mpz_sgn (Ni) == 0) continue ;
        const int32_t col = ctx->q [i] ;
        for (int32_t a = A->p [col] ; a < A->p [col + 1] ; a++) mpz_addmul (acc [A->i [a]], A->x [a], Ni) ;
    }
    for (int32_t r = 0 ; r < n && ok ; r++)
    {
        mpz_mul (t, ctx->det, ctx->b->x [r][c]) ;
        if (mpz_cmp (t, acc [r]) != 0) ok = 0 ;
    }
    mpz_clear (t) ;
    SLIP_delete_mpz_array (&acc, n) ;
    return ok ;
}

static __thread double t_sink_import = 0, t_sink_verify = 0, t_sink_canon = 0 ;

static int rational_column (void *user, int c, int cnt, int stride, const uint32_t *limbs,
    const int32_t *nl, const int8_t *sign)
{
    rational_ctx *ctx = (rational_ctx *) user ;
    double tt = now_s () ;
    #pragma omp parallel for schedule(dynamic, 8) if (cnt > 16)
    for (int t = 0 ; t < cnt ; t++)
        slip_mpz_from_words (mpq_numref (ctx->x [t][c]), limbs + (size_t) t * stride, nl [t], sign [t]) ;
    t_sink_import += now_s () - tt ; tt = now_s () ;
    if (ctx->A && !ctx->failed && !numerators_verify (ctx, c, cnt)) ctx->failed = 1 ;
    t_sink_verify += now_s () - tt ; tt = now_s () ;
    if (ctx->failed) return 0 ;
    if (getenv ("SLIP_B200_CANON_GMP"))
    {   /* the per-entry route (tests compare the two) */
        #pragma omp parallel for schedule(dynamic, 8) if (cnt > 16)
        for (int t = 0 ; t < cnt ; t++)
        {
            mpq_ptr q = ctx->x [t][c] ;
            mpz_set (mpq_denref (q), ctx->det) ;
            mpq_canonicalize (q) ;
        }
    }
    else canonical_column (ctx, c, cnt) ;
    t_sink_canon += now_s () - tt ;
    return 0 ;
}

/* mode 1: the channels behind r are proven to hold the result (Cramer/Hadamard bound of det*x).
 * mode 0: they are an estimate (factors uploaded without A): reconstruct over every channel and
 *         report SLIP_INCORRECT if the two top channels were needed.
 * Bound-mode sessions (fewer channels than the a-priori bound) whose channels do not cover the
 * Cramer bound reconstruct over every channel and VERIFY the numerators exactly against A. */
static SLIP_info solve_on_device (mpq_t **x, SLIP_dense *b, slip_resident *r, const int32_t *pinv, int verified,
    const SLIP_sparse *A, const int32_t *q)
{
    SLIP_info status = SLIP_OK ;
    const int32_t n = r->n, nrhs = b->n ;
    slip_limbs bl = {0} ;
    if (b->m != n || nrhs <= 0) return SLIP_INCORRECT_INPUT ;
    if (mpz_sgn (r->det) == 0) return SLIP_SINGULAR ;
    {
        int64_t words = 0 ;
        for (int32_t i = 0 ; i < n ; i++)
            for (int32_t c = 0 ; c < nrhs ; c++) words += slip_mpz_words (b->x [i][c]) ;
        SLIP_TRY (slip_limbs_begin (&bl, (int64_t) n * nrhs, words)) ;
        /* right-hand side c occupies entries c*n .. c*n+n-1 (a batch of them is one contiguous range) */
        for (int32_t c = 0 ; c < nrhs ; c++)
            for (int32_t i = 0 ; i < n ; i++) slip_limbs_put (&bl, (int64_t) c * n + i, b->x [i][c]) ;
    }
    {
        const int have = slipcu_factor_channels (r->dev) ;
        int s = have, check = 0 ;
        if (verified)
        {   /* Cramer: det*x_i is the determinant of A with one column replaced by b */
            s = slip_channels_for_bits (r->total_bits - r->min_col_bits + slip_dense_max_column_bits (b)) ;
            if (s > have)
            {
                if (r->proven_channels || !A || !q) { status = SLIP_B200_NEED_CHANNELS ; goto cleanup ; }
                s = have ; check = 1 ;
            }
        }
        int32_t top = -1 ;
        rational_ctx ctx ;
        memset (&ctx, 0, sizeof (ctx)) ;
        ctx.x = x ; ctx.det = r->det ; ctx.A = check ? A : NULL ; ctx.q = q ; ctx.b = b ;
        const int rc_dev = slipcu_solve (r->dev, nrhs, bl.limbs, bl.off, bl.sign, pinv, s, rational_column, &ctx, &top) ;
        rational_ctx_clear (&ctx) ;
        SLIP_TRY (slip_from_device_status (rc_dev)) ;
        if (!verified && top >= have - 2) status = SLIP_INCORRECT ;
        if (check) { slip_last_stats.verified_solves += nrhs ; if (ctx.failed) status = SLIP_B200_NEED_CHANNELS ; }
    }
cleanup:
    slip_limbs_free (&bl) ;
    return status ;
}

SLIP_info slip_solve_resident (mpq_t **x, SLIP_dense *b, slip_resident *r, const int32_t *pinv,
    const SLIP_sparse *A, const int32_t *q)
{
    return solve_on_device (x, b, r, pinv, 1, A, q) ;
}

SLIP_sparse *slip_sparse_copy (const SLIP_sparse *A)
{
    SLIP_sparse *C = SLIP_create_sparse () ;
    if (!C) return NULL ;
    const int32_t n = A->n, nz = A->p [n] ;
    C->m = A->m ; C->n = n ; C->nz = nz ; C->nzmax = nz ;
    C->p = (int32_t *) SLIP_malloc ((size_t) (n + 1) * sizeof (int32_t)) ;
    C->i = (int32_t *) SLIP_malloc ((size_t) (nz > 0 ? nz : 1) * sizeof (int32_t)) ;
    C->x = SLIP_create_mpz_array (nz) ;
    if (!C->p || !C->i || !C->x) { SLIP_delete_sparse (&C) ; return NULL ; }
    memcpy (C->p, A->p, (size_t) (n + 1) * sizeof (int32_t)) ;
    memcpy (C->i, A->i, (size_t) nz * sizeof (int32_t)) ;
    for (int32_t a = 0 ; a < nz ; a++) mpz_set (C->x [a], A->x [a]) ;
    return C ;
}

/* resident session from host-side factors (L, U, rhos given as mpz_t) */
static SLIP_info upload_factors (slip_resident **out, const SLIP_sparse *L, const SLIP_sparse *U,
    const mpz_t *rhos, int channels)
{
    SLIP_info status = SLIP_OK ;
    const int32_t n = L->n ;
    slip_limbs vl = {0} ;
    slip_resident *res = NULL ;
    int32_t *cnt = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    int32_t *nU = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    int32_t *piv = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    const int64_t total = (int64_t) L->p [n] + U->p [n] - n ;
    int32_t *rows = (int32_t *) SLIP_malloc ((size_t) (total > 0 ? total : 1) * sizeof (int32_t)) ;
    if (!cnt || !nU || !piv || !rows) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
    {
        int64_t words = 0 ;
        for (int32_t m = 0 ; m < L->p [n] ; m++) words += slip_mpz_words (L->x [m]) ;
        for (int32_t m = 0 ; m < U->p [n] ; m++) words += slip_mpz_words (U->x [m]) ;
        SLIP_TRY (slip_limbs_begin (&vl, total, words)) ;
    }
    int64_t e = 0 ;
    for (int32_t k = 0 ; k < n ; k++)
    {
        const int32_t u0 = U->p [k], u1 = U->p [k + 1] - 1 ;     /* the diagonal closes the column */
        const int32_t l0 = L->p [k], l1 = L->p [k + 1] ;
        if (u1 < u0 || U->i [u1] != k) { status = SLIP_INCORRECT_INPUT ; goto cleanup ; }
        nU [k] = u1 - u0 ; cnt [k] = (u1 - u0) + (l1 - l0) ; piv [k] = -1 ;
        for (int32_t m = u0 ; m < u1 ; m++) { rows [e] = U->i [m] ; slip_limbs_put (&vl, e, U->x [m]) ; e++ ; }
        for (int32_t m = l0 ; m < l1 ; m++)
        {
            if (L->i [m] == k) piv [k] = nU [k] + (m - l0) ;
            rows [e] = L->i [m] ; slip_limbs_put (&vl, e, L->x [m]) ; e++ ;
        }
        if (piv [k] < 0) { status = SLIP_INCORRECT_INPUT ; goto cleanup ; }
    }
    res = (slip_resident *) SLIP_calloc (1, sizeof (slip_resident)) ;
    if (!res) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
    mpz_init_set (res->det, rhos [n - 1]) ;
    res->n = n ;
    res->total_bits = -1 ;               /* no A here: the solve reconstructs over every channel */
    SLIP_TRY (slip_from_device_status (slipcu_factor_upload (&res->dev, n, channels, cnt, nU, piv, rows,
        vl.limbs, vl.off, vl.sign))) ;
    *out = res ; res = NULL ;
cleanup:
    if (res) slip_resident_free (res) ;
    slip_limbs_free (&vl) ;
    SLIP_free (cnt) ; SLIP_free (nU) ; SLIP_free (piv) ; SLIP_free (rows) ;
    return status ;
}

/* largest entry of a SLIP_sparse / mpz array in bits */
static double max_bits_sparse (const SLIP_sparse *M)
{
    size_t best = 0 ;
    for (int32_t m = 0 ; m < M->p [M->n] ; m++)
    {
        size_t b = mpz_sgn (M->x [m]) ? mpz_sizeinbase (M->x [m], 2) : 0 ;
        if (b > best) best = b ;
    }
    return (double) best ;
}

SLIP_info SLIP_LU_solve (mpq_t **x, SLIP_dense *b, const mpz_t *rhos, const SLIP_sparse *L,
    const SLIP_sparse *U, const int32_t *pinv)
{
    if (!x || !b || !rhos || !pinv || !L || !U || !b->x || !L->p || !L->i || !L->x || !U->p || !U->i || !U->x)
        return SLIP_INCORRECT_INPUT ;
    SLIP_info status = SLIP_OK ;
    const int32_t n = L->n ;
    for (int32_t r0 = 0 ; r0 < n ; r0++) if (pinv [r0] < 0 || pinv [r0] >= n) return SLIP_INCORRECT_INPUT ;
    slip_resident *r = slip_resident_acquire (L->x, U->x, n, rhos [n - 1]), *tmp = NULL ;
    const double bbits = slip_dense_max_column_bits (b) ;
    int need = -1 ;
    double tot_bits = 0, min_bits = 0 ;
    if (r)
    {   /* do the resident factors carry enough channels for this right-hand side?  (bound-mode
           factors: the result is verified against the copy of A kept with them) */
        need = slip_channels_for_bits (r->total_bits - r->min_col_bits + bbits) ;
        tot_bits = r->total_bits ; min_bits = r->min_col_bits ;
        status = solve_on_device (x, b, r, pinv, 1, r->A_copy, r->q_copy) ;
        slip_resident_release (r) ; r = NULL ;
        if (status != SLIP_B200_NEED_CHANNELS) return status ;
        status = SLIP_OK ;
    }
    /* Factors that are not resident (or too narrow) are re-encoded from the host copies.  With
     * the bound of the resident factorization at hand the channel count is exact.  Without it
     * (L, U from elsewhere) there is no A to bound det*x with, so the count is estimated from
     * the factors, the result is reconstructed over all channels, and the two top channels must
     * stay empty; otherwise the channels are doubled and the solve repeated. */
    int channels = need > 0 ? need + SLIP_B200_SPARE_CHANNELS
        : slip_channels_for_bits (max_bits_sparse (L) + max_bits_sparse (U) + bbits + 2.0 * log2 ((double) n + 1.0) + 64.0) + 2 ;
    for (int attempt = 0 ; attempt < 8 ; attempt++)
    {
        SLIP_TRY (upload_factors (&tmp, L, U, rhos, channels)) ;
        if (need > 0) { tmp->total_bits = tot_bits ; tmp->min_col_bits = min_bits ; tmp->proven_channels = 1 ; }
        status = solve_on_device (x, b, tmp, pinv, need > 0, NULL, NULL) ;
        slip_resident_free (tmp) ; tmp = NULL ;
        if (status != SLIP_INCORRECT || need > 0) break ;
        channels *= 2 ;
    }
    if (status == SLIP_B200_NEED_CHANNELS) status = SLIP_INCORRECT ;
cleanup:
    if (tmp) slip_resident_free (tmp) ;
    return status ;
}

SLIP_info SLIP_permute_x (mpq_t **x, int32_t n, int32_t numRHS, SLIP_LU_analysis *S)
{
    if (!x || !S || !S->q) return SLIP_INCORRECT_INPUT ;
    /* x <- Q x: row i of the factor ordering is row q[i] of the original system.  Only the row
       pointers move. */
    mpq_t **tmp = (mpq_t **) SLIP_malloc ((size_t) n * sizeof (mpq_t *)) ;
    if (!tmp) return SLIP_OUT_OF_MEMORY ;
    (void) numRHS ;
    for (int32_t i = 0 ; i < n ; i++) tmp [S->q [i]] = x [i] ;
    memcpy (x, tmp, (size_t) n * sizeof (mpq_t *)) ;
    SLIP_free (tmp) ;
    return SLIP_OK ;
}

SLIP_info SLIP_scale_x (mpq_t **x, SLIP_sparse *A, SLIP_dense *b)
{
    if (!x || !A || !b) return SLIP_INCORRECT_INPUT ;
    const int32_t n = A->m, nrhs = b->n ;
    if (mpq_cmp_ui (A->scale, 1, 1) != 0 && mpq_cmp_ui (A->scale, 0, 1) != 0)
        for (int32_t i = 0 ; i < n ; i++)
            for (int32_t j = 0 ; j < nrhs ; j++) mpq_mul (x [i][j], x [i][j], A->scale) ;
    if (mpq_cmp_ui (b->scale, 1, 1) != 0 && mpq_cmp_ui (b->scale, 0, 1) != 0)
        for (int32_t i = 0 ; i < n ; i++)
            for (int32_t j = 0 ; j < nrhs ; j++) mpq_div (x [i][j], x [i][j], b->scale) ;
    return SLIP_OK ;
}

static __thread int32_t *last_pinv = NULL ;
static __thread int32_t last_pinv_n = 0 ;

/* row permutation of the last SLIP_solve_* call of this thread (the factors of that path never
 * reach the host, so this is how tests compare its pivoting with the reference's pinv) */
int SLIP_B200_last_pinv (int32_t *out, int cap)
{
    int k = 0 ;
    for ( ; k < last_pinv_n && k < cap ; k++) out [k] = last_pinv [k] ;
    return k ;
}

/* factor on the GPU, solve on the GPU; L and U never come to the host */
static SLIP_info solve_exact (mpq_t **x, SLIP_sparse *A, SLIP_LU_analysis *S, SLIP_dense *b, SLIP_options *option)
{
    SLIP_info status = SLIP_OK ;
    slip_resident *r = NULL ;
    const int32_t n = A->n ;
    int32_t *pinv = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    if (!pinv) return SLIP_OUT_OF_MEMORY ;
    int min_channels = 0 ;
    for (int attempt = 0 ; ; attempt++)
    {
        double t0 = now_s () ;
        SLIP_TRY (slip_factorize_driver (NULL, NULL, A, S, NULL, pinv, option, 0, &r, slip_dense_max_column_bits (b), min_channels)) ;
        double t1 = now_s () ;
        status = slip_solve_resident (x, b, r, pinv, A, S->q) ;
        if (getenv ("SLIP_B200_TIMING"))
        {
            fprintf (stderr, "slip_lu_b200 timing: factor %.3fs solve %.3fs (host sink: limbs->mpz %.3fs, exact verification %.3fs, canonical rationals %.3fs)\n",
                t1 - t0, now_s () - t1, t_sink_import, t_sink_verify, t_sink_canon) ;
            t_sink_import = t_sink_verify = t_sink_canon = 0 ;
        }
        if (status != SLIP_B200_NEED_CHANNELS) break ;
        /* the factors fitted the channels but det*x does not (a large right-hand side): again
           with four times the channels */
        min_channels = 4 * slipcu_factor_channels (r->dev) ;
        slip_resident_free (r) ; r = NULL ;
        if (attempt >= 6) { status = SLIP_INCORRECT ; goto cleanup ; }
    }
    if (status != SLIP_OK) goto cleanup ;
    {
        int32_t *keep = (int32_t *) realloc (last_pinv, (size_t) n * sizeof (int32_t)) ;
        if (keep) { last_pinv = keep ; last_pinv_n = n ; memcpy (last_pinv, pinv, (size_t) n * sizeof (int32_t)) ; }
    }
    SLIP_TRY (SLIP_permute_x (x, n, b->n, S)) ;
    SLIP_TRY (SLIP_scale_x (x, A, b)) ;
cleanup:
    if (r) slip_resident_free (r) ;
    SLIP_free (pinv) ;
    return status ;
}

SLIP_info SLIP_solve_mpq (mpq_t **x_mpq, SLIP_sparse *A, SLIP_LU_analysis *S, SLIP_dense *b, SLIP_options *option)
{
    if (!x_mpq || !A || !A->p || !A->i || !A->x || !S || !S->q || !b || !b->x || !option)
        return SLIP_INCORRECT_INPUT ;
    return solve_exact (x_mpq, A, S, b, option) ;
}

SLIP_info SLIP_get_double_soln (double **x_doub, mpq_t **x_mpq, int32_t n, int32_t numRHS)
{
    if (!x_doub || !x_mpq) return SLIP_INCORRECT_INPUT ;
    for (int32_t i = 0 ; i < n ; i++)
        for (int32_t j = 0 ; j < numRHS ; j++) x_doub [i][j] = mpq_get_d (x_mpq [i][j]) ;
    return SLIP_OK ;
}

SLIP_info SLIP_solve_double (double **x_doub, SLIP_sparse *A, SLIP_LU_analysis *S, SLIP_dense *b, SLIP_options *option)
{
    if (!x_doub || !A || !A->p || !A->i || !A->x || !S || !S->q || !b || !b->x || !option)
        return SLIP_INCORRECT_INPUT ;
    mpq_t **x = SLIP_create_mpq_mat (A->n, b->n) ;
    if (!x) return SLIP_OUT_OF_MEMORY ;
    SLIP_info status = solve_exact (x, A, S, b, option) ;
    if (status == SLIP_OK) status = SLIP_get_double_soln (x_doub, x, A->n, b->n) ;
    SLIP_delete_mpq_mat (&x, A->n, b->n) ;
    return status ;
}

/* SLIP_get_mpfr_soln.c:27-49: x_mpfr[i][j] = x_mpq[i][j] rounded per option->SLIP_MPFR_ROUND (the
 * precision is the one x_mpfr was created with) */
SLIP_info SLIP_get_mpfr_soln (mpfr_t **x_mpfr, mpq_t **x_mpq, int32_t n, int32_t numRHS, SLIP_options *option)
{
    if (!x_mpfr || !x_mpq || !option) return SLIP_INCORRECT_INPUT ;
    for (int32_t i = 0 ; i < n ; i++)
        for (int32_t j = 0 ; j < numRHS ; j++)
            mpfr_set_q (x_mpfr [i][j], x_mpq [i][j], option->SLIP_MPFR_ROUND) ;
    return SLIP_OK ;
}

/* SLIP_solve_mpfr.c:35-77: factor, solve, permute, scale, then round to mpfr */
SLIP_info SLIP_solve_mpfr (mpfr_t **x_mpfr, SLIP_sparse *A, SLIP_LU_analysis *S, SLIP_dense *b, SLIP_options *option)
{
    if (!x_mpfr || !A || !A->p || !A->i || !A->x || !S || !S->q || !b || !b->x || !option)
        return SLIP_INCORRECT_INPUT ;
    mpq_t **x = SLIP_create_mpq_mat (A->n, b->n) ;
    if (!x) return SLIP_OUT_OF_MEMORY ;
    SLIP_info status = solve_exact (x, A, S, b, option) ;
    if (status == SLIP_OK) status = SLIP_get_mpfr_soln (x_mpfr, x, A->n, b->n, option) ;
    SLIP_delete_mpq_mat (&x, A->n, b->n) ;
    return status ;
}

/* exact residual check A x == b in rational arithmetic, for the integer system (before scaling) */
SLIP_info SLIP_check_solution (SLIP_sparse *A, mpq_t **x, SLIP_dense *b)
{
    if (!A || !x || !b || !b->x || !A->p || !A->i || !A->x) return SLIP_INCORRECT_INPUT ;
    const int32_t n = A->n, nrhs = b->n ;
    SLIP_info status = SLIP_OK ;
    mpq_t **acc = SLIP_create_mpq_mat (n, nrhs) ;
    if (!acc) return SLIP_OUT_OF_MEMORY ;
    mpq_t t ;
    mpq_init (t) ;
    for (int32_t c = 0 ; c < nrhs ; c++)
        for (int32_t j = 0 ; j < n ; j++)
            for (int32_t a = A->p [j] ; a < A->p [j + 1] ; a++)
            {
                mpq_set_z (t, A->x [a]) ;
                mpq_mul (t, t, x [j][c]) ;
                mpq_add (acc [A->i [a]][c], acc [A->i [a]][c], t) ;
            }
    for (int32_t c = 0 ; c < nrhs && status == SLIP_OK ; c++)
        for (int32_t i = 0 ; i < n ; i++)
        {
            mpq_set_z (t, b->x [i][c]) ;
            if (!mpq_equal (t, acc [i][c])) { status = SLIP_INCORRECT ; break ; }
        }
    mpq_clear (t) ;
    SLIP_delete_mpq_mat (&acc, n, nrhs) ;
    return status ;
}

/* structural check of a CSC matrix (the reference also prints it; SLIP_spok.c) */
SLIP_info SLIP_spok (SLIP_sparse *A, SLIP_options *option)
{
    if (!A || !option || !A->p || !A->i || !A->x) return SLIP_INCORRECT_INPUT ;
    const int32_t n = A->n, m = A->m ;
    if (n < 0 || m < 0 || A->nzmax < 0 || A->p [0] != 0 || A->p [n] < 0 || A->p [n] > A->nzmax)
        return SLIP_INCORRECT_INPUT ;
    int32_t *seen = (int32_t *) SLIP_malloc ((size_t) (m > 0 ? m : 1) * sizeof (int32_t)) ;
    if (!seen) return SLIP_OUT_OF_MEMORY ;
    for (int32_t i = 0 ; i < m ; i++) seen [i] = -1 ;
    SLIP_info status = SLIP_OK ;
    for (int32_t j = 0 ; j < n && status == SLIP_OK ; j++)
    {
        if (A->p [j] > A->p [j + 1]) { status = SLIP_INCORRECT_INPUT ; break ; }
        for (int32_t a = A->p [j] ; a < A->p [j + 1] ; a++)
        {
            const int32_t i = A->i [a] ;
            if (i < 0 || i >= m || seen [i] == j) { status = SLIP_INCORRECT_INPUT ; break ; }
            seen [i] = j ;
        }
    }
    SLIP_free (seen) ;
    return status ;
}
