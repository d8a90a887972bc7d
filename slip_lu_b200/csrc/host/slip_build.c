/* slip_build.c -- input builders: user data (mpz / int / double / mpq; CSC, triplet, dense)
 * into the integer SLIP_sparse / SLIP_dense containers plus their rational scale.
 * Mirrors SLIP_LU/Source/SLIP_build_{sparse_ccf,sparse_trip,dense}_{mpz,int,double,mpq}.c,
 * slip_mpz_populate_mat.c, slip_trip_to_mat.c, slip_expand_{double,mpq}_{array,mat}.c. */
#include "slip_internal.h"

/* ---- containers ---- */
static SLIP_info sparse_alloc (SLIP_sparse *A, int32_t n, int32_t nzmax)
{
    if (!A || n <= 0 || nzmax <= 0) return SLIP_INCORRECT_INPUT ;
    A->m = A->n = n ;
    A->nz = 0 ;
    A->nzmax = nzmax ;
    A->x = SLIP_create_mpz_array (nzmax) ;
    A->p = (int32_t *) SLIP_calloc ((size_t) n + 1, sizeof (int32_t)) ;
    A->i = (int32_t *) SLIP_calloc ((size_t) nzmax, sizeof (int32_t)) ;
    return (A->x && A->p && A->i) ? SLIP_OK : SLIP_OUT_OF_MEMORY ;
}

SLIP_info slip_sparse_from_ccf (SLIP_sparse *A, const int32_t *p, const int32_t *I, mpz_t *x,
    int32_t n, int32_t nz)
{
    SLIP_info status = sparse_alloc (A, n, nz) ;
    if (status != SLIP_OK) return status ;
    A->nz = nz ;
    memcpy (A->p, p, ((size_t) n + 1) * sizeof (int32_t)) ;
    for (int32_t k = 0 ; k < nz ; k++)
    {
        if (I [k] < 0) return SLIP_INCORRECT_INPUT ;
        A->i [k] = I [k] ;
        mpz_set (A->x [k], x [k]) ;
    }
    return SLIP_OK ;
}

SLIP_info slip_sparse_from_trip (SLIP_sparse *A, const int32_t *I, const int32_t *J, mpz_t *x,
    int32_t n, int32_t nz)
{
    /* bucket by column, keeping input order inside a column (slip_trip_to_mat.c) */
    SLIP_info status = sparse_alloc (A, n, nz) ;
    if (status != SLIP_OK) return status ;
    int32_t *next = (int32_t *) SLIP_calloc ((size_t) n + 1, sizeof (int32_t)) ;
    if (!next) return SLIP_OUT_OF_MEMORY ;
    for (int32_t k = 0 ; k < nz ; k++) A->p [J [k] + 1]++ ;
    for (int32_t j = 0 ; j < n ; j++) A->p [j + 1] += A->p [j] ;
    memcpy (next, A->p, (size_t) n * sizeof (int32_t)) ;
    for (int32_t k = 0 ; k < nz ; k++)
    {
        int32_t dst = next [J [k]]++ ;
        if (I [k] < 0) { SLIP_free (next) ; return SLIP_INCORRECT_INPUT ; }
        A->i [dst] = I [k] ;
        mpz_set (A->x [dst], x [k]) ;
    }
    A->nz = nz ;
    SLIP_free (next) ;
    return SLIP_OK ;
}

/* ---- double -> integer: x_out = round (10^17 * x) / gcd, scale = 10^17 / gcd.
 * Same MPFR calls at option->prec / option->SLIP_MPFR_ROUND as slip_expand_double_array.c, so the
 * integers (including the sign convention when a single nonzero makes the "gcd" negative) match. */
static SLIP_info expand_doubles (mpz_t **slot, const double *x, int64_t count, mpq_t scale,
    SLIP_options *option)
{
    if (!option) return SLIP_INCORRECT_INPUT ;
    const double ten17 = pow (10, 17) ;
    mpfr_t w ;
    mpz_t g ;
    mpfr_init2 (w, (mpfr_prec_t) option->prec) ;
    mpz_init (g) ;
    mpq_set_d (scale, ten17) ;
    for (int64_t k = 0 ; k < count ; k++)
    {
        mpfr_set_d (w, x [k], option->SLIP_MPFR_ROUND) ;
        mpfr_mul_d (w, w, ten17, option->SLIP_MPFR_ROUND) ;
        mpfr_get_z (*slot [k], w, option->SLIP_MPFR_ROUND) ;
    }
    int64_t first = -1 ;
    int reduced_to_one = 0 ;
    for (int64_t k = 0 ; k < count && !reduced_to_one ; k++)
    {
        if (first < 0)
        {
            if (mpz_sgn (*slot [k]) != 0) { first = k ; mpz_set (g, *slot [k]) ; }
        }
        else
        {
            mpz_gcd (g, g, *slot [k]) ;
            if (mpz_cmp_ui (g, 1) == 0) reduced_to_one = 1 ;
        }
    }
    if (first < 0) mpq_set_ui (scale, 1, 1) ;          /* all zero */
    else if (!reduced_to_one)
    {
        mpq_t t ;
        mpq_init (t) ;
        for (int64_t k = first ; k < count ; k++) mpz_divexact (*slot [k], *slot [k], g) ;
        mpq_set_z (t, g) ;
        mpq_div (scale, scale, t) ;
        mpq_clear (t) ;
    }
    mpz_clear (g) ;
    mpfr_clear (w) ;
    return SLIP_OK ;
}

SLIP_info slip_expand_double_array (mpz_t *x_out, double *x, mpq_t scale, int32_t n, SLIP_options *option)
{
    mpz_t **slot = (mpz_t **) SLIP_malloc ((size_t) n * sizeof (mpz_t *)) ;
    if (!slot) return SLIP_OUT_OF_MEMORY ;
    for (int32_t k = 0 ; k < n ; k++) slot [k] = &x_out [k] ;
    SLIP_info status = expand_doubles (slot, x, n, scale, option) ;
    SLIP_free (slot) ;
    return status ;
}

SLIP_info slip_expand_double_mat (mpz_t **x_out, double **x, mpq_t scale, int32_t m, int32_t n,
    SLIP_options *option)
{
    const int64_t count = (int64_t) m * n ;
    mpz_t **slot = (mpz_t **) SLIP_malloc ((size_t) count * sizeof (mpz_t *)) ;
    double *flat = (double *) SLIP_malloc ((size_t) count * sizeof (double)) ;
    if (!slot || !flat) { SLIP_free (slot) ; SLIP_free (flat) ; return SLIP_OUT_OF_MEMORY ; }
    for (int32_t i = 0 ; i < m ; i++)
        for (int32_t j = 0 ; j < n ; j++)
        {
            slot [(int64_t) i * n + j] = &x_out [i][j] ;
            flat [(int64_t) i * n + j] = x [i][j] ;
        }
    SLIP_info status = expand_doubles (slot, flat, count, scale, option) ;
    SLIP_free (slot) ; SLIP_free (flat) ;
    return status ;
}

/* ---- mpfr -> integer: x_out = round (10^prec * x) / gcd, scale = 10^prec / gcd, with 10^prec
 * itself rounded to option->prec bits as in slip_expand_mpfr_array.c / slip_expand_mpfr_mat.c
 * (same MPFR calls, same row-major gcd scan that stops at 1 and divides from the first nonzero) ---- */
static SLIP_info expand_mpfrs (mpz_t **slot, mpfr_t **src, int64_t count, mpq_t scale, SLIP_options *option)
{
    if (!option) return SLIP_INCORRECT_INPUT ;
    mpfr_t ten, w ;
    mpz_t g ;
    mpfr_init2 (ten, (mpfr_prec_t) option->prec) ;
    mpfr_init2 (w, (mpfr_prec_t) option->prec) ;
    mpz_init (g) ;
    mpfr_ui_pow_ui (ten, 10, (unsigned long) option->prec, option->SLIP_MPFR_ROUND) ;
    for (int64_t k = 0 ; k < count ; k++)
    {
        mpfr_mul (w, *src [k], ten, option->SLIP_MPFR_ROUND) ;
        mpfr_get_z (*slot [k], w, option->SLIP_MPFR_ROUND) ;
    }
    mpfr_get_z (g, ten, option->SLIP_MPFR_ROUND) ;
    mpq_set_z (scale, g) ;
    int64_t first = -1 ;
    int reduced_to_one = 0 ;
    for (int64_t k = 0 ; k < count && !reduced_to_one ; k++)
    {
        if (first < 0)
        {
            if (mpz_sgn (*slot [k]) != 0) { first = k ; mpz_set (g, *slot [k]) ; }
        }
        else
        {
            mpz_gcd (g, g, *slot [k]) ;
            if (mpz_cmp_ui (g, 1) == 0) reduced_to_one = 1 ;
        }
    }
    if (first < 0) mpq_set_ui (scale, 1, 1) ;          /* all zero */
    else if (!reduced_to_one)
    {
        mpq_t t ;
        mpq_init (t) ;
        for (int64_t k = first ; k < count ; k++) mpz_divexact (*slot [k], *slot [k], g) ;
        mpq_set_z (t, g) ;
        mpq_div (scale, scale, t) ;
        mpq_clear (t) ;
    }
    mpz_clear (g) ;
    mpfr_clear (w) ; mpfr_clear (ten) ;
    return SLIP_OK ;
}

SLIP_info slip_expand_mpfr_array (mpz_t *x_out, mpfr_t *x, mpq_t scale, int32_t n, SLIP_options *option)
{
    if (!x || !x_out || n <= 0) return SLIP_INCORRECT_INPUT ;
    mpz_t **slot = (mpz_t **) SLIP_malloc ((size_t) n * sizeof (mpz_t *)) ;
    mpfr_t **src = (mpfr_t **) SLIP_malloc ((size_t) n * sizeof (mpfr_t *)) ;
    if (!slot || !src) { SLIP_free (slot) ; SLIP_free (src) ; return SLIP_OUT_OF_MEMORY ; }
    for (int32_t k = 0 ; k < n ; k++) { slot [k] = &x_out [k] ; src [k] = &x [k] ; }
    SLIP_info status = expand_mpfrs (slot, src, n, scale, option) ;
    SLIP_free (slot) ; SLIP_free (src) ;
    return status ;
}

SLIP_info slip_expand_mpfr_mat (mpz_t **x_out, mpfr_t **x, mpq_t scale, int32_t m, int32_t n, SLIP_options *option)
{
    const int64_t count = (int64_t) m * n ;
    mpz_t **slot = (mpz_t **) SLIP_malloc ((size_t) count * sizeof (mpz_t *)) ;
    mpfr_t **src = (mpfr_t **) SLIP_malloc ((size_t) count * sizeof (mpfr_t *)) ;
    if (!slot || !src) { SLIP_free (slot) ; SLIP_free (src) ; return SLIP_OUT_OF_MEMORY ; }
    for (int32_t i = 0 ; i < m ; i++)
        for (int32_t j = 0 ; j < n ; j++)
        {
            slot [(int64_t) i * n + j] = &x_out [i][j] ;
            src [(int64_t) i * n + j] = &x [i][j] ;
        }
    SLIP_info status = expand_mpfrs (slot, src, count, scale, option) ;
    SLIP_free (slot) ; SLIP_free (src) ;
    return status ;
}

/* ---- rationals -> integer: scale = lcm of denominators, x_out = scale * x ---- */
static SLIP_info expand_rationals (mpz_t **slot, mpq_t **src, int64_t count, mpq_t scale)
{
    mpz_t l ;
    mpq_t t ;
    mpz_init (l) ; mpq_init (t) ;
    mpz_set (l, mpq_denref (*src [0])) ;
    for (int64_t k = 1 ; k < count ; k++) mpz_lcm (l, mpq_denref (*src [k]), l) ;
    mpq_set_z (scale, l) ;
    for (int64_t k = 0 ; k < count ; k++)
    {
        mpq_mul (t, *src [k], scale) ;
        mpz_set_q (*slot [k], t) ;
    }
    mpz_clear (l) ; mpq_clear (t) ;
    return SLIP_OK ;
}

SLIP_info slip_expand_mpq_array (mpz_t *x_out, mpq_t *x, mpq_t scale, int32_t n)
{
    mpz_t **slot = (mpz_t **) SLIP_malloc ((size_t) n * sizeof (mpz_t *)) ;
    mpq_t **src = (mpq_t **) SLIP_malloc ((size_t) n * sizeof (mpq_t *)) ;
    if (!slot || !src) { SLIP_free (slot) ; SLIP_free (src) ; return SLIP_OUT_OF_MEMORY ; }
    for (int32_t k = 0 ; k < n ; k++) { slot [k] = &x_out [k] ; src [k] = &x [k] ; }
    SLIP_info status = expand_rationals (slot, src, n, scale) ;
    SLIP_free (slot) ; SLIP_free (src) ;
    return status ;
}

SLIP_info slip_expand_mpq_mat (mpz_t **x_out, mpq_t **x, mpq_t scale, int32_t m, int32_t n)
{
    const int64_t count = (int64_t) m * n ;
    mpz_t **slot = (mpz_t **) SLIP_malloc ((size_t) count * sizeof (mpz_t *)) ;
    mpq_t **src = (mpq_t **) SLIP_malloc ((size_t) count * sizeof (mpq_t *)) ;
    if (!slot || !src) { SLIP_free (slot) ; SLIP_free (src) ; return SLIP_OUT_OF_MEMORY ; }
    for (int32_t i = 0 ; i < m ; i++)
        for (int32_t j = 0 ; j < n ; j++)
        {
            slot [(int64_t) i * n + j] = &x_out [i][j] ;
            src [(int64_t) i * n + j] = &x [i][j] ;
        }
    SLIP_info status = expand_rationals (slot, src, count, scale) ;
    SLIP_free (slot) ; SLIP_free (src) ;
    return status ;
}

/* ---- sparse builders ---- */
#define BAD_SPARSE_ARGS(idx1, idx2, x, A) (!(idx1) || !(idx2) || !(x) || !(A) || !mpq_denref ((A)->scale)->_mp_d)

SLIP_info SLIP_build_sparse_ccf_mpz (SLIP_sparse *A, int32_t *p, int32_t *I, mpz_t *x, int32_t n, int32_t nz)
{
    if (BAD_SPARSE_ARGS (p, I, x, A)) return SLIP_INCORRECT_INPUT ;
    SLIP_info status = slip_sparse_from_ccf (A, p, I, x, n, nz) ;
    if (status == SLIP_OK) mpq_set_ui (A->scale, 1, 1) ;
    return status ;
}

SLIP_info SLIP_build_sparse_trip_mpz (SLIP_sparse *A, int32_t *I, int32_t *J, mpz_t *x, int32_t n, int32_t nz)
{
    if (BAD_SPARSE_ARGS (I, J, x, A) || n <= 0 || nz <= 0) return SLIP_INCORRECT_INPUT ;
    SLIP_info status = slip_sparse_from_trip (A, I, J, x, n, nz) ;
    if (status == SLIP_OK) mpq_set_ui (A->scale, 1, 1) ;
    return status ;
}

typedef enum { SRC_INT, SRC_DOUBLE, SRC_MPQ, SRC_MPFR } src_kind ;

static SLIP_info build_sparse_any (SLIP_sparse *A, int32_t *a, int32_t *b, void *x, int32_t n, int32_t nz,
    src_kind kind, int triplet, SLIP_options *option)
{
    if (BAD_SPARSE_ARGS (a, b, x, A) || n <= 0 || nz <= 0) return SLIP_INCORRECT_INPUT ;
    mpz_t *xi = SLIP_create_mpz_array (nz) ;
    if (!xi) return SLIP_OUT_OF_MEMORY ;
    SLIP_info status = SLIP_OK ;
    if (kind == SRC_INT)
    {
        for (int32_t k = 0 ; k < nz ; k++) mpz_set_si (xi [k], ((int32_t *) x) [k]) ;
        mpq_set_ui (A->scale, 1, 1) ;
    }
    else if (kind == SRC_DOUBLE) status = slip_expand_double_array (xi, (double *) x, A->scale, nz, option) ;
    else if (kind == SRC_MPFR) status = slip_expand_mpfr_array (xi, (mpfr_t *) x, A->scale, nz, option) ;
    else status = slip_expand_mpq_array (xi, (mpq_t *) x, A->scale, nz) ;
    if (status == SLIP_OK)
        status = triplet ? slip_sparse_from_trip (A, a, b, xi, n, nz) : slip_sparse_from_ccf (A, a, b, xi, n, nz) ;
    SLIP_delete_mpz_array (&xi, nz) ;
    return status ;
}

SLIP_info SLIP_build_sparse_ccf_int (SLIP_sparse *A, int32_t *p, int32_t *I, int32_t *x, int32_t n, int32_t nz)
{ return build_sparse_any (A, p, I, x, n, nz, SRC_INT, 0, NULL) ; }
SLIP_info SLIP_build_sparse_ccf_double (SLIP_sparse *A, int32_t *p, int32_t *I, double *x, int32_t n, int32_t nz, SLIP_options *option)
{ return build_sparse_any (A, p, I, x, n, nz, SRC_DOUBLE, 0, option) ; }
SLIP_info SLIP_build_sparse_ccf_mpq (SLIP_sparse *A, int32_t *p, int32_t *I, mpq_t *x, int32_t n, int32_t nz)
{ return build_sparse_any (A, p, I, x, n, nz, SRC_MPQ, 0, NULL) ; }
SLIP_info SLIP_build_sparse_trip_int (SLIP_sparse *A, int32_t *I, int32_t *J, int32_t *x, int32_t n, int32_t nz)
{ return build_sparse_any (A, I, J, x, n, nz, SRC_INT, 1, NULL) ; }
SLIP_info SLIP_build_sparse_trip_double (SLIP_sparse *A, int32_t *I, int32_t *J, double *x, int32_t n, int32_t nz, SLIP_options *option)
{ return build_sparse_any (A, I, J, x, n, nz, SRC_DOUBLE, 1, option) ; }
SLIP_info SLIP_build_sparse_trip_mpq (SLIP_sparse *A, int32_t *I, int32_t *J, mpq_t *x, int32_t n, int32_t nz)
{ return build_sparse_any (A, I, J, x, n, nz, SRC_MPQ, 1, NULL) ; }

SLIP_info SLIP_build_sparse_ccf_mpfr (SLIP_sparse *A, int32_t *p, int32_t *I, mpfr_t *x, int32_t n, int32_t nz, SLIP_options *option)
{ return option ? build_sparse_any (A, p, I, x, n, nz, SRC_MPFR, 0, option) : SLIP_INCORRECT_INPUT ; }
SLIP_info SLIP_build_sparse_trip_mpfr (SLIP_sparse *A, int32_t *I, int32_t *J, mpfr_t *x, int32_t n, int32_t nz, SLIP_options *option)
{ return option ? build_sparse_any (A, I, J, x, n, nz, SRC_MPFR, 1, option) : SLIP_INCORRECT_INPUT ; }

/* ---- dense builders ---- */
static SLIP_info dense_alloc (SLIP_dense *A, int32_t m, int32_t n)
{
    if (m <= 0 || n <= 0) return SLIP_INCORRECT_INPUT ;
    A->m = m ; A->n = n ;
    A->x = SLIP_create_mpz_mat (m, n) ;
    return A->x ? SLIP_OK : SLIP_OUT_OF_MEMORY ;
}

#define BAD_DENSE_ARGS(b, A) (!(b) || !(A) || !mpq_denref ((A)->scale)->_mp_d)

SLIP_info SLIP_build_dense_mpz (SLIP_dense *A, mpz_t **b, int32_t m, int32_t n)
{
    if (BAD_DENSE_ARGS (b, A)) return SLIP_INCORRECT_INPUT ;
    SLIP_info status = dense_alloc (A, m, n) ;
    if (status != SLIP_OK) return status ;
    for (int32_t i = 0 ; i < m ; i++)
        for (int32_t j = 0 ; j < n ; j++) mpz_set (A->x [i][j], b [i][j]) ;
    mpq_set_ui (A->scale, 1, 1) ;
    return SLIP_OK ;
}

SLIP_info SLIP_build_dense_int (SLIP_dense *A, int32_t **b, int32_t m, int32_t n)
{
    if (BAD_DENSE_ARGS (b, A)) return SLIP_INCORRECT_INPUT ;
    SLIP_info status = dense_alloc (A, m, n) ;
    if (status != SLIP_OK) return status ;
    for (int32_t i = 0 ; i < m ; i++)
        for (int32_t j = 0 ; j < n ; j++) mpz_set_si (A->x [i][j], b [i][j]) ;
    mpq_set_ui (A->scale, 1, 1) ;
    return SLIP_OK ;
}

SLIP_info SLIP_build_dense_double (SLIP_dense *A, double **b, int32_t m, int32_t n, SLIP_options *option)
{
    if (BAD_DENSE_ARGS (b, A)) return SLIP_INCORRECT_INPUT ;
    SLIP_info status = dense_alloc (A, m, n) ;
    if (status != SLIP_OK) return status ;
    return slip_expand_double_mat (A->x, b, A->scale, m, n, option) ;
}

SLIP_info SLIP_build_dense_mpq (SLIP_dense *A, mpq_t **b, int32_t m, int32_t n)
{
    if (BAD_DENSE_ARGS (b, A)) return SLIP_INCORRECT_INPUT ;
    SLIP_info status = dense_alloc (A, m, n) ;
    if (status != SLIP_OK) return status ;
    return slip_expand_mpq_mat (A->x, b, A->scale, m, n) ;
}

SLIP_info SLIP_build_dense_mpfr (SLIP_dense *A, mpfr_t **b, int32_t m, int32_t n, SLIP_options *option)
{
    if (BAD_DENSE_ARGS (b, A) || !option) return SLIP_INCORRECT_INPUT ;
    SLIP_info status = dense_alloc (A, m, n) ;
    if (status != SLIP_OK) return status ;
    return slip_expand_mpfr_mat (A->x, b, A->scale, m, n, option) ;
}
