/* slip_objects.c -- constructors and destructors of the interface's containers.
 * Mirrors SLIP_LU/Source/SLIP_create_*.c and SLIP_delete_*.c (one reference file per function). */
#include "slip_internal.h"

SLIP_options *SLIP_create_default_options (void)
{
    /* defaults of SLIP_LU_internal.h:103-135 */
    SLIP_options *o = (SLIP_options *) SLIP_malloc (sizeof (SLIP_options)) ;
    if (!o) return NULL ;
    o->pivot = SLIP_TOL_SMALLEST ;
    o->order = SLIP_COLAMD ;
    o->tol = 1 ;
    o->print_level = 0 ;
    o->prec = 128 ;
    o->SLIP_MPFR_ROUND = MPFR_RNDN ;
    return o ;
}

SLIP_sparse *SLIP_create_sparse (void)
{
    SLIP_sparse *A = (SLIP_sparse *) SLIP_calloc (1, sizeof (SLIP_sparse)) ;
    if (!A) return NULL ;
    mpq_init (A->scale) ;
    mpq_set_ui (A->scale, 1, 1) ;
    return A ;
}

void SLIP_delete_sparse (SLIP_sparse **A)
{
    if (!A || !*A) return ;
    SLIP_sparse *M = *A ;
    if (M->x) slip_resident_drop (M->x) ;         /* GPU-resident factors keyed by this object */
    SLIP_delete_mpz_array (&M->x, M->nzmax) ;
    SLIP_FREE (M->i) ;
    SLIP_FREE (M->p) ;
    if (mpq_denref (M->scale)->_mp_d) mpq_clear (M->scale) ;
    SLIP_free (M) ;
    *A = NULL ;
}

SLIP_dense *SLIP_create_dense (void)
{
    SLIP_dense *A = (SLIP_dense *) SLIP_calloc (1, sizeof (SLIP_dense)) ;
    if (!A) return NULL ;
    mpq_init (A->scale) ;
    mpq_set_ui (A->scale, 1, 1) ;
    return A ;
}

void SLIP_delete_dense (SLIP_dense **A)
{
    if (!A || !*A) return ;
    SLIP_delete_mpz_mat (&(*A)->x, (*A)->m, (*A)->n) ;
    if (mpq_denref ((*A)->scale)->_mp_d) mpq_clear ((*A)->scale) ;
    SLIP_free (*A) ;
    *A = NULL ;
}

SLIP_LU_analysis *SLIP_create_LU_analysis (int32_t n)
{
    SLIP_LU_analysis *S = (SLIP_LU_analysis *) SLIP_malloc (sizeof (SLIP_LU_analysis)) ;
    if (!S) return NULL ;
    S->q = (int32_t *) SLIP_malloc ((size_t) (n > 0 ? n : 1) * sizeof (int32_t)) ;
    if (!S->q) { SLIP_free (S) ; return NULL ; }
    S->lnz = S->unz = 0 ;
    return S ;
}

void SLIP_delete_LU_analysis (SLIP_LU_analysis **S)
{
    if (!S || !*S) return ;
    SLIP_FREE ((*S)->q) ;
    SLIP_free (*S) ;
    *S = NULL ;
}

/* ---- 1D arrays ---- */
mpz_t *SLIP_create_mpz_array (int32_t n)
{
    if (n <= 0) return NULL ;
    mpz_t *x = (mpz_t *) SLIP_calloc ((size_t) n, sizeof (mpz_t)) ;
    if (!x) return NULL ;
    for (int32_t k = 0 ; k < n ; k++) mpz_init (x [k]) ;
    return x ;
}

void SLIP_delete_mpz_array (mpz_t **x, int32_t n)
{
    if (!x || !*x) return ;
    for (int32_t k = 0 ; k < n ; k++)
        if ((*x) [k]->_mp_d) { mpz_clear ((*x) [k]) ; (*x) [k]->_mp_d = NULL ; }
    SLIP_free (*x) ;
    *x = NULL ;
}

mpq_t *SLIP_create_mpq_array (int32_t n)
{
    if (n <= 0) return NULL ;
    mpq_t *x = (mpq_t *) SLIP_calloc ((size_t) n, sizeof (mpq_t)) ;
    if (!x) return NULL ;
    for (int32_t k = 0 ; k < n ; k++) mpq_init (x [k]) ;
    return x ;
}

void SLIP_delete_mpq_array (mpq_t **x, int32_t n)
{
    if (!x || !*x) return ;
    for (int32_t k = 0 ; k < n ; k++)
        if (mpq_denref ((*x) [k])->_mp_d) mpq_clear ((*x) [k]) ;
    SLIP_free (*x) ;
    *x = NULL ;
}

/* ---- 2D matrices: A[i] is row i ---- */
mpz_t **SLIP_create_mpz_mat (int32_t m, int32_t n)
{
    if (m <= 0 || n <= 0) return NULL ;
    mpz_t **A = (mpz_t **) SLIP_calloc ((size_t) m, sizeof (mpz_t *)) ;
    if (!A) return NULL ;
    for (int32_t i = 0 ; i < m ; i++)
    {
        A [i] = SLIP_create_mpz_array (n) ;
        if (!A [i]) { SLIP_delete_mpz_mat (&A, m, n) ; return NULL ; }
    }
    return A ;
}

void SLIP_delete_mpz_mat (mpz_t ***A, int32_t m, int32_t n)
{
    if (!A || !*A) return ;
    for (int32_t i = 0 ; i < m ; i++) SLIP_delete_mpz_array (&(*A) [i], n) ;
    SLIP_free (*A) ;
    *A = NULL ;
}

mpq_t **SLIP_create_mpq_mat (int32_t m, int32_t n)
{
    if (m <= 0 || n <= 0) return NULL ;
    mpq_t **A = (mpq_t **) SLIP_calloc ((size_t) m, sizeof (mpq_t *)) ;
    if (!A) return NULL ;
    for (int32_t i = 0 ; i < m ; i++)
    {
        A [i] = SLIP_create_mpq_array (n) ;
        if (!A [i]) { SLIP_delete_mpq_mat (&A, m, n) ; return NULL ; }
    }
    return A ;
}

void SLIP_delete_mpq_mat (mpq_t ***A, int32_t m, int32_t n)
{
    if (!A || !*A) return ;
    for (int32_t i = 0 ; i < m ; i++) SLIP_delete_mpq_array (&(*A) [i], n) ;
    SLIP_free (*A) ;
    *A = NULL ;
}

/* ---- mpfr containers at option->prec (SLIP_create_mpfr_array.c, SLIP_create_mpfr_mat.c,
 * SLIP_delete_mpfr_array.c, SLIP_delete_mpfr_mat.c) ---- */
mpfr_t *SLIP_create_mpfr_array (int32_t n, SLIP_options *option)
{
    if (n <= 0 || !option) return NULL ;
    mpfr_t *x = (mpfr_t *) SLIP_calloc ((size_t) n, sizeof (mpfr_t)) ;
    if (!x) return NULL ;
    for (int32_t k = 0 ; k < n ; k++) mpfr_init2 (x [k], (mpfr_prec_t) option->prec) ;
    return x ;
}

void SLIP_delete_mpfr_array (mpfr_t **x, int32_t n)
{
    if (!x || !*x) return ;
    for (int32_t k = 0 ; k < n ; k++)
        if ((*x) [k]->_mpfr_d) mpfr_clear ((*x) [k]) ;
    SLIP_free (*x) ;
    *x = NULL ;
}

mpfr_t **SLIP_create_mpfr_mat (int32_t m, int32_t n, SLIP_options *option)
{
    if (m <= 0 || n <= 0 || !option) return NULL ;
    mpfr_t **A = (mpfr_t **) SLIP_calloc ((size_t) m, sizeof (mpfr_t *)) ;
    if (!A) return NULL ;
    for (int32_t i = 0 ; i < m ; i++)
    {
        A [i] = SLIP_create_mpfr_array (n, option) ;
        if (!A [i]) { SLIP_delete_mpfr_mat (&A, m, n) ; return NULL ; }
    }
    return A ;
}

void SLIP_delete_mpfr_mat (mpfr_t ***A, int32_t m, int32_t n)
{
    if (!A || !*A) return ;
    for (int32_t i = 0 ; i < m ; i++) SLIP_delete_mpfr_array (&(*A) [i], n) ;
    SLIP_free (*A) ;
    *A = NULL ;
}

double **SLIP_create_double_mat (int32_t m, int32_t n)
{
    if (m <= 0 || n <= 0) return NULL ;
    double **A = (double **) SLIP_calloc ((size_t) m, sizeof (double *)) ;
    if (!A) return NULL ;
    for (int32_t i = 0 ; i < m ; i++)
    {
        A [i] = (double *) SLIP_calloc ((size_t) n, sizeof (double)) ;
        if (!A [i]) { SLIP_delete_double_mat (&A, m, n) ; return NULL ; }
    }
    return A ;
}

void SLIP_delete_double_mat (double ***A, int32_t m, int32_t n)
{
    (void) n ;
    if (!A || !*A) return ;
    for (int32_t i = 0 ; i < m ; i++) SLIP_free ((*A) [i]) ;
    SLIP_free (*A) ;
    *A = NULL ;
}

int32_t **SLIP_create_int_mat (int32_t m, int32_t n)
{
    if (m <= 0 || n <= 0) return NULL ;
    int32_t **A = (int32_t **) SLIP_calloc ((size_t) m, sizeof (int32_t *)) ;
    if (!A) return NULL ;
    for (int32_t i = 0 ; i < m ; i++)
    {
        A [i] = (int32_t *) SLIP_calloc ((size_t) n, sizeof (int32_t)) ;
        if (!A [i]) { SLIP_delete_int_mat (&A, m, n) ; return NULL ; }
    }
    return A ;
}

void SLIP_delete_int_mat (int32_t ***A, int32_t m, int32_t n)
{
    (void) n ;
    if (!A || !*A) return ;
    for (int32_t i = 0 ; i < m ; i++) SLIP_free ((*A) [i]) ;
    SLIP_free (*A) ;
    *A = NULL ;
}
