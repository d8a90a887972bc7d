/* slip_limbs.c -- GMP integers <-> limb strings of the device ABI, sizing bounds, and the
 * registry of GPU-resident factorizations. */
#include <pthread.h>
#include "slip_internal.h"

/* ---- export ---- */
int64_t slip_mpz_words (mpz_srcptr z)
{
    size_t nl = mpz_size (z) ;
    if (nl == 0) return 0 ;
    mp_limb_t top = mpz_getlimbn (z, (mp_size_t) nl - 1) ;
    return (int64_t) (2 * nl) - ((top >> 32) == 0 ? 1 : 0) ;
}

SLIP_info slip_limbs_begin (slip_limbs *s, int64_t count, int64_t words)
{
    s->count = count ;
    s->limbs = (uint32_t *) SLIP_malloc ((size_t) (words > 0 ? words : 1) * sizeof (uint32_t)) ;
    s->off = (int64_t *) SLIP_malloc ((size_t) (count + 1) * sizeof (int64_t)) ;
    s->sign = (int8_t *) SLIP_malloc ((size_t) (count > 0 ? count : 1)) ;
    if (!s->limbs || !s->off || !s->sign) { slip_limbs_free (s) ; return SLIP_OUT_OF_MEMORY ; }
    s->off [0] = 0 ;
    return SLIP_OK ;
}

void slip_limbs_put (slip_limbs *s, int64_t k, mpz_srcptr z)
{
    const int64_t w = slip_mpz_words (z) ;
    const int64_t at = s->off [k] ;
    if (w > 0) memcpy (s->limbs + at, mpz_limbs_read (z), (size_t) w * sizeof (uint32_t)) ;
    s->off [k + 1] = at + w ;
    s->sign [k] = (int8_t) mpz_sgn (z) ;
}

void slip_limbs_free (slip_limbs *s)
{
    SLIP_free (s->limbs) ; SLIP_free (s->off) ; SLIP_free (s->sign) ;
    s->limbs = NULL ; s->off = NULL ; s->sign = NULL ;
}

/* ---- import ---- */
void slip_mpz_from_words (mpz_ptr z, const uint32_t *limbs, int32_t n32, int sign)
{
    if (n32 <= 0 || sign == 0) { mpz_set_ui (z, 0) ; return ; }
    const mp_size_t nl = ((mp_size_t) n32 + 1) / 2 ;
    mp_limb_t *d = mpz_limbs_write (z, nl) ;
    memcpy (d, limbs, (size_t) nl * sizeof (mp_limb_t)) ;   /* the odd top word is zero padded */
    if (n32 & 1) d [nl - 1] &= 0xffffffffUL ;
    mpz_limbs_finish (z, sign < 0 ? -nl : nl) ;
}

/* ---- sizing ----
 * Every entry produced while eliminating the first k+1 columns is a minor of those columns, so
 * by Hadamard |entry| <= prod_j ||A(:,j)||_2 (column norms are >= 1 for integer columns).  The
 * per-column bound is log2 of the norm, rounded up. */
static double log2_upper (mpz_srcptr v)
{
    if (mpz_sgn (v) == 0) return 0.0 ;
    signed long e ;
    double d = mpz_get_d_2exp (&e, v) ;        /* v = d * 2^e, 0.5 <= |d| < 1, truncated */
    return log2 (fabs (d) + 1e-15) + (double) e + 1e-9 ;
}

SLIP_info slip_column_bits (const SLIP_sparse *A, double *bits)
{
    mpz_t ss ;
    mpz_init (ss) ;
    for (int32_t j = 0 ; j < A->n ; j++)
    {
        mpz_set_ui (ss, 0) ;
        for (int32_t a = A->p [j] ; a < A->p [j + 1] ; a++) mpz_addmul (ss, A->x [a], A->x [a]) ;
        double b = 0.5 * log2_upper (ss) ;
        bits [j] = b > 0.0 ? b : 0.0 ;
    }
    mpz_clear (ss) ;
    return SLIP_OK ;
}

double slip_dense_max_column_bits (const SLIP_dense *b)
{
    mpz_t ss ;
    mpz_init (ss) ;
    double best = 0.0 ;
    for (int32_t c = 0 ; c < b->n ; c++)
    {
        mpz_set_ui (ss, 0) ;
        for (int32_t r = 0 ; r < b->m ; r++) mpz_addmul (ss, b->x [r][c], b->x [r][c]) ;
        double v = 0.5 * log2_upper (ss) ;
        if (v > best) best = v ;
    }
    mpz_clear (ss) ;
    return best ;
}

int slip_channels_for_bits (double bits)
{
    /* the channel product must exceed twice the bound (signed range): +1 bit, +1 guard */
    double c = ceil ((bits + 2.0) / SLIP_B200_CHANNEL_BITS) ;
    if (c < 1) c = 1 ;
    return (int) c ;
}

/* ---- resident factorizations ----
 * SLIP_LU_factorize leaves L, U and the pivots on the GPU; SLIP_LU_solve finds them again through
 * this registry.  An entry is keyed by the addresses of L->x AND U->x together with n, and a hit
 * is only used if rhos[n-1] equals the determinant the entry was built with, so that an array
 * released by other means and reallocated at the same address is not mistaken for the factors.
 * Entries are reference counted: a solve holds its entry, and a concurrent SLIP_delete_sparse only
 * unlinks it; the last holder frees the device memory.  SLIP_finalize drops everything. */
static pthread_mutex_t reg_lock = PTHREAD_MUTEX_INITIALIZER ;
static slip_resident *reg_head = NULL ;

slip_resident *slip_resident_acquire (const void *Lx, const void *Ux, int32_t n, mpz_srcptr det)
{
    slip_resident *hit = NULL ;
    if (!Lx) return NULL ;
    pthread_mutex_lock (&reg_lock) ;
    for (slip_resident *r = reg_head ; r ; r = r->next)
        if (r->Lx == Lx && r->Ux == Ux && r->n == n && det && mpz_cmp (r->det, det) == 0) { hit = r ; r->holders++ ; break ; }
    pthread_mutex_unlock (&reg_lock) ;
    return hit ;
}

void slip_resident_release (slip_resident *r)
{
    if (!r) return ;
    int last ;
    pthread_mutex_lock (&reg_lock) ;
    last = (--r->holders == 0 && r->unlinked) ;
    pthread_mutex_unlock (&reg_lock) ;
    if (last) slip_resident_free (r) ;
}

void slip_resident_add (slip_resident *r)
{
    slip_resident *old = NULL ;
    pthread_mutex_lock (&reg_lock) ;
    /* a second factorization into the same L object replaces the first */
    for (slip_resident **pp = &reg_head ; *pp ; pp = &(*pp)->next)
        if ((*pp)->Lx == r->Lx) { old = *pp ; *pp = old->next ; old->unlinked = 1 ; if (old->holders) old = NULL ; break ; }
    r->next = reg_head ;
    reg_head = r ;
    pthread_mutex_unlock (&reg_lock) ;
    slip_resident_free (old) ;
}

void slip_resident_free (slip_resident *r)
{
    if (!r) return ;
    if (r->dev) slipcu_factor_free (r->dev) ;
    if (r->det->_mp_d) mpz_clear (r->det) ;
    if (r->A_copy) SLIP_delete_sparse (&r->A_copy) ;
    SLIP_free (r->q_copy) ;
    SLIP_free (r) ;
}

void slip_resident_drop (const void *Lx)
{
    if (!Lx) return ;
    slip_resident *victim = NULL ;
    pthread_mutex_lock (&reg_lock) ;
    for (slip_resident **pp = &reg_head ; *pp ; pp = &(*pp)->next)
        if ((*pp)->Lx == Lx)
        {
            victim = *pp ; *pp = victim->next ; victim->unlinked = 1 ;
            if (victim->holders) victim = NULL ;          /* the solve that holds it frees it */
            break ;
        }
    pthread_mutex_unlock (&reg_lock) ;
    slip_resident_free (victim) ;
}

void slip_resident_drop_all (void)
{
    for (;;)
    {
        slip_resident *victim = NULL ;
        pthread_mutex_lock (&reg_lock) ;
        if (reg_head) { victim = reg_head ; reg_head = victim->next ; victim->unlinked = 1 ; if (victim->holders) victim = (slip_resident *) 1 ; }
        pthread_mutex_unlock (&reg_lock) ;
        if (!victim) break ;
        if (victim != (slip_resident *) 1) slip_resident_free (victim) ;
    }
}
