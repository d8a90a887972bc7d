/* slip_internal.h -- internal declarations of the C host layer of slip_lu_b200.
 * Not for client code; include SLIP_LU.h instead. */
#ifndef SLIP_B200_INTERNAL_H
#define SLIP_B200_INTERNAL_H

#include <math.h>
#include <stdarg.h>
#include "SLIP_LU.h"
#include "slip_b200_device.h"

#define SLIP_TRY(call) do { SLIP_info ok_ = (call) ; if (ok_ != SLIP_OK) { status = ok_ ; goto cleanup ; } } while (0)

/* bits carried by one residue channel, rounded down: every channel prime is within 2^21 of 2^31 */
#define SLIP_B200_CHANNEL_BITS 30.999
/* spare channels kept beyond the factorization bound, so that right-hand sides whose entries are
 * up to ~250 bits larger than the smallest column of A solve without re-encoding the factors */
#define SLIP_B200_SPARE_CHANNELS 8

/* limb-string export of GMP integers for the device layer */
typedef struct
{
    uint32_t *limbs ;     /* all values back to back, little-endian 32-bit words */
    int64_t *off ;        /* count+1 offsets into limbs */
    int8_t *sign ;        /* -1, 0, 1 */
    int64_t count ;
} slip_limbs ;

void slip_limbs_free (slip_limbs *s) ;
SLIP_info slip_limbs_begin (slip_limbs *s, int64_t count, int64_t words) ;
void slip_limbs_put (slip_limbs *s, int64_t k, mpz_srcptr z) ;     /* call with k = 0,1,2,... */
int64_t slip_mpz_words (mpz_srcptr z) ;
/* value <- sign * limbs[0..n32) (stride-padded source: n32 rounded up to even is readable) */
void slip_mpz_from_words (mpz_ptr z, const uint32_t *limbs, int32_t n32, int sign) ;

SLIP_info slip_from_device_status (int rc) ;

/* statistics of the last factorization of this thread (see SLIP_B200_last_stats) */
typedef struct
{
    double n, nnz_L, nnz_U, channels ;
    double updates ;           /* REF entry updates (one per L entry per elimination step) */
    double limb_mul_equiv ;    /* schoolbook-equivalent 32-bit limb multiplies of those updates */
    double t_symbolic, t_device, t_begin, t_factor_total ;
    double channels_hadamard ; /* channels the a-priori (Hadamard) bound asks for */
    double bound_restarts ;    /* bound-mode restarts with more channels */
    double verified_solves ;   /* solves whose numerators were verified exactly (A N = det b) */
} slip_b200_stats ;
extern __thread slip_b200_stats slip_last_stats ;
void slip_set_error (const char *msg) ;

/* sizing: upper bound of log2 ||A(:,j)||_2 for every column, and of a dense matrix' columns */
SLIP_info slip_column_bits (const SLIP_sparse *A, double *bits) ;
double slip_dense_max_column_bits (const SLIP_dense *b) ;
int slip_channels_for_bits (double bits) ;

/* resident factorizations (GPU-side L, U, rho), keyed by the host L object */
typedef struct slip_resident
{
    const void *Lx, *Ux ;         /* L->x, U->x of the owning factorization (NULL: anonymous) */
    int holders, unlinked ;       /* registry reference count (see slip_limbs.c) */
    slipcu_factor *dev ;
    int32_t n ;
    double total_bits ;           /* sum of the column bounds of A */
    double min_col_bits ;
    int proven_channels ;         /* 1: the session's channels cover the Hadamard bound */
    SLIP_sparse *A_copy ;         /* bound-mode factors kept for SLIP_LU_solve: the input, to verify solutions */
    int32_t *q_copy ;
    mpz_t det ;                   /* rho[n-1] */
    struct slip_resident *next ;
} slip_resident ;

slip_resident *slip_resident_acquire (const void *Lx, const void *Ux, int32_t n, mpz_srcptr det) ;
void slip_resident_release (slip_resident *r) ;
void slip_resident_drop_all (void) ;
void slip_resident_add (slip_resident *r) ;
void slip_resident_drop (const void *Lx) ;          /* frees device memory */
void slip_resident_free (slip_resident *r) ;

/* the shared driver behind SLIP_LU_factorize and SLIP_solve_* */
SLIP_info slip_factorize_driver (SLIP_sparse *L, SLIP_sparse *U, SLIP_sparse *A, SLIP_LU_analysis *S,
    mpz_t *rhos, int32_t *pinv, SLIP_options *option, int want_host_factors, slip_resident **resident,
    double rhs_bits, int min_channels) ;
/* solve with resident factors; A, q (may be NULL) allow a bound-mode session to verify its result.
 * Returns SLIP_B200_NEED_CHANNELS when the session's channels could not be shown to suffice. */
#define SLIP_B200_NEED_CHANNELS ((SLIP_info) (-100))
SLIP_info slip_solve_resident (mpq_t **x, SLIP_dense *b, slip_resident *r, const int32_t *pinv,
    const SLIP_sparse *A, const int32_t *q) ;
SLIP_sparse *slip_sparse_copy (const SLIP_sparse *A) ;

SLIP_info slip_expand_double_array (mpz_t *x_out, double *x, mpq_t scale, int32_t n, SLIP_options *option) ;
SLIP_info slip_expand_double_mat (mpz_t **x_out, double **x, mpq_t scale, int32_t m, int32_t n, SLIP_options *option) ;
SLIP_info slip_expand_mpq_array (mpz_t *x_out, mpq_t *x, mpq_t scale, int32_t n) ;
SLIP_info slip_expand_mpq_mat (mpz_t **x_out, mpq_t **x, mpq_t scale, int32_t m, int32_t n) ;
SLIP_info slip_expand_mpfr_array (mpz_t *x_out, mpfr_t *x, mpq_t scale, int32_t n, SLIP_options *option) ;
SLIP_info slip_expand_mpfr_mat (mpz_t **x_out, mpfr_t **x, mpq_t scale, int32_t m, int32_t n, SLIP_options *option) ;
SLIP_info slip_sparse_from_ccf (SLIP_sparse *A, const int32_t *p, const int32_t *I, mpz_t *x, int32_t n, int32_t nz) ;
SLIP_info slip_sparse_from_trip (SLIP_sparse *A, const int32_t *I, const int32_t *J, mpz_t *x, int32_t n, int32_t nz) ;

#endif
