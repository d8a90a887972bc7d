/* SLIP_LU.h -- public C interface of slip_lu_b200 (libslip_lu_b200.so).
 *
 * Drop-in for the factor/solve path of the reference header
 * (cjh10644/SLIP_LU, SLIP_LU/Include/SLIP_LU.h): same type names, struct layouts, enum
 * values, function names, argument order and return codes, so that a program written
 * against the reference links against this library unchanged.  Each declaration cites the
 * line of the reference header it mirrors ("ref:NNN").  The exact arithmetic behind
 * SLIP_LU_factorize / SLIP_LU_solve / SLIP_solve_* runs on an NVIDIA B200 (sm_100a); the
 * library has no CPU path for it and returns an error if no device is present.
 *
 * Not provided: this fork's experimental SLIP_LU_analyze_and_factorize{,1} (ref:866-886).  See
 * DESIGN.md "scope".
 */
#ifndef SLIP_Include
#define SLIP_Include

#include <stdio.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <gmp.h>
#include <mpfr.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIP_LU_VERSION "1.0.0"          /* interface version mirrored (ref:141-144) */
#define SLIP_LU_VERSION_MAJOR 1
#define SLIP_LU_VERSION_MINOR 0
#define SLIP_LU_VERSION_SUB   0
#define SLIP_LU_B200 1                   /* lets client code detect this implementation */

/* ---- status codes (ref:160-168) ---- */
typedef enum
{
    SLIP_OK = 0,
    SLIP_OUT_OF_MEMORY = -1,
    SLIP_SINGULAR = -2,
    SLIP_INCORRECT_INPUT = -3,
    SLIP_INCORRECT = -4
}
SLIP_info ;

/* ---- pivoting rules (ref:176-187) ---- */
typedef enum
{
    SLIP_SMALLEST = 0,
    SLIP_DIAGONAL = 1,
    SLIP_FIRST_NONZERO = 2,
    SLIP_TOL_SMALLEST = 3,      /* default */
    SLIP_TOL_LARGEST = 4,
    SLIP_LARGEST = 5
}
SLIP_pivot ;

/* ---- column orderings (ref:195-201) ---- */
typedef enum
{
    SLIP_NO_ORDERING = 0,
    SLIP_COLAMD = 1,            /* default */
    SLIP_AMD = 2
}
SLIP_col_order ;

/* ---- options (ref:212-223) ---- */
typedef struct SLIP_options
{
    SLIP_pivot pivot ;
    SLIP_col_order order ;
    double tol ;
    int32_t print_level ;
    uint64_t prec ;
    mpfr_rnd_t SLIP_MPFR_ROUND ;
} SLIP_options ;

SLIP_options *SLIP_create_default_options (void) ;                 /* ref:229 */

/* ---- compressed-column matrix of mpz_t (ref:246-256) ---- */
typedef struct
{
    int32_t m ;
    int32_t n ;
    int32_t nzmax ;
    int32_t nz ;
    int32_t *p ;
    int32_t *i ;
    mpz_t *x ;
    mpq_t scale ;
} SLIP_sparse ;

SLIP_sparse *SLIP_create_sparse (void) ;                           /* ref:261 */
void SLIP_delete_sparse (SLIP_sparse **A) ;                        /* ref:264 */

/* ---- dense matrix of mpz_t, x[i][j] (ref:277-284) ---- */
typedef struct
{
    int32_t m ;
    int32_t n ;
    mpz_t **x ;
    mpq_t scale ;
} SLIP_dense ;

SLIP_dense *SLIP_create_dense (void) ;                             /* ref:287 */
void SLIP_delete_dense (SLIP_dense **A) ;                          /* ref:290 */

/* ---- symbolic analysis (ref:303-310) ---- */
typedef struct
{
    int32_t *q ;
    int32_t lnz ;
    int32_t unz ;
} SLIP_LU_analysis ;

SLIP_LU_analysis *SLIP_create_LU_analysis (int32_t n) ;            /* ref:316 */
void SLIP_delete_LU_analysis (SLIP_LU_analysis **S) ;              /* ref:326 */

/* ---- memory (ref:340-387) ---- */
void *SLIP_calloc (size_t n, size_t size) ;
void *SLIP_malloc (size_t size) ;
void *SLIP_realloc (void *p, size_t old_size, size_t new_size) ;
void SLIP_free (void *p) ;
#define SLIP_FREE(p) { SLIP_free (p) ; (p) = NULL ; }

/* ---- input builders (ref:405-438, 478-509, 548-573) ---- */
SLIP_info SLIP_build_sparse_ccf_mpz (SLIP_sparse *A_output, int32_t *p, int32_t *I, mpz_t *x,
    int32_t n, int32_t nz) ;
SLIP_info SLIP_build_sparse_ccf_double (SLIP_sparse *A_output, int32_t *p, int32_t *I, double *x,
    int32_t n, int32_t nz, SLIP_options *option) ;
SLIP_info SLIP_build_sparse_ccf_int (SLIP_sparse *A_output, int32_t *p, int32_t *I, int32_t *x,
    int32_t n, int32_t nz) ;
SLIP_info SLIP_build_sparse_ccf_mpq (SLIP_sparse *A_output, int32_t *p, int32_t *I, mpq_t *x,
    int32_t n, int32_t nz) ;
SLIP_info SLIP_build_sparse_trip_mpz (SLIP_sparse *A_output, int32_t *I, int32_t *J, mpz_t *x,
    int32_t n, int32_t nz) ;
SLIP_info SLIP_build_sparse_trip_double (SLIP_sparse *A_output, int32_t *I, int32_t *J, double *x,
    int32_t n, int32_t nz, SLIP_options *option) ;
SLIP_info SLIP_build_sparse_trip_int (SLIP_sparse *A_output, int32_t *I, int32_t *J, int32_t *x,
    int32_t n, int32_t nz) ;
SLIP_info SLIP_build_sparse_trip_mpq (SLIP_sparse *A_output, int32_t *I, int32_t *J, mpq_t *x,
    int32_t n, int32_t nz) ;
SLIP_info SLIP_build_sparse_ccf_mpfr (SLIP_sparse *A_output, int32_t *p, int32_t *I, mpfr_t *x,
    int32_t n, int32_t nz, SLIP_options *option) ;                 /* ref:454-463 */
SLIP_info SLIP_build_sparse_trip_mpfr (SLIP_sparse *A_output, int32_t *I, int32_t *J, mpfr_t *x,
    int32_t n, int32_t nz, SLIP_options *option) ;                 /* ref:523-532 */
SLIP_info SLIP_build_dense_mpz (SLIP_dense *A_output, mpz_t **b, int32_t m, int32_t n) ;
SLIP_info SLIP_build_dense_double (SLIP_dense *A_output, double **b, int32_t m, int32_t n,
    SLIP_options *option) ;
SLIP_info SLIP_build_dense_int (SLIP_dense *A_output, int32_t **b, int32_t m, int32_t n) ;
SLIP_info SLIP_build_dense_mpq (SLIP_dense *A_output, mpq_t **b, int32_t m, int32_t n) ;
SLIP_info SLIP_build_dense_mpfr (SLIP_dense *A_output, mpfr_t **b, int32_t m, int32_t n,
    SLIP_options *option) ;                                        /* ref:585-592 */

/* ---- 2D and 1D helper containers (ref:602-797) ---- */
double **SLIP_create_double_mat (int32_t m, int32_t n) ;
void SLIP_delete_double_mat (double ***A, int32_t m, int32_t n) ;
int32_t **SLIP_create_int_mat (int32_t m, int32_t n) ;
void SLIP_delete_int_mat (int32_t ***A, int32_t m, int32_t n) ;
mpq_t **SLIP_create_mpq_mat (int32_t m, int32_t n) ;
void SLIP_delete_mpq_mat (mpq_t ***A, int32_t m, int32_t n) ;
mpz_t **SLIP_create_mpz_mat (int32_t m, int32_t n) ;
void SLIP_delete_mpz_mat (mpz_t ***A, int32_t m, int32_t n) ;
mpfr_t **SLIP_create_mpfr_mat (int32_t m, int32_t n, SLIP_options *option) ;
void SLIP_delete_mpfr_mat (mpfr_t ***A, int32_t m, int32_t n) ;
mpfr_t *SLIP_create_mpfr_array (int32_t n, SLIP_options *option) ;
void SLIP_delete_mpfr_array (mpfr_t **x, int32_t n) ;
mpq_t *SLIP_create_mpq_array (int32_t n) ;
void SLIP_delete_mpq_array (mpq_t **x, int32_t n) ;
mpz_t *SLIP_create_mpz_array (int32_t n) ;
void SLIP_delete_mpz_array (mpz_t **x, int32_t n) ;

/* ---- environment (ref:809-823) ---- */
void SLIP_initialize (void) ;
void SLIP_initialize_expert (void *(*MyMalloc) (size_t),
    void *(*MyRealloc) (void *, size_t, size_t), void (*MyFree) (void *, size_t)) ;
void SLIP_finalize (void) ;

/* ---- the factor / solve path (ref:834-993) ---- */
SLIP_info SLIP_LU_analyze (SLIP_LU_analysis *S, SLIP_sparse *A, SLIP_options *option) ;
SLIP_info SLIP_LU_factorize (SLIP_sparse *L, SLIP_sparse *U, SLIP_sparse *A, SLIP_LU_analysis *S,
    mpz_t *rhos, int32_t *pinv, SLIP_options *option) ;
SLIP_info SLIP_LU_solve (mpq_t **x, SLIP_dense *b, const mpz_t *rhos, const SLIP_sparse *L,
    const SLIP_sparse *U, const int32_t *pinv) ;
SLIP_info SLIP_solve_mpq (mpq_t **x_mpq, SLIP_sparse *A, SLIP_LU_analysis *S, SLIP_dense *b,
    SLIP_options *option) ;
SLIP_info SLIP_solve_double (double **x_doub, SLIP_sparse *A, SLIP_LU_analysis *S, SLIP_dense *b,
    SLIP_options *option) ;
SLIP_info SLIP_solve_mpfr (mpfr_t **x_mpfr, SLIP_sparse *A, SLIP_LU_analysis *S, SLIP_dense *b,
    SLIP_options *option) ;                                        /* ref:899-907 */
SLIP_info SLIP_permute_x (mpq_t **x, int32_t n, int32_t numRHS, SLIP_LU_analysis *S) ;
SLIP_info SLIP_scale_x (mpq_t **x, SLIP_sparse *A, SLIP_dense *b) ;
SLIP_info SLIP_check_solution (SLIP_sparse *A, mpq_t **x, SLIP_dense *b) ;
SLIP_info SLIP_get_double_soln (double **x_doub, mpq_t **x_mpq, int32_t n, int32_t numRHS) ;
SLIP_info SLIP_get_mpfr_soln (mpfr_t **x_mpfr, mpq_t **x_mpq, int32_t n, int32_t numRHS,
    SLIP_options *option) ;                                        /* ref:975-982 */
SLIP_info SLIP_spok (SLIP_sparse *A, SLIP_options *option) ;

/* ---- GMP / MPFR wrappers (ref:1023-1156, Source/SLIP_gmp.c) ----
 * Kept for source compatibility (the reference's demos read and print through them).  Here they
 * call GMP/MPFR and return SLIP_OK; the reference's setjmp guard against allocation failures inside
 * GMP is not reproduced. */
SLIP_info SLIP_gmp_fprintf (FILE *fp, const char *format, ...) ;
SLIP_info SLIP_gmp_printf (const char *format, ...) ;
SLIP_info SLIP_gmp_fscanf (FILE *fp, const char *format, ...) ;
SLIP_info SLIP_mpfr_fprintf (FILE *fp, const char *format, ...) ;
SLIP_info SLIP_mpz_init (mpz_t x) ;
SLIP_info SLIP_mpz_init2 (mpz_t x, const uint64_t size) ;
SLIP_info SLIP_mpz_init_set (mpz_t x, const mpz_t y) ;
SLIP_info SLIP_mpz_set (mpz_t x, const mpz_t y) ;
SLIP_info SLIP_mpz_set_ui (mpz_t x, const uint64_t y) ;
SLIP_info SLIP_mpz_set_si (mpz_t x, const int32_t y) ;
SLIP_info SLIP_mpz_set_q (mpz_t x, const mpq_t y) ;
SLIP_info SLIP_mpz_mul (mpz_t a, const mpz_t b, const mpz_t c) ;
SLIP_info SLIP_mpz_swap (mpz_t x, mpz_t y) ;
SLIP_info SLIP_mpz_submul (mpz_t x, const mpz_t y, const mpz_t z) ;
SLIP_info SLIP_mpz_divexact (mpz_t x, const mpz_t y, const mpz_t z) ;
SLIP_info SLIP_mpz_gcd (mpz_t x, const mpz_t y, const mpz_t z) ;
SLIP_info SLIP_mpz_lcm (mpz_t lcm, const mpz_t x, const mpz_t y) ;
SLIP_info SLIP_mpz_abs (mpz_t x, const mpz_t y) ;
SLIP_info SLIP_mpz_cmp (int32_t *r, const mpz_t x, const mpz_t y) ;
SLIP_info SLIP_mpz_cmpabs (int32_t *r, const mpz_t x, const mpz_t y) ;
SLIP_info SLIP_mpz_cmp_ui (int32_t *r, const mpz_t x, const uint64_t y) ;
SLIP_info SLIP_mpz_sgn (int32_t *sgn, const mpz_t x) ;
SLIP_info SLIP_mpz_sizeinbase (size_t *size, const mpz_t x, int32_t base) ;
SLIP_info SLIP_mpq_init (mpq_t x) ;
SLIP_info SLIP_mpq_set (mpq_t x, const mpq_t y) ;
SLIP_info SLIP_mpq_set_z (mpq_t x, const mpz_t y) ;
SLIP_info SLIP_mpq_set_d (mpq_t x, const double y) ;
SLIP_info SLIP_mpq_set_ui (mpq_t x, const uint64_t y, const uint64_t z) ;
SLIP_info SLIP_mpq_set_num (mpq_t x, const mpz_t y) ;
SLIP_info SLIP_mpq_set_den (mpq_t x, const mpz_t y) ;
SLIP_info SLIP_mpq_get_den (mpz_t x, const mpq_t y) ;
SLIP_info SLIP_mpq_get_d (double *x, const mpq_t y) ;
SLIP_info SLIP_mpq_abs (mpq_t x, const mpq_t y) ;
SLIP_info SLIP_mpq_add (mpq_t x, const mpq_t y, const mpq_t z) ;
SLIP_info SLIP_mpq_mul (mpq_t x, const mpq_t y, const mpq_t z) ;
SLIP_info SLIP_mpq_div (mpq_t x, const mpq_t y, const mpq_t z) ;
SLIP_info SLIP_mpq_cmp (int32_t *r, const mpq_t x, const mpq_t y) ;
SLIP_info SLIP_mpq_cmp_ui (int32_t *r, const mpq_t x, const uint64_t num, const uint64_t den) ;
SLIP_info SLIP_mpq_equal (int32_t *r, const mpq_t x, const mpq_t y) ;
SLIP_info SLIP_mpfr_init2 (mpfr_t x, const uint64_t size) ;
SLIP_info SLIP_mpfr_set_d (mpfr_t x, const double y, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_set_q (mpfr_t x, const mpq_t y, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_set_z (mpfr_t x, const mpz_t y, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_abs (mpfr_t x, const mpfr_t y, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_get_z (mpz_t x, const mpfr_t y, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_get_d (double *x, const mpfr_t y, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_mul (mpfr_t x, const mpfr_t y, const mpfr_t z, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_mul_d (mpfr_t x, const mpfr_t y, const double z, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_div_d (mpfr_t x, const mpfr_t y, const double z, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_ui_pow_ui (mpfr_t x, const uint64_t y, const uint64_t z, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_log2 (mpfr_t x, const mpfr_t y, const mpfr_rnd_t rnd) ;
SLIP_info SLIP_mpfr_free_cache (void) ;

/* ---- extensions of this implementation (not in the reference header) ----
 * The factors of the most recent SLIP_LU_factorize stay resident in GPU memory, keyed by the
 * L object, so that SLIP_LU_solve on the same L/U does not re-upload them; deleting L with
 * SLIP_delete_sparse releases them.  SLIP_B200_last_error gives the device-layer message
 * behind the last failing call of this thread. */
const char *SLIP_B200_last_error (void) ;
int SLIP_B200_device_count (void) ;
int SLIP_B200_set_device (int device) ;
/* work and timing of the calling thread's last factorization: n, nnz(L), nnz(U), channels, REF
 * entry updates, schoolbook-equivalent 32-bit limb multiplies, seconds (symbolic, device, set-up,
 * total).  Returns the number of values written. */
int SLIP_B200_last_stats (double *out, int cap) ;
/* [10] channels the a-priori (Hadamard) bound asks for, [11] restarts of the bound mode with more
 * channels, [12] right-hand sides whose numerators were verified exactly (A N = det b).
 * SLIP_B200_last_pinv: the row permutation chosen by the calling thread's last SLIP_solve_* call
 * (that path keeps L and U on the GPU; pinv is what it shares with SLIP_LU_factorize's output). */
int SLIP_B200_last_pinv (int32_t *out, int cap) ;

#ifdef __cplusplus
}
#endif
#endif
