/* slip_gmp_wrappers.c -- the SLIP_gmp_ / SLIP_mpz_ / SLIP_mpq_ / SLIP_mpfr_ functions of the
 * public interface (SLIP_LU/Include/SLIP_LU.h:1023-1156, SLIP_LU/Source/SLIP_gmp.c).
 *
 * In the reference these wrap every GMP/MPFR call of the library in a setjmp guard so that a
 * failed allocation inside GMP becomes SLIP_OUT_OF_MEMORY.  This library does not call GMP from
 * its hot path at all (the arithmetic runs on the GPU), so the wrappers exist only because user
 * code written against the reference calls them -- the reference's own demos read their input
 * with SLIP_gmp_fscanf and print with SLIP_gmp_fprintf / SLIP_mpfr_fprintf.  Each one performs the
 * operation and reports SLIP_OK; an allocation failure inside GMP ends as in plain GMP. */
#include <stdarg.h>
#include "slip_internal.h"

/* ---- formatted input / output ----
 * As in the reference (SLIP_gmp.c:310-452) these return the COUNT on success -- characters written,
 * fields stored -- and SLIP_INCORRECT_INPUT (negative) on failure, although the header types them
 * SLIP_info: the demos test `ok < 3` after reading a triplet. */
SLIP_info SLIP_gmp_fprintf (FILE *fp, const char *format, ...)
{
    va_list args ;
    va_start (args, format) ;
    int n = gmp_vfprintf (fp, format, args) ;
    va_end (args) ;
    return n < 0 ? SLIP_INCORRECT_INPUT : (SLIP_info) n ;
}

SLIP_info SLIP_gmp_printf (const char *format, ...)
{
    va_list args ;
    va_start (args, format) ;
    int n = gmp_vprintf (format, args) ;
    va_end (args) ;
    return n < 0 ? SLIP_INCORRECT_INPUT : (SLIP_info) n ;
}

SLIP_info SLIP_gmp_fscanf (FILE *fp, const char *format, ...)
{
    va_list args ;
    va_start (args, format) ;
    int n = gmp_vfscanf (fp, format, args) ;
    va_end (args) ;
    return n == EOF ? SLIP_INCORRECT_INPUT : (SLIP_info) n ;
}

SLIP_info SLIP_mpfr_fprintf (FILE *fp, const char *format, ...)
{
    va_list args ;
    va_start (args, format) ;
    int n = mpfr_vfprintf (fp, format, args) ;
    va_end (args) ;
    mpfr_free_cache () ;
    return n < 0 ? SLIP_INCORRECT_INPUT : (SLIP_info) n ;
}

/* ---- integers ---- */
SLIP_info SLIP_mpz_init (mpz_t x) { mpz_init (x) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_init2 (mpz_t x, const uint64_t size) { mpz_init2 (x, (mp_bitcnt_t) size) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_init_set (mpz_t x, const mpz_t y) { mpz_init_set (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_set (mpz_t x, const mpz_t y) { mpz_set (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_set_ui (mpz_t x, const uint64_t y) { mpz_set_ui (x, (unsigned long) y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_set_si (mpz_t x, const int32_t y) { mpz_set_si (x, (long) y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_set_q (mpz_t x, const mpq_t y) { mpz_set_q (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_mul (mpz_t a, const mpz_t b, const mpz_t c) { mpz_mul (a, b, c) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_swap (mpz_t x, mpz_t y) { mpz_swap (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_submul (mpz_t x, const mpz_t y, const mpz_t z) { mpz_submul (x, y, z) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_divexact (mpz_t x, const mpz_t y, const mpz_t z) { mpz_divexact (x, y, z) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_gcd (mpz_t x, const mpz_t y, const mpz_t z) { mpz_gcd (x, y, z) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_lcm (mpz_t lcm, const mpz_t x, const mpz_t y) { mpz_lcm (lcm, x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_abs (mpz_t x, const mpz_t y) { mpz_abs (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_cmp (int32_t *r, const mpz_t x, const mpz_t y) { *r = mpz_cmp (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_cmpabs (int32_t *r, const mpz_t x, const mpz_t y) { *r = mpz_cmpabs (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_cmp_ui (int32_t *r, const mpz_t x, const uint64_t y) { *r = mpz_cmp_ui (x, (unsigned long) y) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_sgn (int32_t *sgn, const mpz_t x) { *sgn = mpz_sgn (x) ; return SLIP_OK ; }
SLIP_info SLIP_mpz_sizeinbase (size_t *size, const mpz_t x, int32_t base) { *size = mpz_sizeinbase (x, base) ; return SLIP_OK ; }

/* ---- rationals ---- */
SLIP_info SLIP_mpq_init (mpq_t x) { mpq_init (x) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_set (mpq_t x, const mpq_t y) { mpq_set (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_set_z (mpq_t x, const mpz_t y) { mpq_set_z (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_set_d (mpq_t x, const double y) { mpq_set_d (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_set_ui (mpq_t x, const uint64_t y, const uint64_t z)
{ mpq_set_ui (x, (unsigned long) y, (unsigned long) z) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_set_num (mpq_t x, const mpz_t y) { mpq_set_num (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_set_den (mpq_t x, const mpz_t y) { mpq_set_den (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_get_den (mpz_t x, const mpq_t y) { mpq_get_den (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_get_d (double *x, const mpq_t y) { *x = mpq_get_d (y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_abs (mpq_t x, const mpq_t y) { mpq_abs (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_add (mpq_t x, const mpq_t y, const mpq_t z) { mpq_add (x, y, z) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_mul (mpq_t x, const mpq_t y, const mpq_t z) { mpq_mul (x, y, z) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_div (mpq_t x, const mpq_t y, const mpq_t z) { mpq_div (x, y, z) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_cmp (int32_t *r, const mpq_t x, const mpq_t y) { *r = mpq_cmp (x, y) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_cmp_ui (int32_t *r, const mpq_t x, const uint64_t num, const uint64_t den)
{ *r = mpq_cmp_ui (x, (unsigned long) num, (unsigned long) den) ; return SLIP_OK ; }
SLIP_info SLIP_mpq_equal (int32_t *r, const mpq_t x, const mpq_t y) { *r = mpq_equal (x, y) ; return SLIP_OK ; }

/* ---- floating point ---- */
SLIP_info SLIP_mpfr_init2 (mpfr_t x, const uint64_t size) { mpfr_init2 (x, (mpfr_prec_t) size) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_set_d (mpfr_t x, const double y, const mpfr_rnd_t rnd) { mpfr_set_d (x, y, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_set_q (mpfr_t x, const mpq_t y, const mpfr_rnd_t rnd) { mpfr_set_q (x, y, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_set_z (mpfr_t x, const mpz_t y, const mpfr_rnd_t rnd) { mpfr_set_z (x, y, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_abs (mpfr_t x, const mpfr_t y, const mpfr_rnd_t rnd) { mpfr_abs (x, y, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_get_z (mpz_t x, const mpfr_t y, const mpfr_rnd_t rnd) { mpfr_get_z (x, y, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_get_d (double *x, const mpfr_t y, const mpfr_rnd_t rnd) { *x = mpfr_get_d (y, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_mul (mpfr_t x, const mpfr_t y, const mpfr_t z, const mpfr_rnd_t rnd) { mpfr_mul (x, y, z, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_mul_d (mpfr_t x, const mpfr_t y, const double z, const mpfr_rnd_t rnd) { mpfr_mul_d (x, y, z, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_div_d (mpfr_t x, const mpfr_t y, const double z, const mpfr_rnd_t rnd) { mpfr_div_d (x, y, z, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_ui_pow_ui (mpfr_t x, const uint64_t y, const uint64_t z, const mpfr_rnd_t rnd)
{ mpfr_ui_pow_ui (x, (unsigned long) y, (unsigned long) z, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_log2 (mpfr_t x, const mpfr_t y, const mpfr_rnd_t rnd) { mpfr_log2 (x, y, rnd) ; return SLIP_OK ; }
SLIP_info SLIP_mpfr_free_cache (void) { mpfr_free_cache () ; return SLIP_OK ; }
