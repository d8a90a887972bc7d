/* gmp_abi/mpfr.h -- declarations of the public MPFR 4.x ABI (x86-64 SysV).
 *
 * Same purpose as gmp_abi/gmp.h: the image has libmpfr.so.6 (MPFR 4.2.1) without its
 * header.  Only the documented entry points used by the SLIP_LU interface are declared.
 */
#ifndef __MPFR_H
#define __MPFR_H
#include <gmp.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef long mpfr_prec_t;
typedef int  mpfr_sign_t;
typedef long mpfr_exp_t;

typedef enum
{
    MPFR_RNDN = 0,  /* nearest, ties to even */
    MPFR_RNDZ,      /* toward zero */
    MPFR_RNDU,      /* toward +inf */
    MPFR_RNDD,      /* toward -inf */
    MPFR_RNDA,      /* away from zero */
    MPFR_RNDF,      /* faithful */
    MPFR_RNDNA = -1
} mpfr_rnd_t;

typedef struct
{
    mpfr_prec_t _mpfr_prec;
    mpfr_sign_t _mpfr_sign;
    mpfr_exp_t  _mpfr_exp;
    mp_limb_t  *_mpfr_d;
} __mpfr_struct;
typedef __mpfr_struct mpfr_t[1];
typedef __mpfr_struct *mpfr_ptr;
typedef const __mpfr_struct *mpfr_srcptr;

void mpfr_init2 (mpfr_ptr, mpfr_prec_t);
void mpfr_clear (mpfr_ptr);
int  mpfr_set4 (mpfr_ptr, mpfr_srcptr, mpfr_rnd_t, int);
/* in the real header mpfr_set is a macro over mpfr_set4; the function also exists */
int  mpfr_set (mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int  mpfr_set_d (mpfr_ptr, double, mpfr_rnd_t);
int  mpfr_set_q (mpfr_ptr, mpq_srcptr, mpfr_rnd_t);
int  mpfr_set_z (mpfr_ptr, mpz_srcptr, mpfr_rnd_t);
int  mpfr_set_si (mpfr_ptr, long, mpfr_rnd_t);
int  mpfr_set_str (mpfr_ptr, const char *, int, mpfr_rnd_t);
int  mpfr_abs (mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int  mpfr_get_z (mpz_ptr, mpfr_srcptr, mpfr_rnd_t);
double mpfr_get_d (mpfr_srcptr, mpfr_rnd_t);
int  mpfr_mul (mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int  mpfr_mul_d (mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int  mpfr_div_d (mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int  mpfr_ui_pow_ui (mpfr_ptr, unsigned long, unsigned long, mpfr_rnd_t);
int  mpfr_log2 (mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int  mpfr_cmp3 (mpfr_srcptr, mpfr_srcptr, int);
#define mpfr_cmp(a,b) mpfr_cmp3 (a, b, 1)
void mpfr_free_cache (void);
void mpfr_free_str (char *);
char *mpfr_get_str (char *, mpfr_exp_t *, int, size_t, mpfr_srcptr, mpfr_rnd_t);

#define mpfr_vfprintf  __gmpfr_vfprintf
#define mpfr_vasprintf __gmpfr_vasprintf
int __gmpfr_vfprintf (FILE *, const char *, va_list);
int __gmpfr_vasprintf (char **, const char *, va_list);

#ifdef __cplusplus
}
#endif
#endif /* __MPFR_H */
