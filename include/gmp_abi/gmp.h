/* gmp_abi/gmp.h -- declarations of the public GMP 6.x ABI (x86-64 SysV, 64-bit limbs)
 *
 * The build image carries libgmp.so.10 (GMP 6.3.0) but not its development header.
 * This file declares the subset of the documented GMP interface that slip_lu_b200,
 * its oracle and the oracle/_ref build of the reference need, so that they can link
 * against the system libgmp.  It is only put on the include path when <gmp.h> is
 * absent (see slip_lu_b200/build.py and oracle/Makefile); with a real gmp.h installed
 * it is never seen.  Struct layouts and symbol names (__gmpz_*, __gmpq_*) are the
 * stable GMP ABI documented in the GMP manual ("Integer Internals", "Rational
 * Internals", "Custom Allocation").
 */
#ifndef __GMP_H__
#define __GMP_H__
#define SLIP_B200_GMP_ABI_SHIM 1

#include <stddef.h>
#include <stdio.h>
#include <stdarg.h>

#ifdef __cplusplus
extern "C" {
#endif

#define __GNU_MP_VERSION 6
#define GMP_LIMB_BITS 64
#define GMP_NUMB_BITS 64
#define GMP_NAIL_BITS 0

typedef unsigned long int mp_limb_t;
typedef long int mp_limb_signed_t;
typedef unsigned long int mp_bitcnt_t;
typedef long int mp_size_t;
typedef long int mp_exp_t;

typedef struct
{
    int _mp_alloc;      /* limbs allocated at _mp_d */
    int _mp_size;       /* |size| = limbs in use, sign = sign of the number */
    mp_limb_t *_mp_d;   /* least significant limb first */
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

typedef struct
{
    __mpz_struct _mp_num;
    __mpz_struct _mp_den;
} __mpq_struct;
typedef __mpq_struct mpq_t[1];
typedef __mpq_struct *mpq_ptr;
typedef const __mpq_struct *mpq_srcptr;

typedef struct
{
    int _mp_prec;
    int _mp_size;
    mp_exp_t _mp_exp;
    mp_limb_t *_mp_d;
} __mpf_struct;
typedef __mpf_struct mpf_t[1];

#define mpq_numref(Q) (&((Q)->_mp_num))
#define mpq_denref(Q) (&((Q)->_mp_den))
#define mpz_sgn(Z) ((Z)->_mp_size < 0 ? -1 : (Z)->_mp_size > 0)
#define mpq_sgn(Q) ((Q)->_mp_num._mp_size < 0 ? -1 : (Q)->_mp_num._mp_size > 0)

extern const char *const __gmp_version;
#define gmp_version __gmp_version

/* ---- custom allocation ---- */
#define mp_set_memory_functions __gmp_set_memory_functions
#define mp_get_memory_functions __gmp_get_memory_functions
void __gmp_set_memory_functions (void *(*) (size_t),
    void *(*) (void *, size_t, size_t), void (*) (void *, size_t));
void __gmp_get_memory_functions (void *(**) (size_t),
    void *(**) (void *, size_t, size_t), void (**) (void *, size_t));

/* ---- formatted I/O ---- */
#define gmp_printf    __gmp_printf
#define gmp_fprintf   __gmp_fprintf
#define gmp_snprintf  __gmp_snprintf
#define gmp_vprintf   __gmp_vprintf
#define gmp_vfprintf  __gmp_vfprintf
#define gmp_vfscanf   __gmp_vfscanf
#define gmp_sscanf    __gmp_sscanf
int __gmp_printf (const char *, ...);
int __gmp_fprintf (FILE *, const char *, ...);
int __gmp_snprintf (char *, size_t, const char *, ...);
int __gmp_vprintf (const char *, va_list);
int __gmp_vfprintf (FILE *, const char *, va_list);
int __gmp_vfscanf (FILE *, const char *, va_list);
int __gmp_sscanf (const char *, const char *, ...);

/* ---- mpz ---- */
#define mpz_init        __gmpz_init
#define mpz_init2       __gmpz_init2
#define mpz_init_set    __gmpz_init_set
#define mpz_init_set_ui __gmpz_init_set_ui
#define mpz_init_set_si __gmpz_init_set_si
#define mpz_init_set_str __gmpz_init_set_str
#define mpz_clear       __gmpz_clear
#define mpz_realloc2    __gmpz_realloc2
#define mpz_set         __gmpz_set
#define mpz_set_ui      __gmpz_set_ui
#define mpz_set_si      __gmpz_set_si
#define mpz_set_d       __gmpz_set_d
#define mpz_set_q       __gmpz_set_q
#define mpz_set_str     __gmpz_set_str
#define mpz_get_str     __gmpz_get_str
#define mpz_get_d       __gmpz_get_d
#define mpz_get_d_2exp  __gmpz_get_d_2exp
#define mpz_get_ui      __gmpz_get_ui
#define mpz_get_si      __gmpz_get_si
#define mpz_swap        __gmpz_swap
#define mpz_add         __gmpz_add
#define mpz_add_ui      __gmpz_add_ui
#define mpz_sub         __gmpz_sub
#define mpz_sub_ui      __gmpz_sub_ui
#define mpz_mul         __gmpz_mul
#define mpz_mul_ui      __gmpz_mul_ui
#define mpz_mul_si      __gmpz_mul_si
#define mpz_mul_2exp    __gmpz_mul_2exp
#define mpz_addmul      __gmpz_addmul
#define mpz_addmul_ui   __gmpz_addmul_ui
#define mpz_submul      __gmpz_submul
#define mpz_neg         __gmpz_neg
#define mpz_abs         __gmpz_abs
#define mpz_divexact    __gmpz_divexact
#define mpz_divexact_ui __gmpz_divexact_ui
#define mpz_tdiv_q      __gmpz_tdiv_q
#define mpz_tdiv_r      __gmpz_tdiv_r
#define mpz_divisible_ui_p __gmpz_divisible_ui_p
#define mpz_gcd_ui      __gmpz_gcd_ui
#define mpz_fits_ulong_p __gmpz_fits_ulong_p
#define mpz_tdiv_q_2exp __gmpz_tdiv_q_2exp
#define mpz_tdiv_r_2exp __gmpz_tdiv_r_2exp
#define mpz_fdiv_ui     __gmpz_fdiv_ui
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_mod         __gmpz_mod
#define mpz_gcd         __gmpz_gcd
#define mpz_lcm         __gmpz_lcm
#define mpz_pow_ui      __gmpz_pow_ui
#define mpz_ui_pow_ui   __gmpz_ui_pow_ui
#define mpz_cmp         __gmpz_cmp
#define mpz_cmp_ui      __gmpz_cmp_ui
#define mpz_cmp_si      __gmpz_cmp_si
#define mpz_cmpabs      __gmpz_cmpabs
#define mpz_cmpabs_ui   __gmpz_cmpabs_ui
#define mpz_sizeinbase  __gmpz_sizeinbase
#define mpz_size        __gmpz_size
#define mpz_getlimbn    __gmpz_getlimbn
#define mpz_tstbit      __gmpz_tstbit
#define mpz_import      __gmpz_import
#define mpz_export      __gmpz_export
#define mpz_limbs_read  __gmpz_limbs_read
#define mpz_limbs_write __gmpz_limbs_write
#define mpz_limbs_finish __gmpz_limbs_finish
#define mpz_fits_slong_p __gmpz_fits_slong_p

void __gmpz_init (mpz_ptr);
void __gmpz_init2 (mpz_ptr, mp_bitcnt_t);
void __gmpz_init_set (mpz_ptr, mpz_srcptr);
void __gmpz_init_set_ui (mpz_ptr, unsigned long int);
void __gmpz_init_set_si (mpz_ptr, signed long int);
int  __gmpz_init_set_str (mpz_ptr, const char *, int);
void __gmpz_clear (mpz_ptr);
void __gmpz_realloc2 (mpz_ptr, mp_bitcnt_t);
void __gmpz_set (mpz_ptr, mpz_srcptr);
void __gmpz_set_ui (mpz_ptr, unsigned long int);
void __gmpz_set_si (mpz_ptr, signed long int);
void __gmpz_set_d (mpz_ptr, double);
void __gmpz_set_q (mpz_ptr, mpq_srcptr);
int  __gmpz_set_str (mpz_ptr, const char *, int);
char *__gmpz_get_str (char *, int, mpz_srcptr);
double __gmpz_get_d (mpz_srcptr);
double __gmpz_get_d_2exp (signed long int *, mpz_srcptr);
unsigned long int __gmpz_get_ui (mpz_srcptr);
signed long int __gmpz_get_si (mpz_srcptr);
void __gmpz_swap (mpz_ptr, mpz_ptr);
void __gmpz_add (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_add_ui (mpz_ptr, mpz_srcptr, unsigned long int);
void __gmpz_sub (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_sub_ui (mpz_ptr, mpz_srcptr, unsigned long int);
void __gmpz_mul (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_mul_ui (mpz_ptr, mpz_srcptr, unsigned long int);
void __gmpz_mul_si (mpz_ptr, mpz_srcptr, long int);
void __gmpz_mul_2exp (mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void __gmpz_addmul (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_addmul_ui (mpz_ptr, mpz_srcptr, unsigned long int);
void __gmpz_submul (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_neg (mpz_ptr, mpz_srcptr);
void __gmpz_abs (mpz_ptr, mpz_srcptr);
void __gmpz_divexact (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_divexact_ui (mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_tdiv_q (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_tdiv_r (mpz_ptr, mpz_srcptr, mpz_srcptr);
int  __gmpz_divisible_ui_p (mpz_srcptr, unsigned long int);
unsigned long int __gmpz_gcd_ui (mpz_ptr, mpz_srcptr, unsigned long int);
int  __gmpz_fits_ulong_p (mpz_srcptr);
void __gmpz_tdiv_q_2exp (mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void __gmpz_tdiv_r_2exp (mpz_ptr, mpz_srcptr, mp_bitcnt_t);
unsigned long int __gmpz_fdiv_ui (mpz_srcptr, unsigned long int);
void __gmpz_fdiv_q_2exp (mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void __gmpz_mod (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_gcd (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_lcm (mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_pow_ui (mpz_ptr, mpz_srcptr, unsigned long int);
void __gmpz_ui_pow_ui (mpz_ptr, unsigned long int, unsigned long int);
int  __gmpz_cmp (mpz_srcptr, mpz_srcptr);
int  __gmpz_cmp_ui (mpz_srcptr, unsigned long int);
int  __gmpz_cmp_si (mpz_srcptr, signed long int);
int  __gmpz_cmpabs (mpz_srcptr, mpz_srcptr);
int  __gmpz_cmpabs_ui (mpz_srcptr, unsigned long int);
size_t __gmpz_sizeinbase (mpz_srcptr, int);
size_t __gmpz_size (mpz_srcptr);
mp_limb_t __gmpz_getlimbn (mpz_srcptr, mp_size_t);
int  __gmpz_tstbit (mpz_srcptr, mp_bitcnt_t);
void __gmpz_import (mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void *__gmpz_export (void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
const mp_limb_t *__gmpz_limbs_read (mpz_srcptr);
mp_limb_t *__gmpz_limbs_write (mpz_ptr, mp_size_t);
void __gmpz_limbs_finish (mpz_ptr, mp_size_t);
int  __gmpz_fits_slong_p (mpz_srcptr);

/* ---- mpq ---- */
#define mpq_init         __gmpq_init
#define mpq_clear        __gmpq_clear
#define mpq_canonicalize __gmpq_canonicalize
#define mpq_set          __gmpq_set
#define mpq_set_z        __gmpq_set_z
#define mpq_set_d        __gmpq_set_d
#define mpq_set_ui       __gmpq_set_ui
#define mpq_set_si       __gmpq_set_si
#define mpq_set_str      __gmpq_set_str
#define mpq_get_str      __gmpq_get_str
#define mpq_set_num      __gmpq_set_num
#define mpq_set_den      __gmpq_set_den
#define mpq_get_num      __gmpq_get_num
#define mpq_get_den      __gmpq_get_den
#define mpq_get_d        __gmpq_get_d
#define mpq_abs          __gmpq_abs
#define mpq_neg          __gmpq_neg
#define mpq_add          __gmpq_add
#define mpq_sub          __gmpq_sub
#define mpq_mul          __gmpq_mul
#define mpq_div          __gmpq_div
#define mpq_cmp          __gmpq_cmp
#define mpq_cmp_ui       __gmpq_cmp_ui
#define mpq_equal        __gmpq_equal
#define mpq_swap         __gmpq_swap

void __gmpq_init (mpq_ptr);
void __gmpq_clear (mpq_ptr);
void __gmpq_canonicalize (mpq_ptr);
void __gmpq_set (mpq_ptr, mpq_srcptr);
void __gmpq_set_z (mpq_ptr, mpz_srcptr);
void __gmpq_set_d (mpq_ptr, double);
void __gmpq_set_ui (mpq_ptr, unsigned long int, unsigned long int);
void __gmpq_set_si (mpq_ptr, signed long int, unsigned long int);
int  __gmpq_set_str (mpq_ptr, const char *, int);
char *__gmpq_get_str (char *, int, mpq_srcptr);
void __gmpq_set_num (mpq_ptr, mpz_srcptr);
void __gmpq_set_den (mpq_ptr, mpz_srcptr);
void __gmpq_get_num (mpz_ptr, mpq_srcptr);
void __gmpq_get_den (mpz_ptr, mpq_srcptr);
double __gmpq_get_d (mpq_srcptr);
void __gmpq_abs (mpq_ptr, mpq_srcptr);
void __gmpq_neg (mpq_ptr, mpq_srcptr);
void __gmpq_add (mpq_ptr, mpq_srcptr, mpq_srcptr);
void __gmpq_sub (mpq_ptr, mpq_srcptr, mpq_srcptr);
void __gmpq_mul (mpq_ptr, mpq_srcptr, mpq_srcptr);
void __gmpq_div (mpq_ptr, mpq_srcptr, mpq_srcptr);
int  __gmpq_cmp (mpq_srcptr, mpq_srcptr);
int  __gmpq_cmp_ui (mpq_srcptr, unsigned long int, unsigned long int);
int  __gmpq_equal (mpq_srcptr, mpq_srcptr);
void __gmpq_swap (mpq_ptr, mpq_ptr);

#ifdef __cplusplus
}
#endif
#endif /* __GMP_H__ */
