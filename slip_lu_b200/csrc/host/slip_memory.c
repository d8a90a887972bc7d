/* slip_memory.c -- allocation wrappers and the library environment.
 * Mirrors SLIP_LU/Source/SLIP_{malloc,calloc,realloc,free,initialize,initialize_expert,finalize}.c.
 * The reference installs setjmp-guarded GMP allocators to survive out-of-memory inside GMP
 * (SLIP_gmp.c); here GMP is only used for boundary conversions and keeps plain allocators
 * (or the caller's, via SLIP_initialize_expert). */
#include "slip_internal.h"

void *SLIP_malloc (size_t size) { return malloc (size ? size : 1) ; }

void *SLIP_calloc (size_t n, size_t size)
{
    if (n == 0) n = 1 ;
    if (size == 0) size = 1 ;
    return calloc (n, size) ;
}

void *SLIP_realloc (void *p, size_t old_size, size_t new_size)
{
    /* same contract as the reference: on failure the old block is released and NULL returned */
    (void) old_size ;
    void *q = realloc (p, new_size ? new_size : 1) ;
    if (!q) free (p) ;
    return q ;
}

void SLIP_free (void *p) { if (p) free (p) ; }

void SLIP_initialize (void) { SLIP_initialize_expert (NULL, NULL, NULL) ; }

void SLIP_initialize_expert (void *(*MyMalloc) (size_t), void *(*MyRealloc) (void *, size_t, size_t),
    void (*MyFree) (void *, size_t))
{
    mp_set_memory_functions (MyMalloc, MyRealloc, MyFree) ;   /* NULL keeps GMP's default */
}

void SLIP_finalize (void)
{
    slip_resident_drop_all () ;       /* GPU-resident factors of objects the caller never deleted */
    slipcu_release_cached_memory () ; /* cached device blocks and pinned buffers back to the driver */
    mpfr_free_cache () ;
}

static __thread char slip_err [512] ;

void slip_set_error (const char *msg)
{
    strncpy (slip_err, msg ? msg : "", sizeof (slip_err) - 1) ;
    slip_err [sizeof (slip_err) - 1] = 0 ;
}

const char *SLIP_B200_last_error (void) { return slip_err ; }
int SLIP_B200_device_count (void) { return slipcu_device_count () ; }
int SLIP_B200_set_device (int device) { return slipcu_set_device (device) ; }

SLIP_info slip_from_device_status (int rc)
{
    if (rc == SLIPCU_OK) return SLIP_OK ;
    slip_set_error (slipcu_last_error ()) ;
    switch (rc)
    {
        case SLIPCU_OUT_OF_MEMORY: return SLIP_OUT_OF_MEMORY ;
        case SLIPCU_SINGULAR:      return SLIP_SINGULAR ;
        case SLIPCU_BAD_INPUT:     return SLIP_INCORRECT_INPUT ;
        default:
            /* no device / launch failure: there is no CPU path to fall back to */
            fprintf (stderr, "slip_lu_b200: device layer failed: %s\n", slipcu_last_error ()) ;
            return SLIP_INCORRECT ;
    }
}

__thread slip_b200_stats slip_last_stats ;

/* out[0..] = n, nnz(L), nnz(U), channels, updates, limb_mul_equiv, t_symbolic, t_device, t_begin,
 * t_factor_total of the last factorization run by the calling thread; returns the count written */
int SLIP_B200_last_stats (double *out, int cap)
{
    const double v [13] = { slip_last_stats.n, slip_last_stats.nnz_L, slip_last_stats.nnz_U,
        slip_last_stats.channels, slip_last_stats.updates, slip_last_stats.limb_mul_equiv,
        slip_last_stats.t_symbolic, slip_last_stats.t_device, slip_last_stats.t_begin,
        slip_last_stats.t_factor_total,
        /* [10] channels the Hadamard bound asks for, [11] bound-mode restarts, [12] verified solves */
        slip_last_stats.channels_hadamard, slip_last_stats.bound_restarts, slip_last_stats.verified_solves } ;
    int k = 0 ;
    for ( ; k < 13 && k < cap ; k++) out [k] = v [k] ;
    return k ;
}
