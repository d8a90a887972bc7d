/* slip_analyze.c -- SLIP_LU_analyze: column ordering Q and the nnz guesses.
 * Mirrors SLIP_LU/Source/SLIP_LU_analyze.c:24-134.
 *
 * COLAMD and AMD are SuiteSparse packages (separate libraries, libcolamd / libamd; the reference
 * vendors copies of them).  They are host-side preprocessing outside the hot path, so this
 * library does not re-implement them: it calls `colamd` / `amd_order` from the SuiteSparse
 * libraries found at run time: SLIP_B200_ORDERING_LIB if set, else the build of SuiteSparse that
 * slip_lu_b200/build.py places next to this library (_deps/libsuitesparse_ordering.so), else
 * libcolamd.so / libamd.so on the loader path.  SLIP_NO_ORDERING and a caller-supplied S->q need
 * nothing. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include "slip_internal.h"

typedef int (*colamd_fn) (int, int, int, int *, int *, double *, int *) ;
typedef int (*amd_order_fn) (int, const int *, const int *, int *, double *, double *) ;
typedef void (*amd_defaults_fn) (double *) ;

static void *open_ordering_lib (const char *fallback1, const char *fallback2)
{
    const char *env = getenv ("SLIP_B200_ORDERING_LIB") ;
    void *h = NULL ;
    if (env && *env) h = dlopen (env, RTLD_NOW | RTLD_LOCAL) ;
    if (!h)
    {   /* <directory of this shared library>/_deps/libsuitesparse_ordering.so */
        Dl_info me ;
        if (dladdr ((void *) &SLIP_LU_analyze, &me) && me.dli_fname)
        {
            const char *slash = strrchr (me.dli_fname, '/') ;
            size_t dir = slash ? (size_t) (slash - me.dli_fname) : 0 ;
            static const char tail [] = "/_deps/libsuitesparse_ordering.so" ;
            char *path = (char *) SLIP_malloc (dir + sizeof (tail) + 2) ;
            if (path)
            {
                if (dir) memcpy (path, me.dli_fname, dir) ; else { path [0] = '.' ; dir = 1 ; }
                memcpy (path + dir, tail, sizeof (tail)) ;
                h = dlopen (path, RTLD_NOW | RTLD_LOCAL) ;
                SLIP_free (path) ;
            }
        }
    }
    if (!h) h = dlopen (fallback1, RTLD_NOW | RTLD_LOCAL) ;
    if (!h && fallback2) h = dlopen (fallback2, RTLD_NOW | RTLD_LOCAL) ;
    return h ;
}

SLIP_info SLIP_LU_analyze (SLIP_LU_analysis *S, SLIP_sparse *A, SLIP_options *option)
{
    if (!S || !A || !A->i || !A->x || !A->p || !option || A->n != A->m || !S->q)
        return SLIP_INCORRECT_INPUT ;
    const int32_t n = A->n, nz = A->nz ;

    if (option->order == SLIP_NO_ORDERING)
    {
        for (int32_t k = 0 ; k <= n ; k++) S->q [k] = k ;
        S->lnz = S->unz = 10 * nz ;
    }
    else if (option->order == SLIP_AMD)
    {
        void *h = open_ordering_lib ("libamd.so", "libamd.so.3") ;
        amd_order_fn order = h ? (amd_order_fn) dlsym (h, "amd_order") : NULL ;
        amd_defaults_fn defaults = h ? (amd_defaults_fn) dlsym (h, "amd_defaults") : NULL ;
        if (!order || !defaults)
        {
            slip_set_error ("SLIP_AMD needs SuiteSparse AMD (set SLIP_B200_ORDERING_LIB)") ;
            fprintf (stderr, "slip_lu_b200: %s\n", SLIP_B200_last_error ()) ;
            return SLIP_INCORRECT_INPUT ;
        }
        double Control [5], Info [20] ;          /* AMD_CONTROL, AMD_INFO */
        defaults (Control) ;
        order (n, A->p, A->i, S->q, Control, Info) ;
        S->lnz = S->unz = (int32_t) Info [9] ;   /* AMD_LNZ */
    }
    else
    {
        void *h = open_ordering_lib ("libcolamd.so", "libcolamd.so.3") ;
        colamd_fn order = h ? (colamd_fn) dlsym (h, "colamd") : NULL ;
        if (!order)
        {
            slip_set_error ("SLIP_COLAMD needs SuiteSparse COLAMD (set SLIP_B200_ORDERING_LIB)") ;
            fprintf (stderr, "slip_lu_b200: %s\n", SLIP_B200_last_error ()) ;
            return SLIP_INCORRECT_INPUT ;
        }
        /* workspace size used by the reference call (SLIP_LU_analyze.c:88) */
        const int32_t Alen = 2 * nz + 6 * (n + 1) + 6 * (n + 1) + n ;
        int32_t *work = (int32_t *) SLIP_malloc ((size_t) Alen * sizeof (int32_t)) ;
        if (!work) return SLIP_OUT_OF_MEMORY ;
        memcpy (S->q, A->p, ((size_t) n + 1) * sizeof (int32_t)) ;
        memcpy (work, A->i, (size_t) nz * sizeof (int32_t)) ;
        int stats [20] ;                          /* COLAMD_STATS */
        order (n, n, Alen, work, S->q, NULL, stats) ;
        S->lnz = S->unz = 10 * nz ;
        SLIP_free (work) ;
    }
    /* clamp the guesses exactly as the reference does (SLIP_LU_analyze.c:117-132) */
    if (S->lnz > (double) n * n)
    {
        int32_t half = (int32_t) ceil (0.5 * n * n) ;
        S->lnz = S->unz = half ;
    }
    if (S->lnz < n) S->lnz += n ;
    if (S->unz < n) S->unz += n ;
    return SLIP_OK ;
}
