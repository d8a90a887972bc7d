/* slip_factorize.c -- SLIP_LU_factorize: left-looking REF LU, P A Q = L D^-1 U.
 *
 * Mirrors SLIP_LU/Source/SLIP_LU_factorize.c:34-323.  Division of labour:
 *   host (this file)   the symbolic side of each column: reach of A(:,q[k]) in the graph of L
 *                      (slip_reach.c / slip_dfs.c), ordering of the pattern by the current row
 *                      permutation (slip_sort_xi.c), the row-permutation bookkeeping of
 *                      slip_get_pivot.c:152-172 and the final assembly of L, U, rhos as mpz_t.
 *   GPU (slipcu_*)     all arithmetic: the sparse REF triangular solve, exact reconstruction and
 *                      the exact pivot scan.
 * The only arithmetic left on the host is the rational tolerance comparison of
 * SLIP_TOL_SMALLEST / SLIP_TOL_LARGEST on two already-reconstructed integers, done with the very
 * GMP call the reference uses (slip_get_pivot.c:94-143) so that its outcome is reproduced. */
#include <time.h>
#include "slip_internal.h"

/* Hadamard channel counts up to this are carried in full from the start (see slip_factorize_driver) */
#define SLIP_B200_FULL_START_MAX 768

static double now_s (void)
{
    struct timespec t ;
    clock_gettime (CLOCK_MONOTONIC, &t) ;
    return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec ;
}

/* ---- host copy of the patterns of the finished columns ---- */
typedef struct
{
    int32_t *rows ;      /* all patterns back to back (original row indices) */
    int64_t cap, used ;
    int64_t *ptr ;       /* n+1 */
    int32_t *nU ;        /* n: size of the U part of each pattern */
    int32_t *piv ;       /* n: slot of the pivot */
    /* depth-first search view of the L parts (same offsets as rows): a reorderable copy whose
     * first pend[j] rows are the only ones the search has to follow (symmetric pruning) */
    int32_t *prows ;
    int32_t *pend ;      /* n */
    int8_t *pruned ;     /* n */
    /* row-wise view of the L parts: for every row r the finished columns j whose L part holds r,
     * as linked nodes (node m: column rl_col[m], next node rl_next[m]); rl_head[r] = first node.
     * Lets the pruning find the columns that hold the new pivot row without searching them. */
    int32_t *rl_head ;   /* n */
    int32_t *rl_col, *rl_next ;
    int64_t rl_cap, rl_used ;
    int32_t *ustamp ;    /* n: ustamp[j] == k+1 <=> column j is in the U part of column k */
} pattern_store ;

static void patterns_free (pattern_store *P)
{
    SLIP_free (P->rows) ; SLIP_free (P->ptr) ; SLIP_free (P->nU) ; SLIP_free (P->piv) ;
    SLIP_free (P->prows) ; SLIP_free (P->pend) ; SLIP_free (P->pruned) ;
    SLIP_free (P->rl_head) ; SLIP_free (P->rl_col) ; SLIP_free (P->rl_next) ; SLIP_free (P->ustamp) ;
    memset (P, 0, sizeof (*P)) ;
}

static SLIP_info patterns_init (pattern_store *P, int32_t n, int64_t guess)
{
    memset (P, 0, sizeof (*P)) ;
    P->cap = guess > 4 * (int64_t) n ? guess : 4 * (int64_t) n ;
    P->rows = (int32_t *) SLIP_malloc ((size_t) P->cap * sizeof (int32_t)) ;
    P->ptr = (int64_t *) SLIP_calloc ((size_t) n + 1, sizeof (int64_t)) ;
    P->nU = (int32_t *) SLIP_calloc ((size_t) n, sizeof (int32_t)) ;
    P->piv = (int32_t *) SLIP_calloc ((size_t) n, sizeof (int32_t)) ;
    P->prows = (int32_t *) SLIP_malloc ((size_t) P->cap * sizeof (int32_t)) ;
    P->pend = (int32_t *) SLIP_calloc ((size_t) n, sizeof (int32_t)) ;
    P->pruned = (int8_t *) SLIP_calloc ((size_t) n, sizeof (int8_t)) ;
    P->rl_cap = P->cap ; P->rl_used = 0 ;
    P->rl_head = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    P->rl_col = (int32_t *) SLIP_malloc ((size_t) P->rl_cap * sizeof (int32_t)) ;
    P->rl_next = (int32_t *) SLIP_malloc ((size_t) P->rl_cap * sizeof (int32_t)) ;
    P->ustamp = (int32_t *) SLIP_calloc ((size_t) n, sizeof (int32_t)) ;
    if (!P->rows || !P->ptr || !P->nU || !P->piv || !P->prows || !P->pend || !P->pruned
        || !P->rl_head || !P->rl_col || !P->rl_next || !P->ustamp)
    { patterns_free (P) ; return SLIP_OUT_OF_MEMORY ; }
    for (int32_t r = 0 ; r < n ; r++) P->rl_head [r] = -1 ;
    return SLIP_OK ;
}

static SLIP_info patterns_reserve (pattern_store *P, int64_t extra)
{
    if (P->used + extra <= P->cap) return SLIP_OK ;
    int64_t ncap = P->cap ;
    while (ncap < P->used + extra) ncap *= 2 ;
    int32_t *nr = (int32_t *) realloc (P->rows, (size_t) ncap * sizeof (int32_t)) ;
    if (!nr) return SLIP_OUT_OF_MEMORY ;
    P->rows = nr ;
    int32_t *np = (int32_t *) realloc (P->prows, (size_t) ncap * sizeof (int32_t)) ;
    if (!np) return SLIP_OUT_OF_MEMORY ;
    P->prows = np ; P->cap = ncap ;
    return SLIP_OK ;
}

/* Rows reachable from the rows of A(:,col) through the finished columns 0..klim-1 of L, unordered.
 * With klim = k this is the pattern of column k (slip_reach.c / slip_dfs.c).  With klim = k-1 it is
 * the part of that pattern that does not depend on the pivot of column k-1, which is what the host
 * computes while the GPU is still working on column k-1.  mark[r] == stamp flags membership. */
static int32_t reach_unordered (const SLIP_sparse *A, int32_t col, int32_t klim, const pattern_store *P,
    const int32_t *pinv, int32_t *mark, int32_t stamp, int32_t *stack, int32_t *out)
{
    int32_t cnt = 0 ;
    for (int32_t a = A->p [col] ; a < A->p [col + 1] ; a++)
    {
        int32_t r0 = A->i [a] ;
        if (mark [r0] == stamp) continue ;
        int32_t sp = 0 ;
        mark [r0] = stamp ; stack [sp++] = r0 ;
        while (sp > 0)
        {
            const int32_t r = stack [--sp] ;
            const int32_t pos = pinv [r] ;
            out [cnt++] = r ;
            if (pos < klim)
            {   /* row r is the pivot of column pos: follow the (pruned) L part of that column */
                const int32_t *rows = P->prows + P->ptr [pos] + P->nU [pos] ;
                const int32_t len = P->pend [pos] ;
                for (int32_t m = 0 ; m < len ; m++)
                {
                    const int32_t rr = rows [m] ;
                    if (mark [rr] != stamp) { mark [rr] = stamp ; stack [sp++] = rr ; }
                }
            }
        }
    }
    return cnt ;
}

/* Symmetric pruning (Eisenstat & Liu; the pruning of left-looking LU codes such as KLU), after
 * row prow became the pivot of column k: for every column j of the U part of column k whose L part
 * contains prow, the rows of L(:,j) that are not pivotal yet are also rows of L(:,k) (column k
 * reached j, so its pattern holds all of them), hence any later search that arrives at j finds
 * them through prow -> column k.  Only the pivotal rows of L(:,j) stay visible: in the dense
 * trailing part a search then costs O(columns) instead of O(entries).  The reach SET is unchanged,
 * and the order of a pattern is fixed afterwards by order_by_position, so nothing else is affected. */
static SLIP_info prune_columns (pattern_store *P, int32_t k, int32_t prow, int32_t nU, const int32_t *upos,
    const int32_t *pinv)
{
    /* the columns of the U part of column k ... */
    for (int32_t u = 0 ; u < nU ; u++) P->ustamp [upos [u]] = k + 1 ;
    /* ... whose L part holds the new pivot row: read off the row's list (total cost over the
       factorization: nnz(L); searching every L(:,j) for prow cost as much as the numerical work) */
    for (int32_t m = P->rl_head [prow] ; m >= 0 ; m = P->rl_next [m])
    {
        const int32_t j = P->rl_col [m] ;
        if (P->pruned [j] || P->ustamp [j] != k + 1) continue ;
        int32_t *rj = P->prows + P->ptr [j] + P->nU [j] ;
        int32_t head = 0, tail = P->pend [j] ;
        while (head < tail)
        {
            if (pinv [rj [head]] <= k) head++ ;
            else { tail-- ; const int32_t t = rj [head] ; rj [head] = rj [tail] ; rj [tail] = t ; }
        }
        P->pend [j] = tail ;
        P->pruned [j] = 1 ;
    }
    /* column k joins the row lists: every row of its L part except the pivot row itself */
    const int32_t cnt = (int32_t) (P->ptr [k + 1] - P->ptr [k]), nUk = P->nU [k] ;
    if (P->rl_used + cnt > P->rl_cap)
    {
        int64_t ncap = P->rl_cap ;
        while (ncap < P->rl_used + cnt) ncap *= 2 ;
        int32_t *nc = (int32_t *) realloc (P->rl_col, (size_t) ncap * sizeof (int32_t)) ;
        if (!nc) return SLIP_OUT_OF_MEMORY ;
        P->rl_col = nc ;
        int32_t *nn = (int32_t *) realloc (P->rl_next, (size_t) ncap * sizeof (int32_t)) ;
        if (!nn) return SLIP_OUT_OF_MEMORY ;
        P->rl_next = nn ; P->rl_cap = ncap ;
    }
    const int32_t *rk = P->rows + P->ptr [k] ;
    for (int32_t t = nUk ; t < cnt ; t++)
    {
        const int32_t r = rk [t] ;
        if (r == prow) continue ;
        const int32_t m = (int32_t) P->rl_used++ ;
        P->rl_col [m] = k ; P->rl_next [m] = P->rl_head [r] ; P->rl_head [r] = m ;
    }
    return SLIP_OK ;
}

/* order a pattern by current row position (slip_sort_xi.c): positions are a permutation, so a bit
 * per position and one sweep over the words in use replace the sort (posbits: n/64+1 zero words,
 * left zero again) */
static void order_by_position (int32_t n, int32_t cnt, int32_t *pat, const int32_t *pinv,
    const int32_t *row_at, uint64_t *posbits)
{
    (void) n ;
    int32_t lo = INT32_MAX, hi = -1 ;
    for (int32_t t = 0 ; t < cnt ; t++)
    {
        const int32_t pos = pinv [pat [t]] ;
        posbits [pos >> 6] |= (uint64_t) 1 << (pos & 63) ;
        if (pos < lo) lo = pos ;
        if (pos > hi) hi = pos ;
    }
    int32_t w = 0 ;
    for (int32_t q = lo >> 6 ; hi >= 0 && q <= (hi >> 6) ; q++)
    {
        uint64_t word = posbits [q] ;
        posbits [q] = 0 ;
        while (word)
        {
            const int bit = __builtin_ctzll (word) ;
            pat [w++] = row_at [(q << 6) + bit] ;
            word &= word - 1 ;
        }
    }
}

/* ---- the rational tolerance test on two reconstructed entries ---- */
static SLIP_info tolerance_prefers_diagonal (slipcu_factor *dev, int32_t k, int scheme, double tol,
    int32_t best_slot, int32_t diag_slot, int *prefer)
{
    SLIP_info status = SLIP_OK ;
    const int stride = slipcu_factor_column_stride (dev, k) ;
    uint32_t *wb = (uint32_t *) SLIP_malloc ((size_t) (stride + 2) * sizeof (uint32_t)) ;
    uint32_t *wd = (uint32_t *) SLIP_malloc ((size_t) (stride + 2) * sizeof (uint32_t)) ;
    mpz_t vb, vd ; mpq_t ratio, t ;
    mpz_init (vb) ; mpz_init (vd) ; mpq_init (ratio) ; mpq_init (t) ;
    if (!wb || !wd) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
    {
        int32_t nb = 0, nd = 0 ; int8_t sb = 0, sd = 0 ;
        SLIP_TRY (slip_from_device_status (slipcu_factor_fetch_entry (dev, k, best_slot, wb, &nb, &sb))) ;
        SLIP_TRY (slip_from_device_status (slipcu_factor_fetch_entry (dev, k, diag_slot, wd, &nd, &sd))) ;
        slip_mpz_from_words (vb, wb, nb, sb) ;
        slip_mpz_from_words (vd, wd, nd, sd) ;
    }
    if (scheme == SLIP_TOL_SMALLEST)
    {   /* |smallest| / |diagonal| */
        mpz_abs (mpq_numref (ratio), vb) ;
        mpz_abs (mpq_denref (ratio), vd) ;
    }
    else
    {   /* the reference takes mpq_abs of diagonal/largest without canonicalising: the operand
           of the comparison is |diagonal| over the SIGNED largest entry (slip_get_pivot.c:131-137) */
        mpz_abs (mpq_numref (ratio), vd) ;
        mpz_set (mpq_denref (ratio), vb) ;
    }
    mpq_set_d (t, tol) ;
    *prefer = (mpq_cmp (ratio, t) >= 0) ;
cleanup:
    mpz_clear (vb) ; mpz_clear (vd) ; mpq_clear (ratio) ; mpq_clear (t) ;
    SLIP_free (wb) ; SLIP_free (wd) ;
    return status ;
}

static SLIP_info decide_pivot (slipcu_factor *dev, int32_t k, int scheme, double tol, int32_t diag_slot,
    const slipcu_pivot_info *info, int32_t *slot)
{
    const int32_t best = info->best_slot ;
    *slot = best ;
    if (best < 0) return SLIP_SINGULAR ;          /* every candidate is zero */
    if (!info->diag_eligible || diag_slot == best) return SLIP_OK ;
    switch (scheme)
    {
        case SLIP_DIAGONAL:
            *slot = diag_slot ;
            return SLIP_OK ;
        case SLIP_TOL_SMALLEST:
            /* ratio = |smallest|/|diagonal| lies in (0,1]; most tolerances need no arithmetic */
            if (tol != tol) return SLIP_OK ;
            if (tol <= 0.0 || info->diag_vs_best == 0) { if (tol <= 1.0) *slot = diag_slot ; return SLIP_OK ; }
            if (tol >= 1.0) return SLIP_OK ;
            /* fall through: 0 < tol < 1 and |diagonal| > |smallest| */
        case SLIP_TOL_LARGEST:
        {
            if (tol != tol) return SLIP_OK ;
            int prefer = 0 ;
            SLIP_info status = tolerance_prefers_diagonal (dev, k, scheme, tol, best, diag_slot, &prefer) ;
            if (status != SLIP_OK) return status ;
            if (prefer) *slot = diag_slot ;
            return SLIP_OK ;
        }
        default:
            return SLIP_OK ;
    }
}

/* ---- assembling L, U, rhos on the host from the streamed columns ---- */
typedef struct
{
    SLIP_sparse *L, *U ;
    mpz_t *rhos ;
    const pattern_store *P ;
} assemble_ctx ;

static int assemble_column (void *user, int k, int cnt, int stride, const uint32_t *limbs,
    const int32_t *nl, const int8_t *sign)
{
    assemble_ctx *c = (assemble_ctx *) user ;
    const int32_t nU = c->P->nU [k], piv = c->P->piv [k] ;
    mpz_t *Ux = c->U->x + c->U->p [k] ;
    mpz_t *Lx = c->L->x + c->L->p [k] ;
    #pragma omp parallel for schedule(dynamic, 16) if (cnt > 32)
    for (int t = 0 ; t < cnt ; t++)
    {
        mpz_ptr dst = (t < nU) ? Ux [t] : Lx [t - nU] ;
        mpz_init (dst) ;
        slip_mpz_from_words (dst, limbs + (size_t) t * stride, nl [t], sign [t]) ;
    }
    mpz_init_set (Ux [nU], Lx [piv - nU]) ;       /* the pivot closes the column of U */
    mpz_set (c->rhos [k], Lx [piv - nU]) ;
    return 0 ;
}

/* one column whose pattern is being prepared ahead of the column in flight */
typedef struct
{
    int32_t *pat ;       /* rows reached through the columns < klim (unordered), later the full ordered pattern */
    int32_t *mark ;      /* n: mark [r] == stamp <=> r is in pat */
    int32_t cnt, klim, stamp, spec_slot ;
} ahead_col ;

SLIP_info slip_factorize_driver (SLIP_sparse *L, SLIP_sparse *U, SLIP_sparse *A, SLIP_LU_analysis *S,
    mpz_t *rhos, int32_t *pinv, SLIP_options *option, int want_host_factors, slip_resident **resident,
    double rhs_bits, int min_channels)
{
    if (!A || !S || !pinv || !option || !A->p || !A->x || !A->i || !S->q || A->n <= 0 || A->n != A->m)
        return SLIP_INCORRECT_INPUT ;
    if (want_host_factors && (!L || !U || !rhos)) return SLIP_INCORRECT_INPUT ;
    const int32_t n = A->n, nz = A->p [n] ;
    if (nz <= 0) return SLIP_INCORRECT_INPUT ;
    const int scheme = (int) option->pivot ;
    SLIP_info status = SLIP_OK ;
    const int timing = getenv ("SLIP_B200_TIMING") != NULL ;
    const char *prune_env = getenv ("SLIP_B200_PRUNE") ;
    const int use_pruning = !(prune_env && prune_env [0] == '0') ;      /* symmetric pruning of the reach (default on) */
    const char *single_env = getenv ("SLIP_B200_SINGLE") ;
    const int use_single = !(single_env && single_env [0] == '0') ;      /* no round trip for single-candidate columns */
    double t_sym = 0, t_dev = 0, t_piv = 0, t_begin = 0, t0 = now_s (), tt ;
    double work_updates = 0, work_limbmul = 0 ;
    double *cumbits_at = (double *) SLIP_calloc ((size_t) n, sizeof (double)) ;
    slip_limbs Al = {0} ;
    pattern_store P = {0} ;
    slipcu_factor *dev = NULL ;
    slip_resident *res = NULL ;
    ahead_col ring [SLIPCU_SPEC_SLOTS] ;
    memset (ring, 0, sizeof (ring)) ;
    int ring_size = 0, restarts = 0, bound_restarts = 0 ;
    double *colbits = (double *) SLIP_malloc ((size_t) n * sizeof (double)) ;
    int32_t *row_at = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    int32_t *stack = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    int32_t *upos = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    int32_t *spat = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;      /* bulk-part pattern, ordered */
    int32_t *supos = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
    uint64_t *posflag = (uint64_t *) SLIP_calloc ((size_t) n / 64 + 2, sizeof (uint64_t)) ;      /* position bitmap of order_by_position */
    if (!colbits || !row_at || !stack || !upos || !cumbits_at || !posflag || !spat || !supos) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }

    for (int32_t a = 0 ; a < nz ; a++)
        if (A->i [a] < 0 || A->i [a] >= n) { status = SLIP_INCORRECT_INPUT ; goto cleanup ; }
    for (int32_t k = 0 ; k < n ; k++)
        if (S->q [k] < 0 || S->q [k] >= n) { status = SLIP_INCORRECT_INPUT ; goto cleanup ; }

    /* sizing (replaces the allocation bound of SLIP_LU_factorize.c:102-163 by a true bound) */
    SLIP_TRY (slip_column_bits (A, colbits)) ;
    double total_bits = 0, min_bits = 1e300 ;
    for (int32_t j = 0 ; j < n ; j++) { total_bits += colbits [j] ; if (colbits [j] < min_bits) min_bits = colbits [j] ; }
    /* a right-hand side known up front (SLIP_solve_*) may need more room than the factors */
    const double extra = rhs_bits > min_bits ? rhs_bits - min_bits : 0.0 ;
    const int channels_full = slip_channels_for_bits (total_bits + extra) + SLIP_B200_SPARE_CHANNELS ;
    /* The Hadamard bound is what a residue system must carry to be safe a priori, and on real LP
       bases it is one to two orders of magnitude above the sizes that occur (NSR8K: 29 361 bits
       against a 704-bit determinant).  GMP pays for actual operand sizes; to do the same the
       factorization starts with a fraction of the channels in BOUND MODE -- every column's size is
       then proven on the device from the measured sizes of the finished columns, see
       slip_b200_device.h -- and restarts with four times as many when a column does not fit. */
    int channels = channels_full ;
    {
        const char *ad = getenv ("SLIP_B200_ADAPTIVE"), *sc = getenv ("SLIP_B200_START_CHANNELS") ;
        if (!(ad && ad [0] == '0'))
        {
            int start = sc && *sc ? atoi (sc) : channels_full / 16 ;
            if (start < 32) start = 32 ;
            if (start < min_channels) start = min_channels ;
            start = (start + 31) & ~31 ;
            /* Up to a few hundred channels a column costs about the same whatever the count (the
               CTAs of its channel blocks fit one wave of the 148 SMs), so a fraction buys little
               there and a restart costs an attempt plus ~10 ms of set-up: prob159 (88 channels
               needed of 97) and 23 of the 24 BasisLIB bases with Hadamard counts of 64..768
               channels were started on a fraction and always restarted with (nearly) the full
               count.  The fraction pays where the bound is far off: NSR8K 64 of 956, the
               model/newman bases 544-896 of 1900-4300. */
            if (2 * start <= channels_full && ((sc && *sc) || channels_full > SLIP_B200_FULL_START_MAX)) channels = start ;
        }
    }

    {   /* A as limb strings */
        int64_t words = 0 ;
        for (int32_t a = 0 ; a < nz ; a++) words += slip_mpz_words (A->x [a]) ;
        SLIP_TRY (slip_limbs_begin (&Al, nz, words)) ;
        for (int32_t a = 0 ; a < nz ; a++) slip_limbs_put (&Al, a, A->x [a]) ;
    }

    for (int attempt = 0 ; ; attempt++)
    {
        int retry = 0, grow = 0, tight = 0, predict = 0 ;
        /* fewer channels than the a-priori bound: sessions whose factors go to the host prove every
           column's size (mode 1); SLIP_solve_* sessions verify their result exactly at the end, so
           measured sizes are enough there (mode 2, see slip_b200_device.h).  SLIP_B200_BOUND=proven
           keeps mode 1 for both (tests). */
        int bound_mode = 0 ;
        if (channels < channels_full)
        {
            const char *bm = getenv ("SLIP_B200_BOUND") ;
            bound_mode = (want_host_factors || (bm && bm [0] == 'p')) ? 1 : 2 ;
        }
        SLIP_TRY (patterns_init (&P, n, (int64_t) S->lnz + S->unz)) ;
        const double t_attempt = now_s () ;
        tt = now_s () ;
        SLIP_TRY (slip_from_device_status (slipcu_factor_begin (&dev, n, nz, A->p, A->i, Al.limbs, Al.off,
            Al.sign, channels, want_host_factors, bound_mode))) ;
        t_begin += now_s () - tt ;
        slipcu_factor_nowait_singles (dev, use_single) ;
        const int S_dev = slipcu_factor_channels (dev) ;
        const int cap_units = slipcu_factor_capacity_units (dev) ;
        /* lookahead depth: with few channels one column cannot fill the GPU, so the bulk parts of
           the next columns run beside it (NSR8K: 1.5 s without, 0.5 s with six).  With thousands of
           channels a column is HBM-bound on its own, but the GPU idles between a column and its pivot
           (pivot search, host turn): the bulk parts of the next two columns, queued on their own
           streams, fill those gaps -- 2.13 -> 1.87 s on the device at n = 2000 (round 1 measured
           this slower; its second launch per column was three times as expensive) */
        int look = S_dev <= 512 ? 6 : 2 ;
        { const char *lk = getenv ("SLIP_B200_LOOKAHEAD") ; if (lk && *lk) look = atoi (lk) ; }
        double look_min = 1e6 ;          /* element updates (rows x channels) below which a bulk part is not launched */
        { const char *lm = getenv ("SLIP_B200_LOOK_MIN") ; if (lm && *lm) look_min = atof (lm) ; }
        int look_steps = 1 << 30 ;       /* elimination steps from which a bulk part is launched whatever its volume (off: measured neutral to negative) */
        { const char *ls = getenv ("SLIP_B200_LOOK_STEPS") ; if (ls && *ls) look_steps = atoi (ls) ; }
        if (look < 0) look = 0 ;
        if (look > SLIPCU_SPEC_SLOTS - 1) look = SLIPCU_SPEC_SLOTS - 1 ;
        const int D = look > 1 ? look : 1 ;
        ring_size = D + 1 ;
        for (int i = 0 ; i < ring_size ; i++)
        {
            if (!ring [i].pat) ring [i].pat = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
            if (!ring [i].mark) ring [i].mark = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
            if (!ring [i].pat || !ring [i].mark) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
            memset (ring [i].mark, 0, (size_t) n * sizeof (int32_t)) ;
            ring [i].cnt = 0 ; ring [i].spec_slot = -1 ;
        }
        for (int32_t r = 0 ; r < n ; r++) { pinv [r] = r ; row_at [r] = r ; }
        double cum_bits = 0 ;
        work_updates = 0 ; work_limbmul = 0 ;
        /* the patterns of the first D columns start from the entries of A alone */
        for (int32_t c = 0 ; c < D && c < n ; c++)
        {
            ahead_col *e = &ring [c % ring_size] ;
            e->klim = 0 ; e->stamp = c + 1 ; e->spec_slot = -1 ;
            e->cnt = reach_unordered (A, S->q [c], 0, &P, pinv, e->mark, e->stamp, stack, e->pat) ;
        }
        for (int32_t k = 0 ; k < n ; k++)
        {
            const int32_t col = S->q [k] ;
            ahead_col *e = &ring [k % ring_size] ;
            int32_t *pat = e->pat ;
            tt = now_s () ;
            cum_bits += colbits [col] ;
            const int s_had = slip_channels_for_bits (cum_bits) ;
            const int s_k = s_had > S_dev ? S_dev : s_had ;
            /* the columns committed since the pattern was started: when the pivot row of column j
               is in the pattern, its column of L (rows that were not pivotal at time j) joins it */
            int32_t cnt = e->cnt ;
            for (int32_t j = e->klim ; j < k ; j++)
            {
                if (e->mark [row_at [j]] != e->stamp) continue ;
                const int32_t *rj = P.rows + P.ptr [j] ;
                const int32_t cj = (int32_t) (P.ptr [j + 1] - P.ptr [j]) ;
                for (int32_t t = P.nU [j] ; t < cj ; t++)
                {
                    const int32_t rr = rj [t] ;
                    if (e->mark [rr] != e->stamp) { e->mark [rr] = e->stamp ; pat [cnt++] = rr ; }
                }
            }
            order_by_position (n, cnt, pat, pinv, row_at, posflag) ;
            int32_t nU = 0, diag_slot = -1 ;
            for (int32_t t = 0 ; t < cnt ; t++)
            {
                const int32_t pos = pinv [pat [t]] ;
                if (pos < k) upos [nU++] = pos ;
                else if (pat [t] == col) diag_slot = t ;
            }
            if (cnt == nU) { status = SLIP_SINGULAR ; goto cleanup ; }    /* no candidate row at all */
            t_sym += now_s () - tt ; tt = now_s () ;
            int rc = slipcu_factor_column_launch (dev, k, col, cnt, nU, pat, upos, s_k, scheme, diag_slot, e->spec_slot) ;
            e->spec_slot = -1 ;
            if (rc == SLIPCU_BAD_PRIME) { retry = 1 ; break ; }
            SLIP_TRY (slip_from_device_status (rc)) ;
            t_dev += now_s () - tt ; tt = now_s () ;
            /* while the GPU works: the pattern of column k+D as far as the committed columns
               (those before k) determine it, and its bulk part on the device */
            if (k + D < n)
            {
                const int32_t c = k + D ;
                ahead_col *f = &ring [c % ring_size] ;
                f->klim = k ; f->stamp = c + 1 ; f->spec_slot = -1 ;
                f->cnt = reach_unordered (A, S->q [c], k, &P, pinv, f->mark, f->stamp, stack, f->pat) ;
                if (look > 0 && k > 0)
                {
                    memcpy (spat, f->pat, (size_t) f->cnt * sizeof (int32_t)) ;
                    order_by_position (n, f->cnt, spat, pinv, row_at, posflag) ;
                    int32_t snU = 0 ;
                    for (int32_t t = 0 ; t < f->cnt ; t++)
                    {
                        const int32_t pos = pinv [spat [t]] ;
                        if (pos < k) supos [snU++] = pos ;
                    }
                    /* worth a launch of its own only if there is real work in it: three more
                       launches per column cost the host more than a few short steps save the GPU */
                    double bulk = 0 ;
                    for (int32_t u = 0 ; u < snU ; u++)
                        bulk += (double) (P.ptr [supos [u] + 1] - P.ptr [supos [u]] - P.nU [supos [u]]) ;
                    /* ... or a long chain: every step costs the column ~0.3 us of latency whatever
                       its size (barrier, yhat_j, barrier), and a chain run beside the column in
                       flight is off the path to the pivot */
                    if (snU > 0 && (bulk * (double) S_dev >= look_min || snU >= look_steps))
                    {
                        rc = slipcu_factor_spec_launch (dev, c % ring_size, c, S->q [c], f->cnt, snU, spat, supos) ;
                        if (rc == SLIPCU_BAD_PRIME) { retry = 1 ; break ; }
                        SLIP_TRY (slip_from_device_status (rc)) ;
                        f->spec_slot = c % ring_size ;
                    }
                }
            }
            /* bookkeeping that does not depend on pivot k, also while the GPU works */
            for (int32_t u = 0 ; u < nU ; u++)
            {   /* work model: every L entry below the pivot of column upos[u] is updated once */
                const int32_t j = upos [u] ;
                const double len = (double) (P.ptr [j + 1] - P.ptr [j]) - P.nU [j] - 1 ;
                const double w = ceil (cumbits_at [j] / 32.0) ;
                work_updates += len ; work_limbmul += 3.0 * len * w * w ;
            }
            cumbits_at [k] = cum_bits ;
            SLIP_TRY (patterns_reserve (&P, cnt)) ;
            memcpy (P.rows + P.used, pat, (size_t) cnt * sizeof (int32_t)) ;
            memcpy (P.prows + P.used, pat, (size_t) cnt * sizeof (int32_t)) ;
            t_sym += now_s () - tt ; tt = now_s () ;
            /* a single candidate is the pivot whatever its value (zero: singular, reported by
               the device with the next column that is waited for): no round trip for this column */
            const int single = use_single && (cnt - nU == 1) && k < n - 1 ;
            if (single)
            {
                const int32_t prow1 = pat [nU] ;
                const int32_t oldpos = pinv [prow1], displaced = row_at [k] ;
                row_at [k] = prow1 ; row_at [oldpos] = displaced ;
                pinv [prow1] = k ; pinv [displaced] = oldpos ;
                SLIP_TRY (slip_from_device_status (slipcu_factor_set_pivot (dev, k, nU))) ;
                P.used += cnt ;
                P.ptr [k + 1] = P.used ; P.nU [k] = nU ; P.piv [k] = nU ;
                P.pend [k] = cnt - nU ; P.pruned [k] = 0 ;
                if (use_pruning) SLIP_TRY (prune_columns (&P, k, prow1, nU, upos, pinv)) ;
                t_piv += now_s () - tt ;
                continue ;
            }
            slipcu_pivot_info info ;
            rc = slipcu_factor_column_wait (dev, &info) ;
            if (rc == SLIPCU_BAD_PRIME) { retry = 1 ; break ; }
            SLIP_TRY (slip_from_device_status (rc)) ;
            t_dev += now_s () - tt ; tt = now_s () ;
            if (info.singular_col) { status = SLIP_SINGULAR ; goto cleanup ; }
            if (bound_mode && s_had > S_dev && info.bound_units > cap_units)
            {   /* the column is not proven to fit the channels carried: start over with more */
                grow = (int) ceil (((double) info.bound_units / 64.0 + 4.0) / SLIP_B200_CHANNEL_BITS) ;
                /* proven size of this column against its Hadamard prefix bound: when the sizes run
                   close to the bound (random wide entries: ratio ~0.97) no smaller channel count
                   will do, and the next attempt goes straight to the a-priori count */
                tight = cum_bits > 64.0 && (double) info.bound_units / 64.0 >= 0.75 * cum_bits ;
                /* the sizes seen so far, extrapolated along the Hadamard prefix to the last column
                   (they grow roughly in proportion to it), with half as much again on top */
                if (cum_bits > 1.0)
                    predict = (int) ceil (1.5 * ((double) info.bound_units / 64.0) * ((total_bits + extra) / cum_bits)
                                          / SLIP_B200_CHANNEL_BITS) + SLIP_B200_SPARE_CHANNELS ;
                if (timing)
                    fprintf (stderr, "slip_lu_b200 timing: column %d of %d does not fit %d channels (size %.0f bits, Hadamard prefix %.0f of %.0f bits, %.3fs into the attempt)\n",
                        k, n, S_dev, (double) info.bound_units / 64.0, cum_bits, total_bits, now_s () - t_attempt) ;
                retry = 1 ; break ;
            }
            int32_t slot = -1 ;
            SLIP_TRY (decide_pivot (dev, k, scheme, option->tol, diag_slot, &info, &slot)) ;
            const int32_t prow = pat [slot] ;
            {   /* move the pivot row to position k (slip_get_pivot.c:152-172) */
                const int32_t oldpos = pinv [prow], displaced = row_at [k] ;
                row_at [k] = prow ; row_at [oldpos] = displaced ;
                pinv [prow] = k ; pinv [displaced] = oldpos ;
            }
            SLIP_TRY (slip_from_device_status (slipcu_factor_set_pivot (dev, k, slot))) ;
            P.used += cnt ;
            P.ptr [k + 1] = P.used ; P.nU [k] = nU ; P.piv [k] = slot ;
            P.pend [k] = cnt - nU ; P.pruned [k] = 0 ;
            if (use_pruning) SLIP_TRY (prune_columns (&P, k, prow, nU, upos, pinv)) ;
            if (k == n - 1)
            {   /* det = rho[n-1], kept with the resident factors for the rational solve */
                res = (slip_resident *) SLIP_calloc (1, sizeof (slip_resident)) ;
                if (!res) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
                mpz_init (res->det) ;
                const int stride = slipcu_factor_column_stride (dev, k) ;
                uint32_t *w = (uint32_t *) SLIP_malloc ((size_t) (stride + 2) * sizeof (uint32_t)) ;
                if (!w) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
                int32_t nw = 0 ; int8_t sg = 0 ;
                rc = slipcu_factor_fetch_entry (dev, k, slot, w, &nw, &sg) ;
                if (rc == SLIPCU_OK) slip_mpz_from_words (res->det, w, nw, sg) ;
                SLIP_free (w) ;
                SLIP_TRY (slip_from_device_status (rc)) ;
            }
            t_piv += now_s () - tt ;
        }
        if (!retry)
        {   /* the last pivot is only checked against the channel primes here */
            uint32_t badp = 0 ;
            SLIP_TRY (slip_from_device_status (slipcu_factor_bad_prime (dev, &badp))) ;
            if (badp) retry = 1 ;
        }
        if (!retry) break ;
        {
            uint32_t badp = 0 ;
            if (!grow) slipcu_factor_bad_prime (dev, &badp) ;
            slipcu_factor_free (dev) ; dev = NULL ;
            if (res) { slip_resident_free (res) ; res = NULL ; }
            patterns_free (&P) ;
            if (grow)
            {   /* ran out of room: the extrapolated need (lp n=10000: 32 -> 288 channels in one step
                   where quadrupling took 32 -> 128 -> 480 and 40 % of the time in aborted attempts),
                   at least twice the channels and what the failed column asked for, at most the
                   Hadamard count (which needs no proof) */
                int next = 2 * channels ;
                if (next < predict) next = predict ;
                if (next < grow + grow / 4) next = grow + grow / 4 ;
                next = (next + 31) & ~31 ;
                channels = (!tight && 2 * next <= channels_full) ? next : channels_full ;
                bound_restarts++ ;
            }
            else
            {   /* a channel prime divides a pivot (probability ~ n*S/2^31): retire it and start over */
                if (!badp || ++restarts > 4) { slip_set_error ("could not find usable channel primes") ; status = SLIP_INCORRECT ; goto cleanup ; }
                slipcu_retire_prime (badp) ;
            }
        }
    }

    {
        double lnz = 0, unz = 0 ;
        for (int32_t k = 0 ; k < n ; k++) { lnz += (double) (P.ptr [k + 1] - P.ptr [k]) - P.nU [k] ; unz += P.nU [k] + 1 ; }
        slip_last_stats.n = n ; slip_last_stats.nnz_L = lnz ; slip_last_stats.nnz_U = unz ;
        slip_last_stats.channels = slipcu_factor_channels (dev) ;
        slip_last_stats.channels_hadamard = channels_full ;
        slip_last_stats.bound_restarts = bound_restarts ;
        slip_last_stats.updates = work_updates ; slip_last_stats.limb_mul_equiv = work_limbmul ;
        slip_last_stats.t_symbolic = t_sym ; slip_last_stats.t_device = t_dev ; slip_last_stats.t_begin = t_begin ;
        slip_last_stats.t_factor_total = now_s () - t0 ;
    }
    if (timing)
        fprintf (stderr, "slip_lu_b200 timing: setup %.3fs (device begin %.3fs) symbolic %.3fs device columns %.3fs pivot/commit %.3fs total-so-far %.3fs\n",
            0.0, t_begin, t_sym, t_dev, t_piv, now_s () - t0) ;
    if (want_host_factors)
    {
        int64_t lnz = 0, unz = 0 ;
        for (int32_t k = 0 ; k < n ; k++)
        {
            const int32_t cnt = (int32_t) (P.ptr [k + 1] - P.ptr [k]) ;
            lnz += cnt - P.nU [k] ; unz += P.nU [k] + 1 ;
        }
        if (lnz > INT32_MAX || unz > INT32_MAX) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
        SLIP_sparse *M [2] = { L, U } ;
        int64_t mnz [2] = { lnz, unz } ;
        for (int w = 0 ; w < 2 ; w++)
        {
            M [w]->m = M [w]->n = n ;
            M [w]->nz = M [w]->nzmax = (int32_t) mnz [w] ;
            M [w]->p = (int32_t *) SLIP_calloc ((size_t) n + 1, sizeof (int32_t)) ;
            M [w]->i = (int32_t *) SLIP_calloc ((size_t) mnz [w], sizeof (int32_t)) ;
            M [w]->x = (mpz_t *) SLIP_calloc ((size_t) mnz [w], sizeof (mpz_t)) ;
            if (!M [w]->p || !M [w]->i || !M [w]->x) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
        }
        for (int32_t k = 0 ; k < n ; k++)
        {
            const int32_t *rows = P.rows + P.ptr [k] ;
            const int32_t cnt = (int32_t) (P.ptr [k + 1] - P.ptr [k]), nU = P.nU [k] ;
            int32_t *Ui = U->i + U->p [k], *Li = L->i + L->p [k] ;
            for (int32_t t = 0 ; t < nU ; t++) Ui [t] = pinv [rows [t]] ;
            Ui [nU] = k ;
            for (int32_t t = nU ; t < cnt ; t++) Li [t - nU] = pinv [rows [t]] ;
            U->p [k + 1] = U->p [k] + nU + 1 ;
            L->p [k + 1] = L->p [k] + (cnt - nU) ;
        }
        assemble_ctx ctx = { L, U, rhos, &P } ;
        SLIP_TRY (slip_from_device_status (slipcu_factor_download (dev, assemble_column, &ctx))) ;
    }

    res->dev = dev ; dev = NULL ;
    res->n = n ;
    res->proven_channels = channels >= channels_full ;
    res->total_bits = total_bits ;
    res->min_col_bits = min_bits ;
    res->Lx = want_host_factors ? (const void *) L->x : NULL ;
    res->Ux = want_host_factors ? (const void *) U->x : NULL ;
    if (want_host_factors && !res->proven_channels)
    {   /* SLIP_LU_solve has no A: bound-mode factors keep a copy to verify their solutions with */
        res->A_copy = slip_sparse_copy (A) ;
        res->q_copy = (int32_t *) SLIP_malloc ((size_t) n * sizeof (int32_t)) ;
        if (!res->A_copy || !res->q_copy) { status = SLIP_OUT_OF_MEMORY ; goto cleanup ; }
        memcpy (res->q_copy, S->q, (size_t) n * sizeof (int32_t)) ;
    }
    if (resident) { *resident = res ; res = NULL ; }
    else if (want_host_factors) { slip_resident_add (res) ; res = NULL ; }

cleanup:
    if (dev) slipcu_factor_free (dev) ;
    if (res) slip_resident_free (res) ;
    slip_limbs_free (&Al) ;
    patterns_free (&P) ;
    SLIP_free (colbits) ; SLIP_free (row_at) ; SLIP_free (stack) ;
    SLIP_free (upos) ; SLIP_free (cumbits_at) ; SLIP_free (posflag) ;
    SLIP_free (spat) ; SLIP_free (supos) ;
    for (int i = 0 ; i < SLIPCU_SPEC_SLOTS ; i++) { SLIP_free (ring [i].pat) ; SLIP_free (ring [i].mark) ; }
    return status ;
}

SLIP_info SLIP_LU_factorize (SLIP_sparse *L, SLIP_sparse *U, SLIP_sparse *A, SLIP_LU_analysis *S,
    mpz_t *rhos, int32_t *pinv, SLIP_options *option)
{
    if (!A || !L || !U || !S || !rhos || !pinv || !option || !A->p || !A->x || !A->i)
        return SLIP_INCORRECT_INPUT ;
    return slip_factorize_driver (L, U, A, S, rhos, pinv, option, 1, NULL, 0.0, 0) ;
}
