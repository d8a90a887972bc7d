/* slip_b200_device.h -- thin C ABI between the C host code (slip_lu_b200/csrc/host) and the
 * CUDA kernels for sm_100a (slip_lu_b200/csrc/cuda/slipcu.cu).
 *
 * Plain pointers and sizes only.  Big integers cross this boundary as sign + little-endian
 * 32-bit limb strings (two of them make one GMP limb on x86-64).  Everything behind it runs
 * on the GPU in a residue number system (31-bit prime channels, Montgomery form) plus
 * positional 32-bit-limb reconstruction; see DESIGN.md.
 *
 * Each entry point names the reference routine whose work it takes over.
 */
#ifndef SLIP_B200_DEVICE_H
#define SLIP_B200_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIPCU_OK            0
#define SLIPCU_OUT_OF_MEMORY (-1)
#define SLIPCU_SINGULAR      (-2)
#define SLIPCU_BAD_INPUT     (-3)
#define SLIPCU_CUDA_ERROR    (-10)   /* no device, launch failure, ...: the caller must fail loudly */
#define SLIPCU_BAD_PRIME     (-11)   /* a pivot vanished modulo a channel prime: retire it and retry */

typedef struct slipcu_factor slipcu_factor;     /* one factorization; keeps L, U, rho resident in HBM */

/* result of the exact pivot scan of one column (slip_get_pivot.c:46-150 and the three
 * slip_get_*_pivot.c scans).  Slots index the column's pattern (sorted by row position). */
typedef struct
{
    int32_t best_slot;      /* extreme (smallest / largest / first nonzero) eligible entry, -1: none */
    int32_t best_sign;      /* sign of that entry */
    int32_t diag_eligible;  /* 1 if the diagonal candidate is in the L part and nonzero */
    int32_t diag_vs_best;   /* cmp(|diag|, |best|): -1, 0, +1 (valid if diag_eligible) */
    int32_t bad_channel;    /* 0, or 1 + index of a channel whose prime divides an earlier pivot */
    int32_t reserved[3];
    int32_t bound_units;    /* bound mode: proven upper bound of 64*log2 |entry| so far; measured mode: largest measured candidate so far (running maxima over the session) */
    int32_t singular_col;   /* 0, or 1 + the first column whose candidates were all zero (running: the
                               caller may skip slipcu_factor_column_wait for columns with a single
                               candidate, whose pivot needs no search, and learns of a zero there
                               from the next column it does wait for) */
    int32_t seq;            /* written last by the kernel that filled the record (sequence number of the
                               scan): the record lives in mapped host memory and slipcu_factor_column_wait
                               polls this word instead of sleeping on an event */
    int32_t pad[1];
} slipcu_pivot_info;

/* receives column k of the factorization as positional integers.  `limbs` holds `cnt` rows of
 * `stride` 32-bit limbs (|value|, little endian, zero padded, stride even); nlimbs32[e] is the
 * number of significant 32-bit limbs, sign[e] in {-1,0,1}.  The pointers are only valid during
 * the call. */
typedef int (*slipcu_column_sink) (void *user, int k, int cnt, int stride,
                                   const uint32_t *limbs, const int32_t *nlimbs32, const int8_t *sign);

const char *slipcu_last_error (void);
int  slipcu_device_count (void);
/* choose the CUDA device of this PROCESS (multi-GPU sharding: one rank per GPU).  Sessions created
 * afterwards, on any thread, live on that device; every entry point makes its session's device
 * current, and the allocation pool and the channel tables are kept per device. */
int  slipcu_set_device (int device);
/* diagnostics: measured 32-bit integer-multiply peaks of the device (register-resident chains of
 * IMAD.WIDE, IMAD and IMAD.HI), the denominators of bench.py's roofline.int_mul */
int  slipcu_measure_imad_peak (double *wide_per_s, double *lo_per_s, double *hi_per_s);
/* ... and of the modular multiply-subtract of k_trisolve itself (w <- w + l*y mod p with a Shoup
 * product: IMAD.HI + 2 IMAD + three integer ALU operations) on registers only */
int  slipcu_measure_modmul_peak (double *modmul_per_s);
/* all at once: out[0..4] operations per second as run, out[5..9] operations per SM clock cycle
 * (clock64 inside the kernels: independent of the clock the GPU holds under this power-hungry load),
 * for IMAD.WIDE, IMAD, IMAD.HI, the Montgomery multiply-subtract (IMAD.WIDE + IMAD + IMAD.HI) and the
 * Shoup multiply-subtract of k_trisolve (IMAD.HI + 2 IMAD) */
int  slipcu_measure_int_peaks (double *out10);

/* -- factorization session -------------------------------------------------------------------
 * slipcu_factor_begin: uploads A (CSC; values as limb strings) and reduces it into `channels`
 *   residue channels.  Takes over the workspace set-up of SLIP_LU_factorize.c:60-189.
 *   Avalue_off has nz+1 entries: limbs of entry a are Alimbs[Avalue_off[a] .. Avalue_off[a+1]). */
int slipcu_factor_begin (slipcu_factor **F, int n, int nz, const int32_t *Ap, const int32_t *Ai,
                         const uint32_t *Alimbs, const int64_t *Avalue_off, const int8_t *Asign,
                         int channels, int keep_positional, int bound_mode);
/* bound_mode = 0: `channels` covers a proven a-priori bound (Hadamard) of every entry of L and U.
 *   1: `channels` is smaller than that.  The session then proves the size of every column as it
 *   goes (a bound propagated through the elimination from the MEASURED sizes of the finished
 *   columns) and reports it in slipcu_pivot_info.bound_units; the caller compares it with
 *   slipcu_factor_capacity_units and restarts with more channels when a column does not fit.
 *   GMP pays for the actual size of its operands (slip_REF_triangular_solve.c:150-232 on mpz_t);
 *   this is how the residue representation does the same without giving up exactness.
 *   2: MEASURED mode, for callers that verify their final result exactly (SLIP_solve_*: the
 *   numerators N of x are checked against A N = det b over the integers, and x is unique).
 *   Residue arithmetic is exact for the TRUE integers whatever the channel count, and a candidate
 *   with a nonzero residue is truly nonzero, so every pivot sequence the session produces is a valid
 *   factorization; too few channels can only mislead the magnitude comparison of the pivot search.
 *   The session therefore measures the candidates of every column (bound_units = the largest one;
 *   a value that does not fit the channels reconstructs as a uniform residue of the modulus, i.e.
 *   full size, except with probability 2^-35) and the caller restarts with more channels when the
 *   measured size reaches the capacity.  No bound is propagated, so determinants far below their
 *   Hadamard bound (LP bases) run on the channels their values need. */
int slipcu_factor_capacity_units (const slipcu_factor *F);
/* keep_positional = 1: every entry of L and U is reconstructed as a positional integer and kept
 *   for slipcu_factor_download (SLIP_LU_factorize).  0: only what the pivot scan needs is
 *   reconstructed (SLIP_solve_*: the factors never leave the GPU). */

/* slipcu_factor_column: the sparse REF triangular solve of column k (slip_REF_triangular_solve.c:
 *   84-262) on the pattern computed by the host, followed by exact reconstruction of the column and
 *   the exact nonzero/magnitude pivot scan (slip_get_pivot.c).
 *   rows[0..cnt): pattern, original row indices, sorted by current position; the first nU are
 *   already pivotal and upos[u] is the pivot position of rows[u].  recon_channels: number of
 *   channels that bound the column's entries.  scheme: SLIP_pivot code.  diag_slot: slot of row
 *   `col` in the pattern if it is a candidate, else -1.  Blocks until the scan result is back. */
int slipcu_factor_column (slipcu_factor *F, int k, int col, int cnt, int nU,
                          const int32_t *rows, const int32_t *upos, int recon_channels,
                          int scheme, int diag_slot, slipcu_pivot_info *info);

/* the same in two halves, so that the host can prepare the next column's symbolic pattern while
 * the GPU works: _launch enqueues everything and returns, _wait blocks for the scan result. */
int slipcu_factor_column_launch (slipcu_factor *F, int k, int col, int cnt, int nU,
                                 const int32_t *rows, const int32_t *upos, int recon_channels,
                                 int scheme, int diag_slot, int spec_slot);
int slipcu_factor_column_wait (slipcu_factor *F, slipcu_pivot_info *info);
/* A column with exactly one candidate (cnt - nU == 1) has its pivot without a search: the caller may
 * go straight to slipcu_factor_set_pivot (k, nU) and on to the next column without _wait; the
 * column's reconstruction and scan still run (sizes, zero test) and report through the running
 * fields of the next slipcu_pivot_info.  A caller that does so for every such column except the
 * last says so with slipcu_factor_nowait_singles (F, 1): sessions that need neither sizes nor
 * digits then replace the reconstruction of those columns by a zero test of the residues. */
void slipcu_factor_nowait_singles (slipcu_factor *F, int on);

/* lookahead: the bulk part of the column that will be column k, i.e. the steps of
 * slip_REF_triangular_solve.c:150-232 with every pivot committed so far, on the pattern reachable
 * through those columns (rows sorted by position, nU of them pivotal, upos their pivot positions),
 * in lookahead slot `slot` (0 .. SLIPCU_SPEC_SLOTS-1), on the slot's own stream beside the column
 * in flight.  The later slipcu_factor_column_launch (.., spec_slot = slot) for column k starts from
 * its result and applies the remaining steps (those from U slot nU on). */
int slipcu_factor_spec_launch (slipcu_factor *F, int slot, int k, int col, int cnt, int nU,
                               const int32_t *rows, const int32_t *upos);
#define SLIPCU_SPEC_SLOTS 16

/* one reconstructed entry of the current column (used only for the rational tolerance test of
 * SLIP_TOL_SMALLEST / SLIP_TOL_LARGEST, slip_get_pivot.c:94-143).  limbs must hold stride words. */
int slipcu_factor_fetch_entry (slipcu_factor *F, int k, int slot, uint32_t *limbs,
                               int32_t *nlimbs32, int8_t *sign);
int slipcu_factor_column_stride (slipcu_factor *F, int k);

/* commits the pivot of column k: rho_k, its modular inverses, the L column descriptor
 * (slip_get_pivot.c:152-175). */
int slipcu_factor_set_pivot (slipcu_factor *F, int k, int slot);

/* streams every column of the finished factorization to the host (SLIP_LU_factorize.c:224-262,
 * where the reference copies x into L and U). */
int slipcu_factor_download (slipcu_factor *F, slipcu_column_sink sink, void *user);

/* slipcu_factor_upload: a resident session built from factors held by the host (SLIP_LU_solve on
 *   L, U that are not resident, or whose right-hand side needs more channels).  Column k has
 *   colcnt[k] entries in slot order: colnU[k] entries of U(:,k) above the diagonal, then the
 *   entries of L(:,k) (the diagonal, slot colpiv[k], is among them); rows[] are FINAL positions;
 *   values are limb strings like A. */
int slipcu_factor_upload (slipcu_factor **F, int n, int channels, const int32_t *colcnt,
                          const int32_t *colnU, const int32_t *colpiv, const int32_t *rows,
                          const uint32_t *limbs, const int64_t *value_off, const int8_t *sign);

/* -- solve ------------------------------------------------------------------------------------
 * slipcu_solve: exact forward substitution, scaling by det and back substitution
 *   (SLIP_LU_solve.c:74-91, slip_forward_sub.c, slip_array_mul.c, slip_back_sub.c) for nrhs
 *   right-hand sides against the resident factors.  b holds the right-hand sides one after the
 *   other (entry i of right-hand side c at index c*n + i), rows in ORIGINAL order; pinv is the final
 *   inverse row permutation (checked to be a permutation).  The sink receives one "column" per
 *   right-hand side: cnt = n entries in factor (position) order holding det * x_i, the numerators
 *   of slip_array_div.c.  Right-hand sides are processed in batches (one launch of each kernel per
 *   batch, SLIP_B200_SOLVE_BATCH_MB of device memory); the sink of a batch runs on the calling
 *   thread while the GPU works on the next batch. */
int slipcu_solve (slipcu_factor *F, int nrhs, const uint32_t *blimbs, const int64_t *bvalue_off,
                  const int8_t *bsign, const int32_t *pinv, int recon_channels,
                  slipcu_column_sink sink, void *user, int32_t *top_digit_max);
/* top_digit_max (may be NULL): receives the highest mixed-radix digit index any result uses, so a
 * caller that only has an ESTIMATE of the result size can verify that the top channels stayed
 * empty (and retry with more channels otherwise). */

/* channels available in the session (>= the count passed to begin, rounded up) */
int  slipcu_factor_channels (const slipcu_factor *F);
/* log2 of the product of the first `count` channel primes, rounded down (for sizing) */
double slipcu_channel_bits (int count);
void slipcu_factor_free (slipcu_factor *F);

/* after SLIPCU_BAD_PRIME: which channel prime failed (0: none); then retire it, by value, so that
 * the next session does not use it */
int  slipcu_factor_bad_prime (slipcu_factor *F, uint32_t *prime);
int  slipcu_retire_prime (uint32_t prime);

/* counters for bench.py: kernels launched by this library since process start, and the
 * accumulated device time (ms, CUDA events) and algorithmic bytes of the triangular-solve kernel */
typedef struct
{
    uint64_t launches;            /* all kernels */
    uint64_t trisolve_launches;
    double   trisolve_ms;         /* only accumulated while profiling is enabled */
    double   trisolve_bytes;      /* algorithmic bytes: L entries streamed + column written */
    double   trisolve_modmul;     /* 32-bit limb (modular) multiplies: 4 per entry update */
    double   recon_ms;            /* garner + to_limbs */
    double   recon_mac;           /* 32x32 multiply-accumulates in reconstruction */
    double   h2d_bytes, d2h_bytes;/* bytes this library copied host->device / device->host */
    double   device_ms;           /* CUDA-event time from "A resident in HBM" to "solution numerators
                                     reconstructed in HBM", accumulated over slipcu_solve calls; sessions that
                                     never reach a solve (attempts aborted for more channels or a retired prime)
                                     add their time when they are freed */
    double   other_ms;            /* symbolic pre-pass + pivot scan kernels (profiling mode) */
    double   trisolve_union_ms;   /* time during which at least one k_trisolve launch was running (launches
                                     on the lookahead streams overlap: trisolve_ms counts shared time twice) */
} slipcu_counters;
void slipcu_get_counters (slipcu_counters *out);
void slipcu_reset_counters (void);
void slipcu_set_profiling (int enabled);
/* device blocks and pinned host buffers of finished sessions are cached per process (cudaMalloc and
 * cudaHostAlloc cost milliseconds and serialise the device); this hands every cached one back */
void slipcu_release_cached_memory (void);   /* 1: bracket kernels with CUDA events (adds syncs) */

#ifdef __cplusplus
}
#endif
#endif
