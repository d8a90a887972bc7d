// slipcu.cu -- CUDA kernels (sm_100a) and the thin C ABI of include/slip_b200_device.h.
//
// Exact sparse left-looking REF LU on the GPU.  Every integer of the factorization lives in HBM
// as its residues modulo S 31-bit primes ("channels", Montgomery form, channel-blocked layout);
// the REF update  x_i <- (rho_j * x_i - l_ij * x_j) / rho_{j-1}  becomes one fused
// multiply-add-reduce per channel, the exact division being a multiplication by the inverse of
// the previous pivot.  Exact values (for the pivot scan and for the mpz_t output of the C
// interface) are reconstructed per column with a mixed-radix (Garner) pass and a positional
// 32-bit-limb carry-chain pass.  See DESIGN.md for layout, bounds and rooflines.
//
// Reference routines replaced (cjh10644/SLIP_LU):
//   k_trisolve        slip_REF_triangular_solve.c:84-262, slip_forward_sub.c:64-155, and (the U parts
//                     as steps n-1..0) slip_array_mul.c + slip_back_sub.c:30-56
//   k_prep, k_slots   the index arithmetic of the reference's inner loops (pinv[L->i[m]]), once per
//                     column for all channels; k_prep also commits the previous pivot
//   k_garner*, k_limbs  (GMP keeps values positional; here reconstruction is explicit)
//   pivot_scan_body   slip_get_pivot.c:46-150, slip_get_{smallest,largest,nonzero}_pivot.c (run by the
//                     last CTA of the reconstruction launch); k_fraccrt/k_fracselect: the same search
//                     on approximate magnitudes with a proven-choice rule
//   pivot_commit_body slip_get_pivot.c:152-175
//   k_residues        slip_get_column.c (input side)
//   k_backsub         round 1's back substitution (kept behind SLIP_B200_BACKSUB=0)

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>

#include "slip_b200_device.h"

typedef unsigned long long u64;
typedef uint32_t u32;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail (int code, const char *what, const char *detail)
{
    g_err = std::string (what) + ": " + (detail ? detail : "");
    return code;
}
static int g_debug_sync = -1;     // SLIP_B200_DEBUG_SYNC=1: synchronize and check after every launch
static int debug_check (const char *what, cudaStream_t st)
{
    if (g_debug_sync < 0) { const char *v = getenv ("SLIP_B200_DEBUG_SYNC"); g_debug_sync = (v && *v == '1') ? 1 : 0; }
    if (!g_debug_sync) return 0;
    cudaError_t e = cudaStreamSynchronize (st);
    if (e == cudaSuccess) e = cudaGetLastError ();
    if (e != cudaSuccess) { fprintf (stderr, "slipcu debug: %s failed: %s\n", what, cudaGetErrorString (e)); return 1; }
    return 0;
}
#define CU(call)                                                                          \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess)                                  \
         return fail (e_ == cudaErrorMemoryAllocation ? SLIPCU_OUT_OF_MEMORY : SLIPCU_CUDA_ERROR, \
                      #call, cudaGetErrorString (e_)); } while (0)

extern "C" const char *slipcu_last_error (void) { return g_err.c_str (); }

static std::atomic<uint64_t> g_launches{0}, g_tri_launches{0};
static double g_h2d_bytes = 0, g_d2h_bytes = 0, g_device_ms = 0, g_other_ms = 0;
static double g_tri_ms = 0, g_tri_bytes = 0, g_tri_modmul = 0, g_recon_ms = 0, g_recon_mac = 0;
static int g_profiling = 0;
static double g_hw[8] = {0,0,0,0,0,0,0,0};     // host wall-clock per section of slipcu_factor_column (debug)
#include <chrono>
static inline double wall_s () { return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count (); }

// ------------------------------------------------------------------------------------------------
// Montgomery arithmetic modulo a prime p < 2^31, R = 2^32
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 csub (u32 r, u32 p)
{   // r in [0, 2p) -> [0, p): r - p wraps above r exactly when r < p
    const u32 d = r - p;
    return d < r ? d : r;
}
__host__ __device__ __forceinline__ u32 mont_redc (u64 T, u32 p, u32 ninv)
{   // requires T < p * 2^32; returns T / R mod p in [0, p)
    u32 m = (u32) T * ninv;
    u64 t = T + (u64) m * p;
    return csub ((u32) (t >> 32), p);
}
__host__ __device__ __forceinline__ u32 mont_mul (u32 a, u32 b, u32 p, u32 ninv)
{
    return mont_redc ((u64) a * b, p, ninv);
}
__host__ __device__ __forceinline__ u32 mont_pow (u32 a_m, u32 e, u32 one_m, u32 p, u32 ninv)
{
    u32 r = one_m;
    while (e) { if (e & 1) r = mont_mul (r, a_m, p, ninv); a_m = mont_mul (a_m, a_m, p, ninv); e >>= 1; }
    return r;
}
__host__ __device__ __forceinline__ u32 add_mod (u32 a, u32 b, u32 p) { return csub (a + b, p); }
__host__ __device__ __forceinline__ u32 reduce_word (u32 w, u32 p)
{   // w < 2^32 < 4p  (p > 2^30)
    if (w >= p) w -= p;
    if (w >= p) w -= p;
    if (w >= p) w -= p;
    return w;
}
// 64-bit lazy accumulator for sums of products of 31-bit factors.  Invariant: T < 2^63.
__device__ __forceinline__ void lazy_mac (u64 &T, u32 a, u32 b, u32 p)
{
    T += (u64) a * b;                       // < 2^63 + 2^62
    if (T >> 63) T -= ((u64) p << 32);      // p * 2^32 >= 2^62, keeps T < 2^63 and T mod p
}
// two products per range check: T < 2^63 and each product <= (p-1)^2 < 2^62, so the sum cannot
// wrap; if it reaches 2^63, T - p*2^32 <= 2^63 - 1 + 2(p-1)^2 - p*2^32 < 2^63 because 2p < 2^32
__device__ __forceinline__ void lazy_mac2 (u64 &T, u32 a0, u32 b0, u32 a1, u32 b1, u32 p)
{
    T += (u64) a0 * b0;
    T += (u64) a1 * b1;
    if (T >> 63) T -= ((u64) p << 32);
}
__device__ __forceinline__ u32 lazy_redc (u64 T, u32 p, u32 ninv)
{
    u32 hi = reduce_word ((u32) (T >> 32), p);
    return mont_redc (((u64) hi << 32) | (u32) T, p, ninv);
}

// ------------------------------------------------------------------------------------------------
// channel primes and reconstruction tables (shared by all sessions, immutable once built)
// ------------------------------------------------------------------------------------------------
#define FRAC_WMAX 256          // words of the fixed-point reciprocals 1/p_i

struct Tables
{
    int S = 0;                 // channels covered
    std::vector<u32> hp;       // host copy of the primes
    std::vector<double> cumbits;   // cumbits[c] = log2(p_0 ... p_{c-1})
    u32 *p = nullptr, *ninv = nullptr, *r2 = nullptr, *one = nullptr;
    u32 *C = nullptr;          // [S][S]  C[u][t] = (p_0..p_{u-1}) mod p_t, Montgomery form, u < t
    u32 *invB = nullptr;       // [S]     (p_0..p_{t-1})^-1 mod p_t, Montgomery form
    u32 *Bpos = nullptr;       // [S][LB] limbs of p_0..p_{t-1}
    int LB = 0;
    // approximate magnitudes (k_fraccrt): Minv[(s-1)*S + i] = (p_0..p_{s-1} / p_i)^-1 mod p_i in
    // Montgomery form for i < s; Urec[i] = floor(2^(32*FRAC_WMAX) / p_i), most significant word first
    u32 *Minv = nullptr, *Urec = nullptr;
    int32_t *cum_ub = nullptr; // [S+1] upper bound of 64 log2 (p_0..p_{c-1}) (bound mode)
    ~Tables ()
    {
        cudaFree (p); cudaFree (ninv); cudaFree (r2); cudaFree (one);
        cudaFree (C); cudaFree (invB); cudaFree (Bpos); cudaFree (Minv); cudaFree (Urec); cudaFree (cum_ub);
    }
};

static std::mutex g_tab_mutex;
static std::vector<u32> g_primes;          // descending from 2^31 - 1, retired ones removed
static u32 g_next_candidate = 0x7fffffffu;
static std::map<int, std::shared_ptr<Tables>> g_tables_by_device;   // tables live in device memory: one set per device

static u32 powmod_u32 (u32 b, u32 e, u32 m)
{
    u64 r = 1, x = b % m;
    while (e) { if (e & 1) r = r * x % m; x = x * x % m; e >>= 1; }
    return (u32) r;
}
static bool is_prime_u32 (u32 x)
{   // deterministic Miller-Rabin for 32-bit integers (bases 2, 3, 5, 7)
    if (x < 2) return false;
    for (u32 q : {2u, 3u, 5u, 7u, 11u, 13u, 17u, 19u, 23u, 29u, 31u, 37u})
        if (x % q == 0) return x == q;
    u32 d = x - 1; int r = 0;
    while ((d & 1) == 0) { d >>= 1; ++r; }
    for (u32 a : {2u, 3u, 5u, 7u})
    {
        u32 y = powmod_u32 (a, d, x);
        if (y == 1 || y == x - 1) continue;
        bool comp = true;
        for (int i = 1; i < r && comp; ++i) { y = (u32) ((u64) y * y % x); if (y == x - 1) comp = false; }
        if (comp) return false;
    }
    return true;
}
static void extend_primes (size_t count)
{
    while (g_primes.size () < count)
    {
        while (!is_prime_u32 (g_next_candidate)) g_next_candidate -= 2;
        g_primes.push_back (g_next_candidate);
        g_next_candidate -= 2;
    }
}

__global__ void k_build_tables (int S, const u32 *p, const u32 *ninv, const u32 *r2, const u32 *one,
                                u32 *C, u32 *invB)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S) return;
    const u32 pt = p[t], ni = ninv[t], rr = r2[t];
    u32 acc = one[t];
    for (int u = 0; u < t; ++u)
    {
        C[(size_t) u * S + t] = acc;
        u32 pu = mont_mul (reduce_word (p[u], pt), rr, pt, ni);
        acc = mont_mul (acc, pu, pt, ni);
    }
    invB[t] = mont_pow (acc, pt - 2, one[t], pt, ni);
}

// Minv: one thread per channel i walks the prefix lengths s = i+1 .. S
__global__ void k_build_minv (int S, const u32 *p, const u32 *ninv, const u32 *r2, const u32 *one, u32 *Minv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const u32 pi = p[i], ni = ninv[i], rr = r2[i];
    u32 prod = one[i];
    for (int j = 0; j < i; ++j) prod = mont_mul (prod, mont_mul (reduce_word (p[j], pi), rr, pi, ni), pi, ni);
    for (int s = i + 1; s <= S; ++s)
    {
        if (s > i + 1) prod = mont_mul (prod, mont_mul (reduce_word (p[s - 1], pi), rr, pi, ni), pi, ni);
        Minv[(size_t) (s - 1) * S + i] = mont_pow (prod, pi - 2, one[i], pi, ni);
    }
}

static int build_tables (int S, std::shared_ptr<Tables> &out)
{
    auto T = std::make_shared<Tables> ();
    T->S = S;
    extend_primes ((size_t) S);
    T->hp.assign (g_primes.begin (), g_primes.begin () + S);
    std::vector<u32> ninv (S), r2 (S), one (S);
    T->cumbits.resize (S + 1);
    T->cumbits[0] = 0.0;
    for (int c = 0; c < S; ++c)
    {
        u32 p = T->hp[c];
        u32 inv = 1;                          // Newton: inv = p^-1 mod 2^32
        for (int it = 0; it < 5; ++it) inv *= 2u - p * inv;
        ninv[c] = (u32) (0u - inv);
        u64 r = ((u64) 1 << 32) % p;
        one[c] = (u32) r;
        r2[c] = (u32) ((r * r) % p);
        T->cumbits[c + 1] = T->cumbits[c] + log2 ((double) p);
    }
    // prefix products as limb strings: row t = p_0 .. p_{t-1}  (t limbs at most)
    T->LB = S + 1;
    std::vector<u32> B ((size_t) (S + 4) * T->LB, 0u);      // 4 spare zero rows: k_limbs reads rows in fours
    {
        std::vector<u32> cur (T->LB, 0u);
        cur[0] = 1; int len = 1;
        for (int t = 0; t < S; ++t)
        {
            memcpy (&B[(size_t) t * T->LB], cur.data (), (size_t) len * sizeof (u32));
            u64 carry = 0;
            for (int l = 0; l < len; ++l)
            {
                u64 v = (u64) cur[l] * T->hp[t] + carry;
                cur[l] = (u32) v; carry = v >> 32;
            }
            if (carry) cur[len++] = (u32) carry;
        }
    }
    CU (cudaMalloc (&T->p, S * sizeof (u32)));
    CU (cudaMalloc (&T->ninv, S * sizeof (u32)));
    CU (cudaMalloc (&T->r2, S * sizeof (u32)));
    CU (cudaMalloc (&T->one, S * sizeof (u32)));
    CU (cudaMalloc (&T->invB, S * sizeof (u32)));
    CU (cudaMalloc (&T->C, (size_t) S * S * sizeof (u32)));
    CU (cudaMalloc (&T->Bpos, (size_t) (S + 4) * T->LB * sizeof (u32)));
    CU (cudaMemcpy (T->p, T->hp.data (), S * sizeof (u32), cudaMemcpyHostToDevice));
    CU (cudaMemcpy (T->ninv, ninv.data (), S * sizeof (u32), cudaMemcpyHostToDevice));
    CU (cudaMemcpy (T->r2, r2.data (), S * sizeof (u32), cudaMemcpyHostToDevice));
    CU (cudaMemcpy (T->one, one.data (), S * sizeof (u32), cudaMemcpyHostToDevice));
    CU (cudaMemcpy (T->Bpos, B.data (), B.size () * sizeof (u32), cudaMemcpyHostToDevice));
    {
        std::vector<int32_t> cu (S + 1);
        for (int c = 0; c <= S; ++c) cu[c] = (int32_t) ceil (64.0 * T->cumbits[c]) + 1;
        CU (cudaMalloc (&T->cum_ub, (S + 1) * sizeof (int32_t)));
        CU (cudaMemcpy (T->cum_ub, cu.data (), (S + 1) * sizeof (int32_t), cudaMemcpyHostToDevice));
    }
    CU (cudaMemset (T->C, 0, (size_t) S * S * sizeof (u32)));
    k_build_tables<<<(S + 127) / 128, 128>>> (S, T->p, T->ninv, T->r2, T->one, T->C, T->invB);
    g_launches++;
    CU (cudaGetLastError ());
    {   // tables of the approximate magnitude pass
        std::vector<u32> U ((size_t) S * FRAC_WMAX);
        for (int c = 0; c < S; ++c)
        {   // long division of 2^(32*FRAC_WMAX) by p_c, one 32-bit word of the quotient per step
            const u64 pc = T->hp[c];
            u64 rem = 1;
            for (int w = 0; w < FRAC_WMAX; ++w)
            {
                const u64 num = rem << 32;
                U[(size_t) c * FRAC_WMAX + w] = (u32) (num / pc);
                rem = num % pc;
            }
        }
        CU (cudaMalloc (&T->Urec, U.size () * sizeof (u32)));
        CU (cudaMemcpy (T->Urec, U.data (), U.size () * sizeof (u32), cudaMemcpyHostToDevice));
        CU (cudaMalloc (&T->Minv, (size_t) S * S * sizeof (u32)));
        CU (cudaMemset (T->Minv, 0, (size_t) S * S * sizeof (u32)));
        k_build_minv<<<(S + 63) / 64, 64>>> (S, T->p, T->ninv, T->r2, T->one, T->Minv);
        g_launches++;
        CU (cudaGetLastError ());
    }
    CU (cudaDeviceSynchronize ());
    out = T;
    return SLIPCU_OK;
}

static int get_tables (int S, std::shared_ptr<Tables> &out)
{
    std::lock_guard<std::mutex> lk (g_tab_mutex);
    int dev = 0;
    CU (cudaGetDevice (&dev));
    std::shared_ptr<Tables> &g_tables = g_tables_by_device[dev];
    if (g_tables && g_tables->S >= S) { out = g_tables; return SLIPCU_OK; }
    int want = S;
    if (g_tables) want = std::max (S, g_tables->S + g_tables->S / 4);
    want = (want + 31) & ~31;
    std::shared_ptr<Tables> T;
    int rc = build_tables (want, T);
    if (rc) return rc;
    g_tables = T; out = T;
    return SLIPCU_OK;
}

extern "C" double slipcu_channel_bits (int count)
{
    std::lock_guard<std::mutex> lk (g_tab_mutex);
    extend_primes ((size_t) std::max (count, 0));
    double b = 0;
    for (int c = 0; c < count; ++c) b += log2 ((double) g_primes[c]);
    return b;
}

// retires a channel prime BY VALUE (an index would be into the failing session's snapshot of the
// prime list, which another thread may have changed since)
extern "C" int slipcu_retire_prime (uint32_t prime)
{
    std::lock_guard<std::mutex> lk (g_tab_mutex);
    auto it = std::find (g_primes.begin (), g_primes.end (), prime);
    if (it == g_primes.end ()) return SLIPCU_OK;        // someone else retired it already
    g_primes.erase (it);
    g_tables_by_device.clear ();          // live sessions keep their own reference
    return SLIPCU_OK;
}

// ------------------------------------------------------------------------------------------------
// process-wide cache of device allocations.  cudaMalloc costs tens of milliseconds per call on a
// 180 GB part, and a factorization session needs dozens of large blocks; sessions come and go
// (one per SLIP_solve_* call), so freed blocks are kept and handed to the next session instead of
// going back to the driver.  Blocks are rounded to power-of-two sizes >= 1 MB.
// ------------------------------------------------------------------------------------------------
static std::mutex g_pool_mutex;
typedef std::pair<int, size_t> PoolKey;                    // (device, size): a block only serves its own device
static std::multimap<PoolKey, void *> g_pool_free;         // -> block
static std::map<void *, PoolKey> g_pool_size;              // every live or cached block
static size_t g_pool_cached = 0;

static size_t pool_round (size_t bytes)
{
    size_t r = (size_t) 1 << 20;
    while (r < bytes) r <<= 1;
    return r;
}
static void pool_trim_locked ()
{   // cudaFree takes the block's device from the pointer: no device switch needed
    for (auto &kv : g_pool_free) { cudaFree (kv.second); g_pool_size.erase (kv.second); }
    g_pool_free.clear (); g_pool_cached = 0;
}
static cudaError_t pool_alloc (void **out, size_t bytes)
{
    const size_t want = pool_round (std::max<size_t> (bytes, 1));
    int dev = 0;
    { cudaError_t e0 = cudaGetDevice (&dev); if (e0 != cudaSuccess) return e0; }
    std::lock_guard<std::mutex> lk (g_pool_mutex);
    auto it = g_pool_free.find (PoolKey (dev, want));
    if (it != g_pool_free.end ())
    {
        *out = it->second; g_pool_cached -= want; g_pool_free.erase (it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc (out, want);
    if (e != cudaSuccess)
    {   // memory pressure: give the cached blocks back and try once more
        cudaGetLastError ();
        pool_trim_locked ();
        e = cudaMalloc (out, want);
    }
    if (e == cudaSuccess) g_pool_size[*out] = PoolKey (dev, want);
    return e;
}
static void pool_free (void *ptr)
{
    if (!ptr) return;
    std::lock_guard<std::mutex> lk (g_pool_mutex);
    auto it = g_pool_size.find (ptr);
    if (it == g_pool_size.end ()) { cudaFree (ptr); return; }
    g_pool_free.insert ({it->second, ptr});
    g_pool_cached += it->second.second;
    // cached blocks are bounded: beyond the cap (SLIP_B200_POOL_CAP_MB, default 64 GB, about a third
    // of the device) everything cached goes back to the driver
    static const size_t cap = (size_t) [] { const char *v = getenv ("SLIP_B200_POOL_CAP_MB"); return (v && *v) ? atoll (v) : 65536ll; } () << 20;
    if (g_pool_cached > cap) pool_trim_locked ();
}
// returns every cached device block and pinned host buffer to the driver (SLIP_finalize)
extern "C" void slipcu_release_cached_memory (void);
template <typename T> static cudaError_t pool_alloc_t (T **out, size_t bytes) { return pool_alloc ((void **) out, bytes); }

// pinned host buffers are expensive to create (cudaHostAlloc runs at ~0.3 s per GB, and both it and
// cudaFreeHost serialise against every other thread's work on the device): kept per process.
// Blocks of at least 4 KB, power-of-two sizes.
static std::mutex g_hpool_mutex;
static std::multimap<size_t, void *> g_hpool_free;
static std::map<void *, size_t> g_hpool_size;
static cudaError_t host_pool_alloc (void **out, size_t bytes)
{
    size_t want = 4096;
    while (want < bytes) want <<= 1;
    {
        std::lock_guard<std::mutex> lk (g_hpool_mutex);
        auto it = g_hpool_free.find (want);
        if (it != g_hpool_free.end ()) { *out = it->second; g_hpool_free.erase (it); return cudaSuccess; }
    }
    cudaError_t e = cudaHostAlloc (out, want, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lk (g_hpool_mutex); g_hpool_size[*out] = want; }
    return e;
}
static void host_pool_free (void *ptr)
{
    if (!ptr) return;
    std::lock_guard<std::mutex> lk (g_hpool_mutex);
    auto it = g_hpool_size.find (ptr);
    if (it == g_hpool_size.end ()) { cudaFreeHost (ptr); return; }
    g_hpool_free.insert ({it->second, ptr});
}

// ------------------------------------------------------------------------------------------------
// device memory arena (bump allocation in large chunks; columns never straddle a chunk)
// ------------------------------------------------------------------------------------------------
struct Arena
{
    std::vector<std::pair<char *, size_t>> chunks;     // base, bytes
    size_t chunk_bytes; char *cur = nullptr; size_t left = 0; size_t total = 0;
    explicit Arena (size_t cb) : chunk_bytes (cb) {}
    void *alloc (size_t bytes)
    {
        bytes = (bytes + 255) & ~(size_t) 255;
        if (bytes > left)
        {
            size_t cb = std::max (chunk_bytes, bytes);
            char *ptr = nullptr;
            if (pool_alloc ((void **) &ptr, cb) != cudaSuccess) { cudaGetLastError (); return nullptr; }
            chunks.push_back ({ptr, cb}); cur = ptr; left = cb; total += cb;
        }
        void *r = cur; cur += bytes; left -= bytes;
        return r;
    }
    ~Arena () { for (auto &c : chunks) pool_free (c.first); }
};

// ------------------------------------------------------------------------------------------------
// session
// ------------------------------------------------------------------------------------------------
struct ColDesc            // one finished column, as the kernels see it
{
    const u32 *base;      // residues, [S/CH][cnt][CH]
    const int32_t *rows;  // original row index of every slot
    int32_t cnt;          // slots in the column (U part first, then L part)
    int32_t nU;           // size of the U part (the pivot itself is an L-part slot)
    int32_t pivslot;      // slot of the pivot
    int32_t pad;
    const int32_t *mag;   // (bound mode) upper bounds of log2 |entry| in 1/64 bit, [cnt]; candidates hold measured values
};

struct StepInfo;
struct ChunkInfo;
struct TimedRange;

// per-launch symbolic workspace: the row -> slot map, slot lists, step and chunk tables of one
// column, with the stream they are built and consumed on.  The column in flight has one; every
// lookahead slot (bulk part of a later column, running beside it) has its own.
struct WorkCtx
{
    cudaStream_t st = nullptr;
    int32_t *pos = nullptr;                  // [n]
    int32_t *slots = nullptr; size_t slots_cap = 0;
    StepInfo *steps = nullptr; size_t steps_cap = 0;
    ChunkInfo *chunks = nullptr; size_t chunks_cap = 0;
    // pattern packets in mapped host memory (read by k_prep).  The host does not wait for every
    // column (a column with a single candidate has its pivot before the GPU has looked at it), so
    // the packet of the next column must not overwrite one that k_prep has not read yet: a ring
    // of PK_RING buffers, each guarded by an event recorded behind its k_prep.
    int32_t *h_packet = nullptr;             // PK_RING buffers of pk_stride ints
    int32_t *h_packet_dev = nullptr;         // device address of the ring
    size_t pk_stride = 0; int pk_next = 0, pk_cur = 0;
    cudaEvent_t pk_ev[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    bool pk_busy[8] = { false, false, false, false, false, false, false, false };
};
#define PK_RING 8
#define SPEC_SLOTS 16
struct SpecSlot                              // bulk part of a column that is not the current one yet
{
    WorkCtx w;
    u32 *buf = nullptr; size_t words = 0;    // [S/CH][cnt][CH] normalised vector
    int32_t *mag = nullptr; size_t mag_cap = 0;
    int32_t *rows = nullptr;                 // device copy of its pattern
    int cnt = 0, col = -1, nU = 0;           // col: factorization column it belongs to (-1: free)
    cudaEvent_t done = nullptr, consumed = nullptr;
    bool consumed_pending = false;
};

struct HostCol
{
    u32 *base = nullptr; int32_t *rows = nullptr; int cnt = 0, nU = 0, s = 0, stride = 0;
    u32 *limbs = nullptr; int32_t *nl = nullptr; int8_t *sign = nullptr;
    int32_t *mag = nullptr;
};

struct slipcu_factor
{
    int n = 0, nz = 0, S = 0, CH = 16, threads = 256, cpt = 4, garner_e = 0, device = 0;
    std::shared_ptr<Tables> tab;
    cudaStream_t st = nullptr;
    cudaEvent_t ev = nullptr, ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_start = nullptr, ev_end = nullptr;      // A resident ... solution numerators on device
    int32_t *dAp = nullptr, *dAi = nullptr;
    u32 *dA = nullptr;                       // residues of A, [S/CH][nz][CH]
    u32 *rho = nullptr, *invrho = nullptr;    // [n][S]
    ColDesc *desc = nullptr;                 // [n]
    int32_t *bad = nullptr;                  // device flag
    u32 *dig = nullptr; size_t dig_rows = 0; // [dig_rows][S] digit scratch in use (= digbuf[i])
    int32_t *topd = nullptr;                 // [dig_rows]
    // Sessions that keep positional factors reconstruct the U part and convert the whole column
    // to limbs on a second stream, behind the pivot search and the next column's elimination:
    // the digit scratch is then double-buffered and ev_side[i] guards the reuse of buffer i.
    cudaStream_t st2 = nullptr, wst = nullptr;          // side stream; stream run_garner/run_limbs launch on
    cudaEvent_t ev_tri = nullptr, ev_gl = nullptr, ev_side[2] = { nullptr, nullptr };
    bool side_pending[2] = { false, false };
    u32 *digbuf[2] = { nullptr, nullptr };
    int32_t *topdbuf[2] = { nullptr, nullptr };
    int overlap = 0, seq = 0;
    // approximate pivot search (k_fraccrt) of sessions that do not keep positional factors
    int frac = 0, fracW = 8, frac_col = -1;             // enabled, words for the next column, column searched that way
    int frac_margin = 12;                               // words kept beyond the leading zero words of the last winner
    int frac_verify = 0;                                // (tests) re-run every accepted choice through the exact scan
    // lookahead: bulk parts (every step with a pivot that is already committed) of the next columns,
    // each on its own stream beside the column in flight
    SpecSlot spec[SPEC_SLOTS];
    cudaEvent_t ev_commit = nullptr;                    // recorded after every pivot commit
    struct { int cnt, nU, s, mode, diag_slot, W; } fq = { 0, 0, 0, 0, 0, 0 };
    struct FracKey *frackey = nullptr; size_t frac_rows = 0;
    uint64_t frac_cols = 0, frac_retries = 0, frac_fallbacks = 0;
    Arena resid, ints, limbs;
    std::vector<HostCol> cols;
    std::vector<int32_t> hAp;
    WorkCtx mc;                              // workspace of the column in flight (stream = st)
    slipcu_pivot_info *h_info = nullptr;     // pinned
    slipcu_pivot_info *d_info = nullptr;
    size_t smem_limit = 0;
    int keep_positional = 1, rows_are_positions = 0, x_global = 0, cur = -1, cur_launch = 0;
    int slot_sort = 0;                  // k_slots hands rows to threads by bank group (see there)
    std::vector<struct TimedRange> ranges;      // profiling: event pairs not yet read
    u32 *tmp_limbs = nullptr; int32_t *tmp_nl = nullptr; int tmp_stride = 0;
    int sms = 148;
    int garner_mode = 2;
    // bound mode (see tri_mag_cta): fewer channels than the Hadamard bound, every column's size proven
    unsigned *done_ctr = nullptr;            // device counter of the fused reconstruction + scan launches
    int32_t *run_flags = nullptr;            // device: [0] first column (k+1) without a nonzero candidate, [1] largest measured size so far
    struct { int k, slot; } pending_commit = { -1, -1 };     // pivot chosen, commit folded into the next column's first kernel
    int frac_min_s = 64;                     // approximate pivot search only from this many channels on
    int32_t info_seq = 0;                    // sequence number of the last scan / selection launched
    bool counted = false;                    // this session is in g_live_sessions
    bool timed = false;                      // its device time has been added to g_device_ms (slipcu_solve)
    int nowait_singles = 0;                  // the caller does not wait for single-candidate columns (slipcu_factor_nowait_singles)
    int mag_on = 0;
    int measured = 0;                        // measured mode: sizes of the candidates are measured, the result is verified exactly by the caller
    int32_t *Amag = nullptr;                 // [nz] 64 log2 |a| of the input entries, rounded up
    int32_t *rho_mag = nullptr;              // [n]  measured 64 log2 |rho_k|
    int32_t *bound = nullptr;                // device: largest bound of the column in flight
    slipcu_factor () : resid ((size_t) 512 << 20), ints ((size_t) 16 << 20), limbs ((size_t) 256 << 20) {}
};

// timing helper for the optional profiling mode: CUDA events on the session's stream around a
// launch.  Nothing blocks: the event pairs are queued and read after the next stream
// synchronisation (flush_timers), so profiling does not disturb the overlap of host and device.
struct TimedRange { cudaEvent_t a, b; double *acc; };
static std::vector<cudaEvent_t> g_event_pool;
static std::mutex g_event_mutex;
// Launches of k_trisolve may overlap (lookahead streams): the sum of their durations counts shared
// time twice, so the intervals themselves are kept, relative to a base event, and the bench reads
// the length of their union (time during which at least one k_trisolve was running).
static cudaEvent_t g_base_ev = nullptr;
static std::vector<std::pair<float, float>> g_tri_intervals;
static cudaEvent_t take_event ()
{
    {
        std::lock_guard<std::mutex> lk (g_event_mutex);
        if (!g_event_pool.empty ()) { cudaEvent_t e = g_event_pool.back (); g_event_pool.pop_back (); return e; }
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate (&e);
    return e;
}
// Events and streams of finished sessions are kept too (creating and destroying them takes a
// driver-wide lock: eight sessions starting and ending on eight host threads queue up behind it).
static std::mutex g_obj_mutex;
static std::multimap<int, cudaEvent_t> g_ev_free;          // kind = 2 * device + timing
static std::map<cudaEvent_t, int> g_ev_kind;
static std::multimap<int, cudaStream_t> g_st_free;         // device
static std::map<cudaStream_t, int> g_st_dev;
static cudaError_t pooled_event (cudaEvent_t *out, bool timing)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice (&dev);
    if (e != cudaSuccess) return e;
    const int kind = 2 * dev + (timing ? 1 : 0);
    {
        std::lock_guard<std::mutex> lk (g_obj_mutex);
        auto it = g_ev_free.find (kind);
        if (it != g_ev_free.end ()) { *out = it->second; g_ev_free.erase (it); return cudaSuccess; }
    }
    e = timing ? cudaEventCreate (out) : cudaEventCreateWithFlags (out, cudaEventDisableTiming);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lk (g_obj_mutex); g_ev_kind[*out] = kind; }
    return e;
}
static void release_event (cudaEvent_t ev)
{
    if (!ev) return;
    std::lock_guard<std::mutex> lk (g_obj_mutex);
    auto it = g_ev_kind.find (ev);
    if (it == g_ev_kind.end ()) { cudaEventDestroy (ev); return; }
    g_ev_free.insert ({it->second, ev});
}
static cudaError_t pooled_stream (cudaStream_t *out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice (&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lk (g_obj_mutex);
        auto it = g_st_free.find (dev);
        if (it != g_st_free.end ()) { *out = it->second; g_st_free.erase (it); return cudaSuccess; }
    }
    e = cudaStreamCreateWithFlags (out, cudaStreamNonBlocking);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lk (g_obj_mutex); g_st_dev[*out] = dev; }
    return e;
}
static void release_stream (cudaStream_t st)
{   // the caller has synchronised it
    if (!st) return;
    std::lock_guard<std::mutex> lk (g_obj_mutex);
    auto it = g_st_dev.find (st);
    if (it == g_st_dev.end ()) { cudaStreamDestroy (st); return; }
    g_st_free.insert ({it->second, st});
}

struct ScopedTimer
{
    slipcu_factor *F; double *acc; cudaEvent_t a = nullptr; cudaStream_t st;
    ScopedTimer (slipcu_factor *F_, double *acc_, cudaStream_t st_ = nullptr) : F (F_), acc (acc_), st (st_ ? st_ : F_->wst)
    {
        if (g_profiling) { a = take_event (); cudaEventRecord (a, st); }
    }
    ~ScopedTimer ()
    {
        if (!a) return;
        cudaEvent_t b = take_event ();
        cudaEventRecord (b, st);
        F->ranges.push_back ({a, b, acc});
    }
};
// call after the stream has been synchronised
static void flush_timers (slipcu_factor *F)
{
    if (F->ranges.empty ()) return;
    std::lock_guard<std::mutex> lk (g_event_mutex);
    std::vector<TimedRange> later;
    for (auto &r : F->ranges)
    {
        if (cudaEventQuery (r.b) == cudaErrorNotReady) { later.push_back (r); continue; }   // side stream still busy
        float ms = 0;
        if (cudaEventElapsedTime (&ms, r.a, r.b) == cudaSuccess) *r.acc += ms; else cudaGetLastError ();
        if (r.acc == &g_tri_ms)
        {
            if (!g_base_ev) { g_base_ev = r.a; r.a = nullptr; }          // first launch after a reset: kept as the base
            float t0 = 0, t1 = 0;
            if (cudaEventElapsedTime (&t0, g_base_ev, r.a ? r.a : g_base_ev) == cudaSuccess
                && cudaEventElapsedTime (&t1, g_base_ev, r.b) == cudaSuccess) g_tri_intervals.push_back ({t0, t1});
            else cudaGetLastError ();
        }
        if (r.a) g_event_pool.push_back (r.a);
        g_event_pool.push_back (r.b);
    }
    F->ranges.swap (later);
}

// ------------------------------------------------------------------------------------------------
// k_residues: positional limb strings -> Montgomery residues, channel-blocked
//   out[(cb * count + e) * CH + ch]   e = entry, c = cb*CH + ch
// ------------------------------------------------------------------------------------------------
__global__ void k_residues (int count, int CH, const u32 *limbs, const int64_t *off, const int8_t *sign,
                            const u32 *p, const u32 *ninv, const u32 *r2, u32 *out)
{
    const int per = blockDim.x / CH;
    const int e = blockIdx.x * per + threadIdx.x / CH;
    const int ch = threadIdx.x % CH;
    const int cb = blockIdx.y;
    if (e >= count) return;
    const int c = cb * CH + ch;
    const u32 pc = p[c], ni = ninv[c], rr = r2[c];
    u32 r = 0;
    for (int64_t l = off[e + 1] - 1; l >= off[e]; --l)
    {
        r = mont_mul (r, rr, pc, ni);                 // r * 2^32 mod p
        r += reduce_word (limbs[l], pc);
        if (r >= pc) r -= pc;
    }
    r = mont_mul (r, rr, pc, ni);                     // to Montgomery form
    if (sign[e] < 0 && r) r = pc - r;
    out[((size_t) cb * count + e) * CH + ch] = r;
}

// ------------------------------------------------------------------------------------------------
// shared-memory address helper for the cp.async (LDGSTS) copies
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32 (const void *p) { return (u32) __cvta_generic_to_shared (p); }

// ------------------------------------------------------------------------------------------------
// symbolic pre-pass of a column: row -> slot map, then for every elimination step the list of
// target slots (one per entry of the L column used), shared by all channel blocks.
// ------------------------------------------------------------------------------------------------
#define TRI_THREADS 256       // threads of a k_trisolve CTA when a thread owns 4 channels
// Shared-memory chunk buffers and chunk descriptors kept in shared memory.  With thousands of
// channels there are two CTAs per SM and hundreds per launch: three buffers (two chunks in flight
// while one is consumed) keep HBM busy.  Sessions of few channels (4-channel blocks: LP bases whose
// values need a few hundred bits) have a few dozen CTAs on 148 SMs, each alone on its SM and
// paced by memory latency, and their steps are short (a chunk is often 2-4 KB): eight buffers put
// seven chunks in flight per CTA.
static __host__ __device__ constexpr int tri_bufs (int CH) { return CH == 4 ? 8 : 3; }
static __host__ __device__ constexpr int tri_ring (int CH) { return CH == 4 ? 16 : 8; }

// Every thread of k_trisolve owns 4 channels, so CH/4 threads share a row and a CTA covers
// TRI_THREADS/(CH/4) rows at a time; a pipeline chunk is four such row groups (16 KB of L).
static __host__ __device__ inline int tri_row_groups (int CH) { return TRI_THREADS / (CH / 4); }
static __host__ __device__ inline int tri_chunk_rows (int CH) { return 4 * tri_row_groups (CH); }
// A row of the work vector is CH*4 bytes, so 32/CH consecutive row groups (the lanes of a
// quarter-warp) make one 128-byte wavefront of a 16-byte-per-lane shared-memory access: the number
// of 'bank groups' a row can fall into.  The slot lists hand the rows of a chunk to the threads in
// units of that many row groups (see k_slots).
static __host__ __device__ constexpr int tri_bank_groups (int CH) { return 32 / CH; }
// row groups of a chunk of nrows rows whose threads have anything to do
static __host__ __device__ inline int tri_active_groups (int nrows, int CH)
{
    const int G = tri_bank_groups (CH), RG = tri_row_groups (CH), up = (nrows + G - 1) / G * G;
    return up < RG ? up : RG;
}
// entries of the slot list of a step with len rows: full chunks, then 4 entries per thread in use
static __host__ __device__ inline int tri_slot_extent (int len, int CH)
{
    const int R = tri_chunk_rows (CH), rem = len % R;
    return (len / R) * R + 4 * tri_active_groups (rem, CH);
}
// A slot-list entry: target slot of the work vector (the spare row `cnt` when the row has no
// target) in the high 22 bits, the row of the chunk whose L entry goes there in the low 10
// (one shift gives the slot, one mask the row).
#define TRI_SLOT_BITS 22
#define TRI_SLOT_MASK ((1u << TRI_SLOT_BITS) - 1u)
#define TRI_ROW_BITS 10
#define TRI_ROW_MASK ((1u << TRI_ROW_BITS) - 1u)

struct StepInfo           // one elimination step of a column: eliminate with column j of L
{
    const u32 *lbase;     // first L-part residue of column j for channel block 0
    int32_t j;            // pivot position
    int32_t len;          // entries in the L part of column j (the pivot row included)
    int32_t cbstride;     // words between channel blocks of column j
    int32_t slot_off;     // offset of this step's slot list (multiple of 4)
    int32_t chunk0;       // index of the step's first pipeline chunk
    int32_t pad;
};

struct __align__ (16) ChunkInfo   // one pipeline chunk (rows of one step); 32 bytes, read as two 16-byte words
{
    const u32 *lsrc;      // first L residue of the chunk for channel block 0
    int32_t cbstride;     // words between channel blocks
    int32_t slot_off;     // offset of the chunk's slot list
    int32_t j;            // pivot position of the step
    int32_t meta;         // rows | first chunk of its step << 16 | last chunk << 17
    const int32_t *msrc;  // (bound mode) magnitudes of the chunk's L rows
};

__global__ void k_setpos (int cnt, const int32_t *rows, int32_t *pos)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < cnt) pos[rows[t]] = t;
}

// Slot lists, one CTA per pipeline chunk.  Entry 4*g + q of a chunk's list belongs to the thread
// of row group g (one 16-byte load gives a thread its four entries) and names BOTH the row of the
// chunk it takes its L entry from and the slot of the work vector that entry updates (packed, see
// TRI_SLOT_BITS); rows without a target (past the end of the chunk, or the pivot row itself) point
// at the spare row `cnt` of the vector and are skipped by the consumer.
//
// Which row goes to which thread is free, and it decides the shared-memory bank conflicts of
// k_trisolve: the G = 32/CH row groups of a quarter-warp read G rows of L and read-modify-write G
// rows of w in one access each, conflict-free iff the G rows fall into G distinct bank groups
// (row index mod G for L in the stage buffer, slot mod G for w).  With sort == 0 thread g takes
// rows g + q*RG (consecutive rows of L, targets wherever the pattern puts them: about two
// wavefronts per access on scattered patterns).  With sort == 1 warp a of this CTA owns the rows
// r = a (mod G) of the chunk (always 128 of them) and orders them by (slot - a) mod G, stable;
// the i-th row of every class goes to the same quarter-warp access (q = i / 32, g = (i mod 32) G + a):
// the L rows of an access are one of each class, and so are the targets wherever the four classes
// are in the same stretch of their order (all of them, up to the imbalance of the counts).
// upart = 1: the steps use the U parts of the columns (slots 0..nU-1, no pivot among them): the back
// substitution, z_i -= U_ij z_j for the rows above the diagonal of column j.
__global__ void __launch_bounds__ (256) k_slots (int nU, int CH, int cnt, const int32_t *upos, const int32_t *uoff,
                         const int32_t *uchunk, const ColDesc *desc, const int32_t *pos, int32_t *slots,
                         StepInfo *steps, ChunkInfo *chunks, int upart, int sort)
{
    const int ci = blockIdx.x, lane = threadIdx.x & 31, a = threadIdx.x >> 5;    // chunk, lane, class
    const int G = tri_bank_groups (CH), R = tri_chunk_rows (CH);
    int lo = 0, hi = nU - 1;                      // last u with uchunk[u] <= ci: the step of the chunk
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (uchunk[mid] <= ci) lo = mid; else hi = mid - 1; }
    const int u = lo;
    const ColDesc d = desc[upos[u]];
    const int first = upart ? 0 : d.nU;           // first slot of the part the step streams
    const int len = upart ? d.nU : d.cnt - d.nU;
    const int r0 = (ci - uchunk[u]) * R;          // first row of the chunk within the step
    const int nrows = min (R, len - r0);
    const int active = (nrows == R) ? tri_row_groups (CH) : tri_active_groups (nrows, CH);
    const int npass = (((nrows + G - 1) / G) + 31) >> 5;        // passes of 32 rows of a class that hold rows at all
    const unsigned lt = (1u << lane) - 1u;
    int slot[4], key[4], idx[4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
    {
        const int r = a + G * (p * 32 + lane);
        const int m = first + r0 + r;
        const bool target = r < nrows && (upart || m != d.pivslot);
        slot[p] = target ? pos[d.rows[m]] : cnt;
        key[p] = target ? (sort ? ((slot[p] - a) & (G - 1)) : 0) : G;
        idx[p] = p * 32 + lane;                   // passes without rows keep their place at the end
    }
    if (sort || npass < 4)
    {   // stable order by key: targets first (by bank group relative to the class), rows without one last
        int run = 0;
        for (int k = 0; k <= G; ++k)
        {
            if (!sort && k > 0 && k < G) continue;
#pragma unroll
            for (int p = 0; p < 4; ++p)
            {
                if (p >= npass) continue;
                const unsigned b = __ballot_sync (0xffffffffu, key[p] == k);
                if (key[p] == k) idx[p] = run + __popc (b & lt);
                run += __popc (b);
            }
        }
    }
    int32_t *out = slots + (size_t) uoff[u] + r0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
    {
        const int g = (idx[p] & 31) * G + a, q = idx[p] >> 5;
        const int r = a + G * (p * 32 + lane);
        if (g < active) out[4 * g + q] = (int32_t) (((u32) slot[p] << TRI_ROW_BITS) | (u32) r);
    }
    if (threadIdx.x == 0)
    {
        if (r0 == 0)
        {
            StepInfo si;
            si.lbase = d.base + (size_t) first * CH; si.j = upos[u]; si.len = len; si.cbstride = d.cnt * CH;
            si.slot_off = uoff[u]; si.chunk0 = uchunk[u]; si.pad = 0;
            steps[u] = si;
        }
        ChunkInfo c;
        c.lsrc = d.base + (size_t) (first + r0) * CH; c.cbstride = d.cnt * CH;
        c.slot_off = uoff[u] + r0; c.j = upos[u];
        c.meta = nrows | ((r0 == 0) << 16) | ((r0 + R >= len) << 17);
        c.msrc = d.mag ? d.mag + first + r0 : nullptr;
        chunks[ci] = c;
    }
}

// ------------------------------------------------------------------------------------------------
// k_trisolve: sparse REF triangular solve of one column (or of one dense right-hand side).
// One CTA per block of CH channels; channels are independent.
//
// The vector is held in shared memory in NORMALISED form  w_t = x_t / rho_{h_t}
// (h_t = last elimination step applied to x_t, rho_{-1} = 1).  In that form the REF update
//     x_t <- (rho_j * rho_{j-1}/rho_{h_t} * x_t - l_tj * x_j) / rho_{j-1}
// of slip_REF_triangular_solve.c:150-232, including every history update, collapses to
//     w_t <- w_t - l_tj * yhat_j ,   yhat_j = w_j / rho_j
// because division by a pivot is a multiplication by its inverse modulo the channel prime: the
// history vector of the reference is not needed at all.  The true REF values are restored when
// the column is published:  U(j,k) = w_j * rho_{j-1},  L(t,k) = w_t * rho_{k-1}.
// Slots 0..nU-1 of the pattern are rows that are already pivotal (in pivot order); slots
// nU..cnt-1 are the candidate rows; row cnt is a spare that absorbs updates without a target.
// ------------------------------------------------------------------------------------------------

struct TriArgs
{
    int k;                   // level the L part is brought to (column index; n for a rhs)
    int S, cnt, nU;
    const int32_t *rows;     // [cnt] original row of each slot
    const StepInfo *steps;   // [nU]
    const ChunkInfo *chunks; // [nchunks]
    const int32_t *slots;    // slot lists
    // source vector to scatter: src_cnt entries, entry e at residue index src_first + e*src_step,
    // going to the slot of row src_rows[e] (or row e when src_rows == nullptr)
    const u32 *src; size_t src_total; int64_t src_first; int src_step, src_cnt; const int32_t *src_rows;
    size_t src_y_stride;     // added to src_first per blockIdx.y (multiple right-hand sides)
    u32 *out;                // result: channel block cb of right-hand side y at out + cb * out_cb_stride + y * out_y_stride
    size_t out_y_stride, out_cb_stride;      // (one column: out_cb_stride = cnt * CH)
    const u32 *rho, *invrho, *p, *ninv;      // [n][S] pivots and their inverses (Montgomery form)
    const int32_t *pos;      // [n] row -> slot
    int x_in_smem;
    int nchunks;             // total pipeline chunks of this launch
    const int32_t *upos;     // [nU] pivot position of each U slot (levels of the published U part)
    int publish;             // 1: write true REF values; 0: leave the vector normalised (speculative first part)
    int u0;                  // U slot of the first step in the chunk list
    // bound mode (sessions that carry fewer channels than the Hadamard bound asks for): one extra
    // CTA (blockIdx.x == S/CH) propagates upper bounds of log2 |w_t| through the same work list
    int mag_on;
    const int32_t *mag_src;  // magnitudes of the source entries (entry e of the scatter above)
    int32_t *mag_out;        // [cnt] bounds of the published entries (or of the normalised vector)
    const int32_t *rho_mag;  // [n] measured log2 |rho_k| (upper bound; the lower bound is MAG_GAP less)
    int32_t *bound_out;      // largest published bound of this launch
    int smem_bytes;          // dynamic shared memory of the launch
    int rhs_fastest;         // 1: blockIdx.x = right-hand side, blockIdx.y = channel block (see slipcu_solve)
    // back substitution (slots are positions): the slot of a step's pivot is its position j from the
    // chunk descriptor (steps run n-1..0 and columns without a U part have no chunk at all), the
    // source vector is multiplied by src_scale (det) on the way in, and the result by 1/rho_t on
    // the way out (publish == 2):  z = det y;  for j = n-1..0: x_j = z_j / rho_j, z_i -= U_ij x_j
    int j_is_slot;
    const u32 *src_scale;    // [S] or nullptr
};

template <int CH>
struct TriSmem
{
    static constexpr int RG = TRI_THREADS / (CH / 4);            // row groups
    static constexpr int R = 4 * RG;                             // rows per chunk
    static constexpr int L_BYTES = R * CH * 4;                   // 16 KB
    static constexpr int S_BYTES = R * 4;
    static constexpr int STAGE = (L_BYTES + S_BYTES + CH * 4 + 127) & ~127;   // L rows, targets, 1/rho_j
    static constexpr int BUFS = tri_bufs (CH), RING = tri_ring (CH);
    static constexpr int RING_BYTES = RING * (int) sizeof (ChunkInfo);
    static __host__ __device__ size_t x_bytes (int cnt) { return ((size_t) (cnt + 1) * CH * 4 + 127) & ~(size_t) 127; }
    static __host__ __device__ size_t total (int cnt, bool x_in_smem)
    {
        return (x_in_smem ? x_bytes (cnt) : (size_t) 0) + RING_BYTES + (size_t) BUFS * STAGE;
    }
};

__device__ __forceinline__ void cp_async16 (u32 dst_smem, const void *src)
{
    asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst_smem), "l"(src) : "memory");
}
template <int OFF> __device__ __forceinline__ void cp_async16_off (u32 dst_smem, const void *src)
{   // same displacement on both sides, folded into the instruction
    asm volatile ("cp.async.cg.shared.global [%0+%2], [%1+%2], 16;" :: "r"(dst_smem), "l"(src), "n"(OFF) : "memory");
}
__device__ __forceinline__ void cp_async16_if (bool on, u32 dst_smem, const void *src)
{   // predicated in place: a branch would start a new basic block (and ptxas pads the first LDGSTS of a block)
    asm volatile ("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p cp.async.cg.shared.global [%0], [%1], 16;\n}"
                  :: "r"(dst_smem), "l"(src), "r"((u32) on) : "memory");
}
template <int OFF> __device__ __forceinline__ void cp_async16_off_if (bool on, u32 dst_smem, const void *src)
{
    asm volatile ("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p cp.async.cg.shared.global [%0+%2], [%1+%2], 16;\n}"
                  :: "r"(dst_smem), "l"(src), "n"(OFF), "r"((u32) on) : "memory");
}
__device__ __forceinline__ void cp_async_commit () { asm volatile ("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait () { asm volatile ("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128 (u32 addr)
{
    uint4 v;
    asm volatile ("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64 (u32 addr)
{
    uint2 v;
    asm volatile ("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// CPT (2 or 4) consecutive channels of one row, as one vector access
template <int CPT> struct ChanVec { u32 v[CPT]; };
template <int CPT> __device__ __forceinline__ ChanVec<CPT> ldv (const void *p)
{
    ChanVec<CPT> r;
    if (CPT == 4) { const uint4 t = *reinterpret_cast<const uint4 *> (p); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
    else { const uint2 t = *reinterpret_cast<const uint2 *> (p); r.v[0] = t.x; r.v[1] = t.y; }
    return r;
}
template <int CPT> __device__ __forceinline__ void stv (void *p, const ChanVec<CPT> &r)
{
    if (CPT == 4) *reinterpret_cast<uint4 *> (p) = make_uint4 (r.v[0], r.v[1], r.v[2], r.v[3]);
    else *reinterpret_cast<uint2 *> (p) = make_uint2 (r.v[0], r.v[1]);
}
template <int CPT> __device__ __forceinline__ ChanVec<CPT> ldsv (u32 addr)
{
    ChanVec<CPT> r;
    if (CPT == 4) { const uint4 t = lds128 (addr); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
    else { const uint2 t = lds64 (addr); r.v[0] = t.x; r.v[1] = t.y; }
    return r;
}
template <int CPT> __device__ __forceinline__ ChanVec<CPT> mont_mulv (const ChanVec<CPT> &a, const ChanVec<CPT> &b, const ChanVec<CPT> &p, const ChanVec<CPT> &ni)
{
    ChanVec<CPT> r;
#pragma unroll
    for (int i = 0; i < CPT; ++i) r.v[i] = mont_mul (a.v[i], b.v[i], p.v[i], ni.v[i]);
    return r;
}
// w + l*ny  (ny = -yhat): one Montgomery product and a modular add per channel
template <int CPT> __device__ __forceinline__ ChanVec<CPT> sub_mulv (const ChanVec<CPT> &w, const ChanVec<CPT> &l, const ChanVec<CPT> &ny, const ChanVec<CPT> &p, const ChanVec<CPT> &ni)
{
    ChanVec<CPT> r;
#pragma unroll
    for (int i = 0; i < CPT; ++i) r.v[i] = add_mod (w.v[i], mont_mul (l.v[i], ny.v[i], p.v[i], ni.v[i]), p.v[i]);
    return r;
}
// Shoup multiplication by a multiplier that stays fixed over many products (the -yhat_j of an
// elimination step): with the companion yq = floor (y 2^32 / p), q = hi32 (l yq) and
// r = l y - q p lies in [0, 2p), computed in 32-bit wrap-around arithmetic.  One IMAD.HI and two
// IMAD per product where the Montgomery product needs IMAD.WIDE + IMAD + IMAD.HI; IMAD.WIDE and
// IMAD.HI issue at less than half the rate of IMAD (46 / 55 / 113 per SM cycle, measured), and the
// integer multiplies are what bounds k_trisolve.  y is a PLAIN residue; l in Montgomery form gives
// the product in Montgomery form.
template <int CPT> __device__ __forceinline__ ChanVec<CPT> sub_mul_shoup (const ChanVec<CPT> &w, const ChanVec<CPT> &l, const ChanVec<CPT> &ny, const ChanVec<CPT> &nyq, const ChanVec<CPT> &p)
{   // w + l * ny mod p  (ny = -yhat, plain)
    ChanVec<CPT> r;
#pragma unroll
    for (int i = 0; i < CPT; ++i)
    {
        const u32 q = __umulhi (l.v[i], nyq.v[i]);
        const u32 t = l.v[i] * ny.v[i] - q * p.v[i];           // in [0, 2p)
        r.v[i] = csub (w.v[i] + csub (t, p.v[i]), p.v[i]);
    }
    return r;
}
template <int CPT> __device__ __forceinline__ ChanVec<CPT> negv (const ChanVec<CPT> &y, const ChanVec<CPT> &p)
{
    ChanVec<CPT> r;
#pragma unroll
    for (int i = 0; i < CPT; ++i) r.v[i] = y.v[i] ? p.v[i] - y.v[i] : 0u;
    return r;
}


// ------------------------------------------------------------------------------------------------
// Bound mode: a rigorous size bound for every entry of a column, computed alongside the residues.
//
// A session may carry fewer channels than the Hadamard bound of the input asks for (real LP bases
// have determinants of a few hundred bits against Hadamard bounds of tens of thousands).  What
// makes that exact rather than a guess is this pass: with mb_t >= log2 |w_t| for the normalised
// vector, the update  w_t <- w_t - l_tj * w_j / rho_j  gives
//     mb_t <- log2 (2^mb_t + 2^(lmag_tj + mb_j - lb(rho_j)))          (rounded up)
// where lmag and the pivot sizes are MEASURED on the finished columns (from their mixed-radix
// digits).  By induction over the columns: if the bound of column k is below the capacity of the
// channels, its reconstruction is exact, so its measured magnitudes are the true ones, so the bound
// of column k+1 holds.  A column whose bound does not fit aborts the factorization, which restarts
// with more channels.  Magnitudes are int32 in units of 1/64 bit, MAG_NEG = the entry is zero.
// ------------------------------------------------------------------------------------------------
#define MAG_UNIT 64
#define MAG_NEG (-(1 << 30))
#define MAG_GAP 96            // a measured magnitude m means  m - MAG_GAP <= 64 log2 |v| <= m

__device__ __forceinline__ int32_t mag_lse (int32_t a, int32_t b)
{   // upper bound of 64 log2 (2^(a/64) + 2^(b/64))
    if (a == MAG_NEG) return b;
    if (b == MAG_NEG) return a;
    const int32_t hi = max (a, b), d = hi - min (a, b);
    if (d >= 40 * MAG_UNIT) return hi + 1;
    return hi + __float2int_ru (64.0f * log2f (1.0f + exp2f (-(float) d * (1.0f / 64.0f)))) + 1;
}
__device__ __forceinline__ int32_t mag_log2_ub (u32 v)
{   // upper bound of 64 log2 v for v >= 1
    return __float2int_ru (64.0f * log2f (__uint2float_ru (v))) + 1;
}
__device__ __forceinline__ void cp_async4 (u32 dst_smem, const void *src)
{
    asm volatile ("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst_smem), "l"(src) : "memory");
}

// The bound CTA of a k_trisolve launch.  Same chunk descriptors, same slot lists (every entry names
// its chunk row and its target slot, see k_slots) and the same three-stage cp.async ring, but a
// row is one int32 instead of CH residues.
template <int CH, int NT>
__device__ __noinline__ void tri_mag_cta (const TriArgs &a, unsigned char *smem_raw)
{
    constexpr int RG = TRI_THREADS / (CH / 4), R = 4 * RG;
    constexpr int MSTAGE = R * 4 + R * 4 + 16;               // targets, magnitudes, rho_mag[j]
    constexpr int TRI_BUFS = tri_bufs (CH), TRI_RING = tri_ring (CH);
    constexpr int RING_BYTES = TRI_RING * (int) sizeof (ChunkInfo);
    const int tid = threadIdx.x, cnt = a.cnt, nU = a.nU, nchunks = a.nchunks;
    const bool in_smem = (size_t) (cnt + 1) * 4 + 16 + RING_BYTES + TRI_BUFS * MSTAGE <= (size_t) a.smem_bytes;
    const u32 smem0 = smem_u32 (smem_raw);
    const u32 vec_bytes = in_smem ? (u32) ((((size_t) (cnt + 1) * 4) + 15) & ~(size_t) 15) : 0u;
    int32_t *mb = in_smem ? (int32_t *) smem_raw : a.mag_out;      // working vector (global when the pattern is too long)
    const u32 ring = smem0 + vec_bytes;
    const u32 stage0 = ring + RING_BYTES;

    auto fetch_desc = [&] (int X, int half)
    {
        cp_async16 (ring + (u32) (X % TRI_RING) * (u32) sizeof (ChunkInfo) + half * 16,
                    (const unsigned char *) (a.chunks + X) + half * 16);
    };
    auto issue = [&] (int rs, u32 sb)
    {
        const u32 da = ring + (u32) rs * (u32) sizeof (ChunkInfo);
        const uint4 d0 = lds128 (da);                  // lsrc (lo, hi), cbstride, slot_off
        const uint4 d1 = lds128 (da + 16);             // j, meta, msrc (lo, hi)
        const int nrows = (int) (d1.y & 0xffffu);
        const int32_t *msrc = (const int32_t *) (((unsigned long long) d1.w << 32) | d1.z);
        if (tid < tri_active_groups (nrows, CH)) cp_async16 (sb + (u32) tid * 16, a.slots + (size_t) d0.w + (size_t) tid * 4);
        for (int r = tid; r < nrows; r += NT) cp_async4 (sb + R * 4 + (u32) r * 4, msrc + r);
        if ((d1.y & 0x10000u) && tid == 0) cp_async4 (sb + 2 * R * 4, a.rho_mag + (int) d1.x);
    };

    // descriptors 0 .. 2 BUFS - 3: chunk c + BUFS - 1 is issued at iteration c, when only the copy
    // groups up to iteration c - BUFS + 1 are known to have landed
    if (tid < 2 * (2 * TRI_BUFS - 2) && (tid >> 1) < nchunks) fetch_desc (tid >> 1, tid & 1);
    cp_async_commit ();
    cp_async_wait<0> ();
    __syncthreads ();
    for (int c = 0; c < TRI_BUFS - 1; ++c) { if (c < nchunks) issue (c, stage0 + (u32) c * MSTAGE); cp_async_commit (); }
    for (int i = tid; i <= (in_smem ? cnt : cnt - 1); i += NT) mb[i] = MAG_NEG;
    __syncthreads ();
    for (int e = tid; e < a.src_cnt; e += NT)
    {
        const int row = a.src_rows ? a.src_rows[e] : e;
        mb[a.pos[row]] = a.mag_src[e];
    }

    int u = a.u0, bc = 0, bi = TRI_BUFS - 1;
    int32_t y = MAG_NEG;
    for (int c = 0; c < nchunks; ++c)
    {
        cp_async_wait<TRI_BUFS - 2> ();
        __syncthreads ();
        if (c + TRI_BUFS - 1 < nchunks) issue ((c + TRI_BUFS - 1) % TRI_RING, stage0 + (u32) bi * MSTAGE);
        if (tid < 2 && c + 2 * TRI_BUFS - 2 < nchunks) fetch_desc (c + 2 * TRI_BUFS - 2, tid);
        cp_async_commit ();
        const unsigned char *sbp = smem_raw + vec_bytes + RING_BYTES + (size_t) bc * MSTAGE;
        const u32 meta = lds64 (ring + (u32) (c % TRI_RING) * (u32) sizeof (ChunkInfo) + 16).y;
        const int nrows = (int) (meta & 0xffffu);
        if (meta & 0x10000u)
        {   // yhat_j = w_j / rho_j with |rho_j| >= 2^((rho_mag - MAG_GAP)/64)
            const int32_t wj = mb[u];
            const int32_t rm = *reinterpret_cast<const int32_t *> (sbp + 2 * R * 4);
            y = (wj == MAG_NEG) ? MAG_NEG : wj - (rm - MAG_GAP);
        }
        if (y != MAG_NEG)
        {
            const int ext = (nrows == R) ? R : 4 * tri_active_groups (nrows, CH);
            for (int e = tid; e < ext; e += NT)
            {
                const u32 t = reinterpret_cast<const u32 *> (sbp)[e];
                const int slot = (int) (t >> TRI_ROW_BITS), r = (int) (t & TRI_ROW_MASK);
                if (slot != cnt && r < nrows)
                {
                    const int32_t lm = reinterpret_cast<const int32_t *> (sbp + R * 4)[r];
                    if (lm != MAG_NEG) mb[slot] = mag_lse (mb[slot], lm + y);
                }
            }
        }
        if (meta & 0x20000u) ++u;
        bc = (bc == TRI_BUFS - 1) ? 0 : bc + 1;
        bi = (bi == TRI_BUFS - 1) ? 0 : bi + 1;
    }
    __syncthreads ();
    // publish: U(j,k) = w_j rho_{j-1}, candidates = w_t rho_{k-1}; the largest bound goes to the host
    int32_t mx = MAG_NEG;
    for (int t = tid; t < cnt; t += NT)
    {
        int32_t v = mb[t];
        const int lvl = a.publish ? ((t < nU) ? a.upos[t] : a.k) : 0;
        if (v != MAG_NEG && lvl >= 1) v += a.rho_mag[lvl - 1];
        a.mag_out[t] = v;
        mx = max (mx, v);
    }
    if (a.publish && a.bound_out)
    {
        __shared__ int32_t s_mx;
        if (tid == 0) s_mx = MAG_NEG;
        __syncthreads ();
        atomicMax (&s_mx, mx);
        __syncthreads ();
        if (tid == 0) atomicMax (a.bound_out, s_mx);       // running maximum over the session
    }
}

// All threads of the CTA are consumers; every thread also copies its share of the chunk that is
// two positions ahead in the work list with 16-byte cp.async (LDGSTS): TRI_BUFS-1 chunks are in
// flight per CTA while one is consumed.  The chunk descriptors travel through a small ring in
// shared memory, requested 2 BUFS - 2 chunks ahead in the same copy groups, so that the loop has no
// global load of its own.  (A TMA bulk-copy producer was measured first: with three small bulk
// copies per elimination step the copy engine, not HBM, set the pace -- about 1.1 us per step
// regardless of pipeline depth; see profiles/.)  One __syncthreads per chunk both publishes the
// landed chunk and orders the previous chunk's updates before the next yhat.
// A thread owns CPT channels of a row: CPT = 4 gives 256-thread CTAs, CPT = 2 gives 512-thread CTAs
// (twice the warps per SM for the same shared memory, at more instructions per element).
template <int CH, bool XS, int CPT>
__global__ void __launch_bounds__ (TRI_THREADS * 4 / CPT, CH == 4 ? 1 : 2) k_trisolve (TriArgs a)
{
    typedef TriSmem<CH> SM;
    typedef ChanVec<CPT> V;
    constexpr int TRI_BUFS = SM::BUFS, TRI_RING = SM::RING;
    constexpr int NT = TRI_THREADS * 4 / CPT, TPR = CH / CPT, RG = SM::RG;
    extern __shared__ __align__ (128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    if (a.mag_on && blockIdx.x == (unsigned) (a.S / CH))
    {   // the bound CTA of the launch (first right-hand side / the column itself only)
        if (blockIdx.y == 0) tri_mag_cta<CH, NT> (a, smem_raw);
        return;
    }
    // Several right-hand sides: CTAs of the same channel block are neighbours in the grid, run at
    // the same time and stream the same chunks of L, which then come from HBM once and from L2 for
    // the others (DRAM reads of a batch ~ 4 S nnz(L), not that times the batch size).
    const int cb = a.rhs_fastest ? blockIdx.y : blockIdx.x, by = a.rhs_fastest ? blockIdx.x : blockIdx.y;
    const int S = a.S, cnt = a.cnt, nU = a.nU, nchunks = a.nchunks;
    u32 *xg = a.out + (size_t) by * a.out_y_stride + (size_t) cb * a.out_cb_stride;
    u32 *xs = XS ? (u32 *) smem_raw : xg;
    const u32 smem0 = smem_u32 (smem_raw);
    const u32 ring = smem0 + (XS ? (u32) SM::x_bytes (cnt) : 0u);
    const u32 stage0 = ring + SM::RING_BYTES;
    const u32 spare = (u32) cnt * CH * 4;              // byte offset of the spare row

    // request descriptor X (two threads, 16 bytes each)
    auto fetch_desc = [&] (int X, int half)
    {
        cp_async16 (ring + (u32) (X % TRI_RING) * (u32) sizeof (ChunkInfo) + half * 16,
                    (const unsigned char *) (a.chunks + X) + half * 16);
    };
    // issue this thread's share of the copies of the chunk whose descriptor sits in ring slot rs
    auto issue = [&] (int rs, u32 sb)
    {
        const u32 da = ring + (u32) rs * (u32) sizeof (ChunkInfo);
        const uint4 d0 = lds128 (da);                  // lsrc (lo, hi), cbstride, slot_off
        const uint2 d1 = lds64 (da + 16);              // j, meta
        const int nrows = (int) (d1.y & 0xffffu);
        const unsigned char *lsrc = (const unsigned char *) (((unsigned long long) d0.y << 32) | d0.x)
                                  + ((size_t) cb * d0.z) * 4 + (size_t) tid * 16;
        const u32 ldst = sb + (u32) tid * 16;
        const int lp = nrows * (CH / 4);               // 16-byte pieces of L rows, at most CPT per thread
        if (nrows == SM::R)
        {   // full chunk (most of the bytes): no per-piece predicates
            cp_async16_off<0> (ldst, lsrc);
            cp_async16_off<NT * 16> (ldst, lsrc);
            if (CPT == 4)
            {
                cp_async16_off<2 * NT * 16> (ldst, lsrc);
                cp_async16_off<3 * NT * 16> (ldst, lsrc);
            }
        }
        else
        {
            cp_async16_off_if<0> (tid < lp, ldst, lsrc);
            cp_async16_off_if<NT * 16> (tid + NT < lp, ldst, lsrc);
            if (CPT == 4)
            {
                cp_async16_off_if<2 * NT * 16> (tid + 2 * NT < lp, ldst, lsrc);
                cp_async16_off_if<3 * NT * 16> (tid + 3 * NT < lp, ldst, lsrc);
            }
        }
        // one 16-byte group of entries per row group in use; the step's 1/rho_j with its first chunk
        cp_async16_if (tid < tri_active_groups (nrows, CH), sb + SM::L_BYTES + (u32) tid * 16,
                       a.slots + (size_t) d0.w + (size_t) tid * 4);
        cp_async16_if ((d1.y & 0x10000u) && tid < CH / 4, sb + SM::L_BYTES + SM::S_BYTES + (u32) tid * 16,
                       a.invrho + (size_t) d1.x * S + (size_t) cb * CH + 4 * tid);
    };

    // prologue: the first descriptors, then the first chunks, are requested while the vector is initialised
    // descriptors 0 .. 2 BUFS - 3: chunk c + BUFS - 1 is issued at iteration c, when only the copy
    // groups up to iteration c - BUFS + 1 are known to have landed
    if (tid < 2 * (2 * TRI_BUFS - 2) && (tid >> 1) < nchunks) fetch_desc (tid >> 1, tid & 1);
    cp_async_commit ();
    cp_async_wait<0> ();
    __syncthreads ();
    for (int c = 0; c < TRI_BUFS - 1; ++c) { if (c < nchunks) issue (c, stage0 + (u32) c * SM::STAGE); cp_async_commit (); }
    {
        const int words = (XS ? cnt + 1 : cnt) * CH;
        for (int i = tid; i < words; i += NT) xs[i] = 0;
        __syncthreads ();
        const u32 *src = a.src + (size_t) cb * a.src_total * CH;
        const size_t first = (size_t) a.src_first + (size_t) by * a.src_y_stride;
        if (a.src_cnt >= 64)
        {   // long source (dense right-hand side, speculative first part): CPT channels per access,
            // one slot lookup per row
            const int qs = (tid % TPR) * CPT;
            V sc, pq, nq;
            if (a.src_scale) { sc = ldv<CPT> (a.src_scale + cb * CH + qs); pq = ldv<CPT> (a.p + cb * CH + qs); nq = ldv<CPT> (a.ninv + cb * CH + qs); }
            for (int e = tid / TPR; e < a.src_cnt; e += NT / TPR)
            {
                const int row = a.src_rows ? a.src_rows[e] : e;
                V v = ldv<CPT> (src + (first + (size_t) e * a.src_step) * CH + qs);
                if (a.src_scale) v = mont_mulv<CPT> (v, sc, pq, nq);
                stv<CPT> (xs + (size_t) a.pos[row] * CH + qs, v);
            }
        }
        else
            for (int i = tid; i < a.src_cnt * CH; i += NT)
            {
                const int e = i / CH, ch = i % CH;
                const int row = a.src_rows ? a.src_rows[e] : e;
                u32 v = src[(first + (size_t) e * a.src_step) * CH + ch];
                if (a.src_scale) v = mont_mul (v, a.src_scale[cb * CH + ch], a.p[cb * CH + ch], a.ninv[cb * CH + ch]);
                xs[a.pos[row] * CH + ch] = v;
            }
    }

    const int qc = (tid % TPR) * CPT, rg = tid / TPR;  // first channel of this thread inside the block, row group
    const int c0 = cb * CH + qc;
    const V pv = ldv<CPT> (a.p + c0), niv = ldv<CPT> (a.ninv + c0);
    // the multiplier of a step (-yhat_j and its Shoup companion) is worked out ONCE per warp: lane c
    // does channel c of the block, the others fetch theirs with shuffles (every thread doing its own
    // four would spend more instructions on a step's set-up than on a short step's updates)
    const int ch1 = (tid & 31) % CH;
    const u32 p1 = a.p[cb * CH + ch1], ni1 = a.ninv[cb * CH + ch1];
    unsigned char *xb = (unsigned char *) xs + qc * 4;             // this thread's channels of row 0
    V negy, negyq;                                     // -yhat_j as a plain residue, and its Shoup companion
#pragma unroll
    for (int i = 0; i < CPT; ++i) { negy.v[i] = 0; negyq.v[i] = 0; }
    int u = a.u0;                                      // elimination step (= U slot) of the current chunk
    int bc = 0, bi = TRI_BUFS - 1;                     // buffers of the chunk consumed / requested
    for (int c = 0; c < nchunks; ++c)
    {
        cp_async_wait<TRI_BUFS - 2> ();                // this thread's copies of chunk c have landed
        __syncthreads ();                              // ... and everyone's; chunk c-1 is fully applied
        if (c + TRI_BUFS - 1 < nchunks) issue ((c + TRI_BUFS - 1) % TRI_RING, stage0 + (u32) bi * SM::STAGE);
        {   // descriptor 2 BUFS - 2 chunks ahead (two threads, 16 bytes each)
            const int X = c + 2 * TRI_BUFS - 2;
            cp_async16_if (tid < 2 && X < nchunks, ring + (u32) (X % TRI_RING) * (u32) sizeof (ChunkInfo) + (u32) (tid & 1) * 16,
                           (const unsigned char *) (a.chunks + X) + (tid & 1) * 16);
        }
        cp_async_commit ();
        const u32 sb = stage0 + (u32) bc * SM::STAGE;
        const uint2 jm = lds64 (ring + (u32) (c % TRI_RING) * (u32) sizeof (ChunkInfo) + 16);
        const u32 meta = jm.y;
        if (meta & 0x10000u)
        {   // yhat_j = w_j / rho_j
            const int uu = a.j_is_slot ? (int) jm.x : u;
            const u32 wj1 = xs[(size_t) uu * CH + ch1];
            u32 ir1;
            asm volatile ("ld.shared.u32 %0, [%1];" : "=r"(ir1) : "r"(sb + SM::L_BYTES + SM::S_BYTES + (u32) ch1 * 4));
            // -yhat_j as a plain residue and its Shoup companion floor (ny 2^32 / p).  The Montgomery
            // form nym = ny 2^32 mod p is at hand, ny 2^32 - nym is a multiple of p below p 2^32, and an
            // exact division by p is a multiplication by 1/p modulo 2^32: the companion is
            // nym * (-1/p) mod 2^32, one IMAD with the Montgomery constant
            const u32 ym = mont_mul (wj1, ir1, p1, ni1);                           // yhat_j, Montgomery form
            const u32 nym = ym ? p1 - ym : 0u;
            const u32 ny1 = mont_redc (nym, p1, ni1);                              // out of Montgomery form
            const u32 nyq1 = nym * ni1;
#pragma unroll
            for (int i = 0; i < CPT; ++i)
            {
                negy.v[i] = __shfl_sync (0xffffffffu, ny1, qc + i);
                negyq.v[i] = __shfl_sync (0xffffffffu, nyq1, qc + i);
            }
        }
        // rows of one step hit distinct slots: the four rows of a thread are loaded, updated and
        // stored together so that their latencies overlap
        if (rg < tri_active_groups ((int) (meta & 0xffffu), CH))
        {
            // four slot-list entries: the chunk row each L entry comes from and the slot it updates
            const uint4 tq = lds128 (sb + SM::L_BYTES + (u32) rg * 16);
            const u32 en[4] = { tq.x, tq.y, tq.z, tq.w };
            u32 t[4];
            const u32 lrow = sb + qc * 4;
            V l[4], w[4];
            const u32 xsa = smem0 + qc * 4;            // (XS) shared-memory address of this thread's channels of row 0
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                t[q] = (en[q] >> TRI_ROW_BITS) * (CH * 4);             // byte offset of the target row
                l[q] = ldsv<CPT> (lrow + (en[q] & TRI_ROW_MASK) * (CH * 4));
                if (XS && CPT == 4)
                {   // predicated in place (no branch, the four rows stay interleaved): rows without
                    // a target neither load nor store
                    asm volatile ("{\n .reg .pred p;\n setp.ne.u32 p, %4, %5;\n"
                                  " @p ld.shared.v4.u32 {%0,%1,%2,%3}, [%6];\n}"
                                  : "=r"(w[q].v[0]), "=r"(w[q].v[1]), "=r"(w[q].v[2]), "=r"(w[q].v[3])
                                  : "r"(t[q]), "r"(spare), "r"(xsa + t[q]));
                }
                else if (XS || t[q] != spare) w[q] = ldv<CPT> (xb + t[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                if (XS && CPT == 4)
                {
                    const V r = sub_mul_shoup<CPT> (w[q], l[q], negy, negyq, pv);
                    asm volatile ("{\n .reg .pred p;\n setp.ne.u32 p, %4, %5;\n"
                                  " @p st.shared.v4.u32 [%6], {%0,%1,%2,%3};\n}"
                                  :: "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]),
                                     "r"(t[q]), "r"(spare), "r"(xsa + t[q]) : "memory");
                }
                else if (XS || t[q] != spare) stv<CPT> (xb + t[q], sub_mul_shoup<CPT> (w[q], l[q], negy, negyq, pv));
            }
        }
        if (meta & 0x20000u) ++u;
        bc = (bc == TRI_BUFS - 1) ? 0 : bc + 1;
        bi = (bi == TRI_BUFS - 1) ? 0 : bi + 1;
    }
    __syncthreads ();
    // publish the column as true REF values: U(j,k) = w_j rho_{j-1}, candidates = w_t rho_{k-1}
    for (int t = rg; t < cnt; t += RG)
    {
        V v = ldv<CPT> (xs + t * CH + qc);
        if (a.publish == 2) v = mont_mulv<CPT> (v, ldv<CPT> (a.invrho + (size_t) t * S + c0), pv, niv);      // x_t = z_t / rho_t
        else
        {
            const int lvl = a.publish ? ((t < nU) ? a.upos[t] : a.k) : 0;      // level the entry is brought to
            if (lvl >= 1) v = mont_mulv<CPT> (v, ldv<CPT> (a.rho + (size_t) (lvl - 1) * S + c0), pv, niv);
        }
        stv<CPT> (xg + t * CH + qc, v);
    }
}

// ------------------------------------------------------------------------------------------------
// k_backsub: z = det * y, then for j = n-1..0: z_j /= U_jj ; z_i -= U_ij * z_j  (slots = positions)
// ------------------------------------------------------------------------------------------------
struct BackArgs
{
    int n, S;
    u32 *z; size_t z_y_stride, z_cb_stride;     // channel block cb of right-hand side y at z + cb * z_cb_stride + y * z_y_stride
    int rhs_fastest;                // grid order as in k_trisolve
    const ColDesc *desc; const u32 *rho, *invrho, *p, *ninv;
    const int32_t *pos;             // row -> position (final pinv)
};

template <int CH>
__global__ void __launch_bounds__ (512) k_backsub (BackArgs a)
{
    const int tid = threadIdx.x, ch = tid % CH, rg = tid / CH, RG = blockDim.x / CH;
    const int cb = a.rhs_fastest ? blockIdx.y : blockIdx.x, by = a.rhs_fastest ? blockIdx.x : blockIdx.y;
    const int c = cb * CH + ch, S = a.S, n = a.n;
    const u32 p = a.p[c], ni = a.ninv[c];
    u32 *z = a.z + (size_t) by * a.z_y_stride + (size_t) cb * a.z_cb_stride;
    const u32 det = a.rho[(size_t) (n - 1) * S + c];
    for (int t = rg; t < n; t += RG) z[t * CH + ch] = mont_mul (z[t * CH + ch], det, p, ni);
    __syncthreads ();
    u32 pend_val = 0; int pend_slot = -1;
    for (int j = n - 1; j >= 0; --j)
    {
        if (pend_slot >= 0) { z[pend_slot * CH + ch] = pend_val; pend_slot = -1; }
        const ColDesc d = a.desc[j];
        const u32 zj = mont_mul (z[j * CH + ch], a.invrho[(size_t) j * S + c], p, ni);
        if (rg == 0) { pend_slot = j; pend_val = zj; }
        const u32 negz = zj ? p - zj : 0u;
        const u32 *Ub = d.base + (size_t) cb * d.cnt * CH;
        for (int m = rg; m < d.nU; m += RG)
        {
            const int t = a.pos[d.rows[m]];
            const u32 uv = Ub[(size_t) m * CH + ch];
            u32 zv = z[t * CH + ch] + mont_mul (uv, negz, p, ni);      // zv - uv * zj
            if (zv >= p) zv -= p;
            z[t * CH + ch] = zv;
        }
        __syncthreads ();
    }
    if (pend_slot >= 0) z[pend_slot * CH + ch] = pend_val;
}

// ------------------------------------------------------------------------------------------------
// k_pivot_scan: exact nonzero / magnitude scan over the candidate slots (nU..cnt-1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cmp_mag (const u32 *dig, size_t ds, const int32_t *topd, int e1, int e2)
{
    const int t1 = topd[e1], t2 = topd[e2];
    if (t1 != t2) return t1 < t2 ? -1 : 1;
    for (int t = t1; t >= 0; --t)
    {
        const u32 d1 = dig[(size_t) e1 * ds + t], d2 = dig[(size_t) e2 * ds + t];
        if (d1 != d2) return d1 < d2 ? -1 : 1;
    }
    return 0;
}
// mode: 0 smallest, 1 largest, 2 first nonzero.  Ties go to the earlier slot.
__device__ __forceinline__ int better (const u32 *dig, size_t ds, const int32_t *topd, int mode, int x, int y)
{
    if (x < 0) return y;
    if (y < 0) return x;
    if (mode == 2) return x < y ? x : y;
    int c = cmp_mag (dig, ds, topd, x, y);
    if (mode == 1) c = -c;
    if (c < 0) return x;
    if (c > 0) return y;
    return x < y ? x : y;
}

struct ScanArgs
{
    int cnt, nU, mode, diag_slot;
    const u32 *dig; size_t ds; const int32_t *topd; const int8_t *sign; const int32_t *bad;
    slipcu_pivot_info *info;          // mapped host memory
    int32_t *mag; const int32_t *cum_ub; const int32_t *bound; int measured;
    int k; int32_t *run_flags;        // see slipcu_factor::run_flags
    int32_t seq;                      // written to info->seq last
};

// the scan itself, by all threads of one CTA of any size (<= 512 threads)
__device__ __noinline__ void pivot_scan_body (const ScanArgs &a)
{
    __shared__ int sbest[512];
    __shared__ int32_t s_meas;
    const int cnt = a.cnt, nU = a.nU, mode = a.mode;
    const u32 *dig = a.dig; const size_t ds = a.ds; const int32_t *topd = a.topd;
    int best = -1;
    if (a.measured)
    {   // measured mode: the largest candidate, B_top * d <= |v| < B_top * (d + 1) with d the top
        // mixed-radix digit.  A value that does not fit the channels reconstructs as a random
        // residue of the modulus and shows up here as (almost) full size.
        if (threadIdx.x == 0) s_meas = MAG_NEG;
        __syncthreads ();
        int32_t mx = MAG_NEG;
        for (int e = nU + threadIdx.x; e < cnt; e += blockDim.x)
        {
            const int t = topd[e];
            if (t >= 0) mx = max (mx, a.cum_ub[t] + mag_log2_ub (dig[(size_t) e * ds + t] + 1u));
        }
        atomicMax (&s_meas, mx);
    }
    if (a.mag)
    {   // bound mode: the candidates' bounds are replaced by their measured sizes
        for (int e = nU + threadIdx.x; e < cnt; e += blockDim.x)
        {
            const int t = topd[e];
            a.mag[e] = t < 0 ? MAG_NEG : a.cum_ub[t] + mag_log2_ub (dig[(size_t) e * ds + t] + 1u);
        }
    }
    for (int e = nU + threadIdx.x; e < cnt; e += blockDim.x)
        if (topd[e] >= 0) best = better (dig, ds, topd, mode, best, e);
    sbest[threadIdx.x] = best;
    __syncthreads ();
    if (threadIdx.x < 32)
    {
        int b = -1;
        for (int i = threadIdx.x; i < (int) blockDim.x; i += 32) b = better (dig, ds, topd, mode, b, sbest[i]);
        for (int off = 16; off > 0; off >>= 1)
        {
            const int o = __shfl_down_sync (0xffffffffu, b, off);
            b = better (dig, ds, topd, mode, b, o);
        }
        if (threadIdx.x == 0)
        {
            best = b;
            slipcu_pivot_info *info = a.info;
            info->best_slot = best;
            info->best_sign = best >= 0 ? a.sign[best] : 0;
            const int de = (a.diag_slot >= nU && a.diag_slot < cnt && topd[a.diag_slot] >= 0) ? 1 : 0;
            info->diag_eligible = de;
            info->diag_vs_best = (de && best >= 0) ? cmp_mag (dig, ds, topd, a.diag_slot, best) : 0;
            info->bad_channel = *a.bad;
            // running values: the host does not look at every column (single-candidate columns go
            // on without waiting), so what it must not miss accumulates on the device
            if (best < 0) atomicCAS (&a.run_flags[0], 0, a.k + 1);
            info->singular_col = a.run_flags[0];
            if (a.measured) { atomicMax (&a.run_flags[1], s_meas); info->bound_units = a.run_flags[1]; }
            else info->bound_units = a.bound ? *a.bound : 0;
            __threadfence_system ();
            *(volatile int32_t *) &info->seq = a.seq;
        }
    }
}

__global__ void __launch_bounds__ (256) k_pivot_scan (ScanArgs a) { pivot_scan_body (a); }

// Tail of the reconstruction kernels: the last CTA of the grid to finish runs the pivot scan, so a
// column costs one launch less on the path to its pivot.
__device__ __forceinline__ void scan_by_last_block (const ScanArgs &sc, unsigned *done_ctr)
{
    __shared__ int s_last;
    __threadfence ();
    __syncthreads ();
    if (threadIdx.x == 0)
    {
        const unsigned prev = atomicAdd (done_ctr, 1u);
        s_last = (prev == gridDim.x - 1);
        if (s_last) *done_ctr = 0;                     // ready for the next launch
    }
    __syncthreads ();
    if (!s_last) return;
    __threadfence ();
    pivot_scan_body (sc);
}

// ------------------------------------------------------------------------------------------------
// k_garner: residues -> mixed-radix digits of |x| (+ sign), one warp per entry.
//   x mod M = d_0 + d_1 p_0 + d_2 p_0 p_1 + ...,   d_t = (x_t - sum_{u<t} d_u C[u][t]) / B_t mod p_t
// ------------------------------------------------------------------------------------------------
struct GarnerArgs
{
    int cnt;                    // entries per channel block in the region
    int e0, ne;                 // entries e0 .. e0+ne-1 are reconstructed
    int s, CH, S;
    const u32 *base;            // [S/CH][cnt][CH]
    u32 *dig; size_t dstride;   // digits out, row per entry
    int32_t *topd;              // highest nonzero digit index of |x| (-1 for zero)
    int8_t *sign;
    const u32 *p, *ninv, *C, *invB;
    int scan_on; unsigned *done_ctr;      // fused pivot scan (see scan_by_last_block)
    ScanArgs sc;
};
__global__ void __launch_bounds__ (128) k_garner (GarnerArgs a)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w < a.ne)
    {
    const int e = a.e0 + w;
    const int s = a.s, S = a.S, CH = a.CH;
    u32 *dg = a.dig + (size_t) e * a.dstride;
    const unsigned full = 0xffffffffu;

    for (int t0 = 0; t0 < s; t0 += 32)
    {
        const int t = t0 + lane;
        const bool valid = t < s;
        const int tt = valid ? t : s - 1;
        const u32 p = a.p[tt], ni = a.ninv[tt];
        const u32 vm = a.base[((size_t) (tt / CH) * a.cnt + e) * CH + (tt % CH)];
        const u32 v = mont_redc (vm, p, ni);                 // plain residue
        u64 T = 0;
        const u32 *Ccol = a.C + tt;
        for (int u = 0; u < t0; u += 4)
        {
            const uint4 d4 = *reinterpret_cast<const uint4 *> (dg + u);
            const u32 c0 = Ccol[(size_t) (u + 0) * S], c1 = Ccol[(size_t) (u + 1) * S];
            const u32 c2 = Ccol[(size_t) (u + 2) * S], c3 = Ccol[(size_t) (u + 3) * S];
            lazy_mac (T, d4.x, c0, p); lazy_mac (T, d4.y, c1, p);
            lazy_mac (T, d4.z, c2, p); lazy_mac (T, d4.w, c3, p);
        }
        u32 mine = 0;
        for (int i = 0; i < 32; ++i)
        {
            u32 di = 0;
            if (lane == i)
            {
                const u32 acc = lazy_redc (T, p, ni);
                const u32 diff = v >= acc ? v - acc : v + p - acc;
                di = mont_mul (diff, a.invB[tt], p, ni);
                mine = di;
            }
            di = __shfl_sync (full, di, i);
            if (lane > i && t0 + i < s) lazy_mac (T, di, Ccol[(size_t) (t0 + i) * S], p);
        }
        if (valid) dg[t] = mine;
        __syncwarp ();
    }
    // sign: x is negative iff x mod M > (M-1)/2, whose digits are (p_t - 1)/2
    bool neg = false;
    for (int t0 = ((s - 1) / 32) * 32; t0 >= 0; t0 -= 32)
    {
        const int t = t0 + lane;
        u32 d = 0, h = 0;
        if (t < s) { d = dg[t]; h = (a.p[t] - 1) >> 1; }
        const unsigned ne = __ballot_sync (full, d != h);
        if (ne)
        {
            const int top = 31 - __clz (ne);
            neg = __shfl_sync (full, (int) (d > h), top) != 0;
            break;
        }
    }
    if (neg)
    {   // |x| = M - (x mod M): complement every digit, then add one
        for (int t = lane; t < s; t += 32) dg[t] = a.p[t] - 1 - dg[t];
        __syncwarp ();
        if (lane == 0)
        {
            for (int t = 0; t < s; ++t)
            {
                u32 d = dg[t] + 1;
                if (d == a.p[t]) dg[t] = 0; else { dg[t] = d; break; }
            }
        }
        __syncwarp ();
    }
    int top = -1;
    for (int t0 = ((s - 1) / 32) * 32; t0 >= 0; t0 -= 32)
    {
        const int t = t0 + lane;
        const u32 d = (t < s) ? dg[t] : 0u;
        const unsigned nzm = __ballot_sync (full, d != 0);
        if (nzm) { top = t0 + 31 - __clz (nzm); break; }
    }
    // zero-pad the digit row up to the next multiple of 4 (vector loads in later passes)
    for (int t = s + lane; t < ((s + 3) & ~3); t += 32) dg[t] = 0;
    if (lane == 0) { a.topd[e] = top; a.sign[e] = top < 0 ? 0 : (neg ? -1 : 1); }
    }
    if (a.scan_on) scan_by_last_block (a.sc, a.done_ctr);
}

// ------------------------------------------------------------------------------------------------
// k_garner_small: the same digits for sessions of few channels (s <= 128).  With a few dozen
// channels the reconstruction is nothing but the latency of its serial chain, and k_garner pays a
// global load of C inside every step of it (~30 us per column at s = 64, the largest item on the
// path to the pivot of an LP basis).  Here the CTA first copies the s x s corner of C into shared
// memory (coalesced, one round trip), then one warp per entry runs the chain out of registers and
// shared memory: lane l owns the digits l, l+32, l+64, l+96; step i = the owner finishes d_i, a
// shuffle hands it round, every lane adds d_i C[i][t] to its later digits.
// ------------------------------------------------------------------------------------------------
#define GS_MAX 128
__global__ void __launch_bounds__ (128) k_garner_small (GarnerArgs a)
{
    extern __shared__ u32 gs_c[];                       // [s][s]  C[u][t], u < t
    const int lane = threadIdx.x & 31;
    const int s = a.s, S = a.S, CH = a.CH;
    const unsigned full = 0xffffffffu;
    for (int i = threadIdx.x; i < s * s; i += blockDim.x)
    {
        const int u = i / s, t = i - u * s;
        gs_c[i] = (u < t) ? a.C[(size_t) u * S + t] : 0u;
    }
    __syncthreads ();
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w < a.ne)
    {
    const int e = a.e0 + w;
    u32 *dg = a.dig + (size_t) e * a.dstride;
    u32 pt[4], nit[4], ib[4], v[4], dgt[4];
    u64 T[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        const int t = lane + 32 * j, tt = t < s ? t : s - 1;
        pt[j] = a.p[tt]; nit[j] = a.ninv[tt]; ib[j] = a.invB[tt];
        v[j] = mont_redc (a.base[((size_t) (tt / CH) * a.cnt + e) * CH + (tt % CH)], pt[j], nit[j]);
        T[j] = 0; dgt[j] = 0;
    }
#pragma unroll
    for (int jo = 0; jo < 4; ++jo)
    {
        if (32 * jo >= s) break;
        const int iend = min (32, s - 32 * jo);
        for (int il = 0; il < iend; ++il)
        {
            const int i = 32 * jo + il;
            u32 di = 0;
            if (lane == il)
            {
                const u32 acc = lazy_redc (T[jo], pt[jo], nit[jo]);
                const u32 diff = v[jo] >= acc ? v[jo] - acc : v[jo] + pt[jo] - acc;
                di = mont_mul (diff, ib[jo], pt[jo], nit[jo]);
                dgt[jo] = di;
            }
            di = __shfl_sync (full, di, il);
            const u32 *crow = gs_c + i * s;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const int t = lane + 32 * j;
                if (j >= jo && t > i && t < s) lazy_mac (T[j], di, crow[t], pt[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int t = lane + 32 * j; if (t < s) dg[t] = dgt[j]; }
    __syncwarp ();
    // sign: x is negative iff x mod M > (M-1)/2, whose digits are (p_t - 1)/2
    bool neg = false;
    for (int t0 = ((s - 1) / 32) * 32; t0 >= 0; t0 -= 32)
    {
        const int t = t0 + lane;
        u32 d = 0, h = 0;
        if (t < s) { d = dg[t]; h = (a.p[t] - 1) >> 1; }
        const unsigned ne = __ballot_sync (full, d != h);
        if (ne)
        {
            const int top = 31 - __clz (ne);
            neg = __shfl_sync (full, (int) (d > h), top) != 0;
            break;
        }
    }
    if (neg)
    {   // |x| = M - (x mod M): complement every digit, then add one
        for (int t = lane; t < s; t += 32) dg[t] = a.p[t] - 1 - dg[t];
        __syncwarp ();
        if (lane == 0)
        {
            for (int t = 0; t < s; ++t)
            {
                u32 d = dg[t] + 1;
                if (d == a.p[t]) dg[t] = 0; else { dg[t] = d; break; }
            }
        }
        __syncwarp ();
    }
    int top = -1;
    for (int t0 = ((s - 1) / 32) * 32; t0 >= 0; t0 -= 32)
    {
        const int t = t0 + lane;
        const u32 d = (t < s) ? dg[t] : 0u;
        const unsigned nzm = __ballot_sync (full, d != 0);
        if (nzm) { top = t0 + 31 - __clz (nzm); break; }
    }
    for (int t = s + lane; t < ((s + 3) & ~3); t += 32) dg[t] = 0;
    if (lane == 0) { a.topd[e] = top; a.sign[e] = top < 0 ? 0 : (neg ? -1 : 1); }
    }
    if (a.scan_on) scan_by_last_block (a.sc, a.done_ctr);
}

// ------------------------------------------------------------------------------------------------
// k_garner_tiled: same result as k_garner, organised like a blocked triangular solve so that the
// table C is read once per E entries and the long dependency chain is one CTA-wide step per 32
// digits.  A CTA of W warps reconstructs E entries.  Digit block b (positions 32b..32b+31) is owned
// by warp b mod W, which keeps the running sums  T[e] = sum_u d_e[u] C[u][t]  of its blocks in
// registers.  Per block: the owner finishes its 32 digits (in-warp shuffles), publishes them, and
// all warps add their contribution to the blocks they still own (right-looking).  Digits of
// earlier tiles (s > 32*W*BPW) are applied first, left-looking, from global memory.
// ------------------------------------------------------------------------------------------------
template <int E, int BPW>
__global__ void __launch_bounds__ (512) k_garner_tiled (GarnerArgs a)
{
    __shared__ __align__ (16) u32 dcur[2][E][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int s = a.s, S = a.S, CH = a.CH;
    const int B = (s + 31) >> 5;
    const unsigned full = 0xffffffffu;
    const int g0 = blockIdx.x * E;
    int ent[E];                         // entry indices (clamped; invalid ones recompute the last)
#pragma unroll
    for (int e = 0; e < E; ++e) ent[e] = a.e0 + min (g0 + e, a.ne - 1);

    for (int tb0 = 0; tb0 < B; tb0 += W * BPW)
    {
        u64 acc[E][BPW];
        u32 pt[BPW], nit[BPW];
        int tpos[BPW];
#pragma unroll
        for (int i = 0; i < BPW; ++i)
        {
            const int t = 32 * (tb0 + w + i * W) + lane;
            tpos[i] = t < s ? t : s - 1;
            pt[i] = a.p[tpos[i]]; nit[i] = a.ninv[tpos[i]];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e][i] = 0;
        }
        // (1) digits of the previous tiles
        for (int u = 0; u < tb0 * 32; u += 4)
        {
            uint4 d4[E];
#pragma unroll
            for (int e = 0; e < E; ++e) d4[e] = *reinterpret_cast<const uint4 *> (a.dig + (size_t) ent[e] * a.dstride + u);
#pragma unroll
            for (int i = 0; i < BPW; ++i)
            {
                const u32 *Cc = a.C + (size_t) u * S + tpos[i];
                const u32 c0 = Cc[0], c1 = Cc[S], c2 = Cc[2 * (size_t) S], c3 = Cc[3 * (size_t) S];
#pragma unroll
                for (int e = 0; e < E; ++e)
                {
                    lazy_mac (acc[e][i], d4[e].x, c0, pt[i]); lazy_mac (acc[e][i], d4[e].y, c1, pt[i]);
                    lazy_mac (acc[e][i], d4[e].z, c2, pt[i]); lazy_mac (acc[e][i], d4[e].w, c3, pt[i]);
                }
            }
        }
        // (2) the blocks of this tile, in order
#pragma unroll
        for (int li = 0; li < BPW; ++li)
        {
            for (int ow = 0; ow < W; ++ow)
            {
                const int b = tb0 + li * W + ow;
                if (b >= B) break;                       // uniform across the CTA
                const int buf = b & 1;
                if (w == ow)
                {   // finish the 32 digits of block b for the E entries
                    const int t = 32 * b + lane;
                    const bool valid = t < s;
                    const int tt = tpos[li];
                    const u32 p = pt[li], ni = nit[li];
                    const u32 ib = a.invB[tt];
                    u32 v[E], mine[E];
                    u64 T[E];
#pragma unroll
                    for (int e = 0; e < E; ++e)
                    {
                        v[e] = mont_redc (a.base[((size_t) (tt / CH) * a.cnt + ent[e]) * CH + (tt % CH)], p, ni);
                        T[e] = acc[e][li]; mine[e] = 0;
                    }
                    // the 32x32 diagonal block of C for this lane's column, loaded up front so that
                    // the serial 32-step elimination below runs out of registers
                    const u32 *Ccol = a.C + tt + (size_t) (32 * b) * S;
                    u32 cc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) cc[i] = Ccol[(size_t) i * S];
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                    {
#pragma unroll
                        for (int e = 0; e < E; ++e)
                        {
                            u32 di = 0;
                            if (lane == i)
                            {
                                const u32 r = lazy_redc (T[e], p, ni);
                                const u32 diff = v[e] >= r ? v[e] - r : v[e] + p - r;
                                di = mont_mul (diff, ib, p, ni);
                                mine[e] = di;
                            }
                            di = __shfl_sync (full, di, i);
                            if (lane > i) lazy_mac (T[e], di, cc[i], p);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < E; ++e)
                    {
                        if (valid && g0 + e < a.ne) a.dig[(size_t) ent[e] * a.dstride + t] = mine[e];
                        dcur[buf][e][lane] = valid ? mine[e] : 0u;
                    }
                }
                __syncthreads ();
                // every warp: add block b's digits to the blocks it still owns
#pragma unroll 2
                for (int q = 0; q < 32; q += 4)
                {
                    uint4 d4[E];
#pragma unroll
                    for (int e = 0; e < E; ++e) d4[e] = *reinterpret_cast<const uint4 *> (&dcur[buf][e][q]);
#pragma unroll
                    for (int i = 0; i < BPW; ++i)
                    {
                        // local block i is block tb0 + w + i*W; it is later than b iff i > li or (i == li and w > ow)
                        if (i > li || (i == li && w > ow))
                        {
                            const u32 *Cc = a.C + (size_t) (32 * b + q) * S + tpos[i];
                            const u32 c0 = Cc[0], c1 = Cc[S], c2 = Cc[2 * (size_t) S], c3 = Cc[3 * (size_t) S];
#pragma unroll
                            for (int e = 0; e < E; ++e)
                            {
                                lazy_mac (acc[e][i], d4[e].x, c0, pt[i]); lazy_mac (acc[e][i], d4[e].y, c1, pt[i]);
                                lazy_mac (acc[e][i], d4[e].z, c2, pt[i]); lazy_mac (acc[e][i], d4[e].w, c3, pt[i]);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads ();
    }
    __syncthreads ();
    // sign, magnitude digits and top digit: one warp per entry
    if (w < E && g0 + w < a.ne)
    {
    const int e = a.e0 + g0 + w;
    u32 *dg = a.dig + (size_t) e * a.dstride;
    bool neg = false;
    for (int t0 = ((s - 1) / 32) * 32; t0 >= 0; t0 -= 32)
    {
        const int t = t0 + lane;
        u32 d = 0, h = 0;
        if (t < s) { d = dg[t]; h = (a.p[t] - 1) >> 1; }
        const unsigned ne = __ballot_sync (full, d != h);
        if (ne)
        {
            const int top = 31 - __clz (ne);
            neg = __shfl_sync (full, (int) (d > h), top) != 0;
            break;
        }
    }
    if (neg)
    {
        for (int t = lane; t < s; t += 32) dg[t] = a.p[t] - 1 - dg[t];
        __syncwarp ();
        if (lane == 0)
        {
            for (int t = 0; t < s; ++t)
            {
                u32 d = dg[t] + 1;
                if (d == a.p[t]) dg[t] = 0; else { dg[t] = d; break; }
            }
        }
        __syncwarp ();
    }
    int top = -1;
    for (int t0 = ((s - 1) / 32) * 32; t0 >= 0; t0 -= 32)
    {
        const int t = t0 + lane;
        const u32 d = (t < s) ? dg[t] : 0u;
        const unsigned nzm = __ballot_sync (full, d != 0);
        if (nzm) { top = t0 + 31 - __clz (nzm); break; }
    }
    for (int t = s + lane; t < ((s + 3) & ~3); t += 32) dg[t] = 0;
    if (lane == 0) { a.topd[e] = top; a.sign[e] = top < 0 ? 0 : (neg ? -1 : 1); }
    }
    if (a.scan_on) scan_by_last_block (a.sc, a.done_ctr);
}

// ------------------------------------------------------------------------------------------------
// k_garner_flow: the tiled reconstruction as a dataflow inside the CTA (no CTA-wide barriers).
// Block b (32 digits) is owned by warp b mod W.  A warp walks the blocks in order; when it reaches
// a block it owns it finishes its 32 digits and raises ready[b]; otherwise it waits for ready[b].
// Either way it then adds block b's digits to the later blocks it owns, NEAREST FIRST, and if that
// nearest block is b+1 it finishes and publishes it before touching the others: the serial chain
// (finish b -> apply to b+1 -> finish b+1 ...) never waits for the bulk of the updates, which the
// other warps carry out behind it.  Requires all digit blocks to fit one tile (B <= W*BPW).
// ------------------------------------------------------------------------------------------------
template <int E, int BPW>
__global__ void __launch_bounds__ (384, 1) k_garner_flow (GarnerArgs a)
{
    extern __shared__ __align__ (16) unsigned char gsm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int s = a.s, S = a.S, CH = a.CH;
    const int B = (s + 31) >> 5;
    u32 *digs = (u32 *) gsm;                               // [B][E][32]
    volatile int *ready = (volatile int *) (digs + (size_t) B * E * 32);
    const unsigned full = 0xffffffffu;
    const int g0 = blockIdx.x * E;
    int ent[E];
#pragma unroll
    for (int e = 0; e < E; ++e) ent[e] = a.e0 + min (g0 + e, a.ne - 1);
    for (int b = threadIdx.x; b < B; b += blockDim.x) ready[b] = 0;
    __syncthreads ();

    u64 acc[BPW][E];
    u32 pt[BPW], nit[BPW];
    int tpos[BPW];
#pragma unroll
    for (int i = 0; i < BPW; ++i)
    {
        const int t = 32 * (w + i * W) + lane;
        tpos[i] = t < s ? t : s - 1;
        pt[i] = a.p[tpos[i]]; nit[i] = a.ninv[tpos[i]];
#pragma unroll
        for (int e = 0; e < E; ++e) acc[i][e] = 0;
    }

    u32 *ccs = (u32 *) (ready + B) + (size_t) w * 1024;          // this warp's 32x32 block of C
    // add the digits of block b to one owned block
    auto apply = [&] (int b, u64 (&T)[E], u32 p, int tt)
    {
        const u32 *Cc = a.C + (size_t) (32 * b) * S + tt;
        const u32 *db = digs + (size_t) b * E * 32;
#pragma unroll 1
        for (int q = 0; q < 32; q += 16)
        {
            u32 c[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) c[r] = Cc[(size_t) (q + r) * S];      // 16 loads in flight
#pragma unroll
            for (int r = 0; r < 16; r += 4)
            {
#pragma unroll
                for (int e = 0; e < E; ++e)
                {
                    const uint4 d4 = *reinterpret_cast<const uint4 *> (db + e * 32 + q + r);
                    lazy_mac2 (T[e], d4.x, c[r], d4.y, c[r + 1], p);
                    lazy_mac2 (T[e], d4.z, c[r + 2], d4.w, c[r + 3], p);
                }
            }
        }
    };

    for (int b = -1; b < B - 1 || b < 0; ++b)
    {
        if (b >= 0)
        {
            if (b % W != w) { while (ready[b] == 0) __nanosleep (40); }   // back off: a spinning warp steals issue slots
            __threadfence_block ();
            __syncwarp ();
        }
        // the block that continues the chain, b+1, first -- if it is mine
        u64 T[E];
        u32 fp = 0, fni = 0; int ftt = 0; bool mine_next = false;
#pragma unroll
        for (int i = 0; i < BPW; ++i)
        {
            const int bo = w + i * W;
            if (bo == b + 1 && bo < B)
            {
                if (b >= 0) apply (b, acc[i], pt[i], tpos[i]);
#pragma unroll
                for (int e = 0; e < E; ++e) T[e] = acc[i][e];
                fp = pt[i]; fni = nit[i]; ftt = tpos[i]; mine_next = true;
            }
        }
        if (mine_next)
        {   // finish the 32 digits of block b+1 for the E entries and publish them
            const int bo = b + 1;
            const int t = 32 * bo + lane;
            const bool valid = t < s;
            const u32 ib = a.invB[ftt];
            u32 v[E], mine[E];
#pragma unroll
            for (int e = 0; e < E; ++e)
            {
                v[e] = mont_redc (a.base[((size_t) (ftt / CH) * a.cnt + ent[e]) * CH + (ftt % CH)], fp, fni);
                mine[e] = 0;
            }
            const u32 *Ccol = a.C + ftt + (size_t) (32 * bo) * S;
            // diagonal block of C, pre-multiplied by 1/B_t:  d_t = (v_t - sum)/B_t  then becomes a
            // running value q_t that loses d_i * (C[i][t]/B_t) per finished digit -- one Montgomery
            // product per step on the serial chain instead of two
#pragma unroll 8
            for (int i = 0; i < 32; ++i) ccs[i * 32 + lane] = mont_mul (Ccol[(size_t) i * S], ib, fp, fni);
            __syncwarp ();
            u32 qv[E];
#pragma unroll
            for (int e = 0; e < E; ++e)
            {
                const u32 r = lazy_redc (T[e], fp, fni);
                const u32 diff = v[e] >= r ? v[e] - r : v[e] + fp - r;
                qv[e] = mont_mul (diff, ib, fp, fni);
            }
#pragma unroll 2
            for (int i = 0; i < 32; ++i)
            {
                const u32 cci = ccs[i * 32 + lane];
#pragma unroll
                for (int e = 0; e < E; ++e)
                {
                    const u32 di = __shfl_sync (full, qv[e], i);
                    if (lane == i) mine[e] = di;
                    if (lane > i)
                    {
                        const u32 sub = mont_mul (di, cci, fp, fni);
                        qv[e] = qv[e] >= sub ? qv[e] - sub : qv[e] + fp - sub;
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < E; ++e)
            {
                if (valid && g0 + e < a.ne) a.dig[(size_t) ent[e] * a.dstride + t] = mine[e];
                digs[((size_t) bo * E + e) * 32 + lane] = valid ? mine[e] : 0u;
            }
            __threadfence_block ();
            __syncwarp ();
            if (lane == 0) ready[bo] = 1;
        }
        // then the rest of my later blocks
        if (b >= 0)
        {
#pragma unroll
            for (int i = 0; i < BPW; ++i)
            {
                const int bo = w + i * W;
                if (bo > b + 1 && bo < B) apply (b, acc[i], pt[i], tpos[i]);
            }
        }
    }
    __syncthreads ();
    // sign, magnitude digits and top digit: one warp per entry
    if (w < E && g0 + w < a.ne)
    {
    const int e = a.e0 + g0 + w;
    u32 *dg = a.dig + (size_t) e * a.dstride;
    bool neg = false;
    for (int t0 = ((s - 1) / 32) * 32; t0 >= 0; t0 -= 32)
    {
        const int t = t0 + lane;
        u32 d = 0, h = 0;
        if (t < s) { d = dg[t]; h = (a.p[t] - 1) >> 1; }
        const unsigned ne = __ballot_sync (full, d != h);
        if (ne)
        {
            const int top = 31 - __clz (ne);
            neg = __shfl_sync (full, (int) (d > h), top) != 0;
            break;
        }
    }
    if (neg)
    {
        for (int t = lane; t < s; t += 32) dg[t] = a.p[t] - 1 - dg[t];
        __syncwarp ();
        if (lane == 0)
        {
            for (int t = 0; t < s; ++t)
            {
                u32 d = dg[t] + 1;
                if (d == a.p[t]) dg[t] = 0; else { dg[t] = d; break; }
            }
        }
        __syncwarp ();
    }
    int top = -1;
    for (int t0 = ((s - 1) / 32) * 32; t0 >= 0; t0 -= 32)
    {
        const int t = t0 + lane;
        const u32 d = (t < s) ? dg[t] : 0u;
        const unsigned nzm = __ballot_sync (full, d != 0);
        if (nzm) { top = t0 + 31 - __clz (nzm); break; }
    }
    for (int t = s + lane; t < ((s + 3) & ~3); t += 32) dg[t] = 0;
    if (lane == 0) { a.topd[e] = top; a.sign[e] = top < 0 ? 0 : (neg ? -1 : 1); }
    }
    if (a.scan_on) scan_by_last_block (a.sc, a.done_ctr);
}

// ------------------------------------------------------------------------------------------------
// k_limbs: mixed-radix digits -> positional 32-bit limbs, one warp per entry.
//   limb_l = sum_t d_t * B_t[l] accumulated in 96 bits per lane, then a carry chain across lanes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mac96 (u32 &a0, u32 &a1, u32 &a2, u32 d, u32 b)
{
    asm ("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
         "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
         "addc.u32 %2, %2, 0;"
         : "+r"(a0), "+r"(a1), "+r"(a2) : "r"(d), "r"(b));
}

struct LimbArgs
{
    int e0, ne;                 // digit rows e0 .. e0+ne-1
    int out0;                   // limb row of entry e0 (rows follow consecutively)
    int stride, LB;
    const u32 *dig; size_t dstride; const int32_t *topd;
    const u32 *Bpos;
    u32 *limbs; int32_t *nl;
};

// E entries per warp share every load of the table Bpos (the same idea as in the Garner kernels)
template <int E>
__global__ void __launch_bounds__ (128) k_limbs (LimbArgs a)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w * E >= a.ne) return;
    const unsigned full = 0xffffffffu;
    int ent[E], eo[E], top[E];
    const u32 *dg[E];
    int maxtop = -1;
#pragma unroll
    for (int i = 0; i < E; ++i)
    {
        const int k = min (w * E + i, a.ne - 1);          // surplus lanes of the last group redo the last entry
        ent[i] = a.e0 + k; eo[i] = a.out0 + k;
        top[i] = a.topd[ent[i]];
        dg[i] = a.dig + (size_t) ent[i] * a.dstride;
        maxtop = max (maxtop, top[i]);
    }
    u32 in_a1[E], in_a2[E], in_b2[E], carry_in[E];
    int nl[E];
#pragma unroll
    for (int i = 0; i < E; ++i) { in_a1[i] = in_a2[i] = in_b2[i] = carry_in[i] = 0; nl[i] = 0; }
    for (int l0 = 0; l0 < a.stride; l0 += 32)
    {
        const int l = l0 + lane;
        const int lr = l < a.LB ? l : a.LB - 1;      // lanes past the row read a zero column
        u32 a0[E], a1[E], a2[E];
#pragma unroll
        for (int i = 0; i < E; ++i) a0[i] = a1[i] = a2[i] = 0;
        // B_t has no limb l for t < l (p_i < 2^32); digits above an entry's top digit are zero and
        // the digit rows are zero-padded to a multiple of four
        for (int t = l0 & ~3; t <= maxtop; t += 4)
        {
            const u32 *Bp = a.Bpos + (size_t) t * a.LB + lr;
            const u32 b0 = Bp[0], b1 = Bp[a.LB], b2 = Bp[2 * (size_t) a.LB], b3 = Bp[3 * (size_t) a.LB];
#pragma unroll
            for (int i = 0; i < E; ++i)
            {
                const uint4 d4 = *reinterpret_cast<const uint4 *> (dg[i] + t);
                mac96 (a0[i], a1[i], a2[i], d4.x, b0);
                mac96 (a0[i], a1[i], a2[i], d4.y, b1);
                mac96 (a0[i], a1[i], a2[i], d4.z, b2);
                mac96 (a0[i], a1[i], a2[i], d4.w, b3);
            }
        }
#pragma unroll
        for (int i = 0; i < E; ++i)
        {
            // column sum for limb l: a0[l] + a1[l-1] + a2[l-2]
            u32 p1 = __shfl_up_sync (full, a1[i], 1), p2 = __shfl_up_sync (full, a2[i], 2);
            if (lane == 0) { p1 = in_a1[i]; p2 = in_b2[i]; }
            if (lane == 1) { p2 = in_a2[i]; }
            const u64 sum = (u64) a0[i] + p1 + p2;
            // carries are < 4: iterate the ripple until it settles
            u32 cin = (lane == 0) ? carry_in[i] : 0u, cout;
            for (;;)
            {
                cout = (u32) ((sum + cin) >> 32);
                u32 nin = __shfl_up_sync (full, cout, 1);
                if (lane == 0) nin = carry_in[i];
                const bool changed = nin != cin;
                cin = nin;
                if (!__any_sync (full, changed)) break;
            }
            const u32 limb = (u32) (sum + cin);
            cout = (u32) ((sum + cin) >> 32);
            if (l < a.stride && w * E + i < a.ne) a.limbs[(size_t) eo[i] * a.stride + l] = limb;
            const unsigned nzm = __ballot_sync (full, limb != 0 && l < a.stride);
            if (nzm) nl[i] = l0 + 32 - __clz (nzm);
            carry_in[i] = __shfl_sync (full, cout, 31);
            in_a1[i] = __shfl_sync (full, a1[i], 31);
            in_a2[i] = __shfl_sync (full, a2[i], 31);
            in_b2[i] = __shfl_sync (full, a2[i], 30);
        }
    }
#pragma unroll
    for (int i = 0; i < E; ++i)
        if (lane == 0 && w * E + i < a.ne) a.nl[eo[i]] = top[i] < 0 ? 0 : nl[i];
}

// ------------------------------------------------------------------------------------------------
// k_fraccrt / k_fracselect: pivot search without reconstructing the candidates.
//
// For X = x mod M (M = p_0..p_{s-1}, |x| < M/4) the Chinese remainder theorem gives
//     X / M = frac ( sum_i c_i / p_i ),   c_i = x_i * (M/p_i)^-1 mod p_i .
// With u_i = floor (2^(32W) / p_i) the W-word integer  F = sum_i c_i u_i  mod 2^(32W)  satisfies
//     2^(32W) * frac(X/M)  in  [F, F + s*2^31)   (circularly),
// so g = min (F, 2^(32W) - F) approximates 2^(32W) |x| / M within s*2^31 < 2^44 units: s*W
// multiply-adds per entry instead of the s^2/2 of the mixed-radix reconstruction, and no serial
// chain.  k_fracselect takes the extreme g and accepts it only if every other candidate differs
// from it by far more than the error bound (first differing word at least five words above the
// bottom, 96-bit window difference >= 2); otherwise -- ties, near ties, or too few words for the
// size of the entries -- the caller adds words or falls back to the exact scan.  Zero entries are
// recognised exactly (all residues zero).  The decision is therefore the exact one whenever it is
// accepted, which is what keeps the factorization bit-identical to the reference.
// ------------------------------------------------------------------------------------------------
struct FracKey                          // approximate magnitude of one candidate
{
    int32_t lead;                       // leading zero words of g (W if none is set); -1: the entry is exactly zero
    uint32_t k[4];                      // the four words of g from the leading one down (zero padded)
    int32_t neg;                        // the entry is negative
};

struct FracArgs
{
    int cnt, e0, ne, s, CH, S, W;       // region rows, first entry, entries, channels, table stride, words
    const u32 *base, *p, *ninv, *minv, *urec;
    FracKey *key;                       // [ne]
};

#define FRAC_WARPS 8
// One CTA per E entries; the channels are dealt to the 8 warps in chunks of 32, each warp keeps the
// 96-bit column sums of its share (lane l owns word positions l, l+32, ...), warp e adds the
// eight partial sums of entry e, ripples the carries and extracts the key.
template <int E, int NW>
__global__ void __launch_bounds__ (FRAC_WARPS * 32) k_fraccrt (FracArgs a)
{
    extern __shared__ u32 fsm[];                 // [FRAC_WARPS][E][NW*32][3] partial sums, then flags
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g0 = blockIdx.x * E;
    const unsigned full = 0xffffffffu;
    const int s = a.s, W = a.W, CH = a.CH;
    int ent[E];
#pragma unroll
    for (int e = 0; e < E; ++e) ent[e] = a.e0 + min (g0 + e, a.ne - 1);
    u32 a0[E][NW], a1[E][NW], a2[E][NW];
    bool nz[E];
#pragma unroll
    for (int e = 0; e < E; ++e)
    {
        nz[e] = false;
#pragma unroll
        for (int j = 0; j < NW; ++j) a0[e][j] = a1[e][j] = a2[e][j] = 0;
    }
    for (int i0 = w * 32; i0 < s; i0 += FRAC_WARPS * 32)
    {
        const int i = i0 + lane;
        u32 c[E];
#pragma unroll
        for (int e = 0; e < E; ++e) c[e] = 0;
        if (i < s)
        {
            const u32 pi = a.p[i], ni = a.ninv[i], mi = a.minv[i];
#pragma unroll
            for (int e = 0; e < E; ++e)
            {
                const u32 x = a.base[((size_t) (i / CH) * a.cnt + ent[e]) * CH + (i % CH)];
                nz[e] = nz[e] || (x != 0);
                c[e] = mont_redc (mont_mul (x, mi, pi, ni), pi, ni);        // standard form, < p_i
            }
        }
        // the reciprocal words of eight channels are requested together (the loop is bound by the
        // latency of these L2 loads); channels past s have c = 0 and read row s-1 again
#pragma unroll 1
        for (int j0 = 0; j0 < 32; j0 += 8)
        {
            if (i0 + j0 >= s) break;
            u32 u[8][NW];
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const u32 *ur = a.urec + (size_t) min (i0 + j0 + j, s - 1) * FRAC_WMAX;
#pragma unroll
                for (int q = 0; q < NW; ++q)
                {
                    const int l = lane + 32 * q;              // word position, 0 = least significant
                    u[j][q] = (l < W) ? ur[W - 1 - l] : 0u;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
#pragma unroll
                for (int e = 0; e < E; ++e)
                {
                    const u32 cj = __shfl_sync (full, c[e], j0 + j);
#pragma unroll
                    for (int q = 0; q < NW; ++q) mac96 (a0[e][q], a1[e][q], a2[e][q], cj, u[j][q]);
                }
            }
        }
    }
    int *nzflag = (int *) (fsm + (size_t) FRAC_WARPS * E * NW * 32 * 3);       // [FRAC_WARPS][E]
#pragma unroll
    for (int e = 0; e < E; ++e)
    {
#pragma unroll
        for (int q = 0; q < NW; ++q)
        {
            u32 *d = fsm + (((size_t) (w * E + e) * NW + q) * 32 + lane) * 3;
            d[0] = a0[e][q]; d[1] = a1[e][q]; d[2] = a2[e][q];
        }
        const bool any = __any_sync (full, nz[e]);
        if (lane == 0) nzflag[w * E + e] = any ? 1 : 0;
    }
    __syncthreads ();
    if (w >= E || g0 + w >= a.ne) return;
    const int e = w;
    bool nonzero = false;
    for (int v = 0; v < FRAC_WARPS; ++v) nonzero = nonzero || nzflag[v * E + e];
    u32 word[NW];
    u32 in_a1 = 0, in_a2 = 0, in_b2 = 0, carry_in = 0;
#pragma unroll
    for (int q = 0; q < NW; ++q)
    {
        u64 T0 = 0, T1 = 0, T2 = 0;
        for (int v = 0; v < FRAC_WARPS; ++v)
        {
            const u32 *d = fsm + (((size_t) (v * E + e) * NW + q) * 32 + lane) * 3;
            T0 += d[0]; T1 += d[1]; T2 += d[2];
        }
        T1 += T0 >> 32; T2 += T1 >> 32;
        const u32 b0 = (u32) T0, b1 = (u32) T1, b2 = (u32) T2;
        u32 p1 = __shfl_up_sync (full, b1, 1), p2 = __shfl_up_sync (full, b2, 2);
        if (lane == 0) { p1 = in_a1; p2 = in_b2; }
        if (lane == 1) { p2 = in_a2; }
        const u64 sum = (u64) b0 + p1 + p2;
        u32 cin = (lane == 0) ? carry_in : 0u, cout;
        for (;;)
        {
            cout = (u32) ((sum + cin) >> 32);
            u32 nin = __shfl_up_sync (full, cout, 1);
            if (lane == 0) nin = carry_in;
            const bool changed = nin != cin;
            cin = nin;
            if (!__any_sync (full, changed)) break;
        }
        word[q] = (u32) (sum + cin);
        cout = (u32) ((sum + cin) >> 32);
        carry_in = __shfl_sync (full, cout, 31);
        in_a1 = __shfl_sync (full, b1, 31);
        in_a2 = __shfl_sync (full, b2, 31);
        in_b2 = __shfl_sync (full, b2, 30);
    }
    // sign = top bit of the top word (position W-1); g = F or its complement (one unit off at most)
    const int tp = W - 1;
    u32 topw = 0;
#pragma unroll
    for (int q = 0; q < NW; ++q) { const u32 v = __shfl_sync (full, word[q], tp & 31); if (q == (tp >> 5)) topw = v; }
    const bool neg = (topw >> 31) != 0;
    int hp = -1;                                 // highest position with a nonzero word of g
#pragma unroll
    for (int q = 0; q < NW; ++q)
    {
        const int l = lane + 32 * q;
        word[q] = (l < W) ? (neg ? ~word[q] : word[q]) : 0u;
        const unsigned m = __ballot_sync (full, word[q] != 0);
        if (m) hp = 32 * q + 31 - __clz (m);
    }
    u32 key[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        const int pos = hp - j;
        u32 val = 0;
#pragma unroll
        for (int q = 0; q < NW; ++q) { const u32 v = __shfl_sync (full, word[q], pos & 31); if (pos >= 0 && q == (pos >> 5)) val = v; }
        key[j] = val;
    }
    if (lane == 0)
    {
        FracKey k;
        k.lead = nonzero ? (hp < 0 ? W : W - 1 - hp) : -1;
        k.k[0] = key[0]; k.k[1] = key[1]; k.k[2] = key[2]; k.k[3] = key[3];
        k.neg = neg ? 1 : 0;
        a.key[g0 + e] = k;
    }
}

struct FracSel
{
    int ne, nU, mode, diag_slot, W;     // candidates (slots nU .. nU+ne-1), 0 smallest / 1 largest
    const FracKey *key; const int32_t *bad;
    slipcu_pivot_info *info;
    int k; int32_t *run_flags;
    int32_t seq;
};
// order of the approximate magnitudes: more leading zero words is smaller, then the key words
__device__ __forceinline__ int frac_cmp (const FracKey &x, const FracKey &y)
{
    if (x.lead != y.lead) return x.lead > y.lead ? -1 : 1;
    for (int j = 0; j < 4; ++j) if (x.k[j] != y.k[j]) return x.k[j] < y.k[j] ? -1 : 1;
    return 0;
}
__device__ __forceinline__ int frac_better (const FracKey *key, int mode, int x, int y)
{
    if (x < 0) return y;
    if (y < 0) return x;
    int c = frac_cmp (key[x], key[y]);
    if (mode == 1) c = -c;
    if (c < 0) return x;
    if (c > 0) return y;
    return x < y ? x : y;
}
// |small| < |large| proven: the 96-bit windows at the leading word of `large`, five or more words
// above the bottom where the error of the sums lives, differ by at least two units
__device__ __forceinline__ bool frac_proven_less (const FracKey &small, const FracKey &large, int W)
{
    if (large.lead > W - 6 || small.lead < large.lead) return false;
    const int d = small.lead - large.lead;
    unsigned __int128 L = ((unsigned __int128) large.k[0] << 64) | ((unsigned __int128) large.k[1] << 32) | large.k[2];
    unsigned __int128 S = 0;
    if (d == 0) S = ((unsigned __int128) small.k[0] << 64) | ((unsigned __int128) small.k[1] << 32) | small.k[2];
    else if (d == 1) S = ((unsigned __int128) small.k[0] << 32) | small.k[1];
    else if (d == 2) S = small.k[0];
    return L > S && L - S >= 2;
}
__global__ void __launch_bounds__ (256) k_fracselect (FracSel a)
{
    __shared__ int sbest[256];
    __shared__ int s_uncertain;
    if (threadIdx.x == 0) s_uncertain = 0;
    int best = -1;
    for (int r = threadIdx.x; r < a.ne; r += blockDim.x)
        if (a.key[r].lead >= 0) best = frac_better (a.key, a.mode, best, r);
    sbest[threadIdx.x] = best;
    __syncthreads ();
    for (int h = blockDim.x >> 1; h > 0; h >>= 1)
    {
        if ((int) threadIdx.x < h)
            sbest[threadIdx.x] = frac_better (a.key, a.mode, sbest[threadIdx.x], sbest[threadIdx.x + h]);
        __syncthreads ();
    }
    best = sbest[0];
    if (best >= 0)
    {
        const FracKey kb = a.key[best];
        for (int r = threadIdx.x; r < a.ne; r += blockDim.x)
        {
            if (r == best) continue;
            const FracKey kr = a.key[r];
            if (kr.lead < 0) continue;
            const bool ok = (a.mode == 0) ? frac_proven_less (kb, kr, a.W) : frac_proven_less (kr, kb, a.W);
            if (!ok) s_uncertain = 1;
        }
    }
    __syncthreads ();
    if (threadIdx.x == 0)
    {
        slipcu_pivot_info *info = a.info;
        info->best_slot = best >= 0 ? a.nU + best : -1;
        info->best_sign = best >= 0 ? (a.key[best].neg ? -1 : 1) : 0;
        const int dr = a.diag_slot - a.nU;
        const int de = (dr >= 0 && dr < a.ne && a.key[dr].lead >= 0) ? 1 : 0;
        info->diag_eligible = de;
        info->diag_vs_best = (de && best >= 0 && dr != best) ? (a.mode == 0 ? 1 : -1) : 0;
        info->bad_channel = *a.bad;
        if (best < 0) atomicCAS (&a.run_flags[0], 0, a.k + 1);
        info->singular_col = a.run_flags[0];
        info->bound_units = 0;
        info->reserved[0] = s_uncertain ? 0 : 1;      // 1: the choice is proven
        info->reserved[1] = best >= 0 ? a.key[best].lead : 0;
        info->reserved[2] = a.W;
        __threadfence_system ();
        *(volatile int32_t *) &info->seq = a.seq;
    }
}

// ------------------------------------------------------------------------------------------------
// k_pivot_commit: rho_k, rho_k^-1, rho_k / rho_{k-1} in every channel; column descriptor
// ------------------------------------------------------------------------------------------------
struct CommitArgs
{
    int k, S, CH, slot; ColDesc d; ColDesc *desc; u32 *rho, *invrho;
    const u32 *p, *ninv, *one; int32_t *bad; int32_t *rho_mag;
};
__device__ __forceinline__ void pivot_commit_body (const CommitArgs &a, int c)
{
    if (c == 0) { ColDesc d = a.d; d.pivslot = a.slot; a.desc[a.k] = d; if (a.rho_mag) a.rho_mag[a.k] = d.mag[a.slot]; }
    if (c >= a.S) return;
    const u32 pc = a.p[c], ni = a.ninv[c];
    const u32 v = a.d.base[((size_t) (c / a.CH) * a.d.cnt + a.slot) * a.CH + (c % a.CH)];
    if (v == 0) atomicCAS (a.bad, 0, c + 1);
    a.rho[(size_t) a.k * a.S + c] = v;
    a.invrho[(size_t) a.k * a.S + c] = mont_pow (v, pc - 2, a.one[c], pc, ni);
}
__global__ void k_pivot_commit (CommitArgs a) { pivot_commit_body (a, blockIdx.x * blockDim.x + threadIdx.x); }

// Zero test of a single candidate from its residues (the value is below the product of the channel
// primes, so it is zero iff every residue is): all a column with one candidate needs when neither
// sizes nor digits are wanted.
__global__ void __launch_bounds__ (256) k_zero_test (int k, int S, int CH, int cnt, int slot, const u32 *base, int32_t *run_flags)
{
    __shared__ int s_nz;
    if (threadIdx.x == 0) s_nz = 0;
    __syncthreads ();
    int nz = 0;
    for (int c = threadIdx.x; c < S; c += blockDim.x)
        nz |= base[((size_t) (c / CH) * cnt + slot) * CH + (c % CH)] != 0;
    if (nz) s_nz = 1;
    __syncthreads ();
    if (threadIdx.x == 0 && !s_nz) atomicCAS (&run_flags[0], 0, k + 1);
}

// First kernel of a column: brings the pattern packet over from mapped host memory (no copy-engine
// operation), fills the row -> slot map, and commits the PREVIOUS column's pivot in its spare
// blocks (the host decided it a moment ago; nothing of this kernel depends on it).
struct PrepArgs
{
    int cnt, pk_ints, copy_blocks, has_commit;
    const int32_t *h_packet; int32_t *d_packet; int32_t *pos;
    CommitArgs c;
};
__global__ void __launch_bounds__ (256) k_prep (PrepArgs a)
{
    if ((int) blockIdx.x < a.copy_blocks)
    {
        const int i = blockIdx.x * 256 + threadIdx.x;
        if (i < a.pk_ints)
        {
            const int32_t v = a.h_packet[i];
            a.d_packet[i] = v;
            if (i < a.cnt) a.pos[v] = i;
        }
    }
    else if (a.has_commit) pivot_commit_body (a.c, ((int) blockIdx.x - a.copy_blocks) * 256 + threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// host side of the C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int slipcu_device_count (void)
{
    int n = 0;
    if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError (); return 0; }
    return n;
}
// The device is a property of the PROCESS (one process per GPU), not of the calling thread: CUDA's
// current device is per thread, so a session created on a worker thread would otherwise land on
// device 0.  Sessions take g_device when it is set, and every entry point makes the session's own
// device current before it touches the runtime.
static std::atomic<int> g_device{-1};
static std::atomic<int> g_live_sessions{0};       // sessions alive in the process (polling is for a lone session only)
extern "C" int slipcu_set_device (int device)
{
    CU (cudaSetDevice (device));
    g_device = device;
    return SLIPCU_OK;
}
#define USE_DEVICE(F) CU (cudaSetDevice ((F)->device))

// ------------------------------------------------------------------------------------------------
// 32-bit integer-multiply peak of the device, measured: register-resident chains of one multiply
// form (kind 0: mad.wide.u32 = IMAD.WIDE, 1: mad.lo.u32 = IMAD, 2: mul.hi.u32 = IMAD.HI), eight
// independent chains per thread, enough warps to fill every scheduler.  bench.py reports the
// kernels' multiply rates against these numbers (roofline.int_mul).
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__ (256) k_imad_peak (int iters, u32 *out)
{
    u32 a[8], b = blockIdx.x * 40503u + 12345u;
    u64 w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = (threadIdx.x + 1u) * 2654435761u + i; w[i] = a[i]; }
    __syncthreads ();
    const long long c0 = clock64 ();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            // every multiplicand depends on the previous result, so nothing is loop-invariant
            if (KIND == 0) w[i] = (u64) (u32) w[i] * b + w[i];
            else if (KIND == 1) a[i] = a[i] * b + a[(i + 1) & 7];
            else if (KIND == 2) a[i] = __umulhi (a[i], b) + 0x9e3779b9u;
            else if (KIND == 3)
            {   // the Montgomery multiply-subtract w <- w + l * ny mod p (k_trisolve of round 1; still the
                // operation of the other kernels), on registers.  Operands are not kept reduced: only
                // the instruction mix matters here.
                const u32 p = 0x7fffffffu - 18u, ninv = b | 1u;
                a[i] = add_mod (a[i], mont_mul (a[i], b, p, ninv), p);
            }
            else
            {   // KIND 4: the update of k_trisolve itself, the same with a Shoup product (sub_mul_shoup)
                const u32 p = 0x7fffffffu - 18u, ny = b >> 2, nyq = b | 1u;
                const u32 q = __umulhi (a[i], nyq);
                const u32 t = a[i] * ny - q * p;
                a[i] = csub (a[i] + csub (t, p), p);
            }
        }
    }
    const long long c1 = clock64 ();
    u64 t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += w[i] + a[i];
    if ((u32) t == 0xdeadbeefu) out[0] = (u32) (t >> 32);    // practically never true: keeps the chains alive
    if (threadIdx.x == 0) out[1 + blockIdx.x] = (u32) (c1 - c0);          // SM cycles of this CTA's loop
}

// Runs one kind: returns operations per second (wall) and per SM clock cycle (all 8 CTAs of an SM run
// side by side for the same number of cycles, so the per-cycle rate does not depend on the clock the
// GPU happens to hold under this -- power-hungry -- load).
static int run_imad_kind (int kind, double *per_s, double *per_sm_cycle)
{
    int dev = 0, sms = 148;
    if (g_device.load () >= 0) CU (cudaSetDevice (g_device.load ()));
    CU (cudaGetDevice (&dev));
    cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = 8, grid = sms * per_sm, iters = 1 << 13;
    u32 *out = nullptr;
    CU (cudaMalloc (&out, (size_t) (grid + 1) * sizeof (u32)));
    cudaEvent_t e0, e1;
    CU (cudaEventCreate (&e0)); CU (cudaEventCreate (&e1));
    const double ops = (double) grid * 256.0 * (double) iters * 8.0;
    double best = 0, best_cyc = 0;
    std::vector<u32> cyc (grid + 1);
    for (int rep = 0; rep < 4; ++rep)
    {
        CU (cudaEventRecord (e0, 0));
        if (kind == 0) k_imad_peak<0><<<grid, 256>>> (iters, out);
        else if (kind == 1) k_imad_peak<1><<<grid, 256>>> (iters, out);
        else if (kind == 2) k_imad_peak<2><<<grid, 256>>> (iters, out);
        else if (kind == 3) k_imad_peak<3><<<grid, 256>>> (iters, out);
        else k_imad_peak<4><<<grid, 256>>> (iters, out);
        CU (cudaEventRecord (e1, 0));
        CU (cudaEventSynchronize (e1));
        float ms = 0;
        CU (cudaEventElapsedTime (&ms, e0, e1));
        g_launches++;
        if (rep == 0 || ms <= 0) continue;
        CU (cudaMemcpy (cyc.data (), out, (size_t) (grid + 1) * sizeof (u32), cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 1; i <= grid; ++i) mean += cyc[i];
        mean /= grid;
        const double rate = ops / (ms * 1e-3), pc = (double) per_sm * 256.0 * iters * 8.0 / mean;
        if (rate > best) best = rate;
        if (pc > best_cyc) best_cyc = pc;
    }
    if (per_s) *per_s = best;
    if (per_sm_cycle) *per_sm_cycle = best_cyc;
    cudaEventDestroy (e0); cudaEventDestroy (e1); cudaFree (out);
    return SLIPCU_OK;
}

// out[0..4] = operations per second, out[5..9] = operations per SM clock cycle, for IMAD.WIDE,
// IMAD, IMAD.HI, the Montgomery multiply-subtract and the Shoup multiply-subtract of k_trisolve
extern "C" int slipcu_measure_int_peaks (double *out10)
{
    if (!out10) return fail (SLIPCU_BAD_INPUT, "slipcu_measure_int_peaks", "bad argument");
    for (int k = 0; k < 5; ++k) { int rc = run_imad_kind (k, &out10[k], &out10[5 + k]); if (rc) return rc; }
    return SLIPCU_OK;
}
extern "C" int slipcu_measure_imad_peak (double *wide_per_s, double *lo_per_s, double *hi_per_s)
{
    double o[10];
    int rc = slipcu_measure_int_peaks (o);
    if (rc) return rc;
    if (wide_per_s) *wide_per_s = o[0];
    if (lo_per_s) *lo_per_s = o[1];
    if (hi_per_s) *hi_per_s = o[2];
    return SLIPCU_OK;
}
extern "C" int slipcu_measure_modmul_peak (double *modmul_per_s)
{
    return run_imad_kind (4, modmul_per_s, nullptr);
}

static int env_int (const char *name, int dflt)
{
    const char *v = getenv (name);
    return (v && *v) ? atoi (v) : dflt;
}

static void free_workctx (WorkCtx &w)
{
    pool_free (w.pos); pool_free (w.slots); pool_free (w.steps); pool_free (w.chunks);
    host_pool_free (w.h_packet);
    for (cudaEvent_t e : w.pk_ev) if (e) release_event (e);
    w = WorkCtx ();
}
static int init_workctx (WorkCtx &w, int n, cudaStream_t st)
{
    w.st = st;
    CU (pool_alloc_t (&w.pos, (size_t) n * sizeof (int32_t)));
    CU (cudaMemsetAsync (w.pos, 0, (size_t) n * sizeof (int32_t), st));
    w.pk_stride = ((size_t) 4 * n + 8 + 31) & ~(size_t) 31;
    CU (host_pool_alloc ((void **) &w.h_packet, w.pk_stride * PK_RING * sizeof (int32_t)));
    CU (cudaHostGetDevicePointer ((void **) &w.h_packet_dev, w.h_packet, 0));
    for (int i = 0; i < PK_RING; ++i) CU (pooled_event (&w.pk_ev[i], false));
    return SLIPCU_OK;
}

extern "C" void slipcu_factor_free (slipcu_factor *F)
{
    if (!F) return;
    if (F->counted) g_live_sessions--;
    if (getenv ("SLIP_B200_TIMING"))
        fprintf (stderr, "slipcu host wall: alloc %.3f packet+prepass %.3f trisolve-launch %.3f garner-launch %.3f scan-launch %.3f wait %.3f\n",
                 g_hw[0], g_hw[1], g_hw[2], g_hw[3], g_hw[4], g_hw[5]);
    for (double &v : g_hw) v = 0;
    cudaSetDevice (F->device);
    if (F->st && F->ev_start && F->ev_end && !F->timed && !F->rows_are_positions && F->cur >= 0)
    {   // a session that never reached its solve (an attempt aborted for more channels, a prime
        // retired, a singular matrix): its device time belongs to the job as well
        if (cudaEventRecord (F->ev_end, F->st) == cudaSuccess && cudaEventSynchronize (F->ev_end) == cudaSuccess)
        {
            float ms = 0;
            if (cudaEventElapsedTime (&ms, F->ev_start, F->ev_end) == cudaSuccess) g_device_ms += ms; else cudaGetLastError ();
        }
        else cudaGetLastError ();
    }
    if (F->st) cudaStreamSynchronize (F->st);
    if (F->st2) cudaStreamSynchronize (F->st2);
    flush_timers (F);
    pool_free (F->dAp); pool_free (F->dAi); pool_free (F->dA);
    pool_free (F->rho); pool_free (F->invrho);
    pool_free (F->desc); pool_free (F->bad);
    pool_free (F->digbuf[0]); pool_free (F->topdbuf[0]); pool_free (F->digbuf[1]); pool_free (F->topdbuf[1]);
    pool_free (F->frackey);
    pool_free (F->Amag); pool_free (F->rho_mag); pool_free (F->bound); pool_free (F->done_ctr); pool_free (F->run_flags);
    if (getenv ("SLIP_B200_TIMING") && F->frac)
        fprintf (stderr, "slipcu pivot search: %llu columns by approximate magnitudes, %llu word-count retries, %llu exact fallbacks\n",
                 (unsigned long long) F->frac_cols, (unsigned long long) F->frac_retries, (unsigned long long) F->frac_fallbacks);
    pool_free (F->tmp_limbs); pool_free (F->tmp_nl);
    free_workctx (F->mc);
    for (SpecSlot &sl : F->spec)
    {
        if (sl.w.st) cudaStreamSynchronize (sl.w.st);
        free_workctx (sl.w);
        pool_free (sl.buf); pool_free (sl.mag);
        if (sl.done) release_event (sl.done);
        if (sl.consumed) release_event (sl.consumed);
        if (sl.w.st) release_stream (sl.w.st);
    }
    if (F->ev_commit) release_event (F->ev_commit);
    host_pool_free (F->h_info);
    if (F->ev) release_event (F->ev);
    if (F->ev0) release_event (F->ev0);
    if (F->ev1) release_event (F->ev1);
    if (F->ev_start) release_event (F->ev_start);
    if (F->ev_end) release_event (F->ev_end);
    if (F->ev_tri) release_event (F->ev_tri);
    if (F->ev_gl) release_event (F->ev_gl);
    if (F->ev_side[0]) release_event (F->ev_side[0]);
    if (F->ev_side[1]) release_event (F->ev_side[1]);
    if (F->st2) release_stream (F->st2);
    if (F->st) release_stream (F->st);
    delete F;
}

extern "C" int slipcu_factor_channels (const slipcu_factor *F) { return F ? F->S : 0; }
// largest bound (units of 1/64 bit) a column may report and still be reconstructed exactly from
// the session's channels: |v| < M/2 with two bits to spare
extern "C" int slipcu_factor_capacity_units (const slipcu_factor *F)
{
    if (!F) return 0;
    // measured mode: 36 bits of headroom, so that a value that wrapped around the modulus (uniform
    // below it) passes as fitting with probability 2^-35 -- and then the caller's exact check decides
    return (int) floor (64.0 * F->tab->cumbits[F->S]) - (F->measured ? 36 : 3) * 64;
}

static int ensure_digits (slipcu_factor *F, size_t rows)
{
    if (rows <= F->dig_rows) return SLIPCU_OK;
    size_t want = std::max (rows, F->dig_rows * 2);
    want = std::min<size_t> (std::max<size_t> (want, 64), std::max<size_t> (rows, (size_t) F->n));
    CU (cudaStreamSynchronize (F->st));
    if (F->st2) CU (cudaStreamSynchronize (F->st2));
    F->side_pending[0] = F->side_pending[1] = false;
    F->dig = nullptr; F->topd = nullptr; F->dig_rows = 0;
    for (int i = 0; i < 2; ++i)
    {
        pool_free (F->digbuf[i]); pool_free (F->topdbuf[i]); F->digbuf[i] = nullptr; F->topdbuf[i] = nullptr;
        if (i == 1 && !F->overlap) break;
        CU (pool_alloc_t (&F->digbuf[i], want * (size_t) (F->S + 4) * sizeof (u32)));
        CU (pool_alloc_t (&F->topdbuf[i], want * sizeof (int32_t)));
    }
    F->dig = F->digbuf[0]; F->topd = F->topdbuf[0];
    F->dig_rows = want;
    return SLIPCU_OK;
}

// opt the k_trisolve instances in to the full shared memory of the SM
template <int CH, bool XS, int CPT> static cudaError_t tri_configure_one (int smem_optin)
{
    cudaError_t e = cudaFuncSetAttribute (k_trisolve<CH, XS, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute (k_trisolve<CH, XS, CPT>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
}
template <int CPT> static cudaError_t tri_configure_cpt (int smem_optin)
{
    cudaError_t e;
    if ((e = tri_configure_one<4, true, CPT> (smem_optin)) != cudaSuccess) return e;
    if ((e = tri_configure_one<4, false, CPT> (smem_optin)) != cudaSuccess) return e;
    if ((e = tri_configure_one<8, true, CPT> (smem_optin)) != cudaSuccess) return e;
    if ((e = tri_configure_one<16, true, CPT> (smem_optin)) != cudaSuccess) return e;
    if ((e = tri_configure_one<32, true, CPT> (smem_optin)) != cudaSuccess) return e;
    if ((e = tri_configure_one<8, false, CPT> (smem_optin)) != cudaSuccess) return e;
    if ((e = tri_configure_one<16, false, CPT> (smem_optin)) != cudaSuccess) return e;
    return tri_configure_one<32, false, CPT> (smem_optin);
}
static int tri_configure (int smem_optin)
{
    CU (tri_configure_cpt<4> (smem_optin));
    CU (tri_configure_cpt<2> (smem_optin));
    return SLIPCU_OK;
}

// everything a session needs apart from the input matrix
static int session_common_init (slipcu_factor *F, int n, int channels)
{
    int ndev = 0;
    if (cudaGetDeviceCount (&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError ();
        return fail (SLIPCU_CUDA_ERROR, "slip_lu_b200", "no CUDA device: this library has no CPU path");
    }
    if (g_device.load () >= 0) { F->device = g_device.load (); CU (cudaSetDevice (F->device)); }
    else CU (cudaGetDevice (&F->device));
    g_live_sessions++; F->counted = true;
    F->n = n;
    const int S = (channels + 31) & ~31;
    int rc = get_tables (S, F->tab);
    if (rc) return rc;
    F->S = S;
    // channel block width: 8-channel blocks give the most CTAs per wave and the smallest work
    // vector per CTA (measured best at n = 2000 and n = 3000: 3.1 vs 2.7 TB/s against 16-channel
    // blocks); wider blocks only when even they fill many waves of two CTAs per SM
    int sms = 148;
    cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, F->device);
    int CH = 8;
    // few channels (LP bases, measured mode): 4-channel blocks double the CTAs of a column, whose
    // single chain of steps is otherwise streamed by S/8 < 64 CTAs on 148 SMs
    if (S <= 512) CH = 4;
    if (S / 16 >= 16 * sms) CH = 16;
    if (S / 32 >= 16 * sms) CH = 32;
    CH = env_int ("SLIP_B200_CH", CH);
    if (CH != 4 && CH != 8 && CH != 16 && CH != 32) CH = 16;
    F->CH = CH;
    F->sms = sms;
    F->x_global = env_int ("SLIP_B200_X_GLOBAL", 0);
    F->garner_mode = env_int ("SLIP_B200_GARNER", 2);
    F->frac = env_int ("SLIP_B200_FRAC", 1);
    F->frac_margin = std::max (0, env_int ("SLIP_B200_FRAC_MARGIN", 12));      // 0 in tests: forces the add-words retry
    F->frac_verify = env_int ("SLIP_B200_FRAC_VERIFY", 0);
    static std::mutex attr_mutex;
    static std::vector<int> attr_done;
    bool configure = false;
    {
        std::lock_guard<std::mutex> lk (attr_mutex);
        if (std::find (attr_done.begin (), attr_done.end (), F->device) == attr_done.end ()) { attr_done.push_back (F->device); configure = true; }
    }
    if (configure) {
    CU (cudaFuncSetAttribute (k_garner_small, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_MAX * GS_MAX * (int) sizeof (u32)));
    CU (cudaFuncSetAttribute (k_garner_flow<1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU (cudaFuncSetAttribute (k_garner_flow<2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU (cudaFuncSetAttribute (k_garner_flow<3, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU (cudaFuncSetAttribute (k_garner_flow<4, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU (cudaFuncSetAttribute (k_garner_flow<5, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU (cudaFuncSetAttribute (k_garner_flow<6, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    }
    F->garner_e = env_int ("SLIP_B200_GARNER_E", 0);       // 0: chosen per launch
    int smem_optin = 0;
    cudaDeviceGetAttribute (&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, F->device);
    F->smem_limit = (size_t) smem_optin;
    F->cpt = env_int ("SLIP_B200_CPT", 4);             // channels per thread of k_trisolve: 4 or 2
    if (F->cpt != 2 && F->cpt != 4) F->cpt = 4;
    F->threads = TRI_THREADS * 4 / F->cpt;
    // sessions of thousands of channels are bound by k_trisolve's shared-memory traffic: their slot lists
    // are ordered by bank group; sessions of few channels are bound by the latency of a column,
    // where the plain order (no ballots in k_slots) is the shorter path
    F->slot_sort = env_int ("SLIP_B200_SLOT_SORT", F->S > 512 ? 1 : 0) ? 1 : 0;
    if (F->n >= (int) TRI_SLOT_MASK) return fail (SLIPCU_BAD_INPUT, "session", "more than 4M rows");
    if (configure)
    {
        rc = tri_configure (smem_optin - 1024);      // static shared memory of the kernels (a few words) comes out of the same budget
        if (rc) return rc;
    }

    CU (pooled_stream (&F->st));
    F->wst = F->st;
    CU (pooled_event (&F->ev, false));
    CU (pooled_event (&F->ev0, true)); CU (pooled_event (&F->ev1, true));
    CU (pooled_event (&F->ev_start, true)); CU (pooled_event (&F->ev_end, true));
    CU (pool_alloc_t (&F->rho, (size_t) n * S * sizeof (u32)));
    CU (pool_alloc_t (&F->invrho, (size_t) n * S * sizeof (u32)));
    CU (pool_alloc_t (&F->desc, (size_t) n * sizeof (ColDesc)));
    CU (pool_alloc_t (&F->bad, sizeof (int32_t)));
    CU (cudaMemset (F->bad, 0, sizeof (int32_t)));
    CU (pool_alloc_t (&F->done_ctr, sizeof (unsigned)));
    CU (cudaMemset (F->done_ctr, 0, sizeof (unsigned)));
    {
        const int32_t init[2] = { 0, MAG_NEG };
        CU (pool_alloc_t (&F->run_flags, 2 * sizeof (int32_t)));
        CU (cudaMemcpy (F->run_flags, init, sizeof (init), cudaMemcpyHostToDevice));
    }
    F->frac_min_s = std::max (16, env_int ("SLIP_B200_FRAC_MIN_S", 64));
    rc = init_workctx (F->mc, n, F->st);
    if (rc) return rc;
    CU (pooled_event (&F->ev_commit, false));
    // the scan result goes straight into mapped host memory (a 48-byte store over PCIe instead of a
    // copy-engine operation per column); the event behind the kernel publishes it to the host
    CU (host_pool_alloc ((void **) &F->h_info, sizeof (slipcu_pivot_info)));
    memset (F->h_info, 0, sizeof (slipcu_pivot_info));
    CU (cudaHostGetDevicePointer ((void **) &F->d_info, F->h_info, 0));
    F->cols.resize (n);
    return SLIPCU_OK;
}

extern "C" int slipcu_factor_begin (slipcu_factor **out, int n, int nz, const int32_t *Ap,
                                    const int32_t *Ai, const u32 *Alimbs, const int64_t *Aoff,
                                    const int8_t *Asign, int channels, int keep_positional, int bound_mode)
{
    if (!out || n <= 0 || nz <= 0 || !Ap || !Ai || !Alimbs || !Aoff || !Asign || channels <= 0)
        return fail (SLIPCU_BAD_INPUT, "slipcu_factor_begin", "bad argument");
    slipcu_factor *F = new slipcu_factor ();
    *out = F;
    int rc = session_common_init (F, n, channels);
    if (rc) return rc;
    F->nz = nz;
    F->keep_positional = keep_positional ? 1 : 0;
    if (bound_mode == 2)
    {   // measured mode (see slip_b200_device.h)
        F->measured = 1;
        F->frac = 0;                     // measured sizes come from the exact digits
    }
    else if (bound_mode)
    {   // input magnitudes from the limb strings: 64 log2 |a| rounded up
        F->mag_on = 1;
        F->frac = 0;                     // measured sizes come from the exact digits
        std::vector<int32_t> am (nz);
        for (int e = 0; e < nz; ++e)
        {
            const int64_t l0 = Aoff[e], l1 = Aoff[e + 1];
            if (Asign[e] == 0 || l1 == l0) { am[e] = MAG_NEG; continue; }
            int64_t top = l1 - 1;
            while (top > l0 && Alimbs[top] == 0) --top;
            double v = (double) Alimbs[top];
            if (top > l0) v += ((double) Alimbs[top - 1] + 1.0) / 4294967296.0; else v += 0.0;
            if (v <= 0) { am[e] = MAG_NEG; continue; }
            am[e] = (int32_t) ceil (64.0 * (log2 (v) + 32.0 * (double) (top - l0)) + 1e-6) + 1;
        }
        CU (pool_alloc_t (&F->Amag, (size_t) nz * sizeof (int32_t)));
        CU (pool_alloc_t (&F->rho_mag, (size_t) n * sizeof (int32_t)));
        CU (pool_alloc_t (&F->bound, sizeof (int32_t)));
        CU (cudaMemcpy (F->Amag, am.data (), (size_t) nz * sizeof (int32_t), cudaMemcpyHostToDevice));
        CU (cudaMemset (F->bound, 0, sizeof (int32_t)));
    }
    if (F->keep_positional && env_int ("SLIP_B200_OVERLAP", 1))
    {   // second stream for the positional reconstruction (see the session struct)
        F->overlap = 1;
        CU (pooled_stream (&F->st2));
        CU (pooled_event (&F->ev_tri, false));
        CU (pooled_event (&F->ev_gl, false));
        CU (pooled_event (&F->ev_side[0], false));
        CU (pooled_event (&F->ev_side[1], false));
    }
    const int S = F->S, CH = F->CH;
    CU (pool_alloc_t (&F->dAp, (size_t) (n + 1) * sizeof (int32_t)));
    CU (pool_alloc_t (&F->dAi, (size_t) nz * sizeof (int32_t)));
    CU (pool_alloc_t (&F->dA, (size_t) nz * S * sizeof (u32)));
    F->hAp.assign (Ap, Ap + n + 1);
    CU (cudaMemcpy (F->dAp, Ap, (size_t) (n + 1) * sizeof (int32_t), cudaMemcpyHostToDevice));
    CU (cudaMemcpy (F->dAi, Ai, (size_t) nz * sizeof (int32_t), cudaMemcpyHostToDevice));
    // reduce A into the channels
    {
        u32 *dl = nullptr; int64_t *doff = nullptr; int8_t *dsg = nullptr;
        const size_t nl = (size_t) Aoff[nz];
        CU (pool_alloc_t (&dl, std::max<size_t> (nl, 1) * sizeof (u32)));
        CU (pool_alloc_t (&doff, (size_t) (nz + 1) * sizeof (int64_t)));
        CU (pool_alloc_t (&dsg, (size_t) nz));
        CU (cudaMemcpy (dl, Alimbs, nl * sizeof (u32), cudaMemcpyHostToDevice));
        CU (cudaMemcpy (doff, Aoff, (size_t) (nz + 1) * sizeof (int64_t), cudaMemcpyHostToDevice));
        CU (cudaMemcpy (dsg, Asign, (size_t) nz, cudaMemcpyHostToDevice));
        const int per = 256 / CH;
        dim3 grid ((nz + per - 1) / per, S / CH);
        k_residues<<<grid, 256, 0, F->st>>> (nz, CH, dl, doff, dsg, F->tab->p, F->tab->ninv, F->tab->r2, F->dA);
        g_launches++;
        CU (cudaGetLastError ());
        CU (cudaStreamSynchronize (F->st));
        pool_free (dl); pool_free (doff); pool_free (dsg);
        g_h2d_bytes += (double) nl * 4 + (double) (nz + 1) * 8 + (double) nz * 5 + (double) (n + 1) * 4;
    }
    CU (cudaEventRecord (F->ev_start, F->st));
    return SLIPCU_OK;
}

template <int CH, int CPT>
static cudaError_t launch_tri (const TriArgs &a, dim3 grid, size_t smem, cudaStream_t st)
{
    if (a.x_in_smem) k_trisolve<CH, true, CPT><<<grid, TRI_THREADS * 4 / CPT, smem, st>>> (a);
    else k_trisolve<CH, false, CPT><<<grid, TRI_THREADS * 4 / CPT, smem, st>>> (a);
    return cudaGetLastError ();
}
static size_t tri_smem_bytes (int CH, int cnt, bool x_in_smem)
{
    if (CH == 4) return TriSmem<4>::total (cnt, x_in_smem);
    if (CH == 8) return TriSmem<8>::total (cnt, x_in_smem);
    if (CH == 16) return TriSmem<16>::total (cnt, x_in_smem);
    return TriSmem<32>::total (cnt, x_in_smem);
}
static cudaError_t launch_tri_any (int CH, int cpt, const TriArgs &a, dim3 grid, size_t smem, cudaStream_t st)
{
    g_launches++; g_tri_launches++;
    if (cpt == 2)
    {
        if (CH == 4) return launch_tri<4, 2> (a, grid, smem, st);
        if (CH == 8) return launch_tri<8, 2> (a, grid, smem, st);
        if (CH == 16) return launch_tri<16, 2> (a, grid, smem, st);
        return launch_tri<32, 2> (a, grid, smem, st);
    }
    if (CH == 4) return launch_tri<4, 4> (a, grid, smem, st);
    if (CH == 8) return launch_tri<8, 4> (a, grid, smem, st);
    if (CH == 16) return launch_tri<16, 4> (a, grid, smem, st);
    return launch_tri<32, 4> (a, grid, smem, st);
}

static CommitArgs make_commit_args (slipcu_factor *F, int k, int slot)
{
    const HostCol &hc = F->cols[k];
    const Tables &T = *F->tab;
    CommitArgs c;
    c.k = k; c.S = F->S; c.CH = F->CH; c.slot = slot;
    c.d.base = hc.base; c.d.rows = hc.rows; c.d.cnt = hc.cnt; c.d.nU = hc.nU; c.d.pivslot = slot; c.d.pad = 0;
    c.d.mag = F->mag_on ? hc.mag : nullptr;
    c.desc = F->desc; c.rho = F->rho; c.invrho = F->invrho; c.p = T.p; c.ninv = T.ninv; c.one = T.one;
    c.bad = F->bad; c.rho_mag = F->mag_on ? F->rho_mag : nullptr;
    return c;
}
// a pivot whose commit has not been enqueued yet (it normally rides in the next column's k_prep)
static int flush_commit (slipcu_factor *F)
{
    if (F->pending_commit.k < 0) return SLIPCU_OK;
    const CommitArgs c = make_commit_args (F, F->pending_commit.k, F->pending_commit.slot);
    F->pending_commit.k = -1;
    k_pivot_commit<<<(F->S + 255) / 256, 256, 0, F->st>>> (c);
    g_launches++;
    CU (cudaGetLastError ());
    if (debug_check ("k_pivot_commit", F->st)) return fail (SLIPCU_CUDA_ERROR, "k_pivot_commit", "debug");
    if (!F->rows_are_positions) CU (cudaEventRecord (F->ev_commit, F->st));
    return SLIPCU_OK;
}

// symbolic pre-pass on the device: pos[], the slot lists and the step table of a column whose
// pattern (rows), step positions (upos) and slot-list offsets (uoff, nU+1 entries) are on the device
static int prepare_steps (slipcu_factor *F, WorkCtx &w, int cnt, int nU, const int32_t *rows, const int32_t *upos,
                          const int32_t *uoff, const int32_t *uchunk, int total, int nchunks,
                          int packet_ints = 0, bool with_commit = false, int upart = 0)
{
    if ((size_t) nchunks > w.chunks_cap)
    {
        CU (cudaStreamSynchronize (w.st));
        pool_free (w.chunks); w.chunks = nullptr;
        size_t want = std::max ((size_t) nchunks, std::max<size_t> (w.chunks_cap * 2, 1024));
        CU (pool_alloc_t (&w.chunks, want * sizeof (ChunkInfo)));
        w.chunks_cap = want;
    }
    if ((size_t) total > w.slots_cap)
    {
        CU (cudaStreamSynchronize (w.st));
        pool_free (w.slots); w.slots = nullptr;
        size_t want = std::max ((size_t) total, w.slots_cap * 2);
        CU (pool_alloc_t (&w.slots, want * sizeof (int32_t)));
        w.slots_cap = want;
    }
    if ((size_t) nU > w.steps_cap)
    {
        CU (cudaStreamSynchronize (w.st));
        pool_free (w.steps); w.steps = nullptr;
        size_t want = std::max ((size_t) nU, std::max<size_t> (w.steps_cap * 2, 256));
        CU (pool_alloc_t (&w.steps, want * sizeof (StepInfo)));
        w.steps_cap = want;
    }
    ScopedTimer tm_other (F, &g_other_ms, w.st);
    if (packet_ints > 0)
    {   // the packet sits in mapped host memory: one kernel copies it to `rows`, fills pos[] and
        // (main stream only) commits the pivot that is still pending
        PrepArgs pa; memset (&pa, 0, sizeof (pa));
        pa.cnt = cnt; pa.pk_ints = packet_ints; pa.copy_blocks = (packet_ints + 255) / 256;
        pa.h_packet = w.h_packet_dev + (size_t) w.pk_cur * w.pk_stride; pa.d_packet = const_cast<int32_t *> (rows); pa.pos = w.pos;
        int blocks = pa.copy_blocks;
        if (with_commit && F->pending_commit.k >= 0)
        {
            pa.has_commit = 1;
            pa.c = make_commit_args (F, F->pending_commit.k, F->pending_commit.slot);
            F->pending_commit.k = -1;
            blocks += (F->S + 255) / 256;
        }
        k_prep<<<blocks, 256, 0, w.st>>> (pa);
        g_launches++;
        CU (cudaGetLastError ());
        if (debug_check ("k_prep", w.st)) return fail (SLIPCU_CUDA_ERROR, "k_prep", "debug");
        if (pa.has_commit && !F->rows_are_positions) CU (cudaEventRecord (F->ev_commit, w.st));
        CU (cudaEventRecord (w.pk_ev[w.pk_cur], w.st));
        w.pk_busy[w.pk_cur] = true;
    }
    else
    {
        k_setpos<<<(cnt + 255) / 256, 256, 0, w.st>>> (cnt, rows, w.pos);
        g_launches++;
        CU (cudaGetLastError ());
        if (debug_check ("k_setpos", w.st)) return fail (SLIPCU_CUDA_ERROR, "k_setpos", "debug");
    }
    if (nU > 0 && total > 0 && nchunks > 0)
    {
        k_slots<<<nchunks, 32 * tri_bank_groups (F->CH), 0, w.st>>> (nU, F->CH, cnt, upos, uoff, uchunk, F->desc, w.pos, w.slots, w.steps, w.chunks, upart, F->slot_sort);
        g_launches++;
        CU (cudaGetLastError ());
        if (debug_check ("k_slots", w.st)) return fail (SLIPCU_CUDA_ERROR, "k_slots", "debug");
    }
    return SLIPCU_OK;
}

// fills the launch geometry of a triangular solve; returns the dynamic shared memory size
static int tri_geometry (slipcu_factor *F, TriArgs &a, size_t *smem)
{
    a.x_in_smem = !F->x_global;
    size_t need = tri_smem_bytes (F->CH, a.cnt, a.x_in_smem);
    if (need + 1024 > F->smem_limit && a.x_in_smem)
    {   // pattern too long for shared memory: x stays in the (L2-resident) output region
        a.x_in_smem = 0;
        need = tri_smem_bytes (F->CH, a.cnt, false);
    }
    if (need + 1024 > F->smem_limit) return fail (SLIPCU_BAD_INPUT, "tri_geometry", "pattern too large for shared memory");
    *smem = need;
    return SLIPCU_OK;
}

static int check_channels (slipcu_factor *F);


// mixed-radix digits + sign of entries e0..e0+ne-1 of a residue region (digit row = entry index)
static int run_garner (slipcu_factor *F, const u32 *base, int region_cnt, int e0, int ne, int s, int8_t *sign,
                       const ScanArgs *scan = nullptr)
{
    if (ne <= 0) return SLIPCU_OK;
    const Tables &T = *F->tab;
    GarnerArgs g;
    g.scan_on = scan ? 1 : 0; g.done_ctr = F->done_ctr;
    if (scan) g.sc = *scan; else memset (&g.sc, 0, sizeof (g.sc));
    g.cnt = region_cnt; g.e0 = e0; g.ne = ne; g.s = s; g.CH = F->CH; g.S = T.S;
    g.base = base; g.dig = F->dig; g.dstride = (size_t) F->S + 4; g.topd = F->topd; g.sign = sign;
    g.p = T.p; g.ninv = T.ninv; g.C = T.C; g.invB = T.invB;
    ScopedTimer tm (F, &g_recon_ms);
    if (s <= GS_MAX && F->garner_mode >= 2)
        k_garner_small<<<(ne + 3) / 4, 128, (size_t) s * s * sizeof (u32), F->wst>>> (g);
    else if (F->garner_mode == 0 || s < 64)
        k_garner<<<(ne + 3) / 4, 128, 0, F->wst>>> (g);
    else
    {
        // pick the CTA width so that one tile covers s when possible (W warps x 5 blocks x 32 digits)
        const int blocks = (s + 31) / 32;
        int W = (blocks + 4) / 5;
        W = std::max (4, std::min (16, W));      // >= E warps: the epilogue uses one warp per entry
        const int Wf = std::max (6, (blocks + 5) / 6);          // dataflow variant: 6 blocks per warp
        if (F->garner_mode >= 2 && Wf <= 12)
        {
            // A CTA reconstructs E entries; its run time grows with E (measured, relative, in
            // garner_cost) and the grid runs in waves of one CTA per SM: take the E with the least
            // waves x cost.  (Cutting a launch into pieces with different E was measured slower:
            // separate launches do not overlap their tails.)
            static const int garner_cost[7] = { 0, 180, 260, 345, 420, 518, 625 };
            int E = 1;
            if (F->garner_e) E = std::max (1, std::min (6, F->garner_e));
            else
            {
                long bestc = -1;
                for (int e = 1; e <= 6; ++e)
                {
                    const long waves = ((ne + e - 1) / e + F->sms - 1) / F->sms;
                    const long c = waves * garner_cost[e];
                    if (bestc < 0 || c < bestc) { bestc = c; E = e; }
                }
            }
            const size_t fsm = (size_t) blocks * E * 32 * sizeof (u32) + (size_t) blocks * sizeof (int) + (size_t) Wf * 4096 + 16;
            const int grid = (ne + E - 1) / E;
            switch (E)
            {
                case 1: k_garner_flow<1, 6><<<grid, Wf * 32, fsm, F->wst>>> (g); break;
                case 2: k_garner_flow<2, 6><<<grid, Wf * 32, fsm, F->wst>>> (g); break;
                case 3: k_garner_flow<3, 6><<<grid, Wf * 32, fsm, F->wst>>> (g); break;
                case 4: k_garner_flow<4, 6><<<grid, Wf * 32, fsm, F->wst>>> (g); break;
                case 5: k_garner_flow<5, 6><<<grid, Wf * 32, fsm, F->wst>>> (g); break;
                default: k_garner_flow<6, 6><<<grid, Wf * 32, fsm, F->wst>>> (g); break;
            }
        }
        else
            k_garner_tiled<4, 5><<<(ne + 3) / 4, W * 32, 0, F->wst>>> (g);
    }
    g_launches++;
    CU (cudaGetLastError ());
    if (debug_check ("k_garner", F->wst)) return fail (SLIPCU_CUDA_ERROR, "k_garner", "debug");
    g_recon_mac += (double) ne * ((double) s * s * 0.5);
    return SLIPCU_OK;
}

// positional limbs of digit rows e0..e0+ne-1 into limb rows out0.. of (limbs, nl)
static int run_limbs (slipcu_factor *F, int e0, int ne, int out0, int stride, int s, u32 *limbs, int32_t *nl)
{
    if (ne <= 0) return SLIPCU_OK;
    const Tables &T = *F->tab;
    LimbArgs l;
    l.e0 = e0; l.ne = ne; l.out0 = out0; l.stride = stride; l.LB = T.LB;
    l.dig = F->dig; l.dstride = (size_t) F->S + 4; l.topd = F->topd; l.Bpos = T.Bpos;
    l.limbs = limbs; l.nl = nl;
    ScopedTimer tm (F, &g_recon_ms);
    if (ne >= 64) k_limbs<4><<<((ne + 3) / 4 + 3) / 4, 128, 0, F->wst>>> (l);
    else k_limbs<1><<<(ne + 3) / 4, 128, 0, F->wst>>> (l);
    g_launches++;
    CU (cudaGetLastError ());
    if (debug_check ("k_limbs", F->wst)) return fail (SLIPCU_CUDA_ERROR, "k_limbs", "debug");
    g_recon_mac += (double) ne * ((double) s * s * 0.5);
    return SLIPCU_OK;
}

// approximate magnitudes of the candidates (slots nU .. cnt-1) with W words, then the selection
static int run_frac (slipcu_factor *F, const HostCol &hc, int cnt, int nU, int s, int mode, int diag_slot, int W)
{
    const Tables &T = *F->tab;
    const int ne = cnt - nU;
    if ((size_t) ne > F->frac_rows)
    {
        CU (cudaStreamSynchronize (F->st));
        pool_free (F->frackey); F->frackey = nullptr;
        const size_t want = std::max<size_t> ((size_t) ne, std::min<size_t> ((size_t) F->n, std::max<size_t> (64, F->frac_rows * 2)));
        CU (pool_alloc_t (&F->frackey, want * sizeof (FracKey)));
        F->frac_rows = want;
    }
    FracArgs a;
    a.cnt = cnt; a.e0 = nU; a.ne = ne; a.s = s; a.CH = F->CH; a.S = T.S; a.W = W;
    a.base = hc.base; a.p = T.p; a.ninv = T.ninv; a.minv = T.Minv + (size_t) (s - 1) * T.S; a.urec = T.Urec;
    a.key = F->frackey;
    {
        ScopedTimer tm (F, &g_recon_ms);
        int NW = (W + 31) / 32;
        int E = ne >= 4 * F->sms ? 4 : (ne >= F->sms ? 2 : 1);            // entries per CTA (they share the table loads)
        if (E == 4 && NW == 4) E = 2;                                     // keeps the partial sums within 48 KB
        if (NW > 4) { E = 1; NW = NW <= 6 ? 6 : 8; }                      // long fractions: one entry per CTA
        const int grid = (ne + E - 1) / E;
        const size_t fsm = (size_t) FRAC_WARPS * E * NW * 32 * 3 * sizeof (u32) + (size_t) FRAC_WARPS * E * sizeof (int);
#define FRAC_LAUNCH(EE, NN) k_fraccrt<EE, NN><<<grid, FRAC_WARPS * 32, fsm, F->st>>> (a)
        if (NW == 6) FRAC_LAUNCH (1, 6);
        else if (NW == 8) FRAC_LAUNCH (1, 8);
        else if (E == 4) { if (NW == 1) FRAC_LAUNCH (4, 1); else if (NW == 2) FRAC_LAUNCH (4, 2); else if (NW == 3) FRAC_LAUNCH (4, 3); else FRAC_LAUNCH (4, 4); }
        else if (E == 2) { if (NW == 1) FRAC_LAUNCH (2, 1); else if (NW == 2) FRAC_LAUNCH (2, 2); else if (NW == 3) FRAC_LAUNCH (2, 3); else FRAC_LAUNCH (2, 4); }
        else { if (NW == 1) FRAC_LAUNCH (1, 1); else if (NW == 2) FRAC_LAUNCH (1, 2); else if (NW == 3) FRAC_LAUNCH (1, 3); else FRAC_LAUNCH (1, 4); }
#undef FRAC_LAUNCH
        g_launches++;
        CU (cudaGetLastError ());
        if (debug_check ("k_fraccrt", F->st)) return fail (SLIPCU_CUDA_ERROR, "k_fraccrt", "debug");
        g_recon_mac += (double) ne * (double) s * (double) W;
    }
    FracSel q;
    q.ne = ne; q.nU = nU; q.mode = mode; q.diag_slot = diag_slot; q.W = W;
    q.key = F->frackey; q.bad = F->bad; q.info = F->d_info; q.k = F->cur_launch; q.run_flags = F->run_flags; q.seq = ++F->info_seq;
    {
        ScopedTimer tm (F, &g_other_ms);
        k_fracselect<<<1, 256, 0, F->st>>> (q);
        g_launches++;
        CU (cudaGetLastError ());
        if (debug_check ("k_fracselect", F->st)) return fail (SLIPCU_CUDA_ERROR, "k_fracselect", "debug");
    }
    F->fq.cnt = cnt; F->fq.nU = nU; F->fq.s = s; F->fq.mode = mode; F->fq.diag_slot = diag_slot; F->fq.W = W;
    return SLIPCU_OK;
}
static int frac_word_cap (int s) { return std::min (FRAC_WMAX, s + 2); }

static int alloc_column (slipcu_factor *F, HostCol &hc, int cnt, int s)
{
    hc.cnt = cnt; hc.s = s;
    hc.stride = (s + 1) & ~1;
    hc.base = (u32 *) F->resid.alloc ((size_t) cnt * F->S * sizeof (u32));
    hc.sign = (int8_t *) F->ints.alloc ((size_t) cnt);
    if (F->mag_on)
    {
        hc.mag = (int32_t *) F->ints.alloc ((size_t) cnt * sizeof (int32_t));
        if (!hc.mag) return fail (SLIPCU_OUT_OF_MEMORY, "alloc_column", "device memory exhausted");
    }
    if (F->keep_positional)
    {
        hc.limbs = (u32 *) F->limbs.alloc ((size_t) cnt * hc.stride * sizeof (u32));
        hc.nl = (int32_t *) F->ints.alloc ((size_t) cnt * sizeof (int32_t));
    }
    if (!hc.base || !hc.sign || (F->keep_positional && (!hc.limbs || !hc.nl)))
        return fail (SLIPCU_OUT_OF_MEMORY, "alloc_column", "device memory exhausted");
    return SLIPCU_OK;
}

static ScanArgs make_scan_args (slipcu_factor *F, const HostCol &hc, int cnt, int nU, int mode, int diag_slot)
{
    ScanArgs a;
    a.cnt = cnt; a.nU = nU; a.mode = mode; a.diag_slot = diag_slot;
    a.dig = F->dig; a.ds = (size_t) F->S + 4; a.topd = F->topd; a.sign = hc.sign; a.bad = F->bad; a.info = F->d_info;
    a.mag = F->mag_on ? hc.mag : nullptr; a.cum_ub = F->tab->cum_ub; a.bound = F->mag_on ? F->bound : nullptr;
    a.measured = F->measured;
    a.k = F->cur_launch; a.run_flags = F->run_flags;
    a.seq = ++F->info_seq;
    return a;
}

static int run_exact_scan (slipcu_factor *F, const HostCol &hc, int cnt, int nU, int mode, int diag_slot)
{
    ScopedTimer tm_scan (F, &g_other_ms);
    k_pivot_scan<<<1, 256, 0, F->st>>> (make_scan_args (F, hc, cnt, nU, mode, diag_slot));
    g_launches++;
    CU (cudaGetLastError ());
    if (debug_check ("k_pivot_scan", F->st)) return fail (SLIPCU_CUDA_ERROR, "k_pivot_scan", "debug");
    return SLIPCU_OK;
}

// uploads a pattern packet (rows, U positions, slot-list offsets, chunk offsets of the steps from
// first_step on) and runs the symbolic pre-pass; returns the device copy and the chunk count
static int upload_pattern (slipcu_factor *F, WorkCtx &w, int cnt, int nU, const int32_t *rows, const int32_t *upos,
                           int first_step, int32_t **dev_rows, int *nchunks_out)
{
    const int CH = F->CH;
    const int pk = w.pk_next;
    w.pk_next = (w.pk_next + 1) % PK_RING; w.pk_cur = pk;
    if (w.pk_busy[pk]) { CU (cudaEventSynchronize (w.pk_ev[pk])); w.pk_busy[pk] = false; }
    int32_t *staging = w.h_packet + (size_t) pk * w.pk_stride;
    memcpy (staging, rows, (size_t) cnt * sizeof (int32_t));
    if (nU) memcpy (staging + cnt, upos, (size_t) nU * sizeof (int32_t));
    int32_t *uoff = staging + cnt + nU;
    int32_t *uchunk = uoff + nU + 1;
    int64_t total = 0, nchunks = 0;
    const int chunk_rows = tri_chunk_rows (CH);
    for (int u = 0; u < nU; ++u)
    {
        uoff[u] = (int32_t) total; uchunk[u] = (int32_t) nchunks;
        if (u < first_step) continue;                    // already applied by the bulk part
        const HostCol &lj = F->cols[upos[u]];
        const int len = lj.cnt - lj.nU;
        total += tri_slot_extent (len, CH);
        nchunks += (len + chunk_rows - 1) / chunk_rows;
        if (total > INT32_MAX) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_column", "column has too many updates");
    }
    uoff[nU] = (int32_t) total; uchunk[nU] = (int32_t) nchunks;
    const size_t pk_ints = (size_t) cnt + 3 * (size_t) nU + 2;
    int32_t *d = (int32_t *) F->ints.alloc (pk_ints * sizeof (int32_t));
    if (!d) return fail (SLIPCU_OUT_OF_MEMORY, "slipcu_factor_column", "device memory exhausted");
    g_h2d_bytes += (double) pk_ints * sizeof (int32_t);
    int rc = prepare_steps (F, w, cnt, nU, d, d + cnt, d + cnt + nU, d + cnt + 2 * nU + 1, (int) total, (int) nchunks,
                            (int) pk_ints, &w == &F->mc);
    if (rc) return rc;
    *dev_rows = d; *nchunks_out = (int) nchunks;
    return SLIPCU_OK;
}

// Bulk part of column `col` (to become column k of the factorization) in lookahead slot `slot`:
// eliminates with every pivot that is already committed, on the pattern reachable through those
// columns, and leaves the normalised vector in the slot's buffer.  Runs on the slot's own stream
// beside the column in flight (it only reads finished columns), so that several columns advance at
// once when the channel count is too small to fill the GPU with one.
extern "C" int slipcu_factor_spec_launch (slipcu_factor *F, int slot, int k, int col, int cnt, int nU,
                                          const int32_t *rows, const int32_t *upos)
{
    if (!F || slot < 0 || slot >= SPEC_SLOTS || k < 1 || k >= F->n || cnt <= 0 || cnt > F->n || nU < 0 || nU > cnt || !rows)
        return fail (SLIPCU_BAD_INPUT, "slipcu_factor_spec_launch", "bad argument");
    USE_DEVICE (F);
    const Tables &T = *F->tab;
    const int S = F->S, CH = F->CH;
    SpecSlot &sl = F->spec[slot];
    if (!sl.w.st)
    {
        cudaStream_t st = nullptr;
        CU (pooled_stream (&st));
        int rc0 = init_workctx (sl.w, F->n, st);
        if (rc0) return rc0;
        CU (pooled_event (&sl.done, false));
        CU (pooled_event (&sl.consumed, false));
    }
    WorkCtx &w = sl.w;
    // the slot's previous vector may still be read by the column launch that consumed it
    if (sl.consumed_pending) { CU (cudaStreamWaitEvent (w.st, sl.consumed, 0)); sl.consumed_pending = false; }
    // every pivot committed so far is visible to this stream
    CU (cudaStreamWaitEvent (w.st, F->ev_commit, 0));
    const size_t words = (size_t) cnt * S;
    if (words > sl.words)
    {
        CU (cudaStreamSynchronize (w.st));
        pool_free (sl.buf); sl.buf = nullptr;
        const size_t want = std::max (words, std::min ((size_t) F->n * S, sl.words * 2));
        CU (pool_alloc_t (&sl.buf, want * sizeof (u32)));
        sl.words = want;
    }
    if (F->mag_on && (size_t) cnt > sl.mag_cap)
    {
        CU (cudaStreamSynchronize (w.st));
        pool_free (sl.mag); sl.mag = nullptr;
        const size_t want = std::max ((size_t) cnt, std::min ((size_t) F->n, sl.mag_cap * 2));
        CU (pool_alloc_t (&sl.mag, want * sizeof (int32_t)));
        sl.mag_cap = want;
    }
    int32_t *d = nullptr; int nchunks = 0;
    int rc = upload_pattern (F, w, cnt, nU, rows, upos, 0, &d, &nchunks);
    if (rc) return rc;
    TriArgs a; memset (&a, 0, sizeof (a));
    a.k = k; a.S = S; a.cnt = cnt; a.nU = nU;
    a.rows = d; a.steps = w.steps; a.chunks = w.chunks; a.slots = w.slots;
    a.src = F->dA; a.src_total = F->nz; a.src_first = F->hAp[col]; a.src_step = 1;
    a.src_cnt = F->hAp[col + 1] - F->hAp[col]; a.src_rows = F->dAi + F->hAp[col];
    a.src_y_stride = 0;
    a.out = sl.buf; a.out_y_stride = 0; a.out_cb_stride = (size_t) cnt * CH;
    a.rho = F->rho; a.invrho = F->invrho;
    a.p = T.p; a.ninv = T.ninv; a.pos = w.pos;
    a.nchunks = nchunks; a.upos = d + cnt; a.publish = 0; a.u0 = 0;
    a.mag_on = F->mag_on; a.mag_src = F->mag_on ? F->Amag + F->hAp[col] : nullptr; a.mag_out = sl.mag;
    a.rho_mag = F->rho_mag; a.bound_out = nullptr;
    size_t smem = 0;
    rc = tri_geometry (F, a, &smem);
    if (rc) return rc;
    a.smem_bytes = (int) smem;
    {
        ScopedTimer tm (F, &g_tri_ms, w.st);
        CU (launch_tri_any (CH, F->cpt, a, dim3 (S / CH + (F->mag_on ? 1 : 0), 1), smem, w.st));
        if (debug_check ("k_trisolve(bulk part)", w.st)) return fail (SLIPCU_CUDA_ERROR, "k_trisolve", "debug");
    }
    CU (cudaEventRecord (sl.done, w.st));
    double upd = 0;
    for (int u = 0; u < nU; ++u) { const HostCol &lj = F->cols[upos[u]]; upd += (double) (lj.cnt - lj.nU - 1); }
    g_tri_bytes += upd * (double) S * 4.0;
    g_tri_modmul += upd * (double) S * 4.0;
    sl.rows = d; sl.cnt = cnt; sl.col = k; sl.nU = nU;
    return SLIPCU_OK;
}

extern "C" int slipcu_factor_column_launch (slipcu_factor *F, int k, int col, int cnt, int nU,
                                            const int32_t *rows, const int32_t *upos, int recon_channels,
                                            int scheme, int diag_slot, int spec_slot)
{
    if (!F || k < 0 || k >= F->n || cnt <= 0 || cnt > F->n || nU < 0 || nU >= cnt || !rows || spec_slot >= SPEC_SLOTS)
        return fail (SLIPCU_BAD_INPUT, "slipcu_factor_column", "bad argument");
    USE_DEVICE (F);
    // a bulk part exists for this column: start from its vector, apply the remaining steps
    SpecSlot *sl = (spec_slot >= 0 && F->spec[spec_slot].col == k) ? &F->spec[spec_slot] : nullptr;
    const bool from_spec = sl != nullptr;
    const int first_step = from_spec ? sl->nU : 0;
    if (from_spec && first_step > nU) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_column", "bulk part does not match");
    const Tables &T = *F->tab;
    const int S = F->S, CH = F->CH;
    WorkCtx &w = F->mc;
    double tw = wall_s ();
    int s = std::min (std::max (recon_channels, 1), S);
    HostCol &hc = F->cols[k];
    hc.nU = nU;
    int rc = alloc_column (F, hc, cnt, s);
    if (rc) return rc;
    rc = ensure_digits (F, (size_t) cnt);
    if (rc) return rc;
    g_hw[0] += wall_s () - tw; tw = wall_s ();
    // packet: pattern rows, pivot positions of the U part, slot-list offsets (padded to 4)
    int nchunks = 0;
    rc = upload_pattern (F, w, cnt, nU, rows, upos, first_step, &hc.rows, &nchunks);
    if (rc) return rc;

    g_hw[1] += wall_s () - tw; tw = wall_s ();
    TriArgs a; memset (&a, 0, sizeof (a));
    a.k = k; a.S = S; a.cnt = cnt; a.nU = nU;
    a.rows = hc.rows; a.steps = w.steps; a.chunks = w.chunks; a.slots = w.slots;
    if (from_spec)
    {   // the vector of the bulk part, row by row into the final slots
        CU (cudaStreamWaitEvent (F->st, sl->done, 0));
        a.src = sl->buf; a.src_total = sl->cnt; a.src_first = 0; a.src_step = 1;
        a.src_cnt = sl->cnt; a.src_rows = sl->rows;
        a.mag_src = sl->mag;
    }
    else
    {
        a.src = F->dA; a.src_total = F->nz; a.src_first = F->hAp[col]; a.src_step = 1;
        a.src_cnt = F->hAp[col + 1] - F->hAp[col]; a.src_rows = F->dAi + F->hAp[col];
        a.mag_src = F->mag_on ? F->Amag + F->hAp[col] : nullptr;
    }
    a.src_y_stride = 0;
    a.out = hc.base; a.out_y_stride = 0; a.out_cb_stride = (size_t) cnt * CH;
    a.rho = F->rho; a.invrho = F->invrho;
    a.p = T.p; a.ninv = T.ninv; a.pos = w.pos;
    a.nchunks = nchunks; a.upos = hc.rows + cnt; a.publish = 1; a.u0 = first_step;
    a.mag_on = F->mag_on; a.mag_out = hc.mag; a.rho_mag = F->rho_mag; a.bound_out = F->bound;
    size_t smem = 0;
    rc = tri_geometry (F, a, &smem);
    if (rc) return rc;
    a.smem_bytes = (int) smem;
    {
        ScopedTimer tm (F, &g_tri_ms);
        CU (launch_tri_any (CH, F->cpt, a, dim3 (S / CH + (F->mag_on ? 1 : 0), 1), smem, F->st));
        if (debug_check ("k_trisolve(column)", F->st)) return fail (SLIPCU_CUDA_ERROR, "k_trisolve", "debug");
    }
    if (from_spec)
    {
        CU (cudaEventRecord (sl->consumed, F->st));
        sl->consumed_pending = true; sl->col = -1;
    }
    g_hw[2] += wall_s () - tw; tw = wall_s ();
    {   // algorithmic work of this launch
        double upd = 0;
        for (int u = first_step; u < nU; ++u) { const HostCol &lj = F->cols[upos[u]]; upd += (double) (lj.cnt - lj.nU - 1); }
        g_tri_bytes += (upd + (double) cnt) * (double) S * 4.0;
        g_tri_modmul += upd * (double) S * 4.0;
        if (const char *tf = getenv ("SLIP_B200_TRACE_FILE"))
        {   // per-launch algorithmic bytes, to set beside an ncu capture of the same launch
            if (FILE *fp = fopen (tf, "a"))
            {
                fprintf (fp, "%d %d %d %.0f %.0f\n", k, cnt, nU, upd, (upd + (double) cnt) * (double) S * 4.0);
                fclose (fp);
            }
        }
    }
    // exact values: candidates always (pivot scan); the U part only if the factors go to the host,
    // and then on the side stream, off the path to the pivot
    const bool ov = F->keep_positional && F->overlap;
    int bufi = 0;
    if (ov)
    {
        CU (cudaEventRecord (F->ev_tri, F->st));
        bufi = F->seq++ & 1;
        F->dig = F->digbuf[bufi]; F->topd = F->topdbuf[bufi];
        if (F->side_pending[bufi]) { CU (cudaStreamWaitEvent (F->st, F->ev_side[bufi], 0)); F->side_pending[bufi] = false; }
    }
    const int mode = (scheme == 2) ? 2 : ((scheme == 4 || scheme == 5) ? 1 : 0);
    F->frac_col = -1;
    F->cur_launch = k;
    const bool single = (cnt - nU == 1);             // no search needed: exact zero test and size only
    if (single && F->nowait_singles && k < F->n - 1 && !F->keep_positional && !F->measured && !F->mag_on)
    {   // one candidate, a-priori channel count, factors stay on the device: its pivot needs a zero
        // test and nothing else (the caller does not wait for this column)
        k_zero_test<<<1, 256, 0, F->st>>> (k, S, CH, cnt, nU, hc.base, F->run_flags);
        g_launches++;
        CU (cudaGetLastError ());
        if (debug_check ("k_zero_test", F->st)) return fail (SLIPCU_CUDA_ERROR, "k_zero_test", "debug");
        g_hw[3] += wall_s () - tw; tw = wall_s ();
    }
    else if (F->frac && !F->keep_positional && mode != 2 && s >= F->frac_min_s && !single)
    {   // magnitudes only: no digits unless the choice turns out to be too close to call
        const int W = std::min (std::max (F->fracW, 8), frac_word_cap (s));
        rc = run_frac (F, hc, cnt, nU, s, mode, diag_slot, W);
        if (rc) return rc;
        F->frac_col = k;
        g_hw[3] += wall_s () - tw; tw = wall_s ();
    }
    else
    {
        // reconstruction of the candidates with the pivot scan in its last CTA
        const int e0 = (F->keep_positional && !ov) ? 0 : nU;
        const ScanArgs sc = make_scan_args (F, hc, cnt, nU, mode, diag_slot);
        rc = run_garner (F, hc.base, cnt, e0, cnt - e0, s, hc.sign, &sc);
        if (rc) return rc;
        if (ov) CU (cudaEventRecord (F->ev_gl, F->st));
        g_hw[3] += wall_s () - tw; tw = wall_s ();
    }
    CU (cudaEventRecord (F->ev, F->st));
    if (ov)
    {   // side stream: digits of the U part, then positional limbs of the whole column
        F->wst = F->st2;
        CU (cudaStreamWaitEvent (F->st2, F->ev_tri, 0));
        rc = run_garner (F, hc.base, cnt, 0, nU, s, hc.sign);
        if (rc == SLIPCU_OK)
        {
            cudaStreamWaitEvent (F->st2, F->ev_gl, 0);
            rc = run_limbs (F, 0, cnt, 0, hc.stride, s, hc.limbs, hc.nl);
        }
        F->wst = F->st;
        if (rc) return rc;
        CU (cudaEventRecord (F->ev_side[bufi], F->st2));
        F->side_pending[bufi] = true;
    }
    else if (F->keep_positional)
    {   // positional limbs of the whole column; overlaps the host's pivot decision
        rc = run_limbs (F, 0, cnt, 0, hc.stride, s, hc.limbs, hc.nl);
        if (rc) return rc;
    }
    g_hw[4] += wall_s () - tw;
    F->cur = k;
    return SLIPCU_OK;
}

extern "C" void slipcu_factor_nowait_singles (slipcu_factor *F, int on) { if (F) F->nowait_singles = on ? 1 : 0; }

extern "C" int slipcu_factor_column_wait (slipcu_factor *F, slipcu_pivot_info *info)
{
    if (!F || !info || F->cur < 0) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_column_wait", "bad argument");
    USE_DEVICE (F);
    double tw = wall_s ();
    {   // the record is complete when its sequence word (written last, after a system-wide fence) shows
        // the number of the launch: polled here, which returns microseconds before an event wait would
        const int32_t want = F->info_seq;
        volatile int32_t *sq = &F->h_info->seq;
        bool seen = false;
        // (several sessions in flight on host threads: spinning threads get in each other's way --
        // 128 small systems on 8 threads took twice as long -- so those sleep on the event)
        const int max_spins = g_live_sessions.load () <= 1 ? (1 << 22) : 0;
        for (int spins = 0; spins < max_spins; ++spins)
        {
            if (*sq == want) { seen = true; break; }
            if ((spins & 2047) == 2047 && cudaEventQuery (F->ev) != cudaErrorNotReady) break;      // finished, or failed
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause ();
#endif
        }
        if (!seen) CU (cudaEventSynchronize (F->ev));
        else std::atomic_thread_fence (std::memory_order_acquire);
    }
    g_d2h_bytes += sizeof (slipcu_pivot_info);
    *info = *F->h_info;
    if (F->frac_col == F->cur)
    {   // approximate search: accept a proven choice, add words if the entries are smaller than
        // the words resolve, otherwise (ties, near ties) run the exact scan
        HostCol &hc = F->cols[F->cur];
        const int cap = frac_word_cap (F->fq.s);
        F->frac_cols++;
        while (!info->reserved[0] && info->best_slot >= 0 && info->reserved[1] > F->fq.W - 10 && F->fq.W < cap)
        {
            const int W = std::min (cap, std::max (2 * F->fq.W, info->reserved[1] + 12));      // (always the full margin here)
            F->frac_retries++;
            int rc = run_frac (F, hc, F->fq.cnt, F->fq.nU, F->fq.s, F->fq.mode, F->fq.diag_slot, W);
            if (rc) return rc;
            CU (cudaStreamSynchronize (F->st));
            g_d2h_bytes += sizeof (slipcu_pivot_info);
            *info = *F->h_info;
        }
        if (!info->reserved[0])
        {
            F->frac_fallbacks++;
            F->frac_col = -1;
            int rc = run_garner (F, hc.base, F->fq.cnt, F->fq.nU, F->fq.cnt - F->fq.nU, F->fq.s, hc.sign);
            if (rc == SLIPCU_OK) rc = run_exact_scan (F, hc, F->fq.cnt, F->fq.nU, F->fq.mode, F->fq.diag_slot);
            if (rc) return rc;
            CU (cudaStreamSynchronize (F->st));
            g_d2h_bytes += sizeof (slipcu_pivot_info);
            *info = *F->h_info;
        }
        else
        {
            if (info->best_slot >= 0)
                F->fracW = std::min (FRAC_WMAX, std::max (8, info->reserved[1] + F->frac_margin));  // words for the next column
            if (F->frac_verify)
            {   // self-check: the accepted choice must be what the exact scan finds
                const slipcu_pivot_info got = *info;
                int rc = run_garner (F, hc.base, F->fq.cnt, F->fq.nU, F->fq.cnt - F->fq.nU, F->fq.s, hc.sign);
                if (rc == SLIPCU_OK) rc = run_exact_scan (F, hc, F->fq.cnt, F->fq.nU, F->fq.mode, F->fq.diag_slot);
                if (rc) return rc;
                CU (cudaStreamSynchronize (F->st));
                const slipcu_pivot_info &ex = *F->h_info;
                if (ex.best_slot != got.best_slot || ex.diag_eligible != got.diag_eligible
                    || ex.diag_vs_best != got.diag_vs_best || (got.best_slot >= 0 && ex.best_sign != got.best_sign))
                    return fail (SLIPCU_CUDA_ERROR, "approximate pivot search", "accepted choice differs from the exact scan");
                F->frac_col = -1;
            }
        }
        info->reserved[0] = info->reserved[1] = info->reserved[2] = 0;
        // matrices full of ties (small integer entries) send most columns to the exact scan or to a
        // repeat with more words: the exact path alone is then cheaper than a failed approximate
        // search in front of it
        if (F->frac_cols >= 24 && 4 * (F->frac_fallbacks + F->frac_retries) > F->frac_cols) F->frac = 0;
    }
    g_hw[5] += wall_s () - tw;
    // a zero pivot of a column the host did not wait for also raises the bad-channel flag: singular first
    if (info->bad_channel && !info->singular_col) return fail (SLIPCU_BAD_PRIME, "slipcu_factor_column", "channel prime divides a pivot");
    return SLIPCU_OK;
}

extern "C" int slipcu_factor_column (slipcu_factor *F, int k, int col, int cnt, int nU,
                                     const int32_t *rows, const int32_t *upos, int recon_channels,
                                     int scheme, int diag_slot, slipcu_pivot_info *info)
{
    int rc = slipcu_factor_column_launch (F, k, col, cnt, nU, rows, upos, recon_channels, scheme, diag_slot, -1);
    if (rc) return rc;
    return slipcu_factor_column_wait (F, info);
}

extern "C" int slipcu_factor_column_stride (slipcu_factor *F, int k)
{
    return (F && k >= 0 && k < F->n) ? F->cols[k].stride : 0;
}

extern "C" int slipcu_factor_fetch_entry (slipcu_factor *F, int k, int slot, u32 *limbs,
                                          int32_t *nlimbs32, int8_t *sign)
{
    if (!F || k < 0 || k >= F->n) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_fetch_entry", "bad column");
    HostCol &hc = F->cols[k];
    if (slot < 0 || slot >= hc.cnt) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_fetch_entry", "bad slot");
    USE_DEVICE (F);
    if (F->keep_positional)
    {
        if (F->st2) CU (cudaStreamSynchronize (F->st2));       // the column's limbs come from the side stream
        CU (cudaMemcpyAsync (limbs, hc.limbs + (size_t) slot * hc.stride, (size_t) hc.stride * sizeof (u32),
                             cudaMemcpyDeviceToHost, F->st));
        CU (cudaMemcpyAsync (nlimbs32, hc.nl + slot, sizeof (int32_t), cudaMemcpyDeviceToHost, F->st));
    }
    else
    {   // digits of the current column are still in the scratch: convert this one entry on demand
        if (k != F->cur || slot < hc.nU)
            return fail (SLIPCU_BAD_INPUT, "slipcu_factor_fetch_entry", "entry no longer reconstructible");
        if (!F->tmp_limbs || F->tmp_stride < hc.stride)
        {
            pool_free (F->tmp_limbs); pool_free (F->tmp_nl);
            F->tmp_limbs = nullptr; F->tmp_nl = nullptr;
            CU (pool_alloc_t (&F->tmp_limbs, (size_t) (F->S + 2) * sizeof (u32)));
            CU (pool_alloc_t (&F->tmp_nl, sizeof (int32_t)));
            F->tmp_stride = F->S + 2;
        }
        if (F->frac_col == k)
        {   // the column was searched by approximate magnitudes: no digits yet for this entry
            int rcg = run_garner (F, hc.base, hc.cnt, slot, 1, hc.s, hc.sign);
            if (rcg) return rcg;
        }
        int rc = run_limbs (F, slot, 1, 0, hc.stride, hc.s, F->tmp_limbs, F->tmp_nl);
        if (rc) return rc;
        CU (cudaMemcpyAsync (limbs, F->tmp_limbs, (size_t) hc.stride * sizeof (u32), cudaMemcpyDeviceToHost, F->st));
        CU (cudaMemcpyAsync (nlimbs32, F->tmp_nl, sizeof (int32_t), cudaMemcpyDeviceToHost, F->st));
    }
    CU (cudaMemcpyAsync (sign, hc.sign + slot, 1, cudaMemcpyDeviceToHost, F->st));
    CU (cudaStreamSynchronize (F->st));
    g_d2h_bytes += (double) hc.stride * 4 + 5;
    return SLIPCU_OK;
}

extern "C" int slipcu_factor_set_pivot (slipcu_factor *F, int k, int slot)
{
    if (!F || k < 0 || k >= F->n) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_set_pivot", "bad column");
    HostCol &hc = F->cols[k];
    if (slot < hc.nU || slot >= hc.cnt) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_set_pivot", "bad slot");
    USE_DEVICE (F);
    int rc = flush_commit (F);               // at most one pivot is ever pending
    if (rc) return rc;
    F->pending_commit.k = k; F->pending_commit.slot = slot;
    // sessions built from host factors have no next column launch to carry the commit
    if (F->rows_are_positions) return flush_commit (F);
    return SLIPCU_OK;
}

// ------------------------------------------------------------------------------------------------
// slipcu_factor_upload: build a resident session from factors that live on the host (a
// SLIP_LU_solve call whose L and U did not come from this process' last factorization, or whose
// right-hand side needs more channels than the factorization kept).  Columns are given in slot
// order (U part, then L part); rows are FINAL positions.
// ------------------------------------------------------------------------------------------------
extern "C" int slipcu_factor_upload (slipcu_factor **out, int n, int channels, const int32_t *colcnt,
                                     const int32_t *colnU, const int32_t *colpiv, const int32_t *rows,
                                     const u32 *limbs, const int64_t *off, const int8_t *sign)
{
    if (!out || n <= 0 || channels <= 0 || !colcnt || !colnU || !colpiv || !rows || !limbs || !off || !sign)
        return fail (SLIPCU_BAD_INPUT, "slipcu_factor_upload", "bad argument");
    slipcu_factor *F = new slipcu_factor ();
    *out = F;
    int rc = session_common_init (F, n, channels);
    if (rc) return rc;
    F->rows_are_positions = 1;
    F->keep_positional = 0;
    const Tables &T = *F->tab;
    const int S = F->S, CH = F->CH;
    size_t total = 0;
    for (int k = 0; k < n; ++k) total += (size_t) colcnt[k];
    u32 *dl = nullptr; int64_t *doff = nullptr; int8_t *dsg = nullptr; int32_t *drows = nullptr;
    const size_t nl = (size_t) off[total];
    CU (pool_alloc_t (&dl, std::max<size_t> (nl, 1) * sizeof (u32)));
    CU (pool_alloc_t (&doff, (total + 1) * sizeof (int64_t)));
    CU (pool_alloc_t (&dsg, total));
    drows = (int32_t *) F->ints.alloc (total * sizeof (int32_t));
    if (!drows) return fail (SLIPCU_OUT_OF_MEMORY, "slipcu_factor_upload", "device memory exhausted");
    CU (cudaMemcpyAsync (dl, limbs, nl * sizeof (u32), cudaMemcpyHostToDevice, F->st));
    CU (cudaMemcpyAsync (doff, off, (total + 1) * sizeof (int64_t), cudaMemcpyHostToDevice, F->st));
    CU (cudaMemcpyAsync (dsg, sign, total, cudaMemcpyHostToDevice, F->st));
    CU (cudaMemcpyAsync (drows, rows, total * sizeof (int32_t), cudaMemcpyHostToDevice, F->st));
    size_t e = 0;
    for (int k = 0; k < n; ++k)
    {
        HostCol &hc = F->cols[k];
        const int cnt = colcnt[k];
        hc.cnt = cnt; hc.nU = colnU[k]; hc.s = S; hc.stride = 0;
        hc.rows = drows + e;
        hc.base = (u32 *) F->resid.alloc ((size_t) cnt * S * sizeof (u32));
        if (!hc.base) return fail (SLIPCU_OUT_OF_MEMORY, "slipcu_factor_upload", "device memory exhausted");
        const int per = 256 / CH;
        dim3 grid ((cnt + per - 1) / per, S / CH);
        k_residues<<<grid, 256, 0, F->st>>> (cnt, CH, dl, doff + e, dsg + e, T.p, T.ninv, T.r2, hc.base);
        g_launches++;
        CU (cudaGetLastError ());
        rc = slipcu_factor_set_pivot (F, k, colpiv[k]);
        if (rc) return rc;
        e += (size_t) cnt;
    }
    CU (cudaStreamSynchronize (F->st));
    pool_free (dl); pool_free (doff); pool_free (dsg);
    return check_channels (F);
}

// a channel prime that divides a pivot makes that channel's inverses meaningless
static int check_channels (slipcu_factor *F)
{
    int32_t bad = 0;
    if (F->st2) CU (cudaStreamSynchronize (F->st2));
    CU (cudaMemcpyAsync (&bad, F->bad, sizeof (bad), cudaMemcpyDeviceToHost, F->st));
    CU (cudaStreamSynchronize (F->st));
    if (bad) return fail (SLIPCU_BAD_PRIME, "check_channels", "channel prime divides a pivot");
    return SLIPCU_OK;
}

extern "C" int slipcu_factor_bad_prime (slipcu_factor *F, uint32_t *prime)
{
    int32_t bad = 0;
    if (!F) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_bad_prime", "bad argument");
    USE_DEVICE (F);
    { int rcf = flush_commit (F); if (rcf) return rcf; }
    CU (cudaMemcpyAsync (&bad, F->bad, sizeof (bad), cudaMemcpyDeviceToHost, F->st));
    CU (cudaStreamSynchronize (F->st));
    if (prime) *prime = (bad >= 1 && bad <= F->tab->S) ? F->tab->hp[bad - 1] : 0u;
    return SLIPCU_OK;
}

// The columns come to the host in batches of up to DL_BATCH_BYTES of limbs, two pinned buffer sets
// alternating: while the sink turns batch i into mpz_t (the host's share: allocation and copy of
// every entry), the copy engine brings batch i + 1 (round 1: one copy, one stream synchronisation
// and one sink call per column, strictly one after the other).
#define DL_BATCH_BYTES ((size_t) 64 << 20)
extern "C" int slipcu_factor_download (slipcu_factor *F, slipcu_column_sink sink, void *user)
{
    if (!F || !sink) return fail (SLIPCU_BAD_INPUT, "slipcu_factor_download", "bad argument");
    USE_DEVICE (F);
    { int rcf = flush_commit (F); if (rcf) return rcf; }
    { int rcb = check_channels (F); if (rcb) return rcb; }
    const int n = F->n;
    // batches of consecutive columns
    std::vector<int> first;              // first column of every batch, then n
    size_t maxw = 0, maxc = 0;
    {
        size_t w = 0, c = 0;
        for (int k = 0; k < n; ++k)
        {
            const size_t cw = (size_t) F->cols[k].cnt * F->cols[k].stride;
            if (first.empty () || (w > 0 && (w + cw) * sizeof (u32) > DL_BATCH_BYTES))
            {
                maxw = std::max (maxw, w); maxc = std::max (maxc, c);
                first.push_back (k); w = 0; c = 0;
            }
            w += cw; c += (size_t) F->cols[k].cnt;
        }
        maxw = std::max (maxw, w); maxc = std::max (maxc, c);
        first.push_back (n);
    }
    const int nbatch = (int) first.size () - 1;
    struct Buf { u32 *limbs = nullptr; int32_t *nl = nullptr; int8_t *sign = nullptr; cudaEvent_t done = nullptr; } buf[2];
    int rc = SLIPCU_OK;
    const int nbuf = nbatch > 1 ? 2 : 1;
    for (int i = 0; i < nbuf && rc == SLIPCU_OK; ++i)
    {
        if (host_pool_alloc ((void **) &buf[i].limbs, std::max<size_t> (maxw, 1) * sizeof (u32)) != cudaSuccess
            || host_pool_alloc ((void **) &buf[i].nl, std::max<size_t> (maxc, 1) * sizeof (int32_t)) != cudaSuccess
            || host_pool_alloc ((void **) &buf[i].sign, std::max<size_t> (maxc, 1)) != cudaSuccess
            || pooled_event (&buf[i].done, false) != cudaSuccess)
        { cudaGetLastError (); rc = fail (SLIPCU_OUT_OF_MEMORY, "slipcu_factor_download", "pinned host memory"); }
    }
    auto enqueue = [&] (int bi) -> int
    {
        Buf &B = buf[bi % nbuf];
        size_t w = 0, c = 0;
        for (int k = first[bi]; k < first[bi + 1]; ++k)
        {
            const HostCol &hc = F->cols[k];
            CU (cudaMemcpyAsync (B.limbs + w, hc.limbs, (size_t) hc.cnt * hc.stride * sizeof (u32), cudaMemcpyDeviceToHost, F->st));
            CU (cudaMemcpyAsync (B.nl + c, hc.nl, (size_t) hc.cnt * sizeof (int32_t), cudaMemcpyDeviceToHost, F->st));
            CU (cudaMemcpyAsync (B.sign + c, hc.sign, (size_t) hc.cnt, cudaMemcpyDeviceToHost, F->st));
            g_d2h_bytes += (double) hc.cnt * hc.stride * 4 + (double) hc.cnt * 5;
            w += (size_t) hc.cnt * hc.stride; c += (size_t) hc.cnt;
        }
        CU (cudaEventRecord (B.done, F->st));
        return SLIPCU_OK;
    };
    if (rc == SLIPCU_OK && nbatch > 0) rc = enqueue (0);
    for (int bi = 0; bi < nbatch && rc == SLIPCU_OK; ++bi)
    {
        if (bi + 1 < nbatch) { rc = enqueue (bi + 1); if (rc) break; }
        Buf &B = buf[bi % nbuf];
        if (cudaEventSynchronize (B.done) != cudaSuccess) { rc = fail (SLIPCU_CUDA_ERROR, "slipcu_factor_download", "copy failed"); break; }
        size_t w = 0, c = 0;
        for (int k = first[bi]; k < first[bi + 1]; ++k)
        {
            const HostCol &hc = F->cols[k];
            const int rs = sink (user, k, hc.cnt, hc.stride, B.limbs + w, B.nl + c, B.sign + c);
            if (rs) { rc = fail (rs, "column sink", "host sink failed"); break; }
            w += (size_t) hc.cnt * hc.stride; c += (size_t) hc.cnt;
        }
    }
    cudaStreamSynchronize (F->st);
    for (int i = 0; i < 2; ++i)
    {
        host_pool_free (buf[i].limbs); host_pool_free (buf[i].nl); host_pool_free (buf[i].sign);
        release_event (buf[i].done);
    }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// solve
// ------------------------------------------------------------------------------------------------
// b: nrhs right-hand sides of n entries, entry (row i, right-hand side c) at index c * n + i.
//
// The right-hand sides go through in batches.  A batch is one launch of every kernel: the residues
// of its b, the forward substitution (k_trisolve with a dense source, one CTA per channel block and
// right-hand side, CTAs of one channel block adjacent in the grid so that they share their reads of
// L through L2), the back substitution, ONE reconstruction launch and ONE limb launch for all its
// numerators, one copy to the host.  Two sets of buffers alternate: while the host turns batch i
// into rationals (the sink: limb import, exact verification, canonical form -- the part that costs
// the reference the same GMP time per entry), the GPU works on batch i + 1.
extern "C" int slipcu_solve (slipcu_factor *F, int nrhs, const u32 *blimbs, const int64_t *boff,
                             const int8_t *bsign, const int32_t *pinv, int recon_channels,
                             slipcu_column_sink sink, void *user, int32_t *top_digit_max)
{
    if (!F || nrhs <= 0 || !blimbs || !boff || !bsign || !pinv || !sink)
        return fail (SLIPCU_BAD_INPUT, "slipcu_solve", "bad argument");
    CU (cudaSetDevice (F->device));
    { int rcf = flush_commit (F); if (rcf) return rcf; }
    { int rcb = check_channels (F); if (rcb) return rcb; }
    const Tables &T = *F->tab;
    const int n = F->n, S = F->S, CH = F->CH;
    const int s = std::min (std::max (recon_channels, 1), S);
    if (recon_channels > S) return fail (SLIPCU_BAD_INPUT, "slipcu_solve", "right-hand side needs more channels than the session holds");
    const int stride = (s + 1) & ~1;
    std::vector<int32_t> row_at (n, -1), ident (n);
    for (int r = 0; r < n; ++r)
    {   // pinv must be a permutation of 0..n-1
        if (pinv[r] < 0 || pinv[r] >= n || row_at[pinv[r]] >= 0) return fail (SLIPCU_BAD_INPUT, "slipcu_solve", "pinv is not a permutation");
        row_at[pinv[r]] = r; ident[r] = r;
    }
    // batch size: work vectors, digit scratch and limb buffers of a batch within the memory budget,
    // and at least four batches when there are many right-hand sides (so that host and GPU overlap)
    const size_t budget = (size_t) std::max (1, env_int ("SLIP_B200_SOLVE_BATCH_MB", 1024)) << 20;
    const size_t per_rhs = (size_t) n * (2 * (size_t) S * 4 + ((size_t) S + 4) * 4 + 2 * (size_t) stride * 4);
    int batch = (int) std::max<size_t> (1, std::min<size_t> ((size_t) nrhs, budget / std::max<size_t> (per_rhs, 1)));
    if (nrhs >= 8) batch = std::min (batch, (nrhs + 3) / 4);
    if ((int64_t) batch * n > (int64_t) INT32_MAX / 2) batch = std::max (1, (int) ((int64_t) (INT32_MAX / 2) / n));
    const int nbatches = (nrhs + batch - 1) / batch;
    const size_t bn = (size_t) batch * n;

    u32 *dl = nullptr, *dB = nullptr, *dz = nullptr; int64_t *doff = nullptr; int8_t *dsg = nullptr;
    int32_t *drow_at = nullptr, *dident = nullptr, *dpinv = nullptr, *duoff = nullptr, *dboff = nullptr, *drev = nullptr;
    u32 *dz2 = nullptr;
    WorkCtx bw;                           // step lists of the back substitution (pos[] shared with F->mc)
    int bwd_chunks = 0;
    const bool legacy_backsub = env_int ("SLIP_B200_BACKSUB", 1) == 0;      // k_backsub: one serial loop over the columns per CTA
    struct Buf { u32 *dlimbs = nullptr; int32_t *dnl = nullptr; int8_t *dsign = nullptr;
                 u32 *h_limbs = nullptr; int32_t *h_nl = nullptr; int8_t *h_sign = nullptr; int32_t *h_topd = nullptr;
                 cudaEvent_t done = nullptr; int r0 = 0, nb = 0; bool busy = false; } buf[2];
    int rc = SLIPCU_OK, fwd_chunks = 0;
    double tw_gpu_launch = 0, tw_wait = 0, tw_sink = 0;
    const bool timing = getenv ("SLIP_B200_TIMING") != nullptr;
    if (top_digit_max) *top_digit_max = -1;
    const size_t total = (size_t) n * nrhs;
    const size_t nl = (size_t) boff[total];
    const int nbuf = nbatches > 1 ? 2 : 1;
#define CUG(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail (e_ == cudaErrorMemoryAllocation ? SLIPCU_OUT_OF_MEMORY : SLIPCU_CUDA_ERROR, #call, cudaGetErrorString (e_)); goto done; } } while (0)
    CUG (pool_alloc_t (&dl, std::max<size_t> (nl, 1) * sizeof (u32)));
    CUG (pool_alloc_t (&doff, (total + 1) * sizeof (int64_t)));
    CUG (pool_alloc_t (&dsg, total));
    CUG (pool_alloc_t (&dB, bn * S * sizeof (u32)));
    CUG (pool_alloc_t (&dz, bn * S * sizeof (u32)));
    CUG (pool_alloc_t (&drow_at, (size_t) n * sizeof (int32_t)));
    CUG (pool_alloc_t (&dident, (size_t) n * sizeof (int32_t)));
    CUG (pool_alloc_t (&dpinv, (size_t) n * sizeof (int32_t)));
    for (int i = 0; i < nbuf; ++i)
    {
        CUG (pool_alloc_t (&buf[i].dlimbs, bn * stride * sizeof (u32)));
        CUG (pool_alloc_t (&buf[i].dnl, bn * sizeof (int32_t)));
        CUG (pool_alloc_t (&buf[i].dsign, bn));
        CUG (host_pool_alloc ((void **) &buf[i].h_limbs, bn * stride * sizeof (u32)));
        CUG (host_pool_alloc ((void **) &buf[i].h_nl, bn * sizeof (int32_t)));
        CUG (host_pool_alloc ((void **) &buf[i].h_sign, bn));
        CUG (host_pool_alloc ((void **) &buf[i].h_topd, bn * sizeof (int32_t)));
        CUG (pooled_event (&buf[i].done, false));
    }
    CUG (cudaMemcpyAsync (dl, blimbs, nl * sizeof (u32), cudaMemcpyHostToDevice, F->st));
    CUG (cudaMemcpyAsync (doff, boff, (total + 1) * sizeof (int64_t), cudaMemcpyHostToDevice, F->st));
    CUG (cudaMemcpyAsync (dsg, bsign, total, cudaMemcpyHostToDevice, F->st));
    CUG (cudaMemcpyAsync (drow_at, row_at.data (), (size_t) n * sizeof (int32_t), cudaMemcpyHostToDevice, F->st));
    CUG (cudaMemcpyAsync (dident, ident.data (), (size_t) n * sizeof (int32_t), cudaMemcpyHostToDevice, F->st));
    CUG (cudaMemcpyAsync (dpinv, pinv, (size_t) n * sizeof (int32_t), cudaMemcpyHostToDevice, F->st));
    g_h2d_bytes += (double) nl * 4 + (double) (total + 1) * 8 + (double) total + 3.0 * n * 4;
    rc = ensure_digits (F, bn);
    if (rc) goto done;
    {   // symbolic pre-pass of the forward substitution: every column of L is one step
        std::vector<int32_t> uoff (2 * (size_t) n + 2);
        int64_t tot = 0, nch = 0;
        for (int k = 0; k < n; ++k)
        {
            const int len = F->cols[k].cnt - F->cols[k].nU;
            uoff[k] = (int32_t) tot; uoff[n + 1 + k] = (int32_t) nch;
            tot += tri_slot_extent (len, CH); nch += (len + tri_chunk_rows (CH) - 1) / tri_chunk_rows (CH);
        }
        uoff[n] = (int32_t) tot; uoff[2 * n + 1] = (int32_t) nch;
        fwd_chunks = (int) nch;
        if (tot > INT32_MAX) { rc = fail (SLIPCU_BAD_INPUT, "slipcu_solve", "L has too many entries"); goto done; }
        CUG (pool_alloc_t (&duoff, (2 * (size_t) n + 2) * sizeof (int32_t)));
        CUG (cudaMemcpyAsync (duoff, uoff.data (), (2 * (size_t) n + 2) * sizeof (int32_t), cudaMemcpyHostToDevice, F->st));
        CUG (cudaStreamSynchronize (F->st));
        rc = prepare_steps (F, F->mc, n, n, F->rows_are_positions ? dident : drow_at, dident, duoff, duoff + n + 1, (int) tot, (int) nch);
        if (rc) goto done;
    }
    if (!legacy_backsub)
    {   // ... and of the back substitution: the U part of every column is one step, columns n-1..0
        std::vector<int32_t> boff (2 * (size_t) n + 2), rev (n);
        int64_t tot = 0, nch = 0;
        for (int u = 0; u < n; ++u)
        {
            const int k = n - 1 - u, len = F->cols[k].nU;
            rev[u] = k;
            boff[u] = (int32_t) tot; boff[n + 1 + u] = (int32_t) nch;
            tot += tri_slot_extent (len, CH); nch += (len + tri_chunk_rows (CH) - 1) / tri_chunk_rows (CH);
        }
        boff[n] = (int32_t) tot; boff[2 * n + 1] = (int32_t) nch;
        bwd_chunks = (int) nch;
        if (tot > INT32_MAX) { rc = fail (SLIPCU_BAD_INPUT, "slipcu_solve", "U has too many entries"); goto done; }
        CUG (pool_alloc_t (&dboff, (2 * (size_t) n + 2) * sizeof (int32_t)));
        CUG (pool_alloc_t (&drev, (size_t) n * sizeof (int32_t)));
        CUG (pool_alloc_t (&dz2, bn * S * sizeof (u32)));
        CUG (cudaMemcpyAsync (dboff, boff.data (), (2 * (size_t) n + 2) * sizeof (int32_t), cudaMemcpyHostToDevice, F->st));
        CUG (cudaMemcpyAsync (drev, rev.data (), (size_t) n * sizeof (int32_t), cudaMemcpyHostToDevice, F->st));
        CUG (cudaStreamSynchronize (F->st));
        bw.st = F->st; bw.pos = F->mc.pos;
        rc = prepare_steps (F, bw, n, n, F->rows_are_positions ? dident : drow_at, drev, dboff, dboff + n + 1, (int) tot, (int) nch, 0, false, 1);
        if (rc) goto done;
    }
    for (int bi = 0; bi <= nbatches; ++bi)
    {
        double tw0 = wall_s ();
        if (bi < nbatches)
        {   // enqueue batch bi
            Buf &B = buf[bi % nbuf];
            const int r0 = bi * batch, nb = std::min (batch, nrhs - r0);
            const size_t cntb = (size_t) nb * n;            // entries of the batch; region [S/CH][cntb][CH]
            B.r0 = r0; B.nb = nb;
            {
                const int per = 256 / CH;
                dim3 grid ((unsigned) ((cntb + per - 1) / per), S / CH);
                k_residues<<<grid, 256, 0, F->st>>> ((int) cntb, CH, dl, doff + (size_t) r0 * n, dsg + (size_t) r0 * n, T.p, T.ninv, T.r2, dB);
                g_launches++;
                CUG (cudaGetLastError ());
            }
            TriArgs a; memset (&a, 0, sizeof (a));
            a.k = n; a.S = S; a.cnt = n; a.nU = n;
            // slots are positions.  Resident sessions store original rows (slot of row r = pinv[r]);
            // uploaded sessions store positions (identity map), b rows are then routed through pinv.
            a.rows = F->rows_are_positions ? dident : drow_at;
            a.steps = F->mc.steps; a.chunks = F->mc.chunks; a.slots = F->mc.slots;
            a.src = dB; a.src_total = cntb; a.src_first = 0; a.src_step = 1; a.src_cnt = n;
            a.src_rows = F->rows_are_positions ? dpinv : nullptr;
            a.src_y_stride = (size_t) n;
            a.out = dz; a.out_y_stride = (size_t) n * CH; a.out_cb_stride = cntb * CH;
            a.rho = F->rho; a.invrho = F->invrho;
            a.p = T.p; a.ninv = T.ninv; a.pos = F->mc.pos;
            a.nchunks = fwd_chunks; a.upos = dident; a.publish = 1; a.u0 = 0;
            a.rhs_fastest = 1;
            size_t smem = 0;
            rc = tri_geometry (F, a, &smem);
            if (rc) goto done;
            CUG (launch_tri_any (CH, F->cpt, a, dim3 (nb, S / CH), smem, F->st));
            if (debug_check ("k_trisolve(forward)", F->st)) { rc = fail (SLIPCU_CUDA_ERROR, "k_trisolve(forward)", "debug"); goto done; }
            u32 *dres = dz;                 // where the numerators end up
            if (!legacy_backsub)
            {   // back substitution through the same chunked pipeline: z = det y on the way in, the U
                // parts as steps n-1..0, x_t = z_t / rho_t on the way out
                TriArgs bq = a;
                bq.chunks = bw.chunks; bq.slots = bw.slots; bq.steps = bw.steps; bq.nchunks = bwd_chunks;
                bq.src = dz; bq.src_total = cntb; bq.src_first = 0; bq.src_step = 1; bq.src_cnt = n;
                bq.src_rows = nullptr;                        // dz is in slot (position) order already
                bq.pos = dident;                              // ... so entry e goes to slot e
                bq.src_y_stride = (size_t) n;
                bq.out = dz2; bq.src_scale = F->rho + (size_t) (n - 1) * S; bq.j_is_slot = 1; bq.publish = 2;
                CUG (launch_tri_any (CH, F->cpt, bq, dim3 (nb, S / CH), smem, F->st));
                if (debug_check ("k_trisolve(backward)", F->st)) { rc = fail (SLIPCU_CUDA_ERROR, "k_trisolve(backward)", "debug"); goto done; }
                dres = dz2;
            }
            else
            {
            BackArgs b;
            b.n = n; b.S = S; b.z = dz; b.z_y_stride = (size_t) n * CH; b.z_cb_stride = cntb * CH; b.rhs_fastest = 1;
            b.desc = F->desc; b.rho = F->rho; b.invrho = F->invrho; b.p = T.p; b.ninv = T.ninv; b.pos = F->mc.pos;
            g_launches++;
            if (CH == 4) k_backsub<4><<<dim3 (nb, S / CH), 256, 0, F->st>>> (b);
            else if (CH == 8) k_backsub<8><<<dim3 (nb, S / CH), 256, 0, F->st>>> (b);
            else if (CH == 16) k_backsub<16><<<dim3 (nb, S / CH), 256, 0, F->st>>> (b);
            else k_backsub<32><<<dim3 (nb, S / CH), 256, 0, F->st>>> (b);
            CUG (cudaGetLastError ());
            if (debug_check ("k_backsub", F->st)) { rc = fail (SLIPCU_CUDA_ERROR, "k_backsub", "debug"); goto done; }
            }
            // numerators of the whole batch: one reconstruction launch, one limb launch
            rc = run_garner (F, dres, (int) cntb, 0, (int) cntb, s, B.dsign);
            if (rc) goto done;
            rc = run_limbs (F, 0, (int) cntb, 0, stride, s, B.dlimbs, B.dnl);
            if (rc) goto done;
            CUG (cudaEventRecord (F->ev_end, F->st));
            CUG (cudaMemcpyAsync (B.h_limbs, B.dlimbs, cntb * stride * sizeof (u32), cudaMemcpyDeviceToHost, F->st));
            CUG (cudaMemcpyAsync (B.h_nl, B.dnl, cntb * sizeof (int32_t), cudaMemcpyDeviceToHost, F->st));
            CUG (cudaMemcpyAsync (B.h_sign, B.dsign, cntb, cudaMemcpyDeviceToHost, F->st));
            if (top_digit_max) CUG (cudaMemcpyAsync (B.h_topd, F->topd, cntb * sizeof (int32_t), cudaMemcpyDeviceToHost, F->st));
            CUG (cudaEventRecord (B.done, F->st));
            B.busy = true;
            g_d2h_bytes += (double) cntb * stride * 4 + (double) cntb * 5;
        }
        tw_gpu_launch += wall_s () - tw0; tw0 = wall_s ();
        if (bi > 0)
        {   // hand batch bi - 1 to the sink while the GPU works on batch bi
            Buf &B = buf[(bi - 1) % nbuf];
            CUG (cudaEventSynchronize (B.done));
            tw_wait += wall_s () - tw0; tw0 = wall_s ();
            B.busy = false;
            for (int r = 0; r < B.nb; ++r)
            {
                const size_t e0 = (size_t) r * n;
                if (top_digit_max) for (int t = 0; t < n; ++t) *top_digit_max = std::max (*top_digit_max, B.h_topd[e0 + t]);
                const int rs = sink (user, B.r0 + r, n, stride, B.h_limbs + e0 * stride, B.h_nl + e0, B.h_sign + e0);
                if (rs) { rc = fail (rs, "column sink", "host sink failed"); goto done; }
            }
            tw_sink += wall_s () - tw0;
        }
    }
    if (timing) fprintf (stderr, "slipcu_solve wall: enqueue %.3f wait-for-gpu %.3f host sink %.3f (nrhs %d in %d batches of %d, n %d, channels %d, recon channels %d)\n",
                         tw_gpu_launch, tw_wait, tw_sink, nrhs, nbatches, batch, n, S, s);
done:
    cudaStreamSynchronize (F->st);
    flush_timers (F);
    if (rc == SLIPCU_OK && !F->rows_are_positions)
    {   // device-side job time: A resident -> solution numerators reconstructed on the device
        float ms = 0;
        if (cudaEventElapsedTime (&ms, F->ev_start, F->ev_end) == cudaSuccess) { g_device_ms += ms; F->timed = true; } else cudaGetLastError ();
    }
    pool_free (dl); pool_free (doff); pool_free (dsg); pool_free (dB); pool_free (dz);
    pool_free (drow_at); pool_free (dident); pool_free (dpinv); pool_free (duoff);
    pool_free (dboff); pool_free (drev); pool_free (dz2);
    pool_free (bw.slots); pool_free (bw.steps); pool_free (bw.chunks);
    for (int i = 0; i < 2; ++i)
    {
        pool_free (buf[i].dlimbs); pool_free (buf[i].dnl); pool_free (buf[i].dsign);
        host_pool_free (buf[i].h_limbs); host_pool_free (buf[i].h_nl); host_pool_free (buf[i].h_sign); host_pool_free (buf[i].h_topd);
        if (buf[i].done) release_event (buf[i].done);
    }
#undef CUG
    return rc;
}

// ------------------------------------------------------------------------------------------------
// counters
// ------------------------------------------------------------------------------------------------
extern "C" void slipcu_get_counters (slipcu_counters *o)
{
    if (!o) return;
    o->launches = g_launches.load (); o->trisolve_launches = g_tri_launches.load ();
    o->trisolve_ms = g_tri_ms; o->trisolve_bytes = g_tri_bytes; o->trisolve_modmul = g_tri_modmul;
    o->recon_ms = g_recon_ms; o->recon_mac = g_recon_mac;
    o->h2d_bytes = g_h2d_bytes; o->d2h_bytes = g_d2h_bytes; o->device_ms = g_device_ms; o->other_ms = g_other_ms;
    {   // union of the k_trisolve intervals
        std::lock_guard<std::mutex> lk (g_event_mutex);
        std::vector<std::pair<float, float>> iv = g_tri_intervals;
        std::sort (iv.begin (), iv.end ());
        double total = 0; float lo = 0, hi = -1;
        for (auto &p : iv)
        {
            if (hi < lo || p.first > hi) { if (hi >= lo) total += hi - lo; lo = p.first; hi = p.second; }
            else hi = std::max (hi, p.second);
        }
        if (hi >= lo) total += hi - lo;
        o->trisolve_union_ms = total;
    }
}
extern "C" void slipcu_reset_counters (void)
{
    g_launches = 0; g_tri_launches = 0;
    g_tri_ms = g_tri_bytes = g_tri_modmul = g_recon_ms = g_recon_mac = 0;
    g_h2d_bytes = g_d2h_bytes = g_device_ms = g_other_ms = 0;
    std::lock_guard<std::mutex> lk (g_event_mutex);
    g_tri_intervals.clear ();
    if (g_base_ev) { g_event_pool.push_back (g_base_ev); g_base_ev = nullptr; }
}
extern "C" void slipcu_set_profiling (int enabled) { g_profiling = enabled; }

extern "C" void slipcu_release_cached_memory (void)
{
    { std::lock_guard<std::mutex> lk (g_pool_mutex); pool_trim_locked (); }
    std::lock_guard<std::mutex> lk (g_hpool_mutex);
    for (auto &kv : g_hpool_free) { cudaFreeHost (kv.second); g_hpool_size.erase (kv.second); }
    g_hpool_free.clear ();
}
