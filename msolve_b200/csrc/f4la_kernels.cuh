/* f4la_kernels.cuh -- sm_100a kernels of the F4 GF(p) linear algebra.
 *
 * Device algorithm (see DESIGN.md for the derivation and the roofline):
 *
 *   The reference reduces every lower row r by the known pivots with a
 *   left-to-right chain of sparse AXPYs into a dense int64 accumulator
 *   (reference src/neogb/la_ff_32.c:965-1242, driven by :2637-2698).  The
 *   multiplier of pivot i in row r, m[r][i], and the remainder of row r in the
 *   non-pivot columns, D'[r][:], satisfy
 *
 *        V[j][r] = X[r][j] - sum_{i<j, a_ij != 0} a_ij * V[i][r]        (*)
 *
 *   where X = [C D] are the lower rows, a_ij the entries of the (unit upper
 *   triangular) pivot rows, V[j] = m[:,j] for j < ncl and V[j] = D'[:,j-ncl]
 *   for j >= ncl.  (*) is a sparse triangular solve with nrl right-hand
 *   sides.  We run it TRANSPOSED: V[j] is a dense vector over the lower rows,
 *   a column j pulls the vectors of its source pivots (CSC of the pivot
 *   rows), so every memory access is a coalesced dense vector load, there is
 *   no scatter and no atomics in the hot loop.  Columns are level-scheduled
 *   (level[j] = 1 + max level of its sources).  The lower rows are split into
 *   chunks of 128*T rows so that the V block of the chunks in flight stays
 *   L2-resident (126 MB on B200).
 *
 *   The Schur complement D' (ncr x nrl, column major) is then brought to
 *   reduced row echelon form by Gauss-Jordan elimination (trailing block,
 *   reference la_ff_32.c:3390-3500 + :3354-3388 compute the same unique RREF).
 *
 * Arithmetic is exact: residues are kept canonical in [0,p) as uint32.  For p >= 2^16 a dot
 * product is accumulated in 96 bits (one IMAD.WIDE.U32 with carry-out per product plus an
 * add-with-carry into a third word, see struct Acc below), for p < 2^16 in a plain uint64; one
 * Barrett reduction (floor(2^64/p)) per row and column at the end.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace f4la {

constexpr int kSolveThreads   = 256;   /* 8 warps */
constexpr int kWarpsPerBlock  = kSolveThreads / 32;
constexpr int kRowsPerT       = 128;   /* rows covered by one uint4 per lane */
constexpr int kMaxT           = 4;     /* chunk = 128*T rows, T in 1..4      */
constexpr uint32_t kScatterSplit = 8;  /* warps per lower row in k_scatter_lower */
#ifndef F4LA_LONG_COL_LOG2
#define F4LA_LONG_COL_LOG2 7           /* columns with >= 2^7 = 128 pull-list entries are solved by a whole CTA */
#endif
constexpr uint32_t kLongColumn = 1u << F4LA_LONG_COL_LOG2;  /* >= this many entries: one CTA/col  */
constexpr int kLenClasses     = F4LA_LONG_COL_LOG2;         /* class 0 = long, then the short columns by halving length:
                                                               [L/2, L), [L/4, L/2), ..., [4, 8), [0, 4) */

__host__ __device__ __forceinline__ uint32_t len_class(uint32_t len)
{
    if (len >= kLongColumn) return 0;
    uint32_t c = 1, lim = kLongColumn >> 1;
    while (len < lim && lim > 4) { ++c; lim >>= 1; }
    return len < 4 ? (uint32_t)kLenClasses - 1 : c;
}

/* x mod p for any 64-bit x, pinv = floor((2^64-1)/p).  The quotient estimate
 * is short by at most 2, the loop repairs it. */
__device__ __forceinline__ uint32_t mod_u64(uint64_t x, uint32_t p, uint64_t pinv)
{
    const uint64_t q = __umul64hi(x, pinv);
    uint64_t r = x - q * (uint64_t)p;
    while (r >= p) r -= p;
    return (uint32_t)r;
}

__device__ __forceinline__ uint32_t mulmod(uint32_t a, uint32_t b, uint32_t p, uint64_t pinv)
{
    return mod_u64((uint64_t)a * b, p, pinv);
}

/* Shoup multiplication by a fixed scalar sc < p < 2^31: w = floor(sc * 2^32 / p) once, then
 * f * sc mod p costs one mul.hi, two mul.lo and a conditional subtraction (the estimate of the
 * quotient is short by at most one, so the remainder lies in [0, 2p) and fits 32 bits). */
__device__ __forceinline__ uint32_t shoup_pre(uint32_t sc, uint32_t p)
{
    return (uint32_t)(((uint64_t)sc << 32) / p);
}

/* the same without a 64-bit division: pinv = floor((2^64 - 1) / p) */
__device__ __forceinline__ uint32_t shoup_pre_fast(uint32_t sc, uint32_t p, uint64_t pinv)
{
    const uint64_t x = (uint64_t)sc << 32;
    uint64_t q = __umul64hi(x, pinv);
    uint64_t r = x - q * p;
    while (r >= p) { r -= p; ++q; }
    return (uint32_t)q;
}

__device__ __forceinline__ uint32_t shoup_mul(uint32_t f, uint32_t sc, uint32_t w, uint32_t p)
{
    const uint32_t q = __umulhi(f, w);
    const uint32_t r = f * sc - q * p;
    return r >= p ? r - p : r;
}

/* Modular inverse by extended Euclid (same recurrence as the reference's
 * mod_p_inverse_32, tools.h:95-122). */
__device__ __forceinline__ uint32_t modinv(uint32_t val, uint32_t p)
{
    /* p < 2^31: remainders fit 32 bits (cheap divisions), Bezout coefficients fit int64 */
    uint32_t a = p, b = val % p;
    int64_t c = 1, d = 0;
    while (b != 0) {
        const uint32_t e = a / b;
        const uint32_t r = a - e * b;
        a = b; b = r;
        const int64_t f = d - (int64_t)e * c;
        d = c; c = f;
    }
    if (d < 0) d += p;
    return (uint32_t)d;
}

__device__ __forceinline__ uint32_t load_cf(const void *cf, uint32_t cf_bytes, uint64_t idx)
{
    if (cf_bytes == 4) return ((const uint32_t *)cf)[idx];
    if (cf_bytes == 2) return ((const uint16_t *)cf)[idx];
    return ((const uint8_t *)cf)[idx];
}

/* ------------------------------------------------------------------ */
/* layout kernels                                                      */
/* ------------------------------------------------------------------ */

/* error bits written by the layout kernels */
constexpr uint32_t kErrLeadNotFirst  = 1u;  /* pivot row i does not start with column i */
constexpr uint32_t kErrNotTriangular = 2u;  /* pivot entry left of / on its lead        */
constexpr uint32_t kErrColumnRange   = 4u;  /* column id >= nc                          */
constexpr uint32_t kErrLeadNotOne    = 8u;  /* pivot row not monic                      */
constexpr uint32_t kErrDuplicatePivot = 32u; /* two pivot rows with the same lead column */

/* Pass 1 over the pivot rows: histogram of the target columns of all
 * non-lead entries.  One warp per pivot row. */
__global__ void k_pull_count(const uint64_t *__restrict__ row_ptr,
                             const uint32_t *__restrict__ cols,
                             const uint32_t *__restrict__ row_cf,
                             const uint64_t *__restrict__ cf_ptr,
                             const void *__restrict__ cf, uint32_t cf_bytes,
                             uint32_t nru, uint32_t ncl, uint32_t nc,
                             const uint32_t *__restrict__ colmap,
                             uint32_t *__restrict__ col_cnt, uint32_t *__restrict__ err)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < nru; i += nwarps) {
        const uint64_t a = row_ptr[i], b = row_ptr[i + 1];
        /* pivot id = (renumbered) lead column; without a column map it must equal the row index */
        uint32_t id = i;
        if (b == a) { if (lane == 0) atomicOr(err, kErrLeadNotFirst); continue; }
        if (colmap) {
            const uint32_t lead = cols[a];
            if (lead >= nc) { if (lane == 0) atomicOr(err, kErrColumnRange); continue; }
            id = colmap[lead];
        }
        if (lane == 0) {
            if (!colmap && cols[a] != i) atomicOr(err, kErrLeadNotFirst);
            else if (load_cf(cf, cf_bytes, cf_ptr[row_cf[i]]) != 1) atomicOr(err, kErrLeadNotOne);
        }
        for (uint64_t k = a + 1 + lane; k < b; k += 32) {
            uint32_t j = cols[k];
            if (j >= nc) { atomicOr(err, kErrColumnRange); continue; }
            if (colmap) j = colmap[j];
            if (j <= id) { atomicOr(err, kErrNotTriangular); continue; }
            atomicAdd(&col_cnt[j], 1u);
        }
    }
}

/* Exclusive prefix sum of n counters into out[0..n] (out[n] = total), one
 * block.  n is at most a few 10^5, this is not a hot kernel. */
__global__ void k_exclusive_scan(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n)
{
    __shared__ uint32_t part[1024];
    const uint32_t t = threadIdx.x, nt = blockDim.x;
    const uint32_t per = (n + nt - 1) / nt;
    const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
    uint32_t s = 0;
    for (uint32_t k = lo; k < hi; ++k) s += in[k];
    part[t] = s;
    __syncthreads();
    for (uint32_t off = 1; off < nt; off <<= 1) {
        const uint32_t v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    for (uint32_t k = lo; k < hi; ++k) { const uint32_t c = in[k]; out[k] = run; run += c; }
    if (t == nt - 1) out[n] = part[t];
}

/* Pass 2: fill the pull lists.  Entry = { source pivot i, p - a_ij }. */
__global__ void k_pull_fill(const uint64_t *__restrict__ row_ptr,
                            const uint32_t *__restrict__ cols,
                            const uint32_t *__restrict__ row_cf,
                            const uint64_t *__restrict__ cf_ptr,
                            const void *__restrict__ cf, uint32_t cf_bytes,
                            uint32_t nru, uint32_t nc, uint32_t p,
                            const uint32_t *__restrict__ colmap,
                            const uint32_t *__restrict__ col_ptr,
                            uint32_t *__restrict__ cursor, uint2 *__restrict__ ent,
                            uint32_t *__restrict__ ent_src)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < nru; i += nwarps) {
        const uint64_t a = row_ptr[i], b = row_ptr[i + 1];
        if (b == a) continue;
        uint32_t id = i;
        if (colmap) {
            const uint32_t lead = cols[a];
            if (lead >= nc) continue;
            id = colmap[lead];
        }
        const uint64_t c0 = cf_ptr[row_cf[i]];
        for (uint64_t k = a + 1 + lane; k < b; k += 32) {
            uint32_t j = cols[k];
            if (j >= nc) continue;
            if (colmap) j = colmap[j];
            if (j <= id) continue;
            const uint32_t v = load_cf(cf, cf_bytes, c0 + (k - a)) % p;
            const uint32_t pos = col_ptr[j] + atomicAdd(&cursor[j], 1u);
            ent[pos] = make_uint2(id, v == 0 ? 0u : p - v);
            if (ent_src) ent_src[pos] = (uint32_t)(c0 + (k - a));      /* where the coefficient lives in the pool */
        }
    }
}

/* Same structure, other coefficients (f4la_b200_resident_update: another prime of the multi-modular loop): only the
 * values of the pull lists are refreshed, from the recorded pool positions. */
__global__ void k_ent_refill(uint2 *__restrict__ ent, const uint32_t *__restrict__ ent_src, const void *__restrict__ cf,
                             uint32_t cf_bytes, uint32_t p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = load_cf(cf, cf_bytes, ent_src[i]) % p;
    ent[i].y = v == 0 ? 0u : p - v;
}

/* Column renumbering for matrices whose known pivots are keyed by their lead
 * column (F4LA_FLAG_PIVOT_BY_LEAD): pivot columns first, both groups in their
 * original relative order, so the pivot rows stay upper triangular. */
__global__ void k_mark_leads(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ cols,
                             uint32_t nru, uint32_t nc, uint32_t *__restrict__ is_piv, uint32_t *__restrict__ err)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nru) return;
    const uint64_t a = row_ptr[i];
    if (row_ptr[i + 1] == a) { atomicOr(err, kErrLeadNotFirst); return; }
    const uint32_t lead = cols[a];
    if (lead >= nc) { atomicOr(err, kErrColumnRange); return; }
    if (atomicExch(&is_piv[lead], 1u)) atomicOr(err, kErrDuplicatePivot);
}

/* pividx = exclusive scan of is_piv; colmap[c] = new id, invmap[new id] = c */
__global__ void k_build_colmap(const uint32_t *__restrict__ is_piv, const uint32_t *__restrict__ pividx,
                               uint32_t nc, uint32_t npiv, uint32_t *__restrict__ colmap,
                               uint32_t *__restrict__ invmap)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const uint32_t k = is_piv[c] ? pividx[c] : npiv + (c - pividx[c]);
    colmap[c] = k;
    invmap[k] = c;
}

/* Exact longest-path levels in ONE pass: a persistent grid hands out the
 * columns in increasing order (a topological order, sources are always to the
 * left), a warp waits until the levels of its sources have been published.
 * level[] must be pre-set to kLevelUnset.  Same no-deadlock argument as the
 * dataflow solve below: the oldest unfinished column never waits. */
constexpr uint32_t kLevelUnset = 0xffffffffu;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *ptr)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}

__global__ void k_level_flow(const uint32_t *__restrict__ col_ptr, const uint2 *__restrict__ ent,
                             uint32_t ncl, uint32_t *level, uint32_t *__restrict__ counter,
                             uint32_t *__restrict__ err)
{
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(counter, 1u);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= ncl) return;
        const uint32_t a = col_ptr[j], b = col_ptr[j + 1];
        uint32_t m = 0;
        for (uint32_t k = a + lane; k < b; k += 32) {
            const uint32_t src = __ldg(&ent[k]).x;
            uint32_t l = ld_acquire_u32(&level[src]);
            uint32_t spins = 0;
            while (l == kLevelUnset) {
                __nanosleep(32);
                l = ld_acquire_u32(&level[src]);
                if (++spins > (1u << 24)) { atomicOr(err, 16u); l = 0; }
            }
            m = max(m, l + 1);
        }
        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(level + j), "r"(m) : "memory");
    }
}

/* Histogram of (level, length class): slot = kLenClasses * level + len_class(len).
 * Short columns of similar length share an item, so the 8 warps of a CTA finish together. */
__global__ void k_level_count(const uint32_t *__restrict__ level, const uint32_t *__restrict__ col_ptr,
                              uint32_t ncl, uint32_t *__restrict__ slot_cnt, uint32_t *__restrict__ max_level)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncl) return;
    const uint32_t len = col_ptr[j + 1] - col_ptr[j];
    atomicAdd(&slot_cnt[kLenClasses * level[j] + len_class(len)], 1u);
    atomicMax(max_level, level[j]);
}

__global__ void k_level_fill(const uint32_t *__restrict__ level, const uint32_t *__restrict__ col_ptr,
                             uint32_t ncl, const uint32_t *__restrict__ slot_ptr,
                             uint32_t *__restrict__ cursor, uint32_t *__restrict__ order)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncl) return;
    const uint32_t len = col_ptr[j + 1] - col_ptr[j];
    const uint32_t s = kLenClasses * level[j] + len_class(len);
    order[slot_ptr[s] + atomicAdd(&cursor[s], 1u)] = j;
}

/* Scatter the lower rows of the chunks [chunk0, chunk0+nchunks) into V.
 * V layout: [chunk][column][Rc].  One warp per lower row. */
__global__ void k_scatter_lower(const uint64_t *__restrict__ row_ptr,
                                const uint32_t *__restrict__ cols,
                                const uint32_t *__restrict__ row_cf,
                                const uint64_t *__restrict__ cf_ptr,
                                const void *__restrict__ cf, uint32_t cf_bytes,
                                uint32_t nru, uint32_t nrl, uint32_t nc, uint32_t p,
                                uint32_t chunk0, uint32_t nchunks, uint32_t Rc,
                                const uint32_t *__restrict__ colmap, uint32_t *__restrict__ flags,
                                uint32_t *__restrict__ V, uint32_t *__restrict__ err)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t r0 = chunk0 * Rc;
    const uint32_t r1 = min(nrl, (chunk0 + nchunks) * Rc);
    /* kScatterSplit warps share a row: a chunk has only a few hundred rows */
    const uint32_t nwork = (r1 - r0) * kScatterSplit;
    for (uint32_t wk = warp; wk < nwork; wk += nwarps) {
        const uint32_t r = r0 + wk / kScatterSplit, part = wk % kScatterSplit;
        const uint64_t a = row_ptr[nru + r], b = row_ptr[nru + r + 1];
        const uint64_t c0 = cf_ptr[row_cf[nru + r]];
        const uint32_t lc = (r - r0) / Rc, lr = (r - r0) % Rc;
        uint32_t *Vc = V + (size_t)lc * nc * Rc;
        for (uint64_t k = a + part * 32 + lane; k < b; k += 32 * kScatterSplit) {
            uint32_t j = cols[k];
            if (j >= nc) { atomicOr(err, kErrColumnRange); continue; }
            if (colmap) j = colmap[j];
            const uint32_t v = load_cf(cf, cf_bytes, c0 + (k - a)) % p;
            Vc[(size_t)j * Rc + lr] = v;
            if (flags && v && flags[(size_t)lc * nc + j] == 0xfffffffeu)      /* kFlagAlwaysZ -> kFlagAlwaysNZ */
                flags[(size_t)lc * nc + j] = 0xffffffffu;
        }
    }
}

/* counter-based RNG (splitmix64 finaliser): coefficient / class of (seed, family, row) */
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

/* Probabilistic mode: V <- R * X for a sparse random R (nvr x nrl).  Lower row r enters one
 * combination of each of the `weight` families (family f owns the combinations
 * [f*kw, (f+1)*kw), kw = nvr / weight; the class inside the family and the non-zero
 * coefficient come from the hash), so every row is represented `weight` times.  One warp per
 * (row, family); several rows add into the same cell, hence the compare-and-swap modular add. */
__global__ void k_scatter_combos(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ cols,
                                 const uint32_t *__restrict__ row_cf, const uint64_t *__restrict__ cf_ptr,
                                 const void *__restrict__ cf, uint32_t cf_bytes, uint32_t nru, uint32_t nrl,
                                 uint32_t nc, uint32_t p, uint64_t pinv, uint32_t chunk0, uint32_t nchunks,
                                 uint32_t Rc, uint32_t nvr, uint32_t weight, uint64_t seed,
                                 uint32_t *__restrict__ flags, uint32_t *__restrict__ V, uint32_t *__restrict__ err)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t kw = nvr / weight;
    const uint32_t v0 = chunk0 * Rc, v1 = min(nvr, (chunk0 + nchunks) * Rc);
    for (uint32_t wk = warp; wk < nrl * weight; wk += nwarps) {
        const uint32_t r = wk / weight, fam = wk % weight;
        const uint64_t h = mix64(seed ^ ((uint64_t)fam << 40) ^ r);
        const uint32_t vr = fam * kw + (uint32_t)(h % kw);                 /* the combination this row joins */
        if (vr < v0 || vr >= v1) continue;
        const uint32_t rho = 1u + (uint32_t)((h >> 20) % (p - 1));          /* non-zero multiplier */
        const uint64_t a = row_ptr[nru + r], b = row_ptr[nru + r + 1];
        const uint64_t c0 = cf_ptr[row_cf[nru + r]];
        const uint32_t lc = (vr - v0) / Rc, lr = (vr - v0) % Rc;
        uint32_t *Vc = V + (size_t)lc * nc * Rc;
        for (uint64_t k = a + lane; k < b; k += 32) {
            const uint32_t j = cols[k];
            if (j >= nc) { atomicOr(err, kErrColumnRange); continue; }
            const uint32_t x = load_cf(cf, cf_bytes, c0 + (k - a)) % p;
            const uint32_t t = mulmod(x, rho, p, pinv);
            if (!t) continue;
            uint32_t *cell = &Vc[(size_t)j * Rc + lr];
            uint32_t old = *cell, assumed;
            do {
                assumed = old;
                uint32_t nv = assumed + t;                 /* < 2^32: both < p < 2^31 */
                if (nv >= p) nv -= p;
                old = atomicCAS(cell, assumed, nv);
            } while (old != assumed);
            /* "non-zero" is only a hint that keeps the loads: a sum that cancels to zero is still correct */
            if (flags[(size_t)lc * nc + j] == 0xfffffffeu) flags[(size_t)lc * nc + j] = 0xffffffffu;
        }
    }
}

/* ------------------------------------------------------------------ */
/* arithmetic helpers of the hot kernel                                 */
/* ------------------------------------------------------------------ */

/* Accumulator of one (row, column) dot product.
 * WIDE (p >= 2^16): 96 bits.  A product v*c < 2^62 is added with ONE IMAD.WIDE.U32 (carry-out to a
 * predicate) and one add-with-carry into the third word (IADD3.X, ALU pipe): the multiplier pipe
 * issues IMAD.WIDE at 32 lanes/clk/SM, so the former scheme (coefficient split into 16-bit halves,
 * two IMAD.WIDE into two u64 sums) topped out at 16 MAC/clk/SM; this one measures 28.8
 * (tools/ubench/imad_wide.cu).  a2 counts carries: good for 2^32 products, more than a column
 * can hold (nnz < 2^32), so no intermediate folds are needed.
 * !WIDE (p < 2^16): products < 2^32, a plain 64-bit sum (one IMAD.WIDE) holds 2^32 of them. */
struct Acc { uint64_t lo; uint32_t a2; };     /* lo stays ONE 64-bit register: an aligned pair, accumulated in place */

template <bool WIDE>
__device__ __forceinline__ void mac(Acc &a, uint32_t v, uint32_t c)
{
    if (WIDE)
        asm("{\n\t.reg .u32 l, h;\n\tmov.b64 {l, h}, %0;\n\tmad.lo.cc.u32 l, %2, %3, l;\n\tmadc.hi.cc.u32 h, %2, %3, h;\n\t"
            "addc.u32 %1, %1, 0;\n\tmov.b64 %0, {l, h};\n\t}"
            : "+l"(a.lo), "+r"(a.a2) : "r"(v), "r"(c));
    else
        a.lo += (uint64_t)v * c;
}

/* canonical residue of the accumulator plus extra (< 2^32); two64 = 2^64 mod p */
template <bool WIDE>
__device__ __forceinline__ uint32_t fold_acc(const Acc &a, uint32_t extra, uint32_t p, uint64_t pinv, uint32_t two64)
{
    uint64_t x = (uint64_t)mod_u64(a.lo, p, pinv) + extra;
    if (WIDE) x += mod_u64((uint64_t)a.a2 * two64, p, pinv);
    return mod_u64(x, p, pinv);
}

/* Peak of the multiply-add the hot kernel is made of (mac<true>: one IMAD.WIDE.U32 with carry-out + one
 * add-with-carry), operands in registers: the ceiling bench.py's roofline.binding quotes, measured on the GPU the
 * bench runs on instead of a constant. */
__global__ void __launch_bounds__(256) k_mac_peak(uint32_t *out, uint32_t seed, int iters)
{
    uint32_t v[8], c = seed | 1u;
    Acc a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = seed * (i + 3) + threadIdx.x; a[i] = Acc{0ull, 0u}; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mac<true>(a[i], v[i], c);
        c = c * 1664525u + 1013904223u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= (uint32_t)a[i].lo ^ (uint32_t)(a[i].lo >> 32) ^ a[i].a2;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

/* ------------------------------------------------------------------ */
/* the hot kernel, dataflow-scheduled (one launch per chunk group)      */
/* ------------------------------------------------------------------ */

/* Instead of one launch per dependency level, a persistent grid pulls work
 * items from an atomic counter in topological (level) order and waits per
 * SOURCE COLUMN on a ready flag, as in synchronisation-free sparse triangular
 * solves.  An item is one long column (whole CTA) or up to 8 short columns
 * (one warp each).  Items are handed out in increasing order, a column only
 * depends on columns of earlier items, hence the oldest unfinished item never
 * waits: no deadlock whatever the number of resident CTAs. */

/* ready word of a column: (epoch << 1) | nonzero; columns without sources (V[j] = X[j])
 * are always ready: kFlagAlwaysZ when no lower row of the chunk touches them, else kFlagAlwaysNZ */
constexpr uint32_t kFlagAlwaysZ  = 0xfffffffeu;
constexpr uint32_t kFlagAlwaysNZ = 0xffffffffu;
constexpr uint32_t kSpinLimit  = 1u << 24;      /* polls before the level pass gives up (seconds) */
constexpr uint32_t kErrWatchdog = 16u;

/* compile-time tuning knobs of k_pull_flow (A/B builds: tools/build_variants.sh) */
#ifndef F4LA_FLOW_OCC
#define F4LA_FLOW_OCC 4      /* resident CTAs per SM for T <= 2 (register budget 64) */
#endif
#ifndef F4LA_FLOW_OFF32
#define F4LA_FLOW_OFF32 1
#endif
#ifndef F4LA_FLOW_LOAD
#define F4LA_FLOW_LOAD __ldcg   /* source vectors: L2 only (A/B: ld_ca_u4 keeps them in L1 as well) */
#endif
#ifndef F4LA_FLOW_ACQUIRE_FENCE
#define F4LA_FLOW_ACQUIRE_FENCE 0
#endif
#ifndef F4LA_FLOW_U
#define F4LA_FLOW_U 4        /* entries (source vectors) in flight per warp for T <= 2 */
#endif

struct FlowArgs {
    const uint32_t *col_ptr;
    const uint2    *ent;
    uint32_t       *V;          /* [G][nc][Rc]                                   */
    uint32_t        nc, Rc;
    const uint32_t *sched;      /* column ids in schedule order                   */
    const uint32_t *items;      /* (pos << 5) | (count << 1) | is_long            */
    uint32_t        n_items, G;
    uint32_t       *flags;      /* [G][nc] ready word of each column              */
    uint32_t        epoch;
    uint32_t       *counter;    /* work counter, zeroed before the launch         */
    uint32_t       *err;        /* error bits (kErrWatchdog)                      */
    uint32_t        ncl;        /* columns >= ncl go to out                       */
    uint32_t       *out;
    uint64_t        out_ld;
    uint32_t        chunk0;
    uint32_t        p;
    uint32_t        two64;      /* 2^64 mod p                                     */
    uint64_t        pinv;
    uint64_t        watchdog_ns; /* a wait for a source column longer than this sets kErrWatchdog */
    const uint64_t *piv_row_ptr; /* row_ptr of the pivot rows (pivot j = row j): row length = cost of one application; or NULL */
    unsigned long long *ref_units; /* [kRefSlots][4]: [.][0] += reference-equivalent coefficient applications, [.][1] += pivot applications */
};

/* Ready words are polled with a relaxed, L1-bypassing load.  An acquire load would
 * invalidate the whole L1 on every poll (CCTL.IVALL, 5 % of the stall samples in ncu) although
 * nothing read through L1 can be stale: the vectors V[src] are fetched with ld.cg straight from
 * L2, the point of coherence, and only after the branch that consumed the ready word, while the
 * producer makes its stores visible (fence) before the release store of that word. */
__device__ __forceinline__ uint32_t ld_relaxed_flag(const uint32_t *ptr)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}

__device__ __forceinline__ uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release(uint32_t *ptr, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(ptr), "r"(v) : "memory");
}

__device__ __forceinline__ uint4 ld_ca_u4(const uint4 *ptr)
{
    uint4 v;
    asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
    return v;
}

__device__ __forceinline__ uint4 ld_cg_u4(const uint4 *ptr)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
    return v;
}

/* Coefficient applications of the REFERENCE algorithm: a known pivot j is applied to every lower row with a
 * non-zero multiplier V[j][r], at the cost of its row length (the reference counts len per application,
 * la_ff_32.c:1214-1216).  Fire-and-forget reductions into kRefSlots x {units, applications} spread over
 * separate 32-byte sectors (one address would serialise ~10^6 of them); the host adds the slots up.
 * NOTE: the kernel must keep its single `return` inside the work loop -- with a flush after the loop (a second,
 * predicated EXIT) ptxas 12.9 produced code that faulted at run time ("illegal instruction"). */
constexpr uint32_t kRefSlots = 64;
struct FlowArgs;
__device__ __forceinline__ void count_ref_units(const FlowArgs &A, uint32_t j, uint32_t nnz);

/* Entries [b0,b1) of one column in batches of 32; waits for every source. */
template <int T, bool WIDE>
__device__ __forceinline__ void flow_accumulate(const FlowArgs &A, const uint32_t *__restrict__ Vc,
                                                const uint32_t *__restrict__ flagc,
                                                uint32_t b0, uint32_t b1,
                                                uint32_t lane, Acc (&acc)[T][4])
{
    constexpr int U = (T <= 2) ? F4LA_FLOW_U : 2;
    constexpr uint32_t Rc = T * kRowsPerT;         /* == A.Rc; a constant turns the address products into shifts */
    for (uint32_t b = b0; b < b1; b += 32) {
        const uint32_t n = min(32u, b1 - b);
        uint2 en = make_uint2(0, 0);
        uint32_t f = kFlagAlwaysZ;
        if (lane < n) {
            en = __ldg(&A.ent[b + lane]);
            f = ld_relaxed_flag(&flagc[en.x]);
        }
        bool ready = f >= kFlagAlwaysZ || (f >> 1) == A.epoch;
        uint32_t spins = 0;
        uint64_t t_wait0 = 0;
        while (!__all_sync(0xffffffffu, ready)) {
            if (!ready) {
                __nanosleep(64);
                f = ld_relaxed_flag(&flagc[en.x]);
                ready = f >= kFlagAlwaysZ || (f >> 1) == A.epoch;
#ifdef F4LA_OLD_WATCHDOG
                if (++spins > kSpinLimit) { atomicOr(A.err, kErrWatchdog); ready = true; }
                if (false) {
#else
                if ((++spins & 1023u) == 0) {        /* watchdog on wall time: never hang the GPU */
#endif
                    const uint64_t now = global_timer_ns();
                    if (t_wait0 == 0) t_wait0 = now;
                    else if (now - t_wait0 > A.watchdog_ns) {
                        atomicOr(A.err, kErrWatchdog);
                        ready = true;
                    }
                }
            }
        }
        /* -DF4LA_FLOW_ACQUIRE_FENCE=1: relaxed polls + ONE fence per batch make the formal acquire pattern of the
         * PTX memory model.  Off by default: the fence compiles to MEMBAR.SC.GPU + CCTL.IVALL (an L1 invalidation
         * per batch) and costs 5.4 % of the katsura-12 step (134.6 vs 127.7 ms, same GPU, round 2), while nothing
         * here can be stale in L1: the ready word is polled with ld.relaxed.gpu and the vectors are read with ld.cg,
         * both served by L2 (the point of coherence), the vector loads are issued after the branch that consumed
         * the ready word (the SM does not speculate loads past a branch), and the producer orders its vector stores
         * before the st.release of the word with a fence. */
#if F4LA_FLOW_ACQUIRE_FENCE
        __threadfence();
#endif
        __syncwarp();
        /* Sources whose vector is entirely zero in this chunk contribute nothing:
         * the lower rows come clustered, 15-60 % of the (pivot, chunk) vectors are
         * zero, so their loads and multiplies are skipped altogether. */
        uint32_t live = __ballot_sync(0xffffffffu, lane < n && (f & 1u));
        while (live) {
            uint32_t src[U], cf[U];
            uint4 v[U][T];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                /* next live entry; when the group runs out, repeat the last one with coefficient 0 */
                const bool ok = live != 0;
                const uint32_t idx = ok ? (uint32_t)__ffs(live) - 1 : 0;
                live &= live - (ok ? 1u : 0u);
                const uint32_t sx = __shfl_sync(0xffffffffu, en.x, idx);
                const uint32_t sy = __shfl_sync(0xffffffffu, en.y, idx);
                src[u] = ok ? sx : (u ? src[u - 1] : 0u);
                cf[u]  = ok ? sy : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#if F4LA_FLOW_OFF32
                /* 32-bit element offset (nc * Rc < 2^32, checked by the host): shifts on the uniform / ALU pipes instead of
                 * a 64-bit IMAD.WIDE on the multiplier pipe */
                const uint4 *sp = reinterpret_cast<const uint4 *>(Vc + (uint32_t)(src[u] * Rc)) + lane;
#else
                const uint4 *sp = reinterpret_cast<const uint4 *>(Vc + (size_t)src[u] * Rc) + lane;
#endif
#pragma unroll
                for (int t = 0; t < T; ++t) v[u][t] = F4LA_FLOW_LOAD(sp + t * 32);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    mac<WIDE>(acc[t][0], v[u][t].x, cf[u]); mac<WIDE>(acc[t][1], v[u][t].y, cf[u]);
                    mac<WIDE>(acc[t][2], v[u][t].z, cf[u]); mac<WIDE>(acc[t][3], v[u][t].w, cf[u]);
                }
            }
        }
    }
}

__device__ __forceinline__ void count_ref_units(const FlowArgs &A, uint32_t j, uint32_t nnz)
{
    unsigned long long *slot = A.ref_units + (size_t)(j % kRefSlots) * 4;
    atomicAdd(slot, (unsigned long long)nnz * (A.piv_row_ptr[j + 1] - A.piv_row_ptr[j]));
    atomicAdd(slot + 1, (unsigned long long)nnz);
}

template <int T, bool WIDE>
__global__ void __launch_bounds__(kSolveThreads, (T <= 2) ? F4LA_FLOW_OCC : 2)
k_pull_flow(const FlowArgs A)
{
    __shared__ uint32_t part[kWarpsPerBlock][T * kRowsPerT];
    __shared__ uint32_t s_id;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t total = A.n_items * A.G;
    for (;;) {
        if (threadIdx.x == 0) s_id = atomicAdd(A.counter, 1u);
        __syncthreads();
        const uint32_t id = s_id;
        __syncthreads();
        if (id >= total) return;
        const uint32_t chunk = id % A.G;
        const uint32_t it = A.items[id / A.G];
        const uint32_t pos = it >> 5, cnt = (it >> 1) & 15u;
        const bool is_long = it & 1u;
        uint32_t *Vc = A.V + (size_t)chunk * A.nc * A.Rc;
        uint32_t *flagc = A.flags + (size_t)chunk * A.nc;

        Acc acc[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[t][k] = Acc{0ull, 0u};

        if (!is_long) {
            if (warp < cnt) {
                const uint32_t j = A.sched[pos + warp];
                const uint32_t e0 = A.col_ptr[j], e1 = A.col_ptr[j + 1];
                flow_accumulate<T, WIDE>(A, Vc, flagc, e0, e1, lane, acc);
                const bool to_out = j >= A.ncl;
                uint32_t *dst = to_out ? A.out + (size_t)(j - A.ncl) * A.out_ld + (size_t)(A.chunk0 + chunk) * A.Rc
                                       : Vc + (size_t)j * A.Rc;
                const uint4 *init = reinterpret_cast<const uint4 *>(Vc + (size_t)j * A.Rc) + lane;
                uint32_t nonzero = 0, nnz = 0;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const uint4 x = init[t * 32];
                    uint4 r;
                    r.x = fold_acc<WIDE>(acc[t][0], x.x, A.p, A.pinv, A.two64);
                    r.y = fold_acc<WIDE>(acc[t][1], x.y, A.p, A.pinv, A.two64);
                    r.z = fold_acc<WIDE>(acc[t][2], x.z, A.p, A.pinv, A.two64);
                    r.w = fold_acc<WIDE>(acc[t][3], x.w, A.p, A.pinv, A.two64);
                    reinterpret_cast<uint4 *>(dst)[lane + t * 32] = r;
                    nonzero |= r.x | r.y | r.z | r.w;
                    nnz += (r.x != 0) + (r.y != 0) + (r.z != 0) + (r.w != 0);
                }
                if (!to_out) {
                    if (A.piv_row_ptr) {
                        nnz = __reduce_add_sync(0xffffffffu, nnz);
                        if (lane == 0 && nnz) count_ref_units(A, j, nnz);
                    }
                    const bool any = __any_sync(0xffffffffu, nonzero != 0);
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) st_release(&flagc[j], (A.epoch << 1) | (any ? 1u : 0u));
                }
            }
            continue;       /* the two barriers at the loop head keep the CTA together */
        }

        const uint32_t j = A.sched[pos];
        const uint32_t e0 = A.col_ptr[j], e1 = A.col_ptr[j + 1];
        {
            /* equal contiguous shares of the entries for the 8 warps */
            const uint32_t seg = (e1 - e0 + kWarpsPerBlock - 1) / kWarpsPerBlock;
            const uint32_t b0 = min(e1, e0 + warp * seg), b1 = min(e1, b0 + seg);
            flow_accumulate<T, WIDE>(A, Vc, flagc, b0, b1, lane, acc);
        }
#pragma unroll
        for (int t = 0; t < T; ++t) {
            uint4 r;
            r.x = fold_acc<WIDE>(acc[t][0], 0u, A.p, A.pinv, A.two64);
            r.y = fold_acc<WIDE>(acc[t][1], 0u, A.p, A.pinv, A.two64);
            r.z = fold_acc<WIDE>(acc[t][2], 0u, A.p, A.pinv, A.two64);
            r.w = fold_acc<WIDE>(acc[t][3], 0u, A.p, A.pinv, A.two64);
            reinterpret_cast<uint4 *>(&part[warp][t * kRowsPerT])[lane] = r;
        }
        __syncthreads();
        const bool to_out = j >= A.ncl;
        uint32_t *dst = to_out ? A.out + (size_t)(j - A.ncl) * A.out_ld + (size_t)(A.chunk0 + chunk) * A.Rc
                               : Vc + (size_t)j * A.Rc;
        int nonzero = 0;
        for (uint32_t r = threadIdx.x; r < (uint32_t)(T * kRowsPerT); r += kSolveThreads) {
            uint64_t s = Vc[(size_t)j * A.Rc + r];
#pragma unroll
            for (int w = 0; w < kWarpsPerBlock; ++w) s += part[w][r];
            const uint32_t v = mod_u64(s, A.p, A.pinv);
            dst[r] = v;
            nonzero += v != 0;
        }
        if (!to_out) {
            if (A.piv_row_ptr) {
                const uint32_t cnt = __reduce_add_sync(0xffffffffu, (uint32_t)nonzero);
                if (lane == 0 && cnt) count_ref_units(A, j, cnt);
            }
            __threadfence();
            nonzero = __syncthreads_or(nonzero);
            if (threadIdx.x == 0) st_release(&flagc[j], (A.epoch << 1) | (nonzero ? 1u : 0u));
        }
    }
}

/* ready words of a launch group (g chunk slots): columns without sources are "always ready, zero"
 * (k_scatter_lower upgrades the touched ones to "non-zero"), every other word is reset to "not ready";
 * also zeroes the work counter of the dataflow kernel (one launch instead of two memsets and a kernel). */
__global__ void k_flow_reset_flags(const uint32_t *__restrict__ col_ptr, uint32_t ncl, uint32_t nc, uint32_t g,
                                   uint32_t *__restrict__ flags, uint32_t *__restrict__ counter)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) *counter = 0;                    /* work counter of the dataflow kernel that follows */
    if (j >= nc) return;
    /* every ready word of the launch group: "not ready", except pivot columns without sources (never solved: zero) */
    const uint32_t v = (j < ncl && col_ptr[j + 1] == col_ptr[j]) ? kFlagAlwaysZ : 0u;
    for (uint32_t k = 0; k < g; ++k) flags[(size_t)k * nc + j] = v;
}

/* schedule = pivot columns with sources in level order, then the non-pivot columns */
__global__ void k_flow_sched(const uint32_t *__restrict__ order, uint32_t first, uint32_t ncl, uint32_t ncr,
                             uint32_t *__restrict__ sched)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t na = ncl - first;
    if (k < na) sched[k] = order[first + k];
    else if (k < na + ncr) sched[k] = ncl + (k - na);
}

/* ------------------------------------------------------------------ */
/* trailing block: Gauss-Jordan on D' (column major, ld rows per col)   */
/* ------------------------------------------------------------------ */

struct GEState {
    uint32_t next_col;   /* first column not yet examined                    */
    uint32_t rank;
    uint32_t done;
    uint32_t cur_col;    /* pivot of the current elimination step            */
    uint32_t cur_row;
    uint32_t cur_inv;
    uint32_t active;     /* 1 when cur_* describe a pivot to be eliminated   */
    uint32_t visits;     /* columns loaded by the discovery (statistics)     */
    uint32_t batches;    /* non-empty batches (statistics)                   */
    uint32_t pad;
    unsigned long long t[10];  /* k_ge_fused, CTA 0, clock cycles: gather, load, select + inverse, eliminate, write-back,
                                  update, barrier waits, total; [8] longest update of any CTA, [9] longest first wait */
};

/* colflag[c] = 1 iff column c has a non-zero entry in a row that is not a
 * pivot row yet.  Initially: any non-zero. */
__global__ void k_ge_init_flags(const uint32_t *__restrict__ D, uint64_t ld, uint32_t nrows,
                                uint32_t ncr, uint8_t *__restrict__ colflag)
{
    const uint32_t c = blockIdx.x;
    if (c >= ncr) return;
    int any = 0;
    for (uint32_t r = threadIdx.x; r < nrows; r += blockDim.x) any |= D[(size_t)c * ld + r] != 0;
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) colflag[c] = (uint8_t)(any != 0);
}

/* Next pivot: first flagged column >= next_col, in it the first row that is
 * not yet a pivot row and non-zero.  One block. */
__global__ void k_ge_find(const uint32_t *__restrict__ D, uint64_t ld, uint32_t nrows, uint32_t ncr,
                          const uint8_t *__restrict__ colflag, uint8_t *__restrict__ used,
                          uint8_t *__restrict__ ispiv, GEState *st,
                          uint32_t *__restrict__ pivcol, uint32_t *__restrict__ pivrow,
                          uint32_t *__restrict__ pivinv, uint32_t p)
{
    __shared__ uint32_t s_min;
    if (st->done) return;
    uint32_t c = st->next_col;
    __syncthreads();                      /* everybody has read the state */
    for (;;) {
        /* first flagged column at or after c */
        uint32_t cand = 0xffffffffu;
        for (uint32_t base = c; base < ncr; base += blockDim.x) {
            const uint32_t k = base + threadIdx.x;
            const int hit = k < ncr && colflag[k];
            if (threadIdx.x == 0) s_min = 0xffffffffu;
            __syncthreads();
            if (hit) atomicMin(&s_min, k);
            __syncthreads();
            cand = s_min;
            __syncthreads();
            if (cand != 0xffffffffu) break;
        }
        if (cand == 0xffffffffu) {
            if (threadIdx.x == 0) { st->done = 1; st->active = 0; st->next_col = ncr; }
            return;
        }
        c = cand;
        if (threadIdx.x == 0) s_min = 0xffffffffu;
        __syncthreads();
        for (uint32_t r = threadIdx.x; r < nrows; r += blockDim.x)
            if (!used[r] && D[(size_t)c * ld + r] != 0) { atomicMin(&s_min, r); break; }
        __syncthreads();
        const uint32_t piv = s_min;
        __syncthreads();
        if (piv != 0xffffffffu) {
            if (threadIdx.x == 0) {
                const uint32_t k = st->rank;
                const uint32_t inv = modinv(D[(size_t)c * ld + piv], p);
                pivcol[k] = c; pivrow[k] = piv; pivinv[k] = inv;
                used[piv] = 1; ispiv[c] = 1;
                st->rank = k + 1; st->cur_col = c; st->cur_row = piv; st->cur_inv = inv;
                st->active = 1; st->next_col = c + 1;
            }
            return;
        }
        c = c + 1;                        /* stale flag (cannot happen, but be safe) */
    }
}

/* Eliminate the current pivot from every other row, columns right of it.
 * One block per column.  Also refreshes colflag for the touched columns. */
__global__ void k_ge_eliminate(uint32_t *__restrict__ D, uint64_t ld, uint32_t nrows, uint32_t ncr,
                               uint8_t *__restrict__ colflag, const uint8_t *__restrict__ used,
                               const GEState *__restrict__ st, uint32_t p, uint64_t pinv)
{
    if (st->done || !st->active) return;
    const uint32_t c = st->cur_col, piv = st->cur_row;
    const uint32_t cc = blockIdx.x;
    if (cc <= c || cc >= ncr) return;
    uint32_t *col = D + (size_t)cc * ld;
    const uint32_t *fcol = D + (size_t)c * ld;
    const uint32_t s = mulmod(col[piv], st->cur_inv, p, pinv);
    if (s == 0) return;                   /* column untouched, flag still valid */
    int any = 0;
    for (uint32_t r = threadIdx.x; r < nrows; r += blockDim.x) {
        uint32_t v = col[r];
        if (r != piv) {
            const uint32_t f = fcol[r];
            if (f) {
                const uint32_t t = mulmod(f, s, p, pinv);
                v = v >= t ? v - t : v + p - t;
                col[r] = v;
            }
            any |= (v != 0) && !used[r];
        }
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) colflag[cc] = (uint8_t)(any != 0);
}

constexpr int kUpdateThreads = 256;

/* ---- batched Gauss-Jordan: discover up to kGeBatch pivots, then one update pass ---- */
/* k_ge_discover (one CTA) walks the flagged columns left to right; a visited column is loaded
 * into shared memory, brought up to date with the pivots found EARLIER IN THIS BATCH (the
 * pivots of previous batches were applied by k_ge_batch_update), and searched for a pivot.
 * A pivot column is written back (it is final and carries the multipliers of its pivot);
 * a column without pivot is discarded, the update pass recomputes it.
 * k_ge_batch_update (one CTA per column) then applies the batch to every non-pivot column
 * right of the batch's first pivot and refreshes the column flags.  Compared with one
 * find + one full elimination pass per pivot this divides the number of passes over the
 * block by the batch size. */
constexpr uint32_t kGeBatch = 8;
constexpr int kDiscoverThreads = 1024;

/* dynamic smem: [nrows] column | [nrows] used bytes */
__global__ void __launch_bounds__(kDiscoverThreads)
k_ge_discover(uint32_t *D, uint64_t ld, uint32_t nrows, uint32_t ncr, const uint8_t *colflag, uint8_t *used,
              uint8_t *ispiv, GEState *st, uint32_t *pivcol, uint32_t *pivrow, uint32_t *pivinv,
              uint32_t p, uint64_t pinv)
{
    extern __shared__ uint32_t s_col[];
    uint8_t *s_used = reinterpret_cast<uint8_t *>(s_col + nrows);
    __shared__ uint32_t s_min, s_inv;
    const uint32_t tid = threadIdx.x;
    if (st->done) {                        /* nothing left: make the following update pass a no-op */
        if (tid == 0) st->cur_col = st->rank;
        return;
    }
    const uint32_t k0 = st->rank;
    uint32_t c = st->next_col;
    for (uint32_t r = tid; r < nrows; r += kDiscoverThreads) s_used[r] = used[r];
    __syncthreads();
    uint32_t np = 0, misses = 0;
    bool exhausted = false;
    /* the flags are as of the last update pass: a flagged column may turn out empty once the
     * pivots of this batch are applied.  Two such misses in a row end the batch, the update
     * pass then refreshes every flag (this is what ends the elimination when the rank is
     * exhausted: no column stays flagged). */
    while (np < kGeBatch && misses < 2) {
        /* next flagged column */
        if (tid == 0) s_min = 0xffffffffu;
        __syncthreads();
        for (uint32_t base = c; base < ncr; base += kDiscoverThreads) {
            const uint32_t k = base + tid;
            if (k < ncr && colflag[k]) atomicMin(&s_min, k);
            __syncthreads();
            if (s_min != 0xffffffffu) break;           /* uniform: read after the barrier */
            __syncthreads();
        }
        __syncthreads();
        c = s_min;
        __syncthreads();
        if (c == 0xffffffffu) { exhausted = true; break; }
        uint32_t *g = D + (size_t)c * ld;
        for (uint32_t r = tid; r < nrows; r += kDiscoverThreads) s_col[r] = g[r];
        __syncthreads();
        for (uint32_t k = k0; k < k0 + np; ++k) {       /* pivots of this batch, all left of c */
            const uint32_t piv = pivrow[k];
            const uint32_t sc = mulmod(s_col[piv], pivinv[k], p, pinv);
            __syncthreads();                            /* everybody has read s_col[piv] */
            if (sc != 0) {
                const uint32_t *fcol = D + (size_t)pivcol[k] * ld;
                const uint32_t w = shoup_pre(sc, p);
                for (uint32_t r = tid; r < nrows; r += kDiscoverThreads) {
                    const uint32_t f = __ldcg(&fcol[r]);
                    if (f && r != piv) {
                        const uint32_t t = shoup_mul(f, sc, w, p);
                        const uint32_t v = s_col[r];
                        s_col[r] = v >= t ? v - t : v + p - t;
                    }
                }
            }
            __syncthreads();
        }
        if (tid == 0) s_min = 0xffffffffu;
        __syncthreads();
        for (uint32_t r = tid; r < nrows; r += kDiscoverThreads)
            if (s_col[r] != 0 && !s_used[r]) { atomicMin(&s_min, r); break; }
        __syncthreads();
        const uint32_t piv = s_min;
        if (piv != 0xffffffffu) {
            for (uint32_t r = tid; r < nrows; r += kDiscoverThreads) g[r] = s_col[r];
            if (tid == 0) {
                const uint32_t inv = modinv(s_col[piv], p);
                pivcol[k0 + np] = c; pivrow[k0 + np] = piv; pivinv[k0 + np] = inv;
                used[piv] = 1; s_used[piv] = 1; ispiv[c] = 1;
            }
            __threadfence();
            np++;
            misses = 0;
        } else {
            misses++;
        }
        __syncthreads();
        c = c + 1;
    }
    if (tid == 0) {
        st->cur_col = k0;                 /* batch = pivots [cur_col, rank) */
        st->rank = k0 + np;
        st->next_col = exhausted ? ncr : c;
        if (exhausted) st->done = 1;
    }
}

__global__ void __launch_bounds__(kUpdateThreads)
k_ge_batch_update(uint32_t *__restrict__ D, uint64_t ld, uint32_t nrows, uint32_t ncr,
                  uint8_t *__restrict__ colflag, const uint8_t *__restrict__ used,
                  const uint8_t *__restrict__ ispiv, const GEState *__restrict__ st,
                  const uint32_t *__restrict__ pivcol, const uint32_t *__restrict__ pivrow,
                  const uint32_t *__restrict__ pivinv, uint32_t p, uint64_t pinv)
{
    extern __shared__ uint32_t s_col[];
    const uint32_t k0 = st->cur_col, k1 = st->rank;
    if (k0 >= k1) return;
    const uint32_t c = blockIdx.x;
    if (c >= ncr || ispiv[c] || c <= pivcol[k0]) return;
    uint32_t *g = D + (size_t)c * ld;
    for (uint32_t r = threadIdx.x; r < nrows; r += kUpdateThreads) s_col[r] = g[r];
    __syncthreads();
    bool touched = false;
    for (uint32_t k = k0; k < k1; ++k) {
        if (pivcol[k] >= c) break;                       /* pivot columns ascend within a batch */
        const uint32_t piv = pivrow[k];
        const uint32_t sc = mulmod(s_col[piv], pivinv[k], p, pinv);
        __syncthreads();
        if (sc != 0) {                                   /* uniform over the CTA */
            touched = true;
            const uint32_t *fcol = D + (size_t)pivcol[k] * ld;
            const uint32_t w = shoup_pre(sc, p);
            for (uint32_t r = threadIdx.x; r < nrows; r += kUpdateThreads) {
                const uint32_t f = __ldg(&fcol[r]);
                if (f && r != piv) {
                    const uint32_t t = shoup_mul(f, sc, w, p);
                    const uint32_t v = s_col[r];
                    s_col[r] = v >= t ? v - t : v + p - t;
                }
            }
        }
        __syncthreads();
    }
    int any = 0;
    for (uint32_t r = threadIdx.x; r < nrows; r += kUpdateThreads) {
        const uint32_t v = s_col[r];
        if (touched) g[r] = v;
        any |= (v != 0) && !used[r];
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) colflag[c] = (uint8_t)(any != 0);
}

/* ---- the same two passes with the batch's pivot columns held in shared memory ---- */
/* The pivot columns of a batch are the only data both passes re-read: every visited column of
 * the discovery and every column of the update pass needs all of them.  Here they live in
 * shared memory (nb * nrows words, nb <= kGeBatch chosen by the host so that they fit):
 *  - k_ge_discover_cached applies the batch's earlier pivots to a visited column from there;
 *  - k_ge_batch_update_cached loads them once per CTA (one CTA per SM, 32 warps), then every warp
 *    streams whole columns through registers: the nb multipliers of a column come from an
 *    nb-step scalar recurrence on its pivot-row entries, after which each element is
 *    v + sum_k (p - s_k) * piv_k[r] in one 96-bit accumulation and one reduction. */
constexpr int kUpdate2Threads = 1024;
constexpr uint32_t kGeBatchCached = 32;      /* most pivots per batch (when nrows is small enough for shared memory) */
constexpr uint32_t kGeMissLimit = 32;        /* consecutive empty visited columns that end a batch (a visit costs a few us, a batch a full update pass) */

/* dynamic smem: [nrows4] column | [nb][nrows4] pivot columns of the batch (own pivot-row entry zeroed) |
 * [nrows4] used bytes.  Per visited column the serial chain is: next flagged column (flags cached in a
 * shared-memory window), column load, the multipliers of the batch's earlier pivots by one warp (the
 * same lane-distributed recurrence as in the update pass), ONE fused pass applying them, pivot search. */
constexpr uint32_t kFlagWindow = 4096;

__global__ void __launch_bounds__(kDiscoverThreads)
k_ge_discover_cached(uint32_t *D, uint64_t ld, uint32_t nrows, uint32_t nrows4, uint32_t ncr, uint32_t nb, uint32_t miss_limit,
                     const uint8_t *colflag, uint8_t *used, uint8_t *ispiv, GEState *st,
                     uint32_t *pivcol, uint32_t *pivrow, uint32_t *pivinv, uint32_t p, uint64_t pinv, uint32_t two64)
{
    extern __shared__ __align__(16) uint32_t s_col[];
    uint32_t *s_piv = s_col + nrows4;
    uint8_t *s_used = reinterpret_cast<uint8_t *>(s_piv + (size_t)nb * nrows4);
    __shared__ uint32_t s_min, s_prow[kGeBatchCached], s_pinv[kGeBatchCached], s_neg[kGeBatchCached];
    __shared__ uint32_t s_x[kGeBatchCached][kGeBatchCached + 1];      /* x[k][j] = piv_k[row of pivot j], k < j */
    __shared__ uint8_t s_flag[kFlagWindow];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    if (st->done) {
        if (tid == 0) st->cur_col = st->rank;
        return;
    }
    const uint32_t k0 = st->rank;
    uint32_t c = st->next_col;
    for (uint32_t r = tid; r < nrows; r += kDiscoverThreads) s_used[r] = used[r];
    uint32_t np = 0, misses = 0, win0 = c, win1 = c, visits = 0;
    bool exhausted = false;
    while (np < nb && misses < miss_limit) {
        /* next flagged column at or after c */
        for (;;) {
            if (c >= ncr) break;
            if (c >= win1) {
                win0 = c; win1 = min(ncr, c + kFlagWindow);
                __syncthreads();
                for (uint32_t i = tid; i < win1 - win0; i += kDiscoverThreads) s_flag[i] = colflag[win0 + i];
            }
            if (tid == 0) s_min = 0xffffffffu;
            __syncthreads();
            for (uint32_t k = c + tid; k < win1; k += kDiscoverThreads)
                if (s_flag[k - win0]) { atomicMin(&s_min, k); break; }
            __syncthreads();
            const uint32_t found = s_min;
            __syncthreads();
            if (found != 0xffffffffu) { c = found; break; }
            c = win1;
        }
        if (c >= ncr) { exhausted = true; break; }
        ++visits;
        uint32_t *g = D + (size_t)c * ld;
        for (uint32_t r = tid; r < nrows; r += kDiscoverThreads) s_col[r] = __ldcg(&g[r]);
        __syncthreads();
        if (np) {
            if (tid < 32) {
                uint32_t e = lane < np ? s_col[s_prow[lane]] : 0u, neg = 0;
                for (uint32_t k = 0; k < np; ++k) {
                    const uint32_t ek = __shfl_sync(0xffffffffu, e, k);
                    if (ek == 0) continue;
                    const uint32_t ng = p - mulmod(ek, s_pinv[k], p, pinv);
                    if (lane == k) neg = ng;
                    if (lane > k && lane < np) e = mod_u64((uint64_t)e + (uint64_t)ng * s_x[k][lane], p, pinv);
                }
                s_neg[lane] = neg;
            }
            __syncthreads();
            for (uint32_t r = tid; r < nrows; r += kDiscoverThreads) {
                Acc a = {(uint64_t)s_col[r], 0u};
                for (uint32_t k = 0; k < np; ++k) {
                    const uint32_t nk = s_neg[k];
                    if (nk) mac<true>(a, s_piv[(size_t)k * nrows4 + r], nk);
                }
                s_col[r] = fold_acc<true>(a, 0u, p, pinv, two64);
            }
            __syncthreads();
        }
        if (tid == 0) s_min = 0xffffffffu;
        __syncthreads();
        for (uint32_t r = tid; r < nrows; r += kDiscoverThreads)
            if (s_col[r] != 0 && !s_used[r]) { atomicMin(&s_min, r); break; }
        __syncthreads();
        const uint32_t piv = s_min;
        if (piv != 0xffffffffu) {
            uint32_t *keep = s_piv + (size_t)np * nrows4;
            for (uint32_t r = tid; r < nrows; r += kDiscoverThreads) {
                const uint32_t v = s_col[r];
                g[r] = v;
                keep[r] = r == piv ? 0u : v;
            }
            if (tid < np) s_x[tid][np] = s_piv[(size_t)tid * nrows4 + piv];
            if (tid == 32) {                        /* a warp with no share in the s_x update above */
                const uint32_t inv = modinv(s_col[piv], p);
                pivcol[k0 + np] = c; pivrow[k0 + np] = piv; pivinv[k0 + np] = inv;
                s_prow[np] = piv; s_pinv[np] = inv;
                used[piv] = 1; s_used[piv] = 1; ispiv[c] = 1;
            }
            np++;
            misses = 0;
        } else {
            misses++;
        }
        __syncthreads();
        c = c + 1;
    }
    if (tid == 0) {
        st->cur_col = k0;
        st->rank = k0 + np;
        st->next_col = exhausted ? ncr : c;
        if (exhausted) st->done = 1;
        st->visits += visits;
        st->batches += np != 0;
    }
}

/* dynamic smem: [nb][nrows4] pivot columns (own pivot-row entry zeroed) | [nrows4] used bytes */
__global__ void __launch_bounds__(kUpdate2Threads, 1)
k_ge_batch_update_cached(uint32_t *__restrict__ D, uint64_t ld, uint32_t nrows, uint32_t nrows4, uint32_t ncr,
                         uint32_t nbmax, uint8_t *__restrict__ colflag, const uint8_t *__restrict__ used,
                         const uint8_t *__restrict__ ispiv, const GEState *__restrict__ st,
                         const uint32_t *__restrict__ pivcol, const uint32_t *__restrict__ pivrow,
                         const uint32_t *__restrict__ pivinv, uint32_t p, uint64_t pinv, uint32_t two64)
{
    extern __shared__ __align__(16) uint32_t s_piv[];
    uint8_t *s_used = reinterpret_cast<uint8_t *>(s_piv + (size_t)nbmax * nrows4);
    __shared__ uint32_t s_pc[kGeBatchCached], s_pr[kGeBatchCached], s_pi[kGeBatchCached];
    __shared__ uint32_t s_x[kGeBatchCached][kGeBatchCached + 1];      /* x[k][j] = piv_k[row of pivot j] */
    __shared__ uint32_t s_neg[kUpdate2Threads / 32][kGeBatchCached]; /* per warp: negated multipliers of its column */
    const uint32_t k0 = st->cur_col, k1 = st->rank;
    if (k0 >= k1) return;
    const uint32_t nb = k1 - k0;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t first = pivcol[k0];
    const uint32_t nwarps = gridDim.x * (kUpdate2Threads / 32);
    const uint32_t gwarp = blockIdx.x * (kUpdate2Threads / 32) + (tid >> 5);
    if (first + 1 + blockIdx.x * (kUpdate2Threads / 32) >= ncr) return;      /* no column for this CTA */
    if (tid < nb) { s_pc[tid] = pivcol[k0 + tid]; s_pr[tid] = pivrow[k0 + tid]; s_pi[tid] = pivinv[k0 + tid]; }
    __syncthreads();
    for (uint32_t k = 0; k < nb; ++k) {
        const uint32_t *src = D + (size_t)s_pc[k] * ld;
        const uint32_t own = s_pr[k];
        for (uint32_t r = tid; r < nrows4; r += kUpdate2Threads)
            s_piv[(size_t)k * nrows4 + r] = (r < nrows && r != own) ? __ldcg(&src[r]) : 0u;
    }
    for (uint32_t r = tid; r < nrows4; r += kUpdate2Threads) s_used[r] = r < nrows ? used[r] : (uint8_t)1;
    __syncthreads();
    for (uint32_t q = tid; q < kGeBatchCached * kGeBatchCached; q += kUpdate2Threads) {
        const uint32_t k = q / kGeBatchCached, j = q % kGeBatchCached;
        s_x[k][j] = (k < nb && j < nb) ? s_piv[(size_t)k * nrows4 + s_pr[j]] : 0u;
    }
    __syncthreads();

    for (uint32_t c = first + 1 + gwarp; c < ncr; c += nwarps) {
        if (ispiv[c]) continue;
        uint32_t *g = D + (size_t)c * ld;
        /* multipliers of the batch for this column: lane j tracks the column's entry in the row of
         * pivot j; pivot k's multiplier ends up (negated) in lane k */
        uint32_t e = lane < nb ? g[s_pr[lane]] : 0u, neg = 0;
        bool touched = false;
        for (uint32_t k = 0; k < nb; ++k) {
            const uint32_t ek = __shfl_sync(0xffffffffu, e, k);
            if (s_pc[k] >= c || ek == 0) continue;           /* pivot columns ascend within a batch; uniform */
            const uint32_t ng = p - mulmod(ek, s_pi[k], p, pinv);
            touched = true;
            if (lane == k) neg = ng;
            if (lane > k && lane < nb) e = mod_u64((uint64_t)e + (uint64_t)ng * s_x[k][lane], p, pinv);
        }
        /* the lanes run different numbers of row iterations below: no warp collectives in that loop,
         * the multipliers go through shared memory (the shuffles above re-converge the warp before
         * the next column overwrites them) */
        uint32_t *my_neg = s_neg[tid >> 5];
        my_neg[lane] = neg;
        __syncwarp();
        int any = 0;
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (lane * 4 < nrows4) nxt = *reinterpret_cast<const uint4 *>(g + lane * 4);
        for (uint32_t r4 = lane * 4; r4 < nrows4; r4 += 128) {
            uint4 v = nxt;
            if (r4 + 128 < nrows4) nxt = *reinterpret_cast<const uint4 *>(g + r4 + 128);      /* next piece in flight */
            if (touched) {
                Acc a[4] = {{(uint64_t)v.x, 0u}, {(uint64_t)v.y, 0u}, {(uint64_t)v.z, 0u}, {(uint64_t)v.w, 0u}};
                for (uint32_t k = 0; k < nb; ++k) {
                    const uint32_t nk = my_neg[k];
                    if (nk) {
                        const uint4 f = *reinterpret_cast<const uint4 *>(s_piv + (size_t)k * nrows4 + r4);
                        mac<true>(a[0], f.x, nk); mac<true>(a[1], f.y, nk);
                        mac<true>(a[2], f.z, nk); mac<true>(a[3], f.w, nk);
                    }
                }
                v.x = fold_acc<true>(a[0], 0u, p, pinv, two64); v.y = fold_acc<true>(a[1], 0u, p, pinv, two64);
                v.z = fold_acc<true>(a[2], 0u, p, pinv, two64); v.w = fold_acc<true>(a[3], 0u, p, pinv, two64);
                *reinterpret_cast<uint4 *>(g + r4) = v;
            }
            const uchar4 u = *reinterpret_cast<const uchar4 *>(s_used + r4);
            any |= (v.x && !u.x) | (v.y && !u.y) | (v.z && !u.z) | (v.w && !u.w);
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) colflag[c] = (uint8_t)(any != 0);
    }
}

/* ---- compress-and-verify for tall trailing blocks (exact, Las Vegas) ---------------------- */
/* D' has nrl rows but a small rank r.  Instead of eliminating on all nrl rows:
 *   1. D'' = R * D' for a sparse random R (kc x nrl, every row of D' enters `weight` combinations,
 *      as in the probabilistic mode) -- kc rows only;
 *   2. Gauss-Jordan on D'' gives pivot columns pc_k and reduced rows B (r x ncr);
 *   3. VERIFY, exactly, that every row of D' lies in the row space of B: since B is reduced with
 *      pivots at pc_k, row i can only be sum_k D'[i][pc_k] * B_k, so the test is the matrix identity
 *      D' == D'[:, pc] * B.  rowspace(B) is inside rowspace(D') by construction, so a passed test
 *      proves equality and B is THE reduced echelon form (unique) -- bit-identical output.  A failed
 *      test (rank lost in the random compression, probability ~ 1/p per degree of slack) falls back
 *      to the elimination on D' itself, which is untouched.
 * The identity is a dense (nrl x r) x (r x ncr) product on the CUDA cores (96-bit accumulators):
 * fully parallel, no pivot-by-pivot latency chain -- what bounds the direct elimination. */

/* D''[c][q] = sum over the members e of combination q of rho[e] * D'[c][row[e]].  One CTA per column,
 * the column staged in shared memory (nrows4 words). */
__global__ void __launch_bounds__(256)
k_ge_compress(const uint32_t *__restrict__ D, uint64_t ld, uint32_t nrows, uint32_t nrows4, uint32_t kc,
              const uint32_t *__restrict__ slot_ptr, const uint32_t *__restrict__ slot_row,
              const uint32_t *__restrict__ slot_rho, uint32_t *__restrict__ out,
              uint32_t p, uint64_t pinv, uint32_t two64)
{
    extern __shared__ __align__(16) uint32_t s_col[];
    const uint32_t c = blockIdx.x;
    const uint32_t *g = D + (size_t)c * ld;
    int any = 0;
    for (uint32_t r = threadIdx.x; r < nrows4; r += blockDim.x) {
        const uint32_t v = r < nrows ? g[r] : 0u;
        s_col[r] = v;
        any |= v != 0;
    }
    any = __syncthreads_or(any);
    uint32_t *o = out + (size_t)c * kc;
    if (!any) {
        for (uint32_t q = threadIdx.x; q < kc; q += blockDim.x) o[q] = 0;
        return;
    }
    for (uint32_t q = threadIdx.x; q < kc; q += blockDim.x) {
        Acc a = {0ull, 0u};
        const uint32_t e1 = slot_ptr[q + 1];
        for (uint32_t e = slot_ptr[q]; e < e1; ++e) mac<true>(a, s_col[slot_row[e]], slot_rho[e]);
        o[q] = fold_acc<true>(a, 0u, p, pinv, two64);
    }
}

/* mismatch |= any(D'[i][c] != sum_t D'[i][pc(t)] * B[t][c]);  B row t belongs to pivot rank-1-t.
 * 64 x 64 tile per CTA, 4 x 4 per thread, 16 pivots per shared-memory stage. */
constexpr int kVerTile = 64, kVerK = 16;
__global__ void __launch_bounds__(256)
k_ge_verify(const uint32_t *__restrict__ D, uint64_t ld, uint32_t nrows_pad, uint32_t ncr, uint32_t rank,
            const uint32_t *__restrict__ pivcol, const uint32_t *__restrict__ B,
            uint32_t p, uint64_t pinv, uint32_t two64, uint32_t *__restrict__ mismatch)
{
    __shared__ __align__(16) uint32_t sP[kVerK][kVerTile];
    __shared__ __align__(16) uint32_t sB[kVerK][kVerTile];
    const uint32_t i0 = blockIdx.x * kVerTile, c0 = blockIdx.y * kVerTile;
    const uint32_t tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    Acc acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = Acc{0ull, 0u};
    for (uint32_t t0 = 0; t0 < rank; t0 += kVerK) {
        /* stage: 16 x 64 of each operand, 4 words per thread */
        {
            const uint32_t kk = threadIdx.x >> 4, w = (threadIdx.x & 15) * 4;
            const uint32_t t = t0 + kk;
            uint4 pv = make_uint4(0, 0, 0, 0), bv = make_uint4(0, 0, 0, 0);
            if (t < rank) {
                const uint32_t pc = pivcol[rank - 1 - t];
                pv = *reinterpret_cast<const uint4 *>(D + (size_t)pc * ld + i0 + w);
                const uint32_t *brow = B + (size_t)t * ncr + c0 + w;
                if (c0 + w + 3 < ncr && (((size_t)t * ncr + c0 + w) & 3) == 0) bv = *reinterpret_cast<const uint4 *>(brow);
                else {
                    if (c0 + w < ncr) bv.x = brow[0];
                    if (c0 + w + 1 < ncr) bv.y = brow[1];
                    if (c0 + w + 2 < ncr) bv.z = brow[2];
                    if (c0 + w + 3 < ncr) bv.w = brow[3];
                }
            }
            *reinterpret_cast<uint4 *>(&sP[kk][w]) = pv;
            *reinterpret_cast<uint4 *>(&sB[kk][w]) = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kVerK; ++kk) {
            const uint4 pv = *reinterpret_cast<const uint4 *>(&sP[kk][ty * 4]);
            const uint4 bv = *reinterpret_cast<const uint4 *>(&sB[kk][tx * 4]);
            const uint32_t pa[4] = {pv.x, pv.y, pv.z, pv.w}, ba[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) mac<true>(acc[a][b], pa[a], ba[b]);
        }
        __syncthreads();
    }
    uint32_t bad = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const uint32_t c = c0 + tx * 4 + b;
        if (c >= ncr) continue;
        const uint4 dv = *reinterpret_cast<const uint4 *>(D + (size_t)c * ld + i0 + ty * 4);
        bad |= fold_acc<true>(acc[0][b], 0u, p, pinv, two64) != dv.x;
        bad |= fold_acc<true>(acc[1][b], 0u, p, pinv, two64) != dv.y;
        bad |= fold_acc<true>(acc[2][b], 0u, p, pinv, two64) != dv.z;
        bad |= fold_acc<true>(acc[3][b], 0u, p, pinv, two64) != dv.w;
    }
    if (bad) atomicOr(mismatch, 1u);
}

/* Reduced echelon rows, dense: row t = pivot rank-1-t (decreasing lead). */
__global__ void k_ge_extract(const uint32_t *__restrict__ D, uint64_t ld, uint32_t ncr, uint32_t rank,
                             const uint32_t *__restrict__ pivcol, const uint32_t *__restrict__ pivrow,
                             const uint32_t *__restrict__ pivinv, const uint8_t *__restrict__ ispiv,
                             uint32_t *__restrict__ out, uint32_t p, uint64_t pinv,
                             const GEState *__restrict__ expect_state)
{
    const uint32_t t = blockIdx.x;                     /* rank can exceed the 65535 limit of grid.y */
    const uint32_t c = blockIdx.y * blockDim.x + threadIdx.x;
    if (t >= rank || c >= ncr) return;
    /* replay without a host round trip: `rank` is the EXPECTED rank; if the elimination found another one the pivot
     * arrays do not describe `rank` pivots -- write nothing (k_replay_compact reports the mismatch) */
    if (expect_state && expect_state->rank != rank) return;
    const uint32_t k = rank - 1 - t;
    const uint32_t ck = pivcol[k];
    uint32_t v;
    if (c < ck) v = 0;
    else if (c == ck) v = 1;
    else if (ispiv[c]) v = 0;
    else v = mulmod(D[(size_t)c * ld + pivrow[k]], pivinv[k], p, pinv);
    out[(size_t)t * ncr + c] = v;
}

/* Reducer bit arrays (trace learning): bit (i % 32) of rba[row][i / 32] is set
 * iff pivot i was applied to the row, i.e. its multiplier V[i][row] != 0.
 * grid = (words, chunks * Rc / 128), block = 128 rows. */
__global__ void k_rba_bits(const uint32_t *__restrict__ V, uint32_t nc, uint32_t Rc, uint32_t ncl,
                           uint32_t chunk0, uint32_t nrl, uint32_t words, uint32_t *__restrict__ rba)
{
    const uint32_t w = blockIdx.x;
    const uint32_t per = Rc / 128;
    const uint32_t lc = blockIdx.y / per, lr = (blockIdx.y % per) * 128 + threadIdx.x;
    const uint32_t row = (chunk0 + lc) * Rc + lr;
    if (row >= nrl) return;
    const uint32_t *Vc = V + (size_t)lc * nc * Rc;
    uint32_t bits = 0;
    for (uint32_t b = 0; b < 32; ++b) {
        const uint32_t col = 32 * w + b;
        if (col < ncl && Vc[(size_t)col * Rc + lr] != 0) bits |= 1u << b;
    }
    rba[(size_t)row * words + w] = bits;
}

/* Pivot-by-lead mode: lower row i is represented by the unit vector of its (renumbered) lead
 * column; the solve then yields row i of A^-1 (multipliers) and of -A^-1 B (trailing block). */
__global__ void k_scatter_units(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ cols,
                                uint32_t nru, uint32_t nrl, uint32_t nc, uint32_t chunk0, uint32_t nchunks, uint32_t Rc,
                                const uint32_t *__restrict__ colmap, uint32_t *__restrict__ flags,
                                uint32_t *__restrict__ V, uint32_t *__restrict__ err)
{
    const uint32_t r0 = chunk0 * Rc;
    const uint32_t r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrl || r >= (chunk0 + nchunks) * Rc) return;
    const uint64_t a = row_ptr[nru + r];
    if (row_ptr[nru + r + 1] == a) return;                   /* empty row: stays zero */
    const uint32_t lead = cols[a];
    if (lead >= nc) { atomicOr(err, kErrColumnRange); return; }
    const uint32_t j = colmap[lead];
    const uint32_t lc = (r - r0) / Rc, lr = (r - r0) % Rc;
    V[((size_t)lc * nc + j) * Rc + lr] = 1u;
    if (flags[(size_t)lc * nc + j] == 0xfffffffeu)           /* kFlagAlwaysZ -> kFlagAlwaysNZ (sourceless column) */
        flags[(size_t)lc * nc + j] = 0xffffffffu;
}

/* Lead (first stored) column of every lower row; 0xffffffff for an empty row. */
__global__ void k_lower_leads(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ cols,
                              uint32_t nru, uint32_t nrl, uint32_t *__restrict__ leads)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrl) return;
    const uint64_t a = row_ptr[nru + i];
    leads[i] = row_ptr[nru + i + 1] > a ? cols[a] : 0xffffffffu;
}

/* Normal-form mode: row r of D' as is (no elimination); pivot-by-lead mode: negated (negate_mod = p). */
__global__ void k_transpose_rows(const uint32_t *__restrict__ D, uint64_t ld, uint32_t ncr, uint32_t nrows,
                                 uint32_t negate_mod, uint32_t *__restrict__ out)
{
    __shared__ uint32_t tile[32][33];
    const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x, ty = threadIdx.y;
    for (uint32_t k = ty; k < 32; k += blockDim.y) {
        const uint32_t c = c0 + k, r = r0 + tx;
        tile[k][tx] = (c < ncr && r < nrows) ? D[(size_t)c * ld + r] : 0;
    }
    __syncthreads();
    for (uint32_t k = ty; k < 32; k += blockDim.y) {
        const uint32_t r = r0 + k, c = c0 + tx;
        if (r < nrows && c < ncr) {
            const uint32_t v = tile[tx][k];
            out[(size_t)r * ncr + c] = (negate_mod && v) ? negate_mod - v : v;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* trailing block, one persistent launch: panel factorisation + rank-b update behind grid barriers  */
/* ------------------------------------------------------------------------------------------------ */
/* The batched Gauss-Jordan above costs two launches per 8 pivots plus a host round trip every four batches
 * (K12: ~270 batches x (40 us single-CTA discovery + 32 us update + gaps) = 29 of the 127 ms of a step).
 * k_ge_fused runs the whole elimination in ONE cooperative launch (one CTA per SM, all co-resident):
 *   factor  (CTA 0)   gathers the next <= 32 flagged columns into shared memory AT ONCE and eliminates inside
 *                     the panel: a warp per column keeps "first non-zero row that is not a pivot row yet",
 *                     the leftmost column with a candidate yields the next pivot, one pass of all warps
 *                     applies it to the panel columns to its right and refreshes their candidates -- about a
 *                     microsecond per pivot instead of one visited column at a time;
 *   update  (all CTAs) the rank-b update of every column right of the panel's first pivot, the batch's pivot
 *                     columns in shared memory, multipliers by the lane-distributed recurrence (as in
 *                     k_ge_batch_update_cached), flags refreshed;
 * separated by grid barriers (a monotone arrival counter; data written by other SMs is read with ld.cg, L1
 * is not coherent).  The host reads the rank back once.  A watchdog on the barrier makes a stuck grid give up
 * with an error bit instead of hanging the GPU. */

constexpr uint32_t kPanelMax = 32;
constexpr int kFusedThreads = 1024;
constexpr uint32_t kErrGeBarrier = 128u;

struct GeFusedArgs {
    uint32_t *D; uint64_t ld; uint32_t nrows, nrows4, ncr, pw;     /* pw = panel width (<= kPanelMax, fits shared memory) */
    uint8_t *colflag, *used, *ispiv; GEState *st;
    uint32_t *pivcol, *pivrow, *pivinv;
    uint32_t p, two64; uint64_t pinv;
    uint32_t *bar;            /* [0] arrival counter, [1] abort flag; zeroed before the launch */
    uint32_t *err; uint64_t watchdog_ns; uint32_t max_iter; uint32_t debug;
};

/* true = abort (watchdog expired somewhere) */
__device__ __forceinline__ bool ge_grid_barrier(const GeFusedArgs &A, uint32_t &round)
{
    __shared__ uint32_t s_abort;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t target = (++round) * gridDim.x;
        atomicAdd(A.bar, 1u);
        uint32_t spins = 0, ab = 0;
        uint64_t t0 = 0;
        while (ld_relaxed_flag(A.bar) < target) {
            if (ld_relaxed_flag(A.bar + 1)) { ab = 1; break; }
            __nanosleep(32);
            if ((++spins & 1023u) == 0) {
                const uint64_t now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > A.watchdog_ns) { atomicOr(A.err, kErrGeBarrier); atomicExch(A.bar + 1, 1u); ab = 1; break; }
            }
        }
        __threadfence();
        s_abort = ab;
    } else {
        ++round;
    }
    __syncthreads();
    return s_abort != 0;
}

/* CTA 0: the next panel.  dynamic smem: [pw][nrows4] panel columns | [nrows4] used bytes */
__device__ void ge_panel_factor(const GeFusedArgs &A, uint32_t *s_panel, uint8_t *s_used)
{
    __shared__ uint32_t s_cols[kPanelMax], s_cand[kPanelMax], s_pj[kPanelMax], s_prow[kPanelMax], s_pinv[kPanelMax];
    __shared__ uint32_t s_wcnt[kFusedThreads / 32], s_m, s_next, s_scale[kPanelMax], s_isc[kPanelMax];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nrows = A.nrows, nrows4 = A.nrows4, ncr = A.ncr, pw = A.pw, p = A.p;
    GEState *st = A.st;
    if (__ldcg(&st->done)) {
        if (tid == 0) st->cur_col = __ldcg(&st->rank);          /* nothing pending: the update is a no-op */
        return;
    }
    const uint32_t k0 = __ldcg(&st->rank);
    const uint32_t c0 = __ldcg(&st->next_col);
    long long tk0 = clock64(), tk1, tsel = 0, telim = 0;
    for (uint32_t r = tid; r < nrows4; r += kFusedThreads) s_used[r] = r < nrows ? __ldcg(&A.used[r]) : (uint8_t)1;
    if (tid == 0) { s_m = 0; s_next = ncr; }
    __syncthreads();
    /* gather the next <= pw flagged columns, in order */
    for (uint32_t base = c0; base < ncr; base += kFusedThreads) {
        const uint32_t k = base + tid;
        const bool f = k < ncr && __ldcg(&A.colflag[k]) != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t off = s_m;
        for (uint32_t w = 0; w < warp; ++w) off += s_wcnt[w];
        const uint32_t idx = off + __popc(bal & ((1u << lane) - 1));
        if (f && idx < pw) s_cols[idx] = k;
        if (f && idx == pw - 1) s_next = k + 1;                   /* the panel is full after this column */
        __syncthreads();
        if (tid == 0) { uint32_t t = s_m; for (uint32_t w = 0; w < kFusedThreads / 32; ++w) t += s_wcnt[w]; s_m = t; }
        __syncthreads();
        if (s_m >= pw) break;
    }
    const uint32_t m = min(s_m, pw);
    const uint32_t next_col = s_m >= pw ? s_next : ncr;
    const bool exhausted = next_col >= ncr;
    if (m == 0) {
        if (tid == 0) { st->cur_col = k0; st->rank = k0; st->next_col = ncr; st->done = 1; }
        return;
    }
    tk1 = clock64();
    if (tid == 0) st->t[0] += tk1 - tk0;
    tk0 = tk1;
    /* load the panel; candidates = first non-zero row that is not a pivot row yet */
    for (uint32_t j = warp; j < m; j += kFusedThreads / 32) {
        const uint32_t *g = A.D + (size_t)s_cols[j] * A.ld;
        uint32_t cand = 0xffffffffu;
        for (uint32_t r = lane; r < nrows4; r += 32) {
            const uint32_t v = r < nrows ? __ldcg(&g[r]) : 0u;
            s_panel[(size_t)j * nrows4 + r] = v;
            if (v && !s_used[r]) cand = min(cand, r);
        }
        for (int o = 16; o; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0) { s_cand[j] = cand; s_scale[j] = 1u; }
    }
    __syncthreads();
    tk1 = clock64();
    if (tid == 0) st->t[1] += tk1 - tk0;
    uint32_t np = 0, cur = 0;
    /* Elimination inside the panel is FRACTION FREE: column j <- piv * column j - e_j * pivot column, i.e. the usual
     * step times the constant piv.  Scaling a panel column by a non-zero constant is harmless: a column that never
     * becomes a pivot column is discarded (the update pass recomputes it from memory), and a pivot column is only
     * ever used relative to its own pivot entry (multiplier = entry x inverse of the pivot entry) -- except by the
     * final extraction, which normalises ROWS with the inverse of the pivot entry; so the accumulated scale of every
     * column is tracked and divided out when a pivot column is written back.  No modular inverse sits on the
     * per-pivot critical path; the inverses are computed once per panel, in parallel. */
    /* One warp per panel column (at most 32 columns, 32 warps), and every warp repeats the tiny pivot selection for
     * itself: ONE block barrier per pivot.  (A former version gave each column to a group of 256 threads, four columns
     * at a time behind three more barriers: 4.7k cycles per pivot on a 127-row block, now a few hundred.) */
    for (;;) {
        tk0 = clock64();
        /* leftmost column at or after `cur` with a candidate */
        uint32_t js;
        {
            const uint32_t j = cur + lane;
            const bool has = j < m && s_cand[j] != 0xffffffffu;
            const uint32_t bal = __ballot_sync(0xffffffffu, has);
            js = bal ? cur + (uint32_t)__ffs(bal) - 1 : 0xffffffffu;
        }
        if (js == 0xffffffffu) break;
        const uint32_t rs = s_cand[js];
        const uint32_t *pc = s_panel + (size_t)js * nrows4;
        const uint32_t piv = pc[rs];
        if (tid == kFusedThreads - 1) { s_pj[np] = js; s_prow[np] = rs; s_used[rs] = 1; }   /* warp 31 never owns a column */
        tk1 = clock64(); tsel += tk1 - tk0; tk0 = tk1;
        /* column j <- piv * column j - e_j * pivot column (entry of the pivot row: piv * e_j); new candidates */
        const uint32_t j = js + 1 + warp;
        if (j < m) {
            uint32_t *col = s_panel + (size_t)j * nrows4;
            const uint32_t e = col[rs];
            const uint32_t neg = e ? p - e : 0u;
            const uint32_t negw = shoup_pre_fast(neg, p, A.pinv), pivw = shoup_pre_fast(piv, p, A.pinv);
            uint32_t cand = 0xffffffffu;
#pragma unroll 4
            for (uint32_t r = lane; r < nrows4; r += 32) {
                const uint32_t f = r == rs ? 0u : pc[r];
                uint32_t v = shoup_mul(col[r], piv, pivw, p) + shoup_mul(f, neg, negw, p);
                v = v >= p ? v - p : v;
                col[r] = v;
                if (v && r != rs && !s_used[r]) cand = min(cand, r);
            }
            for (int o = 16; o; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
            if (lane == 0) { s_cand[j] = cand; s_scale[j] = mulmod(s_scale[j], piv, p, A.pinv); }
        }
        __syncthreads();
        tk1 = clock64(); telim += tk1 - tk0;
        ++np;
        cur = js + 1;
    }
    __syncthreads();
    /* un-scaling factors and the inverses of the (true) pivot entries, all at once */
    if (tid < np) {
        const uint32_t isc = modinv(s_scale[s_pj[tid]], p);
        s_isc[tid] = isc;
        s_pinv[tid] = modinv(mulmod(s_panel[(size_t)s_pj[tid] * nrows4 + s_prow[tid]], isc, p, A.pinv), p);
    }
    __syncthreads();
    tk0 = clock64();
    /* pivot columns are final (they carry the multipliers of their pivot): write them back */
    for (uint32_t k = warp; k < np; k += kFusedThreads / 32) {
        const uint32_t j = s_pj[k];
        uint32_t *g = A.D + (size_t)s_cols[j] * A.ld;
        const uint32_t *col = s_panel + (size_t)j * nrows4;
        const uint32_t isc = s_isc[k];
        for (uint32_t r = lane; r < nrows; r += 32) g[r] = isc == 1u ? col[r] : mulmod(col[r], isc, p, A.pinv);
        if (lane == 0) {
            A.pivcol[k0 + k] = s_cols[j]; A.pivrow[k0 + k] = s_prow[k]; A.pivinv[k0 + k] = s_pinv[k];
            A.used[s_prow[k]] = 1; A.ispiv[s_cols[j]] = 1;
        }
    }
    if (tid == 0) {
        st->cur_col = k0; st->rank = k0 + np; st->next_col = next_col;
        if (exhausted) st->done = 1;
        st->visits += m; st->batches += np != 0;
        st->t[2] += tsel; st->t[3] += telim; st->t[4] += clock64() - tk0;
    }
}

/* all CTAs: rank-b update with the pivots [k0, k1), the body of k_ge_batch_update_cached with L2 loads of
 * everything other CTAs may have written.  dynamic smem: [nb][nrows4] pivot columns | [nrows4] used bytes */
__device__ void ge_fused_update(const GeFusedArgs &A, uint32_t k0, uint32_t k1, uint32_t *s_piv, uint8_t *s_used)
{
    __shared__ uint32_t s_pc[kPanelMax], s_pr[kPanelMax], s_pi[kPanelMax];
    __shared__ uint32_t s_x[kPanelMax][kPanelMax + 1];
    __shared__ uint32_t s_neg[kFusedThreads / 32][kPanelMax];
    const uint32_t nb = k1 - k0, nrows = A.nrows, nrows4 = A.nrows4, ncr = A.ncr, p = A.p;
    const uint64_t pinv = A.pinv;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t first = __ldcg(&A.pivcol[k0]);
    const uint32_t nwarps = gridDim.x * (kFusedThreads / 32);
    const uint32_t gwarp = blockIdx.x * (kFusedThreads / 32) + (tid >> 5);
    if (first + 1 + blockIdx.x * (kFusedThreads / 32) >= ncr) return;      /* no column for this CTA (uniform) */
    if (tid < nb) { s_pc[tid] = __ldcg(&A.pivcol[k0 + tid]); s_pr[tid] = __ldcg(&A.pivrow[k0 + tid]); s_pi[tid] = __ldcg(&A.pivinv[k0 + tid]); }
    __syncthreads();
    for (uint32_t k = 0; k < nb; ++k) {
        const uint32_t *src = A.D + (size_t)s_pc[k] * A.ld;
        const uint32_t own = s_pr[k];
        for (uint32_t r = tid; r < nrows4; r += kFusedThreads)
            s_piv[(size_t)k * nrows4 + r] = (r < nrows && r != own) ? __ldcg(&src[r]) : 0u;
    }
    for (uint32_t r = tid; r < nrows4; r += kFusedThreads) s_used[r] = r < nrows ? __ldcg(&A.used[r]) : (uint8_t)1;
    __syncthreads();
    for (uint32_t q = tid; q < kPanelMax * kPanelMax; q += kFusedThreads) {
        const uint32_t k = q / kPanelMax, j = q % kPanelMax;
        s_x[k][j] = (k < nb && j < nb) ? s_piv[(size_t)k * nrows4 + s_pr[j]] : 0u;
    }
    __syncthreads();
    for (uint32_t c = first + 1 + gwarp; c < ncr; c += nwarps) {
        if (__ldcg(&A.ispiv[c])) continue;
        uint32_t *g = A.D + (size_t)c * A.ld;
        uint32_t e = lane < nb ? __ldcg(&g[s_pr[lane]]) : 0u, neg = 0;
        bool touched = false;
        for (uint32_t k = 0; k < nb; ++k) {
            const uint32_t ek = __shfl_sync(0xffffffffu, e, k);
            if (s_pc[k] >= c || ek == 0) continue;
            const uint32_t ng = p - mulmod(ek, s_pi[k], p, pinv);
            touched = true;
            if (lane == k) neg = ng;
            if (lane > k && lane < nb) e = mod_u64((uint64_t)e + (uint64_t)ng * s_x[k][lane], p, pinv);
        }
        uint32_t *my_neg = s_neg[tid >> 5];
        my_neg[lane] = neg;
        __syncwarp();
        int any = 0;
        for (uint32_t r4 = lane * 4; r4 < nrows4; r4 += 128) {
            uint4 v = __ldcg(reinterpret_cast<const uint4 *>(g + r4));
            if (touched) {
                Acc a[4] = {{(uint64_t)v.x, 0u}, {(uint64_t)v.y, 0u}, {(uint64_t)v.z, 0u}, {(uint64_t)v.w, 0u}};
                for (uint32_t k = 0; k < nb; ++k) {
                    const uint32_t nk = my_neg[k];
                    if (nk) {
                        const uint4 f = *reinterpret_cast<const uint4 *>(s_piv + (size_t)k * nrows4 + r4);
                        mac<true>(a[0], f.x, nk); mac<true>(a[1], f.y, nk);
                        mac<true>(a[2], f.z, nk); mac<true>(a[3], f.w, nk);
                    }
                }
                v.x = fold_acc<true>(a[0], 0u, p, pinv, A.two64); v.y = fold_acc<true>(a[1], 0u, p, pinv, A.two64);
                v.z = fold_acc<true>(a[2], 0u, p, pinv, A.two64); v.w = fold_acc<true>(a[3], 0u, p, pinv, A.two64);
                *reinterpret_cast<uint4 *>(g + r4) = v;
            }
            const uchar4 u = *reinterpret_cast<const uchar4 *>(s_used + r4);
            any |= (v.x && !u.x) | (v.y && !u.y) | (v.z && !u.z) | (v.w && !u.w);
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) A.colflag[c] = (uint8_t)(any != 0);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kFusedThreads, 1)
k_ge_fused(const GeFusedArgs A)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint8_t *s_used = reinterpret_cast<uint8_t *>(s_dyn + (size_t)A.pw * A.nrows4);
    uint32_t round = 0;
    const bool t0 = blockIdx.x == 0 && threadIdx.x == 0;
    const long long t_begin = clock64();
    long long upd_cycles = 0, wait1_cycles = 0;
    for (uint32_t it = 0; it < A.max_iter; ++it) {
        if (blockIdx.x == 0) ge_panel_factor(A, s_dyn, s_used);
        long long ta = clock64();
        if (ge_grid_barrier(A, round)) return;
        long long tb = clock64();
        const uint32_t k0 = __ldcg(&A.st->cur_col), k1 = __ldcg(&A.st->rank), done = __ldcg(&A.st->done);
        if (k0 < k1) ge_fused_update(A, k0, k1, s_dyn, s_used);
        else if (done) {
            if (t0) A.st->t[7] += clock64() - t_begin;
            if (A.debug && threadIdx.x == 0) { atomicMax(&A.st->t[8], (unsigned long long)upd_cycles); if (blockIdx.x) atomicMax(&A.st->t[9], (unsigned long long)wait1_cycles); }
            return;
        }
        long long tc = clock64();
        upd_cycles += tc - tb; wait1_cycles += tb - ta;
        if (ge_grid_barrier(A, round)) return;
        if (t0) { A.st->t[5] += tc - tb; A.st->t[6] += (tb - ta) + (clock64() - tc); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(A.err, kErrGeBarrier);       /* did not terminate */
}

/* ====================================================================================================== */
/* Device-resident replay of a recorded call (f4la_b200_replay_call): no host round trip                    */
/* ====================================================================================================== */
constexpr uint32_t kReplayLead = 0xffffffffu;
constexpr uint32_t kReplayBadRank = 1u, kReplayBadSupport = 2u, kReplayDeviceError = 4u;

/* coefficient pool of the call <- coefficient arrays of the basis elements (arena); one warp per slot */
__global__ void k_replay_gather(const uint32_t *__restrict__ arena, uint32_t *__restrict__ pool,
                                const uint64_t *__restrict__ src, const uint64_t *__restrict__ dst,
                                const uint32_t *__restrict__ len, uint32_t nslots)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t sl = warp; sl < nslots; sl += nwarps) {
        const uint32_t *a = arena + src[sl];
        uint32_t *b = pool + dst[sl];
        const uint32_t n = len[sl];
        for (uint32_t i = lane; i < n; i += 32) b[i] = a[i];
    }
}

/* Dense result rows -> coefficient arrays of the new basis elements, at the RECORDED support: row i has entries at
 * the dense columns idx[row_ptr[i] .. row_ptr[i+1]) (kReplayLead = the lead term of a pivot-by-lead row, value 1).
 * The support is the recorded one iff the row has exactly that many non-zeros and all of them sit at recorded
 * positions; a different rank, another support or a device error is reported in *status (the prime then takes the
 * ordinary path on the host, like a bad prime of the reference, f4.c:433-441).  One warp per row. */
__global__ void k_replay_compact(const uint32_t *__restrict__ dense, uint32_t ncr, uint32_t nrows,
                                 const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ idx,
                                 const uint64_t *__restrict__ row_dst, uint32_t *__restrict__ arena,
                                 const GEState *__restrict__ st, const uint32_t *__restrict__ d_err,
                                 uint32_t *__restrict__ status)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool bad_rank = st && (st->rank != nrows || !st->done);
    if (warp == 0 && lane == 0) {
        if (bad_rank) atomicOr(status, kReplayBadRank);
        if (*d_err) atomicOr(status, kReplayDeviceError | (*d_err << 8));
    }
    if (bad_rank) return;
    for (uint32_t i = warp; i < nrows; i += nwarps) {
        const uint32_t *row = dense + (size_t)i * ncr;
        const uint64_t a = row_ptr[i], b = row_ptr[i + 1];
        uint32_t cnt = 0;
        for (uint32_t c = lane; c < ncr; c += 32) cnt += row[c] != 0;
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        uint32_t leads = 0, bad = 0;
        uint32_t *out = arena + row_dst[i];
        for (uint64_t t = a + lane; t < b; t += 32) {
            const uint32_t c = idx[t];
            uint32_t v = 1;
            if (c == kReplayLead) ++leads; else v = row[c];
            bad |= v == 0;
            out[t - a] = v;
        }
        for (int o = 16; o; o >>= 1) { leads += __shfl_xor_sync(0xffffffffu, leads, o); bad |= __shfl_xor_sync(0xffffffffu, bad, o); }
        if (lane == 0 && (bad || (uint64_t)cnt + leads != b - a)) atomicOr(status, kReplayBadSupport);
    }
}

} /* namespace f4la */
