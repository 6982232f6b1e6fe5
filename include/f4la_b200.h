/* f4la_b200 -- B200-native GF(p) linear algebra for F4 (C ABI).
 *
 * Two levels, both plain C (pointers + sizes, no torch / C++ types):
 *
 *  1. "flat" level: a Macaulay matrix [A B; C D] given as CSR-like host (or
 *     device-resident) arrays, reduced to the unique reduced row echelon form
 *     of its lower rows modulo the known pivots.  This is the computational
 *     content of neogb's  exact_sparse_reduced_echelon_form_ff_{8,16,32}
 *     (reference src/neogb/la_ff_32.c:2595, la_ff_16.c:1046, la_ff_8.c:1319).
 *
 *  2. "neogb" level (include/f4la_b200_neogb.h): drop-in replacements for the
 *     reference's global linear-algebra function pointers
 *     (reference src/neogb/data.h:529-595), operating on mat_t/bs_t/md_t.
 *
 * There is no CPU fallback: every entry point that computes fails loudly
 * (message on stderr + non-zero return / exit) when no sm_100 device is usable.
 */
#ifndef F4LA_B200_H
#define F4LA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F4LA_OK            0
#define F4LA_ERR_CUDA      1   /* CUDA runtime error (message on stderr) */
#define F4LA_ERR_NODEVICE  2   /* no usable GPU */
#define F4LA_ERR_ARG       3   /* malformed input */
#define F4LA_ERR_BADPRIME  4   /* apply mode: a row reduced to zero */

/* A Macaulay matrix in flat form.
 *
 * Rows 0 .. nru-1 are the known pivot rows ("rr" of mat_t, reference
 * data.h:203-232): row i has its lead term in column i, lead coefficient 1
 * and the lead term stored FIRST; the remaining entries may come in any
 * order (reference convert.c:405-413 does not sort them).
 * Rows nru .. nru+nrl-1 are the rows to be reduced ("tr").
 * Columns 0 .. ncl-1 are the known pivot columns (ncl == nru in F4 rounds),
 * columns ncl .. ncl+ncr-1 have no known pivot.
 *
 * Coefficients are shared between rows exactly as in the reference
 * (row = monomial x basis polynomial, so many rows alias one coefficient
 * array, reference hash.c:1369-1390): row r uses coefficient array row_cf[r],
 * whose length equals the row's length, entry k of the row has column
 * cols[row_ptr[r]+k] and coefficient cf[cf_ptr[row_cf[r]]+k].
 */
typedef struct f4la_flat_matrix {
    uint32_t fc;        /* field characteristic p, prime, p < 2^31            */
    uint32_t cf_bytes;  /* width of the coefficient pool entries: 1, 2 or 4   */
    uint32_t nru;       /* number of known pivot rows                         */
    uint32_t nrl;       /* number of rows to be reduced                       */
    uint32_t ncl;       /* number of known pivot columns                      */
    uint32_t ncr;       /* number of columns without known pivot              */
    uint32_t ncf;       /* number of distinct coefficient arrays              */
    uint32_t flags;     /* F4LA_FLAG_*                                        */
    const uint64_t *row_ptr; /* [nru+nrl+1] offsets into cols                 */
    const uint32_t *cols;    /* [row_ptr[nru+nrl]] column indices             */
    const uint32_t *row_cf;  /* [nru+nrl] coefficient array id of each row    */
    const uint64_t *cf_ptr;  /* [ncf+1] offsets (in elements) into cf         */
    const void     *cf;      /* coefficient pool, cf_bytes wide entries       */
} f4la_flat_matrix;

/* flags */
#define F4LA_FLAG_NONE        0u
/* keep every lower row as its own (non inter-reduced, non normalised) normal
 * form modulo the known pivots -- the nf>0 / final-reduction contract of the
 * reference (la_ff_32.c:2676-2684, 2764-2766).  Result row i corresponds to
 * input lower row i; rows reducing to zero have length 0. */
#define F4LA_FLAG_NORMALFORM  1u
/* known pivots are keyed by their lead column instead of by row index
 * (in_final_reduction_step, la_ff_32.c:2618-2622; interreduce_matrix_rows, :4254-4326).
 * Contract: every row (upper and lower) is monic with its lead first, all leads are distinct,
 * and the lower rows come in the reference's order (sort_matrix_rows_increasing, f4.c:603:
 * a lower row only contains leads of lower rows that come BEFORE it).  The reference then
 * reduces the lower rows one after the other by everything seen so far, which under that
 * order is the full inter-reduction: result row i = reduced form of lower row i, in place.
 * The device computes exactly that as one triangular solve of unit vectors (no elimination). */
#define F4LA_FLAG_PIVOT_BY_LEAD 2u
/* also return, for every lower row, the bit set of known pivots that were
 * applied to it (non-zero multipliers): the `rba` arrays the reference fills
 * while learning a trace (la_ff_32.c:1027-1036, data.h:209-213). */
#define F4LA_FLAG_WANT_RBA    4u
/* probabilistic echelon form (the reference's -l 42/43/44, la_ff_32.c:2305-2506): the lower
 * rows are replaced by a few hundred random linear combinations when that is enough to span
 * their space; same unique reduced rows with overwhelming probability (p > 2^15 only, else
 * the exact path runs).  Ignored together with NORMALFORM / PIVOT_BY_LEAD / WANT_RBA. */
#define F4LA_FLAG_PROBABILISTIC 8u
/* f4la_b200_reduce() only: the column indices are already on the device, sent piecewise with
 * f4la_b200_stage_cols() while the caller was still assembling them; `cols` is not read. */
#define F4LA_FLAG_COLS_STAGED 16u
/* likewise for the coefficient pool (f4la_b200_stage_commit with F4LA_STAGE_CF); `cf` is not read. */
#define F4LA_FLAG_CF_STAGED 32u

/* Result: rows of the reduced row echelon form, as CSR over ALL columns
 * (column ids in 0..ncl+ncr-1, ascending inside a row).  In echelon mode the
 * rows come by DECREASING lead column, lead coefficient 1, fully
 * inter-reduced (reference la_ff_32.c:2732-2762).  All arrays are malloc()ed
 * by the library and released with f4la_b200_free_result(). */
typedef struct f4la_flat_result {
    uint32_t  nrows;     /* number of result rows (new pivots)                */
    uint32_t  cf_bytes;  /* same as the input                                 */
    uint64_t *row_ptr;   /* [nrows+1]                                         */
    uint32_t *cols;      /* [row_ptr[nrows]]                                  */
    void     *cf;        /* [row_ptr[nrows]] coefficients, cf_bytes wide      */
    uint32_t *src_row;   /* [nrows] index of an input lower row whose
                            reduction produced this pivot (BINDEX/MULT donor); when the
                            pivot came out of a random combination of lower rows
                            (probabilistic mode, compress-and-verify: never while
                            F4LA_FLAG_WANT_RBA / tracing is on) one of the rows that
                            entered the combination                                  */
    uint32_t *rba;       /* F4LA_FLAG_WANT_RBA: [nrl][rba_words] bit i%32 of
                            word i/32 set iff pivot i was used; else NULL     */
    uint32_t  rba_words; /* ceil(nru / 32)                                    */
    uint32_t  reserved;
} f4la_flat_result;

/* Per-call measurements (device times from CUDA events on the ctx stream). */
typedef struct f4la_stats {
    double   ms_total;        /* whole call, host wall clock                  */
    double   ms_h2d;          /* host->device copies                          */
    double   ms_prep;         /* device-side layout kernels                   */
    double   ms_reduce;       /* sparse reduction by known pivots (hot kernel)*/
    double   ms_echelon;      /* trailing block echelon form                  */
    double   ms_d2h;          /* device->host copies + host materialisation   */
    uint64_t units;           /* coefficient applications of the REFERENCE algorithm for this reduction by the
                                 known pivots: sum over the pivots of (row length) x (lower rows with a non-zero
                                 multiplier) -- what la_ff_32.c:1214-1216 counts per application, 1 mul + 1 add
                                 each; the order-dependent new-pivot part of the reference's count is not in it */
    uint64_t pivot_applications; /* entries of the pull lists (non-lead non-zeros of the pivot rows) */
    uint64_t bytes_h2d;
    uint64_t bytes_d2h;
    uint32_t kernel_launches; /* kernels launched by this library in the call */
    uint32_t rank;            /* number of new pivots                         */
    double   ms_hot_kernel;   /* sum over the launches of k_pull_flow (CUDA
                                 events around each launch)                   */
    uint32_t hot_kernel_launches;
    uint32_t echelon_path;    /* trailing block: bit 0 result from compress-and-verify, bit 1 a verification
                                 failed (direct elimination used instead), bit 2 the number of combinations was
                                 escalated at least once                                                   */
    uint64_t dense_macs;      /* multiply-adds of the transposed formulation: (pull-list entries) x (lower rows),
                                 all-zero source vectors included (the hot kernel skips those)             */
    uint64_t ref_applications;/* (known pivot, lower row) pairs with a non-zero multiplier: the reference's
                                 application_nr_red for the reduction by the known pivots                 */
    double   ms_schur;        /* part of ms_hot_kernel spent in the tensor-core product for the columns without
                                 pivot (k_schur_a_split + k_schur_gemm + k_schur_combine); 0 when they are
                                 solved by the dataflow kernel                                              */
    uint64_t schur_macs;      /* part of dense_macs computed by that product                                */
} f4la_stats;

typedef struct f4la_ctx f4la_ctx;           /* one per host thread / stream   */
typedef struct f4la_resident f4la_resident; /* a matrix already in HBM        */

/* One recorded call of a device-resident replay (f4la_b200_replay_call): everything the call needs is on the device.
 * All pointers are DEVICE pointers (f4la_b200_device_alloc).  "arena" holds the coefficient arrays of the basis
 * elements of the prime being replayed; the call gathers its coefficient pool from it, runs the pipeline on the
 * resident structure WITHOUT any host round trip, and writes the coefficients of the result rows back into the arena
 * at the recorded support.  A rank or support other than the recorded one, or a device error, sets bits in *status
 * (1 rank, 2 support, 4 device error with the error bits << 8). */
typedef struct f4la_replay_call {
    uint32_t        expect_rows;   /* rows of the recorded result (new pivots; all lower rows for pivot-by-lead)   */
    uint32_t        nslots;        /* coefficient arrays of the matrix                                            */
    const uint64_t *gather_src;    /* [nslots] arena offset of the array (words)                                  */
    const uint64_t *gather_dst;    /* [nslots] offset in the coefficient pool of the resident matrix (= cf_ptr)    */
    const uint32_t *gather_len;    /* [nslots]                                                                    */
    const uint64_t *row_ptr;       /* [expect_rows + 1] recorded support                                          */
    const uint32_t *idx;           /* per recorded entry: column of the dense result row (global column - ncl_effective,
                                      through the column map for pivot-by-lead); 0xffffffff = lead term, value 1   */
    const uint64_t *row_dst;       /* [expect_rows] arena offset that receives row i's coefficients                */
    uint32_t       *arena;
    uint32_t       *status;
} f4la_replay_call;

/* Library / device management. */
const char *f4la_b200_version(void);
int  f4la_b200_device_count(void);
/* Creates a context bound to CUDA device `device` with its own stream.
 * Returns NULL (and prints why) when no device is usable. */
f4la_ctx *f4la_b200_create(int device);
void f4la_b200_destroy(f4la_ctx *ctx);
const char *f4la_b200_last_error(const f4la_ctx *ctx);

/* Seed of the random combinations of F4LA_FLAG_PROBABILISTIC (default: a fixed constant, so a
 * run is reproducible; the reference seeds rand() with --random-seed). */
void f4la_b200_set_seed(f4la_ctx *ctx, uint64_t seed);
/* How this context uses the host: `host_threads` bounds the OpenMP team of the result compaction (0 = OpenMP's default;
 * 1 when the caller already runs one context per host thread, as msolve's multi-modular loop does, msolve.c:1815-1824);
 * `blocking_sync` != 0 makes every wait for the device sleep on a blocking event instead of spinning in the driver, so
 * that more host threads than cores can each drive a context. */
void f4la_b200_set_host_mode(f4la_ctx *ctx, int host_threads, int blocking_sync);
/* Upper bound on the CTAs of this context's persistent dataflow kernel (0 = fill the device, the default).  When many
 * contexts share one GPU (one per host thread in the multi-modular loop) each of their small latency-bound launches
 * would otherwise occupy every SM and serialise the others; 64 measured best for 16-32 contexts. */
void f4la_b200_set_grid_cap(f4la_ctx *ctx, uint32_t ctas);

/* Page-locked host memory for the flat arrays handed to f4la_b200_reduce():
 * host->device copies from such buffers run at full PCIe rate and
 * asynchronously.  Plain malloc() memory works too, only slower. */
void *f4la_b200_pinned_alloc(size_t bytes);
void  f4la_b200_pinned_free(void *p);

/* Echelon form from HOST buffers: copies in, reduces, copies the result out.
 * Replaces the computational body of exact_sparse_reduced_echelon_form_ff_*
 * (reference la_ff_32.c:2595-2770). */
int f4la_b200_reduce(f4la_ctx *ctx, const f4la_flat_matrix *in,
                     f4la_flat_result *out, f4la_stats *stats);

/* Overlap of the caller's assembly of `cols` (the bulk of the input, 4 B per non-zero) with the
 * host->device copy: sends cols[first .. first+count) of a matrix with total_nnz non-zeros from a
 * page-locked buffer, asynchronously, on the context's stream.  Call it for consecutive pieces as
 * they become ready, then f4la_b200_reduce() with F4LA_FLAG_COLS_STAGED.  Used by the neogb
 * adapter, which gathers the rows of a mat_t (one malloc block each) piece by piece. */
int f4la_b200_stage_cols(f4la_ctx *ctx, const uint32_t *cols, uint64_t first, uint64_t count, uint64_t total_nnz);

/* Staging through a ring of page-locked pieces owned by the context (4 x 16 MB, pinned once): for callers whose
 * data lies in many small heap blocks (neogb: one malloc block per row / per polynomial) and is gathered piece by
 * piece.  f4la_b200_stage_acquire() hands out the next free piece (waiting for the copy that last used it);
 * the caller fills it with elements [first, first+count) of the column indices (which = F4LA_STAGE_COLS,
 * elem_bytes 4) or of the coefficient pool (F4LA_STAGE_CF, elem_bytes = cf_bytes) and commits: the piece is
 * copied asynchronously to its place on the device.  Then f4la_b200_reduce() with F4LA_FLAG_COLS_STAGED and / or
 * F4LA_FLAG_CF_STAGED.  No staging memory is ever re-pinned, whatever the size of the matrix. */
#define F4LA_STAGE_COLS 0
#define F4LA_STAGE_CF   1
int f4la_b200_stage_acquire(f4la_ctx *ctx, void **piece, size_t *capacity_bytes);
int f4la_b200_stage_commit(f4la_ctx *ctx, int which, uint32_t elem_bytes, uint64_t first, uint64_t count, uint64_t total);

/* Same computation with the input already resident in HBM (bench: `value`). */
f4la_resident *f4la_b200_upload(f4la_ctx *ctx, const f4la_flat_matrix *in);
int  f4la_b200_reduce_resident(f4la_ctx *ctx, f4la_resident *m,
                               f4la_flat_result *out, f4la_stats *stats);

/* Device-resident replay (used by the tape of f4la_b200_neogb.h; see f4la_replay_call above).
 * f4la_b200_replay_call needs a resident matrix that went through f4la_b200_resident_update and ONE ordinary
 * f4la_b200_reduce_resident (which fills its layout cache); it only enqueues work on the context's stream.
 * f4la_b200_device_copy: host<->device copy on that stream (to_device != 0: host -> device), waits for the stream when
 * `wait` != 0.  f4la_b200_resident_layout: the column map the pipeline uses for this matrix (identity unless
 * F4LA_FLAG_PIVOT_BY_LEAD) and the effective number of pivot columns. */
void *f4la_b200_device_alloc(f4la_ctx *ctx, size_t bytes);
void  f4la_b200_device_free(f4la_ctx *ctx, void *p);
int   f4la_b200_device_copy(f4la_ctx *ctx, void *dst, const void *src, size_t bytes, int to_device, int wait);
int   f4la_b200_device_zero(f4la_ctx *ctx, void *dst, size_t bytes);
int   f4la_b200_resident_layout(f4la_ctx *ctx, const f4la_resident *m, uint32_t *colmap, uint32_t *ncl_effective);
int   f4la_b200_replay_call(f4la_ctx *ctx, f4la_resident *m, uint32_t fc, const f4la_replay_call *rc);
void f4la_b200_release(f4la_ctx *ctx, f4la_resident *m);
/* A resident matrix with the same structure but other coefficients (and another prime): overwrites the coefficient
 * pool from host memory (same layout and size as at upload time) and sets the field characteristic.  This is how
 * the application phase of the multi-modular tracer reuses one structure for many primes (f4la_b200_neogb.h). */
int  f4la_b200_resident_update(f4la_ctx *ctx, f4la_resident *m, uint32_t fc, const void *cf);

void f4la_b200_free_result(f4la_flat_result *res);

/* Exact dense product mod p on the tensor cores (tcgen05.mma.kind::i8 on 8-bit limbs, TMEM accumulators; see
 * csrc/f4la_tc.cuh): C = A * B mod p with A (M x K) and C (M x N) column major with leading dimension M, B (K x N)
 * row major, all host arrays of residues < p < 2^31, K <= 8192.  This is the kernel that certifies the echelon form
 * of tall trailing blocks inside f4la_b200_reduce(); the entry point exists for tests and measurements.
 * *kernel_ms (optional) receives the duration of the kernel alone. */
int f4la_b200_modp_gemm(f4la_ctx *ctx, const uint32_t *A, const uint32_t *B, uint32_t M, uint32_t N, uint32_t K,
                        uint32_t p, uint32_t *C, double *kernel_ms);

/* Measurement aid (bench.py): sustained rate, in multiply-adds per second, of the 32 x 32 -> 96-bit
 * multiply-accumulate of the hot kernel with register operands, run for about `milliseconds` on the
 * context's device -- the issue ceiling of the multiplier pipe the kernel is bound by. */
double f4la_b200_measure_mac_peak(f4la_ctx *ctx, double milliseconds);

#ifdef __cplusplus
}
#endif
#endif /* F4LA_B200_H */
