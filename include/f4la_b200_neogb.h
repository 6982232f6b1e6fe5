/* f4la_b200_neogb.h -- drop-in replacements for neogb's linear-algebra
 * function pointers (reference src/neogb/data.h:529-595, defaults in
 * data.c:71-103, dispatchers in io.c:868-1044).
 *
 * The reference selects its GF(p) linear algebra through WRITABLE GLOBAL
 * FUNCTION POINTERS; assigning them once before core_f4()/core_gba() takes
 * over the path.  This header declares functions with
 * exactly those signatures.  mat_t / bs_t / md_t are NOT re-declared here:
 * the reference-side stub passes their field offsets (computed by its own
 * compiler with offsetof) in an f4la_neogb_layout, so the backend stays ABI
 * correct if the reference adds or reorders fields.  See INTEGRATION.md for
 * the stub a maintainer adds to the reference.
 */
#ifndef F4LA_B200_NEOGB_H
#define F4LA_B200_NEOGB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field offsets (bytes) of the reference structs, filled in with offsetof()
 * on the reference side.  Row prelude indices are the OFFSET/LENGTH/... macros
 * of data.h:49-59. */
typedef struct f4la_neogb_layout {
    uint32_t struct_size;       /* sizeof(f4la_neogb_layout), version check   */
    /* mat_t, data.h:203-232 */
    uint32_t mat_tr, mat_rba, mat_rr, mat_cf_8, mat_cf_16, mat_cf_32;
    uint32_t mat_sz, mat_np, mat_nr, mat_nc, mat_nru, mat_nrl, mat_ncl, mat_ncr, mat_rbal;
    /* bs_t, data.h:178-200 */
    uint32_t bs_ld, bs_cf_8, bs_cf_16, bs_cf_32;
    /* md_t, data.h:327-437 */
    uint32_t md_trace_level, md_np, md_la_ctime, md_la_rtime, md_num_zerored;
    uint32_t md_fc, md_laopt, md_nthrds, md_ff_bits, md_nf, md_in_final_reduction_step;
    uint32_t md_info_level, md_application_nr_mult, md_application_nr_add, md_application_nr_red;
    uint32_t md_tr;             /* trace_t *tr, data.h:330 */
    /* row prelude, data.h:49-59 */
    uint32_t row_offset, row_length, row_preloop, row_coeffs, row_mult, row_bindex, row_deg;
    uint32_t unroll;            /* UNROLL (data.h:47), PRELOOP = len % UNROLL */
} f4la_neogb_layout;

/* Must be called once before any of the entry points below; copies *layout. */
int f4la_b200_neogb_configure(const f4la_neogb_layout *layout);

/* Learning a trace (LEARN_TRACER, f4.c:360-369) needs the reference's own
 * construct_trace(trace_t *, mat_t *) (tools.c:63-183, static in the unity
 * build): the stub hands its address over.  Without it LEARN_TRACER aborts. */
void f4la_b200_neogb_set_construct_trace(void (*fn)(void *trace, void *mat));

/* CUDA device used by the calling host thread (default: $F4LA_B200_DEVICE or
 * 0).  The QQ multi-modular loop runs one F4 instance per OpenMP thread
 * (reference src/msolve/msolve.c:1815-1824); bind thread t to GPU t mod G. */
void f4la_b200_neogb_set_device(int device);

/* Replaces `linear_algebra` (data.h:545-550; default dispatch_linear_algebra,
 * io.c:878): st->laopt 1, 2, 42, 43, 44 all yield the same unique reduced row
 * echelon form, which is what this computes (exactly, every time). */
void f4la_b200_linear_algebra(void *mat, const void *tbr, const void *bs, void *st);

/* Replaces `exact_linear_algebra` (data.h:531-536; io.c:962). */
void f4la_b200_exact_linear_algebra(void *mat, const void *tbr, const void *bs, void *st);

/* Replaces `application_linear_algebra` (data.h:559-563; io.c dispatch_application_linear_algebra
 * -> exact_application_sparse_linear_algebra_ff_*, la_ff_32.c:4041/3135): same echelon form,
 * returns 1 when a row reduced to zero (unlucky prime), else 0. */
int f4la_b200_application_linear_algebra(void *mat, const void *bs, void *st);

/* Replaces `trace_linear_algebra` (data.h:571-576; exact_trace_sparse_linear_algebra_ff_*,
 * la_ff_32.c:4075/3004): echelon form + reducer bit arrays + construct_trace(trace, mat). */
void f4la_b200_trace_linear_algebra(void *trace, void *mat, const void *bs, void *st);

/* Replaces `interreduce_matrix_rows` (data.h:584-589; interreduce_matrix_rows_ff_*,
 * la_ff_32.c:4254-4326).  free_basis != 0 needs the reference's (static) free_basis_elements,
 * handed over with f4la_b200_neogb_set_free_basis_elements(). */
void f4la_b200_interreduce_matrix_rows(void *mat, void *bs, void *st, int free_basis);
void f4la_b200_neogb_set_free_basis_elements(void (*fn)(void *bs));

/* Accumulated device statistics of the calling thread since the last reset. */
typedef struct f4la_neogb_totals {
    uint64_t calls;
    double   ms_total, ms_flatten, ms_h2d, ms_prep, ms_reduce, ms_echelon, ms_d2h, ms_materialize;
    uint64_t units, bytes_h2d, bytes_d2h, kernel_launches;
} f4la_neogb_totals;
void f4la_b200_neogb_get_totals(f4la_neogb_totals *t, int reset);

/* ---------------------------------------------------------------------------------------------------------
 * Structure-cached replay of the tracer APPLICATION phase (the QQ multi-modular loop, src/msolve/msolve.c:1815-1824;
 * neogb f4.c:416-478 with trace_level == APPLY_TRACER).
 *
 * In the application phase every prime builds matrices of IDENTICAL structure (the trace fixes which multiples of
 * which basis elements form the rows, symbol.c:642-713); only the coefficients differ.  The reference nevertheless
 * rebuilds the structure for every prime on the host (hash-table insertions, column maps), which is what bounds the
 * loop once the linear algebra runs on a GPU.  A tape records, during ONE application run through the drop-in
 * pointers, for every linear-algebra call: the flat structure of the matrix, the basis element behind every
 * coefficient array, and the support of the result rows.  Replaying the tape for another prime needs no symbolic
 * work at all: coefficient pools are assembled from the basis elements computed so far for that prime, every call
 * runs on the structure already resident in HBM, new rows become the next basis elements exactly as
 * convert_sparse_matrix_rows_to_basis_elements(-1, ...) numbers them (convert.c:660-705: element bl + k is result
 * row np-1-k), and the rows of the final reduction step (f4.c:568-634) are the reduced basis.  Primes are
 * independent: one replay per host thread / context, sharded over the GPUs.
 *
 * A replay stops with F4LA_ERR_BADPRIME when a call yields another number of rows than recorded (the reference's
 * own bad-prime test, f4.c:433-441), and with F4LA_ERR_PATTERN when a result row has another support than on the
 * recorded prime (a coefficient vanishing mod p): such primes go through the ordinary drop-in path.
 * --------------------------------------------------------------------------------------------------------- */
#define F4LA_ERR_PATTERN 5
typedef struct f4la_tape f4la_tape;

/* Between begin and end every call of the calling thread through the entry points above is recorded.
 * end() returns NULL when nothing usable was recorded (no call, or a mode that cannot be replayed). */
void       f4la_b200_neogb_tape_begin(void);
f4la_tape *f4la_b200_neogb_tape_end(void);
void       f4la_b200_tape_free(f4la_tape *tape);

/* number of recorded calls; number of initial basis elements (indices 0 .. n-1, the input system); length of
 * initial element i as the tape uses it (0: never referenced); words of the reduced basis a replay returns */
uint32_t f4la_b200_tape_calls(const f4la_tape *tape);
uint32_t f4la_b200_tape_initial_elements(const f4la_tape *tape);
uint32_t f4la_b200_tape_initial_length(const f4la_tape *tape, uint32_t i);
uint64_t f4la_b200_tape_residue_words(const f4la_tape *tape);

/* Replays the tape for the prime p on the context ctx (struct f4la_ctx of f4la_b200.h).  init_cf / init_off: the
 * coefficient arrays of the initial basis elements modulo p, normalised as normalize_initial_basis does (lead
 * coefficient 1), element i at init_cf[init_off[i] .. init_off[i+1]).  residues receives the coefficients of the
 * reduced basis, element after element (f4la_b200_tape_residue_words() words).  la_ms (optional) += device +
 * host milliseconds spent.  Thread safe for distinct contexts. */
int f4la_b200_tape_replay(const f4la_tape *tape, void *ctx, uint32_t p, const uint32_t *init_cf,
                          const uint64_t *init_off, uint32_t *residues, double *la_ms);

/* The same replay WITHOUT host round trips: the coefficient arrays of all basis elements of the prime stay on the
 * device (an arena per context), every recorded call gathers its coefficient pool from there, runs the pipeline on the
 * resident structure with the recorded rank and support as expectation, and writes its result rows back into the
 * arena; one device->host copy of the residues at the end.  _enqueue only queues work on the context's stream (up to
 * 16 primes per context between two _collect calls) -- `residues` must stay valid until _collect, which waits for the
 * stream and returns, in enqueue order, F4LA_OK / F4LA_ERR_BADPRIME (another rank) / F4LA_ERR_PATTERN (another
 * support) / F4LA_ERR_CUDA per prime (return value: number of results, -1 on bad arguments).  The first replay on a
 * context, and every replay of a tape that cannot be expressed this way (blocks that need compress-and-verify,
 * 8/16-bit coefficients), run through f4la_b200_tape_replay inside _enqueue; F4LA_B200_TAPE_ASYNC=0 forces that.
 * Page-locked `residues` (f4la_b200_pinned_alloc) keep the final copy asynchronous; `residues` may also point into
 * DEVICE memory of the context's GPU (the batch then stays in HBM, e.g. for an NCCL gather). */
int f4la_b200_tape_replay_enqueue(const f4la_tape *tape, void *ctx, uint32_t p, const uint32_t *init_cf,
                                  const uint64_t *init_off, uint32_t *residues);
int f4la_b200_tape_replay_collect(const f4la_tape *tape, void *ctx, int *rcs, int max_rcs);

#ifdef __cplusplus
}
#endif
#endif
