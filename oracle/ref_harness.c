/* ref_harness.c -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * Compiles the UNMODIFIED reference neogb (unity build, reference
 * src/neogb/gb.c:27-51) together with a small driver into
 * oracle/_ref/libneogb_ref.so.  The sources are compiled from where they lie
 * under /root/reference; nothing of them is copied into this repository.
 *
 * What it offers to tests/ and to bench.py's cpu_baseline / reference arm:
 *   - refh_f4():          run the reference F4 (export_f4, f4.c:769) on a
 *                         polynomial system, with the linear algebra served
 *                         by (0) the reference itself, (1) an installed
 *                         backend, (2) both, compared step by step (shadow),
 *                         (3) the reference, recording every matrix.
 *   - refh_reference_la(): run the reference LA (dispatch_linear_algebra,
 *                         io.c:878) on one flat matrix.
 *   - refh_struct_layout(): offsets of the reference's mat_t/bs_t/md_t, so a
 *                         CPU test can check include/f4la_b200_neogb.h.
 *
 * Because this file #includes the reference's gb.c it sees its static
 * functions; the global LA function pointers (data.c:71-103) are the hooks.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#include "gb.c"                 /* the reference, unmodified (-I at build) */
#include "f4la_b200.h"          /* flat exchange format only               */
#include "f4la_b200_neogb.h"    /* layout descriptor of the drop-in ABI    */

/* ------------------------------------------------------------------ */
/* records                                                             */
/* ------------------------------------------------------------------ */

typedef struct refh_record {
    /* input */
    f4la_flat_matrix in;        /* owned copies                            */
    uint32_t *bindex;           /* [nru+nrl] row[BINDEX]                   */
    uint32_t *mult;             /* [nru+nrl] row[MULT]                     */
    /* reference output */
    f4la_flat_result out;
    uint32_t *out_bindex;
    uint32_t *out_mult;
    /* meta */
    int32_t  kind;              /* 0 linear_algebra, 1 exact_linear_algebra,
                                   2 interreduce_matrix_rows               */
    int32_t  laopt, ff_bits, nf, final_step, trace_level;
    double   la_seconds;        /* reference wall time of this call        */
    double   units;             /* coefficient applications (31-bit only)  */
    double   pivot_apps;
} refh_record;

static refh_record *g_rec   = NULL;
static int          g_nrec  = 0;
static int          g_caprec = 0;

static int    g_mode        = 0;   /* 0 ref, 1 backend, 2 shadow, 3 record (reference serves), 4 record (backend serves) */
static int    g_max_records = 1 << 30;
static int64_t g_shadow_mismatch = 0;
static int64_t g_shadow_calls    = 0;
static double g_la_seconds       = 0.0;   /* LA wall time, whichever served */
static int    g_la_calls         = 0;
static int    g_verbose          = 0;

typedef void (*la_fn_t)(mat_t *, const bs_t *const, const bs_t *const, md_t *);
typedef void (*ir_fn_t)(mat_t *, bs_t *, md_t *, int);
typedef int  (*app_fn_t)(mat_t *, const bs_t *const, md_t *);
typedef void (*tr_fn_t)(trace_t *, mat_t *, const bs_t *const, md_t *);
static la_fn_t  g_backend_la       = NULL;
static la_fn_t  g_backend_exact_la = NULL;
static ir_fn_t  g_backend_interred = NULL;
static app_fn_t g_backend_app      = NULL;
static tr_fn_t  g_backend_trace    = NULL;

static size_t cf_width(int ff_bits) { return ff_bits == 8 ? 1 : ff_bits == 16 ? 2 : 4; }

static const void *cf_array_of(const bs_t *b, int ff_bits, hm_t idx)
{
    switch (ff_bits) {
        case 8:  return b->cf_8[idx];
        case 16: return b->cf_16[idx];
        default: return b->cf_32[idx];
    }
}

/* Flatten the LA input (mat_t before the call) into owned arrays.
 * Coefficient arrays are deduplicated by (basis, COEFFS index). */
static void flatten_input(refh_record *rec, const mat_t *mat, const bs_t *tbr,
                          const bs_t *bs, const md_t *st, int rows_only_rr)
{
    const uint32_t nru = mat->nru, nrl = rows_only_rr ? 0 : mat->nrl;
    const uint32_t n   = nru + nrl;
    const size_t   w   = cf_width(st->ff_bits);
    f4la_flat_matrix *m = &rec->in;
    memset(m, 0, sizeof(*m));
    m->fc = st->fc; m->cf_bytes = (uint32_t)w;
    m->nru = nru; m->nrl = nrl; m->ncl = mat->ncl; m->ncr = mat->ncr;
    uint64_t *rp = malloc((n + 1) * sizeof(uint64_t));
    uint32_t *rcf = malloc((n ? n : 1) * sizeof(uint32_t));
    rec->bindex = malloc((n ? n : 1) * sizeof(uint32_t));
    rec->mult   = malloc((n ? n : 1) * sizeof(uint32_t));
    uint64_t nnz = 0;
    for (uint32_t r = 0; r < n; ++r) {
        const hm_t *row = r < nru ? mat->rr[r] : mat->tr[r - nru];
        rp[r] = nnz; nnz += row[LENGTH];
    }
    rp[n] = nnz;
    uint32_t *cols = malloc((nnz ? nnz : 1) * sizeof(uint32_t));
    /* dedupe coefficient arrays: map basis index -> pool id, separately for
     * bs (pivots) and tbr (rows to be reduced) unless they are the same */
    const uint32_t bs_ld  = bs->ld;
    const uint32_t tbr_ld = tbr->ld;
    int32_t *map_bs  = malloc((bs_ld + 1) * sizeof(int32_t));
    int32_t *map_tbr = (tbr == bs) ? map_bs : malloc((tbr_ld + 1) * sizeof(int32_t));
    for (uint32_t i = 0; i < bs_ld; ++i) map_bs[i] = -1;
    if (tbr != bs) for (uint32_t i = 0; i < tbr_ld; ++i) map_tbr[i] = -1;
    uint32_t ncf = 0; uint64_t ncoef = 0;
    uint64_t *cfp = malloc((n + 2) * sizeof(uint64_t));
    const void **cfsrc = malloc((n + 1) * sizeof(void *));
    for (uint32_t r = 0; r < n; ++r) {
        const hm_t *row = r < nru ? mat->rr[r] : mat->tr[r - nru];
        const bs_t *b   = r < nru ? bs : tbr;
        int32_t *map    = r < nru ? map_bs : map_tbr;
        memcpy(cols + rp[r], row + OFFSET, row[LENGTH] * sizeof(uint32_t));
        rec->bindex[r] = row[BINDEX];
        rec->mult[r]   = row[MULT];
        const hm_t ci = row[COEFFS];
        if (map[ci] < 0) {
            map[ci] = (int32_t)ncf;
            cfp[ncf] = ncoef;
            cfsrc[ncf] = cf_array_of(b, st->ff_bits, ci);
            ncoef += row[LENGTH];
            ncf++;
        }
        rcf[r] = (uint32_t)map[ci];
    }
    cfp[ncf] = ncoef;
    unsigned char *pool = malloc((ncoef ? ncoef : 1) * w);
    for (uint32_t c = 0; c < ncf; ++c)
        memcpy(pool + cfp[c] * w, cfsrc[c], (cfp[c + 1] - cfp[c]) * w);
    free(cfsrc); free(map_bs); if (tbr != bs) free(map_tbr);
    m->ncf = ncf; m->row_ptr = rp; m->cols = cols; m->row_cf = rcf;
    m->cf_ptr = cfp; m->cf = pool;
    if (st->nf > 0) m->flags |= F4LA_FLAG_NORMALFORM;
    if (st->in_final_reduction_step) m->flags |= F4LA_FLAG_PIVOT_BY_LEAD;
}

/* Flatten the LA output: mat->tr[0..np) (entries may be NULL in nf mode). */
static void flatten_output(f4la_flat_result *res, uint32_t **obi, uint32_t **omu,
                           const mat_t *mat, const md_t *st)
{
    const uint32_t np = mat->np;
    const size_t w = cf_width(st->ff_bits);
    memset(res, 0, sizeof(*res));
    res->nrows = np; res->cf_bytes = (uint32_t)w;
    res->row_ptr = malloc((np + 1) * sizeof(uint64_t));
    *obi = malloc((np ? np : 1) * sizeof(uint32_t));
    *omu = malloc((np ? np : 1) * sizeof(uint32_t));
    uint64_t nnz = 0;
    for (uint32_t i = 0; i < np; ++i) {
        res->row_ptr[i] = nnz;
        if (mat->tr[i]) nnz += mat->tr[i][LENGTH];
    }
    res->row_ptr[np] = nnz;
    res->cols = malloc((nnz ? nnz : 1) * sizeof(uint32_t));
    res->cf   = malloc((nnz ? nnz : 1) * w);
    res->src_row = NULL;
    for (uint32_t i = 0; i < np; ++i) {
        const hm_t *row = mat->tr[i];
        if (!row) { (*obi)[i] = 0; (*omu)[i] = 0; continue; }
        memcpy(res->cols + res->row_ptr[i], row + OFFSET, row[LENGTH] * sizeof(uint32_t));
        const void *cf = st->ff_bits == 8 ? (void *)mat->cf_8[row[COEFFS]]
                       : st->ff_bits == 16 ? (void *)mat->cf_16[row[COEFFS]]
                       : (void *)mat->cf_32[row[COEFFS]];
        memcpy((unsigned char *)res->cf + res->row_ptr[i] * w, cf, row[LENGTH] * w);
        (*obi)[i] = row[BINDEX]; (*omu)[i] = row[MULT];
    }
}

static refh_record *new_record(void)
{
    if (g_nrec == g_caprec) {
        g_caprec = g_caprec ? 2 * g_caprec : 64;
        g_rec = realloc(g_rec, (size_t)g_caprec * sizeof(refh_record));
    }
    refh_record *r = &g_rec[g_nrec++];
    memset(r, 0, sizeof(*r));
    return r;
}

static void free_record(refh_record *r)
{
    free((void *)r->in.row_ptr); free((void *)r->in.cols); free((void *)r->in.row_cf);
    free((void *)r->in.cf_ptr); free((void *)r->in.cf);
    free(r->bindex); free(r->mult);
    free(r->out.row_ptr); free(r->out.cols); free(r->out.cf); free(r->out.src_row);
    free(r->out_bindex); free(r->out_mult);
    memset(r, 0, sizeof(*r));
}

/* Deep copy of a mat_t's rows (for shadow mode: the LA consumes its input). */
static mat_t *clone_matrix(const mat_t *mat)
{
    mat_t *c = calloc(1, sizeof(mat_t));
    *c = *mat;
    c->rr = malloc((mat->nru ? mat->nru : 1) * sizeof(hm_t *));
    c->tr = malloc((mat->nrl ? mat->nrl : 1) * sizeof(hm_t *));
    for (len_t i = 0; i < mat->nru; ++i) {
        size_t sz = (mat->rr[i][LENGTH] + OFFSET) * sizeof(hm_t);
        c->rr[i] = malloc(sz); memcpy(c->rr[i], mat->rr[i], sz);
    }
    for (len_t i = 0; i < mat->nrl; ++i) {
        size_t sz = (mat->tr[i][LENGTH] + OFFSET) * sizeof(hm_t);
        c->tr[i] = malloc(sz); memcpy(c->tr[i], mat->tr[i], sz);
    }
    c->cf_8 = NULL; c->cf_16 = NULL; c->cf_32 = NULL; c->cf_qq = NULL; c->cf_ab_qq = NULL;
    c->rba = NULL; c->rbal = 0;
    if (mat->rba && mat->rbal) {
        /* rba[i] has ceil(nru/32) words (symbol.c:625-629) */
        const size_t wl = mat->nru / 32 + ((mat->nru % 32) != 0);
        c->rba = malloc(mat->rbal * sizeof(rba_t *));
        c->rbal = mat->rbal;
        for (len_t i = 0; i < mat->rbal; ++i) c->rba[i] = calloc(wl ? wl : 1, sizeof(rba_t));
    }
    return c;
}

static void free_clone_after_la(mat_t *c, const md_t *st)
{
    /* rows in c->tr[0..np) and their coefficient arrays are owned by us */
    for (len_t i = 0; i < c->np; ++i) {
        if (!c->tr || !c->tr[i]) continue;
        hm_t ci = c->tr[i][COEFFS];
        if (st->ff_bits == 8 && c->cf_8) free(c->cf_8[ci]);
        else if (st->ff_bits == 16 && c->cf_16) free(c->cf_16[ci]);
        else if (c->cf_32) free(c->cf_32[ci]);
        free(c->tr[i]);
    }
    for (len_t i = 0; i < c->rbal; ++i) free(c->rba[i]);
    free(c->rba);
    free(c->rr); free(c->tr); free(c->cf_8); free(c->cf_16); free(c->cf_32);
    free(c);
}

static int64_t compare_results(const f4la_flat_result *a, const f4la_flat_result *b)
{
    if (a->nrows != b->nrows) return 1 + (int64_t)(a->nrows > b->nrows ? a->nrows - b->nrows : b->nrows - a->nrows);
    int64_t bad = 0;
    for (uint32_t i = 0; i < a->nrows; ++i) {
        uint64_t la = a->row_ptr[i + 1] - a->row_ptr[i];
        uint64_t lb = b->row_ptr[i + 1] - b->row_ptr[i];
        if (la != lb) { bad++; continue; }
        if (memcmp(a->cols + a->row_ptr[i], b->cols + b->row_ptr[i], la * 4)) { bad++; continue; }
        if (memcmp((unsigned char *)a->cf + a->row_ptr[i] * a->cf_bytes,
                   (unsigned char *)b->cf + b->row_ptr[i] * b->cf_bytes, la * a->cf_bytes)) bad++;
    }
    return bad;
}

/* ------------------------------------------------------------------ */
/* the hooks                                                           */
/* ------------------------------------------------------------------ */

static void hook_common(int kind, mat_t *mat, const bs_t *tbr, const bs_t *bs, md_t *st)
{
    la_fn_t ref_fn = kind == 0 ? dispatch_linear_algebra : dispatch_exact_linear_algebra;
    la_fn_t be_fn  = kind == 0 ? g_backend_la : g_backend_exact_la;
    const double t0 = realtime();
    g_la_calls++;
    if ((g_mode == 1 || (g_mode == 4 && g_nrec >= g_max_records)) && be_fn) {
        be_fn(mat, tbr, bs, st);
        g_la_seconds += realtime() - t0;
        return;
    }
    if (g_mode == 2 && be_fn) {
        mat_t *c = clone_matrix(mat);
        md_t stc = *st;
        be_fn(c, tbr, bs, &stc);
        f4la_flat_result rb, rr; uint32_t *b1, *m1, *b2, *m2;
        flatten_output(&rb, &b1, &m1, c, &stc);
        const double t1 = realtime();
        ref_fn(mat, tbr, bs, st);
        g_la_seconds += realtime() - t1;
        flatten_output(&rr, &b2, &m2, mat, st);
        int64_t bad = compare_results(&rb, &rr);
        if (bad && g_verbose)
            fprintf(stderr, "[refh] shadow mismatch kind %d call %d: %ld (np backend %u, ref %u)\n",
                    kind, g_la_calls, (long)bad, rb.nrows, rr.nrows);
        g_shadow_mismatch += bad; g_shadow_calls++;
        free(rb.row_ptr); free(rb.cols); free(rb.cf); free(b1); free(m1);
        free(rr.row_ptr); free(rr.cols); free(rr.cf); free(b2); free(m2);
        free_clone_after_la(c, &stc);
        return;
    }
    if ((g_mode == 3 || (g_mode == 4 && be_fn)) && g_nrec < g_max_records) {
        if (g_mode == 4) ref_fn = be_fn;      /* record the inputs, serve with the backend */
        refh_record *rec = new_record();
        flatten_input(rec, mat, tbr, bs, st, 0);
        rec->kind = kind; rec->laopt = st->laopt; rec->ff_bits = st->ff_bits;
        rec->nf = st->nf; rec->final_step = st->in_final_reduction_step;
        rec->trace_level = st->trace_level;
        const double m0 = st->application_nr_mult, r0 = (double)st->application_nr_red;
        const double t1 = realtime();
        ref_fn(mat, tbr, bs, st);
        rec->la_seconds = realtime() - t1;
        g_la_seconds += rec->la_seconds;
        rec->units = (st->application_nr_mult - m0) * 1000.0;
        rec->pivot_apps = (double)st->application_nr_red - r0;
        flatten_output(&rec->out, &rec->out_bindex, &rec->out_mult, mat, st);
        return;
    }
    ref_fn(mat, tbr, bs, st);
    g_la_seconds += realtime() - t0;
}

static void hook_la(mat_t *mat, const bs_t *const tbr, const bs_t *const bs, md_t *st)
{ hook_common(0, mat, tbr, bs, st); }
static void hook_exact_la(mat_t *mat, const bs_t *const tbr, const bs_t *const bs, md_t *st)
{ hook_common(1, mat, tbr, bs, st); }

static void hook_interred(mat_t *mat, bs_t *bs, md_t *st, int free_basis)
{
    const double t0 = realtime();
    if ((g_mode == 1 || g_mode == 4) && g_backend_interred) g_backend_interred(mat, bs, st, free_basis);
    else dispatch_interreduce_matrix_rows(mat, bs, st, free_basis);
    g_la_seconds += realtime() - t0;
}

/* application_linear_algebra / trace_linear_algebra (data.h:559-576): only the legacy tracer
 * functions of modular.c call them; served by the backend whenever one is installed */
static int hook_app(mat_t *mat, const bs_t *const bs, md_t *st)
{
    const double t0 = realtime();
    const int r = ((g_mode == 1 || g_mode == 4) && g_backend_app) ? g_backend_app(mat, bs, st)
                                                                   : dispatch_application_linear_algebra(mat, bs, st);
    g_la_seconds += realtime() - t0; g_la_calls++;
    return r;
}

static void hook_trace(trace_t *trace, mat_t *mat, const bs_t *const bs, md_t *st)
{
    const double t0 = realtime();
    if ((g_mode == 1 || g_mode == 4) && g_backend_trace) g_backend_trace(trace, mat, bs, st);
    else dispatch_trace_linear_algebra(trace, mat, bs, st);
    g_la_seconds += realtime() - t0; g_la_calls++;
}

/* ------------------------------------------------------------------ */
/* exported API                                                        */
/* ------------------------------------------------------------------ */

void refh_set_mode(int mode, int max_records, int verbose)
{
    g_mode = mode; g_max_records = max_records > 0 ? max_records : (1 << 30);
    g_verbose = verbose;
    /* all five matrix-level pointers of data.h:529-595 go through the harness */
    linear_algebra = hook_la;
    exact_linear_algebra = hook_exact_la;
    interreduce_matrix_rows = hook_interred;
    application_linear_algebra = hook_app;
    trace_linear_algebra = hook_trace;
}

void refh_set_backend(void *la, void *exact_la, void *interred)
{
    g_backend_la = (la_fn_t)la;
    g_backend_exact_la = (la_fn_t)exact_la;
    g_backend_interred = (ir_fn_t)interred;
}

void refh_set_backend_tracer(void *app, void *trace)
{
    g_backend_app = (app_fn_t)app;
    g_backend_trace = (tr_fn_t)trace;
}

/* which of the five pointers currently point at the harness hooks (bit i set), for the tests */
int refh_hooked_pointers(void)
{
    return (linear_algebra == hook_la) | (exact_linear_algebra == hook_exact_la) << 1 |
           (interreduce_matrix_rows == hook_interred) << 2 | (application_linear_algebra == hook_app) << 3 |
           (trace_linear_algebra == hook_trace) << 4;
}

void refh_reset(void)
{
    for (int i = 0; i < g_nrec; ++i) free_record(&g_rec[i]);
    g_nrec = 0;
    g_shadow_mismatch = 0; g_shadow_calls = 0; g_la_seconds = 0.0; g_la_calls = 0;
}

int     refh_num_records(void)      { return g_nrec; }
int64_t refh_shadow_mismatches(void){ return g_shadow_mismatch; }
int64_t refh_shadow_calls(void)     { return g_shadow_calls; }
double  refh_la_seconds(void)       { return g_la_seconds; }
int     refh_la_calls(void)         { return g_la_calls; }
const refh_record *refh_get_record(int i) { return (i >= 0 && i < g_nrec) ? &g_rec[i] : NULL; }
void    refh_drop_record(int i)     { if (i >= 0 && i < g_nrec) free_record(&g_rec[i]); }
size_t  refh_record_size(void)      { return sizeof(refh_record); }

/* Result of one F4 run, owned by the harness until the next run. */
typedef struct refh_gb {
    int32_t  nelts;
    int64_t  nterms;
    int32_t *lens;      /* [nelts]            */
    int32_t *exps;      /* [nterms * nvars]   */
    int32_t *cfs;       /* [nterms]           */
    double   f4_seconds;
    double   la_seconds;
    int32_t  la_calls;
} refh_gb;

static refh_gb g_gb;

/* Runs the reference F4 (export_f4, f4.c:769-899) on the given system. */
const refh_gb *refh_f4(const int32_t *lens, const int32_t *exps, const int32_t *cfs,
                       uint32_t field_char, int32_t nvars, int32_t ngens,
                       int32_t threads, int32_t laopt, int32_t info_level)
{
    free(g_gb.lens); free(g_gb.exps); free(g_gb.cfs);
    memset(&g_gb, 0, sizeof(g_gb));
    g_la_seconds = 0.0; g_la_calls = 0;
    int32_t bld = 0; int32_t *blen = NULL, *bexp = NULL; void *bcf = NULL;
    const double t0 = realtime();
    int64_t nterms = export_f4(malloc, &bld, &blen, &bexp, &bcf, lens, exps, cfs,
                               field_char, 0 /* DRL */, 0, nvars, ngens, 17,
                               threads, 0, 0, laopt, 1 /* reduce gb */, 0, info_level);
    g_gb.f4_seconds = realtime() - t0;
    g_gb.la_seconds = g_la_seconds; g_gb.la_calls = g_la_calls;
    g_gb.nelts = bld; g_gb.nterms = nterms;
    g_gb.lens = blen; g_gb.exps = bexp; g_gb.cfs = (int32_t *)bcf;
    return &g_gb;
}

/* Runs the reference LA on one flat matrix: builds mat_t/bs_t/md_t exactly
 * as the F4 driver would hand them over (rows malloc()ed one by one with the
 * 6-word prelude, data.h:49-59) and calls the dispatcher (io.c:878/962). */
int refh_reference_la(const f4la_flat_matrix *in, int32_t laopt, int32_t threads,
                      f4la_flat_result *out, double *seconds, double *units,
                      double *pivot_apps)
{
    const uint32_t n = in->nru + in->nrl;
    const size_t w = in->cf_bytes;
    const int ff_bits = w == 1 ? 8 : w == 2 ? 16 : 32;
    mat_t *mat = calloc(1, sizeof(mat_t));
    bs_t  *bs  = calloc(1, sizeof(bs_t));
    md_t  *st  = calloc(1, sizeof(md_t));
    st->fc = in->fc; st->gfc = in->fc; st->nthrds = threads; st->laopt = laopt;
    st->ff_bits = ff_bits; st->trace_level = NO_TRACER; st->info_level = 0;
    st->nf = (in->flags & F4LA_FLAG_NORMALFORM) ? 1 : 0;
    st->in_final_reduction_step = (in->flags & F4LA_FLAG_PIVOT_BY_LEAD) ? 1 : 0;
    reset_function_pointers(in->fc, laopt);
    bs->ld = bs->sz = in->ncf;
    void **cfarr = calloc(in->ncf ? in->ncf : 1, sizeof(void *));
    for (uint32_t c = 0; c < in->ncf; ++c)
        cfarr[c] = (unsigned char *)in->cf + in->cf_ptr[c] * w;   /* never freed by the LA */
    bs->cf_8 = (cf8_t **)cfarr; bs->cf_16 = (cf16_t **)cfarr; bs->cf_32 = (cf32_t **)cfarr;
    mat->nru = in->nru; mat->nrl = in->nrl; mat->ncl = in->ncl; mat->ncr = in->ncr;
    mat->nc = in->ncl + in->ncr; mat->nr = mat->sz = n;
    mat->rr = malloc((in->nru ? in->nru : 1) * sizeof(hm_t *));
    mat->tr = malloc((in->nrl ? in->nrl : 1) * sizeof(hm_t *));
    for (uint32_t r = 0; r < n; ++r) {
        const uint64_t len = in->row_ptr[r + 1] - in->row_ptr[r];
        hm_t *row = malloc((len + OFFSET) * sizeof(hm_t));
        row[DEG] = 0; row[BINDEX] = r; row[MULT] = 0; row[COEFFS] = in->row_cf[r];
        row[PRELOOP] = (hm_t)(len % UNROLL); row[LENGTH] = (hm_t)len;
        memcpy(row + OFFSET, in->cols + in->row_ptr[r], len * sizeof(hm_t));
        if (r < in->nru) mat->rr[r] = row; else mat->tr[r - in->nru] = row;
    }
    const double m0 = st->application_nr_mult;
    const double t0 = realtime();
    if (in->flags & (F4LA_FLAG_NORMALFORM | F4LA_FLAG_PIVOT_BY_LEAD))
        dispatch_exact_linear_algebra(mat, bs, bs, st);
    else
        dispatch_linear_algebra(mat, bs, bs, st);
    if (seconds) *seconds = realtime() - t0;
    if (units) *units = (st->application_nr_mult - m0) * 1000.0;
    if (pivot_apps) *pivot_apps = (double)st->application_nr_red;
    uint32_t *bi, *mu;
    flatten_output(out, &bi, &mu, mat, st);
    out->src_row = bi;      /* BINDEX was set to the input row index above */
    free(mu);
    /* release what the LA left behind */
    for (len_t i = 0; i < mat->np; ++i) {
        if (!mat->tr[i]) continue;
        hm_t ci = mat->tr[i][COEFFS];
        if (ff_bits == 8) free(mat->cf_8[ci]); else if (ff_bits == 16) free(mat->cf_16[ci]); else free(mat->cf_32[ci]);
        free(mat->tr[i]);
    }
    /* -l 44 frees mat->rr itself and NULLs it (la_ff_32.c:2495-2496) */
    free(mat->rr);
    free(mat->tr); free(mat->cf_8); free(mat->cf_16); free(mat->cf_32);
    free(mat); free(cfarr); free(bs); free(st);
    return 0;
}

/* side channels of the last refh_call_entry() (SURVEY 8b "side channels to honour"):
 * st->np, mat->np, mat->nr, mat->sz, st->num_zerored, st->la_rtime, st->la_ctime,
 * st->application_nr_mult, _add, _red */
static double g_side[11];      /* [10]: wall seconds of the entry-point call alone */
static double g_call_t0;
static void snapshot_side(const mat_t *mat, const md_t *st)
{
    g_side[0] = st->np; g_side[1] = mat->np; g_side[2] = mat->nr; g_side[3] = mat->sz;
    g_side[4] = (double)st->num_zerored; g_side[5] = st->la_rtime; g_side[6] = st->la_ctime;
    g_side[7] = st->application_nr_mult; g_side[8] = st->application_nr_add; g_side[9] = (double)st->application_nr_red;
}
void refh_last_side(double *o) { memcpy(o, g_side, sizeof(g_side)); }

/* Generic driver for the matrix-level entry points on a flat matrix.
 *   which 0: fn(mat, tbr, bs, st)            linear_algebra / exact_linear_algebra signature
 *   which 1: int fn(mat, bs, st)             application_linear_algebra signature
 *   which 2: fn(mat, bs, st, free_basis=0)   interreduce_matrix_rows signature; the LOWER rows of
 *                                            the flat matrix are handed over as mat->rr
 * fn == NULL runs the reference's own dispatcher.  *ret receives the int result (which 1). */
int refh_call_entry(const f4la_flat_matrix *in, int32_t which, void *fn, int32_t threads,
                    f4la_flat_result *out, int32_t *ret, uint32_t *out_coeffs_index)
{
    const uint32_t n = in->nru + in->nrl;
    const size_t w = in->cf_bytes;
    const int ff_bits = w == 1 ? 8 : w == 2 ? 16 : 32;
    mat_t *mat = calloc(1, sizeof(mat_t));
    bs_t  *bs  = calloc(1, sizeof(bs_t));
    md_t  *st  = calloc(1, sizeof(md_t));
    st->fc = in->fc; st->gfc = in->fc; st->nthrds = threads; st->laopt = 2;
    st->ff_bits = ff_bits; st->trace_level = NO_TRACER; st->info_level = 0;
    st->nf = (in->flags & F4LA_FLAG_NORMALFORM) ? 1 : 0;
    st->in_final_reduction_step = (in->flags & F4LA_FLAG_PIVOT_BY_LEAD) ? 1 : 0;
    bs->ld = bs->sz = in->ncf;
    void **cfarr = calloc(in->ncf ? in->ncf : 1, sizeof(void *));
    for (uint32_t c = 0; c < in->ncf; ++c) cfarr[c] = (unsigned char *)in->cf + in->cf_ptr[c] * w;
    bs->cf_8 = (cf8_t **)cfarr; bs->cf_16 = (cf16_t **)cfarr; bs->cf_32 = (cf32_t **)cfarr;
    mat->ncl = in->ncl; mat->ncr = in->ncr; mat->nc = in->ncl + in->ncr;
    hm_t **rows = malloc((n ? n : 1) * sizeof(hm_t *));
    for (uint32_t r = 0; r < n; ++r) {
        const uint64_t len = in->row_ptr[r + 1] - in->row_ptr[r];
        hm_t *row = malloc((len + OFFSET) * sizeof(hm_t));
        row[DEG] = 0; row[BINDEX] = r; row[MULT] = 7 * r + 1; row[COEFFS] = in->row_cf[r];
        row[PRELOOP] = (hm_t)(len % UNROLL); row[LENGTH] = (hm_t)len;
        memcpy(row + OFFSET, in->cols + in->row_ptr[r], len * sizeof(hm_t));
        rows[r] = row;
    }
    if (which == 2) {
        if (in->nru) { free(rows); return -1; }
        mat->rr = rows; mat->tr = NULL; mat->nr = mat->nru = in->nrl; mat->nrl = 0; mat->sz = in->nrl;
        g_call_t0 = realtime();
        if (fn) ((ir_fn_t)fn)(mat, bs, st, 0); else dispatch_interreduce_matrix_rows(mat, bs, st, 0);
    } else {
        mat->nru = in->nru; mat->nrl = in->nrl; mat->nr = mat->sz = n;
        mat->rr = malloc((in->nru ? in->nru : 1) * sizeof(hm_t *));
        mat->tr = malloc((in->nrl ? in->nrl : 1) * sizeof(hm_t *));
        memcpy(mat->rr, rows, in->nru * sizeof(hm_t *));
        memcpy(mat->tr, rows + in->nru, in->nrl * sizeof(hm_t *));
        free(rows);
        g_call_t0 = realtime();
        if (which == 1) {
            const int r = fn ? ((app_fn_t)fn)(mat, bs, st) : 0;
            if (!fn) dispatch_linear_algebra(mat, bs, bs, st);      /* the reference's own app variant is stale */
            if (ret) *ret = fn ? r : (mat->np != in->nrl);
        } else {
            if (fn) ((la_fn_t)fn)(mat, bs, bs, st);
            else if (in->flags & (F4LA_FLAG_NORMALFORM | F4LA_FLAG_PIVOT_BY_LEAD)) dispatch_exact_linear_algebra(mat, bs, bs, st);
            else dispatch_linear_algebra(mat, bs, bs, st);
        }
    }
    g_side[10] = realtime() - g_call_t0;
    snapshot_side(mat, st);
    uint32_t *bi, *mu;
    flatten_output(out, &bi, &mu, mat, st);
    out->src_row = bi;
    if (out_coeffs_index)
        for (len_t i = 0; i < mat->np; ++i) out_coeffs_index[i] = mat->tr[i] ? mat->tr[i][COEFFS] : 0xffffffffu;
    /* MULT was set to 7*BINDEX+1: a backend must carry both over together (the reference's own
     * 8/16-bit reducers leave them unset, la_ff_16.c:407-409, so it is not checked itself) */
    int bad = 0;
    if (fn) for (len_t i = 0; i < mat->np; ++i) if (mat->tr[i] && mu[i] != 7 * bi[i] + 1) bad = 1;
    free(mu);
    for (len_t i = 0; i < mat->np; ++i) {
        if (!mat->tr || !mat->tr[i]) continue;
        hm_t ci = mat->tr[i][COEFFS];
        if (ff_bits == 8) free(mat->cf_8[ci]); else if (ff_bits == 16) free(mat->cf_16[ci]); else free(mat->cf_32[ci]);
        free(mat->tr[i]);
    }
    free(mat->rr); free(mat->tr); free(mat->cf_8); free(mat->cf_16); free(mat->cf_32);
    free(mat); free(cfarr); free(bs); free(st);
    return bad ? -2 : 0;
}

/* trace_linear_algebra signature (data.h:571-576) on a flat matrix: mat_t with the reducer bit arrays
 * symbolic preprocessing allocates for a learning run (symbol.c:625-629) and a fresh trace_t
 * (initialize_trace, modular.c:45-64).  fn == NULL runs the reference's dispatcher.  Besides the rows the
 * trace entry td[0] written by construct_trace (tools.c:63-183) is returned in tinfo (capacity tcap words):
 *   tinfo[0] = ntr (kept lower rows), tinfo[1] = nrr (reducers used by them), tinfo[2] = words per row,
 *   then ntr BINDEX values of the kept rows, nrr BINDEX values of the reducers, ntr x words bit arrays. */
int refh_call_trace_entry(const f4la_flat_matrix *in, void *fn, int32_t threads,
                          f4la_flat_result *out, uint32_t *tinfo, uint64_t tcap)
{
    const uint32_t n = in->nru + in->nrl;
    const size_t w = in->cf_bytes;
    const int ff_bits = w == 1 ? 8 : w == 2 ? 16 : 32;
    mat_t *mat = calloc(1, sizeof(mat_t));
    bs_t  *bs  = calloc(1, sizeof(bs_t));
    md_t  *st  = calloc(1, sizeof(md_t));
    st->fc = in->fc; st->gfc = in->fc; st->nthrds = threads; st->laopt = 2;
    st->ff_bits = ff_bits; st->trace_level = LEARN_TRACER; st->info_level = 0;
    bs->ld = bs->sz = in->ncf;
    void **cfarr = calloc(in->ncf ? in->ncf : 1, sizeof(void *));
    for (uint32_t c = 0; c < in->ncf; ++c) cfarr[c] = (unsigned char *)in->cf + in->cf_ptr[c] * w;
    bs->cf_8 = (cf8_t **)cfarr; bs->cf_16 = (cf16_t **)cfarr; bs->cf_32 = (cf32_t **)cfarr;
    mat->ncl = in->ncl; mat->ncr = in->ncr; mat->nc = in->ncl + in->ncr;
    mat->nru = in->nru; mat->nrl = in->nrl; mat->nr = mat->sz = n;
    mat->rr = malloc((in->nru ? in->nru : 1) * sizeof(hm_t *));
    mat->tr = malloc((in->nrl ? in->nrl : 1) * sizeof(hm_t *));
    for (uint32_t r = 0; r < n; ++r) {
        const uint64_t len = in->row_ptr[r + 1] - in->row_ptr[r];
        hm_t *row = malloc((len + OFFSET) * sizeof(hm_t));
        row[DEG] = 0; row[BINDEX] = r; row[MULT] = 7 * r + 1; row[COEFFS] = in->row_cf[r];
        row[PRELOOP] = (hm_t)(len % UNROLL); row[LENGTH] = (hm_t)len;
        memcpy(row + OFFSET, in->cols + in->row_ptr[r], len * sizeof(hm_t));
        if (r < in->nru) mat->rr[r] = row; else mat->tr[r - in->nru] = row;
    }
    const uint64_t lrba = in->nru / 32 + ((in->nru % 32) != 0);
    mat->rbal = in->nrl;
    mat->rba = malloc((in->nrl ? in->nrl : 1) * sizeof(rba_t *));
    for (uint32_t i = 0; i < in->nrl; ++i) mat->rba[i] = calloc(lrba ? lrba : 1, sizeof(rba_t));
    trace_t *trace = calloc(1, sizeof(trace_t));
    trace->std = 8; trace->sts = 8;
    trace->td = calloc(trace->std, sizeof(td_t));
    trace->ts = calloc(trace->sts, sizeof(ts_t));

    if (fn) ((tr_fn_t)fn)(trace, mat, bs, st); else dispatch_trace_linear_algebra(trace, mat, bs, st);

    uint32_t *bi, *mu;
    flatten_output(out, &bi, &mu, mat, st);
    out->src_row = bi;
    int bad = 0;
    for (len_t i = 0; i < mat->np; ++i) if (fn && mat->tr[i] && mu[i] != 7 * bi[i] + 1) bad = 1;
    free(mu);
    /* the trace entry of this matrix (the caller of trace_linear_algebra bumps ltd afterwards) */
    const td_t *td = &trace->td[0];
    const uint32_t ntr = td->tld / 2, nrr = td->rld / 2;
    const uint32_t words = nrr / 32 + ((nrr % 32) != 0);
    if (tinfo && tcap >= 3 + (uint64_t)ntr + nrr + (uint64_t)ntr * words) {
        tinfo[0] = ntr; tinfo[1] = nrr; tinfo[2] = words;
        uint32_t *q = tinfo + 3;
        for (uint32_t i = 0; i < ntr; ++i) { *q++ = td->tri[2 * i]; if (td->tri[2 * i + 1] != 7 * td->tri[2 * i] + 1) bad = 1; }
        for (uint32_t i = 0; i < nrr; ++i) { *q++ = td->rri[2 * i]; if (td->rri[2 * i + 1] != 7 * td->rri[2 * i] + 1) bad = 1; }
        for (uint32_t i = 0; i < ntr; ++i) { memcpy(q, td->rba[i], (size_t)words * sizeof(rba_t)); q += words; }
    } else if (tinfo && tcap >= 3) {
        tinfo[0] = ntr; tinfo[1] = nrr; tinfo[2] = 0xffffffffu;
    }
    /* construct_trace compacts mat->rba to the kept rows (or leaves it alone when every row vanished) */
    for (len_t i = 0; i < mat->rbal; ++i) free(mat->rba[i]);
    free(mat->rba);
    for (uint32_t i = 0; i < ntr; ++i) free(td->rba[i]);
    free(td->rba); free(td->tri); free(td->rri);
    free(trace->td); free(trace->ts); free(trace);
    for (len_t i = 0; i < mat->np; ++i) {
        if (!mat->tr || !mat->tr[i]) continue;
        hm_t ci = mat->tr[i][COEFFS];
        if (ff_bits == 8) free(mat->cf_8[ci]); else if (ff_bits == 16) free(mat->cf_16[ci]); else free(mat->cf_32[ci]);
        free(mat->tr[i]);
    }
    free(mat->rr); free(mat->tr); free(mat->cf_8); free(mat->cf_16); free(mat->cf_32);
    free(mat); free(cfarr); free(bs); free(st);
    return bad ? -2 : 0;
}

void refh_free_result(f4la_flat_result *r)
{
    free(r->row_ptr); free(r->cols); free(r->cf); free(r->src_row);
    memset(r, 0, sizeof(*r));
}

/* Offsets of the reference structs, for the ABI check of the product header.
 * Order is fixed; see tests/test_abi_layout.py. */
int refh_struct_layout(uint64_t *o, int cap)
{
    int k = 0;
#define PUSH(x) do { if (k < cap) o[k] = (uint64_t)(x); k++; } while (0)
    PUSH(sizeof(mat_t));
    PUSH(offsetof(mat_t, tr)); PUSH(offsetof(mat_t, rba)); PUSH(offsetof(mat_t, rr));
    PUSH(offsetof(mat_t, cf_8)); PUSH(offsetof(mat_t, cf_16)); PUSH(offsetof(mat_t, cf_32));
    PUSH(offsetof(mat_t, cf_qq)); PUSH(offsetof(mat_t, cf_ab_qq));
    PUSH(offsetof(mat_t, sz)); PUSH(offsetof(mat_t, np)); PUSH(offsetof(mat_t, nr));
    PUSH(offsetof(mat_t, nc)); PUSH(offsetof(mat_t, nru)); PUSH(offsetof(mat_t, nrl));
    PUSH(offsetof(mat_t, ncl)); PUSH(offsetof(mat_t, ncr)); PUSH(offsetof(mat_t, rbal));
    PUSH(offsetof(mat_t, cd));
    PUSH(sizeof(bs_t));
    PUSH(offsetof(bs_t, ld)); PUSH(offsetof(bs_t, sz)); PUSH(offsetof(bs_t, hm));
    PUSH(offsetof(bs_t, cf_8)); PUSH(offsetof(bs_t, cf_16)); PUSH(offsetof(bs_t, cf_32));
    PUSH(sizeof(md_t));
    PUSH(offsetof(md_t, tr)); PUSH(offsetof(md_t, trace_level)); PUSH(offsetof(md_t, np));
    PUSH(offsetof(md_t, la_ctime)); PUSH(offsetof(md_t, la_rtime));
    PUSH(offsetof(md_t, num_zerored)); PUSH(offsetof(md_t, fc)); PUSH(offsetof(md_t, laopt));
    PUSH(offsetof(md_t, nthrds)); PUSH(offsetof(md_t, ff_bits)); PUSH(offsetof(md_t, nf));
    PUSH(offsetof(md_t, in_final_reduction_step)); PUSH(offsetof(md_t, info_level));
    PUSH(offsetof(md_t, application_nr_mult)); PUSH(offsetof(md_t, application_nr_add));
    PUSH(offsetof(md_t, application_nr_red));
    PUSH(OFFSET); PUSH(LENGTH); PUSH(PRELOOP); PUSH(COEFFS); PUSH(MULT); PUSH(BINDEX); PUSH(DEG);
#undef PUSH
    return k;
}

/* The reference-side half of the drop-in binding: describe mat_t/bs_t/md_t to
 * the backend with this compiler's offsetof. */
void refh_fill_layout(f4la_neogb_layout *l)
{
    memset(l, 0, sizeof(*l));
    l->struct_size = sizeof(*l);
    l->mat_tr = offsetof(mat_t, tr); l->mat_rba = offsetof(mat_t, rba); l->mat_rr = offsetof(mat_t, rr);
    l->mat_cf_8 = offsetof(mat_t, cf_8); l->mat_cf_16 = offsetof(mat_t, cf_16); l->mat_cf_32 = offsetof(mat_t, cf_32);
    l->mat_sz = offsetof(mat_t, sz); l->mat_np = offsetof(mat_t, np); l->mat_nr = offsetof(mat_t, nr);
    l->mat_nc = offsetof(mat_t, nc); l->mat_nru = offsetof(mat_t, nru); l->mat_nrl = offsetof(mat_t, nrl);
    l->mat_ncl = offsetof(mat_t, ncl); l->mat_ncr = offsetof(mat_t, ncr); l->mat_rbal = offsetof(mat_t, rbal);
    l->bs_ld = offsetof(bs_t, ld); l->bs_cf_8 = offsetof(bs_t, cf_8); l->bs_cf_16 = offsetof(bs_t, cf_16);
    l->bs_cf_32 = offsetof(bs_t, cf_32);
    l->md_trace_level = offsetof(md_t, trace_level); l->md_np = offsetof(md_t, np);
    l->md_la_ctime = offsetof(md_t, la_ctime); l->md_la_rtime = offsetof(md_t, la_rtime);
    l->md_num_zerored = offsetof(md_t, num_zerored); l->md_fc = offsetof(md_t, fc);
    l->md_laopt = offsetof(md_t, laopt); l->md_nthrds = offsetof(md_t, nthrds);
    l->md_ff_bits = offsetof(md_t, ff_bits); l->md_nf = offsetof(md_t, nf);
    l->md_in_final_reduction_step = offsetof(md_t, in_final_reduction_step);
    l->md_info_level = offsetof(md_t, info_level);
    l->md_application_nr_mult = offsetof(md_t, application_nr_mult);
    l->md_application_nr_add = offsetof(md_t, application_nr_add);
    l->md_application_nr_red = offsetof(md_t, application_nr_red);
    l->md_tr = offsetof(md_t, tr);
    l->row_offset = OFFSET; l->row_length = LENGTH; l->row_preloop = PRELOOP; l->row_coeffs = COEFFS;
    l->row_mult = MULT; l->row_bindex = BINDEX; l->row_deg = DEG; l->unroll = UNROLL;
}

/* construct_trace() is static in the reference's unity build; the drop-in
 * stub passes its address to the backend. */
static void construct_trace_thunk(void *trace, void *mat) { construct_trace((trace_t *)trace, (mat_t *)mat); }
void *refh_construct_trace_ptr(void) { return (void *)construct_trace_thunk; }

/* likewise free_basis_elements() (basis.c:24-70), which interreduce_matrix_rows(free_basis != 0) calls */
static void free_basis_elements_thunk(void *bs) { free_basis_elements((bs_t *)bs); }
void *refh_free_basis_elements_ptr(void) { return (void *)free_basis_elements_thunk; }

/* ------------------------------------------------------------------ */
/* QQ multi-modular tracer flow (learn once, apply per prime)          */
/* ------------------------------------------------------------------ */
/* This is what msolve does for rational input (src/msolve/msolve.c:2145 ff.):
 *   setup   : validate_input_data, check_and_set_meta_data_trace,
 *             initialize_basis, import_input_data, calculate_divmask, sort,
 *             remove_content_of_initial_basis          (msolve.c:2189-2240)
 *   learn   : st->f4_qq_round = 1; core_gba(bs_qq, st, &err, p0)   (msolve.c:1628-1632)
 *   apply   : st->f4_qq_round = 2; st->nthrds = 1; one core_gba per prime,
 *             primes in parallel on OpenMP threads          (msolve.c:1815-1824)
 * FGLM / CRT (which need FLINT) are not part of it. */

static md_t *g_qq_st = NULL;
static bs_t *g_qq_bs = NULL;
static int   g_qq_nvars = 0;
static void (*g_set_device)(int) = NULL;

void refh_set_device_callback(void *fn) { g_set_device = (void (*)(int))fn; }

/* FNV-1a over the reduced basis of a modular F4 run: lengths, coefficients, exponents */
static uint64_t basis_digest(const bs_t *bs, int32_t *nelts, int64_t *nterms)
{
    uint64_t h = 1469598103934665603ULL;
#define MIX(v) do { uint64_t x_ = (uint64_t)(v); for (int b_ = 0; b_ < 8; ++b_) { h ^= (x_ >> (8 * b_)) & 0xff; h *= 1099511628211ULL; } } while (0)
    const ht_t *ht = bs->ht;
    int64_t nt = 0;
    for (bl_t i = 0; i < bs->lml; ++i) {
        const hm_t *hm = bs->hm[bs->lmps[i]];
        const cf32_t *cf = bs->cf_32[hm[COEFFS]];
        MIX(hm[LENGTH]);
        for (len_t k = 0; k < hm[LENGTH]; ++k) {
            MIX(cf[k]);
            const exp_t *e = ht->ev[hm[OFFSET + k]];
            for (len_t v = 0; v < ht->evl; ++v) MIX(e[v]);
        }
        nt += hm[LENGTH];
    }
#undef MIX
    if (nelts) *nelts = (int32_t)bs->lml;
    if (nterms) *nterms = nt;
    return h;
}

int refh_qq_setup(const int32_t *lens, const int32_t *exps, const int32_t *icfs,
                  int32_t nvars, int32_t ngens, int32_t threads, int32_t laopt, int32_t info_level)
{
    int64_t nterms = 0;
    for (int i = 0; i < ngens; ++i) nterms += lens[i];
    /* rational coefficients: numerator, denominator per term (io.c:380-405) */
    mpz_t **cfs = malloc((size_t)(2 * nterms) * sizeof(mpz_t *));
    for (int64_t k = 0; k < nterms; ++k) {
        cfs[2 * k] = malloc(sizeof(mpz_t)); cfs[2 * k + 1] = malloc(sizeof(mpz_t));
        mpz_init(*cfs[2 * k]); mpz_set_si(*cfs[2 * k], icfs[k]);
        mpz_init(*cfs[2 * k + 1]); mpz_set_si(*cfs[2 * k + 1], 1);
    }
    uint32_t field_char = 0;
    int32_t mon_order = 0, elim_block_len = 0, nr_vars = nvars, nr_gens = ngens, nr_nf = 0, ht_size = 17;
    int32_t nr_threads = threads, max_nr_pairs = 0, reset_ht = 0, la_option = laopt, use_signatures = 0;
    int32_t reduce_gb = 1, truncate_lifting = 0, info = info_level;
    int *invalid_gens = NULL;
    if (validate_input_data(&invalid_gens, cfs, lens, &field_char, &mon_order, &elim_block_len, &nr_vars,
                            &nr_gens, &nr_nf, &ht_size, &nr_threads, &max_nr_pairs, &reset_ht, &la_option,
                            &use_signatures, &reduce_gb, &truncate_lifting, &info) == -1) return -1;
    md_t *st = allocate_meta_data();
    if (check_and_set_meta_data_trace(st, lens, exps, cfs, invalid_gens, field_char, mon_order, elim_block_len,
                                      nr_vars, nr_gens, nr_nf, ht_size, nr_threads, max_nr_pairs, reset_ht,
                                      la_option, use_signatures, reduce_gb, (uint32_t)1 << 30, nr_threads,
                                      0, 0, info)) { free(st); return -2; }
    bs_t *bs = initialize_basis(st, NULL);
    import_input_data(bs, st, 0, st->ngens_input, lens, exps, cfs, invalid_gens);
    calculate_divmask(bs->ht);
    sort_r(bs->hm, (unsigned long)bs->ld, sizeof(hm_t *), initial_input_cmp, bs->ht);
    remove_content_of_initial_basis(bs);
    for (int64_t k = 0; k < 2 * nterms; ++k) { mpz_clear(*cfs[k]); free(cfs[k]); }
    free(cfs); free(invalid_gens);
    g_qq_st = st; g_qq_bs = bs; g_qq_nvars = nvars;
    return 0;
}

typedef struct refh_qq_result {
    uint64_t digest;
    int32_t  nelts;
    int64_t  nterms;
    int32_t  err;
    double   seconds;
    double   la_seconds;
} refh_qq_result;

/* learning phase with prime p0 (all st->nthrds threads inside F4) */
int refh_qq_learn(uint32_t p0, refh_qq_result *res)
{
    if (!g_qq_st) return -1;
    memset(res, 0, sizeof(*res));
    g_la_seconds = 0.0; g_la_calls = 0;
    g_qq_st->f4_qq_round = 1;
    int32_t err = 0;
    const double t0 = realtime();
    bs_t *bs = core_gba(g_qq_bs, g_qq_st, &err, p0);
    res->seconds = realtime() - t0;
    res->la_seconds = g_la_seconds;
    res->err = err;
    if (err || !bs) return 1;
    res->digest = basis_digest(bs, &res->nelts, &res->nterms);
    /* in the learning round the modular basis shares bs_qq's hash table
     * (basis.c:398-402), which the application rounds still need */
    free_basis_without_hash_table(&bs);
    return 0;
}

/* application phase: one F4 replay per prime, `threads` primes at a time, host
 * thread t bound to GPU t % ngpus when a backend with a device callback is installed */
/* coefficients of the reduced modular basis, element after element (the residues msolve lifts by CRT,
 * msolve.c:2694-2712); returns the number of words written (<= cap) */
static int64_t basis_residues(const bs_t *bs, uint32_t *out, int64_t cap)
{
    int64_t n = 0;
    for (bl_t i = 0; i < bs->lml; ++i) {
        const hm_t *hm = bs->hm[bs->lmps[i]];
        const cf32_t *cf = bs->cf_32[hm[COEFFS]];
        for (len_t k = 0; k < hm[LENGTH]; ++k) { if (n < cap) out[n] = cf[k]; ++n; }
    }
    return n;
}

int refh_qq_apply_residues(const uint32_t *primes, int32_t nprimes, int32_t threads, int32_t ngpus,
                           refh_qq_result *res, double *wall_seconds, uint32_t *residues, int64_t stride);

int refh_qq_apply(const uint32_t *primes, int32_t nprimes, int32_t threads, int32_t ngpus,
                  refh_qq_result *res, double *wall_seconds)
{
    return refh_qq_apply_residues(primes, nprimes, threads, ngpus, res, wall_seconds, NULL, 0);
}

int refh_qq_apply_residues(const uint32_t *primes, int32_t nprimes, int32_t threads, int32_t ngpus,
                           refh_qq_result *res, double *wall_seconds, uint32_t *residues, int64_t stride)
{
    if (!g_qq_st || g_qq_st->trace_level != APPLY_TRACER) return -1;
    g_qq_st->f4_qq_round = 2;
    const int32_t old_thr = g_qq_st->nthrds, old_info = g_qq_st->info_level;
    g_qq_st->nthrds = 1; g_qq_st->info_level = 0;
    const double t0 = realtime();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int32_t i = 0; i < nprimes; ++i) {
        if (g_set_device && ngpus > 0) g_set_device(omp_get_thread_num() % ngpus);
        int32_t err = 0;
        const double t1 = realtime();
        bs_t *bs = core_gba(g_qq_bs, g_qq_st, &err, primes[i]);
        memset(&res[i], 0, sizeof(res[i]));
        res[i].err = err;
        if (!err && bs) res[i].digest = basis_digest(bs, &res[i].nelts, &res[i].nterms);
        if (!err && bs && residues) basis_residues(bs, residues + (int64_t)i * stride, stride);
        res[i].seconds = realtime() - t1;
        if (bs) {
            if (err) free(bs); else free_basis_and_only_local_hash_table_data(&bs);
        }
    }
    if (wall_seconds) *wall_seconds = realtime() - t0;
    g_qq_st->nthrds = old_thr; g_qq_st->info_level = old_info;
    return 0;
}

/* The initial basis of an application run for the prime p, as initialize_f4 prepares it (f4.c:362-371):
 * copy_basis_mod_p (basis.c:382-464: mpz_fdiv_ui of every coefficient) + normalize_initial_basis (basis.c:340-379:
 * times the inverse of the lead coefficient), without the copies of the hash tables.  Element i (basis index i) at
 * out_cf[out_off[i] .. out_off[i+1]).  Returns the number of elements, or -1 when cap is too small. */
int refh_qq_initial_basis(uint32_t p, uint32_t *out_cf, uint64_t *out_off, int64_t cap)
{
    if (!g_qq_bs) return -1;
    const bs_t *gbs = g_qq_bs;
    const bl_t ld = gbs->ld;
    /* coefficient arrays are addressed by their COEFFS index (the rows of the matrices carry it), which the
     * initial sort of bs->hm does not change: lay the arrays out by that index */
    len_t *len_of = calloc(ld ? ld : 1, sizeof(len_t));
    for (bl_t i = 0; i < ld; ++i) {
        const hm_t idx = gbs->hm[i][COEFFS];
        if (idx >= ld) { free(len_of); return -1; }
        len_of[idx] = gbs->hm[i][LENGTH];
    }
    uint64_t n = 0;
    for (bl_t idx = 0; idx < ld; ++idx) { out_off[idx] = n; n += len_of[idx]; }
    out_off[ld] = n;
    if ((int64_t)n > cap) { free(len_of); return -1; }
    for (bl_t idx = 0; idx < ld; ++idx) {
        const len_t len = len_of[idx];
        if (!len) continue;
        uint32_t *row = out_cf + out_off[idx];
        for (len_t j = 0; j < len; ++j) row[j] = (uint32_t)mpz_fdiv_ui(gbs->cf_qq[idx][j], (unsigned long)p);
        const uint32_t inv = mod_p_inverse_32((int64_t)row[0], (int64_t)p);
        for (len_t j = 0; j < len; ++j) {
            int64_t t = ((int64_t)row[j] * inv) % p;
            t += (t >> 63) & p;
            row[j] = (uint32_t)t;
        }
    }
    free(len_of);
    return (int)ld;
}
