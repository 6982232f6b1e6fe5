/* f4la_oracle.c -- plain-C restatement of neogb's exact sparse echelon form
 * over GF(p).  TEST INFRASTRUCTURE ONLY (see f4la_oracle.h).
 *
 * The algorithm is the reference's, restated on the flat matrix format:
 *   1. pivs[c] = known pivot row with lead column c   (la_ff_32.c:2614-2628)
 *   2. every lower row is scattered into a dense int64 accumulator, reduced
 *      left to right by the pivots, normalised and entered as a new pivot
 *                                                     (la_ff_32.c:2637-2698)
 *   3. the new pivots are inter-reduced from the highest column down and
 *      emitted by decreasing lead column              (la_ff_32.c:2732-2762)
 * Arithmetic: int64 accumulators; 32-bit fields use the subtract-and-fix-up
 * scheme of reduce_dense_row_by_known_pivots_sparse_31_bit
 * (la_ff_32.c:1013-1016, 1203-1210); 16/8-bit fields the add-only scheme of
 * la_ff_16.c:392-400 / la_ff_8.c.  Single threaded, row index order, which is
 * the reference at -t 1.
 */
#include <stdlib.h>
#include <string.h>
#include "f4la_oracle.h"

typedef struct orow {
    uint32_t  len;
    uint32_t  src;      /* input lower row this row descends from */
    uint32_t *cols;     /* owned unless `borrowed`                */
    uint32_t *cf;       /* coefficients widened to 32 bit          */
    int       borrowed;
} orow;

/* reference tools.h:95-122 (mod_p_inverse_32) */
uint32_t f4la_oracle_mod_inverse(uint32_t val, uint32_t p)
{
    int64_t a = p, b = (int64_t)val % (int64_t)p, c = 1, d = 0, e, f;
    while (b != 0) {
        f = b; e = a / f; b = a - e * f; a = f;
        f = c; c = d - e * f; d = f;
    }
    if (d < 0) d += p;
    return (uint32_t)d;
}

static uint32_t cf_at(const f4la_flat_matrix *m, uint64_t idx)
{
    switch (m->cf_bytes) {
        case 1:  return ((const uint8_t  *)m->cf)[idx];
        case 2:  return ((const uint16_t *)m->cf)[idx];
        default: return ((const uint32_t *)m->cf)[idx];
    }
}

static void orow_free(orow *r)
{
    if (!r) return;
    if (!r->borrowed) { free(r->cols); }
    free(r->cf);
    free(r);
}

static orow *orow_from_input(const f4la_flat_matrix *m, uint32_t r)
{
    orow *o = malloc(sizeof(orow));
    const uint64_t a = m->row_ptr[r];
    o->len = (uint32_t)(m->row_ptr[r + 1] - a);
    o->src = r >= m->nru ? r - m->nru : 0;
    o->cols = (uint32_t *)(m->cols + a);
    o->borrowed = 1;
    o->cf = malloc((o->len ? o->len : 1) * sizeof(uint32_t));
    const uint64_t c0 = m->cf_ptr[m->row_cf[r]];
    for (uint32_t k = 0; k < o->len; ++k) o->cf[k] = cf_at(m, c0 + k);
    return o;
}

/* K1: reference la_ff_32.c:965-1242 (31-bit) / la_ff_16.c:97-412.
 * Reduces dr[start..nc) by pivs; returns the surviving entries in columns
 * >= first_out as a fresh row, or NULL when the row vanished. */
static orow *reduce_dense_row(int64_t *dr, orow *const *pivs, uint32_t start,
                              uint32_t nc, uint32_t first_out, uint32_t p,
                              int wide, uint32_t src, f4la_oracle_counters *cnt)
{
    const int64_t mod = p, mod2 = (int64_t)p * p;
    uint32_t k = 0;
    for (uint32_t i = start; i < nc; ++i) {
        if (dr[i] != 0) dr[i] %= mod;
        if (dr[i] == 0) continue;
        const orow *piv = pivs[i];
        if (!piv) { k++; continue; }
        if (wide) {
            const int64_t mul = dr[i];
            for (uint32_t j = 0; j < piv->len; ++j) {
                int64_t *d = &dr[piv->cols[j]];
                *d -= mul * (int64_t)piv->cf[j];
                *d += (*d >> 63) & mod2;
            }
        } else {
            const int64_t mul = mod - dr[i];
            for (uint32_t j = 0; j < piv->len; ++j)
                dr[piv->cols[j]] += mul * (int64_t)piv->cf[j];
        }
        dr[i] = 0;
        if (cnt) { cnt->units += piv->len; cnt->pivot_applications++; }
    }
    if (k == 0) return NULL;
    orow *o = malloc(sizeof(orow));
    o->len = k; o->src = src; o->borrowed = 0;
    o->cols = malloc(k * sizeof(uint32_t));
    o->cf   = malloc(k * sizeof(uint32_t));
    uint32_t j = 0;
    for (uint32_t i = first_out; i < nc; ++i)
        if (dr[i] != 0) { o->cols[j] = i; o->cf[j] = (uint32_t)dr[i]; j++; }
    o->len = j;
    return o;
}

/* K2: reference la_ff_32.c:231-255 */
static void normalize_row(orow *r, uint32_t p)
{
    if (r->cf[0] == 1) return;
    const uint64_t inv = f4la_oracle_mod_inverse(r->cf[0], p);
    for (uint32_t i = 0; i < r->len; ++i) r->cf[i] = (uint32_t)(((uint64_t)r->cf[i] * inv) % p);
    r->cf[0] = 1;
}

static void scatter(int64_t *dr, uint32_t nc, const orow *r)
{
    memset(dr, 0, (size_t)nc * sizeof(int64_t));
    for (uint32_t j = 0; j < r->len; ++j) dr[r->cols[j]] = (int64_t)r->cf[j];
}

static void emit(f4la_flat_result *out, orow *const *rows, uint32_t n, uint32_t cf_bytes)
{
    memset(out, 0, sizeof(*out));
    out->nrows = n; out->cf_bytes = cf_bytes;
    out->row_ptr = malloc(((size_t)n + 1) * sizeof(uint64_t));
    out->src_row = malloc(((size_t)n ? n : 1) * sizeof(uint32_t));
    uint64_t nnz = 0;
    for (uint32_t i = 0; i < n; ++i) { out->row_ptr[i] = nnz; if (rows[i]) nnz += rows[i]->len; }
    out->row_ptr[n] = nnz;
    out->cols = malloc((nnz ? nnz : 1) * sizeof(uint32_t));
    out->cf   = malloc((nnz ? nnz : 1) * (size_t)cf_bytes);
    for (uint32_t i = 0; i < n; ++i) {
        const orow *r = rows[i];
        out->src_row[i] = r ? r->src : 0;
        if (!r) continue;
        const uint64_t a = out->row_ptr[i];
        memcpy(out->cols + a, r->cols, (size_t)r->len * sizeof(uint32_t));
        for (uint32_t k = 0; k < r->len; ++k) {
            switch (cf_bytes) {
                case 1:  ((uint8_t  *)out->cf)[a + k] = (uint8_t)r->cf[k]; break;
                case 2:  ((uint16_t *)out->cf)[a + k] = (uint16_t)r->cf[k]; break;
                default: ((uint32_t *)out->cf)[a + k] = r->cf[k]; break;
            }
        }
    }
}

int f4la_oracle_reduce(const f4la_flat_matrix *in, f4la_flat_result *out,
                       f4la_oracle_counters *cnt)
{
    const uint32_t nc = in->ncl + in->ncr, nru = in->nru, nrl = in->nrl, p = in->fc;
    const int wide = in->cf_bytes == 4;
    const int nf = (in->flags & F4LA_FLAG_NORMALFORM) != 0;
    const int by_lead = (in->flags & F4LA_FLAG_PIVOT_BY_LEAD) != 0;
    if (cnt) memset(cnt, 0, sizeof(*cnt));
    if (p < 2 || (in->cf_bytes != 1 && in->cf_bytes != 2 && in->cf_bytes != 4)) return F4LA_ERR_ARG;

    orow **pivs = calloc(nc ? nc : 1, sizeof(orow *));
    for (uint32_t i = 0; i < nru; ++i) {                 /* la_ff_32.c:2614-2628 */
        orow *r = orow_from_input(in, i);
        const uint32_t lead = by_lead ? r->cols[0] : i;
        if (lead >= nc || pivs[lead]) { orow_free(r); free(pivs); return F4LA_ERR_ARG; }
        pivs[lead] = r;
    }
    int64_t *dr = malloc(((size_t)nc ? nc : 1) * sizeof(int64_t));
    orow **tr = calloc(nrl ? nrl : 1, sizeof(orow *));

    for (uint32_t i = 0; i < nrl; ++i) {                 /* la_ff_32.c:2640-2698 */
        orow *row = orow_from_input(in, nru + i);
        scatter(dr, nc, row);
        const uint32_t sc = nf ? 0 : row->cols[0];       /* la_ff_32.c:2663-2665 */
        orow_free(row);
        /* rows enter pivs only in echelon / final-step mode, so the output
         * starts at ncl exactly as in the reference's gather loop (:1226) */
        orow *np = reduce_dense_row(dr, pivs, sc, nc, in->ncl, p, wide, i, cnt);
        tr[i] = np;
        if (!np || nf) continue;
        normalize_row(np, p);                            /* la_ff_32.c:2689-2692 */
        pivs[np->cols[0]] = np;                          /* single thread: CAS always wins */
    }

    /* ownership: rows entered into pivs[] belong to pivs[]; in nf mode the
     * result rows are never entered and belong to tr[] */
    if (nf || by_lead) {                                 /* la_ff_32.c:2764-2766 */
        emit(out, tr, nrl, in->cf_bytes);
        if (nf) for (uint32_t i = 0; i < nrl; ++i) orow_free(tr[i]);
    } else {
        /* inter-reduction from the right, la_ff_32.c:2732-2762 */
        orow **res = calloc(in->ncr ? in->ncr : 1, sizeof(orow *));
        uint32_t npiv = 0;
        for (uint32_t i = 0; i < in->ncr; ++i) {
            const uint32_t k = nc - 1 - i;
            orow *piv = pivs[k];
            if (!piv) continue;
            scatter(dr, nc, piv);
            const uint32_t src = piv->src, sc = piv->cols[0];
            pivs[k] = NULL;
            orow_free(piv);
            orow *red = reduce_dense_row(dr, pivs, sc, nc, in->ncl, p, wide, src, cnt);
            pivs[k] = red;
            res[npiv++] = red;
        }
        emit(out, res, npiv, in->cf_bytes);
        free(res);
    }
    for (uint32_t c = 0; c < nc; ++c) orow_free(pivs[c]);
    free(tr); free(pivs); free(dr);
    return F4LA_OK;
}

void f4la_oracle_free_result(f4la_flat_result *res)
{
    if (!res) return;
    free(res->row_ptr); free(res->cols); free(res->cf); free(res->src_row);
    memset(res, 0, sizeof(*res));
}
