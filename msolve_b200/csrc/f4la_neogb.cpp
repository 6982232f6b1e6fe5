/* f4la_neogb.cpp -- neogb-level drop-in entry points (include/f4la_b200_neogb.h).
 *
 * Mirrors the wrapper + body pair exact_sparse_linear_algebra_ff_32 /
 * exact_sparse_reduced_echelon_form_ff_32 of the reference
 * (src/neogb/la_ff_32.c:3961-3990 and :2595-2770; 16/8-bit twins in
 * la_ff_16.c:2011/1046, la_ff_8.c:2203/1319) on the host side:
 *   - same inputs consumed (every tr[i] and rr[i] is free()d, :2666/:2718-2721),
 *   - same outputs produced (individually malloc()ed rows with the 6-word
 *     prelude and individually malloc()ed coefficient arrays, :1223-1240;
 *     rows by decreasing lead column, :2732-2762),
 *   - same side channels (st->np, mat->np/nr/sz, num_zerored, la_ctime,
 *     la_rtime, the info_level > 1 print, :3968-3989).
 * The arithmetic happens on the GPU (f4la_device.cu); there is no CPU path.
 */
#include <sys/time.h>
#include <time.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <emmintrin.h>

#include "f4la_b200.h"
#include "f4la_b200_neogb.h"

namespace {

f4la_neogb_layout L;
bool g_configured = false;
void (*g_construct_trace)(void *, void *) = nullptr;
void (*g_free_basis_elements)(void *) = nullptr;

enum class Entry { Normal, Application, Trace };

/* One recorded linear-algebra call of an application-phase run (see f4la_b200_neogb.h). */
struct TapeCall {
    uint32_t flags = 0, nru = 0, nrl = 0, ncl = 0, ncr = 0, cf_bytes = 4, bs_ld = 0;
    bool final_step = false;
    std::vector<uint64_t> row_ptr, cf_ptr;
    std::vector<uint32_t> cols, row_cf, slot_bindex;      /* slot_bindex[s]: basis element behind coefficient array s */
    std::vector<uint64_t> out_row_ptr;                    /* support of the result rows on the recorded prime */
    std::vector<uint32_t> out_cols;
};
} // namespace

struct f4la_tape {
    std::vector<TapeCall> calls;
    uint32_t ninit = 0;
    std::vector<uint32_t> init_len;
    uint64_t residue_words = 0;
    bool valid = true;
    /* per context: the structures resident in HBM (uploaded at the first replay on that context) */
    std::mutex mu;
    static constexpr int kInflight = 16;             /* replays enqueued per context before the host has to wait */
    struct PerCtx {
        std::vector<f4la_resident *> res;
        void *pool = nullptr;                         /* pinned, largest call (host-assembled replay) */
        /* device-resident replay (f4la_b200_tape_replay_enqueue): */
        int async_state = 0;                          /* 0 not built, 1 ready, -1 not possible for this tape */
        bool warmed = false;                          /* one host-assembled replay ran here: layout caches are filled */
        std::vector<f4la_replay_call> plan;           /* one per recorded call, device pointers */
        std::vector<void *> dev;                      /* everything allocated on the device for the plan */
        std::vector<uint64_t> init_off;               /* layout of the initial elements the plan was built for */
        uint32_t *arena = nullptr, *d_status = nullptr;
        uint64_t arena_words = 0, residue_off = 0;
        uint32_t *h_init = nullptr, *h_status = nullptr;   /* pinned: kInflight slots of initial coefficients / status words */
        int inflight = 0;
        std::vector<int> done_rc;                     /* results of replays that ran synchronously, in enqueue order ... */
        std::vector<int> order;                       /* ... 1 = next result comes from done_rc, 0 = from an in-flight slot */
    };
    std::map<void *, PerCtx> resident;
};

namespace {

struct ThreadState {
    f4la_ctx *ctx = nullptr;
    int device = -1;
    /* pinned staging for the flat matrix (grow-only) */
    void *buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[5] = {0, 0, 0, 0, 0};
    f4la_neogb_totals tot = {};
    f4la_tape *tape = nullptr;          /* recording (f4la_b200_neogb_tape_begin) */
    ~ThreadState()
    {
        for (auto &b : buf) if (b) f4la_b200_pinned_free(b);
        delete tape;
        if (ctx) f4la_b200_destroy(ctx);
    }
};

thread_local ThreadState T;

/* Gather copy into a page-locked staging piece.  The piece is written once and then read by the copy engine, never by
 * this CPU: non-temporal stores skip the read-for-ownership of the destination lines (a third of the DRAM traffic of a
 * plain memcpy at these sizes).  F4LA_B200_NT_COPY=0 restores memcpy. */
static const bool g_nt_copy = !(getenv("F4LA_B200_NT_COPY") && atoi(getenv("F4LA_B200_NT_COPY")) == 0);

static inline void copy_to_staging(void *dst, const void *src, size_t n)
{
    unsigned char *d = (unsigned char *)dst;
    const unsigned char *s = (const unsigned char *)src;
    if (!g_nt_copy || n < 512) { memcpy(d, s, n); return; }
    const size_t head = (64 - ((uintptr_t)d & 63)) & 63;
    memcpy(d, s, head);
    d += head; s += head; n -= head;
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
        const __m128i a0 = _mm_loadu_si128((const __m128i *)(s + 64 * i));
        const __m128i a1 = _mm_loadu_si128((const __m128i *)(s + 64 * i + 16));
        const __m128i a2 = _mm_loadu_si128((const __m128i *)(s + 64 * i + 32));
        const __m128i a3 = _mm_loadu_si128((const __m128i *)(s + 64 * i + 48));
        _mm_stream_si128((__m128i *)(d + 64 * i), a0);
        _mm_stream_si128((__m128i *)(d + 64 * i + 16), a1);
        _mm_stream_si128((__m128i *)(d + 64 * i + 32), a2);
        _mm_stream_si128((__m128i *)(d + 64 * i + 48), a3);
    }
    memcpy(d + 64 * blocks, s + 64 * blocks, n - 64 * blocks);
    _mm_sfence();
}

template <typename V> V &fld(void *base, uint32_t off) { return *reinterpret_cast<V *>((char *)base + off); }
template <typename V> const V &fld(const void *base, uint32_t off) { return *reinterpret_cast<const V *>((const char *)base + off); }

double wall_s() { struct timeval t; gettimeofday(&t, nullptr); return t.tv_sec + 1e-6 * t.tv_usec; }
double cpu_s() { return (double)clock() / CLOCKS_PER_SEC; }
double now_ms()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
}

[[noreturn]] void die(const char *msg)
{
    fprintf(stderr, "[f4la_b200] fatal: %s\n", msg);
    exit(1);                       /* same style as the reference, f4.c:828-830 */
}

void *stage(int k, size_t bytes)
{
    if (bytes > T.cap[k]) {
        if (T.buf[k]) f4la_b200_pinned_free(T.buf[k]);
        size_t want = 2 * bytes + ((size_t)1 << 20);      /* geometric growth: few re-pinnings per run */
        T.buf[k] = f4la_b200_pinned_alloc(want);
        if (!T.buf[k]) die("pinned host allocation failed");
        T.cap[k] = want;
    }
    return T.buf[k];
}

f4la_ctx *ctx()
{
    if (!T.ctx) {
        if (T.device < 0) {
            const char *e = getenv("F4LA_B200_DEVICE");
            T.device = e ? atoi(e) : 0;
        }
        T.ctx = f4la_b200_create(T.device);
        if (!T.ctx) die("no usable B200 device for the linear algebra backend (there is no CPU fallback)");
    }
    return T.ctx;
}

/* Echelon form of mat->tr modulo mat->rr; the body shared by the four matrix-level pointers.
 * Returns 1 for an unlucky prime in Entry::Application (la_ff_32.c:3215-3217), else 0. */
int echelon(void *mat, const void *tbr, const void *bs, void *st, Entry entry = Entry::Normal, void *trace = nullptr)
{
    if (!g_configured) die("f4la_b200_neogb_configure() was not called");
    const double ct0 = cpu_s(), rt0 = wall_s();
    const double t_begin = now_ms();

    const int32_t ff_bits = fld<int32_t>(st, L.md_ff_bits);
    const uint32_t fc = fld<uint32_t>(st, L.md_fc);
    const int32_t trace_level = fld<int32_t>(st, L.md_trace_level);
    const uint32_t w = ff_bits == 8 ? 1 : ff_bits == 16 ? 2 : 4;
    if (ff_bits != 8 && ff_bits != 16 && ff_bits != 32) die("unsupported field (ff_bits must be 8, 16 or 32)");
    const bool nf_mode = fld<int32_t>(st, L.md_nf) > 0;
    const bool final_step = fld<int32_t>(st, L.md_in_final_reduction_step) != 0;
    /* the reference records no trace in the final reduction step (la_ff_32.c:2713) */
    /* trace_linear_algebra always learns into the trace it is given (la_ff_32.c:3004-3133);
     * linear_algebra learns when the F4 driver runs in LEARN_TRACER mode (la_ff_32.c:2669, 2713) */
    const bool learn = entry == Entry::Trace || (trace_level == 1 /* LEARN_TRACER */ && !final_step && !nf_mode);
    if (entry != Entry::Normal && (nf_mode || final_step))
        die("application / trace linear algebra in nf or final-reduction mode is not defined by the reference");
    if (learn && !g_construct_trace)
        die("LEARN_TRACER needs f4la_b200_neogb_set_construct_trace() (see INTEGRATION.md)");

    const uint32_t nru = fld<uint32_t>(mat, L.mat_nru), nrl = fld<uint32_t>(mat, L.mat_nrl);
    const uint32_t ncl = fld<uint32_t>(mat, L.mat_ncl), ncr = fld<uint32_t>(mat, L.mat_ncr);
    uint32_t **rr = fld<uint32_t **>(mat, L.mat_rr);
    uint32_t **tr = fld<uint32_t **>(mat, L.mat_tr);
    const uint32_t n = nru + nrl;
    const uint32_t cf_off = w == 1 ? L.mat_cf_8 : w == 2 ? L.mat_cf_16 : L.mat_cf_32;
    const uint32_t bcf_off = w == 1 ? L.bs_cf_8 : w == 2 ? L.bs_cf_16 : L.bs_cf_32;

    /* la_ff_32.c:3975: room for one coefficient array pointer per row */
    void **&mcf = fld<void **>(mat, cf_off);
    mcf = (void **)realloc(mcf, (size_t)fld<uint32_t>(mat, L.mat_nr) * sizeof(void *));

    /* ---- flatten ---- */
    /* The small arrays (row offsets, coefficient-array ids and offsets) go through small page-locked buffers.  The
     * two big ones -- the column indices (4 B per non-zero, ~1 GB per katsura-12 run) and the coefficient pool --
     * are gathered from the rows' individual malloc blocks straight into the library's ring of page-locked pieces
     * (pinned once, never re-pinned) and sent piece by piece, so the gather of piece k+1 by st->nthrds host threads
     * (like the reference's own row loops, e.g. convert.c:365-397) overlaps the host->device copy of piece k. */
    static const bool dbg = getenv("F4LA_B200_DEBUG") && atoi(getenv("F4LA_B200_DEBUG")) != 0;
    double dbg_t[6] = {0, 0, 0, 0, 0, 0}, dbg_t0 = now_ms();
    uint64_t *row_ptr = (uint64_t *)stage(0, ((size_t)n + 1) * 8);
    uint32_t *row_cf = (uint32_t *)stage(2, ((size_t)n + 1) * 4);
    uint64_t nnz = 0;
    for (uint32_t r = 0; r < n; ++r) {
        const uint32_t *row = r < nru ? rr[r] : tr[r - nru];
        row_ptr[r] = nnz; nnz += row[L.row_length];
    }
    row_ptr[n] = nnz;
    void *const *bs_cf = fld<void *const *>(bs, bcf_off);
    void *const *tbr_cf = fld<void *const *>(tbr, bcf_off);
    const uint32_t bs_ld = fld<uint32_t>(bs, L.bs_ld), tbr_ld = fld<uint32_t>(tbr, L.bs_ld);
    std::vector<int32_t> map_bs(bs_ld + 1, -1), map_tbr_v;
    if (tbr != bs) map_tbr_v.assign(tbr_ld + 1, -1);
    std::vector<int32_t> &map_tbr = (tbr == bs) ? map_bs : map_tbr_v;
    std::vector<const void *> cfsrc; cfsrc.reserve(1024);
    std::vector<uint64_t> cfp; cfp.reserve(1025);
    std::vector<uint32_t> src_bindex(nrl), src_mult(nrl);
    uint64_t ncoef = 0;
    const int nthr = fld<int32_t>(st, L.md_nthrds) > 1 ? fld<int32_t>(st, L.md_nthrds) : 1;
    f4la_ctx *cx = ctx();
    {   /* the host side of the call uses the threads this F4 run was given (st->nthrds), not OpenMP's default */
        static const bool blocking = getenv("F4LA_B200_BLOCKING_SYNC") && atoi(getenv("F4LA_B200_BLOCKING_SYNC")) != 0;
        f4la_b200_set_host_mode(cx, nthr, blocking);
    }
    dbg_t[0] = now_ms() - dbg_t0;
    for (uint32_t r = 0; r < n; ++r) {
        const uint32_t *row = r < nru ? rr[r] : tr[r - nru];
        const uint32_t len = row[L.row_length];
        std::vector<int32_t> &map = r < nru ? map_bs : map_tbr;
        const uint32_t ci = row[L.row_coeffs];
        if ((size_t)ci >= map.size()) die("row refers to a coefficient array beyond the basis load");
        if (map[ci] < 0) {
            map[ci] = (int32_t)cfsrc.size();
            cfp.push_back(ncoef);
            cfsrc.push_back((r < nru ? bs_cf : tbr_cf)[ci]);
            ncoef += len;
        }
        row_cf[r] = (uint32_t)map[ci];
        if (r >= nru) { src_bindex[r - nru] = row[L.row_bindex]; src_mult[r - nru] = row[L.row_mult]; }
    }
    cfp.push_back(ncoef);
    const uint32_t ncf = (uint32_t)cfsrc.size();
    uint64_t *cf_ptr = (uint64_t *)stage(3, ((size_t)ncf + 1) * 8);
    memcpy(cf_ptr, cfp.data(), ((size_t)ncf + 1) * 8);
    dbg_t[1] = now_ms() - dbg_t0;
    {
        /* column indices: elements [a, b) of the concatenation of all rows per piece */
        uint32_t r_lo = 0;
        for (uint64_t a = 0; a < nnz;) {
            void *piece = nullptr; size_t cap = 0;
            const double ta = now_ms();
            if (f4la_b200_stage_acquire(cx, &piece, &cap) != F4LA_OK) die(f4la_b200_last_error(cx));
            dbg_t[2] += now_ms() - ta;
            const uint64_t b = std::min<uint64_t>(nnz, a + cap / 4);
            while (row_ptr[r_lo + 1] <= a) ++r_lo;                           /* first row reaching into [a, b) */
            uint32_t r_hi = r_lo;
            while (r_hi < n && row_ptr[r_hi] < b) ++r_hi;
            uint32_t *dst = (uint32_t *)piece;
#pragma omp parallel for num_threads(nthr) schedule(static) if (r_hi - r_lo > 64)
            for (uint32_t r = r_lo; r < r_hi; ++r) {
                const uint32_t *row = r < nru ? rr[r] : tr[r - nru];
                const uint64_t lo = std::max<uint64_t>(row_ptr[r], a), hi = std::min<uint64_t>(row_ptr[r + 1], b);
                if (hi > lo) copy_to_staging(dst + (lo - a), row + L.row_offset + (lo - row_ptr[r]), (size_t)(hi - lo) * 4);
            }
            const double tc = now_ms();
            if (f4la_b200_stage_commit(cx, F4LA_STAGE_COLS, 4, a, b - a, nnz) != F4LA_OK) die(f4la_b200_last_error(cx));
            dbg_t[3] += now_ms() - tc;
            a = b;
        }
        dbg_t[4] = now_ms() - dbg_t0;
        /* coefficient pool: arrays [cfp[c], cfp[c+1]) likewise */
        uint32_t c_lo = 0;
        for (uint64_t a = 0; a < ncoef;) {
            void *piece = nullptr; size_t cap = 0;
            if (f4la_b200_stage_acquire(cx, &piece, &cap) != F4LA_OK) die(f4la_b200_last_error(cx));
            const uint64_t b = std::min<uint64_t>(ncoef, a + cap / w);
            while (cfp[c_lo + 1] <= a) ++c_lo;
            uint32_t c_hi = c_lo;
            while (c_hi < ncf && cfp[c_hi] < b) ++c_hi;
            unsigned char *dst = (unsigned char *)piece;
#pragma omp parallel for num_threads(nthr) schedule(static) if (c_hi - c_lo > 64)
            for (uint32_t c = c_lo; c < c_hi; ++c) {
                const uint64_t lo = std::max<uint64_t>(cfp[c], a), hi = std::min<uint64_t>(cfp[c + 1], b);
                if (hi > lo) copy_to_staging(dst + (lo - a) * w, (const unsigned char *)cfsrc[c] + (lo - cfp[c]) * w, (size_t)(hi - lo) * w);
            }
            if (f4la_b200_stage_commit(cx, F4LA_STAGE_CF, w, a, b - a, ncoef) != F4LA_OK) die(f4la_b200_last_error(cx));
            a = b;
        }
    }

    if (dbg)
        fprintf(stderr, "[f4la debug] flatten %u rows %llu nnz: small staging %.2f, dedupe %.2f, columns %.2f (acquire %.2f, "
                "commit %.2f), coefficients %.2f ms\n", n, (unsigned long long)nnz, dbg_t[0], dbg_t[1] - dbg_t[0],
                dbg_t[4] - dbg_t[1], dbg_t[2], dbg_t[3], now_ms() - dbg_t0 - dbg_t[4]);
    f4la_flat_matrix fm;
    memset(&fm, 0, sizeof(fm));
    fm.fc = fc; fm.cf_bytes = w; fm.nru = nru; fm.nrl = nrl; fm.ncl = ncl; fm.ncr = ncr; fm.ncf = ncf;
    fm.flags = nf_mode ? F4LA_FLAG_NORMALFORM : final_step ? F4LA_FLAG_PIVOT_BY_LEAD : F4LA_FLAG_NONE;
    if (learn) fm.flags |= F4LA_FLAG_WANT_RBA;
    if (nnz) fm.flags |= F4LA_FLAG_COLS_STAGED;
    if (ncoef) fm.flags |= F4LA_FLAG_CF_STAGED;
    /* -l 42 / 43 / 44: the reference's probabilistic variants (io.c:885-957); no tracer runs with
     * them (f4.c:365), the device decides whether random combinations pay off */
    const int32_t laopt = fld<int32_t>(st, L.md_laopt);
    if ((laopt == 42 || laopt == 43 || laopt == 44) && entry == Entry::Normal && !learn && !nf_mode && !final_step &&
        trace_level == 0)
        fm.flags |= F4LA_FLAG_PROBABILISTIC;
    fm.row_ptr = row_ptr; fm.cols = nullptr; fm.row_cf = row_cf; fm.cf_ptr = cf_ptr; fm.cf = nullptr;
    if (T.tape) {
        /* tape of an application-phase run: structure, basis element behind every coefficient array */
        f4la_tape &tp = *T.tape;
        if (tbr != bs || nf_mode || learn || entry != Entry::Normal) tp.valid = false;
        tp.calls.emplace_back();
        TapeCall &tc = tp.calls.back();
        tc.flags = fm.flags & ~(F4LA_FLAG_COLS_STAGED | F4LA_FLAG_CF_STAGED);
        tc.nru = nru; tc.nrl = nrl; tc.ncl = ncl; tc.ncr = ncr; tc.cf_bytes = w; tc.bs_ld = bs_ld; tc.final_step = final_step;
        tc.row_ptr.assign(row_ptr, row_ptr + n + 1);
        tc.row_cf.assign(row_cf, row_cf + n);
        tc.cf_ptr.assign(cfp.begin(), cfp.end());
        tc.cols.resize(nnz);
        for (uint32_t r = 0; r < n; ++r) {
            const uint32_t *row = r < nru ? rr[r] : tr[r - nru];
            memcpy(tc.cols.data() + row_ptr[r], row + L.row_offset, (size_t)row[L.row_length] * 4);
        }
        tc.slot_bindex.assign(ncf, 0);
        for (uint32_t ci = 0; ci < map_bs.size(); ++ci) if (map_bs[ci] >= 0) tc.slot_bindex[map_bs[ci]] = ci;
    }
    const double t_flat = now_ms();

    /* ---- device ---- */
    f4la_flat_result res;
    f4la_stats stats;
    /* The LA consumes its inputs (la_ff_32.c:2666, 2718-2721).  Everything the device needs has
     * been copied to the staging buffers, so a helper thread frees the ~10^4-10^5 input rows
     * (tens of milliseconds of free()) while the GPU works.  Not while learning a trace:
     * construct_trace() still reads the rows. */
    std::thread reaper;
    bool reaped = false;
    if (!learn && (uint64_t)n > 2048) {
        reaper = std::thread([=]() {
            for (uint32_t i = 0; i < nrl; ++i) free(tr[i]);
            for (uint32_t i = 0; i < nru; ++i) free(rr[i]);
        });
        reaped = true;
    }
    const int rc = f4la_b200_reduce(cx, &fm, &res, &stats);
    if (reaped) {
        reaper.join();
        for (uint32_t i = 0; i < nrl; ++i) tr[i] = nullptr;
        for (uint32_t i = 0; i < nru; ++i) rr[i] = nullptr;
    }
    if (rc != F4LA_OK) die(f4la_b200_last_error(cx));
    if (T.tape) {
        TapeCall &tc = T.tape->calls.back();
        tc.out_row_ptr.assign(res.row_ptr, res.row_ptr + res.nrows + 1);
        tc.out_cols.assign(res.cols, res.cols + res.row_ptr[res.nrows]);
    }
    const double t_dev = now_ms();

    /* ---- learning: hand the reference's trace builder what it reads -------------
     * construct_trace() (tools.c:63-183) looks at mat->tr[i] == NULL (row vanished),
     * tr[i][BINDEX/MULT], rr[i][BINDEX/MULT] and the reducer bit arrays mat->rba[i].
     * A lower row is kept iff it became a pivot row of the trailing block: these
     * rows are linearly independent modulo the known pivots and span the new
     * elements, which is all the application phase relies on (any maximal
     * independent subset gives the same reduced rows). */
    if (learn) {
        std::vector<char> kept(nrl, 0);
        for (uint32_t t = 0; t < res.nrows; ++t) {
            if (res.src_row[t] >= nrl) die("tracing needs a donor row for every new pivot");
            kept[res.src_row[t]] = 1;
        }
        uint32_t **rba = fld<uint32_t **>(mat, L.mat_rba);
        const uint32_t words = res.rba_words;
        for (uint32_t i = 0; i < nrl; ++i) {
            if (rba && rba[i] && res.rba) memcpy(rba[i], res.rba + (size_t)i * words, (size_t)words * 4);
            if (!kept[i]) { free(tr[i]); tr[i] = nullptr; }
        }
        g_construct_trace(entry == Entry::Trace ? trace : fld<void *>(st, L.md_tr), mat);
    }

    /* ---- consume the inputs like the reference (la_ff_32.c:2666, 2718-2721) ---- */
    /* serial on purpose: the rows were all allocated by the F4 driver's main thread, freeing them
     * from many threads only contends on that arena's lock (measured slower than one thread) */
    if (!reaped) {
        for (uint32_t i = 0; i < nrl; ++i) { free(tr[i]); tr[i] = nullptr; }
        for (uint32_t i = 0; i < nru; ++i) { free(rr[i]); rr[i] = nullptr; }
    }

    uint32_t np = res.nrows;
    int bad_prime = 0;
    if (nf_mode || final_step) {
        /* rows are returned in place, one per input row, NULL when it vanished;
         * np = nrl (la_ff_32.c:2672-2684, 2764-2766) */
        for (uint32_t t = 0; t < nrl; ++t) {
            const uint64_t a = res.row_ptr[t];
            const uint32_t len = (uint32_t)(res.row_ptr[t + 1] - a);
            if (len == 0) { tr[t] = nullptr; continue; }
            uint32_t *row = (uint32_t *)malloc(((size_t)len + L.row_offset) * sizeof(uint32_t));
            void *cf = malloc((size_t)len * w);
            memcpy(row + L.row_offset, res.cols + a, (size_t)len * 4);
            memcpy(cf, (unsigned char *)res.cf + a * w, (size_t)len * w);
            row[L.row_deg] = 0;
            row[L.row_bindex] = src_bindex[t];
            row[L.row_mult] = src_mult[t];
            row[L.row_coeffs] = t;
            row[L.row_preloop] = len % L.unroll;
            row[L.row_length] = len;
            mcf[t] = cf;
            tr[t] = row;
        }
        f4la_b200_free_result(&res);
        np = nrl;
        fld<uint32_t>(st, L.md_np) = np;
        fld<uint32_t>(mat, L.mat_np) = np;
        fld<uint32_t>(mat, L.mat_nr) = np;
        fld<uint32_t>(mat, L.mat_sz) = np;
    } else if ((trace_level == 2 /* APPLY_TRACER */ || entry == Entry::Application) && np != nrl) {
        /* a row reduced to zero while applying the tracer: bad prime
         * (la_ff_32.c:2679-2683, 2700-2709) */
        np = 0;
        bad_prime = 1;
        if (fld<int32_t>(st, L.md_info_level) > 0)
            fprintf(stderr, "Zero reduction while applying tracer, bad prime.\n");
        fld<uint32_t>(mat, L.mat_np) = 0;
        f4la_b200_free_result(&res);
    } else {
        /* ---- materialise the new pivots as neogb rows (la_ff_32.c:1223-1240) ---- */
        tr = (uint32_t **)realloc(tr, (size_t)(np ? np : 1) * sizeof(uint32_t *));
        fld<uint32_t **>(mat, L.mat_tr) = tr;
        for (uint32_t t = 0; t < np; ++t) {
            const uint64_t a = res.row_ptr[t];
            const uint32_t len = (uint32_t)(res.row_ptr[t + 1] - a);
            uint32_t *row = (uint32_t *)malloc(((size_t)len + L.row_offset) * sizeof(uint32_t));
            void *cf = malloc((size_t)(len ? len : 1) * w);
            memcpy(row + L.row_offset, res.cols + a, (size_t)len * 4);
            memcpy(cf, (unsigned char *)res.cf + a * w, (size_t)len * w);
            /* BINDEX / MULT of the donor row (la_ff_32.c:1234-1235); rows born from random combinations have
             * none -- the reference's 16-bit reducer leaves the two words unset as well (la_ff_16.c:407-409),
             * only the tracer reads them and it never runs in those modes */
            const uint32_t src = res.src_row[t];
            row[L.row_deg] = 0;
            row[L.row_bindex] = src < nrl ? src_bindex[src] : 0;
            row[L.row_mult] = src < nrl ? src_mult[src] : 0;
            row[L.row_coeffs] = t;
            row[L.row_preloop] = len % L.unroll;
            row[L.row_length] = len;
            mcf[t] = cf;
            tr[t] = row;
        }
        f4la_b200_free_result(&res);
        fld<uint32_t>(st, L.md_np) = np;
        fld<uint32_t>(mat, L.mat_np) = np;
        fld<uint32_t>(mat, L.mat_nr) = np;
        fld<uint32_t>(mat, L.mat_sz) = np;
    }
    const double t_end = now_ms();

    /* ---- side channels of the wrapper (la_ff_32.c:3978-3989) ---- */
    fld<double>(st, L.md_la_ctime) += cpu_s() - ct0;
    fld<double>(st, L.md_la_rtime) += wall_s() - rt0;
    fld<int64_t>(st, L.md_num_zerored) += (int64_t)nrl - (int64_t)np;
    /* application_nr_* (la_ff_32.c:1214-1216): per applied pivot its row length / 1000 and one reduction.
     * The device counts exactly these for the known pivots (stats.units); msolve prints them as Gops/sec
     * (msolve.c:2666-2673) */
    fld<double>(st, L.md_application_nr_mult) += (double)stats.units / 1000.0;
    fld<double>(st, L.md_application_nr_add) += (double)stats.units / 1000.0;
    fld<uint64_t>(st, L.md_application_nr_red) += stats.ref_applications;
    if (fld<int32_t>(st, L.md_info_level) > 1) {
        fprintf(stderr, "%9d new %7d zero", (int)np, (int)(nrl - np));
        fflush(stderr);
    }

    f4la_neogb_totals &t = T.tot;
    t.calls++;
    t.ms_total += t_end - t_begin; t.ms_flatten += t_flat - t_begin; t.ms_materialize += t_end - t_dev;
    t.ms_h2d += stats.ms_h2d; t.ms_prep += stats.ms_prep; t.ms_reduce += stats.ms_reduce;
    t.ms_echelon += stats.ms_echelon; t.ms_d2h += stats.ms_d2h;
    t.units += stats.units; t.bytes_h2d += stats.bytes_h2d; t.bytes_d2h += stats.bytes_d2h;
    t.kernel_launches += stats.kernel_launches;
    return entry == Entry::Application ? bad_prime : 0;
}

/* interreduce_matrix_rows_ff_* (la_ff_32.c:4254-4326): every row of mat->rr is a pivot
 * with its own lead column; the rows are reduced by each other (unique result: the reduced
 * row echelon form of their span), returned in mat->tr by INCREASING lead column with
 * COEFFS = lead column (mat->cf_* has one slot per column), mat->np = mat->nr. */
void interreduce(void *mat, void *bs, void *st, int free_basis)
{
    if (!g_configured) die("f4la_b200_neogb_configure() was not called");
    const int32_t ff_bits = fld<int32_t>(st, L.md_ff_bits);
    const uint32_t w = ff_bits == 8 ? 1 : ff_bits == 16 ? 2 : 4;
    if (ff_bits != 8 && ff_bits != 16 && ff_bits != 32) die("unsupported field (ff_bits must be 8, 16 or 32)");
    const uint32_t nrows = fld<uint32_t>(mat, L.mat_nr), ncols = fld<uint32_t>(mat, L.mat_nc);
    uint32_t **rr = fld<uint32_t **>(mat, L.mat_rr);
    const uint32_t cf_off = w == 1 ? L.mat_cf_8 : w == 2 ? L.mat_cf_16 : L.mat_cf_32;
    const uint32_t bcf_off = w == 1 ? L.bs_cf_8 : w == 2 ? L.bs_cf_16 : L.bs_cf_32;
    if (fld<int32_t>(st, L.md_info_level) > 1) fprintf(stderr, "                          ");

    uint32_t **&tr = fld<uint32_t **>(mat, L.mat_tr);
    tr = (uint32_t **)realloc(tr, (size_t)(ncols ? ncols : 1) * sizeof(uint32_t *));
    void **&mcf = fld<void **>(mat, cf_off);
    mcf = (void **)realloc(mcf, (size_t)(ncols ? ncols : 1) * sizeof(void *));
    memset(mcf, 0, (size_t)ncols * sizeof(void *));

    /* flatten: no known pivots, the rows themselves are the lower rows, keyed by lead */
    std::vector<uint64_t> row_ptr(nrows + 1), cf_ptr(nrows + 1);
    uint64_t nnz = 0;
    for (uint32_t r = 0; r < nrows; ++r) { row_ptr[r] = nnz; cf_ptr[r] = nnz; nnz += rr[r][L.row_length]; }
    row_ptr[nrows] = nnz; cf_ptr[nrows] = nnz;
    uint32_t *cols = (uint32_t *)stage(1, (nnz + 1) * 4);
    unsigned char *pool = (unsigned char *)stage(4, (nnz + 1) * w);
    std::vector<uint32_t> row_cf(nrows);
    void *const *bs_cf = fld<void *const *>(bs, bcf_off);
    for (uint32_t r = 0; r < nrows; ++r) {
        const uint32_t len = rr[r][L.row_length];
        memcpy(cols + row_ptr[r], rr[r] + L.row_offset, (size_t)len * 4);
        memcpy(pool + row_ptr[r] * w, bs_cf[rr[r][L.row_coeffs]], (size_t)len * w);   /* not shared here, as :4279 notes */
        row_cf[r] = r;
    }
    f4la_flat_matrix fm;
    memset(&fm, 0, sizeof(fm));
    fm.fc = fld<uint32_t>(st, L.md_fc); fm.cf_bytes = w; fm.nru = 0; fm.nrl = nrows; fm.ncl = 0; fm.ncr = ncols;
    fm.ncf = nrows; fm.flags = F4LA_FLAG_PIVOT_BY_LEAD;
    fm.row_ptr = row_ptr.data(); fm.cols = cols; fm.row_cf = row_cf.data(); fm.cf_ptr = cf_ptr.data(); fm.cf = pool;
    f4la_flat_result res;
    f4la_stats stats;
    if (f4la_b200_reduce(ctx(), &fm, &res, &stats) != F4LA_OK) die(f4la_b200_last_error(ctx()));

    /* position of every row in the output = rank of its lead column */
    std::vector<uint32_t> order(nrows);
    for (uint32_t r = 0; r < nrows; ++r) order[r] = r;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return rr[a][L.row_offset] < rr[b][L.row_offset]; });
    for (uint32_t k = 0; k < nrows; ++k) {
        const uint32_t r = order[k];
        const uint64_t a = res.row_ptr[r];
        const uint32_t len = (uint32_t)(res.row_ptr[r + 1] - a);
        const uint32_t lead = rr[r][L.row_offset];
        uint32_t *row = (uint32_t *)malloc(((size_t)len + L.row_offset) * sizeof(uint32_t));
        void *cf = malloc((size_t)(len ? len : 1) * w);
        memcpy(row + L.row_offset, res.cols + a, (size_t)len * 4);
        memcpy(cf, (unsigned char *)res.cf + a * w, (size_t)len * w);
        row[L.row_deg] = 0;
        row[L.row_bindex] = rr[r][L.row_bindex];
        row[L.row_mult] = rr[r][L.row_mult];
        row[L.row_coeffs] = lead;
        row[L.row_preloop] = len % L.unroll;
        row[L.row_length] = len;
        mcf[lead] = cf;
        tr[k] = row;
    }
    f4la_b200_free_result(&res);
    for (uint32_t r = 0; r < nrows; ++r) free(rr[r]);
    if (free_basis != 0) {
        if (!g_free_basis_elements) die("interreduce_matrix_rows(free_basis != 0) needs f4la_b200_neogb_set_free_basis_elements()");
        g_free_basis_elements(bs);
    }
    free(rr);
    fld<uint32_t **>(mat, L.mat_rr) = nullptr;
    fld<uint32_t>(mat, L.mat_np) = nrows;
    T.tot.calls++;
    T.tot.kernel_launches += stats.kernel_launches;
    T.tot.ms_total += stats.ms_total;
}

} // namespace

extern "C" {

int f4la_b200_neogb_configure(const f4la_neogb_layout *layout)
{
    if (!layout || layout->struct_size != sizeof(f4la_neogb_layout)) {
        fprintf(stderr, "[f4la_b200] layout descriptor has the wrong size (header / library mismatch)\n");
        return F4LA_ERR_ARG;
    }
    L = *layout;
    g_configured = true;
    /* CUDA context, kernel module and the staging ring of the calling thread are set up here, once, before the
     * reference starts its F4 run -- not inside the first linear-algebra call.  Without a device the handshake
     * still succeeds (the layout can be checked anywhere); the first linear-algebra call then fails loudly, there
     * is no CPU fallback. */
    if (f4la_b200_device_count() <= 0) return F4LA_OK;
    if (!T.ctx) {
        if (T.device < 0) { const char *e = getenv("F4LA_B200_DEVICE"); T.device = e ? atoi(e) : 0; }
        T.ctx = f4la_b200_create(T.device);
        if (!T.ctx) return F4LA_OK;
    }
    f4la_ctx *cx = T.ctx;
    void *piece = nullptr; size_t cap = 0;
    for (int k = 0; k < 4; ++k) {
        if (f4la_b200_stage_acquire(cx, &piece, &cap) != F4LA_OK) return F4LA_ERR_CUDA;
        if (f4la_b200_stage_commit(cx, F4LA_STAGE_COLS, 4, 0, 0, 0) != F4LA_OK) return F4LA_ERR_CUDA;
    }
    return F4LA_OK;
}

void f4la_b200_neogb_set_construct_trace(void (*fn)(void *trace, void *mat)) { g_construct_trace = fn; }

void f4la_b200_neogb_set_device(int device)
{
    if (T.ctx && T.device != device) { f4la_b200_destroy(T.ctx); T.ctx = nullptr; }
    T.device = device;
}

void f4la_b200_linear_algebra(void *mat, const void *tbr, const void *bs, void *st)
{
    echelon(mat, tbr, bs, st);
}

void f4la_b200_exact_linear_algebra(void *mat, const void *tbr, const void *bs, void *st)
{
    echelon(mat, tbr, bs, st);
}

int f4la_b200_application_linear_algebra(void *mat, const void *bs, void *st)
{
    return echelon(mat, bs, bs, st, Entry::Application);
}

void f4la_b200_trace_linear_algebra(void *trace, void *mat, const void *bs, void *st)
{
    echelon(mat, bs, bs, st, Entry::Trace, trace);
}

void f4la_b200_interreduce_matrix_rows(void *mat, void *bs, void *st, int free_basis)
{
    interreduce(mat, bs, st, free_basis);
}

void f4la_b200_neogb_set_free_basis_elements(void (*fn)(void *bs)) { g_free_basis_elements = fn; }

void f4la_b200_neogb_tape_begin(void)
{
    delete T.tape;
    T.tape = new f4la_tape();
}

f4la_tape *f4la_b200_neogb_tape_end(void)
{
    f4la_tape *tp = T.tape;
    T.tape = nullptr;
    if (!tp) return nullptr;
    /* usable: every call recorded with pivots and rows from one basis, the last call is the final reduction step */
    if (!tp->valid || tp->calls.empty() || !tp->calls.back().final_step) { delete tp; return nullptr; }
    for (const TapeCall &tc : tp->calls)
        if (tc.cf_bytes != 4) { delete tp; return nullptr; }      /* multi-modular primes are 31-bit (cf32_t) */
    tp->ninit = tp->calls.front().bs_ld;
    tp->init_len.assign(tp->ninit, 0);
    uint32_t known = tp->ninit;                      /* basis elements defined so far */
    for (size_t c = 0; c < tp->calls.size(); ++c) {
        const TapeCall &tc = tp->calls[c];
        for (size_t sl = 0; sl < tc.slot_bindex.size(); ++sl) {
            const uint32_t bi = tc.slot_bindex[sl];
            if (bi >= known) { delete tp; return nullptr; }      /* refers to an element no recorded call produced */
            if (bi < tp->ninit) tp->init_len[bi] = (uint32_t)(tc.cf_ptr[sl + 1] - tc.cf_ptr[sl]);
        }
        const uint32_t np = (uint32_t)tc.out_row_ptr.size() - 1;
        if (!tc.final_step) {
            if (tc.bs_ld != known) { delete tp; return nullptr; }
            known = tc.bs_ld + np;
        } else if (c + 1 != tp->calls.size()) { delete tp; return nullptr; }
    }
    tp->residue_words = tp->calls.back().out_row_ptr.back();
    return tp;
}

void f4la_b200_tape_free(f4la_tape *tape)
{
    if (!tape) return;
    /* the contexts the structures were uploaded through may be gone already: release without touching them */
    for (auto &kv : tape->resident) {
        for (f4la_resident *r : kv.second.res) f4la_b200_release(nullptr, r);
        if (kv.second.pool) f4la_b200_pinned_free(kv.second.pool);
        for (void *d : kv.second.dev) f4la_b200_device_free(nullptr, d);
        if (kv.second.h_init) f4la_b200_pinned_free(kv.second.h_init);
        if (kv.second.h_status) f4la_b200_pinned_free(kv.second.h_status);
    }
    delete tape;
}

uint32_t f4la_b200_tape_calls(const f4la_tape *tape) { return tape ? (uint32_t)tape->calls.size() : 0; }
uint32_t f4la_b200_tape_initial_elements(const f4la_tape *tape) { return tape ? tape->ninit : 0; }
uint32_t f4la_b200_tape_initial_length(const f4la_tape *tape, uint32_t i) { return tape && i < tape->ninit ? tape->init_len[i] : 0; }
uint64_t f4la_b200_tape_residue_words(const f4la_tape *tape) { return tape ? tape->residue_words : 0; }

int f4la_b200_tape_replay(const f4la_tape *ctape, void *vctx, uint32_t p, const uint32_t *init_cf, const uint64_t *init_off,
                          uint32_t *residues, double *la_ms)
{
    f4la_tape *tape = const_cast<f4la_tape *>(ctape);
    f4la_ctx *cx = (f4la_ctx *)vctx;
    if (!tape || !cx || !init_cf || !init_off || !residues) return F4LA_ERR_ARG;
    const double t0 = now_ms();
    /* the structures of all calls, resident in HBM once per context */
    std::vector<f4la_resident *> *res_vec;
    f4la_tape::PerCtx *pcx;
    {
        std::lock_guard<std::mutex> lk(tape->mu);
        pcx = &tape->resident[vctx];
        res_vec = &pcx->res;
    }
    if (!pcx->pool) {
        /* one page-locked coefficient pool per context, sized for the largest call: allocating (or growing) page-locked
         * memory synchronises the whole device, which every other context's replay would feel */
        uint64_t mx = 0;
        for (const TapeCall &tc : tape->calls) mx = std::max<uint64_t>(mx, tc.cf_ptr.back());
        pcx->pool = f4la_b200_pinned_alloc((mx + 1) * 4);
        if (!pcx->pool) return F4LA_ERR_CUDA;
    }
    if (res_vec->empty()) {
        for (const TapeCall &tc : tape->calls) {
            f4la_flat_matrix fm;
            memset(&fm, 0, sizeof(fm));
            fm.fc = p; fm.cf_bytes = tc.cf_bytes; fm.nru = tc.nru; fm.nrl = tc.nrl; fm.ncl = tc.ncl; fm.ncr = tc.ncr;
            fm.ncf = (uint32_t)tc.slot_bindex.size(); fm.flags = tc.flags;
            fm.row_ptr = tc.row_ptr.data(); fm.cols = tc.cols.data(); fm.row_cf = tc.row_cf.data(); fm.cf_ptr = tc.cf_ptr.data();
            std::vector<uint32_t> zero(tc.cf_ptr.back() + 1, 1u);        /* placeholder pool, overwritten per prime */
            fm.cf = zero.data();
            f4la_resident *r = f4la_b200_upload(cx, &fm);
            if (!r) return F4LA_ERR_CUDA;
            res_vec->push_back(r);
        }
    }
    /* basis elements of this prime, by basis index */
    std::vector<std::vector<uint32_t>> cf(tape->ninit);
    for (uint32_t i = 0; i < tape->ninit; ++i) {
        const uint64_t len = init_off[i + 1] - init_off[i];
        if (tape->init_len[i] && len != tape->init_len[i]) return F4LA_ERR_ARG;
        cf[i].assign(init_cf + init_off[i], init_cf + init_off[i + 1]);
    }
    static const bool dbg = getenv("F4LA_B200_TAPE_DEBUG") != nullptr;
    double d_pool = 0, d_upd = 0, d_red = 0, d_post = 0, d_gpu[4] = {0, 0, 0, 0};
    uint32_t d_launch = 0;
    for (size_t c = 0; c < tape->calls.size(); ++c) {
        const TapeCall &tc = tape->calls[c];
        const uint64_t ncoef = tc.cf_ptr.back();
        const double tc0 = now_ms();
        uint32_t *pool = (uint32_t *)pcx->pool;
        for (size_t sl = 0; sl < tc.slot_bindex.size(); ++sl) {
            const std::vector<uint32_t> &src = cf[tc.slot_bindex[sl]];
            const uint64_t len = tc.cf_ptr[sl + 1] - tc.cf_ptr[sl];
            if (src.size() != len) return F4LA_ERR_PATTERN;
            memcpy(pool + tc.cf_ptr[sl], src.data(), len * 4);
        }
        f4la_resident *rm = (*res_vec)[c];
        const double tc1 = now_ms();
        if (f4la_b200_resident_update(cx, rm, p, pool) != F4LA_OK) return F4LA_ERR_CUDA;
        const double tc2 = now_ms();
        f4la_flat_result res;
        f4la_stats st;
        if (f4la_b200_reduce_resident(cx, rm, &res, dbg ? &st : nullptr) != F4LA_OK) return F4LA_ERR_CUDA;
        const double tc3 = now_ms();
        if (dbg) {
            d_gpu[0] += st.ms_prep; d_gpu[1] += st.ms_reduce; d_gpu[2] += st.ms_echelon; d_gpu[3] += st.ms_d2h;
            d_launch += st.kernel_launches;
            static const bool dbg2 = atoi(getenv("F4LA_B200_TAPE_DEBUG")) > 1;
            if (dbg2)
                fprintf(stderr, "[f4la tape]   call %zu: %u+%u rows, %u+%u cols, %llu coefs, rank %u: prep %.3f reduce %.3f "
                        "echelon %.3f d2h %.3f ms, %u launches, wall %.3f\n", c, tc.nru, tc.nrl, tc.ncl, tc.ncr,
                        (unsigned long long)ncoef, st.rank, st.ms_prep, st.ms_reduce, st.ms_echelon, st.ms_d2h,
                        st.kernel_launches, tc3 - tc2);
        }
        const uint32_t np = (uint32_t)tc.out_row_ptr.size() - 1;
        int rc = F4LA_OK;
        if (res.nrows != np) rc = F4LA_ERR_BADPRIME;                                   /* f4.c:433-441 */
        else if (memcmp(res.row_ptr, tc.out_row_ptr.data(), ((size_t)np + 1) * 8) ||
                 memcmp(res.cols, tc.out_cols.data(), tc.out_cols.size() * 4)) rc = F4LA_ERR_PATTERN;
        if (rc == F4LA_OK) {
            const uint32_t *rcf = (const uint32_t *)res.cf;
            if (!tc.final_step) {
                /* convert_sparse_matrix_rows_to_basis_elements(-1, ...): element bs_ld + k is result row np-1-k */
                cf.resize((size_t)tc.bs_ld + np);
                for (uint32_t k = 0; k < np; ++k) {
                    const uint32_t i = np - 1 - k;
                    cf[tc.bs_ld + k].assign(rcf + res.row_ptr[i], rcf + res.row_ptr[i + 1]);
                }
            } else {
                memcpy(residues, rcf, (size_t)res.row_ptr[np] * 4);                   /* f4.c:612-622: element i = row i */
            }
        }
        f4la_b200_free_result(&res);
        if (rc != F4LA_OK) return rc;
        d_pool += tc1 - tc0; d_upd += tc2 - tc1; d_red += tc3 - tc2; d_post += now_ms() - tc3;
    }
    if (dbg)
        fprintf(stderr, "[f4la tape] prime %u: %.2f ms (setup %.2f, pools %.2f, update %.2f, reduce %.2f [device: prep %.2f "
                "reduce %.2f echelon %.2f d2h %.2f; %u launches], results %.2f)\n", p, now_ms() - t0, 0.0, d_pool, d_upd,
                d_red, d_gpu[0], d_gpu[1], d_gpu[2], d_gpu[3], d_launch, d_post);
    if (la_ms) *la_ms += now_ms() - t0;
    return F4LA_OK;
}

} /* extern "C" */

namespace {

template <typename V> const void *to_device(f4la_ctx *cx, f4la_tape::PerCtx &pc, const std::vector<V> &v)
{
    void *d = f4la_b200_device_alloc(cx, v.size() * sizeof(V) + 16);
    if (!d) return nullptr;
    pc.dev.push_back(d);
    if (!v.empty() && f4la_b200_device_copy(cx, d, v.data(), v.size() * sizeof(V), 1, 1) != F4LA_OK) return nullptr;
    return d;
}

/* Everything a replay needs, on the device: where every coefficient array of every call comes from in the arena of
 * basis elements, the recorded support of every result row as dense-column indices, and where the rows go. */
bool build_plan(f4la_tape *tape, f4la_tape::PerCtx &pc, f4la_ctx *cx, const uint64_t *init_off)
{
    std::vector<uint64_t> elem_off(tape->ninit), elem_len(tape->ninit);
    for (uint32_t i = 0; i < tape->ninit; ++i) { elem_off[i] = init_off[i]; elem_len[i] = init_off[i + 1] - init_off[i]; }
    uint64_t words = init_off[tape->ninit];
    pc.init_off.assign(init_off, init_off + tape->ninit + 1);
    pc.plan.clear();
    for (size_t c = 0; c < tape->calls.size(); ++c) {
        const TapeCall &tc = tape->calls[c];
        if (tc.cf_bytes != 4) return false;
        f4la_replay_call rc;
        memset(&rc, 0, sizeof(rc));
        const uint32_t nslots = (uint32_t)tc.slot_bindex.size();
        std::vector<uint64_t> src(nslots), dst(nslots);
        std::vector<uint32_t> len(nslots);
        for (uint32_t sl = 0; sl < nslots; ++sl) {
            const uint32_t e = tc.slot_bindex[sl];
            const uint64_t l = tc.cf_ptr[sl + 1] - tc.cf_ptr[sl];
            if (e >= elem_off.size() || elem_len[e] != l) return false;
            src[sl] = elem_off[e]; dst[sl] = tc.cf_ptr[sl]; len[sl] = (uint32_t)l;
        }
        const uint32_t np = (uint32_t)tc.out_row_ptr.size() - 1;
        const uint32_t nc = tc.ncl + tc.ncr;
        const bool by_lead = (tc.flags & F4LA_FLAG_PIVOT_BY_LEAD) != 0;
        std::vector<uint32_t> colmap(nc);
        uint32_t ncl_eff = 0;
        if (f4la_b200_resident_layout(cx, pc.res[c], colmap.data(), &ncl_eff) != F4LA_OK) return false;
        std::vector<uint32_t> idx(tc.out_cols.size());
        for (uint32_t i = 0; i < np; ++i)
            for (uint64_t t = tc.out_row_ptr[i]; t < tc.out_row_ptr[i + 1]; ++t) {
                if (by_lead && t == tc.out_row_ptr[i]) { idx[t] = 0xffffffffu; continue; }      /* the lead term, value 1 */
                const uint32_t col = tc.out_cols[t];
                if (col >= nc || colmap[col] < ncl_eff) return false;
                idx[t] = colmap[col] - ncl_eff;
            }
        std::vector<uint64_t> row_dst(np);
        if (tc.final_step) {
            /* f4.c:612-622: element i of the reduced basis = row i; the rows, one after the other, are the residues */
            pc.residue_off = words;
            for (uint32_t i = 0; i < np; ++i) row_dst[i] = words + tc.out_row_ptr[i];
            words += tc.out_row_ptr[np];
        } else {
            /* convert_sparse_matrix_rows_to_basis_elements(-1, ...): element bs_ld + k is result row np-1-k */
            if (elem_off.size() < (size_t)tc.bs_ld + np) { elem_off.resize((size_t)tc.bs_ld + np); elem_len.resize((size_t)tc.bs_ld + np); }
            for (uint32_t k = 0; k < np; ++k) {
                const uint32_t i = np - 1 - k;
                const uint64_t l = tc.out_row_ptr[i + 1] - tc.out_row_ptr[i];
                elem_off[tc.bs_ld + k] = words; elem_len[tc.bs_ld + k] = l;
                row_dst[i] = words;
                words += l;
            }
        }
        rc.expect_rows = np; rc.nslots = nslots;
        rc.gather_src = (const uint64_t *)to_device(cx, pc, src);
        rc.gather_dst = (const uint64_t *)to_device(cx, pc, dst);
        rc.gather_len = (const uint32_t *)to_device(cx, pc, len);
        rc.row_ptr = (const uint64_t *)to_device(cx, pc, tc.out_row_ptr);
        rc.idx = (const uint32_t *)to_device(cx, pc, idx);
        rc.row_dst = (const uint64_t *)to_device(cx, pc, row_dst);
        if (!rc.gather_src || !rc.gather_dst || !rc.gather_len || !rc.row_ptr || !rc.idx || !rc.row_dst) return false;
        pc.plan.push_back(rc);
    }
    if (tape->calls.empty() || !tape->calls.back().final_step) return false;
    pc.arena_words = words;
    pc.arena = (uint32_t *)f4la_b200_device_alloc(cx, words * 4 + 16);
    pc.d_status = (uint32_t *)f4la_b200_device_alloc(cx, f4la_tape::kInflight * 4);
    if (!pc.arena || !pc.d_status) return false;
    pc.dev.push_back(pc.arena); pc.dev.push_back(pc.d_status);
    pc.h_init = (uint32_t *)f4la_b200_pinned_alloc((size_t)f4la_tape::kInflight * (init_off[tape->ninit] + 1) * 4);
    pc.h_status = (uint32_t *)f4la_b200_pinned_alloc(f4la_tape::kInflight * 4);
    if (!pc.h_init || !pc.h_status) return false;
    for (auto &rc : pc.plan) rc.arena = pc.arena;
    return true;
}

f4la_tape::PerCtx *per_ctx(f4la_tape *tape, void *vctx)
{
    std::lock_guard<std::mutex> lk(tape->mu);
    return &tape->resident[vctx];
}

int status_to_rc(uint32_t st)
{
    if (st & 4u) return F4LA_ERR_CUDA;
    if (st & 1u) return F4LA_ERR_BADPRIME;
    if (st & 2u) return F4LA_ERR_PATTERN;
    return F4LA_OK;
}

} // namespace

extern "C" {

int f4la_b200_tape_replay_collect(const f4la_tape *ctape, void *vctx, int *rcs, int max_rcs)
{
    f4la_tape *tape = const_cast<f4la_tape *>(ctape);
    f4la_ctx *cx = (f4la_ctx *)vctx;
    if (!tape || !cx) return -1;
    f4la_tape::PerCtx *pc = per_ctx(tape, vctx);
    if (pc->inflight && f4la_b200_device_copy(cx, nullptr, nullptr, 0, 0, 1) != F4LA_OK) {
        /* the stream failed: every replay in flight is lost */
        for (int k = 0; k < pc->inflight; ++k) pc->h_status[k] = 4u;
    }
    int n = 0, slot = 0;
    size_t sync_k = 0;
    for (int which : pc->order) {
        const int rc = which ? pc->done_rc[sync_k++] : status_to_rc(pc->h_status[slot++]);
        if (rcs && n < max_rcs) rcs[n] = rc;
        ++n;
    }
    pc->order.clear(); pc->done_rc.clear(); pc->inflight = 0;
    return n;
}

int f4la_b200_tape_replay_enqueue(const f4la_tape *ctape, void *vctx, uint32_t p, const uint32_t *init_cf,
                                  const uint64_t *init_off, uint32_t *residues)
{
    f4la_tape *tape = const_cast<f4la_tape *>(ctape);
    f4la_ctx *cx = (f4la_ctx *)vctx;
    if (!tape || !cx || !init_cf || !init_off || !residues) return F4LA_ERR_ARG;
    f4la_tape::PerCtx *pc = per_ctx(tape, vctx);
    static const bool no_async = getenv("F4LA_B200_TAPE_ASYNC") && atoi(getenv("F4LA_B200_TAPE_ASYNC")) == 0;
    if (pc->warmed && pc->async_state == 0 && !no_async)
        pc->async_state = build_plan(tape, *pc, cx, init_off) ? 1 : -1;
    const bool same_layout = pc->async_state == 1 &&
        memcmp(pc->init_off.data(), init_off, ((size_t)tape->ninit + 1) * 8) == 0;
    if (!same_layout || no_async) {
        /* host-assembled replay (also the first one on a context: it uploads the structures and fills the layout
         * caches the device-resident replay needs); everything enqueued before it must be done first */
        if (pc->inflight == f4la_tape::kInflight) return F4LA_ERR_ARG;      /* collect first */
        const int rc = f4la_b200_tape_replay(tape, vctx, p, init_cf, init_off, residues, nullptr);
        pc->warmed = true;
        pc->done_rc.push_back(rc);
        pc->order.push_back(1);
        return F4LA_OK;
    }
    if (pc->inflight == f4la_tape::kInflight) return F4LA_ERR_ARG;          /* collect first */
    const int slot = pc->inflight;
    const uint64_t init_words = init_off[tape->ninit];
    uint32_t *h_init = pc->h_init + (size_t)slot * (init_words + 1);
    memcpy(h_init, init_cf, init_words * 4);
    uint32_t *d_status = pc->d_status + slot;
    if (f4la_b200_device_copy(cx, pc->arena, h_init, init_words * 4, 1, 0) != F4LA_OK) return F4LA_ERR_CUDA;
    if (f4la_b200_device_zero(cx, d_status, 4) != F4LA_OK) return F4LA_ERR_CUDA;
    for (size_t c = 0; c < pc->plan.size(); ++c) {
        f4la_replay_call rc = pc->plan[c];
        rc.status = d_status;
        if (f4la_b200_replay_call(cx, pc->res[c], p, &rc) != F4LA_OK) {
            /* not expressible without host round trips (e.g. a block that needs compress-and-verify): from now on the
             * host-assembled replay; this prime's partial work is abandoned */
            pc->async_state = -1;
            f4la_b200_device_copy(cx, nullptr, nullptr, 0, 0, 1);
            const int r2 = f4la_b200_tape_replay(tape, vctx, p, init_cf, init_off, residues, nullptr);
            pc->done_rc.push_back(r2);
            pc->order.push_back(1);
            return F4LA_OK;
        }
    }
    if (f4la_b200_device_copy(cx, residues, pc->arena + pc->residue_off, tape->residue_words * 4, 0, 0) != F4LA_OK) return F4LA_ERR_CUDA;
    if (f4la_b200_device_copy(cx, pc->h_status + slot, d_status, 4, 0, 0) != F4LA_OK) return F4LA_ERR_CUDA;
    pc->inflight++;
    pc->order.push_back(0);
    return F4LA_OK;
}

} /* extern "C" */

extern "C" {

void f4la_b200_neogb_get_totals(f4la_neogb_totals *t, int reset)
{
    if (t) *t = T.tot;
    if (reset) T.tot = f4la_neogb_totals{};
}

} /* extern "C" */
