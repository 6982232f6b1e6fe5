/* f4la_device.cu -- host side of the flat C ABI (include/f4la_b200.h):
 * device memory, streams, launch orchestration, result materialisation.
 * The kernels live in f4la_kernels.cuh.  No CPU fallback: without a usable
 * sm_100 device f4la_b200_create() fails and says so. */
#include <cuda_runtime.h>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <omp.h>

#include "f4la_b200.h"
#include "f4la_kernels.cuh"
#include "f4la_tc.cuh"

using namespace f4la;

namespace {

struct DevBuf {
    void  *p = nullptr;
    size_t cap = 0;
};

struct HostBuf {            /* pinned */
    void  *p = nullptr;
    size_t cap = 0;
};

/* bytes of V (vectors of the chunks in flight).  Measured on K12: keeping V inside the 126 MB L2
 * (88 MB) is SLOWER than letting all chunks run in one launch (1.3 GB of V, most hot vectors
 * still hit L2): the kernel is latency bound and wants the parallelism. */
constexpr size_t kL2BudgetDefault = (size_t)1024 << 20;

double now_ms()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
}

} // namespace

constexpr int kMaxFlowEvents = 510;                /* 3 per launch group: before / after the dataflow kernel / after the Schur product */
constexpr int kRingSlots = 4;
constexpr size_t kRingPieceBytes = (size_t)16 << 20;

struct f4la_ctx {
    int           device = 0;
    cudaStream_t  stream = nullptr;
    cudaStream_t  copy_stream = nullptr;  /* f4la_b200_reduce(): the lower rows travel while the pull lists are built */
    cudaEvent_t   ev_copy = nullptr;
    cudaEvent_t   ev_side[2] = {nullptr, nullptr};   /* fork / join of the side stream that builds the Schur B tiles */
    bool          copy_pending = false;
    int           split_h2d = 1;          /* F4LA_B200_SPLIT_H2D=0: one copy on the compute stream */
    cudaEvent_t   ev[8] = {};
    cudaEvent_t   dbg_ev[2] = {};                /* F4LA_B200_DEBUG only */
    cudaEvent_t   flow_ev[kMaxFlowEvents] = {};   /* around every launch of the hot kernel */
    int           sm_count = 0;
    char          err[512] = {0};
    uint32_t      launches = 0;
    /* scratch (grow-only) */
    DevBuf col_cnt, col_ptr, cursor, ent, level, slots, slot_ptr, order, flags32;
    DevBuf V, D, colflag, gebytes, pivcol, pivrow, pivinv, dense_out;
    DevBuf used, ispiv, gestate;                  /* views into gebytes (never allocated or freed themselves) */
    size_t        fused_smem_optin = 0;
    int           coop_ok = 0;          /* device supports cooperative launches (k_ge_fused) */
    DevBuf sched, items, flowflags, tcA, tcB, scB, scFlagA, scFlagB, scList, scCnt, scP;
    int           ge_mode = 0;          /* F4LA_B200_GE: (default) one persistent cooperative launch (k_ge_fused); "batched": two launches per
                                           batch of 8 pivots with the batch's pivot columns in shared memory (the round-1 default);
                                           "uncached": batched, pivot columns re-read from L2; "columnwise": one find + one
                                           elimination pass per pivot (the fallbacks for very tall blocks, kept selectable for A-B) */
    size_t        cached_smem_optin = 0, compress_smem_optin = 0;
    int           ge_compress = 1;      /* F4LA_B200_COMPRESS=0: never use compress-and-verify on tall trailing blocks */
    DevBuf        D2, combo;            /* compressed block, combination member lists */
    uint64_t      d2_ld = 0;
    uint64_t      compress_attempts = 0, compress_rejected = 0, compress_escalated = 0;
    uint32_t      compress_rows = 768;  /* F4LA_B200_COMPRESS_ROWS: combinations of the first attempt (x4 per escalation) */
    uint32_t      ge_batch = 8, ge_miss_limit = 4;   /* F4LA_B200_GE_BATCH (<= 32), F4LA_B200_GE_MISSES: pivots per batch, empty visits that end one */
    int           schur_tc = 1;         /* F4LA_B200_SCHUR_TC=0: non-pivot columns in the dataflow kernel instead of the tensor-core Schur product */
    double        schur_min_macs = 2e9; /* F4LA_B200_SCHUR_MIN: smallest (non-pivot entries x lower rows) for the tensor path */
    bool          schur_attr_set = false;
    int           tc_verify = 1;        /* F4LA_B200_TC=0: verification product on the CUDA cores (k_ge_verify) instead of tcgen05 */
    bool          tc_attr_set = false;
    uint64_t      tc_launches = 0;
    int           zerocopy = 1;         /* F4LA_B200_ZEROCOPY=0: result rows through a device buffer + copy */
    int           debug = 0;            /* F4LA_B200_DEBUG=1: statistics on stderr */
    int           debug_sync = 0;       /* F4LA_B200_SYNC=1: synchronise after every kernel launch and name a failing one */
    bool          debug_sync_reported = false;
    int           compress_fault = 0;   /* F4LA_B200_COMPRESS_FAULT=1 (tests): treat every verification as failed */
    size_t        batch_smem_optin = 0;
    DevBuf in_row_ptr, in_cols, in_row_cf, in_cf_ptr, in_cf;   /* staging of f4la_b200_reduce() inputs */
    DevBuf is_piv, pividx, colmap, invmap, leads, rba;
    int           flow_blocks_per_sm[2][kMaxT + 1] = {};
    size_t        l2_budget = 0;
    double        watchdog_ms = 20000.0; /* F4LA_B200_WATCHDOG_MS: longest wait for a source column in the dataflow kernel */
    uint32_t      ge_panel = 0;                   /* panel width cap of k_ge_fused (0: kPanelMax) */
    uint32_t      grid_cap = 0;                   /* CTAs a persistent kernel of this context may occupy (0: the device) */
    int           host_threads = 0;               /* threads of the host-side result compaction (0: OpenMP default) */
    bool          blocking_sync = false;          /* waits sleep on a blocking event instead of spinning */
    cudaEvent_t   ev_block = nullptr;
    uint64_t      seed = 0x9e3779b97f4a7c15ull;   /* random combinations of the probabilistic mode */
    int           tmax = 2;             /* F4LA_B200_TMAX: rows per chunk = 128 * T (T = 2 measured best on K12: 125 ms vs 142 / 140 ms for T = 3 / 4) */
    HostBuf h_small, h_dense, h_combo;
    /* ring of page-locked pieces for the big input arrays (f4la_b200_stage_acquire / _commit): pinned ONCE, so a
     * growing matrix never re-pins staging memory on the critical path (the cold-start cost of round 1) */
    void         *ring_buf[kRingSlots] = {};
    cudaEvent_t   ring_ev[kRingSlots] = {};
    bool          ring_busy[kRingSlots] = {};
    int           ring_next = 0, ring_held = -1;
    std::vector<uint32_t> donor;          /* per random combination: one lower row that entered it (BINDEX / MULT donor) */
};

/* What the layout kernels derive from the STRUCTURE of a matrix (pull lists, schedule, column maps): kept with a
 * resident matrix whose coefficients are replaced (f4la_b200_resident_update), so that a run for another prime
 * skips those kernels and their host round trips and only refreshes the values of the pull lists. */
struct PrepCache {
    bool      valid = false;
    uint32_t  nnz_pull = 0, nnz_schur = 0, n_items = 0, schur_nseg = 0;
    bool      schur = false;
    uint32_t *col_ptr = nullptr, *ent_src = nullptr, *sched = nullptr, *items = nullptr, *colmap = nullptr, *invmap = nullptr;
    uint2    *ent = nullptr;
};

struct f4la_resident {
    f4la_flat_matrix dims;        /* pointers below are DEVICE pointers */
    bool      cache_enabled = false;
    mutable PrepCache cache;
    uint64_t *row_ptr = nullptr;
    uint32_t *cols = nullptr;
    uint32_t *row_cf = nullptr;
    uint64_t *cf_ptr = nullptr;
    void     *cf = nullptr;
    uint64_t  nnz = 0, ncoef = 0;
    uint64_t  bytes = 0;
    bool      owned = true;
};

namespace {

int fail(f4la_ctx *ctx, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[f4la_b200] error: %s\n", ctx->err);
    return code;
}

#define CU(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess)                                                         \
            return fail(ctx, F4LA_ERR_CUDA, "%s failed: %s (%s:%d)", #call,            \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                   \
    } while (0)

#define OK(call)                                                                       \
    do {                                                                               \
        int rc_ = (call);                                                              \
        if (rc_ != F4LA_OK) return rc_;                                                \
    } while (0)

/* counts the kernels launched; with F4LA_B200_SYNC=1 every launch is followed by a synchronisation, so that a
 * faulting kernel is reported with the source line of its launch (compute-sanitizer is not always available) */
void note_launch(f4la_ctx *ctx, uint32_t n, int line)
{
    ctx->launches += n;
    if (ctx->debug_sync) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess && !ctx->debug_sync_reported) {
            ctx->debug_sync_reported = true;
            fprintf(stderr, "[f4la_b200] F4LA_B200_SYNC: kernel launched at f4la_device.cu:%d failed: %s\n", line,
                    cudaGetErrorString(e));
        }
    }
}

int ensure(f4la_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return F4LA_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    /* the matrices of an F4 run grow from call to call: doubling keeps the number of (synchronising) re-allocations of
     * a cold process small; beyond 2 GB only a quarter of head-room */
    size_t want = (bytes < ((size_t)2 << 30) ? 2 * bytes : bytes + bytes / 4) + 256;
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return F4LA_OK;
}

int ensure_host(f4la_ctx *ctx, HostBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return F4LA_OK;
    if (b.p) CU(cudaFreeHost(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = (bytes < ((size_t)256 << 20) ? 2 * bytes : bytes + bytes / 4) + 256;     /* page-locking is slow: re-pin rarely */
    CU(cudaMallocHost(&b.p, want));
    b.cap = want;
    return F4LA_OK;
}

/* Wait for the context's stream.  With several host threads per core (f4la_b200_set_host_mode) the wait sleeps on a
 * blocking event instead of spinning in the driver, so that the core serves another thread's prime meanwhile. */
int stream_sync(f4la_ctx *ctx)
{
    if (!ctx->blocking_sync) { CU(cudaStreamSynchronize(ctx->stream)); return F4LA_OK; }
    if (!ctx->ev_block) CU(cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev_block, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev_block));
    return F4LA_OK;
}

/* (re)allocate one buffer of a resident matrix's layout cache */
int cache_alloc(f4la_ctx *ctx, void **p, size_t bytes)
{
    if (*p) CU(cudaFree(*p));
    *p = nullptr;
    CU(cudaMalloc(p, bytes + 16));
    return F4LA_OK;
}

void free_dev(DevBuf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

template <typename T> T *as(DevBuf &b) { return reinterpret_cast<T *>(b.p); }

inline unsigned blocks_for_warps(uint64_t nwarps, int threads = 256)
{
    uint64_t b = (nwarps * 32 + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 148 * 64) b = 148 * 64;
    return (unsigned)b;
}

template <int T, bool WIDE>
int flow_occupancy()
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_pull_flow<T, WIDE>, kSolveThreads, 0) != cudaSuccess) n = 2;
    return n < 1 ? 1 : n;
}

int flow_blocks(f4la_ctx *ctx, int T, bool wide)
{
    int &c = ctx->flow_blocks_per_sm[wide ? 1 : 0][T];
    if (c == 0) {
#define OCC(TT) (wide ? flow_occupancy<TT, true>() : flow_occupancy<TT, false>())
        switch (T) {
            case 1: c = OCC(1); break;
            case 2: c = OCC(2); break;
            case 3: c = OCC(3); break;
            default: c = OCC(4); break;
        }
#undef OCC
    }
    return c;
}

void launch_flow(f4la_ctx *ctx, int T, const FlowArgs &a, bool wide)
{
    const uint64_t total = (uint64_t)a.n_items * a.G;
    if (total == 0) return;
    uint64_t g = (uint64_t)ctx->sm_count * flow_blocks(ctx, T, wide);
    if (g > total) g = total;
    if (ctx->grid_cap && g > ctx->grid_cap) g = ctx->grid_cap;   /* several contexts share the device */
    const unsigned grid = (unsigned)g;
#define FLOW(TT)                                                                                   \
    if (wide) k_pull_flow<TT, true><<<grid, kSolveThreads, 0, ctx->stream>>>(a);                    \
    else      k_pull_flow<TT, false><<<grid, kSolveThreads, 0, ctx->stream>>>(a);
    switch (T) {
        case 1: FLOW(1); break;
        case 2: FLOW(2); break;
        case 3: FLOW(3); break;
        default: FLOW(4); break;
    }
#undef FLOW
    note_launch(ctx, 1, __LINE__);
}

void empty_result(f4la_flat_result *out, uint32_t cf_bytes)
{
    memset(out, 0, sizeof(*out));
    out->cf_bytes = cf_bytes;
    out->row_ptr = (uint64_t *)calloc(1, sizeof(uint64_t));
    out->cols = (uint32_t *)malloc(sizeof(uint32_t));
    out->cf = malloc(4);
    out->src_row = (uint32_t *)malloc(sizeof(uint32_t));
}

/* same finaliser as the device-side mix64 */
uint64_t host_mix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

/* C = A * B mod p (store) or D == A * B mod p (verify) on the tensor cores, see f4la_tc.cuh. */
int launch_tc_gemm(f4la_ctx *ctx, TcGemmArgs g)
{
    if (g.K == 0 || g.K > kTcMaxK || g.M % kTcBM != 0 || g.M == 0 || g.N == 0)
        return fail(ctx, F4LA_ERR_ARG, "tensor-core product: unsupported shape %u x %u x %u", g.M, g.N, g.K);
    if (!ctx->tc_attr_set) {
        CU(cudaFuncSetAttribute(k_tc_gemm_modp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
        ctx->tc_attr_set = true;
    }
    g.w32 = (uint32_t)((1ull << 32) % g.p);
    g.pinv = ~0ULL / g.p;
    const uint32_t mblk = g.M / kTcBM, nblk = (g.N + kTcBN - 1) / kTcBN, nkb = (g.K + kTcBK - 1) / kTcBK;
    /* both operands once, as limb planes in the shared-memory image of the pipeline stages */
    OK(ensure(ctx, ctx->tcA, (size_t)mblk * nkb * kTcStageA));
    OK(ensure(ctx, ctx->tcB, (size_t)nblk * nkb * kTcStageB));
    g.Asplit = as<uint8_t>(ctx->tcA); g.Bsplit = as<uint8_t>(ctx->tcB);
    k_tc_split_a<<<dim3(mblk, nkb), 128, 0, ctx->stream>>>(g);
    k_tc_split_b<<<dim3(nblk, nkb), 128, 0, ctx->stream>>>(g);
    note_launch(ctx, 2, __LINE__);
    dim3 grid(mblk, nblk);
    k_tc_gemm_modp<<<grid, kTcThreads, kTcSmemBytes, ctx->stream>>>(g);
    note_launch(ctx, 1, __LINE__);
    ctx->tc_launches++;
    return F4LA_OK;
}

/* Gauss-Jordan on a column-major block (ncr columns of ld words, nrl rows used): pivot columns, rows and
 * inverses are left in ctx->pivcol / pivrow / pivinv, the block holds the reduced rows unnormalised
 * (k_ge_extract scales them). */
int gauss_jordan(f4la_ctx *ctx, uint32_t *Dm, uint64_t ld, uint32_t nrl, uint32_t ncr, uint32_t p, uint64_t pinv,
                 uint32_t *rank_out, const uint32_t *expect_rank = nullptr)
{
    cudaStream_t s = ctx->stream;
    const uint32_t maxrank = nrl < ncr ? nrl : ncr;
    OK(ensure(ctx, ctx->colflag, (size_t)ncr + 16));
    /* used bytes | pivot-column bytes | GEState + grid barrier words of k_ge_fused: one buffer, ONE memset
     * (ctx->used / ispiv / gestate are views into it) */
    const size_t off_ispiv = ((size_t)ld + 16 + 255) & ~(size_t)255;
    const size_t off_state = off_ispiv + (((size_t)ncr + 16 + 255) & ~(size_t)255);
    OK(ensure(ctx, ctx->gebytes, off_state + 256 + 64));
    ctx->used.p = ctx->gebytes.p;
    ctx->ispiv.p = static_cast<char *>(ctx->gebytes.p) + off_ispiv;
    ctx->gestate.p = static_cast<char *>(ctx->gebytes.p) + off_state;
    OK(ensure(ctx, ctx->pivcol, ((size_t)maxrank + 1) * 4));
    OK(ensure(ctx, ctx->pivrow, ((size_t)maxrank + 1) * 4));
    OK(ensure(ctx, ctx->pivinv, ((size_t)maxrank + 1) * 4));
    static_assert(sizeof(GEState) <= 256, "GEState");
    CU(cudaMemsetAsync(ctx->gebytes.p, 0, off_state + 256 + 64, s));
    GEState *h_state = (GEState *)ctx->h_small.p;
    const size_t smem_cap = (size_t)216 * 1024;
    /* pivots per batch such that the batch's pivot columns (+ one column, + the used bytes) fit in
     * the shared memory of one SM */
    const uint32_t nrows4 = (nrl + 3u) & ~3u;
    uint32_t nb_fit = 0;
    if ((size_t)nrows4 * 5 + 64 <= smem_cap)
        nb_fit = (uint32_t)((smem_cap - 64 - nrows4) / ((size_t)nrows4 * 4)) - 1;
    if (nb_fit > ctx->ge_batch) nb_fit = ctx->ge_batch;
    /* one persistent cooperative launch: panels of up to 32 pivots factored in shared memory, rank-b updates
     * behind grid barriers (k_ge_fused); needs a panel of >= 2 columns of nrows4 words in shared memory */
    uint32_t pw = 0;
    {
        const size_t dyn_cap = (size_t)200 * 1024;
        if ((size_t)nrows4 * 9 <= dyn_cap) pw = (uint32_t)((dyn_cap - nrows4) / ((size_t)nrows4 * 4));
        if (pw > kPanelMax) pw = kPanelMax;
        if (ctx->ge_panel && pw > ctx->ge_panel) pw = ctx->ge_panel;
    }
    if (expect_rank && !(pw >= 2 && ctx->ge_mode == 0 && ctx->coop_ok))
        return fail(ctx, F4LA_ERR_ARG, "a replay without host round trips needs the persistent Gauss-Jordan kernel");
    if (pw >= 2 && ctx->ge_mode == 0 && ctx->coop_ok) {
        const size_t dyn = (size_t)pw * nrows4 * 4 + nrows4;
        if (dyn > ctx->fused_smem_optin) {
            CU(cudaFuncSetAttribute(k_ge_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)200 * 1024)));
            ctx->fused_smem_optin = (size_t)200 * 1024;
        }
        k_ge_init_flags<<<ncr, 256, 0, s>>>(Dm, ld, nrl, ncr, as<uint8_t>(ctx->colflag));
        note_launch(ctx, 1, __LINE__);
        GeFusedArgs a;
        memset(&a, 0, sizeof(a));
        a.D = Dm; a.ld = ld; a.nrows = nrl; a.nrows4 = nrows4; a.ncr = ncr; a.pw = pw;
        a.colflag = as<uint8_t>(ctx->colflag); a.used = as<uint8_t>(ctx->used); a.ispiv = as<uint8_t>(ctx->ispiv);
        a.st = as<GEState>(ctx->gestate);
        a.pivcol = as<uint32_t>(ctx->pivcol); a.pivrow = as<uint32_t>(ctx->pivrow); a.pivinv = as<uint32_t>(ctx->pivinv);
        a.p = p; a.two64 = (uint32_t)(((~0ull) % p + 1) % p); a.pinv = pinv;
        a.bar = as<uint32_t>(ctx->gestate) + 64; a.err = as<uint32_t>(ctx->flags32);
        a.watchdog_ns = (uint64_t)(ctx->watchdog_ms * 1e6); a.max_iter = ncr + 8; a.debug = ctx->debug ? 1u : 0u;
        unsigned grid = (ncr + 31) / 32;
        if (grid > (unsigned)ctx->sm_count) grid = (unsigned)ctx->sm_count;
        if (grid < 1) grid = 1;
        void *kargs[] = {&a};
        if (ctx->debug) CU(cudaEventRecord(ctx->dbg_ev[0], s));
        CU(cudaLaunchCooperativeKernel((const void *)k_ge_fused, dim3(grid), dim3(kFusedThreads), kargs, dyn, s));
        if (ctx->debug) CU(cudaEventRecord(ctx->dbg_ev[1], s));
        note_launch(ctx, 1, __LINE__);
        if (expect_rank) {          /* replay: no read-back; the state stays on the device for k_replay_compact */
            *rank_out = *expect_rank;
            return F4LA_OK;
        }
        CU(cudaMemcpyAsync(h_state, ctx->gestate.p, sizeof(GEState), cudaMemcpyDeviceToHost, s));
        CU(cudaGetLastError());
        OK(stream_sync(ctx));
        if (ctx->debug) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->dbg_ev[0], ctx->dbg_ev[1]);
            fprintf(stderr, "[f4la debug] k_ge_fused %u x %u: %.3f ms, grid %u, panel width %u; CTA 0 kilo-cycles: gather %llu load %llu "
                    "select+inverse %llu eliminate %llu write-back %llu update %llu barriers %llu total %llu; any CTA: longest update %llu, "
                    "longest wait for the panel %llu\n", nrl, ncr, ms, grid, pw,
                    h_state->t[0] / 1000, h_state->t[1] / 1000, h_state->t[2] / 1000, h_state->t[3] / 1000, h_state->t[4] / 1000,
                    h_state->t[5] / 1000, h_state->t[6] / 1000, h_state->t[7] / 1000, h_state->t[8] / 1000, h_state->t[9] / 1000);
        }
        if (!h_state->done)
            return fail(ctx, F4LA_ERR_CUDA, "fused Gauss-Jordan did not finish (rank %u, next column %u of %u)", h_state->rank,
                        h_state->next_col, ncr);
    } else if (nb_fit >= 2 && (ctx->ge_mode == 0 || ctx->ge_mode == 3)) {
        /* batched, pivot columns of the batch cached in shared memory by both passes */
        const size_t disc_smem = ((size_t)nb_fit + 1) * nrows4 * 4 + nrows4, upd_smem = (size_t)nb_fit * nrows4 * 4 + nrows4;
        if (disc_smem > ctx->cached_smem_optin) {
            CU(cudaFuncSetAttribute(k_ge_discover_cached, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            CU(cudaFuncSetAttribute(k_ge_batch_update_cached, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            ctx->cached_smem_optin = smem_cap;
        }
        const uint32_t two64 = (uint32_t)(((~0ull) % p + 1) % p);
        uint32_t upd_grid = (ncr + 31) / 32;
        if (upd_grid > (uint32_t)ctx->sm_count) upd_grid = (uint32_t)ctx->sm_count;
        k_ge_init_flags<<<ncr, 256, 0, s>>>(Dm, ld, nrl, ncr, as<uint8_t>(ctx->colflag));
        note_launch(ctx, 1, __LINE__);
        for (uint32_t batches = 0;;) {
            const int group = 4;
            for (int k = 0; k < group; ++k) {
                k_ge_discover_cached<<<1, kDiscoverThreads, disc_smem, s>>>(
                    Dm, ld, nrl, nrows4, ncr, nb_fit, ctx->ge_miss_limit, as<uint8_t>(ctx->colflag), as<uint8_t>(ctx->used),
                    as<uint8_t>(ctx->ispiv), as<GEState>(ctx->gestate), as<uint32_t>(ctx->pivcol),
                    as<uint32_t>(ctx->pivrow), as<uint32_t>(ctx->pivinv), p, pinv, two64);
                k_ge_batch_update_cached<<<upd_grid, kUpdate2Threads, upd_smem, s>>>(
                    Dm, ld, nrl, nrows4, ncr, nb_fit, as<uint8_t>(ctx->colflag), as<uint8_t>(ctx->used),
                    as<uint8_t>(ctx->ispiv), as<GEState>(ctx->gestate), as<uint32_t>(ctx->pivcol),
                    as<uint32_t>(ctx->pivrow), as<uint32_t>(ctx->pivinv), p, pinv, two64);
                note_launch(ctx, 2, __LINE__);
            }
            batches += group;
            CU(cudaMemcpyAsync(h_state, ctx->gestate.p, sizeof(GEState), cudaMemcpyDeviceToHost, s));
            CU(cudaGetLastError());
            OK(stream_sync(ctx));
            if (h_state->done) break;
            if ((uint64_t)batches * 1 > (uint64_t)maxrank + 2 * group)
                return fail(ctx, F4LA_ERR_CUDA, "batched Gauss-Jordan did not terminate (rank %u)", h_state->rank);
        }
    } else if ((size_t)nrl * 5 + 64 <= smem_cap && ctx->ge_mode != 2) {
        /* batched: discover up to kGeBatch pivots on one SM, then one update pass over the block */
        const size_t disc_smem = (size_t)nrl * 5 + 64, upd_smem = (size_t)nrl * 4;
        if (disc_smem > ctx->batch_smem_optin) {
            CU(cudaFuncSetAttribute(k_ge_discover, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            CU(cudaFuncSetAttribute(k_ge_batch_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            ctx->batch_smem_optin = smem_cap;
        }
        k_ge_init_flags<<<ncr, 256, 0, s>>>(Dm, ld, nrl, ncr, as<uint8_t>(ctx->colflag));
        note_launch(ctx, 1, __LINE__);
        for (uint32_t batches = 0;;) {
            const int group = 4;
            for (int k = 0; k < group; ++k) {
                k_ge_discover<<<1, kDiscoverThreads, disc_smem, s>>>(
                    Dm, ld, nrl, ncr, as<uint8_t>(ctx->colflag), as<uint8_t>(ctx->used), as<uint8_t>(ctx->ispiv),
                    as<GEState>(ctx->gestate), as<uint32_t>(ctx->pivcol), as<uint32_t>(ctx->pivrow),
                    as<uint32_t>(ctx->pivinv), p, pinv);
                k_ge_batch_update<<<ncr, kUpdateThreads, upd_smem, s>>>(
                    Dm, ld, nrl, ncr, as<uint8_t>(ctx->colflag), as<uint8_t>(ctx->used), as<uint8_t>(ctx->ispiv),
                    as<GEState>(ctx->gestate), as<uint32_t>(ctx->pivcol), as<uint32_t>(ctx->pivrow),
                    as<uint32_t>(ctx->pivinv), p, pinv);
                note_launch(ctx, 2, __LINE__);
            }
            batches += group;
            CU(cudaMemcpyAsync(h_state, ctx->gestate.p, sizeof(GEState), cudaMemcpyDeviceToHost, s));
            CU(cudaGetLastError());
            OK(stream_sync(ctx));
            if (h_state->done) break;
            if ((uint64_t)batches * 1 > (uint64_t)maxrank + 2 * group)
                return fail(ctx, F4LA_ERR_CUDA, "batched Gauss-Jordan did not terminate (rank %u)", h_state->rank);
        }
    } else {
    k_ge_init_flags<<<ncr, 256, 0, s>>>(Dm, ld, nrl, ncr, as<uint8_t>(ctx->colflag));
    note_launch(ctx, 1, __LINE__);
    for (uint32_t done_steps = 0;;) {
        const int batch = 16;
        for (int k = 0; k < batch; ++k) {
            k_ge_find<<<1, 1024, 0, s>>>(Dm, ld, nrl, ncr, as<uint8_t>(ctx->colflag), as<uint8_t>(ctx->used),
                                         as<uint8_t>(ctx->ispiv), as<GEState>(ctx->gestate),
                                         as<uint32_t>(ctx->pivcol), as<uint32_t>(ctx->pivrow),
                                         as<uint32_t>(ctx->pivinv), p);
            k_ge_eliminate<<<ncr, 256, 0, s>>>(Dm, ld, nrl, ncr, as<uint8_t>(ctx->colflag), as<uint8_t>(ctx->used),
                                               as<GEState>(ctx->gestate), p, pinv);
            note_launch(ctx, 2, __LINE__);
        }
        done_steps += batch;
        CU(cudaMemcpyAsync(h_state, ctx->gestate.p, sizeof(GEState), cudaMemcpyDeviceToHost, s));
        CU(cudaGetLastError());
        OK(stream_sync(ctx));
        if (h_state->done) break;
        if (done_steps > maxrank + 2 * batch)
            return fail(ctx, F4LA_ERR_CUDA, "Gauss-Jordan did not terminate (rank %u)", h_state->rank);
    }
    }
    if (ctx->debug)
        fprintf(stderr, "[f4la debug] gauss-jordan %u x %u: rank %u, %u batches, %u columns visited\n", nrl, ncr,
                h_state->rank, h_state->batches, h_state->visits);
    *rank_out = h_state->rank;
    return F4LA_OK;
}

/* Tall trailing block: eliminate on kc random combinations of its rows, then verify exactly that every
 * row of the block lies in the span of the result (see k_ge_compress / k_ge_verify).  *ok = false
 * means "not applicable or not confirmed": the caller eliminates on the block itself, which this
 * function never modifies.  On success the reduced rows are in ctx->D2 (leading dimension ctx->d2_ld). */
constexpr uint32_t kCompressRows = 768, kCompressWeight = 8, kCompressSlack = 48;   /* defaults; $F4LA_B200_COMPRESS_ROWS for tests */

int compress_and_verify(f4la_ctx *ctx, uint32_t *Dm, uint64_t ld, uint32_t nrl, uint32_t ncr, uint32_t p,
                        uint64_t pinv, uint32_t *rank_out, bool *ok)
{
    cudaStream_t s = ctx->stream;
    *ok = false;
    const uint32_t nrows4 = (nrl + 3u) & ~3u;
    if ((size_t)nrows4 * 4 > (size_t)216 * 1024 || p < 3) return F4LA_OK;
    const uint32_t two64 = (uint32_t)(((~0ull) % p + 1) % p);
    for (uint32_t kc = ctx->compress_rows; 2 * kc <= nrl; kc *= 4) {
        const uint32_t slack = kc >= 8 * kCompressSlack ? kCompressSlack : kc / 8;
        /* members of every combination: row r joins one combination of each family (hash of seed, family, row) */
        const uint32_t kw = kc / kCompressWeight;
        const size_t nmem = (size_t)nrl * kCompressWeight;
        OK(ensure_host(ctx, ctx->h_combo, ((size_t)kc + 1 + 2 * nmem) * 4 + 64));
        uint32_t *h_ptr = (uint32_t *)ctx->h_combo.p, *h_row = h_ptr + kc + 1, *h_rho = h_row + nmem;
        std::vector<uint32_t> slot(nmem), cnt(kc + 1, 0);
        for (uint32_t r = 0; r < nrl; ++r)
            for (uint32_t f = 0; f < kCompressWeight; ++f) {
                const uint64_t h = host_mix64(ctx->seed ^ 0x5bd1e995u ^ ((uint64_t)f << 40) ^ ((uint64_t)kc << 48) ^ r);
                const uint32_t q = f * kw + (uint32_t)(h % kw);
                slot[(size_t)r * kCompressWeight + f] = q;
                cnt[q + 1]++;
            }
        for (uint32_t q = 0; q < kc; ++q) cnt[q + 1] += cnt[q];
        memcpy(h_ptr, cnt.data(), ((size_t)kc + 1) * 4);
        for (uint32_t r = 0; r < nrl; ++r)
            for (uint32_t f = 0; f < kCompressWeight; ++f) {
                const uint64_t h = host_mix64(ctx->seed ^ 0x5bd1e995u ^ ((uint64_t)f << 40) ^ ((uint64_t)kc << 48) ^ r);
                const uint32_t e = cnt[slot[(size_t)r * kCompressWeight + f]]++;
                h_row[e] = r;
                h_rho[e] = 1u + (uint32_t)((h >> 20) % (p - 1));
            }
        OK(ensure(ctx, ctx->combo, ((size_t)kc + 1 + 2 * nmem) * 4));
        OK(ensure(ctx, ctx->D2, (size_t)ncr * kc * 4));
        uint32_t *d_ptr = as<uint32_t>(ctx->combo), *d_row = d_ptr + kc + 1, *d_rho = d_row + nmem;
        CU(cudaMemcpyAsync(d_ptr, h_ptr, ((size_t)kc + 1 + 2 * nmem) * 4, cudaMemcpyHostToDevice, s));
        if ((size_t)nrows4 * 4 > ctx->compress_smem_optin) {
            CU(cudaFuncSetAttribute(k_ge_compress, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
            ctx->compress_smem_optin = (size_t)216 * 1024;
        }
        uint32_t *D2 = as<uint32_t>(ctx->D2);
        k_ge_compress<<<ncr, 256, (size_t)nrows4 * 4, s>>>(Dm, ld, nrl, nrows4, kc, d_ptr, d_row, d_rho, D2, p, pinv, two64);
        note_launch(ctx, 1, __LINE__);
        uint32_t rank = 0;
        OK(gauss_jordan(ctx, D2, kc, kc, ncr, p, pinv, &rank));
        if (rank + slack > kc) { ctx->compress_escalated++; continue; }   /* too close to the number of combinations */
        OK(ensure(ctx, ctx->dense_out, (size_t)(rank ? rank : 1) * ncr * 4 + 16));
        if (rank) {
            dim3 grid(rank, (ncr + 255) / 256);
            k_ge_extract<<<grid, 256, 0, s>>>(D2, kc, ncr, rank, as<uint32_t>(ctx->pivcol), as<uint32_t>(ctx->pivrow),
                                              as<uint32_t>(ctx->pivinv), as<uint8_t>(ctx->ispiv),
                                              as<uint32_t>(ctx->dense_out), p, pinv, nullptr);
            note_launch(ctx, 1, __LINE__);
        }
        uint32_t *d_bad = as<uint32_t>(ctx->flags32) + 8;
        CU(cudaMemsetAsync(d_bad, 0, 4, s));
        if (ctx->tc_verify && rank >= 1 && rank <= kTcMaxK && ld % kTcBM == 0) {
            /* D' == D'[:, pivot columns] * B on the tensor cores (u8-limb tcgen05.mma, f4la_tc.cuh); B row t
             * belongs to pivot rank-1-t.  ld is a multiple of 128, the rows beyond nrl are zero. */
            TcGemmArgs g;
            memset(&g, 0, sizeof(g));
            g.A = Dm; g.lda = ld; g.acol = as<uint32_t>(ctx->pivcol); g.acol_reverse = 1;
            g.B = as<uint32_t>(ctx->dense_out); g.ldb = ncr;
            g.M = (nrl + kTcBM - 1) / kTcBM * kTcBM; g.N = ncr; g.K = rank; g.p = p;
            g.D = Dm; g.ldd = ld; g.mismatch = d_bad; g.err = as<uint32_t>(ctx->flags32);
            OK(launch_tc_gemm(ctx, g));
        } else {
            const uint32_t nrows_pad = (nrl + kVerTile - 1) / kVerTile * kVerTile;     /* <= ld, padding rows are zero */
            dim3 vgrid(nrows_pad / kVerTile, (ncr + kVerTile - 1) / kVerTile);
            k_ge_verify<<<vgrid, 256, 0, s>>>(Dm, ld, nrows_pad, ncr, rank, as<uint32_t>(ctx->pivcol),
                                              as<uint32_t>(ctx->dense_out), p, pinv, two64, d_bad);
            note_launch(ctx, 1, __LINE__);
        }
        uint32_t *h_bad = (uint32_t *)ctx->h_small.p + 128;
        CU(cudaMemcpyAsync(h_bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaGetLastError());
        OK(stream_sync(ctx));
        ctx->compress_attempts++;
        if (*h_bad || ctx->compress_fault) { ctx->compress_rejected++; return F4LA_OK; }      /* rank lost in the compression: eliminate directly */
        ctx->d2_ld = kc;
        ctx->donor.assign(kc, 0xffffffffu);
        for (uint32_t q = 0; q < kc; ++q)
            if (h_ptr[q + 1] > h_ptr[q]) ctx->donor[q] = h_row[h_ptr[q]];
        *rank_out = rank;
        *ok = true;
        return F4LA_OK;
    }
    return F4LA_OK;
}

/* The whole pipeline on a device-resident matrix. */
int run_pipeline(f4la_ctx *ctx, const f4la_resident *m, f4la_flat_result *out, f4la_stats *st,
                 const f4la_replay_call *ax = nullptr)
{
    const f4la_flat_matrix &d = m->dims;
    const uint32_t nru = d.nru, nrl_in = d.nrl, p = d.fc;
    uint32_t nrl = nrl_in;          /* rows the solve / elimination work on: the lower rows, or in
                                       probabilistic mode the random combinations standing in for them */
    const uint32_t nc = d.ncl + d.ncr;
    const bool by_lead = (d.flags & F4LA_FLAG_PIVOT_BY_LEAD) != 0;
    const bool nf_mode = (d.flags & F4LA_FLAG_NORMALFORM) != 0;
    /* effective split: with pivots keyed by lead column the columns are renumbered
     * (pivot columns first), so that ncl == nru holds for the rest of the pipeline */
    /* Pivots keyed by lead column (final reduction step, interreduce_matrix_rows): the lower rows are
     * pivots too (distinct leads, monic), so ALL rows form the unit upper triangular system
     * [A B] and the wanted rows are rows of its reduced echelon form [I  A^-1 B].  Row i of A^-1 B
     * is -D' of the unit vector e_lead(i) pushed through the same pull solve: no elimination, no
     * pivot search (the former version ran a Gauss-Jordan with one pivot per basis element). */
    const uint32_t npr = by_lead ? nru + nrl_in : nru;      /* rows that act as pivots */
    const uint32_t ncl = by_lead ? npr : d.ncl;
    if (ncl > nc)
        return fail(ctx, F4LA_ERR_ARG, "more pivot rows (%u) than columns (%u)", ncl, nc);
    const uint32_t ncr = nc - ncl;
    const uint64_t pinv = ~0ULL / p;
    const bool wide = p >= 65536u;
    cudaStream_t s = ctx->stream;
    const uint32_t launches0 = ctx->launches;

    const bool want_rba = (d.flags & F4LA_FLAG_WANT_RBA) != 0;
    if (want_rba && by_lead)
        return fail(ctx, F4LA_ERR_ARG, "F4LA_FLAG_WANT_RBA needs pivots keyed by row index");
    if (d.flags & ~(F4LA_FLAG_PIVOT_BY_LEAD | F4LA_FLAG_NORMALFORM | F4LA_FLAG_WANT_RBA | F4LA_FLAG_PROBABILISTIC))
        return fail(ctx, F4LA_ERR_ARG, "unknown flags 0x%x", d.flags);
    if (by_lead && nf_mode)
        return fail(ctx, F4LA_ERR_ARG, "F4LA_FLAG_PIVOT_BY_LEAD and F4LA_FLAG_NORMALFORM cannot be combined");
    if (!by_lead && nru != d.ncl)
        return fail(ctx, F4LA_ERR_ARG, "pivots keyed by row index need nru == ncl (got %u, %u)", nru, d.ncl);
    if (m->nnz >= 0xfffffff0ull)
        return fail(ctx, F4LA_ERR_ARG, "matrix too large: %llu non-zeros", (unsigned long long)m->nnz);
    if (ax && (nrl == 0 || ncr == 0))
        return fail(ctx, F4LA_ERR_ARG, "f4la_b200_replay_call: degenerate matrix (no lower rows or no column without pivot)");
    if (nrl == 0 || (ncr == 0 && !by_lead && !nf_mode)) { empty_result(out, d.cf_bytes); return F4LA_OK; }
    if (ncr == 0 && !by_lead) {           /* every lower row reduces to zero */
        empty_result(out, d.cf_bytes);
        free(out->row_ptr); free(out->src_row);
        out->nrows = nrl;
        out->row_ptr = (uint64_t *)calloc((size_t)nrl + 1, sizeof(uint64_t));
        out->src_row = (uint32_t *)malloc((size_t)nrl * sizeof(uint32_t));
        for (uint32_t i = 0; i < nrl; ++i) out->src_row[i] = i;
        return F4LA_OK;
    }

    const bool timed = st != nullptr || ctx->debug;
    if (timed) CU(cudaEventRecord(ctx->ev[0], s));

    OK(ensure(ctx, ctx->flags32, 64 + (size_t)kRefSlots * 32));     /* error/flag words | unit counters: one memset, one read-back */
    OK(ensure_host(ctx, ctx->h_small, 1 << 20));
    uint32_t *h_small = (uint32_t *)ctx->h_small.p;
    uint32_t *d_err = as<uint32_t>(ctx->flags32);           /* [0] err, [1] changed, [2] max level */
    unsigned long long *d_ref = reinterpret_cast<unsigned long long *>(d_err + 16);
    CU(cudaMemsetAsync(d_err, 0, 64 + (size_t)kRefSlots * 32, s));
    const uint32_t *colmap = nullptr;
    PrepCache *pc = m->cache_enabled ? &m->cache : nullptr;
    const bool cached = pc && pc->valid;
    if (ax && (!cached || want_rba || nf_mode || (d.flags & F4LA_FLAG_PROBABILISTIC) || d.cf_bytes != 4))
        return fail(ctx, F4LA_ERR_ARG, "f4la_b200_replay_call needs a resident matrix with a filled layout cache (one ordinary "
                    "run after f4la_b200_resident_update), 32-bit coefficients, and no trace / normal-form / probabilistic flag");
    uint32_t nnz_pull = 0, nnz_schur = 0, nlev = 0;
    uint32_t *d_col_ptr = nullptr, *d_invmap = nullptr, *d_sched = nullptr, *d_items = nullptr;
    uint2 *d_ent = nullptr;
    std::vector<uint32_t> slot_ptr_h;
    if (cached) {
        /* same structure as the run that filled the cache: only the values of the pull lists change */
        nnz_pull = pc->nnz_pull; nnz_schur = pc->nnz_schur;
        d_col_ptr = pc->col_ptr; d_ent = pc->ent; d_invmap = pc->invmap;
        if (by_lead) colmap = pc->colmap;
        if (nnz_pull) {
            k_ent_refill<<<(nnz_pull + 255) / 256, 256, 0, s>>>(d_ent, pc->ent_src, m->cf, d.cf_bytes, p, nnz_pull);
            note_launch(ctx, 1, __LINE__);
        }
    } else {
    /* ---- 1. pull lists (CSC of the strictly upper part of the pivot rows) ---- */
    OK(ensure(ctx, ctx->col_cnt, ((size_t)nc + 2) * 4));
    OK(ensure(ctx, ctx->col_ptr, ((size_t)nc + 2) * 4));
    CU(cudaMemsetAsync(ctx->col_cnt.p, 0, ((size_t)nc + 2) * 4, s));
    if (by_lead) {
        OK(ensure(ctx, ctx->is_piv, ((size_t)nc + 2) * 4));
        OK(ensure(ctx, ctx->pividx, ((size_t)nc + 2) * 4));
        OK(ensure(ctx, ctx->colmap, ((size_t)nc + 2) * 4));
        OK(ensure(ctx, ctx->invmap, ((size_t)nc + 2) * 4));
        CU(cudaMemsetAsync(ctx->is_piv.p, 0, ((size_t)nc + 2) * 4, s));
        if (npr) k_mark_leads<<<(npr + 255) / 256, 256, 0, s>>>(m->row_ptr, m->cols, npr, nc, as<uint32_t>(ctx->is_piv), d_err);
        k_exclusive_scan<<<1, 1024, 0, s>>>(as<uint32_t>(ctx->is_piv), as<uint32_t>(ctx->pividx), nc);
        k_build_colmap<<<(nc + 255) / 256, 256, 0, s>>>(as<uint32_t>(ctx->is_piv), as<uint32_t>(ctx->pividx), nc, npr,
                                                       as<uint32_t>(ctx->colmap), as<uint32_t>(ctx->invmap));
        note_launch(ctx, 3, __LINE__);
        colmap = as<uint32_t>(ctx->colmap);
    }
    if (npr) {
        k_pull_count<<<blocks_for_warps(npr), 256, 0, s>>>(m->row_ptr, m->cols, m->row_cf, m->cf_ptr, m->cf,
                                                           d.cf_bytes, npr, ncl, nc, colmap,
                                                           as<uint32_t>(ctx->col_cnt), d_err);
        note_launch(ctx, 1, __LINE__);
    }
    k_exclusive_scan<<<1, 1024, 0, s>>>(as<uint32_t>(ctx->col_cnt), as<uint32_t>(ctx->col_ptr), nc);
    note_launch(ctx, 1, __LINE__);
    CU(cudaMemcpyAsync(h_small, as<uint32_t>(ctx->col_ptr) + nc, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_small + 1, d_err, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_small + 2, as<uint32_t>(ctx->col_ptr) + ncl, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaGetLastError());
    OK(stream_sync(ctx));
    nnz_pull = h_small[0];
    nnz_schur = nnz_pull - h_small[2];                       /* pull-list entries of the columns without pivot */
    if (h_small[1])
        return fail(ctx, F4LA_ERR_ARG, "malformed pivot rows (error bits 0x%x: 1 lead not first, 2 not upper "
                    "triangular, 4 column out of range, 8 lead coefficient != 1, 32 duplicate pivot column)", h_small[1]);
    OK(ensure(ctx, ctx->ent, ((size_t)nnz_pull + 1) * sizeof(uint2)));
    if (pc) OK(cache_alloc(ctx, (void **)&pc->ent_src, ((size_t)nnz_pull + 1) * 4));
    CU(cudaMemsetAsync(ctx->col_cnt.p, 0, ((size_t)nc + 2) * 4, s));   /* reuse as cursor */
    if (npr) {
        k_pull_fill<<<blocks_for_warps(npr), 256, 0, s>>>(m->row_ptr, m->cols, m->row_cf, m->cf_ptr, m->cf, d.cf_bytes,
                                                          npr, nc, p, colmap, as<uint32_t>(ctx->col_ptr),
                                                          as<uint32_t>(ctx->col_cnt), as<uint2>(ctx->ent),
                                                          pc ? pc->ent_src : nullptr);
        note_launch(ctx, 1, __LINE__);
    }

    /* ---- 2. levels of the pivot columns ---- */
    if (ncl) {
        OK(ensure(ctx, ctx->level, (size_t)ncl * 4));
        CU(cudaMemsetAsync(ctx->level.p, 0xff, (size_t)ncl * 4, s));      /* kLevelUnset */
        CU(cudaMemsetAsync(d_err + 1, 0, 4, s));
        {
            unsigned g = (ncl + 7) / 8;
            const unsigned gmax = (unsigned)ctx->sm_count * 8;
            if (g > gmax) g = gmax;
            k_level_flow<<<g, 256, 0, s>>>(as<uint32_t>(ctx->col_ptr), as<uint2>(ctx->ent), ncl,
                                           as<uint32_t>(ctx->level), d_err + 1, d_err);
            note_launch(ctx, 1, __LINE__);
        }
        /* class histogram: slot 2l = long columns of level l, 2l+1 = short ones */
        OK(ensure(ctx, ctx->slots, ((size_t)kLenClasses * ncl + 8) * 4));
        OK(ensure(ctx, ctx->slot_ptr, ((size_t)kLenClasses * ncl + 8) * 4));
        OK(ensure(ctx, ctx->order, (size_t)ncl * 4));
        CU(cudaMemsetAsync(ctx->slots.p, 0, ((size_t)kLenClasses * ncl + 8) * 4, s));
        k_level_count<<<(ncl + 255) / 256, 256, 0, s>>>(as<uint32_t>(ctx->level), as<uint32_t>(ctx->col_ptr), ncl,
                                                        as<uint32_t>(ctx->slots), d_err + 2);
        note_launch(ctx, 1, __LINE__);
        CU(cudaMemcpyAsync(h_small, d_err + 2, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaGetLastError());
    OK(stream_sync(ctx));
        nlev = h_small[0] + 1;
        const uint32_t nslots = kLenClasses * nlev;
        k_exclusive_scan<<<1, 1024, 0, s>>>(as<uint32_t>(ctx->slots), as<uint32_t>(ctx->slot_ptr), nslots);
        note_launch(ctx, 1, __LINE__);
        CU(cudaMemsetAsync(ctx->slots.p, 0, ((size_t)nslots + 1) * 4, s));  /* reuse as cursor */
        k_level_fill<<<(ncl + 255) / 256, 256, 0, s>>>(as<uint32_t>(ctx->level), as<uint32_t>(ctx->col_ptr), ncl,
                                                       as<uint32_t>(ctx->slot_ptr), as<uint32_t>(ctx->slots),
                                                       as<uint32_t>(ctx->order));
        note_launch(ctx, 1, __LINE__);
        slot_ptr_h.resize(nslots + 1);
        OK(ensure_host(ctx, ctx->h_small, ((size_t)nslots + 16) * 4));
        h_small = (uint32_t *)ctx->h_small.p;
        CU(cudaMemcpyAsync(h_small, ctx->slot_ptr.p, ((size_t)nslots + 1) * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaGetLastError());
    OK(stream_sync(ctx));
        memcpy(slot_ptr_h.data(), h_small, ((size_t)nslots + 1) * 4);
    }
    d_col_ptr = as<uint32_t>(ctx->col_ptr); d_ent = as<uint2>(ctx->ent);
    if (by_lead) d_invmap = as<uint32_t>(ctx->invmap);
    }   /* !cached */
    if (timed) CU(cudaEventRecord(ctx->ev[1], s));

    /* Probabilistic mode (the reference's -l 42/43/44, la_ff_32.c:2305-2506): the nrl lower rows
     * are replaced by k << nrl random linear combinations of them; when the echelon form of the
     * combinations has a rank well below k they span the same space (w.h.p., p > 2^15) and the
     * reduced rows are the same unique ones.  Every lower row enters kProbWeight combinations (one
     * per family, so no row can be missed); k starts at 512 and is quadrupled while the rank comes
     * within kProbSlack of it; k >= nrl/2 switches to the exact path. */
    constexpr uint32_t kProbWeight = 8, kProbSlack = 48;
    const bool prob = (d.flags & F4LA_FLAG_PROBABILISTIC) && !by_lead && !nf_mode && !want_rba && p > 32768u;
    uint32_t nvr = nrl_in;
    if (prob && nrl_in >= 3 * 512) nvr = 512;
    uint32_t rank = 0, Rc = 0, nchunks = 0, G = 0, n_items = 0, maxrank = 0;
    uint64_t ld = 0;
    int T = 1, n_flow_ev = 0;
    bool items_built = false, schur = false, schur_b_built = false, schur_side_pending = false;
    uint32_t *Dm = nullptr;
    uint64_t ge_ld = 0;               /* leading dimension of the block the reduced rows are read from */
    uint32_t echelon_path = 0;
    uint32_t schur_nseg = 1;
    bool rows_are_combinations = false;   /* the block that was eliminated holds random combinations, not lower rows */
    const uint32_t rba_words = (nru + 31) / 32;

    for (;;) {
    nrl = nvr;
    const bool combos = nvr < nrl_in;
    /* ---- 3. chunking of the lower rows ---- */
    T = (int)((nrl + kRowsPerT - 1) / kRowsPerT);
    if (T > ctx->tmax) T = ctx->tmax;
    const size_t kL2Budget = ctx->l2_budget;
    while (T > 1 && (size_t)nc * kRowsPerT * T * 4 > kL2Budget) --T;
    Rc = kRowsPerT * T;
    nchunks = (nrl + Rc - 1) / Rc;
    G = (uint32_t)(kL2Budget / ((size_t)nc * Rc * 4));
    if (G < 1) G = 1;
    if (G > nchunks) G = nchunks;
    ld = (uint64_t)nchunks * Rc;
    if ((uint64_t)nc * Rc >= (1ull << 32))
        return fail(ctx, F4LA_ERR_ARG, "matrix too wide: %u columns x %u rows per chunk", nc, Rc);
    OK(ensure(ctx, ctx->V, (size_t)G * nc * Rc * 4));
    OK(ensure(ctx, ctx->D, (size_t)ncr * ld * 4));
    OK(ensure(ctx, ctx->flowflags, (size_t)G * nc * 4));
    if (!ncl) CU(cudaMemsetAsync(ctx->flowflags.p, 0, (size_t)G * nc * 4, s));   /* else: written by k_flow_reset_flags */

    /* The columns without pivot have no dependents: V[ncl + j] = X + M * Bneg is a dense-times-sparse product.
     * When it is large it runs on the tensor cores over the non-zero 64 x 64 tiles of B (f4la_tc.cuh), and the
     * dataflow kernel only solves the pivot columns. */
    if (!items_built && cached) {
        items_built = true;
        n_items = pc->n_items; schur = pc->schur;
        d_sched = pc->sched; d_items = pc->items;
    }
    if (!items_built)
        schur = ctx->schur_tc && ncl >= 1 && ncr >= 1 && nnz_schur >= 1 && (ncl + kTcBK - 1) / kTcBK < 65536 &&
                (double)nnz_schur * nrl >= ctx->schur_min_macs &&
                (ctx->schur_min_macs == 0 || (ncl >= 1024 && ncr >= 64 && nrl >= 256));   /* F4LA_B200_SCHUR_MIN=0 (tests): always */
    /* dataflow schedule: items in (level, length class) order, then the non-pivot columns */
    if (!items_built) {
        items_built = true;
        const uint32_t first = nlev ? slot_ptr_h[kLenClasses] : ncl;      /* columns of level 0 have no sources */
        const uint32_t na = ncl - first;
        std::vector<uint32_t> items;
        items.reserve((size_t)na / 4 + ncr + 16);
        for (uint32_t l = 1; l < nlev; ++l) {
            const uint32_t *sp = &slot_ptr_h[(size_t)kLenClasses * l];
            for (uint32_t q = sp[0]; q < sp[1]; ++q)                       /* long columns: one CTA each */
                items.push_back(((q - first) << 5) | (1u << 1) | 1u);
            for (int c = 1; c < kLenClasses; ++c)                          /* short ones: 8 of a class per CTA */
                for (uint32_t q = sp[c]; q < sp[c + 1]; q += kWarpsPerBlock) {
                    const uint32_t cnt = sp[c + 1] - q < (uint32_t)kWarpsPerBlock ? sp[c + 1] - q : (uint32_t)kWarpsPerBlock;
                    items.push_back(((q - first) << 5) | (cnt << 1));
                }
        }
        if (!schur)
            for (uint32_t k = 0; k < ncr; ++k) items.push_back(((na + k) << 5) | (1u << 1) | 1u);
        n_items = (uint32_t)items.size();
        OK(ensure(ctx, ctx->sched, ((size_t)na + ncr + 1) * 4));
        OK(ensure(ctx, ctx->items, ((size_t)n_items + 1) * 4));
        OK(ensure_host(ctx, ctx->h_dense, (size_t)n_items * 4 + 64));
        memcpy(ctx->h_dense.p, items.data(), (size_t)n_items * 4);
        CU(cudaMemcpyAsync(ctx->items.p, ctx->h_dense.p, (size_t)n_items * 4, cudaMemcpyHostToDevice, s));
        k_flow_sched<<<(na + ncr + 255) / 256, 256, 0, s>>>(as<uint32_t>(ctx->order), first, ncl, ncr,
                                                           as<uint32_t>(ctx->sched));
        note_launch(ctx, 1, __LINE__);
        d_sched = as<uint32_t>(ctx->sched); d_items = as<uint32_t>(ctx->items);
        if (pc) {
            /* keep everything derived from the structure with the resident matrix */
            struct { void **dst; const void *src; size_t bytes; } keep[] = {
                {(void **)&pc->col_ptr, d_col_ptr, ((size_t)nc + 2) * 4},
                {(void **)&pc->ent, d_ent, ((size_t)nnz_pull + 1) * sizeof(uint2)},
                {(void **)&pc->sched, d_sched, ((size_t)na + ncr + 1) * 4},
                {(void **)&pc->items, d_items, ((size_t)n_items + 1) * 4},
                {(void **)&pc->colmap, by_lead ? ctx->colmap.p : nullptr, ((size_t)nc + 2) * 4},
                {(void **)&pc->invmap, by_lead ? ctx->invmap.p : nullptr, ((size_t)nc + 2) * 4},
            };
            for (auto &k : keep) {
                if (!k.src) continue;
                OK(cache_alloc(ctx, k.dst, k.bytes));
                CU(cudaMemcpyAsync(*k.dst, k.src, k.bytes, cudaMemcpyDeviceToDevice, s));
            }
            pc->nnz_pull = nnz_pull; pc->nnz_schur = nnz_schur; pc->n_items = n_items; pc->schur = schur;
            pc->valid = true;
        }
        /* the staging buffer is reused for the result later: the copy must be done first */
        OK(stream_sync(ctx));
    }

    if (want_rba && rba_words) OK(ensure(ctx, ctx->rba, (size_t)nrl * rba_words * 4 + 16));

    if (schur && !schur_b_built) {
        /* The tiles of B (stage images of the non-pivot pull lists) depend on the pivot rows only: they are built once
         * per matrix, on the side stream, while the main stream scatters the lower rows and runs the dataflow kernel. */
        schur_b_built = true;
        SchurArgs sb;
        memset(&sb, 0, sizeof(sb));
        sb.nc = nc; sb.ncl = ncl; sb.ncr = ncr; sb.col_ptr = d_col_ptr; sb.ent = d_ent;
        sb.nblk = (ncr + kTcBN - 1) / kTcBN; sb.nkb = (ncl + kTcBK - 1) / kTcBK;
        OK(ensure(ctx, ctx->scB, (size_t)sb.nblk * sb.nkb * kTcStageB));
        OK(ensure(ctx, ctx->scFlagB, (size_t)sb.nblk * sb.nkb + 16));
        OK(ensure(ctx, ctx->scList, (size_t)sb.nblk * sb.nkb * 4 + 16));
        OK(ensure(ctx, ctx->scCnt, (size_t)sb.nblk * 4 + 16));
        OK(ensure_host(ctx, ctx->h_small, ((size_t)sb.nblk + 1024) * 4));
        sb.Bsplit = as<uint8_t>(ctx->scB); sb.flagB = as<uint8_t>(ctx->scFlagB);
        sb.list = as<uint32_t>(ctx->scList); sb.list_cnt = as<uint32_t>(ctx->scCnt);
        cudaStream_t s2 = ctx->copy_stream;
        CU(cudaEventRecord(ctx->ev_side[0], s));
        CU(cudaStreamWaitEvent(s2, ctx->ev_side[0], 0));
        CU(cudaMemsetAsync(ctx->scB.p, 0, (size_t)sb.nblk * sb.nkb * kTcStageB, s2));
        CU(cudaMemsetAsync(ctx->scFlagB.p, 0, (size_t)sb.nblk * sb.nkb, s2));
        k_schur_b_build<<<blocks_for_warps(ncr), 256, 0, s2>>>(sb);
        k_schur_lists<<<sb.nblk, 256, 0, s2>>>(sb);
        note_launch(ctx, 2, __LINE__);
        if (!(pc && pc->schur_nseg))
            CU(cudaMemcpyAsync((uint32_t *)ctx->h_small.p + 512, ctx->scCnt.p, (size_t)sb.nblk * 4, cudaMemcpyDeviceToHost, s2));
        CU(cudaEventRecord(ctx->ev_side[1], s2));
        schur_side_pending = true;
    }

    uint32_t epoch = 0;
    for (uint32_t c0 = 0; c0 < nchunks; c0 += G) {
        const uint32_t g = (c0 + G <= nchunks) ? G : nchunks - c0;
        CU(cudaMemsetAsync(ctx->V.p, 0, (size_t)g * nc * Rc * 4, s));
        if (ncl) {
            k_flow_reset_flags<<<(nc + 255) / 256, 256, 0, s>>>(d_col_ptr, ncl, nc, g,
                                                               as<uint32_t>(ctx->flowflags), d_err + 4);
            note_launch(ctx, 1, __LINE__);
        }
        if (ctx->copy_pending) {                  /* the lower rows of a split host->device copy */
            CU(cudaStreamWaitEvent(s, ctx->ev_copy, 0));
            ctx->copy_pending = false;
        }
        const uint32_t rows_here = ((c0 + g) * Rc <= nrl ? (c0 + g) * Rc : nrl) - c0 * Rc;
        if (by_lead)
            k_scatter_units<<<(rows_here + 255) / 256, 256, 0, s>>>(m->row_ptr, m->cols, nru, nrl, nc, c0, g, Rc, colmap,
                                                                    as<uint32_t>(ctx->flowflags), as<uint32_t>(ctx->V), d_err);
        else if (combos)
            k_scatter_combos<<<blocks_for_warps((uint64_t)nrl_in * kProbWeight), 256, 0, s>>>(
                m->row_ptr, m->cols, m->row_cf, m->cf_ptr, m->cf, d.cf_bytes, nru, nrl_in, nc, p, pinv, c0, g, Rc,
                nvr, kProbWeight, ctx->seed, as<uint32_t>(ctx->flowflags), as<uint32_t>(ctx->V), d_err);
        else
            k_scatter_lower<<<blocks_for_warps((uint64_t)rows_here * kScatterSplit), 256, 0, s>>>(
                m->row_ptr, m->cols, m->row_cf, m->cf_ptr, m->cf, d.cf_bytes, nru, nrl, nc, p, c0, g, Rc, colmap,
                as<uint32_t>(ctx->flowflags), as<uint32_t>(ctx->V), d_err);
        note_launch(ctx, 1, __LINE__);
        FlowArgs f;
        memset(&f, 0, sizeof(f));
        f.col_ptr = d_col_ptr; f.ent = d_ent;
        f.V = as<uint32_t>(ctx->V); f.nc = nc; f.Rc = Rc;
        f.sched = d_sched; f.items = d_items;
        f.n_items = n_items; f.G = g;
        f.flags = as<uint32_t>(ctx->flowflags); f.epoch = ++epoch;
        f.counter = d_err + 4; f.err = d_err;
        f.ncl = ncl; f.out = as<uint32_t>(ctx->D); f.out_ld = ld; f.chunk0 = c0;
        f.p = p; f.pinv = pinv;
        f.two64 = (uint32_t)(((~0ull) % p + 1) % p);
        f.watchdog_ns = (uint64_t)(ctx->watchdog_ms * 1e6);
        /* reference-equivalent unit count: pivots keyed by row index, true lower rows only */
        f.piv_row_ptr = (!by_lead && !combos && !getenv("F4LA_B200_NOUNITS")) ? m->row_ptr : nullptr;
        f.ref_units = d_ref;
        if (!ncl) CU(cudaMemsetAsync(d_err + 4, 0, 4, s));      /* else: zeroed by k_flow_reset_flags */
        const bool ev_ok = st && n_flow_ev + 3 <= kMaxFlowEvents;
        if (ev_ok) CU(cudaEventRecord(ctx->flow_ev[n_flow_ev], s));
        launch_flow(ctx, T, f, wide);
        if (ev_ok) CU(cudaEventRecord(ctx->flow_ev[n_flow_ev + 1], s));
        if (schur) {
            SchurArgs sa;
            memset(&sa, 0, sizeof(sa));
            sa.V = as<uint32_t>(ctx->V); sa.nc = nc; sa.ncl = ncl; sa.ncr = ncr; sa.Rc = Rc; sa.G = g;
            sa.col_ptr = d_col_ptr; sa.ent = d_ent;
            sa.mblk = g * (Rc / kTcBM); sa.nblk = (ncr + kTcBN - 1) / kTcBN; sa.nkb = (ncl + kTcBK - 1) / kTcBK;
            sa.ldp = (uint64_t)sa.mblk * kTcBM;
            sa.out = as<uint32_t>(ctx->D); sa.out_ld = ld; sa.chunk0 = c0;
            sa.p = p; sa.w32 = (uint32_t)((1ull << 32) % p); sa.pinv = pinv; sa.err = d_err;
            OK(ensure(ctx, ctx->scB, (size_t)sa.nblk * sa.nkb * kTcStageB));
            OK(ensure(ctx, ctx->scFlagB, (size_t)sa.nblk * sa.nkb + 16));
            OK(ensure(ctx, ctx->scList, (size_t)sa.nblk * sa.nkb * 4 + 16));
            OK(ensure(ctx, ctx->scCnt, (size_t)sa.nblk * 4 + 16));
            OK(ensure(ctx, ctx->tcA, (size_t)sa.mblk * sa.nkb * kTcStageA));
            OK(ensure(ctx, ctx->scFlagA, (size_t)sa.mblk * sa.nkb + 16));
            sa.Asplit = as<uint8_t>(ctx->tcA); sa.Bsplit = as<uint8_t>(ctx->scB);
            sa.flagA = as<uint8_t>(ctx->scFlagA); sa.flagB = as<uint8_t>(ctx->scFlagB);
            sa.list = as<uint32_t>(ctx->scList); sa.list_cnt = as<uint32_t>(ctx->scCnt);
            if (schur_side_pending) {
                /* join the side stream that built the tiles of B while the dataflow kernel ran */
                schur_side_pending = false;
                if (!(pc && pc->schur_nseg)) {
                    CU(cudaStreamSynchronize(ctx->copy_stream));
                    const uint32_t *h_cnt = (const uint32_t *)ctx->h_small.p + 512;
                    uint32_t mx = 0;
                    for (uint32_t k = 0; k < sa.nblk; ++k) mx = h_cnt[k] > mx ? h_cnt[k] : mx;
                    schur_nseg = (mx + kSchurSeg - 1) / kSchurSeg;
                    if (schur_nseg < 1) schur_nseg = 1;
                    if (pc) pc->schur_nseg = schur_nseg;
                } else {
                    schur_nseg = pc->schur_nseg;              /* the tile lists only depend on the structure */
                }
                CU(cudaStreamWaitEvent(s, ctx->ev_side[1], 0));
            }
            sa.nseg = schur_nseg;
            OK(ensure(ctx, ctx->scP, (size_t)sa.nseg * ncr * sa.ldp * 4));
            sa.P = as<uint32_t>(ctx->scP);
            if (!ctx->schur_attr_set) {
                CU(cudaFuncSetAttribute(k_schur_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
                ctx->schur_attr_set = true;
            }
            k_schur_a_split<<<dim3(sa.mblk, sa.nkb), 128, 0, s>>>(sa);
            k_schur_gemm<<<dim3(sa.mblk, sa.nblk, sa.nseg), kTcThreads, kTcSmemBytes, s>>>(sa);
            k_schur_combine<<<dim3(ncr, (g * Rc + 255) / 256), 256, 0, s>>>(sa);
            note_launch(ctx, 3, __LINE__);
        }
        /* "hot kernel" time = the reduction by the known pivots: dataflow solve + (when used) the tensor-core Schur product */
        if (ev_ok) { CU(cudaEventRecord(ctx->flow_ev[n_flow_ev + 2], s)); n_flow_ev += 3; }
        if (want_rba && rba_words) {
            dim3 grid(rba_words, g * (Rc / 128));
            k_rba_bits<<<grid, 128, 0, s>>>(as<uint32_t>(ctx->V), nc, Rc, ncl, c0, nrl, rba_words,
                                            as<uint32_t>(ctx->rba));
            note_launch(ctx, 1, __LINE__);
        }
    }
    if (timed) CU(cudaEventRecord(ctx->ev[2], s));

    /* ---- 4. Gauss-Jordan on the trailing block (not in normal-form mode) ---- */
    maxrank = nrl < ncr ? nrl : ncr;
    Dm = as<uint32_t>(ctx->D);
    rank = 0;
    if (!nf_mode && !by_lead) {
        bool compressed = false;
        if (ax && ctx->ge_compress && nrl >= 2 * ctx->compress_rows)
            return fail(ctx, F4LA_ERR_ARG, "a replay without host round trips does not cover compress-and-verify blocks");
        if (!want_rba && !combos && ctx->ge_compress && nrl >= 2 * ctx->compress_rows) {
            const uint64_t rej0 = ctx->compress_rejected, esc0 = ctx->compress_escalated;
            int rc = compress_and_verify(ctx, Dm, ld, nrl, ncr, p, pinv, &rank, &compressed);
            if (rc != F4LA_OK) return rc;
            echelon_path = (compressed ? 1u : 0u) | (ctx->compress_rejected != rej0 ? 2u : 0u) |
                           (ctx->compress_escalated != esc0 ? 4u : 0u);
        }
        ge_ld = ld;
        rows_are_combinations = compressed || combos;
        if (combos) {                     /* same hash as k_scatter_combos: first lower row of every combination */
            const uint32_t kw = nvr / kProbWeight;
            ctx->donor.assign(nvr, 0xffffffffu);
            for (uint32_t r = nrl_in; r-- > 0;)
                for (uint32_t f = 0; f < kProbWeight; ++f) {
                    const uint64_t h = host_mix64(ctx->seed ^ ((uint64_t)f << 40) ^ r);
                    ctx->donor[f * kw + (uint32_t)(h % kw)] = r;
                }
        }
        if (compressed) { Dm = as<uint32_t>(ctx->D2); ge_ld = ctx->d2_ld; }
        else OK(gauss_jordan(ctx, Dm, ld, nrl, ncr, p, pinv, &rank, ax ? &ax->expect_rows : nullptr));
    }
    /* probabilistic mode: accept when the rank stays clear of the number of combinations */
    if (!combos || rank + kProbSlack <= nvr) break;
    nvr = (uint64_t)nvr * 4 >= nrl_in / 2 ? nrl_in : nvr * 4;
    }   /* attempts */
    uint64_t ref_units = 0, ref_apps = 0;
    if (ax) {
        /* ---- 5'. replay: dense rows stay on the device, coefficients go to the arena at the recorded support ---- */
        const uint32_t nrows_x = ax->expect_rows;
        if (by_lead && nrows_x != nrl)
            return fail(ctx, F4LA_ERR_ARG, "replay: %u rows expected, the matrix has %u lower rows", nrows_x, nrl);
        OK(ensure(ctx, ctx->dense_out, (size_t)(nrows_x ? nrows_x : 1) * (ncr ? ncr : 1) * 4 + 16));
        if (by_lead && ncr && nrows_x) {
            dim3 grid((ncr + 31) / 32, (nrl + 31) / 32), block(32, 8);
            k_transpose_rows<<<grid, block, 0, s>>>(Dm, ld, ncr, nrl, p, as<uint32_t>(ctx->dense_out));
            note_launch(ctx, 1, __LINE__);
        } else if (!by_lead && nrows_x) {
            dim3 grid(nrows_x, (ncr + 255) / 256);
            k_ge_extract<<<grid, 256, 0, s>>>(Dm, ge_ld, ncr, nrows_x, as<uint32_t>(ctx->pivcol), as<uint32_t>(ctx->pivrow),
                                              as<uint32_t>(ctx->pivinv), as<uint8_t>(ctx->ispiv),
                                              as<uint32_t>(ctx->dense_out), p, pinv, as<GEState>(ctx->gestate));
            note_launch(ctx, 1, __LINE__);
        }
        unsigned blocks = (nrows_x + 7) / 8;
        if (blocks < 1) blocks = 1;
        k_replay_compact<<<blocks, 256, 0, s>>>(as<uint32_t>(ctx->dense_out), ncr, nrows_x, ax->row_ptr, ax->idx, ax->row_dst,
                                                ax->arena, by_lead ? nullptr : as<GEState>(ctx->gestate), d_err, ax->status);
        note_launch(ctx, 1, __LINE__);
        CU(cudaGetLastError());
        return F4LA_OK;
    }
    {
        uint32_t *h_err = (uint32_t *)ctx->h_small.p + 64;
        uint64_t *h_ref = (uint64_t *)(h_err + 16);
        CU(cudaMemcpyAsync(h_err, d_err, 64 + (size_t)kRefSlots * 32, cudaMemcpyDeviceToHost, s));
        OK(stream_sync(ctx));
        for (uint32_t k = 0; k < kRefSlots; ++k) { ref_units += h_ref[4 * k]; ref_apps += h_ref[4 * k + 1]; }
        if (*h_err)
            return fail(ctx, F4LA_ERR_CUDA, "device pipeline reported error bits 0x%x (4: column out of range in a lower "
                        "row, 16: dataflow watchdog expired)", *h_err);
    }
    if (timed) CU(cudaEventRecord(ctx->ev[3], s));

    /* ---- 5. result rows back to the host ---- */
    const uint32_t ndense = (nf_mode || by_lead) ? nrl : rank;       /* dense rows produced on the device */
    const uint32_t nout = (nf_mode || by_lead) ? nrl : rank;
    uint64_t d2h_bytes = 0;
    double dbg_copy_ms = 0, dbg_t1 = now_ms();
    if (nout == 0) {
        empty_result(out, d.cf_bytes);
    } else {
        const size_t dense_bytes = (size_t)ndense * ncr * 4;
        const size_t aux_words = 2 * (size_t)rank + (by_lead ? (size_t)nc + nrl : 0) + 16;
        OK(ensure(ctx, ctx->dense_out, dense_bytes + 16));
        OK(ensure_host(ctx, ctx->h_dense, dense_bytes + aux_words * 4));
        /* The dense result rows are written by the extraction kernel straight into the page-locked host buffer
         * (it is mapped into the device's address space): 128-byte coalesced stores over PCIe instead of a
         * device buffer plus a copy-engine transfer, which ran at 5-10 GB/s here after the link had idled
         * through the compute phases.  F4LA_B200_ZEROCOPY=0 restores the copy. */
        uint32_t *dense_dst = ctx->zerocopy ? (uint32_t *)ctx->h_dense.p : as<uint32_t>(ctx->dense_out);
        if ((nf_mode || by_lead) && ncr) {
            /* by_lead: the rows of A^-1 B are the NEGATED solve results */
            dim3 grid((ncr + 31) / 32, (nrl + 31) / 32), block(32, 8);
            k_transpose_rows<<<grid, block, 0, s>>>(Dm, ld, ncr, nrl, by_lead ? p : 0u, dense_dst);
            note_launch(ctx, 1, __LINE__);
        } else if (rank) {
            dim3 grid(rank, (ncr + 255) / 256);
            k_ge_extract<<<grid, 256, 0, s>>>(Dm, ge_ld, ncr, rank, as<uint32_t>(ctx->pivcol), as<uint32_t>(ctx->pivrow),
                                              as<uint32_t>(ctx->pivinv), as<uint8_t>(ctx->ispiv),
                                              dense_dst, p, pinv, nullptr);
            note_launch(ctx, 1, __LINE__);
        }
        if (ctx->debug) CU(cudaEventRecord(ctx->dbg_ev[0], s));
        uint32_t *h_dense = (uint32_t *)ctx->h_dense.p;
        uint32_t *h_pivrow = h_dense + (size_t)ndense * ncr;
        uint32_t *h_pivcol = h_pivrow + rank;
        uint32_t *h_invmap = h_pivcol + rank;
        uint32_t *h_leads = h_invmap + (by_lead ? nc : 0);
        if (dense_bytes && !ctx->zerocopy)
            CU(cudaMemcpyAsync(h_dense, ctx->dense_out.p, dense_bytes, cudaMemcpyDeviceToHost, s));
        if (rank) {
            CU(cudaMemcpyAsync(h_pivrow, ctx->pivrow.p, (size_t)rank * 4, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(h_pivcol, ctx->pivcol.p, (size_t)rank * 4, cudaMemcpyDeviceToHost, s));
        }
        if (by_lead) {
            OK(ensure(ctx, ctx->leads, (size_t)nrl * 4 + 16));
            k_lower_leads<<<(nrl + 255) / 256, 256, 0, s>>>(m->row_ptr, m->cols, nru, nrl, as<uint32_t>(ctx->leads));
            note_launch(ctx, 1, __LINE__);
            CU(cudaMemcpyAsync(h_invmap, d_invmap, (size_t)nc * 4, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(h_leads, ctx->leads.p, (size_t)nrl * 4, cudaMemcpyDeviceToHost, s));
        }
        CU(cudaGetLastError());
        if (ctx->debug) CU(cudaEventRecord(ctx->dbg_ev[1], s));
        const double t_copy0 = now_ms();
        OK(stream_sync(ctx));
        const double t_copy1 = now_ms();
        if (ctx->debug) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ctx->ev[3], ctx->dbg_ev[0]);
            cudaEventElapsedTime(&b, ctx->dbg_ev[0], ctx->dbg_ev[1]);
            fprintf(stderr, "[f4la debug] extract kernel %.3f ms, copies %.3f ms (%zu bytes)\n", a, b, dense_bytes);
        }
        dbg_copy_ms = t_copy1 - t_copy0; dbg_t1 = t_copy1;
        d2h_bytes = dense_bytes + aux_words * 4;

        /* which dense row answers output row i (or -1: the row is zero) */
        std::vector<int64_t> pick(nout, -1);
        if (nf_mode) {
            for (uint32_t i = 0; i < nrl; ++i) pick[i] = i;
        } else if (by_lead) {
            for (uint32_t i = 0; i < nrl; ++i)
                if (h_leads[i] < nc) pick[i] = i;                  /* empty input rows stay empty */
        } else {
            for (uint32_t t = 0; t < rank; ++t) pick[t] = t;
        }
        memset(out, 0, sizeof(*out));
        out->nrows = nout; out->cf_bytes = d.cf_bytes;
        out->row_ptr = (uint64_t *)malloc(((size_t)nout + 1) * sizeof(uint64_t));
        out->src_row = (uint32_t *)malloc((size_t)nout * sizeof(uint32_t));
        std::vector<uint64_t> rnz(nout, 0);
        const int host_thr = ctx->host_threads > 0 ? ctx->host_threads : omp_get_max_threads();
        /* (a thread that never enters an OpenMP construct is never touched by the runtime: with OMP_PROC_BIND set,
         * libgomp pins every foreign thread that does -- even for a serialised region -- to the first place) */
        const bool par = nout > 16 && host_thr > 1;
        auto count_row = [&](uint32_t t) {
            if (pick[t] < 0) return;
            const uint32_t *row = h_dense + (size_t)pick[t] * ncr;
            uint64_t n = by_lead ? 1 : 0;                          /* by_lead: the lead term itself */
            for (uint32_t c = 0; c < ncr; ++c) n += row[c] != 0;
            rnz[t] = n;
        };
        if (par) {
#pragma omp parallel for num_threads(host_thr) schedule(static)
            for (uint32_t t = 0; t < nout; ++t) count_row(t);
        } else {
            for (uint32_t t = 0; t < nout; ++t) count_row(t);
        }
        uint64_t nnz = 0;
        for (uint32_t t = 0; t < nout; ++t) {
            out->row_ptr[t] = nnz;
            /* a pivot row of a block of random combinations: one of the lower rows that entered that combination */
            out->src_row[t] = (nf_mode || by_lead) ? t
                              : rows_are_combinations ? ctx->donor[h_pivrow[rank - 1 - t]] : h_pivrow[rank - 1 - t];
            nnz += rnz[t];
        }
        out->row_ptr[nout] = nnz;
        out->cols = (uint32_t *)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
        out->cf = malloc((nnz ? nnz : 1) * (size_t)d.cf_bytes);
        auto fill_row = [&](uint32_t t) {
            if (pick[t] < 0) return;
            uint64_t k = out->row_ptr[t];
            const uint32_t *row = h_dense + (size_t)pick[t] * ncr;
            if (by_lead) {
                out->cols[k] = h_leads[t];
                if (d.cf_bytes == 4) ((uint32_t *)out->cf)[k] = 1;
                else if (d.cf_bytes == 2) ((uint16_t *)out->cf)[k] = 1;
                else ((uint8_t *)out->cf)[k] = 1;
                ++k;
            }
            for (uint32_t c = 0; c < ncr; ++c) {
                const uint32_t v = row[c];
                if (!v) continue;
                out->cols[k] = by_lead ? h_invmap[ncl + c] : ncl + c;
                if (d.cf_bytes == 4) ((uint32_t *)out->cf)[k] = v;
                else if (d.cf_bytes == 2) ((uint16_t *)out->cf)[k] = (uint16_t)v;
                else ((uint8_t *)out->cf)[k] = (uint8_t)v;
                ++k;
            }
        };
        if (par) {
#pragma omp parallel for num_threads(host_thr) schedule(static)
            for (uint32_t t = 0; t < nout; ++t) fill_row(t);
        } else {
            for (uint32_t t = 0; t < nout; ++t) fill_row(t);
        }
    }
    if (ctx->debug) fprintf(stderr, "[f4la debug] result: %u rows, wait for extract + D2H %.3f ms, host compaction %.3f ms\n", nout, dbg_copy_ms, now_ms() - dbg_t1);
    if (want_rba && rba_words) {
        out->rba_words = rba_words;
        out->rba = (uint32_t *)malloc((size_t)nrl * rba_words * 4);
        CU(cudaMemcpyAsync(out->rba, ctx->rba.p, (size_t)nrl * rba_words * 4, cudaMemcpyDeviceToHost, s));
        OK(stream_sync(ctx));
        d2h_bytes += (uint64_t)nrl * rba_words * 4;
    }
    if (timed) {
        CU(cudaEventRecord(ctx->ev[4], s));
        CU(cudaEventSynchronize(ctx->ev[4]));
    }

    if (st) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); st->ms_prep = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); st->ms_reduce = ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); st->ms_echelon = ms;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]); st->ms_d2h = ms;
        for (int k = 0; k < n_flow_ev; k += 3) {
            cudaEventElapsedTime(&ms, ctx->flow_ev[k], ctx->flow_ev[k + 2]);
            st->ms_hot_kernel += ms;
            cudaEventElapsedTime(&ms, ctx->flow_ev[k + 1], ctx->flow_ev[k + 2]);
            if (schur) st->ms_schur += ms;
        }
        st->hot_kernel_launches = (uint32_t)(n_flow_ev / 3);
        st->schur_macs = schur ? (uint64_t)nnz_schur * nrl : 0;
        st->units = ref_units;
        st->ref_applications = ref_apps;
        st->dense_macs = (uint64_t)nnz_pull * nrl;
        st->pivot_applications = (uint64_t)nnz_pull;
        st->bytes_d2h = d2h_bytes;
        st->kernel_launches = ctx->launches - launches0;
        st->rank = rank;
        st->echelon_path = echelon_path;
    }
    return F4LA_OK;
}

/* Copies the flat arrays to the device.  own == true: fresh allocations that
 * belong to the resident handle; own == false: the context's grow-only staging
 * buffers (no cudaMalloc/cudaFree on the per-call path). */
int upload(f4la_ctx *ctx, const f4la_flat_matrix *in, f4la_resident *r, bool own)
{
    const uint64_t n = (uint64_t)in->nru + in->nrl;
    if (in->fc < 2 || in->fc >= 0x80000000u)
        return fail(ctx, F4LA_ERR_ARG, "field characteristic %u out of range (need 2 <= p < 2^31)", in->fc);
    if (in->cf_bytes != 1 && in->cf_bytes != 2 && in->cf_bytes != 4)
        return fail(ctx, F4LA_ERR_ARG, "cf_bytes must be 1, 2 or 4");
    if ((in->flags & (F4LA_FLAG_COLS_STAGED | F4LA_FLAG_CF_STAGED)) && own)
        return fail(ctx, F4LA_ERR_ARG, "F4LA_FLAG_COLS_STAGED / F4LA_FLAG_CF_STAGED are for f4la_b200_reduce() only");
    r->dims = *in;
    r->dims.flags &= ~(F4LA_FLAG_COLS_STAGED | F4LA_FLAG_CF_STAGED);
    r->owned = own;
    r->nnz = n ? in->row_ptr[n] : 0;
    r->ncoef = in->ncf ? in->cf_ptr[in->ncf] : 0;
    if (own) {
        CU(cudaMalloc(&r->row_ptr, (n + 1) * 8));
        CU(cudaMalloc(&r->cols, (r->nnz + 1) * 4));
        CU(cudaMalloc(&r->row_cf, (n + 1) * 4));
        CU(cudaMalloc(&r->cf_ptr, ((uint64_t)in->ncf + 1) * 8));
        CU(cudaMalloc(&r->cf, (r->ncoef + 1) * in->cf_bytes));
    } else {
        OK(ensure(ctx, ctx->in_row_ptr, (n + 1) * 8));
        OK(ensure(ctx, ctx->in_cols, (r->nnz + 1) * 4));
        OK(ensure(ctx, ctx->in_row_cf, (n + 1) * 4));
        OK(ensure(ctx, ctx->in_cf_ptr, ((uint64_t)in->ncf + 1) * 8));
        OK(ensure(ctx, ctx->in_cf, (r->ncoef + 1) * in->cf_bytes));
        r->row_ptr = as<uint64_t>(ctx->in_row_ptr); r->cols = as<uint32_t>(ctx->in_cols);
        r->row_cf = as<uint32_t>(ctx->in_row_cf); r->cf_ptr = as<uint64_t>(ctx->in_cf_ptr); r->cf = ctx->in_cf.p;
    }
    cudaStream_t s = ctx->stream;
    if (ctx->copy_pending) { CU(cudaStreamSynchronize(ctx->copy_stream)); ctx->copy_pending = false; }
    CU(cudaMemcpyAsync(r->row_ptr, in->row_ptr, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    if (n) CU(cudaMemcpyAsync(r->row_cf, in->row_cf, n * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(r->cf_ptr, in->cf_ptr, ((uint64_t)in->ncf + 1) * 8, cudaMemcpyHostToDevice, s));
    /* The pull lists and levels (steps 1-2 of the pipeline) read the pivot rows only.  From host buffers the
     * lower rows' column indices and coefficient arrays therefore travel on a second stream while those
     * steps run; the pipeline waits for them just before it scatters the lower rows. */
    const bool split = !own && ctx->split_h2d && in->nrl && in->nru && r->nnz >= (1u << 20) &&
                       !(in->flags & (F4LA_FLAG_PIVOT_BY_LEAD | F4LA_FLAG_COLS_STAGED | F4LA_FLAG_CF_STAGED));
    if (split) {
        const uint64_t nz_up = in->row_ptr[in->nru];
        uint32_t k_up = 0;                                   /* coefficient arrays 0 .. k_up-1 serve the pivot rows */
        for (uint32_t i = 0; i < in->nru; ++i) k_up = in->row_cf[i] >= k_up ? in->row_cf[i] + 1 : k_up;
        if (k_up > in->ncf) return fail(ctx, F4LA_ERR_ARG, "row_cf out of range");
        const uint64_t cf_up = in->cf_ptr[k_up];
        const size_t w = in->cf_bytes;
        if (nz_up) CU(cudaMemcpyAsync(r->cols, in->cols, nz_up * 4, cudaMemcpyHostToDevice, s));
        if (cf_up) CU(cudaMemcpyAsync(r->cf, in->cf, cf_up * w, cudaMemcpyHostToDevice, s));
        cudaStream_t cs = ctx->copy_stream;
        if (r->nnz > nz_up)
            CU(cudaMemcpyAsync(r->cols + nz_up, in->cols + nz_up, (r->nnz - nz_up) * 4, cudaMemcpyHostToDevice, cs));
        if (r->ncoef > cf_up)
            CU(cudaMemcpyAsync((char *)r->cf + cf_up * w, (const char *)in->cf + cf_up * w, (r->ncoef - cf_up) * w,
                               cudaMemcpyHostToDevice, cs));
        CU(cudaEventRecord(ctx->ev_copy, cs));
        ctx->copy_pending = true;
    } else {
        if (r->nnz && !(in->flags & F4LA_FLAG_COLS_STAGED))
            CU(cudaMemcpyAsync(r->cols, in->cols, r->nnz * 4, cudaMemcpyHostToDevice, s));
        if (r->ncoef && !(in->flags & F4LA_FLAG_CF_STAGED))
            CU(cudaMemcpyAsync(r->cf, in->cf, r->ncoef * in->cf_bytes, cudaMemcpyHostToDevice, s));
    }
    r->bytes = (n + 1) * 8 + r->nnz * 4 + n * 4 + ((uint64_t)in->ncf + 1) * 8 + r->ncoef * in->cf_bytes;
    return F4LA_OK;
}

void release_resident(f4la_resident *r)
{
    if (!r) return;
    if (r->owned) { cudaFree(r->row_ptr); cudaFree(r->cols); cudaFree(r->row_cf); cudaFree(r->cf_ptr); cudaFree(r->cf); }
    PrepCache &pc = r->cache;
    cudaFree(pc.col_ptr); cudaFree(pc.ent_src); cudaFree(pc.sched); cudaFree(pc.items); cudaFree(pc.colmap); cudaFree(pc.invmap);
    cudaFree(pc.ent);
    delete r;
}

} // namespace

/* ------------------------------------------------------------------ */
/* C ABI                                                               */
/* ------------------------------------------------------------------ */

extern "C" {

const char *f4la_b200_version(void) { return "f4la_b200 0.1.0 (sm_100a)"; }

int f4la_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

f4la_ctx *f4la_b200_create(int device)
{
    int n = f4la_b200_device_count();
    if (n <= 0) {
        fprintf(stderr, "[f4la_b200] no CUDA device available; this library has no CPU fallback\n");
        return nullptr;
    }
    if (device < 0 || device >= n) {
        fprintf(stderr, "[f4la_b200] device %d out of range (0..%d)\n", device, n - 1);
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        fprintf(stderr, "[f4la_b200] device %d is not an sm_100 class GPU (found sm_%d%d); kernels are built "
                        "for sm_100a only\n", device, prop.major, prop.minor);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    f4la_ctx *ctx = new f4la_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->coop_ok = prop.cooperativeLaunch;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return nullptr; }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return nullptr; }
    cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming);
    for (auto &e : ctx->ev_side) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    { const char *sp = getenv("F4LA_B200_SPLIT_H2D"); if (sp) ctx->split_h2d = atoi(sp) != 0; }
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    for (auto &e : ctx->dbg_ev) cudaEventCreate(&e);
    for (auto &e : ctx->flow_ev) cudaEventCreate(&e);
    const char *ge = getenv("F4LA_B200_GE");
    ctx->ge_mode = !ge ? 0 : !strcmp(ge, "columnwise") ? 2 : !strcmp(ge, "uncached") ? 1 : !strcmp(ge, "batched") ? 3 : 0;
    const char *cm = getenv("F4LA_B200_COMPRESS");
    if (cm) ctx->ge_compress = atoi(cm) != 0;
    const char *cr = getenv("F4LA_B200_COMPRESS_ROWS");
    if (cr && atoi(cr) >= 16) ctx->compress_rows = (uint32_t)atoi(cr) / 8 * 8;
    const char *cf = getenv("F4LA_B200_COMPRESS_FAULT");
    ctx->compress_fault = cf && atoi(cf) != 0;
    const char *gb = getenv("F4LA_B200_GE_BATCH");
    if (gb && atoi(gb) >= 2 && atoi(gb) <= (int)kGeBatchCached) ctx->ge_batch = (uint32_t)atoi(gb);
    const char *gm = getenv("F4LA_B200_GE_MISSES");
    if (gm && atoi(gm) >= 1) ctx->ge_miss_limit = (uint32_t)atoi(gm);
    const char *sct = getenv("F4LA_B200_SCHUR_TC");
    if (sct) ctx->schur_tc = atoi(sct) != 0;
    const char *scm = getenv("F4LA_B200_SCHUR_MIN");
    if (scm && atof(scm) >= 0) ctx->schur_min_macs = atof(scm);
    const char *tcv = getenv("F4LA_B200_TC");
    if (tcv) ctx->tc_verify = atoi(tcv) != 0;
    const char *zc = getenv("F4LA_B200_ZEROCOPY");
    if (zc) ctx->zerocopy = atoi(zc) != 0;
    const char *dbg = getenv("F4LA_B200_DEBUG");
    ctx->debug = dbg && atoi(dbg) != 0;
    { const char *gp = getenv("F4LA_B200_GE_PANEL"); if (gp && atoi(gp) >= 2) ctx->ge_panel = (uint32_t)atoi(gp); }
    { const char *gc = getenv("F4LA_B200_GRID_CAP"); if (gc) ctx->grid_cap = (uint32_t)atoi(gc); }
    const char *dsy = getenv("F4LA_B200_SYNC");
    ctx->debug_sync = dsy && atoi(dsy) != 0;
    const char *wd = getenv("F4LA_B200_WATCHDOG_MS");
    ctx->watchdog_ms = wd && atof(wd) > 0 ? atof(wd) : 20000.0;
    const char *tm = getenv("F4LA_B200_TMAX");
    if (tm && atoi(tm) >= 1 && atoi(tm) <= kMaxT) ctx->tmax = atoi(tm);
    const char *bud = getenv("F4LA_B200_L2_MB");
    ctx->l2_budget = bud && atoi(bud) > 0 ? (size_t)atoi(bud) << 20 : kL2BudgetDefault;
    return ctx;
}

void f4la_b200_destroy(f4la_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->col_cnt, &ctx->col_ptr, &ctx->cursor, &ctx->ent, &ctx->level, &ctx->slots, &ctx->slot_ptr,
                      &ctx->order, &ctx->flags32, &ctx->V, &ctx->D, &ctx->colflag, &ctx->gebytes,
                      &ctx->pivcol, &ctx->pivrow, &ctx->pivinv, &ctx->dense_out,
                      &ctx->sched, &ctx->items, &ctx->flowflags, &ctx->tcA, &ctx->tcB, &ctx->scB, &ctx->scFlagA, &ctx->scFlagB, &ctx->scList, &ctx->scCnt, &ctx->scP, &ctx->in_row_ptr, &ctx->in_cols,
                      &ctx->in_row_cf, &ctx->in_cf_ptr, &ctx->in_cf, &ctx->is_piv, &ctx->pividx, &ctx->colmap,
                      &ctx->invmap, &ctx->leads, &ctx->rba, &ctx->D2, &ctx->combo};
    for (auto b : bufs) free_dev(*b);
    if (ctx->h_small.p) cudaFreeHost(ctx->h_small.p);
    if (ctx->h_dense.p) cudaFreeHost(ctx->h_dense.p);
    if (ctx->h_combo.p) cudaFreeHost(ctx->h_combo.p);
    for (int k = 0; k < kRingSlots; ++k) {
        if (ctx->ring_buf[k]) cudaFreeHost(ctx->ring_buf[k]);
        if (ctx->ring_ev[k]) cudaEventDestroy(ctx->ring_ev[k]);
    }
    for (auto &e : ctx->ev) cudaEventDestroy(e);
    for (auto &e : ctx->dbg_ev) cudaEventDestroy(e);
    for (auto &e : ctx->flow_ev) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->ev_copy);
    for (auto &e : ctx->ev_side) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *f4la_b200_last_error(const f4la_ctx *ctx) { return ctx ? ctx->err : "no context"; }

void f4la_b200_set_host_mode(f4la_ctx *ctx, int host_threads, int blocking_sync)
{
    if (!ctx) return;
    ctx->host_threads = host_threads > 0 ? host_threads : 0;
    ctx->blocking_sync = blocking_sync != 0;
}

void f4la_b200_set_grid_cap(f4la_ctx *ctx, uint32_t ctas) { if (ctx) ctx->grid_cap = ctas; }

void f4la_b200_set_seed(f4la_ctx *ctx, uint64_t seed) { if (ctx) ctx->seed = seed ? seed : 0x9e3779b97f4a7c15ull; }

void *f4la_b200_pinned_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void f4la_b200_pinned_free(void *p) { if (p) cudaFreeHost(p); }

f4la_resident *f4la_b200_upload(f4la_ctx *ctx, const f4la_flat_matrix *in)
{
    if (!ctx || !in) return nullptr;
    cudaSetDevice(ctx->device);
    f4la_resident *r = new f4la_resident();
    if (upload(ctx, in, r, true) != F4LA_OK) { release_resident(r); return nullptr; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { release_resident(r); return nullptr; }
    return r;
}

int f4la_b200_resident_update(f4la_ctx *ctx, f4la_resident *m, uint32_t fc, const void *cf)
{
    if (!ctx || !m || !cf) return F4LA_ERR_ARG;
    if (fc < 2 || fc >= 0x80000000u) return fail(ctx, F4LA_ERR_ARG, "field characteristic %u out of range", fc);
    cudaSetDevice(ctx->device);
    m->dims.fc = fc;
    m->cache_enabled = true;          /* the structure will be reused: keep what the layout kernels derive from it */
    if (m->ncoef)
        CU(cudaMemcpyAsync(m->cf, cf, m->ncoef * m->dims.cf_bytes, cudaMemcpyHostToDevice, ctx->stream));
    return F4LA_OK;
}

void f4la_b200_release(f4la_ctx *ctx, f4la_resident *m)
{
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    release_resident(m);
}

int f4la_b200_reduce_resident(f4la_ctx *ctx, f4la_resident *m, f4la_flat_result *out, f4la_stats *stats)
{
    if (!ctx || !m || !out) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (stats) memset(stats, 0, sizeof(*stats));
    const double t0 = now_ms();
    memset(out, 0, sizeof(*out));
    int rc = run_pipeline(ctx, m, out, stats);
    if (rc != F4LA_OK) f4la_b200_free_result(out);      /* nothing half-built is handed back */
    if (stats) stats->ms_total = now_ms() - t0;
    return rc;
}

/* ---- device-resident replay (the tape of f4la_neogb.cpp drives these) ---- */
void *f4la_b200_device_alloc(f4la_ctx *ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void f4la_b200_device_free(f4la_ctx *ctx, void *p)
{
    if (ctx) cudaSetDevice(ctx->device);
    if (p) cudaFree(p);
}

int f4la_b200_device_copy(f4la_ctx *ctx, void *dst, const void *src, size_t bytes, int to_device, int wait)
{
    if (!ctx || (!dst && bytes) || (!src && bytes)) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    /* to_device == 0: the destination may be host OR device memory (unified addressing tells which) */
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDefault, ctx->stream));
    if (wait) OK(stream_sync(ctx));
    return F4LA_OK;
}

int f4la_b200_device_zero(f4la_ctx *ctx, void *dst, size_t bytes)
{
    if (!ctx || !dst) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaMemsetAsync(dst, 0, bytes, ctx->stream));
    return F4LA_OK;
}

int f4la_b200_resident_layout(f4la_ctx *ctx, const f4la_resident *m, uint32_t *colmap, uint32_t *ncl_effective)
{
    if (!ctx || !m) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    const f4la_flat_matrix &d = m->dims;
    const bool by_lead = (d.flags & F4LA_FLAG_PIVOT_BY_LEAD) != 0;
    if (ncl_effective) *ncl_effective = by_lead ? d.nru + d.nrl : d.ncl;
    if (colmap) {
        const uint32_t nc = d.ncl + d.ncr;
        if (!by_lead) { for (uint32_t c = 0; c < nc; ++c) colmap[c] = c; return F4LA_OK; }
        if (!m->cache.valid || !m->cache.colmap) return fail(ctx, F4LA_ERR_ARG, "the layout cache of this matrix is empty");
        CU(cudaMemcpyAsync(colmap, m->cache.colmap, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
        OK(stream_sync(ctx));
    }
    return F4LA_OK;
}

int f4la_b200_replay_call(f4la_ctx *ctx, f4la_resident *m, uint32_t fc, const f4la_replay_call *rc)
{
    if (!ctx || !m || !rc || !rc->arena || !rc->status) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    m->dims.fc = fc;
    if (rc->nslots) {
        unsigned blocks = (rc->nslots + 7) / 8;
        if (blocks > 1024) blocks = 1024;
        k_replay_gather<<<blocks, 256, 0, ctx->stream>>>(rc->arena, reinterpret_cast<uint32_t *>(m->cf), rc->gather_src,
                                                        rc->gather_dst, rc->gather_len, rc->nslots);
        note_launch(ctx, 1, __LINE__);
    }
    return run_pipeline(ctx, m, nullptr, nullptr, rc);
}

int f4la_b200_stage_cols(f4la_ctx *ctx, const uint32_t *cols, uint64_t first, uint64_t count, uint64_t total_nnz)
{
    if (!ctx || !cols || first + count > total_nnz) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    /* same size as upload() asks for: the buffer is not reallocated between the pieces and the reduce call */
    OK(ensure(ctx, ctx->in_cols, (total_nnz + 1) * 4));
    if (count)
        CU(cudaMemcpyAsync(as<uint32_t>(ctx->in_cols) + first, cols + first, count * 4, cudaMemcpyHostToDevice, ctx->stream));
    return F4LA_OK;
}

int f4la_b200_stage_acquire(f4la_ctx *ctx, void **piece, size_t *capacity_bytes)
{
    if (!ctx || !piece || !capacity_bytes) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    const int k = ctx->ring_next;
    if (!ctx->ring_buf[k]) {
        CU(cudaMallocHost(&ctx->ring_buf[k], kRingPieceBytes));
        CU(cudaEventCreateWithFlags(&ctx->ring_ev[k], cudaEventDisableTiming));
    }
    if (ctx->ring_busy[k]) { CU(cudaEventSynchronize(ctx->ring_ev[k])); ctx->ring_busy[k] = false; }
    ctx->ring_held = k;
    *piece = ctx->ring_buf[k];
    *capacity_bytes = kRingPieceBytes;
    return F4LA_OK;
}

int f4la_b200_stage_commit(f4la_ctx *ctx, int which, uint32_t elem_bytes, uint64_t first, uint64_t count, uint64_t total)
{
    if (!ctx || ctx->ring_held < 0 || first + count > total || count * elem_bytes > kRingPieceBytes) return F4LA_ERR_ARG;
    if ((which == F4LA_STAGE_COLS && elem_bytes != 4) ||
        (which == F4LA_STAGE_CF && elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4) ||
        (which != F4LA_STAGE_COLS && which != F4LA_STAGE_CF)) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    const int k = ctx->ring_held;
    /* same sizes as upload() asks for: the buffers are not reallocated between the pieces and the reduce call */
    DevBuf &dst = which == F4LA_STAGE_COLS ? ctx->in_cols : ctx->in_cf;
    OK(ensure(ctx, dst, (total + 1) * elem_bytes));
    if (count) {
        CU(cudaMemcpyAsync((char *)dst.p + first * elem_bytes, ctx->ring_buf[k], count * elem_bytes, cudaMemcpyHostToDevice,
                           ctx->stream));
        CU(cudaEventRecord(ctx->ring_ev[k], ctx->stream));
        ctx->ring_busy[k] = true;
    }
    ctx->ring_held = -1;
    ctx->ring_next = (k + 1) % kRingSlots;
    return F4LA_OK;
}

int f4la_b200_reduce(f4la_ctx *ctx, const f4la_flat_matrix *in, f4la_flat_result *out, f4la_stats *stats)
{
    if (!ctx || !in || !out) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (stats) memset(stats, 0, sizeof(*stats));
    const double t0 = now_ms();
    cudaEventRecord(ctx->ev[6], ctx->stream);
    f4la_resident r;
    int rc = upload(ctx, in, &r, false);
    /* the pipeline runs on the same stream: no synchronisation needed before it;
     * the H2D time is measured with events */
    cudaEventRecord(ctx->ev[5], ctx->stream);
    memset(out, 0, sizeof(*out));
    if (rc == F4LA_OK) rc = run_pipeline(ctx, &r, out, stats);
    if (rc != F4LA_OK) f4la_b200_free_result(out);      /* nothing half-built is handed back */
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_pending) {                  /* the pipeline returned before it needed the lower rows */
        cudaStreamSynchronize(ctx->copy_stream);
        ctx->copy_pending = false;
    }
    if (stats) {
        float ms = 0;
        if (rc == F4LA_OK && cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[5]) == cudaSuccess) stats->ms_h2d = ms;
        stats->bytes_h2d = r.bytes;
        stats->ms_total = now_ms() - t0;
    }
    return rc;
}

int f4la_b200_modp_gemm(f4la_ctx *ctx, const uint32_t *A, const uint32_t *B, uint32_t M, uint32_t N, uint32_t K, uint32_t p,
                        uint32_t *C, double *kernel_ms)
{
    if (!ctx || !A || !B || !C || p < 2 || p >= 0x80000000u) return F4LA_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const uint64_t Mp = ((uint64_t)M + kTcBM - 1) / kTcBM * kTcBM;
    uint32_t *dA = nullptr, *dB = nullptr, *dC = nullptr;
    int rc = F4LA_OK;
    auto body = [&]() -> int {
        CU(cudaMalloc(&dA, Mp * (K ? K : 1) * 4));
        CU(cudaMalloc(&dB, (size_t)(K ? K : 1) * N * 4));
        CU(cudaMalloc(&dC, Mp * N * 4));
        CU(cudaMemsetAsync(dA, 0, Mp * (K ? K : 1) * 4, s));
        /* A arrives column major with leading dimension M: pad every column to Mp rows */
        CU(cudaMemcpy2DAsync(dA, Mp * 4, A, (size_t)M * 4, (size_t)M * 4, K, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dB, B, (size_t)K * N * 4, cudaMemcpyHostToDevice, s));
        OK(ensure(ctx, ctx->flags32, 64));
        CU(cudaMemsetAsync(ctx->flags32.p, 0, 64, s));
        TcGemmArgs g;
        memset(&g, 0, sizeof(g));
        g.A = dA; g.lda = Mp; g.B = dB; g.ldb = N; g.M = (uint32_t)Mp; g.N = N; g.K = K; g.p = p;
        g.C = dC; g.ldc = Mp; g.err = as<uint32_t>(ctx->flags32); g.mismatch = as<uint32_t>(ctx->flags32) + 8;
        CU(cudaEventRecord(ctx->ev[0], s));
        OK(launch_tc_gemm(ctx, g));
        CU(cudaEventRecord(ctx->ev[1], s));
        uint32_t h_err = 0;
        CU(cudaMemcpyAsync(&h_err, ctx->flags32.p, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpy2DAsync(C, (size_t)M * 4, dC, Mp * 4, (size_t)M * 4, N, cudaMemcpyDeviceToHost, s));
        OK(stream_sync(ctx));
        if (h_err) return fail(ctx, F4LA_ERR_CUDA, "tensor-core product reported error bits 0x%x (64: pipeline wait expired)", h_err);
        if (kernel_ms) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); *kernel_ms = ms; }
        return F4LA_OK;
    };
    rc = body();
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return rc;
}

double f4la_b200_measure_mac_peak(f4la_ctx *ctx, double milliseconds)
{
    if (!ctx) return 0.0;
    cudaSetDevice(ctx->device);
    const int blocks = ctx->sm_count * 8;
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * 256 * 4) != cudaSuccess) return 0.0;
    cudaStream_t s = ctx->stream;
    double best = 0.0;
    int iters = 4000;
    const double t_end = now_ms() + (milliseconds > 0 ? milliseconds : 200.0);
    for (int rep = 0; rep < 64; ++rep) {
        cudaEventRecord(ctx->ev[0], s);
        k_mac_peak<<<blocks, 256, 0, s>>>(d, 12345u + rep, iters);
        cudaEventRecord(ctx->ev[1], s);
        if (cudaEventSynchronize(ctx->ev[1]) != cudaSuccess) { best = 0.0; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        const double rate = (double)blocks * 256 * 8 * iters / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;          /* the first launch pays the module load */
        if (rep > 0 && now_ms() > t_end) break;
        if (ms < 2.0f) iters *= 2;
    }
    cudaFree(d);
    return best;
}

void f4la_b200_free_result(f4la_flat_result *res)
{
    if (!res) return;
    free(res->row_ptr); free(res->cols); free(res->cf); free(res->src_row); free(res->rba);
    memset(res, 0, sizeof(*res));
}

} /* extern "C" */
