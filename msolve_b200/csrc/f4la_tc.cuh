/* f4la_tc.cuh -- exact dense products mod p on the 5th-generation tensor cores (sm_100a: tcgen05 + TMEM).
 *
 *     C[M x N] = A[M x K] * B[K x N]  (mod p),   residues < p < 2^31
 *
 * This is the one truly dense contraction of the path (north_star (3)): the product that certifies the
 * echelon form of a tall trailing block, D' == D'[:, pivot columns] * B (compress-and-verify, DESIGN.md
 * section 4; the reference eliminates the same block row by row on the CPU, la_ff_32.c:3390-3500 with the
 * dense AXPY kernel :2182-2303).
 *
 * Arithmetic.  Every residue is split into four unsigned 8-bit limbs, x = sum_a x_a 2^(8a).  Then
 *     sum_k A[i][k] B[k][j] = sum_{s=0..6} 2^(8s) * S_s[i][j],   S_s = sum_{a+b=s} sum_k A_a[i][k] B_b[k][j]
 * and every S_s is an unsigned-8-bit x unsigned-8-bit matrix product with 32-bit accumulation:
 * tcgen05.mma.kind::i8 (u8 x u8 -> s32).  NO-OVERFLOW BOUND: a product is <= 255^2, the top limbs are <= 127
 * (x < 2^31), at most four (a,b) pairs feed one S_s, so per k the worst S_s (s = 2: three full pairs) grows by
 * <= 195 075 and K <= 8192 keeps every accumulator below 1.6e9 < 2^31.  The host only uses this path for
 * K <= kTcMaxK.  The seven S_s are recombined exactly in 64-bit integer arithmetic and reduced mod p once.
 *
 * Tensor-core mapping.  One CTA computes a 128 x 64 tile of C.  B is staged LIMB-MAJOR: the shared-memory
 * operand has 4 x 64 = 256 "rows" (limb b of column j at row 64 b + j), so ONE instruction with N = 256
 * multiplies limb a of A with all four limbs of B; its 256 accumulator columns are placed at TMEM column
 * 64 a, which makes the column block [64 s, 64 s + 64) collect exactly the pairs with a + b = s:
 *
 *     for a in 0..3:  tcgen05.mma  D[tmem + 64 a : + 256] += A_a[128 x 32] * B'[32 x 256]      (M=128, N=256, K=32)
 *
 * 7 x 64 = 448 TMEM columns (512 allocated), 4 MMAs per 32 deep k-step, every MMA at the full N = 256 shape.
 *
 * Operands.  Splitting u32 residues into limb planes is a byte shuffle the copy engine cannot do, and in a tiled
 * product every A tile is needed N / 64 times and every B tile M / 128 times.  The first version split inside the
 * product kernel (four producer warps: LDG, PRMT, STS): 3 % tensor-pipe activity, the warps sat on the L2 latency
 * of their loads (ncu: long scoreboard 4.9 of 9.6 stall cycles per issue, profiles/r02_ncu_tc_gemm_v1_*).  Now two
 * small pre-passes (k_tc_split_a / _b) write both operands ONCE in exactly the shared-memory image of a pipeline
 * stage -- K-major, no swizzle, 8 x 16-byte core matrices, [k-block][limb][16-byte k-group][row] -- and the product
 * kernel's producer is one thread issuing bulk copies (cp.async.bulk.shared.global, UBLKCP in SASS, completion on
 * the stage's mbarrier) of whole 32 KB + 16 KB stages; a 3-stage ring decouples it from the single MMA-issuing
 * thread; tcgen05.commit hands the stages back and signals the epilogue warps, which read the accumulators with
 * tcgen05.ld, recombine the limbs and compare with / store the result.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "f4la_kernels.cuh"

namespace f4la {

constexpr int kTcBM = 128, kTcBN = 64, kTcBK = 64, kTcStages = 3;
constexpr int kTcEpilogue = 128;                  /* warps 0-3: epilogue (TMEM lanes 32 w .. 32 w + 31) */
constexpr int kTcThreads = kTcEpilogue + 64;      /* warp 4: TMEM allocation + MMA issue, warp 5: bulk-copy producer */
constexpr uint32_t kTcMaxK = 8192;                /* accumulation depth bound of the s32 accumulators (see above) */
constexpr uint32_t kTcPlaneA = kTcBM * kTcBK;     /* bytes of one limb plane of the A tile */
constexpr uint32_t kTcStageA = 4 * kTcPlaneA, kTcStageB = 4 * kTcBN * kTcBK;
constexpr uint32_t kTcStageBytes = kTcStageA + kTcStageB;
constexpr size_t   kTcSmemBytes = (size_t)kTcStages * kTcStageBytes + 256;
constexpr uint32_t kTcTmemCols = 512;
constexpr uint32_t kTcSpinLimit = 1u << 22;       /* mbarrier polls before a wait gives up (never hang the GPU) */
constexpr uint32_t kErrTcTimeout = 64u;

/* instruction descriptor of tcgen05.mma.kind::i8: D s32 (bits 4-5 = 2), A and B unsigned 8 bit (bits 7-9, 10-12 = 0),
 * both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28 */
constexpr uint32_t kTcIdesc = (2u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

struct TcGemmArgs {
    const uint32_t *A;        /* column major: A(i, t) = A[acol(t) * lda + i]                      */
    uint64_t        lda;
    const uint32_t *acol;     /* optional column index list (NULL: identity)                      */
    uint32_t        acol_reverse;   /* acol(t) = acol[K - 1 - t] (pivots are stored by increasing column) */
    const uint32_t *B;        /* row major: B(t, j) = B[t * ldb + j]                              */
    uint64_t        ldb;
    uint8_t        *Asplit;   /* [M / 128][K blocks] stage images of A: 4 limb planes x 4 k-groups x 128 rows x 16 B */
    uint8_t        *Bsplit;   /* [N blocks][K blocks] stage images of B: 4 k-groups x (4 limbs x 64 columns) x 16 B  */
    uint32_t        M, N, K;  /* M must be a multiple of 128 (rows beyond the data are zero)      */
    uint32_t        p, w32;   /* w32 = 2^32 mod p                                                 */
    uint64_t        pinv;
    const uint32_t *D;        /* verify mode: compared with D[j * ldd + i]                        */
    uint64_t        ldd;
    uint32_t       *C;        /* store mode (C != NULL): C[j * ldc + i] = product                 */
    uint64_t        ldc;
    uint32_t       *mismatch; /* verify mode: |= 1 when an entry differs                          */
    uint32_t       *err;      /* |= kErrTcTimeout when a barrier wait expired                     */
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

/* false when the wait expired (an error bit is set, the caller stops using the pipeline) */
__device__ __forceinline__ bool tc_mbar_wait(uint32_t bar, uint32_t parity, uint32_t *err)
{
    for (uint32_t tries = 0; tries < kTcSpinLimit; ++tries) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    atomicOr(err, kErrTcTimeout);
    return false;
}

/* shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; LBO = byte distance of
 * the core matrices adjacent in K, SBO = byte distance of the 8-row groups; descriptor version 1 (sm_100) */
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(1u) : "memory");
}

__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tc_st8_zero(uint32_t taddr)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(taddr), "r"(0u) : "memory");
}

/* 16 residues -> 4 limb planes x 16 bytes (plane a = byte a of every residue, in order) */
__device__ __forceinline__ void tc_split16(const uint32_t (&x)[16], uint4 (&out)[4])
{
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const uint32_t sel = (uint32_t)a | ((4u + a) << 4);
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t t01 = __byte_perm(x[4 * q], x[4 * q + 1], sel);
            const uint32_t t23 = __byte_perm(x[4 * q + 2], x[4 * q + 3], sel);
            w[q] = __byte_perm(t01, t23, 0x5410);
        }
        out[a] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

constexpr uint32_t kTcBulkBytes = 8192;           /* bytes per bulk copy (a stage = 4 + 2 of them) */

__device__ __forceinline__ void tc_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

/* Pre-pass: the stage images of A.  grid (M / 128, K blocks), 128 threads: thread i owns row i of the block. */
__global__ void __launch_bounds__(128)
k_tc_split_a(const TcGemmArgs G)
{
    const uint32_t tid = threadIdx.x, i0 = blockIdx.x * kTcBM, t0 = blockIdx.y * kTcBK;
    uint8_t *dst = G.Asplit + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * kTcStageA;
#pragma unroll 1
    for (uint32_t q = 0; q < 4; ++q) {
        uint32_t x[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const uint32_t t = t0 + 16 * q + kk;
            uint32_t v = 0;
            if (t < G.K) {
                const uint32_t col = G.acol ? G.acol[G.acol_reverse ? G.K - 1 - t : t] : t;
                v = __ldg(G.A + (size_t)col * G.lda + i0 + tid);
            }
            x[kk] = v;
        }
        uint4 pl[4];
        tc_split16(x, pl);
#pragma unroll
        for (int a = 0; a < 4; ++a)
            *reinterpret_cast<uint4 *>(dst + a * kTcPlaneA + q * (kTcBM * 16) + tid * 16) = pl[a];
    }
}

/* Pre-pass: the stage images of B, limb major.  grid (N blocks, K blocks), 128 threads: column tid & 63, half tid >> 6. */
__global__ void __launch_bounds__(128)
k_tc_split_b(const TcGemmArgs G)
{
    const uint32_t tid = threadIdx.x, bc = tid & 63, bhalf = tid >> 6;
    const uint32_t j0 = blockIdx.x * kTcBN, t0 = blockIdx.y * kTcBK;
    const bool b_in = j0 + bc < G.N;
    uint8_t *dst = G.Bsplit + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * kTcStageB;
#pragma unroll 1
    for (uint32_t q = 2 * bhalf; q < 2 * bhalf + 2; ++q) {
        uint32_t x[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const uint32_t t = t0 + 16 * q + kk;
            x[kk] = (b_in && t < G.K) ? __ldg(G.B + (size_t)t * G.ldb + j0 + bc) : 0u;
        }
        uint4 pl[4];
        tc_split16(x, pl);
#pragma unroll
        for (int b = 0; b < 4; ++b)
            *reinterpret_cast<uint4 *>(dst + q * (4 * kTcBN * 16) + (b * kTcBN + bc) * 16) = pl[b];
    }
}

__global__ void __launch_bounds__(kTcThreads, 1)
k_tc_gemm_modp(const TcGemmArgs G)
{
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = tc_smem_u32(tc_smem);
    const uint32_t bar0 = smem0 + kTcStages * kTcStageBytes;      /* full[3] | empty[3] | accum */
    const uint32_t i0 = blockIdx.x * kTcBM, j0 = blockIdx.y * kTcBN;
    const uint32_t nkb = (G.K + kTcBK - 1) / kTcBK;

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc_mbar_init(bar0 + 8 * s, 1);
            tc_mbar_init(bar0 + 8 * (kTcStages + s), 1);
        }
        tc_mbar_init(bar0 + 8 * (2 * kTcStages), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tc_smem_u32(&s_tmem)), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp < 4) {
        /* the accumulators start at zero: the column blocks overlap between the four MMAs of a k-step */
        const uint32_t tl = tmem + ((32u * warp) << 16);
#pragma unroll 8
        for (uint32_t c = 0; c < 7 * 64; c += 8) tc_st8_zero(tl + c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    bool alive = true;
    if (warp == 5) {
        /* ===== producer: one thread, bulk copies of whole pre-split stages ===== */
        if (lane == 0) {
            const uint8_t *srcA = G.Asplit + (size_t)blockIdx.x * nkb * kTcStageA;
            const uint8_t *srcB = G.Bsplit + (size_t)blockIdx.y * nkb * kTcStageB;
            for (uint32_t kb = 0; kb < nkb && alive; ++kb) {
                const uint32_t s = kb % kTcStages, ph = (kb / kTcStages) & 1u;
                alive = tc_mbar_wait(bar0 + 8 * (kTcStages + s), ph ^ 1u, G.err);
                if (!alive) break;
                const uint32_t full = bar0 + 8 * s, dst = smem0 + s * kTcStageBytes;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(full), "r"(kTcStageBytes) : "memory");
#pragma unroll
                for (uint32_t c = 0; c < kTcStageA; c += kTcBulkBytes)
                    tc_bulk_g2s(dst + c, srcA + (size_t)kb * kTcStageA + c, kTcBulkBytes, full);
#pragma unroll
                for (uint32_t c = 0; c < kTcStageB; c += kTcBulkBytes)
                    tc_bulk_g2s(dst + kTcStageA + c, srcB + (size_t)kb * kTcStageB + c, kTcBulkBytes, full);
            }
        }
    } else if (warp == 4 && lane == 0) {
        /* ===== one thread issues every MMA ===== */
        for (uint32_t kb = 0; kb < nkb && alive; ++kb) {
            const uint32_t s = kb % kTcStages, ph = (kb / kTcStages) & 1u;
            alive = tc_mbar_wait(bar0 + 8 * s, ph, G.err);
            if (!alive) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem0 + s * kTcStageBytes, b0 = a0 + kTcStageA;
#pragma unroll
            for (uint32_t ks = 0; ks < kTcBK / 32; ++ks) {
                const uint64_t db = tc_smem_desc(b0 + ks * 2 * (4 * kTcBN * 16), 4 * kTcBN * 16, 128);
#pragma unroll
                for (uint32_t a = 0; a < 4; ++a) {
                    const uint64_t da = tc_smem_desc(a0 + a * kTcPlaneA + ks * 2 * (kTcBM * 16), kTcBM * 16, 128);
                    tc_mma_i8(tmem + 64 * a, da, db, kTcIdesc);
                }
            }
            tc_commit(bar0 + 8 * (kTcStages + s));                   /* stage free again once these MMAs have read it */
        }
        tc_commit(bar0 + 8 * (2 * kTcStages));                       /* accumulators complete */
    }

    if (warp < 4) {
        /* ===== epilogue: TMEM -> registers -> recombine mod p -> compare / store ===== */
        /* verify mode: the 64 entries of D this thread compares with are fetched while the MMAs still run */
        uint32_t dv[kTcBN];
        if (!G.C) {
#pragma unroll
            for (int e = 0; e < kTcBN; ++e)
                dv[e] = (j0 + e < G.N) ? __ldg(G.D + (size_t)(j0 + e) * G.ldd + i0 + tid) : 0u;
        }
        bool ok = tc_mbar_wait(bar0 + 8 * (2 * kTcStages), 0, G.err);
        ok = __all_sync(0xffffffffu, ok);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tl = tmem + ((32u * warp) << 16);
        const uint32_t row = i0 + tid;
        uint32_t bad = 0;
#pragma unroll
        for (uint32_t cb = 0; cb < kTcBN; cb += 8) {
            if (!ok) break;
            uint32_t S[7][8];
#pragma unroll
            for (int s = 0; s < 7; ++s) tc_ld8(tl + 64 * s + cb, S[s]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t j = j0 + cb + e;
                if (j >= G.N) continue;
                const uint64_t lo = (uint64_t)S[0][e] + ((uint64_t)S[1][e] << 8) + ((uint64_t)S[2][e] << 16) + ((uint64_t)S[3][e] << 24);
                const uint64_t hi = (uint64_t)S[4][e] + ((uint64_t)S[5][e] << 8) + ((uint64_t)S[6][e] << 16);
                const uint64_t t = (uint64_t)mod_u64(lo, G.p, G.pinv) + (uint64_t)mod_u64(hi, G.p, G.pinv) * G.w32;
                const uint32_t r = mod_u64(t, G.p, G.pinv);
                if (G.C) G.C[(size_t)j * G.ldc + row] = r;
                else bad |= r != dv[cb + e];
            }
        }
        if (bad) atomicOr(G.mismatch, 1u);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTcTmemCols) : "memory");
    }
}

/* ====================================================================================================== */
/* Schur complement on the tensor cores:  D' = X_D + M * Bneg  (mod p)                                      */
/* ====================================================================================================== */
/* After the dataflow solve has produced the multipliers M = V[:, pivot columns] (nrl x ncl), the columns without
 * pivot are  V[ncl + j] = X[ncl + j] + sum_t (p - a_tj) V[t]  -- a dense-times-sparse product with NO dependencies.
 * B = (p - a_tj) (ncl x ncr) is only ~5 % dense, but its non-zeros cluster: on katsura-12 round 9 only 44 % of its
 * 64 x 64 tiles contain a non-zero.  The product therefore runs as a tiled int8-limb GEMM (same pipeline as
 * k_tc_gemm_modp) over the NON-ZERO tiles only:
 *   k_schur_b_build  scatters the CSC entries of the non-pivot columns into zeroed stage images of B and flags the
 *                    (column block, k-block) tiles that received an entry;
 *   k_schur_lists    compacts, per column block, the list of flagged k-blocks;
 *   k_schur_a_split  writes the stage images of M (from V, chunk layout [chunk][column][Rc]) and flags all-zero tiles
 *                    (the lower rows come clustered: many multiplier tiles are zero);
 *   k_schur_gemm     one CTA per (row block, column block, list segment of <= 128 k-blocks = 8192 deep, the
 *                    accumulator bound): bulk-copies only the listed stages whose M tile is non-zero, 4 MMAs per
 *                    32-deep k-step, writes the partial product;
 *   k_schur_combine  D'[j] = X[j] + sum of the partial products (mod p).
 * The executed multiply-adds are 16 int8 products per residue product over the non-zero tiles only. */

constexpr uint32_t kSchurSeg = 128;              /* k-blocks per list segment: 128 x 64 = 8192 = kTcMaxK */

struct SchurArgs {
    const uint32_t *V;          /* [G][nc][Rc]                                                  */
    uint32_t        nc, ncl, ncr, Rc, G;
    const uint32_t *col_ptr;    /* CSC of the pivot rows                                        */
    const uint2    *ent;        /* { pivot t, p - a_tj }                                        */
    uint8_t        *Asplit;     /* [mblk][nkb] stage images of M                                */
    uint8_t        *Bsplit;     /* [nblk][nkb] stage images of B (zeroed before k_schur_b_build) */
    uint8_t        *flagA;      /* [mblk][nkb]                                                  */
    uint8_t        *flagB;      /* [nblk][nkb] (zeroed)                                         */
    uint32_t       *list;       /* [nblk][nkb] compacted k-blocks                               */
    uint32_t       *list_cnt;   /* [nblk]                                                       */
    uint32_t        mblk, nblk, nkb, nseg;
    uint32_t       *P;          /* [nseg][ncr][ldp] partial products                            */
    uint64_t        ldp;        /* mblk * 128                                                   */
    uint32_t       *out;        /* D' [ncr][out_ld]                                             */
    uint64_t        out_ld;
    uint32_t        chunk0;     /* first chunk of the group (row offset chunk0 * Rc in out)     */
    uint32_t        p, w32;
    uint64_t        pinv;
    uint32_t       *err;
};

/* one warp per non-pivot column: its CSC entries -> limb bytes in the stage images of B */
__global__ void k_schur_b_build(const SchurArgs S)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t jj = warp; jj < S.ncr; jj += nwarps) {
        const uint32_t nb = jj / kTcBN, c = jj % kTcBN;
        const uint32_t e0 = S.col_ptr[S.ncl + jj], e1 = S.col_ptr[S.ncl + jj + 1];
        for (uint32_t e = e0 + lane; e < e1; e += 32) {
            const uint2 en = __ldg(&S.ent[e]);
            const uint32_t t = en.x, kb = t / kTcBK, q = (t % kTcBK) / 16, kk = t % 16;
            uint8_t *dst = S.Bsplit + ((size_t)nb * S.nkb + kb) * kTcStageB + q * (4 * kTcBN * 16) + c * 16 + kk;
#pragma unroll
            for (int b = 0; b < 4; ++b) dst[b * kTcBN * 16] = (uint8_t)(en.y >> (8 * b));
            S.flagB[(size_t)nb * S.nkb + kb] = 1;
        }
    }
}

/* one CTA per column block: ordered list of its flagged k-blocks */
__global__ void __launch_bounds__(256) k_schur_lists(const SchurArgs S)
{
    __shared__ uint32_t s_base;
    const uint32_t nb = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint32_t s_w[8];
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint32_t k0 = 0; k0 < S.nkb; k0 += 256) {
        const uint32_t kb = k0 + tid;
        const bool f = kb < S.nkb && S.flagB[(size_t)nb * S.nkb + kb];
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_w[warp] = __popc(bal);
        __syncthreads();
        uint32_t off = s_base;
        for (uint32_t w = 0; w < warp; ++w) off += s_w[w];
        if (f) S.list[(size_t)nb * S.nkb + off + __popc(bal & ((1u << lane) - 1))] = kb;
        __syncthreads();
        if (tid == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += s_w[w]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) S.list_cnt[nb] = s_base;
}

/* stage images of M = V[:, pivot columns]; grid (mblk, nkb), 128 threads */
__global__ void __launch_bounds__(128) k_schur_a_split(const SchurArgs S)
{
    const uint32_t tid = threadIdx.x, mb = blockIdx.x, kb = blockIdx.y, t0 = kb * kTcBK;
    const uint32_t per = S.Rc / kTcBM, chunk = mb / per, half = mb % per;
    const uint32_t *Vc = S.V + (size_t)chunk * S.nc * S.Rc + half * kTcBM + tid;
    uint8_t *dst = S.Asplit + ((size_t)mb * S.nkb + kb) * kTcStageA;
    uint32_t any = 0;
#pragma unroll 1
    for (uint32_t q = 0; q < 4; ++q) {
        uint32_t x[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const uint32_t t = t0 + 16 * q + kk;
            x[kk] = t < S.ncl ? __ldcg(Vc + (size_t)t * S.Rc) : 0u;
            any |= x[kk];
        }
        uint4 pl[4];
        tc_split16(x, pl);
#pragma unroll
        for (int a = 0; a < 4; ++a)
            *reinterpret_cast<uint4 *>(dst + a * kTcPlaneA + q * (kTcBM * 16) + tid * 16) = pl[a];
    }
    any = __syncthreads_or(any != 0);
    if (tid == 0) S.flagA[(size_t)mb * S.nkb + kb] = (uint8_t)(any != 0);
}

__global__ void __launch_bounds__(kTcThreads, 1)
k_schur_gemm(const SchurArgs S)
{
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ uint32_t s_tmem, s_n;
    __shared__ uint16_t s_kb[kSchurSeg];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = tc_smem_u32(tc_smem);
    const uint32_t bar0 = smem0 + kTcStages * kTcStageBytes;
    const uint32_t mb = blockIdx.x, nb = blockIdx.y, seg = blockIdx.z;
    const uint32_t j0 = nb * kTcBN;
    uint32_t *Pz = S.P + (size_t)seg * S.ncr * S.ldp;

    /* the k-blocks of this segment whose M tile is non-zero (ordered compaction by warp 0) */
    if (warp == 0) {
        const uint32_t cnt = S.list_cnt[nb];
        const uint32_t a = min(cnt, seg * kSchurSeg), b = min(cnt, a + kSchurSeg);
        uint32_t n = 0;
        for (uint32_t k0 = a; k0 < b; k0 += 32) {
            const uint32_t idx = k0 + lane;
            uint32_t kb = 0;
            bool f = false;
            if (idx < b) { kb = S.list[(size_t)nb * S.nkb + idx]; f = S.flagA[(size_t)mb * S.nkb + kb] != 0; }
            const uint32_t bal = __ballot_sync(0xffffffffu, f);
            if (f) s_kb[n + __popc(bal & ((1u << lane) - 1))] = (uint16_t)kb;
            n += __popc(bal);
        }
        if (lane == 0) s_n = n;
    }
    __syncthreads();
    const uint32_t nk = s_n;
    if (nk == 0) {                                   /* nothing to multiply: the partial product is zero */
        if (tid < kTcBM)
            for (uint32_t e = 0; e < kTcBN; ++e)
                if (j0 + e < S.ncr) Pz[(size_t)(j0 + e) * S.ldp + mb * kTcBM + tid] = 0u;
        return;
    }

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc_mbar_init(bar0 + 8 * s, 1);
            tc_mbar_init(bar0 + 8 * (kTcStages + s), 1);
        }
        tc_mbar_init(bar0 + 8 * (2 * kTcStages), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tc_smem_u32(&s_tmem)), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (warp < 4) {
        const uint32_t tl = tmem + ((32u * warp) << 16);
#pragma unroll 8
        for (uint32_t c = 0; c < 7 * 64; c += 8) tc_st8_zero(tl + c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    bool alive = true;
    if (warp == 5) {
        if (lane == 0) {
            const uint8_t *srcA = S.Asplit + (size_t)mb * S.nkb * kTcStageA;
            const uint8_t *srcB = S.Bsplit + (size_t)nb * S.nkb * kTcStageB;
            for (uint32_t it = 0; it < nk && alive; ++it) {
                const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
                alive = tc_mbar_wait(bar0 + 8 * (kTcStages + s), ph ^ 1u, S.err);
                if (!alive) break;
                const uint32_t kb = s_kb[it];
                const uint32_t full = bar0 + 8 * s, dst = smem0 + s * kTcStageBytes;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(full), "r"(kTcStageBytes) : "memory");
#pragma unroll
                for (uint32_t c = 0; c < kTcStageA; c += kTcBulkBytes)
                    tc_bulk_g2s(dst + c, srcA + (size_t)kb * kTcStageA + c, kTcBulkBytes, full);
#pragma unroll
                for (uint32_t c = 0; c < kTcStageB; c += kTcBulkBytes)
                    tc_bulk_g2s(dst + kTcStageA + c, srcB + (size_t)kb * kTcStageB + c, kTcBulkBytes, full);
            }
        }
    } else if (warp == 4 && lane == 0) {
        for (uint32_t it = 0; it < nk && alive; ++it) {
            const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
            alive = tc_mbar_wait(bar0 + 8 * s, ph, S.err);
            if (!alive) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem0 + s * kTcStageBytes, b0 = a0 + kTcStageA;
#pragma unroll
            for (uint32_t ks = 0; ks < kTcBK / 32; ++ks) {
                const uint64_t db = tc_smem_desc(b0 + ks * 2 * (4 * kTcBN * 16), 4 * kTcBN * 16, 128);
#pragma unroll
                for (uint32_t a = 0; a < 4; ++a) {
                    const uint64_t da = tc_smem_desc(a0 + a * kTcPlaneA + ks * 2 * (kTcBM * 16), kTcBM * 16, 128);
                    tc_mma_i8(tmem + 64 * a, da, db, kTcIdesc);
                }
            }
            tc_commit(bar0 + 8 * (kTcStages + s));
        }
        tc_commit(bar0 + 8 * (2 * kTcStages));
    }

    if (warp < 4) {
        bool ok = tc_mbar_wait(bar0 + 8 * (2 * kTcStages), 0, S.err);
        ok = __all_sync(0xffffffffu, ok);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tl = tmem + ((32u * warp) << 16);
        const uint32_t row = mb * kTcBM + tid;
#pragma unroll 1
        for (uint32_t cb = 0; cb < kTcBN; cb += 8) {
            if (!ok) break;
            uint32_t A7[7][8];
#pragma unroll
            for (int s = 0; s < 7; ++s) tc_ld8(tl + 64 * s + cb, A7[s]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t j = j0 + cb + e;
                if (j >= S.ncr) continue;
                const uint64_t lo = (uint64_t)A7[0][e] + ((uint64_t)A7[1][e] << 8) + ((uint64_t)A7[2][e] << 16) + ((uint64_t)A7[3][e] << 24);
                const uint64_t hi = (uint64_t)A7[4][e] + ((uint64_t)A7[5][e] << 8) + ((uint64_t)A7[6][e] << 16);
                const uint64_t t = (uint64_t)mod_u64(lo, S.p, S.pinv) + (uint64_t)mod_u64(hi, S.p, S.pinv) * S.w32;
                Pz[(size_t)j * S.ldp + row] = mod_u64(t, S.p, S.pinv);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTcTmemCols) : "memory");
    }
}

/* D'[j][row] = X[j][row] + sum over the segments of the partial products; grid (ncr, rows / 256) */
__global__ void __launch_bounds__(256) k_schur_combine(const SchurArgs S)
{
    const uint32_t row = blockIdx.y * 256 + threadIdx.x, j = blockIdx.x;
    if (row >= S.G * S.Rc) return;
    const uint32_t chunk = row / S.Rc, r = row % S.Rc;
    uint64_t acc = S.V[((size_t)chunk * S.nc + S.ncl + j) * S.Rc + r];
    for (uint32_t z = 0; z < S.nseg; ++z) acc += S.P[((size_t)z * S.ncr + j) * S.ldp + row];
    S.out[(size_t)j * S.out_ld + (size_t)S.chunk0 * S.Rc + row] = mod_u64(acc, S.p, S.pinv);
}

} /* namespace f4la */
