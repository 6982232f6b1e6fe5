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
 * Operands are K-major, no swizzle (8 x 16-byte core matrices): filling them is a byte shuffle of u32 residues
 * that TMA cannot do, so four producer warps load the residues (coalesced), split the limbs with PRMT and store
 * 16-byte rows (conflict free); a 3-stage mbarrier ring decouples them from the single MMA-issuing thread;
 * tcgen05.commit hands the stages back and signals the epilogue, which reads the accumulators with tcgen05.ld.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "f4la_kernels.cuh"

namespace f4la {

constexpr int kTcBM = 128, kTcBN = 64, kTcBK = 64, kTcStages = 3;
constexpr int kTcProducers = 128;                 /* warps 0-3: producers, then epilogue */
constexpr int kTcThreads = kTcProducers + 32;     /* warp 4: TMEM allocation + MMA issue */
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
            tc_mbar_init(bar0 + 8 * s, kTcProducers);
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
    if (warp < 4) {
        /* ===== producers: residues -> limb planes of the stage ===== */
        const uint32_t bc = tid & 63, bhalf = tid >> 6;             /* B: column of the tile, which half of the k-block */
        const bool b_in = j0 + bc < G.N;
        for (uint32_t kb = 0; kb < nkb && alive; ++kb) {
            const uint32_t s = kb % kTcStages, ph = (kb / kTcStages) & 1u;
            alive = tc_mbar_wait(bar0 + 8 * (kTcStages + s), ph ^ 1u, G.err);
            if (!alive) break;
            uint8_t *stA = tc_smem + (size_t)s * kTcStageBytes, *stB = stA + kTcStageA;
            const uint32_t t0 = kb * kTcBK;
#pragma unroll 1
            for (uint32_t q = 0; q < 4; ++q) {                      /* A: row tid, k-group q (16 columns of A) */
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
                    *reinterpret_cast<uint4 *>(stA + a * kTcPlaneA + q * (kTcBM * 16) + tid * 16) = pl[a];
            }
#pragma unroll 1
            for (uint32_t q = 2 * bhalf; q < 2 * bhalf + 2; ++q) {  /* B: column bc, k-groups of this half */
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
                    *reinterpret_cast<uint4 *>(stB + q * (4 * kTcBN * 16) + (b * kTcBN + bc) * 16) = pl[b];
            }
            /* generic-proxy stores -> visible to the tensor core (async proxy), then hand the stage over */
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_mbar_arrive(bar0 + 8 * s);
        }
    } else if (lane == 0) {
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
        bool ok = tc_mbar_wait(bar0 + 8 * (2 * kTcStages), 0, G.err);
        ok = __all_sync(0xffffffffu, ok);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tl = tmem + ((32u * warp) << 16);
        const uint32_t row = i0 + tid;
        uint32_t bad = 0;
        for (uint32_t cb = 0; cb < kTcBN && ok; cb += 8) {
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
                else bad |= r != __ldg(G.D + (size_t)j * G.ldd + row);
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

} /* namespace f4la */
