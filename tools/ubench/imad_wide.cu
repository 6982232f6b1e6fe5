// Micro-benchmark: throughput of the two candidate multiply-accumulate schemes of k_pull_flow
//   A: two IMAD.WIDE.U32 per product (coefficient split in 16-bit halves, u64 accumulators)
//   B: one IMAD.WIDE.U32 with carry-out + one add-with-carry into a third word (96-bit accumulator)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mac96(uint32_t &lo, uint32_t &hi, uint32_t &top, uint32_t v, uint32_t c)
{
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(lo), "+r"(hi), "+r"(top) : "r"(v), "r"(c));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, int iters)
{
    uint32_t v[8], c = seed | 1u;
    for (int i = 0; i < 8; ++i) v[i] = seed * (i + 3) + threadIdx.x;
    uint64_t lo[8] = {0}, hi[8] = {0};
    uint32_t a0[8] = {0}, a1[8] = {0}, a2[8] = {0};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                lo[i] += (uint64_t)v[i] * (c & 0xffffu);
                hi[i] += (uint64_t)v[i] * (c >> 16);
            } else {
                mac96(a0[i], a1[i], a2[i], v[i], c);
            }
        }
        c = c * 1664525u + 1013904223u;
    }
    uint32_t r = 0;
    for (int i = 0; i < 8; ++i) r ^= (uint32_t)lo[i] ^ (uint32_t)(hi[i] >> 7) ^ a0[i] ^ a1[i] ^ a2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main()
{
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, 12345u + rep, iters);
            else k<1><<<148 * 8, 256>>>(d, 12345u + rep, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double macs = 148.0 * 8 * 256 * 8 * iters;
            printf("mode %d: %.3f ms  %.3e MAC/s  (%.2f MAC/clk/SM at 1.9 GHz)\n", mode, ms, macs / ms * 1e3,
                   macs / ms * 1e3 / 148 / 1.9e9);
        }
    return 0;
}
