// tcgen05.mma issue-to-completion rate for M = 128, K = 16, kind::f16 (fp16 operands, fp32 accumulate) as a function of N, with the
// operands where the production kernels keep them: A in shared memory (SS) or tensor memory (TS), B in shared memory (SWIZZLE_128B,
// K-major).  One CTA per SM, one thread issues `iters` x 4 MMAs (the 4 k-steps of a 64-wide chunk) back to back, commits, waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I said_b200/csrc -o gpurun_out/mma_rate profiles/tools/mma_rate.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "attention_h.cuh"
using namespace said;
using namespace said::hx;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int ts, int pattern, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = base, a_lo = base + 16384, b_hi = base + 32768, b_lo = b_hi + 32768, bar = b_lo + 32768, slot = bar + 8;
    if (threadIdx.x == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_mbar_init();
    }
    if (threadIdx.x < 32) tc::tmem_alloc(slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t dah = tc::make_desc(a_hi), dal = tc::make_desc(a_lo), dbh = tc::make_desc(b_hi), dbl = tc::make_desc(b_lo);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t adv = (uint64_t)(k4 * 2);
                if (pattern == 0) {          // one product per k-step
                    if (ts) mma_f16_ts(tmem, tmem + 448 + 8 * k4, dbh + adv, idesc, 1u);
                    else mma_f16(tmem, dah + adv, dbh + adv, idesc, 1u);
                } else if (pattern == 2) {   // the triple ordered so that consecutive MMAs share one operand: (lo,hi) (hi,hi) (hi,lo)
                    mma_f16(tmem, dal + adv, dbh + adv, idesc, 1u);
                    mma_f16(tmem, dah + adv, dbh + adv, idesc, 1u);
                    mma_f16(tmem, dah + adv, dbl + adv, idesc, 1u);
                } else if (pattern == 3) {   // plane order of the fused feed-forward: (hi,hi) (lo,hi) per k-step, the (hi,lo) products after the chunk
                    mma_f16(tmem, dah + adv, dbh + adv, idesc, 1u);
                    mma_f16(tmem, dal + adv, dbh + adv, idesc, 1u);
                    if (k4 == 3) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) mma_f16(tmem, dah + (uint64_t)(kk * 2), dbl + (uint64_t)(kk * 2), idesc, 1u);
                    }
                } else {                     // the fp16x3 triple
                    if (ts) {
                        mma_f16_ts(tmem, tmem + 448 + 8 * k4, dbh + adv, idesc, 1u);
                        mma_f16_ts(tmem, tmem + 480 + 8 * k4, dbh + adv, idesc, 1u);
                        mma_f16_ts(tmem, tmem + 448 + 8 * k4, dbl + adv, idesc, 1u);
                    } else {
                        mma_f16(tmem, dah + adv, dbh + adv, idesc, 1u);
                        mma_f16(tmem, dal + adv, dbh + adv, idesc, 1u);
                        mma_f16(tmem, dah + adv, dbl + adv, idesc, 1u);
                    }
                }
            }
        }
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, 512);
    }
}

int main() {
    const int smem = 3 * 32768 + 16384 * 2 + 2048;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int Ns[] = {32, 64, 96, 128, 160, 192, 224, 256};
    printf("| N | SS cycles / MMA | TS cycles / MMA | SS fp16x3 triple, cycles / MMA | TS triple | SS triple (lo,hi)(hi,hi)(hi,lo) | SS plane order |\n|---:|---:|---:|---:|---:|---:|---:|\n");
    for (int N : Ns) {
        double r[6];
        for (int mode = 0; mode < 6; ++mode) {
            const int ts = mode < 4 ? (mode & 1) : 0, pattern = mode < 4 ? (mode >> 1) : mode - 2, iters = 256;
            for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<<<148, 128, smem>>>(N, iters, ts, pattern, d);
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0;
            for (int i = 0; i < 148; ++i) s += (double)h[i];
            r[mode] = s / 148 / (iters * 4 * (pattern ? 3 : 1));
        }
        printf("| %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |\n", N, r[0], r[1], r[2], r[3], r[4], r[5]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}
