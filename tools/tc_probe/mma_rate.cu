// tcgen05.mma issue-rate microbenchmark: cycles per kind::tf32 MMA (M=128, K=8) for N in {32,64,128,256}, A from smem
// (SS) or from TMEM (TS).  One CTA, one issuing thread, R back-to-back MMAs on the same accumulator, one commit.
#include <stdio.h>
#include <stdlib.h>
#include "tc05.cuh"
using namespace tc05;

__global__ void __launch_bounds__(128) rate_kernel(int N, int ts, int reps, int nacc, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (128 * 8 + 256 * 8); e += 128) reinterpret_cast<float*>(smem)[e] = 1.0f;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(2, 128, N, 0, 0);
        const uint64_t ad = make_smem_desc(smem_u32(smem), 128 * 16, 128, 0);
        const uint64_t bd = make_smem_desc(smem_u32(smem) + 128 * 8 * 4, N * 16, 128, 0);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t dcol = (uint32_t)((r % nacc) * 64);
            if (ts) mma_tf32_ts(tb + dcol, tb + 256 + (r & 7) * 8, bd, idesc, 1);
            else mma_tf32_ss(tb + dcol, ad, bd, idesc, 1);
        }
        long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int ts = 0; ts < 2; ++ts)
      for (int nacc : {1, 2, 4})
        for (int N : {32, 64}) {
            int reps = 2000;
            rate_kernel<<<1, 128, 64 * 1024>>>(N, ts, reps, nacc, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%s nacc=%d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (math floor %d)\n", ts ? "TS" : "SS", nacc, N,
                   (double)h[0] / reps, (double)h[1] / reps, 128 * N / 256);
        }
    return 0;
}
