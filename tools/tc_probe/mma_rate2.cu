// tcgen05.mma issue cost with a warp-uniform issue loop (all lanes run the loop, elect.sync guards the MMA), fully
// unrolled descriptor stepping.  Variants: N=32 TS (as in G1 of the mean-shift kernel), 16 k-steps x 3 terms.
#include <stdio.h>
#include "tc05.cuh"
using namespace tc05;

// (elect_one comes from the shipped tc05.cuh)
__device__ __forceinline__ void mma_ts_nopred(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int MODE>
__global__ void __launch_bounds__(128) rate_kernel(int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 16 * 1024; e += 128) reinterpret_cast<float*>(smem)[e] = 1.0f;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = tmem_base_s;
    if (warp == 1) {
        const uint32_t idesc = make_idesc(2, 128, N, 0, 0);
        const uint64_t bd0 = make_smem_desc(smem_u32(smem), N * 16, 128, 0);
        const uint64_t bs0 = make_smem_desc(smem_u32(smem) + 16384, N * 16, 128, 0);
        long long t0 = clock64();
        if (MODE == 0) {                 // lane 0 only, runtime loop (as in the v4 kernel)
            if ((tid & 31) == 0)
                for (int r = 0; r < reps; ++r)
#pragma unroll
                    for (int ks = 0; ks < 16; ++ks) {
                        const uint64_t db = bd0 + (uint64_t)(ks * ((2 * N * 16) >> 4)), ds = bs0 + (uint64_t)(ks * ((2 * N * 16) >> 4));
                        mma_tf32_ts(tb, tb + 256 + ks * 8, db, idesc, 1);
                        mma_tf32_ts(tb, tb + 128 + ks * 8, ds, idesc, 1);
                        mma_tf32_ts(tb, tb + 128 + ks * 8, db, idesc, 1);
                    }
        } else {                         // whole warp runs the loop, one elected lane issues
            const bool leader = elect_one();
            for (int r = 0; r < reps; ++r)
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    const uint64_t db = bd0 + (uint64_t)(ks * ((2 * N * 16) >> 4)), ds = bs0 + (uint64_t)(ks * ((2 * N * 16) >> 4));
                    if (leader) {
                        mma_ts_nopred(tb, tb + 256 + ks * 8, db, idesc, 1);
                        mma_ts_nopred(tb, tb + 128 + ks * 8, ds, idesc, 1);
                        mma_ts_nopred(tb, tb + 128 + ks * 8, db, idesc, 1);
                    }
                }
        }
        long long t1 = clock64();
        if ((tid & 31) == 0) { mma_commit(&bar); mbar_wait(&bar, 0); out[0] = t1 - t0; out[1] = clock64() - t0; }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int MODE> void run(long long* d) {
    int reps = 50;
    cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    rate_kernel<N, MODE><<<1, 128, 200 * 1024>>>(reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d mode=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (math floor %d)\n", N, MODE, (double)h[0] / (reps * 48),
           (double)h[1] / (reps * 48), 128 * N / 256);
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    run<16, 1>(d); run<32, 0>(d); run<32, 1>(d); run<64, 0>(d); run<64, 1>(d); run<128, 1>(d); run<256, 1>(d);
    return 0;
}
