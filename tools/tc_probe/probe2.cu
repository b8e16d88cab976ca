// tcgen05 cta_group::2 probe: D[256 x N] = A[256 x K] . B[N x K]^T with a CTA PAIR (cluster of 2).
//   * A: each CTA holds its 128 rows in its own TMEM (tcgen05.st), like the mean-shift kernels' Y operand
//   * B: each CTA stages N/2 rows in its own shared memory as 128B-swizzled K-major slabs (what TMA writes)
//   * the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 (M = 256) and a multicast commit to a barrier in BOTH CTAs
//   * each CTA reads its 128 rows of D from its own TMEM
// Isolates the 2-CTA primitives used by csrc/meanshift_tma.cu (PN_MS_TMA_CG=2): alloc / dealloc cta_group::2, remote
// mbarrier arrive through mapa, MMA with the B operand split over two CTAs, commit .multicast::cluster.
//   mode 0: B staged by ordinary threads;  mode 1: B fetched by cp.async.bulk.tensor .cta_group::2 (completion on the
//   LEADER's barrier from both CTAs)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>
#include "tc05.cuh"
using namespace tc05;

constexpr int N = 64, K = 64;     // per CTA: 32 rows of B, 2 slabs of [32 rows][128 B]

__device__ __forceinline__ uint32_t sw128_off(int r, int c, int rows) {
    const int slab = c >> 5, cc = c & 31;
    return (uint32_t)(slab * rows * 128 + r * 128 + ((((cc >> 2) ^ (r & 7)) & 7) << 4) + (cc & 3) * 4);
}

struct Bars { uint64_t b_ready, b_full, done; };

__global__ void __launch_bounds__(128) probe2_kernel(const __grid_constant__ CUtensorMap mB, const float* A, const float* B,
                                                     float* out, int mode) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sB = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0) tmem_alloc2(&tmem_base_s, 256);
    if (tid == 0) {
        mbar_init(&bars.b_ready, 2 * 128);      // leader: both CTAs' threads (A in TMEM, B in smem)
        mbar_init(&bars.b_full, 1);             // leader: TMA bytes of both CTAs (mode 1)
        mbar_init(&bars.done, 1);               // every CTA: multicast commit
        mbar_fence_init();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    // A rows of this CTA -> TMEM columns [128, 128 + K)
    for (int c0 = 0; c0 < K; c0 += 32) {
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(A[(size_t)(rank * 128 + tid) * K + c0 + j]);
        tmem_st32(tb + 128 + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tmem_st_wait();
    // B rows [32 rank, 32 rank + 32) of the N x K matrix -> two 128B-swizzled slabs
    if (mode == 0) {
        for (int e = tid; e < (N / 2) * K; e += 128) {
            const int r = e / K, k = e % K;
            *reinterpret_cast<float*>(sB + sw128_off(r, k, N / 2)) = B[(size_t)(rank * (N / 2) + r) * K + k];
        }
        fence_async_smem();
    } else if (tid == 0) {
        const uint32_t full = mapa_u32(smem_u32(&bars.b_full), 0);
        if (rank == 0) mbar_arrive_expect_tx(&bars.b_full, N * K * 4);
        for (int sl = 0; sl < K / 32; ++sl)
            tma_load_3d_2sm(sB + sl * (N / 2) * 128, &mB, full, 32 * sl, (int)rank * (N / 2), 0);
    }
    tc_fence_before();
    mbar_arrive_cluster(mapa_u32(smem_u32(&bars.b_ready), 0));
    if (rank == 0 && warp == 0) {
        mbar_wait_guarded(&bars.b_ready, 0);
        if (mode == 1) mbar_wait_guarded(&bars.b_full, 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t idesc = make_idesc(2, 256, N, 0, 0);
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint32_t off = (uint32_t)((ks >> 2) * (N / 2) * 128 + (ks & 3) * 32);
                mma_tf32_ts2(tb, tb + 128 + ks * 8, make_smem_desc(smem_u32(sB) + off, 16, 1024, 2), idesc, ks > 0);
            }
            mma_commit2_mc(&bars.done, (uint16_t)3);
        }
        __syncwarp();
    }
    mbar_wait_guarded(&bars.done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(size_t)(rank * 128 + tid) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tb, 256);
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int M = 256;
    std::vector<float> hA(M * K), hB(N * K), hO(M * N), ref(M * N);
    srand(2);
    auto rnd = []() { return (float)((rand() % 17) - 8) / 8.0f; };
    for (auto& v : hA) v = rnd();
    for (auto& v : hB) v = rnd();
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)hA[i * K + k] * hB[j * K + k];
            ref[i * N + j] = (float)s;
        }
    float *dA, *dB, *dO;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dO, hO.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0xff, hO.size() * 4);
    // B as a 3-D tensor {K, N, 1}, box {32, N/2, 1}, 128B swizzle
    CUtensorMap mB;
    {
        typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
            printf("cuTensorMapEncodeTiled unavailable\n");
            return 1;
        }
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, 1}, strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)N * K * 4};
        cuuint32_t box[3] = {32, N / 2, 1}, es[3] = {1, 1, 1};
        CUresult rc = ((Enc)p)(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc); return 1; }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = (N / 2) * K * 4 + 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe2_kernel, mB, (const float*)dA, (const float*)dB, dO, mode);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe2 mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0, bad_hi = 0;
    for (int i = 0; i < M * N; ++i) {
        double d = fabs(hO[i] - ref[i]);
        if (!(d <= 1e-3)) { ++bad; if (i >= 128 * N) ++bad_hi; }
        if (d > maxerr) maxerr = d;
    }
    printf("probe2 (cta_group::2) mode %d: max err %.4g, mismatches %d / %d (%d of them in the peer CTA's rows)  "
           "out[0..1] = %g %g, out[128*N..] = %g %g, ref %g %g / %g %g\n", mode, maxerr, bad, M * N, bad_hi, hO[0], hO[1],
           hO[128 * N], hO[128 * N + 1], ref[0], ref[1], ref[128 * N], ref[128 * N + 1]);
    return 0;
}
