// tcgen05 probe: D[128 x N] = A[128 x K] . B[N x K]^T (kind::tf32, fp32 accumulate), operands staged by ordinary
// threads in the canonical no-swizzle K-major core-matrix layout.  Validates descriptor encoding, TMEM alloc/ld,
// commit/mbarrier, and (mode 1) the A-from-TMEM + MN-major-B path used by the fused mean-shift kernel.
//   mode bit0: A from TMEM (tcgen05.st) instead of smem;  bit1: B given as Bt[K x N] (MN-major descriptor)
//   mode bit2: operands are FULL-mantissa fp32 values; the host computes the product once with operands truncated to
//              tf32 (low 13 bits cleared) and once rounded to nearest: tells whether the tensor core truncates, i.e.
//              whether raw fp32 data can serve as the "big" part of a split-TF32 operand without a separate masked copy
//   mode bit3: B is staged the way TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B: slabs of [rows][32 floats = 128 B],
//              16-byte chunk index XOR (row & 7); read with a 128B-swizzle descriptor -- K-major (rows = n, slab = 32 k;
//              SBO = 1024, K step = +32 B inside the atom, next slab after 4 steps) or, with bit1, MN-major (rows = k,
//              slab = 32 n; LBO = slab stride, SBO = 1024, K step = 8 rows = +1024 B).  Validates the round-2 design
//              where ONE TMA-loaded X tile serves both products of the mean-shift kernels.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>
#include <cstring>
#include <algorithm>
#include "tc05.cuh"
using namespace tc05;

constexpr int M = 128;

// K-major no-swizzle staging: element (r, k) of a [rows x K] fp32 matrix -> byte offset
//   chunk c = k/4 (16 B), layout [c][r/8][r%8][16B]:  LBO (K direction) = rows*16, SBO (8-row groups) = 128
__device__ __forceinline__ uint32_t kmaj_off(int r, int k, int rows) {
    return (uint32_t)((k >> 2) * rows * 16 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4);
}
// MN-major no-swizzle staging for B given as Bt[K x N] row-major (N contiguous): element (k, n)
//   core matrix = 8 k-rows x 16 B (4 n): layout [n/4][k/8][k%8][16B]: stride between n-chunks = K*16 ("leading", MN
//   direction), stride between 8-k groups = 128 ("stride", K direction)
__device__ __forceinline__ uint32_t mnmaj_off(int k, int n, int K) {
    return (uint32_t)((n >> 2) * K * 16 + (k >> 3) * 128 + (k & 7) * 16 + (n & 3) * 4);
}

// 128B-swizzled slab layout (what TMA SWIZZLE_128B produces for a box of [rows][32 floats]): element (r, c) of slab s
__device__ __forceinline__ uint32_t sw128_off(int r, int c, int rows) {
    const int slab = c >> 5, cc = c & 31;
    return (uint32_t)(slab * rows * 128 + r * 128 + ((((cc >> 2) ^ (r & 7)) & 7) << 4) + (cc & 3) * 4);
}

__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* out, int N, int K, int mode,
                                                    uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b,
                                                    uint32_t kstep_a, uint32_t kstep_b) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);                 // 128*K*4
    float* sB = reinterpret_cast<float*>(smem + M * K * 4);     // N*K*4
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    // stage operands
    for (int e = tid; e < M * K; e += 128) {
        int r = e / K, k = e % K;
        *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sA) + kmaj_off(r, k, M)) = A[e];
    }
    if (mode & 8) {
        if (!(mode & 2)) {                           // B[N][K]: rows = n, columns = k
            for (int e = tid; e < N * K; e += 128) {
                int r = e / K, k = e % K;
                *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sB) + sw128_off(r, k, N)) = B[e];
            }
        } else {                                     // Bt[K][N]: rows = k, columns = n
            for (int e = tid; e < K * N; e += 128) {
                int k = e / N, n = e % N;
                *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sB) + sw128_off(k, n, K)) = B[e];
            }
        }
    } else if (!(mode & 2)) {
        for (int e = tid; e < N * K; e += 128) {
            int r = e / K, k = e % K;
            *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sB) + kmaj_off(r, k, N)) = B[e];
        }
    } else {
        for (int e = tid; e < K * N; e += 128) {     // B is Bt[K][N]
            int k = e / N, n = e % N;
            *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sB) + mnmaj_off(k, n, K)) = B[e];
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t d_tmem = tbase;            // columns [0, N)
    const uint32_t a_tmem = tbase + 128;      // columns [128, 128+K) for mode 1
    if (mode & 1) {
        // every thread writes its row of A into TMEM (lane = 32*warp + lane)
        for (int c0 = 0; c0 < K; c0 += 32) {
            uint32_t v[32];
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(A[tid * K + c0 + j]);
            tmem_st32(a_tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (tid == 0) {
        const uint32_t idesc = make_idesc(2, M, N, 0, (mode & 2) ? 1 : 0);
        for (int ks = 0; ks < K / 8; ++ks) {
            uint32_t boff = ks * kstep_b;
            if ((mode & 8) && !(mode & 2)) boff = (ks >> 2) * (N * 128) + (ks & 3) * 32;   // K-major SW128: slab, then +32 B
            uint64_t bd = make_smem_desc(smem_u32(sB) + boff, lbo_b, sbo_b, (mode & 8) ? 2 : 0);
            if (!(mode & 1)) {
                uint64_t ad = make_smem_desc(smem_u32(sA) + ks * kstep_a, lbo_a, sbo_a, 0);
                mma_tf32_ss(d_tmem, ad, bd, idesc, ks > 0);
            } else {
                mma_tf32_ts(d_tmem, a_tmem + ks * 8, bd, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(d_tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    int variant = argc > 2 ? atoi(argv[2]) : 0;
    int N = 64, K = 64;
    std::vector<float> hA(M * K), hB(N * K), hO(M * N), ref(M * N);
    srand(1);
    auto rnd = [&]() {
        if (mode & 4) return (float)rand() / (float)RAND_MAX + 0.5f;   // full 24-bit mantissas, all positive (no cancellation)
        return (float)((rand() % 17) - 8) / 8.0f;                      // exactly representable in tf32
    };
    for (auto& v : hA) v = rnd();
    for (auto& v : hB) v = rnd();
    auto trunc13 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; };
    auto rna13 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&x, &u, 4); return x; };
    std::vector<float> ref_rna(M * N);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0, s2 = 0;
            for (int k = 0; k < K; ++k) {
                const float a = hA[i * K + k], bb = (!(mode & 2) ? hB[j * K + k] : hB[k * N + j]);
                if (mode & 4) { s += (double)trunc13(a) * trunc13(bb); s2 += (double)rna13(a) * rna13(bb); }
                else s += (double)a * bb;
            }
            ref[i * N + j] = (float)s;
            ref_rna[i * N + j] = (float)s2;
        }
    float *dA, *dB, *dO;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dO, hO.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, hO.size() * 4);
    // descriptor parameters: LBO = stride between core matrices along K, SBO = stride between 8-row groups (variant 0),
    // swapped in variant 1.  K step = 2 core matrices along K (8 tf32) -> 2*LBO bytes.
    uint32_t lbo_a = M * 16, sbo_a = 128, lbo_b, sbo_b, kstep_a = 2 * M * 16, kstep_b;
    if (!(mode & 2)) { lbo_b = N * 16; sbo_b = 128; kstep_b = 2 * N * 16; }
    else { lbo_b = 128; sbo_b = K * 16; kstep_b = 128; }     // MN-major: LBO = next 8 k-rows (+128 B), SBO = next 4 n
    if (mode & 8) {
        if (!(mode & 2)) { lbo_b = 16; sbo_b = 1024; kstep_b = 32; }            // K-major SW128 (LBO unused)
        else { lbo_b = K * 128; sbo_b = 1024; kstep_b = 1024; }                 // MN-major SW128: LBO = next 32-n slab
    }
    if (variant == 1 && (mode & 2)) std::swap(lbo_b, sbo_b);
    size_t smem = (size_t)(M * K + N * K) * 4;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dO, N, K, mode, lbo_a, sbo_a, lbo_b, sbo_b, kstep_a, kstep_b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d variant %d: CUDA error %s\n", mode, variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs(hO[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
    printf("mode %d variant %d: max err %.4g, mismatches %d / %d  (out[0..3] = %g %g %g %g, ref %g %g %g %g)\n", mode,
           variant, maxerr, bad, M * N, hO[0], hO[1], hO[2], hO[3], ref[0], ref[1], ref[2], ref[3]);
    if (mode & 4) {
        double e_tr = 0, e_rn = 0;
        for (int i = 0; i < M * N; ++i) { e_tr = fmax(e_tr, fabs(hO[i] - ref[i]) / fabs(ref[i])); e_rn = fmax(e_rn, fabs(hO[i] - ref_rna[i]) / fabs(ref_rna[i])); }
        printf("  full-mantissa operands: max rel diff vs TRUNCATED-operand product %.3e, vs ROUNDED-operand product %.3e "
               "(the smaller one is what the tensor core does; accumulation error itself is ~1e-6)\n", e_tr, e_rn);
    }
    return 0;
}
