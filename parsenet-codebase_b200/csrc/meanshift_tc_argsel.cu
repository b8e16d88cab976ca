// Arg-selects of MeanShift.nms on tcgen05: the similarity tiles S = A B^T come from the split-TF32 tensor-core pipeline
// of meanshift_tc_kth.cu (A rows resident in TMEM, 32-row B tiles through shared memory, 48 MMAs per tile), the
// epilogue keeps a running (best value, lowest index) per row instead of a histogram.
// Replaces (reference src/mean_shift.py):
//   mode 0  :146-149  membership = argmin_j (2 - 2 A B^T)                       (A = points, B = shifted points)
//   mode 1  :163-171  argmax_j [ (2 - 2 A B^T < thr) * cnt_j ]                  (A = B = shifted points)
// Same selection rule as ms_argsel_kernel in meanshift.cu: strict improvement while j increases, i.e. ties go to the
// lowest j.  (mode 2, the arg-max over the few kept centres, stays on the FP32-pipe kernel.)
#include "common.cuh"
#include "tc05.cuh"

namespace pn {
namespace mstca {
using namespace tc05;

constexpr int D = 128, BM = 128, BN = 32, NT = 416, NSTAGE = 4;
constexpr int EPI_WARPS = 8, LOAD_WARP0 = 8, MMA_WARP = 12, EPI_THREADS = 256;
constexpr uint32_t C_AB = 0, C_AS = 128, C_D0 = 256, TMEM_COLS = 512;
constexpr int XA_BYTES = BN * D * 4, STAGE_BYTES = 2 * XA_BYTES;     // big + small, 32 KB
constexpr uint32_t XA_LBO = BN * 16, SBO = 128;

struct Bars { uint64_t x_full[NSTAGE], x_empty[NSTAGE], s_full[2], s_empty[2], a_ready; };

// grid (ceil(Ma/128), B)
template <int MODE>
__global__ void __launch_bounds__(NT, 1)
ms_argsel_tc_kernel(const float* __restrict__ A, long long a_stride, int Ma, const float* __restrict__ Bm,
                    long long b_stride, int Nb, const float* __restrict__ cnt, const float* __restrict__ thr,
                    int* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ float part_v[BM];
    __shared__ int part_j[BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * BM;
    const float* Ab = A + (long long)b * a_stride;
    const float* Bb = Bm + (long long)b * b_stride;
    const int ntiles = (Nb + BN - 1) / BN;

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars.x_full[s], 128); mbar_init(&bars.x_empty[s], 1); }
        for (int k = 0; k < 2; ++k) { mbar_init(&bars.s_full[k], 1); mbar_init(&bars.s_empty[k], EPI_THREADS); }
        mbar_init(&bars.a_ready, EPI_THREADS);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;

    if (warp < EPI_WARPS) {
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t la = (uint32_t)(q * 32) << 16;
        const bool ok = (i0 + row) < Ma;
        const float* xr = Ab + (long long)(ok ? i0 + row : 0) * D + 64 * h;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t vb[16], vs[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                float4 v = ok ? *reinterpret_cast<const float4*>(xr + c0 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float big = tf32_hi(f[u]);
                    vb[e + u] = __float_as_uint(big);
                    vs[e + u] = __float_as_uint(f[u] - big);
                }
            }
            tmem_st16(tb + la + C_AB + 64 * h + c0, vb);
            tmem_st16(tb + la + C_AS + 64 * h + c0, vs);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars.a_ready);
        const float* cb = (MODE == 1) ? cnt + (long long)b * Nb : nullptr;
        const float th = (MODE == 1) ? thr[b] : 0.f;
        float best = (MODE == 0) ? INFINITY : -INFINITY;
        int bj = 0x7fffffff;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            mbar_wait(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t sv[16];
            tmem_ld16(tb + la + C_D0 + 32 * k + 16 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty[k]);
            const int j0 = t * BN + 16 * h;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int jj = j0 + u;
                if (jj < Nb) {
                    const float dist = 2.0f - 2.0f * __uint_as_float(sv[u]);
                    float v;
                    bool better;
                    if (MODE == 0) { v = dist; better = v < best; }
                    else { v = (dist < th) ? __ldg(cb + jj) : 0.f; better = v > best; }
                    best = better ? v : best;      // jj increases: the first occurrence of the extreme is kept
                    bj = better ? jj : bj;
                }
            }
        }
        // merge the two column halves of every row (ties to the lower index)
        if (h == 1) { part_v[row] = best; part_j[row] = bj; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (h == 0 && ok) {
            const float ov = part_v[row];
            const int oj = part_j[row];
            const bool take = (MODE == 0) ? (ov < best || (ov == best && oj < bj)) : (ov > best || (ov == best && oj < bj));
            out[(long long)b * Ma + i0 + row] = take ? oj : bj;
        }
    } else if (warp < MMA_WARP) {
        const int lw = warp - LOAD_WARP0;
        const int j = 8 * lw + (lane & 7), cq = lane >> 3;       // coalesced 8-row x 4-float4 mapping, see meanshift_tc.cu
        float4 vin[8], vnx[8];
        auto load_tile = [&](int t, float4 (&v)[8]) {
            const int r = t * BN + j;
            const bool ok = (t < ntiles) && (r < Nb);
            const float* p = Bb + (long long)(ok ? r : 0) * D;
#pragma unroll
            for (int it = 0; it < 8; ++it)
                v[it] = ok ? *reinterpret_cast<const float4*>(p + 4 * (4 * it + cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        load_tile(0, vin);
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % NSTAGE;
            load_tile(t + 1, vnx);
            mbar_wait(&bars.x_empty[s], ((t / NSTAGE) & 1) ^ 1);
            unsigned char* xa_b = smem + s * STAGE_BYTES;
            unsigned char* xa_s = xa_b + XA_BYTES;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int c4 = 4 * it + cq;
                const float f0 = vin[it].x, f1 = vin[it].y, f2 = vin[it].z, f3 = vin[it].w;
                const float b0 = tf32_hi(f0), b1 = tf32_hi(f1), b2 = tf32_hi(f2), b3 = tf32_hi(f3);
                const uint32_t oa = (uint32_t)(c4 * XA_LBO + (j >> 3) * 128 + (j & 7) * 16);
                *reinterpret_cast<float4*>(xa_b + oa) = make_float4(b0, b1, b2, b3);
                *reinterpret_cast<float4*>(xa_s + oa) = make_float4(f0 - b0, f1 - b1, f2 - b2, f3 - b3);
            }
            fence_async_smem();
            mbar_arrive(&bars.x_full[s]);
#pragma unroll
            for (int it = 0; it < 8; ++it) vin[it] = vnx[it];
        }
    } else {
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc(2, BM, BN, 0, 0);
        const uint32_t sbase = smem_u32(smem);
        mbar_wait(&bars.a_ready, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % NSTAGE, k = t & 1;
            mbar_wait(&bars.x_full[s], (t / NSTAGE) & 1);
            mbar_wait(&bars.s_empty[k], ((t >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t st = sbase + s * STAGE_BYTES;
            const uint64_t db0 = make_smem_desc(st, XA_LBO, SBO, 0);
            const uint64_t ds0 = make_smem_desc(st + XA_BYTES, XA_LBO, SBO, 0);
            const uint32_t d_s = tb + C_D0 + 32 * k;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < D / 8; ++ks) {
                    const uint64_t db = db0 + (uint64_t)(ks * ((2 * XA_LBO) >> 4));
                    const uint64_t ds = ds0 + (uint64_t)(ks * ((2 * XA_LBO) >> 4));
                    mma_tf32_ts(d_s, tb + C_AS + ks * 8, db, idesc_s, ks > 0 ? 1u : 0u);
                    mma_tf32_ts(d_s, tb + C_AB + ks * 8, ds, idesc_s, 1);
                    mma_tf32_ts(d_s, tb + C_AB + ks * 8, db, idesc_s, 1);
                }
                mma_commit(&bars.s_full[k]);
                mma_commit(&bars.x_empty[s]);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tb, TMEM_COLS);
}

}  // namespace mstca
}  // namespace pn

using namespace pn;

extern "C" int pn_ms_argsel_tc(int mode, const float* A, long long a_stride, int Ma, const float* Bm, long long b_stride,
                               int Nb, int B, int d, const float* cnt, const float* thr, int* out, void* stream) {
    PN_REQUIRE(A && Bm && out, "pn_ms_argsel_tc: null pointer");
    PN_REQUIRE(d == mstca::D, "pn_ms_argsel_tc: embedding width must be %d (got %d)", mstca::D, d);
    PN_REQUIRE((mode == 0) || (mode == 1 && cnt && thr), "pn_ms_argsel_tc: modes 0 and 1 only (mode 1 needs cnt, thr)");
    PN_REQUIRE(Ma > 0 && Nb > 0 && B > 0, "pn_ms_argsel_tc: empty input");
    size_t sm = (size_t)mstca::NSTAGE * mstca::STAGE_BYTES + 1024;
    dim3 grid(cdiv(Ma, mstca::BM), B);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        PN_CUDA(cudaFuncSetAttribute(mstca::ms_argsel_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        mstca::ms_argsel_tc_kernel<0><<<grid, mstca::NT, sm, st>>>(A, a_stride, Ma, Bm, b_stride, Nb, cnt, thr, out);
    } else {
        PN_CUDA(cudaFuncSetAttribute(mstca::ms_argsel_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        mstca::ms_argsel_tc_kernel<1><<<grid, mstca::NT, sm, st>>>(A, a_stride, Ma, Bm, b_stride, Nb, cnt, thr, out);
    }
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_argsel_tc_kernel");
    return PN_OK;
}
