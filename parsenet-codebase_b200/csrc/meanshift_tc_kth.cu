// K-th smallest pairwise distance per row (mean-shift bandwidth) on tcgen05: the distance tiles S = X X^T come from
// the split-TF32 tensor-core pipeline of meanshift_tc.cu (first product only), the selection is the exact 4 x 8-bit
// radix select of meanshift.cu::ms_kth_kernel, with per-row histograms in shared memory (row pitch 257 words: lanes of
// a warp own different rows, so equal digits land in different banks).
// Replaces MeanShift.compute_bandwidth (reference src/mean_shift.py:130-135): dist = 2 - 2 X X^T, topk(K, smallest)[-1].
#include "common.cuh"
#include "tc05.cuh"

namespace pn {
namespace mstck {
using namespace tc05;

constexpr int D = 128, BM = 128, BN = 32, NT = 416, NSTAGE = 2;
constexpr int EPI_WARPS = 8, LOAD_WARP0 = 8, MMA_WARP = 12, EPI_THREADS = 256;
constexpr uint32_t C_AB = 0, C_AS = 128, C_D0 = 256, TMEM_COLS = 512;
constexpr int XA_BYTES = BN * D * 4, STAGE_BYTES = 2 * XA_BYTES;     // big + small, 32 KB
constexpr uint32_t XA_LBO = BN * 16, SBO = 128;
constexpr int HP = 257;                                              // histogram row pitch (words)
constexpr int CAND = 8;                                              // distinct last-pass bins remembered per row

struct Bars { uint64_t x_full[NSTAGE], x_empty[NSTAGE], s_full[2], s_empty[2], a_ready; };

// grid (ceil(S/128), B).  rows (optional): [B][S] indices into the N points of each shape.
// flags (optional): [B][S]; a CTA none of whose 128 rows is flagged exits at once (the bracketed path of meanshift_tma.cu
// uses this kernel as the exact fall-back for the rows whose bracket failed).
__global__ void __launch_bounds__(NT, 1)
ms_kth_tc_kernel(const float* __restrict__ X, const int* __restrict__ rows, int S, long long shape_stride, int K,
                 const int* __restrict__ flags, float* __restrict__ kth) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned* hist = reinterpret_cast<unsigned*>(smem + NSTAGE * STAGE_BYTES);      // [BM][HP]
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned prefix[BM];
    __shared__ int krem[BM];
    // last pass: the (few) columns whose key shares the top 24 bits with the K-th one, packed (bin << 24) | column
    // (rows with more than CAND of them, i.e. many identical points, keep the tensor-core value).  The K-th element's
    // column lets the epilogue recompute that ONE distance with an fp32 FMA chain: the tensor core accumulates with
    // truncation, which biases every S by the same ~1e-6 (harmless for the ranking, visible in the bandwidth mean).
    __shared__ unsigned cand[BM][CAND];
    __shared__ int ncand[BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * BM;
    if (flags) {
        const int f = (tid < BM && i0 + tid < S) ? flags[(long long)b * S + i0 + tid] : 0;
        if (!__syncthreads_or(f)) return;
    }
    const float* Xb = X + (long long)b * shape_stride;
    const int* rb = rows ? rows + (long long)b * S : nullptr;
    const int ntiles = (S + BN - 1) / BN;
    const int total = 4 * ntiles;

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars.x_full[s], 128); mbar_init(&bars.x_empty[s], 1); }
        for (int k = 0; k < 2; ++k) { mbar_init(&bars.s_full[k], 1); mbar_init(&bars.s_empty[k], EPI_THREADS); }
        mbar_init(&bars.a_ready, EPI_THREADS);
        mbar_fence_init();
    }
    for (int e = tid; e < BM * HP; e += NT) hist[e] = 0u;
    if (tid < BM) { prefix[tid] = 0u; krem[tid] = K; ncand[tid] = 0; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    auto row_ptr = [&](int r) -> const float* {          // r < S assumed
        return Xb + (long long)(rb ? rb[r] : r) * D;
    };

    if (warp < EPI_WARPS) {
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t la = (uint32_t)(q * 32) << 16;
        const bool ok = (i0 + row) < S;
        const float* xr = ok ? row_ptr(i0 + row) + 64 * h : nullptr;
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
        unsigned* hrow = hist + row * HP;
#pragma unroll 1
        for (int tt = 0; tt < total; ++tt) {
            const int pass = tt / ntiles, t = tt - pass * ntiles;
            const int k = tt & 1;
            const int shift = 24 - 8 * pass;
            const unsigned pf = prefix[row];
            mbar_wait(&bars.s_full[k], (tt >> 1) & 1);
            tc_fence_after();
            uint32_t sv[16];
            tmem_ld16(tb + la + C_D0 + 32 * k + 16 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty[k]);
            const int j0 = t * BN + 16 * h;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (j0 + u < S) {
                    const float dist = 2.0f - 2.0f * __uint_as_float(sv[u]);
                    const unsigned key = f2ord(dist);
                    const bool match = (pass == 0) || ((key >> (shift + 8)) == (pf >> (shift + 8)));
                    if (match) {
                        const unsigned bin = (key >> shift) & 255u;
                        atomicAdd(&hrow[bin], 1u);              // result unused: a fire-and-forget shared-memory RED
                        if (pass == 3) {                        // rare: shares the top 24 key bits with the K-th element
                            const int slot = atomicAdd(&ncand[row], 1);
                            if (slot < CAND) cand[row][slot] = (bin << 24) | (unsigned)(j0 + u);
                        }
                    }
                }
            }
            if (t == ntiles - 1) {
                // end of a pass: pick the bin holding the krem-th smallest, extend the prefix, clear the row histogram
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (h == 0) {
                    int kk = krem[row];
                    unsigned run = 0, digit = 255u;
                    bool found = false;
                    for (int bin = 0; bin < 256; ++bin) {
                        unsigned cnt = hrow[bin];
                        hrow[bin] = 0u;
                        if (!found && (int)(run + cnt) >= kk) { digit = (unsigned)bin; kk -= (int)run; found = true; }
                        run += cnt;
                    }
                    prefix[row] = pf | (digit << shift);
                    krem[row] = kk;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        }
        if (h == 0 && ok) {
            float val = ord2f(prefix[row]);
            const unsigned digit = prefix[row] & 255u;
            const int nc = min(ncand[row], CAND);
            int jsel = -1;
            for (int c = 0; c < nc; ++c)
                if ((cand[row][c] >> 24) == digit) jsel = (int)(cand[row][c] & 0xffffffu);
            if (jsel >= 0) {
                // exact fp32 value of the selected pair (4 interleaved FMA chains)
                const float4* xi = reinterpret_cast<const float4*>(row_ptr(i0 + row));
                const float4* xj = reinterpret_cast<const float4*>(row_ptr(jsel));
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
                for (int c = 0; c < D / 4; ++c) {
                    const float4 u4 = xi[c], v4 = xj[c];
                    a0 = fmaf(u4.x, v4.x, a0); a1 = fmaf(u4.y, v4.y, a1);
                    a2 = fmaf(u4.z, v4.z, a2); a3 = fmaf(u4.w, v4.w, a3);
                }
                val = 2.0f - 2.0f * ((a0 + a1) + (a2 + a3));
            }
            kth[(long long)b * S + i0 + row] = val;
        }
    } else if (warp < MMA_WARP) {
        const int lw = warp - LOAD_WARP0;
        const int j = 8 * lw + (lane & 7), cq = lane >> 3;       // coalesced 8-row x 4-float4 mapping, see meanshift_tc.cu
        float4 vin[8], vnx[8];
        auto load_tile = [&](int tt, float4 (&v)[8]) {
            const int t = tt % ntiles;
            const int r = t * BN + j;
            const bool ok = (tt < total) && (r < S);
            const float* p = ok ? row_ptr(r) : nullptr;
#pragma unroll
            for (int it = 0; it < 8; ++it)
                v[it] = ok ? *reinterpret_cast<const float4*>(p + 4 * (4 * it + cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        load_tile(0, vin);
#pragma unroll 1
        for (int tt = 0; tt < total; ++tt) {
            const int s = tt % NSTAGE;
            load_tile(tt + 1, vnx);
            mbar_wait(&bars.x_empty[s], ((tt / NSTAGE) & 1) ^ 1);
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
        for (int tt = 0; tt < total; ++tt) {
            const int s = tt % NSTAGE, k = tt & 1;
            mbar_wait(&bars.x_full[s], (tt / NSTAGE) & 1);
            mbar_wait(&bars.s_empty[k], ((tt >> 1) & 1) ^ 1);
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

}  // namespace mstck
}  // namespace pn

using namespace pn;

static int kth_launch(const char* who, const float* X, const int* rows, int B, int S, long long shape_stride, int d, int K,
                      const int* flags, float* kth, void* stream) {
    PN_REQUIRE(X && kth, "%s: null pointer", who);
    PN_REQUIRE(d == mstck::D, "%s: embedding width must be %d (got %d)", who, mstck::D, d);
    PN_REQUIRE(K >= 1 && K <= S && S < (1 << 24), "%s: need 1 <= K <= S < 2^24 (K=%d S=%d)", who, K, S);
    size_t sm = mstck::NSTAGE * mstck::STAGE_BYTES + (size_t)mstck::BM * mstck::HP * sizeof(unsigned) + 1024;
    PN_CUDA(cudaFuncSetAttribute(mstck::ms_kth_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid(cdiv(S, mstck::BM), B);
    mstck::ms_kth_tc_kernel<<<grid, mstck::NT, sm, (cudaStream_t)stream>>>(X, rows, S, shape_stride, K, flags, kth);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_kth_tc_kernel");
    return PN_OK;
}

extern "C" int pn_ms_kth_dist_tc(const float* X, const int* rows, int B, int S, long long shape_stride, int d, int K,
                                 float* kth, void* stream) {
    return kth_launch("pn_ms_kth_dist_tc", X, rows, B, S, shape_stride, d, K, nullptr, kth, stream);
}

// the same selection restricted to the 128-row blocks that contain a row with flags[b][row] != 0 (other blocks return at
// once and their kth entries are left untouched): exact fall-back of pn_ms_kth_dist_tma
extern "C" int pn_ms_kth_dist_tc_flagged(const float* X, const int* rows, int B, int S, long long shape_stride, int d, int K,
                                         const int* flags, float* kth, void* stream) {
    PN_REQUIRE(flags, "pn_ms_kth_dist_tc_flagged: null flags");
    return kth_launch("pn_ms_kth_dist_tc_flagged", X, rows, B, S, shape_stride, d, K, flags, kth, stream);
}
