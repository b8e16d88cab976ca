// Mean-shift iteration on the 5th-gen tensor cores (tcgen05 + TMEM), split-TF32 (3 MMAs per product) for fp32-level
// accuracy.  Same contract as ms_fwd_kernel in meanshift.cu (reference src/mean_shift.py:58-77).
//
// One CTA = 128 query rows of Y, 13 warps, warp-specialised and mbarrier-pipelined:
//   warps 0-7  epilogue : thread = (row, half of the columns).  S tile TMEM -> regs, P = exp(clamp((S-1)/b^2)), row sum, tf32 split,
//                         P -> TMEM; every FLUSH tiles the O accumulator is drained into fp32 registers (the tensor
//                         core accumulates with truncation; short chains keep the bias at the 1e-6 level)
//   warps 8-11 loaders  : X tile (32 rows) LDG -> tf32 split -> two K-major no-swizzle core-matrix layouts in smem
//                         XA[j][d] (B operand of S = Y.X^T) and the transposed XB[d][j] (B operand of O += P.X),
//                         2 stages x 64 KB
//   warp 12    MMA      : one thread issues tcgen05.mma (kind::tf32, M=128): 48 x (N=32,K=8) per tile for S, 12 x
//                         (N=128,K=8) for O; every A operand comes from TENSOR MEMORY (Y split once, P per tile)
//   TMEM cols: Y_big [0,128) Y_small [128,256) S0 [256,288) S1 [288,320) P_big [320,352) P_small [352,384) O [384,512)
// Issue order G1(0) G1(1) G2(0) G1(2) G2(1) ... so the exp of tile t overlaps the S-MMA of tile t+1.
#include "common.cuh"
#include "tc05.cuh"
#include <stdlib.h>

// Loader lane mapping (all tcgen05 mean-shift kernels): at step `it` a loader warp covers an 8-row x 4-float4 block of the
// 32 x 128 tile: row j = 8*lw + (lane & 7), float4 column c4 = 4*it + (lane >> 3).  A warp-wide LDG then touches 8
// 128-byte lines (64 contiguous bytes per row) instead of 32 (lane = row, one 16-byte piece of 32 different rows), the
// XA store stays at its 4-wavefront minimum and the transposed XB store drops from 8 to 4 wavefronts.  ncu before the
// change: l1tex LSU data pipe 79-81 % busy (the limiter), 1024 of ~1900 LSU wavefronts per tile were those global loads.

namespace pn {
namespace mstc {
using namespace tc05;

constexpr int D = 128;
constexpr int BM = 128;       // query rows per CTA
constexpr int BN = 32;        // X rows per tile
constexpr int NT = 416;       // 8 epilogue + 4 loader + 1 MMA warp
constexpr int EPI_WARPS = 8, LOAD_WARP0 = 8, MMA_WARP = 12, EPI_THREADS = 256;
constexpr int NSTAGE = 3;
constexpr int FLUSH = 16;     // tiles per O accumulation chain
constexpr float CLAMP = 75.f;
constexpr uint32_t C_YB = 0, C_YS = 128, C_S0 = 256, C_PB = 320, C_PS = 352, C_O = 384, TMEM_COLS = 512;
constexpr int XA_BYTES = BN * D * 4;     // layout [d/4][j/8][8][16B]   LBO = BN*16 = 512, SBO = 128
constexpr int XB_BYTES = D * BN * 4;     // layout [j/4][d/8][8][16B]   LBO = D*16 = 2048, SBO = 128
constexpr int STAGE_BYTES = 2 * XA_BYTES + 2 * XB_BYTES;   // 64 KB
constexpr uint32_t XA_LBO = BN * 16, XB_LBO = D * 16, SBO = 128;

constexpr uint32_t C_PS2 = 320;   // PDB: P_small buffers [320,352), [352,384); P_big of tile t aliases S buffer t&1 (C_S0 + 32k)

struct Bars {
    uint64_t x_full[NSTAGE], x_empty[NSTAGE], s_full[2], s_empty[2], p_full, p_empty, o_flush, o_done;
    uint64_t p_full2[2], a_ready;    // PDB: per-buffer "P ready", separate "Y rows in TMEM"
};

// grid (ceil(N/BM), B).  PDB = double-buffered P (P_big written in place over the S columns the thread has just loaded,
// see meanshift_tc_bwd.cu): the epilogue of tile t no longer waits for the second product of tile t-1.
template <bool PDB>
__global__ void __launch_bounds__(NT, 1) ms_fwd_tc_kernel(const float* __restrict__ Y, const float* __restrict__ X, int N,
                                                          const float* __restrict__ cinv, float* __restrict__ Ynew,
                                                          float* __restrict__ den_out, float* __restrict__ unorm_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ float part[4][BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * BM;
    const float* Yb = Y + (long long)b * N * D;
    const float* Xb = X + (long long)b * N * D;
    const int ntiles = (N + BN - 1) / BN;

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars.x_full[s], 128); mbar_init(&bars.x_empty[s], 1); }
        for (int k = 0; k < 2; ++k) { mbar_init(&bars.s_full[k], 1); mbar_init(&bars.s_empty[k], EPI_THREADS); }
        mbar_init(&bars.p_full, EPI_THREADS); mbar_init(&bars.p_empty, 1);
        mbar_init(&bars.o_flush, EPI_THREADS); mbar_init(&bars.o_done, 1);
        mbar_init(&bars.p_full2[0], EPI_THREADS); mbar_init(&bars.p_full2[1], EPI_THREADS);
        mbar_init(&bars.a_ready, EPI_THREADS);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;

    if (warp < EPI_WARPS) {
        // =============================================================================== epilogue warps
        // thread = (row, column half h): TMEM lane quarter q = warp & 3, h = warp >> 2
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t la = (uint32_t)(q * 32) << 16;
        const bool ok = (i0 + row) < N;
        const float c2 = cinv[b] * 1.4426950408889634f;          // exp(e) = 2^(e * log2 e); clamp moves to the log2 domain
        const float CL2 = CLAMP * 1.4426950408889634f;
        const float* yr = Yb + (long long)(i0 + row) * D + 64 * h;
        // Y row half -> TMEM (tf32 big / small)
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t vb[16], vs[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                float4 v = ok ? *reinterpret_cast<const float4*>(yr + c0 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float big = tf32_hi(f[u]);
                    vb[e + u] = __float_as_uint(big);
                    vs[e + u] = __float_as_uint(f[u] - big);
                }
            }
            tmem_st16(tb + la + C_YB + 64 * h + c0, vb);
            tmem_st16(tb + la + C_YS + 64 * h + c0, vs);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(PDB ? &bars.a_ready : &bars.p_full);   // (non-PDB: phase 0 of p_full doubles as "Y is in TMEM")
        float oacc[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) oacc[e] = 0.f;
        float den = 0.f;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            const int j0 = t * BN + 16 * h;
            mbar_wait(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t sv[16], pb[16], ps[16];
            tmem_ld16(tb + la + C_S0 + 32 * k + 16 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty[k]);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                float e = (__uint_as_float(sv[u]) - 1.0f) * c2;
                e = fminf(fmaxf(e, -CL2), CL2);
                float p = (j0 + u < N) ? ex2_approx(e) : 0.f;
                den += p;
                float big = tf32_hi(p);
                pb[u] = __float_as_uint(big);
                ps[u] = __float_as_uint(p - big);
            }
            // non-PDB: P buffer free?  (G2 of tile t-1 finished)   p_empty completes once per tile; first wait passes (fresh)
            // PDB: s_full(t) already implies G2(t-2) done; G2(t-1) only has to be complete before O is drained
            const bool flush_now = (t > 0 && (t % FLUSH) == 0);
            if (!PDB || flush_now) {
                mbar_wait(&bars.p_empty, (t & 1) ^ 1);
                tc_fence_after();
            }
            if (flush_now) {
                // drain this thread's half of the O accumulator (sum over the previous FLUSH tiles) into registers
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t ov[16];
                    tmem_ld16(tb + la + C_O + 64 * h + c0, ov);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) oacc[c0 + e] += __uint_as_float(ov[e]);
                }
                tc_fence_before();
                mbar_arrive(&bars.o_flush);
            }
            tmem_st16(tb + la + (PDB ? C_S0 + 32 * k : C_PB) + 16 * h, pb);
            tmem_st16(tb + la + (PDB ? C_PS2 + 32 * k : C_PS) + 16 * h, ps);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(PDB ? &bars.p_full2[k] : &bars.p_full);      // (non-PDB: phase t+1)
        }
        // ---- final: last O chain, then u = y + (O/den - y), Y' = u/|u|
        mbar_wait(&bars.o_done, 0);
        tc_fence_after();
        part[h][row] = den;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float dn = part[0][row] + part[1][row];
        const float dinv = 1.0f / dn;
        float n2 = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t ov[16];
            tmem_ld16(tb + la + C_O + 64 * h + c0, ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                float o = oacc[c0 + e] + __uint_as_float(ov[e]);
                float y = ok ? yr[c0 + e] : 0.f;
                float m = o * dinv - y;
                float uu = y + m;
                oacc[c0 + e] = uu;
                n2 = fmaf(uu, uu, n2);
            }
        }
        part[2 + h][row] = n2;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float nr = sqrtf(part[2][row] + part[3][row]);
        if (ok) {
            float* dst = Ynew + ((long long)b * N + i0 + row) * D + 64 * h;
#pragma unroll
            for (int e = 0; e < 64; e += 4)
                *reinterpret_cast<float4*>(dst + e) =
                    make_float4(oacc[e] / nr, oacc[e + 1] / nr, oacc[e + 2] / nr, oacc[e + 3] / nr);
            if (h == 0) {
                den_out[(long long)b * N + i0 + row] = dn;
                unorm_out[(long long)b * N + i0 + row] = nr;
            }
        }
        tc_fence_before();
    } else if (warp < MMA_WARP) {
        // =============================================================================== loader warps
        const int lw = warp - LOAD_WARP0;        // 0..3
        const int j = 8 * lw + (lane & 7);       // row inside the tile (8-row block per warp)
        const int cq = lane >> 3;                // float4 column inside the 4-column block of step `it`
        const int l4 = lane & 3;                 // == j & 3
        const int jg = j >> 2;                   // 4-row group for the transposed copy
        // software prefetch: the global loads of tile t+1 are issued before tile t is processed / before the stage
        // of tile t+1 is known to be free, so only the split + shared-memory stores sit on the stage-release path
        float4 vin[8], vnx[8];
        {
            const bool ok0 = j < N;
#pragma unroll
            for (int it = 0; it < 8; ++it)
                vin[it] = ok0 ? *reinterpret_cast<const float4*>(Xb + (long long)j * D + 4 * (4 * it + cq))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % NSTAGE;
            {
                const int jn = (t + 1) * BN + j;
                const bool okn = (t + 1 < ntiles) && (jn < N);
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    vnx[it] = okn ? *reinterpret_cast<const float4*>(Xb + (long long)jn * D + 4 * (4 * it + cq))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(&bars.x_empty[s], ((t / NSTAGE) & 1) ^ 1);
            unsigned char* st = smem + s * STAGE_BYTES;
            unsigned char* xa_b = st;
            unsigned char* xa_s = st + XA_BYTES;
            unsigned char* xb_b = st + 2 * XA_BYTES;
            unsigned char* xb_s = st + 2 * XA_BYTES + XB_BYTES;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int c4 = 4 * it + cq;      // float4 column 0..31
                float f0 = vin[it].x, f1 = vin[it].y, f2 = vin[it].z, f3 = vin[it].w;
                // XA: chunk (j, c4) as is
                {
                    const float b0 = tf32_hi(f0), b1 = tf32_hi(f1), b2 = tf32_hi(f2), b3 = tf32_hi(f3);
                    const uint32_t oa = (uint32_t)(c4 * XA_LBO + (j >> 3) * 128 + (j & 7) * 16);
                    *reinterpret_cast<float4*>(xa_b + oa) = make_float4(b0, b1, b2, b3);
                    *reinterpret_cast<float4*>(xa_s + oa) = make_float4(f0 - b0, f1 - b1, f2 - b2, f3 - b3);
                }
                // branch-free 4x4 transpose inside each group of 4 lanes (two butterfly steps):
                // afterwards lane (jg, l4) holds X[4jg + 0..3][4c4 + l4]
                {
                    const bool hi = (l4 & 2) != 0;
                    float sa = hi ? f0 : f2, sb = hi ? f1 : f3;
                    sa = __shfl_xor_sync(0xffffffffu, sa, 2);
                    sb = __shfl_xor_sync(0xffffffffu, sb, 2);
                    f0 = hi ? sa : f0; f1 = hi ? sb : f1; f2 = hi ? f2 : sa; f3 = hi ? f3 : sb;
                    const bool od = (l4 & 1) != 0;
                    float sc = od ? f0 : f1, sd = od ? f2 : f3;
                    sc = __shfl_xor_sync(0xffffffffu, sc, 1);
                    sd = __shfl_xor_sync(0xffffffffu, sd, 1);
                    f0 = od ? sc : f0; f1 = od ? f1 : sc; f2 = od ? sd : f2; f3 = od ? f3 : sd;
                }
                {
                    const float b0 = tf32_hi(f0), b1 = tf32_hi(f1), b2 = tf32_hi(f2), b3 = tf32_hi(f3);
                    const int d = 4 * c4 + l4;
                    const uint32_t ob = (uint32_t)(jg * XB_LBO + (d >> 3) * 128 + (d & 7) * 16);
                    *reinterpret_cast<float4*>(xb_b + ob) = make_float4(b0, b1, b2, b3);
                    *reinterpret_cast<float4*>(xb_s + ob) = make_float4(f0 - b0, f1 - b1, f2 - b2, f3 - b3);
                }
            }
            fence_async_smem();
            mbar_arrive(&bars.x_full[s]);
#pragma unroll
            for (int it = 0; it < 8; ++it) vin[it] = vnx[it];
        }
    } else {
        // =============================================================================== MMA warp
        // The whole warp runs this warp-uniform code (descriptors stay in uniform registers); one elected lane issues.
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc(2, BM, BN, 0, 0);
        const uint32_t idesc_o = make_idesc(2, BM, D, 0, 0);
        const uint32_t sbase = smem_u32(smem);
        auto gemm2 = [&](int u) {
            if (PDB) mbar_wait(&bars.p_full2[u & 1], (u >> 1) & 1);
            else mbar_wait(&bars.p_full, (u + 1) & 1);         // phase u+1 (phase 0 was the Y-ready arrival)
            const uint32_t pb_col = PDB ? (C_S0 + 32 * (u & 1)) : C_PB, ps_col = PDB ? (C_PS2 + 32 * (u & 1)) : C_PS;
            const bool fresh = (u % FLUSH) == 0;
            if (u > 0 && fresh) mbar_wait(&bars.o_flush, ((u / FLUSH) - 1) & 1);
            tc_fence_after();
            const uint32_t st = sbase + (u % NSTAGE) * STAGE_BYTES;
            const uint64_t db0 = make_smem_desc(st + 2 * XA_BYTES, XB_LBO, SBO, 0);
            const uint64_t ds0 = make_smem_desc(st + 2 * XA_BYTES + XB_BYTES, XB_LBO, SBO, 0);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < BN / 8; ++ks) {
                    const uint64_t db = db0 + (uint64_t)(ks * ((2 * XB_LBO) >> 4));
                    const uint64_t ds = ds0 + (uint64_t)(ks * ((2 * XB_LBO) >> 4));
                    mma_tf32_ts(tb + C_O, tb + ps_col + ks * 8, db, idesc_o, (fresh && ks == 0) ? 0u : 1u);
                    mma_tf32_ts(tb + C_O, tb + pb_col + ks * 8, ds, idesc_o, 1);
                    mma_tf32_ts(tb + C_O, tb + pb_col + ks * 8, db, idesc_o, 1);
                }
                mma_commit(&bars.x_empty[u % NSTAGE]);
                mma_commit(&bars.p_empty);
            }
            __syncwarp();
        };
        mbar_wait(PDB ? &bars.a_ready : &bars.p_full, 0);       // Y rows are in TMEM
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
            const uint32_t d_s = tb + C_S0 + 32 * k;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < D / 8; ++ks) {
                    const uint64_t db = db0 + (uint64_t)(ks * ((2 * XA_LBO) >> 4));
                    const uint64_t ds = ds0 + (uint64_t)(ks * ((2 * XA_LBO) >> 4));
                    mma_tf32_ts(d_s, tb + C_YS + ks * 8, db, idesc_s, ks > 0 ? 1u : 0u);
                    mma_tf32_ts(d_s, tb + C_YB + ks * 8, ds, idesc_s, 1);
                    mma_tf32_ts(d_s, tb + C_YB + ks * 8, db, idesc_s, 1);
                }
                mma_commit(&bars.s_full[k]);
            }
            __syncwarp();
            if (t > 0) gemm2(t - 1);
        }
        gemm2(ntiles - 1);
        if (leader) mma_commit(&bars.o_done);
        __syncwarp();
    }
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tb, TMEM_COLS);
}

}  // namespace mstc
}  // namespace pn

using namespace pn;

extern "C" int pn_ms_iter_fwd_tc(const float* Y, const float* X, int B, int N, int d, const float* cinv, float* Ynew,
                                 float* den, float* unorm, void* stream) {
    PN_REQUIRE(Y && X && cinv && Ynew && den && unorm, "pn_ms_iter_fwd_tc: null pointer");
    PN_REQUIRE(d == mstc::D, "pn_ms_iter_fwd_tc: embedding width must be %d (got %d)", mstc::D, d);
    size_t sm = mstc::NSTAGE * mstc::STAGE_BYTES + 1024;
    static const bool pdb = [] { const char* e = getenv("PN_MS_FWD_PDB"); return !(e && e[0] == '0'); }();
    auto kern = pdb ? mstc::ms_fwd_tc_kernel<true> : mstc::ms_fwd_tc_kernel<false>;
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid(cdiv(N, mstc::BM), B);
    kern<<<grid, mstc::NT, sm, (cudaStream_t)stream>>>(Y, X, N, cinv, Ynew, den, unorm);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_fwd_tc_kernel");
    return PN_OK;
}
