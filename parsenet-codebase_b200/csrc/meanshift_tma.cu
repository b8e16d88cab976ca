// Mean-shift iteration kernels whose streamed operand tiles are fetched by TMA instead of being staged by loader warps.
// Default path since round 2 (first run on a B200: bit-identical to the loader-warp kernels of meanshift_tc*.cu, forward
// 3.88 -> 3.47 ms, dense backward 15.8 -> 13.5 ms per iteration at B = 16, N = 10^4; profiles/r02_first_call.md).
//
// Why (profiles/r01_tc_probe.md, r01_ncu_meanshift_tc.md, r01_ms_bwd_ablation_sweep.md): in ms_fwd_tc_kernel /
// ms_bwd_tc_kernel<rows> four loader warps LDG a 32 x 128 tile of X, split it into tf32 big / small parts and store it
// in two layouts (as is for S = Y.X^T, transposed for O += P.X): 16 KB of LDG + 64 KB of STS + shuffle transposes per
// tile, which ncu shows as the l1tex / LSU limiter (79-81 % busy, tensor pipe 50-57 %).  The probes established that
//   * the tensor core TRUNCATES fp32 operands to tf32, so raw X *is* the "big" part (no masked copy needed),
//   * a 128B-swizzled [rows][32 floats] slab as TMA writes it is read exactly by a K-major kind::tf32 descriptor
//     (layout type 2, SBO = 1024, K step = +32 B inside the atom),
//   * MN-major tf32 operands are not executed, i.e. the second product needs the tile in [d][j] order.
// X is constant over all iterations of one mean_shift call (reference src/mean_shift.py:58-77 shifts new_X against the
// fixed X), so ONE pass per call (ms_prep_operands_kernel) writes Xs = X - tf32_hi(X) and the transposes Xt, Xst, and
// every tile step fetches its four operand tiles with 10 TMA box loads on one mbarrier.  Loader warps, their
// global loads, shared-memory stores and shuffles are gone; the epilogue / MMA protocol is that of the PDB variant of
// meanshift_tc.cu (double-buffered S and P in TMEM).  Every mbarrier wait is the watchdog version (trap after ~2 s): a
// protocol or tensor-map mistake surfaces as a launch error, not as a hung GPU.
//
// Stage layout (64 KB, 1024-byte aligned):  XA_b | XA_s | XB_b | XB_s, 16 KB each
//   XA_* : 4 slabs (d in [32 s, 32 s + 32)) of [32 j rows][128 B], TMA box {32 d, 32 j, 1} of X / Xs       (K = d)
//   XB_* : 1 slab of [128 d rows][128 B = 32 j], TMA box {32 j, 128 d, 1} of Xt / Xst                      (K = j)
#include "common.cuh"
#include "tc05.cuh"
#include "kth_select.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace pn {
namespace mstma {
using namespace tc05;

constexpr int D = 128, BM = 128, BN = 32;
constexpr int EPI_WARPS = 8, TMA_WARP = 8, MMA_WARP = 9, NT = 320, EPI_THREADS = 256;
constexpr int NSTAGE = 3, MAX_STAGES = 6, FLUSH = 16;      // NSTAGE: 64 KB stages of one CTA; pairs use 6 x 32 KB
constexpr float CLAMP = 75.f;
constexpr uint32_t C_YB = 0, C_YS = 128, C_S0 = 256, C_PS2 = 320, C_O = 384, TMEM_COLS = 512;
constexpr int PART_BYTES = BN * D * 4;            // 16 KB per operand part
constexpr int STAGE_BYTES = 4 * PART_BYTES;       // 64 KB
constexpr uint32_t SW128 = 2, SBO128 = 1024;      // descriptor layout type / 8-row group stride of a 128B-swizzled slab

struct Bars {
    uint64_t x_full[MAX_STAGES], x_empty[MAX_STAGES], s_full[2], s_empty[2], p_full2[2], p_empty, o_flush, o_done, a_ready;
};

// ---------------------------------------------------------------------------------------------- operand preparation
// X [B][N][128] -> Xs [B][N][128] (small split part), Xt / Xst [B][128][Np] (transposes of X / Xs, zero beyond N).
// grid (Np / 32, B), 256 threads: one 32 x 128 tile, transposed through shared memory.
__global__ void __launch_bounds__(256) ms_prep_operands_kernel(const float* __restrict__ X, int N, int Np,
                                                               float* __restrict__ Xs, float* __restrict__ Xt,
                                                               float* __restrict__ Xst) {
    __shared__ float tb[32][D + 1], ts[32][D + 1];
    const int b = blockIdx.y, j0 = blockIdx.x * 32, tid = threadIdx.x;
    const float* Xb = X + (long long)b * N * D;
    float* Xsb = Xs + (long long)b * N * D;
    for (int e = tid; e < 32 * D; e += 256) {
        const int r = e >> 7, c = e & 127, j = j0 + r;
        const float v = (j < N) ? Xb[(long long)j * D + c] : 0.f;
        const float sm = v - tf32_hi(v);
        if (j < N) Xsb[(long long)j * D + c] = sm;
        tb[r][c] = v;
        ts[r][c] = sm;
    }
    __syncthreads();
    float* Xtb = Xt + (long long)b * D * Np;
    float* Xstb = Xst + (long long)b * D * Np;
    for (int e = tid; e < D * 32; e += 256) {
        const int dd = e >> 5, r = e & 31;
        Xtb[(long long)dd * Np + j0 + r] = tb[r][dd];
        Xstb[(long long)dd * Np + j0 + r] = ts[r][dd];
    }
}

// ---------------------------------------------------------------------------------------------- shared warp roles
// CG = 1: one CTA per 128 rows.  CG = 2 (tcgen05 cta_group::2, cluster of 2 CTAs along x): the pair shares every streamed
// tile -- each CTA stages HALF of both B operands (16 of the 32 tile rows for the first product, 64 of the 128 d rows for
// the second), the leader (cluster rank 0) issues M = 256 MMAs that write S / O into the TMEM of both CTAs, and each CTA
// runs its own epilogue on its own 128 rows.  Per CTA and tile: 32 KB instead of 64 KB of operand traffic into shared
// memory and half of the MMA operand reads -- the shared-memory co-limiter of the 1-CTA kernels.
// Barriers the MMA issuer waits on (s_empty, p_full2, o_flush, a_ready, x_full) live in the leader and collect the
// arrivals of both CTAs; barriers the epilogues / producers wait on (s_full, p_empty, o_done, x_empty) are local and are
// signalled in both CTAs by multicast commits.
template <int CG>
struct Geo {
    static constexpr int ROWS = BN / CG;              // tile rows staged by one CTA (first product)
    static constexpr int SLAB = ROWS * 128;           // one [ROWS][128 B] slab
    static constexpr int PART = PART_BYTES / CG;      // one operand part per CTA
    static constexpr int STAGE = 4 * PART;            // per-CTA stage
    static constexpr int DROWS = D / CG;              // d rows staged by one CTA (second product)
    static constexpr int NSTG = NSTAGE * CG;          // pipeline depth (same shared-memory footprint per CTA)
};

template <int CG>
__device__ __forceinline__ void arrive_to_issuer(uint64_t* bar) {
    if (CG == 1) mbar_arrive(bar);
    else mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
}

// TMA producer (one elected lane of its warp): per tile step 4 + 4 slab boxes of the row-major forms and one box each of
// the transposed forms, all completing on the leader's x_full[stage].
template <int CG>
__device__ __forceinline__ void tma_producer(unsigned char* smem, Bars& bars, const CUtensorMap* mA, const CUtensorMap* mAs,
                                             const CUtensorMap* mT, const CUtensorMap* mTs, int ntiles, int b, int rank) {
    using G = Geo<CG>;
    if (elect_one()) {
        tma_prefetch_desc(mA); tma_prefetch_desc(mAs); tma_prefetch_desc(mT); tma_prefetch_desc(mTs);
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % G::NSTG;
            mbar_wait_guarded(&bars.x_empty[s], ((t / G::NSTG) & 1) ^ 1);
            unsigned char* st = smem + s * G::STAGE;
            if (rank == 0) mbar_arrive_expect_tx(&bars.x_full[s], STAGE_BYTES);      // bytes of BOTH CTAs' halves
            const int row = t * BN + rank * G::ROWS;
            if (CG == 1) {
#pragma unroll
                for (int sl = 0; sl < 4; ++sl) {
                    tma_load_3d(st + sl * G::SLAB, mA, &bars.x_full[s], 32 * sl, row, b);
                    tma_load_3d(st + G::PART + sl * G::SLAB, mAs, &bars.x_full[s], 32 * sl, row, b);
                }
                tma_load_3d(st + 2 * G::PART, mT, &bars.x_full[s], t * BN, 0, b);
                tma_load_3d(st + 3 * G::PART, mTs, &bars.x_full[s], t * BN, 0, b);
            } else {
                const uint32_t full = mapa_u32(smem_u32(&bars.x_full[s]), 0);
#pragma unroll
                for (int sl = 0; sl < 4; ++sl) {
                    tma_load_3d_2sm(st + sl * G::SLAB, mA, full, 32 * sl, row, b);
                    tma_load_3d_2sm(st + G::PART + sl * G::SLAB, mAs, full, 32 * sl, row, b);
                }
                tma_load_3d_2sm(st + 2 * G::PART, mT, full, t * BN, rank * G::DROWS, b);
                tma_load_3d_2sm(st + 3 * G::PART, mTs, full, t * BN, rank * G::DROWS, b);
            }
        }
    }
    __syncwarp();
}

// MMA issue warp of the leader (warp-uniform code, one elected lane issues): per tile the first product D(t) = A . tile^T
// (48 MMAs, N = 32, K = d through the 4 slabs) and, one tile behind, the second product O += P(t-1) . tile (12 MMAs,
// N = 128, K = 32 tile rows), split-TF32 (small.big + big.small + big.big), A operands in TMEM.
// Issue order G1(0) G1(1) G2(0) G1(2) ...
template <int CG>
__device__ __forceinline__ void mma_issuer(unsigned char* smem, Bars& bars, uint32_t tb, int ntiles) {
    using G = Geo<CG>;
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc(2, 128 * CG, BN, 0, 0);
    const uint32_t idesc_o = make_idesc(2, 128 * CG, D, 0, 0);
    const uint32_t sbase = smem_u32(smem);
    auto mma = [&](uint32_t dcol, uint32_t acol, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
        if (CG == 1) mma_tf32_ts(dcol, acol, bdesc, idesc, acc);
        else mma_tf32_ts2(dcol, acol, bdesc, idesc, acc);
    };
    auto commit = [&](uint64_t* bar) {
        if (CG == 1) mma_commit(bar);
        else mma_commit2_mc(bar, (uint16_t)3);
    };
    auto gemm2 = [&](int u) {
        mbar_wait_guarded(&bars.p_full2[u & 1], (u >> 1) & 1);
        const uint32_t pb_col = C_S0 + 32 * (u & 1), ps_col = C_PS2 + 32 * (u & 1);
        const bool fresh = (u % FLUSH) == 0;
        if (u > 0 && fresh) mbar_wait_guarded(&bars.o_flush, ((u / FLUSH) - 1) & 1);
        tc_fence_after();
        const uint32_t st = sbase + (u % G::NSTG) * G::STAGE;
        const uint64_t db0 = make_smem_desc(st + 2 * G::PART, 16, SBO128, SW128);
        const uint64_t ds0 = make_smem_desc(st + 3 * G::PART, 16, SBO128, SW128);
        if (leader) {
#pragma unroll
            for (int ks = 0; ks < BN / 8; ++ks) {            // K = tile row: 4 steps of 8 inside the one 128 B atom row
                const uint64_t db = db0 + (uint64_t)((ks * 32) >> 4);
                const uint64_t ds = ds0 + (uint64_t)((ks * 32) >> 4);
                mma(tb + C_O, tb + ps_col + ks * 8, db, idesc_o, (fresh && ks == 0) ? 0u : 1u);
                mma(tb + C_O, tb + pb_col + ks * 8, ds, idesc_o, 1);
                mma(tb + C_O, tb + pb_col + ks * 8, db, idesc_o, 1);
            }
            commit(&bars.x_empty[u % G::NSTG]);
            commit(&bars.p_empty);
        }
        __syncwarp();
    };
    mbar_wait_guarded(&bars.a_ready, 0);
    tc_fence_after();
#pragma unroll 1
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % G::NSTG, k = t & 1;
        mbar_wait_guarded(&bars.x_full[s], (t / G::NSTG) & 1);
        mbar_wait_guarded(&bars.s_empty[k], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t st = sbase + s * G::STAGE;
        const uint64_t db0 = make_smem_desc(st, 16, SBO128, SW128);
        const uint64_t ds0 = make_smem_desc(st + G::PART, 16, SBO128, SW128);
        const uint32_t d_s = tb + C_S0 + 32 * k;
        if (leader) {
#pragma unroll
            for (int ks = 0; ks < D / 8; ++ks) {             // K = d: slab ks / 4, +32 B per step inside the slab
                const uint32_t off = (uint32_t)((ks >> 2) * G::SLAB + (ks & 3) * 32);
                const uint64_t db = db0 + (uint64_t)(off >> 4);
                const uint64_t ds = ds0 + (uint64_t)(off >> 4);
                mma(d_s, tb + C_YS + ks * 8, db, idesc_s, ks > 0 ? 1u : 0u);
                mma(d_s, tb + C_YB + ks * 8, ds, idesc_s, 1);
                mma(d_s, tb + C_YB + ks * 8, db, idesc_s, 1);
            }
            commit(&bars.s_full[k]);
        }
        __syncwarp();
        if (t > 0) gemm2(t - 1);
    }
    gemm2(ntiles - 1);
    if (leader) commit(&bars.o_done);
    __syncwarp();
}

template <int CG>
__device__ __forceinline__ void init_bars(Bars& bars) {
    for (int s = 0; s < Geo<CG>::NSTG; ++s) { mbar_init(&bars.x_full[s], 1); mbar_init(&bars.x_empty[s], 1); }
    for (int k = 0; k < 2; ++k) {
        mbar_init(&bars.s_full[k], 1);
        mbar_init(&bars.s_empty[k], CG * EPI_THREADS);
        mbar_init(&bars.p_full2[k], CG * EPI_THREADS);
    }
    mbar_init(&bars.p_empty, 1); mbar_init(&bars.o_flush, CG * EPI_THREADS); mbar_init(&bars.o_done, 1);
    mbar_init(&bars.a_ready, CG * EPI_THREADS);
    mbar_fence_init();
}

// kernel prologue / epilogue shared by the three kernels: TMEM allocation, barrier init, (cluster) sync
template <int CG>
__device__ __forceinline__ uint32_t prologue(Bars& bars, uint32_t* tmem_base_s, int warp, int tid) {
    if (warp == MMA_WARP) { if (CG == 1) tmem_alloc(tmem_base_s, TMEM_COLS); else tmem_alloc2(tmem_base_s, TMEM_COLS); }
    if (tid == 0) init_bars<CG>(bars);
    tc_fence_before();
    if (CG == 1) __syncthreads(); else cluster_sync_all();
    tc_fence_after();
    return *tmem_base_s;
}
template <int CG>
__device__ __forceinline__ void finale(uint32_t tb, int warp) {
    tc_fence_before();
    if (CG == 1) __syncthreads(); else cluster_sync_all();      // the leader's MMAs read the peer's shared memory until o_done
    if (warp == MMA_WARP) { if (CG == 1) tmem_dealloc(tb, TMEM_COLS); else tmem_dealloc2(tb, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------- forward iteration
// grid (ceil(N / 128), B), 320 threads: warps 0-7 epilogue, warp 8 TMA producer, warp 9 MMA issue.
template <int CG>
__global__ void __launch_bounds__(NT, 1)
ms_fwd_tma_kernel(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mXs,
                  const __grid_constant__ CUtensorMap mXt, const __grid_constant__ CUtensorMap mXst,
                  const float* __restrict__ Y, int N, const float* __restrict__ cinv, float* __restrict__ Ynew,
                  float* __restrict__ den_out, float* __restrict__ unorm_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ float part[4][BM];
    // the dynamic segment is only guaranteed 16-byte aligned by the ABI: round up to the 1024 B the swizzle needs
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * BM;
    const float* Yb = Y + (long long)b * N * D;
    const int ntiles = (N + BN - 1) / BN;

    const int rank = (CG == 1) ? 0 : (int)cluster_ctarank();
    const uint32_t tb = prologue<CG>(bars, &tmem_base_s, warp, tid);

    if (warp < EPI_WARPS) {
        // =============================================================================== epilogue warps
        // (identical to the PDB path of ms_fwd_tc_kernel: thread = (row, column half h); P_big(t) overwrites S(t) in place)
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t la = (uint32_t)(q * 32) << 16;
        const bool ok = (i0 + row) < N;
        const float c2 = cinv[b] * 1.4426950408889634f;
        const float CL2 = CLAMP * 1.4426950408889634f;
        const float* yr = Yb + (long long)(i0 + row) * D + 64 * h;
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
        arrive_to_issuer<CG>(&bars.a_ready);
        float oacc[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) oacc[e] = 0.f;
        float den = 0.f;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            const int j0 = t * BN + 16 * h;
            mbar_wait_guarded(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t sv[16], pb[16], ps[16];
            tmem_ld16(tb + la + C_S0 + 32 * k + 16 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            arrive_to_issuer<CG>(&bars.s_empty[k]);
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
            const bool flush_now = (t > 0 && (t % FLUSH) == 0);
            if (flush_now) {
                mbar_wait_guarded(&bars.p_empty, (t & 1) ^ 1);       // second product of tile t-1 complete
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t ov[16];
                    tmem_ld16(tb + la + C_O + 64 * h + c0, ov);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) oacc[c0 + e] += __uint_as_float(ov[e]);
                }
                tc_fence_before();
                arrive_to_issuer<CG>(&bars.o_flush);
            }
            tmem_st16(tb + la + C_S0 + 32 * k + 16 * h, pb);
            tmem_st16(tb + la + C_PS2 + 32 * k + 16 * h, ps);
            tmem_st_wait();
            tc_fence_before();
            arrive_to_issuer<CG>(&bars.p_full2[k]);
        }
        mbar_wait_guarded(&bars.o_done, 0);
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
    } else if (warp == TMA_WARP) {
        tma_producer<CG>(smem, bars, &mX, &mXs, &mXt, &mXst, ntiles, b, rank);
    } else if (rank == 0) {
        mma_issuer<CG>(smem, bars, tb, ntiles);
    }
    finale<CG>(tb, warp);
}

// ---------------------------------------------------------------------------------------------- backward, rows
// gY_i = sum_j gS_ij x_j  with gS = (G + gd_i) K c (0 where clamped), S = Y X^T, K = exp(clamp((S-1)c)), G = Gn X^T.
// Same math and epilogue as ms_bwd_tc_kernel<MODE_ROWS> with PDB (meanshift_tc_bwd.cu): the CTA owns 64 rows; TMEM lanes
// 0-63 hold Y_i, lanes 64-127 hold Gn_i ("virtual rows"), so one MMA chain against the streamed X tile yields S and G.
// The streamed tile is X (constant): its four operand forms come from TMA exactly as in the forward kernel.
// grid (ceil(N / 64), B), 320 threads; dynamic smem = stages + [2][64][32] floats for the G hand-over.
template <int CG>
__global__ void __launch_bounds__(NT, 1)
ms_bwd_rows_tma_kernel(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mXs,
                       const __grid_constant__ CUtensorMap mXt, const __grid_constant__ CUtensorMap mXst,
                       const float* __restrict__ Yp, const float* __restrict__ Gn, const float* __restrict__ gd, int N,
                       const float* __restrict__ cinv, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* exch = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const long long off = (long long)b * N * D;
    const float* Ypb = Yp + off;
    const float* Gb = Gn + off;
    const float* gdb = gd + (long long)b * N;
    const int r0 = blockIdx.x * 64;
    const int ntiles = (N + BN - 1) / BN;

    const int rank = (CG == 1) ? 0 : (int)cluster_ctarank();
    const uint32_t tb = prologue<CG>(bars, &tmem_base_s, warp, tid);

    if (warp < EPI_WARPS) {
        // =============================================================================== epilogue warps
        const int q = warp & 3, h = warp >> 2;
        const int vrow = q * 32 + lane;                         // TMEM lane: 0-63 real rows (Y), 64-127 virtual rows (Gn)
        const uint32_t la = (uint32_t)(q * 32) << 16;
        const float c = cinv[b];
        const float c2 = c * 1.4426950408889634f, CL2 = CLAMP * 1.4426950408889634f;
        const int r = r0 + (vrow & 63);
        const bool aok = r < N;
        const float* arow = ((vrow < 64) ? Ypb : Gb) + (long long)r * D + 64 * h;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t vb[16], vs[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                float4 v = aok ? *reinterpret_cast<const float4*>(arow + c0 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
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
        if (q >= 2) {                                           // P_small rows of the virtual half stay zero forever
            uint32_t z[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) z[u] = 0u;
            tmem_st16(tb + la + C_PS2 + 16 * h, z);
            tmem_st16(tb + la + C_PS2 + 32 + 16 * h, z);
        }
        tmem_st_wait();
        tc_fence_before();
        arrive_to_issuer<CG>(&bars.a_ready);
        const bool owner = q < 2;
        float oacc[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) oacc[e] = 0.f;
        const float gd_row = (owner && aok) ? gdb[r0 + vrow] : 0.f;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            mbar_wait_guarded(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t pb[16], ps[16], sv[16];
            tmem_ld16(tb + la + C_S0 + 32 * k + 16 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            arrive_to_issuer<CG>(&bars.s_empty[k]);
            float* ex = exch + (t & 1) * 64 * 32;
            if (q >= 2) {
#pragma unroll
                for (int u = 0; u < 16; ++u) ex[(vrow - 64) * 32 + ((16 * h + u + vrow) & 31)] = __uint_as_float(sv[u]);
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (q < 2) {
                const int j0 = t * 32 + 16 * h;
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    float g = ex[vrow * 32 + ((16 * h + u + vrow) & 31)];
                    float e = (__uint_as_float(sv[u]) - 1.0f) * c2;
                    const bool cl = (e > CL2) || (e < -CL2);
                    e = fminf(fmaxf(e, -CL2), CL2);
                    float kk = ex2_approx(e);
                    float p = (!cl && (j0 + u < N)) ? (g + gd_row) * kk * c : 0.f;
                    float big = tf32_hi(p);
                    pb[u] = __float_as_uint(big);
                    ps[u] = __float_as_uint(p - big);
                }
            }
            const bool flush_now = (t > 0 && (t % FLUSH) == 0);
            if (flush_now) {
                mbar_wait_guarded(&bars.p_empty, (t & 1) ^ 1);
                tc_fence_after();
                if (owner) {
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 16) {
                        uint32_t ov[16];
                        tmem_ld16(tb + la + C_O + 64 * h + c0, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) oacc[c0 + e] += __uint_as_float(ov[e]);
                    }
                }
                tc_fence_before();
                arrive_to_issuer<CG>(&bars.o_flush);
            }
            const uint32_t pb_col = C_S0 + 32 * k, ps_col = C_PS2 + 32 * k;
            if (q < 2) {
                tmem_st16(tb + la + pb_col + 16 * h, pb);
                tmem_st16(tb + la + ps_col + 16 * h, ps);
            } else {                                 // the G values of the virtual half must not act as P rows
                uint32_t z[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) z[u] = 0u;
                tmem_st16(tb + la + pb_col + 16 * h, z);
            }
            tmem_st_wait();
            tc_fence_before();
            arrive_to_issuer<CG>(&bars.p_full2[k]);
        }
        mbar_wait_guarded(&bars.o_done, 0);
        tc_fence_after();
        if (owner) {       // warp-uniform: tcgen05.ld is warp-collective, only the global stores are per-lane guarded
            float* dst = out + off + (long long)(r0 + vrow) * D + 64 * h;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t ov[16];
                tmem_ld16(tb + la + C_O + 64 * h + c0, ov);
                tmem_ld_wait();
                if (aok) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        *reinterpret_cast<float4*>(dst + c0 + e) =
                            make_float4(oacc[c0 + e] + __uint_as_float(ov[e]), oacc[c0 + e + 1] + __uint_as_float(ov[e + 1]),
                                        oacc[c0 + e + 2] + __uint_as_float(ov[e + 2]),
                                        oacc[c0 + e + 3] + __uint_as_float(ov[e + 3]));
                }
            }
        }
        tc_fence_before();
    } else if (warp == TMA_WARP) {
        tma_producer<CG>(smem, bars, &mX, &mXs, &mXt, &mXst, ntiles, b, rank);
    } else if (rank == 0) {
        mma_issuer<CG>(smem, bars, tb, ntiles);
    }
    finale<CG>(tb, warp);
}

// ---------------------------------------------------------------------------------------------- backward, cols
// gX_j (+)= sum_i gS_ij y_i + K_ij Gn_i: the CTA owns 128 rows j of X (A operand, TMEM); the streamed 32-row tile is the
// CONCATENATION of 16 rows of Y and the same 16 rows of Gn, so D[:, 0:16] = S^T, D[:, 16:32] = G^T, P = [gS^T | K^T] and the
// second product with the same tile gives gS^T Y + K^T Gn in one chain (ms_bwd_tc_kernel<MODE_COLS>, meanshift_tc_bwd.cu).
// [Y; Gn] changes every iteration: ms_prep_concat_kernel writes the interleaved matrix C (tile t = rows 32t..32t+31) and
// its three other operand forms once per backward iteration (50 MB per shape against 7 N^2 d flop), after which the
// streaming side is exactly that of the forward kernel.
// grid (Nq / 16, B), 256 threads; Nq = N rounded up to 16; C, Cs [B][2 Nq][128]; Ct, Cst [B][128][2 Nq]
__global__ void __launch_bounds__(256) ms_prep_concat_kernel(const float* __restrict__ Y, const float* __restrict__ Gn, int N,
                                                             int Nq, float* __restrict__ C, float* __restrict__ Cs,
                                                             float* __restrict__ Ct, float* __restrict__ Cst) {
    __shared__ float tb[32][D + 1], ts[32][D + 1];
    const int b = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
    const long long src_off = (long long)b * N * D;
    const long long rows2 = 2LL * Nq;
    float* Cb = C + (long long)b * rows2 * D;
    float* Csb = Cs + (long long)b * rows2 * D;
    for (int e = tid; e < 32 * D; e += 256) {
        const int r = e >> 7, c = e & 127;
        const int i = 16 * t + (r & 15);
        const float* src = (r < 16) ? Y : Gn;
        const float v = (i < N) ? src[src_off + (long long)i * D + c] : 0.f;
        const float sm = v - tf32_hi(v);
        Cb[(32LL * t + r) * D + c] = v;
        Csb[(32LL * t + r) * D + c] = sm;
        tb[r][c] = v;
        ts[r][c] = sm;
    }
    __syncthreads();
    float* Ctb = Ct + (long long)b * D * rows2;
    float* Cstb = Cst + (long long)b * D * rows2;
    for (int e = tid; e < D * 32; e += 256) {
        const int dd = e >> 5, r = e & 31;
        Ctb[(long long)dd * rows2 + 32LL * t + r] = tb[r][dd];
        Cstb[(long long)dd * rows2 + 32LL * t + r] = ts[r][dd];
    }
}

// grid (ceil(N / 128), B), 320 threads
template <int CG>
__global__ void __launch_bounds__(NT, 1)
ms_bwd_cols_tma_kernel(const __grid_constant__ CUtensorMap mC, const __grid_constant__ CUtensorMap mCs,
                       const __grid_constant__ CUtensorMap mCt, const __grid_constant__ CUtensorMap mCst,
                       const float* __restrict__ X, const float* __restrict__ gd, int N, const float* __restrict__ cinv,
                       float* __restrict__ out, int accumulate) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const long long off = (long long)b * N * D;
    const float* Xb = X + off;
    const float* gdb = gd + (long long)b * N;
    const int r0 = blockIdx.x * 128;
    const int ntiles = (N + 15) / 16;

    const int rank = (CG == 1) ? 0 : (int)cluster_ctarank();
    const uint32_t tb = prologue<CG>(bars, &tmem_base_s, warp, tid);

    if (warp < EPI_WARPS) {
        const int q = warp & 3, h = warp >> 2;
        const int vrow = q * 32 + lane;
        const uint32_t la = (uint32_t)(q * 32) << 16;
        const float c = cinv[b];
        const float c2 = c * 1.4426950408889634f, CL2 = CLAMP * 1.4426950408889634f;
        const bool aok = (r0 + vrow) < N;
        const float* arow = Xb + (long long)(r0 + vrow) * D + 64 * h;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t vb[16], vs[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                float4 v = aok ? *reinterpret_cast<const float4*>(arow + c0 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
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
        arrive_to_issuer<CG>(&bars.a_ready);
        float oacc[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) oacc[e] = 0.f;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            mbar_wait_guarded(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t s8[8], g8[8], pb[16], ps[16];
            tmem_ld8(tb + la + C_S0 + 32 * k + 8 * h, s8);              // S^T: tile rows 8h .. 8h+7 (Y half)
            tmem_ld8(tb + la + C_S0 + 32 * k + 16 + 8 * h, g8);         // G^T: the same rows of the Gn half
            tmem_ld_wait();
            tc_fence_before();
            arrive_to_issuer<CG>(&bars.s_empty[k]);
            const int i0t = t * 16 + 8 * h;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool iv = (i0t + u) < N;
                const float gdi = iv ? gdb[i0t + u] : 0.f;
                float e = (__uint_as_float(s8[u]) - 1.0f) * c2;
                const bool cl = (e > CL2) || (e < -CL2);
                e = fminf(fmaxf(e, -CL2), CL2);
                float kk = iv ? ex2_approx(e) : 0.f;
                float p1 = (!cl && iv) ? (__uint_as_float(g8[u]) + gdi) * kk * c : 0.f;
                float b1 = tf32_hi(p1), b2 = tf32_hi(kk);
                pb[u] = __float_as_uint(b1);       ps[u] = __float_as_uint(p1 - b1);       // gS^T -> cols 8h..
                pb[8 + u] = __float_as_uint(b2);   ps[8 + u] = __float_as_uint(kk - b2);   // K^T  -> cols 16+8h..
            }
            const bool flush_now = (t > 0 && (t % FLUSH) == 0);
            if (flush_now) {
                mbar_wait_guarded(&bars.p_empty, (t & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t ov[16];
                    tmem_ld16(tb + la + C_O + 64 * h + c0, ov);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) oacc[c0 + e] += __uint_as_float(ov[e]);
                }
                tc_fence_before();
                arrive_to_issuer<CG>(&bars.o_flush);
            }
            const uint32_t pb_col = C_S0 + 32 * k, ps_col = C_PS2 + 32 * k;
            {
                uint32_t a[8], dd[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { a[u] = pb[u]; dd[u] = pb[8 + u]; }
                tmem_st8(tb + la + pb_col + 8 * h, a);
                tmem_st8(tb + la + pb_col + 16 + 8 * h, dd);
#pragma unroll
                for (int u = 0; u < 8; ++u) { a[u] = ps[u]; dd[u] = ps[8 + u]; }
                tmem_st8(tb + la + ps_col + 8 * h, a);
                tmem_st8(tb + la + ps_col + 16 + 8 * h, dd);
                tmem_st_wait();
            }
            tc_fence_before();
            arrive_to_issuer<CG>(&bars.p_full2[k]);
        }
        mbar_wait_guarded(&bars.o_done, 0);
        tc_fence_after();
        float* dst = out + off + (long long)(r0 + vrow) * D + 64 * h;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t ov[16];
            tmem_ld16(tb + la + C_O + 64 * h + c0, ov);
            tmem_ld_wait();
            if (aok) {
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    float4 v = make_float4(oacc[c0 + e] + __uint_as_float(ov[e]), oacc[c0 + e + 1] + __uint_as_float(ov[e + 1]),
                                           oacc[c0 + e + 2] + __uint_as_float(ov[e + 2]),
                                           oacc[c0 + e + 3] + __uint_as_float(ov[e + 3]));
                    if (accumulate) {
                        float4 a = *reinterpret_cast<const float4*>(dst + c0 + e);
                        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                    }
                    *reinterpret_cast<float4*>(dst + c0 + e) = v;
                }
            }
        }
        tc_fence_before();
    } else if (warp == TMA_WARP) {
        tma_producer<CG>(smem, bars, &mC, &mCs, &mCt, &mCst, ntiles, b, rank);
    } else if (rank == 0) {
        mma_issuer<CG>(smem, bars, tb, ntiles);
    }
    finale<CG>(tb, warp);
}

// ---------------------------------------------------------------------------------------------- K-th distance, bracketed
// MeanShift.compute_bandwidth (reference src/mean_shift.py:115-137) needs, per row, the K-th smallest of dist = 2 - 2 X X^T
// (K = quantile * N, 150 of 10^4).  The radix kernel of meanshift_tc_kth.cu recomputes the N x N distance tiles four times;
// here they are computed once:
//   1. collect<ALL>  : distances of every row to a strided SAMPLE of m <= cap columns (a tensor map whose row pitch is
//                      stride * 128 floats), all m keys stored per row
//   2. select        : hi_i = b-th smallest of row i's sample (b = K m / N + 7 sigma + 2: an upper bracket of the K-th
//                      smallest of the full row unless the sample is 7 sigma unlucky)
//   3. collect       : ONE pass over all N columns, (key, column) of every distance <= hi_i appended to row i's list
//                      (~ b N / m entries, 4 % of the row)
//   4. select<EXACT> : K-th smallest key of the list, that pair's distance recomputed with the fp32 FMA chain of the radix
//                      kernel; rows whose list holds fewer than K entries (bracket too low) or overflowed are FLAGGED
//   5. the radix kernel re-does the 128-row blocks that contain a flagged row (pn_ms_kth_dist_tc_flagged; exits at once
//      otherwise), so the result never depends on the sample.
// The tiles are 64 columns wide (48 MMAs of N = 64 per tile step), operands by TMA as in the forward kernel (row-major
// forms only), S double-buffered in TMEM, epilogue thread = (row, 32-column half).
constexpr int KBN = 64;
constexpr int KSLAB = KBN * 128;                  // one [64 j rows][128 B] slab (d in [32 s, 32 s + 32))
constexpr int KPART = 4 * KSLAB;                  // 32 KB per operand part
constexpr int KSTAGE = 2 * KPART;                 // X | Xs
constexpr int KNST = 3;
constexpr uint32_t C_KS0 = 256;                   // S(t) at columns 256 + 64 (t & 1)

struct KBars { uint64_t x_full[KNST], x_empty[KNST], s_full[2], s_empty[2], a_ready; };

// grid (ceil(N / 128), B), 320 threads.  mC / mCs: maps of the column set (ncols rows; column r is point r * cstride).
// ALL: store the key of column r at slot r of the row's list (ncols <= cap); otherwise append (key, point) when dist <= hi.
template <bool ALL>
__global__ void __launch_bounds__(NT, 1)
ms_kth_collect_kernel(const __grid_constant__ CUtensorMap mC, const __grid_constant__ CUtensorMap mCs,
                      const float* __restrict__ X, int N, int ncols, int cstride, const float* __restrict__ hi, int cap,
                      unsigned* __restrict__ ckey, unsigned short* __restrict__ ccol, int* __restrict__ cnt_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ KBars bars;
    __shared__ uint32_t tmem_base_s;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * BM;
    const float* Xb = X + (long long)b * N * D;
    const int ntiles = (ncols + KBN - 1) / KBN;

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < KNST; ++s) { mbar_init(&bars.x_full[s], 1); mbar_init(&bars.x_empty[s], 1); }
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
        const bool ok = (i0 + row) < N;
        const float* xr = Xb + (long long)(i0 + row) * D + 64 * h;
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
            tmem_st16(tb + la + C_YB + 64 * h + c0, vb);
            tmem_st16(tb + la + C_YS + 64 * h + c0, vs);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars.a_ready);
        // admission threshold on S: every s with fl(2 - 2 s) <= hi satisfies s >= 1 - hi / 2 up to two roundings
        float sthr = INFINITY;                                                                   // rows >= N admit nothing
        if (!ALL && ok) {
            sthr = 1.0f - 0.5f * hi[(long long)b * N + i0 + row];
            sthr -= 4e-7f * fmaxf(1.0f, fabsf(sthr));
        }
        const long long lbase = ((long long)b * N + i0 + row) * cap;
        // appended entries: this thread's half of the tile columns goes to ITS half of the row's list (no atomics)
        const int half_cap = cap >> 1;
        unsigned* kdst = ckey + lbase + h * half_cap;
        unsigned short* cdst = ccol + lbase + h * half_cap;
        int mine = 0;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            const int j0 = t * KBN + 32 * h;
            mbar_wait_guarded(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t sv[32];
            tmem_ld32(tb + la + C_KS0 + 64 * k + 32 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty[k]);
            if (ALL) {
                if (ok) {
#pragma unroll
                    for (int u = 0; u < 32; u += 4) {
                        uint4 kk;                       // raw float bits of the distances (the select kernel orders them)
                        kk.x = __float_as_uint(2.0f - 2.0f * __uint_as_float(sv[u]));
                        kk.y = __float_as_uint(2.0f - 2.0f * __uint_as_float(sv[u + 1]));
                        kk.z = __float_as_uint(2.0f - 2.0f * __uint_as_float(sv[u + 2]));
                        kk.w = __float_as_uint(2.0f - 2.0f * __uint_as_float(sv[u + 3]));
                        // columns beyond ncols are zero-filled by TMA (dist = 2): the select kernel only reads ncols slots
                        if (j0 + u + 3 < cap) *reinterpret_cast<uint4*>(ckey + lbase + j0 + u) = kk;
                    }
                }
            } else {
                // admitted = { s >= sthr }: fl(2 - 2 s) is monotone in s, so this set is closed towards smaller distances and
                // contains every distance <= hi -- all the selection needs
                uint32_t mask = 0u;
#pragma unroll
                for (int u = 0; u < 32; ++u) mask |= (__uint_as_float(sv[u]) >= sthr) ? (1u << u) : 0u;
                if (j0 + 32 > ncols) mask &= (j0 < ncols) ? (0xffffffffu >> (32 - (ncols - j0))) : 0u;
                if (mask) {
                    const int add = __popc(mask);
                    if (mine + add > half_cap) {
                        mine = half_cap + 1;            // overflow: the row is flagged by the select kernel
                        sthr = INFINITY;
                    } else {
                        // one predicated store pair per column (no divergent branches, no dynamic register indexing)
                        int slot = mine;
#pragma unroll
                        for (int u = 0; u < 32; ++u) {
                            const uint32_t bit = (mask >> u) & 1u;
                            const uint32_t key = __float_as_uint(2.0f - 2.0f * __uint_as_float(sv[u]));
                            asm volatile(
                                "{\n"
                                ".reg .pred p;\n"
                                "setp.ne.u32 p, %0, 0;\n"
                                "@p st.global.u32 [%1], %2;\n"
                                "@p st.global.u16 [%3], %4;\n"
                                "}\n" ::"r"(bit), "l"(kdst + slot), "r"(key), "l"(cdst + slot), "h"((unsigned short)(j0 + u))
                                : "memory");
                            slot += (int)bit;
                        }
                        mine += add;
                    }
                }
            }
        }
        if (!ALL && ok) cnt_out[2 * ((long long)b * N + i0 + row) + h] = mine;
        tc_fence_before();
    } else if (warp == TMA_WARP) {
        if (elect_one()) {
            tma_prefetch_desc(&mC); tma_prefetch_desc(&mCs);
#pragma unroll 1
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % KNST;
                mbar_wait_guarded(&bars.x_empty[s], ((t / KNST) & 1) ^ 1);
                unsigned char* st = smem + s * KSTAGE;
                mbar_arrive_expect_tx(&bars.x_full[s], KSTAGE);
#pragma unroll
                for (int sl = 0; sl < 4; ++sl) {
                    tma_load_3d(st + sl * KSLAB, &mC, &bars.x_full[s], 32 * sl, t * KBN, b);
                    tma_load_3d(st + KPART + sl * KSLAB, &mCs, &bars.x_full[s], 32 * sl, t * KBN, b);
                }
            }
        }
        __syncwarp();
    } else {
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc(2, BM, KBN, 0, 0);
        const uint32_t sbase = smem_u32(smem);
        mbar_wait_guarded(&bars.a_ready, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % KNST, k = t & 1;
            mbar_wait_guarded(&bars.x_full[s], (t / KNST) & 1);
            mbar_wait_guarded(&bars.s_empty[k], ((t >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t st = sbase + s * KSTAGE;
            const uint64_t db0 = make_smem_desc(st, 16, SBO128, SW128);
            const uint64_t ds0 = make_smem_desc(st + KPART, 16, SBO128, SW128);
            const uint32_t d_s = tb + C_KS0 + 64 * k;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < D / 8; ++ks) {
                    const uint32_t off = (uint32_t)((ks >> 2) * KSLAB + (ks & 3) * 32);
                    const uint64_t db = db0 + (uint64_t)(off >> 4);
                    const uint64_t ds = ds0 + (uint64_t)(off >> 4);
                    mma_tf32_ts(d_s, tb + C_YS + ks * 8, db, idesc_s, ks > 0 ? 1u : 0u);
                    mma_tf32_ts(d_s, tb + C_YB + ks * 8, ds, idesc_s, 1);
                    mma_tf32_ts(d_s, tb + C_YB + ks * 8, db, idesc_s, 1);
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

// ---------------------------------------------------------------------------------------------- nms arg-selects, TMA-fed
// MeanShift.nms (reference src/mean_shift.py:146-149, :163-171): per row of A the column of Bm with the smallest distance
// (mode 0), or with the largest occupancy count among the columns closer than the bandwidth (mode 1); first occurrence on ties.
// Same tensor-core products and the same epilogue rule as ms_argsel_tc_kernel (meanshift_tc_argsel.cu), with the tile pipeline
// of ms_kth_collect_kernel: 64-column tiles fetched by TMA from Bm and its small split part (written by split_small_kernel),
// N = 64 MMAs: half as many tile hand-overs (barrier round trips, TMEM loads) per column as the 32-column loader-warp kernel.
__global__ void split_small_kernel(const float* __restrict__ x, long long n, float* __restrict__ xs) {
    const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (e + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(x + e);
        *reinterpret_cast<float4*>(xs + e) = make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
    } else {
        for (long long i = e; i < n; ++i) xs[i] = x[i] - tf32_hi(x[i]);
    }
}

// grid (ceil(Ma / 128), B), 320 threads
template <int MODE>
__global__ void __launch_bounds__(NT, 1)
ms_argsel_tma_kernel(const __grid_constant__ CUtensorMap mC, const __grid_constant__ CUtensorMap mCs, const float* __restrict__ A,
                     long long a_stride, int Ma, int Nb, const float* __restrict__ cnt, const float* __restrict__ thr,
                     int* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ KBars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ float part_v[BM];
    __shared__ int part_j[BM];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * BM;
    const float* Ab = A + (long long)b * a_stride;
    const int ntiles = (Nb + KBN - 1) / KBN;

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < KNST; ++s) { mbar_init(&bars.x_full[s], 1); mbar_init(&bars.x_empty[s], 1); }
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
            tmem_st16(tb + la + C_YB + 64 * h + c0, vb);
            tmem_st16(tb + la + C_YS + 64 * h + c0, vs);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars.a_ready);
        const float* cb = (MODE == 1) ? cnt + (long long)b * Nb : nullptr;
        const float th = (MODE == 1) ? thr[b] : 0.f;
        const bool vec_cnt = (MODE == 1) && ((reinterpret_cast<uintptr_t>(cb) & 15u) == 0);
        float best = (MODE == 0) ? INFINITY : -INFINITY;
        int bj = 0x7fffffff;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            const int j0 = t * KBN + 32 * h;
            float cv[32];
            if (MODE == 1) {                       // occupancy counts of this thread's 32 columns (the same for every row)
                if (vec_cnt && j0 + 32 <= Nb) {
#pragma unroll
                    for (int u = 0; u < 32; u += 4) {
                        const float4 c4 = __ldg(reinterpret_cast<const float4*>(cb + j0 + u));
                        cv[u] = c4.x; cv[u + 1] = c4.y; cv[u + 2] = c4.z; cv[u + 3] = c4.w;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 32; ++u) cv[u] = (j0 + u < Nb) ? __ldg(cb + j0 + u) : 0.f;
                }
            }
            mbar_wait_guarded(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t sv[32];
            tmem_ld32(tb + la + C_KS0 + 64 * k + 32 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty[k]);
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const int jj = j0 + u;
                if (jj < Nb) {
                    const float dist = 2.0f - 2.0f * __uint_as_float(sv[u]);
                    float v;
                    bool better;
                    if (MODE == 0) { v = dist; better = v < best; }
                    else { v = (dist < th) ? cv[u] : 0.f; better = v > best; }
                    best = better ? v : best;      // jj increases: the first occurrence of the extreme is kept
                    bj = better ? jj : bj;
                }
            }
        }
        // merge the two column halves of every row: within a tile the half h = 0 holds the lower columns, but over the tiles
        // the halves interleave -- ties go to the lower index
        if (h == 1) { part_v[row] = best; part_j[row] = bj; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (h == 0 && ok) {
            const float ov = part_v[row];
            const int oj = part_j[row];
            const bool take = (MODE == 0) ? (ov < best || (ov == best && oj < bj)) : (ov > best || (ov == best && oj < bj));
            out[(long long)b * Ma + i0 + row] = take ? oj : bj;
        }
        tc_fence_before();
    } else if (warp == TMA_WARP) {
        if (elect_one()) {
            tma_prefetch_desc(&mC); tma_prefetch_desc(&mCs);
#pragma unroll 1
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % KNST;
                mbar_wait_guarded(&bars.x_empty[s], ((t / KNST) & 1) ^ 1);
                unsigned char* st = smem + s * KSTAGE;
                mbar_arrive_expect_tx(&bars.x_full[s], KSTAGE);
#pragma unroll
                for (int sl = 0; sl < 4; ++sl) {
                    tma_load_3d(st + sl * KSLAB, &mC, &bars.x_full[s], 32 * sl, t * KBN, b);
                    tma_load_3d(st + KPART + sl * KSLAB, &mCs, &bars.x_full[s], 32 * sl, t * KBN, b);
                }
            }
        }
        __syncwarp();
    } else {
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc(2, BM, KBN, 0, 0);
        const uint32_t sbase = smem_u32(smem);
        mbar_wait_guarded(&bars.a_ready, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % KNST, k = t & 1;
            mbar_wait_guarded(&bars.x_full[s], (t / KNST) & 1);
            mbar_wait_guarded(&bars.s_empty[k], ((t >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t st = sbase + s * KSTAGE;
            const uint64_t db0 = make_smem_desc(st, 16, SBO128, SW128);
            const uint64_t ds0 = make_smem_desc(st + KPART, 16, SBO128, SW128);
            const uint32_t d_s = tb + C_KS0 + 64 * k;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < D / 8; ++ks) {
                    const uint32_t off = (uint32_t)((ks >> 2) * KSLAB + (ks & 3) * 32);
                    const uint64_t db = db0 + (uint64_t)(off >> 4);
                    const uint64_t ds = ds0 + (uint64_t)(off >> 4);
                    mma_tf32_ts(d_s, tb + C_YS + ks * 8, db, idesc_s, ks > 0 ? 1u : 0u);
                    mma_tf32_ts(d_s, tb + C_YB + ks * 8, ds, idesc_s, 1);
                    mma_tf32_ts(d_s, tb + C_YB + ks * 8, db, idesc_s, 1);
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

// one warp per row: K-th smallest of the row's keys by a bitwise radix select on bit-transposed registers (kth_select.cuh).
// EXACT = false: the row holds nfix keys in slots [0, nfix) (the sample pass), out = that key as a float.
// EXACT = true : the row holds cnt[2 row] keys from slot 0 and cnt[2 row + 1] keys from slot CAP / 2 (the two column halves
// of the collect pass); the selected pair's distance is recomputed with the fp32 FMA chain of ms_kth_tc_kernel; flags[row] = 1
// when the lists cannot hold the answer (fewer than K entries in total, or a half overflowed).
template <bool EXACT, int NW>
__global__ void __launch_bounds__(256)
ms_kth_select_kernel(const unsigned* __restrict__ ckey, const unsigned short* __restrict__ ccol, const int* __restrict__ cnt,
                     int nfix, int K, long long rows_total, int N, const float* __restrict__ X, float* __restrict__ out,
                     int* __restrict__ flags) {
    constexpr int CAP = 1024 * NW, HALF = CAP / 2;          // NW words of 32 keys per lane
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows_total) return;
    int n0 = nfix, n1 = 0;
    if (EXACT) {
        n0 = cnt[2 * row]; n1 = cnt[2 * row + 1];
        const bool bad = n0 > HALF || n1 > HALF || n0 + n1 < K;
        if (lane == 0) flags[row] = bad ? 1 : 0;
        if (bad) return;
    }
    const unsigned* kr = ckey + row * CAP;
    unsigned Bt[NW][32];
    unsigned act[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        act[w] = 0u;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int p = w * 1024 + r * 32 + lane;
            const bool valid = EXACT ? (p < HALF ? p < n0 : p - HALF < n1) : p < n0;
            Bt[w][r] = valid ? f2ord(__uint_as_float(kr[p])) : 0xffffffffu;
            act[w] |= valid ? kthsel::reg_bit(r) : 0u;
        }
        kthsel::bit_transpose32(Bt[w]);
    }
    const int n = n0 + n1;
    int need = K < n ? K : n;
    unsigned prefix = 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        int c = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) c += __popc(kthsel::step_zeros(act[w], Bt[w][i]));
        c = __reduce_add_sync(FULL, c);
        const bool zero = c >= need;
        if (!zero) { need -= c; prefix |= 1u << (31 - i); }
#pragma unroll
        for (int w = 0; w < NW; ++w) act[w] = kthsel::step_next(act[w], Bt[w][i], zero);
    }
    if (!EXACT) {
        if (lane == 0) out[row] = ord2f(prefix);
        return;
    }
    // lowest column among the entries that carry the selected key (act marks exactly those)
    int jbest = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            if (act[w] & kthsel::reg_bit(r)) {
                const int j = (int)ccol[row * CAP + w * 1024 + r * 32 + lane];
                jbest = j < jbest ? j : jbest;
            }
        }
    }
    jbest = __reduce_min_sync(FULL, jbest);
    if (lane == 0) {
        const long long b = row / N;
        const float4* xi = reinterpret_cast<const float4*>(X + row * D);
        const float4* xj = reinterpret_cast<const float4*>(X + (b * N + jbest) * D);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
        for (int c = 0; c < D / 4; ++c) {
            const float4 u4 = xi[c], v4 = xj[c];
            a0 = fmaf(u4.x, v4.x, a0); a1 = fmaf(u4.y, v4.y, a1);
            a2 = fmaf(u4.z, v4.z, a2); a3 = fmaf(u4.w, v4.w, a3);
        }
        out[row] = 2.0f - 2.0f * ((a0 + a1) + (a2 + a3));
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {       // resolved through the runtime: the library keeps no link dependency on libcuda
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// fp32 tensor [dim2][dim1][dim0] (dim0 contiguous), box {b0, b1, 1}, 128B swizzle, zero fill out of bounds
static bool make_map(CUtensorMap* m, const float* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t pitch1_elems,
                     uint64_t pitch2_elems, uint32_t b0, uint32_t b1) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {dim0, dim1, dim2};
    cuuint64_t strides[2] = {pitch1_elems * 4, pitch2_elems * 4};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mstma
}  // namespace pn

using namespace pn;

// X [B][N][128] -> Xs [B][N][128], Xt / Xst [B][128][Np] with Np = N rounded up to a multiple of 32 (caller allocates).
// Once per mean_shift call: X is constant over the iterations (reference src/mean_shift.py:58-77).
extern "C" int pn_ms_prepare_operands(const float* X, int B, int N, int d, int Np, float* Xs, float* Xt, float* Xst,
                                      void* stream) {
    PN_REQUIRE(X && Xs && Xt && Xst, "pn_ms_prepare_operands: null pointer");
    PN_REQUIRE(d == mstma::D, "pn_ms_prepare_operands: embedding width must be %d (got %d)", mstma::D, d);
    PN_REQUIRE(B > 0 && N > 0, "pn_ms_prepare_operands: empty batch (B=%d, N=%d)", B, N);
    PN_REQUIRE(Np >= N && Np % 32 == 0, "pn_ms_prepare_operands: Np must be N rounded up to a multiple of 32");
    mstma::ms_prep_operands_kernel<<<dim3(Np / 32, B), 256, 0, (cudaStream_t)stream>>>(X, N, Np, Xs, Xt, Xst);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_prep_operands_kernel");
    return PN_OK;
}

// CTA-group size.  The CTA-pair instantiations (cluster of 2, tcgen05 cta_group::2) were brought up on a B200 in round 2:
// bit-identical to the 1-CTA kernels but SLOWER (forward 4.12 vs 3.47 ms, backward 18.3 vs 13.5 ms at B = 16, N = 10^4;
// profiles/r02_first_call.md) -- with N = 32 MMAs the pair doubles the issue latency per tile without relieving the
// epilogue.  They are therefore compiled only with -DPN_MS_TMA_PAIRS (tools/exp_ms_tma.py documents the A/B); the
// shipped library contains the 1-CTA kernels only.
#ifdef PN_MS_TMA_PAIRS
static int cta_group() {      // read per call: tools/exp_ms_tma.py flips it inside one process
    const char* e = getenv("PN_MS_TMA_CG");
    return (e && e[0] == '2') ? 2 : 1;
}
#define PN_TMA_PICK(kern, cg) (((cg) == 2) ? kern<2> : kern<1>)
#else
static int cta_group() { return 1; }
#define PN_TMA_PICK(kern, cg) (kern<1>)
#endif

// the four operand-form maps of a [rows][128] matrix M (row-major M, Ms; transposed Mt, Mst with row pitch `tp`), boxes
// sized for one CTA of a group of `cg`
static bool make_forms(CUtensorMap* m, const float* M, const float* Ms, const float* Mt, const float* Mst, uint64_t rows,
                       uint64_t tp, uint64_t B, int cg) {
    const uint64_t dd = (uint64_t)mstma::D;
    return mstma::make_map(&m[0], M, dd, rows, B, dd, rows * dd, 32, 32 / cg) &&
           mstma::make_map(&m[1], Ms, dd, rows, B, dd, rows * dd, 32, 32 / cg) &&
           mstma::make_map(&m[2], Mt, tp, dd, B, tp, dd * tp, 32, 128 / cg) &&
           mstma::make_map(&m[3], Mst, tp, dd, B, tp, dd * tp, 32, 128 / cg);
}

template <typename... Args>
static cudaError_t launch(void (*kern)(Args...), dim3 grid, int cg, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    if (cg == 2) grid.x = (grid.x + 1) & ~1u;              // whole pairs; the padding CTA owns rows >= N only
    cfg.gridDim = grid;
    cfg.blockDim = dim3(mstma::NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// same contract as pn_ms_iter_fwd_tc (one mean-shift iteration, reference src/mean_shift.py:58-77) with the operand
// forms of pn_ms_prepare_operands
extern "C" int pn_ms_iter_fwd_tma(const float* Y, const float* X, const float* Xs, const float* Xt, const float* Xst,
                                  int B, int N, int d, int Np, const float* cinv, float* Ynew, float* den, float* unorm,
                                  void* stream) {
    PN_REQUIRE(Y && X && Xs && Xt && Xst && cinv && Ynew && den && unorm, "pn_ms_iter_fwd_tma: null pointer");
    PN_REQUIRE(d == mstma::D, "pn_ms_iter_fwd_tma: embedding width must be %d (got %d)", mstma::D, d);
    PN_REQUIRE(B > 0 && N > 0, "pn_ms_iter_fwd_tma: empty batch (B=%d, N=%d)", B, N);
    PN_REQUIRE(Np >= N && Np % 32 == 0, "pn_ms_iter_fwd_tma: Np must be N rounded up to a multiple of 32");
    PN_REQUIRE((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Xs) | reinterpret_cast<uintptr_t>(Xt) |
                reinterpret_cast<uintptr_t>(Xst)) % 16 == 0, "pn_ms_iter_fwd_tma: operand forms must be 16-byte aligned (TMA)");
    const int cg = cta_group();
    CUtensorMap m[4];
    if (!make_forms(m, X, Xs, Xt, Xst, (uint64_t)N, (uint64_t)Np, (uint64_t)B, cg)) {
        set_error("pn_ms_iter_fwd_tma: cuTensorMapEncodeTiled failed or is unavailable");
        return PN_ERR_CUDA;
    }
    size_t sm = mstma::NSTAGE * mstma::STAGE_BYTES + 1024;
    auto kern = PN_TMA_PICK(mstma::ms_fwd_tma_kernel, cg);
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_CUDA(launch(kern, dim3(cdiv(N, mstma::BM), B), cg, sm, (cudaStream_t)stream, m[0], m[1], m[2], m[3], Y, N, cinv, Ynew,
                   den, unorm));
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_fwd_tma_kernel");
    return PN_OK;
}

// K-th smallest of dist = 2 - 2 X X^T per row, every row of every shape against all N points of its shape
// (MeanShift.compute_bandwidth, reference src/mean_shift.py:115-137, with num_samples >= N), computed with ONE pass over
// the N x N distance tiles (see the comment above ms_kth_collect_kernel).  Xs = X - tf32_hi(X) (pn_ms_prepare_operands).
// Workspaces: ws_key [B*N][cap] u32, ws_col [B*N][cap] u16, ws_cnt [B*N][2] i32, ws_hi [B*N] f32; cap = 1024 or 2048.
// stride / b_sample: the column sample {0, stride, 2 stride, ...} (at most cap columns) and the order statistic of it that
// brackets the K-th smallest from above.  flags [B*N]: 1 for the rows whose bracket failed -- their kth entry is NOT
// written; the caller runs pn_ms_kth_dist_tc_flagged with the same flags afterwards (no host round trip).
extern "C" int pn_ms_kth_dist_tma(const float* X, const float* Xs, int B, int N, int d, int K, int stride, int b_sample,
                                  unsigned* ws_key, unsigned short* ws_col, int* ws_cnt, float* ws_hi, int cap, int* flags,
                                  float* kth, void* stream) {
    PN_REQUIRE(X && Xs && ws_key && ws_col && ws_cnt && ws_hi && flags && kth, "pn_ms_kth_dist_tma: null pointer");
    PN_REQUIRE(d == mstma::D, "pn_ms_kth_dist_tma: embedding width must be %d (got %d)", mstma::D, d);
    PN_REQUIRE(cap == 1024 || cap == 2048, "pn_ms_kth_dist_tma: candidate lists are built for cap = 1024 or 2048 (got %d)", cap);
    PN_REQUIRE(B > 0 && N >= 128 && N < 65536, "pn_ms_kth_dist_tma: need 128 <= N < 65536 (B=%d, N=%d)", B, N);
    PN_REQUIRE(stride >= 1, "pn_ms_kth_dist_tma: bad sample stride %d", stride);
    const int m = (N + stride - 1) / stride;
    PN_REQUIRE(m <= 1024 && (m % 4 == 0 || m + 3 < 1024), "pn_ms_kth_dist_tma: the column sample must hold at most 1024 columns "
               "(N=%d stride=%d -> %d columns)", N, stride, m);
    PN_REQUIRE(K >= 1 && K <= N && b_sample >= 1 && b_sample <= m, "pn_ms_kth_dist_tma: need 1 <= K <= N and 1 <= b_sample <= m "
               "(K=%d b_sample=%d m=%d)", K, b_sample, m);
    PN_REQUIRE((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Xs) | reinterpret_cast<uintptr_t>(ws_key)) % 16 == 0,
               "pn_ms_kth_dist_tma: X, Xs and ws_key must be 16-byte aligned");
    const uint64_t dd = (uint64_t)mstma::D;
    CUtensorMap ms[2], mf[2];
    if (!(mstma::make_map(&ms[0], X, dd, (uint64_t)m, (uint64_t)B, dd * stride, (uint64_t)N * dd, 32, mstma::KBN) &&
          mstma::make_map(&ms[1], Xs, dd, (uint64_t)m, (uint64_t)B, dd * stride, (uint64_t)N * dd, 32, mstma::KBN) &&
          mstma::make_map(&mf[0], X, dd, (uint64_t)N, (uint64_t)B, dd, (uint64_t)N * dd, 32, mstma::KBN) &&
          mstma::make_map(&mf[1], Xs, dd, (uint64_t)N, (uint64_t)B, dd, (uint64_t)N * dd, 32, mstma::KBN))) {
        set_error("pn_ms_kth_dist_tma: cuTensorMapEncodeTiled failed or is unavailable");
        return PN_ERR_CUDA;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sm = mstma::KNST * mstma::KSTAGE + 1024;
    auto k_all = mstma::ms_kth_collect_kernel<true>;
    auto k_thr = mstma::ms_kth_collect_kernel<false>;
    PN_CUDA(cudaFuncSetAttribute(k_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_CUDA(cudaFuncSetAttribute(k_thr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const dim3 grid(cdiv(N, mstma::BM), B);
    const long long rows = (long long)B * N;
    const int* no_cnt = nullptr;
    const unsigned short* no_col = nullptr;
    const float* no_hi = nullptr;
    int* no_flags = nullptr;
    PN_CUDA(launch(k_all, grid, 1, sm, st, ms[0], ms[1], X, N, m, stride, no_hi, cap, ws_key, ws_col, ws_cnt));
    PN_COUNT_LAUNCH();
    // (the sample lists use the first 1024 slots of a row whatever cap is: the row pitch of the select kernel is its CAP)
    if (cap == 1024)
        mstma::ms_kth_select_kernel<false, 1><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_key, no_col, no_cnt, m, b_sample, rows, N,
                                                                                   X, ws_hi, no_flags);
    else
        mstma::ms_kth_select_kernel<false, 2><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_key, no_col, no_cnt, m, b_sample, rows, N,
                                                                                   X, ws_hi, no_flags);
    PN_COUNT_LAUNCH();
    PN_CUDA(launch(k_thr, grid, 1, sm, st, mf[0], mf[1], X, N, N, 1, (const float*)ws_hi, cap, ws_key, ws_col, ws_cnt));
    PN_COUNT_LAUNCH();
    if (cap == 1024)
        mstma::ms_kth_select_kernel<true, 1><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_key, ws_col, ws_cnt, 0, K, rows, N, X, kth,
                                                                                  flags);
    else
        mstma::ms_kth_select_kernel<true, 2><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_key, ws_col, ws_cnt, 0, K, rows, N, X, kth,
                                                                                  flags);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_kth bracketed kernels");
    return PN_OK;
}

// same contract as pn_ms_argsel_tc (modes 0 and 1 of MeanShift.nms, reference src/mean_shift.py:146-149, :163-171) with the
// column tiles fetched by TMA; Bm must be contiguous [B][Nb][128] and 16-byte aligned, ws_Bms [B][Nb][128] receives its small
// split part (one extra launch).  Identical picks (same tensor-core products, same first-occurrence rule).
extern "C" int pn_ms_argsel_tma_supported(const float* Bm, long long b_stride, int Nb, int d) {
    return (d == mstma::D && Nb >= mstma::KBN && b_stride == (long long)Nb * d && (reinterpret_cast<uintptr_t>(Bm) & 15u) == 0) ? 1 : 0;
}

extern "C" int pn_ms_argsel_tma(int mode, const float* A, long long a_stride, int Ma, const float* Bm, long long b_stride, int Nb,
                                int B, int d, const float* cnt, const float* thr, float* ws_Bms, int* out, void* stream) {
    PN_REQUIRE(A && Bm && ws_Bms && out, "pn_ms_argsel_tma: null pointer");
    PN_REQUIRE((mode == 0) || (mode == 1 && cnt && thr), "pn_ms_argsel_tma: modes 0 and 1 only (mode 1 needs cnt, thr)");
    PN_REQUIRE(Ma > 0 && B > 0 && pn_ms_argsel_tma_supported(Bm, b_stride, Nb, d), "pn_ms_argsel_tma: unsupported problem (Nb=%d d=%d)",
               Nb, d);
    PN_REQUIRE((reinterpret_cast<uintptr_t>(ws_Bms) & 15u) == 0, "pn_ms_argsel_tma: ws_Bms must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)B * Nb * d;
    mstma::split_small_kernel<<<(unsigned)cdiv(cdiv(n, 4), 256), 256, 0, st>>>(Bm, n, ws_Bms);
    PN_COUNT_LAUNCH();
    const uint64_t dd = (uint64_t)mstma::D;
    CUtensorMap m[2];
    if (!(mstma::make_map(&m[0], Bm, dd, (uint64_t)Nb, (uint64_t)B, dd, (uint64_t)Nb * dd, 32, mstma::KBN) &&
          mstma::make_map(&m[1], ws_Bms, dd, (uint64_t)Nb, (uint64_t)B, dd, (uint64_t)Nb * dd, 32, mstma::KBN))) {
        set_error("pn_ms_argsel_tma: cuTensorMapEncodeTiled failed or is unavailable");
        return PN_ERR_CUDA;
    }
    const size_t sm = mstma::KNST * mstma::KSTAGE + 1024;
    const dim3 grid(cdiv(Ma, mstma::BM), B);
    if (mode == 0) {
        auto kern = mstma::ms_argsel_tma_kernel<0>;
        PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        PN_CUDA(launch(kern, grid, 1, sm, st, m[0], m[1], A, a_stride, Ma, Nb, cnt, thr, out));
    } else {
        auto kern = mstma::ms_argsel_tma_kernel<1>;
        PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        PN_CUDA(launch(kern, grid, 1, sm, st, m[0], m[1], A, a_stride, Ma, Nb, cnt, thr, out));
    }
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_argsel_tma_kernel");
    return PN_OK;
}

// same contract as pn_ms_iter_bwd_tc (backward of one mean-shift iteration) with the operand forms of
// pn_ms_prepare_operands for X and a workspace ws_C of 4 * B * 2 Nq * 128 floats (Nq = N rounded up to 16) for the
// operand forms of the interleaved [Yprev; Gn] matrix, which are rebuilt here every iteration.
extern "C" int pn_ms_bwd_prep_tc(const float* gout, const float* Ynew, const float* den, const float* unorm, int B, int N,
                                 int d, float* ws_Gn, float* ws_gd, void* stream);

extern "C" int pn_ms_iter_bwd_tma(const float* gout, const float* Ynew, const float* Yprev, const float* X,
                                  const float* Xs, const float* Xt, const float* Xst, const float* den,
                                  const float* unorm, int B, int N, int d, int Np, const float* cinv, float* ws_Gn,
                                  float* ws_gd, float* ws_C, float* gYprev, float* gX, int accumulate_gX, void* stream) {
    PN_REQUIRE(gout && Ynew && Yprev && X && Xs && Xt && Xst && den && unorm && cinv && ws_Gn && ws_gd && ws_C && gYprev && gX,
               "pn_ms_iter_bwd_tma: null pointer");
    PN_REQUIRE(d == mstma::D, "pn_ms_iter_bwd_tma: embedding width must be %d (got %d)", mstma::D, d);
    PN_REQUIRE(B > 0 && N > 0, "pn_ms_iter_bwd_tma: empty batch (B=%d, N=%d)", B, N);
    PN_REQUIRE(Np >= N && Np % 32 == 0, "pn_ms_iter_bwd_tma: Np must be N rounded up to a multiple of 32");
    PN_REQUIRE((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Xs) | reinterpret_cast<uintptr_t>(Xt) |
                reinterpret_cast<uintptr_t>(Xst) | reinterpret_cast<uintptr_t>(ws_C)) % 16 == 0,
               "pn_ms_iter_bwd_tma: operand forms / workspace must be 16-byte aligned (TMA)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = pn_ms_bwd_prep_tc(gout, Ynew, den, unorm, B, N, d, ws_Gn, ws_gd, stream);
    if (rc != PN_OK) return rc;
    const int cg = cta_group();
    const int Nq = (N + 15) / 16 * 16;
    const uint64_t r2 = 2ull * (uint64_t)Nq;
    const size_t form = (size_t)B * r2 * mstma::D;
    float *C = ws_C, *Cs = ws_C + form, *Ct = ws_C + 2 * form, *Cst = ws_C + 3 * form;
    mstma::ms_prep_concat_kernel<<<dim3(Nq / 16, B), 256, 0, st>>>(Yprev, ws_Gn, N, Nq, C, Cs, Ct, Cst);
    PN_COUNT_LAUNCH();
    CUtensorMap mx[4], mc[4];
    if (!make_forms(mx, X, Xs, Xt, Xst, (uint64_t)N, (uint64_t)Np, (uint64_t)B, cg) ||
        !make_forms(mc, C, Cs, Ct, Cst, r2, r2, (uint64_t)B, cg)) {
        set_error("pn_ms_iter_bwd_tma: cuTensorMapEncodeTiled failed or is unavailable");
        return PN_ERR_CUDA;
    }
    size_t sm = mstma::NSTAGE * mstma::STAGE_BYTES + 2 * 64 * 32 * sizeof(float) + 1024;
    auto rows_k = PN_TMA_PICK(mstma::ms_bwd_rows_tma_kernel, cg);
    auto cols_k = PN_TMA_PICK(mstma::ms_bwd_cols_tma_kernel, cg);
    PN_CUDA(cudaFuncSetAttribute(rows_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_CUDA(cudaFuncSetAttribute(cols_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_CUDA(launch(rows_k, dim3(cdiv(N, 64), B), cg, sm, st, mx[0], mx[1], mx[2], mx[3], Yprev, (const float*)ws_Gn,
                   (const float*)ws_gd, N, cinv, gYprev));
    PN_COUNT_LAUNCH();
    PN_CUDA(launch(cols_k, dim3(cdiv(N, 128), B), cg, sm, st, mc[0], mc[1], mc[2], mc[3], X, (const float*)ws_gd, N, cinv, gX,
                   accumulate_gX));
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_bwd_*_tma kernels");
    return PN_OK;
}
