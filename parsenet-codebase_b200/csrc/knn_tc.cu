// Feature-space kNN graph (reference `knn`, src/PointNet.py:9-26, src/model.py:9-22) with the tensor cores as a FILTER:
// the result is the bit-exact graph of knn.cu / knn_tma.cu (fp32 fmaf chain in channel order, ties to the lower index), but
// only ~k candidates per row are evaluated that way.
//
//   cost_ij = (xx_j - 2 <x_i, x_j>) + xx_i      (smaller = nearer; the kernels above rank D = -cost)
//   tensor-core value  cost~_ij from a split-TF32 tcgen05 product; |cost~_ij - cost_ij| <= eps_ij := c0 (xx_i + xx_j)
//   (c0 bounds the dropped small x small terms, the operand truncations, the tensor core's accumulation and the fmaf chain's own
//   rounding, with |x_i||x_j| <= (xx_i + xx_j) / 2; see knn_tc_c0 below), hence  lo_ij := cost~ - eps <= cost_ij <= hi_ij := cost~ + eps.
//
//   1. prep      : xx (fmaf chain), Xs = X - tf32_hi(X), per-column admission constants
//   2. collect<ALL> over a strided column SAMPLE: hi_ij of every sample column, T_i := b-th smallest of them (kth_select.cuh).
//                  b sits 7 sigma above the hypergeometric mean k m / N, so at least k columns of the full row have hi <= T_i
//                  (unless the sample is unlucky: detected in step 4)
//   3. collect   : ONE tensor-core pass over all columns; (raw product, column) of every lo_ij <= T_i appended to row i's lists
//   4. final     : L_i := k-th smallest hi of the list (>= the exact k-th smallest cost, because k columns have cost <= hi <= L_i).
//                  L_i > T_i, a short or an overflowed list -> the row is FLAGGED.  Otherwise every column with lo <= L_i is in
//                  the list and the exact top k are among them: those survivors (k plus the few inside the 2 eps window) get
//                  the exact fp32 cost, are sorted by (cost, index) and the first k are written.
//   5. knn_tma_kernel re-does the 64-row tiles that contain a flagged row (pn_knn_tma_flagged; other tiles exit at once).
#include "common.cuh"
#include "tc05.cuh"
#include "kth_select.cuh"
#include "knn_select.cuh"
#include <cuda.h>

namespace pn {
namespace knntc {
using namespace tc05;

constexpr int BM = 128, KBN = 64, NT = 320, EPI_WARPS = 8, TMA_WARP = 8, MMA_WARP = 9, EPI_THREADS = 256, KNST = 3;
constexpr int KSLAB = KBN * 128;                      // one [64 columns][128 B] slab (32 channels)
constexpr uint32_t SW128 = 2, SBO128 = 1024;
constexpr int MAXSURV = 128;                          // survivors per row the final kernel sorts

struct Bars { uint64_t x_full[KNST], x_empty[KNST], s_full[2], s_empty[2], a_ready; };

// relative half-width of the tensor-core value's error interval, in units of (xx_i + xx_j):
//   3 * 2^-20   operands: x = hi + lo_t + r with |r| <= 2^-20 |x| (hi, lo truncated to tf32), the lo x lo product is dropped
//   (3 C / 8 + 8) * 2^-22   fp32 accumulation of 3 C / 8 MMAs (truncating adds, K = 8 products aligned per MMA)
//   C * 2^-24   the reference-order fmaf chain's own distance from the exact dot product
//   8 * 2^-24   the two subtractions forming the cost, the norms
// all times |x_i||x_j| <= (xx_i + xx_j) / 2 for the product terms and doubled by the factor 2 of the cost; 1.25 safety factor.
__host__ __device__ inline float knn_tc_c0(int C) {
    const double e = 3.0 / 1048576.0 + (3.0 * C / 8.0 + 8.0) / 4194304.0 + C / 16777216.0 + 8.0 / 16777216.0;
    return (float)(1.25 * e);
}

// ---------------------------------------------------------------------------------------------- preparation
// x [rows][ld] -> xx [rows] (fmaf chain), Xs [rows][C] = x - tf32_hi(x), an [B][Np] = xx (1 - c0a) (admission, all columns),
// ap [B][mp] = xx (1 + c0) of the sample columns {0, stride, ...} (Np, mp: row pitches, multiples of 64, padding = +inf / unused)
__global__ void __launch_bounds__(256) prep_kernel(const float* __restrict__ x, int ld, int C, int N, int stride, int Np, int mp,
                                                   float c0, float c0a, float* __restrict__ xx, float* __restrict__ Xs,
                                                   float* __restrict__ an, float* __restrict__ ap) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= Np) return;
    if (i >= N) {
        if (lane == 0) an[(long long)b * Np + i] = INFINITY;          // never admitted
        return;
    }
    const float* p = x + ((long long)b * N + i) * ld;
    float* ps = Xs + ((long long)b * N + i) * C;
    for (int c = lane; c < C; c += 32) { const float v = p[c]; ps[c] = v - tf32_hi(v); }
    if (lane == 0) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc = fmaf(p[c], p[c], acc);
        xx[(long long)b * N + i] = acc;
        an[(long long)b * Np + i] = acc * (1.0f - c0a);
        if (i % stride == 0) ap[(long long)b * mp + i / stride] = acc * (1.0f + c0);
    }
}

// ---------------------------------------------------------------------------------------------- tensor-core pass
// grid (ceil(N / 128), B), 320 threads: warps 0-7 epilogue (thread = row, 32-column half), warp 8 TMA, warp 9 MMA issue.
// ALL : columns = the sample (ncols = m, map row pitch = stride rows); hi_ij stored at slot j of row i's list.
// !ALL: columns = all N points; (raw product bits, column) appended to the list of (row, column half) when lo_ij <= T_i.
template <int C, bool ALL>
__global__ void __launch_bounds__(NT, C <= 64 ? 2 : 1)
collect_kernel(const __grid_constant__ CUtensorMap mC, const __grid_constant__ CUtensorMap mCs, const float* __restrict__ X,
               int ld, int N, int ncols, const float* __restrict__ xx, const float* __restrict__ colc, int colp,
               const float* __restrict__ T, float c0, float c0a, int cap, unsigned* __restrict__ cval,
               unsigned short* __restrict__ ccol, int* __restrict__ cnt_out) {
    constexpr int NSL = C / 32, KPART = NSL * KSLAB, KSTAGE = 2 * KPART;
    constexpr uint32_t C_AB = 0, C_AS = C, C_S0 = 2 * C, TMEM_COLS = (2 * C + 2 * KBN) <= 256 ? 256 : 512;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * BM;
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
        const float* xr = X + ((long long)b * N + i0 + row) * ld + (C / 2) * h;
#pragma unroll 1
        for (int c0_ = 0; c0_ < C / 2; c0_ += 16) {
            uint32_t vb[16], vs[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                float4 v = ok ? *reinterpret_cast<const float4*>(xr + c0_ + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float big = tf32_hi(f[u]);
                    vb[e + u] = __float_as_uint(big);
                    vs[e + u] = __float_as_uint(f[u] - big);
                }
            }
            tmem_st16(tb + la + C_AB + (C / 2) * h + c0_, vb);
            tmem_st16(tb + la + C_AS + (C / 2) * h + c0_, vs);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars.a_ready);
        const long long grow = (long long)b * N + i0 + row;
        const float xxi = ok ? xx[grow] : 0.f;
        // ALL : hi_ij = fmaf(-2, s, xx_j (1 + c0)) + xx_i (1 + c0)
        // !ALL: lo_ij <= T_i  <=>  fmaf(-2, s, xx_j (1 - c0a)) <= T_i - xx_i (1 - c0a)   (c0a > c0 covers this test's own rounding)
        const float rowc = ALL ? xxi * (1.0f + c0) : ((ok ? T[grow] : -INFINITY) - xxi * (1.0f - c0a));
        const float* cc = colc + (long long)b * colp;
        const long long lbase = grow * cap;
        const int half_cap = cap >> 1;
        unsigned* kdst = cval + lbase + h * half_cap;
        unsigned short* cdst = ccol + lbase + h * half_cap;
        int mine = 0;
        bool dead = !ok;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            const int j0 = t * KBN + 32 * h;
            float cj[32];
#pragma unroll
            for (int u = 0; u < 32; u += 4) {
                const float4 v = *reinterpret_cast<const float4*>(cc + j0 + u);        // (colp is padded to whole tiles)
                cj[u] = v.x; cj[u + 1] = v.y; cj[u + 2] = v.z; cj[u + 3] = v.w;
            }
            mbar_wait_guarded(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t sv[32];
            tmem_ld32(tb + la + C_S0 + KBN * k + 32 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty[k]);
            if (ALL) {
                if (ok) {
#pragma unroll
                    for (int u = 0; u < 32; u += 4) {
                        uint4 kk;
                        kk.x = __float_as_uint(fmaf(-2.0f, __uint_as_float(sv[u]), cj[u]) + rowc);
                        kk.y = __float_as_uint(fmaf(-2.0f, __uint_as_float(sv[u + 1]), cj[u + 1]) + rowc);
                        kk.z = __float_as_uint(fmaf(-2.0f, __uint_as_float(sv[u + 2]), cj[u + 2]) + rowc);
                        kk.w = __float_as_uint(fmaf(-2.0f, __uint_as_float(sv[u + 3]), cj[u + 3]) + rowc);
                        if (j0 + u + 3 < 1024) *reinterpret_cast<uint4*>(cval + lbase + j0 + u) = kk;
                    }
                }
            } else if (!dead) {
                uint32_t mask = 0u;
#pragma unroll
                for (int u = 0; u < 32; ++u) mask |= (fmaf(-2.0f, __uint_as_float(sv[u]), cj[u]) <= rowc) ? (1u << u) : 0u;
                // (padding columns carry +inf constants and are never admitted)
                if (mask) {
                    const int add = __popc(mask);
                    if (mine + add > half_cap) {
                        mine = half_cap + 1;            // overflow: the row is flagged by the final kernel
                        dead = true;
                    } else {
                        int slot = mine;
#pragma unroll
                        for (int u = 0; u < 32; ++u) {
                            const uint32_t bit = (mask >> u) & 1u;
                            asm volatile(
                                "{\n"
                                ".reg .pred p;\n"
                                "setp.ne.u32 p, %0, 0;\n"
                                "@p st.global.u32 [%1], %2;\n"
                                "@p st.global.u16 [%3], %4;\n"
                                "}\n" ::"r"(bit), "l"(kdst + slot), "r"(sv[u]), "l"(cdst + slot), "h"((unsigned short)(j0 + u))
                                : "memory");
                            slot += (int)bit;
                        }
                        mine += add;
                    }
                }
            }
        }
        if (!ALL && ok) cnt_out[2 * grow + h] = mine;
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
                for (int sl = 0; sl < NSL; ++sl) {
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
            const uint32_t d_s = tb + C_S0 + KBN * k;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < C / 8; ++ks) {
                    const uint32_t off = (uint32_t)((ks >> 2) * KSLAB + (ks & 3) * 32);
                    const uint64_t db = db0 + (uint64_t)(off >> 4);
                    const uint64_t ds = ds0 + (uint64_t)(off >> 4);
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

// ---------------------------------------------------------------------------------------------- tensor-core pass, C = 256
// Both split parts of 128 query rows x 256 channels would fill all 512 TMEM columns and leave no room for the accumulator.
// Here a CTA owns 64 query rows and stacks the two parts along the LANES ("virtual rows"): lanes 0-63 hold tf32_hi(x_i), lanes
// 64-127 hold x_i - tf32_hi(x_i), 256 columns.  One MMA chain against the raw column tile (the tensor core truncates it to its
// big part) and one against its small part accumulate into the same D: lanes 0-63 end up with hi.hi + hi.lo, lanes 64-127 with
// lo.hi + lo.lo, and the product of row i is D[i] + D[64 + i] (all four terms, two MMAs per K step instead of three).  The two
// halves live in different warps: warps 2, 3 (+4) hand theirs over through shared memory, warps 0, 1 (+4) add and run the
// admission.  A stage holds 128 of the 256 channels of a 64-column tile (64 KB, as for C = 128): every tile consumes two.
// grid (ceil(N / 64), B), 320 threads.
constexpr int QR = 64;                                 // query rows per CTA of the C = 256 kernel

template <bool ALL>
__global__ void __launch_bounds__(NT, 1)
collect256_kernel(const __grid_constant__ CUtensorMap mC, const __grid_constant__ CUtensorMap mCs, const float* __restrict__ X,
                  int ld, int N, int ncols, const float* __restrict__ xx, const float* __restrict__ colc, int colp,
                  const float* __restrict__ T, float c0, float c0a, int cap, unsigned* __restrict__ cval,
                  unsigned short* __restrict__ ccol, int* __restrict__ cnt_out) {
    constexpr int C = 256, NSL = 4, KPART = NSL * KSLAB, KSTAGE = 2 * KPART;          // one stage = 128 channels, X | Xs
    constexpr uint32_t C_A = 0, C_S0 = 256, TMEM_COLS = 512;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* xchg = reinterpret_cast<float*>(smem + KNST * KSTAGE);                      // [64 rows][64 columns + 1]
    constexpr int XP = KBN + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * QR;
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
        const int L = q * 32 + lane;                   // TMEM lane
        const int row = L & 63;                        // query row of the CTA
        const bool lo_part = L >= 64;
        const uint32_t la = (uint32_t)(q * 32) << 16;
        const bool ok = (i0 + row) < N;
        const float* xr = X + ((long long)b * N + i0 + row) * ld + 128 * h;
#pragma unroll 1
        for (int c0_ = 0; c0_ < 128; c0_ += 16) {
            uint32_t va[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                float4 v = ok ? *reinterpret_cast<const float4*>(xr + c0_ + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float big = tf32_hi(f[u]);
                    va[e + u] = __float_as_uint(lo_part ? f[u] - big : big);
                }
            }
            tmem_st16(tb + la + C_A + 128 * h + c0_, va);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars.a_ready);
        const long long grow = (long long)b * N + i0 + row;
        const float xxi = ok ? xx[grow] : 0.f;
        const float rowc = ALL ? xxi * (1.0f + c0) : ((ok ? T[grow] : -INFINITY) - xxi * (1.0f - c0a));
        const float* cc = colc + (long long)b * colp;
        const long long lbase = grow * cap;
        const int half_cap = cap >> 1;
        unsigned* kdst = cval + lbase + h * half_cap;
        unsigned short* cdst = ccol + lbase + h * half_cap;
        int mine = 0;
        bool dead = !ok;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            const int j0 = t * KBN + 32 * h;
            mbar_wait_guarded(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            uint32_t sv[32];
            tmem_ld32(tb + la + C_S0 + KBN * k + 32 * h, sv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty[k]);
            if (lo_part) {
#pragma unroll
                for (int u = 0; u < 32; ++u) xchg[row * XP + 32 * h + u] = __uint_as_float(sv[u]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (!lo_part) {
#pragma unroll
                for (int u = 0; u < 32; ++u) sv[u] = __float_as_uint(__uint_as_float(sv[u]) + xchg[row * XP + 32 * h + u]);
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (lo_part) continue;
            float cj[32];
#pragma unroll
            for (int u = 0; u < 32; u += 4) {
                const float4 v = *reinterpret_cast<const float4*>(cc + j0 + u);
                cj[u] = v.x; cj[u + 1] = v.y; cj[u + 2] = v.z; cj[u + 3] = v.w;
            }
            if (ALL) {
                if (ok) {
#pragma unroll
                    for (int u = 0; u < 32; u += 4) {
                        uint4 kk;
                        kk.x = __float_as_uint(fmaf(-2.0f, __uint_as_float(sv[u]), cj[u]) + rowc);
                        kk.y = __float_as_uint(fmaf(-2.0f, __uint_as_float(sv[u + 1]), cj[u + 1]) + rowc);
                        kk.z = __float_as_uint(fmaf(-2.0f, __uint_as_float(sv[u + 2]), cj[u + 2]) + rowc);
                        kk.w = __float_as_uint(fmaf(-2.0f, __uint_as_float(sv[u + 3]), cj[u + 3]) + rowc);
                        if (j0 + u + 3 < 1024) *reinterpret_cast<uint4*>(cval + lbase + j0 + u) = kk;
                    }
                }
            } else if (!dead) {
                uint32_t mask = 0u;
#pragma unroll
                for (int u = 0; u < 32; ++u) mask |= (fmaf(-2.0f, __uint_as_float(sv[u]), cj[u]) <= rowc) ? (1u << u) : 0u;
                if (mask) {
                    const int add = __popc(mask);
                    if (mine + add > half_cap) {
                        mine = half_cap + 1;
                        dead = true;
                    } else {
                        int slot = mine;
#pragma unroll
                        for (int u = 0; u < 32; ++u) {
                            const uint32_t bit = (mask >> u) & 1u;
                            asm volatile(
                                "{\n"
                                ".reg .pred p;\n"
                                "setp.ne.u32 p, %0, 0;\n"
                                "@p st.global.u32 [%1], %2;\n"
                                "@p st.global.u16 [%3], %4;\n"
                                "}\n" ::"r"(bit), "l"(kdst + slot), "r"(sv[u]), "l"(cdst + slot), "h"((unsigned short)(j0 + u))
                                : "memory");
                            slot += (int)bit;
                        }
                        mine += add;
                    }
                }
            }
        }
        if (!ALL && ok && !lo_part) cnt_out[2 * grow + h] = mine;
        tc_fence_before();
    } else if (warp == TMA_WARP) {
        if (elect_one()) {
            tma_prefetch_desc(&mC); tma_prefetch_desc(&mCs);
#pragma unroll 1
            for (int u = 0; u < 2 * ntiles; ++u) {       // stage use u = (tile, channel half)
                const int s = u % KNST, t = u >> 1, kh = u & 1;
                mbar_wait_guarded(&bars.x_empty[s], ((u / KNST) & 1) ^ 1);
                unsigned char* st = smem + s * KSTAGE;
                mbar_arrive_expect_tx(&bars.x_full[s], KSTAGE);
#pragma unroll
                for (int sl = 0; sl < NSL; ++sl) {
                    tma_load_3d(st + sl * KSLAB, &mC, &bars.x_full[s], 128 * kh + 32 * sl, t * KBN, b);
                    tma_load_3d(st + KPART + sl * KSLAB, &mCs, &bars.x_full[s], 128 * kh + 32 * sl, t * KBN, b);
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
        for (int u = 0; u < 2 * ntiles; ++u) {
            const int s = u % KNST, t = u >> 1, kh = u & 1, k = t & 1;
            mbar_wait_guarded(&bars.x_full[s], (u / KNST) & 1);
            if (kh == 0) mbar_wait_guarded(&bars.s_empty[k], ((t >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t st = sbase + s * KSTAGE;
            const uint64_t db0 = make_smem_desc(st, 16, SBO128, SW128);
            const uint64_t ds0 = make_smem_desc(st + KPART, 16, SBO128, SW128);
            const uint32_t d_s = tb + C_S0 + KBN * k;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    const uint32_t off = (uint32_t)((ks >> 2) * KSLAB + (ks & 3) * 32);
                    const uint64_t db = db0 + (uint64_t)(off >> 4);
                    const uint64_t ds = ds0 + (uint64_t)(off >> 4);
                    const uint32_t acol = tb + C_A + 128 * kh + ks * 8;
                    mma_tf32_ts(d_s, acol, ds, idesc_s, (kh == 0 && ks == 0) ? 0u : 1u);
                    mma_tf32_ts(d_s, acol, db, idesc_s, 1);
                }
                if (kh == 1) mma_commit(&bars.s_full[k]);
                mma_commit(&bars.x_empty[s]);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tb, TMEM_COLS);
    (void)C;
}

// ---------------------------------------------------------------------------------------------- final: select, refine, sort
// one warp per row.  Lists as written by collect_kernel<C, false>: cnt[2 row + h] entries from slot h * CAP / 2.
template <int NW, typename IdxT>
__global__ void __launch_bounds__(256)
final_kernel(const unsigned* __restrict__ cval, const unsigned short* __restrict__ ccol, const int* __restrict__ cnt,
             const float* __restrict__ T, const float* __restrict__ xx, const float* __restrict__ X, int ld, int C, int N, int k,
             float c0, long long rows_total, IdxT* __restrict__ idx_out, float* __restrict__ dist_out, int* __restrict__ flags) {
    constexpr int CAP = 1024 * NW, HALF = CAP / 2;
    extern __shared__ float fsm[];                       // per warp: query row [C] | survivor columns [MAXSURV] (as ints)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * 8 + wid;
    if (row >= rows_total) return;
    float* xq = fsm + (size_t)wid * (C + MAXSURV);
    int* scol = reinterpret_cast<int*>(xq + C);
    const int n0 = cnt[2 * row], n1 = cnt[2 * row + 1];
    if (n0 > HALF || n1 > HALF || n0 + n1 < k) {
        if (lane == 0) flags[row] = 1;
        return;
    }
    const long long b = row / N;
    const float xxi = xx[row];
    const float* xxb = xx + b * N;
    unsigned Bt[NW][32];
    float lo[NW][32];
    unsigned act[NW], valid[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        act[w] = 0u;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int p = w * 1024 + r * 32 + lane;
            const bool v = p < HALF ? p < n0 : p - HALF < n1;
            float hi = INFINITY;
            lo[w][r] = INFINITY;
            if (v) {
                const float s = __uint_as_float(cval[row * CAP + p]);
                const float xxj = xxb[ccol[row * CAP + p]];
                const float cost = fmaf(-2.0f, s, xxj) + xxi;
                const float eps = c0 * (xxi + xxj);
                hi = cost + eps;
                lo[w][r] = cost - eps;
            }
            Bt[w][r] = v ? f2ord(hi) : 0xffffffffu;
            act[w] |= v ? kthsel::reg_bit(r) : 0u;
        }
        valid[w] = act[w];
        kthsel::bit_transpose32(Bt[w]);
    }
    int need = k;
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
    const float L = ord2f(prefix);                       // k-th smallest upper bound >= the exact k-th smallest cost
    // survivors: lower bound <= L
    int mycount = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int r = 0; r < 32; ++r) mycount += ((valid[w] & kthsel::reg_bit(r)) && lo[w][r] <= L) ? 1 : 0;
    int incl = mycount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (!(L <= T[row]) || total > MAXSURV) {             // bracket too low (possible misses) or too many inside the window
        if (lane == 0) flags[row] = 1;
        return;
    }
    if (lane == 0) flags[row] = 0;
    int pos = incl - mycount;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int r = 0; r < 32; ++r)
            if ((valid[w] & kthsel::reg_bit(r)) && lo[w][r] <= L) scol[pos++] = (int)ccol[row * CAP + w * 1024 + r * 32 + lane];
    const float* xi = X + row * ld;
    for (int c = lane; c < C; c += 32) xq[c] = xi[c];
    __syncwarp();
    // exact cost of the survivors: the fmaf chain in channel order and the cost expression of knn.cu / knn_tma.cu
    unsigned long long key[MAXSURV / 32];
#pragma unroll
    for (int e = 0; e < MAXSURV / 32; ++e) {
        const int p = e * 32 + lane;
        key[e] = ~0ull;
        if (p < total) {
            const int j = scol[p];
            const float4* xj = reinterpret_cast<const float4*>(X + (b * N + j) * ld);
            float acc = 0.f;
            for (int c4 = 0; c4 < C / 4; ++c4) {
                const float4 v = xj[c4];
                acc = fmaf(xq[4 * c4], v.x, acc);
                acc = fmaf(xq[4 * c4 + 1], v.y, acc);
                acc = fmaf(xq[4 * c4 + 2], v.z, acc);
                acc = fmaf(xq[4 * c4 + 3], v.w, acc);
            }
            const float inner = __fmul_rn(-2.0f, acc);
            const float d = __fsub_rn(__fsub_rn(-xxb[j], inner), xxi);
            key[e] = knn::make_key(d, j);
        }
    }
    knn::warp_bitonic_sort<MAXSURV / 32>(key, lane);
#pragma unroll
    for (int e = 0; e < MAXSURV / 32; ++e) {
        const int p = e * 32 + lane;
        if (p < k) {
            idx_out[row * k + p] = (IdxT)(uint32_t)(key[e] & 0xffffffffu);
            if (dist_out) dist_out[row * k + p] = ord2f(~(uint32_t)(key[e] >> 32));
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// fp32 [B][rows][C] with row pitch `pitch` elements and shape pitch `bpitch`, box {32 channels, 64 rows, 1}, 128B swizzle
static bool make_map(CUtensorMap* m, const float* base, uint64_t C, uint64_t rows, uint64_t B, uint64_t pitch, uint64_t bpitch) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {C, rows, B};
    cuuint64_t strides[2] = {pitch * 4, bpitch * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)KBN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C>
static int run_collect(const CUtensorMap* ms, const CUtensorMap* mf, const float* x, int ld, int B, int N, int m, const float* xx,
                       const float* an, int Np, const float* ap, int mp, float* T, int b_sample, float c0, float c0a, int cap,
                       unsigned* ws_val, unsigned short* ws_col, int* ws_cnt, cudaStream_t st) {
    constexpr size_t sm = (size_t)KNST * 2 * (C / 32) * KSLAB + 1024;
    auto k_all = collect_kernel<C, true>;
    auto k_thr = collect_kernel<C, false>;
    PN_CUDA(cudaFuncSetAttribute(k_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_CUDA(cudaFuncSetAttribute(k_thr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const dim3 grid(cdiv(N, BM), B);
    const long long rows = (long long)B * N;
    const float* noT = nullptr;
    k_all<<<grid, NT, sm, st>>>(ms[0], ms[1], x, ld, N, m, xx, ap, mp, noT, c0, c0a, cap, ws_val, ws_col, ws_cnt);
    PN_COUNT_LAUNCH();
    if (cap == 1024) kthsel::kth_smallest_rows_kernel<1><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_val, m, b_sample, rows, T);
    else kthsel::kth_smallest_rows_kernel<2><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_val, m, b_sample, rows, T);
    PN_COUNT_LAUNCH();
    k_thr<<<grid, NT, sm, st>>>(mf[0], mf[1], x, ld, N, N, xx, an, Np, (const float*)T, c0, c0a, cap, ws_val, ws_col, ws_cnt);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn_tc collect kernels");
    return PN_OK;
}

static int run_collect256(const CUtensorMap* ms, const CUtensorMap* mf, const float* x, int ld, int B, int N, int m, const float* xx,
                          const float* an, int Np, const float* ap, int mp, float* T, int b_sample, float c0, float c0a, int cap,
                          unsigned* ws_val, unsigned short* ws_col, int* ws_cnt, cudaStream_t st) {
    const size_t sm = (size_t)KNST * 2 * 4 * KSLAB + (size_t)QR * (KBN + 1) * sizeof(float) + 1024;
    auto k_all = collect256_kernel<true>;
    auto k_thr = collect256_kernel<false>;
    PN_CUDA(cudaFuncSetAttribute(k_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_CUDA(cudaFuncSetAttribute(k_thr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const dim3 grid(cdiv(N, QR), B);
    const long long rows = (long long)B * N;
    const float* noT = nullptr;
    k_all<<<grid, NT, sm, st>>>(ms[0], ms[1], x, ld, N, m, xx, ap, mp, noT, c0, c0a, cap, ws_val, ws_col, ws_cnt);
    PN_COUNT_LAUNCH();
    if (cap == 1024) kthsel::kth_smallest_rows_kernel<1><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_val, m, b_sample, rows, T);
    else kthsel::kth_smallest_rows_kernel<2><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_val, m, b_sample, rows, T);
    PN_COUNT_LAUNCH();
    k_thr<<<grid, NT, sm, st>>>(mf[0], mf[1], x, ld, N, N, xx, an, Np, (const float*)T, c0, c0a, cap, ws_val, ws_col, ws_cnt);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn_tc collect256 kernels");
    return PN_OK;
}

template <typename IdxT>
static int run_final(int cap, const unsigned* ws_val, const unsigned short* ws_col, const int* ws_cnt, const float* T,
                     const float* xx, const float* x, int ld, int C, int N, int k, float c0, long long rows, void* idx, float* dist,
                     int* flags, cudaStream_t st) {
    const size_t sm = (size_t)8 * (C + MAXSURV) * sizeof(float);
    if (cap == 1024)
        final_kernel<1, IdxT><<<(unsigned)cdiv(rows, 8), 256, sm, st>>>(ws_val, ws_col, ws_cnt, T, xx, x, ld, C, N, k, c0, rows,
                                                                      (IdxT*)idx, dist, flags);
    else
        final_kernel<2, IdxT><<<(unsigned)cdiv(rows, 8), 256, sm, st>>>(ws_val, ws_col, ws_cnt, T, xx, x, ld, C, N, k, c0, rows,
                                                                      (IdxT*)idx, dist, flags);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn_tc final_kernel");
    return PN_OK;
}

}  // namespace knntc
}  // namespace pn

using namespace pn;

// 1 if pn_knn_tc takes this problem: feature-space metric, C = 64, 128 or 256, TMA-compatible addressing, 16-bit column indices
extern "C" int pn_knn_tc_supported(const float* x, int N, int C, int ld, int k, int metric) {
    if (metric != 0 || (C != 64 && C != 128 && C != 256) || ld % 4 || (reinterpret_cast<uintptr_t>(x) & 15u)) return 0;
    if (N >= 65536 || N < 2048 || k < 1 || k > 96) return 0;
    return 1;
}

// kNN graph of pn_knn / pn_knn_tma (same arguments, same bit-exact result) through the tensor-core filter described at the top.
// stride / b_sample: the column sample {0, stride, ...} (at most 1024 columns) and the order statistic of its upper bounds used
// as bracket.  Workspaces: ws_norms [B*N]; ws_xs [B*N*C]; ws_colc [B*(Np + mp)] with Np = N rounded up to 64 and mp = the sample
// size rounded up to 64; ws_T [B*N]; ws_val [B*N][cap] u32; ws_col [B*N][cap] u16; ws_cnt [B*N][2]; cap = 1024 or 2048.
// flags [B*N] is written: rows with flag 1 are NOT written -- run pn_knn_tma_flagged with the same flags next.
extern "C" int pn_knn_tc(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64,
                         float* dist_out, int stride, int b_sample, float* ws_norms, float* ws_xs, float* ws_colc, float* ws_T,
                         unsigned* ws_val, unsigned short* ws_col, int* ws_cnt, int cap, int* flags, void* stream) {
    PN_REQUIRE(x && idx_out && ws_norms && ws_xs && ws_colc && ws_T && ws_val && ws_col && ws_cnt && flags, "pn_knn_tc: null pointer");
    PN_REQUIRE(B > 0 && pn_knn_tc_supported(x, N, C, ld, k, metric), "pn_knn_tc: unsupported problem (N=%d C=%d ld=%d k=%d)", N, C,
               ld, k);
    PN_REQUIRE(cap == 1024 || cap == 2048, "pn_knn_tc: lists are built for cap = 1024 or 2048 (got %d)", cap);
    PN_REQUIRE(stride >= 1, "pn_knn_tc: bad sample stride %d", stride);
    const int m = (N + stride - 1) / stride;
    PN_REQUIRE(m <= 1024 && b_sample >= 1 && b_sample <= m, "pn_knn_tc: need at most 1024 sample columns and 1 <= b_sample <= m "
               "(N=%d stride=%d m=%d b_sample=%d)", N, stride, m, b_sample);
    PN_REQUIRE((reinterpret_cast<uintptr_t>(ws_xs) | reinterpret_cast<uintptr_t>(ws_val) | reinterpret_cast<uintptr_t>(ws_colc)) % 16 == 0,
               "pn_knn_tc: workspaces must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int Np = (N + 63) / 64 * 64, mp = (m + 63) / 64 * 64;
    float* an = ws_colc;
    float* ap = ws_colc + (size_t)B * Np;
    const float c0 = knntc::knn_tc_c0(C), c0a = 1.05f * c0;
    PN_CUDA(cudaMemsetAsync(ap, 0, (size_t)B * mp * sizeof(float), st));
    knntc::prep_kernel<<<dim3(cdiv(Np, 8), B), 256, 0, st>>>(x, ld, C, N, stride, Np, mp, c0, c0a, ws_norms, ws_xs, an, ap);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn_tc prep_kernel");
    CUtensorMap ms[2], mf[2];
    const uint64_t uC = (uint64_t)C, uN = (uint64_t)N, uB = (uint64_t)B, uld = (uint64_t)ld;
    if (!(knntc::make_map(&ms[0], x, uC, (uint64_t)m, uB, uld * stride, uN * uld) &&
          knntc::make_map(&ms[1], ws_xs, uC, (uint64_t)m, uB, uC * stride, uN * uC) &&
          knntc::make_map(&mf[0], x, uC, uN, uB, uld, uN * uld) && knntc::make_map(&mf[1], ws_xs, uC, uN, uB, uC, uN * uC))) {
        set_error("pn_knn_tc: cuTensorMapEncodeTiled failed or is unavailable");
        return PN_ERR_CUDA;
    }
    int rc = (C == 64) ? knntc::run_collect<64>(ms, mf, x, ld, B, N, m, ws_norms, an, Np, ap, mp, ws_T, b_sample, c0, c0a, cap,
                                                ws_val, ws_col, ws_cnt, st)
           : (C == 128) ? knntc::run_collect<128>(ms, mf, x, ld, B, N, m, ws_norms, an, Np, ap, mp, ws_T, b_sample, c0, c0a, cap,
                                                  ws_val, ws_col, ws_cnt, st)
                        : knntc::run_collect256(ms, mf, x, ld, B, N, m, ws_norms, an, Np, ap, mp, ws_T, b_sample, c0, c0a, cap,
                                                ws_val, ws_col, ws_cnt, st);
    if (rc != PN_OK) return rc;
    const long long rows = (long long)B * N;
    return idx_is_i64 ? knntc::run_final<long long>(cap, ws_val, ws_col, ws_cnt, ws_T, ws_norms, x, ld, C, N, k, c0, rows, idx_out,
                                                    dist_out, flags, st)
                      : knntc::run_final<int>(cap, ws_val, ws_col, ws_cnt, ws_T, ws_norms, x, ld, C, N, k, c0, rows, idx_out,
                                              dist_out, flags, st);
}
