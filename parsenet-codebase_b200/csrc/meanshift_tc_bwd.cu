// Backward of one mean-shift iteration on tcgen05 (split-TF32), same pipeline as meanshift_tc.cu.
// Math (see meanshift.cu): S = Y X^T, K = exp(clamp((S-1)c)), G = Gn X^T, gS = (G + gd_i) K c (0 where clamped)
//   rows:  gY_i  = sum_j gS_ij x_j                    (CTA owns query rows i)
//   cols:  gX_j += sum_i gS_ij y_i + K_ij Gn_i        (CTA owns rows j of X)
// Both reuse the forward pipeline (A operands in TMEM, 32-row B tiles through smem, 48 + 12 MMAs per tile):
//   MODE_ROWS: the CTA owns 64 real rows; TMEM lanes 0-63 hold Y_i and lanes 64-127 hold Gn_i ("virtual rows"), so ONE
//              MMA chain yields S (lanes 0-63) and G (lanes 64-127) against the streamed X tile; G is handed to the
//              lanes that own S through shared memory; P = gS on lanes 0-63, zero on lanes 64-127.
//   MODE_COLS: the CTA owns 128 rows of X; the streamed 32-row tile is the CONCATENATION of 16 rows of Y and the
//              same 16 rows of Gn, so D[:,0:16] = S^T and D[:,16:32] = G^T; P = [gS^T | K^T] and the second product
//              with the transposed copy of the same tile gives gS^T Y + K^T Gn in one chain.
#include "common.cuh"
#include "tc05.cuh"
#include <stdlib.h>

namespace pn {
namespace mstcb {
using namespace tc05;

constexpr int D = 128, BN = 32, NT = 416, NSTAGE = 3, FLUSH = 16;
constexpr int EPI_WARPS = 8, LOAD_WARP0 = 8, MMA_WARP = 12, EPI_THREADS = 256;
constexpr float CLAMP = 75.f;
constexpr uint32_t C_AB = 0, C_AS = 128, C_D0 = 256, C_PB = 320, C_PS = 352, C_O = 384, TMEM_COLS = 512;
constexpr int XA_BYTES = BN * D * 4, XB_BYTES = D * BN * 4, STAGE_BYTES = 2 * XA_BYTES + 2 * XB_BYTES;
constexpr uint32_t XA_LBO = BN * 16, XB_LBO = D * 16, SBO = 128;
enum { MODE_ROWS = 0, MODE_COLS = 1 };

__device__ volatile int* g_dbg = nullptr;   // bring-up aid: host-mapped progress words (set by pn_debug_set_progress)
#ifdef PN_MS_DEBUG
#define DBG(slot, val) do { if (g_dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) { g_dbg[slot] = (val); __threadfence_system(); } } while (0)
#else
#define DBG(slot, val) do { } while (0)      // (each live DBG site costs an LDG of g_dbg on the critical path)
#endif

// PDB (VAR bit 7, the default): P is double-buffered.  P_big of tile t is written IN PLACE over the S / G accumulator
// columns of buffer t&1 (every epilogue thread overwrites exactly the columns it has just loaded), P_small has its own
// two 32-column buffers.  The epilogue of tile t then only needs s_full(t) -- tcgen05.mma executes in issue order, so
// "first product of tile t complete" implies "second product of tile t-2 complete" -- and no longer waits for the second
// product of tile t-1 before it can hand over P(t): a whole tile step of slack on the MMA -> epilogue -> MMA round trip
// (the ablation sweep in profiles/r01_ms_bwd_ablation_sweep.md showed the kernel bound by exactly that latency chain).
constexpr uint32_t C_PS2 = 320;      // PDB: P_small buffers [320,352) and [352,384); P_big aliases C_D0 + 32k

struct Bars {
    uint64_t x_full[NSTAGE], x_empty[NSTAGE], s_full[2], s_empty[2], p_full, p_empty, o_flush, o_done;
    uint64_t p_full2[2], a_ready;    // PDB: per-buffer "P ready", separate "A operands in TMEM"
};

// LITE: the gradient-side operand P (= gS, resp. [gS^T | K^T]) of the second product is rounded to tf32 (round to nearest)
// instead of split, i.e. 2 instead of 3 MMAs per k-step for P.X (the streamed operand keeps its exact split).  The
// exponent argument S and G = Gn.X^T (a cancellation against gd) keep the full 3-MMA split.  Simulated effect on the
// gradients (float64 model with the same rounding, tests/test_cpu_host_logic.py): 1-4e-4 of the largest entry.
// Opt-in (PN_MS_BWD_LITE=1); the default is the exact split.
// VAR bit 0 = LITE.  The other bits are TIMING-ONLY ablations (results are wrong by construction; tools/exp_ms_bwd.py):
//   2 no exp in the epilogue, 4 loaders skip the "small" split stores, 8 no second-product MMAs, 16 one instead of three
//   MMAs in the first product, 32 loaders skip the transposed copy, 64 epilogue skips the P stores to TMEM.
template <int MODE, int VAR>
// (13 warps are allocated as 16: 128 registers per thread is the hardware ceiling for this CTA shape -- __maxnreg__(144/152)
// fails to launch; lifting it needs setmaxnreg re-balancing between the loader / MMA and the epilogue warpgroups)
__global__ void __launch_bounds__(NT, 1)
ms_bwd_tc_kernel(const float* __restrict__ Yp, const float* __restrict__ X, const float* __restrict__ Gn,
                 const float* __restrict__ gd, int N, const float* __restrict__ cinv, float* __restrict__ out,
                 int accumulate) {
    constexpr bool LITE = (VAR & 1) != 0, A_NOEXP = (VAR & 2) != 0, A_NOSMALL = (VAR & 4) != 0, A_NOG2 = (VAR & 8) != 0,
                   A_G1ONE = (VAR & 16) != 0, A_NOXB = (VAR & 32) != 0, A_NOPST = (VAR & 64) != 0, PDB = (VAR & 128) != 0;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    float* exch = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);     // MODE_ROWS: [2][64][32] G hand-over

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const long long off = (long long)b * N * D;
    const float* Ypb = Yp + off;
    const float* Xb = X + off;
    const float* Gb = Gn + off;
    const float* gdb = gd + (long long)b * N;
    constexpr int OWN = (MODE == MODE_ROWS) ? 64 : 128;      // real rows owned by the CTA
    constexpr int TROWS = (MODE == MODE_ROWS) ? 32 : 16;     // real rows of the streamed matrices per tile
    const int r0 = blockIdx.x * OWN;
    const int ntiles = (N + TROWS - 1) / TROWS;

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
        const int q = warp & 3, h = warp >> 2;
        const int vrow = q * 32 + lane;                         // TMEM lane
        const uint32_t la = (uint32_t)(q * 32) << 16;
        const float c = cinv[b];
        const float c2 = c * 1.4426950408889634f, CL2 = CLAMP * 1.4426950408889634f;
        // ---- A operand rows -> TMEM (split)
        const float* arow;
        bool aok;
        if (MODE == MODE_ROWS) {
            const int r = r0 + (vrow & 63);
            aok = r < N;
            arow = ((vrow < 64) ? Ypb : Gb) + (long long)r * D + 64 * h;
        } else {
            aok = (r0 + vrow) < N;
            arow = Xb + (long long)(r0 + vrow) * D + 64 * h;
        }
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
            tmem_st16(tb + la + C_AB + 64 * h + c0, vb);
            tmem_st16(tb + la + C_AS + 64 * h + c0, vs);
        }
        if (MODE == MODE_ROWS && q >= 2) {                       // P rows of the "virtual" half stay zero forever
            uint32_t z[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) z[u] = 0u;
            if (PDB) {                                           // (P_big of the virtual half is re-zeroed every tile)
                tmem_st16(tb + la + C_PS2 + 16 * h, z);
                tmem_st16(tb + la + C_PS2 + 32 + 16 * h, z);
            } else {
                tmem_st16(tb + la + C_PB + 16 * h, z);
                tmem_st16(tb + la + C_PS + 16 * h, z);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(PDB ? &bars.a_ready : &bars.p_full);         // (non-PDB: phase 0 of p_full) A operands are in TMEM
        const bool owner = (MODE == MODE_COLS) || (q < 2);       // threads that own output rows
        float oacc[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) oacc[e] = 0.f;
        const float gd_row = (MODE == MODE_ROWS && owner && aok) ? gdb[r0 + vrow] : 0.f;
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int k = t & 1;
            DBG(warp, t * 10 + 1);
            mbar_wait(&bars.s_full[k], (t >> 1) & 1);
            tc_fence_after();
            DBG(warp, t * 10 + 2);
            uint32_t pb[16], ps[16];
            if (MODE == MODE_ROWS) {
                uint32_t sv[16];
                tmem_ld16(tb + la + C_D0 + 32 * k + 16 * h, sv);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars.s_empty[k]);
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
                        float kk = A_NOEXP ? e : ex2_approx(e);
                        float p = (!cl && (j0 + u < N)) ? (g + gd_row) * kk * c : 0.f;
                        float big = LITE ? to_tf32(p) : tf32_hi(p);
                        pb[u] = __float_as_uint(big);
                        ps[u] = __float_as_uint(p - big);
                    }
                }
            } else {
                uint32_t s8[8], g8[8];
                tmem_ld8(tb + la + C_D0 + 32 * k + 8 * h, s8);
                tmem_ld8(tb + la + C_D0 + 32 * k + 16 + 8 * h, g8);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars.s_empty[k]);
                const int i0t = t * 16 + 8 * h;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool iv = (i0t + u) < N;
                    const float gdi = iv ? gdb[i0t + u] : 0.f;
                    float e = (__uint_as_float(s8[u]) - 1.0f) * c2;
                    const bool cl = (e > CL2) || (e < -CL2);
                    e = fminf(fmaxf(e, -CL2), CL2);
                    float kk = iv ? (A_NOEXP ? e : ex2_approx(e)) : 0.f;
                    float p1 = (!cl && iv) ? (__uint_as_float(g8[u]) + gdi) * kk * c : 0.f;
                    float b1 = LITE ? to_tf32(p1) : tf32_hi(p1), b2 = LITE ? to_tf32(kk) : tf32_hi(kk);
                    pb[u] = __float_as_uint(b1);       ps[u] = __float_as_uint(p1 - b1);       // gS^T -> cols 8h..
                    pb[8 + u] = __float_as_uint(b2);   ps[8 + u] = __float_as_uint(kk - b2);   // K^T  -> cols 16+8h..
                }
            }
            DBG(warp, t * 10 + 3);
            const bool flush_now = (t > 0 && (t % FLUSH) == 0);
            if (!PDB || flush_now) {                 // second product of tile t-1 complete (PDB: only needed to drain O)
                mbar_wait(&bars.p_empty, (t & 1) ^ 1);
                tc_fence_after();
            }
            DBG(warp, t * 10 + 4);
            if (flush_now) {
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
                mbar_arrive(&bars.o_flush);
            }
            const uint32_t pb_col = PDB ? (C_D0 + 32 * k) : C_PB, ps_col = PDB ? (C_PS2 + 32 * k) : C_PS;
            if (MODE == MODE_ROWS) {
                if (q < 2 && !A_NOPST) {
                    tmem_st16(tb + la + pb_col + 16 * h, pb);
                    if (!LITE) tmem_st16(tb + la + ps_col + 16 * h, ps);
                    tmem_st_wait();
                }
                if (PDB && q >= 2) {                 // the G values of the virtual half must not act as P rows
                    uint32_t z[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) z[u] = 0u;
                    tmem_st16(tb + la + pb_col + 16 * h, z);
                    tmem_st_wait();
                }
            } else if (!A_NOPST) {
                uint32_t a[8], d[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { a[u] = pb[u]; d[u] = pb[8 + u]; }
                tmem_st8(tb + la + pb_col + 8 * h, a);
                tmem_st8(tb + la + pb_col + 16 + 8 * h, d);
                if (!LITE) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) { a[u] = ps[u]; d[u] = ps[8 + u]; }
                    tmem_st8(tb + la + ps_col + 8 * h, a);
                    tmem_st8(tb + la + ps_col + 16 + 8 * h, d);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(PDB ? &bars.p_full2[k] : &bars.p_full);
        }
        mbar_wait(&bars.o_done, 0);
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
        }
        tc_fence_before();
    } else if (warp < MMA_WARP) {
        // =============================================================================== loader warps
        const int lw = warp - LOAD_WARP0;
        const int j = 8 * lw + (lane & 7), cq = lane >> 3, l4 = lane & 3, jg = j >> 2;   // see meanshift_tc.cu
        auto src_row = [&](int t) -> const float* {
            if (MODE == MODE_ROWS) {
                const int r = t * 32 + j;
                return (t < ntiles && r < N) ? Xb + (long long)r * D : nullptr;
            } else {
                const int r = t * 16 + (j & 15);
                return (t < ntiles && r < N) ? ((j < 16) ? Ypb : Gb) + (long long)r * D : nullptr;
            }
        };
        float4 vin[8], vnx[8];
        {
            const float* p = src_row(0);
#pragma unroll
            for (int it = 0; it < 8; ++it)
                vin[it] = p ? *reinterpret_cast<const float4*>(p + 4 * (4 * it + cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % NSTAGE;
            {
                const float* p = src_row(t + 1);
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    vnx[it] = p ? *reinterpret_cast<const float4*>(p + 4 * (4 * it + cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            DBG(warp, t * 10 + 1);
            mbar_wait(&bars.x_empty[s], ((t / NSTAGE) & 1) ^ 1);
            DBG(warp, t * 10 + 2);
            unsigned char* st = smem + s * STAGE_BYTES;
            unsigned char* xa_b = st;
            unsigned char* xa_s = st + XA_BYTES;
            unsigned char* xb_b = st + 2 * XA_BYTES;
            unsigned char* xb_s = st + 2 * XA_BYTES + XB_BYTES;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int c4 = 4 * it + cq;
                float f0 = vin[it].x, f1 = vin[it].y, f2 = vin[it].z, f3 = vin[it].w;
                {
                    const float b0 = tf32_hi(f0), b1 = tf32_hi(f1), b2 = tf32_hi(f2), b3 = tf32_hi(f3);
                    const uint32_t oa = (uint32_t)(c4 * XA_LBO + (j >> 3) * 128 + (j & 7) * 16);
                    *reinterpret_cast<float4*>(xa_b + oa) = make_float4(b0, b1, b2, b3);
                    if (!A_NOSMALL) *reinterpret_cast<float4*>(xa_s + oa) = make_float4(f0 - b0, f1 - b1, f2 - b2, f3 - b3);
                }
                if (!A_NOXB) {
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
                if (!A_NOXB) {
                    const float b0 = tf32_hi(f0), b1 = tf32_hi(f1), b2 = tf32_hi(f2), b3 = tf32_hi(f3);
                    const int d = 4 * c4 + l4;
                    const uint32_t ob = (uint32_t)(jg * XB_LBO + (d >> 3) * 128 + (d & 7) * 16);
                    *reinterpret_cast<float4*>(xb_b + ob) = make_float4(b0, b1, b2, b3);
                    if (!A_NOSMALL) *reinterpret_cast<float4*>(xb_s + ob) = make_float4(f0 - b0, f1 - b1, f2 - b2, f3 - b3);
                }
            }
            fence_async_smem();
            mbar_arrive(&bars.x_full[s]);
#pragma unroll
            for (int it = 0; it < 8; ++it) vin[it] = vnx[it];
        }
    } else {
        // =============================================================================== MMA warp (warp-uniform)
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc(2, 128, BN, 0, 0);
        const uint32_t idesc_o = make_idesc(2, 128, D, 0, 0);
        const uint32_t sbase = smem_u32(smem);
        auto gemm2 = [&](int u) {
            DBG(13, u * 10 + 1);
            if (PDB) mbar_wait(&bars.p_full2[u & 1], (u >> 1) & 1);
            else mbar_wait(&bars.p_full, (u + 1) & 1);
            DBG(13, u * 10 + 2);
            const uint32_t pb_col = PDB ? (C_D0 + 32 * (u & 1)) : C_PB, ps_col = PDB ? (C_PS2 + 32 * (u & 1)) : C_PS;
            const bool fresh = (u % FLUSH) == 0;
            if (u > 0 && fresh) mbar_wait(&bars.o_flush, ((u / FLUSH) - 1) & 1);
            tc_fence_after();
            const uint32_t st = sbase + (u % NSTAGE) * STAGE_BYTES;
            const uint64_t db0 = make_smem_desc(st + 2 * XA_BYTES, XB_LBO, SBO, 0);
            const uint64_t ds0 = make_smem_desc(st + 2 * XA_BYTES + XB_BYTES, XB_LBO, SBO, 0);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < (A_NOG2 ? 0 : BN / 8); ++ks) {
                    const uint64_t db = db0 + (uint64_t)(ks * ((2 * XB_LBO) >> 4));
                    const uint64_t ds = ds0 + (uint64_t)(ks * ((2 * XB_LBO) >> 4));
                    if (LITE) {
                        mma_tf32_ts(tb + C_O, tb + pb_col + ks * 8, ds, idesc_o, (fresh && ks == 0) ? 0u : 1u);
                    } else {
                        mma_tf32_ts(tb + C_O, tb + ps_col + ks * 8, db, idesc_o, (fresh && ks == 0) ? 0u : 1u);
                        mma_tf32_ts(tb + C_O, tb + pb_col + ks * 8, ds, idesc_o, 1);
                    }
                    mma_tf32_ts(tb + C_O, tb + pb_col + ks * 8, db, idesc_o, 1);
                }
                mma_commit(&bars.x_empty[u % NSTAGE]);
                mma_commit(&bars.p_empty);
            }
            __syncwarp();
        };
        DBG(14, 1);
        mbar_wait(PDB ? &bars.a_ready : &bars.p_full, 0);
        DBG(14, 2);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % NSTAGE, k = t & 1;
            DBG(12, t * 10 + 1);
            mbar_wait(&bars.x_full[s], (t / NSTAGE) & 1);
            DBG(12, t * 10 + 2);
            mbar_wait(&bars.s_empty[k], ((t >> 1) & 1) ^ 1);
            DBG(12, t * 10 + 3);
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
                    if (A_G1ONE) {
                        mma_tf32_ts(d_s, tb + C_AB + ks * 8, db, idesc_s, ks > 0 ? 1u : 0u);
                        continue;
                    }
                    mma_tf32_ts(d_s, tb + C_AS + ks * 8, db, idesc_s, ks > 0 ? 1u : 0u);
                    mma_tf32_ts(d_s, tb + C_AB + ks * 8, ds, idesc_s, 1);
                    mma_tf32_ts(d_s, tb + C_AB + ks * 8, db, idesc_s, 1);
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

// prep (same as meanshift.cu): g_u = (g - Y'(Y'.g))/|u| ; Gn = g_u/den ; gd = -(g_u . Y'|u|)/den
__global__ void ms_bwd_prep_tc_kernel(const float* __restrict__ gout, const float* __restrict__ Ynew,
                                      const float* __restrict__ den, const float* __restrict__ unorm, long long rows,
                                      float* __restrict__ Gn, float* __restrict__ gd) {
    long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float4 g = *reinterpret_cast<const float4*>(gout + r * D + 4 * lane);
    float4 y = *reinterpret_cast<const float4*>(Ynew + r * D + 4 * lane);
    float dot = warp_sum(g.x * y.x + g.y * y.y + g.z * y.z + g.w * y.w);
    float nr = unorm[r], dn = den[r];
    float4 gu = make_float4((g.x - y.x * dot) / nr, (g.y - y.y * dot) / nr, (g.z - y.z * dot) / nr, (g.w - y.w * dot) / nr);
    float gt = warp_sum((gu.x * y.x + gu.y * y.y + gu.z * y.z + gu.w * y.w) * nr);
    *reinterpret_cast<float4*>(Gn + r * D + 4 * lane) = make_float4(gu.x / dn, gu.y / dn, gu.z / dn, gu.w / dn);
    if (lane == 0) gd[r] = -gt / dn;
}

}  // namespace mstcb
}  // namespace pn

using namespace pn;

extern "C" int pn_ms_iter_bwd_tc(const float* gout, const float* Ynew, const float* Yprev, const float* X,
                                 const float* den, const float* unorm, int B, int N, int d, const float* cinv,
                                 float* ws_Gn, float* ws_gd, float* gYprev, float* gX, int accumulate_gX,
                                 void* stream) {
    PN_REQUIRE(gout && Ynew && Yprev && X && den && unorm && cinv && ws_Gn && ws_gd && gYprev && gX,
               "pn_ms_iter_bwd_tc: null pointer");
    PN_REQUIRE(d == mstcb::D, "pn_ms_iter_bwd_tc: embedding width must be %d (got %d)", mstcb::D, d);
    cudaStream_t st = (cudaStream_t)stream;
    long long rows = (long long)B * N;
    mstcb::ms_bwd_prep_tc_kernel<<<cdiv(rows, 8), 256, 0, st>>>(gout, Ynew, den, unorm, rows, ws_Gn, ws_gd);
    PN_COUNT_LAUNCH();
    size_t sm = mstcb::NSTAGE * mstcb::STAGE_BYTES + 2 * 64 * 32 * sizeof(float) + 1024;
    static const bool lite = [] { const char* e = getenv("PN_MS_BWD_LITE"); return e && e[0] == '1'; }();
    static const bool pdb = [] { const char* e = getenv("PN_MS_BWD_PDB"); return !(e && e[0] == '0'); }();
    int var = (lite ? 1 : 0) | (pdb ? 128 : 0);
    if (const char* e = getenv("PN_MS_BWD_ABLATE")) var = atoi(e);          // timing-only ablations (wrong results)
    using KernelT = void (*)(const float*, const float*, const float*, const float*, int, const float*, float*, int);
    KernelT rows_k = nullptr, cols_k = nullptr;
#define PN_VAR_CASE(V) case V: rows_k = mstcb::ms_bwd_tc_kernel<0, V>; cols_k = mstcb::ms_bwd_tc_kernel<1, V>; break;
    switch (var) {
        PN_VAR_CASE(0) PN_VAR_CASE(1) PN_VAR_CASE(2) PN_VAR_CASE(4) PN_VAR_CASE(8) PN_VAR_CASE(16) PN_VAR_CASE(32)
        PN_VAR_CASE(64) PN_VAR_CASE(36) PN_VAR_CASE(24) PN_VAR_CASE(126) PN_VAR_CASE(128) PN_VAR_CASE(129) PN_VAR_CASE(254)
        default: PN_REQUIRE(false, "pn_ms_iter_bwd_tc: unknown PN_MS_BWD_ABLATE variant %d", var);
    }
#undef PN_VAR_CASE
    PN_CUDA(cudaFuncSetAttribute(rows_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_CUDA(cudaFuncSetAttribute(cols_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const char* dbg = getenv("PN_MS_BWD_TC_ONLY");      // bring-up switch: "rows" / "cols" run a single kernel
    if (!dbg || dbg[0] == 'r') {
        rows_k<<<dim3(cdiv(N, 64), B), mstcb::NT, sm, st>>>(Yprev, X, ws_Gn, ws_gd, N, cinv, gYprev, 0);
        PN_COUNT_LAUNCH();
    }
    if (!dbg || dbg[0] == 'c') {
        cols_k<<<dim3(cdiv(N, 128), B), mstcb::NT, sm, st>>>(Yprev, X, ws_Gn, ws_gd, N, cinv, gX, accumulate_gX);
        PN_COUNT_LAUNCH();
    }
    PN_LAUNCH_CHECK("ms_bwd_tc kernels");
    return PN_OK;
}

// The two halves of pn_ms_iter_bwd_tc as separate entry points (used by pn_ms_iter_bwd_tma in meanshift_tma.cu, which
// replaces only the rows kernel): the prep pass, and the cols kernel (gX) in its default variant.
extern "C" int pn_ms_bwd_prep_tc(const float* gout, const float* Ynew, const float* den, const float* unorm, int B, int N,
                                 int d, float* ws_Gn, float* ws_gd, void* stream) {
    PN_REQUIRE(gout && Ynew && den && unorm && ws_Gn && ws_gd, "pn_ms_bwd_prep_tc: null pointer");
    PN_REQUIRE(d == mstcb::D, "pn_ms_bwd_prep_tc: embedding width must be %d (got %d)", mstcb::D, d);
    long long rows = (long long)B * N;
    mstcb::ms_bwd_prep_tc_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(gout, Ynew, den, unorm, rows, ws_Gn, ws_gd);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_bwd_prep_tc_kernel");
    return PN_OK;
}

extern "C" int pn_ms_bwd_cols_tc(const float* Yprev, const float* X, int B, int N, int d, const float* cinv,
                                 const float* ws_Gn, const float* ws_gd, float* gX, int accumulate_gX, void* stream) {
    PN_REQUIRE(Yprev && X && cinv && ws_Gn && ws_gd && gX, "pn_ms_bwd_cols_tc: null pointer");
    PN_REQUIRE(d == mstcb::D, "pn_ms_bwd_cols_tc: embedding width must be %d (got %d)", mstcb::D, d);
    size_t sm = mstcb::NSTAGE * mstcb::STAGE_BYTES + 2 * 64 * 32 * sizeof(float) + 1024;
    auto cols_k = mstcb::ms_bwd_tc_kernel<1, 128>;           // double-buffered P, exact split (the default variant)
    PN_CUDA(cudaFuncSetAttribute(cols_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    cols_k<<<dim3(cdiv(N, 128), B), mstcb::NT, sm, (cudaStream_t)stream>>>(Yprev, X, ws_Gn, ws_gd, N, cinv, gX, accumulate_gX);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("ms_bwd_tc_kernel<cols>");
    return PN_OK;
}

extern "C" int pn_debug_set_progress(int* host_mapped_words) {
    PN_CUDA(cudaMemcpyToSymbol(pn::mstcb::g_dbg, &host_mapped_words, sizeof(int*)));
    return PN_OK;
}
