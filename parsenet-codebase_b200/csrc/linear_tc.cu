// Per-point linear layer forward on the 5th-gen tensor cores (tcgen05 + TMEM), split-TF32 (3 MMAs per product) for
// fp32-level accuracy.  Same contract as linear_fwd_kernel in linear.cu:
//   Y = act(norm(A)) . W^T + bias (+ per-shape bias),  epilogue accumulates sum / sum^2 of Y per (shape, group)
// Replaces Conv1d/Conv2d(k=1) + GroupNorm + ReLU chains, src/PointNet.py:157-165,194-196,274-284; src/model.py:155-176.
//
// One CTA = one 128 x 128 output tile, 9 warps, warp-specialised and mbarrier-pipelined:
//   warps 0-3 epilogue : thread = output row (TMEM lane).  accumulator TMEM -> regs, + bias, group statistics; each warp
//                        transposes its 32 x 32 chunk through shared memory so that stores are 4 rows x 128 B per STG.128
//   warps 4-7 loaders  : 4 lanes cover the 64 B a row contributes to a stage (a warp reads 8 rows x 64 B per LDG.128).
//                        A: producer norm + activation -> tf32 split (big by truncation, small = exact remainder);
//                        W: split ; both into the K-major no-swizzle core-matrix layout, 3 stages x 32.5 KB (16 k each)
//   warp 8     MMA     : one elected lane issues tcgen05.mma kind::tf32 M=128 N=128 K=8, per k-step
//                        As.Bb, Ab.Bs, Ab.Bb (small terms first) into 128 TMEM columns; tcgen05.commit frees the stage
// 2 CTAs per SM (96 KB smem, 128 TMEM columns each): one CTA's epilogue overlaps the other's main loop.
#include "common.cuh"
#include "tc05.cuh"

namespace pn {
namespace lintc {
using namespace tc05;

constexpr int BM = 128, BN = 128, BK = 16, NSTAGE = 3;
constexpr int EPI_THREADS = 128, LOAD_WARP0 = 4, LOAD_THREADS = 128, MMA_WARP = 8, NT = 288;
// K-major no-swizzle core-matrix layout [k/4][row/8][8][16 B].  The distance between k-chunks (LBO) is padded by 32 B:
// a loader quarter-warp stores 2 rows x 4 k-chunks per STS.128, which then lands in 8 distinct 16-byte bank groups.
constexpr uint32_t LBO = BM * 16 + 32, SBO = 128, TMEM_COLS = 128;
constexpr int OP_BYTES = (BK / 4) * LBO;              // one operand part of one stage
constexpr int STAGE_BYTES = 4 * OP_BYTES;             // A big, A small, W big, W small
constexpr int EPI_PITCH = 36;                         // floats per row of the per-warp 32 x 32 store-transpose tile
constexpr int MAXG = 4;                               // statistics groups per 128-column tile (channels/group >= 32)

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };
__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LRELU) return v > 0.f ? v : 0.2f * v;
    return v;
}

struct Args {
    const float* A; long long lda;      // [B*Np][K]
    const float* W; long long ldw;      // [Nout][K]
    const float* bias;                  // [Nout] or null
    const float* sbias;                 // [B][Nout] or null
    const float* in_scale; const float* in_shift; int in_act;     // [B][K] or null
    float* Y; long long ldy;            // [B*Np][Nout]
    double* stats;                      // [S][G][2] or null
    int B, Np, K, Nout, G, stats_per_shape;
    // ---- backward-data epilogue (BWD instantiation: A = dY, W = W^T of the layer, Y = dZ, Nout = input width of the layer)
    int acc;                            // dZ += instead of overwrite
    int fin;                            // multiply by act'(norm(fA)) and accumulate the norm-backward sums
    const float* fA; long long fld;     // pre-norm input of the layer [B*Np][Nout]
    const float* fsc; const float* fsh; int fact;      // producer scale / shift [B][Nout] (or null), activation
    const float* fgamma;                // [Nout] norm weight of the producer, or null: mask only
    const float* fmr;                   // [S][fG][2] mean / rstd of the producer
    double* fgsum;                      // [S][fG][2]
    int fG, fps;
};
__device__ __forceinline__ float act_grad(float pre, int act) {
    if (act == ACT_RELU) return pre > 0.f ? 1.f : 0.f;
    if (act == ACT_LRELU) return pre > 0.f ? 1.f : 0.2f;
    return 1.f;
}

struct Bars { uint64_t full[NSTAGE], empty[NSTAGE], acc_full; };

// grid (tiles_n, tiles_m * B): the column tiles of one row block are adjacent in launch order, so the A rows they share
// are read from DRAM once and re-used out of L2 (with the row block outermost ncu showed A re-read once per column tile)
template <bool BWD>
__global__ void __launch_bounds__(NT, 2) linear_tc_kernel(Args p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sc_s = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);      // [K] producer scale (or unused)
    float* sh_s = sc_s + p.K;                                                 // [K] producer shift
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[BN];
    __shared__ double gacc[MAXG][2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_m = (p.Np + BM - 1) / BM;
    const int b = blockIdx.y / tiles_m;
    const int m0 = (blockIdx.y % tiles_m) * BM;
    const int n0 = blockIdx.x * BN;
    const int nch = (p.K + BK - 1) / BK;

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars.full[s], LOAD_THREADS); mbar_init(&bars.empty[s], 1); }
        mbar_init(&bars.acc_full, 1);
        mbar_fence_init();
    }
    if (p.in_scale)
        for (int k = tid; k < p.K; k += NT) {
            sc_s[k] = p.in_scale[(long long)b * p.K + k];
            sh_s[k] = p.in_shift[(long long)b * p.K + k];
        }
    if (tid < BN) {
        const int n = n0 + tid;
        float v = 0.f;
        if (n < p.Nout) {
            if (p.bias) v += p.bias[n];
            if (p.sbias) v += p.sbias[(long long)b * p.Nout + n];
        }
        bias_s[tid] = v;
    }
    if (tid < MAXG * 2) gacc[tid >> 1][tid & 1] = 0.0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;

    if (warp < 4) {
        // =============================================================================== epilogue warps
        const int row = 32 * warp + lane;
        const uint32_t la = (uint32_t)(32 * warp) << 16;
        const int r = m0 + row;
        const bool ok = r < p.Np;
        const int cpg = p.stats ? p.Nout / p.G : 1;
        // per-warp store-transpose tile in the (by then idle) pipeline stages: [32 rows][EPI_PITCH]
        float* tr = reinterpret_cast<float*>(smem) + warp * 32 * EPI_PITCH;
        const int sub_r = lane >> 3, sub_c = 4 * (lane & 7);     // store phase: 4 rows x 8 lanes x 16 B per instruction
        mbar_wait(&bars.acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= p.Nout) break;                       // Nout % 32 == 0 (checked on the host)
            uint32_t v[32];
            tmem_ld32(tb + la + c0, v);
            tmem_ld_wait();
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
                float4 o;
                o.x = __uint_as_float(v[e]) + bias_s[c0 + e];
                o.y = __uint_as_float(v[e + 1]) + bias_s[c0 + e + 1];
                o.z = __uint_as_float(v[e + 2]) + bias_s[c0 + e + 2];
                o.w = __uint_as_float(v[e + 3]) + bias_s[c0 + e + 3];
                *reinterpret_cast<float4*>(tr + lane * EPI_PITCH + e) = o;
                if (ok) {
                    s += (o.x + o.y) + (o.z + o.w);
                    q = fmaf(o.x, o.x, q); q = fmaf(o.y, o.y, q); q = fmaf(o.z, o.z, q); q = fmaf(o.w, o.w, q);
                }
            }
            __syncwarp();
            [[maybe_unused]] float f1 = 0.f, f2 = 0.f;
            [[maybe_unused]] float fmean = 0.f, frstd = 0.f;
            if constexpr (BWD) {
                if (p.fin && p.fgamma) {
                    const int g = (n0 + c0) / (p.Nout / p.fG);
                    const float* mr = p.fmr + ((long long)(p.fps ? b : 0) * p.fG + g) * 2;
                    fmean = mr[0]; frstd = mr[1];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = sub_r + 4 * i;                    // row inside this warp's 32
                const int gr = m0 + 32 * warp + rr;
                if (gr < p.Np) {
                    float4 o = *reinterpret_cast<const float4*>(tr + rr * EPI_PITCH + sub_c);
                    const long long row = (long long)b * p.Np + gr;
                    const int col = n0 + c0 + sub_c;
                    float* dst = p.Y + row * p.ldy + col;
                    if constexpr (BWD) {
                        if (p.acc) {
                            const float4 old = *reinterpret_cast<const float4*>(dst);
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        if (p.fin) {
                            const float4 a = *reinterpret_cast<const float4*>(p.fA + row * p.fld + col);
                            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.fsc) {
                                sc = *reinterpret_cast<const float4*>(p.fsc + (long long)b * p.Nout + col);
                                sh = *reinterpret_cast<const float4*>(p.fsh + (long long)b * p.Nout + col);
                            }
                            const float av[4] = {a.x, a.y, a.z, a.w};
                            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
                            float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float pre = p.fsc ? fmaf(av[e], scv[e], shv[e]) : av[e];
                                const float t = ov[e] * act_grad(pre, p.fact);
                                if (p.fgamma) {
                                    const float xh = (av[e] - fmean) * frstd;
                                    const float gt = p.fgamma[col + e] * t;
                                    f1 += gt;
                                    f2 = fmaf(gt, xh, f2);
                                }
                                ov[e] = t;
                            }
                            o = make_float4(ov[0], ov[1], ov[2], ov[3]);
                        }
                    }
                    *reinterpret_cast<float4*>(dst) = o;
                }
            }
            __syncwarp();
            if constexpr (BWD) {
                if (p.fin && p.fgamma) {
                    const double d1 = warp_sum((double)f1), d2 = warp_sum((double)f2);
                    if (lane == 0) {
                        const int cpgf = p.Nout / p.fG;
                        const int gl = (n0 + c0) / cpgf - n0 / cpgf;
                        atomicAdd(&gacc[gl][0], d1);
                        atomicAdd(&gacc[gl][1], d2);
                    }
                }
            }
            if (p.stats) {
                double ds = warp_sum((double)s), dq = warp_sum((double)q);
                if (lane == 0) {
                    const int gl = (n0 + c0) / cpg - n0 / cpg;       // group of this 32-column chunk, local to the tile
                    atomicAdd(&gacc[gl][0], ds);
                    atomicAdd(&gacc[gl][1], dq);
                }
            }
        }
        tc_fence_before();
        if constexpr (BWD) {
            if (p.fin && p.fgamma) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid < MAXG * 2) {
                    const int cpgf = p.Nout / p.fG;
                    const int gl = tid >> 1;
                    const int g = n0 / cpgf + gl;
                    if (g < p.fG && (long long)g * cpgf < (long long)min(p.Nout, n0 + BN) && gacc[gl][tid & 1] != 0.0)
                        atomicAdd(&p.fgsum[((long long)(p.fps ? b : 0) * p.fG + g) * 2 + (tid & 1)], gacc[gl][tid & 1]);
                }
            }
        }
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid < MAXG * 2) {
                const int gl = tid >> 1;
                const int g = n0 / cpg + gl;
                if (g < p.G && (long long)g * cpg < (long long)min(p.Nout, n0 + BN) && gacc[gl][tid & 1] != 0.0)
                    atomicAdd(&p.stats[((long long)(p.stats_per_shape ? b : 0) * p.G + g) * 2 + (tid & 1)],
                              gacc[gl][tid & 1]);
            }
        }
    } else if (warp < MMA_WARP) {
        // =============================================================================== loader warps
        const int lt = tid - LOAD_WARP0 * 32;          // 0..127
        const int c4 = lt & 3;                         // 16-byte k-chunk of the stage handled by this thread
        const int rb = lt >> 2;                        // rows rb, rb+32, rb+64, rb+96 of both operands
        const float* abase = p.A + (long long)b * p.Np * p.lda;
        const bool has_norm = p.in_scale != nullptr;
        const int act = p.in_act;
        float4 va[4], vw[4], na[4], nw[4];
        auto fetch = [&](int kc, float4 (&xa)[4], float4 (&xw)[4]) {
            const int k = kc * BK + 4 * c4;
            const bool kin = (kc < nch) && (k < p.K);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ra = m0 + rb + 32 * i, rw = n0 + rb + 32 * i;
                xa[i] = (kin && ra < p.Np) ? *reinterpret_cast<const float4*>(abase + (long long)ra * p.lda + k)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                xw[i] = (kin && rw < p.Nout) ? *reinterpret_cast<const float4*>(p.W + (long long)rw * p.ldw + k)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        fetch(0, va, vw);
#pragma unroll 1
        for (int kc = 0; kc < nch; ++kc) {
            const int s = kc % NSTAGE;
            fetch(kc + 1, na, nw);
            mbar_wait(&bars.empty[s], ((kc / NSTAGE) & 1) ^ 1);
            unsigned char* st = smem + s * STAGE_BYTES;
            const int k = kc * BK + 4 * c4;
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_norm && k < p.K) {
                sc = *reinterpret_cast<const float4*>(sc_s + k);
                sh = *reinterpret_cast<const float4*>(sh_s + k);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = rb + 32 * i;
                float f0 = va[i].x, f1 = va[i].y, f2 = va[i].z, f3 = va[i].w;
                if ((m0 + row) < p.Np && k < p.K) {      // (padding rows / columns must stay exactly zero)
                    if (has_norm) {
                        f0 = fmaf(f0, sc.x, sh.x); f1 = fmaf(f1, sc.y, sh.y);
                        f2 = fmaf(f2, sc.z, sh.z); f3 = fmaf(f3, sc.w, sh.w);
                    }
                    f0 = act_fwd(f0, act); f1 = act_fwd(f1, act); f2 = act_fwd(f2, act); f3 = act_fwd(f3, act);
                }
                const uint32_t o = (uint32_t)(c4 * LBO + (row >> 3) * 128 + (row & 7) * 16);
                {
                    const float b0 = tf32_hi(f0), b1 = tf32_hi(f1), b2 = tf32_hi(f2), b3 = tf32_hi(f3);
                    *reinterpret_cast<float4*>(st + o) = make_float4(b0, b1, b2, b3);
                    *reinterpret_cast<float4*>(st + OP_BYTES + o) = make_float4(f0 - b0, f1 - b1, f2 - b2, f3 - b3);
                }
                {
                    const float w0 = vw[i].x, w1 = vw[i].y, w2 = vw[i].z, w3 = vw[i].w;
                    const float b0 = tf32_hi(w0), b1 = tf32_hi(w1), b2 = tf32_hi(w2), b3 = tf32_hi(w3);
                    *reinterpret_cast<float4*>(st + 2 * OP_BYTES + o) = make_float4(b0, b1, b2, b3);
                    *reinterpret_cast<float4*>(st + 3 * OP_BYTES + o) = make_float4(w0 - b0, w1 - b1, w2 - b2, w3 - b3);
                }
            }
            fence_async_smem();
            mbar_arrive(&bars.full[s]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { va[i] = na[i]; vw[i] = nw[i]; }
        }
    } else {
        // =============================================================================== MMA warp
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc(2, BM, BN, 0, 0);
        const uint32_t sbase = smem_u32(smem);
#pragma unroll 1
        for (int kc = 0; kc < nch; ++kc) {
            const int s = kc % NSTAGE;
            mbar_wait(&bars.full[s], (kc / NSTAGE) & 1);
            tc_fence_after();
            const uint32_t st = sbase + s * STAGE_BYTES;
            const uint64_t dab = make_smem_desc(st, LBO, SBO, 0);
            const uint64_t das = make_smem_desc(st + OP_BYTES, LBO, SBO, 0);
            const uint64_t dwb = make_smem_desc(st + 2 * OP_BYTES, LBO, SBO, 0);
            const uint64_t dws = make_smem_desc(st + 3 * OP_BYTES, LBO, SBO, 0);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * ((2 * LBO) >> 4));
                    mma_tf32_ss(tb, das + adv, dwb + adv, idesc, (kc | ks) ? 1u : 0u);
                    mma_tf32_ss(tb, dab + adv, dws + adv, idesc, 1u);
                    mma_tf32_ss(tb, dab + adv, dwb + adv, idesc, 1u);
                }
                mma_commit(&bars.empty[s]);
            }
            __syncwarp();
        }
        if (leader) mma_commit(&bars.acc_full);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tb, TMEM_COLS);
}

// ================================================================================================ backward: weight
// dW[n][k] += sum_m dY[m][n] * act(norm(A))[m][k] on the tensor cores: D[128 n][128 k] += (dY tile)^T (X tile), contraction
// over the rows m.  Both operands are row-major in HBM with the contraction index OUTERMOST, while tcgen05 wants it
// innermost (K-major; MN-major tf32 operands are not executed, profiles/r01_tc_probe.md): the loader warps transpose 4 x 4
// blocks in registers on the way from LDG.128 (4 consecutive n / k of one row) to STS.128 (4 consecutive m of one n / k).
// 8-row groups are 144 B apart (SBO) so that a quarter-warp's eight 16-byte stores fall into eight different bank groups.
// The rows of a shape are split over gridDim.z; partial tiles are accumulated with fp32 atomics like the FP32-pipe kernel.
constexpr uint32_t W_SBO = 144, W_LBO = (BM / 8) * W_SBO;               // 2304 B per 4-row (m) chunk
constexpr int W_OP_BYTES = (BK / 4) * W_LBO;                           // 9216
constexpr int W_STAGE_BYTES = 4 * W_OP_BYTES;                          // dY big, dY small, X big, X small
constexpr int W_NSTAGE = 2;                                            // 2 x 36 KB: two CTAs per SM (4 stages in flight per SM)

struct WArgs {
    const float* dY; long long lddy;    // [B*Np][Nout]
    const float* A; long long lda;      // [B*Np][K]
    const float* in_scale; const float* in_shift; int in_act;     // [B][K] or null
    float* dW; long long lddw;          // [Nout][K]
    float* db;                          // [Nout] or null
    float* dsb;                         // [B][Nout] or null
    int B, Np, K, Nout, rows_per_split, splits_per_shape;
};

__global__ void __launch_bounds__(NT, 2) linear_bwd_weight_tc_kernel(WArgs p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ float colsum_s[4][BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * BM;                 // dW rows
    const int k0 = blockIdx.y * BN;                 // dW columns
    const int b = blockIdx.z / p.splits_per_shape;
    const int r_begin = (blockIdx.z % p.splits_per_shape) * p.rows_per_split;
    const int r_end = min(p.Np, r_begin + p.rows_per_split);
    const int nch = (r_end - r_begin + BK - 1) / BK;          // >= 1 (host guarantees r_begin < Np)

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < W_NSTAGE; ++s) { mbar_init(&bars.full[s], LOAD_THREADS); mbar_init(&bars.empty[s], 1); }
        mbar_init(&bars.acc_full, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;

    if (warp < 4) {
        // =============================================================================== epilogue warps
        const uint32_t la = (uint32_t)(32 * warp) << 16;
        float* tr = reinterpret_cast<float*>(smem) + warp * 32 * EPI_PITCH;
        const int sub_r = lane >> 3, sub_c = 4 * (lane & 7);
        mbar_wait(&bars.acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (k0 + c0 >= p.K) break;
            uint32_t v[32];
            tmem_ld32(tb + la + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 4)
                *reinterpret_cast<float4*>(tr + lane * EPI_PITCH + e) =
                    make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = sub_r + 4 * i;
                const int n = n0 + 32 * warp + rr;
                if (n < p.Nout) {
                    const float4 o = *reinterpret_cast<const float4*>(tr + rr * EPI_PITCH + sub_c);
                    float* dst = p.dW + (long long)n * p.lddw + k0 + c0 + sub_c;
                    const int kk = k0 + c0 + sub_c;
                    if (kk < p.K) atomicAdd(dst, o.x);
                    if (kk + 1 < p.K) atomicAdd(dst + 1, o.y);
                    if (kk + 2 < p.K) atomicAdd(dst + 2, o.z);
                    if (kk + 3 < p.K) atomicAdd(dst + 3, o.w);
                }
            }
            __syncwarp();
        }
        tc_fence_before();
    } else if (warp < MMA_WARP) {
        // =============================================================================== loader warps (transposing)
        const int lt = tid - LOAD_WARP0 * 32;          // 0..127
        const int c4 = lt & 31;                        // columns 4 c4 .. 4 c4 + 3 of both tiles (n of dY, k of A)
        const int mq = lt >> 5;                        // rows 4 mq .. 4 mq + 3 of the 16-row stage = k-chunk index of the stage
        const float* dYb = p.dY + (long long)b * p.Np * p.lddy;
        const float* Ab = p.A + (long long)b * p.Np * p.lda;
        const int nn = n0 + 4 * c4, kk = k0 + 4 * c4;
        const bool n_ok = nn < p.Nout, k_ok = kk < p.K;       // (Nout % 4 == 0 and K % 4 == 0: a 4-group is inside or outside)
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool has_norm = p.in_scale != nullptr;
        if (has_norm && k_ok) {
            sc = *reinterpret_cast<const float4*>(p.in_scale + (long long)b * p.K + kk);
            sh = *reinterpret_cast<const float4*>(p.in_shift + (long long)b * p.K + kk);
        }
        const int act = p.in_act;
        const bool want_cs = (p.db || p.dsb) && blockIdx.y == 0;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
        float4 vy[4], va[4], ny[4], na[4];
        auto fetch = [&](int kc, float4 (&xy)[4], float4 (&xa)[4]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r_begin + kc * BK + 4 * mq + i;
                const bool rin = (kc < nch) && (r < r_end);
                xy[i] = (rin && n_ok) ? *reinterpret_cast<const float4*>(dYb + (long long)r * p.lddy + nn)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                xa[i] = (rin && k_ok) ? *reinterpret_cast<const float4*>(Ab + (long long)r * p.lda + kk)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        fetch(0, vy, va);
#pragma unroll 1
        for (int kc = 0; kc < nch; ++kc) {
            const int s = kc % W_NSTAGE;
            fetch(kc + 1, ny, na);
            mbar_wait(&bars.empty[s], ((kc / W_NSTAGE) & 1) ^ 1);
            unsigned char* st = smem + s * W_STAGE_BYTES;
            // X = act(norm(A)); rows beyond the split stay exactly zero
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r_begin + kc * BK + 4 * mq + i;
                if (r < r_end && k_ok) {
                    float f0 = va[i].x, f1 = va[i].y, f2 = va[i].z, f3 = va[i].w;
                    if (has_norm) {
                        f0 = fmaf(f0, sc.x, sh.x); f1 = fmaf(f1, sc.y, sh.y); f2 = fmaf(f2, sc.z, sh.z); f3 = fmaf(f3, sc.w, sh.w);
                    }
                    va[i] = make_float4(act_fwd(f0, act), act_fwd(f1, act), act_fwd(f2, act), act_fwd(f3, act));
                }
            }
            if (want_cs) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { cs[0] += vy[i].x; cs[1] += vy[i].y; cs[2] += vy[i].z; cs[3] += vy[i].w; }
            }
            // 4 x 4 register transpose: element e of the four rows -> one 16-byte chunk (4 consecutive m) of row (4 c4 + e)
            const float ye[4][4] = {{vy[0].x, vy[1].x, vy[2].x, vy[3].x}, {vy[0].y, vy[1].y, vy[2].y, vy[3].y},
                                    {vy[0].z, vy[1].z, vy[2].z, vy[3].z}, {vy[0].w, vy[1].w, vy[2].w, vy[3].w}};
            const float ae[4][4] = {{va[0].x, va[1].x, va[2].x, va[3].x}, {va[0].y, va[1].y, va[2].y, va[3].y},
                                    {va[0].z, va[1].z, va[2].z, va[3].z}, {va[0].w, va[1].w, va[2].w, va[3].w}};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = 4 * c4 + e;
                const uint32_t o = (uint32_t)(mq * W_LBO + (row >> 3) * W_SBO + (row & 7) * 16);
                {
                    const float b0 = tf32_hi(ye[e][0]), b1 = tf32_hi(ye[e][1]), b2 = tf32_hi(ye[e][2]), b3 = tf32_hi(ye[e][3]);
                    *reinterpret_cast<float4*>(st + o) = make_float4(b0, b1, b2, b3);
                    *reinterpret_cast<float4*>(st + W_OP_BYTES + o) =
                        make_float4(ye[e][0] - b0, ye[e][1] - b1, ye[e][2] - b2, ye[e][3] - b3);
                }
                {
                    const float b0 = tf32_hi(ae[e][0]), b1 = tf32_hi(ae[e][1]), b2 = tf32_hi(ae[e][2]), b3 = tf32_hi(ae[e][3]);
                    *reinterpret_cast<float4*>(st + 2 * W_OP_BYTES + o) = make_float4(b0, b1, b2, b3);
                    *reinterpret_cast<float4*>(st + 3 * W_OP_BYTES + o) =
                        make_float4(ae[e][0] - b0, ae[e][1] - b1, ae[e][2] - b2, ae[e][3] - b3);
                }
            }
            fence_async_smem();
            mbar_arrive(&bars.full[s]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { vy[i] = ny[i]; va[i] = na[i]; }
        }
        if (want_cs) {
#pragma unroll
            for (int e = 0; e < 4; ++e) colsum_s[mq][4 * c4 + e] = cs[e];
            asm volatile("bar.sync 2, 128;" ::: "memory");
            const int n = n0 + lt;
            if (n < p.Nout) {
                const float t = (colsum_s[0][lt] + colsum_s[1][lt]) + (colsum_s[2][lt] + colsum_s[3][lt]);
                if (p.db) atomicAdd(&p.db[n], t);
                if (p.dsb) atomicAdd(&p.dsb[(long long)b * p.Nout + n], t);
            }
        }
    } else {
        // =============================================================================== MMA warp
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc(2, BM, BN, 0, 0);
        const uint32_t sbase = smem_u32(smem);
#pragma unroll 1
        for (int kc = 0; kc < nch; ++kc) {
            const int s = kc % W_NSTAGE;
            mbar_wait(&bars.full[s], (kc / W_NSTAGE) & 1);
            tc_fence_after();
            const uint32_t st = sbase + s * W_STAGE_BYTES;
            const uint64_t dyb = make_smem_desc(st, W_LBO, W_SBO, 0);
            const uint64_t dys = make_smem_desc(st + W_OP_BYTES, W_LBO, W_SBO, 0);
            const uint64_t dab = make_smem_desc(st + 2 * W_OP_BYTES, W_LBO, W_SBO, 0);
            const uint64_t das = make_smem_desc(st + 3 * W_OP_BYTES, W_LBO, W_SBO, 0);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * ((2 * W_LBO) >> 4));
                    mma_tf32_ss(tb, dys + adv, dab + adv, idesc, (kc | ks) ? 1u : 0u);
                    mma_tf32_ss(tb, dyb + adv, das + adv, idesc, 1u);
                    mma_tf32_ss(tb, dyb + adv, dab + adv, idesc, 1u);
                }
                mma_commit(&bars.empty[s]);
            }
            __syncwarp();
        }
        if (leader) mma_commit(&bars.acc_full);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tb, TMEM_COLS);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace lintc
}  // namespace pn

using namespace pn;

// 1 if pn_linear_fwd_tc accepts this problem (the caller uses pn_linear_fwd otherwise)
extern "C" int pn_linear_fwd_tc_supported(const float* A, long long lda, const float* W, long long ldw, const float* Y,
                                          long long ldy, int Np, int K, int Nout, int G, int has_stats) {
    using namespace lintc;
    if (!aligned16(A) || !aligned16(W) || !aligned16(Y) || lda % 4 || ldw % 4 || ldy % 4) return 0;
    if (K % 4 || K < 16 || K > 4096 || Nout % 32 || Nout < 32 || Np < 1) return 0;
    if (has_stats && (G <= 0 || Nout % G || (Nout / G) % 32)) return 0;
    return 1;
}

extern "C" int pn_linear_fwd_tc(const float* A, long long lda, const float* W, long long ldw, const float* bias,
                                const float* sbias, const float* in_scale, const float* in_shift, int in_act, float* Y,
                                long long ldy, double* stats, int B, int Np, int K, int Nout, int G,
                                int stats_per_shape, void* stream) {
    using namespace lintc;
    PN_REQUIRE(A && W && Y, "pn_linear_fwd_tc: null pointer");
    PN_REQUIRE(B > 0 && Np > 0, "pn_linear_fwd_tc: bad shape");
    PN_REQUIRE(pn_linear_fwd_tc_supported(A, lda, W, ldw, Y, ldy, Np, K, Nout, G, stats != nullptr),
               "pn_linear_fwd_tc: unsupported problem (K=%d Nout=%d G=%d; need 16-byte aligned rows, K %% 4 == 0, "
               "Nout %% 32 == 0, channels/group %% 32 == 0)", K, Nout, G);
    Args p{};
    p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.bias = bias; p.sbias = sbias; p.in_scale = in_scale; p.in_shift = in_shift;
    p.in_act = in_act; p.Y = Y; p.ldy = ldy; p.stats = stats; p.B = B; p.Np = Np; p.K = K; p.Nout = Nout; p.G = G;
    p.stats_per_shape = stats_per_shape;
    size_t sm = (size_t)NSTAGE * STAGE_BYTES + 2 * (size_t)K * sizeof(float) + 1024;
    PN_CUDA(cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_REQUIRE((long long)cdiv(Np, BM) * B <= 65535, "pn_linear_fwd_tc: too many row tiles (%d x %d)", cdiv(Np, BM), B);
    dim3 grid(cdiv(Nout, BN), cdiv(Np, BM) * B);
    linear_tc_kernel<false><<<grid, NT, sm, (cudaStream_t)stream>>>(p);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("linear_tc_kernel<fwd>");
    return PN_OK;
}

// Backward w.r.t. the layer input on the tensor cores: dZ[m][k] (+)= sum_n dY[m][n] W[n][k], same epilogue contract as
// pn_linear_bwd_data (activation mask of the layer input + norm-backward sums).  It is the forward GEMM with A = dY and the
// TRANSPOSED weight Wt [K][Nout] as the row-major "weight" operand (contraction over Nout), plus the finalize epilogue.
extern "C" int pn_linear_bwd_data_tc_supported(const float* dY, long long lddy, const float* Wt, long long ldwt, const float* dZ,
                                               long long lddz, const float* A, long long lda, int Np, int K, int Nout, int G,
                                               int has_gamma) {
    using namespace lintc;
    if (!aligned16(dY) || !aligned16(Wt) || !aligned16(dZ) || lddy % 4 || ldwt % 4 || lddz % 4) return 0;
    if (A && (!aligned16(A) || lda % 4)) return 0;
    if (Nout % 4 || Nout < 16 || Nout > 4096 || K % 32 || K < 32 || Np < 1) return 0;
    if (has_gamma && (G <= 0 || K % G || (K / G) % 32)) return 0;
    return 1;
}

extern "C" int pn_linear_bwd_data_tc(const float* dY, long long lddy, const float* Wt, long long ldwt, float* dZ, long long lddz,
                                     int accumulate, int finalize, const float* A, long long lda, const float* in_scale,
                                     const float* in_shift, int in_act, const float* gamma, const float* mean_rstd,
                                     double* gsum, int B, int Np, int K, int Nout, int G, int stats_per_shape, void* stream) {
    using namespace lintc;
    PN_REQUIRE(dY && Wt && dZ, "pn_linear_bwd_data_tc: null pointer");
    PN_REQUIRE(!finalize || A, "pn_linear_bwd_data_tc: finalize needs A");
    PN_REQUIRE(!(finalize && gamma) || (mean_rstd && gsum), "pn_linear_bwd_data_tc: norm args");
    PN_REQUIRE(B > 0 && Np > 0, "pn_linear_bwd_data_tc: bad shape");
    PN_REQUIRE(pn_linear_bwd_data_tc_supported(dY, lddy, Wt, ldwt, dZ, lddz, finalize ? A : nullptr, lda, Np, K, Nout, G,
                                               finalize && gamma),
               "pn_linear_bwd_data_tc: unsupported problem (K=%d Nout=%d G=%d)", K, Nout, G);
    Args p{};
    p.A = dY; p.lda = lddy; p.W = Wt; p.ldw = ldwt; p.Y = dZ; p.ldy = lddz; p.B = B; p.Np = Np;
    p.K = Nout;                  // contraction length of the GEMM
    p.Nout = K;                  // output columns = input width of the layer
    p.G = 1; p.stats_per_shape = 1;
    p.acc = accumulate; p.fin = finalize; p.fA = A; p.fld = lda; p.fsc = in_scale; p.fsh = in_shift; p.fact = in_act;
    p.fgamma = finalize ? gamma : nullptr; p.fmr = mean_rstd; p.fgsum = gsum; p.fG = G > 0 ? G : 1; p.fps = stats_per_shape;
    size_t sm = (size_t)NSTAGE * STAGE_BYTES + 2 * (size_t)p.K * sizeof(float) + 1024;
    PN_CUDA(cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_REQUIRE((long long)cdiv(Np, BM) * B <= 65535, "pn_linear_bwd_data_tc: too many row tiles (%d x %d)", cdiv(Np, BM), B);
    dim3 grid(cdiv(p.Nout, BN), cdiv(Np, BM) * B);
    linear_tc_kernel<true><<<grid, NT, sm, (cudaStream_t)stream>>>(p);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("linear_tc_kernel<bwd_data>");
    return PN_OK;
}


// dW += dY^T act(norm(A)) (+ bias gradients) on the tensor cores; same contract as pn_linear_bwd_weight
extern "C" int pn_linear_bwd_weight_tc_supported(const float* dY, long long lddy, const float* A, long long lda, int Np, int K,
                                                 int Nout) {
    using namespace lintc;
    if (!aligned16(dY) || !aligned16(A) || lddy % 4 || lda % 4) return 0;
    if (K % 4 || Nout % 4 || K < 32 || Nout < 32 || Np < 64) return 0;
    return 1;
}

extern "C" int pn_linear_bwd_weight_tc(const float* dY, long long lddy, const float* A, long long lda, const float* in_scale,
                                       const float* in_shift, int in_act, float* dW, long long lddw, float* db, float* dsb,
                                       int B, int Np, int K, int Nout, void* stream) {
    using namespace lintc;
    PN_REQUIRE(dY && A && dW, "pn_linear_bwd_weight_tc: null pointer");
    PN_REQUIRE(B > 0 && pn_linear_bwd_weight_tc_supported(dY, lddy, A, lda, Np, K, Nout),
               "pn_linear_bwd_weight_tc: unsupported problem (K=%d Nout=%d Np=%d)", K, Nout, Np);
    PN_REQUIRE(!in_scale || (aligned16(in_scale) && aligned16(in_shift)), "pn_linear_bwd_weight_tc: scale / shift alignment");
    WArgs p{dY, lddy, A, lda, in_scale, in_shift, in_act, dW, lddw, db, dsb, B, Np, K, Nout, 0, 0};
    // aim for ~2 waves of 296 CTAs; a split never straddles two shapes and is a multiple of the 16-row stage
    const int tiles = cdiv(Nout, BM) * cdiv(K, BN);
    const int want = max(1, (592 + tiles * B - 1) / (tiles * B));
    int rows = cdiv(Np, want);
    rows = max(BK * 8, ((rows + BK - 1) / BK) * BK);
    p.rows_per_split = rows;
    p.splits_per_shape = cdiv(Np, rows);
    const size_t sm = (size_t)W_NSTAGE * W_STAGE_BYTES + 1024;
    PN_CUDA(cudaFuncSetAttribute(linear_bwd_weight_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    PN_REQUIRE((long long)p.splits_per_shape * B <= 65535, "pn_linear_bwd_weight_tc: too many row splits");
    dim3 grid(cdiv(Nout, BM), cdiv(K, BN), p.splits_per_shape * B);
    linear_bwd_weight_tc_kernel<<<grid, NT, sm, (cudaStream_t)stream>>>(p);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("linear_bwd_weight_tc_kernel");
    return PN_OK;
}
