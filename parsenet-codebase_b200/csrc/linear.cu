// Per-point linear layers (Conv1d / Conv2d with kernel 1) on point-major activations, fp32.
//
// Replaces the reference's Conv1d/Conv2d(k=1)+GroupNorm/BatchNorm+ReLU chains
//   src/PointNet.py:157-165,194-196,274-284 (encoder mlp1 + segmentation head), src/model.py:74-99,155-176
// with three GEMM kernels that keep only PRE-norm activations in HBM:
//   forward      Y = act(norm(A)) . W^T + bias (+ per-shape bias)   [norm/act of the producer applied on operand load]
//                epilogue accumulates sum / sum^2 of Y per (shape, group) for Y's own normalisation
//   bwd-data     dZ = dY . W            epilogue: x activation mask of A, accumulates the two GroupNorm-backward sums
//   bwd-weight   dW += dY^T . act(norm(A)), db += colsum(dY)     (split over rows, fp32 atomics)
// The normalised/activated tensors are never written.  Activations are (rows = B*Np points, channels contiguous).
//
// v1 math: FP32 FMA pipe, 128x128x16 tiles, 256 threads x (8x8) register tile, double-buffered shared memory.
// (fp32-exact products; the tcgen05 3xTF32 path replaces the main loop, the epilogues stay.)
#include "common.cuh"
#include <stdlib.h>

namespace pn {
namespace lin {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PADM = BM + 4;   // smem row pitch (floats), keeps float4 alignment

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LRELU) return v > 0.f ? v : 0.2f * v;
    return v;
}
__device__ __forceinline__ float act_grad(float pre, int act) {
    if (act == ACT_RELU) return pre > 0.f ? 1.f : 0.f;
    if (act == ACT_LRELU) return pre > 0.f ? 1.f : 0.2f;
    return 1.f;
}

struct Tile {
    float a[2][BK][PADM];
    float b[2][BK][PADM];
};

// 8x8 micro-kernel on one BK slab
__device__ __forceinline__ void mma_slab(const float (*As)[PADM], const float (*Bs)[PADM], int ty, int tx,
                                         float (&acc)[8][8]) {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[kk][4 * ty]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + 4 * ty]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][4 * tx]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + 4 * tx]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// ---- operand loaders -------------------------------------------------------------------------------------------
// "K-contiguous": element (r, k) at p[r*ld + k]; fetches a [128 rows][BK] slab into registers (8 floats / thread)
// thread t: rows r = t/4 and 64 + t/4, k-offset 4*(t%4)
struct KcFrag { float v[2][4]; };

template <typename F>
__device__ __forceinline__ void load_kc(KcFrag& f, const float* __restrict__ p, long long ld, int r0, int rmax,
                                        int k0, int K, bool vec, F&& xform) {
    const int t = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int r = r0 + (t >> 2) + 64 * h;
        int k = k0 + 4 * (t & 3);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < rmax) {
            const float* q = p + (long long)r * ld + k;
            if (vec && k + 3 < K) {
                float4 w = *reinterpret_cast<const float4*>(q);
                v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (k + e < K) v[e] = q[e];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (k + e < K) ? xform(v[e], r, k + e) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) f.v[h][e] = v[e];
    }
}
__device__ __forceinline__ void store_kc(const KcFrag& f, float (*S)[PADM]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 4; ++e) S[4 * (t & 3) + e][(t >> 2) + 64 * h] = f.v[h][e];
}

// "row-contiguous": element (k, c) at p[k*ld + c]; fetches a [BK][128 cols] slab; thread t: k = t/32 (+8), cols 4*(t%32)
struct RcFrag { float v[2][4]; };

template <typename F>
__device__ __forceinline__ void load_rc(RcFrag& f, const float* __restrict__ p, long long ld, int k0, int K,
                                        int c0, int cmax, bool vec, F&& xform) {
    const int t = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int k = k0 + (t >> 5) + 8 * h;
        int c = c0 + 4 * (t & 31);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < K) {
            const float* q = p + (long long)k * ld + c;
            if (vec && c + 3 < cmax) {
                float4 w = *reinterpret_cast<const float4*>(q);
                v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (c + e < cmax) v[e] = q[e];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (c + e < cmax) ? xform(v[e], k, c + e) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) f.v[h][e] = v[e];
    }
}
__device__ __forceinline__ void store_rc(const RcFrag& f, float (*S)[PADM]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h)
        *reinterpret_cast<float4*>(&S[(t >> 5) + 8 * h][4 * (t & 31)]) =
            make_float4(f.v[h][0], f.v[h][1], f.v[h][2], f.v[h][3]);
}

struct Ident {
    __device__ __forceinline__ float operator()(float v, int, int) const { return v; }
};

// producer normalisation + activation applied while loading A:  act(v * scale[b][k] + shift[b][k])
struct NormAct {
    const float* scale;   // [B][K] or nullptr
    const float* shift;
    int K, act;
    long long shape_off;  // b * K
    __device__ __forceinline__ float operator()(float v, int, int k) const {
        if (scale) v = fmaf(v, scale[shape_off + k], shift[shape_off + k]);
        return act_fwd(v, act);
    }
};

// ================================================================================================ forward
// grid: (tiles_n, tiles_m_per_shape * B): column tiles of one row block are adjacent in launch order (A rows re-used out of
// L2 instead of re-read from DRAM once per column tile).  Rows of one CTA never straddle two shapes.
struct FwdArgs {
    const float* A; long long lda;     // [B*Np][K]
    const float* W; long long ldw;     // [Nout][K]
    const float* bias;                 // [Nout] or null
    const float* sbias;                // [B][Nout] or null (per-shape bias)
    const float* in_scale; const float* in_shift; int in_act;   // producer norm (per shape, per K channel) + activation
    float* Y; long long ldy;           // [B*Np][Nout]
    double* stats;                     // [S][G][2] sum,sumsq (S = B if stats_per_shape else 1) or null
    int B, Np, K, Nout, G, stats_per_shape;
    int vecA, vecW, vecY;
};

__global__ void __launch_bounds__(NT, 2) linear_fwd_kernel(FwdArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tile& S = *reinterpret_cast<Tile*>(smem_raw);
    __shared__ float csum[BN], csq[BN];
    const int tiles_m = (p.Np + BM - 1) / BM;
    const int b = blockIdx.y / tiles_m;
    const int m0 = (blockIdx.y % tiles_m) * BM;            // row inside the shape
    const int n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const float* Ab = p.A + (long long)b * p.Np * p.lda;
    NormAct xf{p.in_scale, p.in_shift, p.K, p.in_act, (long long)b * p.K};
    Ident id;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    KcFrag fa, fb;
    load_kc(fa, Ab, p.lda, m0, p.Np, 0, p.K, p.vecA, xf);
    load_kc(fb, p.W, p.ldw, n0, p.Nout, 0, p.K, p.vecW, id);
    store_kc(fa, S.a[0]); store_kc(fb, S.b[0]);
    __syncthreads();
    const int nk = (p.K + BK - 1) / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_kc(fa, Ab, p.lda, m0, p.Np, (kt + 1) * BK, p.K, p.vecA, xf);
            load_kc(fb, p.W, p.ldw, n0, p.Nout, (kt + 1) * BK, p.K, p.vecW, id);
        }
        mma_slab(S.a[cur], S.b[cur], ty, tx, acc);
        if (kt + 1 < nk) {
            store_kc(fa, S.a[cur ^ 1]); store_kc(fb, S.b[cur ^ 1]);
        }
        __syncthreads();
    }
    // ---- epilogue
    if (p.stats) {
        if (tid < BN) { csum[tid] = 0.f; csq[tid] = 0.f; }
        __syncthreads();
    }
    float ps[8], pq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ps[j] = 0.f; pq[j] = 0.f; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int r = m0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + i - 4);
        if (r >= p.Np) continue;
        float* yrow = p.Y + ((long long)b * p.Np + r) * p.ldy;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int c = n0 + 64 * h + 4 * tx;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int cc = c + e;
                float t = acc[i][4 * h + e];
                if (cc < p.Nout) {
                    if (p.bias) t += p.bias[cc];
                    if (p.sbias) t += p.sbias[(long long)b * p.Nout + cc];
                    ps[4 * h + e] += t;
                    pq[4 * h + e] = fmaf(t, t, pq[4 * h + e]);
                }
                v[e] = t;
            }
            if (p.vecY && c + 3 < p.Nout) {
                *reinterpret_cast<float4*>(yrow + c) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (c + e < p.Nout) yrow[c + e] = v[e];
            }
        }
    }
    if (p.stats) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int cl = (j < 4 ? 0 : 60) + 4 * tx + j;
            atomicAdd(&csum[cl], ps[j]);
            atomicAdd(&csq[cl], pq[j]);
        }
        __syncthreads();
        if (tid < BN) {
            int c = n0 + tid;
            const int cpg = p.Nout / p.G;
            double s = (c < p.Nout) ? (double)csum[tid] : 0.0, q = (c < p.Nout) ? (double)csq[tid] : 0.0;
            double* st = p.stats + (long long)(p.stats_per_shape ? b : 0) * p.G * 2;
            if (cpg % 32 == 0) {   // a warp's 32 columns sit in one group
                s = warp_sum(s); q = warp_sum(q);
                if ((tid & 31) == 0 && c < p.Nout) {
                    int g = c / cpg;
                    atomicAdd(&st[2 * g], s); atomicAdd(&st[2 * g + 1], q);
                }
            } else if (c < p.Nout) {
                int g = c / cpg;
                atomicAdd(&st[2 * g], s); atomicAdd(&st[2 * g + 1], q);
            }
        }
    }
}

// ================================================================================================ backward: data
// dZ[m][k] (+)= sum_n dY[m][n] W[n][k];  if finalize: dZ *= act'(pre(A)), and accumulate the GroupNorm-backward sums
//   gsum[b][g][0] += sum gamma_k dZ,  gsum[b][g][1] += sum gamma_k dZ xhat     (xhat = (A - mean) rstd)
struct BwdDataArgs {
    const float* dY; long long lddy;   // [B*Np][Nout]
    const float* W; long long ldw;     // [Nout][K]
    float* dZ; long long lddz;         // [B*Np][K]
    int accumulate;                    // dZ += (read-modify-write) instead of overwrite
    int finalize;                      // apply activation mask + GN sums using A below
    const float* A; long long lda;     // pre-norm input of this layer [B*Np][K]
    const float* in_scale; const float* in_shift; int in_act;   // [B][K]
    const float* gamma;                // [K] (norm weight of the producer) or null -> no norm: only mask
    const float* mean_rstd;            // [S][G][2] of the producer
    double* gsum;                      // [S][G][2]
    int B, Np, K, Nout, G, stats_per_shape;
    int vecdY, vecW, vecdZ;
};

// MINB = resident CTAs per SM the register allocation is capped for: 2 -> 128 registers (bwd_data spills 108 B, bwd_weight
// 584 B per thread), 1 -> no cap (159 / 207 registers, no spills, 8 warps per SM).  PN_LIN_BWD_OCC selects (default 2, measured faster: profiles/r01_linear_bwd_occupancy.md).
template <int MINB>
__global__ void __launch_bounds__(NT, MINB) linear_bwd_data_kernel(BwdDataArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tile& S = *reinterpret_cast<Tile*>(smem_raw);
    __shared__ float c1[BN], c2[BN];
    const int tiles_m = (p.Np + BM - 1) / BM;
    const int b = blockIdx.y / tiles_m;
    const int m0 = (blockIdx.y % tiles_m) * BM;
    const int k0 = blockIdx.x * BN;                 // output column block (over K)
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const float* dYb = p.dY + (long long)b * p.Np * p.lddy;
    Ident id;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    KcFrag fa; RcFrag fb;
    // reduction runs over n (Nout); A-operand = dY (n contiguous), B-operand = W[n][k] (k contiguous -> row-contig)
    load_kc(fa, dYb, p.lddy, m0, p.Np, 0, p.Nout, p.vecdY, id);
    load_rc(fb, p.W, p.ldw, 0, p.Nout, k0, p.K, p.vecW, id);
    store_kc(fa, S.a[0]); store_rc(fb, S.b[0]);
    __syncthreads();
    const int nk = (p.Nout + BK - 1) / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_kc(fa, dYb, p.lddy, m0, p.Np, (kt + 1) * BK, p.Nout, p.vecdY, id);
            load_rc(fb, p.W, p.ldw, (kt + 1) * BK, p.Nout, k0, p.K, p.vecW, id);
        }
        mma_slab(S.a[cur], S.b[cur], ty, tx, acc);
        if (kt + 1 < nk) { store_kc(fa, S.a[cur ^ 1]); store_rc(fb, S.b[cur ^ 1]); }
        __syncthreads();
    }
    const bool do_sums = p.finalize && p.gamma && p.gsum;
    if (do_sums) {
        if (tid < BN) { c1[tid] = 0.f; c2[tid] = 0.f; }
        __syncthreads();
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    const int cpg = p.G > 0 ? p.K / p.G : p.K;
    const long long so = (long long)b * p.K;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int r = m0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + i - 4);
        if (r >= p.Np) continue;
        long long row = (long long)b * p.Np + r;
        float* zrow = p.dZ + row * p.lddz;
        const float* arow = p.A ? p.A + row * p.lda : nullptr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int c = k0 + 64 * h + 4 * tx;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int cc = c + e;
                float t = acc[i][4 * h + e];
                if (cc < p.K) {
                    if (p.accumulate) t += zrow[cc];
                    if (p.finalize) {
                        float a = arow[cc];
                        float pre = p.in_scale ? fmaf(a, p.in_scale[so + cc], p.in_shift[so + cc]) : a;
                        t *= act_grad(pre, p.in_act);
                        if (do_sums) {
                            int g = cc / cpg;
                            const float* mr = p.mean_rstd + ((long long)(p.stats_per_shape ? b : 0) * p.G + g) * 2;
                            float xh = (a - mr[0]) * mr[1];
                            float gt = p.gamma[cc] * t;
                            s1[4 * h + e] += gt;
                            s2[4 * h + e] = fmaf(gt, xh, s2[4 * h + e]);
                        }
                    }
                }
                v[e] = t;
            }
            if (p.vecdZ && c + 3 < p.K) {
                *reinterpret_cast<float4*>(zrow + c) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (c + e < p.K) zrow[c + e] = v[e];
            }
        }
    }
    if (do_sums) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int cl = (j < 4 ? 0 : 60) + 4 * tx + j;
            atomicAdd(&c1[cl], s1[j]);
            atomicAdd(&c2[cl], s2[j]);
        }
        __syncthreads();
        if (tid < BN) {
            int c = k0 + tid;
            double a = (c < p.K) ? (double)c1[tid] : 0.0, q = (c < p.K) ? (double)c2[tid] : 0.0;
            double* st = p.gsum + (long long)(p.stats_per_shape ? b : 0) * p.G * 2;
            if (cpg % 32 == 0) {
                a = warp_sum(a); q = warp_sum(q);
                if ((tid & 31) == 0 && c < p.K) {
                    int g = c / cpg;
                    atomicAdd(&st[2 * g], a); atomicAdd(&st[2 * g + 1], q);
                }
            } else if (c < p.K) {
                int g = c / cpg;
                atomicAdd(&st[2 * g], a); atomicAdd(&st[2 * g + 1], q);
            }
        }
    }
}

// ================================================================================================ backward: weight
// dW[n][k] += sum_m dY[m][n] * act(norm(A))[m][k]   (reduction over rows m, split across grid.z, fp32 atomics)
// db[n] += sum_m dY[m][n]; dsb[b][n] += sum_{m in shape b} dY[m][n]   (only CTAs with blockIdx.y == 0)
struct BwdWArgs {
    const float* dY; long long lddy;
    const float* A; long long lda;
    const float* in_scale; const float* in_shift; int in_act;
    float* dW; long long lddw;     // [Nout][K]
    float* db;                     // [Nout] or null
    float* dsb;                    // [B][Nout] or null
    int B, Np, K, Nout;
    int rows_per_split;            // multiple of BK; a split never straddles two shapes
    int splits_per_shape;
    int vecdY, vecA;
};

template <int MINB>
__global__ void __launch_bounds__(NT, MINB) linear_bwd_weight_kernel(BwdWArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tile& S = *reinterpret_cast<Tile*>(smem_raw);
    const int n0 = blockIdx.x * BM;      // dW rows (Nout)
    const int k0 = blockIdx.y * BN;      // dW cols (K)
    const int b = blockIdx.z / p.splits_per_shape;
    const int r_begin = (blockIdx.z % p.splits_per_shape) * p.rows_per_split;
    const int r_end = min(p.Np, r_begin + p.rows_per_split);
    if (r_begin >= r_end) return;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const float* dYb = p.dY + (long long)b * p.Np * p.lddy;
    const float* Ab = p.A + (long long)b * p.Np * p.lda;
    const long long so = (long long)b * p.K;
    Ident id;
    auto xf = [&](float v, int /*m*/, int k) {
        if (p.in_scale) v = fmaf(v, p.in_scale[so + k], p.in_shift[so + k]);
        return act_fwd(v, p.in_act);
    };
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    RcFrag fa, fb;
    // reduction index = row m; both operands are "row-contiguous" slabs [BK rows][128 cols]
    load_rc(fa, dYb, p.lddy, r_begin, r_end, n0, p.Nout, p.vecdY, id);
    load_rc(fb, Ab, p.lda, r_begin, r_end, k0, p.K, p.vecA, xf);
    store_rc(fa, S.a[0]); store_rc(fb, S.b[0]);
    __syncthreads();
    const int nk = (r_end - r_begin + BK - 1) / BK;
    float colsum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) colsum[j] = 0.f;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_rc(fa, dYb, p.lddy, r_begin + (kt + 1) * BK, r_end, n0, p.Nout, p.vecdY, id);
            load_rc(fb, Ab, p.lda, r_begin + (kt + 1) * BK, r_end, k0, p.K, p.vecA, xf);
        }
        mma_slab(S.a[cur], S.b[cur], ty, tx, acc);
        if ((p.db || p.dsb) && blockIdx.y == 0 && tx == 0) {
            // column sums of dY for this CTA's 128 n-columns: thread (ty,0) owns n = 4ty..4ty+3, 64+4ty..
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float4 a0 = *reinterpret_cast<const float4*>(&S.a[cur][kk][4 * ty]);
                float4 a1 = *reinterpret_cast<const float4*>(&S.a[cur][kk][64 + 4 * ty]);
                colsum[0] += a0.x; colsum[1] += a0.y; colsum[2] += a0.z; colsum[3] += a0.w;
                colsum[4] += a1.x; colsum[5] += a1.y; colsum[6] += a1.z; colsum[7] += a1.w;
            }
        }
        if (kt + 1 < nk) { store_rc(fa, S.a[cur ^ 1]); store_rc(fb, S.b[cur ^ 1]); }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int n = n0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + i - 4);
        if (n >= p.Nout) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int k = k0 + (j < 4 ? 4 * tx + j : 64 + 4 * tx + j - 4);
            if (k < p.K) atomicAdd(&p.dW[(long long)n * p.lddw + k], acc[i][j]);
        }
    }
    if ((p.db || p.dsb) && blockIdx.y == 0 && tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int n = n0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + i - 4);
            if (n >= p.Nout) continue;
            if (p.db) atomicAdd(&p.db[n], colsum[i]);
            if (p.dsb) atomicAdd(&p.dsb[(long long)b * p.Nout + n], colsum[i]);
        }
    }
}

// ================================================================================================ norm helpers
// stats [S][G][2] (double sum, sumsq over `count` elements) -> mean_rstd [S][G][2] and per-channel scale/shift [S][C]
__global__ void norm_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, int S, int G, int C, double count, float eps,
                                     float* __restrict__ mean_rstd, float* __restrict__ scale,
                                     float* __restrict__ shift) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * C) return;
    int s = i / C, c = i % C, g = c / (C / G);
    double sum = stats[((long long)s * G + g) * 2], sq = stats[((long long)s * G + g) * 2 + 1];
    double mean = sum / count;
    double var = sq / count - mean * mean;
    if (var < 0) var = 0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float m = (float)mean;
    if (c % (C / G) == 0) {
        mean_rstd[((long long)s * G + g) * 2] = m;
        mean_rstd[((long long)s * G + g) * 2 + 1] = rstd;
    }
    float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    float sc = ga * rstd;
    scale[i] = sc;
    shift[i] = be - m * sc;
}

// in-place GroupNorm backward on dZ (already multiplied by the activation mask):
//   dA = rstd * (gamma * dZ - (s1 + xhat * s2) / count),  plus dgamma/dbeta accumulation
__global__ void norm_bwd_apply_kernel(float* __restrict__ dZ, long long lddz, const float* __restrict__ A,
                                      long long lda, const float* __restrict__ gamma,
                                      const float* __restrict__ mean_rstd, const double* __restrict__ gsum,
                                      int B, int Np, int C, int G, int stats_per_shape, double count,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
    // block = 256 threads covering (rows_per_block x C) ; thread handles channel c = threadIdx.x % C-chunk
    // layout: blockDim.x = 256, each block handles 64 rows; thread t loops over channels t, t+256, ...
    const int rows_per_block = 64;
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(Np, r0 + rows_per_block);
    const int cpg = C / G;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int g = c / cpg;
        long long si = ((long long)(stats_per_shape ? b : 0) * G + g) * 2;
        float mean = mean_rstd[si], rstd = mean_rstd[si + 1];
        float s1 = (float)(gsum[si] / count), s2 = (float)(gsum[si + 1] / count);
        float ga = gamma[c];
        float dg = 0.f, dbt = 0.f;
        for (int r = r0; r < r1; ++r) {
            long long row = (long long)b * Np + r;
            float dz = dZ[row * lddz + c];
            float xh = (A[row * lda + c] - mean) * rstd;
            dg = fmaf(dz, xh, dg);
            dbt += dz;
            dZ[row * lddz + c] = rstd * (ga * dz - (s1 + xh * s2));
        }
        if (dgamma) { atomicAdd(&dgamma[c], dg); atomicAdd(&dbeta[c], dbt); }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace lin
}  // namespace pn

using namespace pn;
using namespace pn::lin;

static int set_smem(const void* fn) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile));
    return e == cudaSuccess ? 0 : 1;
}

extern "C" int pn_linear_fwd(const float* A, long long lda, const float* W, long long ldw, const float* bias,
                             const float* sbias, const float* in_scale, const float* in_shift, int in_act,
                             float* Y, long long ldy, double* stats, int B, int Np, int K, int Nout, int G,
                             int stats_per_shape, void* stream) {
    PN_REQUIRE(A && W && Y, "pn_linear_fwd: null pointer");
    PN_REQUIRE(B > 0 && Np > 0 && K > 0 && Nout > 0, "pn_linear_fwd: bad shape");
    PN_REQUIRE(!stats || (G > 0 && Nout % G == 0), "pn_linear_fwd: Nout %% G != 0");
    FwdArgs p{A, lda, W, ldw, bias, sbias, in_scale, in_shift, in_act, Y, ldy, stats, B, Np, K, Nout, G,
              stats_per_shape, 0, 0, 0};
    p.vecA = aligned16(A) && lda % 4 == 0;
    p.vecW = aligned16(W) && ldw % 4 == 0;
    p.vecY = aligned16(Y) && ldy % 4 == 0;
    PN_REQUIRE(set_smem((const void*)linear_fwd_kernel) == 0, "pn_linear_fwd: smem attribute");
    PN_REQUIRE((long long)cdiv(Np, BM) * B <= 65535, "pn_linear_fwd: too many row tiles (%d x %d)", cdiv(Np, BM), B);
    dim3 grid(cdiv(Nout, BN), cdiv(Np, BM) * B);
    linear_fwd_kernel<<<grid, NT, sizeof(Tile), (cudaStream_t)stream>>>(p);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("linear_fwd_kernel");
    return PN_OK;
}

static int bwd_occ() {      // read per call: tools/exp_linear_bwd.py flips it inside one process
    const char* e = getenv("PN_LIN_BWD_OCC");
    return (e && e[0] == '1') ? 1 : 2;      // measured: 2 is faster (bwd_data 8.2 vs 11.9 ms, bwd_weight equal)
}

extern "C" int pn_linear_bwd_data(const float* dY, long long lddy, const float* W, long long ldw, float* dZ,
                                  long long lddz, int accumulate, int finalize, const float* A, long long lda,
                                  const float* in_scale, const float* in_shift, int in_act, const float* gamma,
                                  const float* mean_rstd, double* gsum, int B, int Np, int K, int Nout, int G,
                                  int stats_per_shape, void* stream) {
    PN_REQUIRE(dY && W && dZ, "pn_linear_bwd_data: null pointer");
    PN_REQUIRE(!finalize || A, "pn_linear_bwd_data: finalize needs A");
    PN_REQUIRE(!(finalize && gamma) || (mean_rstd && gsum && G > 0 && K % G == 0), "pn_linear_bwd_data: norm args");
    BwdDataArgs p{dY, lddy, W, ldw, dZ, lddz, accumulate, finalize, A, lda, in_scale, in_shift, in_act, gamma,
                  mean_rstd, gsum, B, Np, K, Nout, G, stats_per_shape, 0, 0, 0};
    p.vecdY = aligned16(dY) && lddy % 4 == 0;
    p.vecW = aligned16(W) && ldw % 4 == 0;
    p.vecdZ = aligned16(dZ) && lddz % 4 == 0;
    auto kern = bwd_occ() == 2 ? linear_bwd_data_kernel<2> : linear_bwd_data_kernel<1>;
    PN_REQUIRE(set_smem((const void*)kern) == 0, "pn_linear_bwd_data: smem attribute");
    PN_REQUIRE((long long)cdiv(Np, BM) * B <= 65535, "pn_linear_bwd_data: too many row tiles (%d x %d)", cdiv(Np, BM), B);
    dim3 grid(cdiv(K, BN), cdiv(Np, BM) * B);
    kern<<<grid, NT, sizeof(Tile), (cudaStream_t)stream>>>(p);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("linear_bwd_data_kernel");
    return PN_OK;
}

extern "C" int pn_linear_bwd_weight(const float* dY, long long lddy, const float* A, long long lda,
                                    const float* in_scale, const float* in_shift, int in_act, float* dW,
                                    long long lddw, float* db, float* dsb, int B, int Np, int K, int Nout,
                                    void* stream) {
    PN_REQUIRE(dY && A && dW, "pn_linear_bwd_weight: null pointer");
    BwdWArgs p{dY, lddy, A, lda, in_scale, in_shift, in_act, dW, lddw, db, dsb, B, Np, K, Nout, 0, 0, 0, 0};
    p.vecdY = aligned16(dY) && lddy % 4 == 0;
    p.vecA = aligned16(A) && lda % 4 == 0;
    // aim for ~2 waves of 296 CTAs
    int tiles = cdiv(Nout, BM) * cdiv(K, BN);
    int want = max(1, (592 + tiles * B - 1) / (tiles * B));
    int rows = cdiv(Np, want);
    rows = max(BK * 4, ((rows + BK - 1) / BK) * BK);
    p.rows_per_split = rows;
    p.splits_per_shape = cdiv(Np, rows);
    auto kern = bwd_occ() == 2 ? linear_bwd_weight_kernel<2> : linear_bwd_weight_kernel<1>;
    PN_REQUIRE(set_smem((const void*)kern) == 0, "pn_linear_bwd_weight: smem attribute");
    dim3 grid(cdiv(Nout, BM), cdiv(K, BN), p.splits_per_shape * B);
    kern<<<grid, NT, sizeof(Tile), (cudaStream_t)stream>>>(p);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("linear_bwd_weight_kernel");
    return PN_OK;
}

extern "C" int pn_norm_finalize(const double* stats, const float* gamma, const float* beta, int S, int G, int C,
                                double count, float eps, float* mean_rstd, float* scale, float* shift,
                                void* stream) {
    PN_REQUIRE(stats && mean_rstd && scale && shift && G > 0 && C % G == 0, "pn_norm_finalize: bad args");
    int n = S * C;
    norm_finalize_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(stats, gamma, beta, S, G, C, count, eps,
                                                                          mean_rstd, scale, shift);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("norm_finalize_kernel");
    return PN_OK;
}

extern "C" int pn_norm_bwd_apply(float* dZ, long long lddz, const float* A, long long lda, const float* gamma,
                                 const float* mean_rstd, const double* gsum, int B, int Np, int C, int G,
                                 int stats_per_shape, double count, float* dgamma, float* dbeta, void* stream) {
    PN_REQUIRE(dZ && A && gamma && mean_rstd && gsum && G > 0 && C % G == 0, "pn_norm_bwd_apply: bad args");
    dim3 grid(cdiv(Np, 64), B);
    norm_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dZ, lddz, A, lda, gamma, mean_rstd, gsum, B, Np,
                                                                   C, G, stats_per_shape, count, dgamma, dbeta);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("norm_bwd_apply_kernel");
    return PN_OK;
}
