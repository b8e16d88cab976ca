// Small fused kernels around the GEMMs: global max-pool with GroupNorm, log-softmax head, row L2-normalise,
// triplet (embedding) loss.
//
// Replaces (reference): F.relu(bnmlp1(mlp1(x))).max(dim=2)  src/PointNet.py:194-196
//                       LogSoftmax(dim=1) + NLLLoss          src/PointNet.py:284, src/segment_loss.py:151
//                       F.normalize + EmbeddingLoss.triplet_loss   src/segment_loss.py:31-124
#include "common.cuh"

namespace pn {
namespace pw {

// ------------------------------------------------------------------------------------------------ colmax + norm
// Y [B][N][C] pre-norm; out[b][c] = max_n act(scale*y + shift) * (wts ? wts[b][n] : 1); arg[b][c] = first n attaining it
// grid (C/32, B), block 256 = 8 row-lanes x 32 columns
__global__ void __launch_bounds__(256) colmax_norm_kernel(const float* __restrict__ Y, long long ldy, int N, int C,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int act,
                                                          const float* __restrict__ wts, float* __restrict__ out,
                                                          int* __restrict__ arg) {
    __shared__ float sv[8][32];
    __shared__ int si[8][32];
    const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    float sc = 1.f, sh = 0.f;
    if (c < C) { sc = scale[(long long)b * C + c]; sh = shift[(long long)b * C + c]; }
    float best = -INFINITY;
    int bi = 0;
    if (c < C) {
        const float* yb = Y + (long long)b * N * ldy + c;
        for (int n = rl; n < N; n += 8) {
            float pre = fmaf(yb[(long long)n * ldy], sc, sh);
            float v = act == 1 ? fmaxf(pre, 0.f) : (act == 2 ? (pre > 0.f ? pre : 0.2f * pre) : pre);
            if (wts) v *= wts[(long long)b * N + n];
            bool better = v > best;
            best = better ? v : best;
            bi = better ? n : bi;
        }
    }
    sv[rl][threadIdx.x & 31] = best; si[rl][threadIdx.x & 31] = bi;
    __syncthreads();
    if (rl == 0 && c < C) {
        for (int r = 1; r < 8; ++r) {
            float v = sv[r][threadIdx.x]; int i = si[r][threadIdx.x];
            bool better = v > best || (v == best && i < bi);
            best = better ? v : best; bi = better ? i : bi;
        }
        out[(long long)b * C + c] = best;
        arg[(long long)b * C + c] = bi;
    }
}

// dense backward of (norm + act + max over N):
//   dY[b][n][c] = rstd * ( gamma*gt[b][c]*[n == arg] - (s1[b][g] + xhat * s2[b][g]) )    (s1,s2 already / count)
// gt = upstream grad * act'(pre) ; coef layout: gsum [S][G][2] double (un-normalised), count
__global__ void colmax_bwd_fill_kernel(const float* __restrict__ Y, long long ldy, const float* __restrict__ gt,
                                       const int* __restrict__ arg, const float* __restrict__ gamma,
                                       const float* __restrict__ mean_rstd, const double* __restrict__ gsum,
                                       int N, int C, int G, int stats_per_shape, double count, int dense,
                                       float* __restrict__ dY, long long lddy) {
    const int b = blockIdx.z;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n0 = blockIdx.y * 32;
    if (c >= C) return;
    const int g = c / (C / G);
    long long si = ((long long)(stats_per_shape ? b : 0) * G + g) * 2;
    float mean = mean_rstd[si], rstd = mean_rstd[si + 1];
    float s1 = dense ? (float)(gsum[si] / count) : 0.f, s2 = dense ? (float)(gsum[si + 1] / count) : 0.f;
    float gg = gamma[c] * gt[(long long)b * C + c];
    int a = arg[(long long)b * C + c];
    for (int n = n0; n < min(N, n0 + 32); ++n) {
        long long row = (long long)b * N + n;
        float xh = (Y[row * ldy + c] - mean) * rstd;
        float v = -(s1 + xh * s2);
        if (n == a) v += gg;
        dY[row * lddy + c] = rstd * v;
    }
}

// ------------------------------------------------------------------------------------------------ log-softmax head
// logits [B][N][P] (pitch ldl) -> logp [B][P][N] (channel-major, the reference's output layout)
__global__ void logsoftmax_fwd_kernel(const float* __restrict__ logits, long long ldl, int N, int P,
                                      float* __restrict__ logp) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* l = logits + ((long long)b * N + n) * ldl;
    float m = -INFINITY;
    for (int p = 0; p < P; ++p) m = fmaxf(m, l[p]);
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += expf(l[p] - m);
    float lse = m + logf(s);
    for (int p = 0; p < P; ++p) logp[((long long)b * P + p) * N + n] = l[p] - lse;
}
// dlogits[b][n][p] = dlp[b][p][n] - exp(logp[b][p][n]) * sum_p dlp
__global__ void logsoftmax_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ dlp, int N, int P,
                                      float* __restrict__ dlogits, long long ldd) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += dlp[((long long)b * P + p) * N + n];
    for (int p = 0; p < P; ++p) {
        long long o = ((long long)b * P + p) * N + n;
        dlogits[((long long)b * N + n) * ldd + p] = dlp[o] - expf(logp[o]) * s;
    }
}
// NLL (mean over B*N): loss = -sum logp[b][t][n] / (B*N); dlp = -g / (B*N) at the target, 0 elsewhere
__global__ void nll_fwd_kernel(const float* __restrict__ logp, const long long* __restrict__ target, int B, int N,
                               int P, float* __restrict__ loss) {
    __shared__ float red[32];
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    if (e < (long long)B * N) {
        long long b = e / N, n = e % N;
        v = -logp[(b * P + target[e]) * N + n];
    }
    v = block_sum(v, red);
    if (threadIdx.x == 0) atomicAdd(loss, v / (float)((long long)B * N));
}
__global__ void nll_bwd_kernel(const long long* __restrict__ target, const float* __restrict__ gout, int B, int N,
                               int P, float* __restrict__ dlp) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)B * N) return;
    long long b = e / N, n = e % N;
    dlp[(b * P + target[e]) * N + n] = -gout[0] / (float)((long long)B * N);
}

// ------------------------------------------------------------------------------------------------ row L2 normalise
// y = x / max(||x||, eps)  (F.normalize, p=2); one warp per row, D <= 1024
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, long long ldx, long long rows, int D, float eps,
                                  float* __restrict__ y, long long ldy, float* __restrict__ norms) {
    long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* xr = x + r * ldx;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(xr[d], xr[d], s);
    s = warp_sum(s);
    float nrm = fmaxf(sqrtf(s), eps);
    for (int d = lane; d < D; d += 32) y[r * ldy + d] = xr[d] / nrm;
    if (norms && lane == 0) norms[r] = nrm;
}
// dx = (dy - y (y . dy)) / nrm     (valid when ||x|| > eps)
__global__ void l2norm_bwd_kernel(const float* __restrict__ y, long long ldy, const float* __restrict__ dy,
                                  long long lddy, const float* __restrict__ norms, long long rows, int D,
                                  float* __restrict__ dx, long long lddx, int accumulate) {
    long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(y[r * ldy + d], dy[r * lddy + d], s);
    s = warp_sum(s);
    float inv = 1.f / norms[r];
    for (int d = lane; d < D; d += 32) {
        float v = (dy[r * lddy + d] - y[r * ldy + d] * s) * inv;
        if (accumulate) dx[r * lddx + d] += v; else dx[r * lddx + d] = v;
    }
}

// ------------------------------------------------------------------------------------------------ triplet loss
// For pair t: rows a_idx[t][0..S) (anchors/positives, label k1) and n_idx[t][0..S) (negatives, label k2) of the
// normalised embedding E [rows][D].  c[a][m] = relu(|E_a - E_pm|^2 - |E_a - E_nm|^2 + margin);
// pair_loss = (sum c - trace c) / (count(c > 0) + 1).      (src/segment_loss.py:101-119)
// One CTA per pair; S <= 32, D <= 256.
constexpr int TS = 32;
__global__ void __launch_bounds__(256) triplet_fwd_kernel(const float* __restrict__ E, long long lde, int D,
                                                          const int* __restrict__ a_idx,
                                                          const int* __restrict__ n_idx, int S, float margin,
                                                          float* __restrict__ pair_loss,
                                                          float* __restrict__ pair_sat) {
    extern __shared__ float sm[];          // A [S][D+1], Nn [S][D+1]
    float* A = sm;
    float* Ng = sm + S * (D + 1);
    __shared__ float red[32];
    const int t = blockIdx.x;
    for (int e = threadIdx.x; e < S * D; e += blockDim.x) {
        int s = e / D, d = e % D;
        A[s * (D + 1) + d] = E[(long long)a_idx[t * S + s] * lde + d];
        Ng[s * (D + 1) + d] = E[(long long)n_idx[t * S + s] * lde + d];
    }
    __syncthreads();
    float lsum = 0.f, lcnt = 0.f;
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        int a = e / S, m = e % S;
        float dp = 0.f, dn = 0.f;
        for (int d = 0; d < D; ++d) {
            float x = A[a * (D + 1) + d];
            float u = x - A[m * (D + 1) + d], w = x - Ng[m * (D + 1) + d];
            dp = fmaf(u, u, dp); dn = fmaf(w, w, dn);
        }
        float c = fmaxf(dp - dn + margin, 0.f);
        if (c > 0.f) lcnt += 1.f;
        if (a != m) lsum += c;
    }
    lsum = block_sum(lsum, red);
    lcnt = block_sum(lcnt, red);
    if (threadIdx.x == 0) {
        pair_sat[t] = lcnt + 1.f;
        pair_loss[t] = lsum / (lcnt + 1.f);
    }
}
// dE[row] += w_t / sat_t * d c / d row    for c > 0, a != m:  dc/dA_a = 2(Nn_m - A_m), dc/dA_m(pos) = -2 (A_a - A_m),
//                                                             dc/dNn_m = 2 (A_a - Nn_m)
__global__ void __launch_bounds__(256) triplet_bwd_kernel(const float* __restrict__ E, long long lde, int D,
                                                          const int* __restrict__ a_idx,
                                                          const int* __restrict__ n_idx, int S, float margin,
                                                          const float* __restrict__ pair_sat,
                                                          const float* __restrict__ pair_w,   // upstream weight / pair
                                                          float* __restrict__ dE, long long ldde) {
    extern __shared__ float sm[];
    float* A = sm;
    float* Ng = sm + S * (D + 1);
    float* gA = Ng + S * (D + 1);          // [S][D+1]
    float* gN = gA + S * (D + 1);
    __shared__ unsigned char mask[TS * TS];
    const int t = blockIdx.x;
    const float w = pair_w[t] / pair_sat[t];
    for (int e = threadIdx.x; e < S * D; e += blockDim.x) {
        int s = e / D, d = e % D;
        A[s * (D + 1) + d] = E[(long long)a_idx[t * S + s] * lde + d];
        Ng[s * (D + 1) + d] = E[(long long)n_idx[t * S + s] * lde + d];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        int a = e / S, m = e % S;
        float dp = 0.f, dn = 0.f;
        for (int d = 0; d < D; ++d) {
            float x = A[a * (D + 1) + d];
            float u = x - A[m * (D + 1) + d], v = x - Ng[m * (D + 1) + d];
            dp = fmaf(u, u, dp); dn = fmaf(v, v, dn);
        }
        mask[e] = (a != m) && (dp - dn + margin > 0.f);
    }
    __syncthreads();
    // gA[s][d] = sum_m mask[s][m] 2 (Ng_m - A_m)  +  sum_a mask[a][s] (-2)(A_a - A_s);   gN[s][d] = sum_a mask[a][s] 2 (A_a - Ng_s)
    for (int e = threadIdx.x; e < S * D; e += blockDim.x) {
        int s = e / D, d = e % D;
        float ga = 0.f, gn = 0.f;
        float as = A[s * (D + 1) + d], ns = Ng[s * (D + 1) + d];
        for (int m = 0; m < S; ++m) {
            if (mask[s * S + m]) ga += 2.f * (Ng[m * (D + 1) + d] - A[m * (D + 1) + d]);
            if (mask[m * S + s]) {
                float am = A[m * (D + 1) + d];
                ga -= 2.f * (am - as);
                gn += 2.f * (am - ns);
            }
        }
        atomicAdd(&dE[(long long)a_idx[t * S + s] * ldde + d], w * ga);
        atomicAdd(&dE[(long long)n_idx[t * S + s] * ldde + d], w * gn);
    }
}

}  // namespace pw
}  // namespace pn

using namespace pn;
using namespace pn::pw;

extern "C" int pn_colmax_norm(const float* Y, long long ldy, int B, int N, int C, const float* scale,
                              const float* shift, int act, const float* wts, float* out, int* arg, void* stream) {
    PN_REQUIRE(Y && scale && shift && out && arg, "pn_colmax_norm: null pointer");
    dim3 grid(cdiv(C, 32), B);
    colmax_norm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, ldy, N, C, scale, shift, act, wts, out, arg);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("colmax_norm_kernel");
    return PN_OK;
}

extern "C" int pn_colmax_bwd_fill(const float* Y, long long ldy, const float* gt, const int* arg,
                                  const float* gamma, const float* mean_rstd, const double* gsum, int B, int N,
                                  int C, int G, int stats_per_shape, double count, int dense, float* dY,
                                  long long lddy, void* stream) {
    PN_REQUIRE(Y && gt && arg && gamma && mean_rstd && dY, "pn_colmax_bwd_fill: null pointer");
    PN_REQUIRE(!dense || gsum, "pn_colmax_bwd_fill: dense needs gsum");
    dim3 grid(cdiv(C, 128), cdiv(N, 32), B);
    colmax_bwd_fill_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(Y, ldy, gt, arg, gamma, mean_rstd, gsum, N, C, G,
                                                                    stats_per_shape, count, dense, dY, lddy);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("colmax_bwd_fill_kernel");
    return PN_OK;
}

extern "C" int pn_logsoftmax_fwd(const float* logits, long long ldl, int B, int N, int P, float* logp, void* stream) {
    PN_REQUIRE(logits && logp && P > 0, "pn_logsoftmax_fwd: bad args");
    dim3 grid(cdiv(N, 256), B);
    logsoftmax_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, ldl, N, P, logp);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("logsoftmax_fwd_kernel");
    return PN_OK;
}
extern "C" int pn_logsoftmax_bwd(const float* logp, const float* dlp, int B, int N, int P, float* dlogits,
                                 long long ldd, void* stream) {
    PN_REQUIRE(logp && dlp && dlogits, "pn_logsoftmax_bwd: null pointer");
    dim3 grid(cdiv(N, 256), B);
    logsoftmax_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logp, dlp, N, P, dlogits, ldd);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("logsoftmax_bwd_kernel");
    return PN_OK;
}
extern "C" int pn_nll_fwd(const float* logp, const long long* target, int B, int N, int P, float* loss_zeroed,
                          void* stream) {
    PN_REQUIRE(logp && target && loss_zeroed, "pn_nll_fwd: null pointer");
    nll_fwd_kernel<<<cdiv((long long)B * N, 256), 256, 0, (cudaStream_t)stream>>>(logp, target, B, N, P, loss_zeroed);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("nll_fwd_kernel");
    return PN_OK;
}
extern "C" int pn_nll_bwd(const long long* target, const float* gout, int B, int N, int P, float* dlp_zeroed,
                          void* stream) {
    PN_REQUIRE(target && gout && dlp_zeroed, "pn_nll_bwd: null pointer");
    nll_bwd_kernel<<<cdiv((long long)B * N, 256), 256, 0, (cudaStream_t)stream>>>(target, gout, B, N, P, dlp_zeroed);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("nll_bwd_kernel");
    return PN_OK;
}
extern "C" int pn_l2norm_fwd(const float* x, long long ldx, long long rows, int D, float eps, float* y,
                             long long ldy, float* norms, void* stream) {
    PN_REQUIRE(x && y, "pn_l2norm_fwd: null pointer");
    l2norm_fwd_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, D, eps, y, ldy, norms);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("l2norm_fwd_kernel");
    return PN_OK;
}
extern "C" int pn_l2norm_bwd(const float* y, long long ldy, const float* dy, long long lddy, const float* norms,
                             long long rows, int D, float* dx, long long lddx, int accumulate, void* stream) {
    PN_REQUIRE(y && dy && norms && dx, "pn_l2norm_bwd: null pointer");
    l2norm_bwd_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(y, ldy, dy, lddy, norms, rows, D, dx, lddx,
                                                                       accumulate);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("l2norm_bwd_kernel");
    return PN_OK;
}
extern "C" int pn_triplet_fwd(const float* E, long long lde, int D, const int* a_idx, const int* n_idx, int T, int S,
                              float margin, float* pair_loss, float* pair_sat, void* stream) {
    PN_REQUIRE(E && a_idx && n_idx && pair_loss && pair_sat, "pn_triplet_fwd: null pointer");
    PN_REQUIRE(S > 0 && S <= TS && D <= 256, "pn_triplet_fwd: need S <= 32, D <= 256 (S=%d D=%d)", S, D);
    if (T == 0) return PN_OK;
    size_t sm = sizeof(float) * 2 * S * (D + 1);
    PN_CUDA(cudaFuncSetAttribute(triplet_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    triplet_fwd_kernel<<<T, 256, sm, (cudaStream_t)stream>>>(E, lde, D, a_idx, n_idx, S, margin, pair_loss, pair_sat);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("triplet_fwd_kernel");
    return PN_OK;
}
extern "C" int pn_triplet_bwd(const float* E, long long lde, int D, const int* a_idx, const int* n_idx, int T, int S,
                              float margin, const float* pair_sat, const float* pair_w, float* dE, long long ldde,
                              void* stream) {
    PN_REQUIRE(E && a_idx && n_idx && pair_sat && pair_w && dE, "pn_triplet_bwd: null pointer");
    PN_REQUIRE(S > 0 && S <= TS && D <= 256, "pn_triplet_bwd: need S <= 32, D <= 256");
    if (T == 0) return PN_OK;
    size_t sm = sizeof(float) * 4 * S * (D + 1);
    PN_CUDA(cudaFuncSetAttribute(triplet_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    triplet_bwd_kernel<<<T, 256, sm, (cudaStream_t)stream>>>(E, lde, D, a_idx, n_idx, S, margin, pair_sat, pair_w, dE,
                                                             ldde);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("triplet_bwd_kernel");
    return PN_OK;
}
