// Edge-conv (DGCNN) gather / reduce kernels, factored form.
//
// Replaces get_graph_feature + Conv2d(1x1,no bias) + GroupNorm/BatchNorm + LeakyReLU(0.2) + max over k
//   src/PointNet.py:72-103,157-165,176-192 and src/model.py:25-53,74-85,140-154
// The reference materialises the (B,2C,N,k) edge tensor (6.5 GB at B=16,N=10^4,k=80,C=64) and the (B,Cout,N,k)
// conv output.  Here neither exists:
//   W.[x_j - x_i ; x_i] = W1.x_j + (W2-W1).x_i  =:  P[j] + Q[i]          (two per-point GEMMs, linear.cu)
//   max_k act(norm(e_k)) = act(norm(max_k e_k))  if gamma >= 0,  act(norm(min_k e_k)) otherwise
// so the forward is ONE gather pass over P rows (N*k*Cout*4 bytes, L2-resident) that tracks, per (point, channel),
// the selected extreme (value + neighbour id), sum_k e and the group statistics sum e, sum e^2.
// The backward needs the dense GroupNorm terms; they reduce to per-point quantities plus ONE reverse-graph
// aggregation  SQ[j] = sum_{i : j in nbr(i)} Q[i]  done over a CSR transpose of the kNN graph (no float atomics
// on the dense path).
#include "common.cuh"

namespace pn {
namespace edge {

constexpr int NT = 256;
constexpr int PTS_PER_WARP = 8;

// ------------------------------------------------------------------------------------------------ forward gather
// PQ: [B][N][ldpq] with P = cols [0,Cout), Q = cols [Cout, 2Cout).  idx: [B][N][k] int32.
// gamma: [Cout] (sign selects max/min).  Outputs esel/jsel/esum: [B][N][Cout].  stats: [S][G][2] double.
template <int V>
__global__ void __launch_bounds__(NT) edge_gather_kernel(const float* __restrict__ PQ, long long ldpq,
                                                         const int* __restrict__ idx, int N, int k, int Cout,
                                                         const float* __restrict__ gamma, float* __restrict__ esel,
                                                         int* __restrict__ jsel, float* __restrict__ esum,
                                                         double* __restrict__ stats, int G, int stats_per_shape) {
    extern __shared__ double gacc[];   // [G][2]
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cpg = Cout / G;
    for (int t = threadIdx.x; t < 2 * G; t += NT) gacc[t] = 0.0;
    __syncthreads();
    const float* PQb = PQ + (long long)b * N * ldpq;
    const int* idxb = idx + (long long)b * N * k;
    const int i0 = (blockIdx.x * (NT / 32) + warp) * PTS_PER_WARP;

    for (int cb = 0; cb < Cout; cb += 32 * V) {          // channel chunk handled by the warp
        const int c0 = cb + lane * V;
        const bool cvalid = c0 < Cout;                    // Cout % (V) == 0 guaranteed by the launcher
        bool wantmax[V];
#pragma unroll
        for (int v = 0; v < V; ++v) wantmax[v] = cvalid ? (gamma[c0 + v] >= 0.f) : true;
        float ts[V], tq[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { ts[v] = 0.f; tq[v] = 0.f; }

        for (int pi = 0; pi < PTS_PER_WARP; ++pi) {
            const int i = i0 + pi;
            if (i >= N) break;
            float q[V], best[V], s[V], s2[V];
            int bj[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                q[v] = cvalid ? PQb[(long long)i * ldpq + Cout + c0 + v] : 0.f;
                best[v] = wantmax[v] ? -INFINITY : INFINITY;
                bj[v] = 0; s[v] = 0.f; s2[v] = 0.f;
            }
            for (int t0 = 0; t0 < k; t0 += 32) {
                int myj = (t0 + lane < k) ? idxb[(long long)i * k + t0 + lane] : 0;
                const int tn = min(32, k - t0);
#pragma unroll 4
                for (int t = 0; t < tn; ++t) {
                    int j = __shfl_sync(FULL, myj, t);
                    if (cvalid) {
                        const float* pr = PQb + (long long)j * ldpq + c0;
                        float pv[V];
                        if constexpr (V == 4) {
                            float4 w = *reinterpret_cast<const float4*>(pr);
                            pv[0] = w.x; pv[1] = w.y; pv[2] = w.z; pv[3] = w.w;
                        } else if constexpr (V == 2) {
                            float2 w = *reinterpret_cast<const float2*>(pr);
                            pv[0] = w.x; pv[1] = w.y;
                        } else {
#pragma unroll
                            for (int v = 0; v < V; ++v) pv[v] = pr[v];
                        }
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            float e = pv[v] + q[v];
                            bool better = wantmax[v] ? (e > best[v]) : (e < best[v]);
                            best[v] = better ? e : best[v];
                            bj[v] = better ? j : bj[v];
                            s[v] += e;
                            s2[v] = fmaf(e, e, s2[v]);
                        }
                    }
                }
            }
            if (cvalid) {
                long long o = ((long long)b * N + i) * Cout + c0;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    esel[o + v] = best[v]; jsel[o + v] = bj[v]; esum[o + v] = s[v];
                    ts[v] += s[v]; tq[v] += s2[v];
                }
            }
        }
        if (cvalid && stats) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                int g = (c0 + v) / cpg;
                atomicAdd(&gacc[2 * g], (double)ts[v]);
                atomicAdd(&gacc[2 * g + 1], (double)tq[v]);
            }
        }
    }
    if (stats) {
        __syncthreads();
        double* st = stats + (long long)(stats_per_shape ? b : 0) * G * 2;
        for (int t = threadIdx.x; t < 2 * G; t += NT) atomicAdd(&st[t], gacc[t]);
    }
}

// out[b][i][c] = lrelu_0.2(scale[b][c] * esel + shift[b][c])   (out has row pitch ldo: slice of the concat buffer)
__global__ void edge_apply_kernel(const float* __restrict__ esel, const float* __restrict__ scale,
                                  const float* __restrict__ shift, float* __restrict__ out, long long ldo,
                                  long long rows_total, int N, int Cout) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows_total * Cout) return;
    long long row = e / Cout;
    int c = (int)(e % Cout);
    int b = (int)(row / N);
    float pre = fmaf(esel[e], scale[(long long)b * Cout + c], shift[(long long)b * Cout + c]);
    out[row * ldo + c] = pre > 0.f ? pre : 0.2f * pre;
}

// ------------------------------------------------------------------------------------------------ backward
// prep: dy = g * lrelu'(pre); sums s1 = sum gamma dy, s2 = sum gamma dy xhat per (shape, group); dgamma/dbeta.
// grid (ceil(N/64), B), 256 threads; thread strides over channels.
__global__ void __launch_bounds__(NT) edge_bwd_prep_kernel(const float* __restrict__ g, long long ldg,
                                                           const float* __restrict__ esel,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ mean_rstd,
                                                           const float* __restrict__ gamma, int N, int Cout, int G,
                                                           int stats_per_shape, float* __restrict__ dy,
                                                           double* __restrict__ gsum, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
    extern __shared__ double gacc[];   // [G][2]
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * 64, r1 = min(N, r0 + 64);
    const int cpg = Cout / G;
    for (int t = threadIdx.x; t < 2 * G; t += NT) gacc[t] = 0.0;
    __syncthreads();
    for (int c = threadIdx.x; c < Cout; c += NT) {
        int gi = c / cpg;
        long long si = ((long long)(stats_per_shape ? b : 0) * G + gi) * 2;
        float mean = mean_rstd[si], rstd = mean_rstd[si + 1];
        float sc = scale[(long long)b * Cout + c], sh = shift[(long long)b * Cout + c], ga = gamma[c];
        float a1 = 0.f, a2 = 0.f;
        for (int r = r0; r < r1; ++r) {
            long long row = (long long)b * N + r;
            float e = esel[row * Cout + c];
            float pre = fmaf(e, sc, sh);
            float d = g[row * ldg + c] * (pre > 0.f ? 1.f : 0.2f);
            dy[row * Cout + c] = d;
            float xh = (e - mean) * rstd;
            a1 += d;
            a2 = fmaf(d, xh, a2);
        }
        if (dgamma) { atomicAdd(&dgamma[c], a2); atomicAdd(&dbeta[c], a1); }
        if (gsum) {
            atomicAdd(&gacc[2 * gi], (double)(ga * a1));
            atomicAdd(&gacc[2 * gi + 1], (double)(ga * a2));
        }
    }
    if (gsum) {
        __syncthreads();
        double* st = gsum + (long long)(stats_per_shape ? b : 0) * G * 2;
        for (int t = threadIdx.x; t < 2 * G; t += NT) atomicAdd(&st[t], gacc[t]);
    }
}

// CSR transpose of the kNN graph: cnt[b][j] = in-degree, off = exclusive scan, rev[off[j]..] = sources i
__global__ void csr_count_kernel(const int* __restrict__ idx, long long total, int N, int k, int* __restrict__ cnt) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    long long b = e / ((long long)N * k);
    atomicAdd(&cnt[b * N + idx[e]], 1);
}
// one CTA (1024 threads) per shape; off has N+1 entries per shape; cursor := off (copy used by fill)
__global__ void __launch_bounds__(1024) csr_scan_kernel(const int* __restrict__ cnt, int N, int* __restrict__ off,
                                                        int* __restrict__ cursor) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < N) ? cnt[(long long)b * N + i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += y;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + (warp > 0 ? wsum[warp - 1] : 0) + x - v;
        if (i < N) {
            off[(long long)b * (N + 1) + i] = excl;
            cursor[(long long)b * N + i] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[(long long)b * (N + 1) + N] = carry_s;
}
__global__ void csr_fill_kernel(const int* __restrict__ idx, long long total, int N, int k, int* __restrict__ cursor,
                                int* __restrict__ rev) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    long long b = e / ((long long)N * k);
    int i = (int)((e / k) % N);
    int pos = atomicAdd(&cursor[b * N + idx[e]], 1);
    rev[b * (long long)N * k + pos] = i;
}

// dense part:  dPQ[b][j][0:Cout]   = -(rstd/M) (cnt_j (s1 + s2 rstd (P_j - mu)) + s2 rstd SQ_j)      (SQ_j = sum_{i->j} Q_i)
//              dPQ[b][i][Cout:2C]  =  rstd gamma dy - (rstd/M)(k s1 + s2 rstd (esum_i - k mu))
// one warp per point; `dense` = 0 (frozen BatchNorm) drops the statistics terms.
template <int V>
__global__ void __launch_bounds__(NT) edge_bwd_dense_kernel(const float* __restrict__ PQ, long long ldpq,
                                                            const float* __restrict__ dy,
                                                            const float* __restrict__ esum,
                                                            const int* __restrict__ off, const int* __restrict__ rev,
                                                            const float* __restrict__ mean_rstd,
                                                            const double* __restrict__ gsum,
                                                            const float* __restrict__ scale, int N, int k, int Cout,
                                                            int G, int stats_per_shape, double count, int dense,
                                                            float* __restrict__ dPQ, long long lddpq) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * (NT / 32) + warp;
    if (j >= N) return;
    const int cpg = Cout / G;
    const float* PQb = PQ + (long long)b * N * ldpq;
    const int* offb = off + (long long)b * (N + 1);
    const int* revb = rev + (long long)b * N * k;
    const int beg = dense ? offb[j] : 0, end = dense ? offb[j + 1] : 0;
    const float cntj = (float)(end - beg);
    float* out = dPQ + ((long long)b * N + j) * lddpq;
    for (int cb = 0; cb < Cout; cb += 32 * V) {
        const int c0 = cb + lane * V;
        if (c0 >= Cout) continue;
        float sq[V];
#pragma unroll
        for (int v = 0; v < V; ++v) sq[v] = 0.f;
        for (int t0 = beg; t0 < end; t0 += 32) {
            int myi = (t0 + lane < end) ? revb[t0 + lane] : 0;
            const int tn = min(32, end - t0);
#pragma unroll 4
            for (int t = 0; t < tn; ++t) {
                int i = __shfl_sync(FULL, myi, t);
                const float* qr = PQb + (long long)i * ldpq + Cout + c0;
#pragma unroll
                for (int v = 0; v < V; ++v) sq[v] += qr[v];
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            int c = c0 + v;
            int gi = c / cpg;
            long long si = ((long long)(stats_per_shape ? b : 0) * G + gi) * 2;
            float sc = scale[(long long)b * Cout + c];          // gamma * rstd
            float d = dy[((long long)b * N + j) * Cout + c];
            float dP = 0.f, dQ = sc * d;
            if (dense) {
                float mean = mean_rstd[si], rstd = mean_rstd[si + 1];
                float s1 = (float)(gsum[si] / count), s2 = (float)(gsum[si + 1] / count);
                float p = PQb[(long long)j * ldpq + c];
                dP = -rstd * (cntj * (s1 + s2 * rstd * (p - mean)) + s2 * rstd * sq[v]);
                float es = esum[((long long)b * N + j) * Cout + c];
                dQ -= rstd * ((float)k * s1 + s2 * rstd * (es - (float)k * mean));
            }
            out[c] = dP;
            out[Cout + c] = dQ;
        }
    }
}

// sparse part: dPQ[b][jsel][c] += scale[b][c] * dy[b][i][c]
__global__ void edge_bwd_scatter_kernel(const float* __restrict__ dy, const int* __restrict__ jsel,
                                        const float* __restrict__ scale, long long rows_total, int N, int Cout,
                                        float* __restrict__ dPQ, long long lddpq) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows_total * Cout) return;
    long long row = e / Cout;
    int c = (int)(e % Cout);
    long long b = row / N;
    float v = scale[b * Cout + c] * dy[e];
    if (v != 0.f) atomicAdd(&dPQ[(b * N + jsel[e]) * lddpq + c], v);
}

}  // namespace edge
}  // namespace pn

using namespace pn;
using namespace pn::edge;

extern "C" int pn_edge_gather_fwd(const float* PQ, long long ldpq, const int* idx, int B, int N, int k, int Cout,
                                  const float* gamma, float* esel, int* jsel, float* esum, double* stats, int G,
                                  int stats_per_shape, void* stream) {
    PN_REQUIRE(PQ && idx && gamma && esel && jsel && esum, "pn_edge_gather_fwd: null pointer");
    PN_REQUIRE(Cout % 2 == 0 && G > 0 && Cout % G == 0, "pn_edge_gather_fwd: Cout=%d G=%d", Cout, G);
    dim3 grid(cdiv(N, (NT / 32) * PTS_PER_WARP), B);
    size_t sm = sizeof(double) * 2 * G;
    bool v4 = (Cout % 4 == 0) && (ldpq % 4 == 0) && ((reinterpret_cast<uintptr_t>(PQ) & 15u) == 0) && Cout >= 128;
    bool v2 = (ldpq % 2 == 0) && ((reinterpret_cast<uintptr_t>(PQ) & 7u) == 0);
    PN_REQUIRE(v4 || v2, "pn_edge_gather_fwd: PQ must be 8-byte aligned with even pitch");
    if (v4)
        edge_gather_kernel<4><<<grid, NT, sm, (cudaStream_t)stream>>>(PQ, ldpq, idx, N, k, Cout, gamma, esel, jsel,
                                                                      esum, stats, G, stats_per_shape);
    else
        edge_gather_kernel<2><<<grid, NT, sm, (cudaStream_t)stream>>>(PQ, ldpq, idx, N, k, Cout, gamma, esel, jsel,
                                                                      esum, stats, G, stats_per_shape);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("edge_gather_kernel");
    return PN_OK;
}

extern "C" int pn_edge_apply(const float* esel, const float* scale, const float* shift, float* out, long long ldo,
                             int B, int N, int Cout, void* stream) {
    PN_REQUIRE(esel && scale && shift && out, "pn_edge_apply: null pointer");
    long long rows = (long long)B * N;
    edge_apply_kernel<<<cdiv(rows * Cout, 256), 256, 0, (cudaStream_t)stream>>>(esel, scale, shift, out, ldo, rows, N,
                                                                                Cout);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("edge_apply_kernel");
    return PN_OK;
}

extern "C" int pn_edge_bwd_prep(const float* g, long long ldg, const float* esel, const float* scale,
                                const float* shift, const float* mean_rstd, const float* gamma, int B, int N,
                                int Cout, int G, int stats_per_shape, float* dy, double* gsum, float* dgamma,
                                float* dbeta, void* stream) {
    PN_REQUIRE(g && esel && scale && shift && mean_rstd && gamma && dy, "pn_edge_bwd_prep: null pointer");
    dim3 grid(cdiv(N, 64), B);
    edge_bwd_prep_kernel<<<grid, NT, sizeof(double) * 2 * G, (cudaStream_t)stream>>>(
        g, ldg, esel, scale, shift, mean_rstd, gamma, N, Cout, G, stats_per_shape, dy, gsum, dgamma, dbeta);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("edge_bwd_prep_kernel");
    return PN_OK;
}

// cnt must be zero-filled [B][N]; off [B][N+1]; cursor [B][N]; rev [B][N*k]
extern "C" int pn_knn_csr_transpose(const int* idx, int B, int N, int k, int* cnt, int* off, int* cursor, int* rev,
                                    void* stream) {
    PN_REQUIRE(idx && cnt && off && cursor && rev, "pn_knn_csr_transpose: null pointer");
    long long total = (long long)B * N * k;
    cudaStream_t st = (cudaStream_t)stream;
    csr_count_kernel<<<cdiv(total, 256), 256, 0, st>>>(idx, total, N, k, cnt);
    PN_COUNT_LAUNCH();
    csr_scan_kernel<<<B, 1024, 0, st>>>(cnt, N, off, cursor);
    PN_COUNT_LAUNCH();
    csr_fill_kernel<<<cdiv(total, 256), 256, 0, st>>>(idx, total, N, k, cursor, rev);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("csr transpose");
    return PN_OK;
}

extern "C" int pn_edge_bwd(const float* PQ, long long ldpq, const float* dy, const float* esum, const int* jsel,
                           const int* off, const int* rev, const float* mean_rstd, const double* gsum,
                           const float* scale, int B, int N, int k, int Cout, int G, int stats_per_shape,
                           double count, int dense, float* dPQ, long long lddpq, void* stream) {
    PN_REQUIRE(PQ && dy && jsel && scale && dPQ, "pn_edge_bwd: null pointer");
    PN_REQUIRE(!dense || (esum && off && rev && mean_rstd && gsum), "pn_edge_bwd: dense needs csr + stats");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(cdiv(N, NT / 32), B);
    if (Cout % 128 == 0)
        edge_bwd_dense_kernel<4><<<grid, NT, 0, st>>>(PQ, ldpq, dy, esum, off, rev, mean_rstd, gsum, scale, N, k, Cout,
                                                      G, stats_per_shape, count, dense, dPQ, lddpq);
    else
        edge_bwd_dense_kernel<2><<<grid, NT, 0, st>>>(PQ, ldpq, dy, esum, off, rev, mean_rstd, gsum, scale, N, k, Cout,
                                                      G, stats_per_shape, count, dense, dPQ, lddpq);
    PN_COUNT_LAUNCH();
    long long rows = (long long)B * N;
    edge_bwd_scatter_kernel<<<cdiv(rows * Cout, 256), 256, 0, st>>>(dy, jsel, scale, rows, N, Cout, dPQ, lddpq);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("edge_bwd kernels");
    return PN_OK;
}
