// kNN graph build: tiled fp32 distance + streaming top-k; the N x N matrix never exists in HBM.
//
// Replaces (reference, relative to /root/reference):
//   knn                 src/PointNet.py:9-26  and  src/model.py:9-22
//   knn_points_normals  src/PointNet.py:29-69
// which materialise a (B,N,N) fp32 matrix (6.4 GB at B=16,N=10^4) and run a full-row torch.topk.
//
// Arithmetic contract (bit-exact with oracle/c/knn_oracle.c): the ranked value is
//   metric 0:  D = (-xx_j - (-2*dot(x_i,x_j))) - xx_i
//   metric 1:  D = -(((xx_j - 2*dot(p_i,p_j)) + xx_i) * (1 + (2 - 2*dot(n_i,n_j))))
//   metric 2:  D = -(((dx*dx) + (dy*dy)) + (dz*dz)), d = x_i - x_j, C == 3, every op its own fp32 rounding
//              (up_sample_points_torch, reference src/fitting_utils.py:150-163: sum((p_i - p_j)**2, 2), 5 smallest)
// dot / xx are fmaf chains over channels in increasing order starting from +0; ranking is D descending,
// ties to the lower candidate index.  This stays on the FP32 FMA pipe on purpose: TF32/BF16 tensor-core
// products would change near-tie ordering (north_star: indices bit-exact).
//
// Kernel shape: one CTA = 64 query rows x all candidates, streamed in 128-wide tiles.
//   * 256 threads, each owns a 4(query) x 8(candidate) register tile; K is streamed in 32-channel chunks
//     through shared memory stored channel-major so the inner loop is 3 LDS.128 + 32 FFMA.
//   * the finished 64x128 tile of D goes through shared memory to 8 selection warps (8 rows each):
//     candidates that beat the row's current k-th best are appended to a per-row buffer (CAP entries);
//     when it would overflow the owning warp bitonic-sorts it in registers and keeps the best k.
//   * smem ~105 KB -> 2 CTAs/SM; grid = ceil(N/64) x B  (2512 CTAs at B=16,N=10^4 = 8.5 waves of 296).
#include "common.cuh"
#include "knn_select.cuh"
#include <stdlib.h>

namespace pn {
namespace knn {

constexpr int TQ = 64;    // query rows per CTA
constexpr int TC = 128;   // candidates per tile
constexpr int KC = 32;    // channels per shared-memory chunk
constexpr int NT = 256;

__global__ void norms_kernel(const float* __restrict__ x, long long rows, int ld, int c0, int c1,
                             float* __restrict__ xx) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* p = x + r * ld;
    float acc = 0.f;
    for (int c = c0; c < c1; ++c) acc = fmaf(p[c], p[c], acc);
    xx[r] = acc;
}

// SAMPLED (default since round 2, PN_KNN_SAMPLE=0 switches it off; indices identical, 7.8 -> 6.8 ms at C = 6): a pre-pass over the first SAMPLE_M candidates
// (the clouds are randomly permuted, so this is a random sample) finds the exact r-th best of the sample, r ~ 3 k M / N
// (passed in bits 8.. of `vec`); the main pass then starts with that value as its admission threshold instead of -inf:
// ~230 instead of ~600 admitted candidates per row and ~3 instead of ~12 compactions (tools/exp_knn_cap.py header,
// DESIGN.md section 7).  Exact: every candidate better than the threshold is admitted, so the top k are complete
// whenever at least k candidates pass; a CTA in which some row admits fewer than k re-runs from -inf (third phase).
// With SAMPLED = false the phase logic compiles away (same SASS as before the option existed).
constexpr int SAMPLE_M = 1024;

template <int METRIC, int CAP, typename IdxT, bool SAMPLED = false>
__global__ void __launch_bounds__(NT, 2)
knn_kernel(const float* __restrict__ x, const float* __restrict__ xx, int N, int C, int ld, int k, int vec,
           IdxT* __restrict__ idx_out, float* __restrict__ dist_out, const int* __restrict__ flags) {
    // flags (optional, [B][N]): a CTA none of whose TQ query rows is flagged exits at once (exact fall-back of knn_lowdim.cu)
    if (flags) {
        const int q = blockIdx.x * TQ + threadIdx.x;
        const int f = (threadIdx.x < TQ && q < N) ? flags[(size_t)blockIdx.y * N + q] : 0;
        if (!__syncthreads_or(f)) return;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout
    float* xs = reinterpret_cast<float*>(smem_raw);           // [KC][TC]   (aliased by Dt [TQ][TC])
    float* Dt = xs;
    float* qs = xs + TQ * TC;                                  // [KC][TQ]
    float* xxq = qs + KC * TQ;                                 // [TQ]
    float* xxc = xxq + TQ;                                     // [TC]
    float* tau = xxc + TC;                                     // [TQ]
    int* cnt = reinterpret_cast<int*>(tau + TQ);               // [TQ]
    float* bufv = reinterpret_cast<float*>(cnt + TQ);          // [TQ][CAP]
    int* bufi = reinterpret_cast<int*>(bufv + TQ * CAP);       // [TQ][CAP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * TQ;
    const float* xb = x + (size_t)b * N * ld;
    const float* xxb = xx + (size_t)b * N;
    const int tq = tid >> 4;   // 0..15 -> queries 4*tq..4*tq+3
    const int tc = tid & 15;   // 0..15 -> candidates 4*tc..+3 and 64+4*tc..+3

    if (tid < TQ) {
        int q = q0 + tid;
        xxq[tid] = (q < N) ? xxb[q] : 0.f;
        tau[tid] = -INFINITY;
        cnt[tid] = 0;
    }
    const bool q_resident = (C <= KC);
    if (q_resident) {
        for (int e = tid; e < TQ * C; e += NT) {
            int p = e % TQ, c = e / TQ;
            int q = q0 + p;
            qs[c * TQ + p] = (q < N) ? xb[(size_t)q * ld + c] : 0.f;
        }
    }

    int jend = N, ksel = k;
    [[maybe_unused]] int phase = 1;
    if constexpr (SAMPLED) {
        phase = 0;
        ksel = vec >> 8;                      // r
        jend = min(N, SAMPLE_M);
        vec &= 1;
    }
phase_begin:
    for (int j0 = 0; j0 < jend; j0 += TC) {
        float acc[4][8];
        float accn[METRIC == 1 ? 4 : 1][METRIC == 1 ? 8 : 1];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;
        if (METRIC == 1) {
#pragma unroll
            for (int a = 0; a < (METRIC == 1 ? 4 : 1); ++a)
#pragma unroll
                for (int c = 0; c < (METRIC == 1 ? 8 : 1); ++c) accn[a][c] = 0.f;
        }

        for (int c0 = 0; c0 < C; c0 += KC) {
            const int kc = min(KC, C - c0);
            __syncthreads();   // previous readers of xs / Dt / qs are done
            if (vec) {
                // 16-byte loads along the channels (rows are 16-byte aligned, kc % 4 == 0), transposed on the way into
                // the channel-major tile: consecutive lanes take consecutive rows, so the four scalar stores of a
                // float4 are conflict-free and every 32-byte sector fetched from L2 is consumed
                const int kc4 = kc >> 2;
                for (int e = tid; e < TC * kc4; e += NT) {
                    const int p = e % TC, g = e / TC;
                    const int j = j0 + p;
                    const float4 v = (j < N) ? *reinterpret_cast<const float4*>(xb + (size_t)j * ld + c0 + 4 * g)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    float* d = xs + (4 * g) * TC + p;
                    d[0] = v.x; d[TC] = v.y; d[2 * TC] = v.z; d[3 * TC] = v.w;
                }
                if (!q_resident) {
                    for (int e = tid; e < TQ * kc4; e += NT) {
                        const int p = e % TQ, g = e / TQ;
                        const int q = q0 + p;
                        const float4 v = (q < N) ? *reinterpret_cast<const float4*>(xb + (size_t)q * ld + c0 + 4 * g)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                        float* d = qs + (4 * g) * TQ + p;
                        d[0] = v.x; d[TQ] = v.y; d[2 * TQ] = v.z; d[3 * TQ] = v.w;
                    }
                }
            } else {
                for (int e = tid; e < TC * kc; e += NT) {
                    int p = e % TC, c = e / TC;
                    int j = j0 + p;
                    xs[c * TC + p] = (j < N) ? xb[(size_t)j * ld + c0 + c] : 0.f;
                }
                if (!q_resident) {
                    for (int e = tid; e < TQ * kc; e += NT) {
                        int p = e % TQ, c = e / TQ;
                        int q = q0 + p;
                        qs[c * TQ + p] = (q < N) ? xb[(size_t)q * ld + c0 + c] : 0.f;
                    }
                }
            }
            if (c0 == 0 && tid < TC) {
                int j = j0 + tid;
                xxc[tid] = (j < N) ? xxb[j] : 0.f;
            }
            __syncthreads();
            if constexpr (METRIC == 2) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float4 qv = *reinterpret_cast<const float4*>(qs + c * TQ + 4 * tq);
                    float4 xa = *reinterpret_cast<const float4*>(xs + c * TC + 4 * tc);
                    float4 xc = *reinterpret_cast<const float4*>(xs + c * TC + 64 + 4 * tc);
                    const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
                    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xc.x, xc.y, xc.z, xc.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) {
                            const float d = __fsub_rn(qa[a], xv[cc]);
                            acc[a][cc] = __fadd_rn(acc[a][cc], __fmul_rn(d, d));
                        }
                }
            } else if constexpr (METRIC == 0) {
#pragma unroll 8
                for (int c = 0; c < kc; ++c) {
                    float4 qv = *reinterpret_cast<const float4*>(qs + c * TQ + 4 * tq);
                    float4 xa = *reinterpret_cast<const float4*>(xs + c * TC + 4 * tc);
                    float4 xc = *reinterpret_cast<const float4*>(xs + c * TC + 64 + 4 * tc);
                    const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
                    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xc.x, xc.y, xc.z, xc.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) acc[a][cc] = fmaf(qa[a], xv[cc], acc[a][cc]);
                }
            } else {
                // C == 6: channels 0..2 positions, 3..5 normals (single chunk)
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    float4 qv = *reinterpret_cast<const float4*>(qs + c * TQ + 4 * tq);
                    float4 xa = *reinterpret_cast<const float4*>(xs + c * TC + 4 * tc);
                    float4 xc = *reinterpret_cast<const float4*>(xs + c * TC + 64 + 4 * tc);
                    const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
                    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xc.x, xc.y, xc.z, xc.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) {
                            if (c < 3) acc[a][cc] = fmaf(qa[a], xv[cc], acc[a][cc]);
                            else accn[a][cc] = fmaf(qa[a], xv[cc], accn[a][cc]);
                        }
                }
            }
        }
        __syncthreads();   // all reads of xs done -> reuse as Dt
        // ---- distances into the shared tile
        {
            float xq[4], xc8[8];
#pragma unroll
            for (int a = 0; a < 4; ++a) xq[a] = xxq[4 * tq + a];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) xc8[cc] = xxc[(cc < 4 ? 0 : 60) + 4 * tc + cc];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float dv[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    int j = j0 + (cc < 4 ? 0 : 60) + 4 * tc + cc;
                    float D;
                    if constexpr (METRIC == 2) {
                        D = -acc[a][cc];
                    } else if constexpr (METRIC == 0) {
                        float inner = __fmul_rn(-2.0f, acc[a][cc]);
                        D = __fsub_rn(__fsub_rn(-xc8[cc], inner), xq[a]);
                    } else {
                        float pd = __fadd_rn(__fsub_rn(xc8[cc], __fmul_rn(2.0f, acc[a][cc])), xq[a]);
                        float nd = __fsub_rn(2.0f, __fmul_rn(2.0f, accn[a][cc]));
                        D = -__fmul_rn(pd, __fadd_rn(1.0f, nd));
                    }
                    dv[cc] = (j < N) ? D : -INFINITY;
                }
                float* row = Dt + (4 * tq + a) * TC;
                *reinterpret_cast<float4*>(row + 4 * tc) = make_float4(dv[0], dv[1], dv[2], dv[3]);
                *reinterpret_cast<float4*>(row + 64 + 4 * tc) = make_float4(dv[4], dv[5], dv[6], dv[7]);
            }
        }
        __syncthreads();
        // ---- selection: warp w owns rows 8w..8w+7
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
            const int row = warp * 8 + r;
            float t = tau[row];
            int n = cnt[row];
            float* bv = bufv + row * CAP;
            int* bi = bufi + row * CAP;
#pragma unroll
            for (int ch = 0; ch < TC / 32; ++ch) {
                float d = Dt[row * TC + ch * 32 + lane];
                bool pass = d > t;
                unsigned m = __ballot_sync(FULL, pass);
                if (m) {
                    int c = __popc(m);
                    if (n + c > CAP) n = compact_select<CAP>(bv, bi, n, ksel, lane, &t);
                    if (pass) {
                        int pos = n + __popc(m & ((1u << lane) - 1u));
                        bv[pos] = d;
                        bi[pos] = j0 + ch * 32 + lane;
                    }
                    n += c;
                    __syncwarp();
                }
            }
            if (lane == 0) { tau[row] = t; cnt[row] = n; }
        }
        // next iteration starts with __syncthreads()
    }
    __syncwarp();
    if constexpr (SAMPLED) {
        if (phase == 0) {
            // threshold of the main pass = r-th best of the sample (-inf when the sample holds fewer than r candidates)
#pragma unroll 1
            for (int r = 0; r < 8; ++r) {
                const int row = warp * 8 + r;
                float t;
                compact_row<CAP>(bufv + row * CAP, bufi + row * CAP, cnt[row], ksel, lane, &t);
                if (lane == 0) { tau[row] = t; cnt[row] = 0; }
            }
            __syncwarp();
            phase = 1; ksel = k; jend = N;
            goto phase_begin;
        }
        if (phase == 1) {
            int short_rows = 0;
#pragma unroll 1
            for (int r = 0; r < 8; ++r) short_rows |= (cnt[warp * 8 + r] < k) ? 1 : 0;
            if (__syncthreads_or(short_rows)) {            // rare: some row admitted fewer than k -> exact re-run from -inf
#pragma unroll 1
                for (int r = 0; r < 8; ++r)
                    if (lane == 0) { tau[warp * 8 + r] = -INFINITY; cnt[warp * 8 + r] = 0; }
                __syncwarp();
                phase = 2;
                goto phase_begin;
            }
        }
    }
    // ---- final sort + write-out
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
        const int row = warp * 8 + r;
        const int q = q0 + row;
        float t;
        float* bv = bufv + row * CAP;
        int* bi = bufi + row * CAP;
        int n = compact_row<CAP>(bv, bi, cnt[row], k, lane, &t);
        if (q < N) {
            size_t o = ((size_t)b * N + q) * k;
            for (int p = lane; p < k; p += 32) {
                idx_out[o + p] = (IdxT)((p < n) ? bi[p] : 0);
                if (dist_out) dist_out[o + p] = (p < n) ? bv[p] : -INFINITY;
            }
        }
    }
}

static size_t smem_bytes(int cap) {
    return sizeof(float) * (TQ * TC + KC * TQ + TQ + TC + TQ) + sizeof(int) * TQ +
           (size_t)TQ * cap * (sizeof(float) + sizeof(int));
}

// sample rank r of the SAMPLED variant (0 = do not sample): needs a long row, k of the order of the graph degree and
// r < k; 3 k M / N admits ~2.9 k candidates on average with the minimum well above k (measured on synthetic clouds and
// feature spaces: r = 24 for k = 80, N = 10^4 admits 110 .. 380 per row)
static int sample_rank(int N, int k) {
    const char* e = getenv("PN_KNN_SAMPLE");
    if ((e && e[0] == '0') || N < 4 * SAMPLE_M || k < 32) return 0;      // default on since round 2 (PN_KNN_SAMPLE=0: off)
    const int r = (int)((3.0 * k * SAMPLE_M + N - 1) / N);
    return (r >= 8 && r < k) ? r : 0;
}

template <int METRIC, int CAP, typename IdxT>
static int launch(const float* x, const float* xx, int B, int N, int C, int ld, int k, void* idx, float* dist,
                  cudaStream_t st, const int* flags) {
    const int r = sample_rank(N, k);
    auto kern = r ? knn_kernel<METRIC, CAP, IdxT, true> : knn_kernel<METRIC, CAP, IdxT, false>;
    size_t sm = smem_bytes(CAP);
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid(cdiv(N, TQ), B);
    const int vec = ((C % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0)) | (r << 8);
    kern<<<grid, NT, sm, st>>>(x, xx, N, C, ld, k, vec, (IdxT*)idx, dist, flags);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn_kernel");
    return PN_OK;
}

template <int METRIC, typename IdxT>
static int dispatch_cap(const float* x, const float* xx, int B, int N, int C, int ld, int k, void* idx, float* dist,
                        cudaStream_t st, const int* flags) {
    // experiment knob (read per call): PN_KNN_CAP=256 gives a k = 80 row 176 instead of 48 entries of headroom between
    // compactions (3.7x fewer, 2x larger quickselects) at 1 instead of 2 resident CTAs per SM; results are identical
    // for every capacity >= k (the buffer content after a compaction does not depend on when it happens)
    int cap = (k <= 32) ? 64 : (k <= 96 ? 128 : 256);
    if (const char* e = getenv("PN_KNN_CAP")) {
        const int want = atoi(e);
        if ((want == 128 || want == 256) && want >= cap) cap = want;
    }
    if (cap == 64) return launch<METRIC, 64, IdxT>(x, xx, B, N, C, ld, k, idx, dist, st, flags);
    if (cap == 128) return launch<METRIC, 128, IdxT>(x, xx, B, N, C, ld, k, idx, dist, st, flags);
    return launch<METRIC, 256, IdxT>(x, xx, B, N, C, ld, k, idx, dist, st, flags);
}

}  // namespace knn
}  // namespace pn

using namespace pn;

static int knn_run(const char* who, const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out,
                   int idx_is_i64, float* dist_out, float* ws_norms, const int* flags, void* stream) {
    PN_REQUIRE(x && idx_out && ws_norms, "%s: null pointer", who);
    PN_REQUIRE(B > 0 && N > 0 && C > 0 && ld >= C, "%s: bad shape B=%d N=%d C=%d ld=%d", who, B, N, C, ld);
    PN_REQUIRE(k > 0 && k <= N && k <= 224, "%s: need 0 < k <= min(N,224), got k=%d N=%d", who, k, N);
    PN_REQUIRE(metric == 0 || (metric == 1 && C == 6) || (metric == 2 && C == 3),
               "%s: metric 1 needs C == 6, metric 2 needs C == 3", who);
    cudaStream_t st = (cudaStream_t)stream;
    long long rows = (long long)B * N;
    knn::norms_kernel<<<cdiv(rows, 256), 256, 0, st>>>(x, rows, ld, 0, metric == 1 ? 3 : C, ws_norms);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn norms_kernel");
    if (metric == 0) {
        return idx_is_i64 ? knn::dispatch_cap<0, long long>(x, ws_norms, B, N, C, ld, k, idx_out, dist_out, st, flags)
                          : knn::dispatch_cap<0, int>(x, ws_norms, B, N, C, ld, k, idx_out, dist_out, st, flags);
    }
    if (metric == 2) {
        return idx_is_i64 ? knn::dispatch_cap<2, long long>(x, ws_norms, B, N, C, ld, k, idx_out, dist_out, st, flags)
                          : knn::dispatch_cap<2, int>(x, ws_norms, B, N, C, ld, k, idx_out, dist_out, st, flags);
    }
    return idx_is_i64 ? knn::dispatch_cap<1, long long>(x, ws_norms, B, N, C, ld, k, idx_out, dist_out, st, flags)
                      : knn::dispatch_cap<1, int>(x, ws_norms, B, N, C, ld, k, idx_out, dist_out, st, flags);
}

extern "C" int pn_knn(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out,
                      int idx_is_i64, float* dist_out, float* ws_norms, void* stream) {
    return knn_run("pn_knn", x, B, N, C, ld, k, metric, idx_out, idx_is_i64, dist_out, ws_norms, nullptr, stream);
}

// the same graph for the tiles of query rows that contain a row with flags[b][row] != 0 only (other tiles exit at once and leave
// their rows of idx_out / dist_out untouched): exact fall-back of pn_knn_lowdim
extern "C" int pn_knn_flagged(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out,
                              int idx_is_i64, float* dist_out, float* ws_norms, const int* flags, void* stream) {
    PN_REQUIRE(flags, "pn_knn_flagged: null flags");
    return knn_run("pn_knn_flagged", x, B, N, C, ld, k, metric, idx_out, idx_is_i64, dist_out, ws_norms, flags, stream);
}
