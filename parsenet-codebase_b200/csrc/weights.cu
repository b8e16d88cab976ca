// Membership weights of every cluster of every shape (weights_normalize, reference src/fitting_utils.py:306-325) as two
// launches forward / two backward instead of ~12 + ~25 elementwise / reduction kernels:
//   e = exp(clamp(raw / bw^2 / 2, +-75));  p = e / sum_k e   (over the K clusters of a point)
//   K == 1: out = p                           (early return :318-319)
//   else  : out = (p - min_n p) / (max_n (p - min_n p) + eps)    (min / max over the points of a cluster)
// Layout: raw / out [B][N][S] with S = 64 slots per shape, slots >= K[b] are padding (exact zeros out, no gradient).
// One warp per point row (2 slots per lane).  The extrema are kept as packed (ordered value, point index) 64-bit keys so that
// the backward routes the gradient of min / max to ONE point (lowest index among ties), deterministically.
#include "common.cuh"

namespace pn {
namespace wts {

constexpr int S = 64, NT = 256, ROWS = NT / 32;
constexpr float EPS32 = 1.1920928955078125e-07f;

__device__ __forceinline__ float prob_e(float raw, float bw2, bool live) {
    if (!live) return 0.f;
    float x = __fdiv_rn(__fdiv_rn(raw, bw2), 2.0f);
    x = fminf(fmaxf(x, -75.0f), 75.0f);
    return expf(x);
}

// prob [B][N][S] (written), mnkey / mxkey [B][S] (initialised to ~0ull / 0ull by the caller)
__global__ void __launch_bounds__(NT) wnorm_prob_kernel(const float* __restrict__ raw, const float* __restrict__ bw2,
                                                        const int* __restrict__ K, int N, float* __restrict__ prob,
                                                        unsigned long long* __restrict__ mnkey,
                                                        unsigned long long* __restrict__ mxkey) {
    __shared__ unsigned long long s_mn[S], s_mx[S];
    const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int kb = K[b];
    const float b2 = bw2[b];
    if (threadIdx.x < S) { s_mn[threadIdx.x] = ~0ull; s_mx[threadIdx.x] = 0ull; }
    __syncthreads();
    unsigned long long mn0 = ~0ull, mn1 = ~0ull, mx0 = 0ull, mx1 = 0ull;
    for (int n = blockIdx.x * ROWS + w; n < N; n += gridDim.x * ROWS) {
        const long long o = ((long long)b * N + n) * S;
        const float2 r = *reinterpret_cast<const float2*>(raw + o + 2 * lane);
        const float e0 = prob_e(r.x, b2, 2 * lane < kb), e1 = prob_e(r.y, b2, 2 * lane + 1 < kb);
        const float s = warp_sum(e0 + e1);
        const float p0 = __fdiv_rn(e0, s), p1 = __fdiv_rn(e1, s);
        *reinterpret_cast<float2*>(prob + o + 2 * lane) = make_float2(p0, p1);
        const unsigned long long k0 = (unsigned long long)f2ord(p0) << 32, k1 = (unsigned long long)f2ord(p1) << 32;
        const unsigned long long lo = (unsigned)n, hi = 0xffffffffu - (unsigned)n;
        mn0 = min(mn0, k0 | lo); mn1 = min(mn1, k1 | lo);
        mx0 = max(mx0, k0 | hi); mx1 = max(mx1, k1 | hi);
    }
    atomicMin(&s_mn[2 * lane], mn0); atomicMin(&s_mn[2 * lane + 1], mn1);
    atomicMax(&s_mx[2 * lane], mx0); atomicMax(&s_mx[2 * lane + 1], mx1);
    __syncthreads();
    if (threadIdx.x < S && threadIdx.x < kb) {
        atomicMin(&mnkey[b * S + threadIdx.x], s_mn[threadIdx.x]);
        atomicMax(&mxkey[b * S + threadIdx.x], s_mx[threadIdx.x]);
    }
}

__device__ __forceinline__ float key_val(unsigned long long k) { return ord2f((uint32_t)(k >> 32)); }

// in place: prob -> out
__global__ void __launch_bounds__(NT) wnorm_apply_kernel(float* __restrict__ prob, const int* __restrict__ K, int N,
                                                         const unsigned long long* __restrict__ mnkey,
                                                         const unsigned long long* __restrict__ mxkey) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int kb = K[b];
    if (kb == 1) return;                                    // single cluster: the probabilities are the weights
    float mn[2], den[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int k = 2 * lane + u;
        const bool live = k < kb;
        mn[u] = live ? key_val(mnkey[b * S + k]) : 0.f;
        const float mx = live ? key_val(mxkey[b * S + k]) : 0.f;
        den[u] = __fadd_rn(__fsub_rn(mx, mn[u]), EPS32);
    }
    for (int n = blockIdx.x * ROWS + w; n < N; n += gridDim.x * ROWS) {
        float2* p = reinterpret_cast<float2*>(prob + ((long long)b * N + n) * S + 2 * lane);
        float2 v = *p;
        v.x = __fdiv_rn(__fsub_rn(v.x, mn[0]), den[0]);
        v.y = __fdiv_rn(__fsub_rn(v.y, mn[1]), den[1]);
        *p = v;
    }
}

// backward, pass 1: per (shape, slot) A = sum_n g, Bs = sum_n g (p - mn)      (red [B][S][2] zero-initialised, doubles)
__global__ void __launch_bounds__(NT) wnorm_bwd_reduce_kernel(const float* __restrict__ raw, const float* __restrict__ g,
                                                              const float* __restrict__ bw2, const int* __restrict__ K,
                                                              int N, const unsigned long long* __restrict__ mnkey,
                                                              double* __restrict__ red) {
    __shared__ double s_red[S][2];
    const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int kb = K[b];
    if (kb == 1) return;
    const float b2 = bw2[b];
    if (threadIdx.x < S) { s_red[threadIdx.x][0] = 0.0; s_red[threadIdx.x][1] = 0.0; }
    __syncthreads();
    const float mn0 = (2 * lane < kb) ? key_val(mnkey[b * S + 2 * lane]) : 0.f;
    const float mn1 = (2 * lane + 1 < kb) ? key_val(mnkey[b * S + 2 * lane + 1]) : 0.f;
    double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
    for (int n = blockIdx.x * ROWS + w; n < N; n += gridDim.x * ROWS) {
        const long long o = ((long long)b * N + n) * S;
        const float2 r = *reinterpret_cast<const float2*>(raw + o + 2 * lane);
        const float2 gg = *reinterpret_cast<const float2*>(g + o + 2 * lane);
        const float e0 = prob_e(r.x, b2, 2 * lane < kb), e1 = prob_e(r.y, b2, 2 * lane + 1 < kb);
        const float s = warp_sum(e0 + e1);
        const float p0 = __fdiv_rn(e0, s), p1 = __fdiv_rn(e1, s);
        a0 += gg.x; a1 += gg.y;
        c0 += (double)gg.x * (double)(p0 - mn0); c1 += (double)gg.y * (double)(p1 - mn1);
    }
    atomicAdd(&s_red[2 * lane][0], a0); atomicAdd(&s_red[2 * lane][1], c0);
    atomicAdd(&s_red[2 * lane + 1][0], a1); atomicAdd(&s_red[2 * lane + 1][1], c1);
    __syncthreads();
    if (threadIdx.x < S && threadIdx.x < kb) {
        atomicAdd(&red[(b * S + threadIdx.x) * 2], s_red[threadIdx.x][0]);
        atomicAdd(&red[(b * S + threadIdx.x) * 2 + 1], s_red[threadIdx.x][1]);
    }
}

// backward, pass 2: graw [B][N][S]
__global__ void __launch_bounds__(NT) wnorm_bwd_apply_kernel(const float* __restrict__ raw, const float* __restrict__ g,
                                                             const float* __restrict__ bw2, const int* __restrict__ K,
                                                             int N, const unsigned long long* __restrict__ mnkey,
                                                             const unsigned long long* __restrict__ mxkey,
                                                             const double* __restrict__ red, float* __restrict__ graw) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int kb = K[b];
    const float b2 = bw2[b];
    const bool single = kb == 1;
    float den[2], gmn[2], gmx[2];
    unsigned imn[2], imx[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int k = 2 * lane + u;
        const bool live = (k < kb) && !single;
        den[u] = 1.f; gmn[u] = 0.f; gmx[u] = 0.f; imn[u] = 0xffffffffu; imx[u] = 0xffffffffu;
        if (live) {
            const unsigned long long kn = mnkey[b * S + k], kx = mxkey[b * S + k];
            const float mn = key_val(kn), mx = key_val(kx);
            den[u] = __fadd_rn(__fsub_rn(mx, mn), EPS32);
            const double A = red[(b * S + k) * 2], Bs = red[(b * S + k) * 2 + 1];
            const double d = (double)den[u];
            gmn[u] = (float)(-A / d + Bs / (d * d));
            gmx[u] = (float)(-Bs / (d * d));
            imn[u] = (unsigned)(kn & 0xffffffffu);
            imx[u] = 0xffffffffu - (unsigned)(kx & 0xffffffffu);
        }
    }
    for (int n = blockIdx.x * ROWS + w; n < N; n += gridDim.x * ROWS) {
        const long long o = ((long long)b * N + n) * S;
        const float2 r = *reinterpret_cast<const float2*>(raw + o + 2 * lane);
        const float2 gg = *reinterpret_cast<const float2*>(g + o + 2 * lane);
        const bool l0 = 2 * lane < kb, l1 = 2 * lane + 1 < kb;
        const float x0 = __fdiv_rn(__fdiv_rn(r.x, b2), 2.0f), x1 = __fdiv_rn(__fdiv_rn(r.y, b2), 2.0f);
        const float e0 = prob_e(r.x, b2, l0), e1 = prob_e(r.y, b2, l1);
        const float s = warp_sum(e0 + e1);
        const float p0 = __fdiv_rn(e0, s), p1 = __fdiv_rn(e1, s);
        float gp0, gp1;
        if (single) { gp0 = l0 ? gg.x : 0.f; gp1 = l1 ? gg.y : 0.f; }
        else {
            gp0 = l0 ? gg.x / den[0] + ((unsigned)n == imn[0] ? gmn[0] : 0.f) + ((unsigned)n == imx[0] ? gmx[0] : 0.f) : 0.f;
            gp1 = l1 ? gg.y / den[1] + ((unsigned)n == imn[1] ? gmn[1] : 0.f) + ((unsigned)n == imx[1] ? gmx[1] : 0.f) : 0.f;
        }
        const float dot = warp_sum(gp0 * p0 + gp1 * p1);
        // p = e / s: ge = (gp - dot) / s; e = exp(clamp(x)): passes where -75 <= x <= 75; x = raw / bw2 / 2
        float ge0 = (gp0 - dot) / s, ge1 = (gp1 - dot) / s;
        float gx0 = (l0 && x0 >= -75.0f && x0 <= 75.0f) ? ge0 * e0 : 0.f;
        float gx1 = (l1 && x1 >= -75.0f && x1 <= 75.0f) ? ge1 * e1 : 0.f;
        *reinterpret_cast<float2*>(graw + o + 2 * lane) = make_float2((gx0 / 2.0f) / b2, (gx1 / 2.0f) / b2);
    }
}

static int grid_x(int N) {
    int g = cdiv(N, ROWS);
    return g < 148 * 2 ? g : 148 * 2;
}

}  // namespace wts
}  // namespace pn

using namespace pn;

// raw [B][N][64] centre . point similarities, bw2 [B] squared bandwidths, K [B] clusters per shape -> out [B][N][64];
// keys [2][B][64] u64 workspace kept for the backward (first half must be all-ones, second half zero on entry)
extern "C" int pn_weights_normalize_fwd(const float* raw, const float* bw2, const int* K, int B, int N, int S, float* out,
                                        unsigned long long* keys, void* stream) {
    PN_REQUIRE(raw && bw2 && K && out && keys, "pn_weights_normalize_fwd: null pointer");
    PN_REQUIRE(S == wts::S && B > 0 && N > 0, "pn_weights_normalize_fwd: need S == %d slots, B, N > 0 (S=%d B=%d N=%d)", wts::S,
               S, B, N);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(wts::grid_x(N), B);
    wts::wnorm_prob_kernel<<<grid, wts::NT, 0, st>>>(raw, bw2, K, N, out, keys, keys + (size_t)B * S);
    PN_COUNT_LAUNCH();
    wts::wnorm_apply_kernel<<<grid, wts::NT, 0, st>>>(out, K, N, keys, keys + (size_t)B * S);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("wnorm kernels");
    return PN_OK;
}

// g [B][N][64] gradient w.r.t. out -> graw; red [B][64][2] doubles zero-initialised workspace
extern "C" int pn_weights_normalize_bwd(const float* raw, const float* g, const float* bw2, const int* K, int B, int N, int S,
                                        const unsigned long long* keys, double* red_zeroed, float* graw, void* stream) {
    PN_REQUIRE(raw && g && bw2 && K && keys && red_zeroed && graw, "pn_weights_normalize_bwd: null pointer");
    PN_REQUIRE(S == wts::S && B > 0 && N > 0, "pn_weights_normalize_bwd: need S == %d slots, B, N > 0 (S=%d B=%d N=%d)", wts::S,
               S, B, N);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(wts::grid_x(N), B);
    wts::wnorm_bwd_reduce_kernel<<<grid, wts::NT, 0, st>>>(raw, g, bw2, K, N, keys, red_zeroed);
    PN_COUNT_LAUNCH();
    wts::wnorm_bwd_apply_kernel<<<grid, wts::NT, 0, st>>>(raw, g, bw2, K, N, keys, keys + (size_t)B * S, red_zeroed, graw);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("wnorm backward kernels");
    return PN_OK;
}
