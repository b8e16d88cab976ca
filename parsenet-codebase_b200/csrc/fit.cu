// Weighted polynomial moments of a point set for ALL segments of a shape in one pass (primitive fitting).
//
// Replaces the per-segment python loop of Fit.fit_{plane,sphere,cylinder,cone}_torch
//   src/primitive_forward.py:708-843 (+ LeastSquares.lstsq / CustomSVD, src/fitting_utils.py:36-65,420-455)
// which runs torch.svd / qr / matrix_rank (MAGMA, CPU-hybrid) on an (m x 3) matrix per segment plus host syncs
// (np.linalg.cond).  Every quantity those fits need is a function of the moments below:
//   plane   : smallest eigenvector of  sum w^2 (p-c)(p-c)^T ,  c = sum w p / sum w ;  d = a . c
//   sphere  : normal equations of the linearised fit (4 sum w^2 (c-p)(c-p)^T) x = 2 sum w^2 (c-p)(w|p|^2 - nu)
//   cylinder: axis = smallest eigenvector of sum w^2 n n^T ; sphere fit of the projected points (moments of P p are
//             linear images of the moments of p; third-order tensor sum w^3 p p p covers the |P p|^2 P p term)
//   cone    : apex from (sum w^2 n n^T) c = sum w^2 (n.p) n ; axis = plane fit of the normals
// so the O(m) streaming work is this kernel (double accumulation), and the 3x3 algebra runs on (S,3,3) tensors.
// The backward is the exact derivative of every moment w.r.t. the membership weights.
#include "common.cuh"

namespace pn {
namespace fit {

constexpr int NM = 55;          // number of moments per segment (layout documented in pnb200/fitting.py)
constexpr int PTS = 32;         // points per CTA
constexpr int SEG_T = 64;       // threads per CTA = max segments per launch


// phi_k(p, n) and the power of w multiplying it
__device__ __forceinline__ void eval_phi(const float* p, const float* n, float* phi) {
    const float x = p[0], y = p[1], z = p[2], a = n[0], b = n[1], c = n[2];
    const float np_ = a * x + b * y + c * z;
    phi[0] = 1.f; phi[1] = x; phi[2] = y; phi[3] = z;
    phi[4] = x * x; phi[5] = x * y; phi[6] = x * z; phi[7] = y * y; phi[8] = y * z; phi[9] = z * z;
    phi[10] = a; phi[11] = b; phi[12] = c;
    phi[13] = 1.f; phi[14] = x; phi[15] = y; phi[16] = z;
    phi[17] = phi[4]; phi[18] = phi[5]; phi[19] = phi[6]; phi[20] = phi[7]; phi[21] = phi[8]; phi[22] = phi[9];
    phi[23] = a; phi[24] = b; phi[25] = c;
    phi[26] = a * a; phi[27] = a * b; phi[28] = a * c; phi[29] = b * b; phi[30] = b * c; phi[31] = c * c;
    phi[32] = np_ * a; phi[33] = np_ * b; phi[34] = np_ * c;
    phi[35] = phi[4]; phi[36] = phi[5]; phi[37] = phi[6]; phi[38] = phi[7]; phi[39] = phi[8]; phi[40] = phi[9];
    phi[41] = x * x * x; phi[42] = x * x * y; phi[43] = x * x * z; phi[44] = x * y * y; phi[45] = x * y * z;
    phi[46] = x * z * z; phi[47] = y * y * y; phi[48] = y * y * z; phi[49] = y * z * z; phi[50] = z * z * z;
    phi[51] = a; phi[52] = b; phi[53] = c; phi[54] = 1.f;
}
__device__ __forceinline__ int wdeg(int k) { return k < 13 ? 1 : (k < 35 ? 2 : (k < 51 ? 3 : 0)); }

// P, Nr: [N][3]; W: [N][ldw] (column s = segment); point set: n = start + i*step, i < m; w = W[n][s] + eps
// mom: [S][NM] double (zero-initialised)
// blockIdx.y = shape of a batch: P / Nr advance by sP floats, W by sW floats, mom by S * NM doubles per shape
__global__ void __launch_bounds__(SEG_T) moments_fwd_kernel(const float* __restrict__ P, const float* __restrict__ Nr,
                                                            const float* __restrict__ W, long long ldw, int S,
                                                            int start, int step, int m, float eps,
                                                            double* __restrict__ mom, long long sP, long long sW) {
    __shared__ float sp[PTS][3], sn[PTS][3];
    P += blockIdx.y * sP;
    if (Nr) Nr += blockIdx.y * sP;
    W += blockIdx.y * sW;
    mom += (long long)blockIdx.y * S * NM;
    const int s = threadIdx.x;
    const int i0 = blockIdx.x * PTS;
    const int cnt = min(PTS, m - i0);
    for (int e = threadIdx.x; e < cnt * 3; e += SEG_T) {
        int i = e / 3, c = e % 3;
        long long n = start + (long long)(i0 + i) * step;
        sp[i][c] = P[n * 3 + c];
        sn[i][c] = Nr ? Nr[n * 3 + c] : 0.f;
    }
    __syncthreads();
    if (s >= S) return;
    float acc[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) acc[k] = 0.f;
    for (int i = 0; i < cnt; ++i) {
        long long n = start + (long long)(i0 + i) * step;
        const float w = W[n * ldw + s] + eps;
        const float w2 = w * w, w3 = w2 * w;
        float phi[NM];
        eval_phi(sp[i], sn[i], phi);
#pragma unroll
        for (int k = 0; k < NM; ++k) {
            const float wk = k < 13 ? w : (k < 35 ? w2 : (k < 51 ? w3 : 1.f));
            acc[k] = fmaf(wk, phi[k], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < NM; ++k) atomicAdd(&mom[(long long)s * NM + k], (double)acc[k]);
}

// gW[n][s] = sum_k gmom[s][k] * deg_k * w^(deg_k - 1) * phi_k     (rows outside the point set are left untouched)
__global__ void __launch_bounds__(SEG_T) moments_bwd_kernel(const float* __restrict__ P, const float* __restrict__ Nr,
                                                            const float* __restrict__ W, long long ldw, int S,
                                                            int start, int step, int m, float eps,
                                                            const float* __restrict__ gmom,
                                                            float* __restrict__ gW, long long ldg, long long sP,
                                                            long long sW, long long sG) {
    __shared__ float sp[PTS][3], sn[PTS][3];
    P += blockIdx.y * sP;
    if (Nr) Nr += blockIdx.y * sP;
    W += blockIdx.y * sW;
    gW += blockIdx.y * sG;
    gmom += (long long)blockIdx.y * S * NM;
    const int s = threadIdx.x;
    const int i0 = blockIdx.x * PTS;
    const int cnt = min(PTS, m - i0);
    for (int e = threadIdx.x; e < cnt * 3; e += SEG_T) {
        int i = e / 3, c = e % 3;
        long long n = start + (long long)(i0 + i) * step;
        sp[i][c] = P[n * 3 + c];
        sn[i][c] = Nr ? Nr[n * 3 + c] : 0.f;
    }
    __syncthreads();
    if (s >= S) return;
    float g[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) g[k] = gmom[(long long)s * NM + k];
    for (int i = 0; i < cnt; ++i) {
        long long n = start + (long long)(i0 + i) * step;
        const float w = W[n * ldw + s] + eps;
        float phi[NM];
        eval_phi(sp[i], sn[i], phi);
        float a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int k = 0; k < 13; ++k) a1 = fmaf(g[k], phi[k], a1);
#pragma unroll
        for (int k = 13; k < 35; ++k) a2 = fmaf(g[k], phi[k], a2);
#pragma unroll
        for (int k = 35; k < 51; ++k) a3 = fmaf(g[k], phi[k], a3);
        gW[n * ldg + s] = a1 + 2.f * w * a2 + 3.f * w * w * a3;
    }
}

}  // namespace fit
}  // namespace pn

using namespace pn;

extern "C" int pn_fit_moments_fwd(const float* P, const float* Nr, const float* W, long long ldw, int S, int start,
                                  int step, int m, float eps, double* mom_zeroed, void* stream) {
    PN_REQUIRE(P && W && mom_zeroed, "pn_fit_moments_fwd: null pointer");
    PN_REQUIRE(S > 0 && S <= fit::SEG_T, "pn_fit_moments_fwd: 1 <= segments <= %d (got %d)", fit::SEG_T, S);
    if (m <= 0) return PN_OK;
    fit::moments_fwd_kernel<<<cdiv(m, fit::PTS), fit::SEG_T, 0, (cudaStream_t)stream>>>(P, Nr, W, ldw, S, start, step, m,
                                                                                        eps, mom_zeroed, 0, 0);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("fit moments_fwd_kernel");
    return PN_OK;
}

extern "C" int pn_fit_moments_bwd(const float* P, const float* Nr, const float* W, long long ldw, int S, int start,
                                  int step, int m, float eps, const float* gmom, float* gW, long long ldg,
                                  void* stream) {
    PN_REQUIRE(P && W && gmom && gW, "pn_fit_moments_bwd: null pointer");
    PN_REQUIRE(S > 0 && S <= fit::SEG_T, "pn_fit_moments_bwd: 1 <= segments <= %d (got %d)", fit::SEG_T, S);
    if (m <= 0) return PN_OK;
    fit::moments_bwd_kernel<<<cdiv(m, fit::PTS), fit::SEG_T, 0, (cudaStream_t)stream>>>(P, Nr, W, ldw, S, start, step, m,
                                                                                        eps, gmom, gW, ldg, 0, 0, 0);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("fit moments_bwd_kernel");
    return PN_OK;
}

// Batched over the shapes of a step (one launch instead of one per shape): P / Nr [B][N][3], W [B][N][ldw] (column = segment
// slot), mom [B][S][NM] zero-initialised; gW [B][N][ldg] (rows outside the point set are left untouched).
extern "C" int pn_fit_moments_fwd_batched(const float* P, const float* Nr, const float* W, long long ldw, int B, int N, int S,
                                          int start, int step, int m, float eps, double* mom_zeroed, void* stream) {
    PN_REQUIRE(P && W && mom_zeroed, "pn_fit_moments_fwd_batched: null pointer");
    PN_REQUIRE(S > 0 && S <= fit::SEG_T, "pn_fit_moments_fwd_batched: 1 <= segments <= %d (got %d)", fit::SEG_T, S);
    PN_REQUIRE(B > 0 && N > 0 && start >= 0 && step > 0 && start + (long long)(m - 1) * step < N,
               "pn_fit_moments_fwd_batched: point set outside the shape (B=%d N=%d start=%d step=%d m=%d)", B, N, start, step, m);
    if (m <= 0) return PN_OK;
    fit::moments_fwd_kernel<<<dim3(cdiv(m, fit::PTS), B), fit::SEG_T, 0, (cudaStream_t)stream>>>(
        P, Nr, W, ldw, S, start, step, m, eps, mom_zeroed, (long long)N * 3, (long long)N * ldw);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("fit moments_fwd_kernel (batched)");
    return PN_OK;
}

extern "C" int pn_fit_moments_bwd_batched(const float* P, const float* Nr, const float* W, long long ldw, int B, int N, int S,
                                          int start, int step, int m, float eps, const float* gmom, float* gW,
                                          long long ldg, void* stream) {
    PN_REQUIRE(P && W && gmom && gW, "pn_fit_moments_bwd_batched: null pointer");
    PN_REQUIRE(S > 0 && S <= fit::SEG_T, "pn_fit_moments_bwd_batched: 1 <= segments <= %d (got %d)", fit::SEG_T, S);
    PN_REQUIRE(B > 0 && N > 0 && start >= 0 && step > 0 && start + (long long)(m - 1) * step < N,
               "pn_fit_moments_bwd_batched: point set outside the shape (B=%d N=%d start=%d step=%d m=%d)", B, N, start, step, m);
    if (m <= 0) return PN_OK;
    fit::moments_bwd_kernel<<<dim3(cdiv(m, fit::PTS), B), fit::SEG_T, 0, (cudaStream_t)stream>>>(
        P, Nr, W, ldw, S, start, step, m, eps, gmom, gW, ldg, (long long)N * 3, (long long)N * ldw, (long long)N * ldg);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("fit moments_bwd_kernel (batched)");
    return PN_OK;
}
