// Tensor-product B-spline surface evaluation  S[b] = Nu . P[b] . Nv^T  per coordinate, one CTA per surface.
//
// Replaces sample_points_from_control_points_ src/fitting_utils.py:609-622 and the inlined copies in
// src/loss.py:161-165,179-183 (python loops issuing 6*B tiny GEMMs).  Nu (gu x cu), Nv (gv x cv) are the constant
// basis matrices of uniform_knot_bspline (src/loss.py:190); P is the (cu x cv x 3) control grid.
#include "common.cuh"

namespace pn {
namespace spline {

constexpr int NT = 256;

// out[b][u][v][c] = sum_i sum_j Nu[u][i] P[b][i][j][c] Nv[v][j]
__global__ void __launch_bounds__(NT) eval_fwd_kernel(const float* __restrict__ Nu, const float* __restrict__ Nv,
                                                      const float* __restrict__ P, int gu, int gv, int cu, int cv,
                                                      float* __restrict__ out) {
    extern __shared__ float sm[];
    float* sP = sm;                       // [cu][cv][3]
    float* sT = sP + cu * cv * 3;         // [gu][cv][3]
    float* sNu = sT + gu * cv * 3;        // [gu][cu]
    float* sNv = sNu + gu * cu;           // [gv][cv]
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < cu * cv * 3; e += NT) sP[e] = P[(long long)b * cu * cv * 3 + e];
    for (int e = threadIdx.x; e < gu * cu; e += NT) sNu[e] = Nu[e];
    for (int e = threadIdx.x; e < gv * cv; e += NT) sNv[e] = Nv[e];
    __syncthreads();
    for (int e = threadIdx.x; e < gu * cv * 3; e += NT) {
        int u = e / (cv * 3), r = e % (cv * 3);
        float acc = 0.f;
        for (int i = 0; i < cu; ++i) acc = fmaf(sNu[u * cu + i], sP[i * cv * 3 + r], acc);
        sT[e] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < gu * gv * 3; e += NT) {
        int u = e / (gv * 3), v = (e / 3) % gv, c = e % 3;
        float acc = 0.f;
        for (int j = 0; j < cv; ++j) acc = fmaf(sT[(u * cv + j) * 3 + c], sNv[v * cv + j], acc);
        out[(long long)b * gu * gv * 3 + e] = acc;
    }
}

// dP[b][i][j][c] = sum_u sum_v Nu[u][i] g[b][u][v][c] Nv[v][j]
__global__ void __launch_bounds__(NT) eval_bwd_kernel(const float* __restrict__ Nu, const float* __restrict__ Nv,
                                                      const float* __restrict__ g, int gu, int gv, int cu, int cv,
                                                      float* __restrict__ dP) {
    extern __shared__ float sm[];
    float* sG = sm;                       // [gu][gv][3]
    float* sT = sG + gu * gv * 3;         // [gu][cv][3]
    float* sNu = sT + gu * cv * 3;
    float* sNv = sNu + gu * cu;
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < gu * gv * 3; e += NT) sG[e] = g[(long long)b * gu * gv * 3 + e];
    for (int e = threadIdx.x; e < gu * cu; e += NT) sNu[e] = Nu[e];
    for (int e = threadIdx.x; e < gv * cv; e += NT) sNv[e] = Nv[e];
    __syncthreads();
    for (int e = threadIdx.x; e < gu * cv * 3; e += NT) {
        int u = e / (cv * 3), j = (e / 3) % cv, c = e % 3;
        float acc = 0.f;
        for (int v = 0; v < gv; ++v) acc = fmaf(sG[(u * gv + v) * 3 + c], sNv[v * cv + j], acc);
        sT[e] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < cu * cv * 3; e += NT) {
        int i = e / (cv * 3), r = e % (cv * 3);
        float acc = 0.f;
        for (int u = 0; u < gu; ++u) acc = fmaf(sNu[u * cu + i], sT[u * cv * 3 + r], acc);
        dP[(long long)b * cu * cv * 3 + e] = acc;
    }
}

}  // namespace spline
}  // namespace pn

using namespace pn;

extern "C" int pn_spline_eval_fwd(const float* Nu, const float* Nv, const float* P, int B, int gu, int gv, int cu,
                                  int cv, float* out, void* stream) {
    PN_REQUIRE(Nu && Nv && P && out && B > 0, "pn_spline_eval_fwd: bad args");
    size_t sm = sizeof(float) * (cu * cv * 3 + gu * cv * 3 + gu * cu + gv * cv);
    PN_REQUIRE(sm <= 200 * 1024, "pn_spline_eval_fwd: grid too large for shared memory");
    PN_CUDA(cudaFuncSetAttribute(spline::eval_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    spline::eval_fwd_kernel<<<B, spline::NT, sm, (cudaStream_t)stream>>>(Nu, Nv, P, gu, gv, cu, cv, out);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("spline eval_fwd_kernel");
    return PN_OK;
}

extern "C" int pn_spline_eval_bwd(const float* Nu, const float* Nv, const float* g, int B, int gu, int gv, int cu,
                                  int cv, float* dP, void* stream) {
    PN_REQUIRE(Nu && Nv && g && dP && B > 0, "pn_spline_eval_bwd: bad args");
    size_t sm = sizeof(float) * (gu * gv * 3 + gu * cv * 3 + gu * cu + gv * cv);
    PN_REQUIRE(sm <= 200 * 1024, "pn_spline_eval_bwd: grid too large for shared memory");
    PN_CUDA(cudaFuncSetAttribute(spline::eval_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    spline::eval_bwd_kernel<<<B, spline::NT, sm, (cudaStream_t)stream>>>(Nu, Nv, g, gu, gv, cu, cv, dP);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("spline eval_bwd_kernel");
    return PN_OK;
}
