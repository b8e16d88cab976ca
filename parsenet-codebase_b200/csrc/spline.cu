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
template <typename T>
__global__ void __launch_bounds__(NT) eval_fwd_kernel(const T* __restrict__ Nu, const T* __restrict__ Nv,
                                                      const T* __restrict__ P, int gu, int gv, int cu, int cv,
                                                      T* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* sP = reinterpret_cast<T*>(sm_raw);  // [cu][cv][3]
    T* sT = sP + cu * cv * 3;              // [gu][cv][3]
    T* sNu = sT + gu * cv * 3;             // [gu][cu]
    T* sNv = sNu + gu * cu;                // [gv][cv]
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < cu * cv * 3; e += NT) sP[e] = P[(long long)b * cu * cv * 3 + e];
    for (int e = threadIdx.x; e < gu * cu; e += NT) sNu[e] = Nu[e];
    for (int e = threadIdx.x; e < gv * cv; e += NT) sNv[e] = Nv[e];
    __syncthreads();
    for (int e = threadIdx.x; e < gu * cv * 3; e += NT) {
        int u = e / (cv * 3), r = e % (cv * 3);
        T acc = T(0);
        for (int i = 0; i < cu; ++i) acc = fma(sNu[u * cu + i], sP[i * cv * 3 + r], acc);
        sT[e] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < gu * gv * 3; e += NT) {
        int u = e / (gv * 3), v = (e / 3) % gv, c = e % 3;
        T acc = T(0);
        for (int j = 0; j < cv; ++j) acc = fma(sT[(u * cv + j) * 3 + c], sNv[v * cv + j], acc);
        out[(long long)b * gu * gv * 3 + e] = acc;
    }
}

// dP[b][i][j][c] = sum_u sum_v Nu[u][i] g[b][u][v][c] Nv[v][j]
template <typename T>
__global__ void __launch_bounds__(NT) eval_bwd_kernel(const T* __restrict__ Nu, const T* __restrict__ Nv,
                                                      const T* __restrict__ g, int gu, int gv, int cu, int cv,
                                                      T* __restrict__ dP) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* sG = reinterpret_cast<T*>(sm_raw);  // [gu][gv][3]
    T* sT = sG + gu * gv * 3;              // [gu][cv][3]
    T* sNu = sT + gu * cv * 3;
    T* sNv = sNu + gu * cu;
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < gu * gv * 3; e += NT) sG[e] = g[(long long)b * gu * gv * 3 + e];
    for (int e = threadIdx.x; e < gu * cu; e += NT) sNu[e] = Nu[e];
    for (int e = threadIdx.x; e < gv * cv; e += NT) sNv[e] = Nv[e];
    __syncthreads();
    for (int e = threadIdx.x; e < gu * cv * 3; e += NT) {
        int u = e / (cv * 3), j = (e / 3) % cv, c = e % 3;
        T acc = T(0);
        for (int v = 0; v < gv; ++v) acc = fma(sG[(u * gv + v) * 3 + c], sNv[v * cv + j], acc);
        sT[e] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < cu * cv * 3; e += NT) {
        int i = e / (cv * 3), r = e % (cv * 3);
        T acc = T(0);
        for (int u = 0; u < gu; ++u) acc = fma(sNu[u * cu + i], sT[u * cv * 3 + r], acc);
        dP[(long long)b * cu * cv * 3 + e] = acc;
    }
}

}  // namespace spline
}  // namespace pn

using namespace pn;

namespace {
template <typename T>
int spline_launch(bool bwd, const T* Nu, const T* Nv, const T* in, int B, int gu, int gv, int cu, int cv, T* out,
                  void* stream, const char* what) {
    PN_REQUIRE(Nu && Nv && in && out && B > 0, what);
    size_t sm = sizeof(T) * ((bwd ? gu * gv * 3 : cu * cv * 3) + gu * cv * 3 + gu * cu + gv * cv);
    PN_REQUIRE(sm <= 200 * 1024, "pn_spline_eval: grid too large for shared memory");
    auto kern = bwd ? spline::eval_bwd_kernel<T> : spline::eval_fwd_kernel<T>;
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<B, spline::NT, sm, (cudaStream_t)stream>>>(Nu, Nv, in, gu, gv, cu, cv, out);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK(what);
    return PN_OK;
}
}  // namespace

extern "C" int pn_spline_eval_fwd(const float* Nu, const float* Nv, const float* P, int B, int gu, int gv, int cu,
                                  int cv, float* out, void* stream) {
    return spline_launch<float>(false, Nu, Nv, P, B, gu, gv, cu, cv, out, stream, "pn_spline_eval_fwd");
}

extern "C" int pn_spline_eval_bwd(const float* Nu, const float* Nv, const float* g, int B, int gu, int gv, int cu,
                                  int cv, float* dP, void* stream) {
    return spline_launch<float>(true, Nu, Nv, g, B, gu, gv, cu, cv, dP, stream, "pn_spline_eval_bwd");
}

// float64 instances: the control-point solve P = Nu^+ S (Nv^+)^T of approximation.fit_bezier_surface is float64 in the
// reference (numpy) and amplifies input rounding by ||Nu^+||.||Nv^+|| ~ 8e4, so it cannot be held to 1e-4 in fp32.
extern "C" int pn_spline_eval_fwd_f64(const double* Nu, const double* Nv, const double* P, int B, int gu, int gv,
                                      int cu, int cv, double* out, void* stream) {
    return spline_launch<double>(false, Nu, Nv, P, B, gu, gv, cu, cv, out, stream, "pn_spline_eval_fwd_f64");
}

extern "C" int pn_spline_eval_bwd_f64(const double* Nu, const double* Nv, const double* g, int B, int gu, int gv,
                                      int cu, int cv, double* dP, void* stream) {
    return spline_launch<double>(true, Nu, Nv, g, B, gu, gv, cu, cv, dP, stream, "pn_spline_eval_bwd_f64");
}
