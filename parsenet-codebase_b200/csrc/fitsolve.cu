// Parameters of every analytic primitive of a step from the moment table, with Jacobians, in ONE launch.
// One CTA of 64 threads per segment slot: thread j < 55 carries the tangent along moment j (forward mode, fitsolve.cuh),
// every thread recomputes the (tiny, fp64) primal; slots whose kind is -1 are zero-filled.
// Replaces Fit.fit_{plane,sphere,cylinder,cone}_torch + LeastSquares.lstsq + CustomSVD (reference
// src/primitive_forward.py:708-831, src/fitting_utils.py:36-85,385-455) for all segments of all shapes at once.
#include "common.cuh"
#include "fitsolve.cuh"

namespace pn {
namespace fitsolve {

__global__ void __launch_bounds__(64) fit_solve_kernel(const double* __restrict__ mom, const int* __restrict__ kind, int S,
                                                       int rows, double* __restrict__ par, double* __restrict__ jac,
                                                       float* __restrict__ bad) {
    const int s = blockIdx.x, j = threadIdx.x;
    if (s >= S) return;
    const int kd = kind[s];
    double* pj = jac + (long long)s * NPAR * NM;
    if (kd < 0 || kd > KIND_CONE) {
        for (int e = j; e < NPAR * NM; e += 64) pj[e] = 0.0;
        if (j < NPAR) par[(long long)s * NPAR + j] = 0.0;
        if (j == 0) bad[s] = 0.f;
        return;
    }
    Du m[NM], p[NPAR];
    const double* ms = mom + (long long)s * NM;
#pragma unroll 1
    for (int k = 0; k < NM; ++k) m[k] = mk(ms[k], k == j ? 1.0 : 0.0);
    const bool degenerate = solve_segment(kd, m, rows, p);
    if (j < NM)
        for (int i = 0; i < NPAR; ++i) pj[i * NM + j] = p[i].d;
    if (j == 63) {
        for (int i = 0; i < NPAR; ++i) par[(long long)s * NPAR + i] = p[i].v;
        bad[s] = degenerate ? 1.f : 0.f;
    }
}

}  // namespace fitsolve
}  // namespace pn

using namespace pn;

// mom [S][55] fp64 moments (pn_fit_moments_fwd), kind [S] in {-1 none, 0 plane, 1 sphere, 2 cylinder, 3 cone}, rows = number
// of points behind the moments (enters the rank rule of the least-squares solve) -> par [S][8], jac [S][8][55] (d par / d mom),
// bad [S] (1 = degenerate cone: constant parameters, zero Jacobian)
extern "C" int pn_fit_solve(const double* mom, const int* kind, int S, int rows, double* par, double* jac, float* bad,
                            void* stream) {
    PN_REQUIRE(mom && kind && par && jac && bad, "pn_fit_solve: null pointer");
    PN_REQUIRE(S > 0 && rows > 0, "pn_fit_solve: need S > 0 and rows > 0 (S=%d rows=%d)", S, rows);
    fitsolve::fit_solve_kernel<<<S, 64, 0, (cudaStream_t)stream>>>(mom, kind, S, rows, par, jac, bad);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("fit_solve_kernel");
    return PN_OK;
}
