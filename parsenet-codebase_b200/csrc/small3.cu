// Batched 3x3 algebra of the primitive fits (one problem per thread, fp64): symmetric eigen-decomposition and the
// reference's rank-guarded least-squares solve.  The math lives in small3.cuh (host/device shared).
// Replaces torch.svd / torch.matrix_rank / torch.qr + inverse on per-segment matrices:
//   CustomSVD src/fitting_utils.py:420-455, LeastSquares.lstsq + best_lambda src/fitting_utils.py:36-85.
#include "common.cuh"
#include "small3.cuh"

namespace pn {
namespace small3 {

__global__ void eigh3_kernel(const double* __restrict__ G, int S, double* __restrict__ w, double* __restrict__ V) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double a[9], ww[3], vv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = G[(long long)s * 9 + i];
    eigh3(a, ww, vv);
#pragma unroll
    for (int i = 0; i < 3; ++i) w[(long long)s * 3 + i] = ww[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) V[(long long)s * 9 + i] = vv[i];
}

__global__ void lstsq3_kernel(const double* __restrict__ AtA, const double* __restrict__ AtY, int S, int rows,
                              double eps32, double* __restrict__ x, double* __restrict__ minv,
                              double* __restrict__ lam) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double a[9], y[3], xx[3], mi[9], l;
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = AtA[(long long)s * 9 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = AtY[(long long)s * 3 + i];
    lstsq3(a, y, rows, eps32, xx, mi, &l);
#pragma unroll
    for (int i = 0; i < 3; ++i) x[(long long)s * 3 + i] = xx[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) minv[(long long)s * 9 + i] = mi[i];
    lam[s] = l;
}

}  // namespace small3
}  // namespace pn

using namespace pn;

extern "C" int pn_sym3_eigh(const double* G, int S, double* w_asc, double* V, void* stream) {
    PN_REQUIRE(G && w_asc && V && S > 0, "pn_sym3_eigh: bad arguments");
    small3::eigh3_kernel<<<cdiv(S, 64), 64, 0, (cudaStream_t)stream>>>(G, S, w_asc, V);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("eigh3_kernel");
    return PN_OK;
}

extern "C" int pn_lstsq3(const double* AtA, const double* AtY, int S, int rows, double eps32, double* x, double* Minv,
                         double* lam, void* stream) {
    PN_REQUIRE(AtA && AtY && x && Minv && lam && S > 0, "pn_lstsq3: bad arguments");
    small3::lstsq3_kernel<<<cdiv(S, 64), 64, 0, (cudaStream_t)stream>>>(AtA, AtY, S, rows, eps32, x, Minv, lam);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("lstsq3_kernel");
    return PN_OK;
}
