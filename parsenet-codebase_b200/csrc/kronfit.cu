// Scattered-sample control-point fit of SURVEY 8f-1 (the post-fit optimisers of the inference path):
//   fit_bezier_surface_fit_kronecker   reference src/approximation.py:338-364  least squares on the rows A_i = u_i (x) v_i
//   surface evaluation at scattered parameters (geomdl evaluate_list in src/primitive_forward.py:186,258): S_i = sum_ab u_ia v_ib C_ab
// The reference stacks the M x (n m) Kronecker matrix on the host and calls numpy lstsq once per coordinate.  Here one CTA per
// surface assembles the (n m)-square normal equations G = A^T A, A^T P in shared memory (float64), factors G = L L^T in place
// and solves the three right-hand sides; a batch of surfaces is one launch.  A pivot below 1e-12 of the largest diagonal
// entry (rank-deficient sampling: numpy would return the minimum-norm solution) sets flag[s] and leaves ctrl[s] unwritten.
#include "common.cuh"

namespace pn {
namespace kron {

constexpr int NT = 256, CH = 32, MAXNM = 128, MAXE = MAXNM * MAXNM / NT;

// U [S][M][n], V [S][M][m], P [S][M][3] -> ctrl [S][n m][3], flag [S].  dynamic smem: G[nm][nm+1] | A[CH][nm] | rhs[nm][3] | p[CH][3]
__global__ void __launch_bounds__(NT) kron_fit_kernel(const double* __restrict__ U, const double* __restrict__ V,
                                                      const double* __restrict__ P, int M, int n, int m,
                                                      double* __restrict__ ctrl, int* __restrict__ flag) {
    extern __shared__ double sm[];
    const int nm = n * m, ld = nm + 1;
    double* G = sm;
    double* A = G + (size_t)nm * ld;
    double* rhs = A + (size_t)CH * nm;
    double* pc = rhs + (size_t)nm * 3;
    __shared__ int s_bad;
    __shared__ double s_dmax;
    const int s = blockIdx.x, t = threadIdx.x;
    const double* Us = U + (size_t)s * M * n;
    const double* Vs = V + (size_t)s * M * m;
    const double* Ps = P + (size_t)s * M * 3;
    // thread t owns the entries idx = t + NT e of the nm x nm matrix (lower triangle only is accumulated)
    double acc[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) acc[e] = 0.0;
    double racc[2] = {0.0, 0.0};                      // rhs entries t, t + NT of the nm x 3 right-hand side
    if (t == 0) s_bad = 0;
    for (int i0 = 0; i0 < M; i0 += CH) {
        const int rows = min(CH, M - i0);
        __syncthreads();
        for (int e = t; e < CH * nm; e += NT) {
            const int i = e / nm, a = e - i * nm;
            A[e] = (i < rows) ? Us[(size_t)(i0 + i) * n + a / m] * Vs[(size_t)(i0 + i) * m + a % m] : 0.0;
        }
        for (int e = t; e < CH * 3; e += NT) pc[e] = (e / 3 < rows) ? Ps[(size_t)i0 * 3 + e] : 0.0;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int idx = t + NT * e;
            if (idx < nm * nm) {
                const int a = idx / nm, b = idx - a * nm;
                if (b <= a) {
                    double v = acc[e];
                    for (int i = 0; i < CH; ++i) v = fma(A[i * nm + a], A[i * nm + b], v);
                    acc[e] = v;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = t + NT * e;
            if (idx < nm * 3) {
                const int a = idx / 3, c = idx - a * 3;
                double v = racc[e];
                for (int i = 0; i < CH; ++i) v = fma(A[i * nm + a], pc[i * 3 + c], v);
                racc[e] = v;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        const int idx = t + NT * e;
        if (idx < nm * nm) {
            const int a = idx / nm, b = idx - a * nm;
            if (b <= a) G[a * ld + b] = acc[e];
        }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e)
        if (t + NT * e < nm * 3) rhs[t + NT * e] = racc[e];
    __syncthreads();
    if (t == 0) {
        double d = 0.0;
        for (int a = 0; a < nm; ++a) d = fmax(d, G[a * ld + a]);
        s_dmax = d;
    }
    __syncthreads();
    const double tol = 1e-12 * s_dmax;
    // in-place Cholesky, lower triangle
    for (int k = 0; k < nm; ++k) {
        if (t == 0) {
            const double d = G[k * ld + k];
            if (!(d > tol)) s_bad = 1;
            G[k * ld + k] = sqrt(d > tol ? d : 1.0);
        }
        __syncthreads();
        const double inv = 1.0 / G[k * ld + k];
        for (int i = k + 1 + t; i < nm; i += NT) G[i * ld + k] *= inv;
        __syncthreads();
        const int r = nm - k - 1;
        for (int e = t; e < r * r; e += NT) {
            const int i = k + 1 + e / r, j = k + 1 + e % r;
            if (j <= i) G[i * ld + j] = fma(-G[i * ld + k], G[j * ld + k], G[i * ld + j]);
        }
        __syncthreads();
    }
    if (s_bad) {
        if (t == 0) flag[s] = 1;
        return;
    }
    if (t < 3) {                                       // L y = rhs, L^T x = y, one coordinate per thread
        for (int k = 0; k < nm; ++k) {
            double v = rhs[k * 3 + t];
            for (int j = 0; j < k; ++j) v = fma(-G[k * ld + j], rhs[j * 3 + t], v);
            rhs[k * 3 + t] = v / G[k * ld + k];
        }
        for (int k = nm - 1; k >= 0; --k) {
            double v = rhs[k * 3 + t];
            for (int j = k + 1; j < nm; ++j) v = fma(-G[j * ld + k], rhs[j * 3 + t], v);
            rhs[k * 3 + t] = v / G[k * ld + k];
        }
    }
    __syncthreads();
    for (int e = t; e < nm * 3; e += NT) ctrl[(size_t)s * nm * 3 + e] = rhs[e];
    if (t == 0) flag[s] = 0;
}

// out [S][M][3] = sum_ab U[s][i][a] V[s][i][b] C[s][a][b][:]   (C shared by all surfaces when c_stride == 0)
__global__ void kron_eval_kernel(const double* __restrict__ U, const double* __restrict__ V, const double* __restrict__ C,
                                 long long c_stride, int M, int n, int m, double* __restrict__ out) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const double* u = U + ((size_t)s * M + i) * n;
    const double* v = V + ((size_t)s * M + i) * m;
    const double* c = C + (size_t)s * c_stride;
    double x = 0.0, y = 0.0, z = 0.0;
    for (int a = 0; a < n; ++a) {
        const double ua = u[a];
        if (ua == 0.0) continue;
        for (int b = 0; b < m; ++b) {
            const double w = ua * v[b];
            const double* cc = c + ((size_t)a * m + b) * 3;
            x = fma(w, cc[0], x); y = fma(w, cc[1], y); z = fma(w, cc[2], z);
        }
    }
    double* o = out + ((size_t)s * M + i) * 3;
    o[0] = x; o[1] = y; o[2] = z;
}

}  // namespace kron
}  // namespace pn

using namespace pn;

extern "C" int pn_kron_fit(const double* U, const double* V, const double* P, int S, int M, int n, int m, double* ctrl, int* flag,
                           void* stream) {
    PN_REQUIRE(U && V && P && ctrl && flag, "pn_kron_fit: null pointer");
    PN_REQUIRE(S > 0 && M > 0 && n > 0 && m > 0, "pn_kron_fit: bad sizes (S=%d M=%d n=%d m=%d)", S, M, n, m);
    const int nm = n * m;
    PN_REQUIRE(nm <= kron::MAXNM && nm * 3 <= 2 * kron::NT, "pn_kron_fit: at most %d control points per surface (got %d x %d)",
               kron::MAXNM, n, m);
    const size_t smem = ((size_t)nm * (nm + 1) + (size_t)kron::CH * nm + (size_t)nm * 3 + kron::CH * 3) * sizeof(double);
    PN_CUDA(cudaFuncSetAttribute(kron::kron_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kron::kron_fit_kernel<<<S, kron::NT, smem, (cudaStream_t)stream>>>(U, V, P, M, n, m, ctrl, flag);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("kron_fit_kernel");
    return PN_OK;
}

extern "C" int pn_kron_eval(const double* U, const double* V, const double* C, long long c_stride, int S, int M, int n, int m,
                            double* out, void* stream) {
    PN_REQUIRE(U && V && C && out, "pn_kron_eval: null pointer");
    PN_REQUIRE(S > 0 && M > 0 && n > 0 && m > 0, "pn_kron_eval: bad sizes (S=%d M=%d n=%d m=%d)", S, M, n, m);
    kron::kron_eval_kernel<<<dim3(cdiv(M, 128), S), 128, 0, (cudaStream_t)stream>>>(U, V, C, c_stride, M, n, m, out);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("kron_eval_kernel");
    return PN_OK;
}
