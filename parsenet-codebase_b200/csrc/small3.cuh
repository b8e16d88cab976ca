// 3x3 symmetric eigen-decomposition and regularised normal-equation solve in fp64, one problem per thread.
// Shared by the device kernels in small3.cu and (as plain host functions) by tests/c/small3_host.cpp.
//
// Replaces, on (S,3,3) batches, the LAPACK/MAGMA calls behind the reference's per-segment fits:
//   torch.svd of the weighted (m,3) matrix      src/fitting_utils.py:420-455 (CustomSVD)  -> eigh of its Gram matrix
//   LeastSquares.lstsq + best_lambda            src/fitting_utils.py:36-85                -> lstsq3 below
// (torch.linalg.eigh / solve on CUDA block the host on a cusolver info read-back; these never synchronise).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PN_HD __host__ __device__ __forceinline__
#else
#define PN_HD static inline
#endif

namespace pn {
namespace small3 {

// cyclic Jacobi on a symmetric 3x3 (row-major a[9]); w ascending, columns of v (row-major v[9]) = eigenvectors
PN_HD void eigh3(const double* a_in, double* w, double* v) {
    double a[3][3], q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (a_in[3 * i + j] + a_in[3 * j + i]);
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int r = p + 1; r < 3; ++r) {
                const double apq = a[p][r];
                if (apq == 0.0) continue;
                const double theta = (a[r][r] - a[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {           // A <- A J
                    const double akp = a[k][p], akr = a[k][r];
                    a[k][p] = c * akp - s * akr;
                    a[k][r] = s * akp + c * akr;
                }
                for (int k = 0; k < 3; ++k) {           // A <- J^T A
                    const double apk = a[p][k], ark = a[r][k];
                    a[p][k] = c * apk - s * ark;
                    a[r][k] = s * apk + c * ark;
                }
                for (int k = 0; k < 3; ++k) {           // Q <- Q J
                    const double qkp = q[k][p], qkr = q[k][r];
                    q[k][p] = c * qkp - s * qkr;
                    q[k][r] = s * qkp + c * qkr;
                }
            }
    }
    int o[3] = {0, 1, 2};
    double d[3] = {a[0][0], a[1][1], a[2][2]};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (d[o[j]] > d[o[j + 1]]) { int t = o[j]; o[j] = o[j + 1]; o[j + 1] = t; }
    for (int c = 0; c < 3; ++c) {
        w[c] = d[o[c]];
        for (int k = 0; k < 3; ++k) v[3 * k + c] = q[k][o[c]];
    }
}

// inverse of a symmetric 3x3 via the adjugate (the matrices here are made full rank by the lambda rule first)
PN_HD void inv3(const double* m, double* inv) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    inv[0] = c00 * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = c01 * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = c02 * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// LeastSquares.lstsq in normal-equation form (reference src/fitting_utils.py:36-65):
//   full rank (torch.matrix_rank of the (rows,3) matrix A, tolerance s_max * max(rows,3) * eps32): x = (AtA)^-1 AtY
//   otherwise lambda = 1e-6 * 10^j, first j < 7 making AtA + lambda I full rank (3x3 tolerance s_max * 3 * eps32),
//   x = (AtA + lambda I)^-1 AtY.   Outputs x[3], the inverse used (minv[9], for the backward) and lambda.
PN_HD void lstsq3(const double* AtA, const double* AtY, int rows, double eps32, double* x, double* minv,
                  double* lam_out) {
    double w[3], v[9];
    eigh3(AtA, w, v);
    const double e0 = w[2], e2 = w[0];                                 // largest / smallest eigenvalue of AtA
    const double s0 = sqrt(e0 > 0.0 ? e0 : 0.0), s2 = sqrt(e2 > 0.0 ? e2 : 0.0);
    const double tol = s0 * (double)(rows > 3 ? rows : 3) * eps32;
    double lam = 0.0;
    if (s2 <= tol) {
        double cur = 1e-6;
        bool done = false;
        for (int j = 0; j < 7; ++j) {
            if ((e2 + cur) > (e0 + cur) * 3.0 * eps32) { lam = cur; done = true; break; }
            cur *= 10.0;
        }
        if (!done) lam = cur;
    }
    double m[9];
    for (int i = 0; i < 9; ++i) m[i] = 0.5 * (AtA[i] + AtA[3 * (i % 3) + i / 3]);
    m[0] += lam; m[4] += lam; m[8] += lam;
    inv3(m, minv);
    for (int i = 0; i < 3; ++i) x[i] = minv[3 * i] * AtY[0] + minv[3 * i + 1] * AtY[1] + minv[3 * i + 2] * AtY[2];
    *lam_out = lam;
}

}  // namespace small3
}  // namespace pn
