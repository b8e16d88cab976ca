// Primitive parameters from the 55 weighted moments of a segment, together with their Jacobian w.r.t. the moments, in
// fp64 -- one (segment, moment direction) pair per thread, forward-mode: every quantity is a (value, tangent) pair and
// thread j carries the tangent along moment j, so ONE launch yields the parameters of every segment of every shape of a
// step and the full (8 x 55) Jacobians its backward needs (grad_moments = grad_params . J, one small batched product).
// Shared by the device kernel in fitsolve.cu and (as plain host functions) by tests/c/fitsolve_host.cpp.
//
// Replaces, per matched segment, the python 3x3 algebra behind the reference's fits (previously ~100 sub-10-us torch
// kernels per primitive kind and shape, plus as many again in autograd):
//   Fit.fit_plane_torch     src/primitive_forward.py:708-729      Fit.fit_sphere_torch   :746-769
//   Fit.fit_cylinder_torch  :784-806                               Fit.fit_cone_torch     :808-831 (apex, axis, cond rule)
//   LeastSquares.lstsq + best_lambda  src/fitting_utils.py:36-85   CustomSVD + its backward  :385-455
// The tangent of the singular vectors is the TRANSPOSE of the reference's custom backward rule (only grad_V flows,
// K_ij = 1 / ((s_i - s_j)(s_i + s_j)) with |s_i - s_j| floored at 1e-6): backward gG = V sym(K^T o (V^T gV)) V^T is linear
// in gV, and <gG, dG> = <gV, V (K^T o sym(V^T dG V))>, i.e. dV = V (K^T o sym(V^T dG V)).  The least-squares tangent is
// the one of a linear solve with the regularisation lambda held fixed (what autograd does for the reference's
// torch.inverse path): dx = M^-1 (dAtY - dAtA x).
// Moment layout: csrc/fit.cu eval_phi / pnb200/fitting.py.  Parameter rows (8 doubles) in the layout of the residual
// kernel (csrc/primitives.cu): plane [a(3), d], sphere [c(3), r], cylinder [a(3), c(3), r], cone [apex(3), a(3), theta = 0].
#pragma once
#include "small3.cuh"

namespace pn {
namespace fitsolve {

constexpr int NM = 55, NPAR = 8;
constexpr double EPS32 = 1.1920928955078125e-07;      // np.finfo(np.float32).eps
enum : int { KIND_NONE = -1, KIND_PLANE = 0, KIND_SPHERE = 1, KIND_CYLINDER = 2, KIND_CONE = 3 };

struct Du { double v, d; };
PN_HD Du mk(double v, double d = 0.0) { Du r; r.v = v; r.d = d; return r; }
PN_HD Du operator+(Du a, Du b) { return mk(a.v + b.v, a.d + b.d); }
PN_HD Du operator-(Du a, Du b) { return mk(a.v - b.v, a.d - b.d); }
PN_HD Du operator-(Du a) { return mk(-a.v, -a.d); }
PN_HD Du operator*(Du a, Du b) { return mk(a.v * b.v, a.d * b.v + a.v * b.d); }
PN_HD Du operator/(Du a, Du b) { const double q = a.v / b.v; return mk(q, (a.d - q * b.d) / b.v); }
PN_HD Du operator*(double s, Du a) { return mk(s * a.v, s * a.d); }
PN_HD Du operator+(Du a, double s) { return mk(a.v + s, a.d); }

struct V3 { Du e[3]; };
struct M3 { Du e[3][3]; };

PN_HD Du dot(const V3& a, const V3& b) { return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2]; }
PN_HD V3 mom3(const Du* m, int o) { V3 r; r.e[0] = m[o]; r.e[1] = m[o + 1]; r.e[2] = m[o + 2]; return r; }
// (S,6) [xx,xy,xz,yy,yz,zz] -> symmetric 3x3
PN_HD M3 sym6(const Du* m, int o) {
    const int ix[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.e[i][j] = m[o + ix[i][j]];
    return r;
}
// unique entries of the symmetric 3-tensor sum w^3 p p p (order xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz)
PN_HD Du t10(const Du* m, int o, int i, int j, int k) {
    int a = i, b = j, c = k, t;
    if (a > b) { t = a; a = b; b = t; }
    if (b > c) { t = b; b = c; c = t; }
    if (a > b) { t = a; a = b; b = t; }
    const int code = a * 9 + b * 3 + c;
    int idx = 0;
    switch (code) {
        case 0: idx = 0; break;  case 1: idx = 1; break;  case 2: idx = 2; break;  case 4: idx = 3; break;
        case 5: idx = 4; break;  case 8: idx = 5; break;  case 13: idx = 6; break; case 14: idx = 7; break;
        case 17: idx = 8; break; default: idx = 9; break;
    }
    return m[o + idx];
}

// right singular vectors of a matrix A given its Gram matrix G = A^T A: columns of V ordered by DEcreasing singular value
// (pnb200.fitting.GramSVDFn), tangent = transpose of the reference's custom backward.
PN_HD M3 gram_svd(const M3& G) {
    double a[9], w[3], q[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[3 * i + j] = G.e[i][j].v;
    small3::eigh3(a, w, q);                                  // ascending, eigenvectors in columns
    double V[3][3], sv[3];
    for (int c = 0; c < 3; ++c) {
        const double ev = w[2 - c];
        sv[c] = sqrt(ev > 0.0 ? ev : 0.0);
        for (int k = 0; k < 3; ++k) V[k][c] = q[3 * k + (2 - c)];
    }
    // a null direction comes out of the eigen-solver as 0 (or tiny negative): floor like an fp32 SVD of the (m,3)
    // matrix would return it (pnb200.fitting.floor_singular_values)
    double fl = EPS32 * sv[0];
    if (fl < 1e-30) fl = 1e-30;
    for (int c = 0; c < 3; ++c) if (sv[c] < fl) sv[c] = fl;
    // M = V^T dG V, symmetrised
    double dG[3][3], T[3][3], M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) dG[i][j] = G.e[i][j].d;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i][j] = dG[i][0] * V[0][j] + dG[i][1] * V[1][j] + dG[i][2] * V[2][j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = V[0][i] * T[0][j] + V[1][i] * T[1][j] + V[2][i] * T[2][j];
    double W[3][3];                                          // K^T o sym(M)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            if (i == j) { W[i][j] = 0.0; continue; }
            // K[j][i] = 1 / (guard(s_j - s_i) * max(s_j + s_i, 1e-8))
            const double diff = sv[j] - sv[i];
            const double ad = fabs(diff) > 1e-6 ? fabs(diff) : 1e-6;
            const double kneg = (diff >= 0.0 ? 1.0 : -1.0) * ad;
            const double plus = (sv[j] + sv[i]) > 1e-8 ? (sv[j] + sv[i]) : 1e-8;
            W[i][j] = (1.0 / kneg) * (1.0 / plus) * 0.5 * (M[i][j] + M[j][i]);
        }
    M3 R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            R.e[i][j] = mk(V[i][j], V[i][0] * W[0][j] + V[i][1] * W[1][j] + V[i][2] * W[2][j]);
    return R;
}

// x = (AtA + lambda I)^-1 AtY with the reference's rank rule for lambda (small3::lstsq3)
PN_HD V3 solve_normal(const M3& AtA, const V3& AtY, int rows) {
    double a[9], y[3], x[3], minv[9], lam;
    for (int i = 0; i < 3; ++i) {
        y[i] = AtY.e[i].v;
        for (int j = 0; j < 3; ++j) a[3 * i + j] = AtA.e[i][j].v;
    }
    small3::lstsq3(a, y, rows, EPS32, x, minv, &lam);
    double t[3];
    for (int i = 0; i < 3; ++i)
        t[i] = AtY.e[i].d - (AtA.e[i][0].d * x[0] + AtA.e[i][1].d * x[1] + AtA.e[i][2].d * x[2]);
    V3 r;
    for (int i = 0; i < 3; ++i) r.e[i] = mk(x[i], minv[3 * i] * t[0] + minv[3 * i + 1] * t[1] + minv[3 * i + 2] * t[2]);
    return r;
}

// weighted plane through points with first / second moments (w1, s1p) / (w2, s2p, s2pp): unit normal = smallest right
// singular vector of w (p - c), d = a . c
PN_HD void plane_from(Du w1, const V3& s1p, Du w2, const V3& s2p, const M3& s2pp, V3* a, Du* d) {
    const Du sw = w1 + EPS32;
    V3 c;
    for (int i = 0; i < 3; ++i) c.e[i] = s1p.e[i] / sw;
    M3 G;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            G.e[i][j] = s2pp.e[i][j] - c.e[i] * s2p.e[j] - s2p.e[i] * c.e[j] + w2 * c.e[i] * c.e[j];
    const M3 V = gram_svd(G);
    for (int i = 0; i < 3; ++i) a->e[i] = V.e[i][2];
    *d = dot(*a, s1p) / sw;
}

// linearised weighted sphere fit (primitive_forward.py:746-769) from moments
PN_HD void sphere_from(Du w1, const V3& s1p, Du q1, Du w2, const V3& s2p, const M3& s2pp, Du q3, const V3& r3, int rows,
                       V3* center, Du* radius) {
    const Du sw = w1 + EPS32;
    V3 c;
    for (int i = 0; i < 3; ++i) c.e[i] = s1p.e[i] / sw;
    const Du nu = q1 / sw;
    M3 AtA;
    V3 AtY;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            AtA.e[i][j] = 4.0 * (w2 * c.e[i] * c.e[j] - c.e[i] * s2p.e[j] - s2p.e[i] * c.e[j] + s2pp.e[i][j]);
        AtY.e[i] = 2.0 * (c.e[i] * q3 - r3.e[i] - (nu * w2) * c.e[i] + nu * s2p.e[i]);
    }
    const V3 x = solve_normal(AtA, AtY, rows);
    for (int i = 0; i < 3; ++i) center->e[i] = -x.e[i];
    const Du r2 = (q1 - 2.0 * dot(*center, s1p) + dot(*center, *center) * w1) / sw;
    // guard_sqrt(clamp(r2, min=1e-3)): the gradient passes where r2 >= 1e-3
    if (r2.v >= 1e-3) { const double s = sqrt(r2.v); *radius = mk(s, r2.d / (2.0 * s)); }
    else *radius = mk(sqrt(1e-3), 0.0);
}

enum : int { M1_1 = 0, M1_P = 1, M1_PP = 4, M1_N = 10, M2_1 = 13, M2_P = 14, M2_PP = 17, M2_N = 23, M2_NN = 26,
             M2_NPN = 32, M3_PP = 35, M3_T = 41, M0_N = 51, M0_1 = 54 };

// one segment, one tangent direction: m[55] (value, tangent) -> par[8] (value, tangent); returns the degenerate flag of
// the cone rule (cond(w n) > 1e5 -> zero apex, x axis, zero angle, no gradient; primitive_forward.py:818-823)
PN_HD bool solve_segment(int kind, const Du* m, int rows, Du* par) {
    for (int i = 0; i < NPAR; ++i) par[i] = mk(0.0);
    if (kind == KIND_PLANE) {
        V3 a; Du d;
        plane_from(m[M1_1], mom3(m, M1_P), m[M2_1], mom3(m, M2_P), sym6(m, M2_PP), &a, &d);
        par[0] = a.e[0]; par[1] = a.e[1]; par[2] = a.e[2]; par[3] = d;
        return false;
    }
    if (kind == KIND_SPHERE) {
        const M3 pp1 = sym6(m, M1_PP), pp3 = sym6(m, M3_PP);
        const Du q1 = pp1.e[0][0] + pp1.e[1][1] + pp1.e[2][2], q3 = pp3.e[0][0] + pp3.e[1][1] + pp3.e[2][2];
        V3 r3, c; Du r;
        for (int k = 0; k < 3; ++k) r3.e[k] = t10(m, M3_T, 0, 0, k) + t10(m, M3_T, 1, 1, k) + t10(m, M3_T, 2, 2, k);
        sphere_from(m[M1_1], mom3(m, M1_P), q1, m[M2_1], mom3(m, M2_P), sym6(m, M2_PP), q3, r3, rows, &c, &r);
        par[0] = c.e[0]; par[1] = c.e[1]; par[2] = c.e[2]; par[3] = r;
        return false;
    }
    if (kind == KIND_CYLINDER) {
        // axis = smallest right singular vector of w n; circle fit of the points projected along it (:784-806)
        const M3 V = gram_svd(sym6(m, M2_NN));
        V3 a;
        for (int i = 0; i < 3; ++i) a.e[i] = V.e[i][2];
        const Du nrm2 = dot(a, a);
        const double nv = sqrt(nrm2.v);
        const Du nrm = mk(nv, nv > 0.0 ? nrm2.d / (2.0 * nv) : 0.0) + EPS32;
        for (int i = 0; i < 3; ++i) a.e[i] = a.e[i] / nrm;
        M3 P, M2;                                           // P = I - a a^T, M2 = P^T P
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) P.e[i][j] = mk(i == j ? 1.0 : 0.0) - a.e[i] * a.e[j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M2.e[i][j] = P.e[0][i] * P.e[0][j] + P.e[1][i] * P.e[1][j] + P.e[2][i] * P.e[2][j];
        const M3 pp1 = sym6(m, M1_PP), pp2 = sym6(m, M2_PP), pp3 = sym6(m, M3_PP);
        const V3 p1 = mom3(m, M1_P), p2 = mom3(m, M2_P);
        V3 s1p, s2p, r3;
        M3 T1, s2pp;
        Du q1 = mk(0.0), q3 = mk(0.0);
        for (int i = 0; i < 3; ++i) {
            s1p.e[i] = P.e[i][0] * p1.e[0] + P.e[i][1] * p1.e[1] + P.e[i][2] * p1.e[2];
            s2p.e[i] = P.e[i][0] * p2.e[0] + P.e[i][1] * p2.e[1] + P.e[i][2] * p2.e[2];
            for (int j = 0; j < 3; ++j) {
                T1.e[i][j] = P.e[i][0] * pp2.e[0][j] + P.e[i][1] * pp2.e[1][j] + P.e[i][2] * pp2.e[2][j];
                q1 = q1 + M2.e[i][j] * pp1.e[i][j];
                q3 = q3 + M2.e[i][j] * pp3.e[i][j];
            }
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)                      // (P pp2) P^T
                s2pp.e[i][j] = T1.e[i][0] * P.e[j][0] + T1.e[i][1] * P.e[j][1] + T1.e[i][2] * P.e[j][2];
        for (int k = 0; k < 3; ++k) {
            Du acc = mk(0.0);
            for (int l = 0; l < 3; ++l) {
                Du inner = mk(0.0);
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) inner = inner + M2.e[i][j] * t10(m, M3_T, i, j, l);
                acc = acc + P.e[k][l] * inner;
            }
            r3.e[k] = acc;
        }
        V3 c; Du r;
        sphere_from(m[M1_1], s1p, q1, m[M2_1], s2p, s2pp, q3, r3, rows, &c, &r);
        par[0] = a.e[0]; par[1] = a.e[1]; par[2] = a.e[2]; par[3] = c.e[0]; par[4] = c.e[1]; par[5] = c.e[2]; par[6] = r;
        return false;
    }
    if (kind == KIND_CONE) {
        const M3 nn = sym6(m, M2_NN);
        double a9[9], w[3], q[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) a9[3 * i + j] = nn.e[i][j].v;
        small3::eigh3(a9, w, q);
        const double s0 = sqrt(w[2] > 0.0 ? w[2] : 0.0), s2 = sqrt(w[0] > 0.0 ? w[0] : 0.0);
        const bool degenerate = (s0 / s2) > 1e5;            // (0/0 = NaN compares false, like the torch expression)
        if (degenerate) { par[3] = mk(1.0); return true; }
        const V3 apex = solve_normal(nn, mom3(m, M2_NPN), rows);
        V3 a; Du d;
        plane_from(m[M1_1], mom3(m, M1_N), m[M2_1], mom3(m, M2_N), nn, &a, &d);
        const V3 n0 = mom3(m, M0_N);
        if (dot(n0, a).v > 0.0) for (int i = 0; i < 3; ++i) a.e[i] = -a.e[i];
        par[0] = apex.e[0]; par[1] = apex.e[1]; par[2] = apex.e[2]; par[3] = a.e[0]; par[4] = a.e[1]; par[5] = a.e[2];
        return false;
    }
    return false;
}

// all tangents of one segment on the host / in one thread: par[8], jac[8][55]
PN_HD bool solve_segment_full(int kind, const double* mom, int rows, double* par, double* jac) {
    Du m[NM], p[NPAR];
    bool bad = false;
    for (int j = 0; j < NM; ++j) {
        for (int k = 0; k < NM; ++k) m[k] = mk(mom[k], k == j ? 1.0 : 0.0);
        bad = solve_segment(kind, m, rows, p);
        for (int i = 0; i < NPAR; ++i) jac[i * NM + j] = p[i].d;
        if (j == 0) for (int i = 0; i < NPAR; ++i) par[i] = p[i].v;
    }
    return bad;
}

}  // namespace fitsolve
}  // namespace pn
