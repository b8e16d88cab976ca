// Point-to-primitive squared distances for all matched segments of a shape in ONE launch, with the parameter
// Jacobian accumulated in the same pass (the backward of the per-segment mean is then a scalar times that Jacobian).
//
// Replaces ComputePrimitiveDistance.distance_from_{plane,sphere,cylinder,cone}  src/primitives.py:100-195
// (a python loop over segments issuing ~15 elementwise kernels each) as used by ResidualLoss.residual_loss :36-44.
//   plane    (a, d)          f = (p.a - d)^2
//   sphere   (c, r)          f = (|p - c| - r)^2
//   cylinder (a, c, r)       f = (sqrt(max(|v|^2 - (v.a)^2, 1e-5)) - r)^2 ,  v = p - c
//   cone     (apex, a, th)   f = (|v| sin(min(|acos(clamp(v.a/(|v|+1e-7), +-.999)) - th|, 3.142/2)))^2 , v = p - apex + 1e-8
// Parameter layout per segment: 8 floats  [type-specific, see below];  type ids: 0 plane, 1 sphere, 2 cylinder, 3 cone.
#include "common.cuh"

namespace pn {
namespace prim {

constexpr int NPAR = 8;
constexpr int NT = 256;
constexpr int MAXSEG = 64;

// seg[n] in [0,S) or -1; type[s]; par[s][8].  Outputs (zero-initialised): sumf[S], jac[S][8], cnt[S]
__global__ void __launch_bounds__(NT) residual_kernel(const float* P, const int* seg, int N,
                                                      const int* type, const float* par,
                                                      int S, float* sumf, float* jac,
                                                      float* cnt) {
    __shared__ float s_par[MAXSEG * NPAR];
    __shared__ int s_type[MAXSEG];
    __shared__ float s_sum[MAXSEG], s_cnt[MAXSEG], s_jac[MAXSEG * NPAR];
    // blockIdx.y = shape of a batch (every table has S slots per shape, N points per shape)
    P += (long long)blockIdx.y * N * 3; seg += (long long)blockIdx.y * N;
    type += blockIdx.y * S; par += blockIdx.y * S * NPAR;
    sumf += blockIdx.y * S; jac += blockIdx.y * S * NPAR; cnt += blockIdx.y * S;
    for (int e = threadIdx.x; e < S * NPAR; e += NT) { s_par[e] = par[e]; s_jac[e] = 0.f; }
    for (int e = threadIdx.x; e < S; e += NT) { s_type[e] = type[e]; s_sum[e] = 0.f; s_cnt[e] = 0.f; }
    __syncthreads();
    const int n = blockIdx.x * NT + threadIdx.x;
    if (n < N) {
        const int s = seg[n];
        if (s >= 0 && s < S && s_type[s] >= 0) {
            const float* q = s_par + s * NPAR;
            const float px = P[3 * n], py = P[3 * n + 1], pz = P[3 * n + 2];
            float f = 0.f, g[NPAR] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int t = s_type[s];
            if (t == 0) {                       // plane: a = q[0..2], d = q[3]
                float r = px * q[0] + py * q[1] + pz * q[2] - q[3];
                f = r * r;
                g[0] = 2.f * r * px; g[1] = 2.f * r * py; g[2] = 2.f * r * pz; g[3] = -2.f * r;
            } else if (t == 1) {                // sphere: c = q[0..2], r = q[3]
                float vx = px - q[0], vy = py - q[1], vz = pz - q[2];
                float rho = sqrtf(vx * vx + vy * vy + vz * vz);
                float e = rho - q[3];
                f = e * e;
                float k = rho > 0.f ? 2.f * e / rho : 0.f;
                g[0] = -k * vx; g[1] = -k * vy; g[2] = -k * vz; g[3] = -2.f * e;
            } else if (t == 2) {                // cylinder: a = q[0..2], c = q[3..5], r = q[6]
                float vx = px - q[3], vy = py - q[4], vz = pz - q[5];
                float tt = vx * q[0] + vy * q[1] + vz * q[2];
                float ds = (vx * vx + vy * vy + vz * vz) - tt * tt;
                bool live = ds >= 1e-5f;
                float qq = sqrtf(live ? ds : 1e-5f);
                float e = qq - q[6];
                f = e * e;
                float gds = live ? (2.f * e) / (2.f * qq) : 0.f;
                g[0] = gds * (-2.f * tt * vx); g[1] = gds * (-2.f * tt * vy); g[2] = gds * (-2.f * tt * vz);
                g[3] = gds * (-2.f * vx + 2.f * tt * q[0]);
                g[4] = gds * (-2.f * vy + 2.f * tt * q[1]);
                g[5] = gds * (-2.f * vz + 2.f * tt * q[2]);
                g[6] = -2.f * e;
            } else {                            // cone: apex = q[0..2], a = q[3..5], theta = q[6]
                float vx = px - q[0] + 1e-8f, vy = py - q[1] + 1e-8f, vz = pz - q[2] + 1e-8f;
                float mod = sqrtf(vx * vx + vy * vy + vz * vz);
                float va = vx * q[3] + vy * q[4] + vz * q[5];
                float den = mod + 1e-7f;
                float u = va / den;
                bool uin = (u >= -0.999f) && (u <= 0.999f);
                float ax = fminf(fmaxf(u, -0.999f), 0.999f);
                float alpha = acosf(ax);
                float delta = alpha - q[6];
                float ad = fabsf(delta);
                const float HALF = 3.142f / 2.0f;
                bool din = ad <= HALF;
                float da = din ? ad : HALF;
                float sn = sinf(da), cs = cosf(da);
                float ms = mod * sn;
                f = ms * ms;
                float g_da = 2.f * mod * mod * sn * cs;
                float sgn = delta > 0.f ? 1.f : (delta < 0.f ? -1.f : 0.f);
                float g_alpha = din ? g_da * sgn : 0.f;
                float g_u = uin ? -g_alpha / sqrtf(1.f - ax * ax) : 0.f;
                float gm = mod > 0.f ? 2.f * mod * sn * sn / mod : 0.f;           // d f / d mod * (1/mod) factor for v
                float k1 = g_u / den;
                float k2 = mod > 0.f ? g_u * va / (den * den * mod) : 0.f;
                float gvx = k1 * q[3] - k2 * vx + gm * vx;
                float gvy = k1 * q[4] - k2 * vy + gm * vy;
                float gvz = k1 * q[5] - k2 * vz + gm * vz;
                g[0] = -gvx; g[1] = -gvy; g[2] = -gvz;
                g[3] = k1 * vx; g[4] = k1 * vy; g[5] = k1 * vz;
                g[6] = -g_alpha;
            }
            atomicAdd(&s_sum[s], f);
            atomicAdd(&s_cnt[s], 1.f);
#pragma unroll
            for (int k = 0; k < NPAR; ++k) if (g[k] != 0.f) atomicAdd(&s_jac[s * NPAR + k], g[k]);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < S; e += NT) {
        if (s_cnt[e] != 0.f) { atomicAdd(&sumf[e], s_sum[e]); atomicAdd(&cnt[e], s_cnt[e]); }
    }
    for (int e = threadIdx.x; e < S * NPAR; e += NT) if (s_jac[e] != 0.f) atomicAdd(&jac[e], s_jac[e]);
}

}  // namespace prim
}  // namespace pn

using namespace pn;

extern "C" int pn_residual_fwd(const float* P, const int* seg, int N, const int* type, const float* par, int S,
                               float* sumf_zeroed, float* jac_zeroed, float* cnt_zeroed, void* stream) {
    PN_REQUIRE(P && seg && type && par && sumf_zeroed && jac_zeroed && cnt_zeroed, "pn_residual_fwd: null pointer");
    PN_REQUIRE(S > 0 && S <= prim::MAXSEG, "pn_residual_fwd: 1 <= segments <= %d (got %d)", prim::MAXSEG, S);
    prim::residual_kernel<<<cdiv(N, prim::NT), prim::NT, 0, (cudaStream_t)stream>>>(P, seg, N, type, par, S, sumf_zeroed,
                                                                                   jac_zeroed, cnt_zeroed);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("residual_kernel");
    return PN_OK;
}

// Batched over the shapes of a step: P [B][N][3], seg [B][N] (slot of the point's segment or -1), type / par / outputs
// [B][S](x8) indexed by slot (type -1 = unused slot; its points must carry seg = -1).
extern "C" int pn_residual_fwd_batched(const float* P, const int* seg, int B, int N, const int* type, const float* par, int S,
                                       float* sumf_zeroed, float* jac_zeroed, float* cnt_zeroed, void* stream) {
    PN_REQUIRE(P && seg && type && par && sumf_zeroed && jac_zeroed && cnt_zeroed, "pn_residual_fwd_batched: null pointer");
    PN_REQUIRE(S > 0 && S <= prim::MAXSEG, "pn_residual_fwd_batched: 1 <= slots <= %d (got %d)", prim::MAXSEG, S);
    PN_REQUIRE(B > 0 && N > 0, "pn_residual_fwd_batched: empty batch (B=%d N=%d)", B, N);
    prim::residual_kernel<<<dim3(cdiv(N, prim::NT), B), prim::NT, 0, (cudaStream_t)stream>>>(P, seg, N, type, par, S,
                                                                                            sumf_zeroed, jac_zeroed, cnt_zeroed);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("residual_kernel (batched)");
    return PN_OK;
}
