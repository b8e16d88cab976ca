// Control-grid regression losses of the SplineNet training step (train_open_splines.py:158-176):
//   control_points_permute_reg_loss          src/loss.py:76-97   min over the 8 dihedral re-orderings of the target grid
//   control_points_permute_closed_reg_loss   src/loss.py:100-124 min over (g cyclic shifts along u) x (4 flips)
//   laplacian_loss                           src/loss.py:213-239 zero-padded 4-neighbour Laplacian of both grids, L2 / L1
// The reference materialises (B, 8 | 4g, g, g, 3) candidate stacks (torch.flip / transpose / roll / cat) and runs a 3x3x3x3
// cuDNN convolution; here a candidate is an index map evaluated on the fly and the Laplacian is a 5-point stencil.
#include "common.cuh"

namespace pn {
namespace gridloss {

constexpr int NT = 128;

// source cell of candidate p at (i, j).  mode 0 (open): p < 4 = flips {id, u, v, uv}, p >= 4 = transposes of the same flips.
// mode 1 (closed): p = 4 s + f = roll by s along u (torch.roll: out[i] = x[(i - s) mod g]) followed by flip f.
__device__ __forceinline__ void cand_src(int mode, int p, int g, int i, int j, int* si, int* sj) {
    int f, s = 0;
    if (mode == 0) {
        f = p & 3;
        if (p >= 4) { const int t = i; i = j; j = t; }
    } else {
        f = p & 3;
        s = p >> 2;
    }
    if (f & 1) i = g - 1 - i;
    if (f & 2) j = g - 1 - j;
    if (mode == 1) { i -= s; if (i < 0) i += g; }
    *si = i; *sj = j;
}

// diff[b][p] = sum (out - cand_p(gt))^2 ; grid (P, B)
__global__ void __launch_bounds__(NT) perm_diff_kernel(const float* __restrict__ out, const float* __restrict__ gt, int g,
                                                       int mode, int P, float* __restrict__ diff) {
    __shared__ float red[32];
    const int p = blockIdx.x, b = blockIdx.y;
    const float* o = out + (long long)b * g * g * 3;
    const float* t = gt + (long long)b * g * g * 3;
    float acc = 0.f;
    for (int e = threadIdx.x; e < g * g; e += NT) {
        const int i = e / g, j = e - i * g;
        int si, sj;
        cand_src(mode, p, g, i, j, &si, &sj);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = o[e * 3 + c] - t[(si * g + sj) * 3 + c];
            acc = fmaf(d, d, acc);
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) diff[b * P + p] = acc;
}

// per shape: lowest-index argmin over the candidates, loss_b = min, best[b] = that candidate of gt ; grid (B)
__global__ void __launch_bounds__(NT) perm_pick_kernel(const float* __restrict__ diff, const float* __restrict__ gt, int g,
                                                       int mode, int P, float* __restrict__ loss_b, int* __restrict__ pick,
                                                       float* __restrict__ best) {
    __shared__ int s_p;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        float m = diff[b * P];
        int arg = 0;
        for (int p = 1; p < P; ++p) {
            const float v = diff[b * P + p];
            if (v < m) { m = v; arg = p; }
        }
        loss_b[b] = m; pick[b] = arg; s_p = arg;
    }
    __syncthreads();
    const int p = s_p;
    const float* t = gt + (long long)b * g * g * 3;
    float* o = best + (long long)b * g * g * 3;
    for (int e = threadIdx.x; e < g * g; e += NT) {
        const int i = e / g, j = e - i * g;
        int si, sj;
        cand_src(mode, p, g, i, j, &si, &sj);
#pragma unroll
        for (int c = 0; c < 3; ++c) o[e * 3 + c] = t[(si * g + sj) * 3 + c];
    }
}

// dout = (out - best) * (2 * gscale[0] * inv)   (gscale: device scalar = upstream gradient of the mean loss)
__global__ void perm_bwd_kernel(const float* __restrict__ out, const float* __restrict__ best, long long n,
                                const float* __restrict__ gscale, float inv, float* __restrict__ dout) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) dout[e] = (out[e] - best[e]) * (2.0f * gscale[0] * inv);
}

__device__ __forceinline__ float lap_at(const float* d, int g, int i, int j, int c) {
    float v = d[(i * g + j) * 3 + c];
    float nb = 0.f;
    if (i > 0) nb += d[((i - 1) * g + j) * 3 + c];
    if (i < g - 1) nb += d[((i + 1) * g + j) * 3 + c];
    if (j > 0) nb += d[(i * g + j - 1) * 3 + c];
    if (j < g - 1) nb += d[(i * g + j + 1) * 3 + c];
    return v - 0.25f * nb;
}

// one CTA per shape: l = lap(out) - lap(gt) per cell / channel (kept for the backward), part[b] = sum l^2 (or |l|)
__global__ void __launch_bounds__(NT) lap_fwd_kernel(const float* __restrict__ out, const float* __restrict__ gt, int g, int l1,
                                                     float* __restrict__ l, float* __restrict__ part) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    const float* o = out + (long long)b * g * g * 3;
    const float* t = gt + (long long)b * g * g * 3;
    float* lb = l + (long long)b * g * g * 3;
    float acc = 0.f;
    for (int e = threadIdx.x; e < g * g * 3; e += NT) {
        const int c = e % 3, ij = e / 3, i = ij / g, j = ij - i * g;
        const float v = lap_at(o, g, i, j, c) - lap_at(t, g, i, j, c);
        lb[e] = v;
        acc += l1 ? fabsf(v) : v * v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[b] = acc;
}

// the Laplacian stencil with zero padding is self-adjoint: dout = lap(w), w = 2 l (L2) or sign(l) (L1), times gscale * inv;
// dgt = -dout
__global__ void __launch_bounds__(NT) lap_bwd_kernel(const float* __restrict__ l, int g, int l1, const float* __restrict__ gscale,
                                                     float inv, float* __restrict__ dout, float* __restrict__ dgt) {
    extern __shared__ float w[];
    const int b = blockIdx.x;
    const float* lb = l + (long long)b * g * g * 3;
    for (int e = threadIdx.x; e < g * g * 3; e += NT) {
        const float v = lb[e];
        w[e] = l1 ? (v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f)) : 2.0f * v;
    }
    __syncthreads();
    const float sc = gscale[0] * inv;
    for (int e = threadIdx.x; e < g * g * 3; e += NT) {
        const int c = e % 3, ij = e / 3, i = ij / g, j = ij - i * g;
        const float v = lap_at(w, g, i, j, c) * sc;
        dout[(long long)b * g * g * 3 + e] = v;
        if (dgt) dgt[(long long)b * g * g * 3 + e] = -v;
    }
}

}  // namespace gridloss
}  // namespace pn

using namespace pn;

// out, gt [B][g][g][3]; mode 0: 8 dihedral candidates, mode 1: 4 g (roll along u x flips) -> diff [B][P] workspace,
// loss_b [B] (min squared error per shape), pick [B], best [B][g][g][3] (the best-matching candidate of gt)
extern "C" int pn_grid_perm_fwd(const float* out, const float* gt, int B, int g, int mode, float* diff_ws, float* loss_b,
                                int* pick, float* best, void* stream) {
    PN_REQUIRE(out && gt && diff_ws && loss_b && pick && best, "pn_grid_perm_fwd: null pointer");
    PN_REQUIRE(B > 0 && g > 0 && (mode == 0 || mode == 1), "pn_grid_perm_fwd: bad arguments (B=%d g=%d mode=%d)", B, g, mode);
    const int P = mode == 0 ? 8 : 4 * g;
    cudaStream_t st = (cudaStream_t)stream;
    gridloss::perm_diff_kernel<<<dim3(P, B), gridloss::NT, 0, st>>>(out, gt, g, mode, P, diff_ws);
    PN_COUNT_LAUNCH();
    gridloss::perm_pick_kernel<<<B, gridloss::NT, 0, st>>>(diff_ws, gt, g, mode, P, loss_b, pick, best);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("grid perm kernels");
    return PN_OK;
}

// dout = 2 (out - best) * gscale * inv   (inv = 1 / (B * g * g * 3): the mean over the batch and the reference's normalisation)
extern "C" int pn_grid_perm_bwd(const float* out, const float* best, long long n, const float* gscale, float inv, float* dout,
                                void* stream) {
    PN_REQUIRE(out && best && gscale && dout && n > 0, "pn_grid_perm_bwd: bad arguments");
    gridloss::perm_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(out, best, n, gscale, inv, dout);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("perm_bwd_kernel");
    return PN_OK;
}

extern "C" int pn_grid_laplacian_fwd(const float* out, const float* gt, int B, int g, int l1, float* l_ws, float* part,
                                     void* stream) {
    PN_REQUIRE(out && gt && l_ws && part && B > 0 && g > 0, "pn_grid_laplacian_fwd: bad arguments");
    gridloss::lap_fwd_kernel<<<B, gridloss::NT, 0, (cudaStream_t)stream>>>(out, gt, g, l1, l_ws, part);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("lap_fwd_kernel");
    return PN_OK;
}

extern "C" int pn_grid_laplacian_bwd(const float* l_ws, int B, int g, int l1, const float* gscale, float inv, float* dout,
                                     float* dgt, void* stream) {
    PN_REQUIRE(l_ws && gscale && dout && B > 0 && g > 0, "pn_grid_laplacian_bwd: bad arguments");
    const size_t sm = (size_t)g * g * 3 * sizeof(float);
    PN_REQUIRE(sm <= 48 * 1024, "pn_grid_laplacian_bwd: grid too large for the shared stencil buffer (g=%d)", g);
    gridloss::lap_bwd_kernel<<<B, gridloss::NT, sm, (cudaStream_t)stream>>>(l_ws, g, l1, gscale, inv, dout, dgt);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("lap_bwd_kernel");
    return PN_OK;
}
