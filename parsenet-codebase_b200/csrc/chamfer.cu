// Chamfer distance: tiled brute-force nearest-neighbour search in shared memory, argmin kept for the backward.
//
// Replaces chamfer_distance / chamfer_distance_one_side / chamfer_distance_single_shape
//   src/utils.py:273-358  — the reference broadcasts a (B,M,Np,3) difference tensor (1.4 GB at config 3).
// Here: for every query point of set A the squared distance to (and index of) its nearest point of set B; both
// directions are two launches.  Means / sqrt / one-sided variants are O(Np) glue on the returned vectors.
#include "common.cuh"

namespace pn {
namespace chamfer {

constexpr int NT = 256;
constexpr int TILE = 1024;   // candidate points staged per iteration (12 KB)

// A [B][Na][3], Bs [B][Nb][3] -> mind [B][Na], arg [B][Na];  grid (ceil(Na/NT), B)
__global__ void __launch_bounds__(NT) nn_fwd_kernel(const float* __restrict__ A, int Na, const float* __restrict__ Bs,
                                                    int Nb, float* __restrict__ mind, int* __restrict__ arg) {
    __shared__ float sx[TILE], sy[TILE], sz[TILE];
    const int b = blockIdx.y;
    const int i = blockIdx.x * NT + threadIdx.x;
    const float* Ab = A + (long long)b * Na * 3;
    const float* Bb = Bs + (long long)b * Nb * 3;
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (i < Na) { ax = Ab[3 * i]; ay = Ab[3 * i + 1]; az = Ab[3 * i + 2]; }
    float best = INFINITY; int bj = 0;
    for (int j0 = 0; j0 < Nb; j0 += TILE) {
        const int n = min(TILE, Nb - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += NT) {
            sx[t] = Bb[3 * (j0 + t)]; sy[t] = Bb[3 * (j0 + t) + 1]; sz[t] = Bb[3 * (j0 + t) + 2];
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < n; ++t) {
            float dx = ax - sx[t], dy = ay - sy[t], dz = az - sz[t];
            float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            bool better = d < best;
            best = better ? d : best;
            bj = better ? (j0 + t) : bj;
        }
    }
    if (i < Na) { mind[(long long)b * Na + i] = best; arg[(long long)b * Na + i] = bj; }
}

// d mind_i / d a_i = 2 (a_i - b_arg);  dA[i] += g_i * that ; dB[arg] -= g_i * that (atomic)
__global__ void nn_bwd_kernel(const float* __restrict__ A, int Na, const float* __restrict__ Bs, int Nb,
                              const int* __restrict__ arg, const float* __restrict__ g, long long total,
                              float* __restrict__ dA, float* __restrict__ dB) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    long long b = e / Na;
    const float* a = A + e * 3;
    long long jb = b * Nb + arg[e];
    const float* q = Bs + jb * 3;
    float gi = 2.f * g[e];
    float vx = gi * (a[0] - q[0]), vy = gi * (a[1] - q[1]), vz = gi * (a[2] - q[2]);
    if (dA) { atomicAdd(&dA[e * 3], vx); atomicAdd(&dA[e * 3 + 1], vy); atomicAdd(&dA[e * 3 + 2], vz); }
    if (dB) { atomicAdd(&dB[jb * 3], -vx); atomicAdd(&dB[jb * 3 + 1], -vy); atomicAdd(&dB[jb * 3 + 2], -vz); }
}

}  // namespace chamfer
}  // namespace pn

using namespace pn;

extern "C" int pn_chamfer_nn_fwd(const float* A, int Na, const float* Bs, int Nb, int B, float* mind, int* arg,
                                 void* stream) {
    PN_REQUIRE(A && Bs && mind && arg && Na > 0 && Nb > 0 && B > 0, "pn_chamfer_nn_fwd: bad args");
    dim3 grid(cdiv(Na, chamfer::NT), B);
    chamfer::nn_fwd_kernel<<<grid, chamfer::NT, 0, (cudaStream_t)stream>>>(A, Na, Bs, Nb, mind, arg);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("chamfer nn_fwd_kernel");
    return PN_OK;
}

extern "C" int pn_chamfer_nn_bwd(const float* A, int Na, const float* Bs, int Nb, int B, const int* arg,
                                 const float* g, float* dA_accum, float* dB_accum, void* stream) {
    PN_REQUIRE(A && Bs && arg && g, "pn_chamfer_nn_bwd: null pointer");
    long long total = (long long)B * Na;
    chamfer::nn_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(A, Na, Bs, Nb, arg, g, total, dA_accum,
                                                                              dB_accum);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("chamfer nn_bwd_kernel");
    return PN_OK;
}
