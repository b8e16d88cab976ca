// Segment matching of the fit / evaluation stage on the device:
//   relaxed_iou_fast on one-hot memberships   reference src/segment_utils.py:356-373   -> pn_iou_cost (K x K confusion matrix)
//   lapsolver.solve_dense(1 - iou)            reference src/fitting_utils.py:362-376, src/segment_utils.py:165-174 -> pn_hungarian
// The reference builds (N, 50) one-hot matrices, multiplies them on the GPU, copies the cost to the host and runs lapsolver.
#include "common.cuh"
#include "assign.cuh"

namespace pn {
namespace assign {

// pred, gt [B][N] labels in [0, K) -> cost [B][K][K] = 1 - inter / (|pred = p| + |gt = g| - inter + 1e-7), the fp32 expression of
// the reference (every count is an integer below 2^24, so its fp32 matmul is exact).  One CTA per shape.
__global__ void __launch_bounds__(256) iou_cost_kernel(const int* __restrict__ pred, const int* __restrict__ gt, int N, int K,
                                                       float* __restrict__ cost, int* __restrict__ bad) {
    extern __shared__ int conf[];                 // [K][K] | np[K] | ng[K]
    int* np_ = conf + K * K;
    int* ng = np_ + K;
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < K * K + 2 * K; e += blockDim.x) conf[e] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int p = pred[(long long)b * N + i], g = gt[(long long)b * N + i];
        if (p < 0 || p >= K || g < 0 || g >= K) { *bad = 1; continue; }
        atomicAdd(&conf[p * K + g], 1);
        atomicAdd(&np_[p], 1);
        atomicAdd(&ng[g], 1);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
        const int p = e / K, g = e - p * K;
        const float d = (float)conf[e];
        const float den = __fadd_rn(__fsub_rn(__fadd_rn((float)np_[p], (float)ng[g]), d), 1e-7f);
        cost[(long long)b * K * K + e] = __fsub_rn(1.0f, __fdiv_rn(d, den));
    }
}

// one warp per matrix
__global__ void __launch_bounds__(32) hungarian_kernel(const float* __restrict__ cost, int n, int* __restrict__ col_of_row) {
    __shared__ State st;
    __shared__ float a[MAXN * MAXN];
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int e = lane; e < n * n; e += 32) a[e] = cost[(long long)b * n * n + e];
    __syncwarp();
    auto reduce_min = [&](double& v, int& j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(FULL, v, o);
            const int oj = __shfl_xor_sync(FULL, j, o);
            if (ov < v || (ov == v && oj < j)) { v = ov; j = oj; }
        }
    };
    auto sync = [] { __syncwarp(); };
    hungarian<32>(a, n, st, lane, reduce_min, sync, col_of_row + (long long)b * n);
}

}  // namespace assign
}  // namespace pn

using namespace pn;

extern "C" int pn_iou_cost(const int* pred, const int* gt, int B, int N, int K, float* cost, int* bad, void* stream) {
    PN_REQUIRE(pred && gt && cost && bad, "pn_iou_cost: null pointer");
    PN_REQUIRE(B > 0 && N > 0 && K > 0 && K <= assign::MAXN, "pn_iou_cost: need B, N > 0 and 0 < K <= %d (B=%d N=%d K=%d)",
               assign::MAXN, B, N, K);
    PN_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), (cudaStream_t)stream));
    assign::iou_cost_kernel<<<B, 256, (K * K + 2 * K) * sizeof(int), (cudaStream_t)stream>>>(pred, gt, N, K, cost, bad);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("iou_cost_kernel");
    return PN_OK;
}

extern "C" int pn_hungarian(const float* cost, int B, int n, int* col_of_row, void* stream) {
    PN_REQUIRE(cost && col_of_row, "pn_hungarian: null pointer");
    PN_REQUIRE(B > 0 && n > 0 && n <= assign::MAXN, "pn_hungarian: need B > 0 and 0 < n <= %d (B=%d n=%d)", assign::MAXN, B, n);
    assign::hungarian_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(cost, n, col_of_row);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("hungarian_kernel");
    return PN_OK;
}
