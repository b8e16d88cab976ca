// Shared device/host helpers for the parsenet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace pn {

// ---------------------------------------------------------------- error plumbing (C-ABI returns int)
enum : int {
    PN_OK = 0,
    PN_ERR_ARG = 1,        // bad argument (shape, alignment, unsupported size)
    PN_ERR_CUDA = 2,       // CUDA runtime error at launch
    PN_ERR_UNSUPPORTED = 3
};

void set_error(const char* fmt, ...);
const char* last_error();

#define PN_REQUIRE(cond, ...)                      \
    do {                                           \
        if (!(cond)) {                             \
            ::pn::set_error(__VA_ARGS__);          \
            return ::pn::PN_ERR_ARG;               \
        }                                          \
    } while (0)

#define PN_LAUNCH_CHECK(name)                                                          \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            ::pn::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
            return ::pn::PN_ERR_CUDA;                                                  \
        }                                                                              \
    } while (0)

#define PN_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            ::pn::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
            return ::pn::PN_ERR_CUDA;                                                  \
        }                                                                              \
    } while (0)

// counts kernel launches issued through the C-ABI (bench.py reports it as gpu_launches)
extern unsigned long long g_launch_count;
#define PN_COUNT_LAUNCH() (++::pn::g_launch_count)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// monotone float -> uint mapping (larger float -> larger uint), total order incl. -0 < +0
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}

// block-wide sum for blockDim.x <= 1024 (result valid in all threads)
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 elems */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    T r = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
    if (w == 0) r = warp_sum(r);
    if (threadIdx.x == 0) smem[0] = r;
    __syncthreads();
    r = smem[0];
    return r;
}
#endif

}  // namespace pn
