// K-th smallest of up to 1024 32-bit keys held 32 per lane by one warp: bitwise radix select on a bit-transposed copy.
// Lane-local state: B[i] = one word per key BIT (i = 31 - bit) whose bit (31 - r) is that bit of the lane's r-th key, and an
// active mask over the same positions; a radix step is then AND / POPC on two words instead of a pass over 32 registers.
// PN_KTH_HD lets tests/c/kth_select_host.cpp run the same lane-local code on the host.
#pragma once
#include <stdint.h>

#ifndef PN_KTH_HD
#define PN_KTH_HD __device__ __forceinline__
#endif

namespace pn {
namespace kthsel {

// in-place transpose of the 32 x 32 bit matrix a (Hacker's Delight 7-3: row 0 / bit 31 are the top-left corner):
// afterwards bit (31 - r) of a[31 - bit] is bit `bit` of the original a[r]
PN_KTH_HD void bit_transpose32(uint32_t (&a)[32]) {
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if ((k & j) == 0) {
                const uint32_t t = (a[k] ^ (a[k + j] >> j)) & m;
                a[k] ^= t;
                a[k + j] ^= t << j;
            }
        }
        m ^= m << (j >> 1);
    }
}

// mask position of register r
PN_KTH_HD uint32_t reg_bit(int r) { return 1u << (31 - r); }

// one radix step for key bit (31 - i): `zeros` = this lane's active keys whose bit is 0
PN_KTH_HD uint32_t step_zeros(uint32_t act, uint32_t Bi) { return act & ~Bi; }
PN_KTH_HD uint32_t step_next(uint32_t act, uint32_t Bi, bool keep_zeros) { return keep_zeros ? (act & ~Bi) : (act & Bi); }

}  // namespace kthsel
}  // namespace pn

#ifdef __CUDACC__
// one warp per row: out[row] = the K-th smallest of the row's first n float values (stored as raw float bits with row pitch
// 1024 * NW words), as a float.  The sample stage of the bracketed selections (meanshift_tma.cu keeps its own copy with the
// exact-recompute tail; knn_tc.cu uses this one).
namespace pn {
namespace kthsel {
template <int NW>
__global__ void __launch_bounds__(256) kth_smallest_rows_kernel(const unsigned* __restrict__ vals, int n, int K, long long rows_total,
                                                                float* __restrict__ out) {
    constexpr int CAP = 1024 * NW;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows_total) return;
    const unsigned* kr = vals + row * CAP;
    unsigned Bt[NW][32];
    unsigned act[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        act[w] = 0u;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int p = w * 1024 + r * 32 + lane;
            const bool valid = p < n;
            const unsigned u = valid ? kr[p] : 0u;
            Bt[w][r] = valid ? ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) : 0xffffffffu;      // f2ord
            act[w] |= valid ? reg_bit(r) : 0u;
        }
        bit_transpose32(Bt[w]);
    }
    int need = K < n ? K : n;
    unsigned prefix = 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        int c = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) c += __popc(step_zeros(act[w], Bt[w][i]));
        c = __reduce_add_sync(0xffffffffu, c);
        const bool zero = c >= need;
        if (!zero) { need -= c; prefix |= 1u << (31 - i); }
#pragma unroll
        for (int w = 0; w < NW; ++w) act[w] = step_next(act[w], Bt[w][i], zero);
    }
    if (lane == 0) out[row] = __uint_as_float((prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix);      // ord2f
}
}  // namespace kthsel
}  // namespace pn
#endif
