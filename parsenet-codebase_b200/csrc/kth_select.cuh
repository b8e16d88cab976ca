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
