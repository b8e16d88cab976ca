// host twin of the warp K-th select of csrc/meanshift_tma.cu (ms_kth_select_kernel): the lane-local code of
// csrc/kth_select.cuh run for 32 emulated lanes; the cross-lane reduction is a plain sum here.
#define PN_KTH_HD inline
#include "kth_select.cuh"
#include <bit>

extern "C" unsigned kth_select_host(const unsigned* keys, int n, int K) {
    using namespace pn::kthsel;
    uint32_t B[32][32], act[32];
    for (int lane = 0; lane < 32; ++lane) {
        act[lane] = 0u;
        for (int r = 0; r < 32; ++r) {
            const int p = r * 32 + lane;
            B[lane][r] = p < n ? keys[p] : 0xffffffffu;
            if (p < n) act[lane] |= reg_bit(r);
        }
        bit_transpose32(B[lane]);
    }
    int need = K < n ? K : n;
    unsigned prefix = 0u;
    for (int i = 0; i < 32; ++i) {
        int c = 0;
        for (int lane = 0; lane < 32; ++lane) c += std::popcount(step_zeros(act[lane], B[lane][i]));
        const bool zero = c >= need;
        if (!zero) { need -= c; prefix |= 1u << (31 - i); }
        for (int lane = 0; lane < 32; ++lane) act[lane] = step_next(act[lane], B[lane][i], zero);
    }
    return prefix;
}
