// Per-row candidate buffers of the kNN kernels: warp-cooperative compaction (quickselect on 32-bit ordered keys) and the
// final bitonic sort by (distance descending, index ascending).  Shared by knn.cu (loader-thread kernel) and knn_tma.cu
// (TMA-staged kernel); BI = storage type of the candidate indices in the buffer (int, or unsigned short when N < 65536).
#pragma once
#include "common.cuh"

// the full sort of a row buffer (bitonic network, ~2 k instructions at CAP = 128) is inlined by default; a kernel that calls it
// from several places defines PN_KNN_SORT_ATTR as __noinline__ before including this header to keep ONE copy
#ifndef PN_KNN_SORT_ATTR
#define PN_KNN_SORT_ATTR __forceinline__
#endif

namespace pn {
namespace knn {

// ---- warp bitonic sort of CAP (key desc-distance / asc-index) entries, R = CAP/32 per lane ----
template <int R>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&key)[R], int lane) {
    constexpr int CAP = R * 32;
#pragma unroll
    for (int k = 2; k <= CAP; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int rj = j >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rj) == 0) {
                        const int p = r * 32;  // lane bits do not matter for k>=64 & j>=32 direction test below
                        // direction depends on bit k of the element index p = r*32+lane (k >= 64 here, or k==CAP)
                        bool up = ((p & k) == 0);
                        unsigned long long a = key[r], b = key[r ^ rj];
                        bool sw = up ? (a > b) : (a < b);
                        key[r] = sw ? b : a;
                        key[r ^ rj] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int p = r * 32 + lane;
                    unsigned long long a = key[r];
                    unsigned long long b = __shfl_xor_sync(FULL, a, j);
                    bool up = ((p & k) == 0);
                    bool lower = ((lane & j) == 0);
                    bool keep_min = (up == lower);
                    key[r] = keep_min ? (a < b ? a : b) : (a < b ? b : a);
                }
            }
        }
    }
}

__device__ __forceinline__ unsigned long long make_key(float d, int j) {
    return ((unsigned long long)(~f2ord(d)) << 32) | (unsigned)j;
}

// sort the row buffer, keep the best k, return new count; *tau_out = k-th best value (or -inf)
template <int CAP, typename BI = int>
__device__ PN_KNN_SORT_ATTR int compact_row(float* bv, BI* bi, int n, int k, int lane, float* tau_out) {
    constexpr int R = CAP / 32;
    unsigned long long key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int p = r * 32 + lane;
        key[r] = (p < n) ? make_key(bv[p], (int)bi[p]) : ~0ull;
    }
    __syncwarp();
    warp_bitonic_sort<R>(key, lane);
    int nn = n < k ? n : k;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int p = r * 32 + lane;
        if (p < nn) {
            bv[p] = ord2f(~(uint32_t)(key[r] >> 32));
            bi[p] = (BI)(uint32_t)(key[r] & 0xffffffffu);
        }
    }
    __syncwarp();
    *tau_out = (nn == k) ? bv[k - 1] : -INFINITY;
    return nn;
}

// Mid-stream compaction: keep the best k of the n buffered entries of one row WITHOUT ordering them (the buffer is
// sorted once, at the end).  Warp-cooperative quickselect on the 32-bit ordered distance keys: ~4x fewer
// instructions than the bitonic sort of 64-bit keys.  Exact: entries tied with the k-th best distance are resolved
// by index through the sort path (rare).  Returns the new count, *tau_out = k-th best value.
template <int CAP, typename BI = int>
__device__ __forceinline__ int compact_select(float* bv, BI* bi, int n, int k, int lane, float* tau_out) {
    constexpr int R = CAP / 32;
    if (n <= k) return compact_row<CAP, BI>(bv, bi, n, k, lane, tau_out);
    uint32_t u[R];
    int id[R];
    bool act[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int p = r * 32 + lane;
        act[r] = p < n;
        u[r] = act[r] ? f2ord(bv[p]) : 0u;
        id[r] = act[r] ? (int)bi[p] : 0;
    }
    __syncwarp();
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) valid[r] = act[r];
    int need = k;                 // the threshold T is the need-th largest key among the active entries
    uint32_t T = 0u;
    for (;;) {
        // pivot: first active key in (lane, register) order
        uint32_t mine = 0u;
        bool have = false;
#pragma unroll
        for (int r = R - 1; r >= 0; --r)
            if (act[r]) { mine = u[r]; have = true; }
        const unsigned hm = __ballot_sync(FULL, have);
        const uint32_t piv = __shfl_sync(FULL, mine, __ffs(hm) - 1);
        int cg = 0, ce = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cg += (act[r] && u[r] > piv) ? 1 : 0;
            ce += (act[r] && u[r] == piv) ? 1 : 0;
        }
        cg = __reduce_add_sync(FULL, cg);
        ce = __reduce_add_sync(FULL, ce);
        if (cg >= need) {
#pragma unroll
            for (int r = 0; r < R; ++r) act[r] = act[r] && (u[r] > piv);
        } else if (cg + ce >= need) {
            T = piv;
            break;
        } else {
            need -= cg + ce;
#pragma unroll
            for (int r = 0; r < R; ++r) act[r] = act[r] && (u[r] < piv);
        }
    }
    int G = 0, E = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        G += (valid[r] && u[r] > T) ? 1 : 0;
        E += (valid[r] && u[r] == T) ? 1 : 0;
    }
    G = __reduce_add_sync(FULL, G);
    E = __reduce_add_sync(FULL, E);
    if (G + E != k) return compact_row<CAP, BI>(bv, bi, n, k, lane, tau_out);   // index tie-break needed (buffer untouched)
    int base = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool keep = valid[r] && u[r] >= T;
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            bv[pos] = ord2f(u[r]);
            bi[pos] = (BI)id[r];
        }
        base += __popc(m);
    }
    __syncwarp();
    *tau_out = ord2f(T);
    return k;
}

}  // namespace knn
}  // namespace pn
