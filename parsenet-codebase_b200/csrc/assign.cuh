// Hungarian algorithm (Kuhn-Munkres with potentials, O(n^3)) on one n x n cost matrix, written so that the same code runs as
// one warp on the device (lanes split the columns) and as plain C++ on the host (tests/c/assign_host.cpp runs it with one
// "lane").  Ties: the lowest column index wins, exactly as the serial algorithm picks it.
#pragma once
#include <stdint.h>

#ifndef PN_ASSIGN_HD
#define PN_ASSIGN_HD __device__ __forceinline__
#endif

namespace pn {
namespace assign {

constexpr int MAXN = 64;

struct State {                      // 1-based like the classical formulation; index 0 is the virtual column
    double u[MAXN + 1], v[MAXN + 1], minv[MAXN + 1];
    int p[MAXN + 1], way[MAXN + 1];
    unsigned char used[MAXN + 1];
};

// LANES = 1 on the host, 32 on the device; lane = this thread's lane; reduce_min(value, index) -> the pair with the smallest
// value, lowest index on ties, agreed on by all lanes; sync() orders the lanes' accesses to the shared state.
template <int LANES, typename ReduceMin, typename Sync>
PN_ASSIGN_HD void hungarian(const float* a, int n, State& s, int lane, ReduceMin reduce_min, Sync sync, int* col_of_row) {
    for (int j = lane; j <= n; j += LANES) { s.u[j] = 0.0; s.v[j] = 0.0; s.p[j] = 0; s.way[j] = 0; }
    sync();
    for (int i = 1; i <= n; ++i) {
        if (lane == 0) s.p[0] = i;
        for (int j = lane; j <= n; j += LANES) { s.minv[j] = 1e300; s.used[j] = 0; }
        sync();
        int j0 = 0;
        do {
            if (lane == 0) s.used[j0] = 1;
            sync();
            const int i0 = s.p[j0];
            double best = 1e300;
            int bestj = 0x7fffffff;
            for (int j = 1 + lane; j <= n; j += LANES) {
                if (!s.used[j]) {
                    const double cur = (double)a[(i0 - 1) * n + (j - 1)] - s.u[i0] - s.v[j];
                    if (cur < s.minv[j]) { s.minv[j] = cur; s.way[j] = j0; }
                    if (s.minv[j] < best) { best = s.minv[j]; bestj = j; }      // ascending j: first minimum kept
                }
            }
            reduce_min(best, bestj);
            const double delta = best;
            sync();
            for (int j = lane; j <= n; j += LANES) {
                if (s.used[j]) { s.u[s.p[j]] += delta; s.v[j] -= delta; }       // p[j] distinct over used j: no write conflict
                else s.minv[j] -= delta;
            }
            sync();
            j0 = bestj;
        } while (s.p[j0] != 0);
        if (lane == 0) {
            do { const int j1 = s.way[j0]; s.p[j0] = s.p[j1]; j0 = j1; } while (j0);
        }
        sync();
    }
    for (int j = 1 + lane; j <= n; j += LANES) col_of_row[s.p[j] - 1] = j - 1;
    sync();
}

}  // namespace assign
}  // namespace pn
