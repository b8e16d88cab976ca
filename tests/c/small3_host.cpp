// Host twin of csrc/small3.cuh for the CPU test-suite (tests/test_cpu_small3.py): the same source compiled with g++.
#include "small3.cuh"
extern "C" void small3_eigh_host(const double* G, int S, double* w, double* V) {
    for (int s = 0; s < S; ++s) pn::small3::eigh3(G + 9 * s, w + 3 * s, V + 9 * s);
}
extern "C" void small3_lstsq_host(const double* AtA, const double* AtY, int S, int rows, double eps32, double* x,
                                  double* minv, double* lam) {
    for (int s = 0; s < S; ++s)
        pn::small3::lstsq3(AtA + 9 * s, AtY + 3 * s, rows, eps32, x + 3 * s, minv + 9 * s, lam + s);
}
