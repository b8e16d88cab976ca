// Host twin of csrc/fitsolve.cuh for the CPU test-suite (tests/test_cpu_fitsolve.py): the same source compiled with g++.
#include "fitsolve.cuh"
extern "C" void fitsolve_host(const double* mom, const int* kind, int S, int rows, double* par, double* jac, int* bad) {
    for (int s = 0; s < S; ++s)
        bad[s] = pn::fitsolve::solve_segment_full(kind[s], mom + 55 * s, rows, par + 8 * s, jac + 8 * 55 * s) ? 1 : 0;
}
