// host twin of csrc/assign.cu::hungarian_kernel: the same template with one lane
#define PN_ASSIGN_HD inline
#include "assign.cuh"

extern "C" void hungarian_host(const float* cost, int n, int* col_of_row) {
    pn::assign::State st;
    auto reduce_min = [](double&, int&) {};
    auto sync = [] {};
    pn::assign::hungarian<1>(cost, n, st, 0, reduce_min, sync, col_of_row);
}
