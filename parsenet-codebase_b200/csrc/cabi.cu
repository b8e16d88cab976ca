// C-ABI glue: error string, launch counter, version.  All entry points are declared in include/parsenet_b200.h.
#include "common.cuh"
#include <stdarg.h>

namespace pn {
static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }
}  // namespace pn

extern "C" const char* pn_last_error() { return pn::last_error(); }
extern "C" unsigned long long pn_launch_count() { return pn::g_launch_count; }
extern "C" void pn_reset_launch_count() { pn::g_launch_count = 0; }
extern "C" int pn_abi_version() { return 1; }
