// Error state and library identification for libdcta.so.
#include <stdarg.h>

#include "common.cuh"

namespace dcta {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace dcta

extern "C" const char* dcta_last_error(void) { return dcta::g_err; }
extern "C" int dcta_abi_version(void) { return 2; }
extern "C" int dcta_compiled_arch(void) { return 100; }
