// core.cu -- library-wide state of libmcshade: error reporting and ABI version.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";

void mcs_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {
int mcs_abi_version(void) { return MCS_ABI_VERSION; }
const char *mcs_last_error(void) { return g_err; }
}
