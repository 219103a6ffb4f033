// error.cu -- thread-local error text + version for the C ABI.
#include "gsb_common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void gsb_set_error(int code, const char *what, const char *file, int line) {
    snprintf(g_err, sizeof(g_err), "gsplat_b200 error %d: %s (%s:%d)", code, what, file, line);
}

extern "C" const char *gsb_last_error(void) { return g_err; }
extern "C" int gsb_version(void) { return GSB_VERSION; }
