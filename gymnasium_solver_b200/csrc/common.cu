// common.cu — error plumbing and device queries of the C-ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count(int device) {
    static int cached[64];
    if (device < 0 || device >= 64) return 148;
    if (cached[device] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
        cached[device] = n;
    }
    return cached[device];
}

}  // namespace gs

extern "C" {
int gs_version(void) { return GS_VERSION; }
const char* gs_last_error(void) { return gs::g_err; }
int gs_device_sm_count(int device) { return gs::sm_count(device); }
}
