// Error channel + small queries of the C ABI (include/igcn_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace igcn {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// SM clock in kHz (= cycles per millisecond of clock64()); used to turn timeouts into cycle budgets.
long long sm_clock_khz() {
    static long long cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 1965000;
    if (cached[dev] == 0) {
        int khz = 0;
        if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0) khz = 1965000;
        cached[dev] = khz;
    }
    return cached[dev];
}

// programmatic dependent launch for the library's kernels (common.cuh): IGCN_PDL=1 / 0, default off
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("IGCN_PDL"); return e && e[0] == '1'; }();
    return on;
}
}  // namespace igcn

extern "C" const char* igcn_last_error(void) { return igcn::g_err; }
extern "C" int igcn_version(void) { return 100; }
extern "C" int igcn_sm_count(void) { return igcn::sm_count(); }
