// Version, error reporting and device check of libkgan.so.
#include <stdarg.h>

#include "common.cuh"

namespace kgan {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace kgan

extern "C" int kgan_version(void) { return 100; }
extern "C" const char* kgan_last_error(void) { return kgan::g_err; }
extern "C" int kgan_device_ok(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}
