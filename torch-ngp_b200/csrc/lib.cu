// lib.cu — library-level plumbing of libngp_b200.so (error text, launch counter, device info).
#include "common.cuh"
#include <stdarg.h>

namespace ngp {

thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
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

}  // namespace ngp

extern "C" const char* ngp_last_error(void) { return ngp::g_err; }
extern "C" int ngp_version(void) { return 1; }
extern "C" const char* ngp_build_arch(void) { return "sm_100a"; }
extern "C" uint64_t ngp_launch_count(void) { return ngp::g_launches.load(); }
extern "C" void ngp_reset_launch_count(void) { ngp::g_launches.store(0); }
