// Error plumbing and device queries of the lwb_b200 C ABI.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace lwb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_slot()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int sm_count()
{
    static int cached[kMaxDevices] = {};
    const int slot = device_slot();
    if (cached[slot] <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot) == cudaSuccess && n > 0)
            cached[slot] = n;
        else
            return 148;
    }
    return cached[slot];
}

bool pdl_enabled()
{
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("LWB_PDL"); cached = (e && atoi(e) == 0) ? 0 : 1; }
    return cached != 0;
}

}  // namespace lwb

extern "C" int lwb_version(void) { return 100; }

extern "C" const char* lwb_last_error(void) { return lwb::g_err; }

extern "C" int lwb_device_info(int* sm_count, int* cc_major, int* cc_minor)
{
    int dev = 0, n = 0, ma = 0, mi = 0;
    if (sm_count) *sm_count = 0;
    if (cc_major) *cc_major = 0;
    if (cc_minor) *cc_minor = 0;
    LWB_CUDA_OK(cudaGetDevice(&dev));
    LWB_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    LWB_CUDA_OK(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    LWB_CUDA_OK(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    return LWB_OK;
}
