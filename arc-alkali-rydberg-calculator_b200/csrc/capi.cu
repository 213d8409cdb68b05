// capi.cu -- library-level C-ABI entry points (error reporting, device info).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void arcb200_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int arcb200_abi_version(void) { return ARCB200_ABI_VERSION; }

extern "C" const char *arcb200_last_error(void) { return g_err; }

extern "C" int arcb200_device_info(int device, int64_t *props)
{
    cudaDeviceProp p;
    ARCB200_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    int cl = 0, clock = 0;
    cudaDeviceGetAttribute(&cl, cudaDevAttrClusterLaunch, device);
    cudaDeviceGetAttribute(&clock, cudaDevAttrClockRate, device);
    props[0] = p.multiProcessorCount;
    props[1] = p.major;
    props[2] = p.minor;
    props[3] = p.l2CacheSize;
    props[4] = (int64_t)p.sharedMemPerBlockOptin;
    props[5] = (int64_t)(p.totalGlobalMem >> 20);
    props[6] = cl;
    props[7] = clock;
    if (p.major != 10) {
        arcb200_set_error("libarcb200 is built for sm_100a only; device %d is sm_%d%d",
                          device, p.major, p.minor);
        return -3;
    }
    return 0;
}
