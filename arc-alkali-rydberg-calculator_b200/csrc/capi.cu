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

// ---- helper-stream pools (see common.cuh)
#include <map>
#include <mutex>
static std::recursive_mutex g_helper_mu;
static std::map<int, ArcHelperStreams *> g_helper_pools;

ArcHelperGuard::ArcHelperGuard() { g_helper_mu.lock(); }
ArcHelperGuard::~ArcHelperGuard() { g_helper_mu.unlock(); }

ArcHelperStreams *ArcHelperGuard::pool(int device)
{
    auto it = g_helper_pools.find(device);
    if (it != g_helper_pools.end()) return it->second;
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != device && cudaSetDevice(device) != cudaSuccess) {
        arcb200_set_error("helper streams: cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    ArcHelperStreams *p = new ArcHelperStreams();
    bool ok = true;
    for (int s = 0; s < ArcHelperStreams::MAXS && ok; ++s)
        ok = cudaStreamCreateWithFlags(&p->stream[s], cudaStreamNonBlocking) == cudaSuccess;
    for (int s = 0; s <= ArcHelperStreams::MAXS && ok; ++s)
        ok = cudaEventCreateWithFlags(&p->ev[s], cudaEventDisableTiming) == cudaSuccess;
    for (int s = 0; s < 2 && ok; ++s)
        ok = cudaStreamCreateWithFlags(&p->pipe_stream[s], cudaStreamNonBlocking) == cudaSuccess;
    for (int s = 0; s < 3 && ok; ++s)
        ok = cudaEventCreateWithFlags(&p->pipe_ev[s], cudaEventDisableTiming) == cudaSuccess;
    if (cur >= 0 && cur != device) cudaSetDevice(cur);
    if (!ok) {
        arcb200_set_error("helper streams: stream / event creation failed on device %d: %s", device,
                          cudaGetErrorString(cudaGetLastError()));
        delete p;                      // (the handles created so far are left to the driver)
        return nullptr;
    }
    g_helper_pools[device] = p;
    return p;
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
