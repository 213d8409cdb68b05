// common.cuh -- shared helpers for libarcb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "arcb200.h"

#define ARCB200_FULL_MASK 0xffffffffu

void arcb200_set_error(const char *fmt, ...);

#define ARCB200_CUDA_CHECK(expr)                                                   \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            arcb200_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                              cudaGetErrorString(_e));                             \
            return -1;                                                             \
        }                                                                          \
    } while (0)

// argument validation at the C-ABI: checked on the host before any CUDA call, so a bad call
// returns -1 with a message instead of hanging (batch = 0) or faulting on the device
#define ARCB200_REQUIRE(cond, ...)                                                 \
    do {                                                                           \
        if (!(cond)) {                                                             \
            arcb200_set_error(__VA_ARGS__);                                        \
            return -1;                                                             \
        }                                                                          \
    } while (0)

#define ARCB200_LAUNCH_CHECK(name)                                                 \
    do {                                                                           \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) {                                                   \
            arcb200_set_error("launch of %s failed: %s", name,                     \
                              cudaGetErrorString(_e));                             \
            return -2;                                                             \
        }                                                                          \
    } while (0)

// Helper streams of the library (sub-batches of the band reduction, the opt-in half-batch pipeline): one pool
// per device, created on first use and kept for the life of the process.  A host thread takes
// ArcHelperGuard while it ENQUEUES work that touches the pool (the GPU work itself is asynchronous), so
// concurrent callers on one device cannot interleave their event records / waits.
struct ArcHelperStreams {
    static constexpr int MAXS = 8;
    cudaStream_t stream[MAXS];
    cudaEvent_t ev[MAXS + 1];
    cudaStream_t pipe_stream[2];
    cudaEvent_t pipe_ev[3];
};
struct ArcHelperGuard {
    ArcHelperGuard();
    ~ArcHelperGuard();
    ArcHelperGuard(const ArcHelperGuard &) = delete;
    ArcHelperGuard &operator=(const ArcHelperGuard &) = delete;
    // pool of `device` (nullptr after arcb200_set_error if a stream or event cannot be created)
    ArcHelperStreams *pool(int device);
};

static __device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ARCB200_FULL_MASK, v, o);
    return v;
}

static __device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(ARCB200_FULL_MASK, v, o));
    return v;
}
