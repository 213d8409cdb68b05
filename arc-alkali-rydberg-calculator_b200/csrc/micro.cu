// micro.cu -- FP64 tensor-core (DMMA) peak of the device, measured the way the eigensolver uses
// it: mma.sync.aligned.m8n8k4.f64, register operands only, 16 independent accumulators per
// warp, 8 warps per SM.  MEASURED_PEAKS.json carries no FP64 figure, so bench.py reports the
// roofline of the DMMA-bound kernels against this number (measured in the same process).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double seed)
{
    double c[16][2];
    const double a = seed * (threadIdx.x + 1), b = seed * (threadIdx.x * 3 + 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// plain FP64 FMA pipe: 16 independent chains per thread, 8 warps per SM
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed)
{
    double c[16];
    const double a = 1.0 + seed * (threadIdx.x + 1), b = seed * (threadIdx.x * 3 + 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = seed * i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

}  // namespace

extern "C" int arcb200_dfma_peak(int device, void *stream, double *d_scratch, double *h_tflops)
{
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0;
    ARCB200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaEvent_t e0, e1;
    ARCB200_CUDA_CHECK(cudaEventCreate(&e0));
    ARCB200_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 40000;
    dfma_peak_kernel<<<sms, 256, 0, st>>>(d_scratch, 200, 1e-9);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        ARCB200_CUDA_CHECK(cudaEventRecord(e0, st));
        dfma_peak_kernel<<<sms, 256, 0, st>>>(d_scratch, iters, 1e-9);
        ARCB200_CUDA_CHECK(cudaEventRecord(e1, st));
        ARCB200_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        ARCB200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)sms * 256 * iters * 16 * 2.0;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ARCB200_LAUNCH_CHECK("dfma_peak_kernel");
    *h_tflops = best;
    return 0;
}

extern "C" int arcb200_dmma_peak(int device, void *stream, double *d_scratch, double *h_tflops)
{
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0;
    ARCB200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaEvent_t e0, e1;
    ARCB200_CUDA_CHECK(cudaEventCreate(&e0));
    ARCB200_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 40000;
    dmma_peak_kernel<<<sms, 256, 0, st>>>(d_scratch, 200, 1e-9);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        ARCB200_CUDA_CHECK(cudaEventRecord(e0, st));
        dmma_peak_kernel<<<sms, 256, 0, st>>>(d_scratch, iters, 1e-9);
        ARCB200_CUDA_CHECK(cudaEventRecord(e1, st));
        ARCB200_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        ARCB200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)sms * 8 * iters * 16 * (2.0 * 8 * 8 * 4);
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ARCB200_LAUNCH_CHECK("dmma_peak_kernel");
    *h_tflops = best;
    return 0;
}
