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

// Tile-stream probe: the memory side of the band reduction's symmetric product in isolation.  One CTA of
// eight warps per SM and matrix; every warp streams 32 x 32 FP64 tiles of its matrix into a shared-memory
// ring by cp.async (DEPTH tiles ahead) and spends 128 DMMAs on each (the tile's fragments are the A
// operand, so the data is really consumed).  mode bit 1 (value 2): end every step like the lockstep
// kernel does (wait for the next tile, block barrier).  mode bit 0 = 0: column-major matrix, a tile is 32 segments of 256 B
// `lda` doubles apart (the layout of the sweep workspace); bit 0 = 1: tile-major, a tile is 8 KB contiguous.
// Reports the time per tile against the pure DMMA time.
template <int DEPTH>
__global__ void __launch_bounds__(256, 1) tile_probe_kernel(const double *base, size_t mat_stride, int lda, int ntiles,
                                                            int mode, double *out)
{
    extern __shared__ double psm[];
    constexpr int TS = 36;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, grp = lane >> 2, tig = lane & 3;
    const double *A = base + (size_t)blockIdx.x * mat_stride;
    double *ring = psm + wid * ((DEPTH + 1) * 32 * TS);
    const int NBr = lda / 32;
    auto issue = [&](int t) {
        if (t < ntiles) {
            const int idx = t * 8 + wid;
            const double *src;
            size_t step;
            if ((mode & 1) == 0) {
                const int I = idx % NBr, J = (idx / NBr) % NBr;
                src = A + (size_t)(32 * J + (lane >> 4)) * lda + 32 * I + (lane & 15) * 2;
                step = 2 * (size_t)lda;
            } else {
                src = A + (size_t)idx * 1024 + (lane >> 4) * 32 + (lane & 15) * 2;
                step = 64;
            }
            double *dst = ring + (t % (DEPTH + 1)) * (32 * TS) + (lane >> 4) * TS + (lane & 15) * 2;
#pragma unroll 1
            for (int k = 0; k < 16; ++k) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
                src += step;
                dst += 2 * TS;
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    double acc[4][4][2];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
    for (int t = 0; t < DEPTH; ++t) issue(t);
    for (int t = 0; t < ntiles; ++t) {
        issue(t + DEPTH);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(DEPTH) : "memory");
        __syncwarp();
        const double *pd = ring + (t % (DEPTH + 1)) * (32 * TS) + tig * TS + grp;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            double fa[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) fa[x] = pd[4 * ks * TS + 8 * x];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const double b = 1e-3 * (y + 1);
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(acc[x][y][0]), "+d"(acc[x][y][1])
                                 : "d"(fa[x]), "d"(b));
            }
        }
        __syncwarp();
        if (mode & 2) {            // the band-reduction kernel's step end: next tile landed, block barrier
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();
        }
    }
    double s = 0.0;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) s += acc[x][y][0] + acc[x][y][1];
    if (s == 123.456) out[0] = s;
}

}  // namespace

// h_out[0] = ms per launch, h_out[1] = fraction of the DMMA rate (128 DMMAs per tile), h_out[2] = GB/s
extern "C" int arcb200_tile_stream_probe(int device, void *stream, const double *d_buf, int64_t buf_doubles,
                                         int32_t lda, int32_t ntiles_per_warp, int32_t mode, int32_t depth,
                                         double dmma_tflops, double *h_out)
{
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0;
    ARCB200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const size_t mat_stride = (size_t)lda * lda;
    if ((int64_t)(mat_stride * sms) > buf_doubles || depth < 1 || depth > 2) {
        arcb200_set_error("tile probe: buffer too small or depth not in {1, 2}");
        return -1;
    }
    const size_t smem = (size_t)8 * (depth + 1) * 32 * 36 * sizeof(double);
    auto launch = [&]() {
        if (depth == 1) tile_probe_kernel<1><<<sms, 256, smem, st>>>(d_buf, mat_stride, lda, ntiles_per_warp, mode, (double *)d_buf);
        else tile_probe_kernel<2><<<sms, 256, smem, st>>>(d_buf, mat_stride, lda, ntiles_per_warp, mode, (double *)d_buf);
    };
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(tile_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * 2 * 32 * 36 * 8)));
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(tile_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * 3 * 32 * 36 * 8)));
    cudaEvent_t e0, e1;
    ARCB200_CUDA_CHECK(cudaEventCreate(&e0));
    ARCB200_CUDA_CHECK(cudaEventCreate(&e1));
    launch();
    ARCB200_CUDA_CHECK(cudaEventRecord(e0, st));
    launch();
    ARCB200_CUDA_CHECK(cudaEventRecord(e1, st));
    ARCB200_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    ARCB200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ARCB200_LAUNCH_CHECK("tile_probe_kernel");
    const double tiles = (double)sms * 8 * ntiles_per_warp;
    h_out[0] = ms;
    h_out[1] = (tiles * 128 * 512.0 / (ms * 1e-3) / 1e12) / dmma_tflops;
    h_out[2] = tiles * 8192.0 / (ms * 1e-3) / 1e9;
    return 0;
}

namespace {
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
