// FP64 tensor-core (DMMA) peak microbenchmark for sm_100a: mma.sync f64 shapes m8n8k4, m16n8k4,
// m16n8k8, m16n8k16, and plain DFMA, register-only, NACC independent accumulators per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dmma_peak tools/micro/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double seed)
{
    double c[NACC][4];
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i);
    for (int i = 0; i < 4; ++i) b[i] = seed * (threadIdx.x * 3 + i);
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            } else if (SHAPE == 3) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = fma(a[j], b[j], c[i][j]);
            }
        }
    }
    double s = 0.0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}

template <int SHAPE, int NACC>
void run(const char *name, double flop_per_op, int ctas_per_sm, int threads)
{
    double *d; cudaMalloc(&d, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dmma<SHAPE, NACC><<<148 * ctas_per_sm, threads>>>(d, 100, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_dmma<SHAPE, NACC><<<148 * ctas_per_sm, threads>>>(d, iters, 1e-9);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = 148.0 * ctas_per_sm * threads / 32;
    const double flops = warps * iters * NACC * flop_per_op;
    printf("%-10s nacc=%2d warps/SM=%2d: %8.2f TFLOP/s (%.2f ms)%s\n", name, NACC, ctas_per_sm * threads / 32,
           flops / ms / 1e9, ms, cudaGetLastError() == cudaSuccess ? "" : " ERROR");
    cudaFree(d);
}

int main()
{
    for (int w = 0; w < 3; ++w) {
        const int threads = w == 0 ? 128 : 256, ctas = w == 2 ? 2 : 1;
        run<0, 16>("m8n8k4", 2.0 * 8 * 8 * 4, ctas, threads);
        run<0, 8>("m8n8k4", 2.0 * 8 * 8 * 4, ctas, threads);
        run<1, 8>("m16n8k4", 2.0 * 16 * 8 * 4, ctas, threads);
        run<2, 8>("m16n8k8", 2.0 * 16 * 8 * 8, ctas, threads);
        run<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, ctas, threads);
        run<3, 4>("m16n8k16", 2.0 * 16 * 8 * 16, ctas, threads);
        run<4, 8>("dfma", 2.0 * 4 * 32, ctas, threads);
    }
    return 0;
}
