// eig_gemm.cuh -- grouped FP64 GEMM on the tensor cores (DMMA, mma.sync.m8n8k4.f64)
// used by the divide-and-conquer merges:  C(:, cmap[n]) = A(:, aidx[k]) * B(k, n).
// All operands column-major.  One CTA computes a 64x64 tile, 8 warps (2 x 4), each warp
// 32x16 = 4x2 DMMA tiles; A/B tiles are staged through shared memory with cp.async
// (double buffered, 16-deep K slabs); shared-memory strides (68 / 20 doubles) make the
// fragment loads bank-conflict free.  Problem sizes are read from device memory (they
// depend on deflation), so the grid is sized for the worst case and empty tiles exit.
#pragma once
#include "common.cuh"

struct GemmProblem {
    const double *A; int lda; const int32_t *aidx;   // A(:, aidx[k]) (aidx may be null)
    const double *B; int ldb;
    double *C; int ldc; const int32_t *cmap;         // output column n -> cmap[n] (or null)
    int M, Ncols, K;
};

namespace gemm_detail {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int AS = 68;   // As[k][m] row stride (doubles)
constexpr int BS = 20;   // Bs[n][k] row stride (doubles)

__device__ __forceinline__ void cp_async8(void *dst, const void *src, bool pred)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

}  // namespace gemm_detail

// grid = (tilesM * tilesN worst case, nproblems); block = 256
__global__ void __launch_bounds__(256) grouped_dgemm_kernel(const GemmProblem *__restrict__ probs,
                                                            int tilesN_max)
{
    using namespace gemm_detail;
    const GemmProblem P = probs[blockIdx.y];
    const int tm = blockIdx.x / tilesN_max, tn = blockIdx.x % tilesN_max;
    const int m0 = tm * BM, n0 = tn * BN;
    if (m0 >= P.M || n0 >= P.Ncols || P.K <= 0) {
        return;
    }
    __shared__ double As[2][BK][AS];
    __shared__ double Bs[2][BN][BS];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int grp = lane >> 2, tig = lane & 3;
    const int wm = (wid >> 2) * 32, wn = (wid & 3) * 16;

    auto load_tiles = [&](int buf, int k0) {
        // A tile: 16 columns x 64 rows; thread -> column tid/16, rows (tid%16)*4 .. +3
        {
            const int kk = tid >> 4, mr = (tid & 15) * 4;
            const int k = k0 + kk;
            const bool kok = k < P.K;
            const int col = kok ? (P.aidx ? P.aidx[k] : k) : 0;
            const double *src = P.A + (size_t)col * P.lda + m0 + mr;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                cp_async8(&As[buf][kk][mr + q], src + q, kok && (m0 + mr + q < P.M));
        }
        // B tile: 64 columns x 16 rows; thread -> column tid/4, rows (tid%4)*4 .. +3
        {
            const int nn = tid >> 2, kr = (tid & 3) * 4;
            const int n = n0 + nn;
            const bool nok = n < P.Ncols;
            const double *src = P.B + (size_t)(nok ? n : 0) * P.ldb + k0 + kr;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                cp_async8(&Bs[buf][nn][kr + q], src + q, nok && (k0 + kr + q < P.K));
        }
        cp_async_commit();
    };

    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = (P.K + BK - 1) / BK;
    load_tiles(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            load_tiles(buf ^ 1, (kt + 1) * BK);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double af[4], bf[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[buf][4 * ks + tig][wm + 8 * i + grp];
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[j] = Bs[buf][wn + 8 * j + grp][4 * ks + tig];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm + 8 * i + grp;
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + wn + 8 * j + 2 * tig + h;
                if (n < P.Ncols) {
                    const int cc = P.cmap ? P.cmap[n] : n;
                    if (cc >= 0) P.C[(size_t)cc * P.ldc + m] = acc[i][j][h];
                }
            }
        }
    }
}
