// eig_twostage.cuh -- shared definitions of the two-stage tridiagonalisation kernels
// (eig_twostage.cu: one CTA per matrix; eig_sy2sb_multi.cu: per-panel kernels, many CTAs per matrix).
#pragma once
#include <stdlib.h>
#include "eig.cuh"

#ifdef TS_PROFILE
#define TSP_DECL long long tsp_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tsp_last = clock64();
#define TSP(k) { long long now_ = clock64(); tsp_t[k] += now_ - tsp_last; tsp_last = now_; }
#else
#define TSP_DECL
#define TSP(k)
#endif

namespace ts {

constexpr int TS_THREADS = 256;
constexpr int TS_WARPS = 8;
constexpr int BSTR = 36;          // [n][k] operand tiles: conflict-free DMMA fragment reads

struct TsArgs {
    int N, Npad, NB;
    double *base;
    size_t slot;
    size_t oA, oWp, oVp, oTau, oTfac, oBand, oV2, oTau2, oD, oE, oZ, oZt;
    int nsel, kp;
    const int32_t *sel0;
};

static __device__ __forceinline__ void dmma_t(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

static __device__ __forceinline__ void cp_async8_t(void *dst, const void *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src));
}
static __device__ __forceinline__ void cp_async16_t(void *dst, const void *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
static __device__ __forceinline__ void cp_async_wait_all_t()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

}  // namespace ts

// bulge chasing with every sweep split between an E-chain warp and a D-block warp (eig_sb2st_split.cu)
int eig_sb2st_split(const ts::TsArgs &a, int nmat, cudaStream_t st, int *launches);
// band reduction with per-panel kernels and many CTAs per matrix (large N, small batches)
int eig_sy2sb_multi(const ts::TsArgs &a, int nmat, cudaStream_t st, int *launches);
