// radial_gram.cu -- radial dipole / quadrupole tables as FP64 tensor-core GEMMs.
//
// For two groups of radial states (rows a, columns b) that share the x = sqrt(r) grid
//     out[a][b] = trapezoid(psi_a psi_b r^p, r)[0 : min(La, Lb)]
// (alkali_atom_functions.py:1087-1096, 1191-1201) is a Gram matrix
//     Psi_A diag(w_i r_i^p) Psi_B^T ,  w_i = trapezoid weights of the non-uniform r grid,
// because every wavefunction is zero beyond its own grid length.  The only difference to the
// reference is the trapezoid END weight at i = min(La,Lb)-1, which depends on the pair and is
// restored exactly by gram_fix_kernel.  Each psi is read once per block pair (the one-warp-per-
// pair kernel streams 24 B per grid point per pair); the products run on DMMA
// (mma.sync.m8n8k4.f64), K = grid index, split over CTAs and reduced with FP64 atomics.
#include "common.cuh"

namespace {

constexpr int GB = 64, GK = 16, GS = 20;   // tile, K slab, smem row stride (conflict-free)
constexpr int KCHUNK = 4096;               // grid points per CTA

__device__ __forceinline__ void cp8(void *dst, const void *src, bool pred)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void dmma_g(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// w[0][i] = wt_i * r_i, w[1][i] = wt_i * r_i^2 with wt the trapezoid weights of grid r[0..L)
__global__ void gram_weights_kernel(int L, const double *__restrict__ r, double *__restrict__ w)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        const double rl = (i > 0) ? r[i - 1] : r[0];
        const double rh = (i < L - 1) ? r[i + 1] : r[L - 1];
        const double wt = 0.5 * (rh - rl);
        w[i] = wt * r[i];
        w[(size_t)L + i] = wt * r[i] * r[i];
    }
}

// grid = (tiles_m * tiles_n * kchunks (worst case), nblocks)
__global__ void __launch_bounds__(256) gram_kernel(
    const arcb200_gram_block *__restrict__ blocks, const int32_t *__restrict__ rows,
    const int32_t *__restrict__ cols, const arcb200_state_t *__restrict__ states,
    const int64_t *__restrict__ offsets, const double *__restrict__ psi,
    const double *__restrict__ wgt, int Lw, double *__restrict__ out, int tiles_max, int kchunks_max)
{
    const arcb200_gram_block B = blocks[blockIdx.y];
    const int kc = blockIdx.x % kchunks_max;
    const int tile = blockIdx.x / kchunks_max;
    const int tm = tile / tiles_max, tn = tile % tiles_max;
    const int m0 = tm * GB, n0 = tn * GB, kbeg = kc * KCHUNK;
    if (m0 >= B.nrows || n0 >= B.ncols || kbeg >= B.kmax) return;
    const int kend = min(kbeg + KCHUNK, B.kmax);
    __shared__ double As[2][GB][GS];
    __shared__ double Bs[2][GB][GS];
    __shared__ double Ws[2][GK];
    __shared__ const double *pa[GB];
    __shared__ const double *pb[GB];
    __shared__ int la[GB], lb[GB];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int grp = lane >> 2, tig = lane & 3;
    const int wm = (wid >> 2) * 32, wn = (wid & 3) * 16;
    if (tid < GB) {
        const bool ok = m0 + tid < B.nrows;
        const int st = ok ? rows[B.row0 + m0 + tid] : 0;
        pa[tid] = psi + offsets[st];
        la[tid] = ok ? states[st].length : 0;
    } else if (tid < 2 * GB) {
        const int t = tid - GB;
        const bool ok = n0 + t < B.ncols;
        const int st = ok ? cols[B.col0 + n0 + t] : 0;
        pb[t] = psi + offsets[st];
        lb[t] = ok ? states[st].length : 0;
    }
    __syncthreads();
    const double *wv = wgt + (size_t)(B.power == 2 ? Lw : 0);

    auto load = [&](int buf, int k0) {
        const int rr = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + kk + q;
            cp8(&As[buf][rr][kk + q], pa[rr] + k, k < la[rr] && k < kend);
            cp8(&Bs[buf][rr][kk + q], pb[rr] + k, k < lb[rr] && k < kend);
        }
        if (tid < GK) cp8(&Ws[buf][tid], wv + k0 + tid, k0 + tid < kend && k0 + tid < Lw);
        asm volatile("cp.async.commit_group;\n" ::);
    };

    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nk = (kend - kbeg + GK - 1) / GK;
    load(0, kbeg);
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            load(buf ^ 1, kbeg + (kt + 1) * GK);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GK / 4; ++ks) {
            const double w = Ws[buf][4 * ks + tig];
            double af[4], bf[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[buf][wm + 8 * i + grp][4 * ks + tig];
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[j] = Bs[buf][wn + 8 * j + grp][4 * ks + tig] * w;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma_g(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    double *o = out + B.out_off;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm + 8 * i + grp;
        if (m >= B.nrows) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + wn + 8 * j + 2 * tig + h;
                if (n < B.ncols) atomicAdd(&o[(size_t)m * B.ncols + n], acc[i][j][h]);
            }
    }
}

// restore the exact trapezoid end weight at i = up-1, up = min(La, Lb)
__global__ void gram_fix_kernel(const arcb200_gram_block *__restrict__ blocks,
                                const int32_t *__restrict__ rows, const int32_t *__restrict__ cols,
                                const arcb200_state_t *__restrict__ states,
                                const int64_t *__restrict__ offsets, const double *__restrict__ psi,
                                const double *__restrict__ r, int Lw, double *__restrict__ out)
{
    const arcb200_gram_block B = blocks[blockIdx.y];
    const int total = B.nrows * B.ncols;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int m = e / B.ncols, n = e % B.ncols;
        const int sa = rows[B.row0 + m], sb = cols[B.col0 + n];
        const int up = min(states[sa].length, states[sb].length);
        if (up < 2 || up >= Lw) continue;     // up == Lw: the table's own end weight is exact
        const double ri = r[up - 1];
        const double f = psi[offsets[sa] + up - 1] * psi[offsets[sb] + up - 1]
                         * (B.power == 2 ? ri * ri : ri);
        out[B.out_off + e] -= 0.5 * (r[up] - ri) * f;
    }
}

}  // namespace

extern "C" int arcb200_radial_gram(int device, void *stream, int32_t nblocks,
                                   const void *d_blocks, const int32_t *d_rows,
                                   const int32_t *d_cols, const arcb200_state_t *d_states,
                                   const int64_t *d_offsets, const double *d_psi,
                                   const double *d_rgrid, int32_t Lgrid, double *d_wgt,
                                   double *d_out, int32_t max_rows, int32_t max_cols,
                                   int32_t max_k)
{
    if (nblocks <= 0) return 0;
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    gram_weights_kernel<<<148, 256, 0, st>>>(Lgrid, d_rgrid, d_wgt);
    const int tiles = (max(max_rows, max_cols) + GB - 1) / GB;
    const int kch = (min(max_k, Lgrid) + KCHUNK - 1) / KCHUNK;
    const arcb200_gram_block *blk = (const arcb200_gram_block *)d_blocks;
    gram_kernel<<<dim3(tiles * tiles * kch, nblocks), 256, 0, st>>>(
        blk, d_rows, d_cols, d_states, d_offsets, d_psi, d_wgt, Lgrid, d_out, tiles, kch);
    gram_fix_kernel<<<dim3((max_rows * max_cols + 255) / 256, nblocks), 256, 0, st>>>(
        blk, d_rows, d_cols, d_states, d_offsets, d_psi, d_rgrid, Lgrid, d_out);
    ARCB200_LAUNCH_CHECK("radial gram kernels");
    return 0;   // 3 kernels launched
}
