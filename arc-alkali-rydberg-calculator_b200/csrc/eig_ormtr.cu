// eig_ormtr.cu -- stage 3 of the batched symmetric eigensolver: back-transformation
//   Z = Q_house * Q_T[:, sel]     (LAPACK dormtr, 'L','N'; compact-WY blocks of 32)
// Householder vector c is stored in A(c+2:N, c) with an implicit 1 at row c+1; tau[c].
//
// larft_kernel : T factor of every 32-reflector block (one CTA per block and matrix).
// ormtr_kernel : one CTA per (matrix, slab of 32 eigenvector columns); for blocks
//                b = last..0:   W = V_b^T Z_slab ; W = T_b W ; Z_slab -= V_b W
//                Both products run on FP64 tensor cores (DMMA m8n8k4) from shared-memory
//                tiles; the slab is the only data written (12 N^2 bytes per slab in total).
#include "eig.cuh"

namespace {

constexpr int OT_THREADS = 256;
constexpr int OT_WARPS = 8;
constexpr int TS = 36;          // smem row stride (doubles): conflict-free DMMA fragments

struct OrmArgs {
    int N, Npad, NBR;           // NBR = number of reflector blocks = ceil((N-1)/32)
    double *base; size_t slot;
    size_t oA, oTau, oTfac, oZ;
    int nsel;
    const int32_t *sel0;        // per matrix first selected column (or null -> 0)
    int off;                    // reflector c has its implicit 1 at row c + off (1: sytrd, 32: sy2sb)
};

__device__ __forceinline__ double vval(const double *__restrict__ A, int lda, int N, int r, int c,
                                       int off)
{
    // element r of Householder vector c
    if (c + off > N - 1 || r < c + off || r >= N) return 0.0;
    if (r == c + off) return 1.0;
    return A[r + (size_t)c * lda];
}

__device__ __forceinline__ void dmma_o(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// grid = (NBR, nmat)
__global__ void __launch_bounds__(OT_THREADS) larft_kernel(OrmArgs a)
{
    const int b = blockIdx.x, mat = blockIdx.y;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *A = slot + a.oA;
    const double *tau = slot + a.oTau;
    double *T = slot + a.oTfac + (size_t)b * 1024;
    const int N = a.N, lda = a.Npad, c0 = 32 * b;
    __shared__ double Vs[32][65];        // [col][row chunk]
    __shared__ double G[32][33];
    __shared__ double Ts[32][33];
    const int tid = threadIdx.x;
    // Gram matrix G(j,i) = v_j . v_i, thread handles 4 (j,i) pairs
    double acc[4] = {0, 0, 0, 0};
    for (int r0 = c0; r0 < N; r0 += 64) {
        for (int t = tid; t < 32 * 64; t += OT_THREADS) {
            const int col = t >> 6, rr = t & 63;
            Vs[col][rr] = vval(A, lda, N, r0 + rr, c0 + col, a.off);
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int pair = tid + p * OT_THREADS;
            const int j = pair >> 5, i = pair & 31;
            if (j < i) {
                double s = 0.0;
                for (int rr = 0; rr < 64; ++rr) s += Vs[j][rr] * Vs[i][rr];
                acc[p] += s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int pair = tid + p * OT_THREADS;
        G[pair >> 5][pair & 31] = acc[p];
    }
    for (int t = tid; t < 1024; t += OT_THREADS) Ts[t >> 5][t & 31] = 0.0;
    __syncthreads();
    // T(0:i, i) = -tau_i T(0:i, 0:i) G(0:i, i); T(i,i) = tau_i   (dlarft, forward/columnwise)
    if (tid < 32) {
        const int j = tid;
        for (int i = 0; i < 32; ++i) {
            const int c = c0 + i;
            const double ti = (c + a.off <= N - 1) ? tau[c] : 0.0;
            double s = 0.0;
            if (j < i) {
                for (int l = j; l < i; ++l) s += Ts[j][l] * G[l][i];
                s *= -ti;
            } else if (j == i) {
                s = ti;
            }
            __syncwarp();
            if (j <= i) Ts[j][i] = s;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int t = tid; t < 1024; t += OT_THREADS) T[t] = Ts[t >> 5][t & 31];
}

// grid = (nslabs, nmat)
__global__ void __launch_bounds__(OT_THREADS) ormtr_kernel(OrmArgs a)
{
    const int slab = blockIdx.x, mat = blockIdx.y;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *A = slot + a.oA;
    const double *Tall = slot + a.oTfac;
    const int N = a.N, lda = a.Npad;
    const int col0 = (a.sel0 ? a.sel0[mat] : 0) + 32 * slab;
    const int ncol = min(32, a.nsel - 32 * slab);
    double *Z = slot + a.oZ + (size_t)col0 * lda;
    extern __shared__ double orm_smem[];
    double(*Vs)[32][TS] = reinterpret_cast<double(*)[32][TS]>(orm_smem);                         // per warp: [col j][row]
    double(*Zs)[32][TS] = reinterpret_cast<double(*)[32][TS]>(orm_smem + OT_WARPS * 32 * TS);     // per warp: [col n][row]
    double(*Wred)[TS] = reinterpret_cast<double(*)[TS]>(orm_smem + 2 * OT_WARPS * 32 * TS);       // W [n][j]
    double(*Tsm)[33] = reinterpret_cast<double(*)[33]>(orm_smem + 2 * OT_WARPS * 32 * TS + 32 * TS);
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int grp = lane >> 2, tig = lane & 3;

    for (int b = a.NBR - 1; b >= 0; --b) {
        const int c0 = 32 * b;
        const int rbeg = c0 + a.off;             // first row with a non-zero V entry
        const int chunk0 = rbeg & ~31;
        for (int t = tid; t < 1024; t += OT_THREADS) Tsm[t >> 5][t & 31] = Tall[(size_t)b * 1024 + t];
        // ---------------- pass 1: W = V^T Z  (32 x 32), K = rows ----------------
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int r0 = chunk0 + 32 * wid; r0 < N; r0 += 32 * OT_WARPS) {
            const int r = r0 + lane;
#pragma unroll 8
            for (int cc = 0; cc < 32; ++cc) {
                Vs[wid][cc][lane] = vval(A, lda, N, r, c0 + cc, a.off);
                Zs[wid][cc][lane] = (r < N && cc < ncol) ? Z[r + (size_t)cc * lda] : 0.0;
            }
            __syncwarp();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                double af[4], bf[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = Vs[wid][8 * i + grp][4 * ks + tig];   // A[m=j][k=r]
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = Zs[wid][8 * j + grp][4 * ks + tig];   // B[k=r][n]
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_o(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
        }
        // reduce the 8 warps' partial W (m = j, n = col) through shared memory
        __syncthreads();
        for (int t = tid; t < 32 * TS; t += OT_THREADS) (&Wred[0][0])[t] = 0.0;
        __syncthreads();
        for (int w = 0; w < OT_WARPS; ++w) {
            if (wid == w) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            Wred[8 * j + 2 * tig + h][8 * i + grp] += acc[i][j][h];   // [n][jrow]
            }
            __syncthreads();
        }
        // W <- T W   (T upper triangular 32x32): thread -> (row jrow = t&31, col n = t>>5, 4 each)
        double wn[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int t = tid + p * OT_THREADS;
            const int n = t >> 5, jr = t & 31;
            double s = 0.0;
            for (int l = jr; l < 32; ++l) s += Tsm[jr][l] * Wred[n][l];
            wn[p] = s;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int t = tid + p * OT_THREADS;
            Wred[t >> 5][t & 31] = wn[p];
        }
        __syncthreads();
        // ---------------- pass 2: Z -= V W  (rows x 32), K = 32 ------------------
        for (int r0 = chunk0 + 32 * wid; r0 < N; r0 += 32 * OT_WARPS) {
            const int r = r0 + lane;
#pragma unroll 8
            for (int cc = 0; cc < 32; ++cc) {
                Vs[wid][cc][lane] = -vval(A, lda, N, r, c0 + cc, a.off);
                Zs[wid][cc][lane] = (r < N && cc < ncol) ? Z[r + (size_t)cc * lda] : 0.0;
            }
            __syncwarp();
            double c2[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) c2[i][j][h] = Zs[wid][8 * j + 2 * tig + h][8 * i + grp];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                double af[4], bf[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = Vs[wid][4 * ks + tig][8 * i + grp];   // A[m=r][k=j]
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = Wred[8 * j + grp][4 * ks + tig];      // B[k=j][n]
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_o(c2[i][j][0], c2[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) Zs[wid][8 * j + 2 * tig + h][8 * i + grp] = c2[i][j][h];
            __syncwarp();
            if (r < N) {
#pragma unroll 8
                for (int cc = 0; cc < 32; ++cc)
                    if (cc < ncol) Z[r + (size_t)cc * lda] = Zs[wid][cc][lane];
            }
            __syncwarp();
        }
        __syncthreads();
    }
}


// Shared-memory resident variant: the whole slab Z[:, NT*8 columns] stays in shared memory
// across all reflector blocks (read once, written once); only V streams from L2.
// LDZ = Npad + 4 keeps the DMMA fragment accesses bank-conflict free.
template <int NT>
__global__ void __launch_bounds__(OT_THREADS) ormtr_smem_kernel(OrmArgs a)
{
    constexpr int SW = NT * 8;
    const int slab = blockIdx.x, mat = blockIdx.y;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *__restrict__ A = slot + a.oA;
    const double *Tall = slot + a.oTfac;
    const int N = a.N, lda = a.Npad, LDZ = a.Npad + 4;
    const int col0 = (a.sel0 ? a.sel0[mat] : 0) + SW * slab;
    const int ncol = min(SW, a.nsel - SW * slab);
    double *Z = slot + a.oZ + (size_t)col0 * lda;
    extern __shared__ double orm_smem[];
    double *Zs = orm_smem;                                     // [SW][LDZ]
    double(*Wred)[TS] = reinterpret_cast<double(*)[TS]>(orm_smem + (size_t)SW * LDZ);   // [SW][36]
    double(*Tsm)[33] = reinterpret_cast<double(*)[33]>(orm_smem + (size_t)SW * LDZ + SW * TS);
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int grp = lane >> 2, tig = lane & 3;

    for (int cc = wid; cc < SW; cc += OT_WARPS)
        for (int r = lane; r < lda; r += 32)
            Zs[(size_t)cc * LDZ + r] = (r < N && cc < ncol) ? Z[r + (size_t)cc * lda] : 0.0;
    __syncthreads();

    for (int b = a.NBR - 1; b >= 0; --b) {
        const int c0 = 32 * b;
        const int chunk0 = (c0 + a.off) & ~31;
        for (int t = tid; t < 1024; t += OT_THREADS) Tsm[t >> 5][t & 31] = Tall[(size_t)b * 1024 + t];
        for (int t = tid; t < SW * TS; t += OT_THREADS) (&Wred[0][0])[t] = 0.0;
        // ---------------- pass 1: W = V^T Z ----------------
        double acc[4][NT][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int r0 = chunk0 + 32 * wid; r0 < N; r0 += 32 * OT_WARPS) {
            double af[8][4];
            const bool top = r0 < c0 + 64;      // rows that touch the unit / zero part of V
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r0 + 4 * ks + tig, c = c0 + 8 * i + grp;
                    af[ks][i] = (top || r >= N || c + a.off > N - 1) ? vval(A, lda, N, r, c, a.off) : A[r + (size_t)c * lda];
                }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                double bf[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = Zs[(size_t)(8 * j + grp) * LDZ + r0 + 4 * ks + tig];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma_o(acc[i][j][0], acc[i][j][1], af[ks][i], bf[j]);
            }
        }
        __syncthreads();
        for (int w = 0; w < OT_WARPS; ++w) {
            if (wid == w) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) Wred[8 * j + 2 * tig + h][8 * i + grp] += acc[i][j][h];
            }
            __syncthreads();
        }
        // W <- T W
        double wn[(SW * 32 + OT_THREADS - 1) / OT_THREADS];
#pragma unroll
        for (int p = 0; p < (SW * 32 + OT_THREADS - 1) / OT_THREADS; ++p) {
            const int t = tid + p * OT_THREADS;
            double sacc = 0.0;
            if (t < SW * 32) {
                const int n = t >> 5, jr = t & 31;
                for (int l = jr; l < 32; ++l) sacc += Tsm[jr][l] * Wred[n][l];
            }
            wn[p] = sacc;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < (SW * 32 + OT_THREADS - 1) / OT_THREADS; ++p) {
            const int t = tid + p * OT_THREADS;
            if (t < SW * 32) Wred[t >> 5][t & 31] = wn[p];
        }
        __syncthreads();
        // ---------------- pass 2: Z -= V W ----------------
        for (int r0 = chunk0 + 32 * wid; r0 < N; r0 += 32 * OT_WARPS) {
            const bool top = r0 < c0 + 64;
            double af[8][4];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r0 + 8 * i + grp, c = c0 + 4 * ks + tig;
                    af[ks][i] = -((top || r >= N || c + a.off > N - 1) ? vval(A, lda, N, r, c, a.off) : A[r + (size_t)c * lda]);
                }
            double c2[4][NT][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        c2[i][j][h] = Zs[(size_t)(8 * j + 2 * tig + h) * LDZ + r0 + 8 * i + grp];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                double bf[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = Wred[8 * j + grp][4 * ks + tig];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma_o(c2[i][j][0], c2[i][j][1], af[ks][i], bf[j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        Zs[(size_t)(8 * j + 2 * tig + h) * LDZ + r0 + 8 * i + grp] = c2[i][j][h];
        }
        __syncthreads();
    }
    for (int cc = wid; cc < ncol; cc += OT_WARPS)
        for (int r = lane; r < N; r += 32) Z[r + (size_t)cc * lda] = Zs[(size_t)cc * LDZ + r];
}

template <int NT>
int launch_ormtr_smem(const OrmArgs &a, int nmat, cudaStream_t st)
{
    const int SW = NT * 8;
    const size_t smem = ((size_t)SW * (a.Npad + 4) + (size_t)SW * TS + 32 * 33) * sizeof(double);
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(ormtr_smem_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nslab = (a.nsel + SW - 1) / SW;
    ormtr_smem_kernel<NT><<<dim3(nslab, nmat), OT_THREADS, smem, st>>>(a);
    return 0;
}

}  // namespace

int eig_ormtr(const EigSlots &S, int nmat, const int32_t *d_sel0, int nsel, cudaStream_t st,
              int *launches)
{
    const EigLayout &L = S.L;
    if (L.N < 3 || nsel <= 0) return 0;
    OrmArgs a;
    a.N = L.N; a.Npad = L.Npad;
    a.off = L.two_stage ? 32 : 1;
    a.NBR = L.two_stage ? L.NB - 1 : (L.N - 1 + 31) / 32;
    a.base = S.dbl; a.slot = L.slot_doubles;
    a.oA = L.oA; a.oTau = L.oTau; a.oTfac = L.oTfac; a.oZ = L.oQ1;
    a.nsel = nsel; a.sel0 = d_sel0;
    if (a.NBR <= 0) return 0;
    // the band reduction (eig_twostage.cu) leaves its T factors in oTfac already
    if (!L.two_stage) larft_kernel<<<dim3(a.NBR, nmat), OT_THREADS, 0, st>>>(a);
    // shared-memory resident slab when it fits (<= 200 KiB): widest slab first
    const size_t budget = 200 * 1024;
    auto fits = [&](int sw) {
        return ((size_t)sw * (L.Npad + 4) + (size_t)sw * TS + 32 * 33) * sizeof(double) <= budget;
    };
    int rcs = -100;
    if (fits(32)) rcs = launch_ormtr_smem<4>(a, nmat, st);
    else if (fits(16)) rcs = launch_ormtr_smem<2>(a, nmat, st);
    else if (fits(8)) rcs = launch_ormtr_smem<1>(a, nmat, st);
    if (rcs != -100) {
        if (rcs) return rcs;
        ARCB200_LAUNCH_CHECK("ormtr_smem kernel");
        if (launches) *launches += 2;
        return 0;
    }
    const int nslab = (nsel + 31) / 32;
    const size_t smem = (size_t)(2 * OT_WARPS * 32 * TS + 32 * TS + 32 * 33) * sizeof(double);
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(ormtr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ormtr_kernel<<<dim3(nslab, nmat), OT_THREADS, smem, st>>>(a);
    ARCB200_LAUNCH_CHECK("ormtr kernels");
    if (launches) *launches += 2;
    return 0;
}
