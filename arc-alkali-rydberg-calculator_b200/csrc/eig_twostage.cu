// eig_twostage.cu -- two-stage tridiagonalisation for the batched symmetric eigensolver
// (replaces numpy.linalg.eigh / scipy eigsh of calculations_atom_single.py:986 and
// calculations_atom_pairstate.py:3444 together with eig_stedc.cu / eig_ormtr.cu).
//
// The one-stage kernel (eig_sytrd.cu) streams the trailing matrix once per column: (4/3) N^3
// bytes of HBM traffic per matrix, which bounds it.  Here the matrix is first reduced to a band
// of width 32 with level-3 operations only -- every pass over the trailing matrix now serves 32
// columns and runs on the FP64 tensor cores (DMMA) -- and the band is then reduced to
// tridiagonal form by bulge chasing on a 72 x N array.
//
//   sy2sb_kernel   one CTA per matrix.  Per block column p: Householder QR of the panel below the
//                  band (in place, global/L2), compact-WY factor T, then the two-sided update
//                      X = A22 V          (symmetric product from the lower tiles, DMMA)
//                      W = X T - 1/2 V (T^T (V^T X) T)
//                      A22 -= V W^T + W V^T   (DMMA)
//                  Reflector q = 32 p + c lives in A[q+33:, q] with an implicit 1 at row q + 32.
//   sb2st_kernel   one CTA per matrix, one warp per sweep, sweeps pipelined two steps apart
//                  (progress counters in shared memory).  Step t of sweep s applies a length-32
//                  reflector two-sidedly to a 32x32 diagonal block, from the right to the block
//                  below, and annihilates the first column of the bulge.  Lane <-> row; the
//                  transposed halves of the products go through a padded shared-memory tile.
//                  All reflectors of sweep s are stored contiguously in V2[s+1:, s].
//   q2_kernel      Z <- Q2 Z on a row-major copy of the selected eigenvectors: lane <-> eigenvector
//                  column, so a reflector is 32 uniform loads + 64 FMAs per lane with no cross-lane
//                  traffic; a 39-row register window slides over 32 consecutive sweeps.
// Index conventions and application orders are modelled in tools/proto_two_stage.py.
#include <stdlib.h>
#include "eig.cuh"

#ifdef TS_PROFILE
#define TSP_DECL long long tsp_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tsp_last = clock64();
#define TSP(k) { long long now_ = clock64(); tsp_t[k] += now_ - tsp_last; tsp_last = now_; }
#else
#define TSP_DECL
#define TSP(k)
#endif

namespace {

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

__device__ __forceinline__ void dmma_t(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// stage 1: dense -> band
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TS_THREADS, 1) sy2sb_kernel(TsArgs a)
{
    const int mat = blockIdx.x;
    const int np = a.Npad, lda = a.Npad, NB = a.NB;
    double *slot = a.base + (size_t)mat * a.slot;
    double *A = slot + a.oA;
    double *__restrict__ Wp = slot + a.oWp;     // [32][np] column-major: X, then W
    double *__restrict__ Vp = slot + a.oVp;     // [32][np] clean V (explicit unit diagonal)
    double *__restrict__ tt = slot + a.oTau;
    double *__restrict__ Tfac = slot + a.oTfac;

    extern __shared__ double sm[];
    double *uni = sm;                                               // 8448 doubles (union)
    double(*tileT)[32 * 33] = reinterpret_cast<double(*)[32 * 33]>(uni);           // [warp][j*33+l]
    double(*Bs)[2][32][BSTR] = reinterpret_cast<double(*)[2][32][BSTR]>(uni);      // [buf][op][n][k]
    double(*Ts)[33] = reinterpret_cast<double(*)[33]>(sm + 8448);
    double(*G)[33] = reinterpret_cast<double(*)[33]>(sm + 8448 + 1056);
    double(*M1)[33] = reinterpret_cast<double(*)[33]>(sm + 8448 + 2 * 1056);
    double(*red)[32] = reinterpret_cast<double(*)[32]>(sm + 8448 + 3 * 1056);      // [warp][j]
    double(*wv)[32] = reinterpret_cast<double(*)[32]>(sm + 8448 + 3 * 1056 + 256); // [warp][j]
    double *sred = sm + 8448 + 3 * 1056 + 512;                                     // [8] + alpha
    double *taus = sred + 16;                                                      // [32]

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int grp = lane >> 2, tig = lane & 3;

    TSP_DECL
    for (int p = 0; p < NB - 1; ++p) {
        TSP(0)
        const int m0 = 32 * (p + 1), col0 = 32 * p;
        const int nch = NB - (p + 1);
        double *P = A + (size_t)col0 * lda;                   // P[r + j*lda], r absolute

        // ---------------------------------------------------------------- panel QR
        {
            double ssq = 0.0;
            for (int ch = wid; ch < nch; ch += TS_WARPS) {
                const int r = m0 + 32 * ch + lane;
                if (r > m0) { const double x = P[r]; ssq += x * x; }
            }
            ssq = warp_sum(ssq);
            if (lane == 0) sred[wid] = ssq;
            if (tid == 0) sred[8] = P[m0];
        }
        for (int c = 0; c < 32; ++c) {
            __syncthreads();                                         // sync #1
            const int rd = m0 + c;
            double xn2 = 0.0;
#pragma unroll
            for (int q = 0; q < TS_WARPS; ++q) xn2 += sred[q];
            const double alpha = sred[8];
            double tau = 0.0, beta = alpha, scale = 0.0;
            if (rd < np - 1 && xn2 != 0.0) {
                beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            if (tid == 0) { tt[col0 + c] = tau; taus[c] = tau; }
            // pass B: v, w_j = v^T P[:, j] (j > c)
            double acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.0;
            for (int ch = wid; ch < nch; ch += TS_WARPS) {
                const int r = m0 + 32 * ch + lane;
                double v = 0.0;
                if (r == rd) {
                    v = 1.0;
                    P[r + (size_t)c * lda] = beta;
                } else if (r > rd) {
                    v = P[r + (size_t)c * lda] * scale;
                    P[r + (size_t)c * lda] = v;
                }
                Vp[(size_t)c * np + r] = v;
                if (r >= rd) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j > c) acc[j] += v * P[r + (size_t)j * lda];
                }
            }
            double *mt = tileT[wid];
#pragma unroll
            for (int j = 0; j < 32; ++j) mt[j * 33 + lane] = acc[j];
            __syncwarp();
            double wpart = 0.0;
#pragma unroll
            for (int l = 0; l < 32; ++l) wpart += mt[lane * 33 + l];
            red[wid][lane] = wpart;
            __syncthreads();                                         // sync #2
            double wj = 0.0;
#pragma unroll
            for (int q = 0; q < TS_WARPS; ++q) wj += red[q][lane];
            wv[wid][lane] = wj * tau;
            __syncwarp();
            // pass C: P[:, j] -= tau v w_j ; norm / pivot of the next column ride along
            double ssq = 0.0;
            for (int ch = wid; ch < nch; ch += TS_WARPS) {
                const int r = m0 + 32 * ch + lane;
                if (r >= rd) {
                    const double v = Vp[(size_t)c * np + r];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j > c) {
                            const double x = P[r + (size_t)j * lda] - v * wv[wid][j];
                            P[r + (size_t)j * lda] = x;
                            if (j == c + 1) {
                                if (r > rd + 1) ssq += x * x;
                                if (r == rd + 1) sred[8] = x;
                            }
                        }
                    }
                }
            }
            ssq = warp_sum(ssq);
            if (lane == 0) sred[wid] = ssq;
        }
        __syncthreads();
        TSP(1)

        // ---------------------------------------------------------------- T factor
        for (int t = tid; t < 1024; t += TS_THREADS) { G[t >> 5][t & 31] = 0.0; Ts[t >> 5][t & 31] = 0.0; }
        __syncthreads();
        {
            double acc[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int ch = wid; ch < nch; ch += TS_WARPS) {
                const int r0 = m0 + 32 * ch;
                double f[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) f[i][ks] = Vp[(size_t)(8 * i + grp) * np + r0 + 4 * ks + tig];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma_t(acc[i][j][0], acc[i][j][1], f[i][ks], f[j][ks]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) atomicAdd(&G[8 * i + grp][8 * j + 2 * tig + h], acc[i][j][h]);
        }
        __syncthreads();
        if (wid == 0) {
            // T(0:i, i) = -tau_i T(0:i, 0:i) G(0:i, i); T(i,i) = tau_i   (dlarft forward/columnwise)
            const int j = lane;
            for (int i = 0; i < 32; ++i) {
                const double ti = taus[i];
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
        for (int t = tid; t < 1024; t += TS_THREADS) Tfac[(size_t)p * 1024 + t] = Ts[t >> 5][t & 31];

        TSP(2)
        // ---------------------------------------------------------------- X = A22 V
        const int Jt = p + 1;
        auto stageV = [&](int buf, int q) {
            const int k = tid & 31, n0 = tid >> 5;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4)
                Bs[buf][0][n0 + 8 * i4][k] = Vp[(size_t)(n0 + 8 * i4) * np + 32 * q + k];
        };
        for (int g0 = Jt; g0 < NB; g0 += TS_WARPS) {
            const int i = g0 + wid;
            const bool act = i < NB;
            double acc[4][4][2];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
            double fa[4][8], fn[4][8];
            auto load_tile = [&](double(&f)[4][8], int q) {
                // symmetric element (32 i + rr, 32 q + cc) from the lower triangle
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const int ri = 32 * i + 8 * x + grp, ci = 32 * q + 4 * ks + tig;
                        const int hi = max(ri, ci), lo = min(ri, ci);
                        f[x][ks] = __ldcg(&A[hi + (size_t)lo * lda]);
                    }
            };
            __syncthreads();
            stageV(0, Jt);
            if (act) load_tile(fa, Jt);
            __syncthreads();
            for (int q = Jt; q < NB; ++q) {
                const int buf = (q - Jt) & 1;
                if (q + 1 < NB) {
                    stageV(buf ^ 1, q + 1);
                    if (act) load_tile(fn, q + 1);
                }
                if (act) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            const double b = Bs[buf][0][8 * y + grp][4 * ks + tig];
#pragma unroll
                            for (int x = 0; x < 4; ++x) dmma_t(acc[x][y][0], acc[x][y][1], fa[x][ks], b);
                        }
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) fa[x][ks] = fn[x][ks];
                }
                __syncthreads();
            }
            if (act) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            Wp[(size_t)(8 * y + 2 * tig + h) * np + 32 * i + 8 * x + grp] = acc[x][y][h];
            }
        }
        __syncthreads();
        TSP(3)

        // ---------------------------------------------------------------- W
        for (int t = tid; t < 1024; t += TS_THREADS) G[t >> 5][t & 31] = 0.0;
        __syncthreads();
        {
            // S = V^T X
            double acc[4][4][2];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
            for (int ch = wid; ch < nch; ch += TS_WARPS) {
                const int r0 = m0 + 32 * ch;
                double fv[4][8], fx[4][8];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        fv[x][ks] = Vp[(size_t)(8 * x + grp) * np + r0 + 4 * ks + tig];
                        fx[x][ks] = Wp[(size_t)(8 * x + grp) * np + r0 + 4 * ks + tig];
                    }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) dmma_t(acc[x][y][0], acc[x][y][1], fv[x][ks], fx[y][ks]);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int h = 0; h < 2; ++h) atomicAdd(&G[8 * x + grp][8 * y + 2 * tig + h], acc[x][y][h]);
        }
        __syncthreads();
        // M1 = S T
        for (int t = tid; t < 1024; t += TS_THREADS) {
            const int row = t >> 5, col = t & 31;
            double s = 0.0;
            for (int l = 0; l <= col; ++l) s += G[row][l] * Ts[l][col];
            M1[row][col] = s;
        }
        __syncthreads();
        // M2 = T^T M1 -> G
        for (int t = tid; t < 1024; t += TS_THREADS) {
            const int row = t >> 5, col = t & 31;
            double s = 0.0;
            for (int l = 0; l <= row; ++l) s += Ts[l][row] * M1[l][col];
            G[row][col] = s;
        }
        __syncthreads();
        // operand tiles [n][k]: T and -M2/2
        for (int t = tid; t < 1024; t += TS_THREADS) {
            const int n = t >> 5, k = t & 31;
            Bs[0][0][n][k] = Ts[k][n];
            Bs[0][1][n][k] = -0.5 * G[k][n];
        }
        __syncthreads();
        for (int ch = wid; ch < nch; ch += TS_WARPS) {
            const int r0 = m0 + 32 * ch;
            double fx[4][8], fv[4][8];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    fx[x][ks] = Wp[(size_t)(4 * ks + tig) * np + r0 + 8 * x + grp];
                    fv[x][ks] = Vp[(size_t)(4 * ks + tig) * np + r0 + 8 * x + grp];
                }
            double acc[4][4][2];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const double bt = Bs[0][0][8 * y + grp][4 * ks + tig];
                    const double bm = Bs[0][1][8 * y + grp][4 * ks + tig];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        dmma_t(acc[x][y][0], acc[x][y][1], fx[x][ks], bt);
                        dmma_t(acc[x][y][0], acc[x][y][1], fv[x][ks], bm);
                    }
                }
            __syncwarp();
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        Wp[(size_t)(8 * y + 2 * tig + h) * np + r0 + 8 * x + grp] = acc[x][y][h];
        }
        __syncthreads();
        TSP(4)

        // ---------------------------------------------------------------- A22 -= V W^T + W V^T
        auto stageWV = [&](int buf, int J) {
            const int k = tid >> 5, n = tid & 31;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                Bs[buf][0][n][k + 8 * i4] = Wp[(size_t)(k + 8 * i4) * np + 32 * J + n];
                Bs[buf][1][n][k + 8 * i4] = Vp[(size_t)(k + 8 * i4) * np + 32 * J + n];
            }
        };
        stageWV(0, Jt);
        __syncthreads();
        for (int J = Jt; J < NB; ++J) {
            const int buf = (J - Jt) & 1;
            if (J + 1 < NB) stageWV(buf ^ 1, J + 1);
            for (int I = J + ((wid - (J & 7) + 8) & 7); I < NB; I += TS_WARPS) {
                const int ar0 = 32 * I + grp;
                double *cp0 = &A[ar0 + (size_t)(32 * J + 2 * tig) * lda];
                double c0[4][4], c1[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        c0[x][y] = __ldcg(cp0 + 8 * x + (size_t)(8 * y) * lda);
                        c1[x][y] = __ldcg(cp0 + 8 * x + (size_t)(8 * y + 1) * lda);
                    }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    double fv[4], fw[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        fv[x] = -Vp[(size_t)(4 * ks + tig) * np + ar0 + 8 * x];
                        fw[x] = -Wp[(size_t)(4 * ks + tig) * np + ar0 + 8 * x];
                    }
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const double bw = Bs[buf][0][8 * y + grp][4 * ks + tig];
                        const double bv = Bs[buf][1][8 * y + grp][4 * ks + tig];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            dmma_t(c0[x][y], c1[x][y], fv[x], bw);
                            dmma_t(c0[x][y], c1[x][y], fw[x], bv);
                        }
                    }
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        cp0[8 * x + (size_t)(8 * y) * lda] = c0[x][y];
                        cp0[8 * x + (size_t)(8 * y + 1) * lda] = c1[x][y];
                    }
            }
            __syncthreads();
        }
        TSP(5)
    }  // panels
#ifdef TS_PROFILE
    if (mat == 0 && tid == 0)
        printf("sy2sb prof (10 kcycles): qr %lld tfac %lld symm %lld w %lld syr2k %lld\n", tsp_t[1] / 10000,
               tsp_t[2] / 10000, tsp_t[3] / 10000, tsp_t[4] / 10000, tsp_t[5] / 10000);
#endif

    // ------------------------------------------------------------ compact lower band (N x N part)
    __syncthreads();
    {
        double *__restrict__ Bd = slot + a.oBand;
        const int N = a.N;
        const int total = EIG_LDB * (np + 64);
        for (int idx = tid; idx < total; idx += TS_THREADS) {
            const int d = idx % EIG_LDB, j = idx / EIG_LDB;
            double x = 0.0;
            if (d <= 32 && j + d < N) x = A[(j + d) + (size_t)j * lda];
            Bd[idx] = x;
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 2: band -> tridiagonal (bulge chasing)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_house(double x, int lane, double &v, double &tau, double &beta)
{
    const double alpha = __shfl_sync(ARCB200_FULL_MASK, x, 0);
    const double xn2 = warp_sum(lane >= 1 ? x * x : 0.0);
    if (xn2 == 0.0) {
        tau = 0.0;
        beta = alpha;
        v = (lane == 0) ? 1.0 : 0.0;
    } else {
        beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
        tau = (beta - alpha) / beta;
        const double sc = 1.0 / (alpha - beta);
        v = (lane == 0) ? 1.0 : x * sc;
    }
}

__global__ void __launch_bounds__(TS_THREADS, 1) sb2st_kernel(TsArgs a)
{
    const int mat = blockIdx.x;
    const int N = a.N, ldv = a.Npad, TM = a.NB + 2;
    double *slot = a.base + (size_t)mat * a.slot;
    double *__restrict__ Bd = slot + a.oBand;
    double *__restrict__ V2 = slot + a.oV2;
    double *__restrict__ tau2 = slot + a.oTau2;
    double *__restrict__ dd = slot + a.oD;
    double *__restrict__ ee = slot + a.oE;

    extern __shared__ double sm[];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    double *tile = sm + (size_t)wid * (32 * 33);
    double *vs = sm + TS_WARPS * 32 * 33 + wid * 32;
    double *ws = sm + TS_WARPS * 32 * 33 + TS_WARPS * 32 + wid * 32;
    __shared__ volatile int prog[64];
    if (tid < 64) prog[tid] = -1;
    for (size_t i = tid; i < (size_t)a.Npad * TM; i += TS_THREADS) tau2[i] = 0.0;
    __syncthreads();

#define BD(i, j) Bd[((i) - (j)) + (size_t)(j) * EIG_LDB]
    for (int s = wid; s < N - 2; s += TS_WARPS) {
        const int prev = (s - 1) & 63;
        const int pbase = (s - 1) * 4096;
        auto wait_for = [&](int t) {
            if (s > 0) {
                const int need = pbase + min(t + 2, 4095);
                while (prog[prev] < need) { }
                __threadfence_block();
            }
        };
        int t = 0;
        wait_for(0);
        int r = s + 1;
        int L = min(32, N - r);
        double v, tau, beta;
        {
            const double x = (lane < L) ? BD(r + lane, s) : 0.0;
            warp_house(x, lane, v, tau, beta);
            if (lane < L) BD(r + lane, s) = (lane == 0) ? beta : 0.0;
        }
        while (true) {
            if (lane < L) V2[(size_t)(r + lane) + (size_t)s * ldv] = v;
            if (lane == 0) tau2[(size_t)s * TM + t] = tau;
            const int r2 = r + L;
            const int L2 = min(32, N - r2);
            double Dl[32], El[32];
            if (tau != 0.0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) Dl[j] = (j <= lane && lane < L) ? BD(r + lane, r + j) : 0.0;
            }
            if (L2 > 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) El[j] = (lane < L2) ? BD(r2 + lane, r + j) : 0.0;
            }
            vs[lane] = v;
            __syncwarp();
            if (tau != 0.0) {
                // D <- H D H on the lower triangle
                double pd = 0.0;
#pragma unroll
                for (int j = 0; j < 32; ++j) pd += Dl[j] * vs[j];
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = (j < lane) ? Dl[j] * v : 0.0;
                __syncwarp();
                double pt = 0.0;
#pragma unroll
                for (int i = 0; i < 32; ++i) pt += tile[i * 33 + lane];
                __syncwarp();
                const double pv = tau * (pd + pt);
                const double al = 0.5 * tau * warp_sum(pv * v);
                const double w = pv - al * v;
                ws[lane] = w;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j <= lane && lane < L) {
                        const double x = Dl[j] - (v * ws[j] + w * vs[j]);
                        BD(r + lane, r + j) = x;
                    }
                }
                __syncwarp();
            }
            if (L2 <= 0) break;
            // E <- E H
            if (tau != 0.0) {
                double y = 0.0;
#pragma unroll
                for (int j = 0; j < 32; ++j) y += El[j] * vs[j];
                y *= tau;
#pragma unroll
                for (int j = 0; j < 32; ++j) El[j] -= y * vs[j];
            }
            // next reflector from the first column of the bulge; E <- H' E
            double v2, tau_n, beta2;
            warp_house(El[0], lane, v2, tau_n, beta2);
            El[0] = (lane == 0) ? beta2 : 0.0;
            if (tau_n != 0.0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = v2 * El[j];
                __syncwarp();
                double sj = 0.0;
#pragma unroll
                for (int i = 0; i < 32; ++i) sj += tile[i * 33 + lane];
                __syncwarp();
                ws[lane] = (lane == 0) ? 0.0 : sj * tau_n;
                __syncwarp();
#pragma unroll
                for (int j = 1; j < 32; ++j) El[j] -= v2 * ws[j];
            }
            if (lane < L2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) BD(r2 + lane, r + j) = El[j];
            }
            __syncwarp();
            r = r2; L = L2; v = v2; tau = tau_n;
            ++t;
            __threadfence_block();
            __syncwarp();
            __threadfence_block();
            if (lane == 0) prog[s & 63] = s * 4096 + t;
            wait_for(t);
        }
        __threadfence_block();
        __syncwarp();
        __threadfence_block();
        if (lane == 0) prog[s & 63] = s * 4096 + 4095;
    }
    __syncthreads();
    for (int i = tid; i < N; i += TS_THREADS) {
        dd[i] = Bd[(size_t)i * EIG_LDB];
        ee[i] = (i < N - 1) ? Bd[1 + (size_t)i * EIG_LDB] : 0.0;
    }
#undef BD
}

// ------------------------------------------------------------------------------------------
// back-transformation with the bulge-chasing reflectors
// ------------------------------------------------------------------------------------------
// Zt[r][c] = Z[r + (sel0 + c) * lda] (zero padded to (Npad + 64) x kp); grid (row tiles, col tiles, nmat)
__global__ void __launch_bounds__(256) zt_in_kernel(TsArgs a)
{
    __shared__ double tl[32][33];
    const int mat = blockIdx.z;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *Z = slot + a.oZ + (size_t)(a.sel0 ? a.sel0[mat] : 0) * a.Npad;
    double *Zt = slot + a.oZt;
    const int r0 = 32 * blockIdx.x, c0 = 32 * blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int cc = ty; cc < 32; cc += 8) {
        const int r = r0 + tx, c = c0 + cc;
        tl[cc][tx] = (r < a.N && c < a.nsel) ? Z[r + (size_t)c * a.Npad] : 0.0;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) Zt[(size_t)(r0 + rr) * a.kp + c0 + tx] = tl[tx][rr];
}

__global__ void __launch_bounds__(256) zt_out_kernel(TsArgs a)
{
    __shared__ double tl[32][33];
    const int mat = blockIdx.z;
    double *slot = a.base + (size_t)mat * a.slot;
    double *Z = slot + a.oZ + (size_t)(a.sel0 ? a.sel0[mat] : 0) * a.Npad;
    const double *Zt = slot + a.oZt;
    const int r0 = 32 * blockIdx.x, c0 = 32 * blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int rr = ty; rr < 32; rr += 8) tl[tx][rr] = Zt[(size_t)(r0 + rr) * a.kp + c0 + tx];
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
        const int r = r0 + tx, c = c0 + cc;
        if (r < a.N && c < a.nsel) Z[r + (size_t)c * a.Npad] = tl[cc][tx];
    }
}

// grid (ceil(nslabs / 8), nmat); warp <-> slab of 32 eigenvector columns, lane <-> column
__global__ void __launch_bounds__(TS_THREADS) q2_kernel(TsArgs a)
{
    const int mat = blockIdx.y;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slab = blockIdx.x * TS_WARPS + wid;
    if (slab * 32 >= a.kp) return;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *__restrict__ V2 = slot + a.oV2;
    const double *__restrict__ tau2 = slot + a.oTau2;
    double *Zc = slot + a.oZt + slab * 32 + lane;
    const int N = a.N, ldv = a.Npad, TM = a.NB + 2;
    const size_t kp = (size_t)a.kp;
    const int nsw = N - 2;
    const int ngroups = (nsw + 31) / 32;
    for (int G = ngroups - 1; G >= 0; --G) {
        const int s0 = 32 * G;
        for (int t = 0; s0 + 1 + 32 * t < N; ++t) {
            double zr[39];
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
                const int s_hi = s0 + 31 - 8 * sub;
                const int Rb = s_hi - 6 + 32 * t;
                if (sub == 0) {
#pragma unroll
                    for (int x = 0; x < 39; ++x) zr[x] = Zc[(size_t)(Rb + x) * kp];
                } else {
#pragma unroll
                    for (int x = 30; x >= 0; --x) zr[x + 8] = zr[x];
#pragma unroll
                    for (int x = 0; x < 8; ++x) zr[x] = Zc[(size_t)(Rb + x) * kp];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int s = s_hi - u;
                    if (s < nsw) {
                        const double tau = tau2[(size_t)s * TM + t];
                        if (tau != 0.0) {
                            const int r = s + 1 + 32 * t;
                            const double *vp = V2 + (size_t)s * ldv + r;
                            double vl[32];
#pragma unroll
                            for (int l = 0; l < 32; ++l) vl[l] = (r + l < N) ? __ldg(vp + l) : 0.0;
                            double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                            for (int l = 0; l < 32; l += 4) {
                                d0 += vl[l] * zr[7 - u + l];
                                d1 += vl[l + 1] * zr[7 - u + l + 1];
                                d2 += vl[l + 2] * zr[7 - u + l + 2];
                                d3 += vl[l + 3] * zr[7 - u + l + 3];
                            }
                            const double dot = tau * ((d0 + d1) + (d2 + d3));
#pragma unroll
                            for (int l = 0; l < 32; ++l) zr[7 - u + l] -= dot * vl[l];
                        }
                    }
                }
                if (sub < 3) {
#pragma unroll
                    for (int x = 31; x < 39; ++x) Zc[(size_t)(Rb + x) * kp] = zr[x];
                } else {
#pragma unroll
                    for (int x = 0; x < 39; ++x) Zc[(size_t)(Rb + x) * kp] = zr[x];
                }
            }
        }
    }
}

TsArgs make_args(const EigSlots &S)
{
    const EigLayout &L = S.L;
    TsArgs a;
    a.N = L.N; a.Npad = L.Npad; a.NB = L.NB;
    a.base = S.dbl; a.slot = L.slot_doubles;
    a.oA = L.oA; a.oWp = L.oWp; a.oVp = L.oVp; a.oTau = L.oTau; a.oTfac = L.oTfac;
    a.oBand = L.oBand; a.oV2 = L.oV2; a.oTau2 = L.oTau2; a.oD = L.oD; a.oE = L.oE;
    a.oZ = L.oQ1; a.oZt = L.oU;
    a.nsel = 0; a.kp = 0; a.sel0 = nullptr;
    return a;
}

}  // namespace

int eig_use_two_stage(int N)
{
    const char *ev = getenv("ARCB200_TWO_STAGE");
    if (ev) return ev[0] == '1' && N >= 96;
    return N >= 1024 && N <= 4096;
}

int eig_sy2sb(const EigSlots &S, int nmat, cudaStream_t st, int *launches)
{
    if (!S.L.two_stage) { arcb200_set_error("sy2sb: layout without two-stage buffers"); return -1; }
    const TsArgs a = make_args(S);
    const size_t smem = (8448 + 3 * 1056 + 512 + 64) * sizeof(double);
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sy2sb_kernel<<<nmat, TS_THREADS, smem, st>>>(a);
    ARCB200_LAUNCH_CHECK("sy2sb_kernel");
    if (launches) *launches += 1;
    return 0;
}

int eig_sb2st(const EigSlots &S, int nmat, cudaStream_t st, int *launches)
{
    if (!S.L.two_stage) { arcb200_set_error("sb2st: layout without two-stage buffers"); return -1; }
    const TsArgs a = make_args(S);
    const size_t smem = (size_t)(TS_WARPS * 32 * 33 + 2 * TS_WARPS * 32) * sizeof(double);
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sb2st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sb2st_kernel<<<nmat, TS_THREADS, smem, st>>>(a);
    ARCB200_LAUNCH_CHECK("sb2st_kernel");
    if (launches) *launches += 1;
    return 0;
}

int eig_q2_apply(const EigSlots &S, int nmat, const int32_t *d_sel0, int nsel, cudaStream_t st,
                 int *launches)
{
    if (S.L.N < 3 || nsel <= 0) return 0;
    TsArgs a = make_args(S);
    a.nsel = nsel;
    a.kp = (nsel + 31) & ~31;
    a.sel0 = d_sel0;
    const int rowtiles = (S.L.Npad + 64) / 32, coltiles = a.kp / 32;
    zt_in_kernel<<<dim3(rowtiles, coltiles, nmat), 256, 0, st>>>(a);
    const int nslab = a.kp / 32;
    q2_kernel<<<dim3((nslab + TS_WARPS - 1) / TS_WARPS, nmat), TS_THREADS, 0, st>>>(a);
    zt_out_kernel<<<dim3(S.L.NB, coltiles, nmat), 256, 0, st>>>(a);
    ARCB200_LAUNCH_CHECK("q2 kernels");
    if (launches) *launches += 3;
    return 0;
}
