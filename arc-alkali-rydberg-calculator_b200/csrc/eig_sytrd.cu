// eig_sytrd.cu -- stage 1 of the batched symmetric eigensolver: blocked Householder
// tridiagonalisation, one thread-block CLUSTER per matrix (sm_100a).
//
// Algorithm: LAPACK dsytrd/dlatrd, UPLO='L', panel width 32, written for a cluster of
// C CTAs x 8 warps.  Block-row I (32 rows) of the lower triangle is owned by cluster-warp
// (I mod 8C); the owner is the only writer of those rows of A, Vp, Wp.
//   per column c of a panel (3 cluster barriers):
//     phase 1  lazy update of column c with the panel's (V, W); norm partials
//     phase 2  v = Householder vector; symmetric mat-vec  w = A v  on the lower
//              triangle: each 32x32 tile is read ONCE (coalesced, lane <-> row, next tile
//              prefetched into registers) and used for both A_IJ v_J (registers) and
//              A_IJ^T v_I (padded shared-memory transpose, lane <-> column, reduced across
//              block rows with FP64 reductions in L2) -- algorithmic HBM bytes 4 N^3 / 3
//              per matrix; panel dot products W^T v, V^T v ride along
//     phase 3  w -= V (W^T v) + W (V^T v); w *= tau; partial w.v
//     phase 4  w -= (tau/2)(w.v) v  -> column of W
//   per panel: trailing update A -= V W^T + W V^T with FP64 tensor-core DMMA tiles
//   (mma.sync.m8n8k4.f64); the panel rows of block column J are staged in shared memory
//   once per CTA (conflict-free stride 36), A fragments come from the warp's own rows.
// Vectors are exchanged through L2 (ld.global.cg) between cluster barriers.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "eig.cuh"

namespace cg = cooperative_groups;

#ifdef SYTRD_PROFILE
#define PROF_DECL long long prof_t[12] = {0,0,0,0,0,0,0,0,0,0,0,0}; long long prof_last = clock64();
#define PROF(k) { long long now_ = clock64(); prof_t[k] += now_ - prof_last; prof_last = now_; }
#else
#define PROF_DECL
#define PROF(k)
#endif

namespace {

constexpr int SY_WARPS = 8;
constexpr int SY_THREADS = SY_WARPS * 32;
constexpr int TPAD = 33;

struct SytrdArgs {
    int N, Npad, NB;
    double *base;            // slot 0
    size_t slot;             // doubles per slot
    size_t oA, oWp, oVp, oTrans, oWdir, oD, oE, oTau, oScr;
};

__device__ __forceinline__ void dmma8x8x4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// sum_j V(r,j)*a[j] + W(r,j)*b[j] over the 32 panel columns (unused columns are zero).
// Panels are column-major [32][Npad]: lane <-> row, so each of the 64 loads is one
// coalesced 256-byte warp request and all of them are in flight together.
__device__ __forceinline__ double panel_row_dot(const double *__restrict__ vcol,
                                                const double *__restrict__ wcol, int ld,
                                                const double *a, const double *b, int ncols)
{
    // only the first `ncols` panel columns are populated (warp-uniform): 8 at a time
    double s0 = 0.0, s1 = 0.0;
    for (int j0 = 0; j0 < ncols; j0 += 8) {
        double x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { x[j] = vcol[(size_t)(j0 + j) * ld]; y[j] = wcol[(size_t)(j0 + j) * ld]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s0 += x[j] * a[j0 + j]; s1 += y[j] * b[j0 + j]; }
    }
    return s0 + s1;
}

template <bool CL>
__device__ __forceinline__ void group_sync(cg::cluster_group &cl)
{
    if (CL) cl.sync(); else __syncthreads();
}

// LEAN: no register double-buffering of the symv tiles -> <= 128 registers, two CTAs per SM.
// Small matrices (few tiles per warp and column) are bound by the per-column latency chain,
// so a second resident matrix per SM hides more than tile prefetching does.
template <bool CL, bool LEAN>
__global__ void __launch_bounds__(SY_THREADS, LEAN ? 2 : 1) sytrd_kernel(SytrdArgs a)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = CL ? (int)cluster.num_blocks() : 1;
    const int rank = CL ? (int)cluster.block_rank() : 0;
    const int mat = blockIdx.x / C;
    const int N = a.N, lda = a.Npad, NB = a.NB;
    double *slot = a.base + (size_t)mat * a.slot;
    double *__restrict__ A = slot + a.oA;
    double *__restrict__ Wp = slot + a.oWp;       // [32][Npad] column-major panels
    double *__restrict__ Vp = slot + a.oVp;
    double *__restrict__ wtr = slot + a.oTrans;   // [Npad] L2-reduced transposed part of the mat-vec
    double *__restrict__ dd = slot + a.oD;
    double *__restrict__ ee = slot + a.oE;
    double *__restrict__ tt = slot + a.oTau;
    double *__restrict__ scr = slot + a.oScr;     // [0]=alpha, [8..8+C)=norm, [32..32+C)=p, [64 + 64*rank ..)=t

    extern __shared__ double smem[];
    double *v = smem;                              // [Npad]
    double *tile = smem + lda;                     // [SY_WARPS][32*TPAD]
    __shared__ double sred[SY_WARPS][64];
    __shared__ __align__(32) double pv[32], pw[32], t1s[32], t2s[32];
    __shared__ double sbc[4];

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int gw = rank * SY_WARPS + wid;
    const int nW = C * SY_WARPS;
    double *mytile = tile + wid * (32 * TPAD);
    const int grp = lane >> 2, tig = lane & 3;

    auto first_owned = [&](int I0) { return I0 + ((gw - I0 % nW) + nW) % nW; };

    double pw_last = 0.0;    // W(c, i-1) computed redundantly by every CTA
    PROF_DECL

    for (int k0 = 0; k0 < N; k0 += 32) {
        const int k1 = min(k0 + 32, N);
        // zero the panels for the rows this warp owns
        for (int I = first_owned(k0 / 32); I < NB; I += nW) {
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                Wp[(size_t)q * lda + 32 * I + lane] = 0.0;
                Vp[(size_t)q * lda + 32 * I + lane] = 0.0;
            }
        }
        __syncwarp();

        for (int c = k0; c < k1; ++c) {
            const int i = c - k0;
            // pivot-row entries V(c, j), W(c, j), j < i
            if (wid == 0) {
                // entries j >= i are zero so that the row updates can run over all 32 columns
                double a_ = 0.0, b_ = 0.0;
                if (lane < i) {
                    a_ = (lane == i - 1) ? 1.0 : __ldcg(&Vp[(size_t)lane * lda + c]);
                    b_ = (lane == i - 1) ? pw_last : __ldcg(&Wp[(size_t)lane * lda + c]);
                }
                pv[lane] = a_;
                pw[lane] = b_;
            }
            __syncthreads();

            PROF(0)
            // ---------------- phase 1: lazy update of column c --------------------
            double nrm = 0.0, u1acc = 0.0, u2acc = 0.0;
            for (int I = first_owned(c / 32); I < NB; I += nW) {
                const int r = 32 * I + lane;
                wtr[r] = 0.0;
                const bool live = (r >= c && r < N);
                // panel row (V(r, 0..i), W(r, 0..i)): loaded ONCE, used for the lazy column
                // update and for the panel dot products W^T a, V^T a of the new reflector
                double x[32], y[32];
#pragma unroll
                for (int j0 = 0; j0 < 32; j0 += 8) {
                    if (j0 < i) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            x[j0 + j] = Vp[(size_t)(j0 + j) * lda + r];
                            y[j0 + j] = Wp[(size_t)(j0 + j) * lda + r];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j0 + j] = y[j0 + j] = 0.0;
                    }
                }
                double av = 0.0;
                if (live) {
                    av = __ldcg(&A[r + (size_t)c * lda]);
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) { s0 += x[j] * pw[j]; s1 += y[j] * pv[j]; }
                    av -= s0 + s1;
                    A[r + (size_t)c * lda] = av;
                    if (r >= c + 2) nrm += av * av;
                    if (r == c) dd[c] = av;
                    if (r == c + 1) scr[0] = av;
                }
                if (i > 0) {
                    const double ar = (live && r >= c + 1) ? av : 0.0;     // row c is the diagonal
#pragma unroll
                    for (int j = 0; j < 32; ++j) mytile[j * TPAD + lane] = y[j] * ar;
                    __syncwarp();
                    {
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int q = 0; q < 32; q += 2) { s0 += mytile[lane * TPAD + q]; s1 += mytile[lane * TPAD + q + 1]; }
                        u1acc += s0 + s1;
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) mytile[j * TPAD + lane] = x[j] * ar;
                    __syncwarp();
                    {
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int q = 0; q < 32; q += 2) { s0 += mytile[lane * TPAD + q]; s1 += mytile[lane * TPAD + q + 1]; }
                        u2acc += s0 + s1;
                    }
                    __syncwarp();
                }
            }
            if (c == N - 1) break;
            nrm = warp_sum(nrm);
            sred[wid][lane] = u1acc;
            sred[wid][32 + lane] = u2acc;
            __syncthreads();
            if (wid == 0) {
                double s1 = 0.0, s2 = 0.0;
                for (int q = 0; q < SY_WARPS; ++q) { s1 += sred[q][lane]; s2 += sred[q][32 + lane]; }
                __stcg(&scr[64 + 64 * rank + lane], s1);
                __stcg(&scr[64 + 64 * rank + 32 + lane], s2);
            }
            __syncthreads();
            if (lane == 0) sred[wid][0] = nrm;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int q = 0; q < SY_WARPS; ++q) s += sred[q][0];
                scr[8 + rank] = s;
            }
            PROF(1)
            group_sync<CL>(cluster);                                   // ---- B1
            PROF(2)

            double alpha = __ldcg(&scr[0]);
            double xn2 = 0.0;
            for (int q = 0; q < C; ++q) xn2 += __ldcg(&scr[8 + q]);
            double tau, beta, scale;
            if (xn2 == 0.0) {
                tau = 0.0; beta = alpha; scale = 0.0;
            } else {
                beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            if (gw == 0 && lane == 0) { ee[c] = beta; tt[c] = tau; }
            if (tau == 0.0) {
                // H = I: V column = unit vector (never multiplied by anything but W=0)
                for (int I = first_owned((c + 1) / 32); I < NB; I += nW) {
                    const int r = 32 * I + lane;
                    if (r == c + 1) Vp[(size_t)i * lda + r] = 1.0;
                }
                pw_last = 0.0;
                __syncthreads();
                continue;
            }

            // panel dot products of the scaled reflector:  t = scale * (P^T a) + (1 - scale*alpha) * P(c+1, :)
            if (wid == 0) {
                double s1 = 0.0, s2 = 0.0;
                for (int q = 0; q < C; ++q) {
                    s1 += __ldcg(&scr[64 + 64 * q + lane]);
                    s2 += __ldcg(&scr[64 + 64 * q + 32 + lane]);
                }
                double t1 = 0.0, t2 = 0.0;
                if (lane < i) {
                    const double corr = 1.0 - scale * alpha;
                    t1 = scale * s1 + corr * __ldcg(&Wp[(size_t)lane * lda + c + 1]);
                    t2 = scale * s2 + corr * __ldcg(&Vp[(size_t)lane * lda + c + 1]);
                }
                t1s[lane] = t1;
                t2s[lane] = t2;
            }
            // ---------------- phase 2: v, symmetric mat-vec ------------------------
            for (int r = tid; r < lda; r += SY_THREADS) {
                double x = 0.0;
                if (r == c + 1) x = 1.0;
                else if (r > c + 1 && r < N) x = __ldcg(&A[r + (size_t)c * lda]) * scale;
                v[r] = x;
            }
            __syncthreads();

            PROF(3)
            const int Jmin = (c + 1) / 32;
            // V panel column of the new reflector (rows this warp owns)
            for (int I = first_owned(Jmin); I < NB; I += nW) {
                const int r = 32 * I + lane;
                if (r >= c + 1) Vp[(size_t)i * lda + r] = v[r];
            }
            PROF(4)
            {
                // symmetric mat-vec over the lower-triangular tiles (I >= J >= Jmin), dealt to
                // the cluster's warps in contiguous, equally sized chunks of the J-major tile
                // list (balanced; neighbouring warps stream neighbouring memory).  Both halves
                // of the product are reduced in L2:  w[32I..] += A_IJ v_J,  w[32J..] += A_IJ^T v_I.
                const int m = NB - Jmin;
                const int T = m * (m + 1) / 2;
                const int t0 = (int)(((long long)T * gw) / nW), t1 = (int)(((long long)T * (gw + 1)) / nW);
                int ntile = t1 - t0;
                // tile t0 -> (I, J): column jj holds m - jj tiles
                int jj = (int)(((2.0 * m + 1.0) - sqrt((2.0 * m + 1.0) * (2.0 * m + 1.0) - 8.0 * t0)) * 0.5);
                while (jj > 0 && jj * m - jj * (jj - 1) / 2 > t0) --jj;
                while ((jj + 1) * m - (jj + 1) * jj / 2 <= t0) ++jj;
                int J = Jmin + jj, I = J + (t0 - (jj * m - jj * (jj - 1) / 2));
                double ta[32], tb[32];
                double sacc = 0.0;
                auto load_tile = [&](double(&t)[32], int I_, int J_) {
                    const double *ap = &A[(32 * I_ + lane) + (size_t)(32 * J_) * lda];
#pragma unroll
                    for (int cc = 0; cc < 32; ++cc) t[cc] = __ldcg(ap + (size_t)cc * lda);
                };
                auto consume = [&](const double(&t)[32], int I_, int J_, bool lastOfJ) {
                    const double *vJ = &v[32 * J_];
                    const double vr = v[32 * I_ + lane];
                    double wr0 = 0.0, wr1 = 0.0;
                    if (J_ < I_) {
#pragma unroll
                        for (int cc = 0; cc < 32; cc += 2) {
                            wr0 += t[cc] * vJ[cc];
                            wr1 += t[cc + 1] * vJ[cc + 1];
                            mytile[cc * TPAD + lane] = t[cc] * vr;
                            mytile[(cc + 1) * TPAD + lane] = t[cc + 1] * vr;
                        }
                    } else {
#pragma unroll
                        for (int cc = 0; cc < 32; ++cc) {
                            if (cc <= lane) wr0 += t[cc] * vJ[cc];
                            mytile[cc * TPAD + lane] = (lane > cc) ? t[cc] * vr : 0.0;
                        }
                    }
                    __syncwarp();
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int q = 0; q < 32; q += 2) {
                        s0 += mytile[lane * TPAD + q];
                        s1 += mytile[lane * TPAD + q + 1];
                    }
                    __syncwarp();
                    atomicAdd(&wtr[32 * I_ + lane], wr0 + wr1);
                    sacc += s0 + s1;
                    if (lastOfJ) { atomicAdd(&wtr[32 * J_ + lane], sacc); sacc = 0.0; }
                };
                auto advance = [&](int &I_, int &J_) {
                    if (++I_ >= NB) { ++J_; I_ = J_; }
                };
                if (LEAN) {
                    while (ntile > 0) {
                        int nI = I, nJ = J;
                        advance(nI, nJ);
                        load_tile(ta, I, J);
                        consume(ta, I, J, ntile == 1 || nJ != J);
                        I = nI; J = nJ; --ntile;
                    }
                }
                if (!LEAN && ntile > 0) load_tile(ta, I, J);
                while (!LEAN && ntile > 0) {
                    int nI = I, nJ = J;
                    advance(nI, nJ);
                    if (ntile > 1) load_tile(tb, nI, nJ);
                    consume(ta, I, J, ntile == 1 || nJ != J);
                    I = nI; J = nJ; --ntile;
                    if (ntile == 0) break;
                    advance(nI, nJ);
                    if (ntile > 1) load_tile(ta, nI, nJ);
                    consume(tb, I, J, ntile == 1 || nJ != J);
                    I = nI; J = nJ; --ntile;
                }
            }
            PROF(5)
            group_sync<CL>(cluster);                                   // ---- B2
            PROF(6)

            // ---------------- phase 3 --------------------------------------------
            double pacc = 0.0;
            for (int I = first_owned(Jmin); I < NB; I += nW) {
                const int r = 32 * I + lane;
                if (r >= c + 1 && r < N) {
                    double w = __ldcg(&wtr[r]);
                    w -= panel_row_dot(Vp + r, Wp + r, lda, t1s, t2s, i);
                    w *= tau;
                    pacc += w * v[r];
                    Wp[(size_t)i * lda + r] = w;
                }
            }
            // every CTA recomputes W(c+1, i) (the next pivot-row entry) redundantly
            double wpiv = 0.0;
            if (wid == 0) {
                const int r = c + 1;
                double part = 0.0;
                if (lane < i)
                    part -= __ldcg(&Vp[(size_t)lane * lda + r]) * t1s[lane]
                            + __ldcg(&Wp[(size_t)lane * lda + r]) * t2s[lane];
                part = warp_sum(part);
                wpiv = (part + __ldcg(&wtr[r])) * tau;
            }
            pacc = warp_sum(pacc);
            if (lane == 0) sred[wid][0] = pacc;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int q = 0; q < SY_WARPS; ++q) s += sred[q][0];
                __stcg(&scr[32 + rank], s);
            }
            PROF(7)
            group_sync<CL>(cluster);                                   // ---- B3
            PROF(8)

            // ---------------- phase 4 --------------------------------------------
            double p = 0.0;
            for (int q = 0; q < C; ++q) p += __ldcg(&scr[32 + q]);
            const double alpha2 = -0.5 * tau * p;
            for (int I = first_owned(Jmin); I < NB; I += nW) {
                const int r = 32 * I + lane;
                if (r >= c + 1 && r < N) {
                    Wp[(size_t)i * lda + r] += alpha2 * v[r];
                    if (r >= c + 2) A[r + (size_t)c * lda] = v[r];
                }
            }
            if (wid == 0) {
                if (lane == 0) sbc[0] = wpiv + alpha2;   // v[c+1] == 1
            }
            __syncthreads();
            pw_last = sbc[0];
            PROF(9)
        }  // columns of the panel

        if (k0 + 32 >= N) break;
        group_sync<CL>(cluster);          // panels complete and visible cluster-wide

        // ---------------- trailing update: A -= V W^T + W V^T  (DMMA) --------------
        const int Jt = k0 / 32 + 1;
        double(*BsW)[36] = reinterpret_cast<double(*)[36]>(tile);            // [n][k], stride 36:
        double(*BsV)[36] = reinterpret_cast<double(*)[36]>(tile + 32 * 36);  // conflict-free frags
        for (int J = Jt; J < NB; ++J) {
            __syncthreads();
            for (int t = tid; t < 1024; t += SY_THREADS) {
                // t -> (k = t >> 5, n = t & 31): coalesced along the rows of block column J
                BsW[t & 31][t >> 5] = __ldcg(&Wp[(size_t)(t >> 5) * lda + 32 * J + (t & 31)]);
                BsV[t & 31][t >> 5] = __ldcg(&Vp[(size_t)(t >> 5) * lda + 32 * J + (t & 31)]);
            }
            __syncthreads();
            for (int I = first_owned(J); I < NB; I += nW) {
                // one 32x32 tile: all 32 C values and the 64 A fragments are requested up
                // front (independent loads), then 256 DMMAs on 16 independent accumulators
                const int ar0 = 32 * I + grp;
                double *cp0 = &A[ar0 + (size_t)(32 * J + 2 * tig) * lda];
                double c0[4][4], c1[4][4];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        c0[mt][nt] = __ldcg(cp0 + 8 * mt + (size_t)(8 * nt) * lda);
                        c1[mt][nt] = __ldcg(cp0 + 8 * mt + (size_t)(8 * nt + 1) * lda);
                    }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    double fv[4], fw[4];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        fv[mt] = -Vp[(size_t)(4 * ks + tig) * lda + ar0 + 8 * mt];
                        fw[mt] = -Wp[(size_t)(4 * ks + tig) * lda + ar0 + 8 * mt];
                    }
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const double bw = BsW[8 * nt + grp][4 * ks + tig];
                        const double bv = BsV[8 * nt + grp][4 * ks + tig];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
                            dmma8x8x4(c0[mt][nt], c1[mt][nt], fv[mt], bw);
                            dmma8x8x4(c0[mt][nt], c1[mt][nt], fw[mt], bv);
                        }
                    }
                }
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        cp0[8 * mt + (size_t)(8 * nt) * lda] = c0[mt][nt];
                        cp0[8 * mt + (size_t)(8 * nt + 1) * lda] = c1[mt][nt];
                    }
            }
        }
        PROF(10)
        // other CTAs read this CTA's panel rows as B fragments: nobody may start zeroing
        // the panels for the next block column before every CTA is done with them
        group_sync<CL>(cluster);
    }  // panels
#ifdef SYTRD_PROFILE
    if (mat == 0 && tid == 0) {
        printf("sytrd prof rank %d: col-setup %lld ph1 %lld B1 %lld vload %lld dots %lld symv %lld B2wait %lld ph3 %lld B3wait %lld ph4 %lld update %lld (units of 10 kcycles)\n", rank,
               prof_t[0] / 10000, prof_t[1] / 10000, prof_t[2] / 10000, prof_t[3] / 10000, prof_t[4] / 10000, prof_t[5] / 10000, prof_t[6] / 10000, prof_t[7] / 10000, prof_t[8] / 10000, prof_t[9] / 10000, prof_t[10] / 10000);
    }
#endif
}

}  // namespace

int eig_pick_cluster(int N, int nmat)
{
    // Measured on B200 (N=2363): 1 or 2 CTAs per matrix are the efficient shapes (4.2-4.4
    // ms/matrix); clusters of 4/8 pay for barriers and short per-warp tile lists (8-10
    // ms/matrix).  So: the smallest cluster that still occupies about half of the 148 SMs.
    int c = 1;
    while (c < 8 && nmat * c < 74 && N >= 128 * c) c *= 2;
    return c;
}

int eig_sytrd(const EigSlots &S, int nmat, int cluster, cudaStream_t st, int *launches)
{
    const EigLayout &L = S.L;
    SytrdArgs a;
    a.N = L.N; a.Npad = L.Npad; a.NB = L.NB;
    a.base = S.dbl; a.slot = L.slot_doubles;
    a.oA = L.oA; a.oWp = L.oWp; a.oVp = L.oVp; a.oTrans = L.oTrans; a.oWdir = L.oWdir;
    a.oD = L.oD; a.oE = L.oE; a.oTau = L.oTau; a.oScr = L.oScr;
    const size_t smem = ((size_t)L.Npad + (size_t)SY_WARPS * 32 * TPAD) * sizeof(double);
    if (smem > 220 * 1024) {
        arcb200_set_error("sytrd: N=%d needs %zu bytes of shared memory (max 220 KiB)", L.N, smem);
        return -1;
    }
    if (cluster <= 1) {
        const char *ev = getenv("ARCB200_SYTRD_LEAN");      // tuning knob: force (1) / forbid (0)
        // two matrices per SM (LEAN, <= 128 registers) whenever the batch has more matrices than
        // SMs and two CTAs' shared memory fits: the HBM-bound symv sweeps of one matrix overlap
        // the latency-bound row phases / DMMA update of the other (measured on B200, N=2363:
        // 4.7 -> 4.1 ms per matrix)
        const bool lean = ev ? (ev[0] == '1') : (nmat > 148 && 2 * smem <= 200 * 1024);
        if (lean) {
            ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sytrd_kernel<false, true>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sytrd_kernel<false, true><<<nmat, SY_THREADS, smem, st>>>(a);
        } else {
            ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sytrd_kernel<false, false>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sytrd_kernel<false, false><<<nmat, SY_THREADS, smem, st>>>(a);
        }
    } else {
        ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sytrd_kernel<true, false>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (cluster > 8)
            ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sytrd_kernel<true, false>,
                                                    cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nmat * cluster);
        cfg.blockDim = dim3(SY_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        ARCB200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sytrd_kernel<true, false>, a));
    }
    ARCB200_LAUNCH_CHECK("sytrd_kernel");
    if (launches) *launches += 1;
    return 0;
}
