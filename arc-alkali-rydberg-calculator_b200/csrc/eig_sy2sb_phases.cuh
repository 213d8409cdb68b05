// eig_sy2sb_phases.cuh -- the five phases of one panel step of the band reduction (stage 1 of the
// two-stage tridiagonalisation) as device functions, shared by the one-CTA-per-matrix kernels
// (eig_twostage.cu) and the per-phase, many-CTAs-per-matrix kernels (eig_sy2sb_multi.cu).
#pragma once
#include "eig_twostage.cuh"

namespace ts {

// ------------------------------------------------------------------------------------------
// stage 1: dense -> band
// ------------------------------------------------------------------------------------------
// panel QR, T factor, symmetric product X = A22 V, W, rank-2k update.
constexpr int S1_TSTR = 36;      // per-warp transit tile [cc][TSTR] (column-major): conflict-free LDS.64
                                 // fragments in both orientations

struct S1Ctx {
    int tid, wid, lane, grp, tig;
    int np, lda, NB;
    double *A, *Wp, *Vp, *tt, *Tfac;
    double(*Bs)[2][32][BSTR];     // operand tiles [buf][op][n][k]
    double(*Ts)[33];
    double(*G)[33];
    double(*M1)[33];
    double(*wv)[32];
    double(*sred)[16];
    double *taus;
    double *Tw;                   // this warp's two transit tiles
};

// Householder QR of the panel below the band (in place) + clean V panel
template <int WARPS>
__device__ __forceinline__ void s1_panel_qr(const S1Ctx &c, int p)
{
    const int tid = c.tid, wid = c.wid, lane = c.lane, grp = c.grp, tig = c.tig;
    const int np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *__restrict__ tt = c.tt;
    double *__restrict__ Tfac = c.Tfac;
    double(*Bs)[2][32][BSTR] = c.Bs;
    double(*Ts)[33] = c.Ts;
    double(*G)[33] = c.G;
    double(*M1)[33] = c.M1;
    double(*wv)[32] = c.wv;
    double(*sred)[16] = c.sred;
    double *taus = c.taus;
    double *Tw = c.Tw;
    constexpr int NT = WARPS * 32;
    constexpr int TSTR = S1_TSTR;
    const int m0 = 32 * (p + 1), col0 = 32 * p;
    const int nch = NB - (p + 1);
    double *P = A + (size_t)col0 * lda;                   // P[r + j*lda], r absolute
    const int Jt = p + 1;
    constexpr int SV = 32 / WARPS;
    (void)tid; (void)grp; (void)tig; (void)np; (void)Wp; (void)Vp; (void)tt; (void)Tfac; (void)Bs; (void)Ts; (void)G;
    (void)M1; (void)wv; (void)sred; (void)taus; (void)Tw; (void)NT; (void)m0; (void)col0; (void)nch; (void)P; (void)Jt;
    (void)SV; (void)lane; (void)wid; (void)lda; (void)NB; (void)A;
    // asynchronous copy of the stored 32x32 tile (Ib, Jb), Ib >= Jb, into this warp's transit
    // tile: 16-byte pieces, whole 256-byte column segments per half warp
    auto issue_tile = [&](int Ib, int Jb, int tb) {
        // a real loop with running pointers: sixteen hoisted 64-bit addresses would cost 48
        // registers in the middle of the tensor-core loops
        const double *src = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
        const size_t sstep = 2 * (size_t)lda;
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            cp_async16_t(dst, src);
            src += sstep;
            dst += 2 * TSTR;
        }
    };
    (void)issue_tile;
    // ---------------------------------------------------------------- panel QR
    // Two phases per column c (c = -1: gather only), lane <-> row everywhere:
    //   A  rows dealt to the warps: finalise column c (v = x * scale, beta on the diagonal),
    //      apply H_c to column cn = c + 1 and gather its norm and pivot  -> block reduce
    //   B  columns dealt to the warps (j = wid + WARPS k): apply H_c to the own columns j > cn
    //      and gather g_j = sum_{r > rd+1} x_cn(r) x_j(r) and the pivot row, so that
    //      tau' w'_j = tau' (x_j(rd+1) + scale' g_j) is known to the owner without another pass.
    // Loads are issued in register batches ahead of the dependent stores.
    {
        constexpr int KO = 32 / WARPS;
        double wreg[KO];
#pragma unroll
        for (int k = 0; k < KO; ++k) wreg[k] = 0.0;
        double tau = 0.0, scale = 0.0, beta = 0.0;
        for (int c = -1; c < 32; ++c) {
            const int rd = m0 + c, cn = c + 1, par = cn & 1;
            // ---------------- phase A
            double ssq = 0.0;
            const double wcn = (c >= 0 && cn < 32) ? wv[c & 1][cn] : 0.0;
            for (int ch0 = wid; ch0 < nch; ch0 += WARPS * 4) {
                double pc[4], pn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ch = ch0 + WARPS * u, r = m0 + 32 * ch + lane;
                    pc[u] = (c >= 0 && ch < nch) ? P[r + (size_t)c * lda] : 0.0;
                    pn[u] = (cn < 32 && ch < nch) ? P[r + (size_t)cn * lda] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ch = ch0 + WARPS * u, r = m0 + 32 * ch + lane;
                    if (ch < nch) {
                        double v = 0.0;
                        if (c >= 0) {
                            if (r == rd) {
                                v = 1.0;
                                P[r + (size_t)c * lda] = beta;
                            } else if (r > rd) {
                                v = pc[u] * scale;
                                P[r + (size_t)c * lda] = v;
                            }
                            Vp[(size_t)c * np + r] = v;
                        }
                        if (cn < 32 && r >= rd) {
                            double x = pn[u];
                            if (c >= 0) {
                                x -= v * wcn;
                                P[r + (size_t)cn * lda] = x;
                            }
                            if (r == rd + 1) sred[par][15] = x;
                            if (r > rd + 1) ssq += x * x;
                        }
                    }
                }
            }
            if (cn >= 32) break;
            ssq = warp_sum(ssq);
            if (lane == 0) sred[par][wid] = ssq;
            __syncthreads();                                     // #1
            double xn2 = 0.0;
#pragma unroll
            for (int q = 0; q < WARPS; ++q) xn2 += sred[par][q];
            const double alpha = sred[par][15];
            double tau_n = 0.0, beta_n = alpha, scale_n = 0.0;
            if (rd + 1 < np - 1 && xn2 != 0.0) {
                beta_n = -copysign(sqrt(alpha * alpha + xn2), alpha);
                tau_n = (beta_n - alpha) / beta_n;
                scale_n = 1.0 / (alpha - beta_n);
            }
            if (tid == 0) { tt[col0 + cn] = tau_n; taus[cn] = tau_n; }
            // ---------------- phase B
            double g[KO], h[KO];
#pragma unroll
            for (int k = 0; k < KO; ++k) g[k] = h[k] = 0.0;
            // the warp streams (column c, column cn, own columns) x all rows through a ring of
            // chunk slots in its transit-tile area with cp.async, RD chunks ahead: the pass is
            // no longer a chain of exposed DRAM round trips.  lane <-> row for the copy and
            // for the arithmetic, so only the thread's own copies have to be waited for.
            {
                constexpr int SLOTW = (2 + KO) * 32;
                constexpr int NSLOT = (2 * 32 * TSTR) / SLOTW;
                constexpr int RD = (NSLOT - 2 < 8) ? NSLOT - 2 : 8;
                double *ring = Tw;
                auto issueB = [&](int ch) {
                    if (ch < nch) {
                        const int r = m0 + 32 * ch + lane;
                        double *dst = ring + (ch % NSLOT) * SLOTW + lane;
                        if (c >= 0) cp_async8_t(dst, Vp + (size_t)c * np + r);
                        cp_async8_t(dst + 32, P + r + (size_t)cn * lda);
#pragma unroll
                        for (int k = 0; k < KO; ++k) {
                            const int j = wid + WARPS * k;
                            if (j > cn) cp_async8_t(dst + 64 + 32 * k, P + r + (size_t)j * lda);
                        }
                    }
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                };
#pragma unroll 1
                for (int ch = 0; ch < RD; ++ch) issueB(ch);
#pragma unroll 1
                for (int ch = 0; ch < nch; ++ch) {
                    issueB(ch + RD);
                    asm volatile("cp.async.wait_group %0;\n" ::"n"(RD) : "memory");
                    const int r = m0 + 32 * ch + lane;
                    if (r >= rd) {
                        const double *sl = ring + (ch % NSLOT) * SLOTW + lane;
                        const double v = (c >= 0) ? sl[0] : 0.0;
                        const double xc = sl[32];
                        const double xg = (r > rd + 1) ? xc : 0.0;
#pragma unroll
                        for (int k = 0; k < KO; ++k) {
                            const int j = wid + WARPS * k;
                            if (j > cn) {
                                double x = sl[64 + 32 * k];
                                if (c >= 0) {
                                    x -= v * wreg[k];
                                    P[r + (size_t)j * lda] = x;
                                }
                                g[k] += xg * x;
                                if (r == rd + 1) h[k] = x;
                            }
                        }
                    }
                }
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
#pragma unroll
            for (int k = 0; k < KO; ++k) {
                g[k] = warp_sum(g[k]);
                h[k] = __shfl_sync(ARCB200_FULL_MASK, h[k], cn & 31);   // row m0 + cn: chunk 0, lane cn
                wreg[k] = tau_n * (h[k] + scale_n * g[k]);
                if (lane == 0) wv[par][wid + WARPS * k] = wreg[k];
            }
            tau = tau_n; beta = beta_n; scale = scale_n;
            __syncthreads();                                     // #2
        }
    }
    __syncthreads();
}

// compact-WY factor T of the panel reflectors (Gram matrix on DMMA), stored to Tfac
template <int WARPS>
__device__ __forceinline__ void s1_tfactor(const S1Ctx &c, int p)
{
    const int tid = c.tid, wid = c.wid, lane = c.lane, grp = c.grp, tig = c.tig;
    const int np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *__restrict__ tt = c.tt;
    double *__restrict__ Tfac = c.Tfac;
    double(*Bs)[2][32][BSTR] = c.Bs;
    double(*Ts)[33] = c.Ts;
    double(*G)[33] = c.G;
    double(*M1)[33] = c.M1;
    double(*wv)[32] = c.wv;
    double(*sred)[16] = c.sred;
    double *taus = c.taus;
    double *Tw = c.Tw;
    constexpr int NT = WARPS * 32;
    constexpr int TSTR = S1_TSTR;
    const int m0 = 32 * (p + 1), col0 = 32 * p;
    const int nch = NB - (p + 1);
    double *P = A + (size_t)col0 * lda;                   // P[r + j*lda], r absolute
    const int Jt = p + 1;
    constexpr int SV = 32 / WARPS;
    (void)tid; (void)grp; (void)tig; (void)np; (void)Wp; (void)Vp; (void)tt; (void)Tfac; (void)Bs; (void)Ts; (void)G;
    (void)M1; (void)wv; (void)sred; (void)taus; (void)Tw; (void)NT; (void)m0; (void)col0; (void)nch; (void)P; (void)Jt;
    (void)SV; (void)lane; (void)wid; (void)lda; (void)NB; (void)A;
    // asynchronous copy of the stored 32x32 tile (Ib, Jb), Ib >= Jb, into this warp's transit
    // tile: 16-byte pieces, whole 256-byte column segments per half warp
    auto issue_tile = [&](int Ib, int Jb, int tb) {
        // a real loop with running pointers: sixteen hoisted 64-bit addresses would cost 48
        // registers in the middle of the tensor-core loops
        const double *src = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
        const size_t sstep = 2 * (size_t)lda;
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            cp_async16_t(dst, src);
            src += sstep;
            dst += 2 * TSTR;
        }
    };
    (void)issue_tile;
    // ---------------------------------------------------------------- T factor
    for (int t = tid; t < 1024; t += NT) { G[t >> 5][t & 31] = 0.0; Ts[t >> 5][t & 31] = 0.0; }
    __syncthreads();
    {
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int ch = wid; ch < nch; ch += WARPS) {
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
    for (int t = tid; t < 1024; t += NT) Tfac[(size_t)p * 1024 + t] = Ts[t >> 5][t & 31];
}

// X = A22 V, all WARPS warps in lockstep over q (one block barrier per step); row groups gfirst,
// gfirst + gstride, ... of WARPS block rows (one CTA per matrix: all of them)
template <int WARPS>
__device__ __forceinline__ void s1_symm_lockstep(const S1Ctx &c, int p, int gfirst = 0, int gstride = 1)
{
    const int tid = c.tid, wid = c.wid, lane = c.lane, grp = c.grp, tig = c.tig;
    const int np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *__restrict__ tt = c.tt;
    double *__restrict__ Tfac = c.Tfac;
    double(*Bs)[2][32][BSTR] = c.Bs;
    double(*Ts)[33] = c.Ts;
    double(*G)[33] = c.G;
    double(*M1)[33] = c.M1;
    double(*wv)[32] = c.wv;
    double(*sred)[16] = c.sred;
    double *taus = c.taus;
    double *Tw = c.Tw;
    constexpr int NT = WARPS * 32;
    constexpr int TSTR = S1_TSTR;
    const int m0 = 32 * (p + 1), col0 = 32 * p;
    const int nch = NB - (p + 1);
    double *P = A + (size_t)col0 * lda;                   // P[r + j*lda], r absolute
    const int Jt = p + 1;
    constexpr int SV = 32 / WARPS;
    (void)tid; (void)grp; (void)tig; (void)np; (void)Wp; (void)Vp; (void)tt; (void)Tfac; (void)Bs; (void)Ts; (void)G;
    (void)M1; (void)wv; (void)sred; (void)taus; (void)Tw; (void)NT; (void)m0; (void)col0; (void)nch; (void)P; (void)Jt;
    (void)SV; (void)lane; (void)wid; (void)lda; (void)NB; (void)A;
    // asynchronous copy of the stored 32x32 tile (Ib, Jb), Ib >= Jb, into this warp's transit
    // tile: 16-byte pieces, whole 256-byte column segments per half warp
    auto issue_tile = [&](int Ib, int Jb, int tb) {
        // a real loop with running pointers: sixteen hoisted 64-bit addresses would cost 48
        // registers in the middle of the tensor-core loops
        const double *src = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
        const size_t sstep = 2 * (size_t)lda;
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            cp_async16_t(dst, src);
            src += sstep;
            dst += 2 * TSTR;
        }
    };
    (void)issue_tile;
    // ---------------------------------------------------------------- X = A22 V
    // Row block i of X = sum_q Asym(i, q) V_q: warp <-> i (groups of WARPS block rows).  The
    // stored tile (lower triangle; transposed use when q > i) arrives in the warp's transit
    // tile by cp.async while the previous product runs; V_q is shared through shared memory.
            for (int g0 = Jt + WARPS * gfirst; g0 < NB; g0 += WARPS * gstride) {    // the caller's share of the row groups
        const int i = g0 + wid;
        const bool act = i < NB;
        double acc[4][4][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
        auto fillV = [&](int buf, int q) {
#pragma unroll
            for (int n = 0; n < SV; ++n)
                cp_async8_t(&Bs[buf][0][wid + WARPS * n][lane], &Vp[(size_t)(wid + WARPS * n) * np + 32 * q + lane]);
        };
        __syncthreads();
        if (act) issue_tile(max(i, Jt), min(i, Jt), 0);
        fillV(0, Jt);
        cp_async_wait_all_t();
        __syncthreads();
        for (int q = Jt; q < NB; ++q) {
            const int buf = (q - Jt) & 1;
            if (q + 1 < NB) {
                if (act) issue_tile(max(i, q + 1), min(i, q + 1), buf ^ 1);
                fillV(buf ^ 1, q + 1);
            }
            if (act) {
                // fragments straight from the transit tile inside the k loop (no register
                // array); diagonal tiles are kept fully symmetric, so q == i reads like a
                // stored tile
                const double *tw = Tw + buf * (32 * TSTR);
                if (q <= i) {
                    const double *pd = tw + tig * TSTR + grp;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        double fa[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) fa[x] = pd[4 * ks * TSTR + 8 * x];
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            const double b = Bs[buf][0][8 * y + grp][4 * ks + tig];
#pragma unroll
                            for (int x = 0; x < 4; ++x) dmma_t(acc[x][y][0], acc[x][y][1], fa[x], b);
                        }
                    }
                } else {
                    const double *pt = tw + grp * TSTR + tig;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        double fa[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) fa[x] = pt[8 * x * TSTR + 4 * ks];
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            const double b = Bs[buf][0][8 * y + grp][4 * ks + tig];
#pragma unroll
                            for (int x = 0; x < 4; ++x) dmma_t(acc[x][y][0], acc[x][y][1], fa[x], b);
                        }
                    }
                }
            }
            cp_async_wait_all_t();
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
}

// W = X T - 1/2 V (T^T (V^T X) T)
template <int WARPS>
__device__ __forceinline__ void s1_wphase(const S1Ctx &c, int p)
{
    const int tid = c.tid, wid = c.wid, lane = c.lane, grp = c.grp, tig = c.tig;
    const int np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *__restrict__ tt = c.tt;
    double *__restrict__ Tfac = c.Tfac;
    double(*Bs)[2][32][BSTR] = c.Bs;
    double(*Ts)[33] = c.Ts;
    double(*G)[33] = c.G;
    double(*M1)[33] = c.M1;
    double(*wv)[32] = c.wv;
    double(*sred)[16] = c.sred;
    double *taus = c.taus;
    double *Tw = c.Tw;
    constexpr int NT = WARPS * 32;
    constexpr int TSTR = S1_TSTR;
    const int m0 = 32 * (p + 1), col0 = 32 * p;
    const int nch = NB - (p + 1);
    double *P = A + (size_t)col0 * lda;                   // P[r + j*lda], r absolute
    const int Jt = p + 1;
    constexpr int SV = 32 / WARPS;
    (void)tid; (void)grp; (void)tig; (void)np; (void)Wp; (void)Vp; (void)tt; (void)Tfac; (void)Bs; (void)Ts; (void)G;
    (void)M1; (void)wv; (void)sred; (void)taus; (void)Tw; (void)NT; (void)m0; (void)col0; (void)nch; (void)P; (void)Jt;
    (void)SV; (void)lane; (void)wid; (void)lda; (void)NB; (void)A;
    // asynchronous copy of the stored 32x32 tile (Ib, Jb), Ib >= Jb, into this warp's transit
    // tile: 16-byte pieces, whole 256-byte column segments per half warp
    auto issue_tile = [&](int Ib, int Jb, int tb) {
        // a real loop with running pointers: sixteen hoisted 64-bit addresses would cost 48
        // registers in the middle of the tensor-core loops
        const double *src = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
        const size_t sstep = 2 * (size_t)lda;
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            cp_async16_t(dst, src);
            src += sstep;
            dst += 2 * TSTR;
        }
    };
    (void)issue_tile;
    // ---------------------------------------------------------------- W
    for (int t = tid; t < 1024; t += NT) G[t >> 5][t & 31] = 0.0;
    __syncthreads();
    {
        // S = V^T X
        double acc[4][4][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
        for (int ch = wid; ch < nch; ch += WARPS) {
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
    for (int t = tid; t < 1024; t += NT) {
        const int row = t >> 5, col = t & 31;
        double s = 0.0;
        for (int l = 0; l <= col; ++l) s += G[row][l] * Ts[l][col];
        M1[row][col] = s;
    }
    __syncthreads();
    // M2 = T^T M1 -> G
    for (int t = tid; t < 1024; t += NT) {
        const int row = t >> 5, col = t & 31;
        double s = 0.0;
        for (int l = 0; l <= row; ++l) s += Ts[l][row] * M1[l][col];
        G[row][col] = s;
    }
    __syncthreads();
    // operand tiles [n][k]: T and -M2/2
    for (int t = tid; t < 1024; t += NT) {
        const int n = t >> 5, k = t & 31;
        Bs[0][0][n][k] = Ts[k][n];
        Bs[0][1][n][k] = -0.5 * G[k][n];
    }
    __syncthreads();
    for (int ch = wid; ch < nch; ch += WARPS) {
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
}

// A22 -= V W^T + W V^T, all WARPS warps in lockstep over J (one block barrier per step); row groups as above
template <int WARPS>
__device__ __forceinline__ void s1_syr2k_lockstep(const S1Ctx &c, int p, int gfirst = 0, int gstride = 1)
{
    const int tid = c.tid, wid = c.wid, lane = c.lane, grp = c.grp, tig = c.tig;
    const int np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *__restrict__ tt = c.tt;
    double *__restrict__ Tfac = c.Tfac;
    double(*Bs)[2][32][BSTR] = c.Bs;
    double(*Ts)[33] = c.Ts;
    double(*G)[33] = c.G;
    double(*M1)[33] = c.M1;
    double(*wv)[32] = c.wv;
    double(*sred)[16] = c.sred;
    double *taus = c.taus;
    double *Tw = c.Tw;
    constexpr int NT = WARPS * 32;
    constexpr int TSTR = S1_TSTR;
    const int m0 = 32 * (p + 1), col0 = 32 * p;
    const int nch = NB - (p + 1);
    double *P = A + (size_t)col0 * lda;                   // P[r + j*lda], r absolute
    const int Jt = p + 1;
    constexpr int SV = 32 / WARPS;
    (void)tid; (void)grp; (void)tig; (void)np; (void)Wp; (void)Vp; (void)tt; (void)Tfac; (void)Bs; (void)Ts; (void)G;
    (void)M1; (void)wv; (void)sred; (void)taus; (void)Tw; (void)NT; (void)m0; (void)col0; (void)nch; (void)P; (void)Jt;
    (void)SV; (void)lane; (void)wid; (void)lda; (void)NB; (void)A;
    // asynchronous copy of the stored 32x32 tile (Ib, Jb), Ib >= Jb, into this warp's transit
    // tile: 16-byte pieces, whole 256-byte column segments per half warp
    auto issue_tile = [&](int Ib, int Jb, int tb) {
        // a real loop with running pointers: sixteen hoisted 64-bit addresses would cost 48
        // registers in the middle of the tensor-core loops
        const double *src = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
        const size_t sstep = 2 * (size_t)lda;
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            cp_async16_t(dst, src);
            src += sstep;
            dst += 2 * TSTR;
        }
    };
    (void)issue_tile;
    // ---------------------------------------------------------------- A22 -= V W^T + W V^T
    // warp <-> block row I (groups of WARPS): the A-operand fragments -V_I, -W_I stay in
    // registers for the whole sweep over J <= I; W_J, V_J are shared through shared memory;
    // the next C tile arrives in the transit tile by cp.async during the 256 DMMAs.
    // Fragment row mapping here: DMMA row (x, grp) <-> tile row 4 grp + x, so that a lane's
    // four rows are contiguous (128-bit shared loads and global stores).
    for (int g0 = Jt + WARPS * gfirst; g0 < NB; g0 += WARPS * gstride) {    // the caller's share of the row groups
        const int I = g0 + wid;
        const bool act = I < NB;
        const int Jmax = min(g0 + WARPS - 1, NB - 1);
        // -W_I in registers; -V_I in the warp's second transit tile ([k][row], pitch TSTR)
        double fw[4][8];
        if (act) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const double2 *pw2 = reinterpret_cast<const double2 *>(Wp + (size_t)(4 * ks + tig) * np + 32 * I + 4 * grp);
                const double2 w0 = pw2[0], w1 = pw2[1];
                fw[0][ks] = -w0.x; fw[1][ks] = -w0.y; fw[2][ks] = -w1.x; fw[3][ks] = -w1.y;
            }
        }
        double *Tv = Tw + 32 * TSTR;
        auto fillWV = [&](int buf, int J) {
#pragma unroll
            for (int n = 0; n < SV; ++n) {
                cp_async8_t(&Bs[buf][0][lane][wid + WARPS * n], &Wp[(size_t)(wid + WARPS * n) * np + 32 * J + lane]);
                cp_async8_t(&Bs[buf][1][lane][wid + WARPS * n], &Vp[(size_t)(wid + WARPS * n) * np + 32 * J + lane]);
            }
        };
        __syncthreads();
        if (act) {
            issue_tile(I, Jt, 0);
            // V_I: panel columns k = 0..31, rows 32 I .. 32 I + 31
            const double *src = Vp + (size_t)(lane >> 4) * np + 32 * I + (lane & 15) * 2;
            double *dst = Tv + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll 1
            for (int k = 0; k < 16; ++k) {
                cp_async16_t(dst, src);
                src += 2 * (size_t)np;
                dst += 2 * TSTR;
            }
        }
        fillWV(0, Jt);
        cp_async_wait_all_t();
        __syncwarp();
        if (act) {
#pragma unroll 4
            for (int r = 0; r < 32; ++r) Tv[lane * TSTR + r] = -Tv[lane * TSTR + r];
        }
        __syncthreads();
        const double *tv = Tv + tig * TSTR + 4 * grp;
        const double *tc = Tw + (2 * tig) * TSTR + 4 * grp;
        for (int J = Jt; J <= Jmax; ++J) {
            const int buf = (J - Jt) & 1;
            const bool work = act && J <= I;
            double c0[4][4], c1[4][4];      // [x][y]: rows 4 grp + x, columns 8 y + 2 tig (+1)
            if (work) {
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const double2 *q0 = reinterpret_cast<const double2 *>(tc + 8 * y * TSTR);
                    const double2 *q1 = reinterpret_cast<const double2 *>(tc + (8 * y + 1) * TSTR);
                    const double2 a0 = q0[0], a1 = q0[1], b0 = q1[0], b1 = q1[1];
                    c0[0][y] = a0.x; c0[1][y] = a0.y; c0[2][y] = a1.x; c0[3][y] = a1.y;
                    c1[0][y] = b0.x; c1[1][y] = b0.y; c1[2][y] = b1.x; c1[3][y] = b1.y;
                }
            }
            __syncwarp();
            if (J + 1 <= Jmax) {
                if (act && J + 1 <= I) issue_tile(I, J + 1, 0);
                fillWV(buf ^ 1, J + 1);
            }
            if (work) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const double2 va = *reinterpret_cast<const double2 *>(tv + 4 * ks * TSTR);
                    const double2 vb = *reinterpret_cast<const double2 *>(tv + 4 * ks * TSTR + 2);
                    const double fv[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const double bw = Bs[buf][0][8 * y + grp][4 * ks + tig];
                        const double bv = Bs[buf][1][8 * y + grp][4 * ks + tig];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            dmma_t(c0[x][y], c1[x][y], fv[x], bw);
                            dmma_t(c0[x][y], c1[x][y], fw[x][ks], bv);
                        }
                    }
                }
                double *cp0 = &A[32 * I + 4 * grp + (size_t)(32 * J + 2 * tig) * lda];
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    double2 *o0 = reinterpret_cast<double2 *>(cp0 + (size_t)(8 * y) * lda);
                    double2 *o1 = reinterpret_cast<double2 *>(cp0 + (size_t)(8 * y + 1) * lda);
                    o0[0] = make_double2(c0[0][y], c0[1][y]);
                    o0[1] = make_double2(c0[2][y], c0[3][y]);
                    o1[0] = make_double2(c1[0][y], c1[1][y]);
                    o1[1] = make_double2(c1[2][y], c1[3][y]);
                }
            }
            cp_async_wait_all_t();
            __syncthreads();
        }
    }
}

}  // namespace ts
