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

// Householder QR of the panel below the band (in place) + clean V panel.  TWD = doubles of shared memory behind
// c.Tw per warp (the chunk ring of phase B lives there)
template <int WARPS, int TWD = 2 * 32 * S1_TSTR>
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
    const int nc2 = (nch + 1) >> 1;                       // 64-row chunks of the panel QR
    double *P = A + (size_t)col0 * lda;                   // P[r + j*lda], r absolute
    const int Jt = p + 1;
    constexpr int SV = 32 / WARPS;
    (void)tid; (void)grp; (void)tig; (void)np; (void)Wp; (void)Vp; (void)tt; (void)Tfac; (void)Bs; (void)Ts; (void)G;
    (void)M1; (void)wv; (void)sred; (void)taus; (void)Tw; (void)NT; (void)m0; (void)col0; (void)nch; (void)P; (void)Jt;
    (void)SV; (void)lane; (void)wid; (void)lda; (void)NB; (void)A;
    // asynchronous copy of the stored 32x32 tile (Ib, Jb), Ib >= Jb, into this warp's transit
    // tile: 16-byte pieces, whole 256-byte column segments per half warp
    auto issue_tile = [&](int Ib, int Jb, int tb) {
        // two running source pointers: a cp.async holds its address registers until the LSU has read
        // them, so a single pointer bumped 16 times serialises the copies (~40 cycles each)
        const size_t sstep = 2 * (size_t)lda;
        const double *s0 = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        const double *s1 = s0 + sstep;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            cp_async16_t(dst, s0);
            cp_async16_t(dst + 2 * TSTR, s1);
            s0 += 2 * sstep;
            s1 += 2 * sstep;
            dst += 4 * TSTR;
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
            // ---------------- phase A   (lane <-> two consecutive rows: 64-row chunks, 16-byte accesses)
            double ssq = 0.0;
            const double wcn = (c >= 0 && cn < 32) ? wv[c & 1][cn] : 0.0;
            for (int ch0 = wid; ch0 < nc2; ch0 += WARPS * 4) {
                double2 pc[4], pn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ch = ch0 + WARPS * u, r = m0 + 64 * ch + 2 * lane;
                    const bool ok = ch < nc2 && r < np;
                    pc[u] = (c >= 0 && ok) ? *reinterpret_cast<const double2 *>(&P[r + (size_t)c * lda]) : make_double2(0.0, 0.0);
                    pn[u] = (cn < 32 && ok) ? *reinterpret_cast<const double2 *>(&P[r + (size_t)cn * lda]) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ch = ch0 + WARPS * u, r = m0 + 64 * ch + 2 * lane;
                    if (ch < nc2 && r < np) {
                        double2 v = make_double2(0.0, 0.0);
                        if (c >= 0) {
                            double2 o = pc[u];                   // rows above the diagonal of R stay as they are
                            if (r == rd) { v.x = 1.0; o.x = beta; } else if (r > rd) { v.x = pc[u].x * scale; o.x = v.x; }
                            if (r + 1 == rd) { v.y = 1.0; o.y = beta; } else if (r + 1 > rd) { v.y = pc[u].y * scale; o.y = v.y; }
                            if (r + 1 >= rd) *reinterpret_cast<double2 *>(&P[r + (size_t)c * lda]) = o;
                            *reinterpret_cast<double2 *>(&Vp[(size_t)c * np + r]) = v;
                        }
                        if (cn < 32 && r + 1 >= rd) {
                            double2 x = pn[u];
                            if (c >= 0) {
                                x.x -= v.x * wcn;                // v = 0 above row rd
                                x.y -= v.y * wcn;
                                *reinterpret_cast<double2 *>(&P[r + (size_t)cn * lda]) = x;
                            }
                            if (r == rd + 1) sred[par][15] = x.x;
                            if (r + 1 == rd + 1) sred[par][15] = x.y;
                            if (r > rd + 1) ssq += x.x * x.x;
                            if (r + 1 > rd + 1) ssq += x.y * x.y;
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
            // no longer a chain of exposed DRAM round trips.  lane <-> row pair for the copy and
            // for the arithmetic, so only the thread's own copies have to be waited for.
            {
                constexpr int SLOTW = (2 + KO) * 64;
                constexpr int NSLOT = TWD / SLOTW;
                constexpr int RD = (NSLOT - 1 < 4) ? NSLOT - 1 : 4;
                static_assert(RD >= 1, "phase-B ring: at least two chunk slots per warp");
                double *ring = Tw;
                auto issueB = [&](int ch) {
                    const int r = m0 + 64 * ch + 2 * lane;
                    if (ch < nc2 && r < np) {
                        double *dst = ring + (ch % NSLOT) * SLOTW + 2 * lane;
                        if (c >= 0) cp_async16_t(dst, Vp + (size_t)c * np + r);
                        cp_async16_t(dst + 64, P + r + (size_t)cn * lda);
#pragma unroll
                        for (int k = 0; k < KO; ++k) {
                            const int j = wid + WARPS * k;
                            if (j > cn) cp_async16_t(dst + 128 + 64 * k, P + r + (size_t)j * lda);
                        }
                    }
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                };
#pragma unroll 1
                for (int ch = 0; ch < RD; ++ch) issueB(ch);
#pragma unroll 1
                for (int ch = 0; ch < nc2; ++ch) {
                    issueB(ch + RD);
                    asm volatile("cp.async.wait_group %0;\n" ::"n"(RD) : "memory");
                    const int r = m0 + 64 * ch + 2 * lane;
                    if (r + 1 >= rd && r < np) {
                        const double *sl = ring + (ch % NSLOT) * SLOTW + 2 * lane;
                        const double2 v = (c >= 0) ? *reinterpret_cast<const double2 *>(sl) : make_double2(0.0, 0.0);
                        const double2 xc = *reinterpret_cast<const double2 *>(sl + 64);
                        const double xg0 = (r > rd + 1) ? xc.x : 0.0;
                        const double xg1 = (r + 1 > rd + 1) ? xc.y : 0.0;
#pragma unroll
                        for (int k = 0; k < KO; ++k) {
                            const int j = wid + WARPS * k;
                            if (j > cn) {
                                double2 x = *reinterpret_cast<const double2 *>(sl + 128 + 64 * k);
                                if (c >= 0) {
                                    x.x -= v.x * wreg[k];        // v = 0 above row rd
                                    x.y -= v.y * wreg[k];
                                    *reinterpret_cast<double2 *>(&P[r + (size_t)j * lda]) = x;
                                }
                                g[k] += xg0 * x.x + xg1 * x.y;
                                if (r == rd + 1) h[k] = x.x;
                                if (r + 1 == rd + 1) h[k] = x.y;
                            }
                        }
                    }
                }
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
#pragma unroll
            for (int k = 0; k < KO; ++k) {
                g[k] = warp_sum(g[k]);
                h[k] = __shfl_sync(ARCB200_FULL_MASK, h[k], (cn >> 1) & 31);   // row m0 + cn: chunk 0, lane cn / 2
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
        // two running source pointers: a cp.async holds its address registers until the LSU has read
        // them, so a single pointer bumped 16 times serialises the copies (~40 cycles each)
        const size_t sstep = 2 * (size_t)lda;
        const double *s0 = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        const double *s1 = s0 + sstep;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            cp_async16_t(dst, s0);
            cp_async16_t(dst + 2 * TSTR, s1);
            s0 += 2 * sstep;
            s1 += 2 * sstep;
            dst += 4 * TSTR;
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
        // two running source pointers: a cp.async holds its address registers until the LSU has read
        // them, so a single pointer bumped 16 times serialises the copies (~40 cycles each)
        const size_t sstep = 2 * (size_t)lda;
        const double *s0 = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        const double *s1 = s0 + sstep;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            cp_async16_t(dst, s0);
            cp_async16_t(dst + 2 * TSTR, s1);
            s0 += 2 * sstep;
            s1 += 2 * sstep;
            dst += 4 * TSTR;
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
        // two running source pointers: a cp.async holds its address registers until the LSU has read
        // them, so a single pointer bumped 16 times serialises the copies (~40 cycles each)
        const size_t sstep = 2 * (size_t)lda;
        const double *s0 = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        const double *s1 = s0 + sstep;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            cp_async16_t(dst, s0);
            cp_async16_t(dst + 2 * TSTR, s1);
            s0 += 2 * sstep;
            s1 += 2 * sstep;
            dst += 4 * TSTR;
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
        // two running source pointers: a cp.async holds its address registers until the LSU has read
        // them, so a single pointer bumped 16 times serialises the copies (~40 cycles each)
        const size_t sstep = 2 * (size_t)lda;
        const double *s0 = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
        const double *s1 = s0 + sstep;
        double *dst = Tw + tb * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            cp_async16_t(dst, s0);
            cp_async16_t(dst + 2 * TSTR, s1);
            s0 += 2 * sstep;
            s1 += 2 * sstep;
            dst += 4 * TSTR;
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
                        // the two products of an accumulator four DMMAs apart (asm volatile keeps this order): issued
                        // back to back the second waits for the first one's result
#pragma unroll
                        for (int x = 0; x < 4; ++x) dmma_t(c0[x][y], c1[x][y], fv[x], bw);
#pragma unroll
                        for (int x = 0; x < 4; ++x) dmma_t(c0[x][y], c1[x][y], fw[x][ks], bv);
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

static __device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one bulk copy (TMA unit): `bytes` contiguous bytes global -> shared, completion counted on `bar`
static __device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
static __device__ __forceinline__ void cp_async_mbar_arrive(void *bar)
{
    // the barrier tracks the completion of this thread's earlier cp.async copies (SASS ARRIVES.LDGSTSBAR)
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// all state spaces: generic-proxy writes to GLOBAL memory (panels, tiles) before bulk copies read them
static __device__ __forceinline__ void fence_async_proxy_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
static __device__ __forceinline__ void wg_bar(int wg) { asm volatile("bar.sync %0, 128;\n" ::"r"(1 + wg) : "memory"); }

static __device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------
// barrier-free tensor phases: per-warp tile rings + a shared operand ring on mbarriers
// ------------------------------------------------------------------------------------------
// arcb200_tile_stream_probe (csrc/micro.cu): eight independent warps per SM, each streaming its own
// tiles by cp.async one tile ahead and spending 128 DMMAs on each, run at 0.90 of the DMMA rate
// (4.2 TB/s); ending every step like the lockstep kernel does -- wait for the next tile, block
// barrier -- drops that to 0.75: a step then takes the slowest of eight tile latencies.  Here no warp
// ever waits for another warp's tile.  The only shared data of a step, the operand tile(s) V_q (or
// W_J, V_J), goes through a four-stage ring: every warp copies its four rows of the operand two steps
// ahead (cp.async, completion counted on the stage's `full` mbarrier, 256 arrivals) and releases a
// stage through its `empty` mbarrier (8 arrivals), so the warps may drift two steps apart.  Tile
// prefetch runs across row-group boundaries.  All warps walk the same (group, step) schedule.
struct S1Ring {
    double(*ring)[2][32][BSTR];    // [stage][op][row][col]
    unsigned long long *full, *empty;   // [4] each
};

static __device__ __forceinline__ void s1_symm_ring(const S1Ctx &c, const S1Ring &r, int p, unsigned &gu, int gfirst = 0,
                                                    int gcount = -1)
{
    constexpr int TSTR = S1_TSTR;
    const int lane = c.lane, grp = c.grp, tig = c.tig, wid = c.wid, np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *Tw = c.Tw;
    const int Jt = p + 1, m = NB - (p + 1);
    const int ng = gcount < 0 ? ((m + 7) >> 3) - gfirst : gcount;     // this CTA's row groups: gfirst .. gfirst + ng - 1
    const int S = ng * m;                       // steps: step u <-> (group gfirst + u / m, q = Jt + u % m)
    const unsigned gu0 = gu;
    // step u <-> (group gfirst + u / m, column block Jt + u % m); one group per CTA (the per-phase kernels): no division
    const bool one = ng == 1;
    auto ugrp = [&](int u) { return one ? 0 : u / m; };
    auto ucol = [&](int u) { return one ? u : u % m; };
    __syncthreads();
    auto fill = [&](int u) {                    // my four rows of V_q of step u (commits one cp.async group)
        if (u < S) {
            const unsigned g = gu0 + (unsigned)u;
            const int st = g & 3;
            if (g >= 4) mbar_wait(&r.empty[st], ((g >> 2) - 1) & 1u);
            const int q = Jt + ucol(u);
            const int row = 4 * wid + (lane >> 4), cc = (lane & 15) * 2;
            cp_async16_t(&r.ring[st][0][row][cc], Vp + (size_t)row * np + 32 * q + cc);
            cp_async16_t(&r.ring[st][0][row + 2][cc], Vp + (size_t)(row + 2) * np + 32 * q + cc);
            cp_async_mbar_arrive(&r.full[st]);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    auto tile = [&](int u) {                    // my tile of step u into buffer u & 1 (commits one group)
        if (u < S) {
            const int i = Jt + 8 * (gfirst + ugrp(u)) + wid, q = Jt + ucol(u);
            if (i < NB) {
                const int Ib = max(i, q), Jb = min(i, q);
                // four running source pointers: a cp.async holds its address registers until the LSU has
                // read them (17 cycles), so one pointer bumped 16 times serialises the 16 copies (640 cycles)
                const size_t sstep = 2 * (size_t)lda;
                const double *s0 = A + (size_t)(32 * Jb + (lane >> 4)) * lda + 32 * Ib + (lane & 15) * 2;
                const double *s1 = s0 + sstep, *s2 = s1 + sstep, *s3 = s2 + sstep;
                double *dst = Tw + (u & 1) * (32 * TSTR) + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    cp_async16_t(dst + (8 * k) * TSTR, s0);
                    cp_async16_t(dst + (8 * k + 2) * TSTR, s1);
                    cp_async16_t(dst + (8 * k + 4) * TSTR, s2);
                    cp_async16_t(dst + (8 * k + 6) * TSTR, s3);
                    s0 += 4 * sstep; s1 += 4 * sstep; s2 += 4 * sstep; s3 += 4 * sstep;
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    fill(0);
    fill(1);
    tile(0);
    double acc[4][4][2];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
    for (int u = 0; u < S; ++u) {
        fill(u + 2);
        tile(u + 1);
        asm volatile("cp.async.wait_group 2;\n" ::: "memory");      // my tile of step u has landed
        __syncwarp();
        const unsigned g = gu0 + (unsigned)u;
        const int st = g & 3;
        mbar_wait(&r.full[st], (g >> 2) & 1u);
        const int i = Jt + 8 * (gfirst + ugrp(u)) + wid, q = Jt + ucol(u);
        if (i < NB) {
            const double *tw = Tw + (u & 1) * (32 * TSTR);
            if (q <= i) {
                const double *pd = tw + tig * TSTR + grp;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    double fa[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) fa[x] = pd[4 * ks * TSTR + 8 * x];
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const double b = r.ring[st][0][8 * y + grp][4 * ks + tig];
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
                        const double b = r.ring[st][0][8 * y + grp][4 * ks + tig];
#pragma unroll
                        for (int x = 0; x < 4; ++x) dmma_t(acc[x][y][0], acc[x][y][1], fa[x], b);
                    }
                }
            }
            if (q == NB - 1) {                  // last step of the row: X_i
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            Wp[(size_t)(8 * y + 2 * tig + h) * np + 32 * i + 8 * x + grp] = acc[x][y][h];
                            acc[x][y][h] = 0.0;
                        }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&r.empty[st]);
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    gu = gu0 + (unsigned)S;
    __syncthreads();
}

static __device__ __forceinline__ void s1_syr2k_ring(const S1Ctx &c, const S1Ring &r, int p, unsigned &gu, int gfirst = 0,
                                                     int gcount = -1)
{
    constexpr int TSTR = S1_TSTR;
    const int lane = c.lane, grp = c.grp, tig = c.tig, wid = c.wid, np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *Tw = c.Tw;
    double *Tv = Tw + 32 * TSTR;
    const int Jt = p + 1, m = NB - (p + 1);
    const int ng = gfirst + (gcount < 0 ? ((m + 7) >> 3) - gfirst : gcount);    // this CTA's row groups: gfirst .. ng - 1
    const unsigned gu0 = gu;
    // group t: block rows Jt + 8 t - pad + wid (those >= Jt), steps J = Jt .. Jt + 8 t + 7 - pad.  The partial group is
    // the FIRST one (the shortest rows): at the end it would cost a full-length pass with most warps idle.
    const int pad = (8 - (m & 7)) & 7;
    auto group_last = [&](int t) { return Jt + 8 * t + 7 - pad; };
    __syncthreads();
    int ft = gfirst, fJ = Jt;                   // schedule position of the next operand fill
    unsigned fg = gu0;
    auto fill = [&]() {                         // my four rows of W_J and V_J (one cp.async group)
        if (ft < ng) {
            const int st = fg & 3;
            if (fg >= 4) mbar_wait(&r.empty[st], ((fg >> 2) - 1) & 1u);
            const int row = 4 * wid + (lane >> 4), cc = (lane & 15) * 2;
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2) {
                const int k = row + 2 * t2;
                cp_async16_t(&r.ring[st][0][k][cc], Wp + (size_t)k * np + 32 * fJ + cc);
                cp_async16_t(&r.ring[st][1][k][cc], Vp + (size_t)k * np + 32 * fJ + cc);
            }
            cp_async_mbar_arrive(&r.full[st]);
            ++fg;
            if (++fJ > group_last(ft)) { ++ft; fJ = Jt; }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    auto tile_to = [&](double *dst0, const double *src0, size_t pitch) {
        // eight running pointers, each used twice: a cp.async holds its address registers until the LSU has read
        // them, and that queue is long in this phase (16 tile stores per thread and step ahead of the copies)
        const size_t sstep = 2 * pitch;
        const double *s[8];
        s[0] = src0 + (size_t)(lane >> 4) * pitch + (lane & 15) * 2;
#pragma unroll
        for (int k = 1; k < 8; ++k) s[k] = s[k - 1] + sstep;
        double *dst = dst0 + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll
        for (int k = 0; k < 8; ++k) cp_async16_t(dst + (2 * k) * TSTR, s[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) cp_async16_t(dst + (16 + 2 * k) * TSTR, s[k] + 8 * sstep);
    };
    fill();
    fill();
    unsigned g = gu0;
    const double *tv = Tv + tig * TSTR + 4 * grp;
    const double *tc = Tw + (2 * tig) * TSTR + 4 * grp;
    for (int t = gfirst; t < ng; ++t) {
        const int I = Jt + 8 * t + wid - pad;
        const bool act = I >= Jt;
        const int Jlast = group_last(t);
        double fw[4][8];
        if (act) {
            // row operands: -W_I to registers, -V_I to the second tile; the first C tile (unless the
            // previous group already prefetched it)
            tile_to(Tv, Vp + 32 * I, np);
            if (t == gfirst || I - 8 < Jt) tile_to(Tw, A + (size_t)(32 * Jt) * lda + 32 * I, lda);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const double2 *pw2 = reinterpret_cast<const double2 *>(Wp + (size_t)(4 * ks + tig) * np + 32 * I + 4 * grp);
                const double2 w0 = pw2[0], w1 = pw2[1];
                fw[0][ks] = -w0.x; fw[1][ks] = -w0.y; fw[2][ks] = -w1.x; fw[3][ks] = -w1.y;
            }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();
        if (act) {
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) Tv[lane * TSTR + rr] = -Tv[lane * TSTR + rr];
        }
        __syncwarp();
        for (int J = Jt; J <= Jlast; ++J, ++g) {
            const int st = g & 3;
            const bool work = act && J <= I;
            fill();                                                    // operands two steps ahead
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");    // everything but that fill: my C tile
            __syncwarp();
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
                // the tile buffer is refilled below: its loads have to have landed
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    asm volatile("" ::"d"(c0[0][y]), "d"(c0[1][y]), "d"(c0[2][y]), "d"(c0[3][y]), "d"(c1[0][y]), "d"(c1[1][y]),
                                 "d"(c1[2][y]), "d"(c1[3][y]));
            }
            __syncwarp();
            // next C tile of this row, or -- on its last step -- the first one of my row in the next group
            if (act && J + 1 <= I) {
                tile_to(Tw, A + (size_t)(32 * (J + 1)) * lda + 32 * I, lda);
            } else if (act && J == I && t + 1 < ng) {
                tile_to(Tw, A + (size_t)(32 * Jt) * lda + 32 * (I + 8), lda);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            mbar_wait(&r.full[st], (g >> 2) & 1u);
            if (work) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const double2 va = *reinterpret_cast<const double2 *>(tv + 4 * ks * TSTR);
                    const double2 vb = *reinterpret_cast<const double2 *>(tv + 4 * ks * TSTR + 2);
                    const double fv[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const double bw = r.ring[st][0][4 * ks + tig][8 * y + grp];
                        const double bv = r.ring[st][1][4 * ks + tig][8 * y + grp];
                        // the two products of an accumulator four DMMAs apart (asm volatile keeps this order): issued
                        // back to back the second waits for the first one's result
#pragma unroll
                        for (int x = 0; x < 4; ++x) dmma_t(c0[x][y], c1[x][y], fv[x], bw);
#pragma unroll
                        for (int x = 0; x < 4; ++x) dmma_t(c0[x][y], c1[x][y], fw[x][ks], bv);
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&r.empty[st]);
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    gu = g;
    __syncthreads();
}

}  // namespace ts
