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
//                  column, so a reflector is 64 FMAs per lane with no cross-lane traffic; a 39-row
//                  register window slides over 32 consecutive sweeps; the 32 reflectors of the
//                  next block are staged in shared memory while the current block is applied.
//   ormtr_t_kernel Z <- Q1 Z on the same row-major window (compact-WY panels, DMMA): the chunk of
//                  V is staged once per CTA and step and shared by all warps, each warp streams
//                  its 32 columns of Z through a transit tile.
// For small batches (large N) the band reduction runs as one launch per phase and panel with many
// CTAs per matrix (eig_sy2sb_multi.cu, same phase bodies).
// Index conventions and application orders are modelled in tools/proto_two_stage.py.
#include "eig_sy2sb_phases.cuh"

namespace {
using namespace ts;

// ------------------------------------------------------------------------------------------
// decoupled warpgroups + TMA-unit bulk copies (the default band-reduction kernel)
// ------------------------------------------------------------------------------------------
// ncu of the lockstep kernel (profiles/r02_sy2sb_lockstep_*): inside the k loops the FP64 tensor
// pipe is saturated (one DMMA per 16 cycles and scheduler, the two warps of a scheduler alternate
// in groups of four), but every step of the symmetric product / rank-2k update ends in a block
// barrier, so the per-step prologue and epilogue (operand waits, fragment loads of the C tile,
// 128-bit stores, next-tile issue) of all eight warps coincide and the pipe idles: 69 % / 59 % of
// the DMMA rate inside those phases.  Here the eight warps form two warpgroups of four that take
// row groups of four block rows from a dispenser and run them independently (named barriers,
// private operand tiles): warps w and w + 4 share a scheduler, so one warpgroup's prologue and
// epilogue overlap the other's DMMA loop, and the triangular tail of a row group is four steps
// instead of eight.  All tile movement of these phases goes through the TMA unit: one
// cp.async.bulk per lane moves a 256-byte tile column (SASS UBLKCP), completion is tracked by
// mbarriers (SYNCS) -- no per-thread cp.async address loops, no cp.async.wait_group.
struct S1Wg {                      // warpgroup-private state of the decoupled phases
    double(*Bs)[2][32][BSTR];      // this warpgroup's operand tiles [buf][op][row][col]
    unsigned long long *tbar;      // this warp's two tile barriers
    unsigned long long *bbar;      // this warpgroup's two operand barriers
    int *next;                     // row-group dispenser of the CTA
    volatile int *cur;             // row group taken by this warpgroup
    int wg, wl;                    // warpgroup, warp inside it
};

// X = A22 V.  Warp <-> block row i of a group of four; the stored tile (i, q) (transposed use when
// q > i) and the shared operand V_q arrive by bulk copies one step ahead.
static __device__ __forceinline__ void s1_symm_wg(const S1Ctx &c, const S1Wg &g, int p, unsigned &tph, unsigned &bph)
{
    constexpr int TSTR = S1_TSTR;
    const int lane = c.lane, grp = c.grp, tig = c.tig, np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *Tw = c.Tw;
    const int Jt = p + 1, nch = NB - (p + 1);
    const int ng = (nch + 3) >> 2;
    if (c.tid == 0) *g.next = 0;
    __syncthreads();
    fence_async_proxy_all();       // generic-proxy writes of the earlier phases before the bulk copies
    auto issue_tile = [&](int Ib, int Jb, int tb) {      // lane <-> tile column (256 contiguous bytes)
        if (lane == 0) mbar_expect_tx(&g.tbar[tb], 32 * 32 * 8);
        __syncwarp();
        bulk_g2s(Tw + tb * (32 * TSTR) + lane * TSTR, A + (size_t)(32 * Jb + lane) * lda + 32 * Ib, 256, &g.tbar[tb]);
    };
    auto fillV = [&](int buf, int q) {                   // V_q as [n][k]: row n = panel column n
        if (g.wl == 0) {
            if (lane == 0) mbar_expect_tx(&g.bbar[buf], 32 * 32 * 8);
            __syncwarp();
            bulk_g2s(&g.Bs[buf][0][lane][0], Vp + (size_t)lane * np + 32 * q, 256, &g.bbar[buf]);
        }
    };
    for (;;) {
        if (g.wl == 0 && lane == 0) *g.cur = atomicAdd(g.next, 1);
        wg_bar(g.wg);
        const int t = *g.cur;
        if (t >= ng) break;
        const int i = Jt + 4 * t + g.wl;
        const bool act = i < NB;
        double acc[4][4][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
        if (act) issue_tile(max(i, Jt), min(i, Jt), 0);
        fillV(0, Jt);
        for (int q = Jt; q < NB; ++q) {
            const int buf = (q - Jt) & 1;
            if (q + 1 < NB) {
                if (act) issue_tile(max(i, q + 1), min(i, q + 1), buf ^ 1);
                fillV(buf ^ 1, q + 1);
            }
            mbar_wait(&g.bbar[buf], (bph >> buf) & 1u);
            bph ^= 1u << buf;
            if (act) {
                mbar_wait(&g.tbar[buf], (tph >> buf) & 1u);
                tph ^= 1u << buf;
                // fragments straight from the transit tile inside the k loop; diagonal tiles are
                // kept fully symmetric, so q == i reads like a stored tile
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
                            const double b = g.Bs[buf][0][8 * y + grp][4 * ks + tig];
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
                            const double b = g.Bs[buf][0][8 * y + grp][4 * ks + tig];
#pragma unroll
                            for (int x = 0; x < 4; ++x) dmma_t(acc[x][y][0], acc[x][y][1], fa[x], b);
                        }
                    }
                }
            }
            wg_bar(g.wg);          // the four warps are done with Bs[buf]
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

// A22 -= V W^T + W V^T.  Warp <-> block row I of a group of four (largest groups first): -W_I
// stays in registers, -V_I in the warp's second transit tile, W_J / V_J ([k][n]: row k = panel
// column) are shared by the warpgroup, the C tile arrives one step ahead.  DMMA row (x, grp) <->
// tile row 4 grp + x, so a lane's four rows are contiguous (128-bit shared loads, global stores).
static __device__ __forceinline__ void s1_syr2k_wg(const S1Ctx &c, const S1Wg &g, int p, unsigned &tph, unsigned &bph,
                                                   int Jcap = 1 << 30)
{
    constexpr int TSTR = S1_TSTR;
    const int lane = c.lane, grp = c.grp, tig = c.tig, np = c.np, lda = c.lda, NB = c.NB;
    double *A = c.A;
    double *__restrict__ Wp = c.Wp;
    double *__restrict__ Vp = c.Vp;
    double *Tw = c.Tw;
    double *Tv = Tw + 32 * TSTR;
    const int Jt = p + 1, nch = NB - (p + 1);
    const int ng = (nch + 3) >> 2;
    if (c.tid == 0) *g.next = 0;
    __syncthreads();
    fence_async_proxy_all();
    auto issue_tile = [&](int Ib, int Jb) {
        if (lane == 0) mbar_expect_tx(&g.tbar[0], 32 * 32 * 8);
        __syncwarp();
        bulk_g2s(Tw + lane * TSTR, A + (size_t)(32 * Jb + lane) * lda + 32 * Ib, 256, &g.tbar[0]);
    };
    auto fillWV = [&](int buf, int J) {
        if (g.wl == 0) {
            if (lane == 0) mbar_expect_tx(&g.bbar[buf], 2 * 32 * 32 * 8);
            __syncwarp();
            bulk_g2s(&g.Bs[buf][0][lane][0], Wp + (size_t)lane * np + 32 * J, 256, &g.bbar[buf]);
            bulk_g2s(&g.Bs[buf][1][lane][0], Vp + (size_t)lane * np + 32 * J, 256, &g.bbar[buf]);
        }
    };
    for (;;) {
        if (g.wl == 0 && lane == 0) *g.cur = atomicAdd(g.next, 1);
        wg_bar(g.wg);
        const int tt_ = *g.cur;
        if (tt_ >= ng) break;
        const int g0 = Jt + 4 * (ng - 1 - tt_);
        const int I = g0 + g.wl;
        const bool act = I < NB;
        const int Jmax = min(min(g0 + 3, NB - 1), Jcap);      // Jcap = Jt: first block column only
        if (act) {
            issue_tile(I, Jt);
            // V_I: panel columns k = 0..31 (lane <-> k), rows 32 I .. 32 I + 31
            if (lane == 0) mbar_expect_tx(&g.tbar[1], 32 * 32 * 8);
            __syncwarp();
            bulk_g2s(Tv + lane * TSTR, Vp + (size_t)lane * np + 32 * I, 256, &g.tbar[1]);
        }
        fillWV(0, Jt);
        double fw[4][8];
        if (act) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const double2 *pw2 = reinterpret_cast<const double2 *>(Wp + (size_t)(4 * ks + tig) * np + 32 * I + 4 * grp);
                const double2 w0 = pw2[0], w1 = pw2[1];
                fw[0][ks] = -w0.x; fw[1][ks] = -w0.y; fw[2][ks] = -w1.x; fw[3][ks] = -w1.y;
            }
            mbar_wait(&g.tbar[1], (tph >> 1) & 1u);
            tph ^= 2u;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) Tv[lane * TSTR + r] = -Tv[lane * TSTR + r];
            __syncwarp();
            fence_async_proxy();   // these generic writes precede the next group's bulk copy into Tv
        }
        const double *tv = Tv + tig * TSTR + 4 * grp;
        const double *tc = Tw + (2 * tig) * TSTR + 4 * grp;
        for (int J = Jt; J <= Jmax; ++J) {
            const int buf = (J - Jt) & 1;
            const bool work = act && J <= I;
            mbar_wait(&g.bbar[buf], (bph >> buf) & 1u);
            bph ^= 1u << buf;
            double c0[4][4], c1[4][4];      // [x][y]: rows 4 grp + x, columns 8 y + 2 tig (+1)
            if (work) {
                mbar_wait(&g.tbar[0], tph & 1u);
                tph ^= 1u;
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
            if (J + 1 <= Jmax) {
                if (act && J + 1 <= I) issue_tile(I, J + 1);
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
                        const double bw = g.Bs[buf][0][4 * ks + tig][8 * y + grp];
                        const double bv = g.Bs[buf][1][4 * ks + tig][8 * y + grp];
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
            wg_bar(g.wg);
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// look-ahead band reduction: the rank-2k update of panel p fused with the symmetric product of
// panel p + 1
// ------------------------------------------------------------------------------------------
// The two tensor phases above are bound by HBM latency, not by the tensor pipe: every warp has one
// 8 KB tile in flight, and at the DMMA rate they would move 4.6 TB/s (each stored tile is read twice by
// the symmetric product and read + written by the update: 32 KB per 512 DMMAs).  Once block column
// p + 1 of the trailing matrix has been updated, panel p + 1 can be factored BEFORE the rest of the
// update of panel p, and every remaining tile then makes a single round trip:
//     C' = C - V_I W_J^T - W_I V_J^T      (256 DMMAs)   -> stored by a TMA bulk copy from shared memory
//     X'_I += C' V'_J                     (128 DMMAs)   registers of the row's warp
//     X'_J += C'^T V'_I   (J < I)         (128 DMMAs)   -> red.global.add.f64 into the X' panel
// 16 KB per 512 DMMAs.  Four compute warps (one per scheduler: a warp alone sustains the pipe's one DMMA
// per 16 cycles) own one block row each -- -W_I in registers, -V_I, V'_I and a two-tile C ring in shared
// memory; a producer warp streams the shared operands W_J, V_J, V'_J through a two-stage ring guarded
// by full / empty mbarriers, so the compute warps never meet at a block barrier inside the pass.


// Work split of the fused pass: compute warp w (0..3, one per scheduler) does nothing but wait for
// its operands, issue DMMAs and shared-memory fragment traffic; helper warp 4 + w does the memory
// side of the same block row: cp.async of the next C tiles (two ahead) and of the row operands,
// the 128-bit global stores of the updated tiles (read back from shared memory), and a quarter of
// the shared W_J / V_J / V'_J ring.  (A first version let the compute warps issue one
// cp.async.bulk per tile column themselves: UBLKCP takes its addresses from uniform registers,
// so the 32 per-lane copies of a tile are serialised in an ELECT / R2UR loop -- ~1500 cycles of
// the only warp that feeds the tensor pipe of its scheduler, profiles/r02_sy2sb_fused_v1_*.)
struct S1Fz {
    double *Vc, *Wc;               // panel p: reflectors and W (np x 32, column-major)
    double *Vn;                    // panel p + 1 reflectors
    double *Xacc;                  // X' accumulation panel (zeroed by the caller)
    double(*Bs3)[3][32][BSTR];     // [buf][W_J | V_J | V'_J][row][col]
    double *tiles;                 // the block row's four tiles: C ring (2), -V_I, V'_I
    unsigned long long *rb;        // the block row's barriers: cfull[2], cready[2], cdone[2], vfull, vfree
    unsigned long long *full, *empty;   // [2] each: operand ring
};
enum { RB_CFULL = 0, RB_CREADY = 2, RB_CDONE = 4, RB_VFULL = 6, RB_VFREE = 7 };

struct S1FzPhase {                 // phase parities, persistent over the passes (one bit per barrier)
    unsigned rb;                   // bits RB_*
    unsigned ring;                 // bit buf: full (compute warps) / empty (helpers)
};

static __device__ __forceinline__ void s1_fused_pass(const S1Ctx &c, const S1Fz &f, int p, S1FzPhase &ph)
{
    constexpr int TSTR = S1_TSTR;
    const int lane = c.lane, grp = c.grp, tig = c.tig, np = c.np, lda = c.lda, NB = c.NB;
    const int w = c.wid & 3;
    const bool helper = c.wid >= 4;
    double *A = c.A;
    const int J0 = p + 2;                  // first block row / column of the pass
    const int ng = (NB - J0 + 3) >> 2;
    double *Tc0 = f.tiles, *Tv = f.tiles + 2 * (32 * TSTR), *Tn = f.tiles + 3 * (32 * TSTR);
    auto rbwait = [&](int which) {
        mbar_wait(&f.rb[which], (ph.rb >> which) & 1u);
        ph.rb ^= 1u << which;
    };
    // the schedule both warps of a block row walk through: groups of four block rows, largest first
    auto group_g0 = [&](int t) { return J0 + 4 * (ng - 1 - t); };
    auto row_of = [&](int t) { return group_g0(t) + ((t & 1) ? 3 - w : w); };    // alternate long and short rows
    __syncthreads();
    if (helper) {
        // ------------------------------------------------------------------ memory side of block row w
        auto load_tile = [&](double *dst, const double *src, size_t pitch) {     // 32 x 32 tile, columns `pitch` apart
            const double *s0 = src + (size_t)(lane >> 4) * pitch + (lane & 15) * 2;
            double *d0 = dst + (lane >> 4) * TSTR + (lane & 15) * 2;
#pragma unroll 4
            for (int k = 0; k < 16; ++k) {
                cp_async16_t(d0, s0);
                s0 += 2 * pitch;
                d0 += 2 * TSTR;
            }
        };
        auto fill_ring = [&](int buf, int J) {           // this helper's quarter: rows 8 w .. 8 w + 7 of the three tiles
            const int k0 = 8 * w + (lane >> 4), cc = (lane & 15) * 2;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + 2 * q;
                cp_async16_t(&f.Bs3[buf][0][k][cc], f.Wc + (size_t)k * np + 32 * J + cc);
                cp_async16_t(&f.Bs3[buf][1][k][cc], f.Vc + (size_t)k * np + 32 * J + cc);
                cp_async16_t(&f.Bs3[buf][2][k][cc], f.Vn + (size_t)k * np + 32 * J + cc);
            }
            cp_async_mbar_arrive(&f.full[buf]);
        };
        // look-ahead iterator over the steps of the pass (for the ring fills two steps ahead)
        int ft = 0, fJ = J0;                              // next step to fill
        auto fill_next = [&](int buf) {
            if (ft < ng) {
                fill_ring(buf, fJ);
                if (++fJ > min(group_g0(ft) + 3, NB - 1)) { ++ft; fJ = J0; }
            }
        };
        fill_next(0);
        fill_next(1);
        int s = 0;
        for (int t = 0; t < ng; ++t) {
            const int I = row_of(t);
            const bool act = I < NB;
            const int Jlast = min(group_g0(t) + 3, NB - 1);
            if (act) {
                load_tile(Tv, f.Vc + 32 * I, np);         // V_I  as [k][row]
                load_tile(Tn, f.Vn + 32 * I, np);         // V'_I as [n][k]
                cp_async_mbar_arrive(&f.rb[RB_VFULL]);
                load_tile(Tc0, A + (size_t)(32 * J0) * lda + 32 * I, lda);
                cp_async_mbar_arrive(&f.rb[RB_CFULL + 0]);
                if (J0 + 1 <= I) {
                    load_tile(Tc0 + 32 * TSTR, A + (size_t)(32 * (J0 + 1)) * lda + 32 * I, lda);
                    cp_async_mbar_arrive(&f.rb[RB_CFULL + 1]);
                }
            }
            for (int J = J0; J <= Jlast; ++J, ++s) {
                if (act && J <= I) {
                    const int cb = (J - J0) & 1;
                    double *tile = Tc0 + cb * (32 * TSTR);
                    // the updated tile: shared memory -> global, 256-byte column segments
                    rbwait(RB_CREADY + cb);
                    {
                        const double *s0 = tile + (lane >> 4) * TSTR + (lane & 15) * 2;
                        double *d0 = A + (size_t)(32 * J + (lane >> 4)) * lda + 32 * I + (lane & 15) * 2;
#pragma unroll 4
                        for (int k = 0; k < 16; ++k) {
                            *reinterpret_cast<double2 *>(d0) = *reinterpret_cast<const double2 *>(s0);
                            s0 += 2 * TSTR;
                            d0 += 2 * (size_t)lda;
                        }
                    }
                    // the compute warp is done with the tile: refill its buffer two tiles ahead
                    rbwait(RB_CDONE + cb);
                    if (J + 2 <= I) {
                        load_tile(tile, A + (size_t)(32 * (J + 2)) * lda + 32 * I, lda);
                        cp_async_mbar_arrive(&f.rb[RB_CFULL + cb]);
                    }
                }
                // operand ring: all four compute warps are done with step s -> its buffer gets step s + 2
                const int buf = s & 1;
                mbar_wait(&f.empty[buf], (ph.ring >> buf) & 1u);
                ph.ring ^= 1u << buf;
                fill_next(buf);
            }
            if (act) rbwait(RB_VFREE);                    // row operands may be overwritten
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
        // ------------------------------------------------------------------ compute warp w
        int s = 0;
        for (int t = 0; t < ng; ++t) {
            const int I = row_of(t);
            const bool act = I < NB;
            const int Jlast = min(group_g0(t) + 3, NB - 1);
            double fw[4][8];
            double xa[4][4][2];
            if (act) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const double2 *pw2 = reinterpret_cast<const double2 *>(f.Wc + (size_t)(4 * ks + tig) * np + 32 * I + 4 * grp);
                    const double2 w0 = pw2[0], w1 = pw2[1];
                    fw[0][ks] = -w0.x; fw[1][ks] = -w0.y; fw[2][ks] = -w1.x; fw[3][ks] = -w1.y;
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) xa[x][y][0] = xa[x][y][1] = 0.0;
                rbwait(RB_VFULL);
#pragma unroll 4
                for (int r = 0; r < 32; ++r) Tv[lane * TSTR + r] = -Tv[lane * TSTR + r];
                __syncwarp();
            }
            const double *tv = Tv + tig * TSTR + 4 * grp;
            for (int J = J0; J <= Jlast; ++J, ++s) {
                const int buf = s & 1;
                mbar_wait(&f.full[buf], (ph.ring >> buf) & 1u);
                ph.ring ^= 1u << buf;
                if (act && J <= I) {
                    const int cb = (J - J0) & 1;
                    double *tile = Tc0 + cb * (32 * TSTR);
                    rbwait(RB_CFULL + cb);
                    double c0[4][4], c1[4][4];      // [x][y]: rows 4 grp + x, columns 8 y + 2 tig (+1)
                    double *tc = tile + (2 * tig) * TSTR + 4 * grp;
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const double2 *q0 = reinterpret_cast<const double2 *>(tc + 8 * y * TSTR);
                        const double2 *q1 = reinterpret_cast<const double2 *>(tc + (8 * y + 1) * TSTR);
                        const double2 a0 = q0[0], a1 = q0[1], b0 = q1[0], b1 = q1[1];
                        c0[0][y] = a0.x; c0[1][y] = a0.y; c0[2][y] = a1.x; c0[3][y] = a1.y;
                        c1[0][y] = b0.x; c1[1][y] = b0.y; c1[2][y] = b1.x; c1[3][y] = b1.y;
                    }
                    // ---- C' = C - V_I W_J^T - W_I V_J^T
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const double2 va = *reinterpret_cast<const double2 *>(tv + 4 * ks * TSTR);
                        const double2 vb = *reinterpret_cast<const double2 *>(tv + 4 * ks * TSTR + 2);
                        const double fv[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            const double bw = f.Bs3[buf][0][4 * ks + tig][8 * y + grp];
                            const double bv = f.Bs3[buf][1][4 * ks + tig][8 * y + grp];
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                dmma_t(c0[x][y], c1[x][y], fv[x], bw);
                                dmma_t(c0[x][y], c1[x][y], fw[x][ks], bv);
                            }
                        }
                    }
                    // ---- C' back into the tile: the helper stores it, the products below read it
                    __syncwarp();
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        double2 *o0 = reinterpret_cast<double2 *>(tc + 8 * y * TSTR);
                        double2 *o1 = reinterpret_cast<double2 *>(tc + (8 * y + 1) * TSTR);
                        o0[0] = make_double2(c0[0][y], c0[1][y]);
                        o0[1] = make_double2(c0[2][y], c0[3][y]);
                        o1[0] = make_double2(c1[0][y], c1[1][y]);
                        o1[1] = make_double2(c1[2][y], c1[3][y]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&f.rb[RB_CREADY + cb]);
                    // ---- X'_I += C' V'_J
                    {
                        const double *pd = tile + tig * TSTR + grp;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            double fa[4];
#pragma unroll
                            for (int x = 0; x < 4; ++x) fa[x] = pd[4 * ks * TSTR + 8 * x];
#pragma unroll
                            for (int y = 0; y < 4; ++y) {
                                const double b = f.Bs3[buf][2][8 * y + grp][4 * ks + tig];
#pragma unroll
                                for (int x = 0; x < 4; ++x) dmma_t(xa[x][y][0], xa[x][y][1], fa[x], b);
                            }
                        }
                    }
                    // ---- X'_J += C'^T V'_I (strictly lower tiles)
                    if (J < I) {
                        double ya[4][4][2];
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y) ya[x][y][0] = ya[x][y][1] = 0.0;
                        const double *pt = tile + grp * TSTR + tig;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            double fa[4];
#pragma unroll
                            for (int x = 0; x < 4; ++x) fa[x] = pt[8 * x * TSTR + 4 * ks];
#pragma unroll
                            for (int y = 0; y < 4; ++y) {
                                const double b = Tn[(8 * y + grp) * TSTR + 4 * ks + tig];
#pragma unroll
                                for (int x = 0; x < 4; ++x) dmma_t(ya[x][y][0], ya[x][y][1], fa[x], b);
                            }
                        }
                        double *xr = f.Xacc + (size_t)(2 * tig) * np + 32 * J + grp;
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y)
#pragma unroll
                                for (int h = 0; h < 2; ++h) atomicAdd(xr + (size_t)(8 * y + h) * np + 8 * x, ya[x][y][h]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&f.rb[RB_CDONE + cb]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&f.empty[buf]);
            }
            if (act) {
                double *xr = f.Xacc + (size_t)(2 * tig) * np + 32 * I + grp;
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
#pragma unroll
                        for (int h = 0; h < 2; ++h) atomicAdd(xr + (size_t)(8 * y + h) * np + 8 * x, xa[x][y][h]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&f.rb[RB_VFREE]);
            }
        }
    }
    __threadfence();               // tile stores and reductions are performed before the block reads them
    __syncthreads();
}

static __device__ __forceinline__ void s1_band_copy(const TsArgs &a, double *slot, double *A, int lda, int tid, int NT)
{
    const int np = a.Npad;
    __syncthreads();
    {
        double *__restrict__ Bd = slot + a.oBand;
        const int N = a.N;
        const int total = EIG_LDB * (np + 64);
        for (int idx = tid; idx < total; idx += NT) {
            const int d = idx % EIG_LDB, j = idx / EIG_LDB;
            double x = 0.0;
            if (d <= 32 && j + d < N) x = A[(j + d) + (size_t)j * lda];
            Bd[idx] = x;
        }
    }
}

// WARPS = 8: one CTA per SM; WARPS = 4: one 4-warp CTA per matrix (kept for A/B runs).  All warps in
// lockstep (one block barrier per tile step), tiles by per-thread cp.async.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) sy2sb_kernel(TsArgs a)
{
    constexpr int TSTR = S1_TSTR;
    constexpr int UNI = 2 * WARPS * 32 * TSTR + 4608;   // 2 transit tiles per warp + operand tiles (doubles)
    const int mat = blockIdx.x;
    double *slot = a.base + (size_t)mat * a.slot;
    extern __shared__ double sm[];
    S1Ctx c;
    c.tid = threadIdx.x; c.wid = c.tid >> 5; c.lane = c.tid & 31; c.grp = c.lane >> 2; c.tig = c.lane & 3;
    c.np = a.Npad; c.lda = a.Npad; c.NB = a.NB;
    c.A = slot + a.oA; c.Wp = slot + a.oWp; c.Vp = slot + a.oVp; c.tt = slot + a.oTau; c.Tfac = slot + a.oTfac;
    c.Bs = reinterpret_cast<double(*)[2][32][BSTR]>(sm + 2 * WARPS * 32 * TSTR);
    c.Ts = reinterpret_cast<double(*)[33]>(sm + UNI);
    c.G = reinterpret_cast<double(*)[33]>(sm + UNI + 1056);
    c.M1 = reinterpret_cast<double(*)[33]>(sm + UNI + 2 * 1056);
    c.wv = reinterpret_cast<double(*)[32]>(sm + UNI + 3 * 1056);            // [par][j]
    c.sred = reinterpret_cast<double(*)[16]>(sm + UNI + 3 * 1056 + 64);     // [par][warp], [par][15] = alpha
    c.taus = sm + UNI + 3 * 1056 + 96;                                      // [32]
    c.Tw = sm + c.wid * (2 * 32 * TSTR);
    TSP_DECL
    for (int p = 0; p < a.NB - 1; ++p) {
        TSP(0)
        s1_panel_qr<WARPS>(c, p);
        TSP(1)
        s1_tfactor<WARPS>(c, p);
        TSP(2)
        s1_symm_lockstep<WARPS>(c, p);
        TSP(3)
        s1_wphase<WARPS>(c, p);
        TSP(4)
        s1_syr2k_lockstep<WARPS>(c, p);
        TSP(5)
    }
#ifdef TS_PROFILE
    if (mat == 0 && c.tid == 0)
        printf("sy2sb prof (10 kcycles): qr %lld tfac %lld symm %lld w %lld syr2k %lld\n", tsp_t[1] / 10000,
               tsp_t[2] / 10000, tsp_t[3] / 10000, tsp_t[4] / 10000, tsp_t[5] / 10000);
#endif
    __syncthreads();
    s1_band_copy(a, slot, c.A, c.lda, c.tid, WARPS * 32);
}

// Default: two decoupled warpgroups, TMA-unit bulk copies + mbarriers (see above).
constexpr int S1WG_SMEM_DOUBLES = 2 * 8 * 32 * S1_TSTR + 2 * 4608 + 128;
constexpr size_t S1WG_SMEM_BYTES = (size_t)S1WG_SMEM_DOUBLES * sizeof(double) + 256;
__global__ void __launch_bounds__(256, 1) sy2sb_wg_kernel(TsArgs a)
{
    constexpr int WARPS = 8;
    constexpr int TSTR = S1_TSTR;
    constexpr int TILES = 2 * WARPS * 32 * TSTR;
    const int mat = blockIdx.x;
    double *slot = a.base + (size_t)mat * a.slot;
    extern __shared__ double sm[];
    S1Ctx c;
    c.tid = threadIdx.x; c.wid = c.tid >> 5; c.lane = c.tid & 31; c.grp = c.lane >> 2; c.tig = c.lane & 3;
    c.np = a.Npad; c.lda = a.Npad; c.NB = a.NB;
    c.A = slot + a.oA; c.Wp = slot + a.oWp; c.Vp = slot + a.oVp; c.tt = slot + a.oTau; c.Tfac = slot + a.oTfac;
    // [transit tiles][operand tiles of warpgroup 0][operand tiles of warpgroup 1][small arrays][barriers]
    // T, G, M1 of the T-factor / W phases live in warpgroup 1's operand tiles (idle in those phases;
    // T is reloaded from global memory after the symmetric product has used them)
    c.Bs = reinterpret_cast<double(*)[2][32][BSTR]>(sm + TILES);
    double *bsB = sm + TILES + 4608;
    c.Ts = reinterpret_cast<double(*)[33]>(bsB);
    c.G = reinterpret_cast<double(*)[33]>(bsB + 1056);
    c.M1 = reinterpret_cast<double(*)[33]>(bsB + 2 * 1056);
    double *small = sm + TILES + 2 * 4608;
    c.wv = reinterpret_cast<double(*)[32]>(small);
    c.sred = reinterpret_cast<double(*)[16]>(small + 64);
    c.taus = small + 96;
    c.Tw = sm + c.wid * (2 * 32 * TSTR);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(small + 128);   // 16 tile + 4 operand barriers
    int *ctl = reinterpret_cast<int *>(bars + 20);                                     // next, cur[2]
    S1Wg g;
    g.wg = c.wid >> 2; g.wl = c.wid & 3;
    g.Bs = g.wg ? reinterpret_cast<double(*)[2][32][BSTR]>(bsB) : c.Bs;
    g.tbar = bars + 2 * c.wid;
    g.bbar = bars + 16 + 2 * g.wg;
    g.next = ctl;
    g.cur = ctl + 1 + g.wg;
    if (c.tid == 0) {
        for (int i = 0; i < 20; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    fence_async_proxy();
    __syncthreads();
    unsigned tph = 0, bph = 0;     // phase parities of this warp's tile / this warpgroup's operand barriers
    TSP_DECL
    for (int p = 0; p < a.NB - 1; ++p) {
        TSP(0)
        s1_panel_qr<WARPS>(c, p);
        TSP(1)
        s1_tfactor<WARPS>(c, p);
        TSP(2)
        s1_symm_wg(c, g, p, tph, bph);
        TSP(3)
        for (int t = c.tid; t < 1024; t += WARPS * 32) c.Ts[t >> 5][t & 31] = c.Tfac[(size_t)p * 1024 + t];
        s1_wphase<WARPS>(c, p);
        TSP(4)
        s1_syr2k_wg(c, g, p, tph, bph);
        TSP(5)
    }
#ifdef TS_PROFILE
    if (mat == 0 && c.tid == 0)
        printf("sy2sb wg prof (10 kcycles): qr %lld tfac %lld symm %lld w %lld syr2k %lld\n", tsp_t[1] / 10000,
               tsp_t[2] / 10000, tsp_t[3] / 10000, tsp_t[4] / 10000, tsp_t[5] / 10000);
#endif
    __syncthreads();
    s1_band_copy(a, slot, c.A, c.lda, c.tid, WARPS * 32);
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
    TSP_DECL
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
        TSP(0)
        wait_for(0);
        TSP(1)
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
            if (L == 32 && L2 == 32 && tau != 0.0) {
                // ---- interior step, all blocks full: one base pointer, constant offsets, no
                // divergent regions.  Element (r + lane, r + j) sits at cp[j * CS]; the block
                // below, element (r + 32 + lane, r + j), at cp[32 + j * CS].
                constexpr int CS = EIG_LDB - 1;
                double *cp = Bd + (size_t)r * EIG_LDB + lane;
                double Dl[32], El[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) Dl[j] = (j <= lane) ? cp[j * CS] : 0.0;
#pragma unroll
                for (int j = 0; j < 32; ++j) El[j] = cp[32 + j * CS];
                vs[lane] = v;
                __syncwarp();
                const double2 *vs2 = reinterpret_cast<const double2 *>(vs);
                const double2 *ws2 = reinterpret_cast<const double2 *>(ws);
                double pd0 = 0.0, pd1 = 0.0, pd2 = 0.0, pd3 = 0.0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const double2 a = vs2[j / 2], b = vs2[j / 2 + 1];
                    pd0 += Dl[j] * a.x;
                    pd1 += Dl[j + 1] * a.y;
                    pd2 += Dl[j + 2] * b.x;
                    pd3 += Dl[j + 3] * b.y;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = Dl[j] * v;      // entries j > lane are zero
                tile[lane * 33 + lane] = 0.0;                                       // the diagonal is in pd
                __syncwarp();
                double pt0 = 0.0, pt1 = 0.0, pt2 = 0.0, pt3 = 0.0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    pt0 += tile[i * 33 + lane];
                    pt1 += tile[(i + 1) * 33 + lane];
                    pt2 += tile[(i + 2) * 33 + lane];
                    pt3 += tile[(i + 3) * 33 + lane];
                }
                __syncwarp();
                const double pv = tau * (((pd0 + pd1) + (pd2 + pd3)) + ((pt0 + pt1) + (pt2 + pt3)));
                const double al = 0.5 * tau * warp_sum(pv * v);
                const double w = pv - al * v;
                ws[lane] = w;
                __syncwarp();
                double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const double2 a = vs2[j / 2], b = ws2[j / 2];
                    const double x0 = Dl[j] - (v * b.x + w * a.x);
                    const double x1 = Dl[j + 1] - (v * b.y + w * a.y);
                    if (j <= lane) cp[j * CS] = x0;
                    if (j + 1 <= lane) cp[(j + 1) * CS] = x1;
                    if (j & 2) { y2 += El[j] * a.x; y3 += El[j + 1] * a.y; } else { y0 += El[j] * a.x; y1 += El[j + 1] * a.y; }
                }
                const double y = tau * ((y0 + y1) + (y2 + y3));
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const double2 a = vs2[j / 2];
                    El[j] -= y * a.x;
                    El[j + 1] -= y * a.y;
                }
                double v2, tau_n, beta2;
                warp_house(El[0], lane, v2, tau_n, beta2);
                El[0] = (lane == 0) ? beta2 : 0.0;
                __syncwarp();
                if (tau_n != 0.0) {
#pragma unroll
                    for (int j = 1; j < 32; ++j) tile[lane * 33 + j] = v2 * El[j];
                    __syncwarp();
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        s0 += tile[i * 33 + lane];
                        s1 += tile[(i + 1) * 33 + lane];
                        s2 += tile[(i + 2) * 33 + lane];
                        s3 += tile[(i + 3) * 33 + lane];
                    }
                    ws[lane] = ((s0 + s1) + (s2 + s3)) * tau_n;
                    __syncwarp();
#pragma unroll
                    for (int j = 2; j < 32; j += 2) {
                        const double2 b = ws2[j / 2];
                        El[j] -= v2 * b.x;
                        El[j + 1] -= v2 * b.y;
                    }
                    El[1] -= v2 * ws[1];
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) cp[32 + j * CS] = El[j];
                __syncwarp();
                r = r2; L = L2; v = v2; tau = tau_n;
                ++t;
                __threadfence_block();
                __syncwarp();
                __threadfence_block();
                if (lane == 0) prog[s & 63] = s * 4096 + t;
                wait_for(t);
                continue;
            }
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
#ifdef TS_PROFILE
            if (Dl[0] + El[0] == 1.2345e300) tsp_t[7]++;      // force the loads to complete here
#endif
            TSP(2)
            if (tau != 0.0) {
                // D <- H D H on the lower triangle
                // four partial sums per reduction: the dependent FP64 chains are the critical path
                double pd0 = 0.0, pd1 = 0.0, pd2 = 0.0, pd3 = 0.0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    pd0 += Dl[j] * vs[j];
                    pd1 += Dl[j + 1] * vs[j + 1];
                    pd2 += Dl[j + 2] * vs[j + 2];
                    pd3 += Dl[j + 3] * vs[j + 3];
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = (j < lane) ? Dl[j] * v : 0.0;
                __syncwarp();
                double pt0 = 0.0, pt1 = 0.0, pt2 = 0.0, pt3 = 0.0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    pt0 += tile[i * 33 + lane];
                    pt1 += tile[(i + 1) * 33 + lane];
                    pt2 += tile[(i + 2) * 33 + lane];
                    pt3 += tile[(i + 3) * 33 + lane];
                }
                __syncwarp();
                const double pv = tau * (((pd0 + pd1) + (pd2 + pd3)) + ((pt0 + pt1) + (pt2 + pt3)));
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
            TSP(3)
            if (L2 <= 0) break;
            // E <- E H
            if (tau != 0.0) {
                double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    y0 += El[j] * vs[j];
                    y1 += El[j + 1] * vs[j + 1];
                    y2 += El[j + 2] * vs[j + 2];
                    y3 += El[j + 3] * vs[j + 3];
                }
                const double y = tau * ((y0 + y1) + (y2 + y3));
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
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    s0 += tile[i * 33 + lane];
                    s1 += tile[(i + 1) * 33 + lane];
                    s2 += tile[(i + 2) * 33 + lane];
                    s3 += tile[(i + 3) * 33 + lane];
                }
                const double sj = (s0 + s1) + (s2 + s3);
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
            TSP(4)
            r = r2; L = L2; v = v2; tau = tau_n;
            ++t;
            __threadfence_block();
            __syncwarp();
            __threadfence_block();
            if (lane == 0) prog[s & 63] = s * 4096 + t;
            TSP(5)
            wait_for(t);
            TSP(1)
        }
        __threadfence_block();
        __syncwarp();
        __threadfence_block();
        if (lane == 0) prog[s & 63] = s * 4096 + 4095;
    }
#ifdef TS_PROFILE
    if (mat == 0 && lane == 0 && (wid == 0 || wid == 5))
        printf("sb2st prof warp %d (10 kcycles): wait %lld loads %lld D %lld E %lld publish %lld other %lld\n", wid,
               tsp_t[1] / 10000, tsp_t[2] / 10000, tsp_t[3] / 10000, tsp_t[4] / 10000, tsp_t[5] / 10000, tsp_t[0] / 10000);
#endif
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

// grid (ceil(nslabs / 8), nmat); warp <-> slab of 32 eigenvector columns, lane <-> column.
// The 8 warps of a CTA walk the same (group, step) sequence of one matrix: the 32 reflectors of
// the next block (8 KB) and their taus are prefetched into registers while the current block is
// applied from shared memory (broadcast reads), one __syncthreads per block.
// Q2_WARPS warps (slabs) per CTA (one warp per CTA was measured slower: nothing hides the
// staging latency)
template <int Q2_WARPS>
__global__ void __launch_bounds__(Q2_WARPS * 32, Q2_WARPS == 4 ? 2 : 1) q2_kernel(TsArgs a)
{
    __shared__ __align__(16) double Vs[2][32][32];
    __shared__ double taus[2][32];
    const int mat = blockIdx.y;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int slab = blockIdx.x * Q2_WARPS + wid;
    const bool act = slab * 32 < a.kp;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *__restrict__ V2 = slot + a.oV2;
    const double *__restrict__ tau2 = slot + a.oTau2;
    double *Zc = slot + a.oZt + (act ? slab * 32 + lane : 0);
    const int N = a.N, ldv = a.Npad, TM = a.NB + 2;
    const size_t kp = (size_t)a.kp;
    const int nsw = N - 2;
    const int ngroups = (nsw + 31) / 32;
    if (ngroups <= 0) return;

    auto fetch = [&](int G, int t, double(&pv)[32 / Q2_WARPS], double &pt) {
        const int s0 = 32 * G;
#pragma unroll
        for (int i = 0; i < 32 / Q2_WARPS; ++i) {
            const int s = s0 + wid + Q2_WARPS * i, row = s + 1 + 32 * t + lane;
            pv[i] = (s < nsw && row < N) ? V2[(size_t)s * ldv + row] : 0.0;
        }
        pt = 0.0;
        if (tid < 32) {
            const int s = s0 + tid;
            if (s < nsw && s + 1 + 32 * t < N) pt = tau2[(size_t)s * TM + t];
        }
    };
    auto stash = [&](int buf, const double(&pv)[32 / Q2_WARPS], double pt) {
#pragma unroll
        for (int i = 0; i < 32 / Q2_WARPS; ++i) Vs[buf][wid + Q2_WARPS * i][lane] = pv[i];
        if (tid < 32) taus[buf][tid] = pt;
    };

    // Row staging (per warp, lane-private columns).  Block (G, t) works on the 63 rows Wb .. Wb + 62,
    // Wb = 32 G + 1 + 32 t, from the bottom up; block (G, t + 1) on Wb + 32 .. Wb + 94.  Its lower 31 rows are
    // results of this block and are forwarded through L (double-buffered), its 32 fresh rows are prefetched
    // into F by cp.async while this block computes: only the first block of a group reads Z with exposed
    // latency (direct loads), and every row is read from HBM once per group step instead of twice.
    extern __shared__ double q2sm[];
    double *F = q2sm + wid * (1024 + 2 * 992) + lane;
    double *L0 = F + 1024;
    int lb = 0;
    int G = ngroups - 1, t = 0, buf = 0;
    double pv[32 / Q2_WARPS], pt;
    fetch(G, t, pv, pt);
    stash(0, pv, pt);
    __syncthreads();
    while (true) {
        int Gn = G, tn = t + 1;
        if (32 * Gn + 1 + 32 * tn >= N) { Gn = G - 1; tn = 0; }
        const bool more = Gn >= 0;
        if (more) fetch(Gn, tn, pv, pt);
        if (act) {
            const int s0 = 32 * G;
            const bool first = t == 0;                       // nothing forwarded / prefetched for this block
            const bool next_same = more && Gn == G;          // the next block continues this group
            const double *Lc = L0 + lb * 992;
            double *Ln = L0 + (lb ^ 1) * 992;
            double zr[39];
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
                const int Rb = s0 + 31 - 8 * sub - 6 + 32 * t;
                if (sub == 0) {
                    if (first) {
#pragma unroll
                        for (int x = 0; x < 39; ++x) zr[x] = Zc[(size_t)(Rb + x) * kp];
                    } else {
                        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#pragma unroll
                        for (int x = 0; x < 7; ++x) zr[x] = Lc[(24 + x) * 32];
#pragma unroll
                        for (int x = 0; x < 32; ++x) zr[7 + x] = F[x * 32];
                    }
                    if (next_same) {
                        // F is refilled: its reads above have to have completed
#pragma unroll
                        for (int x = 7; x < 39; x += 8)
                            asm volatile("" ::"d"(zr[x]), "d"(zr[x + 1]), "d"(zr[x + 2]), "d"(zr[x + 3]), "d"(zr[x + 4]),
                                         "d"(zr[x + 5]), "d"(zr[x + 6]), "d"(zr[x + 7]));
                        const double *src = Zc + (size_t)(Rb + 39) * kp;       // rows Wb + 63 .. Wb + 94
#pragma unroll
                        for (int x = 0; x < 32; ++x) cp_async8_t(F + x * 32, src + (size_t)x * kp);
                    }
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                } else {
#pragma unroll
                    for (int x = 30; x >= 0; --x) zr[x + 8] = zr[x];
                    if (first) {
#pragma unroll
                        for (int x = 0; x < 8; ++x) zr[x] = Zc[(size_t)(Rb + x) * kp];
                    } else {
#pragma unroll
                        for (int x = 0; x < 8; ++x) zr[x] = Lc[(24 - 8 * sub + x) * 32];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int sw = 31 - 8 * sub - u;
                    const double tau = taus[buf][sw];
                    if (tau != 0.0) {
                        const double2 *vv = reinterpret_cast<const double2 *>(&Vs[buf][sw][0]);
                        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                        for (int l2 = 0; l2 < 16; l2 += 2) {
                            const double2 x0 = vv[l2], x1 = vv[l2 + 1];
                            d0 += x0.x * zr[7 - u + 2 * l2];
                            d1 += x0.y * zr[7 - u + 2 * l2 + 1];
                            d2 += x1.x * zr[7 - u + 2 * l2 + 2];
                            d3 += x1.y * zr[7 - u + 2 * l2 + 3];
                        }
                        // (eight partial sums instead of four: no gain, 57.3 against 56.2 ms on C3)
                        const double dot = tau * ((d0 + d1) + (d2 + d3));
#pragma unroll
                        for (int l2 = 0; l2 < 16; ++l2) {
                            const double2 x0 = vv[l2];
                            zr[7 - u + 2 * l2] -= dot * x0.x;
                            zr[7 - u + 2 * l2 + 1] -= dot * x0.y;
                        }
                    }
                }
                // finished rows: the lower 31 go to the next block of the group through L (it stores them when it
                // is done with them), everything else -- and everything of a group's last block -- to global memory
                if (sub < 3) {
                    if (next_same) {
#pragma unroll
                        for (int x = 31; x < 39; ++x) Ln[(x - 8 - 8 * sub) * 32] = zr[x];      // rows 23..30 / 15..22 / 7..14 of the next block
                    } else {
#pragma unroll
                        for (int x = 31; x < 39; ++x) Zc[(size_t)(Rb + x) * kp] = zr[x];
                    }
                } else {
#pragma unroll
                    for (int x = 0; x < 32; ++x) Zc[(size_t)(Rb + x) * kp] = zr[x];
                    if (next_same) {
#pragma unroll
                        for (int x = 32; x < 39; ++x) Ln[(x - 32) * 32] = zr[x];               // rows 0..6 of the next block
                    } else {
#pragma unroll
                        for (int x = 32; x < 39; ++x) Zc[(size_t)(Rb + x) * kp] = zr[x];
                    }
                }
            }
            if (next_same) lb ^= 1;
        }
        if (more) stash(buf ^ 1, pv, pt);
        __syncthreads();
        if (!more) break;
        G = Gn; t = tn; buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// back-transformation with the band-reduction reflectors, on the row-major eigenvector window
// ------------------------------------------------------------------------------------------
// Zt <- Q1 Zt, Q1 = prod_p (I - V_p T_p V_p^T), panels last to first (LAPACK dormtr 'L','N').
// One CTA per (matrix, group of 256 eigenvector columns); warp <-> 32 columns.  The 32x32 chunk of
// V_p is staged ONCE per CTA and step in shared memory and shared by the 8 warps (the one-stage
// kernel keeps an 8-column slab resident and re-reads V from L2 for every slab: 64 passes over V
// per matrix at k = 250); every warp streams its own 32x32 chunk of Zt through a transit tile.
//   pass 1   W = V_p^T Zt      (accumulated over the row chunks, DMMA)
//            W <- -T_p W       (DMMA, through shared memory for the operand layout)
//   pass 2   Zt += V_p W       (per chunk, DMMA, written back with 128-bit stores)
constexpr int OT2_PITCH = 36;
// OTW warps per CTA (8: one CTA covers 256 columns; 4: two CTAs per 256 columns for small batches)
template <int OTW>
__global__ void __launch_bounds__(OTW * 32, 1) ormtr_t_kernel(TsArgs a)
{
    constexpr int PT = OT2_PITCH;
    const int mat = blockIdx.y;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int grp = lane >> 2, tig = lane & 3;
    const int lda = a.Npad, NB = a.NB;
    const size_t kp = (size_t)a.kp;
    const int c0 = 32 * OTW * blockIdx.x + 32 * wid;
    const bool act = c0 < a.kp;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *__restrict__ A = slot + a.oA;
    const double *__restrict__ Tfac = slot + a.oTfac;
    double *Zt = slot + a.oZt;

    extern __shared__ double sm[];
    double(*Vs)[32 * PT] = reinterpret_cast<double(*)[32 * PT]>(sm);        // [buf][row k][col j]
    double *Ts = sm + 2 * 32 * PT;                                          // [j][l]
    double *Tz = sm + 3 * 32 * PT + wid * (2 * 32 * PT);                    // per warp: Z chunk [row][col]
    double *Wt = Tz + 32 * PT;                                              // per warp: W [j][col]

    // V chunk ch of panel p -> Vs[buf]: element (row rr, reflector cc); the first chunk carries the
    // implicit unit diagonal and the zeros above it (the stored entries there belong to R)
    auto stageV = [&](int buf, int p, int ch) {
        const int r0 = 32 * (p + 1) + 32 * ch;
        const int rr = tid & 31;
#pragma unroll
        for (int cc = tid >> 5; cc < 32; cc += OTW) {
            double x = A[(size_t)(r0 + rr) + (size_t)(32 * p + cc) * lda];
            if (ch == 0) x = (rr < cc) ? 0.0 : ((rr == cc) ? 1.0 : x);
            Vs[buf][rr * PT + cc] = x;
        }
    };
    // this warp's 32x32 chunk of Zt (rows r0.., columns c0..) -> Tz, 16-byte pieces
    auto issueZ = [&](int r0) {
        const double *src = Zt + (size_t)(r0 + (lane >> 4)) * kp + c0 + (lane & 15) * 2;
        double *dst = Tz + (lane >> 4) * PT + (lane & 15) * 2;
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            cp_async16_t(dst, src);
            src += 2 * kp;
            dst += 2 * PT;
        }
    };

    for (int p = NB - 2; p >= 0; --p) {
        const int m0 = 32 * (p + 1), nch = NB - (p + 1);
        __syncthreads();
        for (int t = tid; t < 1024; t += OTW * 32) Ts[(t >> 5) * PT + (t & 31)] = Tfac[(size_t)p * 1024 + t];
        // ---------------------------------------------------------------- pass 1: W = V^T Z
        double acc[4][4][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
        stageV(0, p, 0);
        if (act) issueZ(m0);
        cp_async_wait_all_t();
        __syncthreads();
        for (int ch = 0; ch < nch; ++ch) {
            const int buf = ch & 1;
            double fb[4][8];
            if (act) {
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) fb[y][ks] = Tz[(4 * ks + tig) * PT + 8 * y + grp];
            }
            __syncwarp();
            if (ch + 1 < nch) {
                if (act) issueZ(m0 + 32 * (ch + 1));
                stageV(buf ^ 1, p, ch + 1);
            }
            if (act) {
                const double *va = Vs[buf] + tig * PT + grp;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    double fa[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) fa[x] = va[4 * ks * PT + 8 * x];     // V^T: (m = j, k = row)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
#pragma unroll
                        for (int x = 0; x < 4; ++x) dmma_t(acc[x][y][0], acc[x][y][1], fa[x], fb[y][ks]);
                }
            }
            cp_async_wait_all_t();
            __syncthreads();
        }
        // ---------------------------------------------------------------- W <- -T W
        if (act) {
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int h = 0; h < 2; ++h) Wt[(8 * x + grp) * PT + 8 * y + 2 * tig + h] = acc[x][y][h];
            __syncwarp();
            double w2[4][4][2];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) w2[x][y][0] = w2[x][y][1] = 0.0;
            const double *ta = Ts + grp * PT + tig;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                double fa[4], fw[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) fa[x] = ta[8 * x * PT + 4 * ks];            // T (m = j, k = l)
#pragma unroll
                for (int y = 0; y < 4; ++y) fw[y] = Wt[(4 * ks + tig) * PT + 8 * y + grp];
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int x = 0; x < 4; ++x) dmma_t(w2[x][y][0], w2[x][y][1], fa[x], fw[y]);
            }
            __syncwarp();
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int h = 0; h < 2; ++h) Wt[(8 * x + grp) * PT + 8 * y + 2 * tig + h] = -w2[x][y][h];
            __syncwarp();
        }
        // ---------------------------------------------------------------- pass 2: Z += V W
        __syncthreads();
        stageV(0, p, 0);
        if (act) issueZ(m0);
        cp_async_wait_all_t();
        __syncthreads();
        for (int ch = 0; ch < nch; ++ch) {
            const int buf = ch & 1;
            const int r0 = m0 + 32 * ch;
            double c[4][4][2];
            if (act) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const double2 q = *reinterpret_cast<const double2 *>(Tz + (8 * x + grp) * PT + 8 * y + 2 * tig);
                        c[x][y][0] = q.x;
                        c[x][y][1] = q.y;
                    }
            }
            __syncwarp();
            if (ch + 1 < nch) {
                if (act) issueZ(r0 + 32);
                stageV(buf ^ 1, p, ch + 1);
            }
            if (act) {
                const double *va = Vs[buf] + grp * PT + tig;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    double fa[4], fw[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) fa[x] = va[8 * x * PT + 4 * ks];        // V (m = row, k = j)
#pragma unroll
                    for (int y = 0; y < 4; ++y) fw[y] = Wt[(4 * ks + tig) * PT + 8 * y + grp];
#pragma unroll
                    for (int y = 0; y < 4; ++y)
#pragma unroll
                        for (int x = 0; x < 4; ++x) dmma_t(c[x][y][0], c[x][y][1], fa[x], fw[y]);
                }
                double *zo = Zt + (size_t)(r0 + grp) * kp + c0 + 2 * tig;
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
                        *reinterpret_cast<double2 *>(zo + (size_t)(8 * x) * kp + 8 * y) = make_double2(c[x][y][0], c[x][y][1]);
            }
            cp_async_wait_all_t();
            __syncthreads();
        }
    }
}


// Look-ahead kernel (default for batches of whole-SM matrices): per panel p
//   (A) rank-2k update of block column p + 1 only      (s1_syr2k_wg, both warpgroups)
//   (B) panel QR + T factor of panel p + 1
//   (C) fused pass over the remaining tiles: update with panel p, X' = A22 V' of panel p + 1
//   (D) W' = X' T' - 1/2 V' (T'^T V'^T X' T')
// The panels alternate between two buffer pairs; X' is accumulated with global reductions in a
// third panel and copied (L2 loads) into the W buffer of panel p + 1.
constexpr size_t S1FZ_SMEM_BYTES = (size_t)(2 * 8 * 32 * S1_TSTR + 2 * 4608 + 128) * sizeof(double) + 640;
__global__ void __launch_bounds__(256, 1) sy2sb_fused_kernel(TsArgs a)
{
    constexpr int WARPS = 8;
    constexpr int TSTR = S1_TSTR;
    constexpr int TILES = 2 * WARPS * 32 * TSTR;
    const int mat = blockIdx.x;
    double *slot = a.base + (size_t)mat * a.slot;
    extern __shared__ double sm[];
    const int np = a.Npad, NB = a.NB;
    S1Ctx c;
    c.tid = threadIdx.x; c.wid = c.tid >> 5; c.lane = c.tid & 31; c.grp = c.lane >> 2; c.tig = c.lane & 3;
    c.np = np; c.lda = np; c.NB = NB;
    c.A = slot + a.oA; c.tt = slot + a.oTau; c.Tfac = slot + a.oTfac;
    c.Bs = reinterpret_cast<double(*)[2][32][BSTR]>(sm + TILES);
    double *bsB = sm + TILES + 4608;
    c.Ts = reinterpret_cast<double(*)[33]>(bsB);
    c.G = reinterpret_cast<double(*)[33]>(bsB + 1056);
    c.M1 = reinterpret_cast<double(*)[33]>(bsB + 2 * 1056);
    double *small = sm + TILES + 2 * 4608;
    c.wv = reinterpret_cast<double(*)[32]>(small);
    c.sred = reinterpret_cast<double(*)[16]>(small + 64);
    c.taus = small + 96;
    c.Tw = sm + c.wid * (2 * 32 * TSTR);
    // barriers: 16 tile + 4 warpgroup operand (TMA phases); fused pass: 4 block rows x 8, 2 full + 2 empty
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(small + 128);
    int *ctl = reinterpret_cast<int *>(bars + 56);
    S1Wg g;
    g.wg = c.wid >> 2; g.wl = c.wid & 3;
    g.Bs = g.wg ? reinterpret_cast<double(*)[2][32][BSTR]>(bsB) : c.Bs;
    g.tbar = bars + 2 * c.wid;
    g.bbar = bars + 16 + 2 * g.wg;
    g.next = ctl;
    g.cur = ctl + 1 + g.wg;
    if (c.tid == 0) {
        for (int i = 0; i < 20; ++i) mbar_init(&bars[i], 1);
        for (int r = 0; r < 4; ++r) {
            unsigned long long *rb = bars + 20 + 8 * r;
            mbar_init(&rb[RB_CFULL], 32); mbar_init(&rb[RB_CFULL + 1], 32);      // cp.async arrivals of the helper's lanes
            mbar_init(&rb[RB_CREADY], 1); mbar_init(&rb[RB_CREADY + 1], 1);
            mbar_init(&rb[RB_CDONE], 1); mbar_init(&rb[RB_CDONE + 1], 1);
            mbar_init(&rb[RB_VFULL], 32); mbar_init(&rb[RB_VFREE], 1);
        }
        mbar_init(&bars[52], 128); mbar_init(&bars[53], 128);                    // full: 4 helpers x 32 lanes
        mbar_init(&bars[54], 4); mbar_init(&bars[55], 4);                        // empty: the four compute warps
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    fence_async_proxy();
    __syncthreads();
    // panel buffers: even panels in (oVp, oWp), odd panels in the (idle) eigenvector matrix
    double *Vb[2] = {slot + a.oVp, slot + a.oZ};
    double *Wb[2] = {slot + a.oWp, slot + a.oZ + (size_t)32 * np};
    double *Xacc = slot + a.oZ + (size_t)64 * np;
    S1Fz f;
    f.Xacc = Xacc;
    f.Bs3 = reinterpret_cast<double(*)[3][32][BSTR]>(sm + TILES);
    f.tiles = sm + (c.wid & 3) * (4 * 32 * TSTR);
    f.rb = bars + 20 + 8 * (c.wid & 3);
    f.full = bars + 52;
    f.empty = bars + 54;
    unsigned tph = 0, bph = 0;
    S1FzPhase fph;
    fph.rb = 0; fph.ring = 0;
    auto reload_T = [&](int p) {
        for (int t = c.tid; t < 1024; t += WARPS * 32) c.Ts[t >> 5][t & 31] = c.Tfac[(size_t)p * 1024 + t];
    };
    // ---- panel 0: QR, T, X = A22 V (two reads of the lower triangle, once), W
    c.Vp = Vb[0]; c.Wp = Wb[0];
    s1_panel_qr<WARPS>(c, 0);
    s1_tfactor<WARPS>(c, 0);
    s1_symm_wg(c, g, 0, tph, bph);
    reload_T(0);
    s1_wphase<WARPS>(c, 0);
    for (int p = 0; p < NB - 1; ++p) {
        const bool last = (p == NB - 2);
        double *Vn = Vb[(p + 1) & 1], *Wn = Wb[(p + 1) & 1];
        if (!last) {
            // X' accumulator of the rows below the next panel's band
            for (int n = 0; n < 32; ++n)
                for (int r = 32 * (p + 2) + c.tid; r < np; r += WARPS * 32) Xacc[(size_t)n * np + r] = 0.0;
            __threadfence();
        }
        c.Vp = Vb[p & 1]; c.Wp = Wb[p & 1];
        s1_syr2k_wg(c, g, p, tph, bph, p + 1);                       // (A)
        if (last) break;
        c.Vp = Vn; c.Wp = Wn;
        s1_panel_qr<WARPS>(c, p + 1);                                // (B)
        s1_tfactor<WARPS>(c, p + 1);
        f.Vc = Vb[p & 1]; f.Wc = Wb[p & 1]; f.Vn = Vn;
        s1_fused_pass(c, f, p, fph);                                 // (C)
        for (int n = 0; n < 32; ++n)                                 // (D)
            for (int r = 32 * (p + 2) + c.tid; r < np; r += WARPS * 32) Wn[(size_t)n * np + r] = __ldcg(&Xacc[(size_t)n * np + r]);
        reload_T(p + 1);
        s1_wphase<WARPS>(c, p + 1);
    }
    __syncthreads();
    s1_band_copy(a, slot, c.A, c.lda, c.tid, WARPS * 32);
}


// Barrier-free tensor phases (see s1_symm_ring): same algorithm and phase order as the lockstep kernel.
constexpr size_t S1RING_SMEM_BYTES = (size_t)(2 * 8 * 32 * S1_TSTR + 2 * 4608 + 128) * sizeof(double) + 256;
__global__ void __launch_bounds__(256, 1) sy2sb_ring_kernel(TsArgs a)
{
    constexpr int WARPS = 8;
    constexpr int TSTR = S1_TSTR;
    constexpr int TILES = 2 * WARPS * 32 * TSTR;
    const int mat = blockIdx.x;
    double *slot = a.base + (size_t)mat * a.slot;
    extern __shared__ double sm[];
    S1Ctx c;
    c.tid = threadIdx.x; c.wid = c.tid >> 5; c.lane = c.tid & 31; c.grp = c.lane >> 2; c.tig = c.lane & 3;
    c.np = a.Npad; c.lda = a.Npad; c.NB = a.NB;
    c.A = slot + a.oA; c.Wp = slot + a.oWp; c.Vp = slot + a.oVp; c.tt = slot + a.oTau; c.Tfac = slot + a.oTfac;
    // [transit tiles][operand ring: 4 stages x 2 tiles][small arrays][barriers]; the W phase uses the first two
    // ring tiles as its operand tiles, T / G / M1 live in the second half of the ring (idle outside the tensor
    // phases; T is reloaded from global memory after the symmetric product has used the ring)
    c.Bs = reinterpret_cast<double(*)[2][32][BSTR]>(sm + TILES);
    double *half2 = sm + TILES + 4608;
    c.Ts = reinterpret_cast<double(*)[33]>(half2);
    c.G = reinterpret_cast<double(*)[33]>(half2 + 1056);
    c.M1 = reinterpret_cast<double(*)[33]>(half2 + 2 * 1056);
    double *small = sm + TILES + 2 * 4608;
    c.wv = reinterpret_cast<double(*)[32]>(small);
    c.sred = reinterpret_cast<double(*)[16]>(small + 64);
    c.taus = small + 96;
    c.Tw = sm + c.wid * (2 * 32 * TSTR);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(small + 128);
    S1Ring r;
    r.ring = reinterpret_cast<double(*)[2][32][BSTR]>(sm + TILES);
    r.full = bars;
    r.empty = bars + 4;
    if (c.tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&r.full[i], 256); mbar_init(&r.empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    unsigned gu = 0;               // operand-ring steps so far (stage = gu & 3, use = gu >> 2)
    TSP_DECL
    for (int p = 0; p < a.NB - 1; ++p) {
        TSP(0)
        s1_panel_qr<WARPS>(c, p);
        TSP(1)
        s1_tfactor<WARPS>(c, p);
        TSP(2)
        s1_symm_ring(c, r, p, gu);
        TSP(3)
        for (int t = c.tid; t < 1024; t += WARPS * 32) c.Ts[t >> 5][t & 31] = c.Tfac[(size_t)p * 1024 + t];
        s1_wphase<WARPS>(c, p);
        TSP(4)
        s1_syr2k_ring(c, r, p, gu);
        TSP(5)
    }
#ifdef TS_PROFILE
    if (mat == 0 && c.tid == 0)
        printf("sy2sb ring prof (10 kcycles): qr %lld tfac %lld symm %lld w %lld syr2k %lld\n", tsp_t[1] / 10000,
               tsp_t[2] / 10000, tsp_t[3] / 10000, tsp_t[4] / 10000, tsp_t[5] / 10000);
#endif
    __syncthreads();
    s1_band_copy(a, slot, c.A, c.lda, c.tid, WARPS * 32);
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
    // measured (tools/ab_crossover.py, batches of 74 / 296 / 592): the two-stage path is 1.1-1.5x faster than the
    // one-stage cluster kernel for every N from 128 to 1000
    return N >= 128 && N <= 16384;
}

int eig_sy2sb(const EigSlots &S, int nmat, cudaStream_t st, int *launches)
{
    if (!S.L.two_stage) { arcb200_set_error("sy2sb: layout without two-stage buffers"); return -1; }
    const TsArgs a = make_args(S);
    // The per-phase kernels (eig_sy2sb_multi.cu: panel QR at two CTAs per SM, symmetric product and rank-2k update
    // with one CTA per 8 block rows) are the default for every batch: the row groups of all matrices balance over
    // the SMs and the latency-bound QR phases of two matrices overlap on one SM (tools/ab_multi.py, one CTA per
    // matrix against per-phase: 2363 x296 262.0 / 250.9 ms, x250 259 / 217.5, 1932 x296 156.7 / 149.3,
    // 1100 x296 38.6 / 35.7, 651 x286 13.3 / 11.5, 10384 x8 8560 / 1263).  ARCB200_SY2SB_MULTI = 0 selects the
    // one-CTA-per-matrix kernels below.
    const char *em = getenv("ARCB200_SY2SB_MULTI");
    const bool multi = !(em && em[0] == '0');
    if (multi) return eig_sy2sb_multi(a, nmat, st, launches);
    // Four kernels for whole waves of one matrix per SM, measured on C3 (N = 2363, 296 matrices;
    // profiles/r02_ab_sy2sb_variants_c3.json):
    //   ring (default)      barrier-free tensor phases: per-warp tile rings + shared operand ring on mbarriers
    //   lockstep            block barrier per tile step, per-thread cp.async (round 1)
    //   wg                  decoupled warpgroups, TMA-unit bulk copies + mbarriers
    //   fused               look-ahead: rank-2k update fused with the next symmetric product
    // ARCB200_SY2SB_VARIANT = ring | lockstep | wg | fused selects one.
    const char *evv = getenv("ARCB200_SY2SB_VARIANT");
    const bool wgonly = evv && evv[0] == 'w';
    const bool fusedk = evv && evv[0] == 'f';
    const bool lockk = evv && evv[0] == 'l';
    const bool ringk = !wgonly && !fusedk && !lockk;
    const char *ev = getenv("ARCB200_SY2SB_WARPS");      // tuning knob (lockstep kernel only)
    const int warps = ev ? atoi(ev) : 8;
    if (ringk) {
        ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)S1RING_SMEM_BYTES));
        sy2sb_ring_kernel<<<nmat, 256, S1RING_SMEM_BYTES, st>>>(a);
    } else if (fusedk) {
        ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)S1FZ_SMEM_BYTES));
        sy2sb_fused_kernel<<<nmat, 256, S1FZ_SMEM_BYTES, st>>>(a);
    } else if (wgonly) {
        ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_wg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)S1WG_SMEM_BYTES));
        sy2sb_wg_kernel<<<nmat, 256, S1WG_SMEM_BYTES, st>>>(a);
    } else if (warps == 4) {
        const size_t smem = (2 * 4 * 32 * 36 + 4608 + 3 * 1056 + 128) * sizeof(double);
        ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sy2sb_kernel<4><<<nmat, 128, smem, st>>>(a);
    } else {
        const size_t smem = (2 * 8 * 32 * 36 + 4608 + 3 * 1056 + 128) * sizeof(double);
        ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sy2sb_kernel<8><<<nmat, 256, smem, st>>>(a);
    }
    ARCB200_LAUNCH_CHECK("sy2sb_kernel");
    if (launches) *launches += 1;
    return 0;
}

int eig_sb2st(const EigSlots &S, int nmat, cudaStream_t st, int *launches)
{
    if (!S.L.two_stage) { arcb200_set_error("sb2st: layout without two-stage buffers"); return -1; }
    const TsArgs a = make_args(S);
    // two warps per sweep (eig_sb2st_split.cu) up to N = 3072 (tools/ab_sb2st.py: 1027 x3 11.7 -> 10.4 ms, 2363 x296
    // 112.0 -> 107.7, 1932 x148 39.5 -> 35.9, 4096 x20 149 -> 158); ARCB200_SB2ST_SPLIT = 0 / 1 forces one kernel
    const char *es = getenv("ARCB200_SB2ST_SPLIT");
    if (es ? es[0] != '0' : a.N <= 3072) return eig_sb2st_split(a, nmat, st, launches);
    const size_t smem = (size_t)(TS_WARPS * 32 * 33 + 2 * TS_WARPS * 32) * sizeof(double);
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sb2st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sb2st_kernel<<<nmat, TS_THREADS, smem, st>>>(a);
    ARCB200_LAUNCH_CHECK("sb2st_kernel");
    if (launches) *launches += 1;
    return 0;
}

int eig_q2_apply(const EigSlots &S, int nmat, const int32_t *d_sel0, int nsel, cudaStream_t st,
                 int *launches, bool with_q1)
{
    if (S.L.N < 3 || nsel <= 0) return 0;
    TsArgs a = make_args(S);
    a.nsel = nsel;
    a.kp = (nsel + 31) & ~31;
    a.sel0 = d_sel0;
    const int rowtiles = (S.L.Npad + 64) / 32, coltiles = a.kp / 32;
    zt_in_kernel<<<dim3(rowtiles, coltiles, nmat), 256, 0, st>>>(a);
    const int nslab = a.kp / 32;
    // four slabs (128 eigenvector columns) per CTA, two CTAs per SM; batches that would leave SMs without a CTA
    // (C5: 74 matrices x 250 vectors) get two slabs per CTA.  ARCB200_Q2_WARPS = 2 | 4 forces one.
    const char *eqw = getenv("ARCB200_Q2_WARPS");
    const bool q2small = eqw ? eqw[0] == '2' : (long long)nmat * ((nslab + 3) / 4) < 2 * 148;
    if (q2small) {
        const size_t q2smem = (size_t)2 * (1024 + 2 * 992) * sizeof(double);
        ARCB200_CUDA_CHECK(cudaFuncSetAttribute(q2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q2smem));
        q2_kernel<2><<<dim3((nslab + 1) / 2, nmat), 64, q2smem, st>>>(a);
    } else {
        const size_t q2smem = (size_t)4 * (1024 + 2 * 992) * sizeof(double);
        ARCB200_CUDA_CHECK(cudaFuncSetAttribute(q2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q2smem));
        q2_kernel<4><<<dim3((nslab + 3) / 4, nmat), 128, q2smem, st>>>(a);
    }
    int nl = 3;
    if (with_q1 && S.L.NB >= 2) {
        // Q1 on the same row-major window (T factors were left in oTfac by the band reduction)
        if ((long long)nmat * ((a.kp + 255) / 256) < 148) {
            const size_t smem = (size_t)(3 * 32 * OT2_PITCH + 4 * 2 * 32 * OT2_PITCH) * sizeof(double);
            ARCB200_CUDA_CHECK(cudaFuncSetAttribute(ormtr_t_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ormtr_t_kernel<4><<<dim3((a.kp + 127) / 128, nmat), 128, smem, st>>>(a);
        } else {
            const size_t smem = (size_t)(3 * 32 * OT2_PITCH + 8 * 2 * 32 * OT2_PITCH) * sizeof(double);
            ARCB200_CUDA_CHECK(cudaFuncSetAttribute(ormtr_t_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ormtr_t_kernel<8><<<dim3((a.kp + 255) / 256, nmat), 256, smem, st>>>(a);
        }
        ++nl;
    }
    zt_out_kernel<<<dim3(S.L.NB, coltiles, nmat), 256, 0, st>>>(a);
    ARCB200_LAUNCH_CHECK("q2 kernels");
    if (launches) *launches += nl;
    return 0;
}
