// eig_stedc.cu -- stage 2 of the batched symmetric eigensolver: divide and conquer
// for the symmetric tridiagonal eigenproblem (Cuppen / Gu-Eisenstat, the algorithm of
// LAPACK dstedc, which is what numpy.linalg.eigh runs in the reference,
// calculations_atom_single.py:986), restructured for a batch of equally sized problems.
//
//   tear      T -> 2^D leaf blocks (<= 32 rows) by rank-one tearing
//   leaves    implicit-shift QL per leaf, one warp per leaf (eigenvector rows <-> lanes)
//   merge x D per level, for all merges of all matrices at once:
//               prep     z vector, sort, deflation (tiny z, close poles via Givens)
//               secular  one thread per root, two-pole rational iteration in pole-shifted
//                        coordinates, bisection safeguarded
//               zhat     Gu-Eisenstat recomputation of z (orthogonality by construction)
//               umat     secular eigenvectors U, final ordering of roots + deflated
//               gemm     Qnew = Qold[:, nondeflated] * U  on FP64 tensor cores (DMMA)
//               copy     deflated columns
// Node (level l, index q) covers rows [q*N >> l, (q+1)*N >> l).
#include <float.h>
#include "eig.cuh"
#include "eig_gemm.cuh"

namespace {

constexpr int LEAF = 32;
constexpr double EPS = 2.220446049250313e-16;   // DBL_EPSILON

struct DcArgs {
    int N, Npad, D;          // D = tree depth
    double *base; size_t slot;
    int32_t *ibase; size_t islot;
    size_t oQa, oQb, oU, oD, oE, oLam, oAux;
    size_t oPerm, oIdx;
    const int32_t *win0;     // per matrix first selected final position (level 0 only)
    int wink;                // number of selected eigenpairs (0 = all)
};

// per-slot aux double arrays (each Npad long), offsets inside oAux
enum { AX_DLAM = 0, AX_WZ, AX_MU, AX_ZHAT, AX_LAMNEW, AX_DEFL, AX_RHO, AX_LAMROOT, AX_TMP, AX_TMP2 };
// per-slot int arrays (each Npad long) inside oPerm / oIdx
enum { IX_COLSRC = 0, IX_ORIGIN, IX_KP, IX_POSROOT, IX_DEFSRC, IX_DEFPOS, IX_TMP, IX_TMP2 };

__device__ __forceinline__ int node_start(int N, int level, int q)
{
    return (int)(((long long)q * N) >> level);
}

// ------------------------------------------------------------------ tear
__global__ void tear_kernel(DcArgs a, int nmat)
{
    const int mat = blockIdx.x;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    double *d = slot + a.oD;
    const double *e = slot + a.oE;
    double *lam = slot + a.oLam;
    double *rho = slot + a.oAux + (size_t)AX_RHO * a.Npad;
    const int N = a.N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) lam[i] = d[i];
    __syncthreads();
    const int nleaf = 1 << a.D;
    // every internal boundary m = node_start(D, q), q = 1..nleaf-1, tears once
    for (int q = 1 + threadIdx.x; q < nleaf; q += blockDim.x) {
        const int m = node_start(N, a.D, q);
        const double r = e[m - 1];
        rho[m] = r;                       // indexed by the first row of the right block
        // distinct boundaries touch distinct (m-1, m) pairs only if leaves have >= 2 rows;
        // leaves have >= 16 rows by construction
        lam[m - 1] -= fabs(r);
        lam[m] -= fabs(r);
    }
}

// ------------------------------------------------------------------ leaves
// One warp per leaf.  d/e scalars are processed redundantly by all lanes (uniform control
// flow), lane r owns row r of the leaf's eigenvector matrix (shared memory, padded).
constexpr int LEAF_WARPS = 4;
__global__ void __launch_bounds__(LEAF_WARPS * 32) leaf_kernel(DcArgs a, int nmat)
{
    __shared__ double Z[LEAF_WARPS][LEAF][LEAF + 1];
    __shared__ double sd[LEAF_WARPS][LEAF], se[LEAF_WARPS][LEAF];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nleaf = 1 << a.D;
    const long long gl = (long long)blockIdx.x * LEAF_WARPS + wid;
    if (gl >= (long long)nmat * nleaf) return;
    const int mat = (int)(gl / nleaf), q = (int)(gl % nleaf);
    double *slot = a.base + (size_t)mat * a.slot;
    const int s = node_start(a.N, a.D, q), n = node_start(a.N, a.D, q + 1) - s;
    double *lam = slot + a.oLam;
    const double *e = slot + a.oE;
    double *Q = slot + a.oQa;
    double *d_ = sd[wid], *e_ = se[wid];
    double(*z)[LEAF + 1] = Z[wid];
    if (lane < n) {
        d_[lane] = lam[s + lane];
        e_[lane] = (lane < n - 1) ? e[s + lane] : 0.0;
    }
    for (int c = 0; c < n; ++c) z[lane][c] = (lane == c) ? 1.0 : 0.0;
    __syncwarp();

    for (int l = 0; l < n; ++l) {
        int iter = 0;
        while (true) {
            int m = l;
            for (; m < n - 1; ++m) {
                const double dd = fabs(d_[m]) + fabs(d_[m + 1]);
                if (fabs(e_[m]) <= EPS * dd) break;
            }
            if (m == l) break;
            if (++iter > 80) break;
            double g = (d_[l + 1] - d_[l]) / (2.0 * e_[l]);
            double r = sqrt(g * g + 1.0);
            g = d_[m] - d_[l] + e_[l] / (g + copysign(r, g));
            double sn = 1.0, cs = 1.0, p = 0.0;
            int i = m - 1;
            bool under = false;
            __syncwarp();
            for (; i >= l; --i) {
                double f = sn * e_[i];
                const double b = cs * e_[i];
                r = sqrt(f * f + g * g);
                __syncwarp();
                if (lane == 0) e_[i + 1] = r;
                if (r == 0.0) {
                    if (lane == 0) { d_[i + 1] -= p; e_[m] = 0.0; }
                    under = true;
                    __syncwarp();
                    break;
                }
                sn = f / r;
                cs = g / r;
                g = d_[i + 1] - p;
                r = (d_[i] - g) * sn + 2.0 * cs * b;
                p = sn * r;
                __syncwarp();
                if (lane == 0) d_[i + 1] = g + p;
                g = cs * r - b;
                // rotate columns i, i+1 of the eigenvector matrix (lane = row)
                if (lane < n) {
                    f = z[lane][i + 1];
                    z[lane][i + 1] = sn * z[lane][i] + cs * f;
                    z[lane][i] = cs * z[lane][i] - sn * f;
                }
                __syncwarp();
            }
            if (under) continue;
            __syncwarp();
            if (lane == 0) { d_[l] -= p; e_[l] = g; e_[m] = 0.0; }
            __syncwarp();
        }
    }
    __syncwarp();
    // ascending order: rank of each eigenvalue (ties broken by index)
    int rank = 0;
    double mine = 0.0;
    if (lane < n) {
        mine = d_[lane];
        for (int k = 0; k < n; ++k) {
            const double o = d_[k];
            rank += (o < mine) || (o == mine && k < lane);
        }
        lam[s + rank] = mine;
    }
    // column `lane` of z goes to column s+rank of Q
    if (lane < n) {
        double *qc = Q + (size_t)(s + rank) * a.Npad + s;
        for (int r2 = 0; r2 < n; ++r2) qc[r2] = z[r2][lane];
    }
}

// ------------------------------------------------------------------ merge: prep
// One CTA per merge.  Produces (all arrays indexed by global row, merges are disjoint):
//   kp[s]                number of non-deflated poles k'
//   dlam[s..s+k')        poles (ascending), wz[..] z components, colsrc[..] source column
//   defl[s+k'..s+n)      deflated eigenvalues, defsrc[..] their source columns (sorted asc)
//   rho_eff at AX_TMP[s]
// and applies the deflation Givens rotations to Qold in place.
constexpr int PREP_THREADS = 256;
__global__ void __launch_bounds__(PREP_THREADS) merge_prep_kernel(DcArgs a, int level, int nmat,
                                                                  int cur /*0: Qa, 1: Qb*/)
{
    const int nnode = 1 << level;
    const int mat = blockIdx.x / nnode, q = blockIdx.x % nnode;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    int32_t *islot = a.ibase + (size_t)mat * a.islot;
    const int N = a.N, ld = a.Npad;
    const int s = node_start(N, level, q), send = node_start(N, level, q + 1);
    const int mid = node_start(N, level + 1, 2 * q + 1);
    const int n = send - s, n1 = mid - s;
    double *Q = slot + (cur ? a.oQb : a.oQa);
    double *lam = slot + a.oLam;
    double *aux = slot + a.oAux;
    double *dlam = aux + (size_t)AX_DLAM * ld, *wz = aux + (size_t)AX_WZ * ld;
    double *defl = aux + (size_t)AX_DEFL * ld, *tmpd = aux + (size_t)AX_TMP * ld;
    double *zv = aux + (size_t)AX_TMP2 * ld;          // z vector (global index)
    double *dsort = aux + (size_t)AX_LAMNEW * ld;     // scratch: sorted d (reused later)
    int32_t *colsrc = islot + a.oPerm + (size_t)IX_COLSRC * ld;
    int32_t *kp = islot + a.oPerm + (size_t)IX_KP * ld;
    int32_t *defsrc = islot + a.oPerm + (size_t)IX_DEFSRC * ld;
    int32_t *indx = islot + a.oPerm + (size_t)IX_TMP * ld;     // sorted order -> column
    int32_t *rotl = islot + a.oPerm + (size_t)IX_TMP2 * ld;    // rotation list (pairs packed)
    double *rotc = aux + (size_t)AX_ZHAT * ld, *rots = aux + (size_t)AX_MU * ld;  // scratch
    const double rho_in = aux[(size_t)AX_RHO * ld + mid];
    const int tid = threadIdx.x;
    __shared__ double sred[PREP_THREADS];
    __shared__ int s_k, s_nrot, s_ndef;

    // z = [last row of Q1-block ; +-first row of Q2-block] / sqrt(2)
    const double sgn = (rho_in < 0.0) ? -1.0 : 1.0;
    double zmax = 0.0, dmax = 0.0;
    for (int j = tid; j < n; j += PREP_THREADS) {
        double zz;
        if (j < n1) zz = Q[(size_t)(s + j) * ld + (mid - 1)];
        else zz = sgn * Q[(size_t)(s + j) * ld + mid];
        zz *= 0.70710678118654752440;
        zv[s + j] = zz;
        zmax = fmax(zmax, fabs(zz));
        dmax = fmax(dmax, fabs(lam[s + j]));
    }
    sred[tid] = zmax;
    __syncthreads();
    for (int o = PREP_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) sred[tid] = fmax(sred[tid], sred[tid + o]);
        __syncthreads();
    }
    zmax = sred[0];
    __syncthreads();
    sred[tid] = dmax;
    __syncthreads();
    for (int o = PREP_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) sred[tid] = fmax(sred[tid], sred[tid + o]);
        __syncthreads();
    }
    dmax = sred[0];
    const double rho = fabs(2.0 * rho_in);
    const double tol = 8.0 * EPS * fmax(dmax, zmax);

    // merge the two ascending runs: rank by binary search
    for (int j = tid; j < n; j += PREP_THREADS) {
        const double x = lam[s + j];
        int lo, hi, pos;
        if (j < n1) {   // count elements of run 2 strictly less than x
            lo = 0; hi = n - n1;
            while (lo < hi) { const int m2 = (lo + hi) >> 1; if (lam[mid + m2] < x) lo = m2 + 1; else hi = m2; }
            pos = j + lo;
        } else {        // count elements of run 1 less than or equal to x
            lo = 0; hi = n1;
            while (lo < hi) { const int m2 = (lo + hi) >> 1; if (lam[s + m2] <= x) lo = m2 + 1; else hi = m2; }
            pos = (j - n1) + lo;
        }
        indx[s + pos] = s + j;
        dsort[s + pos] = x;
    }
    __syncthreads();

    // sequential deflation scan (LAPACK dlaed2 logic), one thread
    if (tid == 0) {
        int k = 0, ndef = 0, nrot = 0;
        if (rho * zmax <= tol) {
            // everything deflates
            for (int j = 0; j < n; ++j) { defl[s + j] = dsort[s + j]; defsrc[s + j] = indx[s + j]; }
            ndef = n;
        } else {
            int pj = -1;
            double dpj = 0.0;
            // deflated entries are collected at the tail in scan order, sorted afterwards
            for (int j = 0; j < n; ++j) {
                const int nj = indx[s + j];
                const double dj = dsort[s + j];
                if (rho * fabs(zv[nj]) <= tol) {
                    ++ndef;
                    defl[s + n - ndef] = dj;
                    defsrc[s + n - ndef] = nj;
                } else if (pj < 0) {
                    pj = nj; dpj = dj;
                } else {
                    double sv = zv[pj], cv = zv[nj];
                    const double tau = sqrt(cv * cv + sv * sv);
                    const double t = dj - dpj;
                    cv /= tau;
                    sv = -sv / tau;
                    if (fabs(t * cv * sv) <= tol) {
                        zv[nj] = tau;
                        zv[pj] = 0.0;
                        rotl[s + nrot] = (pj - s) | ((nj - s) << 16);
                        rotc[s + nrot] = cv;
                        rots[s + nrot] = sv;
                        ++nrot;
                        const double tt = dpj * cv * cv + dj * sv * sv;
                        const double dnew = dpj * sv * sv + dj * cv * cv;
                        ++ndef;
                        defl[s + n - ndef] = tt;
                        defsrc[s + n - ndef] = pj;
                        pj = nj; dpj = dnew;
                    } else {
                        dlam[s + k] = dpj; wz[s + k] = zv[pj]; colsrc[s + k] = pj; ++k;
                        pj = nj; dpj = dj;
                    }
                }
            }
            if (pj >= 0) { dlam[s + k] = dpj; wz[s + k] = zv[pj]; colsrc[s + k] = pj; ++k; }
        }
        s_k = k; s_nrot = nrot; s_ndef = ndef;
        kp[s] = k;
        tmpd[s] = rho;
    }
    __syncthreads();
    const int k = s_k, nrot = s_nrot, ndef = s_ndef;
    // apply the Givens rotations to the columns of Qold, in order
    for (int t = 0; t < nrot; ++t) {
        const int pk = rotl[s + t];
        const int c1 = s + (pk & 0xffff), c2 = s + (pk >> 16);
        const double cv = rotc[s + t], sv = rots[s + t];
        double *q1 = Q + (size_t)c1 * ld + s, *q2 = Q + (size_t)c2 * ld + s;
        for (int r = tid; r < n; r += PREP_THREADS) {
            const double x = q1[r], y = q2[r];
            q1[r] = cv * x + sv * y;
            q2[r] = cv * y - sv * x;
        }
        __syncthreads();
    }
    // sort the deflated tail ascending (rank by counting; ties by position)
    double *dtmp = dsort;    // reuse
    int32_t *itmp = indx;
    for (int j = tid; j < ndef; j += PREP_THREADS) {
        const double x = defl[s + k + j];
        int rank = 0;
        for (int m2 = 0; m2 < ndef; ++m2) {
            const double o = defl[s + k + m2];
            rank += (o < x) || (o == x && m2 < j);
        }
        dtmp[s + k + rank] = x;
        itmp[s + k + rank] = defsrc[s + k + j];
    }
    __syncthreads();
    for (int j = tid; j < ndef; j += PREP_THREADS) {
        defl[s + k + j] = dtmp[s + k + j];
        defsrc[s + k + j] = itmp[s + k + j];
    }
}

// ------------------------------------------------------------------ merge: secular roots
// f(mu) = 1 + rho * sum_i z_i^2 / (Delta_i - mu),  Delta_i = d_i - d_origin.
__device__ double secular_eval(const double *__restrict__ d, const double *__restrict__ z, int k,
                               double rho, int org, double mu, int split, double &psi,
                               double &dpsi, double &phi, double &dphi)
{
    // psi: poles i <= split (left of the root), phi: poles i > split
    psi = dpsi = phi = dphi = 0.0;
    const double dorg = d[org];
    for (int i = 0; i <= split; ++i) {
        const double t = z[i] / ((d[i] - dorg) - mu);
        psi += z[i] * t;
        dpsi += t * t;
    }
    for (int i = split + 1; i < k; ++i) {
        const double t = z[i] / ((d[i] - dorg) - mu);
        phi += z[i] * t;
        dphi += t * t;
    }
    psi *= rho; dpsi *= rho; phi *= rho; dphi *= rho;
    return 1.0 + psi + phi;
}

__global__ void __launch_bounds__(128) secular_kernel(DcArgs a, int level, int nmat)
{
    const int nnode = 1 << level;
    const int mat = blockIdx.y / nnode, q = blockIdx.y % nnode;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    int32_t *islot = a.ibase + (size_t)mat * a.islot;
    const int ld = a.Npad;
    const int s = node_start(a.N, level, q);
    const int k = islot[a.oPerm + (size_t)IX_KP * ld + s];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    double *aux = slot + a.oAux;
    const double *d = aux + (size_t)AX_DLAM * ld + s;
    const double *z = aux + (size_t)AX_WZ * ld + s;
    const double rho = aux[(size_t)AX_TMP * ld + s];
    double *mu_out = aux + (size_t)AX_MU * ld + s;
    double *lamroot = aux + (size_t)AX_LAMROOT * ld + s;
    int32_t *origin = islot + a.oPerm + (size_t)IX_ORIGIN * ld + s;

    if (k == 1) {
        origin[0] = 0;
        mu_out[0] = rho * z[0] * z[0];
        lamroot[0] = d[0] + mu_out[0];
        return;
    }
    double psi, dpsi, phi, dphi;
    int org;
    double lo, hi, mu;
    const bool last = (j == k - 1);
    if (!last) {
        const double gap = d[j + 1] - d[j];
        // decide the closer pole from the sign of f at the midpoint
        const double fmid = secular_eval(d, z, k, rho, j, 0.5 * gap, j, psi, dpsi, phi, dphi);
        if (fmid >= 0.0) { org = j; lo = 0.0; hi = 0.5 * gap; }
        else { org = j + 1; lo = -0.5 * gap; hi = 0.0; }
        mu = 0.5 * (lo + hi);
    } else {
        double zz = 0.0;
        for (int i = 0; i < k; ++i) zz += z[i] * z[i];
        org = k - 1; lo = 0.0; hi = rho * zz;
        mu = 0.5 * hi;
    }
    const double dorg = d[org];
    const double Da = d[j] - dorg;                       // left pole (shifted)
    const double Db = last ? 0.0 : d[j + 1] - dorg;      // right pole (shifted)
    for (int it = 0; it < 100; ++it) {
        const double f = secular_eval(d, z, k, rho, org, mu, j, psi, dpsi, phi, dphi);
        const double err = 8.0 * EPS * (1.0 + fabs(psi) + fabs(phi)) ;
        if (fabs(f) <= err) break;
        if (f > 0.0) hi = mu; else lo = mu;
        if (!(hi - lo > 2.0 * EPS * fmax(fabs(lo), fabs(hi)))) { mu = 0.5 * (lo + hi); break; }
        double munew;
        const double da = Da - mu;                       // < 0
        if (!last) {
            const double db = Db - mu;                   // > 0
            // psi ~ sA + SA/(Da - x), phi ~ sB + SB/(Db - x)   (value+slope matched at mu)
            const double SA = dpsi * da * da, sA = psi - dpsi * da;
            const double SB = dphi * db * db, sB = phi - dphi * db;
            const double c0 = 1.0 + sA + sB;
            // c0 (Da-x)(Db-x) + SA (Db-x) + SB (Da-x) = 0, unknown eta = x - mu
            // with u = da - eta, w = db - eta:
            const double qa = c0;
            const double qb = -(c0 * (da + db) + SA + SB);
            const double qc = c0 * da * db + SA * db + SB * da;
            double eta;
            if (qa == 0.0) {
                eta = (qb != 0.0) ? -qc / qb : 0.0;
            } else {
                const double disc = qb * qb - 4.0 * qa * qc;
                const double sq = sqrt(fmax(disc, 0.0));
                // two roots; take the one inside (da, db) i.e. da < eta < db
                const double qq = -0.5 * (qb + copysign(sq, qb));
                const double r1 = qq / qa, r2 = (qq != 0.0) ? qc / qq : r1;
                const bool in1 = (r1 > da && r1 < db), in2 = (r2 > da && r2 < db);
                eta = in1 ? r1 : (in2 ? r2 : 0.5 * (lo + hi) - mu);
            }
            munew = mu + eta;
        } else {
            // only left poles: 1 + sA + SA/(Da - x) = 0
            const double SA = dpsi * da * da, sA = psi - dpsi * da;
            const double c0 = 1.0 + sA;
            munew = (c0 != 0.0) ? Da + SA / c0 : 0.5 * (lo + hi);
        }
        if (!(munew > lo && munew < hi)) munew = 0.5 * (lo + hi);
        mu = munew;
    }
    origin[j] = org;
    mu_out[j] = mu;
    lamroot[j] = dorg + mu;
}

// ------------------------------------------------------------------ merge: zhat
__global__ void __launch_bounds__(128) zhat_kernel(DcArgs a, int level, int nmat)
{
    const int nnode = 1 << level;
    const int mat = blockIdx.y / nnode, q = blockIdx.y % nnode;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    int32_t *islot = a.ibase + (size_t)mat * a.islot;
    const int ld = a.Npad;
    const int s = node_start(a.N, level, q);
    const int k = islot[a.oPerm + (size_t)IX_KP * ld + s];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    double *aux = slot + a.oAux;
    const double *d = aux + (size_t)AX_DLAM * ld + s;
    const double *z = aux + (size_t)AX_WZ * ld + s;
    const double *mu = aux + (size_t)AX_MU * ld + s;
    const int32_t *origin = islot + a.oPerm + (size_t)IX_ORIGIN * ld + s;
    double *zhat = aux + (size_t)AX_ZHAT * ld + s;
    if (k <= 2) { zhat[i] = z[i]; return; }   // LAPACK dlaed3: K==1,2 use z directly
    const double di = d[i];
    double w = (di - d[origin[i]]) - mu[i];     // delta_ii = d_i - lambda_i
    for (int j = 0; j < k; ++j) {
        if (j == i) continue;
        const double dij = (di - d[origin[j]]) - mu[j];
        w *= dij / (di - d[j]);
    }
    zhat[i] = copysign(sqrt(fabs(w)), z[i]);
}

// ------------------------------------------------------------------ merge: ordering
// Final order = merge of the ascending roots with the ascending deflated values.
__global__ void __launch_bounds__(128) order_kernel(DcArgs a, int level, int nmat)
{
    const int nnode = 1 << level;
    const int mat = blockIdx.y / nnode, q = blockIdx.y % nnode;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    int32_t *islot = a.ibase + (size_t)mat * a.islot;
    const int ld = a.Npad;
    const int s = node_start(a.N, level, q), n = node_start(a.N, level, q + 1) - s;
    const int k = islot[a.oPerm + (size_t)IX_KP * ld + s];
    const int ndef = n - k;
    double *aux = slot + a.oAux;
    const double *lamroot = aux + (size_t)AX_LAMROOT * ld + s;
    const double *defl = aux + (size_t)AX_DEFL * ld + s + k;
    double *lamnew = aux + (size_t)AX_LAMNEW * ld + s;
    int32_t *posroot = islot + a.oPerm + (size_t)IX_POSROOT * ld + s;
    int32_t *defpos = islot + a.oPerm + (size_t)IX_DEFPOS * ld + s;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < k) {
        // position among deflated values: count deflated < root (ties: roots first)
        const double x = lamroot[t];
        int lo = 0, hi = ndef;
        while (lo < hi) { const int m2 = (lo + hi) >> 1; if (defl[m2] < x) lo = m2 + 1; else hi = m2; }
        posroot[t] = t + lo;
        lamnew[t + lo] = x;
    } else if (t < n) {
        const int jd = t - k;
        const double x = defl[jd];
        int lo = 0, hi = k;
        while (lo < hi) { const int m2 = (lo + hi) >> 1; if (lamroot[m2] <= x) lo = m2 + 1; else hi = m2; }
        defpos[jd] = jd + lo;
        lamnew[jd + lo] = x;
    }
}

// k eigenvalues nearest sigma = a contiguous window of the ascending spectrum
__global__ void window_kernel(int N, const double *base, size_t slot, size_t off, int k,
                              double sigma, int32_t *sel0, int nmat)
{
    const int mat = blockIdx.x * blockDim.x + threadIdx.x;
    if (mat >= nmat) return;
    const double *lam = base + (size_t)mat * slot + off;
    int lo = 0, hi = N;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (lam[m] < sigma) lo = m + 1; else hi = m; }
    int left = lo - 1, right = lo;
    for (int t = 0; t < k; ++t) {
        const bool canL = left >= 0, canR = right < N;
        if (canL && (!canR || fabs(lam[left] - sigma) <= fabs(lam[right] - sigma))) --left;
        else ++right;
    }
    sel0[mat] = left + 1;
}

// ------------------------------------------------------------------ merge: U
// U(:, j) = normalised zhat_i / (d_i - lambda_j), stored at Ubuf[(s+i) + (s+j)*ld].
// One warp per column (coalesced); at level 0 only the columns inside the requested
// eigenvalue window are built.  grid = (ceil(nmax / 4), nmat * nnode), block = 128.
__global__ void __launch_bounds__(128) umat_kernel(DcArgs a, int level, int nmat)
{
    const int nnode = 1 << level;
    const int mat = blockIdx.y / nnode, q = blockIdx.y % nnode;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    int32_t *islot = a.ibase + (size_t)mat * a.islot;
    const int ld = a.Npad;
    const int s = node_start(a.N, level, q);
    const int k = islot[a.oPerm + (size_t)IX_KP * ld + s];
    const int j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= k) return;
    const int32_t *posroot = islot + a.oPerm + (size_t)IX_POSROOT * ld + s;
    if (level == 0 && a.wink > 0) {
        const int w0 = a.win0[mat], p = posroot[j];
        if (p < w0 || p >= w0 + a.wink) return;
    }
    double *aux = slot + a.oAux;
    const double *d = aux + (size_t)AX_DLAM * ld + s;
    const double *mu = aux + (size_t)AX_MU * ld + s;
    const double *zhat = aux + (size_t)AX_ZHAT * ld + s;
    const int32_t *origin = islot + a.oPerm + (size_t)IX_ORIGIN * ld + s;
    const double dorg = d[origin[j]], m = mu[j];
    double *uc = slot + a.oU + (size_t)(s + j) * ld + s;
    double nrm = 0.0;
    for (int i = lane; i < k; i += 32) {
        const double val = zhat[i] / ((d[i] - dorg) - m);
        uc[i] = val;
        nrm += val * val;
    }
    nrm = warp_sum(nrm);
    const double inv = 1.0 / sqrt(nrm);
    __syncwarp();
    for (int i = lane; i < k; i += 32) uc[i] *= inv;
}

// ------------------------------------------------------------------ merge: finish
// builds the GEMM problem descriptor, copies deflated columns, publishes eigenvalues.
__global__ void __launch_bounds__(256) merge_finish_kernel(DcArgs a, int level, int nmat, int cur,
                                                           GemmProblem *probs)
{
    const int nnode = 1 << level;
    const int mat = blockIdx.x / nnode, q = blockIdx.x % nnode;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    int32_t *islot = a.ibase + (size_t)mat * a.islot;
    const int ld = a.Npad;
    const int s = node_start(a.N, level, q), n = node_start(a.N, level, q + 1) - s;
    const int k = islot[a.oPerm + (size_t)IX_KP * ld + s];
    double *Qold = slot + (cur ? a.oQb : a.oQa);
    double *Qnew = slot + (cur ? a.oQa : a.oQb);
    double *aux = slot + a.oAux;
    double *lam = slot + a.oLam;
    const double *lamnew = aux + (size_t)AX_LAMNEW * ld + s;
    if (threadIdx.x == 0) {
        GemmProblem P;
        P.A = Qold + s;                    // rows s..s+n
        P.lda = ld;
        P.aidx = islot + a.oPerm + (size_t)IX_COLSRC * ld + s;
        const int32_t *posroot = islot + a.oPerm + (size_t)IX_POSROOT * ld + s;
        int jlo = 0, jn = k;
        if (level == 0 && a.wink > 0) {
            // roots keep their order, so the selected ones are a contiguous range
            const int w0 = a.win0[mat], w1 = w0 + a.wink;
            int lo = 0, hi = k;
            while (lo < hi) { const int m2 = (lo + hi) >> 1; if (posroot[m2] < w0) lo = m2 + 1; else hi = m2; }
            jlo = lo;
            hi = k;
            while (lo < hi) { const int m2 = (lo + hi) >> 1; if (posroot[m2] < w1) lo = m2 + 1; else hi = m2; }
            jn = lo - jlo;
        }
        P.B = slot + a.oU + (size_t)(s + jlo) * ld + s;
        P.ldb = ld;
        P.C = Qnew + (size_t)s * ld + s;   // cmap holds positions local to the block
        P.ldc = ld;
        P.cmap = posroot + jlo;            // local positions
        P.M = n; P.Ncols = jn; P.K = k;
        probs[blockIdx.x] = P;
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) lam[s + j] = lamnew[j];
}

// deflated eigenvectors are plain column copies Qold(:, src) -> Qnew(:, s + pos); one warp
// per column.  grid = (ceil(nmax / 8), nmat * nnode)
__global__ void __launch_bounds__(256) copy_deflated_kernel(DcArgs a, int level, int nmat, int cur)
{
    const int nnode = 1 << level;
    const int mat = blockIdx.y / nnode, q = blockIdx.y % nnode;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    int32_t *islot = a.ibase + (size_t)mat * a.islot;
    const int ld = a.Npad;
    const int s = node_start(a.N, level, q), n = node_start(a.N, level, q + 1) - s;
    const int k = islot[a.oPerm + (size_t)IX_KP * ld + s];
    const int jd = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (jd >= n - k) return;
    const double *Qold = slot + (cur ? a.oQb : a.oQa);
    double *Qnew = slot + (cur ? a.oQa : a.oQb);
    const int src_col = islot[a.oPerm + (size_t)IX_DEFSRC * ld + s + k + jd];
    const int pos = islot[a.oPerm + (size_t)IX_DEFPOS * ld + s + jd];
    if (level == 0 && a.wink > 0) {
        const int w0 = a.win0[mat];
        if (pos < w0 || pos >= w0 + a.wink) return;
    }
    const double *src = Qold + (size_t)src_col * ld + s;
    double *dst = Qnew + (size_t)(s + pos) * ld + s;
    for (int r = lane; r < n; r += 32) dst[r] = src[r];
}

__global__ void zero_kernel(double *base, size_t slot, size_t off, size_t count, int nmat)
{
    const size_t total = count * (size_t)nmat;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t m = i / count, e = i % count;
        base[m * slot + off + e] = 0.0;
    }
}

}  // namespace

int eig_stedc(const EigSlots &S, int nmat, cudaStream_t st, int *launches, int sel_k,
              double sigma, int32_t *d_sel0)
{
    const EigLayout &L = S.L;
    static_assert(sizeof(GemmProblem) <= EIG_PROB_BYTES, "GemmProblem too large");
    DcArgs a;
    a.N = L.N; a.Npad = L.Npad;
    int D = 0;
    while (((L.N + (1 << D) - 1) >> D) > LEAF) ++D;
    a.D = D;
    if ((1 << D) / 2 > EIG_MAX_NODES) { arcb200_set_error("stedc: N too large"); return -1; }
    a.base = S.dbl; a.slot = L.slot_doubles;
    a.ibase = S.ints; a.islot = L.slot_ints;
    // ping-pong so that the final eigenvectors land in Q1
    const bool swap = (D & 1);
    a.oQa = swap ? L.oQ2 : L.oQ1;
    a.oQb = swap ? L.oQ1 : L.oQ2;
    a.oU = L.oU;
    a.oD = L.oD; a.oE = L.oE; a.oLam = L.oLam; a.oAux = L.oAux;
    a.oPerm = L.oPerm; a.oIdx = L.oIdx;
    const bool window = sel_k > 0 && sel_k < L.N && d_sel0 != nullptr;
    a.win0 = d_sel0;
    a.wink = window ? sel_k : 0;
    int nl = 0;
    const size_t nn = (size_t)L.Npad * L.Npad;
    zero_kernel<<<148 * 8, 256, 0, st>>>(S.dbl, L.slot_doubles, L.oQ1, nn, nmat);
    zero_kernel<<<148 * 8, 256, 0, st>>>(S.dbl, L.slot_doubles, L.oQ2, nn, nmat);
    tear_kernel<<<nmat, 256, 0, st>>>(a, nmat);
    const long long nleaves = (long long)nmat << D;
    leaf_kernel<<<(unsigned)((nleaves + LEAF_WARPS - 1) / LEAF_WARPS), LEAF_WARPS * 32, 0, st>>>(a, nmat);
    nl += 4;
    GemmProblem *probs = (GemmProblem *)S.gemm_probs;
    int cur = 0;
    for (int level = D - 1; level >= 0; --level) {
        const int nnode = 1 << level;
        const int nmax = (L.N + nnode - 1) / nnode + 1;
        const unsigned gy = (unsigned)(nmat * nnode);
        if (gy > 65535u) { arcb200_set_error("stedc: batch*nodes exceeds grid limit"); return -1; }
        merge_prep_kernel<<<gy, PREP_THREADS, 0, st>>>(a, level, nmat, cur);
        const dim3 g1((nmax + 127) / 128, gy);
        secular_kernel<<<g1, 128, 0, st>>>(a, level, nmat);
        zhat_kernel<<<g1, 128, 0, st>>>(a, level, nmat);
        order_kernel<<<g1, 128, 0, st>>>(a, level, nmat);
        if (level == 0 && window) {
            // final spectrum is known now: pick the k eigenvalues nearest sigma, then build
            // only those eigenvectors (U columns, GEMM columns, deflated copies)
            window_kernel<<<(nmat + 127) / 128, 128, 0, st>>>(
                L.N, S.dbl, L.slot_doubles, L.oAux + (size_t)AX_LAMNEW * L.Npad, sel_k, sigma, d_sel0, nmat);
            ++nl;
        }
        umat_kernel<<<dim3((nmax + 3) / 4, gy), 128, 0, st>>>(a, level, nmat);
        merge_finish_kernel<<<gy, 256, 0, st>>>(a, level, nmat, cur, probs);
        copy_deflated_kernel<<<dim3((nmax + 7) / 8, gy), 256, 0, st>>>(a, level, nmat, cur);
        const int tiles = (nmax + 63) / 64;
        grouped_dgemm_kernel<<<dim3(tiles * tiles, gy), 256, 0, st>>>(probs, tiles);
        nl += 8;
        cur ^= 1;
    }
    if (D == 0 && sel_k > 0 && d_sel0 != nullptr) {
        window_kernel<<<(nmat + 127) / 128, 128, 0, st>>>(L.N, S.dbl, L.slot_doubles, L.oLam,
                                                           sel_k < L.N ? sel_k : L.N, sigma, d_sel0, nmat);
        ++nl;
    } else if (!window && sel_k > 0 && d_sel0 != nullptr) {
        window_kernel<<<(nmat + 127) / 128, 128, 0, st>>>(L.N, S.dbl, L.slot_doubles, L.oLam,
                                                           L.N, sigma, d_sel0, nmat);
        ++nl;
    }
    ARCB200_LAUNCH_CHECK("stedc kernels");
    if (launches) *launches += nl;
    return 0;
}
