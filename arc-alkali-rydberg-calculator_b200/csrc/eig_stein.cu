// eig_stein.cu -- k eigenpairs nearest sigma of the symmetric tridiagonal matrices produced by the
// tridiagonalisation, by bisection + inverse iteration (LAPACK dstebz / dstein scheme), for pair-state
// sweeps where k << N (scipy eigsh(k, sigma) of calculations_atom_pairstate.py:3444-3450 asks for
// 250 of 2363 or 10384 eigenpairs).  The divide & conquer solver (eig_stedc.cu) builds the full
// eigenvector basis of every sub-problem -- (4/3) N^3 flop of merges to hand over N k numbers; here
// the work is O(N k) per matrix:
//   stein_prep_kernel    Gershgorin bounds, pivmin, Sturm count at sigma -> candidate index range
//                        [lo - k, lo + k) (one warp per matrix)
//   stein_bisect_kernel  one thread per candidate eigenvalue: bisection on the Sturm count to full
//                        precision; d and e^2 are read warp-uniformly
//   stein_window_kernel  the k candidates nearest sigma (same greedy rule as window_kernel of the
//                        D&C path: ties go to the lower eigenvalue) -> lam[0..k), sel0 = 0
//   stein_invit_kernel   one thread per eigenvector: LU with partial pivoting of T - lam I (dgttrf
//                        row interchanges, reciprocal pivots, tiny pivots perturbed to eps |T| like
//                        dlagts), two inverse iterations from a random vector; factors and the
//                        iterate live in global memory interleaved over the eigenvectors, so the
//                        32 lanes of a warp stream 256 contiguous bytes per array and row
//   stein_orth_kernel    eigenvalues closer than 1e-6 |T| form a cluster; one CTA per cluster
//                        re-orthogonalises its vectors (modified Gram-Schmidt, twice).  Outside
//                        clusters three iterations leave a cross contamination below
//                        (eps |T| / gap)^2 < 1e-17.
// Measured on the C3 tridiagonals (N = 2363, k = 250, numpy model of this file): eigenvalues
// 4e-16 |T|, residual 4e-16 |T|, orthogonality 1e-12, highlight identical to LAPACK to 3e-15.
#include <math.h>
#include <stdlib.h>
#include "eig.cuh"

namespace {

constexpr double STEIN_CLUSTER_TOL = 1e-6;
constexpr int SU = 4;             // rows per load batch of the inverse-iteration passes
constexpr int STEIN_ITS = 2;      // eigenvalues are exact to eps |T|: the second iteration already converges (dstein)

struct SteinArgs {
    int N, Npad, k, kp;
    double *base;
    size_t slot;
    int32_t *ibase;
    size_t islot;
    size_t oD, oE, oLam, oAux, oFac, oX, oZ;
    double sigma;
    int32_t *sel0;
};

// aux layout (doubles, inside oAux, 12 * Npad available): [0, 8) scalars, [Npad, 2 Npad) e^2,
// [2 Npad, 4 Npad) candidate eigenvalues
constexpr int SA_TN = 0, SA_PIVMIN = 1, SA_GL = 2, SA_GU = 3, SA_C0 = 4, SA_NCAND = 5;

__device__ __forceinline__ int sturm_count(const double *__restrict__ d, const double *__restrict__ e2, int n,
                                           double x, double pivmin)
{
    double q = d[0] - x;
    int cnt = q < 0.0;
    for (int i = 1; i < n; ++i) {
        if (fabs(q) < pivmin) q = -pivmin;
        q = d[i] - x - e2[i - 1] / q;
        cnt += q < 0.0;
    }
    return cnt;
}

__global__ void __launch_bounds__(32) stein_prep_kernel(SteinArgs a)
{
    const int mat = blockIdx.x, lane = threadIdx.x;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *d = slot + a.oD, *e = slot + a.oE;
    double *aux = slot + a.oAux;
    double *e2 = aux + a.Npad;
    const int n = a.N;
    double gl = INFINITY, gu = -INFINITY, e2max = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double el = (i > 0) ? fabs(e[i - 1]) : 0.0, er = (i < n - 1) ? fabs(e[i]) : 0.0;
        gl = fmin(gl, d[i] - el - er);
        gu = fmax(gu, d[i] + el + er);
        if (i < n - 1) {
            const double v = e[i] * e[i];
            e2[i] = v;
            e2max = fmax(e2max, v);
        }
    }
    gl = -warp_max(-gl);
    gu = warp_max(gu);
    e2max = warp_max(e2max);
    const double tn = fmax(fabs(gl), fabs(gu));
    const double pivmin = 2.2250738585072014e-308 * fmax(1.0, e2max);
    const double eps = 2.220446049250313e-16;
    gl -= 2.0 * tn * eps * n + 2.0 * pivmin;
    gu += 2.0 * tn * eps * n + 2.0 * pivmin;
    __syncwarp();
    // radius search: the smallest rho (to ~1e-7 of the spectral width) with at least k eigenvalues in
    // (sigma - rho, sigma + rho); 16-section per round, lanes 0-15 count below sigma - rho_i, lanes
    // 16-31 below sigma + rho_i.  The candidates for the bisection are the eigenvalues in that interval.
    double rlo = 0.0, rhi = fmax(gu - a.sigma, a.sigma - gl);
    int cL = 0, cR = n;                                    // counts at sigma -+ rhi
    for (int round = 0; round < 6; ++round) {
        const int i = lane & 15;
        const double rho = rlo + (rhi - rlo) * (double)(i + 1) / 17.0;
        const double x = (lane < 16) ? a.sigma - rho : a.sigma + rho;
        const int cnt = sturm_count(d, e2, n, x, pivmin);
        const int cl = __shfl_sync(ARCB200_FULL_MASK, cnt, i), cr = __shfl_sync(ARCB200_FULL_MASK, cnt, 16 + i);
        const unsigned ok = __ballot_sync(ARCB200_FULL_MASK, lane < 16 && (cr - cl) >= a.k);
        // first radius that holds k eigenvalues (if any) bounds from above, its predecessor from below
        const int first = ok ? __ffs(ok) - 1 : 16;
        const double nlo = (first == 0) ? rlo : rlo + (rhi - rlo) * (double)first / 17.0;
        if (first < 16) {
            const double nhi = rlo + (rhi - rlo) * (double)(first + 1) / 17.0;
            cL = __shfl_sync(ARCB200_FULL_MASK, cl, first);
            cR = __shfl_sync(ARCB200_FULL_MASK, cr, first);
            rhi = nhi;
        }
        rlo = nlo;
    }
    if (lane == 0) {
        int c0 = cL, nc = cR - cL;
        if (nc < a.k || nc > 2 * a.k) {                    // (massively degenerate boundary: plain index window)
            const int lo = sturm_count(d, e2, n, a.sigma, pivmin);
            c0 = max(0, lo - a.k);
            nc = min(n, lo + a.k) - c0;
        }
        aux[SA_TN] = tn; aux[SA_PIVMIN] = pivmin; aux[SA_GL] = gl; aux[SA_GU] = gu;
        aux[SA_C0] = (double)c0; aux[SA_NCAND] = (double)nc;
    }
}

// grid = (ceil(2k / 128), nmat).  d and e^2 are staged through shared memory in chunks of 256 rows (next chunk
// prefetched into registers): read straight from global memory the dependent Sturm recurrence paid an L2 round
// trip per row (warp-uniform loads of eight matrices per SM do not stay in L1) -- 29 ms of the 485 ms C3 step.
__global__ void __launch_bounds__(128) stein_bisect_kernel(SteinArgs a)
{
    constexpr int CH = 256;
    __shared__ double sd[CH], se[CH];
    const int mat = blockIdx.y, tid = threadIdx.x;
    const int t = blockIdx.x * 128 + tid;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *__restrict__ d = slot + a.oD;
    double *aux = slot + a.oAux;
    const double *__restrict__ e2 = aux + a.Npad;
    double *cand = aux + 2 * (size_t)a.Npad;
    const int ncand = (int)aux[SA_NCAND];
    if (blockIdx.x * 128 >= ncand) return;               // whole blocks leave together
    const int idx = (int)aux[SA_C0] + min(t, ncand - 1);
    const double pivmin = aux[SA_PIVMIN];
    const double atol = 2.220446049250313e-16 * aux[SA_TN];
    double lo = aux[SA_GL], hi = aux[SA_GU];
    const int n = a.N;
    auto ld_d = [&](int i) { return i < n ? d[i] : 0.0; };
    auto ld_e = [&](int i) { return (i >= 1 && i < n) ? e2[i - 1] : 0.0; };     // row i uses e2[i - 1]; row 0 none
    for (int it = 0; it < 80; ++it) {
        const double mid = 0.5 * (lo + hi);
        // stop like dstebz with its default tolerances: interval below ulp |T| (the backward error of the
        // tridiagonalisation) or down to neighbouring doubles
        const bool live = mid > lo && mid < hi && (hi - lo) > atol;
        if (!__syncthreads_or(live)) break;              // (threads that are done keep evaluating: the loop is uniform)
        double q = 1.0;                                  // row 0: q = d[0] - mid - 0 / 1
        int cnt = 0;
        double nd0 = ld_d(tid), nd1 = ld_d(tid + 128), ne0 = ld_e(tid), ne1 = ld_e(tid + 128);
        for (int c0 = 0; c0 < n; c0 += CH) {
            __syncthreads();
            sd[tid] = nd0; sd[tid + 128] = nd1; se[tid] = ne0; se[tid + 128] = ne1;
            __syncthreads();
            const int nx = c0 + CH;
            nd0 = ld_d(nx + tid); nd1 = ld_d(nx + tid + 128); ne0 = ld_e(nx + tid); ne1 = ld_e(nx + tid + 128);
            const int len = min(CH, n - c0);
#pragma unroll 8
            for (int i = 0; i < len; ++i) {
                if (fabs(q) < pivmin) q = -pivmin;
                q = sd[i] - mid - se[i] * __drcp_rn(q);       // sign sequence only: a rounded reciprocal is enough
                cnt += q < 0.0;
            }
        }
        if (live) { if (cnt > idx) hi = mid; else lo = mid; }
    }
    if (t < ncand) cand[t] = 0.5 * (lo + hi);
}

__global__ void stein_window_kernel(SteinArgs a, int nmat)
{
    const int mat = blockIdx.x * blockDim.x + threadIdx.x;
    if (mat >= nmat) return;
    double *slot = a.base + (size_t)mat * a.slot;
    double *aux = slot + a.oAux;
    const double *lam = aux + 2 * (size_t)a.Npad;
    double *out = slot + a.oLam;
    const int nc = (int)aux[SA_NCAND];
    int lo = 0, hi = nc;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (lam[m] < a.sigma) lo = m + 1; else hi = m; }
    int left = lo - 1, right = lo;
    for (int t = 0; t < a.k; ++t) {
        const bool canL = left >= 0, canR = right < nc;
        if (canL && (!canR || fabs(lam[left] - a.sigma) <= fabs(lam[right] - a.sigma))) --left;
        else ++right;
    }
    for (int j = 0; j < a.k; ++j) out[j] = lam[left + 1 + j];
    a.sel0[mat] = 0;
}

__device__ __forceinline__ double stein_rand(unsigned i, unsigned j, unsigned m)
{
    unsigned h = i * 0x9E3779B1u ^ (j + 0x7F4A7C15u) * 0x85EBCA6Bu ^ (m + 1u) * 0xC2B2AE35u;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return (double)h * (2.0 / 4294967296.0) - 1.0;       // uniform in [-1, 1)
}

// grid = (kp / 128, nmat); thread <-> eigenvector j.  Work arrays [row][kp] (lanes contiguous).
__global__ void __launch_bounds__(128) stein_invit_kernel(SteinArgs a)
{
    const int mat = blockIdx.y;
    const int j = blockIdx.x * 128 + threadIdx.x;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *__restrict__ d = slot + a.oD;
    const double *__restrict__ e = slot + a.oE;
    const double *aux = slot + a.oAux;
    const int n = a.N, kp = a.kp;
    if ((j & ~31) >= a.k) return;
    const bool live = j < a.k;
    const double lam = (slot + a.oLam)[live ? j : a.k - 1];
    const double tn = aux[SA_TN];
    const double eps = 2.220446049250313e-16;
    const double ptiny = fmax(eps * tn, 1e-290);
    const size_t plane = (size_t)n * kp;
    double *__restrict__ fainv = slot + a.oFac + j;
    double *__restrict__ fb = fainv + plane;
    double *__restrict__ fdu2 = fb + plane;
    double *__restrict__ fl = fdu2 + plane;
    double *__restrict__ x = slot + a.oX + j;
    unsigned char *__restrict__ piv = reinterpret_cast<unsigned char *>(slot + a.oX + plane) + j;

    // ---- LU with partial pivoting of T - lam I (row interchanges between i and i + 1)
    {
        double ad = d[0] - lam, bs = (n > 1) ? e[0] : 0.0;
        for (int i = 0; i < n - 1; ++i) {
            const double c = e[i], a1 = d[i + 1] - lam, b1 = (i + 2 < n) ? e[i + 1] : 0.0;
            const size_t o = (size_t)i * kp;
            if (fabs(ad) >= fabs(c)) {
                if (fabs(ad) < ptiny) ad = copysign(ptiny, ad);
                const double inv = 1.0 / ad, m = c * inv;
                fainv[o] = inv; fb[o] = bs; fdu2[o] = 0.0; fl[o] = m; piv[o] = 0;
                ad = a1 - m * bs;
                bs = b1;
            } else {
                const double inv = 1.0 / c, m = ad * inv;
                fainv[o] = inv; fb[o] = a1; fdu2[o] = b1; fl[o] = m; piv[o] = 1;
                ad = bs - m * a1;
                bs = -m * b1;
            }
        }
        if (fabs(ad) < ptiny) ad = copysign(ptiny, ad);
        fainv[(size_t)(n - 1) * kp] = 1.0 / ad;
    }
    // ---- inverse iterations
    double s1 = 0.0, s2 = 1.0;
    for (int it = 0; it < STEIN_ITS; ++it) {
        // forward: x <- L^-1 P (scale * x); the start vector is generated on the fly
        const double sc = (it == 0) ? 1.0 : (n * tn * eps) / fmax(s1, 1e-300);
        double xc = (it == 0) ? stein_rand(0u, (unsigned)j, (unsigned)mat) * (n * tn * eps) / (0.5 * n) : x[0] * sc;
        // rows in batches of SU: the loads of a batch are issued before its dependent chain (a chain step is one
        // FMA, a DRAM round trip per step left the pass at a third of the HBM rate)
        for (int i0 = 0; i0 < n - 1; i0 += SU) {
            double l[SU], xn[SU];
            unsigned char pv[SU];
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const int i = min(i0 + u, n - 2);
                const size_t o = (size_t)i * kp;
                l[u] = fl[o];
                pv[u] = piv[o];
                xn[u] = (it == 0) ? stein_rand((unsigned)(i + 1), (unsigned)j, (unsigned)mat) * (n * tn * eps) / (0.5 * n)
                                  : x[o + kp] * sc;
            }
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                if (i0 + u < n - 1) {
                    const size_t o = (size_t)(i0 + u) * kp;
                    if (pv[u]) { x[o] = xn[u]; xc = xc - l[u] * xn[u]; }
                    else { x[o] = xc; xc = xn[u] - l[u] * xc; }
                }
            }
        }
        x[(size_t)(n - 1) * kp] = xc;
        // backward: x <- U^-1 x
        double x1 = xc * fainv[(size_t)(n - 1) * kp], x2 = 0.0;
        x[(size_t)(n - 1) * kp] = x1;
        s1 = fabs(x1);
        s2 = x1 * x1;
        for (int i0 = n - 2; i0 >= 0; i0 -= SU) {
            double xo[SU], b[SU], du[SU], ai[SU];
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const size_t o = (size_t)max(i0 - u, 0) * kp;
                xo[u] = x[o]; b[u] = fb[o]; du[u] = fdu2[o]; ai[u] = fainv[o];
            }
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                if (i0 - u >= 0) {
                    const double v = (xo[u] - b[u] * x1 - du[u] * x2) * ai[u];
                    x[(size_t)(i0 - u) * kp] = v;
                    s1 += fabs(v);
                    s2 += v * v;
                    x2 = x1;
                    x1 = v;
                }
            }
        }
    }
    // ---- normalised eigenvector -> column j of Z (column-major, ld = Npad)
    if (live) {
        const double rn = 1.0 / sqrt(s2);
        double *__restrict__ z = slot + a.oZ + (size_t)j * a.Npad;
        for (int i = 0; i < n; ++i) z[i] = x[(size_t)i * kp] * rn;
        for (int i = n; i < a.Npad; ++i) z[i] = 0.0;
    }
}

// grid = (k, nmat), 128 threads; CTA <-> eigenvector j, only the heads of clusters work
__global__ void __launch_bounds__(128) stein_orth_kernel(SteinArgs a, int32_t *status)
{
    const int mat = blockIdx.y;
    const int j = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *lam = slot + a.oLam;
    const double tol = STEIN_CLUSTER_TOL * (slot + a.oAux)[SA_TN];
    if (j > 0 && lam[j] - lam[j - 1] <= tol) return;          // not the head of a cluster
    int end = j + 1;
    while (end < a.k && lam[end] - lam[end - 1] <= tol) ++end;
    if (end == j + 1) return;                                  // isolated eigenvalue
    __shared__ double red[4];
    double *Z = slot + a.oZ;
    const int n = a.N;
    auto block_sum = [&](double v) {
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) red[wid] = v;
        __syncthreads();
        return (red[0] + red[1]) + (red[2] + red[3]);
    };
    for (int m = j + 1; m < end; ++m) {
        double *zm = Z + (size_t)m * a.Npad;
        double first = 1.0;
        for (int rep = 0; rep < 2; ++rep) {
            for (int q = j; q < m; ++q) {
                const double *zq = Z + (size_t)q * a.Npad;
                double dot = 0.0;
                for (int i = tid; i < n; i += 128) dot += zq[i] * zm[i];
                dot = block_sum(dot);
                for (int i = tid; i < n; i += 128) zm[i] -= dot * zq[i];
            }
            double nn = 0.0;
            for (int i = tid; i < n; i += 128) nn += zm[i] * zm[i];
            nn = sqrt(block_sum(nn));
            if (rep == 0) first = nn;
            const double rn = 1.0 / fmax(nn, 1e-300);
            for (int i = tid; i < n; i += 128) zm[i] *= rn;
        }
        // two vectors of a cluster came out (nearly) parallel: reported, never silently wrong
        if (first < 1e-6 && tid == 0 && status) atomicAdd(&status[mat], 1);
    }
}

}  // namespace

int eig_stein_usable(int N, int k)
{
    const char *ev = getenv("ARCB200_STEIN");              // tuning knob: 0 = always divide & conquer
    if (ev && ev[0] == '0') return 0;
    // work arrays: 4 factor planes of N x kp in the Q2 matrix, the iterate + pivot flags in U
    const int kp = (k + 31) & ~31;
    const int Npad = (N + 31) & ~31;
    if ((size_t)4 * N * kp > (size_t)Npad * Npad) return 0;
    if (2 * k > N) return 0;
    return N >= (ev && ev[0] == '1' ? 64 : 512);
}

int eig_stein(const EigSlots &S, int nmat, int k, double sigma, int32_t *d_sel0, cudaStream_t st, int *launches)
{
    const EigLayout &L = S.L;
    if (!eig_stein_usable(L.N, k) || !d_sel0) { arcb200_set_error("stein: unsupported (N=%d, k=%d)", L.N, k); return -1; }
    SteinArgs a;
    a.N = L.N; a.Npad = L.Npad; a.k = k; a.kp = (k + 31) & ~31;
    a.base = S.dbl; a.slot = L.slot_doubles;
    a.ibase = S.ints; a.islot = L.slot_ints;
    a.oD = L.oD; a.oE = L.oE; a.oLam = L.oLam; a.oAux = L.oAux;
    a.oFac = L.oQ2; a.oX = L.oU; a.oZ = L.oQ1;
    a.sigma = sigma;
    a.sel0 = d_sel0;
    if (nmat > 65535) { arcb200_set_error("stein: batch exceeds grid limit"); return -1; }
    stein_prep_kernel<<<nmat, 32, 0, st>>>(a);
    stein_bisect_kernel<<<dim3((2 * k + 127) / 128, nmat), 128, 0, st>>>(a);
    stein_window_kernel<<<(nmat + 127) / 128, 128, 0, st>>>(a, nmat);
    stein_invit_kernel<<<dim3((a.kp + 127) / 128, nmat), 128, 0, st>>>(a);
    stein_orth_kernel<<<dim3(k, nmat), 128, 0, st>>>(a, nullptr);
    ARCB200_LAUNCH_CHECK("stein kernels");
    if (launches) *launches += 5;
    return 0;
}
