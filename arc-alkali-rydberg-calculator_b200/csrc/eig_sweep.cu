// eig_sweep.cu -- subsystem (2) Hamiltonian assembly, the fused result epilogue and the
// C-ABI sweep entry points (include/arcb200.h).
//
//   assemble  H(F) = diag(h0) + F V           (calculations_atom_single.py:984)
//             H(R) = diag(h0) + sum_o M_o / (R 1e-6)^(3+o)   (calculations_atom_pairstate.py:3435-3440)
//             written straight into the eigensolver's batch workspace (lower-triangle
//             32x32 tiles only, coalesced; V / M_o are read through L2 once per batch)
//   solve     eig_sytrd | eig_sy2sb + eig_sb2st -> eig_stedc (all eigenpairs) | eig_stein (k nearest sigma)
//             -> eig_ormtr | eig_q2_apply
//   epilogue  y, highlight and the top-k composition per eigenvector, so eigenvectors
//             never leave the GPU (calculations_atom_single.py:988-1021, 1395-1416;
//             calculations_atom_pairstate.py:3494-3520).
#include <stdlib.h>
#include "eig.cuh"

namespace {

struct WorkView {
    EigSlots S;
    int32_t *sel0;      // per slot: first selected eigenvector
};

size_t work_bytes(const EigLayout &L, int batch)
{
    size_t b = 256;
    b += (size_t)batch * L.slot_doubles * sizeof(double);
    b += (size_t)batch * L.slot_ints * sizeof(int32_t);
    b += (size_t)batch * EIG_MAX_NODES * EIG_PROB_BYTES;
    b += (size_t)batch * sizeof(int32_t) + 256;
    return (b + 255) & ~(size_t)255;          // whatever follows the workspace stays 256-byte aligned
}

WorkView carve(void *d_work, const EigLayout &L, int batch)
{
    WorkView W;
    uintptr_t p = ((uintptr_t)d_work + 255) & ~(uintptr_t)255;
    W.S.L = L;
    W.S.batch = batch;
    W.S.dbl = (double *)p;
    p += (size_t)batch * L.slot_doubles * sizeof(double);
    W.S.ints = (int32_t *)p;
    p += (size_t)batch * L.slot_ints * sizeof(int32_t);
    W.S.gemm_probs = (void *)p;
    p += (size_t)batch * EIG_MAX_NODES * EIG_PROB_BYTES;
    W.sel0 = (int32_t *)p;
    return W;
}

// ---------------------------------------------------------------- assembly
// grid = (ntiles_lower, nmat); one 32x32 tile per CTA of 256 threads (8 columns / pass)
__global__ void __launch_bounds__(256) assemble_stark_kernel(
    int N, int Npad, double *base, size_t slot, size_t oA, const double *__restrict__ h0,
    const double *__restrict__ V, const double *__restrict__ F, int f0)
{
    // tile index -> (I, J), J <= I
    const int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    while (I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int mat = blockIdx.y;
    double *A = base + (size_t)mat * slot + oA;
    const double f = F[f0 + mat];
    const int r = 32 * I + (threadIdx.x & 31);
    for (int cc = threadIdx.x >> 5; cc < 32; cc += 8) {
        const int c = 32 * J + cc;
        double x = 0.0;
        if (r < N && c < N) {
            // mat1 + mat2*F, elementwise like numpy (mat2 symmetric: V[c][r] == V[r][c])
            x = __dmul_rn(V[(size_t)c * N + r], f);
            if (r == c) x = __dadd_rn(h0[r], x);
        }
        A[r + (size_t)c * Npad] = x;
    }
}

__global__ void __launch_bounds__(256) assemble_pair_kernel(
    int N, int Npad, double *base, size_t slot, size_t oA, const double *__restrict__ h0,
    int norders, const double *const *__restrict__ M, const double *__restrict__ R, int p0)
{
    const int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    while (I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int mat = blockIdx.y;
    double *A = base + (size_t)mat * slot + oA;
    const double x1 = __dmul_rn(R[p0 + mat], 1.0e-6);
    const int r = 32 * I + (threadIdx.x & 31);
    for (int cc = threadIdx.x >> 5; cc < 32; cc += 8) {
        const int c = 32 * J + cc;
        double x = 0.0;
        if (r < N && c < N) {
            if (r == c) x = h0[r];
            double rX = __dmul_rn(__dmul_rn(x1, x1), x1);          // (R*1e-6)**3
            for (int o = 0; o < norders; ++o) {
                x = __dadd_rn(x, __ddiv_rn(M[o][(size_t)c * N + r], rX));   // m = m + matRX / rX
                rX = __dmul_rn(rX, x1);
            }
        }
        A[r + (size_t)c * Npad] = x;
    }
}

__global__ void scatter_coo_kernel(int N, int64_t nnz, const int32_t *__restrict__ rows,
                                   const int32_t *__restrict__ cols,
                                   const double *__restrict__ vals, double *__restrict__ dense)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x) {
        // scipy's csr_matrix((v,(r,c))) sums duplicate entries
        atomicAdd(&dense[(size_t)rows[i] * N + cols[i]], vals[i]);
    }
}

__global__ void load_dense_kernel(int N, int Npad, double *base, size_t slot, size_t oA,
                                  const double *__restrict__ src)
{
    const int mat = blockIdx.y;
    double *A = base + (size_t)mat * slot + oA;
    const double *S = src + (size_t)mat * N * N;
    const size_t total = (size_t)Npad * Npad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i % Npad), c = (int)(i / Npad);
        A[i] = (r < N && c < N) ? S[(size_t)c * N + r] : 0.0;
    }
}

__global__ void store_dense_kernel(int N, int Npad, const double *base, size_t slot, size_t oQ,
                                   size_t oLam, double *__restrict__ Zout, double *__restrict__ wout)
{
    const int mat = blockIdx.y;
    const double *Q = base + (size_t)mat * slot + oQ;
    const double *lam = base + (size_t)mat * slot + oLam;
    const size_t total = (size_t)N * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i % N), c = (int)(i / N);
        Zout[(size_t)mat * total + i] = Q[r + (size_t)c * Npad];
        if (i < (size_t)N) wout[(size_t)mat * N + i] = lam[i];
    }
}

// ---------------------------------------------------------------- epilogue
// one warp per eigenvector: eigenvalue, highlight, top-k composition
constexpr int MAXTOP = 8;
__global__ void __launch_bounds__(256) epilogue_kernel(
    int N, int Npad, const double *base, size_t slot, size_t oQ, size_t oLam,
    const int32_t *__restrict__ sel0, int nsel, int idx0, const double *__restrict__ drive_w,
    int topk, double contrib_max, double *__restrict__ y, double *__restrict__ hl,
    double *__restrict__ comp_c, int32_t *__restrict__ comp_i, int point0)
{
    const int mat = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= nsel) return;
    const int s0 = sel0 ? sel0[mat] : 0;
    const double *q = base + (size_t)mat * slot + oQ + (size_t)(s0 + i) * Npad;
    const double *lam = base + (size_t)mat * slot + oLam;
    const size_t orow = (size_t)(point0 + mat) * nsel + i;
    double best_a[MAXTOP];
    int best_i[MAXTOP];
#pragma unroll
    for (int t = 0; t < MAXTOP; ++t) { best_a[t] = -1.0; best_i[t] = -1; }
    double hacc = 0.0;
    for (int r = lane; r < N; r += 32) {
        const double c = q[r];
        const double a = fabs(c);
        if (drive_w) hacc += drive_w[r] * (c * c);
        if (a > best_a[MAXTOP - 1] && topk > 0) {
            // insertion into the lane-local sorted list
            double ca = a; int ci = r;
#pragma unroll
            for (int t = 0; t < MAXTOP; ++t) {
                if (ca > best_a[t]) {
                    const double ta = best_a[t]; const int ti = best_i[t];
                    best_a[t] = ca; best_i[t] = ci; ca = ta; ci = ti;
                }
            }
        }
    }
    if (drive_w) hacc = warp_sum(hacc);
    if (lane == 0) {
        y[orow] = lam[s0 + i];
        hl[orow] = drive_w ? hacc : q[idx0] * q[idx0];
    }
    // warp merge: pop the global maximum topk times
    double total = 0.0;
    const int kk = min(topk, MAXTOP);
    for (int t = 0; t < kk; ++t) {
        const bool go = total < contrib_max;
        double a = best_a[0];
        int src = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double oa = __shfl_xor_sync(ARCB200_FULL_MASK, a, o);
            const int os = __shfl_xor_sync(ARCB200_FULL_MASK, src, o);
            if (oa > a || (oa == a && os < src)) { a = oa; src = os; }
        }
        const int widx = __shfl_sync(ARCB200_FULL_MASK, best_i[0], src);
        if (lane == src) {
#pragma unroll
            for (int u = 0; u < MAXTOP - 1; ++u) { best_a[u] = best_a[u + 1]; best_i[u] = best_i[u + 1]; }
            best_a[MAXTOP - 1] = -1.0; best_i[MAXTOP - 1] = -1;
        }
        if (lane == 0) {
            const bool valid = go && widx >= 0;
            comp_c[orow * topk + t] = valid ? q[widx] : 0.0;
            comp_i[orow * topk + t] = valid ? widx : -1;
        }
        if (go && widx >= 0) total += a * a;
    }
    if (lane == 0)
        for (int t = kk; t < topk; ++t) { comp_c[orow * topk + t] = 0.0; comp_i[orow * topk + t] = -1; }
}

// topk > MAXTOP (upTo = -1 or upTo > 9 in the reference's _stateComposition2,
// calculations_atom_single.py:1399-1407): the whole eigenvector sorted by |c| descending.
// One CTA per eigenvector: bitonic sort of (|c|, index) keys in shared memory (N padded to a
// power of two), then the first topk entries are written while the running sum of c^2 is below
// contrib_max (the reference's loop condition; upTo = -1 passes contrib_max = 2 = no cut).
// Ties in |c| are ordered by ascending index (numpy's heapsort leaves them unspecified).
__global__ void __launch_bounds__(256) full_composition_kernel(
    int N, int Npad, int P2, const double *base, size_t slot, size_t oQ,
    const int32_t *__restrict__ sel0, int nsel, int topk, double contrib_max,
    double *__restrict__ comp_c, int32_t *__restrict__ comp_i, int point0)
{
    extern __shared__ double fc_sm[];
    double *key = fc_sm;                                   // [P2]
    int32_t *idx = reinterpret_cast<int32_t *>(fc_sm + P2);   // [P2]
    __shared__ double part[256];
    const int mat = blockIdx.y, i = blockIdx.x, tid = threadIdx.x;
    const int s0 = sel0 ? sel0[mat] : 0;
    const double *q = base + (size_t)mat * slot + oQ + (size_t)(s0 + i) * Npad;
    const size_t orow = (size_t)(point0 + mat) * nsel + i;
    for (int t = tid; t < P2; t += 256) {
        key[t] = (t < N) ? fabs(q[t]) : -1.0;
        idx[t] = (t < N) ? t : 0x7fffffff;
    }
    __syncthreads();
    // descending by key, ascending by index on ties
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < P2; t += 256) {
                const int u = t ^ j;
                if (u > t) {
                    const double a = key[t], b = key[u];
                    const int ia = idx[t], ib = idx[u];
                    const bool a_first = (a > b) || (a == b && ia < ib);   // a belongs before b
                    const bool up = ((t & k) == 0);
                    if (up != a_first) { key[t] = b; key[u] = a; idx[t] = ib; idx[u] = ia; }
                }
            }
            __syncthreads();
        }
    }
    // exclusive prefix sums of c^2 in sorted order (chunk per thread, then across threads)
    const int chunk = (N + 255) / 256;
    const int lo = min(N, tid * chunk), hi = min(N, lo + chunk);
    double s = 0.0;
    for (int t = lo; t < hi; ++t) s += key[t] * key[t];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        double run = 0.0;
        for (int t = 0; t < 256; ++t) { const double v = part[t]; part[t] = run; run += v; }
    }
    __syncthreads();
    double run = part[tid];
    for (int t = lo; t < hi; ++t) {
        if (t < topk) {
            const bool valid = run < contrib_max;
            comp_c[orow * topk + t] = valid ? q[idx[t]] : 0.0;
            comp_i[orow * topk + t] = valid ? idx[t] : -1;
        }
        run += key[t] * key[t];
    }
    for (int t = N + tid; t < topk; t += 256) { comp_c[orow * topk + t] = 0.0; comp_i[orow * topk + t] = -1; }
}

static int launch_full_composition(int N, int Npad, const double *base, size_t slot, size_t oQ,
                                   const int32_t *sel0, int nsel, int nmat, int topk, double contrib_max,
                                   double *comp_c, int32_t *comp_i, int point0, cudaStream_t st)
{
    int P2 = 1;
    while (P2 < N) P2 <<= 1;
    const size_t smem = (size_t)P2 * (sizeof(double) + sizeof(int32_t));
    if (smem > 200 * 1024) { arcb200_set_error("full composition: N=%d too large", N); return -1; }
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(full_composition_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    full_composition_kernel<<<dim3(nsel, nmat), 256, smem, st>>>(N, Npad, P2, base, slot, oQ, sel0, nsel, topk,
                                                                 contrib_max, comp_c, comp_i, point0);
    return 0;
}

// ---------------------------------------------------------------- adiabatic overlaps
// O[t][i][j] = ( v_t,i . v_(t-1),j )^2 between the selected eigenvectors of consecutive sweep
// points (calculations_atom_pairstate.py:3462-3468).  One CTA per 16x16 output tile.
__global__ void __launch_bounds__(256) overlap_kernel(
    int N, int Npad, const double *base, size_t slot, size_t oQ, const int32_t *__restrict__ sel0,
    int k, const double *__restrict__ prevZ /* N x k, ld N, for matrix 0 (or null) */,
    double *__restrict__ ov, int point0)
{
    const int mat = blockIdx.z;
    const double *Zc = base + (size_t)mat * slot + oQ + (size_t)sel0[mat] * Npad;
    const double *Zp;
    int ldp;
    if (mat == 0) {
        if (!prevZ) { Zp = Zc; ldp = Npad; }          // first point: overlap with itself
        else { Zp = prevZ; ldp = N; }
    } else {
        Zp = base + (size_t)(mat - 1) * slot + oQ + (size_t)sel0[mat - 1] * Npad;
        ldp = Npad;
    }
    __shared__ double A[16][65], B[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.x * 16, j0 = blockIdx.y * 16;
    double acc = 0.0;
    for (int r0 = 0; r0 < N; r0 += 64) {
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            const int c = t >> 6, r = r0 + (t & 63);
            A[c][t & 63] = (r < N && i0 + c < k) ? Zc[r + (size_t)(i0 + c) * Npad] : 0.0;
            B[c][t & 63] = (r < N && j0 + c < k) ? Zp[r + (size_t)(j0 + c) * ldp] : 0.0;
        }
        __syncthreads();
#pragma unroll 16
        for (int r = 0; r < 64; ++r) acc += A[tx][r] * B[ty][r];
        __syncthreads();
    }
    if (i0 + tx < k && j0 + ty < k)
        ov[((size_t)(point0 + mat) * k + (i0 + tx)) * k + (j0 + ty)] = acc * acc;
}

__global__ void save_window_kernel(int N, int Npad, const double *base, size_t slot, size_t oQ,
                                   const int32_t *__restrict__ sel0, int k, int mat, double *out)
{
    const double *Z = base + (size_t)mat * slot + oQ + (size_t)sel0[mat] * Npad;
    const size_t total = (size_t)N * k;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x)
        out[e] = Z[(e % N) + (e / N) * Npad];
}

// Floquet (Shirley) transition probabilities: for every atomic basis state a
//   P[f][a] = sum_k sum_i |Z[k*dim0 + a, i]|^2 * hl[f][i],   hl[f][i] = |Z[target, i]|^2
// (calculations_atom_single.py:3944-3949).  Thread <-> matrix row (coalesced along the rows of
// the column-major eigenvector matrix), the Floquet blocks are folded with atomics.
// grid = (ceil(N/256), nmat); d_tp must be zeroed.
__global__ void __launch_bounds__(256) transprob_kernel(int N, int Npad, const double *base, size_t slot,
                                                        size_t oQ, int dim0, const double *__restrict__ d_hl,
                                                        double *__restrict__ d_tp, int f0)
{
    __shared__ double sh[256];
    const int mat = blockIdx.y;
    const double *Z = base + (size_t)mat * slot + oQ;
    const double *hl = d_hl + (size_t)(f0 + mat) * N;
    const int r = blockIdx.x * 256 + threadIdx.x;
    double acc = 0.0;
    for (int i0 = 0; i0 < N; i0 += 256) {
        __syncthreads();
        sh[threadIdx.x] = (i0 + threadIdx.x < N) ? hl[i0 + threadIdx.x] : 0.0;
        __syncthreads();
        if (r < N) {
            const int n = min(256, N - i0);
            for (int i = 0; i < n; ++i) {
                const double z = Z[r + (size_t)(i0 + i) * Npad];
                acc += z * z * sh[i];
            }
        }
    }
    if (r < N) atomicAdd(&d_tp[(size_t)(f0 + mat) * dim0 + (r % dim0)], acc);
}

int solve_batch(const WorkView &W, int nmat, int nsel, bool window, double sigma,
                cudaStream_t st, int *launches)
{
    const EigLayout &L = W.S.L;
    int rc;
    if (L.two_stage) {
        rc = eig_sy2sb(W.S, nmat, st, launches);
        if (rc) return rc;
        rc = eig_sb2st(W.S, nmat, st, launches);
    } else {
        rc = eig_sytrd(W.S, nmat, eig_pick_cluster(L.N, nmat), st, launches);
    }
    if (rc) return rc;
    if (window && nsel < L.N && eig_stein_usable(L.N, nsel))
        rc = eig_stein(W.S, nmat, nsel, sigma, W.sel0, st, launches);      // k << N: bisection + inverse iteration
    else
        rc = eig_stedc(W.S, nmat, st, launches, window ? nsel : 0, sigma, window ? W.sel0 : nullptr);
    if (rc) return rc;
    const int32_t *sel = window ? W.sel0 : nullptr;
    if (L.two_stage) {
        const char *ev = getenv("ARCB200_ORMTR_T");         // tuning knob: 0 = slab kernel for Q1
        const bool fused = !(ev && ev[0] == '0');
        rc = eig_q2_apply(W.S, nmat, sel, nsel, st, launches, fused);
        if (rc || fused) return rc;
    }
    return eig_ormtr(W.S, nmat, sel, nsel, st, launches);
}

}  // namespace

extern "C" int64_t arcb200_sweep_workspace_bytes(int32_t N, int32_t batch, int32_t nvec)
{
    if (N < 1 || batch < 1) return -1;
    return (int64_t)work_bytes(eig_layout(N, nvec), batch);
}

extern "C" int arcb200_scatter_coo(int device, void *stream, int32_t N, int64_t nnz,
                                   const int32_t *d_rows, const int32_t *d_cols,
                                   const double *d_vals, double *d_dense)
{
    if (nnz <= 0) return 0;
    ARCB200_REQUIRE(N >= 1 && d_rows && d_cols && d_vals && d_dense, "scatter_coo: N < 1 or null pointer");
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    scatter_coo_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(N, nnz, d_rows, d_cols, d_vals, d_dense);
    ARCB200_LAUNCH_CHECK("scatter_coo_kernel");
    return 0;
}

static int stark_sweep_impl(int device, void *stream, int32_t N, const double *d_h0diag,
                            const double *d_V, int32_t nF, const double *d_F, int32_t idx0,
                            const double *d_drive_w, int32_t topk, double contrib_max,
                            double *d_y, double *d_hl, double *d_comp_c, int32_t *d_comp_i,
                            void *d_work, int64_t wbytes, int32_t batch, int32_t dim0,
                            double *d_tprob, int32_t *h_launches)
{
    if (h_launches) *h_launches = 0;
    if (nF <= 0) return 0;
    ARCB200_REQUIRE(N >= 1, "stark sweep: N = %d < 1", N);
    ARCB200_REQUIRE(batch >= 1 && batch <= 65535, "stark sweep: batch = %d outside [1, 65535]", batch);
    ARCB200_REQUIRE(d_h0diag && d_V && d_F && d_y && d_hl && d_work, "stark sweep: null operator, field, result or workspace pointer");
    ARCB200_REQUIRE(idx0 >= 0 && idx0 < N, "stark sweep: tracked state index %d outside [0, %d)", idx0, N);
    ARCB200_REQUIRE(topk >= 0 && (topk == 0 || (d_comp_c && d_comp_i)), "stark sweep: topk < 0 or null composition arrays");
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const EigLayout L = eig_layout(N, N);
    if ((int64_t)work_bytes(L, batch) > wbytes) { arcb200_set_error("workspace too small"); return -1; }
    const WorkView W = carve(d_work, L, batch);
    const int ntiles = L.NB * (L.NB + 1) / 2;
    int launches = 0;
    if (d_tprob) ARCB200_CUDA_CHECK(cudaMemsetAsync(d_tprob, 0, (size_t)nF * dim0 * sizeof(double), st));
    for (int f0 = 0; f0 < nF; f0 += batch) {
        const int nmat = min(batch, nF - f0);
        assemble_stark_kernel<<<dim3(ntiles, nmat), 256, 0, st>>>(
            N, L.Npad, W.S.dbl, L.slot_doubles, L.oA, d_h0diag, d_V, d_F, f0);
        ++launches;
        int rc = solve_batch(W, nmat, N, false, 0.0, st, &launches);
        if (rc) return rc;
        epilogue_kernel<<<dim3((N + 7) / 8, nmat), 256, 0, st>>>(
            N, L.Npad, W.S.dbl, L.slot_doubles, L.oQ1, L.oLam, nullptr, N, idx0, d_drive_w,
            (topk > 0 && topk <= MAXTOP) ? topk : 0, contrib_max, d_y, d_hl, d_comp_c, d_comp_i, f0);
        ++launches;
        if (topk > MAXTOP) {
            rc = launch_full_composition(N, L.Npad, W.S.dbl, L.slot_doubles, L.oQ1, nullptr, N, nmat, topk,
                                         contrib_max, d_comp_c, d_comp_i, f0, st);
            if (rc) return rc;
            ++launches;
        }
        if (d_tprob) {
            transprob_kernel<<<dim3((N + 255) / 256, nmat), 256, 0, st>>>(
                N, L.Npad, W.S.dbl, L.slot_doubles, L.oQ1, dim0, d_hl, d_tprob, f0);
            ++launches;
        }
    }
    ARCB200_LAUNCH_CHECK("stark sweep");
    if (h_launches) *h_launches = launches;
    return 0;
}

extern "C" int arcb200_stark_sweep(int device, void *stream, int32_t N, const double *d_h0diag,
                                   const double *d_V, int32_t nF, const double *d_F, int32_t idx0,
                                   const double *d_drive_w, int32_t topk, double contrib_max,
                                   double *d_y, double *d_hl, double *d_comp_c, int32_t *d_comp_i,
                                   void *d_work, int64_t wbytes, int32_t batch, int32_t *h_launches)
{
    return stark_sweep_impl(device, stream, N, d_h0diag, d_V, nF, d_F, idx0, d_drive_w, topk,
                            contrib_max, d_y, d_hl, d_comp_c, d_comp_i, d_work, wbytes, batch, 0,
                            nullptr, h_launches);
}

extern "C" int arcb200_floquet_sweep(int device, void *stream, int32_t N, int32_t dim0,
                                     const double *d_h0diag, const double *d_B, int32_t nF,
                                     const double *d_F, int32_t idx0, double *d_y, double *d_hl,
                                     double *d_tprob, void *d_work, int64_t wbytes, int32_t batch,
                                     int32_t *h_launches)
{
    if (dim0 <= 0 || N % dim0 != 0) { arcb200_set_error("floquet_sweep: N must be a multiple of dim0"); return -1; }
    ARCB200_REQUIRE(nF <= 0 || d_tprob, "floquet_sweep: null transition-probability array");
    return stark_sweep_impl(device, stream, N, d_h0diag, d_B, nF, d_F, idx0, nullptr, 0, 2.0, d_y,
                            d_hl, nullptr, nullptr, d_work, wbytes, batch, dim0, d_tprob, h_launches);
}

extern "C" int arcb200_pair_sweep(int device, void *stream, int32_t N, const double *d_h0diag,
                                  int32_t norders, const double *const *d_M, int32_t nR,
                                  const double *d_R, int32_t k, double sigma, int32_t opi,
                                  const double *d_drive_w, int32_t topk, double contrib_max,
                                  double *d_y, double *d_hl, double *d_comp_c,
                                  int32_t *d_comp_i, double *d_overlap, void *d_work,
                                  int64_t wbytes, int32_t batch, int32_t *h_launches)
{
    if (h_launches) *h_launches = 0;
    if (nR <= 0) return 0;
    ARCB200_REQUIRE(N >= 1, "pair sweep: N = %d < 1", N);
    if (k < 1 || k > N) { arcb200_set_error("k must be in [1, N]"); return -1; }
    ARCB200_REQUIRE(batch >= 1 && batch <= 65535, "pair sweep: batch = %d outside [1, 65535]", batch);
    ARCB200_REQUIRE(norders >= 0 && (norders == 0 || d_M), "pair sweep: norders < 0 or null operator table");
    ARCB200_REQUIRE(d_h0diag && d_R && d_y && d_hl && d_work, "pair sweep: null operator, distance, result or workspace pointer");
    ARCB200_REQUIRE(opi >= 0 && opi < N, "pair sweep: tracked state index %d outside [0, %d)", opi, N);
    ARCB200_REQUIRE(topk >= 0 && (topk == 0 || (d_comp_c && d_comp_i)), "pair sweep: topk < 0 or null composition arrays");
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const EigLayout L = eig_layout(N, k);
    if ((int64_t)work_bytes(L, batch) > wbytes) { arcb200_set_error("workspace too small"); return -1; }
    const WorkView W = carve(d_work, L, batch);
    const int ntiles = L.NB * (L.NB + 1) / 2;
    int launches = 0;
    // eigenvector window of the last point of the previous batch (for the overlaps): right after
    // the carved workspace, 256-byte aligned like the carve itself (work_bytes() is a multiple of
    // 256 and already contains the slack of the alignment of d_work)
    double *prevZ = (double *)((((uintptr_t)d_work + 255) & ~(uintptr_t)255) + work_bytes(L, batch) - 256);
    if (d_overlap && (int64_t)(work_bytes(L, batch) + (size_t)N * k * sizeof(double)) > wbytes) {
        arcb200_set_error("workspace too small for overlaps (add N*k doubles)");
        return -1;
    }
    // One chunk of sweep points: assemble -> eigensolve -> epilogue in the slots [slot0, slot0 + nmat) on stream cs
    auto run_chunk = [&](int p0, int nmat, int slot0, cudaStream_t cs) -> int {
        WorkView V = W;
        V.S.dbl = W.S.dbl + (size_t)slot0 * L.slot_doubles;
        V.S.ints = W.S.ints + (size_t)slot0 * L.slot_ints;
        V.S.gemm_probs = (char *)W.S.gemm_probs + (size_t)slot0 * EIG_MAX_NODES * EIG_PROB_BYTES;
        V.sel0 = W.sel0 + slot0;
        assemble_pair_kernel<<<dim3(ntiles, nmat), 256, 0, cs>>>(
            N, L.Npad, V.S.dbl, L.slot_doubles, L.oA, d_h0diag, norders, d_M, d_R, p0);
        ++launches;
        int rc = solve_batch(V, nmat, k, true, sigma, cs, &launches);
        if (rc) return rc;
        epilogue_kernel<<<dim3((k + 7) / 8, nmat), 256, 0, cs>>>(
            N, L.Npad, V.S.dbl, L.slot_doubles, L.oQ1, L.oLam, V.sel0, k, opi, d_drive_w,
            topk <= MAXTOP ? topk : 0, contrib_max, d_y, d_hl, d_comp_c, d_comp_i, p0);
        ++launches;
        if (topk > MAXTOP) {
            rc = launch_full_composition(N, L.Npad, V.S.dbl, L.slot_doubles, L.oQ1, V.sel0, k, nmat, topk,
                                         contrib_max, d_comp_c, d_comp_i, p0, cs);
            if (rc) return rc;
            ++launches;
        }
        return 0;
    };
    // Experiment kept behind ARCB200_PIPELINE=1 (off by default): batches smaller than the SM count (C5 holds 74
    // matrices in HBM; strong-scaling shards) leave SMs idle in every kernel that gives a matrix one CTA (bulge
    // chasing, Q2, Q1: 29 % of the C5 step on half of the SMs).  Splitting such a batch into two half-batches on
    // two streams lets the one-CTA-per-matrix stages of a chunk run beside the band reduction of the next -- but
    // those stages take as long for 37 matrices as for 74 (latency bound per matrix), so the sweep got slower:
    // C5 74 points 5.94 -> 7.13 s, C3 125 points 281 -> 283 ms, 74 points 186 -> 217 ms (tools/ab_pipeline.py).
    const char *epl = getenv("ARCB200_PIPELINE");
    const bool pipe = epl && epl[0] == '1' && L.two_stage && !d_overlap && batch >= 2;
    if (pipe) {
        ArcHelperGuard guard;          // held while the chunks are enqueued on the device's pipeline streams
        ArcHelperStreams *hp = guard.pool(device);
        if (!hp) return -1;
        cudaStream_t *pstream = hp->pipe_stream;
        cudaEvent_t *pev = hp->pipe_ev;
        ARCB200_CUDA_CHECK(cudaEventRecord(pev[2], st));
        for (int i = 0; i < 2; ++i) ARCB200_CUDA_CHECK(cudaStreamWaitEvent(pstream[i], pev[2], 0));
        const int half = (batch + 1) / 2;
        int c = 0;
        for (int p0 = 0; p0 < nR; ++c) {
            const int h = c & 1;
            const int nmat = min(h ? batch - half : half, nR - p0);
            const int rc = run_chunk(p0, nmat, h ? half : 0, pstream[h]);
            if (rc) return rc;
            p0 += nmat;
        }
        for (int i = 0; i < 2; ++i) {
            ARCB200_CUDA_CHECK(cudaEventRecord(pev[i], pstream[i]));
            ARCB200_CUDA_CHECK(cudaStreamWaitEvent(st, pev[i], 0));
        }
    } else {
        for (int p0 = 0; p0 < nR; p0 += batch) {
            const int nmat = min(batch, nR - p0);
            const int rc = run_chunk(p0, nmat, 0, st);
            if (rc) return rc;
            if (d_overlap) {
                overlap_kernel<<<dim3((k + 15) / 16, (k + 15) / 16, nmat), 256, 0, st>>>(
                    N, L.Npad, W.S.dbl, L.slot_doubles, L.oQ1, W.sel0, k, p0 == 0 ? nullptr : prevZ,
                    d_overlap, p0);
                save_window_kernel<<<148, 256, 0, st>>>(N, L.Npad, W.S.dbl, L.slot_doubles, L.oQ1, W.sel0,
                                                        k, nmat - 1, prevZ);
                launches += 2;
            }
        }
    }
    ARCB200_LAUNCH_CHECK("pair sweep");
    if (h_launches) *h_launches = launches;
    return 0;
}

extern "C" int arcb200_syevd_batched(int device, void *stream, int32_t N, int32_t batch,
                                     double *d_A, double *d_w, double *d_Z, void *d_work,
                                     int64_t wbytes, int32_t *h_launches)
{
    if (h_launches) *h_launches = 0;
    if (batch <= 0) return 0;
    ARCB200_REQUIRE(N >= 1 && batch <= 65535, "syevd_batched: N < 1 or batch > 65535");
    ARCB200_REQUIRE(d_A && d_w && d_Z && d_work, "syevd_batched: null pointer");
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const EigLayout L = eig_layout(N, N);
    if ((int64_t)work_bytes(L, batch) > wbytes) { arcb200_set_error("workspace too small"); return -1; }
    const WorkView W = carve(d_work, L, batch);
    int launches = 0;
    load_dense_kernel<<<dim3(148, batch), 256, 0, st>>>(N, L.Npad, W.S.dbl, L.slot_doubles, L.oA, d_A);
    int rc = solve_batch(W, batch, N, false, 0.0, st, &launches);
    if (rc) return rc;
    store_dense_kernel<<<dim3(148, batch), 256, 0, st>>>(N, L.Npad, W.S.dbl, L.slot_doubles, L.oQ1,
                                                         L.oLam, d_Z, d_w);
    ARCB200_LAUNCH_CHECK("syevd");
    if (h_launches) *h_launches = launches + 2;
    return 0;
}

// ---------------------------------------------------------------- stage-level entry points
__global__ void copy_vec_kernel(int N, const double *base, size_t slot, size_t off, double *out)
{
    const int mat = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        out[(size_t)mat * N + i] = base[(size_t)mat * slot + off + i];
}
__global__ void put_vec_kernel(int N, double *base, size_t slot, size_t off, const double *in)
{
    const int mat = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        base[(size_t)mat * slot + off + i] = in[(size_t)mat * N + i];
}

extern "C" int arcb200_sytrd_batched(int device, void *stream, int32_t N, int32_t batch,
                                     const double *d_A, double *d_d, double *d_e, double *d_tau,
                                     void *d_work, int64_t wbytes, int32_t cluster, int32_t *h_launches)
{
    if (h_launches) *h_launches = 0;
    if (batch <= 0) return 0;
    ARCB200_REQUIRE(N >= 1 && batch <= 65535, "sytrd_batched: N < 1 or batch > 65535");
    ARCB200_REQUIRE(d_A && d_d && d_e && d_work, "sytrd_batched: null pointer");
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const EigLayout L = eig_layout(N, N);
    if ((int64_t)work_bytes(L, batch) > wbytes) { arcb200_set_error("workspace too small"); return -1; }
    const WorkView W = carve(d_work, L, batch);
    int launches = 0;
    load_dense_kernel<<<dim3(148, batch), 256, 0, st>>>(N, L.Npad, W.S.dbl, L.slot_doubles, L.oA, d_A);
    int rc;
    if (L.two_stage && cluster <= 0) {
        rc = eig_sy2sb(W.S, batch, st, &launches);
        if (rc) return rc;
        rc = eig_sb2st(W.S, batch, st, &launches);
    } else {
        rc = eig_sytrd(W.S, batch, cluster > 0 ? cluster : eig_pick_cluster(N, batch), st, &launches);
    }
    if (rc) return rc;
    copy_vec_kernel<<<dim3(4, batch), 256, 0, st>>>(N, W.S.dbl, L.slot_doubles, L.oD, d_d);
    copy_vec_kernel<<<dim3(4, batch), 256, 0, st>>>(N, W.S.dbl, L.slot_doubles, L.oE, d_e);
    if (d_tau) copy_vec_kernel<<<dim3(4, batch), 256, 0, st>>>(N, W.S.dbl, L.slot_doubles, L.oTau, d_tau);
    ARCB200_LAUNCH_CHECK("sytrd_batched");
    if (h_launches) *h_launches = launches + 3;
    return 0;
}

// band after stage 1 of the two-stage reduction: d_band[b][j][d] = B[j+d][j], d = 0..32
__global__ void copy_band_kernel(int N, const double *base, size_t slot, size_t off, double *out)
{
    const int mat = blockIdx.y;
    const double *Bd = base + (size_t)mat * slot + off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * 33; i += gridDim.x * blockDim.x) {
        const int j = i / 33, d = i % 33;
        out[(size_t)mat * N * 33 + i] = Bd[d + (size_t)j * EIG_LDB];
    }
}

extern "C" int arcb200_sy2sb_batched(int device, void *stream, int32_t N, int32_t batch,
                                     const double *d_A, double *d_band, void *d_work,
                                     int64_t wbytes, int32_t *h_launches)
{
    if (h_launches) *h_launches = 0;
    if (batch <= 0) return 0;
    ARCB200_REQUIRE(N >= 1 && batch <= 65535, "sy2sb_batched: N < 1 or batch > 65535");
    ARCB200_REQUIRE(d_A && d_band && d_work, "sy2sb_batched: null pointer");
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const EigLayout L = eig_layout(N, N);
    if (!L.two_stage) { arcb200_set_error("sy2sb_batched: N=%d is outside the two-stage range", N); return -1; }
    if ((int64_t)work_bytes(L, batch) > wbytes) { arcb200_set_error("workspace too small"); return -1; }
    const WorkView W = carve(d_work, L, batch);
    int launches = 0;
    load_dense_kernel<<<dim3(148, batch), 256, 0, st>>>(N, L.Npad, W.S.dbl, L.slot_doubles, L.oA, d_A);
    int rc = eig_sy2sb(W.S, batch, st, &launches);
    if (rc) return rc;
    copy_band_kernel<<<dim3(16, batch), 256, 0, st>>>(N, W.S.dbl, L.slot_doubles, L.oBand, d_band);
    ARCB200_LAUNCH_CHECK("sy2sb_batched");
    if (h_launches) *h_launches = launches + 2;
    return 0;
}

extern "C" int arcb200_stedc_batched(int device, void *stream, int32_t N, int32_t batch,
                                     const double *d_d, const double *d_e, double *d_w,
                                     double *d_Q, void *d_work, int64_t wbytes, int32_t *h_launches)
{
    if (h_launches) *h_launches = 0;
    if (batch <= 0) return 0;
    ARCB200_REQUIRE(N >= 1 && batch <= 65535, "stedc_batched: N < 1 or batch > 65535");
    ARCB200_REQUIRE(d_d && d_e && d_w && d_Q && d_work, "stedc_batched: null pointer");
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const EigLayout L = eig_layout(N, N);
    if ((int64_t)work_bytes(L, batch) > wbytes) { arcb200_set_error("workspace too small"); return -1; }
    const WorkView W = carve(d_work, L, batch);
    int launches = 0;
    put_vec_kernel<<<dim3(4, batch), 256, 0, st>>>(N, W.S.dbl, L.slot_doubles, L.oD, d_d);
    put_vec_kernel<<<dim3(4, batch), 256, 0, st>>>(N, W.S.dbl, L.slot_doubles, L.oE, d_e);
    int rc = eig_stedc(W.S, batch, st, &launches);
    if (rc) return rc;
    store_dense_kernel<<<dim3(148, batch), 256, 0, st>>>(N, L.Npad, W.S.dbl, L.slot_doubles, L.oQ1,
                                                         L.oLam, d_Q, d_w);
    ARCB200_LAUNCH_CHECK("stedc_batched");
    if (h_launches) *h_launches = launches + 3;
    return 0;
}
