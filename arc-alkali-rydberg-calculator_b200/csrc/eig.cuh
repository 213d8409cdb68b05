// eig.cuh -- workspace layout and internal interfaces of the batched symmetric
// eigensolver (subsystem 3).  All matrices are stored column-major with leading
// dimension Npad = N rounded up to 32; only the lower triangle of A is referenced.
#pragma once
#include "common.cuh"

constexpr int EIG_PB = 32;          // panel width = tile width
constexpr int EIG_MAX_NODES = 512;  // merges per level per matrix (N <= 32768)
constexpr int EIG_PROB_BYTES = 96;  // >= sizeof(GemmProblem)

struct EigLayout {
    int N, Npad, NB, nvec;
    // per-slot element offsets (in doubles) inside the slot
    size_t oA, oQ1, oQ2, oU, oWp, oVp, oTrans, oWdir, oD, oE, oTau, oScr, oTfac, oLam, oAux;
    // two-stage path (eig_twostage.cu): compact band, bulge-chasing reflectors and their taus
    size_t oBand, oV2, oTau2;
    int two_stage;
    size_t slot_doubles;
    // int workspace per slot (in int32)
    size_t oPerm, oIdx, slot_ints;
};

constexpr int EIG_LDB = 72;         // leading dimension of the compact lower band (>= 2*32 + 1)

static inline size_t eig_align(size_t x) { return (x + 31) & ~(size_t)31; }

// Two-stage tridiagonalisation (dense -> band -> tridiagonal) for N >= 1024; smaller matrices
// are latency bound either way and use the one-stage kernel.  ARCB200_TWO_STAGE=0/1 overrides.
int eig_use_two_stage(int N);
// Pair sweeps with k << N take the bisection + inverse iteration path (eig_stein.cu): no N x N
// eigenvector matrices are needed then, and the slot keeps only N x k windows (compact layout).
int eig_stein_usable(int N, int k);

static inline EigLayout eig_layout(int N, int nvec)
{
    EigLayout L;
    L.N = N;
    L.Npad = (N + 31) & ~31;
    L.NB = L.Npad / 32;
    L.nvec = nvec;
    const size_t np = (size_t)L.Npad;
    size_t o = 0;
    // compact: only nvec << N eigenvectors are ever formed (stein path).  Q1 = the N x nvec window
    // (>= 128 columns: the look-ahead band-reduction variant parks three panels there), Q2 = the four
    // LU planes of the inverse iteration, U = the iterate + pivot flags, later the row-major window of
    // the back-transformation.  C5 (N = 10384): 1.8 GB per matrix instead of 4.3 GB, so 2.4 x as many
    // matrices share the kernels that give a matrix one CTA (bulge chasing, panel QR, Q2).
    const bool compact = nvec > 0 && nvec < N && eig_stein_usable(N, nvec);
    const size_t kp = (size_t)((nvec + 31) & ~31), kq = kp < 128 ? 128 : kp;
    L.oA = o;      o += np * np;
    L.oQ1 = o;     o += compact ? np * kq : np * np;
    L.oQ2 = o;     o += compact ? 4 * np * kp : np * np;
    L.oU = o;      o += compact ? (np + 64) * (kq + 64) + np * kp / 8 + 64
                                : np * (np + 64);   // also the row-major Z window of the Q2 pass
    L.oWp = o;     o += np * EIG_PB;
    L.oVp = o;     o += np * EIG_PB;
    L.oTrans = o;  o += (size_t)L.NB * np;
    L.oWdir = o;   o += np;
    L.oD = o;      o += np;
    L.oE = o;      o += np;
    L.oTau = o;    o += np;
    L.oScr = o;    o += 4096;                       // cluster exchange scratch
    L.oTfac = o;   o += (size_t)L.NB * EIG_PB * EIG_PB;   // T factors of the block reflectors
    L.oLam = o;    o += np;                         // eigenvalues
    L.oAux = o;    o += 12 * np;                    // D&C vectors (z, dlamda, w, ...)
    L.two_stage = eig_use_two_stage(N);
    L.oBand = L.oV2 = L.oTau2 = o;
    if (L.two_stage) {
        L.oBand = o;   o += (size_t)EIG_LDB * (np + 64);
        L.oV2 = o;     o += np * np;
        L.oTau2 = o;   o += np * (size_t)(L.NB + 2);
    }
    L.slot_doubles = eig_align(o);
    size_t oi = 0;
    L.oPerm = oi;  oi += 8 * np;
    L.oIdx = oi;   oi += 4 * np;
    L.slot_ints = eig_align(oi);
    return L;
}

struct EigSlots {
    EigLayout L;
    double *dbl;       // batch * slot_doubles
    int32_t *ints;     // batch * slot_ints
    int batch;
    void *gemm_probs;  // batch * EIG_MAX_NODES GemmProblem descriptors
    __host__ __device__ double *slot(int b) const { return dbl + (size_t)b * L.slot_doubles; }
    __host__ __device__ int32_t *islot(int b) const { return ints + (size_t)b * L.slot_ints; }
};

// stage 1: A (lower, column-major, lda=Npad) -> d, e, Householder vectors in A, tau
int eig_sytrd(const EigSlots &S, int nmat, int cluster, cudaStream_t st, int *launches);
// stage 2: (d, e) -> eigenvalues (ascending, in oLam) and eigenvectors of T in Q1
// sel_k > 0: only the sel_k eigenpairs nearest sigma are needed; their first index per
// matrix is written to d_sel0 and only those eigenvector columns of Q1 are valid.
int eig_stedc(const EigSlots &S, int nmat, cudaStream_t st, int *launches, int sel_k = 0,
              double sigma = 0.0, int32_t *d_sel0 = nullptr);
// stage 2 for pair sweeps with k << N (eig_stein.cu): the k eigenpairs nearest sigma by bisection +
// inverse iteration; eigenvalues land in oLam[0..k), eigenvectors of T in Q1 columns 0..k-1,
// d_sel0[mat] = 0.
int eig_stein(const EigSlots &S, int nmat, int k, double sigma, int32_t *d_sel0, cudaStream_t st, int *launches);
// stage 3: Z = Q_house * Q1[:, sel0 : sel0+nsel]  (in place on Q1 columns)
int eig_ormtr(const EigSlots &S, int nmat, const int32_t *d_sel0, int nsel, cudaStream_t st,
              int *launches);
int eig_pick_cluster(int N, int nmat);
// two-stage variants: A -> band (V1 below the band, T factors in oTfac) -> (d, e) with the
// bulge-chasing reflectors in oV2/oTau2; back-transformation Z = Q1 (Q2 Z)
int eig_sy2sb(const EigSlots &S, int nmat, cudaStream_t st, int *launches);
int eig_sb2st(const EigSlots &S, int nmat, cudaStream_t st, int *launches);
// (with_q1: also apply the band-reduction reflectors Q1 on the row-major window, i.e. the whole
// back-transformation Z = Q1 (Q2 Z); otherwise eig_ormtr has to follow)
int eig_q2_apply(const EigSlots &S, int nmat, const int32_t *d_sel0, int nsel, cudaStream_t st,
                 int *launches, bool with_q1);
