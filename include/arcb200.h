/*
 * include/arcb200.h -- C-ABI of libarcb200.so, the B200 (sm_100a) hot path of ARC.
 *
 * Plain pointers and sizes only; no torch / CPython types.  Every entry point
 * names the reference interface (file:line under /root/reference) it replaces.
 * Conventions:
 *   - every `d_*` pointer is DEVICE memory owned by the caller (the Python host
 *     passes torch CUDA tensors' data_ptr()); the library never allocates
 *     result buffers and never frees caller memory;
 *   - `device` is the CUDA ordinal, `stream` a cudaStream_t passed as void*
 *     (NULL = legacy default stream);
 *   - return value 0 = success, negative = error (-1 invalid argument or failed CUDA API call,
 *     -2 kernel launch failure, -3 not an sm_100 device); arcb200_last_error() returns a
 *     thread-local message.  Sizes, index ranges and required pointers are validated on the
 *     host before the first CUDA call (tests/test_abi.py), so a bad call never hangs or faults.  Entry points that launch a variable number of kernels report
 *     the count through a trailing `int32_t *h_launches` (host pointer, may be NULL);
 *     per-item failures come back in device status arrays (d_status), never as aborts.  Nothing calls exit()/abort() (the reference's C
 *     extension returns NULL without setting an exception, arc_c_extensions.c:145,200);
 *   - re-entrant per (device, stream): no integrator or solver state outlives a call (the
 *     reference keeps all integrator state in globals, arc_c_extensions.c:11-24,57-67).  The only
 *     process-wide objects are one pool of helper streams / events per device (sub-batches of the
 *     band reduction), created on first use; a host thread holds the pool's mutex while it enqueues,
 *     so concurrent callers on one device are safe (their sub-batches queue behind each other).
 */
#ifndef ARCB200_H
#define ARCB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARCB200_ABI_VERSION 2

/* ------------------------------------------------------------------------ */
/* library                                                                   */
/* ------------------------------------------------------------------------ */
int arcb200_abi_version(void);
const char *arcb200_last_error(void);
/* fills props[0..7] = {sm_count, cc_major, cc_minor, l2_bytes, smem_per_block_optin,
 *                      total_mem_MiB, cluster_launch_supported, clock_khz} */
int arcb200_device_info(int device, int64_t *props);

/* FP64 tensor-core (mma.sync.m8n8k4.f64) peak of the device in TFLOP/s, register operands
 * only: the denominator of the roofline of the DMMA-bound eigensolver kernels (the driver's
 * MEASURED_PEAKS.json has no FP64 figure).  d_scratch: one device double. */
int arcb200_dmma_peak(int device, void *stream, double *d_scratch, double *h_tflops);
/* Same for the plain FP64 FMA pipe (DFMA), 16 independent chains per thread. */
int arcb200_dfma_peak(int device, void *stream, double *d_scratch, double *h_tflops);

/* Tile-stream probe (diagnostic): every warp of one 8-warp CTA per SM streams 32 x 32 FP64 tiles by
 * cp.async, `depth` (1|2) tiles ahead, and spends 128 DMMAs on each -- the memory side of the band
 * reduction's symmetric product in isolation.  mode 0: tiles of a column-major matrix (32 segments of
 * 256 B, lda doubles apart), mode 1: tile-major (8 KB contiguous).  d_buf: >= sm_count * lda^2 doubles.
 * h_out = {ms per launch, fraction of dmma_tflops, GB/s}. */
int arcb200_tile_stream_probe(int device, void *stream, const double *d_buf, int64_t buf_doubles,
                              int32_t lda, int32_t ntiles_per_warp, int32_t mode, int32_t depth,
                              double dmma_tflops, double *h_out);

/* ------------------------------------------------------------------------ */
/* subsystem (1): Numerov radial wavefunctions + radial integrals            */
/* ------------------------------------------------------------------------ */

/* One radial state = the 18 scalars of the reference's C-extension call
 * NumerovWavefunction(innerLimit, outerLimit, step, init1, init2, l, s, j,
 *   stateEnergy, alphaC, alpha, Z, a1, a2, a3, a4, rc, mu)
 * (arc_c_extensions.c:113-116, format "dddddidddddidddddd"). 144 bytes. */
typedef struct arcb200_state {
    double innerLimit, outerLimit, step, init1, init2;
    double s, j, stateEnergy, alphaC, alpha;
    double a1, a2, a3, a4, rc, mu;
    int32_t l, Z;
    int32_t length;   /* filled by arcb200_numerov_lengths: grid points L */
    int32_t _pad;
} arcb200_state_t;

/* L = (int)((sqrt(outer)-sqrt(inner))/step)  (arc_c_extensions.c:131).
 * Host-side, fills states[i].length and offsets[0..n] (prefix sum, each segment
 * padded to a multiple of 32 doubles); returns total doubles needed, <0 on error. */
int64_t arcb200_numerov_lengths(int32_t nstates, arcb200_state_t *h_states,
                                int64_t *h_offsets);

/* Batched replacement of NumerovWavefunction (arc_c_extensions.c:108-285) fused with
 * the normalisation of AlkaliAtom.radialWavefunction (alkali_atom_functions.py:603-606).
 * One warp per state.  Outputs, per state i at [d_offsets[i], d_offsets[i]+length):
 *   d_psi  = R(r)*r   (normalised so trapezoid(psi^2, r) = 1 if normalise!=0,
 *                      else raw like the C extension's row 0), zero below the
 *                      divergence point
 *   d_r    = r grid   (the C extension's row 1)
 *   d_norm[i] = trapezoid(psi_raw^2, r);  d_dp[i] = divergencePoint
 *   d_status[i] (may be NULL) = 0 ok | 1 the outward walk left the integrated range (the
 *                reference's "Numerov error", arc_c_extensions.c:196-201: returns NULL)
 *                | 2 norm not finite or not positive (no usable wavefunction)
 * d_order = processing order (state indices, longest first) or NULL. */
int arcb200_numerov_batch(int device, void *stream, int32_t nstates,
                          const arcb200_state_t *d_states, const int64_t *d_offsets,
                          const int32_t *d_order, double *d_psi, double *d_r,
                          double *d_norm, int32_t *d_dp, int32_t normalise,
                          int32_t *d_status);

/* Coefficient-array variant for NumerovBack (alkali_atom_functions.py:3939-4063, the pure-Python
 * twin that takes an arbitrary callable kfun): the host evaluates kfun on the grid and passes, per
 * state and grid index i (packed at d_offsets[s] + i, i < d_lengths[s]),
 *   A_i = 2(1 - 5/12 h^2 kfun(x_i)),  B_i = 1 + h^2/12 kfun(x_i + h),  D_i = 1 + h^2/12 kfun(x_i - h);
 * the kernel runs sol[i] = (A_i sol[i+1] - B_i sol[i+2]) / D_i from i = L-1 down (x_{L-1} =
 * d_xtop[s], x_{i-1} = x_i - h) with the twin's divergence control and tail rule.
 * d_inits = double[nstates][2] = {init1, init2}.  Outputs unnormalised: d_psi = sol*sqrt(x),
 * d_r = x^2; d_dp = divergencePoint; d_status as in arcb200_numerov_batch. */
int arcb200_numerov_coeff_batch(int device, void *stream, int32_t nstates,
                                const int64_t *d_offsets, const int32_t *d_lengths,
                                const double *d_A, const double *d_B, const double *d_D,
                                const double *d_xtop, const double *d_step, const double *d_inits,
                                double *d_psi, double *d_r, int32_t *d_dp, int32_t *d_status);

/* Batched radial overlap integrals, one warp per pair:
 *   out[p] = trapezoid(psi_a*psi_b*r_a^power, x=r_a)[0:min(La,Lb)]
 * replaces the integration in AlkaliAtom.getRadialMatrixElement
 * (alkali_atom_functions.py:1087-1096, power=1) and getQuadrupoleMatrixElement
 * (alkali_atom_functions.py:1191-1201, power=2).  d_pairs = int32[npairs][3] =
 * {state a (lower energy; its r grid is used), state b, power}. */
int arcb200_radial_integrals(int device, void *stream, int64_t npairs,
                             const int32_t *d_pairs, const arcb200_state_t *d_states,
                             const int64_t *d_offsets, const double *d_psi,
                             const double *d_r, double *d_out);

/* The same integrals for whole BLOCKS of state pairs as tensor-core Gram matrices:
 * block g computes out[out_off + a*ncols + b] for a in rows[row0..row0+nrows),
 * b in cols[col0..col0+ncols) (state indices), power 1|2.  All states must share the x grid
 * origin and step (same atom, same innerLimit); d_rgrid[0..Lgrid) is the r grid of the LONGEST
 * state of the batch, d_wgt a scratch of 2*Lgrid doubles, d_out must be zero-initialised.
 * arcb200_gram_block_t = {int32 row0, nrows, col0, ncols, power, kmax; int64 out_off} (32 bytes),
 * kmax = min(max La, max Lb).  Replaces the per-element loop over
 * getRadialMatrixElement / getQuadrupoleMatrixElement in StarkMap.defineBasis
 * (calculations_atom_single.py:816-833) and __makeRawMatrix2
 * (calculations_atom_pairstate.py:939-1013). */
typedef struct arcb200_gram_block {
    int32_t row0, nrows, col0, ncols, power, kmax;
    int64_t out_off;
} arcb200_gram_block_t;
int arcb200_radial_gram(int device, void *stream, int32_t nblocks, const void *d_blocks,
                        const int32_t *d_rows, const int32_t *d_cols,
                        const arcb200_state_t *d_states, const int64_t *d_offsets,
                        const double *d_psi, const double *d_rgrid, int32_t Lgrid,
                        double *d_wgt, double *d_out, int32_t max_rows, int32_t max_cols,
                        int32_t max_k);

/* Semiclassical radial elements used by all divalent atoms
 * (alkali_atom_functions.py:2683-2735 dipole, 2737-2802 quadrupole; the Anger
 * functions J_{dnu-+1}(-dnu) replace mpmath.angerj).  One warp per element.
 * d_in = double[n][4] = {nu (state 1), nu1 (state 2), l1, l2}; kind 1|2;
 * d_quad = double[2][nq]: Gauss-Legendre nodes then weights on [-1,1], nq % 32 == 0. */
int arcb200_semiclassical_table(int device, void *stream, int64_t n,
                                const double *d_in, int32_t kind,
                                const double *d_quad, int32_t nq, double *d_out);

/* ------------------------------------------------------------------------ */
/* subsystems (2)+(3): Hamiltonian assembly + batched symmetric eigensolve   */
/* ------------------------------------------------------------------------ */

/* Workspace size (bytes) for a sweep of `batch` matrices of order N solved
 * concurrently with nvec eigenvectors kept (nvec = N for Stark maps). */
int64_t arcb200_sweep_workspace_bytes(int32_t N, int32_t batch, int32_t nvec);

/* Stark-map sweep: for every field F[f]:  H = diag(h0) + F[f]*V   (dense, FP64),
 * full symmetric eigendecomposition, then
 *   y[f][i]  = eigenvalue i (ascending)
 *   hl[f][i] = |v[idx0,i]|^2                 if d_drive_w == NULL
 *            = sum_j drive_w[j] |v[j,i]|^2   otherwise
 *   comp_c/comp_i[f][i][0..topk) = coefficient/index of the largest |v[:,i]|
 *            entries, descending, stopping once sum c^2 >= contrib_max
 *            (unused slots: index -1); any topk <= N: up to 8 entries come from the fused
 *            epilogue, more (upTo = -1 -> topk = N, contrib_max = 2) from a full sort of
 *            the eigenvector by |c| (calculations_atom_single.py:1399-1407)
 * Replaces the loop body of StarkMap.diagonalise
 * (calculations_atom_single.py:976-1021, _stateComposition2 1395-1416,
 *  numpy.linalg.eigh at 986). */
int arcb200_stark_sweep(int device, void *stream, int32_t N, const double *d_h0diag,
                        const double *d_V, int32_t nF, const double *d_F, int32_t idx0,
                        const double *d_drive_w, int32_t topk, double contrib_max,
                        double *d_y, double *d_hl, double *d_comp_c, int32_t *d_comp_i,
                        void *d_work, int64_t work_bytes, int32_t batch, int32_t *h_launches);

/* Floquet (Shirley) sweep: for every field amplitude F[f]:  H = diag(h0) + F[f]*B, N = dim0 *
 * (2 fn + 1) (the caller folds the drive frequency into h0 = tile(bareEnergies) + k*freq),
 * full eigendecomposition, then
 *   y[f][i]  = eigenvalue i (ascending),  hl[f][i] = |v[idx0,i]|^2,
 *   tprob[f][a] = sum_k sum_i |v[k*dim0 + a, i]|^2 |v[idx0, i]|^2   (a < dim0)
 * Replaces the loop body of ShirleyMethod.diagonalise
 * (calculations_atom_single.py:3931-3957; Shirley, Phys. Rev. 138, B979, eq. 19). */
int arcb200_floquet_sweep(int device, void *stream, int32_t N, int32_t dim0,
                          const double *d_h0diag, const double *d_B, int32_t nF,
                          const double *d_F, int32_t idx0, double *d_y, double *d_hl,
                          double *d_tprob, void *d_work, int64_t work_bytes, int32_t batch,
                          int32_t *h_launches);

/* Pair-energy scan of StarkMapResonances.findResonances
 * (calculations_atom_pairstate.py:4753-4789): keeps the pairs (j, jj) of single-atom Stark
 * eigenstates of field i with eMin < y1[i][j] + y2[i][jj] - e0 < eMax and ok1[i][j] && ok2[i][jj],
 * in the order of the reference's nested loops.  Two calls: d_rowoff == NULL counts the hits of
 * every row (i, j) into d_rowcount[nF*N1]; with d_rowoff = exclusive prefix sum of those counts
 * the hits are written to d_out_e / d_out_j / d_out_jj. */
int arcb200_resonance_scan(int device, void *stream, int32_t nF, int32_t N1, int32_t N2,
                           const double *d_y1, const double *d_y2, const uint8_t *d_ok1,
                           const uint8_t *d_ok2, double e0, double eMin, double eMax,
                           const int64_t *d_rowoff, int32_t *d_rowcount, double *d_out_e,
                           int32_t *d_out_j, int32_t *d_out_jj);

/* Pair-state sweep: for every distance R[p] (micrometres):
 *   H = diag(h0) + sum_o M_o / (R*1e-6)^(3+o)      (M_o dense N x N, o < norders)
 * d_M = DEVICE array of norders device pointers to the dense row-major M_o.
 * k eigenpairs nearest sigma (GHz), ascending; hl/comp as above (topk entries).
 * d_overlap (optional, nR x k x k): squared overlaps (v_t,i . v_(t-1),j)^2 between the selected
 * eigenvectors of consecutive points, the input of the reference's adiabatic re-ordering
 * (sortEigenvectors, calculations_atom_pairstate.py:3452-3492); point 0 overlaps with itself.
 * When given, the workspace must hold N*k extra doubles.
 * Replaces the loop body of PairStateInteractions.diagonalise
 * (calculations_atom_pairstate.py:3425-3520; scipy eigsh at 3444-3450). */
int arcb200_pair_sweep(int device, void *stream, int32_t N, const double *d_h0diag,
                       int32_t norders, const double *const *d_M, int32_t nR,
                       const double *d_R, int32_t k, double sigma, int32_t opi,
                       const double *d_drive_w, int32_t topk, double contrib_max,
                       double *d_y, double *d_hl, double *d_comp_c, int32_t *d_comp_i,
                       double *d_overlap, void *d_work, int64_t work_bytes, int32_t batch,
                       int32_t *h_launches);

/* Sparse (COO) -> dense scatter used to stage matR[o] on the device
 * (the reference keeps matR as scipy CSR, calculations_atom_pairstate.py:3230-3239). */
int arcb200_scatter_coo(int device, void *stream, int32_t N, int64_t nnz,
                        const int32_t *d_rows, const int32_t *d_cols,
                        const double *d_vals, double *d_dense);

/* Batched dense symmetric eigensolver on caller-assembled matrices
 * (d_A: batch x N x N, overwritten).  Exposed for tests and for callers with
 * other Hamiltonians (numpy.linalg.eigh replacement, calculations_atom_single.py:986).
 * d_w: batch x N ascending; d_Z: batch x N x N, eigenvector i of matrix b is
 * d_Z[b][i][0..N) (row i = vector i). */
int arcb200_syevd_batched(int device, void *stream, int32_t N, int32_t batch,
                          double *d_A, double *d_w, double *d_Z, void *d_work,
                          int64_t work_bytes, int32_t *h_launches);

/* Stage-level entry points of the eigensolver (LAPACK dsytrd / dstedc equivalents),
 * exposed for the stage parity tests.  d_A: batch x N x N symmetric; outputs d, e, tau
 * (batch x N each; e[N-1], tau[N-1] unused).  cluster = CTAs per matrix (0 = auto). */
int arcb200_sytrd_batched(int device, void *stream, int32_t N, int32_t batch,
                          const double *d_A, double *d_d, double *d_e, double *d_tau,
                          void *d_work, int64_t work_bytes, int32_t cluster, int32_t *h_launches);
/* (d, e) -> eigenvalues d_w (ascending) and eigenvectors of the tridiagonal matrix,
 * d_Q[b][i][0..N) = eigenvector i. */
int arcb200_stedc_batched(int device, void *stream, int32_t N, int32_t batch,
                          const double *d_d, const double *d_e, double *d_w, double *d_Q,
                          void *d_work, int64_t work_bytes, int32_t *h_launches);
/* Stage 1 of the two-stage path alone (dense -> band of width 32, 128 <= N <= 16384; when the
 * two-stage path is active arcb200_sytrd_batched with cluster = 0 runs both of its stages and
 * returns the tridiagonal d, e; tau is then per band-reduction reflector).
 * d_band[b][j][d] = B[j+d][j], d = 0..32 (batch x N x 33). */
int arcb200_sy2sb_batched(int device, void *stream, int32_t N, int32_t batch,
                          const double *d_A, double *d_band, void *d_work, int64_t work_bytes,
                          int32_t *h_launches);

#ifdef __cplusplus
}
#endif
#endif /* ARCB200_H */
