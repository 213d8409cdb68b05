// eig_sy2sb_multi.cu -- band reduction (stage 1 of the two-stage tridiagonalisation) with one
// launch per phase and panel, many CTAs per matrix.  sy2sb_kernel (eig_twostage.cu) gives every
// matrix one CTA, which needs >= 148 matrices in flight; matrices of order ~10^4 (BASELINE
// configs[4], 4.3 GB of workspace each) come in batches of 18, and the one-stage cluster kernel
// streams (4/3) N^3 = 1.5 TB per matrix at a third of the HBM peak (0.67 s per matrix measured).
#include "eig_sy2sb_phases.cuh"

namespace {
using namespace ts;

// One panel step of the band reduction, split into phases that are separate launches:
//   PHASE 0  panel QR + T factor        grid (1, nmat)
//   PHASE 1  X = A22 V                  grid (groups of WARPS block rows, nmat)
//   PHASE 2  W = X T - 1/2 V (...)      grid (1, nmat)
//   PHASE 3  A22 -= V W^T + W V^T       grid (groups, nmat)
// The phase bodies are the shared device functions of eig_sy2sb_phases.cuh; here the block rows of one
// matrix are spread over many CTAs, which is what a batch of a few N ~ 10^4 matrices needs.
template <int WARPS, int PHASE>
__global__ void __launch_bounds__(WARPS * 32, 1) sy2sb_phase_kernel(TsArgs a, int p)
{
    static_assert(WARPS == 8, "the operand ring is laid out for eight warps");
    constexpr int TSTR = S1_TSTR;
    constexpr int TILES = 2 * WARPS * 32 * TSTR;
    const int mat = blockIdx.y;
    double *slot = a.base + (size_t)mat * a.slot;
    extern __shared__ double sm[];
    S1Ctx c;
    c.tid = threadIdx.x; c.wid = c.tid >> 5; c.lane = c.tid & 31; c.grp = c.lane >> 2; c.tig = c.lane & 3;
    c.np = a.Npad; c.lda = a.Npad; c.NB = a.NB;
    c.A = slot + a.oA; c.Wp = slot + a.oWp; c.Vp = slot + a.oVp; c.tt = slot + a.oTau; c.Tfac = slot + a.oTfac;
    // shared-memory carve of sy2sb_ring_kernel: [transit tiles][operand ring / W-phase operand tiles | T, G, M1][small][barriers]
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
    if (PHASE == 0) {
        s1_panel_qr<WARPS>(c, p);
        s1_tfactor<WARPS>(c, p);
    } else if (PHASE == 2) {
        // the T factor of this panel comes back from global memory (phase 0 ran in another launch)
        for (int t = c.tid; t < 1024; t += WARPS * 32) c.Ts[t >> 5][t & 31] = c.Tfac[(size_t)p * 1024 + t];
        s1_wphase<WARPS>(c, p);
    } else {
        // tensor phases: this CTA's row group, barrier-free (per-warp tile rings + shared operand ring)
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
        unsigned gu = 0;
        if (PHASE == 1) s1_symm_ring(c, r, p, gu, (int)blockIdx.x, 1);
        else s1_syr2k_ring(c, r, p, gu, (int)(gridDim.x - 1 - blockIdx.x), 1);      // longest row groups first
    }
}

// Panel QR + T factor alone, two CTAs per SM (<= 128 registers, 93 KB of shared memory): the column loop is a
// chain of block barriers and L2 round trips, so two matrices per SM overlap almost perfectly
constexpr int QR_TWD = 32 * S1_TSTR;
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) sy2sb_qr_kernel(TsArgs a, int p)
{
    const int mat = blockIdx.y;
    double *slot = a.base + (size_t)mat * a.slot;
    extern __shared__ double sm[];
    S1Ctx c;
    c.tid = threadIdx.x; c.wid = c.tid >> 5; c.lane = c.tid & 31; c.grp = c.lane >> 2; c.tig = c.lane & 3;
    c.np = a.Npad; c.lda = a.Npad; c.NB = a.NB;
    c.A = slot + a.oA; c.Wp = slot + a.oWp; c.Vp = slot + a.oVp; c.tt = slot + a.oTau; c.Tfac = slot + a.oTfac;
    c.Bs = nullptr;
    double *half2 = sm + WARPS * QR_TWD;
    c.Ts = reinterpret_cast<double(*)[33]>(half2);
    c.G = reinterpret_cast<double(*)[33]>(half2 + 1056);
    c.M1 = nullptr;
    double *small = half2 + 2 * 1056;
    c.wv = reinterpret_cast<double(*)[32]>(small);
    c.sred = reinterpret_cast<double(*)[16]>(small + 64);
    c.taus = small + 96;
    c.Tw = sm + c.wid * QR_TWD;
    s1_panel_qr<WARPS, QR_TWD>(c, p);
    s1_tfactor<WARPS>(c, p);
}

// compact lower band (N x N part) after the last panel; grid (blocks, nmat)
__global__ void __launch_bounds__(256) band_extract_kernel(TsArgs a)
{
    const int mat = blockIdx.y;
    double *slot = a.base + (size_t)mat * a.slot;
    const double *A = slot + a.oA;
    double *__restrict__ Bd = slot + a.oBand;
    const int N = a.N, np = a.Npad, lda = a.Npad;
    const int total = EIG_LDB * (np + 64);
    for (int idx = blockIdx.x * 256 + threadIdx.x; idx < total; idx += gridDim.x * 256) {
        const int d = idx % EIG_LDB, j = idx / EIG_LDB;
        double x = 0.0;
        if (d <= 32 && j + d < N) x = A[(j + d) + (size_t)j * lda];
        Bd[idx] = x;
    }
}

}  // namespace

int eig_sy2sb_multi(const ts::TsArgs &a, int nmat, cudaStream_t st, int *launches)
{
    constexpr int W = 8;
    const size_t smem = (size_t)(2 * W * 32 * 36 + 2 * 4608 + 128) * sizeof(double) + 256;
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_phase_kernel<W, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_phase_kernel<W, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_phase_kernel<W, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_phase_kernel<W, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t smem_qr = (size_t)(W * QR_TWD + 2 * 1056 + 128) * sizeof(double);
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sy2sb_qr_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_qr));
    // ARCB200_SY2SB_QR2 = 0: panel QR in the one-CTA-per-SM phase kernel (as before)
    const char *eq = getenv("ARCB200_SY2SB_QR2");
    const bool qr2 = !(eq && eq[0] == '0');
    int n = 0;
    const bool dbg = getenv("ARCB200_DEBUG_SYNC") != nullptr;
    auto check = [&](int phase, int p) {
        if (!dbg) return 0;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            arcb200_set_error("sy2sb multi: phase %d of panel %d failed: %s", phase, p, cudaGetErrorString(e));
            return -2;
        }
        return 0;
    };
    if (check(-1, -1)) return -2;
    // Sub-batches on internal streams: the phases of one sub-batch are serial, but the tail of every launch (row
    // groups of different lengths, one CTA per SM) and the latency-bound QR launches are filled by the kernels of
    // the other sub-batches (tools/ab_multi.py, 1 / 2 / 4 / 8 streams: 2363 x296 251.2 / 239.8 / 239.7 / 236.9 ms,
    // 2363 x125 119.5 / 114.6 / 109.4 / 109.0, 651 x286 10.7 / 9.9 / 9.7 / 9.7, 410 x300 5.4 / 4.3 / 3.9 / 3.8).
    // ARCB200_SY2SB_STREAMS = 1..8 overrides the choice below.
    constexpr int MAXS = ArcHelperStreams::MAXS;
    int dev = 0;
    cudaGetDevice(&dev);
    ArcHelperGuard guard;              // held while this call enqueues on the device's helper streams
    ArcHelperStreams *hp = guard.pool(dev);
    if (!hp) return -1;
    cudaStream_t *istream = hp->stream;
    cudaEvent_t *iev = hp->ev;
    // (eight streams only pay for large matrices: 651 x286 and 410 x300 are as fast with four and half the launches)
    int ns = (nmat >= 256 && a.N > 1024) ? 8 : (nmat >= 64 ? 4 : (nmat >= 16 ? 2 : 1));
    const char *esn = getenv("ARCB200_SY2SB_STREAMS");
    if (esn && esn[0] >= '1' && esn[0] <= '8') ns = esn[0] - '0';
    if (ns > nmat) ns = nmat;
    if (dbg) ns = 1;
    cudaStream_t sst[MAXS];
    TsArgs sa[MAXS];
    int sn[MAXS];
    if (ns > 1) ARCB200_CUDA_CHECK(cudaEventRecord(iev[MAXS], st));
    for (int s = 0; s < ns; ++s) {
        const int lo = (int)((long long)nmat * s / ns), hi = (int)((long long)nmat * (s + 1) / ns);
        sa[s] = a;
        sa[s].base = a.base + (size_t)lo * a.slot;
        sn[s] = hi - lo;
        sst[s] = ns > 1 ? istream[s] : st;
        if (ns > 1) ARCB200_CUDA_CHECK(cudaStreamWaitEvent(sst[s], iev[MAXS], 0));
    }
    for (int p = 0; p < a.NB - 1; ++p) {
        const int groups = (a.NB - (p + 1) + W - 1) / W;
        for (int s = 0; s < ns; ++s) {
            if (qr2) sy2sb_qr_kernel<W><<<dim3(1, sn[s]), W * 32, smem_qr, sst[s]>>>(sa[s], p);
            else sy2sb_phase_kernel<W, 0><<<dim3(1, sn[s]), W * 32, smem, sst[s]>>>(sa[s], p);
            if (check(0, p)) return -2;
            sy2sb_phase_kernel<W, 1><<<dim3(groups, sn[s]), W * 32, smem, sst[s]>>>(sa[s], p);
            if (check(1, p)) return -2;
            sy2sb_phase_kernel<W, 2><<<dim3(1, sn[s]), W * 32, smem, sst[s]>>>(sa[s], p);
            if (check(2, p)) return -2;
            sy2sb_phase_kernel<W, 3><<<dim3(groups, sn[s]), W * 32, smem, sst[s]>>>(sa[s], p);
            if (check(3, p)) return -2;
            n += 4;
        }
    }
    if (ns > 1) {
        for (int s = 0; s < ns; ++s) {
            ARCB200_CUDA_CHECK(cudaEventRecord(iev[s], sst[s]));
            ARCB200_CUDA_CHECK(cudaStreamWaitEvent(st, iev[s], 0));
        }
    }
    band_extract_kernel<<<dim3(64, nmat), 256, 0, st>>>(a);
    ARCB200_LAUNCH_CHECK("sy2sb multi-CTA kernels");
    if (launches) *launches += n + 1;
    return 0;
}
