// eig_sb2st_split.cu -- stage 2 of the two-stage tridiagonalisation (band -> tridiagonal by bulge
// chasing) with every sweep split between TWO warps.
//
// A chase step of sweep s at position r = s + 1 + 32 t touches two disjoint 32x32 blocks of the band:
//     D = B[r .. r+31, r .. r+31]      <- H D H            (two-sided, lower triangle)
//     E = B[r+32 .. r+63, r .. r+31]   <- H' (E H)         (creates the bulge, H' annihilates its first column)
// Both need only the reflector H = I - tau v v^T of the step; the NEXT reflector comes out of E alone.  The
// dependent chain of a sweep is therefore E -> v' -> E' -> v'' ...; the D updates are side work.  In
// sb2st_kernel (eig_twostage.cu) one warp does both in sequence, 8 warps per SM at ~200 registers, and the
// SM issues on 13 % of its cycles (profiles/r01_stage23_c3_ncu_full_summary.json).  Here warp q (0..7) runs
// the E chain of sweeps q, q + 8, ... and warp q + 8 the D blocks of the same sweeps, reading the
// reflectors from an 8-deep ring in shared memory: 16 warps at <= 128 registers, the chain per step about
// half as long, the D work overlapped.
//
// Ordering between sweeps (all counters in shared memory, published after a block-level fence):
//     nv[s] = number of reflectors of sweep s published = E steps completed + 1
//     cd[s] = D steps of sweep s completed
//   E step t of sweep s reads rows r+32..r+63, columns r..r+31; sweep s-1 writes there in its E steps t, t+1
//   and its D step t+1                                       -> nv[s-1] >= t + 3, cd[s-1] >= t + 2
//   D step t of sweep s overlaps sweep s-1's D steps t, t+1 and E step t
//                                                            -> nv[s-1] >= t + 2, cd[s-1] >= t + 2
//   the first reflector (column s) needs sweep s-1's first D and E steps (implied by the E-step-0 wait).
// Same arithmetic, step for step, as sb2st_kernel (the reference-parity tests of the two-stage path hold for both).
#include "eig_twostage.cuh"

namespace {

using namespace ts;

constexpr int SP_VRING = 8;                 // reflectors buffered between the E warp and the D warp
constexpr int SP_DONE = 4095;

__device__ __forceinline__ double sp_warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(ARCB200_FULL_MASK, x, o);
    return x;
}

__device__ __forceinline__ void sp_house(double x, int lane, double &v, double &tau, double &beta)
{
    const double alpha = __shfl_sync(ARCB200_FULL_MASK, x, 0);
    const double xn2 = sp_warp_sum(lane >= 1 ? x * x : 0.0);
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

__device__ __forceinline__ void sp_wait(volatile int *p, int need)
{
    while (*p < need) __nanosleep(20);
    __threadfence_block();
}

__device__ __forceinline__ void sp_publish(volatile int *p, int value, int lane)
{
    __threadfence_block();
    __syncwarp();
    if (lane == 0) *p = value;
}

// SP_PAIRS = sweeps in flight per matrix (E warp + D warp each); 8 / SP_PAIRS CTAs share an SM.  A sweep has
// N / 32 steps and trails its predecessor by two, so a small matrix cannot keep 8 sweeps busy: N <= 1024 runs
// 4 pairs x 2 matrices per SM, N <= 512 2 pairs x 4 matrices per SM (when the batch has that many matrices).
template <int SP_PAIRS>
__global__ void __launch_bounds__(2 * SP_PAIRS * 32, 8 / SP_PAIRS) sb2st_split_kernel(TsArgs a)
{
    constexpr int SP_THREADS = 2 * SP_PAIRS * 32;
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
    const int q = wid & (SP_PAIRS - 1);
    const bool dwarp = wid >= SP_PAIRS;
    double *tile = sm + (size_t)wid * (32 * 33);
    double *vs = sm + 2 * SP_PAIRS * 32 * 33 + wid * 32;
    double *ws = sm + 2 * SP_PAIRS * 32 * 33 + 2 * SP_PAIRS * 32 + wid * 32;
    double *vring = sm + 2 * SP_PAIRS * 32 * 33 + 4 * SP_PAIRS * 32 + q * (SP_VRING * 33);    // [k][0..31] v, [k][32] tau
    __shared__ volatile int nv[64], cd[64];
    if (tid < 64) { nv[tid] = -1; cd[tid] = -1; }
    for (size_t i = tid; i < (size_t)a.Npad * TM; i += SP_THREADS) tau2[i] = 0.0;
    __syncthreads();

    constexpr int CS = EIG_LDB - 1;
    for (int s = q; s < N - 2; s += SP_PAIRS) {
        const int me = s & 63, prev = (s - 1) & 63;
        const int base = s * 4096, pbase = (s - 1) * 4096;
        if (!dwarp) {
            // ------------------------------------------------------------ E chain
            // the D warp of this pair has to be out of the sweep that used the reflector ring before
            if (s >= SP_PAIRS) sp_wait(&cd[(s - SP_PAIRS) & 63], (s - SP_PAIRS) * 4096 + SP_DONE);
            if (s > 0) { sp_wait(&nv[prev], pbase + 2); sp_wait(&cd[prev], pbase + 1); }     // column s is final
            int t = 0, r = s + 1, L = min(32, N - r);
            double v, tau, beta;
            {
                double *c0 = Bd + (size_t)s * EIG_LDB + 1 + lane;          // element (r + lane, s)
                const double x = (lane < L) ? *c0 : 0.0;
                sp_house(x, lane, v, tau, beta);
                if (lane < L) *c0 = (lane == 0) ? beta : 0.0;
            }
            while (true) {
                if (lane < L) V2[(size_t)(r + lane) + (size_t)s * ldv] = v;
                if (lane == 0) tau2[(size_t)s * TM + t] = tau;
                // reflector t to the D warp (slot t % 8 is free once D step t - 8 is done)
                if (t >= SP_VRING) sp_wait(&cd[me], base + t - SP_VRING + 1);
                vring[(t & (SP_VRING - 1)) * 33 + lane] = v;
                if (lane == 0) vring[(t & (SP_VRING - 1)) * 33 + 32] = tau;
                sp_publish(&nv[me], base + t + 1, lane);
                const int r2 = r + L, L2 = min(32, N - r2);
                if (L2 <= 0) break;
                if (s > 0) { sp_wait(&nv[prev], pbase + t + 3); sp_wait(&cd[prev], pbase + t + 2); }
                double *cp = Bd + (size_t)r * EIG_LDB + 32 + lane;         // element (r2 + lane, r + j) at cp[j * CS]
                double El[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) El[j] = (lane < L2) ? cp[j * CS] : 0.0;
                vs[lane] = v;
                __syncwarp();
                const double2 *vs2 = reinterpret_cast<const double2 *>(vs);
                const double2 *ws2 = reinterpret_cast<const double2 *>(ws);
                if (tau != 0.0) {
                    // E <- E H
                    double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const double2 p0 = vs2[j / 2], p1 = vs2[j / 2 + 1];
                        y0 += El[j] * p0.x;
                        y1 += El[j + 1] * p0.y;
                        y2 += El[j + 2] * p1.x;
                        y3 += El[j + 3] * p1.y;
                    }
                    const double y = tau * ((y0 + y1) + (y2 + y3));
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const double2 p0 = vs2[j / 2];
                        El[j] -= y * p0.x;
                        El[j + 1] -= y * p0.y;
                    }
                }
                // next reflector from the first column of the bulge; E <- H' E
                double v2, tau_n, beta2;
                sp_house(El[0], lane, v2, tau_n, beta2);
                El[0] = (lane == 0) ? beta2 : 0.0;
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
                    __syncwarp();
                    ws[lane] = (lane == 0) ? 0.0 : ((s0 + s1) + (s2 + s3)) * tau_n;
                    __syncwarp();
#pragma unroll
                    for (int j = 2; j < 32; j += 2) {
                        const double2 b = ws2[j / 2];
                        El[j] -= v2 * b.x;
                        El[j + 1] -= v2 * b.y;
                    }
                    El[1] -= v2 * ws[1];
                }
                if (lane < L2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) cp[j * CS] = El[j];
                }
                __syncwarp();
                r = r2; L = L2; v = v2; tau = tau_n;
                ++t;
            }
            sp_publish(&nv[me], base + SP_DONE, lane);
        } else {
            // ------------------------------------------------------------ D blocks
            const int T = (N - s - 1 + 31) >> 5;
            for (int t = 0; t < T; ++t) {
                const int r = s + 1 + 32 * t, L = min(32, N - r);
                sp_wait(&nv[me], base + t + 1);
                const double v = vring[(t & (SP_VRING - 1)) * 33 + lane];
                const double tau = vring[(t & (SP_VRING - 1)) * 33 + 32];
                if (tau != 0.0) {
                    if (s > 0) { sp_wait(&nv[prev], pbase + t + 2); sp_wait(&cd[prev], pbase + t + 2); }
                    double *cp = Bd + (size_t)r * EIG_LDB + lane;          // element (r + lane, r + j) at cp[j * CS]
                    double Dl[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) Dl[j] = (j <= lane && lane < L) ? cp[j * CS] : 0.0;
                    vs[lane] = v;
                    __syncwarp();
                    const double2 *vs2 = reinterpret_cast<const double2 *>(vs);
                    const double2 *ws2 = reinterpret_cast<const double2 *>(ws);
                    // D <- H D H on the lower triangle; four partial sums per reduction
                    double pd0 = 0.0, pd1 = 0.0, pd2 = 0.0, pd3 = 0.0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const double2 p0 = vs2[j / 2], p1 = vs2[j / 2 + 1];
                        pd0 += Dl[j] * p0.x;
                        pd1 += Dl[j + 1] * p0.y;
                        pd2 += Dl[j + 2] * p1.x;
                        pd3 += Dl[j + 3] * p1.y;
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
                    const double al = 0.5 * tau * sp_warp_sum(pv * v);
                    const double w = pv - al * v;
                    ws[lane] = w;
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const double2 p0 = vs2[j / 2], b = ws2[j / 2];
                        const double x0 = Dl[j] - (v * b.x + w * p0.x);
                        const double x1 = Dl[j + 1] - (v * b.y + w * p0.y);
                        if (j <= lane && lane < L) cp[j * CS] = x0;
                        if (j + 1 <= lane && lane < L) cp[(j + 1) * CS] = x1;
                    }
                    __syncwarp();
                }
                sp_publish(&cd[me], base + t + 1, lane);
            }
            sp_publish(&cd[me], base + SP_DONE, lane);
        }
    }
    __syncthreads();
    for (int i = tid; i < N; i += SP_THREADS) {
        dd[i] = Bd[(size_t)i * EIG_LDB];
        ee[i] = (i < N - 1) ? Bd[1 + (size_t)i * EIG_LDB] : 0.0;
    }
}

template <int SP_PAIRS>
static int launch_split(const ts::TsArgs &a, int nmat, cudaStream_t st)
{
    const size_t smem = (size_t)(2 * SP_PAIRS * 32 * 33 + 4 * SP_PAIRS * 32 + SP_PAIRS * SP_VRING * 33) * sizeof(double);
    ARCB200_CUDA_CHECK(cudaFuncSetAttribute(sb2st_split_kernel<SP_PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sb2st_split_kernel<SP_PAIRS><<<nmat, 2 * SP_PAIRS * 32, smem, st>>>(a);
    ARCB200_LAUNCH_CHECK("sb2st_split_kernel");
    return 0;
}

}  // namespace

int eig_sb2st_split(const ts::TsArgs &a, int nmat, cudaStream_t st, int *launches)
{
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    int pairs = 8;
    if (a.N <= 1024 && nmat > sms) pairs = 4;
    if (a.N <= 512 && nmat > 2 * sms) pairs = 2;
    const char *ep = getenv("ARCB200_SB2ST_PAIRS");        // tuning knob: 2 | 4 | 8
    if (ep && (ep[0] == '2' || ep[0] == '4' || ep[0] == '8')) pairs = ep[0] - '0';
    const int rc = pairs == 8 ? launch_split<8>(a, nmat, st) : pairs == 4 ? launch_split<4>(a, nmat, st) : launch_split<2>(a, nmat, st);
    if (rc) return rc;
    if (launches) *launches += 1;
    return 0;
}
