// resonance.cu -- pair-energy scan of StarkMapResonances.findResonances
// (calculations_atom_pairstate.py:4753-4789): for every field i and every pair (j, jj) of
// single-atom Stark eigenstates,
//     energy = y1[i][j] + y2[i][jj] - pairStateEnergy,
// keep the pair if eMin < energy < eMax and both eigenstates are led by a basis state with
// |l - l_state| = 1 (ok1 / ok2, evaluated by the host from the top-1 composition index).
// The reference walks the O(N1 N2) pairs per field in Python; here one warp owns a row (i, j) and
// compacts its hits in order (ballot prefix), so the output order equals the reference's nested
// loops.  Two passes: count per row, then (after an exclusive scan by the host) fill.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(128) resonance_scan_kernel(
    int nF, int N1, int N2, const double *__restrict__ y1, const double *__restrict__ y2,
    const uint8_t *__restrict__ ok1, const uint8_t *__restrict__ ok2, double e0, double eMin,
    double eMax, const int64_t *__restrict__ rowoff /* null: count pass */,
    int32_t *__restrict__ rowcount, double *__restrict__ out_e, int32_t *__restrict__ out_j,
    int32_t *__restrict__ out_jj)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= (long long)nF * N1) return;
    const int i = (int)(row / N1), j = (int)(row % N1);
    if (!ok1[row]) {
        if (!rowoff && lane == 0) rowcount[row] = 0;
        return;
    }
    const double ya = y1[row];
    const double *yb = y2 + (size_t)i * N2;
    const uint8_t *okb = ok2 + (size_t)i * N2;
    int64_t pos = rowoff ? rowoff[row] : 0;
    int count = 0;
    for (int j0 = 0; j0 < N2; j0 += 32) {
        const int jj = j0 + lane;
        bool hit = false;
        double energy = 0.0;
        if (jj < N2 && okb[jj]) {
            energy = __dsub_rn(__dadd_rn(ya, yb[jj]), e0);      // (y1 + y2) - E0 like the reference
            hit = (energy > eMin) && (energy < eMax);
        }
        const unsigned m = __ballot_sync(ARCB200_FULL_MASK, hit);
        if (rowoff && hit) {
            const int64_t o = pos + __popc(m & ((1u << lane) - 1u));
            out_e[o] = energy;
            out_j[o] = j;
            out_jj[o] = jj;
        }
        pos += __popc(m);
        count += __popc(m);
    }
    if (!rowoff && lane == 0) rowcount[row] = count;
}

}  // namespace

extern "C" int arcb200_resonance_scan(int device, void *stream, int32_t nF, int32_t N1, int32_t N2,
                                      const double *d_y1, const double *d_y2, const uint8_t *d_ok1,
                                      const uint8_t *d_ok2, double e0, double eMin, double eMax,
                                      const int64_t *d_rowoff, int32_t *d_rowcount, double *d_out_e,
                                      int32_t *d_out_j, int32_t *d_out_jj)
{
    if (nF <= 0 || N1 <= 0 || N2 <= 0) return 0;
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    const long long rows = (long long)nF * N1;
    resonance_scan_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        nF, N1, N2, d_y1, d_y2, d_ok1, d_ok2, e0, eMin, eMax, d_rowoff, d_rowcount, d_out_e, d_out_j,
        d_out_jj);
    ARCB200_LAUNCH_CHECK("resonance_scan_kernel");
    return 0;
}
