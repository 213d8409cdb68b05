// anger.cu -- semiclassical radial dipole / quadrupole elements (divalent atoms).
//
// Replaces arc/alkali_atom_functions.py:2683-2735 (_getRadialDipoleSemiClassical) and
// :2737-2802 (_getRadialQuadrupoleSemiClassical), which call mpmath.angerj.  The Anger
// function J_nu(x) = (1/pi) int_0^pi cos(nu t - x sin t) dt is evaluated here by
// Gauss-Legendre quadrature (nodes/weights supplied by the host, nq a multiple of 32),
// one warp per matrix element, both orders nu = dnu -+ 1 sharing sin(t).
#include <math.h>
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(128)
semiclassical_kernel(int64_t n, const double *__restrict__ in, int kind,
                     const double *__restrict__ quad, int nq, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= n) return;
    const double nu = in[4 * w], nu1 = in[4 * w + 1];
    const double l1 = in[4 * w + 2], l2 = in[4 * w + 3];
    const double l_c = (l1 + l2 + 1.0) / 2.0;
    const double nu_c = sqrt(nu * nu1);
    const double dn = nu - nu1;
    const double gamma = ((l2 - l1) * l_c) / nu_c;
    const double PI = 3.14159265358979323846;

    double g0 = 1.0, g1 = 0.0;
    if (dn != 0.0) {
        // J_{dn-1}(-dn), J_{dn+1}(-dn):  phase = nu t + dn sin t
        double sm = 0.0, sp = 0.0;
        for (int q = lane; q < nq; q += 32) {
            const double t = 0.5 * PI * (quad[q] + 1.0);
            const double wq = quad[nq + q];
            const double st = sin(t);
            sm += wq * cos((dn - 1.0) * t + dn * st);
            sp += wq * cos((dn + 1.0) * t + dn * st);
        }
        sm = warp_sum(sm) * 0.5;   // (1/pi) * (pi/2) * sum
        sp = warp_sum(sp) * 0.5;
        g0 = (1.0 / (3.0 * dn)) * (sm - sp);
        g1 = -(1.0 / (3.0 * dn)) * (sm + sp);
    }
    double res;
    if (kind == 1) {
        double g2 = 0.0, g3 = 0.0;
        if (dn != 0.0) {
            g2 = g0 - sin(PI * dn) / (PI * dn);
            g3 = (dn / 2.0) * g0 + g1;
        }
        const double lr = l_c / nu_c;
        res = 1.5 * nu_c * nu_c * sqrt(1.0 - lr * lr)
              * (g0 + gamma * g1 + gamma * gamma * g2 + gamma * gamma * gamma * g3);
    } else {
        double q0 = 1.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
        if (dn != 0.0) {
            q0 = -(6.0 / (5.0 * dn)) * g1;
            q1 = -(6.0 / (5.0 * dn)) * g0 + (6.0 / 5.0) * sin(PI * dn) / (PI * dn * dn);
            q2 = -(3.0 / 4.0) * (6.0 / (5.0 * dn) * g1 + g0);
            q3 = 0.5 * (dn * 0.5 * q0 + q1);
        }
        const double dl = fabs(l2 - l1);
        const double nc2 = nu_c * nu_c;
        if (dl == 0.0) {
            res = 2.5 * nc2 * nc2 * (1.0 - (3.0 * l_c * l_c) / (5.0 * nc2))
                  * (q0 + gamma * gamma * q2);
        } else if (dl == 2.0) {
            res = 2.5 * nc2 * nc2 * sqrt(1.0 - (l_c + 1.0) * (l_c + 1.0) / nc2)
                  * sqrt(1.0 - (l_c + 2.0) * (l_c + 2.0) / nc2)
                  * (q0 + gamma * q1 + gamma * gamma * q2 + gamma * gamma * gamma * q3);
        } else {
            res = 0.0;
        }
    }
    if (lane == 0) out[w] = res;
}

}  // namespace

extern "C" int arcb200_semiclassical_table(int device, void *stream, int64_t n,
                                           const double *d_in, int32_t kind,
                                           const double *d_quad, int32_t nq,
                                           double *d_out)
{
    if (n <= 0) return 0;
    if (kind != 1 && kind != 2) { arcb200_set_error("kind must be 1 or 2"); return -1; }
    if (nq <= 0 || (nq & 31)) { arcb200_set_error("nq must be a positive multiple of 32"); return -1; }
    ARCB200_REQUIRE(d_in && d_quad && d_out, "semiclassical_table: null pointer");
    ARCB200_CUDA_CHECK(cudaSetDevice(device));
    const int64_t grid = (n + 3) / 4;
    semiclassical_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(n, d_in, kind, d_quad, nq, d_out);
    ARCB200_LAUNCH_CHECK("semiclassical_kernel");
    return 0;
}
