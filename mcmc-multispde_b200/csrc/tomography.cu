// the tomography matrix H on the device, FP64.
// Reference: Sinogram.get_measurement_matrix (/root/reference/mcmc/image_cupy.py:364-396) launching the numba kernel
// _calculate_H_Tomography (/root/reference/mcmc/util_cupy.py:445-470), then H *= n_r/(2*(n_r//2)) (377-378) and the
// "temporary fixing" roll of the first n_theta//4 angle blocks (381-385).  One thread per element, stores coalesced
// along the Fourier index; the per-row trigonometry is hoisted (one sincos per row instead of per element).
#include "common.cuh"
#include <math_constants.h>

namespace mspde {

namespace {

__global__ void __launch_bounds__(256) tomography_H_kernel(const double* __restrict__ r, const double* __restrict__ theta,
                                                           const int* __restrict__ ix, const int* __restrict__ iy, int M,
                                                           int N, int n_r, int n_theta, int apply_fixes,
                                                           cplx* __restrict__ H, int ldh) {
    const int m = blockIdx.y;                       // source row of the reference kernel
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double pi = CUDART_PI;
    const double theta_m = theta[M - 1 - m] + 0.5 * pi;          // theta[-(m+1)] + pi/2   (util_cupy.py:457)
    double sT, cT;
    sincos(theta_m, &sT, &cT);
    const double r_m = r[m];
    const double kx = double(ix[n]), ky = double(iy[n]);
    const double ku = kx * cT + ky * sT;                          // 463
    const double kv = -kx * sT + ky * cT;                         // 464
    const double l = sqrt(0.25 - r_m * r_m);                      // 465
    double ps, pc;
    sincos(pi * ((kx + ky) - 2.0 * kv * r_m), &ps, &pc);          // exp(1j*pi*(...))
    double amp;
    if (kv * kv > 0.0) amp = sin(2.0 * pi * ku * l) / (pi * ku);  // 466-468
    else amp = 2.0 * l;                                           // 470
    cplx v = make_double2(pc * amp, ps * amp);
    int row = m;
    if (apply_fixes) {
        const double ratio = double(n_r) / double(2 * (n_r / 2)); // image_cupy.py:377-378
        v.x *= ratio;
        v.y *= ratio;
        const int b = m / n_r, i = m % n_r;
        if (b < n_theta / 4) row = b * n_r + (i + 1) % n_r;       // cp.roll(block, 1, axis=0): out[(i+1) % n_r] = in[i]
    }
    H[size_t(row) * ldh + n] = v;
}

}  // namespace

int tomography_H(cudaStream_t st, int M, int N, const double* r, const double* theta, const int* ix, const int* iy, int n_r,
                 int n_theta, int apply_fixes, cplx* H, int ldh) {
    dim3 grid((N + 255) / 256, M);
    tomography_H_kernel<<<grid, 256, 0, st>>>(r, theta, ix, iy, M, N, n_r, n_theta, apply_fixes, H, ldh);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mspde
