// a2 + a3 + a4 of the hot path: field synthesis and fused assembly of the Galerkin operator L(u).
//
// Reference: Lmatrix_2D.construct_from (/root/reference/mcmc/image_cupy.py:231-233 -> 171-196),
// FourierAnalysis_2D.irfft2 / rfft2 / constructMatexplicit (108-149), util.constructU_from_uSym2D_cuda and
// the _construct_U numba kernel (/root/reference/mcmc/util_cupy.py:323-330, 426-442).
//
// The reference makes >= 5 full N^2 passes (two zero fills, two uncoalesced scatters, multiply, subtract,
// divide).  Here L is written exactly once: each CTA stages the few table runs it needs (the "index map")
// in shared memory and streams out 4 KB row segments with 16-byte stores.  m = 2 n_ext - 1 = 127 is
// prime and the transforms are tiny, so the 2-D transforms are done as separable DFT sums in FP64.
#include <cstdlib>
#include "context.cuh"

namespace mspde {

namespace {

__device__ __forceinline__ int modm(int a, int m) {
    int r = a % m;
    return r < 0 ? r + m : r;
}

// Hermitian-symmetrised spectrum of util.symmetrize_2D(uHalf2D) read straight from the sym vector:
// sym2D[r][c] = u[c*il + r] (reshape 'F'); columns c < n-1 are the conjugate mirror.
__device__ __forceinline__ cplx sfull(const cplx* __restrict__ u, int r, int c, int n, int il) {
    if (c >= n - 1) return u[c * il + r];
    cplx v = u[(il - 1 - c) * il + (il - 1 - r)];
    return make_double2(v.x, -v.y);
}

// T1[x][c] = sum_r S[r][c] W^{(r-(n-1)) x}
__global__ void field_stage1(const cplx* __restrict__ u, const cplx* __restrict__ W, cplx* __restrict__ T1, int n,
                             int il, int m) {
    const int x = blockIdx.x;
    for (int c = threadIdx.x; c < il; c += blockDim.x) {
        cplx acc = make_double2(0.0, 0.0);
        for (int r = 0; r < il; ++r) acc = cfma(sfull(u, r, c, n, il), W[modm((r - (n - 1)) * x, m)], acc);
        T1[x * il + c] = acc;
    }
}

// img[x][y] = (1/m) Re sum_c T1[x][c] W^{(c-(n-1)) y};  fp = exp(img), fm = exp(-img)
__global__ void field_stage2(const cplx* __restrict__ T1, const cplx* __restrict__ W, double* __restrict__ fp,
                             double* __restrict__ fm, double* __restrict__ img_out, int n, int il, int m) {
    extern __shared__ cplx srow[];
    const int x = blockIdx.x;
    for (int c = threadIdx.x; c < il; c += blockDim.x) srow[c] = T1[x * il + c];
    __syncthreads();
    for (int y = threadIdx.x; y < m; y += blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < il; ++c) {
            cplx w = W[modm((c - (n - 1)) * y, m)];
            acc = fma(srow[c].x, w.x, acc);
            acc = fma(-srow[c].y, w.y, acc);
        }
        const double img = acc / double(m);
        fp[x * m + y] = exp(img);
        fm[x * m + y] = exp(-img);
        if (img_out) img_out[x * m + y] = img;
    }
}

// G[s][x][kc] = sum_y f_s[x][y] conj(W^{kc y}),  s = 0 (exp(+u)), 1 (exp(-u)),  kc = 0..n-1
__global__ void field_stage3(const double* __restrict__ fp, const double* __restrict__ fm, const cplx* __restrict__ W,
                             cplx* __restrict__ G, int n, int m) {
    extern __shared__ double frow[];   // 2 x m
    const int x = blockIdx.x;
    for (int y = threadIdx.x; y < m; y += blockDim.x) {
        frow[y] = fp[x * m + y];
        frow[m + y] = fm[x * m + y];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * n; idx += blockDim.x) {
        const int s = idx / n, kc = idx % n;
        double ax = 0.0, ay = 0.0;
        for (int y = 0; y < m; ++y) {
            cplx w = W[modm(kc * y, m)];
            double f = frow[s * m + y];
            ax = fma(f, w.x, ax);
            ay = fma(-f, w.y, ay);
        }
        G[(size_t(s) * m + x) * n + kc] = make_double2(ax, ay);
    }
}

// F_s[kr][kc] = (1/m) sum_x conj(W^{kr x}) G[s][x][kc], written symmetrised (util.symmetrize_2D) into il x il tables
__global__ void field_stage4(const cplx* __restrict__ G, const cplx* __restrict__ W, cplx* __restrict__ tp,
                             cplx* __restrict__ tm, int n, int il, int m) {
    const int r = blockIdx.x;
    const int kr = r - (n - 1);
    for (int idx = threadIdx.x; idx < 2 * n; idx += blockDim.x) {
        const int s = idx / n, kc = idx % n;
        cplx acc = make_double2(0.0, 0.0);
        for (int x = 0; x < m; ++x) acc = cfmac(W[modm(kr * x, m)], G[(size_t(s) * m + x) * n + kc], acc);
        acc.x /= double(m);
        acc.y /= double(m);
        cplx* T = s == 0 ? tp : tm;
        T[r * il + (kc + n - 1)] = acc;
        if (kc >= 1) T[(il - 1 - r) * il + (n - 1 - kc)] = make_double2(acc.x, -acc.y);
    }
}

constexpr int ACOLS = 256;   // columns of L per CTA (4 KB row segments)

// L[i*il+k][j*il+l] = (Tp[k-l+n-1][i-j+n-1] * D[j*il+l] - Tm[k-l+n-1][i-j+n-1]) * (1/sqrt_beta) inside the band, else 0
__global__ void __launch_bounds__(ACOLS) assemble_L_kernel(const cplx* __restrict__ tp, const cplx* __restrict__ tm,
                                                           const double* __restrict__ D, double inv_sb,
                                                           cplx* __restrict__ L, int ldl, int n, int il, int N) {
    extern __shared__ cplx stab[];   // [2][nj][il]: runs Tp[:, i-j+n-1], Tm[:, i-j+n-1] for the block columns j touched
    const int i = blockIdx.y;
    const int c0 = blockIdx.x * ACOLS;
    const int jlo = c0 / il;
    const int jhi = min((c0 + ACOLS - 1) / il, il - 1);
    const int nj = jhi - jlo + 1;
    cplx* sp = stab;
    cplx* sm = stab + nj * il;
    for (int idx = threadIdx.x; idx < nj * il; idx += ACOLS) {
        const int jj = idx / il, d = idx % il;
        const int dij = i - (jlo + jj);
        cplx vp = make_double2(0.0, 0.0), vm = vp;
        if (dij >= -(n - 1) && dij <= n - 1) {
            vp = tp[d * il + dij + n - 1];
            vm = tm[d * il + dij + n - 1];
        }
        sp[idx] = vp;
        sm[idx] = vm;
    }
    __syncthreads();
    const int col = c0 + threadIdx.x;
    if (col >= N) return;
    const int j = col / il, l = col % il;
    const int base = (j - jlo) * il;
    const double Dc = D[col];
    cplx* out = L + size_t(i) * il * ldl + col;
    // the il rows of this block row are split over gridDim.z CTAs (more CTAs in flight, even tail)
    const int kper = (il + gridDim.z - 1) / gridDim.z;
    const int k0 = blockIdx.z * kper, k1 = min(il, k0 + kper);
    for (int k = k0; k < k1; ++k) {
        const int d = k - l + (n - 1);
        cplx v = make_double2(0.0, 0.0);
        if (d >= 0 && d < il) {
            const cplx a = sp[base + d], b = sm[base + d];
            // multiply, subtract, scale -- no FMA contraction, so L is bit-identical to the oracle's
            // (U+ * D - U-) / sqrt_beta given identical tables (numpy divides a complex by a real scalar as x * (1/s))
            v.x = __dmul_rn(__dsub_rn(__dmul_rn(a.x, Dc), b.x), inv_sb);
            v.y = __dmul_rn(__dsub_rn(__dmul_rn(a.y, Dc), b.y), inv_sb);
        }
        __stcs(reinterpret_cast<double2*>(out + size_t(k) * ldl), v);   // streaming store: L is consumed much later
    }
}

}  // namespace

int field_coeffs(Ctx* c, const cplx* u_sym, cplx* tp_out, cplx* tm_out) {
    cudaStream_t st = c->stream;
    const int n = c->n, il = c->il, m = c->m;
    field_stage1<<<m, 64, 0, st>>>(u_sym, c->W, c->T1, n, il, m);
    MSPDE_LAUNCH_CHECK();
    field_stage2<<<m, 128, il * sizeof(cplx), st>>>(c->T1, c->W, c->fp, c->fm, nullptr, n, il, m);
    MSPDE_LAUNCH_CHECK();
    field_stage3<<<m, 64, 2 * m * sizeof(double), st>>>(c->fp, c->fm, c->W, c->G, n, m);
    MSPDE_LAUNCH_CHECK();
    field_stage4<<<il, 64, 0, st>>>(c->G, c->W, tp_out, tm_out, n, il, m);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

// running mean / M2 (Welford) of the field u(x) = irfft2(sample) and of exp(u(x)) on the device, so that
// histories need not be stored for the posterior statistics of /root/reference/post_analysis.py:57-105
// (util.updateWelford, /root/reference/mcmc/util_cupy.py:168-177).  count is the sample count AFTER this update.
__global__ void welford_kernel(const double* __restrict__ img, const double* __restrict__ ell, double count,
                               double* __restrict__ mean_u, double* __restrict__ m2_u, double* __restrict__ mean_ell,
                               double* __restrict__ m2_ell, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    {
        const double x = img[i], mu = mean_u[i];
        const double delta = x - mu;
        const double mu2 = mu + delta / count;
        mean_u[i] = mu2;
        m2_u[i] += delta * (x - mu2);
    }
    {
        const double x = ell[i], mu = mean_ell[i];
        const double delta = x - mu;
        const double mu2 = mu + delta / count;
        mean_ell[i] = mu2;
        m2_ell[i] += delta * (x - mu2);
    }
}

int posterior_stats_update(Ctx* c, const cplx* u_sym, long long count, double* mean_u, double* m2_u, double* mean_ell,
                           double* m2_ell) {
    cudaStream_t st = c->stream;
    const int n = c->n, il = c->il, m = c->m;
    field_stage1<<<m, 64, 0, st>>>(u_sym, c->W, c->T1, n, il, m);
    MSPDE_LAUNCH_CHECK();
    field_stage2<<<m, 128, il * sizeof(cplx), st>>>(c->T1, c->W, c->fp, c->fm, c->img, n, il, m);
    MSPDE_LAUNCH_CHECK();
    welford_kernel<<<(m * m + 255) / 256, 256, 0, st>>>(c->img, c->fp, double(count), mean_u, m2_u, mean_ell, m2_ell, m * m);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

int assemble_L(Ctx* c, const cplx* tp, const cplx* tm, double sqrt_beta, cplx* L, int ldl) {
    const int il = c->il, N = c->N;
    static const int zsplit = getenv("MSPDE_ASM_Z") ? atoi(getenv("MSPDE_ASM_Z")) : 4;
    dim3 grid((N + ACOLS - 1) / ACOLS, il, zsplit);
    const int njmax = ACOLS / il + 2;
    size_t smem = size_t(2) * njmax * il * sizeof(cplx);
    assemble_L_kernel<<<grid, ACOLS, smem, c->stream>>>(tp, tm, c->D_diag, 1.0 / sqrt_beta, L, ldl, c->n, il, N);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mspde
