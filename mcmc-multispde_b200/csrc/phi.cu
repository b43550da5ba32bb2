// a9: Phi(L) = 1/2 y^T Q^-1 y + 1/2 log det Q in the reference's Woodbury form
// (pCN._log_ratio, /root/reference/mcmc/image_cupy.py:633-643):
//   r = L^H L + HtH + delta I ; c = chol(r) ; Ht = c^-1 H^H ; Q_inv = I - Ht^H Ht ; C = chol(Q_inv)
//   Phi = 0.5 * || y @ C ||^2 - sum(log(Re diag C))
// Here both Cholesky factors are kept in upper form (r = U^H U, c = U^H; Q_inv = V^H V, C = V^H) so that
// every product is a k-major "TN" zgemm on the DMMA pipe and only the upper triangles are computed
// (HERK-style, half the flops of the reference's full GEMMs).  || y @ C || = || V y || since y is real.
#include "context.cuh"

namespace mspde {

namespace {

// r[i][j] = HtH[i][j] + (i == j) * delta
__global__ void init_r_kernel(const cplx* __restrict__ HtH, int ldh, double delta, cplx* __restrict__ r, int ldr, int N) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N) return;
    cplx v = HtH[size_t(i) * ldh + j];
    if (i == j) v.x += delta;
    r[size_t(i) * ldr + j] = v;
}

__global__ void add_diag_kernel(cplx* __restrict__ A, int lda, int n, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[size_t(i) * lda + i].x += v;
}

// one warp per row j of the upper factor V: z_j = sum_{i>=j} V[j][i] y_i ; rowsq[j] = |z_j|^2 ; rowlog[j] = log V[j][j]
__global__ void __launch_bounds__(256) vy_kernel(const cplx* __restrict__ V, int ldv, const double* __restrict__ y, int M,
                                                 double* __restrict__ rowsq, double* __restrict__ rowlog) {
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= M) return;
    const cplx* row = V + size_t(j) * ldv;
    double sx = 0.0, sy = 0.0;
    for (int i = j + lane; i < M; i += 32) {
        const cplx v = row[i];
        const double yi = y[i];
        sx = fma(v.x, yi, sx);
        sy = fma(v.y, yi, sy);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if (lane == 0) {
        rowsq[j] = sx * sx + sy * sy;
        rowlog[j] = log(row[j].x);
    }
}

// deterministic single-block reduction: out = {0.5*sum(rowsq) - sum(rowlog), 0.5*sum(rowsq), sum(rowlog)}
__global__ void __launch_bounds__(1024) phi_reduce_kernel(const double* __restrict__ rowsq, const double* __restrict__ rowlog,
                                                          int M, double* __restrict__ out3) {
    __shared__ double s1[1024], s2[1024];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < M; i += 1024) {
        a += rowsq[i];
        b += rowlog[i];
    }
    s1[threadIdx.x] = a;
    s2[threadIdx.x] = b;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if (threadIdx.x < o) {
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out3[0] = 0.5 * s1[0] - s2[0];
        out3[1] = 0.5 * s1[0];
        out3[2] = s2[0];
    }
}

}  // namespace

int phi_eval(Ctx* c, cudaStream_t st, PhiWork& w, const cplx* L, int ldl, double* out3, bool serial) {
    const int N = c->N, M = c->M;
    // r = HtH + delta I + L^H L (upper triangle)
    init_r_kernel<<<dim3((N + 255) / 256, N), 256, 0, st>>>(c->HtH, c->ldHtH, c->delta, w.r, c->ldN, N);
    MSPDE_LAUNCH_CHECK();
    int rc = zgemm_tn(st, true, N, N, N, 1.0, L, ldl, L, ldl, 1.0, w.r, c->ldN, GEMM_UPPER);
    if (rc) return rc;
    rc = potrf_upper(st, w.r, c->ldN, N, w.winv, c->info, serial ? nullptr : &w.la);
    if (rc) return rc;
    // X = U^-H H^H  (= c^-1 H^H = Ht)
    MSPDE_CUDA(cudaMemcpy2DAsync(w.X, size_t(c->ldM) * sizeof(cplx), c->Hct, size_t(c->ldHct) * sizeof(cplx),
                                 size_t(M) * sizeof(cplx), N, cudaMemcpyDeviceToDevice, st));
    rc = trsm_uh(st, w.r, c->ldN, w.winv, N, w.X, c->ldM, M);
    if (rc) return rc;
    // Q_inv = I - X^H X (upper triangle)
    rc = zgemm_tn(st, true, M, M, N, -1.0, w.X, c->ldM, w.X, c->ldM, 0.0, w.Q, c->ldM, GEMM_UPPER);
    if (rc) return rc;
    add_diag_kernel<<<(M + 255) / 256, 256, 0, st>>>(w.Q, c->ldM, M, 1.0);
    MSPDE_LAUNCH_CHECK();
    rc = potrf_upper(st, w.Q, c->ldM, M, w.winv, c->info, serial ? nullptr : &w.la);
    if (rc) return rc;
    vy_kernel<<<(M + 7) / 8, 256, 0, st>>>(w.Q, c->ldM, c->y, M, w.rowsq, w.rowlog);
    MSPDE_LAUNCH_CHECK();
    phi_reduce_kernel<<<1, 1024, 0, st>>>(w.rowsq, w.rowlog, M, out3);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mspde
