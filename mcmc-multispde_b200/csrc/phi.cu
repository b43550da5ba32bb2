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

// one warp per row j of an upper-trapezoidal row panel V (nrows x ncols, diagonal at column 0 of row 0):
// z_j = sum_{i>=j} V[j][i] y_i ; rowsq[j] = |z_j|^2 ; rowlog[j] = log Re V[j][j]   (image_cupy.py:643)
__global__ void __launch_bounds__(256) vy_rows_kernel(const cplx* __restrict__ V, int ldv, const double* __restrict__ y,
                                                      int nrows, int ncols, double* __restrict__ rowsq,
                                                      double* __restrict__ rowlog) {
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= nrows) return;
    const cplx* row = V + size_t(j) * ldv;
    double sx = 0.0, sy = 0.0;
    for (int i = j + lane; i < ncols; i += 32) {
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

// g = H^H y helper input: y as a complex column
__global__ void real_to_cplx_kernel(const double* __restrict__ y, cplx* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(y[i], 0.0);
}

// A = G + delta I (upper), r = G + HtH + delta I (upper) from one Gram matrix G = L^H L stored in r
__global__ void split_gram_kernel(cplx* __restrict__ r, int ldr, const cplx* __restrict__ HtH, int ldh, double delta,
                                  cplx* __restrict__ A, int lda, int N) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N || j < i) return;
    cplx g = r[size_t(i) * ldr + j];
    if (i == j) g.x += delta;
    A[size_t(i) * lda + j] = g;
    const cplx h = HtH[size_t(i) * ldh + j];
    r[size_t(i) * ldr + j] = make_double2(g.x + h.x, g.y + h.y);
}

// out3 = {0.5*(ynorm2 - ||z||^2) + sum log diag(Ur) - sum log diag(Ua), 0.5*(ynorm2 - ||z||^2), -(sum log Ur - sum log Ua)}
__global__ void __launch_bounds__(1024) phi_fast_reduce_kernel(const cplx* __restrict__ Ur, int ldr, const cplx* __restrict__ Ua,
                                                               int lda, const cplx* __restrict__ z, const double* __restrict__ y,
                                                               int N, int M, double* __restrict__ out3) {
    __shared__ double s1[1024], s2[1024], s3[1024];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < N; i += 1024) {
        const cplx zi = z[i];
        a += zi.x * zi.x + zi.y * zi.y;
        b += log(Ur[size_t(i) * ldr + i].x) - log(Ua[size_t(i) * lda + i].x);
    }
    for (int i = threadIdx.x; i < M; i += 1024) c += y[i] * y[i];
    s1[threadIdx.x] = a; s2[threadIdx.x] = b; s3[threadIdx.x] = c;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if (threadIdx.x < o) {
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
            s3[threadIdx.x] += s3[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double quad = 0.5 * (s3[0] - s1[0]);
        out3[0] = quad + s2[0];
        out3[1] = quad;
        out3[2] = -s2[0];
    }
}

}  // namespace

// exported helpers for the block-distributed Phi (mspde/dist.py)
int add_diag(cudaStream_t st, cplx* A, int lda, int n, double v) {
    add_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(A, lda, n, v);
    MSPDE_LAUNCH_CHECK();
    return 0;
}
// rows of an upper-trapezoidal panel V (nrows x ncols, diagonal at column 0) times y: rowsq[j] = |sum_{i>=j} V[j][i] y[i]|^2
int upper_rows_times_y(cudaStream_t st, const cplx* V, int ldv, int nrows, int ncols, const double* y, double* rowsq,
                       double* rowlog) {
    vy_rows_kernel<<<(nrows + 7) / 8, 256, 0, st>>>(V, ldv, y, nrows, ncols, rowsq, rowlog);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

// Results-preserving reformulation of Phi (SURVEY 8d note ii), OPT-IN and reported separately from the reference-faithful
// Woodbury path:  y^T Q_inv y = ||y||^2 - ||U_r^-H H^H y||^2  and, by the matrix determinant lemma,
// -sum log diag C = 1/2 log det(I + H A^-1 H^H) = sum log diag U_r - sum log diag U_A,  A = L^H L + delta I = U_A^H U_A,
// r = A + HtH = U_r^H U_r.  No M-sized product or factorisation is formed: 4N^3 + 2*(4/3)N^3 flops instead of F_Phi.
int phi_fast_prepare(Ctx* c, cudaStream_t st) {
    const int N = c->N, M = c->M;
    for (int s = 0; s < 2; ++s)
        if (!c->pw[s].A) MSPDE_CUDA(cudaMalloc((void**)&c->pw[s].A, size_t(N) * c->ldN * sizeof(cplx)));   // freed with the context
    if (!c->g_valid) {          // g = H^H y, once per set of inputs
        real_to_cplx_kernel<<<(M + 255) / 256, 256, 0, st>>>(c->y, c->pw[0].Q, M);
        MSPDE_LAUNCH_CHECK();
        int rc0 = zgemm_tn(st, true, N, 1, M, 1.0, c->H, c->ldH, c->pw[0].Q, 1, 0.0, c->gvec, 1, GEMM_FULL);
        if (rc0) return rc0;
        c->g_valid = true;
    }
    return 0;
}

int phi_eval_fast(Ctx* c, cudaStream_t st, PhiWork& w, const cplx* L, int ldl, double* out3) {
    const int N = c->N, M = c->M;
    if (!w.A || !c->g_valid) {
        set_error("phi_eval_fast: phi_fast_prepare was not called");
        return -1;
    }
    cplx* A = w.A;
    cplx* z = w.Q;              // first N entries of the Q buffer
    MSPDE_CUDA(cudaMemsetAsync(w.r, 0, size_t(N) * c->ldN * sizeof(cplx), st));
    int rc = zgemm_tn(st, true, N, N, N, 1.0, L, ldl, L, ldl, 0.0, w.r, c->ldN, GEMM_UPPER);
    if (rc) return rc;
    split_gram_kernel<<<dim3((N + 255) / 256, N), 256, 0, st>>>(w.r, c->ldN, c->HtH, c->ldHtH, c->delta, A, c->ldN, N);
    MSPDE_LAUNCH_CHECK();
    rc = potrf_upper(st, A, c->ldN, N, w.winv, c->info, nullptr);
    if (rc) return rc;
    rc = potrf_upper(st, w.r, c->ldN, N, w.winv, c->info, nullptr);
    if (rc) return rc;
    MSPDE_CUDA(cudaMemcpyAsync(z, c->gvec, size_t(N) * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    rc = trsm_uh(st, w.r, c->ldN, w.winv, N, z, 1, 1);
    if (rc) return rc;
    phi_fast_reduce_kernel<<<1, 1024, 0, st>>>(w.r, c->ldN, A, c->ldN, z, c->y, N, M, out3);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

// out3 = {0.5*sum|x|^2 + sum log diag V, 0.5*sum|x|^2, sum log diag V}   (naive variant, image_cupy.py:657)
__global__ void __launch_bounds__(1024) phi_naive_reduce_kernel(const cplx* __restrict__ x, const cplx* __restrict__ V, int ldv,
                                                                int M, double* __restrict__ out3) {
    __shared__ double s1[1024], s2[1024];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < M; i += 1024) {
        const cplx xi = x[i];
        a += xi.x * xi.x + xi.y * xi.y;
        b += log(V[size_t(i) * ldv + i].x);
    }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out3[0] = 0.5 * s1[0] + s2[0]; out3[1] = 0.5 * s1[0]; out3[2] = s2[0]; }
}

// "Naive" Phi of the reference (use_naive_logRatio_implementation, image_cupy.py:644-657), no stabiliser:
//   Z = solve(L^H, H^H)  (pivoted LU of L^H with the M columns of H^H carried along, then U^-1 on all of them)
//   C = chol(Z^H Z + I) ; Phi = 0.5 * || C^-1 y ||^2 + sum(log(Re diag C))
int phi_eval_naive(Ctx* c, cudaStream_t st, PhiWork& w, const cplx* L, int ldl, double* out3) {
    const int N = c->N, M = c->M;
    const int ldz = (N + M + 7) / 8 * 8;
    if (!c->naive_aug) {   // [L^H | H^H] workspace and its LU scratch, allocated on first use (freed with the context)
        MSPDE_CUDA(cudaMalloc((void**)&c->naive_aug, size_t(N) * ldz * sizeof(cplx)));
        MSPDE_CUDA(cudaMalloc((void**)&c->naive_lu_work, lu_work_bytes(N, N + M)));
        MSPDE_CUDA(cudaMalloc((void**)&c->naive_ut, size_t(N) * c->ldN * sizeof(cplx)));
        MSPDE_CUDA(cudaMalloc((void**)&c->naive_xy, size_t(M) * sizeof(cplx)));
    }
    cplx* A = c->naive_aug;
    int rc = transpose(st, L, ldl, N, N, true, A, ldz);                                   // L^H
    if (rc) return rc;
    MSPDE_CUDA(cudaMemcpy2DAsync(A + N, size_t(ldz) * sizeof(cplx), c->Hct, size_t(c->ldHct) * sizeof(cplx),
                                 size_t(M) * sizeof(cplx), N, cudaMemcpyDeviceToDevice, st));   // H^H
    rc = getrf_aug(st, A, ldz, N, M, c->naive_lu_work, c->info, nullptr);
    if (rc) return rc;
    rc = trsm_upper(st, A, ldz, N, A + N, ldz, M, c->naive_ut, c->ldN, w.winv);            // Z (N x M) in the augmented block
    if (rc) return rc;
    rc = zgemm_tn(st, true, M, M, N, 1.0, A + N, ldz, A + N, ldz, 0.0, w.Q, c->ldM, GEMM_UPPER);   // Z^H Z
    if (rc) return rc;
    add_diag_kernel<<<(M + 255) / 256, 256, 0, st>>>(w.Q, c->ldM, M, 1.0);
    MSPDE_LAUNCH_CHECK();
    rc = potrf_upper(st, w.Q, c->ldM, M, w.winv, c->info, nullptr);                       // Q = V^H V, C = V^H
    if (rc) return rc;
    real_to_cplx_kernel<<<(M + 255) / 256, 256, 0, st>>>(c->y, c->naive_xy, M);
    MSPDE_LAUNCH_CHECK();
    rc = trsm_uh(st, w.Q, c->ldM, w.winv, M, c->naive_xy, 1, 1);                          // x = C^-1 y = V^-H y
    if (rc) return rc;
    phi_naive_reduce_kernel<<<1, 1024, 0, st>>>(c->naive_xy, w.Q, c->ldM, M, out3);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

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
    vy_rows_kernel<<<(M + 7) / 8, 256, 0, st>>>(w.Q, c->ldM, c->y, M, M, w.rowsq, w.rowlog);
    MSPDE_LAUNCH_CHECK();
    phi_reduce_kernel<<<1, 1024, 0, st>>>(w.rowsq, w.rowlog, M, out3);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mspde
