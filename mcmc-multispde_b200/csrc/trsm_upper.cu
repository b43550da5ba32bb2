// Multi-right-hand-side solves with general (LU) factors, used by the "naive" Phi variant
// (/root/reference/mcmc/image_cupy.py:644-657: Z = cp.linalg.solve(L^H, H^H), i.e. zgetrf + zgetrs with M right-hand sides):
//   * transpose_kernel        out = op(in)^T (tiled through shared memory, optional conjugation)
//   * triu_inv_kernel         (U_ii^-1)^T of a general complex upper-triangular 64 x 64 diagonal block
//   * trsm_upper              B <- U^-1 B, recursive: all flops in zgemm_tn (the A operand of "U12 X2" is read from U^T)
#include "linalg.cuh"

namespace mspde {

namespace {

constexpr int NB = LEAF;
constexpr int SP = NB + 1;

__global__ void __launch_bounds__(256) transpose_kernel(const cplx* __restrict__ in, int ldi, int rows, int cols, int conj,
                                                        cplx* __restrict__ out, int ldo) {
    __shared__ cplx tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = by + r, j = bx + tx;
        if (i < rows && j < cols) tile[r][tx] = in[size_t(i) * ldi + j];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bx + r, i = by + tx;                    // out[j][i] = in[i][j]
        if (i < rows && j < cols) {
            cplx v = tile[tx][r];
            if (conj) v.y = -v.y;
            out[size_t(j) * ldo + i] = v;
        }
    }
}

// invT[k][i] = (U^-1)[i][k] for the nb x nb upper-triangular block U (general complex diagonal), identity padding to 64.
__global__ void __launch_bounds__(256) triu_inv_kernel(const cplx* __restrict__ U, int ldu, int nb, cplx* __restrict__ invT) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    cplx* S = reinterpret_cast<cplx*>(sm_raw);   // [NB][SP]
    const int tid = threadIdx.x;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        cplx v = make_double2(0.0, 0.0);
        if (i < nb && j < nb && j >= i) v = U[size_t(i) * ldu + j];
        if (i >= nb && i == j) v = make_double2(1.0, 0.0);
        S[i * SP + j] = v;
    }
    __syncthreads();
    // in-place columnwise inversion (ztrti2): X[0:j, j] = -X[j][j] * (X[0:j,0:j] * U[0:j, j]), X[j][j] = 1 / U[j][j]
    const int row = tid >> 2, part = tid & 3;
    for (int j = 0; j < NB; ++j) {
        cplx s = make_double2(0.0, 0.0);
        if (row < j)
            for (int k = row + part; k < j; k += 4) s = cfma(S[row * SP + k], S[k * SP + j], s);
        s.x += __shfl_xor_sync(0xffffffffu, s.x, 1);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, 1);
        s.x += __shfl_xor_sync(0xffffffffu, s.x, 2);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, 2);
        const cplx xjj = cdiv(make_double2(1.0, 0.0), S[j * SP + j]);
        __syncthreads();
        if (part == 0) {
            if (row < j) S[row * SP + j] = cmul(make_double2(-xjj.x, -xjj.y), s);
            else if (row == j) S[row * SP + j] = xjj;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int k = idx >> 6, i = idx & 63;
        invT[idx] = (k >= i) ? S[i * SP + k] : make_double2(0.0, 0.0);
    }
}

int split(int n) {
    int n1 = ((n / 2 + NB - 1) / NB) * NB;
    if (n1 >= n) n1 -= NB;
    if (n1 < NB) n1 = NB;
    return n1;
}

DeviceOnce g_attr;

int trsm_rec(cudaStream_t st, const cplx* Ut, int ldut, const cplx* invT, int n, cplx* B, int ldb, int ncols) {
    if (n <= NB)   // X = U_ii^-1 B in place (single row tile)
        return zgemm_tn(st, false, n, ncols, n, 1.0, invT, NB, B, ldb, 0.0, B, ldb, GEMM_FULL);
    const int n1 = split(n);
    int rc = trsm_rec(st, Ut + size_t(n1) * ldut + n1, ldut, invT + size_t(n1 / NB) * NB * NB, n - n1, B + size_t(n1) * ldb, ldb,
                      ncols);                                                              // X2 = U22^-1 B2
    if (rc) return rc;
    rc = zgemm_tn(st, false, n1, ncols, n - n1, -1.0, Ut + size_t(n1) * ldut, ldut, B + size_t(n1) * ldb, ldb, 1.0, B, ldb,
                  GEMM_FULL);                                                              // B1 -= U12 X2  (A = U12^T rows of U^T)
    if (rc) return rc;
    return trsm_rec(st, Ut, ldut, invT, n1, B, ldb, ncols);                                // X1 = U11^-1 B1
}

}  // namespace

int transpose(cudaStream_t st, const cplx* in, int ldi, int rows, int cols, bool conj, cplx* out, int ldo) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_kernel<<<grid, 256, 0, st>>>(in, ldi, rows, cols, conj ? 1 : 0, out, ldo);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

// B (n x ncols) <- U^-1 B for the upper triangle U of a row-major matrix.  Ut (n x n, ldut) and invT (ceil(n/64) blocks of
// 64 x 64) are workspace.
int trsm_upper(cudaStream_t st, const cplx* U, int ldu, int n, cplx* B, int ldb, int ncols, cplx* Ut, int ldut, cplx* invT) {
    if (g_attr.first()) {
        MSPDE_CUDA(cudaFuncSetAttribute(triu_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NB * SP * sizeof(cplx))));
    }
    int rc = transpose(st, U, ldu, n, n, false, Ut, ldut);
    if (rc) return rc;
    for (int b0 = 0; b0 < n; b0 += NB) {
        triu_inv_kernel<<<1, 256, NB * SP * sizeof(cplx), st>>>(U + size_t(b0) * ldu + b0, ldu, min(NB, n - b0),
                                                               invT + size_t(b0 / NB) * NB * NB);
        MSPDE_LAUNCH_CHECK();
    }
    return trsm_rec(st, Ut, ldut, invT, n, B, ldb, ncols);
}

}  // namespace mspde
