// Blocked complex FP64 Cholesky (upper form, row-major:  A = U^H U) and the matching triangular solve
// B <- U^-H B, both recursive so that almost all flops land in large-K zgemm_tn (DMMA) calls.
//
// Reference call sites: cp.linalg.cholesky at /root/reference/mcmc/image_cupy.py:637 and :642 (the
// reference's lower factor c equals U^H here) and cp.linalg.solve(c, H^H) at :638.
//
// Leaves are 64 x 64 diagonal blocks handled by one CTA: the block is factorised with the trailing
// matrix held in registers (one __syncthreads per column), then its triangular inverse is formed so
// that the leaf solves become  X = inv(U11)^H B  -- again a zgemm_tn call (in place, one row tile).
#include <cstdlib>
#include "common.cuh"
#include "linalg.cuh"

namespace mspde {

namespace {

constexpr int NB = LEAF;   // 64
constexpr int SP = NB + 1; // padded pitch (complex) in shared memory

// One CTA, 256 threads as a 16 x 16 grid; thread (ty, tx) owns elements (ty+16a, tx+16b), a,b in 0..3.
__global__ void __launch_bounds__(256) potrf_diag_kernel(cplx* __restrict__ A, int lda, int nb, int goff,
                                                         cplx* __restrict__ winv, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    cplx* S = reinterpret_cast<cplx*>(sm_raw);          // U (upper) after factorisation   [NB][SP]
    __shared__ cplx rowbuf[2][NB];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;

    cplx r[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = ty + 16 * a, j = tx + 16 * b;
            cplx v = make_double2(0.0, 0.0);
            if (i < nb && j < nb && j >= i) v = A[size_t(i) * lda + j];
            if (i >= nb && i == j) v = make_double2(1.0, 0.0);   // identity padding for edge blocks
            r[a][b] = v;
        }
    for (int k = 0; k < NB; ++k) {
        const int ka = k >> 4, kr = k & 15;
        if (ty == kr) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
                if (a == ka) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) rowbuf[k & 1][tx + 16 * b] = r[a][b];
                }
        }
        __syncthreads();
        const double akk = rowbuf[k & 1][k].x;
        if (!(akk > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, goff + k + 1);   // not positive definite (or NaN): LAPACK zpotrf info = k+1
        }
        const double inv = rsqrt(akk);
        const double d = akk * inv;
        cplx vi[4], vj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) vi[a] = cscale(rowbuf[k & 1][ty + 16 * a], inv);
#pragma unroll
        for (int b = 0; b < 4; ++b) vj[b] = cscale(rowbuf[k & 1][tx + 16 * b], inv);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = ty + 16 * a;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = tx + 16 * b;
                if (i == k) {
                    if (j == k) r[a][b] = make_double2(d, 0.0);
                    else if (j > k) r[a][b] = vj[b];
                } else if (i > k && j >= i) {
                    r[a][b] = cfnmac(vi[a], vj[b], r[a][b]);
                }
            }
        }
    }
    // publish U to shared memory and to A (upper triangle only)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = ty + 16 * a, j = tx + 16 * b;
            S[i * SP + j] = r[a][b];
            if (i < nb && j < nb && j >= i) A[size_t(i) * lda + j] = r[a][b];
        }
    __syncthreads();
    // In-place triangular inverse (ztrti2, columnwise): X[0:j, j] = -X[j][j] * (X[0:j,0:j] * U[0:j, j]), X[j][j] = 1/U[j][j].
    // Row i of column j is a dot product over k = i..j-1, split over 4 consecutive lanes.
    {
        const int row = tid >> 2, part = tid & 3;
        for (int j = 0; j < NB; ++j) {
            cplx s = make_double2(0.0, 0.0);
            if (row < j) {
                for (int k = row + part; k < j; k += 4) s = cfma(S[row * SP + k], S[k * SP + j], s);
            }
            s.x += __shfl_xor_sync(0xffffffffu, s.x, 1);
            s.y += __shfl_xor_sync(0xffffffffu, s.y, 1);
            s.x += __shfl_xor_sync(0xffffffffu, s.x, 2);
            s.y += __shfl_xor_sync(0xffffffffu, s.y, 2);
            const double xjj = 1.0 / S[j * SP + j].x;     // diagonal of a Cholesky factor is real
            __syncthreads();                              // every dot product has read column j before it is overwritten
            if (part == 0) {
                if (row < j) S[row * SP + j] = make_double2(-xjj * s.x, -xjj * s.y);
                else if (row == j) S[row * SP + j] = make_double2(xjj, 0.0);
            }
            __syncthreads();
        }
    }
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int i = idx >> 6, j = idx & 63;
        winv[idx] = (j >= i) ? S[i * SP + j] : make_double2(0.0, 0.0);
    }
}

int split(int n) {
    int n1 = ((n / 2 + NB - 1) / NB) * NB;
    if (n1 >= n) n1 -= NB;
    if (n1 < NB) n1 = NB;
    return n1;
}

DeviceOnce g_attr;

}  // namespace

// B (n x ncols, row-major, ldb) <- U^-H B, U = upper n x n at `U` (ldu), winv = inverses of its 64-blocks.
int trsm_uh(cudaStream_t st, const cplx* U, int ldu, const cplx* winv, int n, cplx* B, int ldb, int ncols) {
    if (n <= 0 || ncols <= 0) return 0;
    if (n <= NB) {
        // leaf: X = inv(U11)^H B, in place (single row tile -> each CTA reads its whole B tile before writing)
        return zgemm_tn(st, true, n, ncols, n, 1.0, winv, NB, B, ldb, 0.0, B, ldb, GEMM_FULL);
    }
    const int n1 = split(n);
    int rc = trsm_uh(st, U, ldu, winv, n1, B, ldb, ncols);
    if (rc) return rc;
    // B2 -= U12^H X1
    rc = zgemm_tn(st, true, n - n1, ncols, n1, -1.0, U + n1, ldu, B, ldb, 1.0, B + size_t(n1) * ldb, ldb, GEMM_FULL);
    if (rc) return rc;
    return trsm_uh(st, U + size_t(n1) * ldu + n1, ldu, winv + size_t(n1 / NB) * NB * NB, n - n1, B + size_t(n1) * ldb,
                   ldb, ncols);
}

// In-place upper Cholesky of the Hermitian matrix whose upper triangle is stored in A (row-major).
// winv must hold ceil(n/64) * 64*64 complex.  *info (device) receives the first failing column (1-based), 0 if ok.
//
// Right-looking over 512-wide panels (well-filled K = 512 trailing updates); each panel's diagonal block and panel
// solve are recursive so that their flops are large-K GEMMs as well.  With a look-ahead context the trailing update
// of panel p is split: the block row of panel p+1 is updated on the main stream, the rest on an auxiliary
// (lower-priority) stream, so that the latency-bound factorisation of panel p+1 runs underneath the big HERK.
int potrf_upper(cudaStream_t st, cplx* A, int lda, int n, cplx* winv, int* info, const LookAhead* la) {
    if (g_attr.first()) {
        MSPDE_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(NB * SP * sizeof(cplx))));
    }
    static const int NBO = getenv("MSPDE_POTRF_NBO") ? atoi(getenv("MSPDE_POTRF_NBO")) : 512;   // outer panel width (multiple of 64)
    const bool ahead = la && la->aux && n > 2 * NBO;
    bool aux_pending = false;
    for (int j0 = 0; j0 < n; j0 += NBO) {
        const int jb = n - j0 < NBO ? n - j0 : NBO;
        cplx* A11 = A + size_t(j0) * lda + j0;
        cplx* wv = winv + size_t(j0 / NB) * NB * NB;
        const int rest = n - j0 - jb;
        // Panel phase, left-looking over the 64-row leaves of the panel: per leaf one wide update of its block row
        // (all columns to the right, panel and trailing part alike), the diagonal-block kernel, and one wide row solve.
        // 3 launches per leaf, every GEMM spans >= rest columns -- this replaces potrf_rec + trsm_uh on the panel.
        int rc = 0;
        for (int i0 = 0; i0 < jb; i0 += NB) {
            const int ib = jb - i0 < NB ? jb - i0 : NB;
            cplx* Aii = A11 + size_t(i0) * lda + i0;
            const int right = (jb - i0) + rest;           // columns from the leaf's diagonal to the end of the matrix
            if (i0 > 0) {
                // A[i, i:] -= U[0:i0, i]^H U[0:i0, i:]   (upper part of the diagonal block only matters; full rows are fine)
                rc = zgemm_tn(st, true, ib, right, i0, -1.0, A11 + i0, lda, A11 + i0, lda, 1.0, Aii, lda, GEMM_UPPER);
                if (rc) return rc;
            }
            potrf_diag_kernel<<<1, 256, NB * SP * sizeof(cplx), st>>>(Aii, lda, ib, j0 + i0, wv + size_t(i0 / NB) * NB * NB, info);
            MSPDE_LAUNCH_CHECK();
            if (right > ib) {
                rc = zgemm_tn(st, true, ib, right - ib, ib, 1.0, wv + size_t(i0 / NB) * NB * NB, NB, Aii + ib, lda, 0.0,
                              Aii + ib, lda, GEMM_FULL);   // row solve, in place (single row tile)
                if (rc) return rc;
            }
        }
        if (rest <= 0) break;
        cplx* A12 = A11 + jb;
        cplx* A22 = A11 + size_t(jb) * lda + jb;
        if (!ahead) {
            rc = zgemm_tn(st, true, rest, rest, jb, -1.0, A12, lda, A12, lda, 1.0, A22, lda, GEMM_UPPER);
            if (rc) return rc;
            continue;
        }
        const int nxt = rest < NBO ? rest : NBO;          // rows of the next panel
        if (aux_pending) MSPDE_CUDA(cudaStreamWaitEvent(st, la->ev_b, 0));   // its block row got the previous (b) update
        rc = zgemm_tn(st, true, nxt, rest, jb, -1.0, A12, lda, A12, lda, 1.0, A22, lda, GEMM_UPPER);          // (a)
        if (rc) return rc;
        if (rest > nxt) {
            MSPDE_CUDA(cudaEventRecord(la->ev_a, st));
            MSPDE_CUDA(cudaStreamWaitEvent(la->aux, la->ev_a, 0));
            rc = zgemm_tn(la->aux, true, rest - nxt, rest - nxt, jb, -1.0, A12 + nxt, lda, A12 + nxt, lda, 1.0,
                          A22 + size_t(nxt) * lda + nxt, lda, GEMM_UPPER);                                    // (b)
            if (rc) return rc;
            MSPDE_CUDA(cudaEventRecord(la->ev_b, la->aux));
            aux_pending = true;
        } else {
            aux_pending = false;
        }
    }
    if (aux_pending) MSPDE_CUDA(cudaStreamWaitEvent(st, la->ev_b, 0));
    return 0;
}

}  // namespace mspde
