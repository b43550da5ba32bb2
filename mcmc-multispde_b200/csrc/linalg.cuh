// Dense linear algebra building blocks of libmspde (all row-major complex128, all on a caller stream).
#pragma once
#include "common.cuh"

namespace mspde {

constexpr int LEAF = 64;   // diagonal-block size of the recursive factorizations (== zgemm_tn BM)

inline int winv_blocks(int n) { return (n + LEAF - 1) / LEAF; }
inline size_t winv_elems(int n) { return size_t(winv_blocks(n)) * LEAF * LEAF; }

// potrf.cu
int potrf_upper(cudaStream_t st, cplx* A, int lda, int n, cplx* winv, int* info);
int trsm_uh(cudaStream_t st, const cplx* U, int ldu, const cplx* winv, int n, cplx* B, int ldb, int ncols);

// getrf.cu
struct LuWork {
    cplx* T;        // transposed-panel scratch, >= n * n/2 complex
    cplx* linv;     // transposed inverses of the unit-lower 64-blocks, winv_elems(n)
    int* piv;       // n pivots (device)
    double* cand;   // panel-kernel scratch
    int* sync;      // grid barrier counters
};
size_t lu_work_bytes(int n);
// In-place LU with partial pivoting (row-major, P A = L U) of the n x n matrix A augmented with nrhs extra
// columns (A is n x (n+nrhs)); the extra columns receive L^-1 P b.  logabsdet (device) <- sum log|u_ii|.
int getrf_aug(cudaStream_t st, cplx* A, int lda, int n, int nrhs, void* work, int* info, double* logabsdet);
// x <- U^-1 y for the upper triangle of A (n x n), y = column n of the augmented matrix.
int trsv_upper_aug(cudaStream_t st, const cplx* A, int lda, int n, int col, cplx* x);

// geqrf.cu
size_t qr_work_bytes(int rows, int cols);
// Least squares min || A x - b ||, A rows x cols (rows >= cols) stored augmented: Aaug = [A | b], rows x (cols+1).
// Overwrites Aaug; x (cols) <- solution.
int lstsq_qr_aug(cudaStream_t st, cplx* Aaug, int lda, int rows, int cols, void* work, cplx* x);

}  // namespace mspde
