// Dense linear algebra building blocks of libmspde (all row-major complex128, all on a caller stream).
#pragma once
#include "common.cuh"

namespace mspde {

constexpr int LEAF = 64;   // diagonal-block size of the recursive factorizations (== zgemm_tn BM)

inline int winv_blocks(int n) { return (n + LEAF - 1) / LEAF; }
inline size_t winv_elems(int n) { return size_t(winv_blocks(n)) * LEAF * LEAF; }

// potrf.cu
struct LookAhead {          // auxiliary stream + events for the panel look-ahead of potrf_upper
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
};
int potrf_upper(cudaStream_t st, cplx* A, int lda, int n, cplx* winv, int* info, const LookAhead* la = nullptr);
int trsm_uh(cudaStream_t st, const cplx* U, int ldu, const cplx* winv, int n, cplx* B, int ldb, int ncols);

// getrf.cu
size_t lu_work_bytes(int n, int ntot);
// In-place P A = L U (partial pivoting, LAPACK pivot rule) of the leading n x n block of A (n x (n+naug), row-major); the
// naug trailing columns receive L^-1 P b.  *info (device) <- first zero pivot (1-based); logabs (device) <- sum log|u_jj|.
int getrf_aug(cudaStream_t st, cplx* A, int lda, int n, int naug, void* work, int* info, double* logabs);
int lu_panel_factor(cudaStream_t st, cplx* Ap, int lda, int mp, int pw, int goff, int n, int ntot, void* work, int* info,
                    double* logabs);
int lu_panel_apply(cudaStream_t st, int mp, int pw, int n, int ntot, void* work, cplx* A12, int lda, int n2);
void lu_work_layout(int n, int ntot, void* work, long long* out6);

// trsm_upper.cu
int transpose(cudaStream_t st, const cplx* in, int ldi, int rows, int cols, bool conj, cplx* out, int ldo);
int trsm_upper(cudaStream_t st, const cplx* U, int ldu, int n, cplx* B, int ldb, int ncols, cplx* Ut, int ldut, cplx* invT);

// geqrf.cu
size_t qr_work_bytes(int rows, int cols);
// Householder QR of the leading `cols` columns of A (rows x (cols+naug), rows >= cols); the naug trailing columns
// receive Q^H applied.  logabs (device, optional) <- sum_j log|r_jj|.
struct QrLookAhead {        // auxiliary stream, events and second panel workspace for the one-panel look-ahead of geqrf_aug
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    void* work2 = nullptr;  // qr_work_bytes(rows, cols) like `work`
    // K = 128 path (two 64-column panels per trailing update): two pair workspaces of qr_pair_work_bytes(pair_rows, pair_cols,
    // pair_naug) bytes each; geqrf_aug uses them for any problem that fits those sizes
    void* pair[2] = {nullptr, nullptr};
    int pair_rows = 0, pair_cols = 0, pair_naug = 0;
};
size_t qr_pair_work_bytes(int rows, int cols, int naug);
int qr_pair_prepare_at(cudaStream_t st, void* work, int wrows, int wcols, int wnaug, cplx* Ap, int lda, int mp, int span,
                       double* logabs, int* kk_host);
int qr_pair_apply_at(cudaStream_t st, void* work, int wrows, int wcols, int wnaug, int kk, int mp, cplx* A2, int lda, int n2);
void qr_pair_work_layout(int wrows, int wcols, int wnaug, void* work, long long* out10);
int qr_block_width();
int set_qr_block_width(int k);
int geqrf_aug(cudaStream_t st, cplx* A, int lda, int rows, int cols, int naug, void* work, double* logabs,
              const QrLookAhead* la = nullptr);
int qr_panel_factor(cudaStream_t st, cplx* Ap, int lda, int mp, int pw, int rows, int cols, void* work, double* logabs);
int qr_panel_apply(cudaStream_t st, int mp, int rows, int cols, void* work, cplx* A2, int lda, int n2);
void qr_work_layout(int rows, int cols, void* work, long long* out6);
// x <- R^-1 d with R = upper triangle of A[0:n,0:n] and d = column dcol of A (d is overwritten).
int trsv_upper_aug(cudaStream_t st, cplx* A, int lda, int n, int dcol, cplx* x);
// Least squares min || A x - b ||: Aaug = [A | b] is rows x (cols+1) and is overwritten; x (cols) <- solution.
int lstsq_qr_aug(cudaStream_t st, cplx* Aaug, int lda, int rows, int cols, void* work, cplx* x, double* logabs,
                 const QrLookAhead* la = nullptr);

}  // namespace mspde
