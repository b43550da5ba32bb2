// Per-(device, n, n_ext, M, n_layers) context: problem sizes, borrowed fixed inputs, library-owned workspace.
#pragma once
#include "common.cuh"
#include "linalg.cuh"
#include "../../include/mspde.h"

namespace mspde {

struct PhiWork {
    cplx* r = nullptr;        // N x ldN
    cplx* X = nullptr;        // N x ldM   Ht = c^-1 H^H
    cplx* Q = nullptr;        // M x ldM
    cplx* A = nullptr;        // N x ldN   L^H L + delta I (fast Phi formulation only; allocated on first use)
    cplx* winv = nullptr;     // inverses of 64-blocks, sized for max(N, M)
    double* rowsq = nullptr;  // M
    double* rowlog = nullptr; // M
    LookAhead la;             // look-ahead stream/events of this workspace's Cholesky factorizations
};

struct Ctx {
    int device = 0;
    int n = 0, n_ext = 0, il = 0, N = 0, n_ravel = 0, m = 0, M = 0, n_layers = 0;
    int ldN = 0, ldM = 0;            // padded leading dimensions of the internal N- and M-wide workspaces
    cudaStream_t stream = nullptr;

    // fixed inputs, borrowed from the caller (device pointers, row-major)
    const cplx* H = nullptr;   int ldH = 0;     // M x N   normalised tomography operator
    const cplx* Hct = nullptr; int ldHct = 0;   // N x M   H^H
    const cplx* HtH = nullptr; int ldHtH = 0;   // N x N   symmetrised H^H H / sigma^2
    const double* y = nullptr;                  // M
    const double* D_diag = nullptr;             // N
    const double* stdev_sym = nullptr;          // N   layer-0 diagonal prior
    double delta = 0.0;                         // Cholesky stabiliser  eps * ||HtH||_F

    // field synthesis scratch
    cplx* W = nullptr;        // m twiddles e^{2 pi i a / m}
    cplx* T1 = nullptr;       // m x il
    double* fp = nullptr;     // m x m   exp(+u)
    double* fm = nullptr;     // m x m   exp(-u)
    double* img = nullptr;    // m x m   u(x) (posterior statistics only)
    cplx* G = nullptr;        // 2 x m x n
    cplx* tp = nullptr;       // il x il  sym2D(rfft2(exp(+u)))
    cplx* tm = nullptr;       // il x il

    // Phi workspaces: two sets so that Phi(L_current) and Phi(L_latest) can run on concurrent streams
    PhiWork pw[2];
    cplx*& r = pw[0].r;       // aliases of set 0 (debug / parity access)
    cplx*& X = pw[0].X;
    cplx*& Q = pw[0].Q;
    cplx*& winv = pw[0].winv;
    double* scal = nullptr;   // small device scalars (phi triples, logdet, flags)
    int* info = nullptr;      // device LAPACK-style info

    // QR workspace (Gibbs least squares and the dense solves L x = w)
    cplx* QRaug = nullptr;    // (M+N) x ldA,  ldA = round8(N+1)
    int ldA = 0;
    void* qr_work = nullptr;
    QrLookAhead qr_la;        // look-ahead of the Householder QR (second panel workspace, stream, events)
    void* lu_work = nullptr;
    cplx* vecN = nullptr;     // scratch vectors (4 x (N+M))
    cplx* gvec = nullptr;     // H^H y (fast Phi formulation)
    cplx* naive_aug = nullptr;        // naive Phi variant: [L^H | H^H], N x (N+M); allocated on first use
    void* naive_lu_work = nullptr;
    cplx* naive_ut = nullptr;         // U^T
    cplx* naive_xy = nullptr;         // C^-1 y
    bool g_valid = false;

    // side streams for the independent heavy tasks of one step (Phi(L_current), Gibbs QR) and their events
    cudaStream_t s0 = nullptr, s1 = nullptr, s2 = nullptr;   // main chain (highest priority), Phi(current), Gibbs
    cudaEvent_t ev_start = nullptr, ev_L = nullptr, ev_s1 = nullptr, ev_s2 = nullptr, ev_out = nullptr;
};

// pcn.cu
int solve_L(Ctx* c, const cplx* L, int ldl, const cplx* rhs, cplx* x, double* logabsdet);      // pivoted LU
int solve_L_qr(Ctx* c, const cplx* L, int ldl, const cplx* rhs, cplx* x, double* logabsdet);   // Householder QR
int gibbs_lstsq(Ctx* c, const cplx* L, int ldl, const cplx* rhs, cplx* v);
int lstsq_general(Ctx* c, const cplx* A, int lda, int rows, int cols, const cplx* b, cplx* x);
int pcn_step(Ctx* c, const mspde_chain* ch, const mspde_step_noise* nz, int flags, double* out_dev, cplx* record_dev);

// assembly.cu
int field_coeffs(Ctx* c, const cplx* u_sym, cplx* tp_out, cplx* tm_out);
int assemble_L(Ctx* c, const cplx* tp, const cplx* tm, double sqrt_beta, cplx* L, int ldl);
int posterior_stats_update(Ctx* c, const cplx* u_sym, long long count, double* mean_u, double* m2_u, double* mean_ell,
                           double* m2_ell);
// phi.cu
int phi_eval(Ctx* c, cudaStream_t st, PhiWork& w, const cplx* L, int ldl, double* out3, bool serial = false);
int add_diag(cudaStream_t st, cplx* A, int lda, int n, double v);
int upper_rows_times_y(cudaStream_t st, const cplx* V, int ldv, int nrows, int ncols, const double* y, double* rowsq,
                       double* rowlog);
int phi_eval_naive(Ctx* c, cudaStream_t st, PhiWork& w, const cplx* L, int ldl, double* out3);
int phi_fast_prepare(Ctx* c, cudaStream_t st);
int phi_eval_fast(Ctx* c, cudaStream_t st, PhiWork& w, const cplx* L, int ldl, double* out3);   // out3 (device): phi, quad, logdet

}  // namespace mspde
