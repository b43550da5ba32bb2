// extern "C" boundary of libmspde.so -- declarations and reference citations live in include/mspde.h.
#include "context.cuh"
#include <cstdarg>
#include <cmath>
#include <vector>
#include "../../include/mspde.h"

namespace mspde {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }
std::atomic<long long> g_launches{0};

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

template <typename T>
static int dmalloc(T** p, size_t count) {
    MSPDE_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    return 0;
}
}  // namespace mspde

using namespace mspde;

struct mspde_ctx {
    Ctx c;
    std::vector<void*> owned;
};

#define CTX_ALLOC(field, count)                                  \
    do {                                                         \
        if (dmalloc(&h->c.field, (count))) { mspde_ctx_destroy(h); return -1; } \
        h->owned.push_back((void*)h->c.field);                   \
    } while (0)

static int g_qr_aux_prio = 0;

extern "C" {

const char* mspde_last_error(void) { return get_error(); }
long long mspde_launch_count(void) { return g_launches.load(); }
int mspde_set_gemm_mode(int mode) { return set_gemm_mode(mode); }
int mspde_get_gemm_mode(void) { return gemm_mode(); }
int mspde_set_qr_block(int k) { return set_qr_block_width(k); }
int mspde_get_qr_block(void) { return qr_block_width(); }
int mspde_set_gemm_pre_min(int v) { return set_gemm_pre_min(v); }
int mspde_get_gemm_pre_min(void) { return gemm_pre_min(); }
long long mspde_release_gemm_workspaces(void) { return (long long)gemm_release_all(); }

int mspde_ctx_create(mspde_ctx** out, int device, int n, int n_ext, int M, int n_layers, void* stream) {
    return mspde_ctx_create_ex(out, device, n, n_ext, M, n_layers, stream, 0);
}

int mspde_ctx_create_ex(mspde_ctx** out, int device, int n, int n_ext, int M, int n_layers, void* stream, int flags) {
    if (!out || n < 2 || n_ext < n || M < 1 || n_layers < 2) {
        set_error("mspde_ctx_create: bad arguments (n=%d n_ext=%d M=%d n_layers=%d)", n, n_ext, M, n_layers);
        return -1;
    }
    MSPDE_CUDA(cudaSetDevice(device));
    mspde_ctx* h = new mspde_ctx();
    Ctx& c = h->c;
    c.device = device;
    c.n = n; c.n_ext = n_ext; c.il = 2 * n - 1; c.N = c.il * c.il; c.n_ravel = 2 * n * n - 2 * n + 1;
    c.m = 2 * n_ext - 1; c.M = M; c.n_layers = n_layers;
    c.ldN = round_up(c.N, 8); c.ldM = round_up(M, 8);
    c.stream = (cudaStream_t)stream;
    const int N = c.N, m = c.m, il = c.il;
    CTX_ALLOC(W, m);
    CTX_ALLOC(T1, size_t(m) * il);
    CTX_ALLOC(fp, size_t(m) * m);
    CTX_ALLOC(fm, size_t(m) * m);
    CTX_ALLOC(img, size_t(m) * m);
    CTX_ALLOC(G, size_t(2) * m * n);
    CTX_ALLOC(tp, size_t(il) * il);
    CTX_ALLOC(tm, size_t(il) * il);
    CTX_ALLOC(scal, 64);
    CTX_ALLOC(info, 4);
    MSPDE_CUDA(cudaMemset(c.info, 0, 4 * sizeof(int)));
    MSPDE_CUDA(cudaMemset(c.scal, 0, 64 * sizeof(double)));
    {
        std::vector<cplx> w(m);
        for (int a = 0; a < m; ++a) {
            const double ang = 2.0 * M_PI * double(a) / double(m);
            w[a] = make_double2(cos(ang), sin(ang));
        }
        MSPDE_CUDA(cudaMemcpy(c.W, w.data(), m * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    if (flags & MSPDE_CTX_ASSEMBLY_ONLY) {   // field synthesis + assembly only (block-distributed path keeps its own buffers)
        *out = h;
        return 0;
    }
    for (int s = 0; s < 2; ++s) {
        CTX_ALLOC(pw[s].r, size_t(N) * c.ldN);
        CTX_ALLOC(pw[s].X, size_t(N) * c.ldM);
        CTX_ALLOC(pw[s].Q, size_t(M) * c.ldM);
        CTX_ALLOC(pw[s].winv, winv_elems(N > M ? N : M));
        CTX_ALLOC(pw[s].rowsq, M);
        CTX_ALLOC(pw[s].rowlog, M);
    }
    {
        int lo = 0, hi = 0;   // numerically lower = higher priority
        MSPDE_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const int mid = (hi + 1 <= lo) ? hi + 1 : lo;
        int order[6] = {0, 1, 2, 3, 4, 0};   // s0, s2, aux(Phi latest), s1, aux(Phi current), aux(QR); 0 = highest
        if (const char* e = getenv("MSPDE_PRIO")) sscanf(e, "%d,%d,%d,%d,%d,%d", &order[0], &order[1], &order[2], &order[3], &order[4], &order[5]);
        auto lvl = [&](int k) { int p = hi + order[k]; return p > lo ? lo : p; };
        g_qr_aux_prio = lvl(5);
        MSPDE_CUDA(cudaStreamCreateWithPriority(&c.s0, cudaStreamNonBlocking, lvl(0)));    // solve -> Phi(latest): critical path
        MSPDE_CUDA(cudaStreamCreateWithPriority(&c.s2, cudaStreamNonBlocking, lvl(1)));    // Gibbs QR
        MSPDE_CUDA(cudaStreamCreateWithPriority(&c.pw[0].la.aux, cudaStreamNonBlocking, lvl(2)));   // trailing HERKs of Phi(latest)
        MSPDE_CUDA(cudaStreamCreateWithPriority(&c.s1, cudaStreamNonBlocking, lvl(3)));    // Phi(current): filler
        MSPDE_CUDA(cudaStreamCreateWithPriority(&c.pw[1].la.aux, cudaStreamNonBlocking, lvl(4)));
        for (int s = 0; s < 2; ++s) {
            MSPDE_CUDA(cudaEventCreateWithFlags(&c.pw[s].la.ev_a, cudaEventDisableTiming));
            MSPDE_CUDA(cudaEventCreateWithFlags(&c.pw[s].la.ev_b, cudaEventDisableTiming));
        }
        (void)mid;
    }
    MSPDE_CUDA(cudaEventCreateWithFlags(&c.ev_out, cudaEventDisableTiming));
    MSPDE_CUDA(cudaEventCreateWithFlags(&c.ev_start, cudaEventDisableTiming));
    MSPDE_CUDA(cudaEventCreateWithFlags(&c.ev_L, cudaEventDisableTiming));
    MSPDE_CUDA(cudaEventCreateWithFlags(&c.ev_s1, cudaEventDisableTiming));
    MSPDE_CUDA(cudaEventCreateWithFlags(&c.ev_s2, cudaEventDisableTiming));
    CTX_ALLOC(vecN, size_t(4) * (N + M));
    CTX_ALLOC(gvec, N);
    c.ldA = round_up(N + 2, 8);
    CTX_ALLOC(QRaug, size_t(M + N) * c.ldA);
    {
        unsigned char* w = nullptr;
        if (dmalloc(&w, qr_work_bytes(M + N, N))) { mspde_ctx_destroy(h); return -1; }
        h->owned.push_back(w);
        c.qr_work = w;
        unsigned char* wq = nullptr;          // second panel workspace + stream + events: look-ahead of the Gibbs QR
        if (dmalloc(&wq, qr_work_bytes(M + N, N))) { mspde_ctx_destroy(h); return -1; }
        h->owned.push_back(wq);
        c.qr_la.work2 = wq;
        for (int q = 0; q < 2; ++q) {         // pair workspaces of the K = 128 block-reflector path of the Householder QR
            unsigned char* wp = nullptr;
            if (dmalloc(&wp, qr_pair_work_bytes(M + N, N, 2))) { mspde_ctx_destroy(h); return -1; }
            h->owned.push_back(wp);
            c.qr_la.pair[q] = wp;
        }
        c.qr_la.pair_rows = M + N; c.qr_la.pair_cols = N; c.qr_la.pair_naug = 2;
        int lo = 0, hi = 0;
        MSPDE_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        MSPDE_CUDA(cudaStreamCreateWithPriority(&c.qr_la.aux, cudaStreamNonBlocking, g_qr_aux_prio));   // panel kernels: latency-critical
        MSPDE_CUDA(cudaEventCreateWithFlags(&c.qr_la.ev_a, cudaEventDisableTiming));
        MSPDE_CUDA(cudaEventCreateWithFlags(&c.qr_la.ev_b, cudaEventDisableTiming));
        unsigned char* w2 = nullptr;
        if (dmalloc(&w2, lu_work_bytes(N, N + 2))) { mspde_ctx_destroy(h); return -1; }
        h->owned.push_back(w2);
        c.lu_work = w2;
    }
    *out = h;
    return 0;
}

int mspde_ctx_destroy(mspde_ctx* h) {
    if (!h) return 0;
    for (cudaStream_t q : {h->c.s0, h->c.s1, h->c.s2, h->c.qr_la.aux, h->c.pw[0].la.aux, h->c.pw[1].la.aux})
        if (q) gemm_release_stream(q);              // packed-operand workspaces these streams own (zgemm_tn.cu)
    if (h->c.s0) cudaStreamDestroy(h->c.s0);
    if (h->c.ev_out) cudaEventDestroy(h->c.ev_out);
    if (h->c.s1) cudaStreamDestroy(h->c.s1);
    if (h->c.s2) cudaStreamDestroy(h->c.s2);
    for (cudaEvent_t e : {h->c.ev_start, h->c.ev_L, h->c.ev_s1, h->c.ev_s2, h->c.qr_la.ev_a, h->c.qr_la.ev_b,
                          h->c.pw[0].la.ev_a, h->c.pw[0].la.ev_b, h->c.pw[1].la.ev_a, h->c.pw[1].la.ev_b})
        if (e) cudaEventDestroy(e);
    for (cudaStream_t q : {h->c.qr_la.aux, h->c.pw[0].la.aux, h->c.pw[1].la.aux})
        if (q) cudaStreamDestroy(q);
    for (int s = 0; s < 2; ++s)
        if (h->c.pw[s].A) cudaFree(h->c.pw[s].A);
    for (void* p : {(void*)h->c.naive_aug, h->c.naive_lu_work, (void*)h->c.naive_ut, (void*)h->c.naive_xy})
        if (p) cudaFree(p);
    for (void* p : h->owned) cudaFree(p);
    delete h;
    return 0;
}

int mspde_ctx_set_stream(mspde_ctx* h, void* stream) {
    h->c.stream = (cudaStream_t)stream;
    return 0;
}

int mspde_ctx_set_inputs(mspde_ctx* h, const void* H, int ldH, const void* Hct, int ldHct, const void* HtH, int ldHtH,
                         const double* y, const double* D_diag, const double* stdev_sym, double delta) {
    Ctx& c = h->c;
    c.H = (const cplx*)H; c.ldH = ldH;
    c.Hct = (const cplx*)Hct; c.ldHct = ldHct;
    c.HtH = (const cplx*)HtH; c.ldHtH = ldHtH;
    c.y = y; c.D_diag = D_diag; c.stdev_sym = stdev_sym; c.delta = delta;
    c.g_valid = false;
    return 0;
}

int mspde_ctx_dims(mspde_ctx* h, int* out8) {
    const Ctx& c = h->c;
    out8[0] = c.n; out8[1] = c.n_ext; out8[2] = c.il; out8[3] = c.N; out8[4] = c.n_ravel; out8[5] = c.m; out8[6] = c.M;
    out8[7] = c.n_layers;
    return 0;
}

void* mspde_ctx_buffer(mspde_ctx* h, const char* name, int* ld) {
    Ctx& c = h->c;
    if (!strcmp(name, "r")) { if (ld) *ld = c.ldN; return c.r; }
    if (!strcmp(name, "X")) { if (ld) *ld = c.ldM; return c.X; }
    if (!strcmp(name, "Q")) { if (ld) *ld = c.ldM; return c.Q; }
    if (!strcmp(name, "Z")) { if (ld) *ld = (c.N + c.M + 7) / 8 * 8; return c.naive_aug ? c.naive_aug + c.N : nullptr; }
    if (!strcmp(name, "tp")) { if (ld) *ld = c.il; return c.tp; }
    if (!strcmp(name, "tm")) { if (ld) *ld = c.il; return c.tm; }
    if (!strcmp(name, "fp")) { if (ld) *ld = c.m; return c.fp; }
    if (!strcmp(name, "scal")) { if (ld) *ld = 1; return c.scal; }
    if (!strcmp(name, "info")) { if (ld) *ld = 1; return c.info; }
    set_error("mspde_ctx_buffer: unknown buffer '%s'", name);
    return nullptr;
}

int mspde_index_maps(int n, int32_t* ix, int32_t* iy, int32_t* kmat, double* D_diag) {
    // FourierAnalysis_2D.__init__ tables (host arrays): ix/iy il x il (row-major), Kmatrix 2 x N, D_diag N ('F' ravel)
    const int il = 2 * n - 1, N = il * il;
    const float two_pi = 2.0f * (float)M_PI;   // (2*util.PI)**2 is float32 arithmetic in the reference
    const float c32 = two_pi * two_pi;
    for (int r = 0; r < il; ++r)
        for (int c = 0; c < il; ++c) {
            if (ix) ix[r * il + c] = c - (n - 1);
            if (iy) iy[r * il + c] = r - (n - 1);
        }
    for (int j = 0; j < N; ++j) {
        const int kx = j / il - (n - 1), ky = j % il - (n - 1);
        if (kmat) { kmat[j] = kx; kmat[N + j] = ky; }
        if (D_diag) D_diag[j] = -(double)c32 * (double)(kx * kx + ky * ky);
    }
    return 0;
}

int mspde_bttb_index_map(int n, int32_t* row_idx, int32_t* col_idx, uint8_t* mask) {
    // host restatement used by the bit-exactness test of the kernel's index arithmetic
    const int il = 2 * n - 1, N = il * il;
    for (int m = 0; m < N; ++m)
        for (int c = 0; c < N; ++c) {
            const int dij = m / il - c / il, dkl = m % il - c % il;
            const size_t o = size_t(m) * N + c;
            row_idx[o] = dkl + (n - 1);
            col_idx[o] = dij + (n - 1);
            mask[o] = (dij >= -(n - 1) && dij < n && dkl >= -(n - 1) && dkl < n) ? 1 : 0;
        }
    return 0;
}

int mspde_field_coeffs(mspde_ctx* h, const void* u_sym, void* coef_plus, void* coef_minus) {
    return field_coeffs(&h->c, (const cplx*)u_sym, (cplx*)coef_plus, (cplx*)coef_minus);
}

int mspde_assemble_L(mspde_ctx* h, const void* coef_plus, const void* coef_minus, double sqrt_beta, void* L, int ldl) {
    return assemble_L(&h->c, (const cplx*)coef_plus, (const cplx*)coef_minus, sqrt_beta, (cplx*)L, ldl);
}

int mspde_construct_L(mspde_ctx* h, const void* u_sym, double sqrt_beta, void* L, int ldl) {
    int rc = field_coeffs(&h->c, (const cplx*)u_sym, h->c.tp, h->c.tm);
    if (rc) return rc;
    return assemble_L(&h->c, h->c.tp, h->c.tm, sqrt_beta, (cplx*)L, ldl);
}

int mspde_posterior_stats_update(mspde_ctx* h, const void* u_sym, long long count, double* mean_u, double* m2_u,
                                 double* mean_ell, double* m2_ell) {
    if (count < 1) {
        set_error("mspde_posterior_stats_update: count must be >= 1 (it is the sample count after this update)");
        return -1;
    }
    return posterior_stats_update(&h->c, (const cplx*)u_sym, count, mean_u, m2_u, mean_ell, m2_ell);
}

int mspde_phi(mspde_ctx* h, const void* L, int ldl, double* out3_dev) {
    return phi_eval(&h->c, h->c.stream, h->c.pw[0], (const cplx*)L, ldl, out3_dev, false);
}

int mspde_solve(mspde_ctx* h, const void* L, int ldl, const void* rhs, void* x, double* logabsdet_dev) {
    return solve_L(&h->c, (const cplx*)L, ldl, (const cplx*)rhs, (cplx*)x, logabsdet_dev ? logabsdet_dev : h->c.scal + 20);
}

int mspde_solve_qr(mspde_ctx* h, const void* L, int ldl, const void* rhs, void* x, double* logabsdet_dev) {
    return solve_L_qr(&h->c, (const cplx*)L, ldl, (const cplx*)rhs, (cplx*)x, logabsdet_dev ? logabsdet_dev : h->c.scal + 20);
}

int mspde_lstsq(mspde_ctx* h, const void* a, int lda, int rows, int cols, const void* b, void* x) {
    return lstsq_general(&h->c, (const cplx*)a, lda, rows, cols, (const cplx*)b, (cplx*)x);
}

int mspde_gibbs_lstsq(mspde_ctx* h, const void* L, int ldl, const void* rhs, void* v) {
    return gibbs_lstsq(&h->c, (const cplx*)L, ldl, (const cplx*)rhs, (cplx*)v);
}

int mspde_pcn_step(mspde_ctx* h, const mspde_chain* chain, const mspde_step_noise* noise, int flags, double* out_dev,
                   void* record_dev) {
    if (!chain || !noise || chain->n_layers != h->c.n_layers) {
        set_error("mspde_pcn_step: chain/noise missing or n_layers mismatch");
        return -1;
    }
    return pcn_step(&h->c, chain, noise, flags, out_dev, (cplx*)record_dev);
}

int mspde_phi_naive(mspde_ctx* h, const void* L, int ldl, double* out3_dev) {
    return phi_eval_naive(&h->c, h->c.stream, h->c.pw[0], (const cplx*)L, ldl, out3_dev);
}

int mspde_phi_fast(mspde_ctx* h, const void* L, int ldl, double* out3_dev) {
    int rc = phi_fast_prepare(&h->c, h->c.stream);
    if (rc) return rc;
    return phi_eval_fast(&h->c, h->c.stream, h->c.pw[0], (const cplx*)L, ldl, out3_dev);
}

int mspde_read_info(mspde_ctx* h, int reset) {
    int v[4] = {0, 0, 0, 0};
    if (cudaMemcpyAsync(v, h->c.info, sizeof(int), cudaMemcpyDeviceToHost, h->c.stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(h->c.stream) != cudaSuccess) {
        set_error("stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    if (reset && v[0]) cudaMemsetAsync(h->c.info, 0, sizeof(int), h->c.stream);
    return v[0];
}

int mspde_zgemm_tn(void* stream, int conjA, int m, int n, int k, double alpha, const void* A, int lda, const void* B,
                   int ldb, double beta, void* C, int ldc, int uplo, void* splitk_ws, int nsplit) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nsplit > 1)
        return zgemm_tn_splitk(st, conjA != 0, m, n, k, alpha, (const cplx*)A, lda, (const cplx*)B, ldb, beta, (cplx*)C,
                               ldc, (cplx*)splitk_ws, nsplit);
    return zgemm_tn(st, conjA != 0, m, n, k, alpha, (const cplx*)A, lda, (const cplx*)B, ldb, beta, (cplx*)C, ldc, uplo);
}

int mspde_tomography_H(void* stream, int M, int N, const double* r_flat, const double* theta_flat, const int32_t* ix,
                       const int32_t* iy, int n_r, int n_theta, int apply_fixes, void* H, int ldh) {
    return tomography_H((cudaStream_t)stream, M, N, r_flat, theta_flat, ix, iy, n_r, n_theta, apply_fixes, (cplx*)H, ldh);
}

/* ---- panel-level entry points for the block-distributed LU / QR (mspde/dist.py) ---- */
long long mspde_qr_work_bytes(int rows, int cols) { return (long long)qr_work_bytes(rows, cols); }
long long mspde_lu_work_bytes(int n, int ntot) { return (long long)lu_work_bytes(n, ntot); }
int mspde_qr_work_layout(int rows, int cols, void* work, long long* out6) { qr_work_layout(rows, cols, work, out6); return 0; }
int mspde_lu_work_layout(int n, int ntot, void* work, long long* out6) { lu_work_layout(n, ntot, work, out6); return 0; }
int mspde_qr_panel(void* stream, void* Ap, int lda, int mp, int pw, int rows, int cols, void* work, double* logabs_dev) {
    return qr_panel_factor((cudaStream_t)stream, (cplx*)Ap, lda, mp, pw, rows, cols, work, logabs_dev);
}
int mspde_qr_apply(void* stream, int mp, int rows, int cols, void* work, void* A2, int lda, int n2) {
    return qr_panel_apply((cudaStream_t)stream, mp, rows, cols, work, (cplx*)A2, lda, n2);
}
long long mspde_qr_pair_work_bytes(int rows, int cols, int naug) { return (long long)qr_pair_work_bytes(rows, cols, naug); }
int mspde_qr_pair_work_layout(int rows, int cols, int naug, void* work, long long* out10) {
    qr_pair_work_layout(rows, cols, naug, work, out10);
    return 0;
}
int mspde_qr_pair_prepare(void* stream, void* work, int wrows, int wcols, int wnaug, void* Ap, int lda, int mp, int span,
                          double* logabs_dev, int* kk_host) {
    return qr_pair_prepare_at((cudaStream_t)stream, work, wrows, wcols, wnaug, (cplx*)Ap, lda, mp, span, logabs_dev, kk_host);
}
int mspde_qr_pair_apply(void* stream, void* work, int wrows, int wcols, int wnaug, int kk, int mp, void* A2, int lda, int n2) {
    return qr_pair_apply_at((cudaStream_t)stream, work, wrows, wcols, wnaug, kk, mp, (cplx*)A2, lda, n2);
}
int mspde_lu_panel(void* stream, void* Ap, int lda, int mp, int pw, int goff, int n, int ntot, void* work, int* info_dev,
                   double* logabs_dev) {
    return lu_panel_factor((cudaStream_t)stream, (cplx*)Ap, lda, mp, pw, goff, n, ntot, work, info_dev, logabs_dev);
}
int mspde_lu_apply(void* stream, int mp, int pw, int n, int ntot, void* work, void* A12, int lda, int n2) {
    return lu_panel_apply((cudaStream_t)stream, mp, pw, n, ntot, work, (cplx*)A12, lda, n2);
}
int mspde_trsv_upper(void* stream, void* A, int lda, int n, int dcol, void* x) {
    return trsv_upper_aug((cudaStream_t)stream, (cplx*)A, lda, n, dcol, (cplx*)x);
}

int mspde_add_diag(void* stream, void* A, int lda, int n, double v) {
    return add_diag((cudaStream_t)stream, (cplx*)A, lda, n, v);
}

int mspde_upper_rows_times_y(void* stream, const void* V, int ldv, int nrows, int ncols, const double* y, double* rowsq,
                             double* rowlog) {
    return upper_rows_times_y((cudaStream_t)stream, (const cplx*)V, ldv, nrows, ncols, y, rowsq, rowlog);
}

int mspde_zpotrf_upper(void* stream, void* A, int lda, int n, void* winv, int* info_dev) {
    return potrf_upper((cudaStream_t)stream, (cplx*)A, lda, n, (cplx*)winv, info_dev);
}

int mspde_ztrsm_uh(void* stream, const void* U, int ldu, const void* winv, int n, void* B, int ldb, int ncols) {
    return trsm_uh((cudaStream_t)stream, (const cplx*)U, ldu, (const cplx*)winv, n, (cplx*)B, ldb, ncols);
}

}  // extern "C"
