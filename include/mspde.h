/* libmspde.so -- C ABI of the B200-native pCN hot path of MCMC-MultiSPDE.
 *
 * The reference has no FFI layer: the path sits behind Python methods on CuPy arrays
 * (SURVEY.md section 8b).  Each entry point below names the reference call site it replaces
 * (paths relative to /root/reference).  Conventions:
 *   - every matrix is row-major (C order) interleaved complex128 with an explicit leading dimension
 *     (in complex elements); vectors are contiguous; all data pointers are DEVICE pointers unless the
 *     name says host;
 *   - every function returns int: 0 ok, >0 LAPACK-style info (host raises numpy.linalg.LinAlgError,
 *     the only error Simulation.run catches, mcmc/image_cupy.py:969), <0 CUDA failure with the text
 *     in mspde_last_error();
 *   - work is enqueued on the context's stream; nothing synchronises unless stated.
 */
#ifndef MSPDE_H
#define MSPDE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mspde_ctx mspde_ctx;

const char* mspde_last_error(void);
/* FP64 tensor-pipe (DMMA m8n8k4) peak of the current device in TFLOP/s, measured now (~0.3 s; synchronises).  bench.py uses it
 * as the roofline denominator of the box it runs on (MEASURED_PEAKS.json carries no FP64 figure). */
int mspde_fp64_peak_probe(void* stream, double* tflops_host);
long long mspde_launch_count(void);   /* kernels launched by the library so far (bench.py gpu_launches) */

/* Complex product scheme of every dense contraction of the path (the cuBLAS ZGEMM calls behind `@`, cp.linalg.solve,
 * cp.linalg.cholesky and util.lstsq: mcmc/image_cupy.py:561, 634-643; mcmc/util_cupy.py:590-609).
 *   MSPDE_GEMM_4M: four real products per complex product, 8 flop per multiply-add (what ZGEMM does);
 *   MSPDE_GEMM_3M: three real products (T1 = ArBr, T2 = AiBi, T3 = (Ar+Ai)(Br+Bi)), 6 flop -- the ZGEMM3M scheme; normwise
 *                  the same O(u) error bound, 1.2x the throughput on the FP64 tensor pipe.  Default.
 * Process-wide; also selectable with the environment variable MSPDE_GEMM=3m|4m before the first launch. */
enum { MSPDE_GEMM_4M = 0, MSPDE_GEMM_3M = 1 };
int mspde_set_gemm_mode(int mode);
int mspde_get_gemm_mode(void);
/* Width of the block reflectors of the Householder QR behind util.lstsq (mcmc/util_cupy.py:590-609): 128 (default) applies two
 * consecutive 64-column panels in one rank-128 trailing update, 64 applies every panel on its own (the round-1 path, kept for
 * comparison; also what the block-distributed driver uses).  Same panel kernel, results agree to rounding.  Process-wide;
 * environment MSPDE_QR_BLOCK=64|128. */
int mspde_set_qr_block(int width);
int mspde_get_qr_block(void);
/* Under MSPDE_GEMM_3M, products with k >= 512 and min(m, n) >= this threshold (default 2048; environment
 * MSPDE_GEMM_PRE_MIN; 0 = never) first repack both operands into the shared-memory image of the kernel's stage tiles, with the
 * two T3 operand planes (Ar +- Ai, Br + Bi) formed once per matrix, and run the kernel variant that consumes those tiles
 * (zgemm3m_pre_tn_kernel) instead of forming the sums per CTA tile inside the main loop.  Same arithmetic, bit-identical
 * results; +6 % on the dominant launch (profiles/r02_gemm_pre_v3.log). */
int mspde_set_gemm_pre_min(int min_mn);
int mspde_get_gemm_pre_min(void);
/* The repacked operands live in one workspace per stream, grown on demand and kept for the stream's next product; a context
 * returns its own streams' workspaces in mspde_ctx_destroy.  This call waits for the device and frees all of them (those of
 * caller-provided streams included); returns the bytes given back.  Use it before a phase that needs the whole HBM. */
long long mspde_release_gemm_workspaces(void);

/* Context = sizes + borrowed fixed inputs + library-owned workspace (replaces the managed-memory pool and the
 * temporaries of pCN._log_ratio, mcmc/image_cupy.py:834-835, 634-643). */
int mspde_ctx_create(mspde_ctx** out, int device, int n, int n_ext, int M, int n_layers, void* stream);
enum { MSPDE_CTX_ASSEMBLY_ONLY = 1 };   /* no Phi/QR workspaces: only field synthesis + assembly (distributed path) */
int mspde_ctx_create_ex(mspde_ctx** out, int device, int n, int n_ext, int M, int n_layers, void* stream, int flags);
int mspde_ctx_destroy(mspde_ctx* ctx);
int mspde_ctx_set_stream(mspde_ctx* ctx, void* stream);
/* H (M x N), H^H (N x M), HtH (N x N), y (M), D_diag (N), stdev_sym (N), delta = chol_epsilon*||HtH||_F
 * (pCN.__init__ 436-452, set_chol_epsilon 473-478, Simulation.__init__ 869-880). */
int mspde_ctx_set_inputs(mspde_ctx* ctx, const void* H, int ldH, const void* Hct, int ldHct, const void* HtH, int ldHtH,
                         const double* y, const double* D_diag, const double* stdev_sym, double delta);
int mspde_ctx_dims(mspde_ctx* ctx, int* out8_host);
void* mspde_ctx_buffer(mspde_ctx* ctx, const char* name, int* ld_host);   /* debug/parity access to workspaces */
int mspde_read_info(mspde_ctx* ctx, int reset);                           /* synchronises; returns device info flag */

/* a1: FourierAnalysis_2D.__init__ index tables (mcmc/image_cupy.py:52-57, 63, 73-74); HOST arrays. */
int mspde_index_maps(int n, int32_t* ix_host, int32_t* iy_host, int32_t* kmat_host, double* D_diag_host);
/* a3: closed form of the _construct_U scatter (mcmc/util_cupy.py:426-442); HOST arrays N x N (bit-exact test). */
int mspde_bttb_index_map(int n, int32_t* row_idx_host, int32_t* col_idx_host, uint8_t* mask_host);

/* a2: sym2D(rfft2(exp(+-irfft2(u)))) -- FourierAnalysis_2D.constructMatexplicit up to the scatter (139-145). */
int mspde_field_coeffs(mspde_ctx* ctx, const void* u_sym, void* coef_plus, void* coef_minus);
/* a3+a4: L = (U[coef_plus]*D_diag - U[coef_minus])/sqrt_beta -- constructU + image_cupy.py:179, one fused pass. */
int mspde_assemble_L(mspde_ctx* ctx, const void* coef_plus, const void* coef_minus, double sqrt_beta, void* L, int ldl);
/* Lmatrix_2D.construct_from / construct_from_with_sqrt_beta (231-261). */
int mspde_construct_L(mspde_ctx* ctx, const void* u_sym, double sqrt_beta, void* L, int ldl);

/* a9: pCN._log_ratio (619-662, default branch); out3 (device) = {Phi, 0.5*||y@C||^2, sum log diag C}. */
int mspde_phi(mspde_ctx* ctx, const void* L, int ldl, double* out3_dev);

/* f2: the reference's "naive" Phi (use_naive_logRatio_implementation, 644-657; no stabiliser): Z = solve(L^H, H^H) by a
 * pivoted LU with M right-hand sides, C = chol(Z^H Z + I), Phi = 0.5*||C^-1 y||^2 + sum log diag C.
 * out3 = {Phi, 0.5*||C^-1 y||^2, sum log diag C}. */
int mspde_phi_naive(mspde_ctx* ctx, const void* L, int ldl, double* out3_dev);

/* Same Phi through the determinant lemma (algorithm note ii): no M-sized work.  Opt-in; results agree with
 * mspde_phi to rounding (validated at 1e-10 in tests); reported separately from the reference-faithful path. */
int mspde_phi_fast(mspde_ctx* ctx, const void* L, int ldl, double* out3_dev);

/* a5 + a8: cp.linalg.solve(L, w) (561, 681, 755, 785) and Lmatrix_2D.logDet (268-276; patch/norms.py:236-293).
 * Blocked LU with partial pivoting (LAPACK pivot rule), right-hand side carried as an extra column.
 * x (N) <- L^-1 rhs; logabsdet_dev (device double, may be NULL) <- log|det L| = sum log|diag| of the factor.
 * L is not modified (it is copied into the workspace). */
int mspde_solve(mspde_ctx* ctx, const void* L, int ldl, const void* rhs, void* x, double* logabsdet_dev);
/* Same solve through Householder QR instead of the pivoted LU (independent check; log|det| = sum log|r_jj|). */
int mspde_solve_qr(mspde_ctx* ctx, const void* L, int ldl, const void* rhs, void* x, double* logabsdet_dev);
/* a11: util.lstsq(a, b) = reduced QR -> q^H b -> solve_triangular (mcmc/util_cupy.py:590-609,
 * mcmc/extra_linalg.py:15-106); a is rows x cols (rows >= cols, within the context's (M+N) x N workspace). */
int mspde_lstsq(mspde_ctx* ctx, const void* a, int lda, int rows, int cols, const void* b, void* x);
/* v = lstsq(vstack(H, L), rhs) with rhs of length M+N (image_cupy.py:596-599). */
int mspde_gibbs_lstsq(mspde_ctx* ctx, const void* L, int ldl, const void* rhs, void* v);

/* a6-a12: pCN.one_step (553-617) with injected noise, accept decision on the device.
 * All pointers inside the structs are device pointers except sqrt_betas (host). */
typedef struct {
    int n_layers;
    double beta, betaZ;             /* pCN step (host-managed adaptation, 528-540) */
    const double* sqrt_betas;       /* host, n_layers */
    void** noise_cur;               /* host arrays of n_layers device pointers, each N complex */
    void** noise_new;
    void** u_cur;                   /* current_sample_sym */
    void** u_new;                   /* new_sample_sym */
    void** L_cur;                   /* per layer k >= 1: N x ldl; entry 0 unused */
    void** L_latest;
    int ldl;
    int phi_cur_valid;              /* phi_cache holds Phi(L_cur[last]) from a previous step of THIS chain, computed with the
                                       context's current inputs and the same Phi flavour (the host tracks both) */
    double* phi_cache;              /* device, 3 doubles owned by the chain: {Phi, quad, logdet} of L_cur[last]; NULL = a
                                       context-owned slot (one chain per context only) */
} mspde_chain;

typedef struct {
    const void** w_half;            /* host array of n_layers-1 device pointers (n_ravel complex each) */
    const void* w_last;             /* n_ravel complex: last layer's Gibbs noise */
    const double* e;                /* M real */
    double log_u;                   /* log of the accept uniform */
} mspde_step_noise;

enum { MSPDE_STEP_CACHE_PHI_CUR = 1, MSPDE_STEP_SKIP_GIBBS = 2, MSPDE_STEP_SERIAL = 4 /* one stream, no task overlap */,
       MSPDE_STEP_FAST_PHI = 8 /* determinant-lemma Phi (mspde_phi_fast) for both evaluations */,
       MSPDE_STEP_NAIVE_PHI = 16 /* the reference's naive Phi (mspde_phi_naive) for both evaluations */ };

/* out_dev (device, 4 doubles) <- {accepted (0/1), logRatio, Phi_cur, Phi_new}; record_dev (optional) <-
 * n_layers x n_ravel current half samples (Layer.record_sample, 792-795). */
int mspde_pcn_step(mspde_ctx* ctx, const mspde_chain* chain, const mspde_step_noise* noise, int flags, double* out_dev,
                   void* record_dev);

/* f1 (setup): Sinogram.get_measurement_matrix (image_cupy.py:364-396) = the _calculate_H_Tomography kernel
 * (mcmc/util_cupy.py:445-470) + ratio scaling (377-378) + roll fix (381-385) when apply_fixes != 0.
 * r_flat / theta_flat are r_grid.ravel('F') / theta_grid.ravel('F') (length M, radians), ix / iy the 'F'-ravelled
 * int32 index vectors (length N); H is M x N un-normalised. */
int mspde_tomography_H(void* stream, int M, int N, const double* r_flat, const double* theta_flat, const int32_t* ix,
                       const int32_t* iy, int n_r, int n_theta, int apply_fixes, void* H, int ldh);

/* Building blocks (exported for the parity tests):
 * zgemm_tn  : C = alpha*opA(A)^T*B + beta*C, A k x m, B k x n  (cuBLAS ZGEMM call sites 634, 639)
 * zpotrf    : upper Cholesky A = U^H U                         (cuSOLVER ZPOTRF call sites 637, 642)
 * ztrsm_uh  : B <- U^-H B                                      (cp.linalg.solve(c, H^H), 638) */
int mspde_zgemm_tn(void* stream, int conjA, int m, int n, int k, double alpha, const void* A, int lda, const void* B,
                   int ldb, double beta, void* C, int ldc, int uplo, void* splitk_ws, int nsplit);
int mspde_zpotrf_upper(void* stream, void* A, int lda, int n, void* winv, int* info_dev);
int mspde_ztrsm_uh(void* stream, const void* U, int ldu, const void* winv, int n, void* B, int ldb, int ncols);

/* f4: Welford running mean / M2 of u(x) = irfft2(sample) and of exp(u(x)) (m x m doubles each, device), one call per
 * recorded sample; count = number of samples including this one (util.updateWelford, mcmc/util_cupy.py:168-177;
 * post_analysis.py:57-105).  variance = M2/count, sample variance = M2/(count-1) (finalizeWelford, 181-187). */
int mspde_posterior_stats_update(mspde_ctx* ctx, const void* u_sym, long long count, double* mean_u, double* m2_u,
                                 double* mean_ell, double* m2_ell);

/* Panel-level pieces of the blocked LU / QR for the block-distributed driver (mspde/dist.py): the owner of a
 * 64-column panel factors it (mspde_*_panel), broadcasts the pieces named by mspde_*_work_layout (byte offset, byte size
 * pairs inside the caller-allocated workspace: QR -> V, V^T, T; LU -> L^T, (L11^-1)^T, row moves), and every rank applies
 * them to its own columns right of the panel (mspde_*_apply).  Same kernels as mspde_solve / mspde_lstsq. */
long long mspde_qr_work_bytes(int rows, int cols);
long long mspde_lu_work_bytes(int n, int ntot);
int mspde_qr_work_layout(int rows, int cols, void* work, long long* out6_host);
int mspde_lu_work_layout(int n, int ntot, void* work, long long* out6_host);
int mspde_qr_panel(void* stream, void* Ap, int lda, int mp, int pw, int rows, int cols, void* work, double* logabs_dev);
int mspde_qr_apply(void* stream, int mp, int rows, int cols, void* work, void* A2, int lda, int n2);
int mspde_lu_panel(void* stream, void* Ap, int lda, int mp, int pw, int goff, int n, int ntot, void* work, int* info_dev,
                   double* logabs_dev);
int mspde_lu_apply(void* stream, int mp, int pw, int n, int ntot, void* work, void* A12, int lda, int n2);
/* Rank-128 variant for the block-distributed QR over 128-column blocks: the owner of a block factors BOTH of its 64-column
 * panels (mspde_qr_pair_prepare: Ap = top-left element of the block, mp rows from there down, span <= 128 pivot columns;
 * *kk_host <- 64 or 128), broadcasts the five pieces named by mspde_qr_pair_work_layout (byte offset, byte size pairs: V
 * (rows x 128), V^T, T1, T2, V1^H V2), and every rank applies the pair to its own columns with ONE rank-kk update
 * (mspde_qr_pair_apply).  The workspace is sized / laid out for (wrows, wcols, wnaug) = the whole problem. */
long long mspde_qr_pair_work_bytes(int rows, int cols, int naug);
int mspde_qr_pair_work_layout(int rows, int cols, int naug, void* work, long long* out10_host);
int mspde_qr_pair_prepare(void* stream, void* work, int wrows, int wcols, int wnaug, void* Ap, int lda, int mp, int span,
                          double* logabs_dev, int* kk_host);
int mspde_qr_pair_apply(void* stream, void* work, int wrows, int wcols, int wnaug, int kk, int mp, void* A2, int lda, int n2);
/* x <- R^-1 d for the upper triangle R of A[0:n,0:n], d = column dcol of A (ztrsm call site, extra_linalg.py:102). */
int mspde_trsv_upper(void* stream, void* A, int lda, int n, int dcol, void* x);

/* Helpers of the block-distributed Phi (mspde/dist.py): A[i][i] += v; and, for an upper-trapezoidal row
 * panel V (nrows x ncols, diagonal at column 0): rowsq[j] = |sum_{i>=j} V[j][i] y[i]|^2, rowlog[j] = log Re V[j][j]
 * (the pieces of 0.5*||y@C||^2 - sum log diag C, image_cupy.py:643). */
int mspde_add_diag(void* stream, void* A, int lda, int n, double v);
int mspde_upper_rows_times_y(void* stream, const void* V, int ldv, int nrows, int ncols, const double* y, double* rowsq,
                             double* rowlog);

#ifdef __cplusplus
}
#endif
#endif
