// a6-a12: the pCN step around the dense kernels -- non-centred proposals, dense solves, accept/reject and the
// Gibbs draw of the last layer (pCN.one_step, /root/reference/mcmc/image_cupy.py:553-617).
// Everything is enqueued on the context stream; the accept decision is taken on the device, so the host only
// synchronises when it wants the flags.
#include <cstdlib>
#include "context.cuh"
#include "../../include/mspde.h"

namespace mspde {

namespace {

__device__ __forceinline__ cplx sym_at(const cplx* __restrict__ half, int j, int nr) {
    // util.symmetrize: concat(conj(half[:0:-1]), half)  (/root/reference/mcmc/util_cupy.py:201-204)
    if (j >= nr - 1) return half[j - (nr - 1)];
    cplx v = half[(nr - 1) - j];
    return make_double2(v.x, -v.y);
}

// Layer.sample_non_centered (image_cupy.py:789-790) fused with the layer-0 diagonal map (563)
__global__ void propose_kernel(const cplx* __restrict__ noise_cur, const cplx* __restrict__ w_half, double beta,
                               double betaZ, int N, int nr, cplx* __restrict__ noise_new,
                               const double* __restrict__ stdev_sym, cplx* __restrict__ u_new0) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const cplx w = sym_at(w_half, j, nr);
    const cplx c = noise_cur[j];
    cplx v;
    v.x = betaZ * c.x + beta * w.x;
    v.y = betaZ * c.y + beta * w.y;
    noise_new[j] = v;
    if (stdev_sym) {
        const double s = stdev_sym[j];
        u_new0[j] = make_double2(s * v.x, s * v.y);
    }
}

// logRatio = phi_cur - phi_new ; accepted = logRatio > log_u  (image_cupy.py:571-573)
__global__ void accept_kernel(const double* __restrict__ phi_cur3, const double* __restrict__ phi_new3, double log_u,
                              double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double lr = phi_cur3[0] - phi_new3[0];
        out[0] = (lr > log_u) ? 1.0 : 0.0;
        out[1] = lr;
        out[2] = phi_cur3[0];
        out[3] = phi_new3[0];
    }
}

// dst <- src when *flag != 0  (Layer.update_current_sample, image_cupy.py:802-807; set_current_L_to_latest 279-280)
__global__ void cond_copy_kernel(const double* __restrict__ flag, const double2* __restrict__ src,
                                 double2* __restrict__ dst, size_t count) {
    if (*flag == 0.0) return;
    size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (; i < count; i += stride) dst[i] = src[i];
}

__global__ void cond_copy_scalar_kernel(const double* __restrict__ flag, const double* __restrict__ src,
                                        double* __restrict__ dst, int count) {
    if (*flag == 0.0) return;
    if (threadIdx.x < count) dst[threadIdx.x] = src[threadIdx.x];
}

// column `col` of Aaug <- b
__global__ void set_col_kernel(cplx* __restrict__ A, int lda, int col, const cplx* __restrict__ b, int rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) A[size_t(i) * lda + col] = b[i];
}

__global__ void record_kernel(const cplx* __restrict__ u_sym, int nr, cplx* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nr) out[j] = u_sym[nr - 1 + j];   // current_sample = sym[n_ravel-1:]  (image_cupy.py:564, 794)
}

int copy2d(cudaStream_t st, cplx* dst, int ldd, const cplx* src, int lds, int rows, int cols) {
    MSPDE_CUDA(cudaMemcpy2DAsync(dst, size_t(ldd) * sizeof(cplx), src, size_t(lds) * sizeof(cplx),
                                 size_t(cols) * sizeof(cplx), rows, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // namespace

// x = L^-1 rhs by LU with partial pivoting (zgetrf + zgetrs semantics of cp.linalg.solve); logabsdet = sum log |u_jj|
// (patch/norms.py:276).  A singular pivot sets the context's info flag like LAPACK's info > 0.
int solve_L(Ctx* c, const cplx* L, int ldl, const cplx* rhs, cplx* x, double* logabsdet) {
    const int N = c->N;
    int rc = copy2d(c->stream, c->QRaug, c->ldA, L, ldl, N, N);
    if (rc) return rc;
    set_col_kernel<<<(N + 255) / 256, 256, 0, c->stream>>>(c->QRaug, c->ldA, N, rhs, N);
    MSPDE_LAUNCH_CHECK();
    rc = getrf_aug(c->stream, c->QRaug, c->ldA, N, 1, c->lu_work, c->info, logabsdet);
    if (rc) return rc;
    return trsv_upper_aug(c->stream, c->QRaug, c->ldA, N, N, x);
}

// Same solve through Householder QR of L (kept as an independent, equally backward-stable path; logabsdet = sum log |r_jj|).
int solve_L_qr(Ctx* c, const cplx* L, int ldl, const cplx* rhs, cplx* x, double* logabsdet) {
    const int N = c->N;
    int rc = copy2d(c->stream, c->QRaug, c->ldA, L, ldl, N, N);
    if (rc) return rc;
    set_col_kernel<<<(N + 255) / 256, 256, 0, c->stream>>>(c->QRaug, c->ldA, N, rhs, N);
    MSPDE_LAUNCH_CHECK();
    return lstsq_qr_aug(c->stream, c->QRaug, c->ldA, N, N, c->qr_work, x, logabsdet, &c->qr_la);
}

int lstsq_general(Ctx* c, const cplx* A, int lda, int rows, int cols, const cplx* b, cplx* x) {
    if (rows > c->M + c->N || cols > c->N || rows < cols) {
        set_error("mspde_lstsq: shape %d x %d exceeds the context workspace (%d x %d)", rows, cols, c->M + c->N, c->N);
        return -1;
    }
    int rc = copy2d(c->stream, c->QRaug, c->ldA, A, lda, rows, cols);
    if (rc) return rc;
    set_col_kernel<<<(rows + 255) / 256, 256, 0, c->stream>>>(c->QRaug, c->ldA, cols, b, rows);
    MSPDE_LAUNCH_CHECK();
    return lstsq_qr_aug(c->stream, c->QRaug, c->ldA, rows, cols, c->qr_work, x, nullptr, &c->qr_la);
}

// v = argmin || [H; L] v - rhs ||  (image_cupy.py:596-599)
int gibbs_lstsq(Ctx* c, const cplx* L, int ldl, const cplx* rhs, cplx* v) {
    const int N = c->N, M = c->M;
    int rc = copy2d(c->stream, c->QRaug, c->ldA, c->H, c->ldH, M, N);
    if (rc) return rc;
    rc = copy2d(c->stream, c->QRaug + size_t(M) * c->ldA, c->ldA, L, ldl, N, N);
    if (rc) return rc;
    set_col_kernel<<<(M + N + 255) / 256, 256, 0, c->stream>>>(c->QRaug, c->ldA, N, rhs, M + N);
    MSPDE_LAUNCH_CHECK();
    return lstsq_qr_aug(c->stream, c->QRaug, c->ldA, M + N, N, c->qr_work, v, nullptr, &c->qr_la);
}

namespace {
// noise_new candidates of the last layer for both outcomes of the accept test (A.5: line 583 uses noise_cur AFTER the
// accept copy): accepted -> base = previous noise_new, rejected -> base = noise_cur.  Column N / N+1 of the stacked
// system get the matching right-hand sides [y - e ; -candidate].
__global__ void gibbs_candidates_kernel(const cplx* __restrict__ noise_cur, const cplx* __restrict__ noise_new_prev,
                                        const cplx* __restrict__ w_half, double beta, double betaZ, int N, int nr, int M,
                                        const double* __restrict__ y, const double* __restrict__ e,
                                        cplx* __restrict__ cand_acc, cplx* __restrict__ cand_rej, cplx* __restrict__ A, int lda) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) {
        const cplx v = make_double2(y[i] - e[i], 0.0);
        A[size_t(i) * lda + N] = v;
        A[size_t(i) * lda + N + 1] = v;
    } else if (i < M + N) {
        const int j = i - M;
        const cplx w = sym_at(w_half, j, nr);
        const cplx ca = noise_new_prev[j], cr = noise_cur[j];
        const cplx va = make_double2(betaZ * ca.x + beta * w.x, betaZ * ca.y + beta * w.y);
        const cplx vr = make_double2(betaZ * cr.x + beta * w.x, betaZ * cr.y + beta * w.y);
        cand_acc[j] = va;
        cand_rej[j] = vr;
        A[size_t(i) * lda + N] = make_double2(0.0 - va.x, 0.0 - va.y);
        A[size_t(i) * lda + N + 1] = make_double2(0.0 - vr.x, 0.0 - vr.y);
    }
}

// dst <- (*flag != 0) ? a : b
__global__ void select_kernel(const double* __restrict__ flag, const cplx* __restrict__ a, const cplx* __restrict__ b,
                              cplx* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (*flag != 0.0) ? a[i] : b[i];
}

}  // namespace

int pcn_step(Ctx* c, const mspde_chain* ch, const mspde_step_noise* nz, int flags, double* out_dev, cplx* record_dev) {
    const bool serial = (flags & MSPDE_STEP_SERIAL) != 0;
    cudaStream_t caller = c->stream;
    cudaStream_t st = serial ? caller : c->s0;
    struct StreamGuard {   // helpers (field_coeffs, assemble_L, solve_L) launch on c->stream: point it at the main chain
        Ctx* c; cudaStream_t saved;
        ~StreamGuard() { c->stream = saved; }
    } guard{c, caller};
    const bool naive = (flags & MSPDE_STEP_NAIVE_PHI) != 0;      // both evaluations share one LU workspace: same stream
    cudaStream_t s1 = (serial || naive) ? st : c->s1, s2 = serial ? st : c->s2;
    if (flags & MSPDE_STEP_FAST_PHI) {
        int rc0 = phi_fast_prepare(c, caller);
        if (rc0) return rc0;
    }
    static const bool dbg = getenv("MSPDE_STEP_TIMING") != nullptr;
    static cudaEvent_t dt[6];
    if (dbg && !dt[0]) for (auto& e : dt) cudaEventCreate(&e);
    if (dbg) cudaEventRecord(dt[0], caller);
    if (!serial) {   // internal streams start after everything already queued on the caller's stream
        MSPDE_CUDA(cudaEventRecord(c->ev_start, caller));
        MSPDE_CUDA(cudaStreamWaitEvent(st, c->ev_start, 0));
        MSPDE_CUDA(cudaStreamWaitEvent(s1, c->ev_start, 0));
    }
    c->stream = st;
    const int K = ch->n_layers, N = c->N, M = c->M, nr = c->n_ravel;
    const int nb = (N + 255) / 256;
    double* phi_cur3 = ch->phi_cache ? ch->phi_cache : c->scal + 8;   // per-chain cache of Phi(L_cur[last])
    double* phi_new3 = c->scal + 12;
    cplx* cand_acc = c->vecN;
    cplx* cand_rej = c->vecN + N;
    cplx* v_acc = c->vecN + 2 * size_t(N);
    cplx* v_rej = c->vecN + 3 * size_t(N);
    const bool do_gibbs = !(flags & MSPDE_STEP_SKIP_GIBBS);
    const bool fast = (flags & MSPDE_STEP_FAST_PHI) != 0;
    int rc;
    // ---- stream s1: Phi(L_current) only depends on the previous step (image_cupy.py:567-568)
    const bool cached = (flags & MSPDE_STEP_CACHE_PHI_CUR) && ch->phi_cur_valid;
    if (!cached) {
        rc = naive ? phi_eval_naive(c, s1, c->pw[1], (const cplx*)ch->L_cur[K - 1], ch->ldl, phi_cur3)
             : fast ? phi_eval_fast(c, s1, c->pw[1], (const cplx*)ch->L_cur[K - 1], ch->ldl, phi_cur3)
                    : phi_eval(c, s1, c->pw[1], (const cplx*)ch->L_cur[K - 1], ch->ldl, phi_cur3, serial);
        if (rc) return rc;
        if (!serial && !naive) MSPDE_CUDA(cudaEventRecord(c->ev_s1, s1));
        if (dbg) cudaEventRecord(dt[1], s1);
    }
    // ---- main stream: layers 0..K-2 (557-564)
    for (int k = 0; k < K - 1; ++k) {
        propose_kernel<<<nb, 256, 0, st>>>((const cplx*)ch->noise_cur[k], (const cplx*)nz->w_half[k], ch->beta, ch->betaZ, N,
                                           nr, (cplx*)ch->noise_new[k], k == 0 ? c->stdev_sym : nullptr,
                                           (cplx*)ch->u_new[0]);
        MSPDE_LAUNCH_CHECK();
        if (k > 0) {
            rc = field_coeffs(c, (const cplx*)ch->u_new[k - 1], c->tp, c->tm);
            if (rc) return rc;
            rc = assemble_L(c, c->tp, c->tm, ch->sqrt_betas[k], (cplx*)ch->L_latest[k], ch->ldl);
            if (rc) return rc;
            rc = solve_L(c, (const cplx*)ch->L_latest[k], ch->ldl, (const cplx*)ch->noise_new[k], (cplx*)ch->u_new[k],
                         c->scal + 20);
            if (rc) return rc;
        }
    }
    rc = field_coeffs(c, (const cplx*)ch->u_new[K - 2], c->tp, c->tm);                                   // 570
    if (rc) return rc;
    rc = assemble_L(c, c->tp, c->tm, ch->sqrt_betas[K - 1], (cplx*)ch->L_latest[K - 1], ch->ldl);
    if (rc) return rc;
    if (dbg) cudaEventRecord(dt[2], st);
    // ---- stream s2: Gibbs least squares on [H; L_latest] with both candidate right-hand sides (583-599)
    if (do_gibbs) {
        if (!serial) {
            MSPDE_CUDA(cudaEventRecord(c->ev_L, st));
            MSPDE_CUDA(cudaStreamWaitEvent(s2, c->ev_L, 0));
        }
        rc = copy2d(s2, c->QRaug, c->ldA, c->H, c->ldH, M, N);
        if (rc) return rc;
        rc = copy2d(s2, c->QRaug + size_t(M) * c->ldA, c->ldA, (const cplx*)ch->L_latest[K - 1], ch->ldl, N, N);
        if (rc) return rc;
        gibbs_candidates_kernel<<<(M + N + 255) / 256, 256, 0, s2>>>((const cplx*)ch->noise_cur[K - 1],
                                                                    (const cplx*)ch->noise_new[K - 1], (const cplx*)nz->w_last,
                                                                    ch->beta, ch->betaZ, N, nr, M, c->y, nz->e, cand_acc,
                                                                    cand_rej, c->QRaug, c->ldA);
        MSPDE_LAUNCH_CHECK();
        rc = geqrf_aug(s2, c->QRaug, c->ldA, M + N, N, 2, c->qr_work, nullptr, &c->qr_la);
        if (rc) return rc;
        rc = trsv_upper_aug(s2, c->QRaug, c->ldA, N, N, v_acc);
        if (rc) return rc;
        rc = trsv_upper_aug(s2, c->QRaug, c->ldA, N, N + 1, v_rej);
        if (rc) return rc;
        if (!serial) MSPDE_CUDA(cudaEventRecord(c->ev_s2, s2));
        if (dbg) cudaEventRecord(dt[3], s2);
    } else {
        gibbs_candidates_kernel<<<(M + N + 255) / 256, 256, 0, st>>>((const cplx*)ch->noise_cur[K - 1],
                                                                    (const cplx*)ch->noise_new[K - 1], (const cplx*)nz->w_last,
                                                                    ch->beta, ch->betaZ, N, nr, M, c->y, nz->e, cand_acc,
                                                                    cand_rej, c->QRaug, c->ldA);
        MSPDE_LAUNCH_CHECK();
    }
    // ---- main stream: Phi(L_latest) (571), then join
    rc = naive ? phi_eval_naive(c, st, c->pw[0], (const cplx*)ch->L_latest[K - 1], ch->ldl, phi_new3)
         : fast ? phi_eval_fast(c, st, c->pw[0], (const cplx*)ch->L_latest[K - 1], ch->ldl, phi_new3)
                : phi_eval(c, st, c->pw[0], (const cplx*)ch->L_latest[K - 1], ch->ldl, phi_new3, serial);
    if (rc) return rc;
    if (dbg) cudaEventRecord(dt[4], st);
    if (!serial) {
        if (!cached && !naive) MSPDE_CUDA(cudaStreamWaitEvent(st, c->ev_s1, 0));
        if (do_gibbs) MSPDE_CUDA(cudaStreamWaitEvent(st, c->ev_s2, 0));
    }
    accept_kernel<<<1, 32, 0, st>>>(phi_cur3, phi_new3, nz->log_u, out_dev);                              // 573
    MSPDE_LAUNCH_CHECK();
    if (dbg) {
        cudaEventRecord(dt[5], st);
        cudaDeviceSynchronize();
        float a, b, d, e, f;
        cudaEventElapsedTime(&a, dt[0], dt[1]); cudaEventElapsedTime(&b, dt[0], dt[2]); cudaEventElapsedTime(&d, dt[0], dt[3]);
        cudaEventElapsedTime(&e, dt[0], dt[4]); cudaEventElapsedTime(&f, dt[0], dt[5]);
        fprintf(stderr, "[step] phi_cur done %.1f | L_latest ready %.1f | gibbs done %.1f | phi_new done %.1f | joined %.1f ms\n", a, b, d, e, f);
    }
    for (int k = 0; k < K; ++k) {                                                                        // 575-579
        cond_copy_kernel<<<nb, 256, 0, st>>>(out_dev, (const double2*)ch->u_new[k], (double2*)ch->u_cur[k], N);
        cond_copy_kernel<<<nb, 256, 0, st>>>(out_dev, (const double2*)ch->noise_new[k], (double2*)ch->noise_cur[k], N);
        if (k >= 1)
            cond_copy_kernel<<<148 * 8, 256, 0, st>>>(out_dev, (const double2*)ch->L_latest[k], (double2*)ch->L_cur[k],
                                                      size_t(N) * ch->ldl);
    }
    cond_copy_scalar_kernel<<<1, 32, 0, st>>>(out_dev, phi_new3, phi_cur3, 3);   // keeps a cached Phi(current) in step
    MSPDE_LAUNCH_CHECK();
    // last layer: noise_new and the Gibbs draw that match the accept outcome
    select_kernel<<<nb, 256, 0, st>>>(out_dev, cand_acc, cand_rej, (cplx*)ch->noise_new[K - 1], N);
    if (do_gibbs) select_kernel<<<nb, 256, 0, st>>>(out_dev, v_acc, v_rej, (cplx*)ch->u_new[K - 1], N);
    MSPDE_LAUNCH_CHECK();
    if (record_dev) {
        for (int k = 0; k < K; ++k) {
            record_kernel<<<(nr + 255) / 256, 256, 0, st>>>((const cplx*)ch->u_cur[k], nr, record_dev + size_t(k) * nr);
        }
        MSPDE_LAUNCH_CHECK();
    }
    if (!serial) {   // the caller's stream continues after the step
        MSPDE_CUDA(cudaEventRecord(c->ev_out, st));
        MSPDE_CUDA(cudaStreamWaitEvent(caller, c->ev_out, 0));
    }
    return 0;
}

}  // namespace mspde
