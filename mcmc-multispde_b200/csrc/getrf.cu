// Blocked LU with partial pivoting (row-major complex128) for the dense solves L x = w of the middle layers and for
// log|det L|.
//
// Reference call sites: cp.linalg.solve(L, w) (/root/reference/mcmc/image_cupy.py:561, 681, 755, 785: cuSOLVER
// zgetrf + zgetrs) and the patched complex slogdet (/root/reference/patch/norms.py:236-293: zgetrf, then
// sum(log(abs(diag)))).  Pivot rule as in LAPACK izamax: largest |re| + |im|, first occurrence.
//
// Right-looking over 64-column panels.  The right-hand side travels as an extra column, so L^-1 P b falls out of the
// trailing updates and only the back substitution with U remains (trsv_upper_aug).
//   1. lu_panel_kernel   cooperative launch; each CTA keeps a slab of panel rows in shared memory; per column one
//                        grid barrier: every CTA publishes its pivot candidate (value, row index, row contents), after
//                        the barrier every CTA picks the winner, the two rows are exchanged and the slab eliminated.
//                        The tail of the kernel emits L^T of the panel and the transposed inverse of its unit-lower
//                        diagonal block, plus the net row permutation of the panel (<= 128 moved rows).
//   2. gather/scatter    apply that permutation to the columns right of the panel (fully parallel, no serial swaps)
//   3. U12 = L11^-1 A12  in-place leaf zgemm_tn with the transposed inverse;  A22 -= L21 U12  zgemm_tn (K = 64)
#include <cstdlib>
#include "linalg.cuh"

namespace mspde {

namespace {

constexpr int PW = 64;
constexpr int SPITCH = PW + 1;
constexpr int MAXG = 148;
constexpr int MAXMOVE = 2 * PW;

struct LuWork {
    cplx* Lt;        // PW x ldlt      L^T of the current panel (strictly lower part, zero elsewhere)
    cplx* linvT;     // PW x PW        (L11^-1)^T
    cplx* tmp;       // MAXMOVE x ldt  gathered rows
    cplx* crowdata;  // 2 x MAXG x PW  candidate pivot rows
    cplx* currow;    // 2 x PW
    double* cval;    // 2 x MAXG
    int* crow;       // 2 x MAXG
    int* moves;      // [0] = count, then MAXMOVE dst, MAXMOVE src (panel-local rows)
    unsigned* bar;
    cplx* gslab;     // panel slabs when they do not fit in shared memory
    int ldlt, ldt;
};

constexpr size_t SMEM_SLAB_LIMIT = 150 * 1024;
size_t gslab_bytes(int n) {
    size_t slab = (size_t(n) + MAXG - 1) / MAXG;
    if (slab < PW) slab = PW;
    const bool need = slab * SPITCH * sizeof(cplx) > SMEM_SLAB_LIMIT || getenv("MSPDE_FORCE_GSLAB") != nullptr;
    return need ? (size_t(n) + MAXG * PW) * SPITCH * sizeof(cplx) : 256;
}

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) {}
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) lu_panel_kernel(cplx* __restrict__ A, int lda, int mp, int pw, int slab, int goff,
                                                       cplx* __restrict__ Lt, int ldlt, cplx* __restrict__ linvT,
                                                       cplx* crowdata, cplx* currow, double* cval, int* crow,
                                                       int* __restrict__ moves, double* __restrict__ logabs,
                                                       int* __restrict__ info, unsigned* barrier_counter, cplx* gslab) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    // slab of panel rows: shared memory when it fits, otherwise a CTA-private region of an L2-resident scratch buffer
    cplx* S = gslab ? gslab + size_t(blockIdx.x) * slab * SPITCH : reinterpret_cast<cplx*>(sm_raw) + PW * SPITCH;
    cplx* X = reinterpret_cast<cplx*>(sm_raw);   // [PW][SPITCH] scratch of CTA 0 for the inverse of the diagonal block
    __shared__ double wval[8];
    __shared__ int wrow[8];
    __shared__ cplx prow[PW], cur[PW];
    __shared__ int s_prow;
    __shared__ double s_pval;
    __shared__ int s_piv[PW];
    __shared__ int sdst[MAXMOVE], ssrc[MAXMOVE];
    const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, lane = tid & 31, wrp = tid >> 5;
    const int G = gridDim.x;
    const int row0 = blockIdx.x * slab;
    const int nrow = max(0, min(slab, mp - row0));

    for (int idx = tid; idx < nrow * PW; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        S[i * SPITCH + j] = (j < pw) ? A[size_t(row0 + i) * lda + j] : make_double2(0.0, 0.0);
    }
    __syncthreads();

    for (int j = 0; j < pw; ++j) {
        const int buf = j & 1;
        // ---- local pivot candidate over rows >= j: largest |re|+|im|, first occurrence
        double best = -1.0;
        int brow = 0x7fffffff;
        for (int i = tid; i < nrow; i += 256) {
            const int gi = row0 + i;
            if (gi >= j) {
                const cplx v = S[i * SPITCH + j];
                const double a = fabs(v.x) + fabs(v.y);
                if (a > best || (a == best && gi < brow)) { best = a; brow = gi; }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
            if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
        }
        if (lane == 0) { wval[wrp] = best; wrow[wrp] = brow; }
        __syncthreads();
        if (tid == 0) {
            double b = wval[0]; int r = wrow[0];
            for (int w = 1; w < 8; ++w)
                if (wval[w] > b || (wval[w] == b && wrow[w] < r)) { b = wval[w]; r = wrow[w]; }
            s_pval = b; s_prow = r;
            cval[buf * MAXG + blockIdx.x] = b;
            crow[buf * MAXG + blockIdx.x] = r;
        }
        __syncthreads();
        if (tid < PW) {
            const int r = s_prow;
            if (s_pval >= 0.0) crowdata[(size_t(buf) * MAXG + blockIdx.x) * PW + tid] = S[(r - row0) * SPITCH + tid];
            if (j >= row0 && j < row0 + nrow) currow[buf * PW + tid] = S[(j - row0) * SPITCH + tid];
        }
        grid_barrier(barrier_counter, unsigned(G) * unsigned(j + 1));
        // ---- global winner (every CTA, same order): warp 0 scans the G candidates
        if (wrp == 0) {
            double b = -1.0; int r = 0x7fffffff, c = 0;
            for (int q = lane; q < G; q += 32) {
                const double v = __ldcg(cval + buf * MAXG + q);
                const int rr = __ldcg(crow + buf * MAXG + q);
                if (v > b || (v == b && rr < r)) { b = v; r = rr; c = q; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, b, o);
                const int orr = __shfl_xor_sync(0xffffffffu, r, o);
                const int oc = __shfl_xor_sync(0xffffffffu, c, o);
                if (ob > b || (ob == b && orr < r)) { b = ob; r = orr; c = oc; }
            }
            if (lane == 0) { s_pval = b; s_prow = r; wrow[0] = c; }
        }
        __syncthreads();
        const double pval = s_pval;
        const int pr = s_prow;
        if (tid < PW) {
            prow[tid] = __ldcg(crowdata + (size_t(buf) * MAXG + wrow[0]) * PW + tid);
            cur[tid] = __ldcg(currow + buf * PW + tid);
        }
        if (tid == 0) s_piv[j] = (pval > 0.0) ? pr : j;
        __syncthreads();
        if (!(pval > 0.0)) {   // exactly singular column (or NaN): LAPACK zgetf2 records info and moves on
            if (blockIdx.x == 0 && tid == 0) {
                atomicCAS(info, 0, goff + j + 1);
                if (logabs) *logabs += log(0.0);
            }
            continue;
        }
        // ---- exchange rows j and pr inside the panel
        if (pr != j && tid < PW) {
            if (pr >= row0 && pr < row0 + nrow) S[(pr - row0) * SPITCH + tid] = cur[tid];
            if (j >= row0 && j < row0 + nrow) S[(j - row0) * SPITCH + tid] = prow[tid];
        }
        __syncthreads();
        const cplx pv = prow[j];
        const cplx rinv = cdiv(make_double2(1.0, 0.0), pv);     // zgetf2 scales by the reciprocal of the pivot
        if (blockIdx.x == 0 && tid == 0 && logabs) *logabs += log(hypot(pv.x, pv.y));
        {
            const int lo = max(0, j + 1 - row0);
            for (int i = lo + ty; i < nrow; i += 4) {
                const cplx l = cmul(S[i * SPITCH + j], rinv);
                if (tx > j && tx < pw) S[i * SPITCH + tx] = cfnma(l, prow[tx], S[i * SPITCH + tx]);
            }
        }
        __syncthreads();
        {
            const int lo = max(0, j + 1 - row0);
            for (int i = lo + tid; i < nrow; i += 256) S[i * SPITCH + j] = cmul(S[i * SPITCH + j], rinv);
        }
        __syncthreads();
    }
    // ---- write back the factored panel and L^T (strictly lower part)
    for (int idx = tid; idx < nrow * PW; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        if (j < pw) A[size_t(row0 + i) * lda + j] = S[i * SPITCH + j];
    }
    for (int idx = tid; idx < nrow * PW; idx += 256) {
        const int j = idx / nrow, i = idx % nrow;
        const int gi = row0 + i;
        Lt[size_t(j) * ldlt + gi] = (j < pw && gi > j) ? S[i * SPITCH + j] : make_double2(0.0, 0.0);
    }
    if (blockIdx.x != 0) return;
    // ---- CTA 0: net row permutation of the panel and (L11^-1)^T
    if (tid == 0) {
        int cnt = 0;
        int* dst = sdst;
        int* src = ssrc;                     // src[k] = row whose ORIGINAL content now lives at dst[k]
        for (int j = 0; j < pw; ++j) {
            const int a = j, b = s_piv[j];
            if (a == b) continue;
            int ia = -1, ib = -1;
            for (int k = 0; k < cnt; ++k) { if (dst[k] == a) ia = k; if (dst[k] == b) ib = k; }
            if (ia < 0) { ia = cnt; dst[cnt] = a; src[cnt] = a; ++cnt; }
            if (ib < 0) { ib = cnt; dst[cnt] = b; src[cnt] = b; ++cnt; }
            const int t = src[ia]; src[ia] = src[ib]; src[ib] = t;
        }
        moves[0] = cnt;
        for (int k = 0; k < cnt; ++k) { moves[1 + k] = dst[k]; moves[1 + MAXMOVE + k] = src[k]; }
    }
    // unit lower L11 sits in rows 0..pw-1 of this CTA's slab (slab >= 64 by construction)
    // X = L11^-1 computed row by row: X[i][:] = e_i - sum_{k<i} L[i][k] X[k][:]; thread t owns column t of X
    __syncthreads();
    {
        if (tid < PW) {
            const int c = tid;
            for (int i = 0; i < PW; ++i) {
                cplx s = (i == c) ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
                if (i < pw && c < pw) {
                    for (int k = c; k < i; ++k) s = cfnma(S[i * SPITCH + k], X[k * SPITCH + c], s);
                }
                X[i * SPITCH + c] = s;
            }
        }
        __syncthreads();
        for (int idx = tid; idx < PW * PW; idx += 256) {
            const int k = idx >> 6, i = idx & 63;           // linvT[k][i] = X[i][k]
            linvT[idx] = X[i * SPITCH + k];
        }
    }
}

// tmp[k][col] = A[src[k]][col]  /  A[dst[k]][col] = tmp[k][col]   for the moved rows of the panel
__global__ void gather_rows_kernel(const cplx* __restrict__ A, int lda, const int* __restrict__ moves, int ncols,
                                   cplx* __restrict__ tmp, int ldt) {
    const int k = blockIdx.y;
    if (k >= moves[0]) return;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    tmp[size_t(k) * ldt + col] = A[size_t(moves[1 + MAXMOVE + k]) * lda + col];
}
__global__ void scatter_rows_kernel(cplx* __restrict__ A, int lda, const int* __restrict__ moves, int ncols,
                                    const cplx* __restrict__ tmp, int ldt) {
    const int k = blockIdx.y;
    if (k >= moves[0]) return;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    A[size_t(moves[1 + k]) * lda + col] = tmp[size_t(k) * ldt + col];
}

LuWork carve(void* work, int n, int ntot) {
    LuWork w;
    unsigned char* p = reinterpret_cast<unsigned char*>(work);
    auto take = [&](size_t bytes) { void* q = p; p += align_up(bytes); return q; };
    w.ldlt = (n + 7) / 8 * 8;
    w.ldt = (ntot + 7) / 8 * 8;
    w.Lt = (cplx*)take(size_t(PW) * w.ldlt * sizeof(cplx));
    w.linvT = (cplx*)take(size_t(PW) * PW * sizeof(cplx));
    w.tmp = (cplx*)take(size_t(MAXMOVE) * w.ldt * sizeof(cplx));
    w.crowdata = (cplx*)take(size_t(2) * MAXG * PW * sizeof(cplx));
    w.currow = (cplx*)take(size_t(2) * PW * sizeof(cplx));
    w.cval = (double*)take(size_t(2) * MAXG * sizeof(double));
    w.crow = (int*)take(size_t(2) * MAXG * sizeof(int));
    w.moves = (int*)take(size_t(1 + 2 * MAXMOVE) * sizeof(int));
    w.bar = (unsigned*)take(256);
    w.gslab = (cplx*)take(gslab_bytes(n));
    return w;
}

DeviceOnce g_attr;

}  // namespace

size_t lu_work_bytes(int n, int ntot) {
    const size_t ldlt = (n + 7) / 8 * 8, ldt = (ntot + 7) / 8 * 8;
    return align_up(size_t(PW) * ldlt * sizeof(cplx)) + align_up(size_t(PW) * PW * sizeof(cplx)) +
           align_up(size_t(MAXMOVE) * ldt * sizeof(cplx)) + align_up(size_t(2) * MAXG * PW * sizeof(cplx)) +
           align_up(size_t(2) * PW * sizeof(cplx)) + align_up(size_t(2) * MAXG * sizeof(double)) +
           align_up(size_t(2) * MAXG * sizeof(int)) + align_up(size_t(1 + 2 * MAXMOVE) * sizeof(int)) + 256 +
           align_up(gslab_bytes(n)) + 4096;
}

// One 64-column panel of the pivoted LU: Ap (mp rows x pw pivot columns) factored in place; the workspace (laid out by
// carve(work, n, ntot)) receives L^T, (L11^-1)^T and the net row permutation of the panel.
int lu_panel_factor(cudaStream_t st, cplx* Ap, int lda, int mp, int pw, int goff, int n, int ntot, void* work, int* info,
                    double* logabs) {
    if (g_attr.first()) {
        MSPDE_CUDA(cudaFuncSetAttribute(lu_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    }
    LuWork w = carve(work, n, ntot);
    int G = min(MAXG, (mp + 63) / 64);
    int slab = (mp + G - 1) / G;
    if (slab < PW) slab = PW;               // CTA 0 must hold the whole diagonal block
    G = (mp + slab - 1) / slab;
    size_t smem = (size_t(slab) + PW) * SPITCH * sizeof(cplx);
    cplx* gslab = nullptr;
    static const bool force_global = getenv("MSPDE_FORCE_GSLAB") != nullptr;   // test hook for the tall-panel path
    if (force_global || size_t(slab) * SPITCH * sizeof(cplx) > SMEM_SLAB_LIMIT) {   // tall panel: slabs live in the L2-resident scratch
        gslab = w.gslab;
        smem = size_t(PW) * SPITCH * sizeof(cplx);
    }
    MSPDE_CUDA(cudaMemsetAsync(w.bar, 0, sizeof(unsigned), st));
    int lda_ = lda, mp_ = mp, pw_ = pw, slab_ = slab, goff_ = goff, ldlt = w.ldlt;
    cplx* Lt = w.Lt; cplx* linvT = w.linvT; cplx* crowdata = w.crowdata; cplx* currow = w.currow;
    double* cval = w.cval; int* crow = w.crow; int* moves = w.moves; double* la = logabs; int* inf = info;
    unsigned* bar = w.bar;
    void* args[] = {&Ap, &lda_, &mp_, &pw_, &slab_, &goff_, &Lt, &ldlt, &linvT, &crowdata, &currow, &cval, &crow,
                    &moves, &la, &inf, &bar, &gslab};
    MSPDE_CUDA(cudaLaunchCooperativeKernel((void*)lu_panel_kernel, dim3(G), dim3(256), args, smem, st));
    ++g_launches;
    return 0;
}

// Apply a factored panel to n2 columns right of it: A12 points at (panel row 0, first of those columns); rows are
// permuted, U12 = L11^-1 A12 (first pw rows), then the mp - pw rows below get -= L21 U12.
int lu_panel_apply(cudaStream_t st, int mp, int pw, int n, int ntot, void* work, cplx* A12, int lda, int n2) {
    if (n2 <= 0) return 0;
    LuWork w = carve(work, n, ntot);
    gather_rows_kernel<<<dim3((n2 + 255) / 256, MAXMOVE), 256, 0, st>>>(A12, lda, w.moves, n2, w.tmp, w.ldt);
    MSPDE_LAUNCH_CHECK();
    scatter_rows_kernel<<<dim3((n2 + 255) / 256, MAXMOVE), 256, 0, st>>>(A12, lda, w.moves, n2, w.tmp, w.ldt);
    MSPDE_LAUNCH_CHECK();
    int rc = zgemm_tn(st, false, pw, n2, pw, 1.0, w.linvT, PW, A12, lda, 0.0, A12, lda, GEMM_FULL);   // in place, one row tile
    if (rc) return rc;
    if (mp > pw)
        rc = zgemm_tn(st, false, mp - pw, n2, pw, -1.0, w.Lt + pw, w.ldlt, A12, lda, 1.0, A12 + size_t(pw) * lda, lda, GEMM_FULL);
    return rc;
}

void lu_work_layout(int n, int ntot, void* work, long long* out6) {
    LuWork w = carve(work, n, ntot);
    unsigned char* base = reinterpret_cast<unsigned char*>(work);
    out6[0] = reinterpret_cast<unsigned char*>(w.Lt) - base;     out6[1] = (long long)PW * w.ldlt * sizeof(cplx);
    out6[2] = reinterpret_cast<unsigned char*>(w.linvT) - base;  out6[3] = (long long)PW * PW * sizeof(cplx);
    out6[4] = reinterpret_cast<unsigned char*>(w.moves) - base;  out6[5] = (long long)(1 + 2 * MAXMOVE) * sizeof(int);
}

// In-place P A = L U of the leading n x n block of A (n x (n + naug)); the naug trailing columns receive L^-1 P b.
// *info (device) <- first exactly-zero pivot (1-based) if any; logabs (device, optional) <- sum log |u_jj|.
int getrf_aug(cudaStream_t st, cplx* A, int lda, int n, int naug, void* work, int* info, double* logabs) {
    const int ntot = n + naug;
    if (logabs) MSPDE_CUDA(cudaMemsetAsync(logabs, 0, sizeof(double), st));
    for (int c0 = 0; c0 < n; c0 += PW) {
        const int pw = min(PW, n - c0);
        const int mp = n - c0;
        cplx* Ap = A + size_t(c0) * lda + c0;
        int rc = lu_panel_factor(st, Ap, lda, mp, pw, c0, n, ntot, work, info, logabs);
        if (rc) return rc;
        rc = lu_panel_apply(st, mp, pw, n, ntot, work, Ap + pw, lda, ntot - (c0 + pw));
        if (rc) return rc;
    }
    return 0;
}

}  // namespace mspde
