// Blocked Householder QR least squares (row-major complex128) for the Gibbs draw of the last layer and for the
// dense solves L x = w of the middle layers.
//
// Reference call sites: util.lstsq = cupy.linalg.qr -> q^H b -> solve_triangular
// (/root/reference/mcmc/util_cupy.py:590-609, /root/reference/mcmc/extra_linalg.py:15-106) and
// cp.linalg.solve(L, w) (/root/reference/mcmc/image_cupy.py:561).  No explicit Q is formed: the right-hand side
// travels as an extra column of the matrix, so Q^H b falls out of the trailing updates.
//
// Per 64-column panel:
//   1. qr_panel_kernel   cooperative launch, each CTA keeps a slab of panel rows in shared memory; one grid
//                        barrier per column: the column norm and all v^H a_j' products are reduced together
//   2. V^H V (split-K zgemm_tn) -> larft_kernel builds the compact-WY factor T
//   3. W = V^H A2 (split-K zgemm_tn), W <- T^H W (in-place leaf zgemm_tn), A2 -= V W (zgemm_tn, K = 64)
#include <cooperative_groups.h>
#include <cstdlib>
#include "linalg.cuh"

namespace cg = cooperative_groups;

namespace mspde {

namespace {

constexpr int PW = 64;           // panel width
constexpr int SPITCH = PW + 1;   // shared-memory row pitch (complex)
constexpr int MAXG = 148;        // CTAs of the cooperative panel kernel (one per SM)
constexpr size_t SMALL_SMEM = (size_t(PW) * (PW + 1) + PW) * sizeof(cplx);

struct QrWork {
    cplx* V;      // rows x PW   explicit reflectors (unit diagonal, zeros above)
    cplx* Vt;     // PW x ldvt   transposed copy
    cplx* W;      // PW x ldw
    cplx* part;   // split-K partials
    cplx* Gm;     // PW x PW  V^H V
    cplx* T;      // PW x PW  compact WY factor
    cplx* tau;    // PW
    cplx* dots;   // 2 x MAXG x PW   per-CTA partial dot products (double buffered)
    cplx* currow; // 2 x PW          current diagonal row (double buffered)
    unsigned* bar; // grid barrier counter
    cplx* gslab;  // rows x SPITCH   panel slabs when they do not fit in shared memory (tall panels, e.g. n = 96)
    int ldvt, ldw, nsplit;
};

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }
// split-K partial products: up to 148 slices of the 64 x 64 Gram matrix, or ~10 slices of the 64 x (cols+1) W panel
constexpr size_t SMEM_SLAB_LIMIT = 200 * 1024;
size_t gslab_bytes(int rows) {   // only needed when ceil(rows/148) rows x 65 complex exceed the shared-memory budget
    const size_t slab = (size_t(rows) + MAXG - 1) / MAXG;
    const bool need = slab * SPITCH * sizeof(cplx) > SMEM_SLAB_LIMIT || getenv("MSPDE_FORCE_GSLAB") != nullptr;
    return need ? (size_t(rows) + MAXG) * SPITCH * sizeof(cplx) : 256;
}
size_t part_elems(int cols) {
    const size_t ldw = (cols + 1 + 7) / 8 * 8;
    const size_t a = size_t(149) * PW * PW, b = size_t(12) * PW * ldw;
    return a > b ? a : b;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Grid-wide barrier on a monotonically increasing counter (zeroed before the launch).  Data exchanged across it is
// written with plain stores before the barrier and read with ld.global.cg (L1 bypass) after it.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) {}
    }
    __syncthreads();
}
__device__ __forceinline__ cplx ldcg(const cplx* p) { return __ldcg(p); }

// Panel: A[c0 + (0..mp-1)][c0 + (0..pw-1)], pw <= PW pivot columns.
__global__ void __launch_bounds__(256) qr_panel_kernel(cplx* __restrict__ A, int lda, int mp, int pw, int slab,
                                                       cplx* __restrict__ V, int ldv, cplx* __restrict__ Vt, int ldvt,
                                                       cplx* __restrict__ tau_out, cplx* dots, cplx* currow,
                                                       double* __restrict__ logabs, unsigned* barrier_counter,
                                                       cplx* gslab) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    // slab of panel rows: shared memory when it fits, otherwise a CTA-private region of an L2-resident scratch buffer
    cplx* S = gslab ? gslab + size_t(blockIdx.x) * slab * SPITCH : reinterpret_cast<cplx*>(sm_raw);   // [slab][SPITCH]
    __shared__ cplx red[4][PW];
    __shared__ cplx sdot[PW];
    __shared__ cplx srow[PW];
    const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
    const int G = gridDim.x;
    const int row0 = blockIdx.x * slab;                 // first panel-local row of this slab
    const int nrow = max(0, min(slab, mp - row0));

    for (int idx = tid; idx < nrow * PW; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        S[i * SPITCH + j] = (j < pw) ? A[size_t(row0 + i) * lda + j] : make_double2(0.0, 0.0);
    }
    __syncthreads();

#ifdef MSPDE_QR_PROFILE
    long long tA = 0, tB = 0, tC = 0, tD = 0, t0, t1;
#define TICK(acc) t1 = clock64(); acc += t1 - t0; t0 = t1;
    t0 = clock64();
#else
#define TICK(acc)
#endif
    for (int j = 0; j < pw; ++j) {
        const int buf = j & 1;
        // ---- partial dots over rows strictly below the diagonal row j:  sum_i conj(a[i][j]) a[i][j'] , j' >= j
        cplx acc = make_double2(0.0, 0.0);
        if (tx >= j && tx < pw) {
            const int lo = max(0, j + 1 - row0);
            cplx acc2 = make_double2(0.0, 0.0);
            int i = lo + ty;
            for (; i + 4 < nrow; i += 8) {
                acc = cfmac(S[i * SPITCH + j], S[i * SPITCH + tx], acc);
                acc2 = cfmac(S[(i + 4) * SPITCH + j], S[(i + 4) * SPITCH + tx], acc2);
            }
            if (i < nrow) acc = cfmac(S[i * SPITCH + j], S[i * SPITCH + tx], acc);
            acc = cadd(acc, acc2);
        }
        red[ty][tx] = acc;
        __syncthreads();
        if (ty == 0) {
            cplx s = cadd(cadd(red[0][tx], red[1][tx]), cadd(red[2][tx], red[3][tx]));
            dots[(size_t(buf) * PW + tx) * MAXG + blockIdx.x] = s;          // [column][cta]: coalesced for the readers
            if (j >= row0 && j < row0 + nrow) currow[buf * PW + tx] = S[(j - row0) * SPITCH + tx];
        }
        TICK(tA)
        grid_barrier(barrier_counter, unsigned(G) * unsigned(j + 1));
        TICK(tB)
        // ---- every CTA reduces the G partials of the live columns in the same fixed order: warp w takes columns
        //      w, w+8, ...; lane l sums partials l, l+32, ... and a butterfly finishes.  All loads are issued first.
        {
            const int wrp = tid >> 5, ln = tid & 31;
            cplx part[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const cplx* src = dots + (size_t(buf) * PW + (wrp + 8 * q)) * MAXG;     // columns < j hold stale, unused data
                cplx v[5];
#pragma unroll
                for (int h = 0; h < 5; ++h) {
                    const int cidx = ln + 32 * h;
                    const bool live = (cidx < G) && (wrp + 8 * q >= j);
                    v[h] = live ? ldcg(src + cidx) : make_double2(0.0, 0.0);
                }
                part[q] = cadd(cadd(cadd(v[0], v[1]), cadd(v[2], v[3])), v[4]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (wrp + 8 * q < j) continue;     // warp-uniform
                cplx a = part[q];
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
                    a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
                }
                if (ln == 0) sdot[wrp + 8 * q] = a;
            }
            if (tid < PW) srow[tid] = ldcg(currow + buf * PW + tid);
        }
        __syncthreads();
        TICK(tC)
        // ---- zlarfg: alpha = a_jj, beta = -sign(Re alpha) * sqrt(|alpha|^2 + xnorm^2), tau = (beta - alpha)/beta
        const cplx alpha = srow[j];
        const double xn2 = sdot[j].x;
        cplx tau = make_double2(0.0, 0.0), scale = make_double2(0.0, 0.0);
        double beta = alpha.x;
        const bool trivial = (xn2 == 0.0 && alpha.y == 0.0);
        if (!trivial) {
            beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2), alpha.x);
            tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
            scale = cdiv(make_double2(1.0, 0.0), make_double2(alpha.x - beta, alpha.y));
        }
        if (blockIdx.x == 0 && tid == 0) {
            tau_out[j] = tau;
            if (logabs) *logabs += log(fabs(beta));
        }
        // w_j' = v^H a_j' = a[j][j'] + conj(scale) * dot_j' ; coefficient applied to rows: conj(tau) * w_j'
        cplx coef = make_double2(0.0, 0.0);
        if (tx > j && tx < pw && !trivial) {
            cplx w = cadd(srow[tx], cmulc(scale, sdot[tx]));
            coef = cmulc(tau, w);
        }
        // ---- apply H_j^H to the slab and store v_j in column j
        {
            const int lo = max(0, j - row0);
            for (int i = lo + ty; i < nrow; i += 4) {
                const int gi = row0 + i;
                if (gi == j) {
                    if (tx > j && tx < pw) S[i * SPITCH + tx] = csub(S[i * SPITCH + tx], coef);
                } else {
                    const cplx vij = cmul(S[i * SPITCH + j], scale);
                    if (tx > j && tx < pw) S[i * SPITCH + tx] = cfnma(coef, vij, S[i * SPITCH + tx]);
                }
            }
        }
        __syncthreads();
        {
            const int lo = max(0, j - row0);
            for (int i = lo + tid; i < nrow; i += 256) {
                const int gi = row0 + i;
                if (gi == j) S[i * SPITCH + j] = make_double2(beta, 0.0);
                else S[i * SPITCH + j] = trivial ? make_double2(0.0, 0.0) : cmul(S[i * SPITCH + j], scale);
            }
        }
        __syncthreads();
        TICK(tD)
    }
#ifdef MSPDE_QR_PROFILE
    if (tid == 0 && blockIdx.x == 0) printf("panel mp=%d G=%d pw=%d cycles/col: dots %lld sync %lld reduce %lld update %lld\n", mp, G, pw, tA / pw, tB / pw, tC / pw, tD / pw);
#endif
    // ---- write back: R on/above the diagonal into A (reflector storage below it), explicit V and V^T to workspace
    for (int idx = tid; idx < nrow * PW; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        const int gi = row0 + i;
        const cplx v = S[i * SPITCH + j];
        if (j < pw) A[size_t(gi) * lda + j] = v;
        cplx e = make_double2(0.0, 0.0);
        if (j < pw) {
            if (gi > j) e = v;
            else if (gi == j) e = make_double2(1.0, 0.0);
        }
        V[size_t(gi) * ldv + j] = e;
    }
    for (int idx = tid; idx < nrow * PW; idx += 256) {
        const int j = idx / nrow, i = idx % nrow;
        const int gi = row0 + i;
        cplx e = make_double2(0.0, 0.0);
        if (j < pw) {
            if (gi > j) e = S[i * SPITCH + j];
            else if (gi == j) e = make_double2(1.0, 0.0);
        }
        Vt[size_t(j) * ldvt + gi] = e;
    }
}

// T[j][j] = tau_j ; T[0:j, j] = -tau_j * T[0:j,0:j] * (V^H V)[0:j, j]      (zlarft, forward, columnwise)
// 256 threads: row a = tid/4, the dot product over b = a..j-1 is split over 4 consecutive lanes.
__global__ void __launch_bounds__(256) larft_kernel(const cplx* __restrict__ Gm, const cplx* __restrict__ tau, int pw,
                                                    cplx* __restrict__ T) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    cplx (*sT)[PW + 1] = reinterpret_cast<cplx (*)[PW + 1]>(sm_raw);
    cplx (*sG)[PW + 1] = reinterpret_cast<cplx (*)[PW + 1]>(sm_raw + sizeof(cplx) * PW * (PW + 1));
    const int tid = threadIdx.x, a = tid >> 2, part = tid & 3;
    for (int idx = tid; idx < PW * PW; idx += 256) {
        sT[idx >> 6][idx & 63] = make_double2(0.0, 0.0);
        sG[idx >> 6][idx & 63] = Gm[idx];
    }
    __syncthreads();
    for (int j = 0; j < pw; ++j) {
        const cplx tj = tau[j];
        cplx s = make_double2(0.0, 0.0);
        if (a < j)
            for (int b = a + part; b < j; b += 4) s = cfma(sT[a][b], sG[b][j], s);
        s.x += __shfl_xor_sync(0xffffffffu, s.x, 1);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, 1);
        s.x += __shfl_xor_sync(0xffffffffu, s.x, 2);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, 2);
        if (part == 0) {
            if (a < j) sT[a][j] = cmul(make_double2(-tj.x, -tj.y), s);
            else if (a == j) sT[a][j] = tj;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < PW * PW; idx += 256) T[idx] = sT[idx >> 6][idx & 63];
}

// x_blk = R_bb^-1 d_blk for one 64-block (single CTA), R upper; d lives in column `dcol` of the same matrix
__global__ void __launch_bounds__(64) trsv_diag_kernel(const cplx* __restrict__ A, int lda, int b0, int nb, int dcol,
                                                       cplx* __restrict__ x) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    cplx (*R)[PW + 1] = reinterpret_cast<cplx (*)[PW + 1]>(sm_raw);
    cplx* xs = reinterpret_cast<cplx*>(sm_raw) + PW * (PW + 1);
    const int t = threadIdx.x;
    for (int i = 0; i < nb; ++i)
        if (t < nb) R[i][t] = A[size_t(b0 + i) * lda + b0 + t];
    if (t < nb) xs[t] = A[size_t(b0 + t) * lda + dcol];
    __syncthreads();
    for (int i = nb - 1; i >= 0; --i) {
        if (t == i) xs[i] = cdiv(xs[i], R[i][i]);
        __syncthreads();
        if (t < i) xs[t] = cfnma(R[t][i], xs[i], xs[t]);
        __syncthreads();
    }
    if (t < nb) x[b0 + t] = xs[t];
}

// d[i] -= sum_{k in block} R[i][b0+k] x[b0+k] for all rows i < b0 (one warp per row)
__global__ void __launch_bounds__(256) trsv_update_kernel(cplx* __restrict__ A, int lda, int b0, int nb, int dcol,
                                                          const cplx* __restrict__ x) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= b0) return;
    cplx s = make_double2(0.0, 0.0);
    for (int k = lane; k < nb; k += 32) s = cfma(A[size_t(i) * lda + b0 + k], x[b0 + k], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    }
    if (lane == 0) {
        cplx* d = A + size_t(i) * lda + dcol;
        *d = csub(*d, s);
    }
}

QrWork carve(void* work, int rows, int cols) {
    QrWork w;
    unsigned char* p = reinterpret_cast<unsigned char*>(work);
    auto take = [&](size_t bytes) { void* q = p; p += align_up(bytes); return q; };
    w.ldvt = (rows + 7) / 8 * 8;
    w.ldw = (cols + 1 + 7) / 8 * 8;
    w.nsplit = 148;
    w.V = (cplx*)take(size_t(rows) * PW * sizeof(cplx));
    w.Vt = (cplx*)take(size_t(PW) * w.ldvt * sizeof(cplx));
    w.W = (cplx*)take(size_t(PW) * w.ldw * sizeof(cplx));
    w.part = (cplx*)take(part_elems(cols) * sizeof(cplx));
    w.Gm = (cplx*)take(size_t(PW) * PW * sizeof(cplx));
    w.T = (cplx*)take(size_t(PW) * PW * sizeof(cplx));
    w.tau = (cplx*)take(size_t(PW) * sizeof(cplx));
    w.dots = (cplx*)take(size_t(2) * MAXG * PW * sizeof(cplx));
    w.currow = (cplx*)take(size_t(2) * PW * sizeof(cplx));
    w.bar = (unsigned*)take(256);
    w.gslab = (cplx*)take(gslab_bytes(rows));
    return w;
}

DeviceOnce g_attr;

}  // namespace

size_t qr_work_bytes(int rows, int cols) {
    const size_t ldvt = (rows + 7) / 8 * 8, ldw = (cols + 1 + 7) / 8 * 8;
    size_t b = 0;
    b += align_up(size_t(rows) * PW * sizeof(cplx));
    b += align_up(size_t(PW) * ldvt * sizeof(cplx));
    b += align_up(size_t(PW) * ldw * sizeof(cplx));
    b += align_up(part_elems(cols) * sizeof(cplx));
    b += 3 * align_up(size_t(PW) * PW * sizeof(cplx));
    b += align_up(size_t(2) * MAXG * PW * sizeof(cplx));
    b += align_up(size_t(2) * PW * sizeof(cplx));
    b += align_up(gslab_bytes(rows));
    return b + 4096;
}

// Householder QR of the leading `cols` columns of Aaug (rows x (cols + naug)), trailing columns get Q^H applied.
// logabs (device, optional) accumulates sum log |r_jj|.
static int init_attrs() {
    if (g_attr.first()) {
        MSPDE_CUDA(cudaFuncSetAttribute(qr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MSPDE_CUDA(cudaFuncSetAttribute(larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(cplx) * PW * (PW + 1))));
        MSPDE_CUDA(cudaFuncSetAttribute(trsv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM));
    }
    return 0;
}

// Factor one 64-column panel Ap (mp rows x pw pivot columns, leading dimension lda): R and the reflectors in place, and in
// the workspace (laid out by carve(work, rows, cols)) the explicit V (mp x 64), V^T (64 x ldvt) and the compact-WY T.
// Launch of the cooperative panel kernel: R and the reflectors in place in Ap, explicit V (leading dimension ldv) and V^T.
static int launch_panel(cudaStream_t st, cplx* Ap, int lda, int mp, int pw, cplx* V, int ldv, cplx* Vt, int ldvt, cplx* tau,
                        cplx* dots, cplx* currow, unsigned* bar, cplx* gslab_buf, double* logabs) {
    int G = min(MAXG, (mp + 63) / 64);
    int slab = (mp + G - 1) / G;
    G = (mp + slab - 1) / slab;
    size_t smem = size_t(slab) * SPITCH * sizeof(cplx);
    cplx* gslab = nullptr;
    static const bool force_global = getenv("MSPDE_FORCE_GSLAB") != nullptr;   // test hook for the tall-panel path
    if (force_global || smem > SMEM_SLAB_LIMIT) {   // tall panel: slabs live in the L2-resident scratch instead
        gslab = gslab_buf;
        smem = 0;
    }
    int mp_ = mp, pw_ = pw, slab_ = slab, lda_ = lda, ldv_ = ldv, ldvt_ = ldvt;
    double* la = logabs;
    MSPDE_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
    void* args[] = {&Ap, &lda_, &mp_, &pw_, &slab_, &V, &ldv_, &Vt, &ldvt_, &tau, &dots, &currow, &la, &bar, &gslab};
    MSPDE_CUDA(cudaLaunchCooperativeKernel((void*)qr_panel_kernel, dim3(G), dim3(256), args, smem, st));
    ++g_launches;
    return 0;
}

int qr_panel_factor(cudaStream_t st, cplx* Ap, int lda, int mp, int pw, int rows, int cols, void* work, double* logabs) {
    if (init_attrs()) return -1;
    QrWork w = carve(work, rows, cols);
    int rc = launch_panel(st, Ap, lda, mp, pw, w.V, PW, w.Vt, w.ldvt, w.tau, w.dots, w.currow, w.bar, w.gslab, logabs);
    if (rc) return rc;
    // T from V^H V
    int ns_g = (mp + 127) / 128;
    if (ns_g > w.nsplit) ns_g = w.nsplit;
    rc = zgemm_tn_splitk(st, true, PW, PW, mp, 1.0, w.V, PW, w.V, PW, 0.0, w.Gm, PW, w.part, ns_g);
    if (rc) return rc;
    larft_kernel<<<1, 256, 2 * sizeof(cplx) * PW * (PW + 1), st>>>(w.Gm, w.tau, pw, w.T);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

// A2 (mp x n2) <- (I - V T^H V^H) A2 with the V, V^T, T held in the workspace.
int qr_panel_apply(cudaStream_t st, int mp, int rows, int cols, void* work, cplx* A2, int lda, int n2) {
    if (n2 <= 0) return 0;
    QrWork w = carve(work, rows, cols);
    const int tiles = (n2 + 63) / 64;
    int ns_w = (592 + tiles - 1) / tiles;                 // ~2 waves of 296 resident CTAs
    const int ns_cap = (mp + 255) / 256;                  // keep >= 256 k per slice
    if (ns_w > ns_cap) ns_w = ns_cap;
    const int ns_fit = int(part_elems(cols) / (size_t(PW) * size_t(n2)));
    if (ns_w > ns_fit) ns_w = ns_fit;                     // partial-product workspace
    if (ns_w < 1) ns_w = 1;
    int rc = zgemm_tn_splitk(st, true, PW, n2, mp, 1.0, w.V, PW, A2, lda, 0.0, w.W, w.ldw, w.part, ns_w);   // W = V^H A2
    if (rc) return rc;
    rc = zgemm_tn(st, true, PW, n2, PW, 1.0, w.T, PW, w.W, w.ldw, 0.0, w.W, w.ldw, GEMM_FULL);   // W <- T^H W (in place)
    if (rc) return rc;
    return zgemm_tn(st, false, mp, n2, PW, -1.0, w.Vt, w.ldvt, w.W, w.ldw, 1.0, A2, lda, GEMM_FULL);   // A2 -= V W
}

// byte offsets / sizes of the pieces of the workspace a distributed driver has to broadcast
void qr_work_layout(int rows, int cols, void* work, long long* out6) {
    QrWork w = carve(work, rows, cols);
    unsigned char* base = reinterpret_cast<unsigned char*>(work);
    out6[0] = reinterpret_cast<unsigned char*>(w.V) - base;   out6[1] = (long long)rows * PW * sizeof(cplx);
    out6[2] = reinterpret_cast<unsigned char*>(w.Vt) - base;  out6[3] = (long long)PW * w.ldvt * sizeof(cplx);
    out6[4] = reinterpret_cast<unsigned char*>(w.T) - base;   out6[5] = (long long)PW * PW * sizeof(cplx);
}

// ------------------------------------------------------------------------------------------------------------------
// K = 128 block reflectors: two consecutive 64-column panels share ONE trailing update.
//   A2 <- (I - V2 T2^H V2^H)(I - V1 T1^H V1^H) A2  =  A2 - [V1 V2] [W1; W2]
//   Y = [V1 V2]^H A2 (one split-K product with m = 128),  W1 = T1^H Y1,  W2 = T2^H (Y2 - (V2^H V1) W1)
// so the rank-64 update (C read-modify-write bound, 27 TFLOP/s class) becomes a rank-128 one (40 TFLOP/s class) and half as
// many passes over the trailing matrix are made.  The look-ahead works on pairs: the pair's reflectors go to the next
// pair's 128 columns first, then the next pair is prepared (panel, narrow update, panel, Gram blocks) on the auxiliary
// stream underneath the big update.  The panel kernel and T are unchanged, so R and the solution agree with the 64-wide
// path to rounding.
namespace {

constexpr int PW2 = 2 * PW;

struct QrPairWork {
    cplx* V;      // rows x 128 (ld 128): [V1 | V2], V2 shifted down by 64 rows (its first 64 rows are zero)
    cplx* Vt;     // 128 x ldvt
    cplx* W;      // 128 x ldw
    cplx* part;   // split-K partials
    cplx* Gm;     // 64 x 64 scratch: V^H V of the panel being factored
    cplx* T[2];   // compact-WY factors of the two panels
    cplx* G12;    // V1^H V2
    cplx* tau;    // 2 x 64
    cplx* dots; cplx* currow; unsigned* bar; cplx* gslab;
    int ldvt, ldw;
};

size_t pair_part_elems(int cols, int naug) {
    const size_t ldw = (cols + naug + 7) / 8 * 8;
    const size_t a = size_t(149) * PW * PW, b = size_t(12) * PW2 * ldw;
    return a > b ? a : b;
}

QrPairWork carve_pair(void* work, int rows, int cols, int naug) {
    QrPairWork w;
    unsigned char* p = reinterpret_cast<unsigned char*>(work);
    auto take = [&](size_t bytes) { void* q = p; p += align_up(bytes); return q; };
    w.ldvt = (rows + 7) / 8 * 8;
    w.ldw = (cols + naug + 7) / 8 * 8;
    w.V = (cplx*)take(size_t(rows) * PW2 * sizeof(cplx));
    w.Vt = (cplx*)take(size_t(PW2) * w.ldvt * sizeof(cplx));
    w.W = (cplx*)take(size_t(PW2) * w.ldw * sizeof(cplx));
    w.part = (cplx*)take(pair_part_elems(cols, naug) * sizeof(cplx));
    w.Gm = (cplx*)take(size_t(PW) * PW * sizeof(cplx));
    w.T[0] = (cplx*)take(size_t(PW) * PW * sizeof(cplx));
    w.T[1] = (cplx*)take(size_t(PW) * PW * sizeof(cplx));
    w.G12 = (cplx*)take(size_t(PW) * PW * sizeof(cplx));
    w.tau = (cplx*)take(size_t(2) * PW * sizeof(cplx));
    w.dots = (cplx*)take(size_t(2) * MAXG * PW * sizeof(cplx));
    w.currow = (cplx*)take(size_t(2) * PW * sizeof(cplx));
    w.bar = (unsigned*)take(256);
    w.gslab = (cplx*)take(gslab_bytes(rows));
    return w;
}

int splits_for(int mp, int n2, int kk, size_t part_cap) {
    const int tiles = (n2 + 63) / 64 * (kk / 64);
    int ns = (592 + tiles - 1) / tiles;                 // ~2 waves of 296 resident CTAs
    const int cap = (mp + 255) / 256;                   // keep >= 256 k per slice
    if (ns > cap) ns = cap;
    const int fit = int(part_cap / (size_t(kk) * size_t(n2)));
    if (ns > fit) ns = fit;
    return ns < 1 ? 1 : ns;
}

// A2 (mp x n2) <- A2 - [V1 V2][W1; W2]  (kk = 128)  or the single-panel update with V1 alone (kk = 64)
int pair_apply(cudaStream_t st, const QrPairWork& w, int kk, int mp, cplx* A2, int lda, int n2, size_t part_cap) {
    if (n2 <= 0) return 0;
    const int ns = splits_for(mp, n2, kk, part_cap);
    int rc = zgemm_tn_splitk(st, true, kk, n2, mp, 1.0, w.V, PW2, A2, lda, 0.0, w.W, w.ldw, w.part, ns);            // Y = V^H A2
    if (rc) return rc;
    rc = zgemm_tn(st, true, PW, n2, PW, 1.0, w.T[0], PW, w.W, w.ldw, 0.0, w.W, w.ldw, GEMM_FULL);                   // W1 = T1^H Y1
    if (rc) return rc;
    if (kk == PW2) {
        cplx* Y2 = w.W + size_t(PW) * w.ldw;
        rc = zgemm_tn(st, true, PW, n2, PW, -1.0, w.G12, PW, w.W, w.ldw, 1.0, Y2, w.ldw, GEMM_FULL);                // Y2 -= (V2^H V1) W1
        if (rc) return rc;
        rc = zgemm_tn(st, true, PW, n2, PW, 1.0, w.T[1], PW, Y2, w.ldw, 0.0, Y2, w.ldw, GEMM_FULL);                 // W2 = T2^H Y2
        if (rc) return rc;
    }
    return zgemm_tn(st, false, mp, n2, kk, -1.0, w.Vt, w.ldvt, w.W, w.ldw, 1.0, A2, lda, GEMM_FULL);                // A2 -= V W
}

// Factor the panel pair whose top-left element is Ap (mp rows below and including it, `span` <= 128 pivot columns): returns the
// reflector width (64 or 128) in *kk.
int pair_prepare(cudaStream_t st, const QrPairWork& w, cplx* Ap, int lda, int mp, int span, size_t part_cap, double* logabs,
                 int* kk) {
    const int pwa = min(PW, span);
    // rows 0..63 of the second column block of V and columns 0..63 of rows 64..127 of V^T stay zero (V2 starts 64 rows lower)
    MSPDE_CUDA(cudaMemset2DAsync(w.V + PW, PW2 * sizeof(cplx), 0, PW * sizeof(cplx), PW, st));
    MSPDE_CUDA(cudaMemset2DAsync(w.Vt + size_t(PW) * w.ldvt, size_t(w.ldvt) * sizeof(cplx), 0, PW * sizeof(cplx), PW, st));
    int rc = launch_panel(st, Ap, lda, mp, pwa, w.V, PW2, w.Vt, w.ldvt, w.tau, w.dots, w.currow, w.bar, w.gslab, logabs);
    if (rc) return rc;
    int ns_g = (mp + 127) / 128;
    if (ns_g > 148) ns_g = 148;
    rc = zgemm_tn_splitk(st, true, PW, PW, mp, 1.0, w.V, PW2, w.V, PW2, 0.0, w.Gm, PW, w.part, ns_g);
    if (rc) return rc;
    larft_kernel<<<1, 256, 2 * sizeof(cplx) * PW * (PW + 1), st>>>(w.Gm, w.tau, pwa, w.T[0]);
    MSPDE_LAUNCH_CHECK();
    if (span <= PW || mp <= PW) { *kk = PW; return 0; }
    const int pwb = span - PW;
    rc = pair_apply(st, w, PW, mp, Ap + PW, lda, pwb, part_cap);             // reflector 1 -> the second panel's columns
    if (rc) return rc;
    cplx* V2 = w.V + size_t(PW) * PW2 + PW;
    cplx* Vt2 = w.Vt + size_t(PW) * w.ldvt + PW;
    rc = launch_panel(st, Ap + size_t(PW) * lda + PW, lda, mp - PW, pwb, V2, PW2, Vt2, w.ldvt, w.tau + PW, w.dots, w.currow, w.bar,
                      w.gslab, logabs);
    if (rc) return rc;
    ns_g = (mp - PW + 127) / 128;
    if (ns_g > 148) ns_g = 148;
    rc = zgemm_tn_splitk(st, true, PW, PW, mp - PW, 1.0, V2, PW2, V2, PW2, 0.0, w.Gm, PW, w.part, ns_g);
    if (rc) return rc;
    larft_kernel<<<1, 256, 2 * sizeof(cplx) * PW * (PW + 1), st>>>(w.Gm, w.tau + PW, pwb, w.T[1]);
    MSPDE_LAUNCH_CHECK();
    ns_g = (mp + 127) / 128;
    if (ns_g > 148) ns_g = 148;
    rc = zgemm_tn_splitk(st, true, PW, PW, mp, 1.0, w.V, PW2, w.V + PW, PW2, 0.0, w.G12, PW, w.part, ns_g);     // G12 = V1^H V2
    if (rc) return rc;
    *kk = PW2;
    return 0;
}

}  // namespace

size_t qr_pair_work_bytes(int rows, int cols, int naug) {
    const size_t ldvt = (rows + 7) / 8 * 8, ldw = (cols + naug + 7) / 8 * 8;
    size_t b = 0;
    b += align_up(size_t(rows) * PW2 * sizeof(cplx));
    b += align_up(size_t(PW2) * ldvt * sizeof(cplx));
    b += align_up(size_t(PW2) * ldw * sizeof(cplx));
    b += align_up(pair_part_elems(cols, naug) * sizeof(cplx));
    b += 4 * align_up(size_t(PW) * PW * sizeof(cplx));
    b += align_up(size_t(2) * PW * sizeof(cplx));
    b += align_up(size_t(2) * MAXG * PW * sizeof(cplx));
    b += align_up(size_t(2) * PW * sizeof(cplx));
    b += 256;
    b += align_up(gslab_bytes(rows));
    return b + 4096;
}

// Pair-level pieces for the block-distributed driver (mspde/dist.py: DistQR over 128-column blocks): the owner of a block
// prepares the pair (both panels, their T factors, V1^H V2), broadcasts V, V^T, T1, T2, G12 (qr_pair_work_layout) and every
// rank applies it to its own columns.  The workspace is laid out for (wrows, wcols, wnaug) like the single-GPU path's.
int qr_pair_prepare_at(cudaStream_t st, void* work, int wrows, int wcols, int wnaug, cplx* Ap, int lda, int mp, int span,
                       double* logabs, int* kk_host) {
    if (init_attrs()) return -1;
    QrPairWork w = carve_pair(work, wrows, wcols, wnaug);
    return pair_prepare(st, w, Ap, lda, mp, span, pair_part_elems(wcols, wnaug), logabs, kk_host);
}
int qr_pair_apply_at(cudaStream_t st, void* work, int wrows, int wcols, int wnaug, int kk, int mp, cplx* A2, int lda, int n2) {
    QrPairWork w = carve_pair(work, wrows, wcols, wnaug);
    return pair_apply(st, w, kk, mp, A2, lda, n2, pair_part_elems(wcols, wnaug));
}
void qr_pair_work_layout(int wrows, int wcols, int wnaug, void* work, long long* out10) {
    QrPairWork w = carve_pair(work, wrows, wcols, wnaug);
    unsigned char* base = reinterpret_cast<unsigned char*>(work);
    out10[0] = reinterpret_cast<unsigned char*>(w.V) - base;    out10[1] = (long long)wrows * PW2 * sizeof(cplx);
    out10[2] = reinterpret_cast<unsigned char*>(w.Vt) - base;   out10[3] = (long long)PW2 * w.ldvt * sizeof(cplx);
    out10[4] = reinterpret_cast<unsigned char*>(w.T[0]) - base; out10[5] = (long long)PW * PW * sizeof(cplx);
    out10[6] = reinterpret_cast<unsigned char*>(w.T[1]) - base; out10[7] = (long long)PW * PW * sizeof(cplx);
    out10[8] = reinterpret_cast<unsigned char*>(w.G12) - base;  out10[9] = (long long)PW * PW * sizeof(cplx);
}

static int geqrf_aug_pairs(cudaStream_t st, cplx* A, int lda, int rows, int cols, int naug, double* logabs,
                           const QrLookAhead* la) {
    if (init_attrs()) return -1;
    const int ntot = cols + naug;
    QrPairWork W[2] = {carve_pair(la->pair[0], la->pair_rows, la->pair_cols, la->pair_naug),
                       carve_pair(la->pair[1], la->pair_rows, la->pair_cols, la->pair_naug)};
    const size_t cap = pair_part_elems(la->pair_cols, la->pair_naug);
    int kk = PW;
    int rc = pair_prepare(st, W[0], A, lda, rows, min(PW2, cols), cap, logabs, &kk);
    if (rc) return rc;
    int q = 0;
    for (int c0 = 0; c0 < cols; c0 += PW2, ++q) {
        const QrPairWork& cur = W[q & 1];
        const int span = min(PW2, cols - c0);
        const int mp = rows - c0;
        cplx* Arow = A + size_t(c0) * lda;
        const int next = c0 + span;
        if (next < cols) {
            const int nspan = min(PW2, cols - next);
            rc = pair_apply(st, cur, kk, mp, Arow + next, lda, nspan, cap);                 // next pair's columns first
            if (rc) return rc;
            MSPDE_CUDA(cudaEventRecord(la->ev_a, st));
            MSPDE_CUDA(cudaStreamWaitEvent(la->aux, la->ev_a, 0));
            int kk_next = PW;
            rc = pair_prepare(la->aux, W[(q + 1) & 1], A + size_t(next) * lda + next, lda, rows - next, nspan, cap, logabs, &kk_next);
            if (rc) return rc;
            MSPDE_CUDA(cudaEventRecord(la->ev_b, la->aux));
            rc = pair_apply(st, cur, kk, mp, Arow + next + nspan, lda, ntot - (next + nspan), cap);
            if (rc) return rc;
            MSPDE_CUDA(cudaStreamWaitEvent(st, la->ev_b, 0));
            kk = kk_next;
        } else {
            rc = pair_apply(st, cur, kk, mp, Arow + next, lda, ntot - next, cap);
            if (rc) return rc;
        }
    }
    return 0;
}

static std::atomic<int> g_qr_block{-1};    // -1: not chosen yet (environment MSPDE_QR_BLOCK = 64 | 128, default 128)
int qr_block_width() {
    if (g_qr_block < 0) {
        const char* e = getenv("MSPDE_QR_BLOCK");
        g_qr_block = (e && atoi(e) == PW) ? PW : 2 * PW;
    }
    return g_qr_block;
}
int set_qr_block_width(int k) {
    if (k != PW && k != 2 * PW) { set_error("QR block-reflector width must be 64 or 128"); return -1; }
    g_qr_block = k;
    return 0;
}

// Householder QR of the leading `cols` columns of Aaug (rows x (cols + naug)), trailing columns get Q^H applied.
// logabs (device, optional) accumulates sum log |r_jj|.
int geqrf_aug(cudaStream_t st, cplx* A, int lda, int rows, int cols, int naug, void* work, double* logabs,
              const QrLookAhead* la) {
    if (logabs) MSPDE_CUDA(cudaMemsetAsync(logabs, 0, sizeof(double), st));
    const int ntot = cols + naug;
    if (qr_block_width() == 2 * PW && la && la->pair[0] && la->pair[1] && cols > 2 * PW && rows <= la->pair_rows && cols <= la->pair_cols &&
        naug <= la->pair_naug)
        return geqrf_aug_pairs(st, A, lda, rows, cols, naug, logabs, la);
    if (la == nullptr || la->work2 == nullptr || cols <= PW) {
        for (int c0 = 0; c0 < cols; c0 += PW) {
            const int pw = min(PW, cols - c0);
            const int mp = rows - c0;
            cplx* Ap = A + size_t(c0) * lda + c0;
            int rc = qr_panel_factor(st, Ap, lda, mp, pw, rows, cols, work, logabs);
            if (rc) return rc;
            rc = qr_panel_apply(st, mp, rows, cols, work, Ap + pw, lda, ntot - (c0 + pw));
            if (rc) return rc;
        }
        return 0;
    }
    // One-panel look-ahead: reflector p goes to the next panel's 64 columns first; the latency-bound factorisation of panel
    // p+1 then runs on the auxiliary stream (into the other workspace) underneath the trailing update with reflector p.
    void* wk[2] = {work, la->work2};
    int rc = qr_panel_factor(st, A, lda, rows, min(PW, cols), rows, cols, wk[0], logabs);
    if (rc) return rc;
    int p = 0;
    for (int c0 = 0; c0 < cols; c0 += PW, ++p) {
        const int pw = min(PW, cols - c0);
        const int mp = rows - c0;
        cplx* Ap = A + size_t(c0) * lda + c0;
        const int next = c0 + pw;
        if (next < cols) {
            const int pwn = min(PW, cols - next);
            rc = qr_panel_apply(st, mp, rows, cols, wk[p & 1], Ap + pw, lda, pwn);
            if (rc) return rc;
            MSPDE_CUDA(cudaEventRecord(la->ev_a, st));
            MSPDE_CUDA(cudaStreamWaitEvent(la->aux, la->ev_a, 0));
            rc = qr_panel_factor(la->aux, A + size_t(next) * lda + next, lda, rows - next, pwn, rows, cols, wk[(p + 1) & 1], logabs);
            if (rc) return rc;
            MSPDE_CUDA(cudaEventRecord(la->ev_b, la->aux));
            rc = qr_panel_apply(st, mp, rows, cols, wk[p & 1], Ap + pw + pwn, lda, ntot - (next + pwn));
            if (rc) return rc;
            MSPDE_CUDA(cudaStreamWaitEvent(st, la->ev_b, 0));
        } else {
            rc = qr_panel_apply(st, mp, rows, cols, wk[p & 1], Ap + pw, lda, ntot - next);
            if (rc) return rc;
        }
    }
    return 0;
}

// x <- R^-1 d, R = upper triangle of A[0:n,0:n], d = column dcol of A
int trsv_upper_aug(cudaStream_t st, cplx* A, int lda, int n, int dcol, cplx* x) {
    if (init_attrs()) return -1;
    for (int b0 = ((n - 1) / PW) * PW; b0 >= 0; b0 -= PW) {
        const int nb = min(PW, n - b0);
        trsv_diag_kernel<<<1, 64, SMALL_SMEM, st>>>(A, lda, b0, nb, dcol, x);
        MSPDE_LAUNCH_CHECK();
        if (b0 > 0) {
            trsv_update_kernel<<<(b0 + 7) / 8, 256, 0, st>>>(A, lda, b0, nb, dcol, x);
            MSPDE_LAUNCH_CHECK();
        }
    }
    return 0;
}

int lstsq_qr_aug(cudaStream_t st, cplx* Aaug, int lda, int rows, int cols, void* work, cplx* x, double* logabs,
                 const QrLookAhead* la) {
    int rc = geqrf_aug(st, Aaug, lda, rows, cols, 1, work, logabs, la);
    if (rc) return rc;
    return trsv_upper_aug(st, Aaug, lda, cols, cols, x);
}

}  // namespace mspde
