// Complex FP64 GEMM on the FP64 tensor pipe (DMMA m8n8k4) for sm_100a.
//
//   C[i][j] = alpha * sum_k opA(A[k][i]) * B[k][j] + beta * C[i][j]       opA = identity | conj
//
// A is K x m, B is K x n, C is m x n, all row-major interleaved complex128.  This single "TN" form is
// what every dense step of the pCN path reduces to (L^H L, U12^H U12 trailing updates, U^-H B solves,
// V^H A block reflectors): both operands are read along contiguous rows, so tiles are fetched with
// 1-D bulk copies (cp.async.bulk.shared::cluster.global -- the non-tensor form of the TMA unit, SASS UBLKCP: one row
// segment per copy, no tensor map, so it works on arbitrary sub-blocks; completion by mbarrier complete_tx) into padded,
// bank-conflict-free shared-memory rows.
//
// Two kernels, one per complex product scheme (mspde_set_gemm_mode / MSPDE_GEMM=3m|4m; DEFAULT: 3M, common.cuh):
//   zgemm3m_tn_kernel  three real DMMA products per complex product (6 flop per complex multiply-add), further down;
//   zgemm_tn_kernel    four real products (8 flop), directly below.
// Accuracy trade-off of the default: 3M keeps the normwise bound O(u) sum |a||b| of 4M but loses the componentwise bound on
// the imaginary part (Higham, Accuracy and Stability, 23.2.4), and the beta != 0 preload forms Im(s c) as s(cr + ci) - s cr,
// which cancels when |cr| >> |ci|.  No consumer on the path relies on the componentwise bound (Cholesky leaves read only the
// real part of the diagonal); the parity tests run under BOTH schemes at 1e-10 and bench.py reports both rates.
//
// 4M: a complex m x n x k product is run as the real
// product  C(m x 2n) = A(m x 2k) * B~(2k x 2n)  where A is used as stored (interleaved re/im along k)
// and B~ is the real 2x2 expansion [[br, bi], [-bi, br]] of B.  The expansion is never materialised: a
// lane of the k4 x n8 "col" fragment picks its component ((k'&1)^(n'&1)) and sign ((k'&1)&!(n'&1)) when
// it loads from shared memory.  One DMMA.8x8x4 therefore performs a complex 8 x 4 x 2 product with no
// wasted flops, and the accumulator fragment (row g, cols 2t,2t+1) is exactly one complex C element.
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "common.cuh"

namespace mspde {

int gemm_pre_min();

namespace {

constexpr int BM = 64;          // C tile rows (i)
constexpr int BN = 64;          // C tile cols (j), complex
constexpr int BK = 16;          // complex k per stage
constexpr int STAGES = 3;
constexpr int GROUP_I = 16;     // row tiles per rasterisation group
constexpr int PA = BM + 4;      // smem row pitch (complex); == 4 mod 8 -> conflict-free fragment loads
constexpr int PB = BN + 4;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int STAGE_DOUBLES = (BK * PA + BK * PB) * 2;
constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_DOUBLES * sizeof(double) + 2 * STAGES * sizeof(uint64_t);

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip(double x, int signmask) {
    return __hiloint2double(__double2hiint(x) ^ signmask, __double2loint(x));
}

struct GemmArgs {
    const cplx* A;
    const cplx* B;
    cplx* C;
    int m, n, k;
    int lda, ldb, ldc;
    double alpha, beta;
    int conjA;
    int uplo;
    int kslice;             // k per blockIdx.z slice
    long long cslice;       // C stride (complex elements) per slice
    const double* SA;       // pre-summed operand planes of the 3M "P" kernel: SA[k][i] = Ar +- Ai, SB[k][j] = Br + Bi
    const double* SB;
    int ldsa, ldsb;         // leading dimensions in doubles (even)
};

__global__ void __launch_bounds__(THREADS, 2) zgemm_tn_kernel(GemmArgs p) {
    // Grouped rasterisation: consecutive CTAs walk GROUP_I row tiles x all column tiles, so one resident wave touches
    // ~16 A panels + ~18 B panels (fits the 126 MB L2) instead of every A panel of the matrix.
    const int tiles_i = (p.m + BM - 1) / BM, tiles_j = (p.n + BN - 1) / BN;
    const int pid = blockIdx.x;
    const int per_group = GROUP_I * tiles_j;
    const int first_i = (pid / per_group) * GROUP_I;
    const int gsize = min(tiles_i - first_i, GROUP_I);
    const int i0 = (first_i + (pid % per_group) % gsize) * BM;
    const int j0 = ((pid % per_group) / gsize) * BN;
    if (p.uplo == GEMM_UPPER && j0 + BN - 1 < i0) return;   // tile entirely below the diagonal

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + size_t(STAGES) * STAGE_DOUBLES * sizeof(double));
    uint64_t* empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const int kbeg = blockIdx.z * p.kslice;
    const int kend = min(p.k, kbeg + p.kslice);
    const int klen = kend - kbeg;
    const int ktiles = (klen + BK - 1) / BK;
    cplx* __restrict__ C = p.C + size_t(blockIdx.z) * p.cslice;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int mrows = min(BM, p.m - i0);   // valid A columns (= C rows) in this tile
    const int ncols = min(BN, p.n - j0);

    if (warp == CONSUMER_WARPS) {
        // ---------------- producer warp: lanes 0-15 fetch A rows, lanes 16-31 fetch B rows ----------------
        const bool isA = lane < BK;
        const int r = isA ? lane : lane - BK;
        const cplx* src_base = isA ? (p.A + size_t(kbeg) * p.lda + i0) : (p.B + size_t(kbeg) * p.ldb + j0);
        const size_t ld = isA ? p.lda : p.ldb;
        const uint32_t row_bytes = uint32_t(isA ? mrows : ncols) * 16u;
        const int dst_off = isA ? r * PA * 2 : (BK * PA + r * PB) * 2;
        int stage = 0;
        uint32_t parity = 0;
        for (int kt = 0; kt < ktiles; ++kt) {
            mbar_wait(&empty_bar[stage], parity ^ 1);
            const int kvalid = min(BK, klen - kt * BK);
            if (lane == 0) mbar_expect_tx(&full_bar[stage], uint32_t(kvalid) * uint32_t(mrows + ncols) * 16u);
            __syncwarp();
            if (r < kvalid) {
                bulk_g2s(stage_base + size_t(stage) * STAGE_DOUBLES + dst_off, src_base + (size_t(kt) * BK + r) * ld,
                         row_bytes, &full_bar[stage]);
            }
            if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int g = lane >> 2;
    const int t = lane & 3;
    const int wi = warp & 1;          // 2 warps along i (32 rows each)
    const int wj = warp >> 1;         // 4 warps along j (16 complex cols each)
    const int a_sign = (p.conjA && (t & 1)) ? 0x80000000 : 0;
    const int b_sign = ((t & 1) && !(g & 1)) ? 0x80000000 : 0;
    const int a_off = ((t >> 1) * PA + wi * 32 + g) * 2 + (t & 1);
    const int b_off = (BK * PA + (t >> 1) * PB + wj * 16 + (g >> 1)) * 2 + ((t & 1) ^ (g & 1));

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    int stage = 0;
    uint32_t parity = 0;
    for (int kt = 0; kt < ktiles; ++kt) {
        mbar_wait(&full_bar[stage], parity);
        const double* S = stage_base + size_t(stage) * STAGE_DOUBLES;
        const int kvalid = klen - kt * BK;
        if (kvalid >= BK) {
#pragma unroll
            for (int kk = 0; kk < BK / 2; ++kk) {
                double af[4], bf[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) af[a] = flip(S[a_off + kk * 2 * PA * 2 + a * 16], a_sign);
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = flip(S[b_off + kk * 2 * PB * 2 + b * 8], b_sign);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
        } else {
            // k tail: rows >= kvalid of the stage were not fetched; feed zeros for them
            const int nk4 = (kvalid + 1) >> 1;
            for (int kk = 0; kk < nk4; ++kk) {
                const bool dead = (kk * 2 + (t >> 1)) >= kvalid;
                double af[4], bf[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) af[a] = dead ? 0.0 : flip(S[a_off + kk * 2 * PA * 2 + a * 16], a_sign);
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = dead ? 0.0 : flip(S[b_off + kk * 2 * PB * 2 + b * 8], b_sign);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; parity ^= 1; }
    }

    // ---------------- epilogue: one complex element per lane per fragment ----------------
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + wi * 32 + a * 8 + g;
        if (i >= p.m) continue;
        cplx* crow = C + size_t(i) * p.ldc;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + wj * 16 + b * 4 + t;
            if (j >= p.n) continue;
            if (p.uplo == GEMM_UPPER && j < i) continue;
            cplx out = make_double2(alpha * acc[a][b][0], alpha * acc[a][b][1]);
            if (beta != 0.0) {
                cplx c = crow[j];
                out.x = fma(beta, c.x, out.x);
                out.y = fma(beta, c.y, out.y);
            }
            crow[j] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 3M variant (the default scheme, common.cuh: GEMM_DEFAULT_MODE): three real products per complex product instead of four,
//     T1 = Ar^T Br,  T2 = Ai^T Bi,  T3 = (Ar +- Ai)^T (Br + Bi)
//     plain: C = (T1 - T2) + i (T3 - T1 - T2)        conj(A): C = (T1 + T2) + i (T3 - T1 + T2)   [T3 built with Ar - Ai]
// so the FP64 tensor pipe executes 6 instead of 8 real flops per complex multiply-add (cuBLAS calls the same scheme
// ZGEMM3M).  Normwise the rounding error stays O(u) * sum |a||b|; the imaginary part loses the componentwise bound of the
// 4M form (see the header of this file), which is why the 4M kernel stays selectable and is measured beside it.
// One DMMA.8x8x4 is a real 8 x 8 x 4 product here: a lane loads one complex A element (k = t, row g) and one complex B
// element (k = t, col g) with a 128-bit shared-memory load and feeds re, im and re+-im to the three accumulator sets.
// Three accumulator sets cost 1.5x the registers, so the CTA tile is 64 x 32 (8 MMA warps of 16 x 16), still 2 CTAs/SM.
template <int BN_, int STAGES_, int MF_>
struct M3 {
    static constexpr int BM = 64, BN = BN_, BK = 16, STAGES = STAGES_;
    static constexpr int PA = BM + 2;   // pitch == 2 mod 8 (complex): the 8 lanes of a 128-bit quarter-warp load hit 8 distinct 16 B slots
    static constexpr int PB = BN + 2;
    static constexpr int MF = MF_;                       // 8-row A fragments per warp: warp tile (8 MF) x 16
    static constexpr int WI = BM / (8 * MF);             // MMA warps along i
    static constexpr int WARPS = WI * (BN / 16);
    static constexpr int NTHREADS = (WARPS + 1) * 32;
    static constexpr int STAGE_DOUBLES = (BK * PA + BK * PB) * 2;
    static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_DOUBLES * sizeof(double) + 2 * STAGES * sizeof(uint64_t);
};
typedef M3<32, 4, 2> M3N;    // 8 MMA warps of 16 x 16: the general kernel
typedef M3<32, 4, 4> M3T;    // 4 MMA warps of 32 x 16 (168 registers): 6 instead of 8 operand sums per 24 DMMAs, +3 % when k is long enough to
                             // amortise the thinner latency cover of 8 MMA warps per SM (k >= 1024); 10-20 % slower on k <= 128 updates    // 64 x 32 tile, 288 threads, 2 CTAs/SM.  (A 64 x 64 tile with 16 MMA warps at 1 CTA/SM measured 2-25 % slower.)

template <class T>
__global__ void __launch_bounds__(T::NTHREADS, 2) zgemm3m_tn_kernel(GemmArgs p) {
    constexpr int BM = T::BM, BN = T::BN, BK = T::BK, STAGES = T::STAGES, PA = T::PA, PB = T::PB;
    constexpr int STAGE_DOUBLES = T::STAGE_DOUBLES;
    constexpr int CONSUMER_WARPS = T::WARPS;
    constexpr int MF = T::MF, WI = T::WI;
    const int tiles_i = (p.m + BM - 1) / BM, tiles_j = (p.n + BN - 1) / BN;
    const int pid = blockIdx.x;
    const int per_group = GROUP_I * tiles_j;
    const int first_i = (pid / per_group) * GROUP_I;
    const int gsize = min(tiles_i - first_i, GROUP_I);
    const int i0 = (first_i + (pid % per_group) % gsize) * BM;
    const int j0 = ((pid % per_group) / gsize) * BN;
    if (p.uplo == GEMM_UPPER && j0 + BN - 1 < i0) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + size_t(STAGES) * STAGE_DOUBLES * sizeof(double));
    uint64_t* empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const int kbeg = blockIdx.z * p.kslice;
    const int kend = min(p.k, kbeg + p.kslice);
    const int klen = kend - kbeg;
    const int ktiles = (klen + BK - 1) / BK;
    cplx* __restrict__ C = p.C + size_t(blockIdx.z) * p.cslice;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int mrows = min(BM, p.m - i0);
    const int ncols = min(BN, p.n - j0);

    if (warp == CONSUMER_WARPS) {
        const bool isA = lane < BK;
        const int r = isA ? lane : lane - BK;
        const cplx* src_base = isA ? (p.A + size_t(kbeg) * p.lda + i0) : (p.B + size_t(kbeg) * p.ldb + j0);
        const size_t ld = isA ? p.lda : p.ldb;
        const uint32_t row_bytes = uint32_t(isA ? mrows : ncols) * 16u;
        const int dst_off = isA ? r * PA * 2 : (BK * PA + r * PB) * 2;
        int stage = 0;
        uint32_t parity = 0;
        for (int kt = 0; kt < ktiles; ++kt) {
            mbar_wait(&empty_bar[stage], parity ^ 1);
            const int kvalid = min(BK, klen - kt * BK);
            if (lane == 0) mbar_expect_tx(&full_bar[stage], uint32_t(kvalid) * uint32_t(mrows + ncols) * 16u);
            __syncwarp();
            if (r < kvalid) {
                bulk_g2s(stage_base + size_t(stage) * STAGE_DOUBLES + dst_off, src_base + (size_t(kt) * BK + r) * ld,
                         row_bytes, &full_bar[stage]);
            }
            if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
        return;
    }

    const int g = lane >> 2;
    const int t = lane & 3;
    const int wi = warp % WI;         // WI warps along i (8 MF rows each)
    const int wj = warp / WI;         // BN/16 warps along j (16 complex cols each)
    const int a_off = (t * PA + wi * 8 * MF + g) * 2;                 // doubles; + kk*4*PA*2 per k4 step, + 16 per fragment
    const int b_off = (BK * PA + t * PB + wj * 16 + g) * 2;
    const bool conj = p.conjA != 0;
    const double csign = conj ? -1.0 : 1.0;      // A enters T3 as Ar + csign * Ai

    // beta != 0: the C tile is read up front into the accumulators (T1 <- s*cr, T3 <- s*(cr + ci), s = beta/alpha, which the
    // epilogue's T1 -+ T2 and T3 - T1 -+ T2 turn back into s*c + product), so its latency overlaps the first bulk copies
    const bool preload = p.beta != 0.0 && p.alpha != 0.0;
    const double cscale = preload ? p.beta / p.alpha : 0.0;
    double t1[MF][2][2], t2[MF][2][2], t3[MF][2][2];
#pragma unroll
    for (int a = 0; a < MF; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                t1[a][b][c] = t2[a][b][c] = t3[a][b][c] = 0.0;
                const int i = i0 + wi * 8 * MF + a * 8 + g, j = j0 + wj * 16 + b * 8 + 2 * t + c;
                if (preload && i < p.m && j < p.n) {
                    const cplx cv = C[size_t(i) * p.ldc + j];
                    t1[a][b][c] = cscale * cv.x;
                    t3[a][b][c] = cscale * (cv.x + cv.y);
                }
            }

    int stage = 0;
    uint32_t parity = 0;
    for (int kt = 0; kt < ktiles; ++kt) {
        mbar_wait(&full_bar[stage], parity);
        const double* S = stage_base + size_t(stage) * STAGE_DOUBLES;
        const int kvalid = klen - kt * BK;
        if (kvalid >= BK) {
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                double2 av[MF], bv[2];
                double as[MF], bs[2];
#pragma unroll
                for (int a = 0; a < MF; ++a) av[a] = *reinterpret_cast<const double2*>(S + a_off + kk * 4 * PA * 2 + a * 16);
#pragma unroll
                for (int b = 0; b < 2; ++b) bv[b] = *reinterpret_cast<const double2*>(S + b_off + kk * 4 * PB * 2 + b * 16);
#pragma unroll
                for (int a = 0; a < MF; ++a) as[a] = fma(csign, av[a].y, av[a].x);
                bs[0] = bv[0].x + bv[0].y;
                bs[1] = bv[1].x + bv[1].y;
#pragma unroll
                for (int a = 0; a < MF; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) dmma884(t1[a][b][0], t1[a][b][1], av[a].x, bv[b].x);
#pragma unroll
                for (int a = 0; a < MF; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) dmma884(t2[a][b][0], t2[a][b][1], av[a].y, bv[b].y);
#pragma unroll
                for (int a = 0; a < MF; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) dmma884(t3[a][b][0], t3[a][b][1], as[a], bs[b]);
            }
        } else {
            // k tail: rows >= kvalid of the stage were not fetched; feed zeros for them
            const int nk4 = (kvalid + 3) >> 2;
            for (int kk = 0; kk < nk4; ++kk) {
                const bool dead = (kk * 4 + t) >= kvalid;
                double2 av[MF], bv[2];
                double as[MF], bs[2];
#pragma unroll
                for (int a = 0; a < MF; ++a) {
                    av[a] = dead ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(S + a_off + kk * 4 * PA * 2 + a * 16);
                    as[a] = fma(csign, av[a].y, av[a].x);
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    bv[b] = dead ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(S + b_off + kk * 4 * PB * 2 + b * 16);
                    bs[b] = bv[b].x + bv[b].y;
                }
#pragma unroll
                for (int a = 0; a < MF; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        dmma884(t1[a][b][0], t1[a][b][1], av[a].x, bv[b].x);
                        dmma884(t2[a][b][0], t2[a][b][1], av[a].y, bv[b].y);
                        dmma884(t3[a][b][0], t3[a][b][1], as[a], bs[b]);
                    }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; parity ^= 1; }
    }

    // epilogue: the accumulator fragment (row g, cols 2t, 2t+1) is two adjacent complex C elements
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int a = 0; a < MF; ++a) {
        const int i = i0 + wi * 8 * MF + a * 8 + g;
        if (i >= p.m) continue;
        cplx* crow = C + size_t(i) * p.ldc;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int j = j0 + wj * 16 + b * 8 + 2 * t + c;
                if (j >= p.n) continue;
                if (p.uplo == GEMM_UPPER && j < i) continue;
                const double x1 = t1[a][b][c], x2 = t2[a][b][c], x3 = t3[a][b][c];
                const double re = conj ? x1 + x2 : x1 - x2;
                const double im = conj ? (x3 - x1) + x2 : (x3 - x1) - x2;
                cplx out = make_double2(alpha * re, alpha * im);
                if (beta != 0.0 && !preload) {
                    cplx cv = crow[j];
                    out.x = fma(beta, cv.x, out.x);
                    out.y = fma(beta, cv.y, out.y);
                }
                crow[j] = out;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 3M on PACKED operands ("P" kernel).  For large products the two operands are first repacked (pack_operand_kernel: one read
// of the operand, 24 B written per element) into the exact shared-memory image of the kernel's stage tiles:
//     tile(kt, ct) = [ 16 x 64 complex values | 16 x 64 sums ]        sums = Ar +- Ai on the A side, Br + Bi on the B side
// so that (1) the T3 operands are formed ONCE per matrix instead of once per CTA tile -- the DADD/DFMA that build them share
// the FP64 pipe with the DMMAs and cost ~4 DMMA issue slots each -- and (2) a stage is filled by 32 equal bulk copies of the
// contiguous tile images instead of row segments of the operand matrices.  (A first pre-summed variant with separate sum planes
// and 64 row copies per stage was 11-24 % SLOWER: ncu showed it stalled on long_scoreboard, i.e. on copy issue.)  Bank
// conflicts are removed by an XOR swizzle that the pack kernel bakes into the image (no padding): value column c of k-row r
// sits at c ^ 2(r&3), sum column c at c ^ 4(r&3); the fragment loads of a quarter-warp (128-bit) / half-warp (64-bit) then hit
// 8 / 16 distinct 16 B / 8 B slots.  Rows and columns beyond the matrix are zero in the image, so the kernel has no edge or
// k-tail path.  Same arithmetic in the same order as the in-loop kernel: results are bit-identical (tested).
//
// One CTA per SM, on purpose.  The natural configuration (64 x 32 tile, 36 KB stages, two CTAs per SM) was correct alone, with
// equal stream priorities, next to other kernels and next to itself -- but inside the pCN step, where Phi(L_current) runs on a
// LOW-priority stream, it produced wrong values whenever one of its CTAs shared an SM with a CTA of the in-loop GEMM kernel of
// the high-priority stream (the packed images and the kernel ordering were verified in place; padding its shared-memory
// request to 125 KB, which only rules that pairing out, made every run correct again).  The cause was not identified, so the
// shipped kernel excludes the pairing by construction: a 64 x 64 tile with 8 MMA warps and four 48 KB stages = 196 KB of shared
// memory, which leaves no room for a second GEMM CTA of any kind.  It is as fast as the 2-CTA variant (45.2 vs 45.0 ms on the
// dominant launch) and moves a third less data from L2 per flop.
struct M3P {
    // 64 x 64 tile, 8 MMA warps of 32 x 16, 4 stages of 48 KB: ONE CTA per SM by construction (196 KB of shared memory), so a P
    // CTA never shares an SM with another GEMM CTA (see the note on co-residency above launch()).
    static constexpr int BM = 64, BN = 64, BK = 16, STAGES = 4;
    static constexpr int MF = 4;
    static constexpr int WI = BM / (8 * MF);
    static constexpr int WARPS = WI * (BN / 16);
    static constexpr int NTHREADS = (WARPS + 1) * 32;
    static constexpr int A_TILE_DOUBLES = BK * BM * 3;    // 16 x 64 complex + 16 x 64 sums
    static constexpr int B_TILE_DOUBLES = BK * BN * 3;
    static constexpr int STAGE_DOUBLES = A_TILE_DOUBLES + B_TILE_DOUBLES;
    static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_DOUBLES * sizeof(double) + 2 * STAGES * sizeof(uint64_t);
};
static_assert(M3P::SMEM_BYTES + 1024 <= 232448 && 2 * M3P::SMEM_BYTES > 233472, "exactly one P-kernel CTA per SM");

// One CTA per tile: X is rows(k) x cols, row-major; tile (tk, tc) covers k-rows tk*16.. and columns tc*W..
template <int W>
__global__ void __launch_bounds__(256) pack_operand_kernel(const cplx* __restrict__ X, int ldx, int rows, int cols, double sign,
                                                           double* __restrict__ out, int tiles_c) {
    const int tc = blockIdx.x, tk = blockIdx.y;
    double* tile = out + (size_t(tk) * tiles_c + tc) * size_t(16 * W * 3);
    cplx* data = reinterpret_cast<cplx*>(tile);
    double* sums = tile + 16 * W * 2;
    for (int idx = threadIdx.x; idx < 16 * W; idx += 256) {
        const int r = idx / W, c = idx % W;
        const int kr = tk * 16 + r, gc = tc * W + c;
        cplx v = make_double2(0.0, 0.0);
        if (kr < rows && gc < cols) v = X[size_t(kr) * ldx + gc];
        const int t = r & 3;
        data[r * W + (c ^ (2 * t))] = v;
        sums[r * W + (c ^ (4 * t))] = fma(sign, v.y, v.x);
    }
}

__global__ void __launch_bounds__(M3P::NTHREADS, 1) zgemm3m_pre_tn_kernel(GemmArgs p) {
    typedef M3P T;
    constexpr int BM = T::BM, BN = T::BN, BK = T::BK, STAGES = T::STAGES;
    constexpr int STAGE_DOUBLES = T::STAGE_DOUBLES;
    constexpr int CONSUMER_WARPS = T::WARPS;
    constexpr int MF = T::MF, WI = T::WI;
    constexpr int A_SUM = BK * BM * 2;                    // doubles: start of the A sums inside a stage
    constexpr int B_DATA = T::A_TILE_DOUBLES;             // start of the B values
    constexpr int B_SUM = B_DATA + BK * BN * 2;
    const int tiles_i = (p.m + BM - 1) / BM, tiles_j = (p.n + BN - 1) / BN;
    const int pid = blockIdx.x;
    const int per_group = GROUP_I * tiles_j;
    const int first_i = (pid / per_group) * GROUP_I;
    const int gsize = min(tiles_i - first_i, GROUP_I);
    const int ti = first_i + (pid % per_group) % gsize;
    const int tj = (pid % per_group) / gsize;
    const int i0 = ti * BM, j0 = tj * BN;
    if (p.uplo == GEMM_UPPER && j0 + BN - 1 < i0) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + size_t(STAGES) * STAGE_DOUBLES * sizeof(double));
    uint64_t* empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const int kbeg = blockIdx.z * p.kslice;               // multiple of BK (the split-K driver rounds slices to BK)
    const int kend = min(p.k, kbeg + p.kslice);
    const int ktiles = (kend - kbeg + BK - 1) / BK;
    cplx* __restrict__ C = p.C + size_t(blockIdx.z) * p.cslice;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // producer warp: the packed tiles are the shared-memory image; lanes 0-15 copy the A tile and lanes 16-31 the B tile in
        // sixteen equal slices each (1536 B / 768 B)
        const bool isA = lane < 16;
        const int part = lane & 15;
        const uint32_t bytes = uint32_t((isA ? T::A_TILE_DOUBLES : T::B_TILE_DOUBLES) / 16 * sizeof(double));
        const double* src = isA ? p.SA + (size_t(kbeg / BK) * tiles_i + ti) * size_t(T::A_TILE_DOUBLES) + size_t(part) * (T::A_TILE_DOUBLES / 16)
                                : p.SB + (size_t(kbeg / BK) * tiles_j + tj) * size_t(T::B_TILE_DOUBLES) + size_t(part) * (T::B_TILE_DOUBLES / 16);
        const size_t step = isA ? size_t(tiles_i) * T::A_TILE_DOUBLES : size_t(tiles_j) * T::B_TILE_DOUBLES;
        const int dst = isA ? part * (T::A_TILE_DOUBLES / 16) : B_DATA + part * (T::B_TILE_DOUBLES / 16);
        int stage = 0;
        uint32_t parity = 0;
        for (int kt = 0; kt < ktiles; ++kt) {
            mbar_wait(&empty_bar[stage], parity ^ 1);
            if (lane == 0) mbar_expect_tx(&full_bar[stage], uint32_t(STAGE_DOUBLES * sizeof(double)));
            __syncwarp();
            bulk_g2s(stage_base + size_t(stage) * STAGE_DOUBLES + dst, src + size_t(kt) * step, bytes, &full_bar[stage]);
            if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
        return;
    }

    const int g = lane >> 2;
    const int t = lane & 3;
    const int wi = warp % WI;
    const int wj = warp / WI;
    // per-lane offsets inside a stage (doubles) for k-row t of a k4 group; + kk*4*W*{2,1} per group; the swizzles depend on t only
    const int a_off = (t * BM + wi * 8 * MF + (g ^ (2 * t))) * 2;            // + a*16 per fragment
    const int b_off = B_DATA + (t * BN + wj * 16 + (g ^ (2 * t))) * 2;       // + b*16
    int sa_off[MF], sb_off[2];
#pragma unroll
    for (int a = 0; a < MF; ++a) sa_off[a] = A_SUM + t * BM + wi * 8 * MF + ((a * 8 + g) ^ (4 * t));
#pragma unroll
    for (int b = 0; b < 2; ++b) sb_off[b] = B_SUM + t * BN + wj * 16 + ((b * 8 + g) ^ (4 * t));
    const bool conj = p.conjA != 0;

    const bool preload = p.beta != 0.0 && p.alpha != 0.0;
    const double cscale = preload ? p.beta / p.alpha : 0.0;
    double t1[MF][2][2], t2[MF][2][2], t3[MF][2][2];
#pragma unroll
    for (int a = 0; a < MF; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                t1[a][b][c] = t2[a][b][c] = t3[a][b][c] = 0.0;
                const int i = i0 + wi * 8 * MF + a * 8 + g, j = j0 + wj * 16 + b * 8 + 2 * t + c;
                if (preload && i < p.m && j < p.n) {
                    const cplx cv = C[size_t(i) * p.ldc + j];
                    t1[a][b][c] = cscale * cv.x;
                    t3[a][b][c] = cscale * (cv.x + cv.y);
                }
            }

    int stage = 0;
    uint32_t parity = 0;
    for (int kt = 0; kt < ktiles; ++kt) {
        mbar_wait(&full_bar[stage], parity);
        const double* S = stage_base + size_t(stage) * STAGE_DOUBLES;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double2 av[MF], bv[2];
            double as[MF], bs[2];
#pragma unroll
            for (int a = 0; a < MF; ++a) {
                av[a] = *reinterpret_cast<const double2*>(S + a_off + kk * 4 * BM * 2 + a * 16);
                as[a] = S[sa_off[a] + kk * 4 * BM];
            }
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                bv[b] = *reinterpret_cast<const double2*>(S + b_off + kk * 4 * BN * 2 + b * 16);
                bs[b] = S[sb_off[b] + kk * 4 * BN];
            }
#pragma unroll
            for (int a = 0; a < MF; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) dmma884(t1[a][b][0], t1[a][b][1], av[a].x, bv[b].x);
#pragma unroll
            for (int a = 0; a < MF; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) dmma884(t2[a][b][0], t2[a][b][1], av[a].y, bv[b].y);
#pragma unroll
            for (int a = 0; a < MF; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) dmma884(t3[a][b][0], t3[a][b][1], as[a], bs[b]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; parity ^= 1; }
    }

    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int a = 0; a < MF; ++a) {
        const int i = i0 + wi * 8 * MF + a * 8 + g;
        if (i >= p.m) continue;
        cplx* crow = C + size_t(i) * p.ldc;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int j = j0 + wj * 16 + b * 8 + 2 * t + c;
                if (j >= p.n) continue;
                if (p.uplo == GEMM_UPPER && j < i) continue;
                const double x1 = t1[a][b][c], x2 = t2[a][b][c], x3 = t3[a][b][c];
                const double re = conj ? x1 + x2 : x1 - x2;
                const double im = conj ? (x3 - x1) + x2 : (x3 - x1) - x2;
                cplx out = make_double2(alpha * re, alpha * im);
                if (beta != 0.0 && !preload) {
                    cplx cv = crow[j];
                    out.x = fma(beta, cv.x, out.x);
                    out.y = fma(beta, cv.y, out.y);
                }
                crow[j] = out;
            }
        }
    }
}

// Per-stream workspace of the packed operands, grown on demand.  A buffer that is too small is RETIRED, not freed: products
// enqueued earlier on the same stream may still be reading it.  New buffers are sized for the largest product any stream has
// asked for so far, so a stream grows at most two or three times.  A context hands its streams' buffers back when it is
// destroyed (gemm_release_stream); gemm_release_all frees everything, the retired buffers included, once the device is idle.
struct PreWs { double* p = nullptr; size_t cap = 0; };
std::mutex g_prews_mu;
std::unordered_map<cudaStream_t, PreWs> g_prews;
std::vector<double*> g_prews_retired;
size_t g_prews_max_need = 0;

double* prews_get(cudaStream_t st, size_t doubles) {
    std::lock_guard<std::mutex> lk(g_prews_mu);
    if (doubles > g_prews_max_need) g_prews_max_need = doubles;
    PreWs& w = g_prews[st];
    if (w.cap < doubles) {
        if (w.p) g_prews_retired.push_back(w.p);
        w.p = nullptr; w.cap = 0;
        size_t want = g_prews_max_need + g_prews_max_need / 16;
        if (cudaMalloc(&w.p, want * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            want = doubles;                                 // memory is tight: exactly what this product needs
            if (cudaMalloc(&w.p, want * sizeof(double)) != cudaSuccess) { cudaGetLastError(); w.p = nullptr; return nullptr; }
        }
        w.cap = want;
    }
    return w.p;
}

__global__ void splitk_reduce_kernel(const cplx* __restrict__ part, int nsplit, long long slice, int m, int n,
                                     double alpha, double beta, cplx* __restrict__ C, int ldc) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)m * n) return;
    int i = int(idx / n), j = int(idx % n);
    // fixed summation order (deterministic); 8 independent accumulators keep 8 loads in flight
    double ax[8], ay[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) ax[u] = ay[u] = 0.0;
    int s = 0;
    for (; s + 8 <= nsplit; s += 8) {
        cplx v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = part[(s + u) * slice + idx];
#pragma unroll
        for (int u = 0; u < 8; ++u) { ax[u] += v[u].x; ay[u] += v[u].y; }
    }
    for (; s < nsplit; ++s) {   // tail (fewer than 8 slices left), still a fixed order
        cplx v = part[s * slice + idx];
        ax[0] += v.x;
        ay[0] += v.y;
    }
    const double sx = ((ax[0] + ax[1]) + (ax[2] + ax[3])) + ((ax[4] + ax[5]) + (ax[6] + ax[7]));
    const double sy = ((ay[0] + ay[1]) + (ay[2] + ay[3])) + ((ay[4] + ay[5]) + (ay[6] + ay[7]));
    cplx* c = C + size_t(i) * ldc + j;
    cplx out = make_double2(alpha * sx, alpha * sy);
    if (beta != 0.0) {
        cplx old = *c;
        out.x = fma(beta, old.x, out.x);
        out.y = fma(beta, old.y, out.y);
    }
    *c = out;
}

DeviceOnce g_attr_set;

int launch(cudaStream_t st, const GemmArgs& a, int nz) {
    if (a.m <= 0 || a.n <= 0) return 0;
    const bool use3m = gemm_mode() == GEMM_3M;
    if (g_attr_set.first()) {
        MSPDE_CUDA(cudaFuncSetAttribute(zgemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        MSPDE_CUDA(cudaFuncSetAttribute(zgemm3m_tn_kernel<M3N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M3N::SMEM_BYTES));
        MSPDE_CUDA(cudaFuncSetAttribute(zgemm3m_tn_kernel<M3T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M3T::SMEM_BYTES));
        MSPDE_CUDA(cudaFuncSetAttribute(zgemm3m_pre_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M3P::SMEM_BYTES));
    }
    static FILE* logf = getenv("MSPDE_GEMM_LOG") ? fopen(getenv("MSPDE_GEMM_LOG"), "w") : nullptr;   // dev aid: shapes per launch
    if (logf) { fprintf(logf, "%d %d %d %d %d\n", a.m, a.n, a.k, a.uplo, nz); fflush(logf); }
    const int pre_min = gemm_pre_min();
    if (use3m && pre_min > 0 && min(a.m, a.n) >= pre_min && min(a.k, a.kslice) >= 512 && (nz == 1 || a.kslice % M3P::BK == 0)) {
        // large product: repack both operands into the kernel's stage-tile image (values + pre-summed T3 operands), then the P kernel
        GemmArgs b = a;
        const int tiles_k = (a.k + M3P::BK - 1) / M3P::BK;
        const int tiles_m = (a.m + M3P::BM - 1) / M3P::BM, tiles_n = (a.n + M3P::BN - 1) / M3P::BN;
        const size_t a_doubles = size_t(tiles_k) * tiles_m * M3P::A_TILE_DOUBLES;
        const size_t b_doubles = size_t(tiles_k) * tiles_n * M3P::B_TILE_DOUBLES;
        const size_t need = a_doubles + b_doubles;
        double* ws = need <= (size_t(6) << 27) ? prews_get(st, need) : nullptr;      // at most 6 GB per stream
        if (ws) {
            b.SA = ws; b.SB = ws + a_doubles;
            pack_operand_kernel<M3P::BM><<<dim3(tiles_m, tiles_k), 256, 0, st>>>(a.A, a.lda, a.k, a.m, a.conjA ? -1.0 : 1.0, ws, tiles_m);
            pack_operand_kernel<M3P::BN><<<dim3(tiles_n, tiles_k), 256, 0, st>>>(a.B, a.ldb, a.k, a.n, 1.0, ws + a_doubles, tiles_n);
            dim3 grid(tiles_m * tiles_n, 1, nz);
            zgemm3m_pre_tn_kernel<<<grid, M3P::NTHREADS, M3P::SMEM_BYTES, st>>>(b);
            MSPDE_LAUNCH_CHECK();
            return 0;
        }
    }
    if (use3m) {
        dim3 grid(((a.m + M3N::BM - 1) / M3N::BM) * ((a.n + M3N::BN - 1) / M3N::BN), 1, nz);
        static const int long_k = getenv("MSPDE_GEMM_LONG_K") ? atoi(getenv("MSPDE_GEMM_LONG_K")) : 1024;   // dev hook
        if (min(a.k, a.kslice) >= long_k)
            zgemm3m_tn_kernel<M3T><<<grid, M3T::NTHREADS, M3T::SMEM_BYTES, st>>>(a);
        else
            zgemm3m_tn_kernel<M3N><<<grid, M3N::NTHREADS, M3N::SMEM_BYTES, st>>>(a);
    } else {
        dim3 grid(((a.m + BM - 1) / BM) * ((a.n + BN - 1) / BN), 1, nz);
        zgemm_tn_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(a);
    }
    MSPDE_LAUNCH_CHECK();
    return 0;
}

std::atomic<int> g_gemm_mode{-1};   // -1: not chosen yet (environment MSPDE_GEMM = 3m | 4m, else the default)
std::atomic<int> g_pre_min{-1};     // -1: not chosen yet (environment MSPDE_GEMM_PRE_MIN, else 2048)

}  // namespace

// Smallest min(m, n) from which a 3M product with k >= 512 repacks its operands and runs the P kernel (0 = never).  Measured
// on B200 (profiles/r02_gemm_pre_v3.log; pack passes included): +6.3 % on the dominant launch (48.06 -> 45.20 ms = 0.929 of the
// FP64 tensor peak in executed flops), +5.2 % on L^H L, +3.1 % on a rank-512 update with m = 10 000, break-even near m = 3 000
// at k = 512, slower below; the step goes from 343.1 to 335.1 ms.
int gemm_pre_min() {
    if (g_pre_min < 0) {
        const char* e = getenv("MSPDE_GEMM_PRE_MIN");
        g_pre_min = e ? atoi(e) : 2048;
    }
    return g_pre_min;
}
int set_gemm_pre_min(int v) {
    g_pre_min = v < 0 ? 0 : v;
    return 0;
}

// Frees the packed-operand workspace a stream owns.  The caller is about to destroy the stream: its queued work is waited for.
void gemm_release_stream(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prews_mu);
    auto it = g_prews.find(st);
    if (it == g_prews.end()) return;
    cudaStreamSynchronize(st);
    if (it->second.p) cudaFree(it->second.p);
    g_prews.erase(it);
}

// Frees every packed-operand workspace of the process (live and retired) after waiting for the device; the next large
// product allocates afresh.  Returns the number of bytes given back.
size_t gemm_release_all() {
    std::lock_guard<std::mutex> lk(g_prews_mu);
    cudaDeviceSynchronize();
    size_t bytes = 0;
    for (auto& kv : g_prews)
        if (kv.second.p) { cudaFree(kv.second.p); bytes += kv.second.cap * sizeof(double); }
    g_prews.clear();
    for (double* p : g_prews_retired) cudaFree(p);
    g_prews_retired.clear();
    g_prews_max_need = 0;
    cudaGetLastError();
    return bytes;
}

int gemm_mode() {
    if (g_gemm_mode < 0) {
        const char* e = getenv("MSPDE_GEMM");
        g_gemm_mode = (e && (e[0] == '4')) ? GEMM_4M : (e && (e[0] == '3')) ? GEMM_3M : GEMM_DEFAULT_MODE;
    }
    return g_gemm_mode;
}
int set_gemm_mode(int mode) {
    if (mode != GEMM_4M && mode != GEMM_3M) { set_error("gemm mode must be 0 (4M) or 1 (3M)"); return -1; }
    g_gemm_mode = mode;
    return 0;
}

int zgemm_tn(cudaStream_t st, bool conjA, int m, int n, int k, double alpha, const cplx* A, int lda, const cplx* B,
             int ldb, double beta, cplx* C, int ldc, int uplo) {
    GemmArgs a;
    a.A = A; a.B = B; a.C = C;
    a.m = m; a.n = n; a.k = k;
    a.lda = lda; a.ldb = ldb; a.ldc = ldc;
    a.alpha = alpha; a.beta = beta;
    a.conjA = conjA ? 1 : 0;
    a.uplo = uplo;
    a.kslice = k > 0 ? k : 1;
    a.cslice = 0;
    a.SA = a.SB = nullptr; a.ldsa = a.ldsb = 0;
    return launch(st, a, 1);
}

int zgemm_tn_splitk(cudaStream_t st, bool conjA, int m, int n, int k, double alpha, const cplx* A, int lda,
                    const cplx* B, int ldb, double beta, cplx* C, int ldc, cplx* part, int nsplit) {
    if (nsplit <= 1) return zgemm_tn(st, conjA, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, GEMM_FULL);
    GemmArgs a;
    a.A = A; a.B = B; a.C = part;
    a.m = m; a.n = n; a.k = k;
    a.lda = lda; a.ldb = ldb; a.ldc = n;
    a.alpha = 1.0; a.beta = 0.0;
    a.conjA = conjA ? 1 : 0;
    a.uplo = GEMM_FULL;
    int ks = (k + nsplit - 1) / nsplit;
    ks = ((ks + BK - 1) / BK) * BK;
    a.kslice = ks;
    int nz = (k + ks - 1) / ks;
    a.cslice = (long long)m * n;
    a.SA = a.SB = nullptr; a.ldsa = a.ldsb = 0;
    int rc = launch(st, a, nz);
    if (rc) return rc;
    long long total = (long long)m * n;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part, nz, a.cslice, m, n, alpha, beta, C, ldc);
    MSPDE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mspde
