// Shared declarations for libmspde (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

typedef double2 cplx;   // complex128, interleaved (re, im) -- same bits as numpy/torch complex128

namespace mspde {

void set_error(const char* fmt, ...);
const char* get_error();
extern std::atomic<long long> g_launches;   // kernels launched by this library (bench.py reports it as gpu_launches)

#define MSPDE_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            mspde::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                             cudaGetErrorString(e__));                                        \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

#define MSPDE_LAUNCH_CHECK()                                                                  \
    do {                                                                                      \
        ++mspde::g_launches;                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) {                                                             \
            mspde::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,         \
                             cudaGetErrorString(e__));                                        \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {   // conj(a) * b
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {   // a*b + c
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
    return c;
}
__device__ __forceinline__ cplx cfmac(cplx a, cplx b, cplx c) {  // conj(a)*b + c
    c.x = fma(a.x, b.x, c.x); c.x = fma(a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(-a.y, b.x, c.y);
    return c;
}
__device__ __forceinline__ cplx cfnma(cplx a, cplx b, cplx c) {  // c - a*b
    c.x = fma(-a.x, b.x, c.x); c.x = fma(a.y, b.y, c.x);
    c.y = fma(-a.x, b.y, c.y); c.y = fma(-a.y, b.x, c.y);
    return c;
}
__device__ __forceinline__ cplx cfnmac(cplx a, cplx b, cplx c) { // c - conj(a)*b
    c.x = fma(-a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(-a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
    return c;
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
// Smith-free complex division; operands are O(1)-scaled in every caller.
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: a process that opens contexts on several GPUs has to repeat
// it on each.  DeviceOnce::first() is true the first time it is asked on the current device.
struct DeviceOnce {
    bool seen[64] = {};
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
        if (seen[d]) return false;
        seen[d] = true;
        return true;
    }
};

// ---- GEMM (zgemm_tn.cu) --------------------------------------------------------------------------
enum { GEMM_FULL = 0, GEMM_UPPER = 1 };
// Complex product scheme of every GEMM launch: 4M = four real products (8 flop per complex multiply-add), 3M = three real
// products (6 flop, the ZGEMM3M scheme).  Process-wide; chosen by mspde_set_gemm_mode() or MSPDE_GEMM=3m|4m.
enum { GEMM_4M = 0, GEMM_3M = 1, GEMM_DEFAULT_MODE = GEMM_3M };
int gemm_mode();
int set_gemm_mode(int mode);
int gemm_pre_min();
int set_gemm_pre_min(int v);
void gemm_release_stream(cudaStream_t st);   // packed-operand workspace of one stream (before the stream is destroyed)
size_t gemm_release_all();                    // every packed-operand workspace; waits for the device
// C[i][j] = alpha * sum_k opA(A[k][i]) * B[k][j] + beta * C[i][j];  A: k x m, B: k x n, C: m x n, row-major.
int zgemm_tn(cudaStream_t st, bool conjA, int m, int n, int k, double alpha, const cplx* A, int lda,
             const cplx* B, int ldb, double beta, cplx* C, int ldc, int uplo);
// Split-K variant: partial products into `part` (nsplit slices of m x n, ld n), then a deterministic reduce.
int zgemm_tn_splitk(cudaStream_t st, bool conjA, int m, int n, int k, double alpha, const cplx* A, int lda,
                    const cplx* B, int ldb, double beta, cplx* C, int ldc, cplx* part, int nsplit);

// tomography.cu
int tomography_H(cudaStream_t st, int M, int N, const double* r, const double* theta, const int* ix, const int* iy, int n_r,
                 int n_theta, int apply_fixes, cplx* H, int ldh);

}  // namespace mspde
