// extern "C" boundary of libmspde.so -- declarations and reference citations live in include/mspde.h.
#include "common.cuh"
#include <cstdarg>
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
}  // namespace mspde

using namespace mspde;

extern "C" {

const char* mspde_last_error(void) { return get_error(); }

int mspde_zgemm_tn(void* stream, int conjA, int m, int n, int k, double alpha, const void* A, int lda, const void* B,
                   int ldb, double beta, void* C, int ldc, int uplo, void* splitk_ws, int nsplit) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nsplit > 1)
        return zgemm_tn_splitk(st, conjA != 0, m, n, k, alpha, (const cplx*)A, lda, (const cplx*)B, ldb, beta, (cplx*)C,
                               ldc, (cplx*)splitk_ws, nsplit);
    return zgemm_tn(st, conjA != 0, m, n, k, alpha, (const cplx*)A, lda, (const cplx*)B, ldb, beta, (cplx*)C, ldc, uplo);
}

}  // extern "C"
