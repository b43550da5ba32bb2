/* libmspde.so -- C ABI of the B200-native pCN hot path (placeholder header; filled in as entry points land). */
#ifndef MSPDE_H
#define MSPDE_H
#ifdef __cplusplus
extern "C" {
#endif
const char* mspde_last_error(void);
int mspde_zgemm_tn(void* stream, int conjA, int m, int n, int k, double alpha, const void* A, int lda, const void* B,
                   int ldb, double beta, void* C, int ldc, int uplo, void* splitk_ws, int nsplit);
#ifdef __cplusplus
}
#endif
#endif
