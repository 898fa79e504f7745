/* mg_b200.h — C ABI of libmg_b200.so (work in progress header; see bottom of file for the model API). */
#ifndef MG_B200_H
#define MG_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* mg_last_error(void);

/* c[M,N] = act(a[M,K] * b[N,K]^T + bias[N]) + residual[M,N]; fp32 device pointers, row-major.
 * planes: 2 = split-bf16 (near-fp32), 1 = plain bf16. swap_out: write c transposed ([N,M] row-major). */
int mg_op_gemm(void* stream, int M, int N, int K, const float* a, const float* b, float* c, const float* bias,
               const float* residual, int act, int planes, int block_n, int ksplit, int swap_out);

#ifdef __cplusplus
}
#endif
#endif
