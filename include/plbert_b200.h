/* plbert_b200 — C ABI of the B200-native (sm_100a) PL-BERT training hot path.
 *
 * The reference (Fadi987/PL-BERT) has no FFI of its own: its hot path is Python calling
 * PyTorch/HuggingFace library kernels.  This header is the boundary a
 * maintainer binds instead of those library calls; every entry point names the reference
 * code (R/ = /root/reference, HF/ = transformers 5.5 models/albert/modeling_albert.py) whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     unless the comment says "host";
 *   - kernels are only ENQUEUED on `stream` (a cudaStream_t passed as void*); nothing here
 *     allocates, synchronises or frees device memory;
 *   - return value: PLB_OK (0) or a negative PLB_ERR_* code; the message is available from
 *     plb_last_error_string() (thread-local).  Nothing throws across this boundary;
 *   - "bf16" is __nv_bfloat16 (2 bytes); row-major matrices carry an explicit leading
 *     dimension in ELEMENTS; bf16 matrices read or written by the tensor-core GEMM must
 *     have 16-byte aligned base pointers and leading dimensions that are multiples of 8.
 */
#ifndef PLBERT_B200_H_
#define PLBERT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLB_ABI_VERSION 1

#define PLB_OK 0
#define PLB_ERR_ARG (-1)         /* bad argument (shape, alignment, null pointer)            */
#define PLB_ERR_CUDA (-2)        /* a CUDA runtime / driver call failed                      */
#define PLB_ERR_UNSUPPORTED (-3) /* valid request this build has no kernel for               */

typedef void* plb_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------ */
int plb_abi_version(void);
const char* plb_last_error_string(void);
/* Number of SMs of the current device (148 on B200); <0 on error. */
int plb_device_sm_count(void);

/* ---- a1: key-padding mask --------------------------------------------------------------------
 * R/train.py:34-44 length_to_mask + R/train.py:386 `(~mask).int()`:
 *   attn_mask[b, j] = ((j + 1) > lengths[b]) ? 0 : 1            (int32, bit-exact)            */
int plb_length_to_mask(const int32_t* lengths, int B, int L, int32_t* attn_mask, plb_stream_t stream);

/* ---- GEMM (tcgen05 / TMEM / TMA) -------------------------------------------------------------
 * Replaces the cuBLAS calls behind nn.Linear forward/backward in HF/:180-182,198-200,239-241,276
 * and R/model.py:15-16 (heads).  bf16 operands, fp32 accumulation in tensor memory.
 *
 *   PLB_GEMM_FWD   : C[M,N]  = A[M,K] * W[N,K]^T        (y = x W^T, nn.Linear weight layout)
 *   PLB_GEMM_DGRAD : C[M,N]  = A[M,K] * W[K,N]          (dx = dy W)
 *   PLB_GEMM_WGRAD : C[M,N] += A[K,M]^T * B[K,N]        (dW += dy^T x; C is FP32, accumulated
 *                                                         with atomics so the 12 applications of
 *                                                         the shared ALBERT layer sum in place)
 * Epilogues (FWD/DGRAD only; C and aux are bf16):
 *   PLB_EPI_NONE      C = acc (+ bias[n] if bias != NULL)
 *   PLB_EPI_RESIDUAL  C = acc (+ bias) + aux_in[m,n]
 *   PLB_EPI_GELU      aux_out = u = acc (+ bias);  C = gelu_new(u)        (HF/activations.py:66)
 *   PLB_EPI_GELU_BWD  C = acc * gelu_new'(aux_in[m,n])
 */
#define PLB_GEMM_FWD 0
#define PLB_GEMM_DGRAD 1
#define PLB_GEMM_WGRAD 2

#define PLB_EPI_NONE 0
#define PLB_EPI_RESIDUAL 1
#define PLB_EPI_GELU 2
#define PLB_EPI_GELU_BWD 3

int plb_gemm_bf16(int op, int epilogue, int M, int N, int K,
                  const void* A, int lda, const void* B, int ldb,
                  void* C, int ldc,
                  const float* bias,
                  const void* aux_in, int ld_aux_in,
                  void* aux_out, int ld_aux_out,
                  plb_stream_t stream);

/* Plain CUDA-core GEMM with the same signature and semantics as plb_gemm_bf16 (fp32 math,
 * one thread per output element).  TEST INFRASTRUCTURE: used by tests/ to cross-check the
 * tensor-core kernel on the device; never called by the product path. */
int plb_gemm_bf16_check(int op, int epilogue, int M, int N, int K,
                        const void* A, int lda, const void* B, int ldb,
                        void* C, int ldc,
                        const float* bias,
                        const void* aux_in, int ld_aux_in,
                        void* aux_out, int ld_aux_out,
                        plb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PLBERT_B200_H_ */
