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
/* dst[0 .. bytes) = 0 on the stream (any alignment).  The step's own zero fills (the row-gradient buffer the masked rows are
 * scattered into, the attention delta accumulator, the loss scalar) — torch.Tensor.zero_() in the reference's autograd —
 * so that a training step launches no library kernel at all and its launches chain with programmatic dependent launch. */
int plb_fill_zero(void* dst, unsigned long long bytes, plb_stream_t stream);
/* Programmatic dependent launch for the kernels of one stream (default: enabled).  The reference's step is a chain of
 * ~1000 stream-ordered library kernels (R/train.py:353-357); here the ~260 kernels of a step are launched so that the
 * next kernel's CTAs take over an SM the moment the previous kernel's CTA leaves it (their prologue runs before
 * griddepcontrol.wait).  Results are identical either way.  Process-wide; returns PLB_OK. */
int plb_set_launch_overlap(int enable);

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
 *   PLB_EPI_GELU      u = bf16(acc (+ bias));  C = gelu_new(u) (HF/activations.py:66);  aux_out = gelu_new'(u) — the
 *                     derivative is what the backward pass needs of u, saved instead of u itself
 *   PLB_EPI_GELU_BWD  C = acc * aux_in[m,n]   (aux_in = the derivative saved by PLB_EPI_GELU); if aux_out != NULL it is an
 *                     fp32 [N] accumulator to which the column sums of the stored (rounded) C are ADDED — the bias
 *                     gradient of the Linear that produced u (autograd of HF/:239 ffn), fused so that no separate pass
 *                     reads d_u again
 *   PLB_EPI_ROWDOT    C = acc;  aux_out is FP32: aux_out[(n / 64) * ld_aux_out + m] += sum over the 64-column block of n
 *                     of bf16(C[m, n']) * aux_in[m, n'] — with C = dO (the O-projection DGRAD, HF:198-200) and aux_in = O this
 *                     is attention's delta = rowsum(dO o O) per (head, row), fused into the GEMM
 *                     that produces dO.  N %% 64 == 0; the caller zeroes aux_out first (two 32-column partials are added).
 */
#define PLB_GEMM_FWD 0
#define PLB_GEMM_DGRAD 1
#define PLB_GEMM_WGRAD 2

#define PLB_EPI_NONE 0
#define PLB_EPI_RESIDUAL 1
#define PLB_EPI_GELU 2
#define PLB_EPI_GELU_BWD 3
#define PLB_EPI_ROWDOT 4

int plb_gemm_bf16(int op, int epilogue, int M, int N, int K,
                  const void* A, int lda, const void* B, int ldb,
                  void* C, int ldc,
                  const float* bias,
                  const void* aux_in, int ld_aux_in,
                  void* aux_out, int ld_aux_out,
                  plb_stream_t stream);

/* ---- (e): weight gradient + its data-parallel all-reduce in ONE kernel -------------------------
 * R/train.py:160,218-221: DistributedDataParallel all-reduces every gradient bucket with NCCL after the
 * kernels that produced it.  For the large head (64 000 x 768 fp32 = 197 MB) this call IS the collective:
 * the same WGRAD as plb_gemm_bf16(PLB_GEMM_WGRAD, ...), but C_multicast is the NVSwitch MULTICAST address of
 * a buffer that every rank of the node has mapped (CUDA multicast object / torch symmetric memory), and the
 * epilogue accumulates with multimem.red.add.v4.f32 — the switch adds each tile into ALL ranks' copies, so
 * when every rank's kernel has completed each copy holds the SUM over ranks (no NCCL kernel, no second pass
 * over the gradient).  The caller orders the ranks (no rank may start adding before every rank has consumed
 * and zeroed the previous step's sum; readers wait for every rank's kernel).  N % 4 == 0, ldc % 4 == 0,
 * C_multicast 16-byte aligned. */
int plb_gemm_wgrad_allreduce(int M, int N, int K,
                             const void* A, int lda, const void* B, int ldb,
                             float* C_multicast, int ldc,
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

/* ---- a3: embeddings -------------------------------------------------------------------------------
 * HF/:71-110 AlbertEmbeddings.forward:  e = LayerNorm(word[ids[t]] + type[0] + pos_table[pos[t]])
 * ids/pos: int32 [T] (pos restarts at 0 per sequence, HF/:85-86).  Tables are fp32 [*, E]; `type0` points at
 * row 0 of token_type_embeddings.  out: bf16 [T, E]; stats: fp32 [T, 2] = (mean, rstd) saved for backward.
 * Backward: `de` = gradient of the LN output (bf16 [T, E]); accumulates (+=, fp32 atomics) into dword [V, E]
 * (row pad_id skipped: padding_idx, HF/:56; V = rows of the word table), dpos, dtype0 [E], dln_w, dln_b. */
int plb_embed_ln_fwd(const int32_t* ids, const int32_t* pos, int T, int E,
                     const float* word, const float* posemb, const float* type0,
                     const float* ln_w, const float* ln_b, float eps,
                     void* out, float* stats, plb_stream_t stream);
int plb_embed_ln_bwd(const void* de, const int32_t* ids, const int32_t* pos, int T, int E, int V,
                     const float* word, const float* posemb, const float* type0,
                     const float* ln_w, const float* stats, int pad_id,
                     float* dword, float* dpos, float* dtype0, float* dln_w, float* dln_b,
                     plb_stream_t stream);

/* ---- a5/a6: LayerNorm over bf16 rows ----------------------------------------------------------------
 * HF/:202 attention.LayerNorm and HF/:235 full_layer_layer_norm (the residual add is fused into the producing
 * GEMM's epilogue, so `z` is already x + sublayer(x)).  y = (z - mean) * rstd * w + b, biased variance,
 * eps inside the sqrt.  stats: fp32 [T, 2] = (mean, rstd).
 * Backward: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w; dw_accum += sum dy * xhat;
 * db_accum += sum dy; if colsum_accum != NULL, colsum_accum += sum_rows dz (the bias gradient of the Linear
 * that produced z). */
int plb_ln_fwd(const void* z, int T, int H, const float* w, const float* b, float eps,
               void* y, float* stats, plb_stream_t stream);
int plb_ln_bwd(const void* dy, const void* z, const float* stats, const float* w, int T, int H,
               void* dz, float* dw_accum, float* db_accum, float* colsum_accum, plb_stream_t stream);

/* ---- north star (c): dropout --------------------------------------------------------------------------
 * HF/:109 (after the embedding LayerNorm) and HF/:201 (attention output projection, before the residual add):
 * nn.Dropout(hidden_dropout_prob).  y = (residual ? residual : 0) + x * m, m = 0 or 1 / (1 - p) from a counter-based
 * generator keyed by (seed, stream, element index r * cols + c) (csrc/rng.cuh; numpy replica oracle/dropout_rng.py),
 * so the SAME call is the backward of the op and no mask tensor exists.  bf16 [rows, cols], cols % 8 == 0, in place
 * allowed.  The reference itself runs with p = 0. */
int plb_dropout(const void* x, int ld_x, const void* residual, int ld_res, void* y, int ld_y, int rows, int cols,
                float p, unsigned long long seed, unsigned int stream, plb_stream_t stream_);

/* ---- a5: fused attention with key padding -----------------------------------------------------------
 * HF/:184-197 (+ HF/integrations/sdpa_attention.py:92, HF/masking_utils.py:495-520).  qkv: bf16 [T, 3*heads*64]
 * = q | k | v as written by the fused QKV GEMM.  Sequence b owns rows [seq_start[b], seq_start[b] + q_len[b]);
 * its first kv_len[b] rows are valid keys, the rest (padded layout) are masked exactly like the reference's
 * -inf key mask.  No [B,H,L,L] tensor exists.  ctx: bf16 [T, heads*64].  lse2: fp32 [heads, total_rows], the
 * row log-sum-exp in the log2 domain (saved for backward).  Backward writes dqkv (bf16, same layout as qkv;
 * rows of masked keys get exact zeros) and uses `delta` (fp32 [heads, total_rows]) as scratch. head_dim must be 64. */
/* tcgen05 / TMEM implementation (sequences up to 512 rows): K and V of a (sequence, head) are TMA-loaded once into
 * shared memory, S and O accumulate in tensor memory. */
int plb_attn_fwd_tc(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                    const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                    void* ctx, int ld_ctx, float* lse2, plb_stream_t stream);

/* Backward on tcgen05 (delta + dQ kernel + dK/dV kernel).  If dbias_qkv_accum != NULL (fp32 [3*heads*64]) the column
 * sums of dqkv — the bias gradient of the fused QKV projection — are accumulated into it by the same kernels.
 * ctx == NULL: delta[heads, total_rows] already holds rowsum(dO o O) (PLB_EPI_ROWDOT produced it) and is only read. */
int plb_attn_bwd_tc(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                    const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                    const void* ctx, const void* dctx, int ld_ctx, const float* lse2, float* delta,
                    void* dqkv, int ld_dqkv, float* dbias_qkv_accum, plb_stream_t stream);

/* Backward as ONE kernel (attn_bwd_fused.cu): every [128 keys x 128 queries] block is visited once — S^T, dP^T and the
 * exponentials are computed once instead of once per backward kernel; dK / dV accumulate in tensor memory, the per-key-tile
 * partials of dQ are summed in `workspace` (fp32, L2-resident, plb_attn_bwd_fused_workspace_bytes(B, heads) bytes,
 * caller-owned, contents irrelevant before and after the call).  Same arguments and results as plb_attn_bwd_tc. */
int plb_attn_bwd_fused_workspace_bytes(int B, int heads);
int plb_attn_bwd_fused(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                       const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                       const void* ctx, const void* dctx, int ld_ctx, const float* lse2, float* delta,
                       void* dqkv, int ld_dqkv, float* dbias_qkv_accum, void* workspace, long long workspace_bytes,
                       plb_stream_t stream);

/* Attention-probability dropout (HF/:134, :194: dropout(softmax(S)) V, active in training when
 * attention_probs_dropout_prob > 0; the reference runs with 0).  Same arguments as plb_attn_fwd_tc / plb_attn_bwd_fused plus
 * (p, seed, rng_stream): the mask of probability (sequence b, head h, query q, key k) comes from the counter-based generator
 * of csrc/rng.cuh (replica: oracle/dropout_rng.attention_mask, item = b * heads + h), the backward call regenerates it
 * from the same triple.  At most 32768 (sequence, head) items.  The forward's ctx is computed from the dropped probabilities,
 * lse2 from the undropped scores. */
int plb_attn_fwd_tc_dropout(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                            const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                            void* ctx, int ld_ctx, float* lse2, float p_drop, unsigned long long seed,
                            unsigned int rng_stream, plb_stream_t stream);
int plb_attn_bwd_fused_dropout(const void* qkv, int ld_qkv, const int32_t* seq_start, const int32_t* q_len,
                               const int32_t* kv_len, int B, int heads, int head_dim, int max_q_len, int total_rows,
                               const void* ctx, const void* dctx, int ld_ctx, const float* lse2, float* delta,
                               void* dqkv, int ld_dqkv, float* dbias_qkv_accum, void* workspace, long long workspace_bytes,
                               float p_drop, unsigned long long seed, unsigned int rng_stream, plb_stream_t stream);

/* ---- a8: masked-row loss ----------------------------------------------------------------------------
 * R/train.py:107-131 calculate_phoneme_loss.  The host flattens masked_indices into rows[M] (= b*L + i, in the
 * reference's order) and weights[M] = 1 / (count * n_b); only those rows go through the heads.
 *   plb_gather_rows       dst[i, :] = src[rows[i], :]                    (bf16, bit-exact)
 *   plb_scatter_add_rows  dst[rows[i], :] += src[i, :]                   (bf16, atomics: duplicates add)
 *   plb_masked_ce         row_loss[i] = w_i * (logsumexp(z_i) - z_i[y_i]); loss_out (+)= sum_i row_loss[i]
 *                         (fixed-order sum); logits are OVERWRITTEN with dz = w_i * (softmax(z_i) - onehot(y_i)),
 *                         columns [V, ld) zeroed.  M == 0 gives loss 0 (R/train.py:129).
 *   plb_colsum            out_accum[n] += sum_t x[t, n]                  (bias gradients)             */
int plb_gather_rows(const void* src, int ld_src, const int32_t* rows, int M, int H,
                    void* dst, int ld_dst, plb_stream_t stream);
int plb_scatter_add_rows(const void* src, int ld_src, const int32_t* rows, int M, int H,
                         void* dst, int ld_dst, plb_stream_t stream);
int plb_masked_ce(void* logits, int ld, int M, int V, const int32_t* labels, const float* weights,
                  float* row_loss, float* loss_out, int accumulate, plb_stream_t stream);
int plb_colsum(const void* x, int ld, int T, int N, float* out_accum, plb_stream_t stream);

/* ---- a10: fused AdamW --------------------------------------------------------------------------------
 * torch.optim.AdamW as built at R/train.py:272 (rule: torch/optim/adam.py:417-419,457,476,531-547) over flat
 * fp32 arenas of n elements (n % 4 == 0).  `step` is the 1-based step count of these parameters; gradients are
 * multiplied by grad_scale first (1/world_size after a SUM all-reduce).  Hyper-parameters are host doubles (torch keeps
 * them as python floats and rounds derived scalars once).  If bf16_shadow != NULL the updated
 * parameters are also written as bf16 (the copy the GEMMs read); if zero_grad != 0 the gradient arena is
 * zeroed for the next step's in-place accumulation. */
int plb_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* bf16_shadow,
                   long n, int step, double lr, double beta1, double beta2, double eps, double weight_decay,
                   double grad_scale, int zero_grad, plb_stream_t stream);
int plb_cast_bf16(const float* src, void* dst, long n, plb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PLBERT_B200_H_ */
