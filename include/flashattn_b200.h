/*
 * flashattn_b200.h -- C ABI of the B200-native fused-attention libraries.
 *
 * Four shared libraries are built by compile_cuda.sh under minitorch/cuda_kernels/,
 * with the SAME file names and the SAME legacy symbols the reference loads through
 * ctypes (reference: minitorch/cuda_kernel_ops.py:26-29), so the reference's own
 * cuda_kernel_ops.py binds them unchanged:
 *
 *   flashattention_kernel.so   replaces src/flashattention_kernel.cu
 *   softmax_kernel.so          replaces src/softmax_kernel.cu
 *   layernorm_kernel.so        replaces src/layernorm_kernel.cu
 *   combine.so                 replaces src/combine.cu (map / zip / reduce / matmul plumbing)
 *
 * Everything is plain pointers and ints: no torch / numpy types cross this boundary.
 * Tensors are row-major; "host" pointers are ordinary CPU memory (numpy storage),
 * "dev" pointers are CUDA device memory.  Legacy symbols keep the reference's `void`
 * return type; they never exit() or throw -- the outcome is read with fa_last_status().
 */
#ifndef FLASHATTN_B200_H_
#define FLASHATTN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* fa_stream_t; /* == cudaStream_t; NULL = default stream */

/* ---- status (all four libraries) ------------------------------------------------- */
enum { FA_OK = 0, FA_ERR_INVALID = 1, FA_ERR_UNSUPPORTED = 2, FA_ERR_CUDA = 3 };
int fa_last_status(void);          /* status of the most recent call into this library */
const char* fa_last_error(void);   /* human-readable message for it ("" when FA_OK)     */

/* ---- device-memory / timing utilities (all four libraries) ----------------------- */
int fa_device_count(void);
int fa_set_device(int dev);
void* fa_malloc(size_t bytes);                 /* NULL on failure */
void* fa_malloc_host(size_t bytes);            /* pinned host memory */
void* fa_malloc_async(size_t bytes, fa_stream_t stream);   /* stream-ordered, from the cached default pool */
int fa_free_async(void* dptr, fa_stream_t stream);
int fa_memset_async(void* dptr, int byte, size_t bytes, fa_stream_t stream);
int fa_free(void* dptr);
int fa_free_host(void* hptr);
int fa_memset(void* dptr, int byte, size_t bytes);
int fa_h2d(void* dst_dev, const void* src_host, size_t bytes);
int fa_d2h(void* dst_host, const void* src_dev, size_t bytes);
int fa_sync(void);                             /* cudaDeviceSynchronize */
void* fa_event_create(void);
int fa_event_record(void* ev, fa_stream_t s);
float fa_event_elapsed_ms(void* ev_start, void* ev_stop); /* synchronises on ev_stop */
int fa_event_destroy(void* ev);
int fa_flush_l2(void);                         /* overwrite a >L2-sized scratch buffer */
unsigned long long fa_launch_count(void);      /* kernels launched by this library so far */

/* =====================================================================================
 * flashattention_kernel.so
 * ===================================================================================== */
enum { FA_DTYPE_F32 = 0, FA_DTYPE_BF16 = 1 };
/* Arithmetic used by the legacy host-pointer entry points (their ABI is fp32 only):
 *   FA_MODE_FP32: fp32 SIMT kernels, <=1e-5 of the composed reference (default);
 *   FA_MODE_BF16: inputs rounded to bf16 on device, tcgen05 tensor-core kernels
 *                 (head_dim 64/128), fp32 accumulation, <=2e-2 max-abs.
 * Also settable with env MINITORCH_FA_MODE=fp32|bf16 read at load time. */
enum { FA_MODE_FP32 = 0, FA_MODE_BF16 = 1 };
void fa_set_mode(int mode);
int fa_get_mode(void);
/* bf16 backward without atomics: the tensor-core backward sums dQ across KV tiles with fp32 add-reductions in varying
 * order (reproducible to rounding, not bitwise); on = 1 (or env MINITORCH_FA_DETERMINISTIC=1) runs the deterministic
 * CUDA-core backward for bf16 tensors instead (much slower).  Forward and fp32 mode are always bitwise reproducible. */
void fa_set_deterministic(int on);
/* The legacy entry points are transfer-bound, so they cut the (batch, head) units into chunks of about this many bytes of
 * fp32 per tensor and overlap H2D of chunk c+1, the kernels of chunk c and D2H of chunk c-1 on three streams (results are
 * independent of the chunking).  Default 16 MiB for direct copies / 64 MiB for staged ones, or env MINITORCH_FA_CHUNK_MB;
 * 0 restores the default.  Caller buffers that are NOT page-locked (numpy storage) are staged through a ring of pinned
 * slots by a persistent pool of host threads (at most 16, env MINITORCH_FA_COPY_THREADS).  In FA_MODE_BF16 the wire format
 * is bf16: the host threads narrow fp32 -> bf16 while staging and widen the results on the way back, halving the bytes
 * on PCIe in both directions.
 * Threading: the legacy entry points keep per-device state (streams, events, pinned rings) and must be called from one
 * host thread at a time per process, like the reference's (single-threaded Python over ctypes). */
void fa_set_legacy_chunk_bytes(size_t bytes);
/* A forward call keeps its device copies of Q, K, V, O, m, l alive so that the matching backward call (same host pointers
 * and shape, same sampled fingerprint of the contents) uploads only dO.  Bounded by this budget (default 4096 MiB, env
 * MINITORCH_FA_KEEP_FWD_MB; 0 disables): oldest entries are dropped first, a backward consumes its entry. */
void fa_set_keep_forward_mb(long long mb);
void fa_forward_cache_stats(unsigned long long* hits, unsigned long long* misses);
/* Times a pinned staging ring could not be allocated (the call that needed it failed with FA_ERR_CUDA and said so;
 * fa_release_staging() or a smaller MINITORCH_FA_CHUNK_MB frees / needs less page-locked memory). */
unsigned long long fa_staging_fallbacks(void);
/* Bytes the legacy entry points have handed to the DMA engine since the library was loaded, per direction.  In bf16 mode
 * a page-locked caller tensor is split between two routes -- fp32 image by DMA + cast on the device, or narrowed /
 * widened by the host threads with bf16 on the wire -- in the proportion that lets the link and the host cores finish
 * together (env MINITORCH_FA_HYBRID=0 disables the split, MINITORCH_FA_HYBRID_COST sets the host/DMA cost ratio). */
void fa_wire_bytes(unsigned long long* h2d, unsigned long long* d2h);
/* The split above, set explicitly: host_cost = time of the host route for one tensor relative to its direct DMA as fp32
 * (> 0: used as given; 0: MINITORCH_FA_HYBRID_COST or the built-in 0.8 scaled by 16 / staging threads; < 0: no split,
 * page-locked tensors always go direct).  Tensors smaller than min_tensor_bytes (< 0: the default 32 MiB) are never split. */
void fa_set_transfer_policy(double host_cost, long long min_tensor_bytes);
/* The split a legacy call with n_in uploaded / n_out downloaded large tensors would get at host_cost (pageable[i] != 0:
 * tensor i is pageable, i.e. staged throughout): quarters of each tensor on the staged route, 0..4.  Host logic only. */
int fa_plan_transfer_preview(int n_in, const int* in_pageable, int n_out, const int* out_pageable, double host_cost,
                             int* in_staged_q, int* out_staged_q);
/* Release the pinned staging rings and the cached forward tensors of the current device. */
int fa_release_staging(void);

/* Legacy ABI -- identical to the reference's
 *   src/flashattention_kernel.cu:259 launch_flashattention_forward
 *   src/flashattention_kernel.cu:352 launch_flashattention_backward
 *   src/flashattention_kernel.cu:694 launch_flashattention_forward_causal
 *   src/flashattention_kernel.cu:761 launch_flashattention_backward_causal
 * Q,K,V,O,dO,dQ,dK,dV: host fp32 (B,nh,N,d) contiguous; l,m: host fp32 (B,nh,N).
 * scale = 1/sqrt(d) applied inside; K is NOT transposed; m = row max of the scaled
 * scores, l = sum exp(s - m).  Outputs are fully overwritten. */
void launch_flashattention_forward(float* Q, float* K, float* V, float* O, float* l, float* m, int B, int nh,
                                   int N, int d);
void launch_flashattention_forward_causal(float* Q, float* K, float* V, float* O, float* l, float* m, int B,
                                          int nh, int N, int d);
void launch_flashattention_backward(float* Q, float* K, float* V, float* O, float* dQ, float* dK, float* dV,
                                    float* dO, float* l, float* m, int B, int nh, int N, int d);
void launch_flashattention_backward_causal(float* Q, float* K, float* V, float* O, float* dQ, float* dK,
                                           float* dV, float* dO, float* l, float* m, int B, int nh, int N,
                                           int d);
/* Same, plus key padding (additive (B,N) host mask, LightSeq semantics of
 * src/softmax_kernel.cu:27-33, may be NULL) -- the reference has no flash+padding entry. */
void launch_flashattention_forward_masked(float* Q, float* K, float* V, float* O, float* l, float* m,
                                          const float* key_mask, int causal, int B, int nh, int N, int d);
void launch_flashattention_backward_masked(float* Q, float* K, float* V, float* O, float* dQ, float* dK,
                                           float* dV, float* dO, float* l, float* m, const float* key_mask,
                                           int causal, int B, int nh, int N, int d);

/* Device-pointer API (no copies, asynchronous on `stream`).
 * dtype selects the element type of Q,K,V,O,dO,dQ,dK,dV (m,l are always fp32).
 * Element (b,h,n,x) of every 4-D tensor lives at  b*stride_b + h*stride_h + n*stride_n + x
 * (in elements); pass 0,0,0 for the contiguous (B,nh,N,d) layout.  (B,N,nh,d) storage --
 * what MultiHeadAttention.project_to_query_key_value produces before its permute,
 * minitorch/modules_transfomer.py:87-100 -- is stride_b=N*nh*d, stride_h=d, stride_n=nh*d.
 * kv_len: optional int32[B] (keys >= kv_len[b] are padding); key_mask: optional fp32 (B,N)
 * additive mask.  Returns FA_OK or an error code (message via fa_last_error()).
 * Workspaces (LSE / D vectors, the fp32 dQ accumulator) are allocated stream-ordered on `stream` from the device's
 * default memory pool, so calls in flight on different streams are independent. */
typedef struct {
  int B, H, N, d;
  int dtype;      /* FA_DTYPE_* */
  int causal;     /* 0/1 */
  long long stride_b, stride_h, stride_n; /* elements; all 0 => contiguous (B,H,N,d) */
  const int* kv_len;       /* device int32[B] or NULL */
  const float* key_mask;   /* device fp32 (B,N) additive or NULL */
} fa_attn_desc;

int fa_flash_fwd_dev(const fa_attn_desc* desc, const void* Q, const void* K, const void* V, void* O, float* m,
                     float* l, fa_stream_t stream);
int fa_flash_bwd_dev(const fa_attn_desc* desc, const void* Q, const void* K, const void* V, const void* O,
                     const void* dO, const float* m, const float* l, void* dQ, void* dK, void* dV,
                     fa_stream_t stream);
/* Decode shapes (additive; SURVEY.md 8(f)-3 -- the reference re-runs the whole prefix per generated token,
 * project/run_machine_translation.py:299-325, and has no cache): ONE query token per (batch, head) against a key / value
 * cache.  q, out: logical (B, H, d); k_cache, v_cache: logical (B, H, L_cap, d) of which the first L (or kv_len[b])
 * positions are valid; element strides as below (0 => contiguous).  Split-KV over the cache so batch-1 decoding still
 * fills the GPU; every visible cache byte is read once.  lse (optional, (B,H) fp32) receives log sum exp of the scaled
 * scores.  dtype FA_DTYPE_F32 (head_dim % 4 == 0, <= 256) or FA_DTYPE_BF16 (head_dim % 8 == 0, <= 512). */
typedef struct {
  int B, H, d;
  int L;          /* valid cache positions when kv_len is NULL */
  int L_cap;      /* capacity of the cache along the position axis (0 => L) */
  int dtype;      /* FA_DTYPE_* of q, caches and out */
  long long q_stride_b, q_stride_h;
  long long o_stride_b, o_stride_h;
  long long cache_stride_b, cache_stride_h, cache_stride_n;
  const int* kv_len;   /* device int32[B] or NULL */
} fa_decode_desc;
int fa_flash_decode_dev(const fa_decode_desc* desc, const void* q, const void* k_cache, const void* v_cache, void* out,
                        float* lse, fa_stream_t stream);
/* fp32 <-> bf16 element conversion on device (n elements). */
int fa_cast_f32_to_bf16_dev(const float* src, void* dst_bf16, size_t n, fa_stream_t stream);
int fa_cast_bf16_to_f32_dev(const void* src_bf16, float* dst, size_t n, fa_stream_t stream);
/* Algorithmic FLOPs of one call (4*B*H*N*Nk*d fwd, 10*... bwd; causal halves; kv_len host
 * array optional) -- the figure bench.py divides by the measured time. */
double fa_attn_flops(int B, int H, int N, int d, int causal, const int* kv_len_host, int backward);

/* =====================================================================================
 * softmax_kernel.so   (reference: src/softmax_kernel.cu:233, :345)
 * ===================================================================================== */
/* In-place masked row softmax of inp (B,nhead,from_len,to_len) fp32; attn_mask (B,to_len)
 * additive or NULL; mask_future masks j>i; denominator is sum+1e-8 like the reference. */
void launch_attn_softmax(float* inp, const float* attn_mask, int batch_size, int nhead, int from_len,
                         int to_len, bool mask_future, fa_stream_t stream);
/* In-place: out_grad <- soft_inp * (out_grad - sum_j out_grad*soft_inp), rows x softmax_len. */
void launch_attn_softmax_bw(float* out_grad, const float* soft_inp, int rows, int softmax_len,
                            fa_stream_t stream);
int fa_attn_softmax_dev(float* inp, const float* attn_mask, int batch_size, int nhead, int from_len, int to_len,
                        int mask_future, fa_stream_t stream);
int fa_attn_softmax_bw_dev(float* out_grad, const float* soft_inp, long long rows, int softmax_len,
                           fa_stream_t stream);

/* =====================================================================================
 * layernorm_kernel.so   (reference: src/layernorm_kernel.cu:101, :370)
 * ===================================================================================== */
/* vars receives var+1e-8 and the backward adds 1e-8 again, exactly like the reference
 * (src/layernorm_kernel.cu:70, :229, :310). */
void launch_layernorm(float* ln_res, float* vars, float* means, const float* inp, const float* scale,
                      const float* bias, int batch_size, int hidden_dim, fa_stream_t stream);
void launch_layernorm_bw(float* gamma_grad, float* betta_grad, float* inp_grad, const float* out_grad,
                         const float* inp, const float* gamma, const float* betta, const float* vars,
                         const float* means, int batch_size, int hidden_dim, fa_stream_t stream_1,
                         fa_stream_t stream_2);
int fa_layernorm_dev(float* ln_res, float* vars, float* means, const float* inp, const float* scale,
                     const float* bias, long long rows, int hidden_dim, fa_stream_t stream);
int fa_layernorm_bw_dev(float* gamma_grad, float* betta_grad, float* inp_grad, const float* out_grad,
                        const float* inp, const float* gamma, const float* betta, const float* vars,
                        const float* means, long long rows, int hidden_dim, fa_stream_t stream);

/* =====================================================================================
 * combine.so   (reference: src/combine.cu:315, :385, :443, :523 -- SURVEY.md 8(f)-1 plumbing row)
 * ===================================================================================== */
/* Strided / broadcasting elementwise map, zip, single-dimension reduce and batched matmul over host
 * fp32 storages described by int32 shape / stride arrays, with the reference's function ids
 * (minitorch/cuda_kernel_ops.py:33-52: 1 add, 2 mul, 3 id, 4 neg, 5 lt, 6 eq, 7 sigmoid, 8 relu,
 * 9 relu_back, 10 log, 11 log_back, 12 exp, 13 inv, 14 inv_back, 15 is_close, 16 max, 17 pow, 18 tanh).
 * tensorReduce takes reduce_value as double: that is what the reference's ctypes binding passes
 * (cuda_kernel_ops.py:217); its own C side declared float and therefore read garbage. */
void tensorMap(float* out, int* out_shape, int* out_strides, int out_size, float* in_storage, int* in_shape,
               int* in_strides, int in_size, int shape_size, int fn_id);
void tensorZip(float* out, int* out_shape, int* out_strides, int out_size, int out_shape_size, float* a_storage,
               int* a_shape, int* a_strides, int a_size, int a_shape_size, float* b_storage, int* b_shape,
               int* b_strides, int b_size, int b_shape_size, int fn_id);
void tensorReduce(float* out, int* out_shape, int* out_strides, int out_size, float* a_storage, int* a_shape,
                  int* a_strides, int reduce_dim, double reduce_value, int shape_size, int fn_id);
void MatrixMultiply(float* out, int* out_shape, int* out_strides, float* a_storage, int* a_shape, int* a_strides,
                    float* b_storage, int* b_shape, int* b_strides, int batch, int m, int p);
/* Device-pointer variants (additive): the same kernels on operands already resident in HBM, asynchronous on
 * `stream`; shape / stride arrays stay host int32.  Operands of lower rank broadcast right-aligned, as in
 * tensorZip.  fa_matmul_dev: 3-D (batch | 1, m, n) @ (batch | 1, n, p) -> (batch, m, p), arbitrary strides. */
int fa_map_dev(float* out, const int* out_shape, const int* out_strides, int out_nd, const float* in,
               const int* in_shape, const int* in_strides, int in_nd, int fn_id, fa_stream_t stream);
int fa_zip_dev(float* out, const int* out_shape, const int* out_strides, int out_nd, const float* a,
               const int* a_shape, const int* a_strides, int a_nd, const float* b, const int* b_shape,
               const int* b_strides, int b_nd, int fn_id, fa_stream_t stream);
int fa_reduce_dev(float* out, const int* out_shape, const int* out_strides, const float* a, const int* a_shape,
                  const int* a_strides, int nd, int reduce_dim, double reduce_value, int fn_id, fa_stream_t stream);
int fa_matmul_dev(float* out, const int* out_shape, const int* out_strides, const float* a, const int* a_shape,
                  const int* a_strides, const float* b, const int* b_shape, const int* b_strides, fa_stream_t stream);
/* bf16 tensor-core GEMM for the Linear layers around the attention core (additive; the reference's
 * MatrixMultiplyKernel, src/combine.cu:148-210, is one fp32 thread per output element):
 *   C[M,N] (fp32, or bf16 when out_bf16) = A[M,K] . B[K,N], bf16 operands, fp32 accumulation on tcgen05.
 *   a_mn = 0: A stored [M][lda], K contiguous;   a_mn = 1: A stored [K][lda], M contiguous (x^T of a row-major x);
 *   b_mn = 1: B stored [K][ldb], N contiguous;   b_mn = 0: B stored [N][ldb], K contiguous (W^T of a row-major W).
 * So y = x.W, dx = dy.W^T and dW = x^T.dy all run on the stored tensors without a transpose copy.  Pointers 16-byte
 * aligned, lda / ldb multiples of 8 elements (FA_ERR_UNSUPPORTED otherwise: use fa_matmul_dev).
 * fa_qkv_proj_bf16_dev: x (M,E) times the concatenated weight (E,3E) in ONE GEMM; the three column blocks land in
 * separate (M,E) buffers = q, k, v in the (B,N,nh,d) layout fa_flash_fwd_dev consumes in place (E % 32 == 0). */
int fa_gemm_bf16_dev(void* out, int out_bf16, long long ldo, const void* a_bf16, int a_mn, long long lda,
                     const void* b_bf16, int b_mn, long long ldb, int M, int N, int K, fa_stream_t stream);
int fa_qkv_proj_bf16_dev(void* q, void* k, void* v, int out_bf16, const void* x_bf16, const void* wqkv_bf16, int M, int E,
                         fa_stream_t stream);
/* Embedding lookup and softmax cross-entropy without one-hot matmuls (additive; the reference
 * builds (tokens, vocab) one-hot matrices, minitorch/modules_basic.py:55-71 and nn.py:251-271, and only declares
 * fused kernels, src/includes/kernels.h:196-215).  ids / targets are fp32 tensors holding integers.
 *   fa_embedding_fw_dev:    out (n,E) = W (V,E) rows selected by ids (n)          == one_hot(ids) @ W
 *   fa_embedding_bw_dev:    dW (V,E)  = sum of dout rows per id, ascending order  == one_hot(ids)^T @ dout, deterministic
 *   fa_softmax_xent_fw_dev: loss (n) = logsumexp(logits (n,C)) - logits[i, t_i]; lse (n) saved for the backward;
 *                           logsumexp = max + log(sum + 1e-6), minitorch's log (operators.py:107-110)
 *   fa_softmax_xent_bw_dev: dlogits = dloss[i] * (exp(logits - lse[i]) - [j == t_i]) */
/* host-pointer variants (staged through the device pool, synchronous, status via fa_last_status) */
void launch_embedding_fw(float* out, const float* ids, const float* W, long long n, int V, int E);
void launch_embedding_bw(float* dW, const float* ids, const float* dout, long long n, int V, int E);
void launch_softmax_xent_fw(float* loss, float* lse, const float* logits, const float* targets, long long n, int C);
void launch_softmax_xent_bw(float* dlogits, const float* dloss, const float* logits, const float* targets,
                            const float* lse, long long n, int C);
int fa_embedding_fw_dev(float* out, const float* ids, const float* W, long long n, int V, int E, fa_stream_t stream);
int fa_embedding_bw_dev(float* dW, const float* ids, const float* dout, long long n, int V, int E, fa_stream_t stream);
int fa_softmax_xent_fw_dev(float* loss, float* lse, const float* logits, const float* targets, long long n, int C,
                           fa_stream_t stream);
int fa_softmax_xent_bw_dev(float* dlogits, const float* dloss, const float* logits, const float* targets,
                           const float* lse, long long n, int C, fa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FLASHATTN_B200_H_ */
