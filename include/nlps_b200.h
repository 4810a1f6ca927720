/*
 * nlps_b200.h -- C ABI of libnlps_b200.so: the B200 (sm_100a) training step of the
 * NLPScratch next-word Transformer.
 *
 * The reference (hydrogendeuteride/NLPScratch) is pure Python and has no FFI; its
 * boundary for this path is the duck-typed model object that
 * utils/train.py::train_transformer drives (utils/train.py:104-140).  Every entry point
 * below names the reference method (file:line under the reference tree) whose arithmetic
 * it replaces.  nlpscratch_b200/transformer.py binds them with ctypes and re-exposes the
 * reference's own method names, so the reference's loop runs unmodified on top.
 *
 * Conventions
 *   - plain C types only; every pointer named d_* is DEVICE memory, h_* is HOST memory.
 *   - every function returns 0 on success, non-zero on failure; nlps_last_error() returns
 *     a thread-local message.  There is NO CPU fallback: without a CUDA device every
 *     compute entry point fails with a message.
 *   - a model handle owns one CUDA stream of work order (nlps_set_stream); calls on one
 *     handle are not thread-safe; launches are asynchronous unless stated otherwise.
 *   - float tensors are fp32 row-major; token ids are int32; PAD id is 0.
 */
#ifndef NLPS_B200_H_
#define NLPS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLPS_ABI_VERSION 1

/* arithmetic mode of the dense contractions */
enum {
  NLPS_PREC_FP32 = 0,   /* CUDA-core FP32 FMA everywhere: the <=1e-4 "FP32-faithful" mode      */
  NLPS_PREC_TF32 = 1,   /* tcgen05 kind::tf32 tensor-core GEMMs, FP32 accumulate in TMEM       */
  NLPS_PREC_TF32X3 = 2  /* 3xTF32 split (hi*hi + lo*hi + hi*lo) on tcgen05: FP32-level error   */
};

/* parameter tensor ids (Transformer.__init__, Transformer/transformer.py:34-63) */
enum {
  NLPS_P_We = 0, NLPS_P_Wq, NLPS_P_Wk, NLPS_P_Wv, NLPS_P_Wo, NLPS_P_W1, NLPS_P_W2, NLPS_P_b1,
  NLPS_P_b2, NLPS_P_W_vocab, NLPS_P_b_vocab, NLPS_P_ln1_gamma, NLPS_P_ln1_beta,
  NLPS_P_ln2_gamma, NLPS_P_ln2_beta, NLPS_P_COUNT
};
/* which flat buffer a view points into */
enum { NLPS_BUF_PARAM = 0, NLPS_BUF_GRAD = 1, NLPS_BUF_ADAM_M = 2, NLPS_BUF_ADAM_V = 3 };

typedef struct nlps_model nlps_model;   /* opaque */
typedef struct nlps_cache nlps_cache;   /* opaque: activations of ONE forward call */

/* Transformer(vocab_size, embed_dim, num_heads, ff_dim, num_layers, max_len, ..., dropout_p,
 * use_adam) -- transformer.py:9-21 */
typedef struct nlps_config {
  int32_t vocab_size, embed_dim, num_heads, ff_dim, num_layers, max_len;
  float dropout_p;
  int32_t use_adam;
  int32_t precision;     /* NLPS_PREC_* */
  uint64_t seed;         /* dropout Philox key */
} nlps_config;

/* strided view of one parameter tensor (or its grad / Adam moment) inside a flat buffer.
 * Wq/Wk/Wv are column blocks of one packed (E,3E) matrix per layer, hence the strides. */
typedef struct nlps_view {
  void* d_ptr;
  int32_t ndim;
  int64_t shape[3];
  int64_t stride[3];     /* in elements */
} nlps_view;

/* ---- library --------------------------------------------------------------------- */
int nlps_abi_version(void);
const char* nlps_last_error(void);
int nlps_device_count(int* count);                 /* 0 devices is not an error here        */
int nlps_device_info(int device, char* name_out, size_t name_cap, int* sm_count, int* cc_major,
                     int* cc_minor, size_t* total_mem);

/* ---- raw device memory for the host-side array shim (replaces the numpy/cupy array
 *      module the reference picks at transformer.py:11-12) ------------------------------- */
int nlps_malloc(void** d_ptr, size_t bytes);
int nlps_free(void* d_ptr);
int nlps_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);   /* async if pinned */
int nlps_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);   /* synchronises    */
int nlps_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream);
int nlps_memset(void* d_dst, int value, size_t bytes, void* stream);
int nlps_stream_sync(void* stream);
/* read-back without stalling the GPU: asynchronous copy into PINNED host memory, ordered by an event the host waits on
 * later (a training loop reads step i's loss while step i+1 runs; utils/train.py:138 reads it at once) */
int nlps_memcpy_d2h_async(void* h_pinned, const void* d_src, size_t bytes, void* stream);
int nlps_event_create(void** event);
int nlps_event_record(void* event, void* stream);
int nlps_event_synchronize(void* event);
int nlps_event_destroy(void* event);
int nlps_model_stream(nlps_model* m, void** stream);              /* the cudaStream_t the model's work is ordered on */
int nlps_host_alloc_pinned(void** h_ptr, size_t bytes);
int nlps_host_free_pinned(void* h_ptr);

/* ---- model lifetime ------------------------------------------------------------------ */
int nlps_create(const nlps_config* cfg, int device, nlps_model** out);
int nlps_destroy(nlps_model* m);
int nlps_set_stream(nlps_model* m, void* cuda_stream);       /* cudaStream_t; NULL = default */
int nlps_set_precision(nlps_model* m, int precision);
int nlps_set_training(nlps_model* m, int training);          /* train()/eval(), transformer.py:96-100 */
int nlps_set_dropout(nlps_model* m, float p, uint64_t seed);
int nlps_sync(nlps_model* m);

/* flat buffers (params, grads, Adam m, Adam v): [We | layer 0 ... layer L-1 | W_vocab b_vocab], every tensor 256-B
 * aligned.  The gradient BUCKETS (nlps_bucket_range) enumerate contiguous ranges of that buffer in
 * backward-completion order -- [vocab] [layer L-1] ... [layer 0] [We] -- so that data-parallel reductions can start
 * as backward finishes them. */
int nlps_flat_size(const nlps_model* m, int64_t* n_floats);
int nlps_flat_ptr(nlps_model* m, int which_buf, void** d_ptr);
int nlps_param_view(nlps_model* m, int which_buf, int param_id, nlps_view* out);
int nlps_pe_ptr(nlps_model* m, void** d_ptr);                /* (max_len,E), transformer.py:609-615 */
int nlps_num_buckets(const nlps_model* m, int* n);
int nlps_bucket_range(const nlps_model* m, int bucket, int64_t* first_float, int64_t* n_floats);
int nlps_stream_wait_bucket(nlps_model* m, int bucket, void* other_stream);  /* grads of bucket final */
int nlps_wait_stream(nlps_model* m, void* other_stream);     /* model stream waits for other_stream */

/* ---- the step --------------------------------------------------------------------------- */
/* Transformer.forward, transformer.py:110-199.  d_x: (B,S) int32 with row stride x_row_stride
 * (the driver passes the slice xb[:, :s_max], utils/train.py:110).  hidden_only!=0 returns the
 * (B,S,E) final hidden state (return_hidden_only=True), else (B,S,V) logits with row stride
 * ld_out >= V.  d_out may be NULL to keep the result only inside the cache. */
int nlps_cache_create(nlps_model* m, int B, int S, nlps_cache** out);
int nlps_cache_release(nlps_model* m, nlps_cache* c);
int nlps_forward(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t x_row_stride,
                 int hidden_only, float* d_out, int64_t ld_out);
int nlps_cache_hidden_ptr(nlps_cache* c, void** d_h_final);  /* (B*S,E), cache['H_final'] */
int nlps_cache_logits_ptr(nlps_cache* c, void** d_logits, int64_t* ld);
/* cache['layers'][layer][...] of the reference's cache dict (transformer.py:163-171): which = 0 'attn_input' (B*S,E),
 * 1 'ffn_input' (B*S,E), 2 'ffn_hidden' (B*S,F); plus 3 = merged heads of P@V (B*S,E) and 4 = packed Q|K|V rows (B*S,3E),
 * which this build stores instead of the (B,h,S,S) attention weights. */
int nlps_cache_layer_ptr(nlps_model* m, nlps_cache* c, int layer, int which, void** d_ptr);
int nlps_logits_ld(const nlps_model* m, int64_t* ld);        /* padded row stride of (N,V) buffers */

/* Transformer.calculate_loss_dlogits, transformer.py:515-549.  rows = B*S (or B for the
 * last-token path).  d_dlogits may alias d_logits.  grad_scale multiplies dlogits on top of the
 * reference's 1/rows (1/world_size under data parallelism, else 1).  d_loss: one float. */
int nlps_loss_dlogits(nlps_model* m, const float* d_logits, int64_t ld_logits, const int32_t* d_y,
                      int64_t rows, float grad_scale, float* d_loss, float* d_dlogits, int64_t ld_dlogits);

/* Transformer.backward, transformer.py:327-459.  Fills the flat grad buffer. */
int nlps_backward(nlps_model* m, nlps_cache* c, const float* d_dlogits, int64_t ld_dlogits);
/* Transformer.backward_last_token, transformer.py:201-325. */
int nlps_backward_last_token(nlps_model* m, nlps_cache* c, const int32_t* d_t_index,
                             const float* d_dlogits_last, int64_t ld_dlogits);

/* clip_grads, utils/function.py:27-38, over the whole flat grad buffer.  The scale factor stays
 * on the device; *h_norm_out (optional) forces a synchronising read-back of the norm. */
int nlps_clip(nlps_model* m, float max_norm, float* h_norm_out);
/* Transformer.step / _adam_update, transformer.py:461-509.  use_clip_scale!=0 multiplies the
 * gradients by the factor the last nlps_clip_compute left on the device (fused clip). */
int nlps_clip_compute(nlps_model* m, float max_norm);
int nlps_step(nlps_model* m, float lr, int use_clip_scale);
int nlps_adam_t(nlps_model* m, int64_t* t, int set, int64_t value);

/* Fused whole step for the full-sequence objective (transformer.py:635-641 with the clip of
 * utils/train.py:122-129): forward, loss, backward; then, unless defer_update!=0 (data parallel:
 * the caller all-reduces the flat grad buffer first and then calls nlps_clip_compute/nlps_step),
 * clip and step.  max_norm<=0 disables the clip.  d_loss: one float on the device. */
int nlps_train_step(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t x_row_stride,
                    const int32_t* d_y, float lr, float max_norm, float grad_scale,
                    int defer_update, float* d_loss);
/* Fused step of the reference driver's hot loop (utils/train.py:104-131): hidden-only forward on the
 * trimmed batch, H_last = H[arange(B), t], logits_last = H_last W_vocab + b_vocab, loss on B rows,
 * backward_last_token, clip, step -- no host synchronisation.  d_t: (B,) int32 = lens - 1. */
int nlps_train_step_last_token(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t x_row_stride,
                               const int32_t* d_t, const int32_t* d_y, float lr, float max_norm,
                               float grad_scale, int defer_update, float* d_loss);
/* The evaluation pass, the learning-rate schedule and the sampling block of the driver loop without per-batch host
 * synchronisation.  nlps_eval_last_token = one batch of compute_eval_loss (utils/train.py:51-75): eval-mode
 * hidden-only forward, last-real-token vocab head, per-row losses; adds their sum to d_accum[0] and the row count to
 * d_accum[1].  nlps_eval_finish closes a pass: d_sched = {previous eval loss, passes so far, learning rate, halved flag,
 * this eval loss}; eval = accum[0]/accum[1]; "if the loss went up, halve the learning rate" (utils/train.py:86-88) is taken
 * on the device; the accumulator is cleared.  The host reads d_sched once per pass (it prints the loss, like the reference).
 * nlps_topk_softmax = `probs = exp(l - max); probs /= sum; argsort()[-k:][::-1]` (utils/train.py:149-159) for one logits row:
 * the k most probable ids, most probable first, and their probabilities; d_work: V floats of scratch. */
int nlps_eval_last_token(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t x_row_stride, const int32_t* d_t,
                         const int32_t* d_y, float* d_accum);
int nlps_eval_finish(nlps_model* m, float* d_accum, float* d_sched);
int nlps_topk_softmax(void* stream, const float* d_logits, int V, int k, float* d_work, int32_t* d_idx, float* d_prob);
/* how the fused steps of this model ran so far: eagerly (first sighting of a signature, or not capturable), captured into a
 * CUDA graph (second sighting), replayed */
int nlps_graph_stats(nlps_model* m, int64_t* eager, int64_t* captured, int64_t* replayed);
int nlps_launch_count(nlps_model* m, int64_t* launches, int reset);  /* kernels launched so far (process-wide counter) */

/* ---- data-parallel exchange (the reference has no distributed code) -------------------------------------
 * One process per GPU, batch rows sharded; the only exchange of the step is the sum of the flat gradient buffer over
 * ranks.  The library owns NCCL (bound at run time with dlopen, so this file loads without it): rank 0 calls
 * nlps_comm_unique_id and ships the 128 bytes to the other ranks by any means (the Python host uses torch.distributed's
 * store); every rank calls nlps_comm_init.  From then on the UNDEFERRED fused steps (nlps_train_step,
 * nlps_train_step_last_token with defer_update = 0) are data-parallel steps: each rank passes grad_scale =
 * n_local / n_global, the gradient buckets are summed with ncclAllReduce on a communication stream as backward completes
 * them (buckets coalesced to >= min_bucket_floats), clip + update run on the summed gradients -- all inside the call and,
 * after its second sighting, inside ONE CUDA graph.  nlps_comm_allreduce_grads serves the unfused forward / backward API. */
int nlps_comm_available(int* nccl_version);                          /* 0 if libnccl could be loaded */
int nlps_comm_unique_id(void* h_id128);
int nlps_comm_init(nlps_model* m, const void* h_id128, int rank, int world, int64_t min_bucket_floats);
int nlps_comm_destroy(nlps_model* m);
int nlps_comm_info(const nlps_model* m, int* rank, int* world, int* n_exchanges);   /* world 0: no communicator */
int nlps_comm_allreduce_grads(nlps_model* m);
int nlps_comm_broadcast_params(nlps_model* m, int root);

/* Roofline denominator, measured: the tcgen05.mma kind::tf32 issue ceiling of this GPU with FP32 operands resident in
 * shared memory and accumulators in TMEM (the GEMM kernel's MMA loop with TMA, epilogue and global traffic removed;
 * csrc/tf32_peak.cu).  cta_group 1: M128 N256 per SM; 2: M256 N256 per SM pair.  iters groups of 16 MMAs per issuing unit,
 * best of `reps` launches, timed with CUDA events inside the call (synchronises the device). */
int nlps_tf32_mma_peak(int device, int cta_group, int iters, int reps, float* h_tflops, float* h_ms);

/* Dispatch statistics (process-wide): counts[k] = calls served by implementation k since the last reset, k =
 * 0 tcgen05 CTA-pair GEMM, 1 tcgen05 single-CTA GEMM, 2 CUDA-core FP32 GEMM, 3 3xTF32 GEMM, 4 tcgen05 attention forward,
 * 5 tcgen05 attention backward, 6 mma.sync attention, 7 CUDA-core FP32 attention, 8 first-generation tcgen05 attention.
 * The parity tests use it to assert that the benchmarked kernels -- not a fallback -- produced the checked numbers. */
int nlps_kernel_stats(int64_t* counts, int n, int reset);

/* Error word of the asynchronous kernels, read and cleared (synchronises the model's stream).  The reference raises
 * IndexError from NumPy's fancy indexing at once; kernels cannot, so they clamp the id, raise a bit, and the host turns
 * it into IndexError at its next synchronisation point (float(loss), synchronize(), get()):
 *   bit 0 (1): token id outside [-V, V) in We[x]                       (transformer.py:114)
 *   bit 1 (2): target id outside [-V, V) in log_probs[rows, y]         (transformer.py:538-546) */
int nlps_check_errors(nlps_model* m, int* flags);

/* Transformer._dropout, transformer.py:102-108.  The reference draws its keep-masks from NumPy's global generator; here a
 * mask is the pure function Philox4x32-10(key = seed; counter = element >> 3, layer*2 + branch, forward-call counter)
 * compared with round(keep * 2^16) per 16-bit field, evaluated in the forward epilogue and again in backward.
 * nlps_dropout_counter reads / sets the forward-call counter (the `step` of the next forward; part of a resume
 * checkpoint); nlps_dropout_mask writes the keep decisions (1.0 / 0.0) of n consecutive elements of the (B*S, E) tensor at
 * (layer, branch: 0 attention output, 1 FFN output) for forward call `step` -- the hook the parity tests use to replay the
 * masks inside the CPU oracle. */
int nlps_dropout_counter(nlps_model* m, int64_t* value, int set, int64_t new_value);
int nlps_dropout_mask(nlps_model* m, int layer, int which, int64_t step, int64_t n, float* d_out);

/* ---- stand-alone operators (tests, the array shim and the benchmarks call these) --------- */
/* C[M,N] = op(A)[M,K] * op(B)[K,N] (+bias[N]) (relu) (*gate>0) (+residual) ; trans flags say the
 * operand is stored transposed (A as [K,M], B as [N,K]).  precision = NLPS_PREC_*.  */
int nlps_gemm(void* stream, int precision, int M, int N, int K,
              const float* d_A, int64_t lda, int trans_a,
              const float* d_B, int64_t ldb, int trans_b,
              float* d_C, int64_t ldc,
              const float* d_bias, int relu,
              const float* d_gate, int64_t ldg,
              const float* d_residual, int64_t ldr,
              int accumulate);
/* causal + key-padding attention over packed (B*S,3E) Q|K|V rows (transformer.py:131-141),
 * fully-masked rows attend to all keys (additive -1e9 semantics, transformer.py:585-606). */
int nlps_attention_forward(void* stream, int precision, int B, int S, int h, int d,
                           const float* d_qkv, const int32_t* d_x, float* d_ctx, float* d_lse);
int nlps_attention_backward(void* stream, int precision, int B, int S, int h, int d,
                            const float* d_qkv, const int32_t* d_x, const float* d_ctx,
                            const float* d_lse, const float* d_dctx, float* d_dqkv);
/* layer_norm_affine_forward/backward, utils/function.py:90-124 (eps 1e-6 inside the sqrt). */
int nlps_layernorm_forward(void* stream, int64_t rows, int E, const float* d_x, const float* d_gamma,
                           const float* d_beta, float* d_y, float* d_mean, float* d_rstd);
int nlps_layernorm_backward(void* stream, int64_t rows, int E, const float* d_dy, const float* d_x,
                            const float* d_mean, const float* d_rstd, const float* d_gamma,
                            float* d_dx, float* d_dgamma, float* d_dbeta);
/* We[x] + pe[:S] (transformer.py:114) and its scatter-add gradient (transformer.py:453-457),
 * summed in ascending token position per vocabulary row: deterministic. */
int nlps_embed_forward(void* stream, int B, int S, int E, int V, const int32_t* d_x,
                       int64_t x_row_stride, const float* d_We, const float* d_pe, float* d_out,
                       int32_t* d_x_flat, int32_t* d_err_flag);
int nlps_embed_backward(void* stream, int64_t N, int E, int V, const int32_t* d_x_flat,
                        const float* d_dh, float* d_dWe);
int nlps_colsum(void* stream, int64_t rows, int cols, const float* d_x, int64_t ld, float* d_out);
int nlps_sumsq(void* stream, int64_t n, const float* d_x, float* d_out);

/* ---- element-wise / indexing helpers behind the array shim (model.np, utils/train.py:46-116) */
int nlps_strided_copy(void* stream, int elem_size, int ndim, const int64_t* shape,
                      const void* d_src, const int64_t* src_stride, void* d_dst, const int64_t* dst_stride);
int nlps_gather_rows(void* stream, int elem_size, int64_t n_out, int64_t row_elems, const void* d_src,
                     int64_t src_row_stride, const int32_t* d_index, int64_t n_src, void* d_dst);
/* H[idx, t, :] on a dense (n0, n1, row_elems) source, each index wrapped on its own axis (utils/train.py:113-115) */
int nlps_gather_rows2(void* stream, int elem_size, int64_t n_out, int64_t row_elems, const void* d_src, int64_t n0,
                      int64_t n1, const int32_t* d_i0, const int32_t* d_i1, void* d_dst);
int nlps_gather_last(void* stream, int B, int S, int E, const float* d_h, const int32_t* d_t, float* d_out);
int nlps_count_nonzero_rows(void* stream, int64_t rows, int64_t cols, const int32_t* d_x,
                            int64_t row_stride, int32_t* d_out);
/* op: 0 add 1 sub 2 mul 3 div 4 max ; b_mode: 0 scalar, 1 same shape, 2 row vector (cols) */
int nlps_binary_f32(void* stream, int op, int64_t n, int64_t cols, const float* d_a, const float* d_b,
                    float b_scalar, int b_mode, float* d_out);
int nlps_binary_i32(void* stream, int op, int64_t n, const int32_t* d_a, const int32_t* d_b,
                    int32_t b_scalar, int b_mode, int32_t* d_out);
/* op: 0 exp 1 sqrt 2 square 3 log 4 neg */
int nlps_unary_f32(void* stream, int op, int64_t n, const float* d_a, float* d_out);
/* op: 0 sum 1 max 2 sum of squares ; result is one element on the device */
int nlps_reduce_f32(void* stream, int op, int64_t n, const float* d_a, float* d_out);
int nlps_reduce_i32(void* stream, int op, int64_t n, const int32_t* d_a, int32_t* d_out);
int nlps_compare_ne_i32(void* stream, int64_t n, const int32_t* d_a, int32_t value, int32_t* d_out);
int nlps_cast_i32_f32(void* stream, int64_t n, const int32_t* d_a, float* d_out);

/* ---- dataset construction in front of the loop ---------------------------------
 * generate_sequence_data / generate_bpe_sequence_data (Transformer/transformer_train.py:47-60,
 * Transformer/transformer_train_bpe.py:52-65) with pad_sequence (Transformer/transformer.py:575-582).
 * Sentences are a ragged int32 array: ids d_ids[d_sent_off[s] .. d_sent_off[s+1]).  Row r of X is the
 * r-th proper prefix in sentence order (d_row_off[s] = sum over earlier sentences of max(len-1, 0),
 * n_sent+1 entries, d_row_off[n_sent] = n_rows), right-padded with `pad`, keeping the LAST max_len
 * ids of longer prefixes; Y[r] is the id that follows the prefix.  X is (n_rows, max_len) int32. */
int nlps_prefix_dataset(void* stream, int n_sent, int64_t n_rows, int max_len, int32_t pad,
                        const int32_t* d_ids, const int64_t* d_sent_off, const int64_t* d_row_off,
                        int32_t* d_X, int32_t* d_Y);
/* BPEEncoder.encode_ids over a batch of sentences (Transformer/bpe_encoder.py:225-249): word-type id w
 * expands to d_tab_ids[d_tab_off[w] .. d_tab_off[w+1]) (the host builds this table once per word type,
 * the reference's encode_word cache); bos/eos < 0 = not added.  Two calls: _offsets writes the n_sent+1
 * output offsets (d_len_scratch: n_sent int32), the caller reads d_out_off[n_sent] to size d_out_ids,
 * then _fill writes the ids. */
int nlps_expand_tokens_offsets(void* stream, int n_sent, const int32_t* d_words, const int64_t* d_sent_off,
                               const int32_t* d_tab_off, int32_t bos, int32_t eos, int32_t* d_len_scratch,
                               int64_t* d_out_off);
int nlps_expand_tokens_fill(void* stream, int n_sent, const int32_t* d_words, const int64_t* d_sent_off,
                            const int32_t* d_tab_off, const int32_t* d_tab_ids, int32_t bos, int32_t eos,
                            const int64_t* d_out_off, int32_t* d_out_ids);

#ifdef __cplusplus
}
#endif
#endif  /* NLPS_B200_H_ */
