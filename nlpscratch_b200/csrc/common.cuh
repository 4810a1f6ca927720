// Shared helpers for libnlps_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>

namespace nlps {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();
void count_launch(int n = 1);         // every kernel launch of ours goes through this counter
// dispatch statistics: how often each implementation of a contraction / attention served a call (nlps_kernel_stats)
enum { KIND_GEMM_PAIR = 0, KIND_GEMM_TC_SINGLE, KIND_GEMM_SIMT, KIND_GEMM_X3, KIND_ATTN_TS_FWD, KIND_ATTN_TS_BWD,
       KIND_ATTN_MMA, KIND_ATTN_FP32, KIND_ATTN_SS, KIND_COUNT };
void count_kind(int kind);
long long kind_count(int kind, bool reset);
long long launch_count(bool reset);
void ensure_mempool_config();
unsigned long long alloc_epoch();      // see util.cu: invalidates captured CUDA graphs after a re-allocation
void bump_alloc_epoch();

#define NLPS_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::nlps::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

#define NLPS_TRY(expr)                \
  do {                                \
    int _s = (expr);                  \
    if (_s != 0) return _s;           \
  } while (0)

#define NLPS_REQUIRE(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      ::nlps::set_error(__VA_ARGS__); \
      return 2;                       \
    }                                 \
  } while (0)

// check the launch itself (not the execution): cheap, no sync
#define NLPS_LAUNCH_CHECK()                                                                \
  do {                                                                                     \
    ::nlps::count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::nlps::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

constexpr int kNumSMs = 148;           // B200
constexpr float kLnEps = 1e-6f;        // transformer.py:150,161

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox4x32-10 (Salmon et al. 2011).  counter = (idx4, stream, step, 0); key = seed.
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

struct DropoutParams {
  float p;             // drop probability; <=0 disables
  float inv_keep;      // 1/(1-p)
  uint32_t threshold;  // keep iff rand16 < threshold  (threshold = round(keep * 2^16))
  uint32_t seed_lo, seed_hi;
  uint32_t stream;     // (layer*2 + which)
  uint32_t step;       // forward-call counter so that every step draws fresh masks
  // when set, the counter is read from device memory instead: the value is then not baked into the launch parameters
  // and a captured CUDA graph of the step can be replayed with a new counter
  const uint32_t* step_ptr;
};

// One Philox4x32-10 call yields the keep decisions of 8 consecutive elements (16 random bits each):
// element idx uses call idx>>3, 16-bit field idx&7.  Inverted dropout, transformer.py:102-108.
__device__ __forceinline__ uint4 dropout_rand8(const DropoutParams& d, uint64_t idx8) {
  const uint32_t step = d.step_ptr != nullptr ? __ldg(d.step_ptr) : d.step;
  return philox4x32((uint32_t)idx8, (uint32_t)(idx8 >> 32), d.stream, step, d.seed_lo, d.seed_hi);
}
__device__ __forceinline__ bool dropout_keep(const DropoutParams& d, const uint4& r, int field) {
  const uint32_t w = (field >> 1) == 0 ? r.x : (field >> 1) == 1 ? r.y : (field >> 1) == 2 ? r.z : r.w;
  return ((w >> ((field & 1) * 16)) & 0xFFFFu) < d.threshold;
}
__device__ __forceinline__ float dropout_one(const DropoutParams& d, uint64_t idx, float v) {
  const uint4 r = dropout_rand8(d, idx >> 3);
  return dropout_keep(d, r, (int)(idx & 7)) ? v * d.inv_keep : 0.f;
}
// 4 consecutive elements starting at idx (idx % 4 == 0)
__device__ __forceinline__ void dropout_four(const DropoutParams& d, uint64_t idx, float4& v) {
  const uint4 r = dropout_rand8(d, idx >> 3);
  const uint32_t w0 = (idx & 4) ? r.z : r.x, w1 = (idx & 4) ? r.w : r.y;
  v.x = (w0 & 0xFFFFu) < d.threshold ? v.x * d.inv_keep : 0.f;
  v.y = (w0 >> 16) < d.threshold ? v.y * d.inv_keep : 0.f;
  v.z = (w1 & 0xFFFFu) < d.threshold ? v.z * d.inv_keep : 0.f;
  v.w = (w1 >> 16) < d.threshold ? v.w * d.inv_keep : 0.f;
}

inline DropoutParams make_dropout(float p, uint64_t seed, uint32_t stream, uint32_t step, const uint32_t* step_ptr = nullptr) {
  DropoutParams d{};
  d.p = p;
  if (p > 0.f) {
    const double keep = 1.0 - (double)p;
    d.inv_keep = (float)(1.0 / keep);
    const double t = keep * 65536.0 + 0.5;
    d.threshold = t >= 65536.0 ? 65536u : (uint32_t)t;
  } else {
    d.inv_keep = 1.f;
    d.threshold = 65536u;
  }
  d.seed_lo = (uint32_t)seed;
  d.seed_hi = (uint32_t)(seed >> 32);
  d.stream = stream;
  d.step = step;
  d.step_ptr = step_ptr;
  return d;
}

// ---- GEMM argument block shared by the tcgen05 and the CUDA-core implementations -------
struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const float* A = nullptr; int64_t lda = 0; bool trans_a = false;   // trans_a: stored [K,M]
  const float* B = nullptr; int64_t ldb = 0; bool trans_b = false;   // trans_b: stored [N,K]
  float* C = nullptr; int64_t ldc = 0;
  const float* bias = nullptr;      // [N], added first
  bool relu = false;                // then max(0,.)
  const float* gate = nullptr; int64_t ldg = 0;       // then *= (gate > 0)       (ReLU backward)
  // bit-packed ReLU mask, one uint32 per (row, 32 columns): written by the ReLU GEMM, read by its
  // backward instead of the float gate (tensor-core path only; needs N % 32 == 0)
  uint32_t* relu_bits_out = nullptr; const uint32_t* gate_bits = nullptr; int64_t ld_bits = 0;
  DropoutParams drop = {0.f, 1.f, 65536u, 0, 0, 0, 0, nullptr};   // then inverted dropout
  const float* residual = nullptr; int64_t ldr = 0;   // then += residual
  bool accumulate = false;          // finally C += result instead of C = result
  // optional by-product (tensor-core CTA-pair kernel only, see gemm_tf32_colsum_fusable): per 32-row block column sums of
  // the stored result, [ceil(M/32)][N] floats -- the bias gradient of the layer below without another pass over C
  float* colsum32_out = nullptr;
  // second by-product (same kernel, same conditions): softmax statistics of every 32-column block of a row,
  // rowstat_out[row * ceil(N/32) + block] = {max, sum exp(v - max)} -- the cross-entropy then skips its first pass
  // over the logits (the part of the vocab-head + CE fusion that pays at V = 8192)
  float2* rowstat_out = nullptr;
  // third by-product (same kernel, same conditions, see gemm_tf32_rowdot_fusable): blocked row dots of the stored result
  // with a second matrix, rowdot_out[((row / rowdot_S) * (N / rowdot_w) + col / rowdot_w) * rowdot_S + row % rowdot_S] =
  // sum over the rowdot_w (32 or 64) columns of the block of C[row, col] * rowdot_in[row, col].  With C = dctx and
  // rowdot_in = ctx this is the row term of softmax_backward moved through P@V, delta[b,h,i] = dO_i . O_i
  // (function.py:59-65), without the extra pass over dctx and ctx.
  const float* rowdot_in = nullptr; int64_t ld_rowdot = 0; float* rowdot_out = nullptr; int rowdot_w = 0, rowdot_S = 0;
};

// bias -> relu -> gate -> dropout -> residual, identical in every GEMM implementation
__device__ __forceinline__ float gemm_epilogue_one(const GemmArgs& g, float acc, int row, int col) {
  if (g.bias) acc += g.bias[col];
  if (g.relu) acc = fmaxf(acc, 0.f);
  if (g.gate) acc = g.gate[(int64_t)row * g.ldg + col] > 0.f ? acc : 0.f;
  if (g.drop.p > 0.f) acc = dropout_one(g.drop, (uint64_t)row * (uint64_t)g.N + (uint64_t)col, acc);
  if (g.residual) acc += g.residual[(int64_t)row * g.ldr + col];
  return acc;
}

int gemm_fp32(cudaStream_t st, const GemmArgs& g);       // gemm_simt.cu
int gemm_tf32(cudaStream_t st, const GemmArgs& g);       // gemm_tcgen05.cu  (1xTF32)
int gemm_tf32x3(cudaStream_t st, const GemmArgs& g);     // 3-pass split, FP32-level error
bool gemm_tf32_supported(const GemmArgs& g);
bool gemm_tf32_colsum_fusable(const GemmArgs& g);   // colsum32_out may be set for this product
bool gemm_tf32_rowdot_fusable(const GemmArgs& g);   // rowdot_out may be set for this product (rowdot_* fields filled in)
int gemm_dispatch(cudaStream_t st, int precision, const GemmArgs& g);
// comm_nccl.cu: libnccl bound at run time (dlopen); all return 0 on success
int comm_available(int* version);
int comm_unique_id(void* out128);
int comm_create(const void* id128, int rank, int world, void** comm_out);
int comm_destroy(void* comm);
int comm_allreduce_sum(void* comm, float* buf, size_t count, cudaStream_t st);
int comm_broadcast(void* comm, float* buf, size_t count, int root, cudaStream_t st);
// tf32_peak.cu: measured tcgen05 kind::tf32 issue ceiling (operands resident in shared memory, accumulators in TMEM)
int tf32_mma_peak(int cta_group, int iters, int reps, float* tflops, float* ms);

// dataset.cu
int prefix_dataset(cudaStream_t st, int n_sent, int64_t n_rows, int max_len, int32_t pad, const int32_t* ids,
                   const int64_t* sent_off, const int64_t* row_off, int32_t* X, int32_t* Y);
int expand_tokens_offsets(cudaStream_t st, int n_sent, const int32_t* words, const int64_t* sent_off, const int32_t* tab_off,
                          int32_t bos, int32_t eos, int32_t* len_scratch, int64_t* out_off);
int expand_tokens_fill(cudaStream_t st, int n_sent, const int32_t* words, const int64_t* sent_off, const int32_t* tab_off,
                       const int32_t* tab_ids, int32_t bos, int32_t eos, const int64_t* out_off, int32_t* out_ids);

// attention.cu
int attention_forward(cudaStream_t st, int precision, int B, int S, int h, int d, const float* qkv,
                      const int32_t* x, const int32_t* first_nonpad, float* ctx, float* lse);
int attention_backward(cudaStream_t st, int precision, int B, int S, int h, int d, const float* qkv,
                       const int32_t* x, const int32_t* first_nonpad, const float* ctx, const float* lse,
                       const float* dctx, float* delta_scratch, float* dqkv, bool delta_ready = false);
// delta_ready: delta_scratch already holds sum(dO*O) per (b,h,i) (by-product of the Wo dgrad GEMM, GemmArgs::rowdot_out)
int first_nonpad(cudaStream_t st, int B, int S, const int32_t* x, int32_t* out);
// which shapes the tcgen05 attention serves (head_dim 32 or 64, TMA-addressable rows); attention.cu
bool attention_tc_supported(int d, int E, const float* qkv);
// attention_ts.cu (tcgen05 with the A operands in TMEM, head_dim 32 or 64)
int attention_forward_ts(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                         const int32_t* fnp, float* ctx, float* lse);
int attention_backward_ts(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                          const int32_t* fnp, const float* dctx, const float* lse, const float* delta, float* dqkv);

// attention_mma.cu (TF32 tensor-core versions, head_dim <= 64)
int attention_forward_mma(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                          const int32_t* first_nonpad, float* ctx, float* lse, bool x3 = false);
int attention_backward_mma(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                           const int32_t* first_nonpad, const float* dctx, const float* lse, const float* delta,
                           float* dqkv, bool x3 = false);   // x3: 3xTF32 split products (FP32-level accuracy)

// rowwise.cu
int layernorm_forward(cudaStream_t st, int64_t rows, int E, const float* x, const float* gamma,
                      const float* beta, float* y, float* mean, float* rstd);
int layernorm_backward(cudaStream_t st, int64_t rows, int E, const float* dy, const float* x,
                       const float* mean, const float* rstd, const float* gamma, float* dx,
                       float* dx_dropped, const DropoutParams* drop,
                       float* dgamma, float* dbeta, float* partial_scratch,
                       float* dbias = nullptr,    // optional: column sums of the branch gradient leaving the kernel (dx_dropped, else dx)
                       bool defer_finish = false);   // true: the partial sums stay in partial_scratch for layernorm_backward_finish
int layernorm_backward_finish(cudaStream_t st, int64_t rows, int E, const float* partial, float* dgamma, float* dbeta,
                              float* dbias = nullptr);
size_t layernorm_backward_scratch_floats(int E);
// rowstat != nullptr: per-32-column {max, sum exp} blocks from the vocab-head GEMM epilogue (see GemmArgs::rowstat_out)
int softmax_xent_rows(cudaStream_t st, const float* logits, int64_t ld, const int32_t* y, int64_t rows, int V, float inv,
                      float* dlogits, int64_t ldd, float* row_loss, int32_t* err = nullptr);
int row_loss_mean(cudaStream_t st, const float* row_loss, int64_t rows, float* loss);
// cross-entropy whose pass also leaves the column-sum partials of dlogits (the vocabulary-bias gradient) behind
int64_t xent_colsum_partial_rows(int64_t rows);
bool xent_colsum_fusable(const float* logits, int64_t ld, const float* dlogits, int64_t ldd, int V);
int softmax_xent_colsum(cudaStream_t st, const float* logits, int64_t ld, const int32_t* y, int64_t rows, int V,
                        float grad_scale, float* loss, float* dlogits, int64_t ldd, float* row_loss, float* partial,
                        int64_t ld_partial, int32_t* err);
// device-side evaluation bookkeeping of the driver loop (utils/train.py:51-75, 86-88, 143-163)
int eval_accumulate(cudaStream_t st, const float* row_loss, int64_t rows, float* accum);
int eval_finish(cudaStream_t st, float* accum, float* sched);
int topk_softmax(cudaStream_t st, const float* logits, int V, int k, float* work, int32_t* out_idx, float* out_prob);
int softmax_xent_stats(cudaStream_t st, const float* logits, int64_t ld, const int32_t* y, int64_t rows, int V,
                       float grad_scale, float* loss, float* dlogits, int64_t ldd, float* row_loss, const float2* rowstat,
                       int32_t* err = nullptr);
int softmax_xent(cudaStream_t st, const float* logits, int64_t ld, const int32_t* y, int64_t rows, int V,
                 float grad_scale, float* loss, float* dlogits, int64_t ldd, float* row_loss_scratch,
                 int32_t* err = nullptr);   // err: model error word, bit 1 = target id out of range
// tickets: ceil(cols/128) zero-initialised counters (self-resetting) -> the reduction finishes inside the one kernel;
// nullptr -> separate finish kernel
int colsum(cudaStream_t st, int64_t rows, int cols, const float* x, int64_t ld, float* out, float* scratch,
           unsigned int* tickets = nullptr);
size_t colsum_scratch_floats(int cols);
int sumsq(cudaStream_t st, int64_t n, const float* x, float* out, float* scratch);
size_t sumsq_scratch_floats();

// embed.cu
int embed_forward(cudaStream_t st, int B, int S, int E, int V, const int32_t* x, int64_t xs,
                  const float* We, const float* pe, float* out, int32_t* x_flat, int32_t* err_flag);
int embed_backward(cudaStream_t st, int64_t N, int E, int V, const int32_t* x_flat, const float* dh,
                   float* dWe, int32_t* scratch_i32);
int embed_backward_sort(cudaStream_t st, int64_t N, int V, const int32_t* x_flat, int32_t* scratch_i32);   // ids only
int embed_backward_apply(cudaStream_t st, int64_t N, int E, int V, const float* dh, float* dWe, const int32_t* scratch_i32);
size_t embed_backward_scratch_ints(int64_t N, int V);
int scatter_rows(cudaStream_t st, int B, int S, int E, const int32_t* t, const float* rows, float* out);

// optimizer.cu
int clip_scale_from_sumsq(cudaStream_t st, const float* sumsq, float max_norm, float* scale_out, float* norm_out);
int scale_inplace(cudaStream_t st, int64_t n, float* x, const float* scale_dev);
int adam_update(cudaStream_t st, int64_t n, float* p, const float* g, float* m, float* v, float lr,
                float b1, float b2, float one_minus_b1, float one_minus_b2, float eps, float bc1, float bc2,
                const float* clip_scale_dev,
                const float* hyper_dev = nullptr);   // {lr, bc1, bc2} on the device override the by-value arguments
int sgd_update(cudaStream_t st, int64_t n, float* p, const float* g, float lr, const float* clip_scale_dev,
               const float* hyper_dev = nullptr);

}  // namespace nlps

namespace nlps {
// array_ops.cu
int strided_copy(cudaStream_t st, int elem_size, int ndim, const int64_t* shape, const void* src,
                 const int64_t* sstr, void* dst, const int64_t* dstr);
int gather_rows(cudaStream_t st, int elem_size, int64_t n_out, int64_t row_elems, const void* src, int64_t src_stride,
                const int32_t* index, int64_t n_src, void* dst);
int gather_rows2(cudaStream_t st, int elem_size, int64_t n_out, int64_t row_elems, const void* src, int64_t n0, int64_t n1,
                 const int32_t* i0, const int32_t* i1, void* dst);
int gather_last(cudaStream_t st, int B, int S, int E, const float* h, const int32_t* t, float* out);
// keep-mask (1.0 / 0.0) of a dropout stream over n elements: the decisions dropout_one() takes in the GEMM epilogues (forward)
// and in the LayerNorm backward (regenerated), exported for the dropout parity tests
int dropout_mask(cudaStream_t st, const DropoutParams& d, int64_t n, float* out);
int count_nonzero_rows(cudaStream_t st, int64_t rows, int64_t cols, const int32_t* x, int64_t stride, int32_t* out);
int binary_f32(cudaStream_t st, int op, int64_t n, int64_t cols, const float* a, const float* b, float bs, int mode, float* out);
int binary_i32(cudaStream_t st, int op, int64_t n, const int32_t* a, const int32_t* b, int32_t bs, int mode, int32_t* out);
int unary_f32(cudaStream_t st, int op, int64_t n, const float* a, float* out);
int reduce_f32(cudaStream_t st, int op, int64_t n, const float* a, float* out);
int reduce_i32(cudaStream_t st, int op, int64_t n, const int32_t* a, int32_t* out);
int compare_ne_i32(cudaStream_t st, int64_t n, const int32_t* a, int32_t value, int32_t* out);
int cast_i32_f32(cudaStream_t st, int64_t n, const int32_t* a, float* out);
}  // namespace nlps
