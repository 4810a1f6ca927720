// Model handle, memory layout and the training-step orchestration behind the C ABI.
//
// HBM layout (all fp32, every tensor 256-B aligned inside ONE flat buffer per kind):
//   params / grads / adam_m / adam_v :  [ We | layer 0 | ... | layer L-1 | W_vocab | b_vocab ]
//   layer l = [ Wqkv (E,3E: Wq|Wk|Wv packed column-wise) | Wo (E,E) | ln1_gamma | ln1_beta |
//               W1 (E,F) | b1 | W2 (F,E) | b2 | ln2_gamma | ln2_beta ]
//   W_vocab is (E, ldV) with ldV = round_up(V,4) so that every row start is 16-B aligned for TMA;
//   pad columns are zero in params and grads and therefore stay zero under Adam/SGD.
// Gradient buckets (for data-parallel all-reduce) are contiguous ranges of the grad buffer in the
// order the backward pass completes them: [vocab] [layer L-1] ... [layer 0] [We].
//
// Activations of one forward call live in one allocation per nlps_cache (pooled by size):
//   per layer: qkv (N,3E), ctx (N,E), lse (B,h,S), r1 (N,E), mean1/rstd1 (N), y1 (N,E),
//              hid (N,F), r2 (N,E), mean2/rstd2 (N), y2 (N,E);  plus h0 (N,E), x_flat, first_nonpad,
//              logits (N,ldV).  The (B,h,S,S) attention tensors and dropout masks are never stored.
#include <cstdlib>
#include "common.cuh"
#include "../../include/nlps_b200.h"
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

using namespace nlps;

namespace {
constexpr int64_t kAlign = 64;   // floats (256 B)

struct LayerOff {
  int64_t wqkv, wo, ln1g, ln1b, w1, b1, w2, b2, ln2g, ln2b, begin, end;
};
}  // namespace

struct nlps_cache {
  int B = 0, S = 0;
  int64_t N = 0;
  size_t bytes = 0;
  char* base = nullptr;
  int32_t* x_flat = nullptr;
  int32_t* fnp = nullptr;
  float* h0 = nullptr;
  struct L { float *qkv, *ctx, *lse, *r1, *mean1, *rstd1, *y1, *hid, *r2, *mean2, *rstd2, *y2; uint32_t* relu_bits; };
  std::vector<L> layers;
  float* logits = nullptr;            // (N, ldV) logits / dlogits + the rowstat block: allocated on first use (ensure_cache_logits) --
  char* logits_base = nullptr;        // the fused vocabulary path (NLPS_FUSE_VOCAB) never needs them
  uint32_t* dev_step = nullptr;       // drop_step on the device: what the kernels of this cache read (graph-replayable)
  float2* rowstat = nullptr;          // [N][ceil(V/32)] softmax statistics left by the vocab-head GEMM epilogue
  bool rowstat_valid = false;
  float drop_p = 0.f;       // dropout probability in effect during this forward
  uint32_t drop_step = 0;
  bool has_logits = false;
  bool valid = false;
  bool relu_bits_valid = false;   // the forward ran the tensor-core FFN and left bit-packed ReLU masks
};

struct nlps_model {
  nlps_config cfg{};
  int device = 0;
  int V = 0, E = 0, h = 0, d = 0, F = 0, L = 0, max_len = 0;
  int64_t ldV = 0;
  int precision = NLPS_PREC_TF32;
  bool training = true;
  float dropout_p = 0.f;
  uint64_t seed = 0;
  uint32_t fwd_counter = 0;
  int64_t adam_t = 0;
  cudaStream_t stream = nullptr;

  int64_t flat = 0;                 // floats per flat buffer
  float *P = nullptr, *G = nullptr, *M = nullptr, *Vv = nullptr, *pe = nullptr;
  int64_t off_We = 0, off_Wvocab = 0, off_bvocab = 0;
  std::vector<LayerOff> lo;
  std::vector<std::pair<int64_t, int64_t>> buckets;   // (first, count) in completion order
  std::vector<cudaEvent_t> bucket_ev;
  cudaEvent_t aux_ev = nullptr;

  // backward workspace (grows with N)
  int64_t ws_N = 0;
  float *g0 = nullptr, *g1 = nullptr, *gdrop = nullptr, *dctx = nullptr, *dqkv = nullptr, *dhid = nullptr;
  float* colsum32 = nullptr;                  // [ceil(N/32)][F] column-sum partials written by the dz GEMM epilogue
  float* vchunk = nullptr;                    // fused vocabulary path: (vchunk_rows, ldV) logits -> dlogits of one row chunk
  int64_t vchunk_rows = 0;
  float* dbv_partial = nullptr;               // [ceil(N/16)][ldV] column-sum partials of dlogits left by the fused cross-entropy
  int64_t dbv_partial_rows = 0;               // > 0: the next nlps_backward takes db_vocab from them
  float *g1b = nullptr, *gdropb = nullptr;    // LayerNorm-2 backward outputs (own buffers so that the weight-gradient stream may lag)
  // weight-gradient stream: dW2, dW1, dWo, dWqkv and dW_vocab are off the critical path of backward; on a second stream
  // they fill the tensor pipe while the main stream runs its bandwidth-bound kernels (LayerNorm backward, column sums,
  // split-K reducers).  ev_fork: main -> wstream; ev_w[k]: "the k-th kind of weight gradient has been formed" (guards the
  // buffers it read); w_pending[k]: recorded in this pass.
  cudaStream_t wstream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_w[8] = {};
  bool w_pending[8] = {};
  float *delta = nullptr, *row_loss = nullptr, *small = nullptr, *rows_tmp = nullptr;
  int32_t* iscratch = nullptr;
  int64_t rows_tmp_cap = 0;
  // fixed scratch
  float *ln_partial = nullptr, *colsum_scratch = nullptr, *sumsq_scratch = nullptr;
  float* colsum_scratch_w = nullptr;          // column sums issued on the weight-gradient stream use their own scratch + tickets
  unsigned int* colsum_tickets = nullptr;     // 4096 self-resetting counters: one per 128-column group of a colsum
  float* scalars = nullptr;          // [0]=sumsq [1]=clip scale [2]=norm [3]=loss
  int32_t* err_flag = nullptr;

  std::vector<nlps_cache*> pool;

  // ---- data-parallel exchange owned by the library (nlps_comm_init): NCCL communicator, its stream, the coalesced bucket
  // groups (ranges of bucket indices in completion order), and whether the current call reduces inside backward ----
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_done = nullptr;
  std::vector<std::pair<int, int>> comm_plan;
  bool comm_active = false;

  // ---- CUDA graph of the fused full-sequence step (nlps_train_step) ----
  // One graph per (cache, x, y, clip, scale, mode) signature; the only node whose parameters change from step to step is
  // the setter that publishes the step's scalars (dropout counter, lr, Adam bias corrections) to device memory.
  struct StepGraph {
    const nlps_cache* cache; const int32_t* x; int64_t xs; const int32_t* y;
    const int32_t* t;                  // last-token step: the (B,) index buffer; nullptr = full-sequence step
    float max_norm, grad_scale, dropout_p; int precision; bool training; bool deferred;
    unsigned long long epoch;
    int state;                         // 1 = signature seen once (ran eagerly), 2 = captured, 3 = not capturable
    cudaGraphExec_t exec = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphNode_t setter = nullptr;
    long long launches = 0;
    bool relu_bits_valid = false;
  };
  std::vector<StepGraph> graphs;
  long long n_eager = 0, n_captured = 0, n_replayed = 0;   // how the fused steps ran (nlps_graph_stats)
  cudaStream_t own_stream = nullptr;   // created when the model would otherwise run on the legacy default stream
  float* hyper = nullptr;              // device {lr, bc1, bc2}
  bool in_graph_body = false;          // the per-call setters are replaced by the one setter node
  bool graph_external_events = false;  // capturing a deferred-update step: bucket events become external record nodes
  int graph_mode = -1;
};

namespace {

// the step's scalars on the device: every other kernel of the step reads them from there
__global__ void set_step_scalars_kernel(uint32_t* step_slot, uint32_t step, float* hyper, float lr, float bc1, float bc2) {
  if (step_slot != nullptr) *step_slot = step;
  if (hyper != nullptr) { hyper[0] = lr; hyper[1] = bc1; hyper[2] = bc2; }
}

inline int64_t take(int64_t& cursor, int64_t n) {
  const int64_t at = cursor;
  cursor = round_up(cursor + n, kAlign);
  return at;
}

int build_layout(nlps_model* m) {
  const int64_t E = m->E, F = m->F, V = m->V;
  m->ldV = round_up(V, 4);
  int64_t c = 0;
  m->off_We = take(c, V * E);
  m->lo.resize(m->L);
  for (int l = 0; l < m->L; ++l) {
    LayerOff& o = m->lo[l];
    o.begin = c;
    o.wqkv = take(c, E * 3 * E);
    o.wo = take(c, E * E);
    o.ln1g = take(c, E);
    o.ln1b = take(c, E);
    o.w1 = take(c, E * F);
    o.b1 = take(c, F);
    o.w2 = take(c, F * E);
    o.b2 = take(c, E);
    o.ln2g = take(c, E);
    o.ln2b = take(c, E);
    o.end = c;
  }
  const int64_t vocab_begin = c;
  m->off_Wvocab = take(c, E * m->ldV);
  m->off_bvocab = take(c, m->ldV);
  m->flat = c;
  m->buckets.clear();
  m->buckets.push_back({vocab_begin, c - vocab_begin});
  for (int l = m->L - 1; l >= 0; --l) m->buckets.push_back({m->lo[l].begin, m->lo[l].end - m->lo[l].begin});
  m->buckets.push_back({m->off_We, m->lo.empty() ? vocab_begin - m->off_We : m->lo[0].begin - m->off_We});
  return 0;
}

int ensure_workspace(nlps_model* m, int B, int S) {
  const int64_t N = (int64_t)B * S;
  if (N <= m->ws_N) return 0;
  NLPS_CUDA(cudaStreamSynchronize(m->stream));
  if (m->wstream) NLPS_CUDA(cudaStreamSynchronize(m->wstream));
  float** bufs[] = {&m->g0, &m->g1, &m->gdrop, &m->dctx, &m->dqkv, &m->dhid, &m->delta, &m->row_loss, &m->colsum32, &m->g1b, &m->gdropb,
                    &m->dbv_partial};
  for (float** b : bufs) { if (*b) cudaFree(*b); *b = nullptr; }
  bump_alloc_epoch();
  if (m->iscratch) { cudaFree(m->iscratch); m->iscratch = nullptr; }
  const int64_t E = m->E, F = m->F;
  NLPS_CUDA(cudaMalloc(&m->g0, sizeof(float) * N * E));
  NLPS_CUDA(cudaMalloc(&m->g1, sizeof(float) * N * E));
  NLPS_CUDA(cudaMalloc(&m->gdrop, sizeof(float) * N * E));
  NLPS_CUDA(cudaMalloc(&m->g1b, sizeof(float) * N * E));
  NLPS_CUDA(cudaMalloc(&m->gdropb, sizeof(float) * N * E));
  NLPS_CUDA(cudaMalloc(&m->dctx, sizeof(float) * N * E));
  NLPS_CUDA(cudaMalloc(&m->dqkv, sizeof(float) * N * 3 * E));
  NLPS_CUDA(cudaMalloc(&m->dhid, sizeof(float) * N * F));
  NLPS_CUDA(cudaMalloc(&m->colsum32, sizeof(float) * ceil_div(N, 32) * F));
  NLPS_CUDA(cudaMalloc(&m->dbv_partial, sizeof(float) * xent_colsum_partial_rows(N) * m->ldV));
  NLPS_CUDA(cudaMalloc(&m->delta, sizeof(float) * N * m->h));
  NLPS_CUDA(cudaMalloc(&m->row_loss, sizeof(float) * N));
  NLPS_CUDA(cudaMalloc(&m->iscratch, sizeof(int32_t) * embed_backward_scratch_ints(N, m->V)));
  m->ws_N = N;
  return 0;
}

int ensure_rows_tmp(nlps_model* m, int64_t floats) {
  if (floats <= m->rows_tmp_cap) return 0;
  NLPS_CUDA(cudaStreamSynchronize(m->stream));
  if (m->rows_tmp) cudaFree(m->rows_tmp);
  bump_alloc_epoch();
  m->rows_tmp = nullptr;
  NLPS_CUDA(cudaMalloc(&m->rows_tmp, sizeof(float) * floats));
  m->rows_tmp_cap = floats;
  return 0;
}

GemmArgs gemm_nn(int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc) {
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  return g;
}

DropoutParams drop_for(const nlps_model* m, const nlps_cache* c, int layer, int which) {
  return make_dropout(c->drop_p, m->seed, (uint32_t)(layer * 2 + which), c->drop_step, c->dev_step);
}

inline unsigned int* colsum_tickets_for(nlps_model* m, int64_t cols) { return ceil_div(cols, 128) <= 4096 ? m->colsum_tickets : nullptr; }
// NLPS_SIDE_SUMS=0 keeps the reductions nobody waits for (LayerNorm parameter-gradient finish, db1, db_vocab) on the main
// stream as in round 1 (A/B); default: they ride on the weight-gradient stream
inline bool side_sums() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NLPS_SIDE_SUMS"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// column sums issued on the weight-gradient stream may run next to main-stream ones: own counters
inline unsigned int* colsum_tickets_w(nlps_model* m, int64_t cols) { return ceil_div(cols, 128) <= 4096 ? m->colsum_tickets + 4096 : nullptr; }

// graphs that replay into a cache that is about to be freed
void drop_graphs_of_cache(nlps_model* m, const nlps_cache* c) {
  for (size_t i = 0; i < m->graphs.size();) {
    if (m->graphs[i].cache == c) {
      if (m->graphs[i].exec) cudaGraphExecDestroy(m->graphs[i].exec);
      if (m->graphs[i].graph) cudaGraphDestroy(m->graphs[i].graph);
      m->graphs.erase(m->graphs.begin() + i);
    } else {
      ++i;
    }
  }
}

void drop_step_graphs(nlps_model* m) {
  for (auto& g : m->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
  }
  m->graphs.clear();
}

inline void adam_bias_corrections(int64_t t, float* bc1, float* bc2) {
  *bc1 = (float)(1.0 - std::pow(0.9, (double)t));
  *bc2 = (float)(1.0 - std::pow(0.999, (double)t));
}

// ---- weight-gradient stream plumbing (no-ops when the model has none) ----
enum { W_DW2 = 0, W_DW1 = 1, W_DWO = 2, W_DWQKV = 3, W_VOCAB = 4, W_JOIN = 5, W_XENT = 6, W_SORT = 7 };
// the stream the next weight-gradient product goes to, ordered after everything enqueued on the main stream so far
int wgrad_fork(nlps_model* m, cudaStream_t* out) {
  if (!m->wstream) { *out = m->stream; return 0; }
  NLPS_CUDA(cudaEventRecord(m->ev_fork, m->stream));
  NLPS_CUDA(cudaStreamWaitEvent(m->wstream, m->ev_fork, 0));
  *out = m->wstream;
  return 0;
}
int wgrad_done(nlps_model* m, int kind) {
  if (!m->wstream) return 0;
  NLPS_CUDA(cudaEventRecord(m->ev_w[kind], m->wstream));
  m->w_pending[kind] = true;
  return 0;
}
// the main stream is about to overwrite what weight gradient `kind` read (or needs its result)
int wgrad_need(nlps_model* m, int kind) {
  if (!m->wstream || !m->w_pending[kind]) return 0;
  NLPS_CUDA(cudaStreamWaitEvent(m->stream, m->ev_w[kind], 0));
  m->w_pending[kind] = false;
  return 0;
}
// everything on the weight-gradient stream so far is ordered before what the main stream does next
int wgrad_join(nlps_model* m) {
  if (!m->wstream) return 0;
  NLPS_CUDA(cudaEventRecord(m->ev_w[W_JOIN], m->wstream));
  NLPS_CUDA(cudaStreamWaitEvent(m->stream, m->ev_w[W_JOIN], 0));
  for (bool& b : m->w_pending) b = false;
  return 0;
}

// contiguous runs of the flat buffer covered by buckets [lo, hi) (adjacent layers are adjacent in memory)
std::vector<std::pair<int64_t, int64_t>> merged_ranges(const nlps_model* m, int lo, int hi) {
  std::vector<std::pair<int64_t, int64_t>> r(m->buckets.begin() + lo, m->buckets.begin() + hi);
  std::sort(r.begin(), r.end());
  std::vector<std::pair<int64_t, int64_t>> out;
  for (auto& q : r) {
    if (!out.empty() && out.back().first + out.back().second == q.first) out.back().second += q.second;
    else out.push_back(q);
  }
  return out;
}

// bucket `bucket` is final (its event has been recorded): if it closes a group of the exchange plan, the communication
// stream waits for that event and sums the group's ranges over ranks while backward goes on
int comm_on_bucket(nlps_model* m, int bucket) {
  for (auto& grp : m->comm_plan) {
    if (grp.second - 1 != bucket) continue;
    NLPS_CUDA(cudaStreamWaitEvent(m->comm_stream, m->bucket_ev[bucket], 0));
    for (auto& r : merged_ranges(m, grp.first, grp.second))
      NLPS_TRY(comm_allreduce_sum(m->comm, m->G + r.first, (size_t)r.second, m->comm_stream));
  }
  return 0;
}
// the main stream continues (clip + update) only after every reduction
int comm_join(nlps_model* m) {
  NLPS_CUDA(cudaEventRecord(m->comm_done, m->comm_stream));
  NLPS_CUDA(cudaStreamWaitEvent(m->stream, m->comm_done, 0));
  return 0;
}

int record_bucket(nlps_model* m, int bucket, bool on_wstream = false) {
  // Who needs the event: the library's own exchange (comm_active), a caller-side exchange (eager step, or a captured
  // deferred step whose records become EXTERNAL event-record nodes that a communication stream waits on after the graph
  // launch).  A captured single-GPU step needs none.
  const bool external = m->in_graph_body && m->graph_external_events && !m->comm_active;
  const bool need_event = m->comm_active || !m->in_graph_body || m->graph_external_events;
  if (!need_event) return 0;
  // a bucket that holds weight gradients is complete when the weight-gradient stream has passed this point too
  cudaStream_t rs = m->stream;
  if (on_wstream && m->wstream) NLPS_TRY(wgrad_fork(m, &rs));
  if (external) {
    NLPS_CUDA(cudaEventRecordWithFlags(m->bucket_ev[bucket], rs, cudaEventRecordExternal));
    return 0;
  }
  NLPS_CUDA(cudaEventRecord(m->bucket_ev[bucket], rs));
  if (m->comm_active) NLPS_TRY(comm_on_bucket(m, bucket));
  return 0;
}

// layers L-1..0 and the embedding scatter, starting from dH in m->g0 (transformer.py:360-457).
// Critical path on the main stream: LN2 bwd -> dz -> d(y1) -> LN1 bwd -> dctx -> attention bwd -> d(x_in); the four weight
// gradients of a layer go to the weight-gradient stream as soon as their inputs exist (wgrad_fork) and the main stream
// only waits for one of them right before it overwrites what that product read (wgrad_need).
int backward_stack(nlps_model* m, nlps_cache* c) {
  const int E = m->E, F = m->F;
  const int N = (int)c->N;
  cudaStream_t st = m->stream;
  cudaStream_t ws = st;
  const int prec = m->precision;
  const bool dropping = c->drop_p > 0.f;
  float* g1b = m->wstream ? m->g1b : m->g1;            // one buffer set is enough without the second stream
  float* gdropb = m->wstream ? m->gdropb : m->gdrop;
  for (int k = 0; k < W_VOCAB; ++k) m->w_pending[k] = false;   // nothing of an earlier (possibly failed) pass is waited for
  // the counting sort of the token ids for the embedding scatter depends on x only: on the second stream it is done long
  // before dH of layer 0 exists instead of sitting at the very end of the step (NLPS_EARLY_SORT=0: A/B)
  static int early_sort = -1;
  if (early_sort < 0) { const char* e = getenv("NLPS_EARLY_SORT"); early_sort = (e && e[0] == '0') ? 0 : 1; }
  const bool sorted_early = m->wstream != nullptr && early_sort;
  if (sorted_early) {
    NLPS_TRY(wgrad_fork(m, &ws));
    NLPS_TRY(embed_backward_sort(ws, N, m->V, c->x_flat, m->iscratch));
    NLPS_TRY(wgrad_done(m, W_SORT));
  }
  for (int l = m->L - 1; l >= 0; --l) {
    const LayerOff& o = m->lo[l];
    const nlps_cache::L& a = c->layers[l];
    const float* x_in = l == 0 ? c->h0 : c->layers[l - 1].y2;
    // LN2 backward (+ dropout backward of the FFN branch)
    DropoutParams dpf = drop_for(m, c, l, 1);
    NLPS_TRY(wgrad_need(m, W_DW2));                    // the layer above read d_f from these buffers
    float* part2 = m->ln_partial + (size_t)(2 * l + 1) * layernorm_backward_scratch_floats(E);
    float* part1 = m->ln_partial + (size_t)(2 * l) * layernorm_backward_scratch_floats(E);
    NLPS_TRY(layernorm_backward(st, N, E, m->g0, a.r2, a.mean2, a.rstd2, m->P + o.ln2g, g1b,
                                dropping ? gdropb : nullptr, dropping ? &dpf : nullptr, m->G + o.ln2g,
                                m->G + o.ln2b, part2, m->G + o.b2, /*defer_finish=*/true));   // db2 = column sums of d_f ride along
    const float* d_f = dropping ? gdropb : g1b;
    {  // dW2 = hid^T d_f ; db2
      GemmArgs g = gemm_nn(F, E, N, a.hid, F, d_f, E, m->G + o.w2, E);
      g.trans_a = true;
      NLPS_TRY(wgrad_fork(m, &ws));
      // the LayerNorm parameter gradients (and db2) are nobody's input: their final sum leaves the critical path
      NLPS_TRY(layernorm_backward_finish(side_sums() ? ws : st, N, E, part2, m->G + o.ln2g, m->G + o.ln2b, m->G + o.b2));
      NLPS_TRY(gemm_dispatch(ws, prec, g));
      NLPS_TRY(wgrad_done(m, W_DW2));
    }
    bool db1_fused = false;
    {  // dz = (d_f W2^T) * (hid > 0)
      GemmArgs g = gemm_nn(N, F, E, d_f, E, m->P + o.w2, E, m->dhid, F);
      g.trans_b = true;
      if (c->relu_bits_valid && prec == NLPS_PREC_TF32) { g.gate_bits = a.relu_bits; g.ld_bits = F / 32; }
      else { g.gate = a.hid; g.ldg = F; }
      // db1 = column sums of dz: 32-row partials come out of this product's epilogue (tensor-core path)
      static int fuse_db1 = -1;             // NLPS_FUSE_DB1=0: separate column-sum pass (A/B measurements)
      if (fuse_db1 < 0) { const char* e = getenv("NLPS_FUSE_DB1"); fuse_db1 = (e && e[0] == '0') ? 0 : 1; }
      if (fuse_db1 && prec == NLPS_PREC_TF32 && m->colsum32 != nullptr && gemm_tf32_colsum_fusable(g)) { g.colsum32_out = m->colsum32; db1_fused = true; }
      NLPS_TRY(wgrad_need(m, W_DW1));                  // the layer above read dz from m->dhid
      NLPS_TRY(gemm_dispatch(st, prec, g));
    }
    {  // dW1 = y1^T dz ; db1
      GemmArgs g = gemm_nn(E, F, N, a.y1, E, m->dhid, F, m->G + o.w1, F);
      g.trans_a = true;
      NLPS_TRY(wgrad_fork(m, &ws));
      NLPS_TRY(gemm_dispatch(ws, prec, g));
      // db1 on the same stream, before the event that guards dhid / the epilogue partials against the next layer's dz product
      {
        const bool side = side_sums();
        cudaStream_t cs = side ? ws : st;
        float* scratch = side ? m->colsum_scratch_w : m->colsum_scratch;
        unsigned int* tickets = side ? colsum_tickets_w(m, F) : colsum_tickets_for(m, F);
        if (db1_fused) NLPS_TRY(colsum(cs, ceil_div(N, 32), F, m->colsum32, F, m->G + o.b1, scratch, tickets));
        else NLPS_TRY(colsum(cs, N, F, m->dhid, F, m->G + o.b1, scratch, tickets));
      }
      NLPS_TRY(wgrad_done(m, W_DW1));
    }
    {  // d(y1) = dz W1^T + d_r2
      GemmArgs g = gemm_nn(N, E, F, m->dhid, F, m->P + o.w1, F, m->g0, E);
      g.trans_b = true;
      g.residual = g1b; g.ldr = E;
      NLPS_TRY(gemm_dispatch(st, prec, g));
    }
    // LN1 backward (+ dropout backward of the attention branch)
    DropoutParams dpa = drop_for(m, c, l, 0);
    NLPS_TRY(wgrad_need(m, W_DWO));                    // the layer above read d_a from these buffers
    NLPS_TRY(layernorm_backward(st, N, E, m->g0, a.r1, a.mean1, a.rstd1, m->P + o.ln1g, m->g1,
                                dropping ? m->gdrop : nullptr, dropping ? &dpa : nullptr, m->G + o.ln1g,
                                m->G + o.ln1b, part1, nullptr, /*defer_finish=*/true));
    const float* d_a = dropping ? m->gdrop : m->g1;
    {  // dWo = ctx^T d_a
      GemmArgs g = gemm_nn(E, E, N, a.ctx, E, d_a, E, m->G + o.wo, E);
      g.trans_a = true;
      NLPS_TRY(wgrad_fork(m, &ws));
      NLPS_TRY(layernorm_backward_finish(side_sums() ? ws : st, N, E, part1, m->G + o.ln1g, m->G + o.ln1b));
      NLPS_TRY(gemm_dispatch(ws, prec, g));
      NLPS_TRY(wgrad_done(m, W_DWO));
    }
    bool delta_ready = false;
    {  // dctx = d_a Wo^T; the epilogue also forms delta = sum(dctx * ctx) per (token, head) when the product allows it
      GemmArgs g = gemm_nn(N, E, E, d_a, E, m->P + o.wo, E, m->dctx, E);
      g.trans_b = true;
      static int fuse_delta = -1;
      if (fuse_delta < 0) { const char* e = getenv("NLPS_FUSE_DELTA"); fuse_delta = (e && e[0] == '0') ? 0 : 1; }
      if (fuse_delta && prec == NLPS_PREC_TF32) {
        GemmArgs t = g;
        t.rowdot_in = a.ctx; t.ld_rowdot = E; t.rowdot_out = m->delta; t.rowdot_w = m->d; t.rowdot_S = c->S;
        if (gemm_tf32_rowdot_fusable(t)) { g = t; delta_ready = true; }
      }
      NLPS_TRY(gemm_dispatch(st, prec, g));
    }
    NLPS_TRY(wgrad_need(m, W_DWQKV));                  // the layer above read m->dqkv
    NLPS_TRY(attention_backward(st, prec, c->B, c->S, m->h, m->d, a.qkv, c->x_flat, c->fnp, a.ctx, a.lse, m->dctx,
                                m->delta, m->dqkv, delta_ready));
    {  // dWqkv = x_in^T dqkv
      GemmArgs g = gemm_nn(E, 3 * E, N, x_in, E, m->dqkv, 3 * E, m->G + o.wqkv, 3 * E);
      g.trans_a = true;
      NLPS_TRY(wgrad_fork(m, &ws));
      NLPS_TRY(gemm_dispatch(ws, prec, g));
      NLPS_TRY(wgrad_done(m, W_DWQKV));
    }
    {  // dx_in = dqkv Wqkv^T + d_r1
      GemmArgs g = gemm_nn(N, E, 3 * E, m->dqkv, 3 * E, m->P + o.wqkv, 3 * E, m->g0, E);
      g.trans_b = true;
      g.residual = m->g1; g.ldr = E;
      NLPS_TRY(gemm_dispatch(st, prec, g));
    }
    NLPS_TRY(record_bucket(m, 1 + (m->L - 1 - l), /*on_wstream=*/true));
  }
  if (sorted_early) {
    NLPS_TRY(wgrad_need(m, W_SORT));
    NLPS_TRY(embed_backward_apply(st, N, E, m->V, m->g0, m->G + m->off_We, m->iscratch));
  } else {
    NLPS_TRY(embed_backward(st, N, E, m->V, c->x_flat, m->g0, m->G + m->off_We, m->iscratch));
  }
  NLPS_TRY(wgrad_join(m));                             // every gradient is complete on the main stream from here on
  NLPS_TRY(record_bucket(m, m->L + 1));
  return 0;
}

// the (N, ldV) logits / dlogits buffer (+ the softmax-statistics block) of a cache, allocated when a call first needs it
int ensure_cache_logits(nlps_model* m, nlps_cache* c) {
  if (c->logits != nullptr) return 0;
  const int64_t n_logits = round_up(c->N * m->ldV, kAlign);
  const int64_t n_stat = round_up(2 * c->N * ((m->V + 31) / 32), kAlign);
  NLPS_CUDA(cudaMalloc(&c->logits_base, sizeof(float) * (size_t)(n_logits + n_stat)));
  c->logits = reinterpret_cast<float*>(c->logits_base);
  c->rowstat = reinterpret_cast<float2*>(c->logits + n_logits);
  return 0;
}

int check_cache(const nlps_model* m, const nlps_cache* c) {
  NLPS_REQUIRE(m && c, "null model or cache");
  NLPS_REQUIRE(c->valid, "cache does not hold a forward pass");
  return 0;
}

}  // namespace

// =========================================================================================
//                                         C ABI
// =========================================================================================
extern "C" {

int nlps_abi_version(void) { return NLPS_ABI_VERSION; }
const char* nlps_last_error(void) { return get_error(); }

int nlps_device_count(int* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { n = 0; cudaGetLastError(); }
  if (count) *count = n;
  return 0;
}

int nlps_device_info(int device, char* name_out, size_t name_cap, int* sm_count, int* cc_major, int* cc_minor,
                     size_t* total_mem) {
  cudaDeviceProp p;
  NLPS_CUDA(cudaGetDeviceProperties(&p, device));
  if (name_out && name_cap) { strncpy(name_out, p.name, name_cap - 1); name_out[name_cap - 1] = 0; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return 0;
}

int nlps_malloc(void** d_ptr, size_t bytes) {
  NLPS_REQUIRE(d_ptr, "nlps_malloc: null out pointer");
  NLPS_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 4));
  return 0;
}
int nlps_free(void* d_ptr) {
  if (d_ptr) NLPS_CUDA(cudaFree(d_ptr));
  return 0;
}
int nlps_memcpy_h2d(void* d, const void* h, size_t bytes, void* stream) {
  if (bytes) NLPS_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
int nlps_memcpy_d2h(void* h, const void* d, size_t bytes, void* stream) {
  if (bytes) NLPS_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  NLPS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
int nlps_memcpy_d2d(void* dd, const void* ds, size_t bytes, void* stream) {
  if (bytes) NLPS_CUDA(cudaMemcpyAsync(dd, ds, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int nlps_memset(void* d, int value, size_t bytes, void* stream) {
  if (bytes) NLPS_CUDA(cudaMemsetAsync(d, value, bytes, (cudaStream_t)stream));
  return 0;
}
int nlps_stream_sync(void* stream) {
  NLPS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
// asynchronous read-back into PINNED host memory + events, so that a loop can read step i's loss while step i+1 runs
int nlps_memcpy_d2h_async(void* h_pinned, const void* d, size_t bytes, void* stream) {
  if (bytes) NLPS_CUDA(cudaMemcpyAsync(h_pinned, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
int nlps_event_create(void** event) {
  NLPS_REQUIRE(event, "nlps_event_create: null argument");
  cudaEvent_t e;
  NLPS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *event = e;
  return 0;
}
int nlps_event_record(void* event, void* stream) {
  NLPS_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return 0;
}
int nlps_event_synchronize(void* event) {
  NLPS_CUDA(cudaEventSynchronize((cudaEvent_t)event));
  return 0;
}
int nlps_event_destroy(void* event) {
  if (event) NLPS_CUDA(cudaEventDestroy((cudaEvent_t)event));
  return 0;
}
int nlps_model_stream(nlps_model* m, void** stream) {
  NLPS_REQUIRE(m && stream, "null argument");
  *stream = m->stream;
  return 0;
}
int nlps_host_alloc_pinned(void** h_ptr, size_t bytes) {
  NLPS_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 4));
  return 0;
}
int nlps_host_free_pinned(void* h_ptr) {
  if (h_ptr) NLPS_CUDA(cudaFreeHost(h_ptr));
  return 0;
}

// ---- model lifetime -----------------------------------------------------------------------
int nlps_create(const nlps_config* cfg, int device, nlps_model** out) {
  NLPS_REQUIRE(cfg && out, "nlps_create: null argument");
  NLPS_REQUIRE(cfg->vocab_size > 0 && cfg->embed_dim > 0 && cfg->num_heads > 0 && cfg->ff_dim > 0 &&
                   cfg->num_layers >= 0 && cfg->max_len > 0, "nlps_create: non-positive dimension");
  NLPS_REQUIRE(cfg->embed_dim % cfg->num_heads == 0, "embed_dim must be divisible by num_heads");
  NLPS_REQUIRE(cfg->embed_dim / cfg->num_heads <= 128, "head_dim > 128 is not supported");
  int ndev = 0;
  nlps_device_count(&ndev);
  NLPS_REQUIRE(ndev > 0, "no CUDA device: libnlps_b200 has no CPU fallback");
  NLPS_REQUIRE(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
  NLPS_CUDA(cudaSetDevice(device));
  ensure_mempool_config();
  nlps_model* m = new nlps_model();
  m->cfg = *cfg;
  m->device = device;
  m->V = cfg->vocab_size; m->E = cfg->embed_dim; m->h = cfg->num_heads; m->d = m->E / m->h;
  m->F = cfg->ff_dim; m->L = cfg->num_layers; m->max_len = cfg->max_len;
  m->precision = cfg->precision;
  m->dropout_p = cfg->dropout_p;
  m->seed = cfg->seed;
  build_layout(m);
  const size_t fb = sizeof(float) * (size_t)m->flat;
  float** bufs[] = {&m->P, &m->G, &m->M, &m->Vv};
  for (float** b : bufs) {
    NLPS_CUDA(cudaMalloc(b, fb));
    NLPS_CUDA(cudaMemset(*b, 0, fb));
  }
  NLPS_CUDA(cudaMalloc(&m->pe, sizeof(float) * (size_t)m->max_len * m->E));
  {  // positional table in f64 then f32 (transformer.py:609-615)
    std::vector<float> pe((size_t)m->max_len * m->E, 0.f);
    for (int pos = 0; pos < m->max_len; ++pos)
      for (int i = 0; i < m->E; i += 2) {
        const double div = std::exp((double)i * -(std::log(10000.0) / (double)m->E));
        pe[(size_t)pos * m->E + i] = (float)std::sin((double)pos * div);
        if (i + 1 < m->E) pe[(size_t)pos * m->E + i + 1] = (float)std::cos((double)pos * div);
      }
    NLPS_CUDA(cudaMemcpy(m->pe, pe.data(), sizeof(float) * pe.size(), cudaMemcpyHostToDevice));
  }
  // one slice per LayerNorm of the stack: the partial sums of a backward are finished on the weight-gradient stream and
  // must survive until then
  NLPS_CUDA(cudaMalloc(&m->ln_partial, sizeof(float) * layernorm_backward_scratch_floats(m->E) * (size_t)std::max(1, 2 * m->L)));
  NLPS_CUDA(cudaMalloc(&m->colsum_scratch, sizeof(float) * colsum_scratch_floats(std::max<int64_t>(std::max(m->E, m->F), m->ldV))));
  NLPS_CUDA(cudaMalloc(&m->colsum_scratch_w, sizeof(float) * colsum_scratch_floats(std::max<int64_t>(std::max(m->E, m->F), m->ldV))));
  NLPS_CUDA(cudaMalloc(&m->colsum_tickets, sizeof(unsigned int) * 8192));
  NLPS_CUDA(cudaMemset(m->colsum_tickets, 0, sizeof(unsigned int) * 8192));
  NLPS_CUDA(cudaMalloc(&m->sumsq_scratch, sizeof(float) * sumsq_scratch_floats()));
  NLPS_CUDA(cudaMalloc(&m->hyper, sizeof(float) * 4));
  NLPS_CUDA(cudaMemset(m->hyper, 0, sizeof(float) * 4));
  NLPS_CUDA(cudaMalloc(&m->scalars, sizeof(float) * 16));
  NLPS_CUDA(cudaMemset(m->scalars, 0, sizeof(float) * 16));
  NLPS_CUDA(cudaMalloc(&m->err_flag, sizeof(int32_t)));
  NLPS_CUDA(cudaMemset(m->err_flag, 0, sizeof(int32_t)));
  {
    // NLPS_GRAPH=0 disables CUDA graphs.  Stream capture is impossible on the legacy default stream, so the model then
    // gets a stream of its own -- a BLOCKING one: it stays ordered with everything the caller does on stream 0.
    const char* e = getenv("NLPS_GRAPH");
    m->graph_mode = (e && e[0] == '0') ? 0 : 1;
    int pri_least = 0, pri_greatest = 0;
    NLPS_CUDA(cudaDeviceGetStreamPriorityRange(&pri_least, &pri_greatest));
    if (m->graph_mode) {
      NLPS_CUDA(cudaStreamCreateWithPriority(&m->own_stream, cudaStreamDefault, pri_greatest));
      m->stream = m->own_stream;
    }
    // NLPS_WGRAD_STREAM=0 keeps the weight gradients on the main stream (A/B measurements, bitwise-equality test)
    const char* w = getenv("NLPS_WGRAD_STREAM");
    if (!(w && w[0] == '0')) {
      const char* pr = getenv("NLPS_WGRAD_PRIO");          // "same": no priority difference between the two streams (A/B)
      NLPS_CUDA(cudaStreamCreateWithPriority(&m->wstream, cudaStreamNonBlocking, (pr && pr[0] == 's') ? pri_greatest : pri_least));
      NLPS_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
      for (auto& e : m->ev_w) NLPS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
  }
  m->bucket_ev.resize(m->buckets.size());
  for (auto& e : m->bucket_ev) NLPS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  NLPS_CUDA(cudaEventCreateWithFlags(&m->aux_ev, cudaEventDisableTiming));
  *out = m;
  return 0;
}

int nlps_destroy(nlps_model* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  for (nlps_cache* c : m->pool) { if (c->base) cudaFree(c->base); if (c->logits_base) cudaFree(c->logits_base); delete c; }
  float* bufs[] = {m->P, m->G, m->M, m->Vv, m->pe, m->g0, m->g1, m->gdrop, m->dctx, m->dqkv, m->dhid, m->delta,
                   m->row_loss, m->rows_tmp, m->ln_partial, m->colsum_scratch, m->sumsq_scratch, m->scalars, m->colsum32,
                   m->g1b, m->gdropb, m->dbv_partial, m->vchunk, m->colsum_scratch_w};
  for (float* b : bufs) if (b) cudaFree(b);
  if (m->iscratch) cudaFree(m->iscratch);
  if (m->err_flag) cudaFree(m->err_flag);
  if (m->colsum_tickets) cudaFree(m->colsum_tickets);
  if (m->hyper) cudaFree(m->hyper);
  drop_step_graphs(m);
  if (m->comm) { comm_destroy(m->comm); m->comm = nullptr; }
  if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
  if (m->comm_done) cudaEventDestroy(m->comm_done);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  if (m->wstream) cudaStreamDestroy(m->wstream);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  for (auto& e : m->ev_w) if (e) cudaEventDestroy(e);
  for (auto& e : m->bucket_ev) cudaEventDestroy(e);
  if (m->aux_ev) cudaEventDestroy(m->aux_ev);
  delete m;
  return 0;
}

int nlps_set_stream(nlps_model* m, void* s) {
  NLPS_REQUIRE(m, "null model");
  drop_step_graphs(m);
  // a null handle means "the default": with graphs enabled that is the model's own blocking stream
  m->stream = (s == nullptr && m->own_stream != nullptr) ? m->own_stream : (cudaStream_t)s;
  return 0;
}
int nlps_set_precision(nlps_model* m, int p) {
  NLPS_REQUIRE(m, "null model");
  NLPS_REQUIRE(p >= 0 && p <= 2, "unknown precision %d", p);
  m->precision = p;
  return 0;
}
int nlps_set_training(nlps_model* m, int t) { NLPS_REQUIRE(m, "null model"); m->training = t != 0; return 0; }
int nlps_set_dropout(nlps_model* m, float p, uint64_t seed) {
  NLPS_REQUIRE(m, "null model");
  NLPS_REQUIRE(p < 1.f, "dropout_p must be < 1");
  m->dropout_p = p; m->seed = seed;
  return 0;
}
int nlps_sync(nlps_model* m) {
  NLPS_REQUIRE(m, "null model");
  NLPS_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

int nlps_flat_size(const nlps_model* m, int64_t* n) { NLPS_REQUIRE(m && n, "null argument"); *n = m->flat; return 0; }
int nlps_flat_ptr(nlps_model* m, int which, void** p) {
  NLPS_REQUIRE(m && p, "null argument");
  float* b[] = {m->P, m->G, m->M, m->Vv};
  NLPS_REQUIRE(which >= 0 && which < 4, "unknown buffer %d", which);
  *p = b[which];
  return 0;
}

int nlps_param_view(nlps_model* m, int which, int id, nlps_view* v) {
  NLPS_REQUIRE(m && v, "null argument");
  void* basev = nullptr;
  NLPS_TRY(nlps_flat_ptr(m, which, &basev));
  float* base = static_cast<float*>(basev);
  const int64_t E = m->E, F = m->F, V = m->V, L = m->L;
  const int64_t ls = m->L > 1 ? m->lo[1].begin - m->lo[0].begin : 0;
  auto set = [&](int64_t off, int nd, int64_t s0, int64_t s1, int64_t s2, int64_t t0, int64_t t1, int64_t t2) {
    v->d_ptr = base + off; v->ndim = nd;
    v->shape[0] = s0; v->shape[1] = s1; v->shape[2] = s2;
    v->stride[0] = t0; v->stride[1] = t1; v->stride[2] = t2;
  };
  const LayerOff zero{};
  const LayerOff& o = m->L > 0 ? m->lo[0] : zero;
  switch (id) {
    case NLPS_P_We: set(m->off_We, 2, V, E, 1, E, 1, 0); break;
    case NLPS_P_Wq: set(o.wqkv, 3, L, E, E, ls, 3 * E, 1); break;
    case NLPS_P_Wk: set(o.wqkv + E, 3, L, E, E, ls, 3 * E, 1); break;
    case NLPS_P_Wv: set(o.wqkv + 2 * E, 3, L, E, E, ls, 3 * E, 1); break;
    case NLPS_P_Wo: set(o.wo, 3, L, E, E, ls, E, 1); break;
    case NLPS_P_W1: set(o.w1, 3, L, E, F, ls, F, 1); break;
    case NLPS_P_W2: set(o.w2, 3, L, F, E, ls, E, 1); break;
    case NLPS_P_b1: set(o.b1, 2, L, F, 1, ls, 1, 0); break;
    case NLPS_P_b2: set(o.b2, 2, L, E, 1, ls, 1, 0); break;
    case NLPS_P_W_vocab: set(m->off_Wvocab, 2, E, V, 1, m->ldV, 1, 0); break;
    case NLPS_P_b_vocab: set(m->off_bvocab, 1, V, 1, 1, 1, 0, 0); break;
    case NLPS_P_ln1_gamma: set(o.ln1g, 2, L, E, 1, ls, 1, 0); break;
    case NLPS_P_ln1_beta: set(o.ln1b, 2, L, E, 1, ls, 1, 0); break;
    case NLPS_P_ln2_gamma: set(o.ln2g, 2, L, E, 1, ls, 1, 0); break;
    case NLPS_P_ln2_beta: set(o.ln2b, 2, L, E, 1, ls, 1, 0); break;
    default: NLPS_REQUIRE(false, "unknown parameter id %d", id);
  }
  return 0;
}

int nlps_pe_ptr(nlps_model* m, void** p) { NLPS_REQUIRE(m && p, "null argument"); *p = m->pe; return 0; }
int nlps_num_buckets(const nlps_model* m, int* n) { NLPS_REQUIRE(m && n, "null argument"); *n = (int)m->buckets.size(); return 0; }
int nlps_bucket_range(const nlps_model* m, int b, int64_t* first, int64_t* count) {
  NLPS_REQUIRE(m && first && count, "null argument");
  NLPS_REQUIRE(b >= 0 && b < (int)m->buckets.size(), "bucket %d out of range", b);
  *first = m->buckets[b].first; *count = m->buckets[b].second;
  return 0;
}
int nlps_stream_wait_bucket(nlps_model* m, int b, void* other) {
  NLPS_REQUIRE(m, "null model");
  NLPS_REQUIRE(b >= 0 && b < (int)m->buckets.size(), "bucket %d out of range", b);
  NLPS_CUDA(cudaStreamWaitEvent((cudaStream_t)other, m->bucket_ev[b], 0));
  return 0;
}
int nlps_wait_stream(nlps_model* m, void* other) {
  NLPS_REQUIRE(m, "null model");
  NLPS_CUDA(cudaEventRecord(m->aux_ev, (cudaStream_t)other));
  NLPS_CUDA(cudaStreamWaitEvent(m->stream, m->aux_ev, 0));
  return 0;
}
int nlps_logits_ld(const nlps_model* m, int64_t* ld) { NLPS_REQUIRE(m && ld, "null argument"); *ld = m->ldV; return 0; }

// ---- caches ------------------------------------------------------------------------------
int nlps_cache_create(nlps_model* m, int B, int S, nlps_cache** out) {
  NLPS_REQUIRE(m && out, "null argument");
  NLPS_REQUIRE(B > 0 && S > 0, "forward needs a non-empty batch (B=%d, S=%d)", B, S);
  NLPS_REQUIRE(S <= m->max_len, "sequence length %d exceeds max_len %d", S, m->max_len);
  for (size_t i = 0; i < m->pool.size(); ++i)
    if (m->pool[i]->B == B && m->pool[i]->S == S) {
      *out = m->pool[i];
      m->pool.erase(m->pool.begin() + i);
      (*out)->valid = false;
      return 0;
    }
  NLPS_CUDA(cudaSetDevice(m->device));
  nlps_cache* c = new nlps_cache();
  c->B = B; c->S = S; c->N = (int64_t)B * S;
  const int64_t N = c->N, E = m->E, F = m->F;
  int64_t cur = 0;   // in floats
  auto carve = [&](int64_t n) { const int64_t at = cur; cur = round_up(cur + n, kAlign); return at; };
  const int64_t o_x = carve(N), o_fnp = carve(B), o_step = carve(1), o_h0 = carve(N * E);
  struct LO { int64_t qkv, ctx, lse, r1, mean1, rstd1, y1, hid, r2, mean2, rstd2, y2, bits; };
  std::vector<LO> lo(m->L);
  for (auto& q : lo) {
    q.qkv = carve(N * 3 * E); q.ctx = carve(N * E); q.lse = carve((int64_t)B * m->h * S);
    q.r1 = carve(N * E); q.mean1 = carve(N); q.rstd1 = carve(N); q.y1 = carve(N * E);
    q.hid = carve(N * F); q.r2 = carve(N * E); q.mean2 = carve(N); q.rstd2 = carve(N); q.y2 = carve(N * E);
    q.bits = carve(N * ((F + 31) / 32));
  }
  c->bytes = (size_t)cur * sizeof(float);
  cudaError_t e = cudaMalloc(&c->base, c->bytes);
  if (e != cudaSuccess) {
    // out of memory: drop pooled caches and retry once
    cudaGetLastError();
    cudaStreamSynchronize(m->stream);
    for (nlps_cache* p : m->pool) { drop_graphs_of_cache(m, p); cudaFree(p->base); if (p->logits_base) cudaFree(p->logits_base); delete p; }
    m->pool.clear();
    e = cudaMalloc(&c->base, c->bytes);
  }
  if (e != cudaSuccess) {
    set_error("cache allocation of %zu bytes failed: %s", c->bytes, cudaGetErrorString(e));
    delete c;
    return 1;
  }
  float* f = reinterpret_cast<float*>(c->base);
  c->x_flat = reinterpret_cast<int32_t*>(f + o_x);
  c->fnp = reinterpret_cast<int32_t*>(f + o_fnp);
  c->dev_step = reinterpret_cast<uint32_t*>(f + o_step);
  c->h0 = f + o_h0;
  c->layers.resize(m->L);
  for (int l = 0; l < m->L; ++l) {
    const LO& q = lo[l];
    c->layers[l] = {f + q.qkv, f + q.ctx, f + q.lse, f + q.r1, f + q.mean1, f + q.rstd1, f + q.y1,
                    f + q.hid, f + q.r2, f + q.mean2, f + q.rstd2, f + q.y2,
                    reinterpret_cast<uint32_t*>(f + q.bits)};
  }
  *out = c;
  return 0;
}

int nlps_cache_release(nlps_model* m, nlps_cache* c) {
  if (!c) return 0;
  NLPS_REQUIRE(m, "null model");
  c->valid = false;
  if (m->pool.size() >= 32) {      // bound the pool (one entry per batch shape; the driver loop has ~10 length buckets)
    nlps_cache* old = m->pool.front();
    m->pool.erase(m->pool.begin());
    cudaStreamSynchronize(m->stream);
    drop_graphs_of_cache(m, old);          // only the graphs that replay into this cache; every other one stays valid
    cudaFree(old->base);
    if (old->logits_base) cudaFree(old->logits_base);
    delete old;
  }
  m->pool.push_back(c);
  return 0;
}

int nlps_cache_hidden_ptr(nlps_cache* c, void** p) {
  NLPS_REQUIRE(c && p, "null argument");
  *p = c->layers.empty() ? c->h0 : c->layers.back().y2;
  return 0;
}
int nlps_cache_layer_ptr(nlps_model* m, nlps_cache* c, int layer, int which, void** p) {
  NLPS_REQUIRE(m && c && p, "null argument");
  NLPS_REQUIRE(layer >= 0 && layer < (int)c->layers.size(), "layer %d out of range", layer);
  const nlps_cache::L& a = c->layers[layer];
  switch (which) {
    case 0: *p = layer == 0 ? c->h0 : c->layers[layer - 1].y2; break;   // attn_input
    case 1: *p = a.y1; break;                                           // ffn_input
    case 2: *p = a.hid; break;                                          // ffn_hidden
    case 3: *p = a.ctx; break;                                          // merged heads of P@V (the reference recomputes it)
    case 4: *p = a.qkv; break;                                          // packed Q|K|V rows
    default: NLPS_REQUIRE(false, "unknown cache tensor %d", which);
  }
  return 0;
}

int nlps_cache_logits_ptr(nlps_cache* c, void** p, int64_t* ld) {
  NLPS_REQUIRE(c && p, "null argument");
  NLPS_REQUIRE(c->logits != nullptr, "this cache holds no logits (hidden-only forward or fused vocabulary path)");
  *p = c->logits;
  (void)ld;
  return 0;
}

// ---- forward -------------------------------------------------------------------------------
int nlps_forward(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, int hidden_only, float* d_out,
                 int64_t ld_out) {
  NLPS_REQUIRE(m && c && d_x, "nlps_forward: null argument");
  NLPS_CUDA(cudaSetDevice(m->device));
  const int E = m->E, F = m->F, N = (int)c->N;
  cudaStream_t st = m->stream;
  const int prec = m->precision;
  c->drop_p = (m->training && m->dropout_p > 0.f) ? m->dropout_p : 0.f;
  c->drop_step = m->fwd_counter++;
  c->has_logits = false;
  m->dbv_partial_rows = 0;             // column-sum partials of an earlier (possibly failed) step are not this step's
  if (!m->in_graph_body && c->drop_p > 0.f) {
    set_step_scalars_kernel<<<1, 1, 0, st>>>(c->dev_step, c->drop_step, nullptr, 0.f, 0.f, 0.f);
    NLPS_LAUNCH_CHECK();
  }
  NLPS_TRY(embed_forward(st, c->B, c->S, E, m->V, d_x, xs, m->P + m->off_We, m->pe, c->h0, c->x_flat, m->err_flag));
  NLPS_TRY(first_nonpad(st, c->B, c->S, c->x_flat, c->fnp));
  const float* x_in = c->h0;
  // 1-bit ReLU masks replace the 4-byte gate read in backward (tensor-core path, F % 32 == 0)
  const bool use_bits = prec == NLPS_PREC_TF32 && F % 32 == 0 && E % 4 == 0;
  c->relu_bits_valid = use_bits;
  for (int l = 0; l < m->L; ++l) {
    const LayerOff& o = m->lo[l];
    nlps_cache::L& a = c->layers[l];
    {  // Q|K|V = x Wqkv                                              (transformer.py:127-129)
      GemmArgs g = gemm_nn(N, 3 * E, E, x_in, E, m->P + o.wqkv, 3 * E, a.qkv, 3 * E);
      NLPS_TRY(gemm_dispatch(st, prec, g));
    }
    NLPS_TRY(attention_forward(st, prec, c->B, c->S, m->h, m->d, a.qkv, c->x_flat, c->fnp, a.ctx, a.lse));
    {  // r1 = x + dropout(ctx Wo)                                    (:142-148)
      GemmArgs g = gemm_nn(N, E, E, a.ctx, E, m->P + o.wo, E, a.r1, E);
      g.drop = drop_for(m, c, l, 0);
      g.residual = x_in; g.ldr = E;
      NLPS_TRY(gemm_dispatch(st, prec, g));
    }
    NLPS_TRY(layernorm_forward(st, N, E, a.r1, m->P + o.ln1g, m->P + o.ln1b, a.y1, a.mean1, a.rstd1));
    {  // hid = relu(y1 W1 + b1)                                      (:153)
      GemmArgs g = gemm_nn(N, F, E, a.y1, E, m->P + o.w1, F, a.hid, F);
      g.bias = m->P + o.b1; g.relu = true;
      if (use_bits) { g.relu_bits_out = a.relu_bits; g.ld_bits = F / 32; }
      NLPS_TRY(gemm_dispatch(st, prec, g));
    }
    {  // r2 = y1 + dropout(hid W2 + b2)                              (:154-159)
      GemmArgs g = gemm_nn(N, E, F, a.hid, F, m->P + o.w2, E, a.r2, E);
      g.bias = m->P + o.b2;
      g.drop = drop_for(m, c, l, 1);
      g.residual = a.y1; g.ldr = E;
      NLPS_TRY(gemm_dispatch(st, prec, g));
    }
    NLPS_TRY(layernorm_forward(st, N, E, a.r2, m->P + o.ln2g, m->P + o.ln2b, a.y2, a.mean2, a.rstd2));
    x_in = a.y2;
  }
  if (!hidden_only) {  // logits = H W_vocab + b_vocab                 (:191)
    NLPS_TRY(ensure_cache_logits(m, c));
    GemmArgs g = gemm_nn(N, m->V, E, x_in, E, m->P + m->off_Wvocab, m->ldV, c->logits, m->ldV);
    g.bias = m->P + m->off_bvocab;
    // NLPS_FUSE_CE=1: softmax statistics come out of this product's epilogue and the loss skips its first pass over
    // the logits.  Off by default: at c4 the 64 extra MUFU/FMNMX per 32-column block cost the K=512 vocab product more
    // (+0.17 ms/step measured, same box) than the saved 1.07 GB read that mostly hits L2 anyway.
    static int fuse_ce = -1;
    if (fuse_ce < 0) { const char* e = getenv("NLPS_FUSE_CE"); fuse_ce = (e && e[0] == '1') ? 1 : 0; }
    c->rowstat_valid = fuse_ce && prec == NLPS_PREC_TF32 && c->rowstat != nullptr && gemm_tf32_colsum_fusable(g);
    if (c->rowstat_valid) g.rowstat_out = c->rowstat;
    NLPS_TRY(gemm_dispatch(st, prec, g));
    c->has_logits = true;
  }
  c->valid = true;
  if (d_out) {
    const int64_t cols = hidden_only ? E : m->V;
    const int64_t sld = hidden_only ? E : m->ldV;
    const int64_t shape[2] = {N, cols}, ss[2] = {sld, 1}, ds[2] = {ld_out, 1};
    NLPS_TRY(strided_copy(st, 4, 2, shape, hidden_only ? x_in : c->logits, ss, d_out, ds));
  }
  return 0;
}

int nlps_loss_dlogits(nlps_model* m, const float* d_logits, int64_t ld, const int32_t* d_y, int64_t rows,
                      float grad_scale, float* d_loss, float* d_dlogits, int64_t ldd) {
  NLPS_REQUIRE(m && d_logits && d_y && d_dlogits, "nlps_loss_dlogits: null argument");
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_TRY(ensure_rows_tmp(m, rows));
  float* loss = d_loss ? d_loss : m->scalars + 3;
  return softmax_xent(m->stream, d_logits, ld, d_y, rows, m->V, grad_scale, loss, d_dlogits, ldd, m->rows_tmp, m->err_flag);
}

int nlps_backward(nlps_model* m, nlps_cache* c, const float* d_dlogits, int64_t ldd) {
  NLPS_TRY(check_cache(m, c));
  NLPS_REQUIRE(d_dlogits, "nlps_backward: null dlogits");
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_TRY(ensure_workspace(m, c->B, c->S));
  const int E = m->E, N = (int)c->N;
  cudaStream_t st = m->stream;
  const float* h_final = m->L > 0 ? c->layers.back().y2 : c->h0;
  {  // dW_vocab = H^T dlogits ; db_vocab                               (transformer.py:352-353)
    GemmArgs g = gemm_nn(E, m->V, N, h_final, E, d_dlogits, ldd, m->G + m->off_Wvocab, m->ldV);
    g.trans_a = true;
    cudaStream_t ws = st;
    NLPS_TRY(wgrad_fork(m, &ws));                      // off the critical path: weight-gradient stream
    // db_vocab first: a bandwidth-bound pass over the dlogits that now runs next to the tensor-bound dH product
    {
      const bool side = side_sums();
      cudaStream_t cs = side ? ws : st;
      float* scratch = side ? m->colsum_scratch_w : m->colsum_scratch;
      unsigned int* tickets = side ? colsum_tickets_w(m, m->V) : colsum_tickets_for(m, m->V);
      if (m->dbv_partial_rows > 0)     // the cross-entropy pass already summed 16-row slabs of these dlogits
        NLPS_TRY(colsum(cs, m->dbv_partial_rows, m->V, m->dbv_partial, m->ldV, m->G + m->off_bvocab, scratch, tickets));
      else
        NLPS_TRY(colsum(cs, N, m->V, d_dlogits, ldd, m->G + m->off_bvocab, scratch, tickets));
    }
    m->dbv_partial_rows = 0;
    NLPS_TRY(gemm_dispatch(ws, m->precision, g));
    NLPS_TRY(wgrad_done(m, W_VOCAB));
  }
  {  // dH = dlogits W_vocab^T                                          (:355)
    GemmArgs g = gemm_nn(N, E, m->V, d_dlogits, ldd, m->P + m->off_Wvocab, m->ldV, m->g0, E);
    g.trans_b = true;
    NLPS_TRY(gemm_dispatch(st, m->precision, g));
  }
  NLPS_TRY(record_bucket(m, 0, /*on_wstream=*/true));
  return backward_stack(m, c);
}

int nlps_backward_last_token(nlps_model* m, nlps_cache* c, const int32_t* d_t, const float* d_dl, int64_t ldd) {
  NLPS_TRY(check_cache(m, c));
  NLPS_REQUIRE(d_t && d_dl, "nlps_backward_last_token: null argument");
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_TRY(ensure_workspace(m, c->B, c->S));
  const int E = m->E, B = c->B;
  cudaStream_t st = m->stream;
  const float* h_final = m->L > 0 ? c->layers.back().y2 : c->h0;
  // H_last = H[arange(B), t]  (B,E)  and the (B,E) gradient rows
  NLPS_TRY(ensure_rows_tmp(m, (int64_t)2 * B * E));
  float* h_last = m->rows_tmp;
  float* dh_rows = m->rows_tmp + (int64_t)B * E;
  NLPS_TRY(gather_last(st, B, c->S, E, h_final, d_t, h_last));
  {  // dW_vocab = H_last^T dlogits_last ; db_vocab                     (transformer.py:235-236)
    GemmArgs g = gemm_nn(E, m->V, B, h_last, E, d_dl, ldd, m->G + m->off_Wvocab, m->ldV);
    g.trans_a = true;
    NLPS_TRY(gemm_dispatch(st, m->precision, g));
    NLPS_TRY(colsum(st, B, m->V, d_dl, ldd, m->G + m->off_bvocab, m->colsum_scratch, colsum_tickets_for(m, m->V)));
  }
  {  // dH[arange(B), t] = dlogits_last W_vocab^T, zero elsewhere       (:239-240)
    GemmArgs g = gemm_nn(B, E, m->V, d_dl, ldd, m->P + m->off_Wvocab, m->ldV, dh_rows, E);
    g.trans_b = true;
    NLPS_TRY(gemm_dispatch(st, m->precision, g));
    NLPS_TRY(scatter_rows(st, B, c->S, E, d_t, dh_rows, m->g0));
  }
  NLPS_TRY(record_bucket(m, 0));
  return backward_stack(m, c);
}

int nlps_clip_compute(nlps_model* m, float max_norm) {
  NLPS_REQUIRE(m, "null model");
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_TRY(sumsq(m->stream, m->flat, m->G, m->scalars + 0, m->sumsq_scratch));
  return clip_scale_from_sumsq(m->stream, m->scalars + 0, max_norm, m->scalars + 1, m->scalars + 2);
}

int nlps_clip(nlps_model* m, float max_norm, float* h_norm_out) {
  NLPS_TRY(nlps_clip_compute(m, max_norm));
  NLPS_TRY(scale_inplace(m->stream, m->flat, m->G, m->scalars + 1));
  if (h_norm_out) NLPS_TRY(nlps_memcpy_d2h(h_norm_out, m->scalars + 2, sizeof(float), m->stream));
  return 0;
}

int nlps_step(nlps_model* m, float lr, int use_clip_scale) {
  NLPS_REQUIRE(m, "null model");
  NLPS_CUDA(cudaSetDevice(m->device));
  const float* cs = use_clip_scale ? m->scalars + 1 : nullptr;
  const float* hyper = m->in_graph_body ? m->hyper : nullptr;    // inside a captured step the values come from the setter node
  if (!m->cfg.use_adam) return sgd_update(m->stream, m->flat, m->P, m->G, lr, cs, hyper);
  m->adam_t += 1;
  const double b1 = 0.9, b2 = 0.999;                       // transformer.py:29-31
  float bc1, bc2;
  adam_bias_corrections(m->adam_t, &bc1, &bc2);
  return adam_update(m->stream, m->flat, m->P, m->G, m->M, m->Vv, lr, (float)b1, (float)b2, (float)(1.0 - b1),
                     (float)(1.0 - b2), 1e-8f, bc1, bc2, cs, hyper);
}

int nlps_adam_t(nlps_model* m, int64_t* t, int set, int64_t value) {
  NLPS_REQUIRE(m, "null model");
  if (set) m->adam_t = value;
  if (t) *t = m->adam_t;
  return 0;
}

// Fused vocabulary head + cross-entropy + vocabulary backward that never materialises the (N, V) logits / dlogits (SURVEY 8f #1;
// transformer.py:191, 515-549, 352-355).  The rows are walked in chunks small enough for the chunk's logits to stay in L2:
//   logits_c = H_c W_vocab + b_vocab  ->  dlogits_c in place (per-row losses aside)  ->  db partial_c = column sums,
//   dW_vocab (+)= H_c^T dlogits_c,  dH_c = dlogits_c W_vocab^T
// so the 5 passes over an (N, V) matrix that the materialising path pays in HBM (write logits, read + write in the loss,
// read for dW_vocab, read for dH) happen in L2, and the activation cache loses its largest tensor (1.07 GB at c4, 4.3 GB
// at V = 32768).  Same kernels and the same per-row arithmetic as the materialising path; dW_vocab and db_vocab are summed
// chunk by chunk in a fixed order (deterministic; not the bits of the one-shot sums).
static int vocab_chunk_rows(const nlps_model* m, int64_t N) {
  static int64_t mb = -1;
  if (mb < 0) { const char* e = getenv("NLPS_VOCAB_CHUNK_MB"); mb = e ? std::max(1, atoi(e)) : 48; }
  int64_t rows = (mb << 20) / (m->ldV * 4);
  rows = std::max<int64_t>(256, rows / 256 * 256);
  return (int)std::min<int64_t>(rows, N);
}

static int fused_vocab_head(nlps_model* m, nlps_cache* c, const int32_t* d_y, float grad_scale, float* d_loss) {
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_TRY(ensure_workspace(m, c->B, c->S));
  NLPS_TRY(ensure_rows_tmp(m, c->N));
  const int64_t N = c->N, E = m->E;
  const int rows = vocab_chunk_rows(m, N);
  if (m->vchunk_rows < rows) {
    NLPS_CUDA(cudaStreamSynchronize(m->stream));
    if (m->vchunk) cudaFree(m->vchunk);
    bump_alloc_epoch();
    m->vchunk = nullptr;
    NLPS_CUDA(cudaMalloc(&m->vchunk, sizeof(float) * (size_t)rows * m->ldV));
    m->vchunk_rows = rows;
  }
  cudaStream_t st = m->stream;
  const float* h_final = m->L > 0 ? c->layers.back().y2 : c->h0;
  const float inv = (float)((double)grad_scale / (double)N);
  int k = 0;
  for (int64_t r0 = 0; r0 < N; r0 += rows, ++k) {
    const int n = (int)std::min<int64_t>(rows, N - r0);
    const float* h_c = h_final + r0 * E;
    {  // logits_c = H_c W_vocab + b_vocab                                 (transformer.py:191)
      GemmArgs g = gemm_nn(n, m->V, (int)E, h_c, E, m->P + m->off_Wvocab, m->ldV, m->vchunk, m->ldV);
      g.bias = m->P + m->off_bvocab;
      NLPS_TRY(gemm_dispatch(st, m->precision, g));
    }
    // dlogits_c in place, per-row losses aside                          (:515-549)
    NLPS_TRY(softmax_xent_rows(st, m->vchunk, m->ldV, d_y + r0, n, m->V, inv, m->vchunk, m->ldV, m->rows_tmp + r0, m->err_flag));
    // db_vocab partial of this chunk                                    (:353)
    NLPS_TRY(colsum(st, n, m->V, m->vchunk, m->ldV, m->dbv_partial + (int64_t)k * m->ldV, m->colsum_scratch, colsum_tickets_for(m, m->V)));
    {  // dW_vocab (+)= H_c^T dlogits_c                                    (:352)
      GemmArgs g = gemm_nn((int)E, m->V, n, h_c, E, m->vchunk, m->ldV, m->G + m->off_Wvocab, m->ldV);
      g.trans_a = true;
      g.accumulate = k > 0;
      NLPS_TRY(gemm_dispatch(st, m->precision, g));
    }
    {  // dH_c = dlogits_c W_vocab^T                                       (:355)
      GemmArgs g = gemm_nn(n, (int)E, m->V, m->vchunk, m->ldV, m->P + m->off_Wvocab, m->ldV, m->g0 + r0 * E, E);
      g.trans_b = true;
      NLPS_TRY(gemm_dispatch(st, m->precision, g));
    }
  }
  NLPS_TRY(row_loss_mean(st, m->rows_tmp, N, d_loss ? d_loss : m->scalars + 3));
  NLPS_TRY(colsum(st, k, m->V, m->dbv_partial, m->ldV, m->G + m->off_bvocab, m->colsum_scratch, colsum_tickets_for(m, m->V)));
  NLPS_TRY(record_bucket(m, 0));
  return 0;
}

// NLPS_FUSE_VOCAB=1 selects the fused vocabulary path.  OPT-IN: it removes the (N, V) tensor from memory (1.07 GB at c4,
// 4.3 GB at V = 32768) but same-box A/Bs measured it slower at every configuration -- c3 2.00 -> 2.32 ms, c4 13.18 -> 13.49,
// c4 with V 32768 20.2 -> 22.3, c5 27.2 -> 27.5 with 96 MB chunks, worse with 64 / 32 MB chunks that do stay in L2: what the
// chunks cost (partial waves of the 256x256 tiles on 1-3 k rows, split-K for the short dH products, the dW_vocab
// accumulator read and written once per chunk) exceeds what the L2-resident passes save.
static bool fused_vocab_wanted(const nlps_model*, const nlps_cache*) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NLPS_FUSE_VOCAB"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// the fused step as it always ran: forward, loss (dlogits in place), backward, clip, update
static int train_step_body(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_y, float lr,
                           float max_norm, float grad_scale, int defer_update, float* d_loss) {
  static int fuse_ce_on = -1;
  if (fuse_ce_on < 0) { const char* e = getenv("NLPS_FUSE_CE"); fuse_ce_on = (e && e[0] == '1') ? 1 : 0; }
  if (fused_vocab_wanted(m, c) && !fuse_ce_on && c->N > 0) {
    NLPS_TRY(nlps_forward(m, c, d_x, xs, 1, nullptr, 0));
    NLPS_TRY(fused_vocab_head(m, c, d_y, grad_scale, d_loss));
    c->has_logits = false;
    c->rowstat_valid = false;
    NLPS_TRY(backward_stack(m, c));
    if (defer_update) return 0;
    if (m->comm_active) NLPS_TRY(comm_join(m));
    if (max_norm > 0.f) NLPS_TRY(nlps_clip_compute(m, max_norm));
    return nlps_step(m, lr, max_norm > 0.f ? 1 : 0);
  }
  NLPS_TRY(nlps_forward(m, c, d_x, xs, 0, nullptr, 0));
  // dlogits overwrite the logits in place: 8V bytes of HBM traffic per token for the loss
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_TRY(ensure_rows_tmp(m, c->N));
  // NLPS_FUSE_DBV=1: the vocabulary-bias gradient rides on the cross-entropy pass.  OPT-IN: same-box A/B at c4 measured it
  // SLOWER (13.65 vs 13.50 ms/step) -- a CTA that walks 16 rows with two block barriers per row hides less latency than
  // 16 one-row CTAs, which costs the bandwidth-bound pass more than the saved 1.07 GB column-sum read returns
  static int fuse_dbv = -1;
  if (fuse_dbv < 0) { const char* e = getenv("NLPS_FUSE_DBV"); fuse_dbv = (e && e[0] == '1') ? 1 : 0; }
  if (fuse_dbv && !c->rowstat_valid && xent_colsum_fusable(c->logits, m->ldV, c->logits, m->ldV, m->V)) {
    NLPS_TRY(ensure_workspace(m, c->B, c->S));
    NLPS_TRY(softmax_xent_colsum(m->stream, c->logits, m->ldV, d_y, c->N, m->V, grad_scale, d_loss ? d_loss : m->scalars + 3,
                                 c->logits, m->ldV, m->rows_tmp, m->dbv_partial, m->ldV, m->err_flag));
    m->dbv_partial_rows = xent_colsum_partial_rows(c->N);
  } else {
    NLPS_TRY(softmax_xent_stats(m->stream, c->logits, m->ldV, d_y, c->N, m->V, grad_scale, d_loss ? d_loss : m->scalars + 3, c->logits,
                                m->ldV, m->rows_tmp, c->rowstat_valid ? c->rowstat : nullptr, m->err_flag));
  }
  c->has_logits = false;
  c->rowstat_valid = false;
  NLPS_TRY(nlps_backward(m, c, c->logits, m->ldV));
  if (defer_update) return 0;
  if (m->comm_active) NLPS_TRY(comm_join(m));          // gradients are now the sum over ranks
  if (max_norm > 0.f) NLPS_TRY(nlps_clip_compute(m, max_norm));
  return nlps_step(m, lr, max_norm > 0.f ? 1 : 0);
}

static int train_step_graphed(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_t,
                              const int32_t* d_y, float lr, float max_norm, float grad_scale, int defer_update, float* d_loss);
static int last_token_body(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_t, const int32_t* d_y,
                           float lr, float max_norm, float grad_scale, int defer_update, float* d_loss);
// the eager form of either fused step: full-sequence objective (d_t == nullptr) or the driver's last-token objective
static int run_body(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_t, const int32_t* d_y,
                    float lr, float max_norm, float grad_scale, int defer_update, float* d_loss) {
  if (d_t != nullptr) return last_token_body(m, c, d_x, xs, d_t, d_y, lr, max_norm, grad_scale, defer_update, d_loss);
  return train_step_body(m, c, d_x, xs, d_y, lr, max_norm, grad_scale, defer_update, d_loss);
}

int nlps_train_step(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_y, float lr,
                    float max_norm, float grad_scale, int defer_update, float* d_loss) {
  NLPS_REQUIRE(m && c && d_x && d_y, "nlps_train_step: null argument");
  // with a communicator attached (nlps_comm_init) the undeferred step is the data-parallel step: the gradient buckets are
  // summed over ranks on the communication stream as backward completes them, and clip + update wait for the sums
  m->comm_active = m->comm != nullptr && !defer_update;
  const int rc = train_step_graphed(m, c, d_x, xs, nullptr, d_y, lr, max_norm, grad_scale, defer_update, d_loss);
  m->comm_active = false;
  return rc;
}

static int train_step_graphed(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_t,
                              const int32_t* d_y, float lr, float max_norm, float grad_scale, int defer_update, float* d_loss) {
  // CUDA graph path, every arithmetic mode (the 3xTF32 operands live in a grow-only per-stream workspace).  First sighting
  // of a signature runs eagerly (it sizes every workspace), the second one is captured, later ones are replayed with one
  // patched node.
  static const bool dp_graph = [] { const char* e = getenv("NLPS_GRAPH_DP"); return !(e && e[0] == '0'); }();
  const bool graphable = m->graph_mode == 1 && (!defer_update || dp_graph) && m->stream != nullptr;
  const bool deferred = defer_update != 0;
  if (!graphable) { m->n_eager++; return run_body(m, c, d_x, xs, d_t, d_y, lr, max_norm, grad_scale, defer_update, d_loss); }
  NLPS_CUDA(cudaSetDevice(m->device));
  const float drop_p = (m->training && m->dropout_p > 0.f) ? m->dropout_p : 0.f;
  nlps_model::StepGraph* e = nullptr;
  for (size_t i = 0; i < m->graphs.size();) {
    nlps_model::StepGraph& g = m->graphs[i];
    if (g.epoch != alloc_epoch()) {            // something it points into was re-allocated
      if (g.exec) cudaGraphExecDestroy(g.exec);
      if (g.graph) cudaGraphDestroy(g.graph);
      m->graphs.erase(m->graphs.begin() + i);
      continue;
    }
    if (g.cache == c && g.x == d_x && g.xs == xs && g.y == d_y && g.t == d_t && g.max_norm == max_norm && g.grad_scale == grad_scale &&
        g.dropout_p == drop_p && g.precision == m->precision && g.training == m->training && g.deferred == deferred) e = &g;
    ++i;
  }
  if (e == nullptr) {
    m->n_eager++;
    const int rc = run_body(m, c, d_x, xs, d_t, d_y, lr, max_norm, grad_scale, defer_update, d_loss);
    if (rc != 0) return rc;
    if (m->graphs.size() >= 48) {              // bounded: forget the oldest signature (the driver loop has one per length bucket)
      if (m->graphs.front().exec) cudaGraphExecDestroy(m->graphs.front().exec);
      if (m->graphs.front().graph) cudaGraphDestroy(m->graphs.front().graph);
      m->graphs.erase(m->graphs.begin());
    }
    nlps_model::StepGraph g{};
    g.cache = c; g.x = d_x; g.xs = xs; g.y = d_y; g.t = d_t; g.max_norm = max_norm; g.grad_scale = grad_scale; g.dropout_p = drop_p;
    g.precision = m->precision; g.training = m->training; g.deferred = deferred; g.epoch = alloc_epoch(); g.state = 1;
    m->graphs.push_back(g);
    return 0;
  }
  if (e->state == 3) { m->n_eager++; return run_body(m, c, d_x, xs, d_t, d_y, lr, max_norm, grad_scale, defer_update, d_loss); }
  // this step's scalars
  const uint32_t step = m->fwd_counter;
  float bc1 = 1.f, bc2 = 1.f;
  if (m->cfg.use_adam) adam_bias_corrections(m->adam_t + 1, &bc1, &bc2);   // unused by a deferred graph (no update inside)
  uint32_t* a_slot = c->dev_step; uint32_t a_step = step; float* a_hyper = m->hyper; float a_lr = lr, a_bc1 = bc1, a_bc2 = bc2;
  void* setter_args[6] = {&a_slot, &a_step, &a_hyper, &a_lr, &a_bc1, &a_bc2};
  float* loss_slot = m->scalars + 3;
  if (e->state == 1) {
    // ---- capture ----
    cudaStream_t st = m->stream;
    const long long launches_before = launch_count(false);
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      e->state = 3;
      return run_body(m, c, d_x, xs, d_t, d_y, lr, max_norm, grad_scale, defer_update, d_loss);
    }
    m->in_graph_body = true;
    m->graph_external_events = deferred;
    set_step_scalars_kernel<<<1, 1, 0, st>>>(a_slot, a_step, a_hyper, a_lr, a_bc1, a_bc2);
    count_launch();
    const int rc = run_body(m, c, d_x, xs, d_t, d_y, lr, max_norm, grad_scale, defer_update, loss_slot);
    m->in_graph_body = false;
    m->graph_external_events = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    const unsigned long long epoch_after = alloc_epoch();
    bool ok = rc == 0 && ce == cudaSuccess && graph != nullptr && epoch_after == e->epoch;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t setter = nullptr;
    if (ok) {
      size_t n_nodes = 0;
      ok = cudaGraphGetNodes(graph, nullptr, &n_nodes) == cudaSuccess && n_nodes > 0;
      std::vector<cudaGraphNode_t> nodes(n_nodes);
      if (ok) ok = cudaGraphGetNodes(graph, nodes.data(), &n_nodes) == cudaSuccess;
      for (size_t i = 0; ok && i < n_nodes && setter == nullptr; ++i) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp{};
        if (cudaGraphKernelNodeGetParams(nodes[i], &kp) == cudaSuccess && kp.func == (void*)set_step_scalars_kernel) setter = nodes[i];
      }
      ok = ok && setter != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    }
    if (!ok) {
      // The host-side bookkeeping of the captured body has happened (counters advanced) but no kernel ran: wind it back
      // and run this step eagerly; the signature stays eager from now on.
      cudaGetLastError();
      if (exec) cudaGraphExecDestroy(exec);
      if (graph) cudaGraphDestroy(graph);
      if (rc == 0) { m->fwd_counter -= 1; if (m->cfg.use_adam && !deferred) m->adam_t -= 1; }
      // e may dangle if the vector changed: look the entry up again
      for (auto& g : m->graphs) if (g.cache == c && g.x == d_x && g.y == d_y && g.t == d_t && g.state == 1) g.state = 3;
      return run_body(m, c, d_x, xs, d_t, d_y, lr, max_norm, grad_scale, defer_update, d_loss);
    }
    e->graph = graph; e->exec = exec; e->setter = setter; e->state = 2;
    m->n_captured++;
    e->launches = launch_count(false) - launches_before;
    e->relu_bits_valid = c->relu_bits_valid;
    // nothing has executed yet: this step's work is the first launch of the graph
    NLPS_CUDA(cudaGraphLaunch(exec, st));
    if (d_loss != nullptr) NLPS_CUDA(cudaMemcpyAsync(d_loss, loss_slot, sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  // ---- replay ----
  {
    cudaKernelNodeParams kp{};
    kp.func = (void*)set_step_scalars_kernel;
    kp.gridDim = dim3(1, 1, 1); kp.blockDim = dim3(1, 1, 1); kp.sharedMemBytes = 0;
    kp.kernelParams = setter_args; kp.extra = nullptr;
    m->n_replayed++;
    NLPS_CUDA(cudaGraphExecKernelNodeSetParams(e->exec, e->setter, &kp));
    NLPS_CUDA(cudaGraphLaunch(e->exec, m->stream));
    if (d_loss != nullptr) NLPS_CUDA(cudaMemcpyAsync(d_loss, loss_slot, sizeof(float), cudaMemcpyDeviceToDevice, m->stream));
    // the host-side state the eager body leaves behind
    c->drop_p = drop_p;
    c->drop_step = m->fwd_counter++;
    c->valid = true;
    c->has_logits = false;
    c->rowstat_valid = false;
    c->relu_bits_valid = e->relu_bits_valid;
    if (m->cfg.use_adam && !deferred) m->adam_t += 1;
    count_launch((int)e->launches);
  }
  return 0;
}

int nlps_train_step_last_token(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_t,
                               const int32_t* d_y, float lr, float max_norm, float grad_scale, int defer_update,
                               float* d_loss) {
  NLPS_REQUIRE(m && c && d_x && d_t && d_y, "nlps_train_step_last_token: null argument");
  // like nlps_train_step: data-parallel when a communicator is attached, replayed as a CUDA graph from the third sighting
  // of a (cache, x, t, y) signature on (the driver loop stages its batches into fixed buffers per length bucket)
  m->comm_active = m->comm != nullptr && !defer_update;
  const int rc = train_step_graphed(m, c, d_x, xs, d_t, d_y, lr, max_norm, grad_scale, defer_update, d_loss);
  m->comm_active = false;
  return rc;
}

static int last_token_body(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_t, const int32_t* d_y,
                           float lr, float max_norm, float grad_scale, int defer_update, float* d_loss) {
  NLPS_TRY(nlps_forward(m, c, d_x, xs, 1, nullptr, 0));
  const int B = c->B, E = m->E;
  const float* h_final = m->L > 0 ? c->layers.back().y2 : c->h0;
  // H_last (B,E) then logits_last (B,ldV) live in the row scratch, behind backward_last_token's own use of it
  const int64_t off_h = round_up(2 * (int64_t)B * E + B, kAlign);
  const int64_t off_l = off_h + round_up((int64_t)B * E, kAlign);
  NLPS_TRY(ensure_rows_tmp(m, off_l + (int64_t)B * m->ldV));
  float* h_last = m->rows_tmp + off_h;
  float* logits = m->rows_tmp + off_l;
  NLPS_TRY(gather_last(m->stream, B, c->S, E, h_final, d_t, h_last));
  {
    GemmArgs g = gemm_nn(B, m->V, E, h_last, E, m->P + m->off_Wvocab, m->ldV, logits, m->ldV);
    g.bias = m->P + m->off_bvocab;
    NLPS_TRY(gemm_dispatch(m->stream, m->precision, g));
  }
  float* loss = d_loss ? d_loss : m->scalars + 3;
  // row-loss scratch: the first B floats of rows_tmp are free until backward_last_token runs
  NLPS_TRY(softmax_xent(m->stream, logits, m->ldV, d_y, B, m->V, grad_scale, loss, logits, m->ldV, m->rows_tmp, m->err_flag));
  NLPS_TRY(nlps_backward_last_token(m, c, d_t, logits, m->ldV));
  if (defer_update) return 0;
  if (m->comm_active) NLPS_TRY(comm_join(m));
  if (max_norm > 0.f) NLPS_TRY(nlps_clip_compute(m, max_norm));
  return nlps_step(m, lr, max_norm > 0.f ? 1 : 0);
}

// ---- data-parallel exchange ------------------------------------------------------------------------------
int nlps_comm_available(int* nccl_version) { return comm_available(nccl_version) ? 0 : (set_error("libnccl.so.2 could not be loaded"), 1); }

int nlps_comm_unique_id(void* h_id128) {
  NLPS_REQUIRE(h_id128, "nlps_comm_unique_id: null argument");
  return comm_unique_id(h_id128);
}

int nlps_comm_init(nlps_model* m, const void* h_id128, int rank, int world, int64_t min_bucket_floats) {
  NLPS_REQUIRE(m && h_id128, "nlps_comm_init: null argument");
  NLPS_REQUIRE(m->comm == nullptr, "nlps_comm_init: the model already has a communicator");
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_TRY(comm_create(h_id128, rank, world, &m->comm));
  m->comm_rank = rank; m->comm_world = world;
  int pri_least = 0, pri_greatest = 0;
  NLPS_CUDA(cudaDeviceGetStreamPriorityRange(&pri_least, &pri_greatest));
  // NLPS_COMM_PRIO=low|high (A/B): with high priority NCCL's CTAs take the first SMs that free up (earliest exchange);
  // with low priority they fill SMs the compute streams leave idle
  const char* cp = getenv("NLPS_COMM_PRIO");
  NLPS_CUDA(cudaStreamCreateWithPriority(&m->comm_stream, cudaStreamNonBlocking, (cp && cp[0] == 'l') ? pri_least : pri_greatest));
  NLPS_CUDA(cudaEventCreateWithFlags(&m->comm_done, cudaEventDisableTiming));
  // coalesce consecutive buckets (completion order) into exchanges of at least min_bucket_floats: NVSwitch collectives
  // are latency-, not link-bound, so small buckets ride with their neighbours
  m->comm_plan.clear();
  int start = 0; int64_t acc = 0;
  for (int b = 0; b < (int)m->buckets.size(); ++b) {
    acc += m->buckets[b].second;
    if (acc >= min_bucket_floats || b == (int)m->buckets.size() - 1) { m->comm_plan.push_back({start, b + 1}); start = b + 1; acc = 0; }
  }
  drop_step_graphs(m);
  return 0;
}

int nlps_comm_destroy(nlps_model* m) {
  NLPS_REQUIRE(m, "null model");
  if (m->comm == nullptr) return 0;
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_CUDA(cudaStreamSynchronize(m->stream));
  NLPS_CUDA(cudaStreamSynchronize(m->comm_stream));
  drop_step_graphs(m);
  NLPS_TRY(comm_destroy(m->comm));
  m->comm = nullptr; m->comm_world = 1; m->comm_rank = 0;
  m->comm_plan.clear();
  return 0;
}

int nlps_comm_info(const nlps_model* m, int* rank, int* world, int* n_exchanges) {
  NLPS_REQUIRE(m, "null model");
  if (rank) *rank = m->comm_rank;
  if (world) *world = m->comm ? m->comm_world : 0;
  if (n_exchanges) *n_exchanges = (int)m->comm_plan.size();
  return 0;
}

// sum of the whole flat gradient buffer over ranks, after everything enqueued on the model's stream (the unfused API:
// forward / backward / [this] / clip / step); the model's stream continues after the sum
int nlps_comm_allreduce_grads(nlps_model* m) {
  NLPS_REQUIRE(m && m->comm, "nlps_comm_allreduce_grads: no communicator (nlps_comm_init)");
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_CUDA(cudaEventRecord(m->aux_ev, m->stream));
  NLPS_CUDA(cudaStreamWaitEvent(m->comm_stream, m->aux_ev, 0));
  NLPS_TRY(comm_allreduce_sum(m->comm, m->G, (size_t)m->flat, m->comm_stream));
  return comm_join(m);
}

// identical replicas: rank `root`'s parameters (and Adam moments) everywhere
int nlps_comm_broadcast_params(nlps_model* m, int root) {
  NLPS_REQUIRE(m && m->comm, "nlps_comm_broadcast_params: no communicator (nlps_comm_init)");
  NLPS_REQUIRE(root >= 0 && root < m->comm_world, "root %d out of range", root);
  NLPS_CUDA(cudaSetDevice(m->device));
  NLPS_CUDA(cudaEventRecord(m->aux_ev, m->stream));
  NLPS_CUDA(cudaStreamWaitEvent(m->comm_stream, m->aux_ev, 0));
  float* bufs[] = {m->P, m->M, m->Vv};
  for (float* b : bufs) NLPS_TRY(comm_broadcast(m->comm, b, (size_t)m->flat, root, m->comm_stream));
  return comm_join(m);
}

int nlps_tf32_mma_peak(int device, int cta_group, int iters, int reps, float* h_tflops, float* h_ms) {
  int ndev = 0;
  nlps_device_count(&ndev);
  NLPS_REQUIRE(ndev > 0, "no CUDA device: libnlps_b200 has no CPU fallback");
  NLPS_CUDA(cudaSetDevice(device));
  return tf32_mma_peak(cta_group, iters, reps, h_tflops, h_ms);
}

int nlps_kernel_stats(int64_t* counts, int n, int reset) {
  NLPS_REQUIRE(counts && n >= 0, "nlps_kernel_stats: null argument");
  for (int k = 0; k < n; ++k) counts[k] = kind_count(k, reset != 0);
  return 0;
}

int nlps_check_errors(nlps_model* m, int* flags) {
  NLPS_REQUIRE(m && flags, "nlps_check_errors: null argument");
  NLPS_CUDA(cudaSetDevice(m->device));
  int32_t word = 0;
  NLPS_CUDA(cudaMemcpyAsync(&word, m->err_flag, sizeof(word), cudaMemcpyDeviceToHost, m->stream));
  NLPS_CUDA(cudaStreamSynchronize(m->stream));
  if (word != 0) NLPS_CUDA(cudaMemsetAsync(m->err_flag, 0, sizeof(word), m->stream));
  *flags = word;
  return 0;
}

int nlps_dropout_counter(nlps_model* m, int64_t* value, int set, int64_t new_value) {
  NLPS_REQUIRE(m, "null model");
  if (set) {
    NLPS_REQUIRE(new_value >= 0 && new_value <= 0xFFFFFFFFll, "dropout counter %lld out of range", (long long)new_value);
    m->fwd_counter = (uint32_t)new_value;
  }
  if (value) *value = m->fwd_counter;
  return 0;
}

int nlps_dropout_mask(nlps_model* m, int layer, int which, int64_t step, int64_t n, float* d_out) {
  NLPS_REQUIRE(m && d_out, "nlps_dropout_mask: null argument");
  NLPS_REQUIRE(layer >= 0 && layer < m->L && (which == 0 || which == 1), "nlps_dropout_mask: no dropout site (layer %d, branch %d)", layer, which);
  NLPS_REQUIRE(step >= 0 && step <= 0xFFFFFFFFll, "nlps_dropout_mask: step out of range");
  NLPS_CUDA(cudaSetDevice(m->device));
  const DropoutParams d = make_dropout(m->dropout_p, m->seed, (uint32_t)(layer * 2 + which), (uint32_t)step);
  return dropout_mask(m->stream, d, n, d_out);
}

// ---- evaluation pass of the driver loop without host synchronisation (utils/train.py:51-75) ------------------------
int nlps_eval_last_token(nlps_model* m, nlps_cache* c, const int32_t* d_x, int64_t xs, const int32_t* d_t,
                         const int32_t* d_y, float* d_accum) {
  NLPS_REQUIRE(m && c && d_x && d_t && d_y && d_accum, "nlps_eval_last_token: null argument");
  const bool was_training = m->training;
  m->training = false;                                    // model.eval(): dropout off (utils/train.py:54)
  const int rc = nlps_forward(m, c, d_x, xs, 1, nullptr, 0);
  m->training = was_training;
  if (rc != 0) return rc;
  const int B = c->B, E = m->E;
  const float* h_final = m->L > 0 ? c->layers.back().y2 : c->h0;
  const int64_t off_h = round_up((int64_t)B, kAlign);
  const int64_t off_l = off_h + round_up((int64_t)B * E, kAlign);
  NLPS_TRY(ensure_rows_tmp(m, off_l + (int64_t)B * m->ldV));
  float* h_last = m->rows_tmp + off_h;
  float* logits = m->rows_tmp + off_l;
  NLPS_TRY(gather_last(m->stream, B, c->S, E, h_final, d_t, h_last));
  GemmArgs g = gemm_nn(B, m->V, E, h_last, E, m->P + m->off_Wvocab, m->ldV, logits, m->ldV);
  g.bias = m->P + m->off_bvocab;
  NLPS_TRY(gemm_dispatch(m->stream, m->precision, g));
  NLPS_TRY(softmax_xent_rows(m->stream, logits, m->ldV, d_y, B, m->V, 1.f, logits, m->ldV, m->rows_tmp, m->err_flag));
  return eval_accumulate(m->stream, m->rows_tmp, B, d_accum);
}

int nlps_eval_finish(nlps_model* m, float* d_accum, float* d_sched) {
  NLPS_REQUIRE(m && d_accum && d_sched, "nlps_eval_finish: null argument");
  return eval_finish(m->stream, d_accum, d_sched);
}

int nlps_topk_softmax(void* stream, const float* d_logits, int V, int k, float* d_work, int32_t* d_idx, float* d_prob) {
  NLPS_REQUIRE(d_logits && d_work && d_idx && d_prob, "nlps_topk_softmax: null argument");
  return topk_softmax((cudaStream_t)stream, d_logits, V, k, d_work, d_idx, d_prob);
}

int nlps_graph_stats(nlps_model* m, int64_t* eager, int64_t* captured, int64_t* replayed) {
  NLPS_REQUIRE(m, "null model");
  if (eager) *eager = m->n_eager;
  if (captured) *captured = m->n_captured;
  if (replayed) *replayed = m->n_replayed;
  return 0;
}

int nlps_launch_count(nlps_model*, int64_t* launches, int reset) {
  const long long n = launch_count(reset != 0);
  if (launches) *launches = n;
  return 0;
}

// ---- stand-alone operators ------------------------------------------------------------------
int nlps_gemm(void* stream, int precision, int M, int N, int K, const float* A, int64_t lda, int ta, const float* B,
              int64_t ldb, int tb, float* C, int64_t ldc, const float* bias, int relu, const float* gate, int64_t ldg,
              const float* residual, int64_t ldr, int accumulate) {
  NLPS_REQUIRE(A && B && C, "nlps_gemm: null operand");
  GemmArgs g = gemm_nn(M, N, K, A, lda, B, ldb, C, ldc);
  g.trans_a = ta != 0; g.trans_b = tb != 0;
  g.bias = bias; g.relu = relu != 0; g.gate = gate; g.ldg = ldg; g.residual = residual; g.ldr = ldr;
  g.accumulate = accumulate != 0;
  if (precision != NLPS_PREC_FP32)
    NLPS_REQUIRE(gemm_tf32_supported(g), "nlps_gemm: tensor-core path needs 16-B aligned operands with ld %% 4 == 0");
  return gemm_dispatch((cudaStream_t)stream, precision, g);
}

int nlps_attention_forward(void* stream, int precision, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                           float* ctx, float* lse) {
  int32_t* fnp = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  ensure_mempool_config();            // keep the scratch cached: a trip to the driver per call stalls the stream
  NLPS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&fnp), sizeof(int32_t) * B, st));
  NLPS_TRY(first_nonpad(st, B, S, x, fnp));
  const int rc = attention_forward(st, precision, B, S, h, d, qkv, x, fnp, ctx, lse);
  NLPS_CUDA(cudaFreeAsync(fnp, st));
  return rc;
}

int nlps_attention_backward(void* stream, int precision, int B, int S, int h, int d, const float* qkv,
                            const int32_t* x, const float* ctx, const float* lse, const float* dctx, float* dqkv) {
  cudaStream_t st = (cudaStream_t)stream;
  char* buf = nullptr;
  const size_t nd = sizeof(float) * (size_t)B * h * S;
  ensure_mempool_config();
  NLPS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&buf), nd + sizeof(int32_t) * B, st));
  float* delta = reinterpret_cast<float*>(buf);
  int32_t* fnp = reinterpret_cast<int32_t*>(buf + nd);
  NLPS_TRY(first_nonpad(st, B, S, x, fnp));
  const int rc = attention_backward(st, precision, B, S, h, d, qkv, x, fnp, ctx, lse, dctx, delta, dqkv);
  NLPS_CUDA(cudaFreeAsync(buf, st));
  return rc;
}

int nlps_layernorm_forward(void* stream, int64_t rows, int E, const float* x, const float* gamma, const float* beta,
                           float* y, float* mean, float* rstd) {
  return layernorm_forward((cudaStream_t)stream, rows, E, x, gamma, beta, y, mean, rstd);
}

int nlps_layernorm_backward(void* stream, int64_t rows, int E, const float* dy, const float* x, const float* mean,
                            const float* rstd, const float* gamma, float* dx, float* dgamma, float* dbeta) {
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = nullptr;
  ensure_mempool_config();
  NLPS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&partial), sizeof(float) * layernorm_backward_scratch_floats(E), st));
  const int rc = layernorm_backward(st, rows, E, dy, x, mean, rstd, gamma, dx, nullptr, nullptr, dgamma, dbeta, partial);
  NLPS_CUDA(cudaFreeAsync(partial, st));
  return rc;
}

int nlps_embed_forward(void* stream, int B, int S, int E, int V, const int32_t* x, int64_t xs, const float* We,
                       const float* pe, float* out, int32_t* x_flat, int32_t* err) {
  return embed_forward((cudaStream_t)stream, B, S, E, V, x, xs, We, pe, out, x_flat, err);
}

int nlps_embed_backward(void* stream, int64_t N, int E, int V, const int32_t* x_flat, const float* dh, float* dWe) {
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* scratch = nullptr;
  ensure_mempool_config();
  NLPS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), sizeof(int32_t) * embed_backward_scratch_ints(N, V), st));
  const int rc = embed_backward(st, N, E, V, x_flat, dh, dWe, scratch);
  NLPS_CUDA(cudaFreeAsync(scratch, st));
  return rc;
}

int nlps_colsum(void* stream, int64_t rows, int cols, const float* x, int64_t ld, float* out) {
  cudaStream_t st = (cudaStream_t)stream;
  float* scratch = nullptr;
  ensure_mempool_config();
  NLPS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), sizeof(float) * colsum_scratch_floats(cols), st));
  const int rc = colsum(st, rows, cols, x, ld, out, scratch);
  NLPS_CUDA(cudaFreeAsync(scratch, st));
  return rc;
}

int nlps_sumsq(void* stream, int64_t n, const float* x, float* out) {
  cudaStream_t st = (cudaStream_t)stream;
  float* scratch = nullptr;
  ensure_mempool_config();
  NLPS_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), sizeof(float) * sumsq_scratch_floats(), st));
  const int rc = sumsq(st, n, x, out, scratch);
  NLPS_CUDA(cudaFreeAsync(scratch, st));
  return rc;
}

int nlps_strided_copy(void* stream, int es, int nd, const int64_t* shape, const void* src, const int64_t* ss, void* dst,
                      const int64_t* ds) {
  return strided_copy((cudaStream_t)stream, es, nd, shape, src, ss, dst, ds);
}
int nlps_gather_rows(void* stream, int es, int64_t n_out, int64_t row_elems, const void* src, int64_t stride,
                     const int32_t* index, int64_t n_src, void* dst) {
  return gather_rows((cudaStream_t)stream, es, n_out, row_elems, src, stride, index, n_src, dst);
}
int nlps_gather_rows2(void* stream, int es, int64_t n_out, int64_t row_elems, const void* src, int64_t n0, int64_t n1,
                      const int32_t* i0, const int32_t* i1, void* dst) {
  return gather_rows2((cudaStream_t)stream, es, n_out, row_elems, src, n0, n1, i0, i1, dst);
}
int nlps_gather_last(void* stream, int B, int S, int E, const float* h, const int32_t* t, float* out) {
  return gather_last((cudaStream_t)stream, B, S, E, h, t, out);
}
int nlps_count_nonzero_rows(void* stream, int64_t rows, int64_t cols, const int32_t* x, int64_t stride, int32_t* out) {
  return count_nonzero_rows((cudaStream_t)stream, rows, cols, x, stride, out);
}
int nlps_binary_f32(void* stream, int op, int64_t n, int64_t cols, const float* a, const float* b, float bs, int mode,
                    float* out) {
  return binary_f32((cudaStream_t)stream, op, n, cols, a, b, bs, mode, out);
}
int nlps_binary_i32(void* stream, int op, int64_t n, const int32_t* a, const int32_t* b, int32_t bs, int mode,
                    int32_t* out) {
  return binary_i32((cudaStream_t)stream, op, n, a, b, bs, mode, out);
}
int nlps_unary_f32(void* stream, int op, int64_t n, const float* a, float* out) {
  return unary_f32((cudaStream_t)stream, op, n, a, out);
}
int nlps_reduce_f32(void* stream, int op, int64_t n, const float* a, float* out) {
  return reduce_f32((cudaStream_t)stream, op, n, a, out);
}
int nlps_reduce_i32(void* stream, int op, int64_t n, const int32_t* a, int32_t* out) {
  return reduce_i32((cudaStream_t)stream, op, n, a, out);
}
int nlps_compare_ne_i32(void* stream, int64_t n, const int32_t* a, int32_t value, int32_t* out) {
  return compare_ne_i32((cudaStream_t)stream, n, a, value, out);
}
int nlps_cast_i32_f32(void* stream, int64_t n, const int32_t* a, float* out) {
  return cast_i32_f32((cudaStream_t)stream, n, a, out);
}

int nlps_prefix_dataset(void* stream, int n_sent, int64_t n_rows, int max_len, int32_t pad, const int32_t* d_ids,
                        const int64_t* d_sent_off, const int64_t* d_row_off, int32_t* d_X, int32_t* d_Y) {
  NLPS_REQUIRE(n_rows == 0 || (d_ids && d_sent_off && d_row_off && d_X && d_Y), "nlps_prefix_dataset: null argument");
  return prefix_dataset((cudaStream_t)stream, n_sent, n_rows, max_len, pad, d_ids, d_sent_off, d_row_off, d_X, d_Y);
}
int nlps_expand_tokens_offsets(void* stream, int n_sent, const int32_t* d_words, const int64_t* d_sent_off,
                               const int32_t* d_tab_off, int32_t bos, int32_t eos, int32_t* d_len_scratch, int64_t* d_out_off) {
  NLPS_REQUIRE(d_words && d_sent_off && d_tab_off && d_len_scratch && d_out_off, "nlps_expand_tokens_offsets: null argument");
  return expand_tokens_offsets((cudaStream_t)stream, n_sent, d_words, d_sent_off, d_tab_off, bos, eos, d_len_scratch, d_out_off);
}
int nlps_expand_tokens_fill(void* stream, int n_sent, const int32_t* d_words, const int64_t* d_sent_off,
                            const int32_t* d_tab_off, const int32_t* d_tab_ids, int32_t bos, int32_t eos,
                            const int64_t* d_out_off, int32_t* d_out_ids) {
  NLPS_REQUIRE(d_words && d_sent_off && d_tab_off && d_tab_ids && d_out_off && d_out_ids, "nlps_expand_tokens_fill: null argument");
  return expand_tokens_fill((cudaStream_t)stream, n_sent, d_words, d_sent_off, d_tab_off, d_tab_ids, bos, eos, d_out_off, d_out_ids);
}

}  // extern "C"
