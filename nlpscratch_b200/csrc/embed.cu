// Token embedding gather (+ positional table) and its deterministic scatter-add gradient.
//
// Forward  H0[b,s,:] = We[x[b,s],:] + pe[s,:]                      (transformer.py:114)
// Backward dWe[v,:]  = sum over positions n with x[n]==v of dH[n,:] (transformer.py:453-457)
//
// np.add.at on the CPU adds occurrences in ascending position order; cupy.add.at uses atomics
// (order undefined).  Here the positions are bucketed by token id with a counting sort whose
// within-bucket order is the ascending position (rank = number of earlier equal ids), and one
// warp per vocabulary row adds them in that order: bit-reproducible, no float atomics.
#include "common.cuh"
#include <algorithm>

namespace nlps {
namespace {

__global__ void __launch_bounds__(256) embed_fwd_kernel(int B, int S, int E, int V, const int32_t* __restrict__ x,
                                                        int64_t xs, const float* __restrict__ We,
                                                        const float* __restrict__ pe, float* __restrict__ out,
                                                        int32_t* __restrict__ x_flat, int32_t* __restrict__ err,
                                                        int vec) {
  // one warp per token
  const int lane = threadIdx.x & 31;
  const int64_t tok = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= (int64_t)B * S) return;
  const int b = (int)(tok / S), s = (int)(tok - (int64_t)b * S);
  int id = x[(int64_t)b * xs + s];
  if (id < 0) id += V;                      // numpy-style negative index
  if (id < 0 || id >= V) {                  // numpy raises IndexError; flag it and clamp
    if (lane == 0 && err) atomicOr(err, 1);
    id = 0;
  }
  if (lane == 0 && x_flat) x_flat[tok] = id;
  const float* w = We + (int64_t)id * E;
  const float* p = pe + (int64_t)s * E;
  float* o = out + tok * E;
  if (vec) {
    for (int c = lane * 4; c < E; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(w + c);
      const float4 q = *reinterpret_cast<const float4*>(p + c);
      *reinterpret_cast<float4*>(o + c) = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
    }
  } else {
    for (int c = lane; c < E; c += 32) o[c] = w[c] + p[c];
  }
}

__global__ void hist_kernel(int64_t N, const int32_t* __restrict__ x, int32_t* __restrict__ count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&count[x[i]], 1);             // integer atomics: the result is order independent
}

// exclusive scan of count[0..V) into start[0..V], single CTA
__global__ void __launch_bounds__(1024) scan_kernel(int V, const int32_t* __restrict__ count, int32_t* __restrict__ start) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < V; base += 1024) {
    const int i = base + threadIdx.x;
    const int32_t v = i < V ? count[i] : 0;
    int32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;                   // inclusive totals of the warps
    }
    __syncthreads();
    const int32_t before = carry + (warp > 0 ? warp_tot[warp - 1] : 0);
    if (i < V) start[i] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) start[V] = carry;
}

// rank[n] = #{n' < n : x[n'] == x[n]};  order[start[x[n]] + rank[n]] = n
// (brute-force version, O(N^2) compares: kept for huge vocabulary x chunk products)
__global__ void __launch_bounds__(256) rank_kernel(int64_t N, const int32_t* __restrict__ x,
                                                   const int32_t* __restrict__ start, int32_t* __restrict__ order) {
  __shared__ int32_t tile[1024];
  const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int32_t mine = n < N ? x[n] : -1;
  int32_t rank = 0;
  const int64_t block_end = min(N, ((int64_t)blockIdx.x + 1) * 256);    // only positions < n matter
  for (int64_t base = 0; base < block_end; base += 1024) {
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 256) tile[i] = (base + i < N) ? x[base + i] : -2;
    __syncthreads();
    const int64_t rem = n - base;
    const int64_t lim = rem < 1024 ? rem : 1024;
    for (int i = 0; i < lim; ++i) rank += (tile[i] == mine);
  }
  if (n < N) order[start[mine] + rank] = (int32_t)n;
}

// Chunked version: positions are cut into chunks of 1024.  (1) rank inside the chunk by comparing
// against the chunk's ids in smem, and a per-(chunk, id) count; (2) per id, an exclusive prefix of
// the counts over chunks (and the id's total); (3) after the scan over ids, every position knows its
// slot: start[id] + earlier-chunks[id] + rank-in-chunk -- ascending position inside every bucket.
constexpr int kChunk = 1024;
__global__ void __launch_bounds__(kChunk) chunk_rank_kernel(int64_t N, int V, const int32_t* __restrict__ x,
                                                            int32_t* __restrict__ local_rank,
                                                            int32_t* __restrict__ chunk_hist) {
  __shared__ int32_t ids[kChunk];
  const int64_t n = (int64_t)blockIdx.x * kChunk + threadIdx.x;
  const int32_t mine = n < N ? x[n] : -1;
  ids[threadIdx.x] = mine;
  __syncthreads();
  int32_t rank = 0;
  for (int i = 0; i < (int)threadIdx.x; ++i) rank += (ids[i] == mine);     // smem broadcast reads
  if (n < N) {
    local_rank[n] = rank;
    atomicAdd(&chunk_hist[(int64_t)blockIdx.x * V + mine], 1);           // integer counts: order independent
  }
}
__global__ void chunk_prefix_kernel(int V, int chunks, int32_t* __restrict__ chunk_hist, int32_t* __restrict__ count) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int32_t run = 0;
  for (int c = 0; c < chunks; ++c) {
    const int32_t t = chunk_hist[(int64_t)c * V + v];
    chunk_hist[(int64_t)c * V + v] = run;                                 // exclusive prefix over chunks
    run += t;
  }
  count[v] = run;
}
__global__ void place_kernel(int64_t N, int V, const int32_t* __restrict__ x, const int32_t* __restrict__ local_rank,
                             const int32_t* __restrict__ chunk_off, const int32_t* __restrict__ start,
                             int32_t* __restrict__ order) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int32_t id = x[n];
  order[start[id] + chunk_off[(n / kChunk) * V + id] + local_rank[n]] = (int32_t)n;
}

__global__ void __launch_bounds__(256) scatter_sum_kernel(int V, int E, const int32_t* __restrict__ start,
                                                          const int32_t* __restrict__ order,
                                                          const float* __restrict__ dh, float* __restrict__ dWe, int vec) {
  // one warp per vocabulary row; occurrences added in ascending position
  const int lane = threadIdx.x & 31;
  const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (v >= V) return;
  const int s0 = start[v], s1 = start[v + 1];
  float* out = dWe + (int64_t)v * E;
  if (vec) {
    for (int c = lane * 4; c < E; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int i = s0;
      for (; i + 3 < s1; i += 4) {          // 4 independent loads in flight, added in order
        const float4 a = *reinterpret_cast<const float4*>(dh + (int64_t)order[i] * E + c);
        const float4 b = *reinterpret_cast<const float4*>(dh + (int64_t)order[i + 1] * E + c);
        const float4 d = *reinterpret_cast<const float4*>(dh + (int64_t)order[i + 2] * E + c);
        const float4 e = *reinterpret_cast<const float4*>(dh + (int64_t)order[i + 3] * E + c);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
        acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
      }
      for (; i < s1; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(dh + (int64_t)order[i] * E + c);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
      *reinterpret_cast<float4*>(out + c) = acc;
    }
  } else {
    for (int c = lane; c < E; c += 32) {
      float acc = 0.f;
      for (int i = s0; i < s1; ++i) acc += dh[(int64_t)order[i] * E + c];
      out[c] = acc;
    }
  }
}

// out = 0 except out[b, t[b], :] = rows[b, :]      (transformer.py:239-240)
__global__ void scatter_rows_kernel(int B, int S, int E, const int32_t* __restrict__ t, const float* __restrict__ rows,
                                    float* __restrict__ out) {
  const int64_t total = (int64_t)B * S * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = i / E;
    const int c = (int)(i - tok * E);
    const int b = (int)(tok / S), s = (int)(tok - (int64_t)b * S);
    int tb = t[b];
    if (tb < 0) tb += S;
    out[i] = (s == tb) ? rows[(int64_t)b * E + c] : 0.f;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int embed_forward(cudaStream_t st, int B, int S, int E, int V, const int32_t* x, int64_t xs, const float* We,
                  const float* pe, float* out, int32_t* x_flat, int32_t* err_flag) {
  const int64_t N = (int64_t)B * S;
  if (N <= 0) return 0;
  const int vec = (E % 4 == 0 && aligned16(We) && aligned16(pe) && aligned16(out)) ? 1 : 0;
  embed_fwd_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(B, S, E, V, x, xs, We, pe, out, x_flat, err_flag, vec);
  NLPS_LAUNCH_CHECK();
  return 0;
}

static bool use_chunked(int64_t N, int V) { return ceil_div(N, kChunk) * (int64_t)V <= ((int64_t)32 << 20); }

size_t embed_backward_scratch_ints(int64_t N, int V) {
  size_t n = (size_t)(2 * ((int64_t)V + 1) + N + 8);
  if (use_chunked(N, V)) n += (size_t)N + (size_t)(ceil_div(N, kChunk) * (int64_t)V);
  return n;
}

// Two halves: the counting sort of the ids depends on x only (it can run long before dH exists, on another stream); the
// scatter reads its result.  embed_backward = sort + apply.
int embed_backward_sort(cudaStream_t st, int64_t N, int V, const int32_t* x_flat, int32_t* scratch) {
  int32_t* count = scratch;
  int32_t* start = scratch + (V + 1);
  int32_t* order = scratch + 2 * ((int64_t)V + 1);
  if (N > 0 && use_chunked(N, V)) {
    const int chunks = (int)ceil_div(N, kChunk);
    int32_t* local_rank = order + N;
    int32_t* chunk_hist = local_rank + N;
    NLPS_CUDA(cudaMemsetAsync(chunk_hist, 0, sizeof(int32_t) * (size_t)chunks * V, st));
    chunk_rank_kernel<<<chunks, kChunk, 0, st>>>(N, V, x_flat, local_rank, chunk_hist);
    NLPS_LAUNCH_CHECK();
    chunk_prefix_kernel<<<(unsigned)ceil_div(V, 256), 256, 0, st>>>(V, chunks, chunk_hist, count);
    NLPS_LAUNCH_CHECK();
    scan_kernel<<<1, 1024, 0, st>>>(V, count, start);
    NLPS_LAUNCH_CHECK();
    place_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(N, V, x_flat, local_rank, chunk_hist, start, order);
    NLPS_LAUNCH_CHECK();
  } else {
    NLPS_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * (V + 1), st));
    if (N > 0) {
      hist_kernel<<<(unsigned)std::min<int64_t>(ceil_div(N, 256), kNumSMs * 8), 256, 0, st>>>(N, x_flat, count);
      NLPS_LAUNCH_CHECK();
    }
    scan_kernel<<<1, 1024, 0, st>>>(V, count, start);
    NLPS_LAUNCH_CHECK();
    if (N > 0) {
      rank_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(N, x_flat, start, order);
      NLPS_LAUNCH_CHECK();
    }
  }
  return 0;
}

int embed_backward_apply(cudaStream_t st, int64_t N, int E, int V, const float* dh, float* dWe, const int32_t* scratch) {
  (void)N;
  const int32_t* start = scratch + (V + 1);
  const int32_t* order = scratch + 2 * ((int64_t)V + 1);
  const int vec = (E % 4 == 0 && aligned16(dh) && aligned16(dWe)) ? 1 : 0;
  scatter_sum_kernel<<<(unsigned)ceil_div(V, 8), 256, 0, st>>>(V, E, start, order, dh, dWe, vec);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int embed_backward(cudaStream_t st, int64_t N, int E, int V, const int32_t* x_flat, const float* dh, float* dWe,
                   int32_t* scratch) {
  NLPS_TRY(embed_backward_sort(st, N, V, x_flat, scratch));
  return embed_backward_apply(st, N, E, V, dh, dWe, scratch);
}

int scatter_rows(cudaStream_t st, int B, int S, int E, const int32_t* t, const float* rows, float* out) {
  const int64_t total = (int64_t)B * S * E;
  if (total <= 0) return 0;
  scatter_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), kNumSMs * 16), 256, 0, st>>>(B, S, E, t, rows, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace nlps
