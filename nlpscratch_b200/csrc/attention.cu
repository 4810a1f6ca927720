// Fused causal + key-padding attention, forward and backward, FP32 CUDA-core version.
//
// Replaces transformer.py:131-141 (forward) and :427-441 (backward).  The (B,h,S,S) score and
// probability tensors of the reference are never materialised: the forward keeps a running
// (max, sum) per query row (online softmax, reductions by warp shuffles over the 16 lanes that
// share a row) and stores only the merged-head context and the log-sum-exp; the backward
// recomputes P = exp(S - lse) tile by tile.  Two backward kernels (one owning key tiles for
// dK/dV, one owning query tiles for dQ) keep every output element owned by exactly one thread:
// no atomics, deterministic.
//
// Mask semantics (transformer.py:585-606, additive -1e9, not -inf): key j is visible to query i
// iff j <= i and x[b,j] != PAD -- masked probabilities underflow to exactly 0 in the reference,
// so skipping them is exact.  A query row with NO visible key (i < first non-PAD position) sees
// the same -1e9 on all S keys, i.e. a softmax over the raw scores of ALL keys, future included.
//
// Q|K|V arrive packed as rows of 3E floats; head j owns columns [j*d,(j+1)*d) of each block.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace nlps {
namespace {

constexpr int BQ = 64, BKV = 64, PLD = BKV + 4;
constexpr float kNegBig = -1e30f;

__global__ void first_nonpad_kernel(int B, int S, const int32_t* __restrict__ x, int32_t* __restrict__ out) {
  const int b = blockIdx.x;
  int best = S;
  for (int j = threadIdx.x; j < S; j += blockDim.x)
    if (x[(int64_t)b * S + j] != 0) { best = j; break; }
  // min over the block (threads scan interleaved positions, each stops at its first hit)
  __shared__ int smin;
  if (threadIdx.x == 0) smin = S;
  __syncthreads();
  atomicMin(&smin, best);
  __syncthreads();
  if (threadIdx.x == 0) out[b] = smin;
}

// load a (rows x d) slab of the packed qkv matrix into smem [rows][D+4], zero padded
template <int D>
__device__ __forceinline__ void load_slab(float* dst, const float* __restrict__ src, int64_t row_stride,
                                          int first_row, int S, int d, int tid) {
  constexpr int LD = D + 4;
  for (int idx = tid; idx < 64 * D; idx += 256) {
    const int r = idx / D, c = idx - r * D;
    const int gr = first_row + r;
    dst[r * LD + c] = (gr < S && c < d) ? src[(int64_t)gr * row_stride + c] : 0.f;
  }
}

template <int D>
__device__ __forceinline__ void tile_dot(const float* __restrict__ As, const float* __restrict__ Bs, int ty, int tx,
                                         float (&acc)[4][4]) {
  constexpr int LD = D + 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < D; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * LD + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + (tx + 16 * j) * LD + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
}

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct AttnDims {
  int B, S, h, d, E;
  float scale;     // 1/sqrt(d)
};

// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnDims p, const float* __restrict__ qkv,
                                                       const int32_t* __restrict__ x,
                                                       const int32_t* __restrict__ fnp, float* __restrict__ ctx,
                                                       float* __restrict__ lse) {
  constexpr int LD = D + 4, NC = (D + 15) / 16;
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* Ks = Qs + BQ * LD;
  float* Vs = Ks + BKV * LD;
  float* Ps = Vs + BKV * LD;
  int* kpad = reinterpret_cast<int*>(Ps + BQ * PLD);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * BQ;
  const int64_t rs = 3 * (int64_t)p.E;
  const float* qbase = qkv + (int64_t)b * p.S * rs + head * p.d;
  const float* kbase = qbase + p.E;
  const float* vbase = qbase + 2 * p.E;
  const int first_real = fnp[b];
  const bool deg_block = q0 < first_real;

  load_slab<D>(Qs, qbase, rs, q0, p.S, p.d, tid);

  float m[4], l[4], o[4][NC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = kNegBig; l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) o[i][c] = 0.f;
  }
  const int n_kv_all = (p.S + BKV - 1) / BKV;
  const int n_kv = deg_block ? n_kv_all : min(n_kv_all, (q0 + BQ - 1) / BKV + 1);

  for (int kb = 0; kb < n_kv; ++kb) {
    const int k0 = kb * BKV;
    __syncthreads();
    load_slab<D>(Ks, kbase, rs, k0, p.S, p.d, tid);
    load_slab<D>(Vs, vbase, rs, k0, p.S, p.d, tid);
    if (tid < BKV) kpad[tid] = (k0 + tid < p.S) ? (x[(int64_t)b * p.S + k0 + tid] == 0) : 1;
    __syncthreads();

    float s[4][4];
    tile_dot<D>(Qs, Ks, ty, tx, s);
    float alpha[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = q0 + ty + 16 * i;
      const bool deg = r < first_real;
      float mx = kNegBig;
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = k0 + tx + 16 * j;
        ok[j] = (c < p.S) && (r < p.S) && (deg || (c <= r && !kpad[tx + 16 * j]));
        s[i][j] *= p.scale;
        if (ok[j]) mx = fmaxf(mx, s[i][j]);
      }
      mx = half_warp_max(mx);
      const float m_new = fmaxf(m[i], mx);
      alpha[i] = expf(m[i] - m_new);
      float rsum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pv = ok[j] ? expf(s[i][j] - m_new) : 0.f;
        Ps[(ty + 16 * i) * PLD + tx + 16 * j] = pv;
        rsum += pv;
      }
      rsum = half_warp_sum(rsum);
      l[i] = l[i] * alpha[i] + rsum;
      m[i] = m_new;
#pragma unroll
      for (int c = 0; c < NC; ++c) o[i][c] *= alpha[i];
    }
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < BKV; j += 4) {
      float4 pa[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pa[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * PLD + j);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = tx + 16 * c;
        if (D >= 16 || col < D) {
          const float v0 = Vs[(j + 0) * LD + col], v1 = Vs[(j + 1) * LD + col];
          const float v2 = Vs[(j + 2) * LD + col], v3 = Vs[(j + 3) * LD + col];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o[i][c] = fmaf(pa[i].x, v0, o[i][c]);
            o[i][c] = fmaf(pa[i].y, v1, o[i][c]);
            o[i][c] = fmaf(pa[i].z, v2, o[i][c]);
            o[i][c] = fmaf(pa[i].w, v3, o[i][c]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + ty + 16 * i;
    if (r >= p.S) continue;
    const float inv = 1.f / l[i];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = tx + 16 * c;
      if (col < p.d) ctx[((int64_t)b * p.S + r) * p.E + head * p.d + col] = o[i][c] * inv;
    }
    if (tx == 0) lse[((int64_t)b * p.h + head) * p.S + r] = m[i] + logf(l[i]);
  }
}

// delta[b,h,i] = sum_c dctx[b,i,h,c] * ctx[b,i,h,c]   (the row term of softmax_backward,
// function.py:59-65, moved through P@V: sum_j dP_ij P_ij = dO_i . O_i)
__global__ void attn_delta_kernel(AttnDims p, const float* __restrict__ ctx, const float* __restrict__ dctx,
                                  float* __restrict__ delta) {
  const int64_t tok = blockIdx.x;             // b*S + i
  const int b = (int)(tok / p.S), i = (int)(tok - (int64_t)b * p.S);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int head = warp; head < p.h; head += nwarp) {
    const int64_t off = tok * p.E + head * p.d;
    float acc = 0.f;
    for (int c = lane; c < p.d; c += 32) acc = fmaf(dctx[off + c], ctx[off + c], acc);
    acc = warp_sum(acc);
    if (lane == 0) delta[((int64_t)b * p.h + head) * p.S + i] = acc;
  }
}

// vectorised variant: one warp per token, float4 per lane; a head is G = d/4 adjacent lanes
// (G a power of two <= 32), reduced with G-wide shuffles.  8 tokens per CTA.
template <int G>
__global__ void __launch_bounds__(256) attn_delta_vec_kernel(AttnDims p, int64_t ntok, const float* __restrict__ ctx,
                                                             const float* __restrict__ dctx,
                                                             float* __restrict__ delta) {
  const int lane = threadIdx.x & 31;
  const int64_t tok = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= ntok) return;
  const int b = (int)(tok / p.S), i = (int)(tok - (int64_t)b * p.S);
  const float4* c4 = reinterpret_cast<const float4*>(ctx + tok * p.E);
  const float4* d4 = reinterpret_cast<const float4*>(dctx + tok * p.E);
  const int n4 = p.E / 4;
  for (int base = 0; base < n4; base += 32) {
    const int idx = base + lane;
    float acc = 0.f;
    if (idx < n4) {
      const float4 a = c4[idx], g = d4[idx];
      acc = (a.x * g.x + a.y * g.y) + (a.z * g.z + a.w * g.w);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (idx < n4 && (lane & (G - 1)) == 0) delta[((int64_t)b * p.h + idx / G) * p.S + i] = acc;
  }
}

// recompute P and dS for one (query tile, key tile) pair; results left in registers
template <int D>
__device__ __forceinline__ void recompute_p_ds(const AttnDims& p, const float* Qs, const float* Ks, const float* dOs,
                                               const float* Vs, const float* lse_s, const float* delta_s,
                                               const int* kpad, int q0, int k0, int first_real, int ty, int tx,
                                               float (&pr)[4][4], float (&ds)[4][4]) {
  float s[4][4], dp[4][4];
  tile_dot<D>(Qs, Ks, ty, tx, s);
  tile_dot<D>(dOs, Vs, ty, tx, dp);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty + 16 * i, r = q0 + rl;
    const bool deg = r < first_real;
    const float ls = lse_s[rl], dl = delta_s[rl];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = tx + 16 * j, c = k0 + cl;
      const bool ok = (c < p.S) && (r < p.S) && (deg || (c <= r && !kpad[cl]));
      const float pv = ok ? expf(s[i][j] * p.scale - ls) : 0.f;
      pr[i][j] = pv;
      ds[i][j] = pv * (dp[i][j] - dl);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256) attn_bwd_dkdv_kernel(AttnDims p, const float* __restrict__ qkv,
                                                            const int32_t* __restrict__ x,
                                                            const int32_t* __restrict__ fnp,
                                                            const float* __restrict__ dctx,
                                                            const float* __restrict__ lse,
                                                            const float* __restrict__ delta,
                                                            float* __restrict__ dqkv) {
  constexpr int LD = D + 4, NC = (D + 15) / 16;
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;
  float* Vs = Ks + BKV * LD;
  float* Qs = Vs + BKV * LD;
  float* dOs = Qs + BQ * LD;
  float* Pt = dOs + BQ * LD;        // [key][query]  (transposed so the second GEMM reads float4 along queries)
  float* dSt = Pt + BKV * PLD;
  float* lse_s = dSt + BKV * PLD;
  float* delta_s = lse_s + BQ;
  int* kpad = reinterpret_cast<int*>(delta_s + BQ);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, head = blockIdx.y, k0 = blockIdx.x * BKV;
  const int64_t rs = 3 * (int64_t)p.E;
  const float* qbase = qkv + (int64_t)b * p.S * rs + head * p.d;
  const float* kbase = qbase + p.E;
  const float* vbase = qbase + 2 * p.E;
  const float* dobase = dctx + (int64_t)b * p.S * p.E + head * p.d;
  const int first_real = fnp[b];

  load_slab<D>(Ks, kbase, rs, k0, p.S, p.d, tid);
  load_slab<D>(Vs, vbase, rs, k0, p.S, p.d, tid);
  if (tid < BKV) kpad[tid] = (k0 + tid < p.S) ? (x[(int64_t)b * p.S + k0 + tid] == 0) : 1;

  float dk[4][NC], dv[4][NC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < NC; ++c) { dk[i][c] = 0.f; dv[i][c] = 0.f; }

  const int n_q = (p.S + BQ - 1) / BQ;
  for (int qb = 0; qb < n_q; ++qb) {
    const int q0 = qb * BQ;
    const bool causal_hit = (q0 + BQ - 1) >= k0;       // some query row at or after the first key
    const bool deg_hit = q0 < first_real;              // fully-masked rows see every key
    if (!causal_hit && !deg_hit) continue;             // block-uniform
    __syncthreads();
    load_slab<D>(Qs, qbase, rs, q0, p.S, p.d, tid);
    load_slab<D>(dOs, dobase, p.E, q0, p.S, p.d, tid);
    if (tid < BQ) {
      const int r = q0 + tid;
      lse_s[tid] = r < p.S ? lse[((int64_t)b * p.h + head) * p.S + r] : 0.f;
      delta_s[tid] = r < p.S ? delta[((int64_t)b * p.h + head) * p.S + r] : 0.f;
    }
    __syncthreads();
    float pr[4][4], ds[4][4];
    recompute_p_ds<D>(p, Qs, Ks, dOs, Vs, lse_s, delta_s, kpad, q0, k0, first_real, ty, tx, pr, ds);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Pt[(tx + 16 * j) * PLD + ty + 16 * i] = pr[i][j];
        dSt[(tx + 16 * j) * PLD + ty + 16 * i] = ds[i][j];
      }
    __syncthreads();
    // dV[key,c] += sum_r P[r,key] dO[r,c] ;  dK[key,c] += sum_r dS[r,key] Q[r,c]
#pragma unroll 2
    for (int r = 0; r < BQ; r += 4) {
      float4 pa[4], da[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pa[i] = *reinterpret_cast<const float4*>(Pt + (ty + 16 * i) * PLD + r);
        da[i] = *reinterpret_cast<const float4*>(dSt + (ty + 16 * i) * PLD + r);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = tx + 16 * c;
        if (D >= 16 || col < D) {
          const float g0 = dOs[(r + 0) * LD + col], g1 = dOs[(r + 1) * LD + col];
          const float g2 = dOs[(r + 2) * LD + col], g3 = dOs[(r + 3) * LD + col];
          const float q0v = Qs[(r + 0) * LD + col], q1v = Qs[(r + 1) * LD + col];
          const float q2v = Qs[(r + 2) * LD + col], q3v = Qs[(r + 3) * LD + col];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dv[i][c] = fmaf(pa[i].x, g0, dv[i][c]);
            dv[i][c] = fmaf(pa[i].y, g1, dv[i][c]);
            dv[i][c] = fmaf(pa[i].z, g2, dv[i][c]);
            dv[i][c] = fmaf(pa[i].w, g3, dv[i][c]);
            dk[i][c] = fmaf(da[i].x, q0v, dk[i][c]);
            dk[i][c] = fmaf(da[i].y, q1v, dk[i][c]);
            dk[i][c] = fmaf(da[i].z, q2v, dk[i][c]);
            dk[i][c] = fmaf(da[i].w, q3v, dk[i][c]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = k0 + ty + 16 * i;
    if (key >= p.S) continue;
    float* out = dqkv + ((int64_t)b * p.S + key) * rs + head * p.d;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = tx + 16 * c;
      if (col < p.d) {
        out[p.E + col] = dk[i][c] * p.scale;
        out[2 * p.E + col] = dv[i][c];
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(AttnDims p, const float* __restrict__ qkv,
                                                          const int32_t* __restrict__ x,
                                                          const int32_t* __restrict__ fnp,
                                                          const float* __restrict__ dctx,
                                                          const float* __restrict__ lse,
                                                          const float* __restrict__ delta,
                                                          float* __restrict__ dqkv) {
  constexpr int LD = D + 4, NC = (D + 15) / 16;
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* dOs = Qs + BQ * LD;
  float* Ks = dOs + BQ * LD;
  float* Vs = Ks + BKV * LD;
  float* dSs = Vs + BKV * LD;       // [query][key]
  float* lse_s = dSs + BQ * PLD;
  float* delta_s = lse_s + BQ;
  int* kpad = reinterpret_cast<int*>(delta_s + BQ);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * BQ;
  const int64_t rs = 3 * (int64_t)p.E;
  const float* qbase = qkv + (int64_t)b * p.S * rs + head * p.d;
  const float* kbase = qbase + p.E;
  const float* vbase = qbase + 2 * p.E;
  const float* dobase = dctx + (int64_t)b * p.S * p.E + head * p.d;
  const int first_real = fnp[b];
  const bool deg_block = q0 < first_real;

  load_slab<D>(Qs, qbase, rs, q0, p.S, p.d, tid);
  load_slab<D>(dOs, dobase, p.E, q0, p.S, p.d, tid);
  if (tid < BQ) {
    const int r = q0 + tid;
    lse_s[tid] = r < p.S ? lse[((int64_t)b * p.h + head) * p.S + r] : 0.f;
    delta_s[tid] = r < p.S ? delta[((int64_t)b * p.h + head) * p.S + r] : 0.f;
  }
  float dq[4][NC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < NC; ++c) dq[i][c] = 0.f;

  const int n_kv_all = (p.S + BKV - 1) / BKV;
  const int n_kv = deg_block ? n_kv_all : min(n_kv_all, (q0 + BQ - 1) / BKV + 1);
  for (int kb = 0; kb < n_kv; ++kb) {
    const int k0 = kb * BKV;
    __syncthreads();
    load_slab<D>(Ks, kbase, rs, k0, p.S, p.d, tid);
    load_slab<D>(Vs, vbase, rs, k0, p.S, p.d, tid);
    if (tid < BKV) kpad[tid] = (k0 + tid < p.S) ? (x[(int64_t)b * p.S + k0 + tid] == 0) : 1;
    __syncthreads();
    float pr[4][4], ds[4][4];
    recompute_p_ds<D>(p, Qs, Ks, dOs, Vs, lse_s, delta_s, kpad, q0, k0, first_real, ty, tx, pr, ds);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dSs[(ty + 16 * i) * PLD + tx + 16 * j] = ds[i][j];
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < BKV; j += 4) {
      float4 da[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) da[i] = *reinterpret_cast<const float4*>(dSs + (ty + 16 * i) * PLD + j);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = tx + 16 * c;
        if (D >= 16 || col < D) {
          const float k0v = Ks[(j + 0) * LD + col], k1v = Ks[(j + 1) * LD + col];
          const float k2v = Ks[(j + 2) * LD + col], k3v = Ks[(j + 3) * LD + col];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dq[i][c] = fmaf(da[i].x, k0v, dq[i][c]);
            dq[i][c] = fmaf(da[i].y, k1v, dq[i][c]);
            dq[i][c] = fmaf(da[i].z, k2v, dq[i][c]);
            dq[i][c] = fmaf(da[i].w, k3v, dq[i][c]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + ty + 16 * i;
    if (r >= p.S) continue;
    float* out = dqkv + ((int64_t)b * p.S + r) * rs + head * p.d;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = tx + 16 * c;
      if (col < p.d) out[col] = dq[i][c] * p.scale;
    }
  }
}

template <int D> constexpr size_t fwd_smem() { return (size_t)(3 * 64 * (D + 4) + BQ * PLD) * 4 + BKV * 4; }
template <int D> constexpr size_t dkdv_smem() { return (size_t)(4 * 64 * (D + 4) + 2 * BKV * PLD + 2 * BQ) * 4 + BKV * 4; }
template <int D> constexpr size_t dq_smem() { return (size_t)(4 * 64 * (D + 4) + BQ * PLD + 2 * BQ) * 4 + BKV * 4; }

template <typename K>
int set_smem(K kern, size_t bytes) {
  NLPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int D>
int fwd_launch(cudaStream_t st, const AttnDims& p, const float* qkv, const int32_t* x, const int32_t* fnp, float* ctx,
               float* lse) {
  static bool once = false;
  if (!once) { NLPS_TRY(set_smem(attn_fwd_kernel<D>, fwd_smem<D>())); once = true; }
  dim3 grid((unsigned)ceil_div(p.S, BQ), (unsigned)p.h, (unsigned)p.B);
  attn_fwd_kernel<D><<<grid, 256, fwd_smem<D>(), st>>>(p, qkv, x, fnp, ctx, lse);
  NLPS_LAUNCH_CHECK();
  return 0;
}

template <int D>
int bwd_launch(cudaStream_t st, const AttnDims& p, const float* qkv, const int32_t* x, const int32_t* fnp,
               const float* dctx, const float* lse, const float* delta, float* dqkv) {
  static bool once = false;
  if (!once) {
    NLPS_TRY(set_smem(attn_bwd_dkdv_kernel<D>, dkdv_smem<D>()));
    NLPS_TRY(set_smem(attn_bwd_dq_kernel<D>, dq_smem<D>()));
    once = true;
  }
  dim3 grid((unsigned)ceil_div(p.S, BQ), (unsigned)p.h, (unsigned)p.B);
  attn_bwd_dkdv_kernel<D><<<grid, 256, dkdv_smem<D>(), st>>>(p, qkv, x, fnp, dctx, lse, delta, dqkv);
  NLPS_LAUNCH_CHECK();
  attn_bwd_dq_kernel<D><<<grid, 256, dq_smem<D>(), st>>>(p, qkv, x, fnp, dctx, lse, delta, dqkv);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int check_dims(int B, int S, int h, int d) {
  NLPS_REQUIRE(B > 0 && S > 0 && h > 0 && d > 0, "attention: empty problem B=%d S=%d h=%d d=%d", B, S, h, d);
  NLPS_REQUIRE(d <= 128, "attention: head_dim %d > 128 is not supported", d);
  NLPS_REQUIRE(B <= 65535 && h <= 65535, "attention: B=%d or h=%d exceeds the grid limits", B, h);
  return 0;
}

}  // namespace

// NLPS_ATTN=mma forces the mma.sync kernels where the tcgen05 ones would run (A/B measurements)
static int attn_impl() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NLPS_ATTN"); v = (e && e[0] == 'm') ? 1 : 0; }
  return v;
}

int first_nonpad(cudaStream_t st, int B, int S, const int32_t* x, int32_t* out) {
  first_nonpad_kernel<<<B, 128, 0, st>>>(B, S, x, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

// the tcgen05 kernels serve head_dim 32 / 64 on TMA-addressable packed rows
bool attention_tc_supported(int d, int E, const float* qkv) {
  return (d == 32 || d == 64) && (E % 4 == 0) && ((reinterpret_cast<uintptr_t>(qkv) & 15u) == 0);
}

static bool x3_attention_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NLPS_X3_ATTN"); v = (e && e[0] == 'f') ? 0 : 1; }
  return v == 1;
}

int attention_forward(cudaStream_t st, int precision, int B, int S, int h, int d, const float* qkv,
                      const int32_t* x, const int32_t* fnp, float* ctx, float* lse) {
  NLPS_TRY(check_dims(B, S, h, d));
  if (precision == 1 && attention_tc_supported(d, h * d, qkv) && attn_impl() != 1) {
    count_kind(KIND_ATTN_TS_FWD);
    return attention_forward_ts(st, B, S, h, d, qkv, x, fnp, ctx, lse);
  }
  if (precision == 1 && d <= 64 && d % 4 == 0) { count_kind(KIND_ATTN_MMA); return attention_forward_mma(st, B, S, h, d, qkv, x, fnp, ctx, lse); }
  // FP32-faithful tensor-core mode: 3xTF32 split products on mma.sync (NLPS_X3_ATTN=fp32 keeps the CUDA-core kernels, A/B)
  if (precision == 2 && d <= 64 && d % 4 == 0 && x3_attention_on()) { count_kind(KIND_ATTN_MMA); return attention_forward_mma(st, B, S, h, d, qkv, x, fnp, ctx, lse, true); }
  count_kind(KIND_ATTN_FP32);
  AttnDims p{B, S, h, d, h * d, (float)(1.0 / sqrt((double)d))};
  if (d <= 8) return fwd_launch<8>(st, p, qkv, x, fnp, ctx, lse);
  if (d <= 16) return fwd_launch<16>(st, p, qkv, x, fnp, ctx, lse);
  if (d <= 32) return fwd_launch<32>(st, p, qkv, x, fnp, ctx, lse);
  if (d <= 64) return fwd_launch<64>(st, p, qkv, x, fnp, ctx, lse);
  return fwd_launch<128>(st, p, qkv, x, fnp, ctx, lse);
}

int attention_backward(cudaStream_t st, int precision, int B, int S, int h, int d, const float* qkv,
                       const int32_t* x, const int32_t* fnp, const float* ctx, const float* lse,
                       const float* dctx, float* delta, float* dqkv, bool delta_ready) {
  NLPS_TRY(check_dims(B, S, h, d));
  AttnDims p{B, S, h, d, h * d, (float)(1.0 / sqrt((double)d))};
  const int64_t ntok = (int64_t)B * S;
  const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(ctx) | reinterpret_cast<uintptr_t>(dctx)) & 15u) == 0;
  const unsigned vgrid = (unsigned)ceil_div(ntok, 8);
  if (delta_ready) goto delta_done;
  if (vec && d == 64) attn_delta_vec_kernel<16><<<vgrid, 256, 0, st>>>(p, ntok, ctx, dctx, delta);
  else if (vec && d == 32) attn_delta_vec_kernel<8><<<vgrid, 256, 0, st>>>(p, ntok, ctx, dctx, delta);
  else if (vec && d == 128) attn_delta_vec_kernel<32><<<vgrid, 256, 0, st>>>(p, ntok, ctx, dctx, delta);
  else if (vec && d == 16) attn_delta_vec_kernel<4><<<vgrid, 256, 0, st>>>(p, ntok, ctx, dctx, delta);
  else if (vec && d == 8) attn_delta_vec_kernel<2><<<vgrid, 256, 0, st>>>(p, ntok, ctx, dctx, delta);
  else attn_delta_kernel<<<(unsigned)ntok, std::min(1024, 32 * h), 0, st>>>(p, ctx, dctx, delta);
  NLPS_LAUNCH_CHECK();
delta_done:
  if (precision == 1 && attention_tc_supported(d, h * d, qkv) && attn_impl() != 1 &&
      (reinterpret_cast<uintptr_t>(dctx) & 15u) == 0) {
    count_kind(KIND_ATTN_TS_BWD);
    return attention_backward_ts(st, B, S, h, d, qkv, x, fnp, dctx, lse, delta, dqkv);
  }
  if (precision == 1 && d <= 64 && d % 4 == 0) { count_kind(KIND_ATTN_MMA); return attention_backward_mma(st, B, S, h, d, qkv, x, fnp, dctx, lse, delta, dqkv); }
  if (precision == 2 && d <= 64 && d % 4 == 0 && x3_attention_on()) { count_kind(KIND_ATTN_MMA); return attention_backward_mma(st, B, S, h, d, qkv, x, fnp, dctx, lse, delta, dqkv, true); }
  count_kind(KIND_ATTN_FP32);
  if (d <= 8) return bwd_launch<8>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
  if (d <= 16) return bwd_launch<16>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
  if (d <= 32) return bwd_launch<32>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
  if (d <= 64) return bwd_launch<64>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
  return bwd_launch<128>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
}

}  // namespace nlps
