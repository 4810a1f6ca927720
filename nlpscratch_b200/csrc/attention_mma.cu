// Fused causal + key-padding attention on the tensor cores (TF32 mma.sync m16n8k8, FP32
// accumulate), forward and backward.  Used by the tf32 / tf32x3 precision modes; the FP32 mode
// keeps the CUDA-core kernels of attention.cu.  Same mask semantics, same outputs (context +
// log-sum-exp; dQ|dK|dV packed), same determinism (every output element has one owner thread).
//
// Shape of the computation (flash-attention style, nothing of size S x S ever leaves the SM):
//   * a warp owns 16 rows of a 64-row tile; scores live in mma accumulator fragments, the online
//     softmax reduces a row over the 4 lanes that share it with two warp shuffles;
//   * the probabilities never go to shared memory: the C-fragment of an 8-column score tile
//     (row g, cols 2t,2t+1) IS the A-fragment of the following product if that product's k index
//     is read as k=t -> col 2t, k=t+4 -> col 2t+1, so the B fragments are simply loaded in that order;
//   * backward: the dK/dV kernel computes the TRANSPOSED tiles S^T = K Q^T and dP^T = V dO^T so that
//     P^T and dS^T come out in the fragment layout the dV / dK products need; the dQ kernel uses the
//     untransposed orientation.  Operands are rounded to TF32 (cvt.rna) once, when staged in smem.
//   * smem rows are padded to D+4 floats: every fragment load pattern below is bank-conflict free.
#include "common.cuh"
#include <algorithm>
#include <cmath>

namespace nlps {
namespace {

constexpr int TQ = 64;          // tile rows (queries or keys) per CTA, 16 per warp
constexpr int kThreads = 128;
constexpr float kNegBig = -1e30f;

struct Dims {
  int B, S, h, d, E;
  float scale;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 3xTF32 on mma.sync (the FP32-faithful `tf32x3` mode): operands stay FP32 in shared memory / registers and every product is
// formed as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (small terms first; a_lo*b_lo ~ 2^-22 is dropped), hi = cvt.rna.tf32(x), lo = x - hi.
__device__ __forceinline__ void split_tf32(uint32_t raw, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(__uint_as_float(raw));
  lo = __float_as_uint(__uint_as_float(raw) - __uint_as_float(hi));     // exact; the tensor core truncates it to TF32
}
template <bool X3>
__device__ __forceinline__ void mma_op(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (!X3) { mma_tf32(c, a, b0, b1); return; }
  uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
  split_tf32(b0, bh0, bl0);
  split_tf32(b1, bh1, bl1);
  mma_tf32(c, al, bh0, bh1);
  mma_tf32(c, ah, bl0, bl1);
  mma_tf32(c, ah, bh0, bh1);
}
// a probability / dS value on its way into an A fragment: rounded to TF32 in the fast mode, raw FP32 bits under 3xTF32
template <bool X3>
__device__ __forceinline__ uint32_t frag(float x) { return X3 ? __float_as_uint(x) : to_tf32(x); }

// Under 3xTF32 the long accumulations (over key tiles in the forward / dQ, over query tiles in dK/dV) do not run through the
// tensor core's accumulator -- it adds with truncation, and the bias of a chain grows with its length (S/8 steps) -- but
// through a per-tile temporary that is added to the running sum with round-to-nearest FP32 adds.
template <bool X3, int NT, int N>
__device__ __forceinline__ float (&pick_acc(float (&tmp)[NT][4], float (&run)[N][4]))[N][4] {
  if constexpr (X3) { static_assert(NT == N, "temporary must match the accumulator"); return tmp; } else return run;
}
template <bool X3, int N>
__device__ __forceinline__ void tile_begin(float (&tmp)[N][4]) {
  if constexpr (X3) {
#pragma unroll
    for (int n = 0; n < N; ++n) tmp[n][0] = tmp[n][1] = tmp[n][2] = tmp[n][3] = 0.f;
  }
}
template <bool X3, int N, int NT>
__device__ __forceinline__ void tile_end_acc(float (&run)[N][4], const float (&tmp)[NT][4]) {
  if constexpr (X3) {
#pragma unroll
    for (int n = 0; n < N; ++n) { run[n][0] += tmp[n][0]; run[n][1] += tmp[n][1]; run[n][2] += tmp[n][2]; run[n][3] += tmp[n][3]; }
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Tiles are staged with 16-byte cp.async (zero-filled outside the tensor / beyond head_dim) into a
// double buffer: the copy of tile i+1 is in flight while the tensor cores work on tile i.  After the
// copy lands, each thread rounds the chunks IT copied to TF32 in place (cvt.rna), then the CTA syncs.
__device__ __forceinline__ void cp_async16(uint32_t* smem_dst, const float* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int D>
__device__ __forceinline__ void issue_tile(uint32_t* dst, const float* __restrict__ src, int64_t row_stride,
                                           int first_row, int S, int d, int tid) {
  constexpr int LD = D + 4, C4 = D / 4;
  for (int idx = tid; idx < TQ * C4; idx += kThreads) {
    const int r = idx / C4, c = (idx - r * C4) * 4;
    const int gr = first_row + r;
    const bool valid = gr < S && c < d;
    cp_async16(dst + r * LD + c, valid ? src + (int64_t)gr * row_stride + c : src, valid);
  }
}
template <int D, bool X3>
__device__ __forceinline__ void round_tile(uint32_t* tile, int tid) {
  if (X3) return;                       // 3xTF32 keeps the FP32 values; mma_op splits them
  constexpr int LD = D + 4, C4 = D / 4;
  for (int idx = tid; idx < TQ * C4; idx += kThreads) {
    const int r = idx / C4, c = (idx - r * C4) * 4;
    uint4* p = reinterpret_cast<uint4*>(tile + r * LD + c);
    uint4 v = *p;
    v.x = to_tf32(__uint_as_float(v.x)); v.y = to_tf32(__uint_as_float(v.y));
    v.z = to_tf32(__uint_as_float(v.z)); v.w = to_tf32(__uint_as_float(v.w));
    *p = v;
  }
}
// synchronous variant for the one-off operand tiles whose fragments live in registers
template <int D, bool X3>
__device__ __forceinline__ void stage_tile(uint32_t* dst, const float* __restrict__ src, int64_t row_stride,
                                           int first_row, int S, int d, int tid) {
  issue_tile<D>(dst, src, row_stride, first_row, S, d, tid);
  cp_async_commit();
  cp_async_wait_all();
  round_tile<D, X3>(dst, tid);
}

// A fragments of the warp's 16 rows for all D/8 k-chunks (standard k order): rows g, g+8; cols t, t+4
template <int D>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[D / 8][4], const uint32_t* tile, int warp, int g, int t) {
  constexpr int LD = D + 4;
  const uint32_t* r0 = tile + (warp * 16 + g) * LD;
  const uint32_t* r1 = r0 + 8 * LD;
#pragma unroll
  for (int kc = 0; kc < D / 8; ++kc) {
    a[kc][0] = r0[kc * 8 + t];
    a[kc][1] = r1[kc * 8 + t];
    a[kc][2] = r0[kc * 8 + t + 4];
    a[kc][3] = r1[kc * 8 + t + 4];
  }
}

// c[4] = A(16 x D) . rows(8j..8j+7 of tile)^T : one 16x8 score tile, B read as [n=row][k=dim]
template <int D, bool X3>
__device__ __forceinline__ void score_tile(float (&c)[4], const uint32_t (&a)[D / 8][4], const uint32_t* tile, int j,
                                           int g, int t) {
  constexpr int LD = D + 4;
  const uint32_t* row = tile + (j * 8 + g) * LD;
  c[0] = c[1] = c[2] = c[3] = 0.f;
#pragma unroll
  for (int kc = 0; kc < D / 8; ++kc) mma_op<X3>(c, a[kc], row[kc * 8 + t], row[kc * 8 + t + 4]);
}

// acc(16 x D) += P(16 x 8 cols of score tile j, given as A fragment) . tile[rows 8j+2t, 8j+2t+1][dims]
template <int D, bool X3>
__device__ __forceinline__ void accum_tile(float (&acc)[D / 8][4], const uint32_t (&pa)[4], const uint32_t* tile, int j,
                                           int g, int t) {
  constexpr int LD = D + 4;
  const uint32_t* r0 = tile + (j * 8 + 2 * t) * LD + g;
  const uint32_t* r1 = r0 + LD;
#pragma unroll
  for (int n = 0; n < D / 8; ++n) mma_op<X3>(acc[n], pa, r0[n * 8], r1[n * 8]);
}

__device__ __forceinline__ bool visible(int q, int k, int S, bool deg, bool key_is_pad) {
  return (k < S) && (q < S) && (deg || (k <= q && !key_is_pad));
}

// =========================================================================================
template <int D, bool X3>
__global__ void __launch_bounds__(kThreads) attn_fwd_mma(Dims p, const float* __restrict__ qkv,
                                                         const int32_t* __restrict__ x,
                                                         const int32_t* __restrict__ fnp, float* __restrict__ ctx,
                                                         float* __restrict__ lse) {
  constexpr int LD = D + 4, TILE = TQ * LD;
  extern __shared__ __align__(16) uint32_t smem_u[];
  // [buf][K | V] then kpad[buf][64]
  int* kpad_all = reinterpret_cast<int*>(smem_u + 4 * TILE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * TQ;
  const int64_t rs = 3 * (int64_t)p.E;
  const float* qbase = qkv + (int64_t)b * p.S * rs + head * p.d;
  const int first_real = fnp[b];
  const bool deg_block = q0 < first_real;

  stage_tile<D, X3>(smem_u, qbase, rs, q0, p.S, p.d, tid);
  __syncthreads();
  uint32_t qa[D / 8][4];
  load_a_frags<D>(qa, smem_u, warp, g, t);
  __syncthreads();

  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
  const bool deg0 = row0 < first_real, deg1 = row1 < first_real;
  float m0 = kNegBig, m1 = kNegBig, l0 = 0.f, l1 = 0.f;
  float o[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;

  const int n_all = (p.S + TQ - 1) / TQ;
  const int n_kv = deg_block ? n_all : min(n_all, (q0 + TQ - 1) / TQ + 1);
  auto issue = [&](int kb) {
    uint32_t* buf = smem_u + (kb & 1) * 2 * TILE;
    const int k0 = kb * TQ;
    issue_tile<D>(buf, qbase + p.E, rs, k0, p.S, p.d, tid);
    issue_tile<D>(buf + TILE, qbase + 2 * p.E, rs, k0, p.S, p.d, tid);
    cp_async_commit();
    if (tid < TQ) kpad_all[(kb & 1) * TQ + tid] = (k0 + tid < p.S) ? (x[(int64_t)b * p.S + k0 + tid] == 0) : 1;
  };
  issue(0);
  for (int kb = 0; kb < n_kv; ++kb) {
    const int k0 = kb * TQ;
    uint32_t* Ks = smem_u + (kb & 1) * 2 * TILE;
    uint32_t* Vs = Ks + TILE;
    const int* kpad = kpad_all + (kb & 1) * TQ;
    cp_async_wait_all();
    round_tile<D, X3>(Ks, tid);
    round_tile<D, X3>(Vs, tid);
    __syncthreads();
    if (kb + 1 < n_kv) issue(kb + 1);

    float s[8][4];
    float mx0 = kNegBig, mx1 = kNegBig;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      score_tile<D, X3>(s[j], qa, Ks, j, g, t);
      const int c0 = k0 + j * 8 + 2 * t;
      const bool pd0 = kpad[j * 8 + 2 * t], pd1 = kpad[j * 8 + 2 * t + 1];
      const bool v00 = visible(row0, c0, p.S, deg0, pd0), v01 = visible(row0, c0 + 1, p.S, deg0, pd1);
      const bool v10 = visible(row1, c0, p.S, deg1, pd0), v11 = visible(row1, c0 + 1, p.S, deg1, pd1);
      s[j][0] = v00 ? s[j][0] * p.scale : kNegBig;
      s[j][1] = v01 ? s[j][1] * p.scale : kNegBig;
      s[j][2] = v10 ? s[j][2] * p.scale : kNegBig;
      s[j][3] = v11 ? s[j][3] * p.scale : kNegBig;
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float al0 = expf(m0 - mn0), al1 = expf(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int n = 0; n < D / 8; ++n) { o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1; }
    float ot[X3 ? D / 8 : 1][4];
    tile_begin<X3>(ot);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // masked entries hold the sentinel: exp underflows to exactly 0 unless the whole row is still
      // at the sentinel (no visible key yet), which the explicit test guards
      const float p00 = s[j][0] > kNegBig ? expf(s[j][0] - mn0) : 0.f;
      const float p01 = s[j][1] > kNegBig ? expf(s[j][1] - mn0) : 0.f;
      const float p10 = s[j][2] > kNegBig ? expf(s[j][2] - mn1) : 0.f;
      const float p11 = s[j][3] > kNegBig ? expf(s[j][3] - mn1) : 0.f;
      rs0 += p00 + p01;
      rs1 += p10 + p11;
      const uint32_t pa[4] = {frag<X3>(p00), frag<X3>(p10), frag<X3>(p01), frag<X3>(p11)};
      accum_tile<D, X3>(pick_acc<X3>(ot, o), pa, Vs, j, g, t);
    }
    tile_end_acc<X3>(o, ot);
    l0 = l0 * al0 + quad_sum(rs0);
    l1 = l1 * al1 + quad_sum(rs1);
  }
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int col = n * 8 + 2 * t;
    if (row0 < p.S) {
      float* dst = ctx + ((int64_t)b * p.S + row0) * p.E + head * p.d;
      if (col < p.d) dst[col] = o[n][0] * inv0;
      if (col + 1 < p.d) dst[col + 1] = o[n][1] * inv0;
    }
    if (row1 < p.S) {
      float* dst = ctx + ((int64_t)b * p.S + row1) * p.E + head * p.d;
      if (col < p.d) dst[col] = o[n][2] * inv1;
      if (col + 1 < p.d) dst[col + 1] = o[n][3] * inv1;
    }
  }
  if (t == 0) {
    if (row0 < p.S) lse[((int64_t)b * p.h + head) * p.S + row0] = m0 + logf(l0);
    if (row1 < p.S) lse[((int64_t)b * p.h + head) * p.S + row1] = m1 + logf(l1);
  }
}

// =========================================================================================
// dK/dV: CTA owns 64 keys; warp owns 16 of them as MMA rows; loops over query tiles.
template <int D, bool X3>
__global__ void __launch_bounds__(kThreads) attn_bwd_dkdv_mma(Dims p, const float* __restrict__ qkv,
                                                              const int32_t* __restrict__ x,
                                                              const int32_t* __restrict__ fnp,
                                                              const float* __restrict__ dctx,
                                                              const float* __restrict__ lse,
                                                              const float* __restrict__ delta,
                                                              float* __restrict__ dqkv) {
  constexpr int LD = D + 4, TILE = TQ * LD;
  extern __shared__ __align__(16) uint32_t smem_u[];
  // [buf][Q | dO] then lse[buf][64], delta[buf][64]
  float* stat = reinterpret_cast<float*>(smem_u + 4 * TILE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, head = blockIdx.y, kbi = blockIdx.x, k0 = kbi * TQ;
  const int64_t rs = 3 * (int64_t)p.E;
  const float* qbase = qkv + (int64_t)b * p.S * rs + head * p.d;
  const float* dobase = dctx + (int64_t)b * p.S * p.E + head * p.d;
  const int first_real = fnp[b];
  const int64_t so = ((int64_t)b * p.h + head) * p.S;

  uint32_t ka[D / 8][4], va[D / 8][4];
  issue_tile<D>(smem_u, qbase + p.E, rs, k0, p.S, p.d, tid);
  issue_tile<D>(smem_u + TILE, qbase + 2 * p.E, rs, k0, p.S, p.d, tid);
  cp_async_commit();
  cp_async_wait_all();
  round_tile<D, X3>(smem_u, tid);
  round_tile<D, X3>(smem_u + TILE, tid);
  __syncthreads();
  load_a_frags<D>(ka, smem_u, warp, g, t);
  load_a_frags<D>(va, smem_u + TILE, warp, g, t);
  __syncthreads();

  const int key0 = k0 + warp * 16 + g, key1 = key0 + 8;
  const bool pad0 = key0 < p.S ? (x[(int64_t)b * p.S + key0] == 0) : true;
  const bool pad1 = key1 < p.S ? (x[(int64_t)b * p.S + key1] == 0) : true;

  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  // contributing query tiles: those at/after this key tile (causal) plus the leading tiles that hold
  // fully-masked rows (they see every key)
  const int n_q = (p.S + TQ - 1) / TQ;
  const int n_deg = min(min((first_real + TQ - 1) / TQ, n_q), kbi);
  const int n_it = n_deg + (n_q - kbi);
  auto tile_of = [&](int it) { return it < n_deg ? it : it - n_deg + kbi; };
  auto issue = [&](int it) {
    const int q0 = tile_of(it) * TQ;
    uint32_t* buf = smem_u + (it & 1) * 2 * TILE;
    issue_tile<D>(buf, qbase, rs, q0, p.S, p.d, tid);
    issue_tile<D>(buf + TILE, dobase, p.E, q0, p.S, p.d, tid);
    cp_async_commit();
    if (tid < TQ) {
      const int r = q0 + tid;
      stat[(it & 1) * 2 * TQ + tid] = r < p.S ? lse[so + r] : 0.f;
      stat[(it & 1) * 2 * TQ + TQ + tid] = r < p.S ? delta[so + r] : 0.f;
    }
  };
  issue(0);
  for (int it = 0; it < n_it; ++it) {
    const int q0 = tile_of(it) * TQ;
    uint32_t* Qs = smem_u + (it & 1) * 2 * TILE;
    uint32_t* dOs = Qs + TILE;
    const float* lse_s = stat + (it & 1) * 2 * TQ;
    const float* delta_s = lse_s + TQ;
    cp_async_wait_all();
    round_tile<D, X3>(Qs, tid);
    round_tile<D, X3>(dOs, tid);
    __syncthreads();
    if (it + 1 < n_it) issue(it + 1);
    float dvt[X3 ? D / 8 : 1][4], dkt[X3 ? D / 8 : 1][4];
    tile_begin<X3>(dvt);
    tile_begin<X3>(dkt);
#pragma unroll 4
    for (int j = 0; j < 8; ++j) {
      // transposed tiles: rows = this warp's keys, cols = queries q0+8j+2t, +1
      float st[4], dpt[4];
      score_tile<D, X3>(st, ka, Qs, j, g, t);           // S^T  = K Q^T
      score_tile<D, X3>(dpt, va, dOs, j, g, t);         // dP^T = V dO^T
      const int qc = q0 + j * 8 + 2 * t;
      const float ls0 = lse_s[j * 8 + 2 * t], ls1 = lse_s[j * 8 + 2 * t + 1];
      const float dl0 = delta_s[j * 8 + 2 * t], dl1 = delta_s[j * 8 + 2 * t + 1];
      const bool dg0 = qc < first_real, dg1 = (qc + 1) < first_real;
      const float p00 = visible(qc, key0, p.S, dg0, pad0) ? expf(st[0] * p.scale - ls0) : 0.f;
      const float p01 = visible(qc + 1, key0, p.S, dg1, pad0) ? expf(st[1] * p.scale - ls1) : 0.f;
      const float p10 = visible(qc, key1, p.S, dg0, pad1) ? expf(st[2] * p.scale - ls0) : 0.f;
      const float p11 = visible(qc + 1, key1, p.S, dg1, pad1) ? expf(st[3] * p.scale - ls1) : 0.f;
      const uint32_t pa[4] = {frag<X3>(p00), frag<X3>(p10), frag<X3>(p01), frag<X3>(p11)};
      const uint32_t da[4] = {frag<X3>(p00 * (dpt[0] - dl0)), frag<X3>(p10 * (dpt[2] - dl0)),
                              frag<X3>(p01 * (dpt[1] - dl1)), frag<X3>(p11 * (dpt[3] - dl1))};
      accum_tile<D, X3>(pick_acc<X3>(dvt, dv), pa, dOs, j, g, t);          // dV += P^T  dO
      accum_tile<D, X3>(pick_acc<X3>(dkt, dk), da, Qs, j, g, t);           // dK += dS^T Q
    }
    tile_end_acc<X3>(dv, dvt);
    tile_end_acc<X3>(dk, dkt);
  }
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int col = n * 8 + 2 * t;
    if (key0 < p.S) {
      float* out = dqkv + ((int64_t)b * p.S + key0) * rs + head * p.d;
      if (col < p.d) { out[p.E + col] = dk[n][0] * p.scale; out[2 * p.E + col] = dv[n][0]; }
      if (col + 1 < p.d) { out[p.E + col + 1] = dk[n][1] * p.scale; out[2 * p.E + col + 1] = dv[n][1]; }
    }
    if (key1 < p.S) {
      float* out = dqkv + ((int64_t)b * p.S + key1) * rs + head * p.d;
      if (col < p.d) { out[p.E + col] = dk[n][2] * p.scale; out[2 * p.E + col] = dv[n][2]; }
      if (col + 1 < p.d) { out[p.E + col + 1] = dk[n][3] * p.scale; out[2 * p.E + col + 1] = dv[n][3]; }
    }
  }
}

// =========================================================================================
// dQ: CTA owns 64 queries; loops over key tiles (same range as the forward).
template <int D, bool X3>
__global__ void __launch_bounds__(kThreads) attn_bwd_dq_mma(Dims p, const float* __restrict__ qkv,
                                                            const int32_t* __restrict__ x,
                                                            const int32_t* __restrict__ fnp,
                                                            const float* __restrict__ dctx,
                                                            const float* __restrict__ lse,
                                                            const float* __restrict__ delta,
                                                            float* __restrict__ dqkv) {
  constexpr int LD = D + 4, TILE = TQ * LD;
  extern __shared__ __align__(16) uint32_t smem_u[];
  int* kpad_all = reinterpret_cast<int*>(smem_u + 4 * TILE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * TQ;
  const int64_t rs = 3 * (int64_t)p.E;
  const float* qbase = qkv + (int64_t)b * p.S * rs + head * p.d;
  const float* dobase = dctx + (int64_t)b * p.S * p.E + head * p.d;
  const int first_real = fnp[b];
  const bool deg_block = q0 < first_real;

  uint32_t qa[D / 8][4], doa[D / 8][4];
  issue_tile<D>(smem_u, qbase, rs, q0, p.S, p.d, tid);
  issue_tile<D>(smem_u + TILE, dobase, p.E, q0, p.S, p.d, tid);
  cp_async_commit();
  cp_async_wait_all();
  round_tile<D, X3>(smem_u, tid);
  round_tile<D, X3>(smem_u + TILE, tid);
  __syncthreads();
  load_a_frags<D>(qa, smem_u, warp, g, t);
  load_a_frags<D>(doa, smem_u + TILE, warp, g, t);
  __syncthreads();

  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
  const bool deg0 = row0 < first_real, deg1 = row1 < first_real;
  const int64_t so = ((int64_t)b * p.h + head) * p.S;
  const float ls0 = row0 < p.S ? lse[so + row0] : 0.f, ls1 = row1 < p.S ? lse[so + row1] : 0.f;
  const float dl0 = row0 < p.S ? delta[so + row0] : 0.f, dl1 = row1 < p.S ? delta[so + row1] : 0.f;

  float dq[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;

  const int n_all = (p.S + TQ - 1) / TQ;
  const int n_kv = deg_block ? n_all : min(n_all, (q0 + TQ - 1) / TQ + 1);
  auto issue = [&](int kb) {
    uint32_t* buf = smem_u + (kb & 1) * 2 * TILE;
    const int k0 = kb * TQ;
    issue_tile<D>(buf, qbase + p.E, rs, k0, p.S, p.d, tid);
    issue_tile<D>(buf + TILE, qbase + 2 * p.E, rs, k0, p.S, p.d, tid);
    cp_async_commit();
    if (tid < TQ) kpad_all[(kb & 1) * TQ + tid] = (k0 + tid < p.S) ? (x[(int64_t)b * p.S + k0 + tid] == 0) : 1;
  };
  issue(0);
  for (int kb = 0; kb < n_kv; ++kb) {
    const int k0 = kb * TQ;
    uint32_t* Ks = smem_u + (kb & 1) * 2 * TILE;
    uint32_t* Vs = Ks + TILE;
    const int* kpad = kpad_all + (kb & 1) * TQ;
    cp_async_wait_all();
    round_tile<D, X3>(Ks, tid);
    round_tile<D, X3>(Vs, tid);
    __syncthreads();
    if (kb + 1 < n_kv) issue(kb + 1);
    float dqt[X3 ? D / 8 : 1][4];
    tile_begin<X3>(dqt);
#pragma unroll 4
    for (int j = 0; j < 8; ++j) {
      float s[4], dp[4];
      score_tile<D, X3>(s, qa, Ks, j, g, t);            // S  = Q K^T
      score_tile<D, X3>(dp, doa, Vs, j, g, t);          // dP = dO V^T
      const int c0 = k0 + j * 8 + 2 * t;
      const bool pd0 = kpad[j * 8 + 2 * t], pd1 = kpad[j * 8 + 2 * t + 1];
      const float p00 = visible(row0, c0, p.S, deg0, pd0) ? expf(s[0] * p.scale - ls0) : 0.f;
      const float p01 = visible(row0, c0 + 1, p.S, deg0, pd1) ? expf(s[1] * p.scale - ls0) : 0.f;
      const float p10 = visible(row1, c0, p.S, deg1, pd0) ? expf(s[2] * p.scale - ls1) : 0.f;
      const float p11 = visible(row1, c0 + 1, p.S, deg1, pd1) ? expf(s[3] * p.scale - ls1) : 0.f;
      const uint32_t da[4] = {frag<X3>(p00 * (dp[0] - dl0)), frag<X3>(p10 * (dp[2] - dl1)),
                              frag<X3>(p01 * (dp[1] - dl0)), frag<X3>(p11 * (dp[3] - dl1))};
      accum_tile<D, X3>(pick_acc<X3>(dqt, dq), da, Ks, j, g, t);           // dQ += dS K
    }
    tile_end_acc<X3>(dq, dqt);
  }
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int col = n * 8 + 2 * t;
    if (row0 < p.S) {
      float* out = dqkv + ((int64_t)b * p.S + row0) * rs + head * p.d;
      if (col < p.d) out[col] = dq[n][0] * p.scale;
      if (col + 1 < p.d) out[col + 1] = dq[n][1] * p.scale;
    }
    if (row1 < p.S) {
      float* out = dqkv + ((int64_t)b * p.S + row1) * rs + head * p.d;
      if (col < p.d) out[col] = dq[n][2] * p.scale;
      if (col + 1 < p.d) out[col + 1] = dq[n][3] * p.scale;
    }
  }
}

template <int D> constexpr size_t smem_bytes() { return (size_t)(4 * TQ * (D + 4)) * 4 + 4 * TQ * 4; }

template <int D, bool X3>
int launch_fwd(cudaStream_t st, const Dims& p, const float* qkv, const int32_t* x, const int32_t* fnp, float* ctx,
               float* lse) {
  static bool once = false;
  if (!once) {
    NLPS_CUDA(cudaFuncSetAttribute(attn_fwd_mma<D, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<D>()));
    once = true;
  }
  dim3 grid((unsigned)ceil_div(p.S, TQ), (unsigned)p.h, (unsigned)p.B);
  attn_fwd_mma<D, X3><<<grid, kThreads, smem_bytes<D>(), st>>>(p, qkv, x, fnp, ctx, lse);
  NLPS_LAUNCH_CHECK();
  return 0;
}

template <int D, bool X3>
int launch_bwd(cudaStream_t st, const Dims& p, const float* qkv, const int32_t* x, const int32_t* fnp,
               const float* dctx, const float* lse, const float* delta, float* dqkv) {
  static bool once = false;
  if (!once) {
    NLPS_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_mma<D, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<D>()));
    NLPS_CUDA(cudaFuncSetAttribute(attn_bwd_dq_mma<D, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<D>()));
    once = true;
  }
  dim3 grid((unsigned)ceil_div(p.S, TQ), (unsigned)p.h, (unsigned)p.B);
  attn_bwd_dkdv_mma<D, X3><<<grid, kThreads, smem_bytes<D>(), st>>>(p, qkv, x, fnp, dctx, lse, delta, dqkv);
  NLPS_LAUNCH_CHECK();
  attn_bwd_dq_mma<D, X3><<<grid, kThreads, smem_bytes<D>(), st>>>(p, qkv, x, fnp, dctx, lse, delta, dqkv);
  NLPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int attention_forward_mma(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                          const int32_t* fnp, float* ctx, float* lse, bool x3) {
  Dims p{B, S, h, d, h * d, (float)(1.0 / std::sqrt((double)d))};
  NLPS_REQUIRE(d <= 64 && d % 4 == 0, "attention_forward_mma: head_dim %d is served by the CUDA-core kernels", d);
  if (x3) {
    if (d <= 16) return launch_fwd<16, true>(st, p, qkv, x, fnp, ctx, lse);
    if (d <= 32) return launch_fwd<32, true>(st, p, qkv, x, fnp, ctx, lse);
    return launch_fwd<64, true>(st, p, qkv, x, fnp, ctx, lse);
  }
  if (d <= 16) return launch_fwd<16, false>(st, p, qkv, x, fnp, ctx, lse);
  if (d <= 32) return launch_fwd<32, false>(st, p, qkv, x, fnp, ctx, lse);
  return launch_fwd<64, false>(st, p, qkv, x, fnp, ctx, lse);
}

int attention_backward_mma(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                           const int32_t* fnp, const float* dctx, const float* lse, const float* delta,
                           float* dqkv, bool x3) {
  Dims p{B, S, h, d, h * d, (float)(1.0 / std::sqrt((double)d))};
  NLPS_REQUIRE(d <= 64 && d % 4 == 0, "attention_backward_mma: head_dim %d is served by the CUDA-core kernels", d);
  if (x3) {
    if (d <= 16) return launch_bwd<16, true>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
    if (d <= 32) return launch_bwd<32, true>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
    return launch_bwd<64, true>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
  }
  if (d <= 16) return launch_bwd<16, false>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
  if (d <= 32) return launch_bwd<32, false>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
  return launch_bwd<64, false>(st, p, qkv, x, fnp, dctx, lse, delta, dqkv);
}

}  // namespace nlps
