// Bandwidth-bound row kernels: affine LayerNorm forward/backward (function.py:90-124), softmax
// cross-entropy with its gradient (transformer.py:515-549), column sums (bias gradients) and the
// sum of squares behind the global-norm clip (function.py:27-38).
//
// All of them are one pass over HBM per tensor where the algorithm allows it: a warp owns a row,
// keeps it in registers (float4 loads, 16 B per lane per request) and reduces with shuffles.
// Cross-row reductions (dgamma/dbeta, bias grads, loss, norms) are two-stage with a fixed
// partition and a fixed summation order: deterministic, no floating-point atomics.
#include "common.cuh"
#include <algorithm>
#include <cmath>

namespace nlps {
namespace {

constexpr int kLnWarps = 8;                 // rows in flight per CTA
constexpr int kLnPartials = kNumSMs * 2;    // CTAs of the LayerNorm backward (fixed => deterministic)

// ---------------- LayerNorm forward: warp per row, row cached in registers ----------------
// VEC = float4 per lane (E == VEC*128).  VEC == 0: generic E, three passes over L1-resident row.
template <int VEC>
__global__ void __launch_bounds__(kLnWarps * 32) ln_fwd_kernel(int64_t rows, int E, const float* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ y,
                                                               float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * E;
  float* yr = y + row * E;
  const float invE = 1.f / (float)E;
  if constexpr (VEC > 0) {
    float4 v[VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mu = warp_sum(s) * invE;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float var = warp_sum(q) * invE;
    const float rstd = 1.f / sqrtf(var + kLnEps);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float4 bt = *reinterpret_cast<const float4*>(beta + c);
      float4 o;
      o.x = (v[i].x - mu) * rstd * g.x + bt.x;
      o.y = (v[i].y - mu) * rstd * g.y + bt.y;
      o.z = (v[i].z - mu) * rstd * g.z + bt.z;
      o.w = (v[i].w - mu) * rstd * g.w + bt.w;
      *reinterpret_cast<float4*>(yr + c) = o;
    }
    if (lane == 0) { mean_out[row] = mu; rstd_out[row] = rstd; }
  } else {
    float s = 0.f;
    for (int c = lane; c < E; c += 32) s += xr[c];
    const float mu = warp_sum(s) * invE;
    float q = 0.f;
    for (int c = lane; c < E; c += 32) { const float a = xr[c] - mu; q += a * a; }
    const float rstd = 1.f / sqrtf(warp_sum(q) * invE + kLnEps);
    for (int c = lane; c < E; c += 32) yr[c] = (xr[c] - mu) * rstd * gamma[c] + beta[c];
    if (lane == 0) { mean_out[row] = mu; rstd_out[row] = rstd; }
  }
}

// ---------------- LayerNorm backward ----------------
// dx = (dxh - mean(dxh) - xh*mean(dxh*xh)) * rstd with dxh = dy*gamma, xh = (x-mean)*rstd.
// Each CTA walks a fixed slab of rows, accumulating dgamma/dbeta per column in registers;
// the per-CTA partials are summed in CTA order by ln_bwd_finish_kernel.
// If drop != nullptr a second output dx_dropped = dx * mask / keep is written (dropout backward of
// the sub-layer branch, transformer.py:377-379 / :416-418) with the forward's Philox mask.
template <int VEC>
__global__ void __launch_bounds__(kLnWarps * 32) ln_bwd_kernel(int64_t rows, int E, const float* __restrict__ dy,
                                                               const float* __restrict__ x,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma,
                                                               float* __restrict__ dx, float* __restrict__ dx_dropped,
                                                               const DropoutParams drop, int use_drop,
                                                               float* __restrict__ partial, int nsum) {
  // nsum == 3: a third column sum rides along -- of the branch gradient that leaves this kernel (dx_dropped, or dx when
  // there is no dropout): the bias gradient of the sub-layer's last projection (db2) without another pass
  extern __shared__ float red[];     // [kLnWarps][nsum][E]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invE = 1.f / (float)E;
  const int64_t rows_per_cta = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  if constexpr (VEC > 0) {
    float4 g[VEC], ag[VEC], ab[VEC], ad[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      g[i] = *reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4);
      ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      ad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t row = r0 + warp; row < r1; row += kLnWarps) {
      const float mu = mean[row], rs = rstd[row];
      float4 xh[VEC], dh[VEC];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 xv = *reinterpret_cast<const float4*>(x + row * E + c);
        const float4 dv = *reinterpret_cast<const float4*>(dy + row * E + c);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        dh[i] = make_float4(dv.x * g[i].x, dv.y * g[i].y, dv.z * g[i].z, dv.w * g[i].w);
        s1 += (dh[i].x + dh[i].y) + (dh[i].z + dh[i].w);
        s2 += (dh[i].x * xh[i].x + dh[i].y * xh[i].y) + (dh[i].z * xh[i].z + dh[i].w * xh[i].w);
        ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
        ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
      }
      s1 = warp_sum(s1) * invE;
      s2 = warp_sum(s2) * invE;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 o;
        o.x = (dh[i].x - s1 - xh[i].x * s2) * rs;
        o.y = (dh[i].y - s1 - xh[i].y * s2) * rs;
        o.z = (dh[i].z - s1 - xh[i].z * s2) * rs;
        o.w = (dh[i].w - s1 - xh[i].w * s2) * rs;
        *reinterpret_cast<float4*>(dx + row * E + c) = o;
        if (use_drop) {
          dropout_four(drop, (uint64_t)row * (uint64_t)E + (uint64_t)c, o);
          *reinterpret_cast<float4*>(dx_dropped + row * E + c) = o;
        }
        ad[i].x += o.x; ad[i].y += o.y; ad[i].z += o.z; ad[i].w += o.w;
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      *reinterpret_cast<float4*>(red + (warp * nsum + 0) * E + c) = ag[i];
      *reinterpret_cast<float4*>(red + (warp * nsum + 1) * E + c) = ab[i];
      if (nsum == 3) *reinterpret_cast<float4*>(red + (warp * nsum + 2) * E + c) = ad[i];
    }
  } else {
    for (int c = lane; c < E; c += 32)
      for (int k = 0; k < nsum; ++k) red[(warp * nsum + k) * E + c] = 0.f;
    for (int64_t row = r0 + warp; row < r1; row += kLnWarps) {
      const float mu = mean[row], rs = rstd[row];
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < E; c += 32) {
        const float xh = (x[row * E + c] - mu) * rs, dv = dy[row * E + c], dh = dv * gamma[c];
        s1 += dh; s2 += dh * xh;
        red[(warp * nsum + 0) * E + c] += dv * xh;
        red[(warp * nsum + 1) * E + c] += dv;
      }
      s1 = warp_sum(s1) * invE;
      s2 = warp_sum(s2) * invE;
      for (int c = lane; c < E; c += 32) {
        const float xh = (x[row * E + c] - mu) * rs, dh = dy[row * E + c] * gamma[c];
        float o = (dh - s1 - xh * s2) * rs;
        dx[row * E + c] = o;
        if (use_drop) { o = dropout_one(drop, (uint64_t)row * (uint64_t)E + (uint64_t)c, o); dx_dropped[row * E + c] = o; }
        if (nsum == 3) red[(warp * nsum + 2) * E + c] += o;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < nsum * E; c += blockDim.x) {
    const int which = c / E, col = c - which * E;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) acc += red[(w * nsum + which) * E + col];
    partial[((int64_t)blockIdx.x * nsum + which) * E + col] = acc;
  }
}

// block = 32 columns x 8 partial-slices; fixed summation order (slice-major) => deterministic
__global__ void __launch_bounds__(256) ln_bwd_finish_kernel(int nparts, int E, int nsum, const float* __restrict__ partial,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ dbias) {
  __shared__ float sm[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;               // column in the concatenated [dgamma | dbeta | dbias] row
  const int W = nsum * E;
  float acc = 0.f;
  if (c < W) {
    int p = warp;
    for (; p + 24 < nparts; p += 32) {                  // four loads in flight, added in slice order
      const float a = partial[(int64_t)p * W + c], b = partial[(int64_t)(p + 8) * W + c];
      const float d = partial[(int64_t)(p + 16) * W + c], e = partial[(int64_t)(p + 24) * W + c];
      acc += a; acc += b; acc += d; acc += e;
    }
    for (; p < nparts; p += 8) acc += partial[(int64_t)p * W + c];
  }
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < W) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    if (c < E) dgamma[c] = t;
    else if (c < 2 * E) dbeta[c - E] = t;
    else dbias[c - 2 * E] = t;
  }
}

// ---------------- softmax cross-entropy: CTA per row, online (max,sum), 2 reads + 1 write ------
__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  s = s * expf(m - mn) + s2 * expf(m2 - mn);
  m = mn;
}

// Target id of a row as NumPy's fancy index reads it (transformer.py:538-546 `log_probs[rows, y]`): negative ids wrap once
// (-V <= y < 0 means y + V); anything else out of range is an IndexError in the reference -- here the id is clamped to 0 so
// that no access leaves the row, and bit 1 of the model's error word is raised (nlps_check_errors turns it into IndexError).
__device__ __forceinline__ int xent_target(const int32_t* __restrict__ y, int64_t row, int V, int32_t* err) {
  int t = y[row];
  if (t < 0) t += V;
  if (t < 0 || t >= V) {
    if (err != nullptr) atomicOr(err, 2);
    t = 0;
  }
  return t;
}

// Cross-entropy whose (max, sum exp) come from the vocab-head GEMM's epilogue as one pair per 32 columns: one read of
// the logits instead of two.  Same merge and the same second phase as xent_kernel.
__global__ void __launch_bounds__(256) xent_stats_kernel(const float* logits, int64_t ld, const int32_t* __restrict__ y, int V,
                                                         float inv_rows_scaled, float* dlogits, int64_t ldd,
                                                         float* __restrict__ row_loss, const float2* __restrict__ rowstat, int32_t* err) {
  __shared__ float sm_m[8], sm_s[8];
  __shared__ float bc[2];
  const int64_t row = blockIdx.x;
  const float* z = logits + row * ld;
  float* dz = dlogits + row * ldd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (V + 31) >> 5;
  float m = -INFINITY, s = 0.f;
  for (int k = tid; k < nblk; k += 256) {
    const float2 st = rowstat[row * nblk + k];
    if (m > -INFINITY) online_merge(m, s, st.x, st.y); else { m = st.x; s = st.y; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    if (m2 > -INFINITY) { if (m > -INFINITY) online_merge(m, s, m2, s2); else { m = m2; s = s2; } }
  }
  if (lane == 0) { sm_m[warp] = m; sm_s[warp] = s; }
  __syncthreads();
  if (tid == 0) {
    float M = sm_m[0], S = sm_s[0];
    for (int w = 1; w < 8; ++w)
      if (sm_m[w] > -INFINITY) { if (M > -INFINITY) online_merge(M, S, sm_m[w], sm_s[w]); else { M = sm_m[w]; S = sm_s[w]; } }
    bc[0] = M; bc[1] = S;
    const int t = xent_target(y, row, V, err);
    row_loss[row] = -(z[t] - M - logf(S));
  }
  __syncthreads();
  const float M = bc[0], invS = 1.f / bc[1];
  const int target = xent_target(y, row, V, nullptr);
  for (int c = tid * 4; c < V; c += 1024) {                // V % 4 == 0 and 16-B aligned rows (checked by the launcher)
    const float4 v = *reinterpret_cast<const float4*>(z + c);
    float4 o;
    o.x = (expf(v.x - M) * invS - (c + 0 == target ? 1.f : 0.f)) * inv_rows_scaled;
    o.y = (expf(v.y - M) * invS - (c + 1 == target ? 1.f : 0.f)) * inv_rows_scaled;
    o.z = (expf(v.z - M) * invS - (c + 2 == target ? 1.f : 0.f)) * inv_rows_scaled;
    o.w = (expf(v.w - M) * invS - (c + 3 == target ? 1.f : 0.f)) * inv_rows_scaled;
    *reinterpret_cast<float4*>(dz + c) = o;
  }
}

__global__ void __launch_bounds__(256) xent_kernel(const float* logits, int64_t ld,
                                                   const int32_t* __restrict__ y, int V, float inv_rows_scaled,
                                                   float* dlogits, int64_t ldd,
                                                   float* __restrict__ row_loss, int vec_ok, int32_t* err) {
  __shared__ float sm_m[8], sm_s[8];
  __shared__ float bc[2];
  const int64_t row = blockIdx.x;
  const float* z = logits + row * ld;
  float* dz = dlogits + row * ldd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -INFINITY, s = 0.f;
  if (vec_ok) {
    for (int c = tid * 4; c < V; c += 1024) {
      const float4 v = *reinterpret_cast<const float4*>(z + c);
      const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      if (mx > m) { s *= expf(m - mx); m = mx; }
      s += (expf(v.x - m) + expf(v.y - m)) + (expf(v.z - m) + expf(v.w - m));
    }
  } else {
    for (int c = tid; c < V; c += 256) {
      const float v = z[c];
      if (v > m) { s *= expf(m - v); m = v; }
      s += expf(v - m);
    }
  }
  // threads that saw no element keep (m=-inf, s=0): merge must not produce NaN
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    if (m2 > -INFINITY) { if (m > -INFINITY) online_merge(m, s, m2, s2); else { m = m2; s = s2; } }
  }
  if (lane == 0) { sm_m[warp] = m; sm_s[warp] = s; }
  __syncthreads();
  if (tid == 0) {
    float M = sm_m[0], S = sm_s[0];
    for (int w = 1; w < 8; ++w)
      if (sm_m[w] > -INFINITY) { if (M > -INFINITY) online_merge(M, S, sm_m[w], sm_s[w]); else { M = sm_m[w]; S = sm_s[w]; } }
    bc[0] = M; bc[1] = S;
    const int t = xent_target(y, row, V, err);
    // loss_row = -(z_t - max - log(sum))
    row_loss[row] = -(z[t] - M - logf(S));
  }
  __syncthreads();
  const float M = bc[0], invS = 1.f / bc[1];
  const int target = xent_target(y, row, V, nullptr);
  if (vec_ok) {
    for (int c = tid * 4; c < V; c += 1024) {
      const float4 v = *reinterpret_cast<const float4*>(z + c);
      float4 o;
      o.x = (expf(v.x - M) * invS - (c + 0 == target ? 1.f : 0.f)) * inv_rows_scaled;
      o.y = (expf(v.y - M) * invS - (c + 1 == target ? 1.f : 0.f)) * inv_rows_scaled;
      o.z = (expf(v.z - M) * invS - (c + 2 == target ? 1.f : 0.f)) * inv_rows_scaled;
      o.w = (expf(v.w - M) * invS - (c + 3 == target ? 1.f : 0.f)) * inv_rows_scaled;
      *reinterpret_cast<float4*>(dz + c) = o;
    }
  } else {
    for (int c = tid; c < V; c += 256)
      dz[c] = (expf(z[c] - M) * invS - (c == target ? 1.f : 0.f)) * inv_rows_scaled;
  }
}

// Cross-entropy with the vocabulary-bias gradient riding along (db_vocab = column sums of dlogits, transformer.py:353):
// a CTA walks kXentRows consecutive rows with the arithmetic of xent_kernel and keeps, per thread, the running column sums
// of the dlogits it writes (thread t owns columns 4t + 1024k); at the end it stores one partial row.  The column-sum pass
// over the (N, V) dlogits (1.07 GB at c4) becomes a column-sum pass over the (N/16, V) partials.  Fixed order: deterministic.
constexpr int kXentRows = 16;
constexpr int kXentMaxChunks = 16;      // V <= 16384
template <int CHUNKS>
__global__ void __launch_bounds__(256) xent_colsum_kernel(const float* logits, int64_t ld, const int32_t* __restrict__ y, int V,
                                                          int64_t rows, float inv_rows_scaled, float* dlogits, int64_t ldd,
                                                          float* __restrict__ row_loss, float* __restrict__ partial,
                                                          int64_t ld_partial, int32_t* err) {
  __shared__ float sm_m[8], sm_s[8];
  __shared__ float bc[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * kXentRows, r1 = min(rows, r0 + kXentRows);
  float4 acc[CHUNKS];
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = r0; row < r1; ++row) {
    const float* z = logits + row * ld;
    float* dz = dlogits + row * ldd;
    float m = -INFINITY, s = 0.f;
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
      const int c = tid * 4 + k * 1024;
      if (c < V) {
        const float4 v = *reinterpret_cast<const float4*>(z + c);
        const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (mx > m) { s *= expf(m - mx); m = mx; }
        s += (expf(v.x - m) + expf(v.y - m)) + (expf(v.z - m) + expf(v.w - m));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      if (m2 > -INFINITY) { if (m > -INFINITY) online_merge(m, s, m2, s2); else { m = m2; s = s2; } }
    }
    if (lane == 0) { sm_m[warp] = m; sm_s[warp] = s; }
    __syncthreads();
    if (tid == 0) {
      float M = sm_m[0], S = sm_s[0];
      for (int w = 1; w < 8; ++w)
        if (sm_m[w] > -INFINITY) { if (M > -INFINITY) online_merge(M, S, sm_m[w], sm_s[w]); else { M = sm_m[w]; S = sm_s[w]; } }
      bc[0] = M; bc[1] = S;
      const int t = xent_target(y, row, V, err);
      row_loss[row] = -(z[t] - M - logf(S));
    }
    __syncthreads();
    const float M = bc[0], invS = 1.f / bc[1];
    const int target = xent_target(y, row, V, nullptr);
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
      const int c = tid * 4 + k * 1024;
      if (c < V) {
        const float4 v = *reinterpret_cast<const float4*>(z + c);
        float4 o;
        o.x = (expf(v.x - M) * invS - (c + 0 == target ? 1.f : 0.f)) * inv_rows_scaled;
        o.y = (expf(v.y - M) * invS - (c + 1 == target ? 1.f : 0.f)) * inv_rows_scaled;
        o.z = (expf(v.z - M) * invS - (c + 2 == target ? 1.f : 0.f)) * inv_rows_scaled;
        o.w = (expf(v.w - M) * invS - (c + 3 == target ? 1.f : 0.f)) * inv_rows_scaled;
        *reinterpret_cast<float4*>(dz + c) = o;
        acc[k].x += o.x; acc[k].y += o.y; acc[k].z += o.z; acc[k].w += o.w;
      }
    }
    __syncthreads();          // bc / sm_* are rewritten by the next row
  }
  float* prow = partial + (int64_t)blockIdx.x * ld_partial;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    const int c = tid * 4 + k * 1024;
    if (c < V) *reinterpret_cast<float4*>(prow + c) = acc[k];
  }
}

// out[0] = scale * sum(x[0..n)) with a fixed tree: one CTA, deterministic
__global__ void __launch_bounds__(1024) sum_to_scalar_kernel(const float* __restrict__ x, int64_t n, float scale,
                                                             float* __restrict__ out) {
  __shared__ float sm[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = sm[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v * scale;
  }
}

// ---------------- column sums (bias gradients): two-stage ----------------
constexpr int kColSlabsMax = 256;
// block = 32 lanes x 8 warps; a lane owns 4 adjacent columns (float4), a warp walks rows of its slab
// with 4 independent loads in flight; grid = (ceil(cols/128), slabs).  Fixed partition => deterministic.
// The LAST slab CTA of a 128-column group (ticket counter in global memory, reset by its taker) adds the slab
// partials in the fixed order of colsum_finish_kernel: same bits, one launch and one drain/fill bubble fewer.
__global__ void __launch_bounds__(256) colsum_partial_kernel(int64_t rows, int cols, const float* __restrict__ x,
                                                             int64_t ld, float* __restrict__ partial, int vec,
                                                             unsigned int* __restrict__ tickets, float* __restrict__ out) {
  __shared__ float4 sm[8][32];
  __shared__ unsigned int ticket_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + lane) * 4;
  const int64_t rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec) {
    if (col < cols) {
      int64_t r = r0 + warp;
      for (; r + 24 < r1; r += 32) {
        const float4 a = *reinterpret_cast<const float4*>(x + r * ld + col);
        const float4 b = *reinterpret_cast<const float4*>(x + (r + 8) * ld + col);
        const float4 c = *reinterpret_cast<const float4*>(x + (r + 16) * ld + col);
        const float4 d = *reinterpret_cast<const float4*>(x + (r + 24) * ld + col);
        acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
        acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
      }
      for (; r < r1; r += 8) {
        const float4 a = *reinterpret_cast<const float4*>(x + r * ld + col);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
    }
  } else {
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      if (col + 0 < cols) acc.x += x[r * ld + col + 0];
      if (col + 1 < cols) acc.y += x[r * ld + col + 1];
      if (col + 2 < cols) acc.z += x[r * ld + col + 2];
      if (col + 3 < cols) acc.w += x[r * ld + col + 3];
    }
  }
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; ++w) { const float4 u = sm[w][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    float* dst = partial + (int64_t)blockIdx.y * cols + col;
    if (col + 0 < cols) dst[0] = t.x;
    if (col + 1 < cols) dst[1] = t.y;
    if (col + 2 < cols) dst[2] = t.z;
    if (col + 3 < cols) dst[3] = t.w;
  }
  if (tickets == nullptr) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket_s = atomicAdd(tickets + blockIdx.x, 1u);
  __syncthreads();
  if (ticket_s != gridDim.y - 1) return;
  __threadfence();
  // slab-major sums exactly like colsum_finish_kernel: warp w adds slabs w, w+8, ..., then the 8 warp sums in order
  const int nparts = (int)gridDim.y;
  float4 acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = warp; p < nparts; p += 8) {
    const volatile float* src = partial + (int64_t)p * cols + col;
    if (col + 0 < cols) acc2.x += src[0];
    if (col + 1 < cols) acc2.y += src[1];
    if (col + 2 < cols) acc2.z += src[2];
    if (col + 3 < cols) acc2.w += src[3];
  }
  __syncthreads();
  sm[warp][lane] = acc2;
  __syncthreads();
  if (warp == 0) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; ++w) { const float4 u = sm[w][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    if (col + 0 < cols) out[col + 0] = t.x;
    if (col + 1 < cols) out[col + 1] = t.y;
    if (col + 2 < cols) out[col + 2] = t.z;
    if (col + 3 < cols) out[col + 3] = t.w;
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;     // ready for the next launch on this scratch
}
// block = 32 columns x 8 partial-slices
__global__ void __launch_bounds__(256) colsum_finish_kernel(int nparts, int cols, const float* __restrict__ partial,
                                                            float* __restrict__ out) {
  __shared__ float sm[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < cols)
    for (int p = warp; p < nparts; p += 8) acc += partial[(int64_t)p * cols + c];
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    out[c] = t;
  }
}

// ---------------- sum of squares: two-stage ----------------
constexpr int kSumsqBlocks = kNumSMs * 4;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(int64_t n, const float* __restrict__ x,
                                                            float* __restrict__ partial) {
  __shared__ float sm[8];
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const int64_t stride = (int64_t)gridDim.x * 256;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {       // 4 independent 16-B loads in flight per thread
    const float4 a = x4[i], b = x4[i + stride], c = x4[i + 2 * stride], d = x4[i + 3 * stride];
    acc += ((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w)) + ((b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w)) +
           ((c.x * c.x + c.y * c.y) + (c.z * c.z + c.w * c.w)) + ((d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w));
  }
  for (; i < n4; i += stride) {
    const float4 v = x4[i];
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += 256) acc += x[i] * x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sm[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void sumsq_scalar_partial_kernel(int64_t n, const float* __restrict__ x, float* __restrict__ partial) {
  // unaligned base: scalar loads
  __shared__ float sm[8];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc += x[i] * x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sm[w];
    partial[blockIdx.x] = t;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int layernorm_forward(cudaStream_t st, int64_t rows, int E, const float* x, const float* gamma, const float* beta,
                      float* y, float* mean, float* rstd) {
  if (rows <= 0) return 0;
  const unsigned grid = (unsigned)ceil_div(rows, kLnWarps);
  const bool vec = (E % 128 == 0) && E <= 1024 && aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta);
#define NLPS_LN_FWD(V) ln_fwd_kernel<V><<<grid, kLnWarps * 32, 0, st>>>(rows, E, x, gamma, beta, y, mean, rstd)
  if (!vec) NLPS_LN_FWD(0);
  else switch (E / 128) {
    case 1: NLPS_LN_FWD(1); break;
    case 2: NLPS_LN_FWD(2); break;
    case 3: NLPS_LN_FWD(3); break;
    case 4: NLPS_LN_FWD(4); break;
    case 5: NLPS_LN_FWD(5); break;
    case 6: NLPS_LN_FWD(6); break;
    case 7: NLPS_LN_FWD(7); break;
    default: NLPS_LN_FWD(8); break;
  }
#undef NLPS_LN_FWD
  NLPS_LAUNCH_CHECK();
  return 0;
}

size_t layernorm_backward_scratch_floats(int E) { return (size_t)kLnPartials * 3 * E; }

// second half of layernorm_backward when it was called with defer_finish: the fixed-order sum of the per-CTA partials
// (the parameter gradients are off the critical path of backward: the model runs this on its weight-gradient stream)
int layernorm_backward_finish(cudaStream_t st, int64_t rows, int E, const float* partial, float* dgamma, float* dbeta,
                              float* dbias) {
  if (rows <= 0) return 0;
  const int grid = (int)std::min<int64_t>(kLnPartials, ceil_div(rows, kLnWarps));
  const int nsum = dbias != nullptr ? 3 : 2;
  ln_bwd_finish_kernel<<<(unsigned)ceil_div(nsum * E, 32), 256, 0, st>>>(grid, E, nsum, partial, dgamma, dbeta, dbias);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int layernorm_backward(cudaStream_t st, int64_t rows, int E, const float* dy, const float* x, const float* mean,
                       const float* rstd, const float* gamma, float* dx, float* dx_dropped,
                       const DropoutParams* drop, float* dgamma, float* dbeta, float* partial, float* dbias,
                       bool defer_finish) {
  if (rows <= 0) return 0;
  const int grid = (int)std::min<int64_t>(kLnPartials, ceil_div(rows, kLnWarps));
  const int use_drop = (drop != nullptr && drop->p > 0.f && dx_dropped != nullptr) ? 1 : 0;
  DropoutParams dp = drop ? *drop : make_dropout(0.f, 0, 0, 0);
  const bool vec = (E % 128 == 0) && E <= 1024 && aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(gamma) &&
                   (!use_drop || aligned16(dx_dropped));
  const int nsum = dbias != nullptr ? 3 : 2;
  const size_t smem = (size_t)kLnWarps * nsum * E * sizeof(float);
#define NLPS_LN_BWD(V)                                                                                         \
  {                                                                                                            \
    static size_t smem_set = 0;                                                                                \
    if (smem > 48 * 1024 && smem > smem_set) {                                                                 \
      NLPS_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      smem_set = smem;                                                                                         \
    }                                                                                                          \
    ln_bwd_kernel<V><<<grid, kLnWarps * 32, smem, st>>>(rows, E, dy, x, mean, rstd, gamma, dx, dx_dropped, dp, \
                                                        use_drop, partial, nsum);                              \
  }
  NLPS_REQUIRE(smem <= 200 * 1024, "layernorm_backward: embed_dim %d too large", E);
  if (!vec) NLPS_LN_BWD(0)
  else switch (E / 128) {
    case 1: NLPS_LN_BWD(1) break;
    case 2: NLPS_LN_BWD(2) break;
    case 3: NLPS_LN_BWD(3) break;
    case 4: NLPS_LN_BWD(4) break;
    case 5: NLPS_LN_BWD(5) break;
    case 6: NLPS_LN_BWD(6) break;
    case 7: NLPS_LN_BWD(7) break;
    default: NLPS_LN_BWD(8) break;
  }
#undef NLPS_LN_BWD
  NLPS_LAUNCH_CHECK();
  if (defer_finish) return 0;
  ln_bwd_finish_kernel<<<(unsigned)ceil_div(nsum * E, 32), 256, 0, st>>>(grid, E, nsum, partial, dgamma, dbeta, dbias);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int softmax_xent(cudaStream_t st, const float* logits, int64_t ld, const int32_t* y, int64_t rows, int V,
                 float grad_scale, float* loss, float* dlogits, int64_t ldd, float* row_loss, int32_t* err) {
  NLPS_REQUIRE(rows > 0 && V > 0, "softmax_xent: empty input");
  const int vec_ok = (V % 4 == 0 && ld % 4 == 0 && ldd % 4 == 0 && aligned16(logits) && aligned16(dlogits)) ? 1 : 0;
  const float inv = (float)((double)grad_scale / (double)rows);
  xent_kernel<<<(unsigned)rows, 256, 0, st>>>(logits, ld, y, V, inv, dlogits, ldd, row_loss, vec_ok, err);
  NLPS_LAUNCH_CHECK();
  sum_to_scalar_kernel<<<1, 1024, 0, st>>>(row_loss, rows, (float)(1.0 / (double)rows), loss);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int64_t xent_colsum_partial_rows(int64_t rows) { return ceil_div(rows, (int64_t)kXentRows); }
bool xent_colsum_fusable(const float* logits, int64_t ld, const float* dlogits, int64_t ldd, int V) {
  return V % 4 == 0 && V <= 1024 * kXentMaxChunks && ld % 4 == 0 && ldd % 4 == 0 && aligned16(logits) && aligned16(dlogits);
}
// softmax_xent that also leaves the (ceil(rows/16), V) column-sum partials of dlogits in `partial` (row stride ld_partial)
int softmax_xent_colsum(cudaStream_t st, const float* logits, int64_t ld, const int32_t* y, int64_t rows, int V,
                        float grad_scale, float* loss, float* dlogits, int64_t ldd, float* row_loss, float* partial,
                        int64_t ld_partial, int32_t* err) {
  NLPS_REQUIRE(rows > 0 && V > 0, "softmax_xent: empty input");
  NLPS_REQUIRE(xent_colsum_fusable(logits, ld, dlogits, ldd, V) && ld_partial % 4 == 0 && aligned16(partial),
               "softmax_xent_colsum: operands not fusable");
  const float inv = (float)((double)grad_scale / (double)rows);
  const unsigned grid = (unsigned)xent_colsum_partial_rows(rows);
  const int chunks = (int)ceil_div(V, 1024);
#define NLPS_XC(C) xent_colsum_kernel<C><<<grid, 256, 0, st>>>(logits, ld, y, V, rows, inv, dlogits, ldd, row_loss, partial, ld_partial, err)
  if (chunks <= 1) NLPS_XC(1); else if (chunks <= 2) NLPS_XC(2); else if (chunks <= 4) NLPS_XC(4);
  else if (chunks <= 8) NLPS_XC(8); else NLPS_XC(16);
#undef NLPS_XC
  NLPS_LAUNCH_CHECK();
  sum_to_scalar_kernel<<<1, 1024, 0, st>>>(row_loss, rows, (float)(1.0 / (double)rows), loss);
  NLPS_LAUNCH_CHECK();
  return 0;
}

// the two halves of softmax_xent for a row-chunked pipeline: dlogits + per-row losses of `rows` rows with an explicit
// factor (grad_scale / total rows), then the mean over all rows -- same kernels, same bits as the one-call form
int softmax_xent_rows(cudaStream_t st, const float* logits, int64_t ld, const int32_t* y, int64_t rows, int V, float inv,
                      float* dlogits, int64_t ldd, float* row_loss, int32_t* err) {
  if (rows <= 0) return 0;
  const int vec_ok = (V % 4 == 0 && ld % 4 == 0 && ldd % 4 == 0 && aligned16(logits) && aligned16(dlogits)) ? 1 : 0;
  xent_kernel<<<(unsigned)rows, 256, 0, st>>>(logits, ld, y, V, inv, dlogits, ldd, row_loss, vec_ok, err);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int row_loss_mean(cudaStream_t st, const float* row_loss, int64_t rows, float* loss) {
  sum_to_scalar_kernel<<<1, 1024, 0, st>>>(row_loss, rows, (float)(1.0 / (double)rows), loss);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int softmax_xent_stats(cudaStream_t st, const float* logits, int64_t ld, const int32_t* y, int64_t rows, int V,
                       float grad_scale, float* loss, float* dlogits, int64_t ldd, float* row_loss, const float2* rowstat, int32_t* err) {
  const bool vec_ok = V % 4 == 0 && ld % 4 == 0 && ldd % 4 == 0 && aligned16(logits) && aligned16(dlogits);
  if (rowstat == nullptr || !vec_ok) return softmax_xent(st, logits, ld, y, rows, V, grad_scale, loss, dlogits, ldd, row_loss, err);
  NLPS_REQUIRE(rows > 0 && V > 0, "softmax_xent: empty input");
  const float inv = (float)((double)grad_scale / (double)rows);
  xent_stats_kernel<<<(unsigned)rows, 256, 0, st>>>(logits, ld, y, V, inv, dlogits, ldd, row_loss, rowstat, err);
  NLPS_LAUNCH_CHECK();
  sum_to_scalar_kernel<<<1, 1024, 0, st>>>(row_loss, rows, (float)(1.0 / (double)rows), loss);
  NLPS_LAUNCH_CHECK();
  return 0;
}

namespace {
// ---- device-side evaluation bookkeeping of the driver loop (utils/train.py:51-75, 86-88, 143-163) ----
// accum[0] += sum of the batch's per-row losses, accum[1] += rows: one CTA, fixed tree => deterministic
__global__ void __launch_bounds__(1024) eval_accumulate_kernel(const float* __restrict__ row_loss, int64_t rows,
                                                               float* __restrict__ accum) {
  __shared__ float sm[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < rows; i += 1024) acc += row_loss[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = warp_sum(sm[threadIdx.x]);
    if (threadIdx.x == 0) { accum[0] += v; accum[1] += (float)rows; }
  }
}
// sched = {previous eval loss, evaluations so far, learning rate, halved-this-time flag, this eval loss}: the mean loss of
// the pass, and the reference's schedule "halve the learning rate when the loss went up" (utils/train.py:86-88)
__global__ void eval_finish_kernel(float* __restrict__ accum, float* __restrict__ sched) {
  const float eval = accum[0] / fmaxf(accum[1], 1.f);
  const bool worse = sched[1] > 0.f && eval > sched[0];
  if (worse) sched[2] *= 0.5f;
  sched[0] = eval; sched[1] += 1.f; sched[3] = worse ? 1.f : 0.f; sched[4] = eval;
  accum[0] = 0.f; accum[1] = 0.f;
}
// softmax probabilities of one logits row and its k most probable ids (descending; ties: the higher id first, like
// argsort()[-k:][::-1]).  One CTA; `work` is a V-float scratch that ends up holding the probabilities with the winners removed.
__global__ void __launch_bounds__(256) topk_softmax_kernel(const float* __restrict__ logits, int V, int k, float* __restrict__ work,
                                                           int32_t* __restrict__ out_idx, float* __restrict__ out_prob) {
  __shared__ float red[8];
  __shared__ int redi[8];
  __shared__ float bc[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -INFINITY;
  for (int c = tid; c < V; c += 256) m = fmaxf(m, logits[c]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) { float t = red[0]; for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]); bc[0] = t; }
  __syncthreads();
  m = bc[0];
  float s = 0.f;
  for (int c = tid; c < V; c += 256) { const float e = expf(logits[c] - m); work[c] = e; s += e; }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; bc[1] = t; }
  __syncthreads();
  const float inv = 1.f / bc[1];
  for (int j = 0; j < k; ++j) {
    float best = -1.f; int bi = -1;
    for (int c = tid; c < V; c += 256) { const float v = work[c]; if (v > best || (v == best && c > bi)) { best = v; bi = c; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi > bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { red[warp] = best; redi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float b = red[0]; int i = redi[0];
      for (int w = 1; w < 8; ++w) if (red[w] > b || (red[w] == b && redi[w] > i)) { b = red[w]; i = redi[w]; }
      out_idx[j] = i; out_prob[j] = b * inv;
      if (i >= 0) work[i] = -1.f;
    }
    __syncthreads();
  }
}

__global__ void dropout_mask_kernel(DropoutParams d, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = dropout_one(d, (uint64_t)i, 1.f) != 0.f ? 1.f : 0.f;
}
}  // namespace

int eval_accumulate(cudaStream_t st, const float* row_loss, int64_t rows, float* accum) {
  if (rows <= 0) return 0;
  eval_accumulate_kernel<<<1, 1024, 0, st>>>(row_loss, rows, accum);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int eval_finish(cudaStream_t st, float* accum, float* sched) {
  eval_finish_kernel<<<1, 1, 0, st>>>(accum, sched);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int topk_softmax(cudaStream_t st, const float* logits, int V, int k, float* work, int32_t* out_idx, float* out_prob) {
  NLPS_REQUIRE(V > 0 && k > 0 && k <= V, "topk_softmax: need 0 < k <= V (k=%d, V=%d)", k, V);
  topk_softmax_kernel<<<1, 256, 0, st>>>(logits, V, k, work, out_idx, out_prob);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int dropout_mask(cudaStream_t st, const DropoutParams& d, int64_t n, float* out) {
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), kNumSMs * 8));
  dropout_mask_kernel<<<grid, 256, 0, st>>>(d, n, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

size_t colsum_scratch_floats(int cols) { return (size_t)kColSlabsMax * cols; }

int colsum(cudaStream_t st, int64_t rows, int cols, const float* x, int64_t ld, float* out, float* scratch,
           unsigned int* tickets) {
  if (cols <= 0) return 0;
  const int gx = (int)ceil_div(cols, 128);
  // enough row slabs to put ~4 CTAs on every SM, at least 32 rows each
  int slabs = (int)std::min<int64_t>(kColSlabsMax, std::max<int64_t>(1, (kNumSMs * 4 + gx - 1) / gx));
  slabs = (int)std::max<int64_t>(1, std::min<int64_t>(slabs, ceil_div(rows, 32)));
  const int vec = (ld % 4 == 0 && cols % 4 == 0 && aligned16(x)) ? 1 : 0;
  dim3 grid((unsigned)gx, (unsigned)slabs);
  if (tickets != nullptr) {
    colsum_partial_kernel<<<grid, 256, 0, st>>>(rows, cols, x, ld, scratch, vec, tickets, out);
    NLPS_LAUNCH_CHECK();
    return 0;
  }
  colsum_partial_kernel<<<grid, 256, 0, st>>>(rows, cols, x, ld, scratch, vec, nullptr, nullptr);
  NLPS_LAUNCH_CHECK();
  colsum_finish_kernel<<<(unsigned)ceil_div(cols, 32), 256, 0, st>>>(slabs, cols, scratch, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

size_t sumsq_scratch_floats() { return kSumsqBlocks; }

int sumsq(cudaStream_t st, int64_t n, const float* x, float* out, float* scratch) {
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(kSumsqBlocks, ceil_div(n, 1024)));
  if (aligned16(x)) sumsq_partial_kernel<<<blocks, 256, 0, st>>>(n, x, scratch);
  else sumsq_scalar_partial_kernel<<<blocks, 256, 0, st>>>(n, x, scratch);
  NLPS_LAUNCH_CHECK();
  sum_to_scalar_kernel<<<1, 1024, 0, st>>>(scratch, blocks, 1.f, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace nlps
