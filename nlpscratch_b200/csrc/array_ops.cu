// Element-wise / indexing kernels behind the host-side array shim (the `model.np` surface that
// utils/train.py:46-116 drives: slicing copies, fancy row gathers, != PAD counts, scalar math).
// Small, generic, bandwidth-bound grid-stride loops.
#include "common.cuh"
#include <algorithm>

namespace nlps {
namespace {

struct Strided4 {
  int64_t shape[4], ss[4], ds[4];
};

template <typename T>
__global__ void strided_copy_kernel(Strided4 a, int64_t total, const T* __restrict__ src, T* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i, so = 0, d0 = 0;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
      const int64_t c = r % a.shape[k];
      r /= a.shape[k];
      so += c * a.ss[k];
      d0 += c * a.ds[k];
    }
    dst[d0] = src[so];
  }
}

template <typename T>
__global__ void gather_rows_kernel(int64_t n_out, int64_t row_elems, const T* __restrict__ src, int64_t src_stride,
                                   const int32_t* __restrict__ index, int64_t n_src, T* __restrict__ dst) {
  const int64_t total = n_out * row_elems;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row_elems, c = i - r * row_elems;
    int64_t s = index[r];
    if (s < 0) s += n_src;
    s = s < 0 ? 0 : (s >= n_src ? n_src - 1 : s);
    dst[i] = src[s * src_stride + c];
  }
}

__global__ void gather_last_kernel(int B, int S, int E, const float* __restrict__ h, const int32_t* __restrict__ t,
                                   float* __restrict__ out) {
  const int64_t total = (int64_t)B * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / E), c = (int)(i - (int64_t)b * E);
    int tb = t[b];
    if (tb < 0) tb += S;
    tb = tb < 0 ? 0 : (tb >= S ? S - 1 : tb);
    out[i] = h[((int64_t)b * S + tb) * E + c];
  }
}

__global__ void count_nonzero_rows_kernel(int64_t rows, int64_t cols, const int32_t* __restrict__ x, int64_t stride,
                                          int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int cnt = 0;
  for (int64_t c = lane; c < cols; c += 32) cnt += x[row * stride + c] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) out[row] = cnt;
}

template <typename T>
__device__ __forceinline__ T apply_bin(int op, T a, T b) {
  switch (op) {
    case 0: return a + b;
    case 1: return a - b;
    case 2: return a * b;
    case 3: return a / b;
    default: return a > b ? a : b;
  }
}

template <typename T>
__global__ void binary_kernel(int op, int64_t n, int64_t cols, const T* __restrict__ a, const T* __restrict__ b,
                              T bs, int mode, T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const T bv = mode == 0 ? bs : (mode == 1 ? b[i] : b[i % cols]);
    out[i] = apply_bin<T>(op, a[i], bv);
  }
}

__global__ void unary_kernel(int op, int64_t n, const float* __restrict__ a, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = a[i];
    float o;
    switch (op) {
      case 0: o = expf(v); break;
      case 1: o = sqrtf(v); break;
      case 2: o = v * v; break;
      case 3: o = logf(v); break;
      default: o = -v; break;
    }
    out[i] = o;
  }
}
// out[r] = src[i0[r], i1[r], :] for a dense (n0, n1, row_elems) source: NumPy's H[idx, t, :] (utils/train.py:113-115),
// each index wrapped on its own axis like NumPy does (t = lens - 1 is -1 for an all-PAD row and selects the row's own
// last position), then clamped into range
template <typename T>
__global__ void gather_rows2_kernel(int64_t n_out, int64_t row_elems, const T* __restrict__ src, int64_t n0, int64_t n1,
                                    const int32_t* __restrict__ i0, const int32_t* __restrict__ i1, T* __restrict__ dst) {
  const int64_t total = n_out * row_elems;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row_elems, c = i - r * row_elems;
    int64_t a = i0[r], b = i1[r];
    if (a < 0) a += n0;
    if (b < 0) b += n1;
    a = a < 0 ? 0 : (a >= n0 ? n0 - 1 : a);
    b = b < 0 ? 0 : (b >= n1 ? n1 - 1 : b);
    dst[i] = src[(a * n1 + b) * row_elems + c];
  }
}

// single-CTA reductions: the shim only reduces small arrays (lens, logits of one row, grads go
// through sumsq()); fixed order => deterministic
template <typename T, int OP>
__global__ void __launch_bounds__(1024) reduce_kernel(int64_t n, const T* __restrict__ a, T* __restrict__ out) {
  __shared__ T sm[32];
  T acc = OP == 1 ? a[0] : (T)0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const T v = a[i];
    if (OP == 0) acc += v;
    else if (OP == 1) acc = v > acc ? v : acc;
    else acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T t = __shfl_xor_sync(0xffffffffu, acc, o);
    acc = OP == 1 ? (t > acc ? t : acc) : acc + t;
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    T v = sm[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T t = __shfl_xor_sync(0xffffffffu, v, o);
      v = OP == 1 ? (t > v ? t : v) : v + t;
    }
    if (threadIdx.x == 0) out[0] = v;
  }
}

__global__ void compare_ne_kernel(int64_t n, const int32_t* __restrict__ a, int32_t value, int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a[i] != value;
}
__global__ void cast_i32_f32_kernel(int64_t n, const int32_t* __restrict__ a, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (float)a[i];
}

inline unsigned grid_for(int64_t n) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), kNumSMs * 8)); }

}  // namespace

int strided_copy(cudaStream_t st, int elem_size, int ndim, const int64_t* shape, const void* src,
                 const int64_t* sstr, void* dst, const int64_t* dstr) {
  NLPS_REQUIRE(elem_size == 4, "strided_copy: only 4-byte elements are supported");
  NLPS_REQUIRE(ndim >= 0 && ndim <= 4, "strided_copy: ndim %d > 4", ndim);
  Strided4 a;
  int64_t total = 1;
  for (int k = 0; k < 4; ++k) {
    const int j = k - (4 - ndim);
    a.shape[k] = j >= 0 ? shape[j] : 1;
    a.ss[k] = j >= 0 ? sstr[j] : 0;
    a.ds[k] = j >= 0 ? dstr[j] : 0;
    total *= a.shape[k];
  }
  if (total <= 0) return 0;
  strided_copy_kernel<uint32_t><<<grid_for(total), 256, 0, st>>>(a, total, static_cast<const uint32_t*>(src),
                                                                 static_cast<uint32_t*>(dst));
  NLPS_LAUNCH_CHECK();
  return 0;
}

int gather_rows(cudaStream_t st, int elem_size, int64_t n_out, int64_t row_elems, const void* src, int64_t src_stride,
                const int32_t* index, int64_t n_src, void* dst) {
  NLPS_REQUIRE(elem_size == 4, "gather_rows: only 4-byte elements are supported");
  if (n_out * row_elems <= 0) return 0;
  gather_rows_kernel<uint32_t><<<grid_for(n_out * row_elems), 256, 0, st>>>(
      n_out, row_elems, static_cast<const uint32_t*>(src), src_stride, index, n_src, static_cast<uint32_t*>(dst));
  NLPS_LAUNCH_CHECK();
  return 0;
}

int gather_rows2(cudaStream_t st, int elem_size, int64_t n_out, int64_t row_elems, const void* src, int64_t n0, int64_t n1,
                 const int32_t* i0, const int32_t* i1, void* dst) {
  NLPS_REQUIRE(elem_size == 4, "gather_rows2: only 4-byte elements are supported");
  if (n_out * row_elems <= 0) return 0;
  NLPS_REQUIRE(n0 > 0 && n1 > 0, "gather_rows2: empty source");
  gather_rows2_kernel<uint32_t><<<grid_for(n_out * row_elems), 256, 0, st>>>(
      n_out, row_elems, static_cast<const uint32_t*>(src), n0, n1, i0, i1, static_cast<uint32_t*>(dst));
  NLPS_LAUNCH_CHECK();
  return 0;
}

int gather_last(cudaStream_t st, int B, int S, int E, const float* h, const int32_t* t, float* out) {
  if ((int64_t)B * E <= 0) return 0;
  gather_last_kernel<<<grid_for((int64_t)B * E), 256, 0, st>>>(B, S, E, h, t, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int count_nonzero_rows(cudaStream_t st, int64_t rows, int64_t cols, const int32_t* x, int64_t stride, int32_t* out) {
  if (rows <= 0) return 0;
  count_nonzero_rows_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(rows, cols, x, stride, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int binary_f32(cudaStream_t st, int op, int64_t n, int64_t cols, const float* a, const float* b, float bs, int mode,
               float* out) {
  if (n <= 0) return 0;
  binary_kernel<float><<<grid_for(n), 256, 0, st>>>(op, n, cols > 0 ? cols : 1, a, b, bs, mode, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int binary_i32(cudaStream_t st, int op, int64_t n, const int32_t* a, const int32_t* b, int32_t bs, int mode,
               int32_t* out) {
  if (n <= 0) return 0;
  binary_kernel<int32_t><<<grid_for(n), 256, 0, st>>>(op, n, 1, a, b, bs, mode, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int unary_f32(cudaStream_t st, int op, int64_t n, const float* a, float* out) {
  if (n <= 0) return 0;
  unary_kernel<<<grid_for(n), 256, 0, st>>>(op, n, a, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int reduce_f32(cudaStream_t st, int op, int64_t n, const float* a, float* out) {
  NLPS_REQUIRE(n > 0, "reduce: empty array");
  if (op == 0) reduce_kernel<float, 0><<<1, 1024, 0, st>>>(n, a, out);
  else if (op == 1) reduce_kernel<float, 1><<<1, 1024, 0, st>>>(n, a, out);
  else reduce_kernel<float, 2><<<1, 1024, 0, st>>>(n, a, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int reduce_i32(cudaStream_t st, int op, int64_t n, const int32_t* a, int32_t* out) {
  NLPS_REQUIRE(n > 0, "reduce: empty array");
  if (op == 0) reduce_kernel<int32_t, 0><<<1, 1024, 0, st>>>(n, a, out);
  else if (op == 1) reduce_kernel<int32_t, 1><<<1, 1024, 0, st>>>(n, a, out);
  else reduce_kernel<int32_t, 2><<<1, 1024, 0, st>>>(n, a, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int compare_ne_i32(cudaStream_t st, int64_t n, const int32_t* a, int32_t value, int32_t* out) {
  if (n <= 0) return 0;
  compare_ne_kernel<<<grid_for(n), 256, 0, st>>>(n, a, value, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}
int cast_i32_f32(cudaStream_t st, int64_t n, const int32_t* a, float* out) {
  if (n <= 0) return 0;
  cast_i32_f32_kernel<<<grid_for(n), 256, 0, st>>>(n, a, out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace nlps
