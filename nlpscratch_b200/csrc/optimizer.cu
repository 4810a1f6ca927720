// Global-norm clip (utils/function.py:27-38) and the Adam / SGD update (transformer.py:461-509)
// over the flat parameter / gradient / moment buffers: one vectorised pass, 28 B per parameter.
// The arithmetic follows the reference's operation order in FP32 (no FMA contraction) so the
// updated weights match NumPy's to the last bit or two.
#include "common.cuh"
#include <algorithm>

namespace nlps {
namespace {

__global__ void clip_scale_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ scale,
                                  float* __restrict__ norm_out) {
  const float norm = sqrtf(sumsq[0]);
  // if total_norm > max_norm: g *= max_norm / (total_norm + 1e-6)
  scale[0] = (max_norm > 0.f && norm > max_norm) ? __fdiv_rn(max_norm, __fadd_rn(norm, 1e-6f)) : 1.f;
  if (norm_out) norm_out[0] = norm;
}

__global__ void scale_kernel(int64_t n, float* __restrict__ x, const float* __restrict__ scale) {
  const float s = scale[0];
  if (s == 1.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = __fmul_rn(x[i], s);
}

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float lr, float b1, float b2,
                                          float omb1, float omb2, float eps, float bc1, float bc2) {
  // m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*(g*g)
  m = __fadd_rn(__fmul_rn(b1, m), __fmul_rn(omb1, g));
  v = __fadd_rn(__fmul_rn(b2, v), __fmul_rn(omb2, __fmul_rn(g, g)));
  const float mhat = __fdiv_rn(m, bc1);
  const float vhat = __fdiv_rn(v, bc2);
  // p -= lr * mhat / (sqrt(vhat) + eps)
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(lr, mhat), __fadd_rn(__fsqrt_rn(vhat), eps)));
  return p;
}

__global__ void __launch_bounds__(256) adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float lr, float b1,
                                                   float b2, float omb1, float omb2, float eps, float bc1, float bc2,
                                                   const float* __restrict__ clip_scale, const float* __restrict__ hyper) {
  if (hyper != nullptr) { lr = hyper[0]; bc1 = hyper[1]; bc2 = hyper[2]; }   // graph replay: the step's values live on the device
  const float cs = clip_scale ? clip_scale[0] : 1.f;
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    if (cs != 1.f) { gg.x = __fmul_rn(gg.x, cs); gg.y = __fmul_rn(gg.y, cs); gg.z = __fmul_rn(gg.z, cs); gg.w = __fmul_rn(gg.w, cs); }
    adam_one(pp.x, gg.x, mm.x, vv.x, lr, b1, b2, omb1, omb2, eps, bc1, bc2);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr, b1, b2, omb1, omb2, eps, bc1, bc2);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr, b1, b2, omb1, omb2, eps, bc1, bc2);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr, b1, b2, omb1, omb2, eps, bc1, bc2);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float gg = g[i];
      if (cs != 1.f) gg = __fmul_rn(gg, cs);
      adam_one(p[i], gg, m[i], v[i], lr, b1, b2, omb1, omb2, eps, bc1, bc2);
    }
}

__global__ void __launch_bounds__(256) sgd_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g,
                                                  float lr, const float* __restrict__ clip_scale, const float* __restrict__ hyper) {
  if (hyper != nullptr) lr = hyper[0];
  const float cs = clip_scale ? clip_scale[0] : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gg = g[i];
    if (cs != 1.f) gg = __fmul_rn(gg, cs);
    p[i] = __fsub_rn(p[i], __fmul_rn(lr, gg));     // p -= lr * g   (transformer.py:494-509)
  }
}

}  // namespace

int clip_scale_from_sumsq(cudaStream_t st, const float* sumsq_dev, float max_norm, float* scale_out, float* norm_out) {
  clip_scale_kernel<<<1, 1, 0, st>>>(sumsq_dev, max_norm, scale_out, norm_out);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int scale_inplace(cudaStream_t st, int64_t n, float* x, const float* scale_dev) {
  if (n <= 0) return 0;
  scale_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), kNumSMs * 8), 256, 0, st>>>(n, x, scale_dev);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int adam_update(cudaStream_t st, int64_t n, float* p, const float* g, float* m, float* v, float lr, float b1,
                float b2, float one_minus_b1, float one_minus_b2, float eps, float bc1, float bc2,
                const float* clip_scale_dev, const float* hyper_dev) {
  if (n <= 0) return 0;
  adam_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 1024), kNumSMs * 8), 256, 0, st>>>(
      n, p, g, m, v, lr, b1, b2, one_minus_b1, one_minus_b2, eps, bc1, bc2, clip_scale_dev, hyper_dev);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int sgd_update(cudaStream_t st, int64_t n, float* p, const float* g, float lr, const float* clip_scale_dev, const float* hyper_dev) {
  if (n <= 0) return 0;
  sgd_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), kNumSMs * 8), 256, 0, st>>>(n, p, g, lr, clip_scale_dev, hyper_dev);
  NLPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace nlps
