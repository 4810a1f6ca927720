// CUDA-core FP32 GEMM: the "FP32-faithful" arithmetic mode (NLPS_PREC_FP32) and the path for
// shapes the TMA descriptors cannot express (unaligned leading dimensions).  Plain FMA in
// FP32, sequential K order per output element.  Same fused epilogue as the tcgen05 kernel.
#include "common.cuh"

namespace nlps {
namespace {

constexpr int TM = 64, TN = 64, TK = 16, PADS = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[TK][TM + PADS];
  __shared__ __align__(16) float Bs[TK][TN + PADS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (TA) { m = idx & (TM - 1); k = idx >> 6; } else { k = idx & (TK - 1); m = idx >> 4; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < g.K) v = TA ? g.A[(int64_t)gk * g.lda + gm] : g.A[(int64_t)gm * g.lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int n, k;
      if (TB) { k = idx & (TK - 1); n = idx >> 4; } else { n = idx & (TN - 1); k = idx >> 6; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < g.K) v = TB ? g.B[(int64_t)gn * g.ldb + gk] : g.B[(int64_t)gk * g.ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= g.N) continue;
      float v = gemm_epilogue_one(g, acc[i][j], row, col);
      float* dst = g.C + (int64_t)row * g.ldc + col;
      if (g.accumulate) v += *dst;
      *dst = v;
    }
  }
}

}  // namespace

int gemm_fp32(cudaStream_t st, const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0) return 0;
  NLPS_REQUIRE(!g.relu_bits_out && !g.gate_bits, "gemm_fp32: bit-packed ReLU masks are a tensor-core-path feature");
  dim3 grid((unsigned)ceil_div(g.M, TM), (unsigned)ceil_div(g.N, TN));
  NLPS_REQUIRE(grid.y <= 65535u, "gemm_fp32: N=%d too large for this grid", g.N);
  if (g.trans_a && g.trans_b) gemm_simt_kernel<true, true><<<grid, 256, 0, st>>>(g);
  else if (g.trans_a) gemm_simt_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (g.trans_b) gemm_simt_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else gemm_simt_kernel<false, false><<<grid, 256, 0, st>>>(g);
  NLPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace nlps
