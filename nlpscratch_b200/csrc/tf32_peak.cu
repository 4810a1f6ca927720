// Measured ceiling of the tensor pipe for the arithmetic this library uses: tcgen05.mma kind::tf32 with FP32 operands
// resident in shared memory and FP32 accumulators in TMEM -- the denominator of every "fraction of tensor peak" the
// benchmark reports (BASELINE.md section 3 promised a measured value instead of "half of BF16").
//
// The probe is the GEMM kernel's MMA-issue loop with everything else removed: operands are written to shared memory once
// (K-major, 128-byte swizzle atoms, pseudo-random values so that the datapath toggles like real data), one elected thread
// per CTA (or per CTA pair) issues `iters` groups of 16 MMAs (4 k-blocks of 32 floats: M x N x 128 per group) into two
// alternating TMEM accumulators, with at most two groups in flight (commit -> mbarrier), exactly like the ring of the
// real kernel.  No TMA, no epilogue, no global traffic: what is measured is MMA issue + the shared-memory operand reads.
//   cta_group 1: M=128, N=256 per instruction, one CTA per SM       (gemm_tf32_kernel's tile)
//   cta_group 2: M=256, N=256 per instruction over a CTA pair      (gemm_tf32_pair_kernel's tile; each SM reads half of B)
#include "tc_common.cuh"

namespace nlps {
namespace {
using namespace tc;

constexpr int kKBlocks = 4;                  // resident k-blocks of 32 floats per operand (192 KB for the M128 N256 tile)
constexpr int kABytes = 128 * 128;           // 128 rows x 128 B per k-block
constexpr int kBBytesSingle = 256 * 128;
constexpr int kBBytesPair = 128 * 128;

__device__ __forceinline__ float hash_unit(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return (float)(int32_t)x * (1.0f / 2147483648.0f);
}

template <int GROUP>
__global__ void __launch_bounds__(128) tf32_mma_peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* aligned = smem_raw + (base - raw);
  constexpr int kBBytes = GROUP == 2 ? kBBytesPair : kBBytesSingle;
  constexpr int kOperandBytes = kKBlocks * (kABytes + kBBytes);
  const uint32_t bar0 = base + kOperandBytes, bar1 = bar0 + 8, tmem_slot = bar0 + 16;
  const uint32_t rank = GROUP == 2 ? cluster_ctarank() : 0u;
  // operands: any bit pattern is a valid swizzled tile; pseudo-random values in (-1, 1), rounded to the TF32 grid
  for (int i = threadIdx.x; i < kOperandBytes / 4; i += blockDim.x) {
    const float v = hash_unit((uint32_t)i * 2654435761u + blockIdx.x * 40503u + 17u);
    reinterpret_cast<uint32_t*>(aligned)[i] = __float_as_uint(v) & 0xFFFFE000u;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  if (threadIdx.x < 32) { if (GROUP == 2) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  tc_fence_before();
  if (GROUP == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(aligned + kOperandBytes + 16);
  if (threadIdx.x == 0 && rank == 0) {
    constexpr uint32_t M = GROUP == 2 ? 256u : 128u, N = 256u;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      const uint32_t bar = (it & 1) ? bar1 : bar0;
      if (it >= 2) { mbar_wait(bar, (uint32_t)((it >> 1) - 1) & 1u); tc_fence_after(); }
      const uint32_t tmem_d = tmem_base + (uint32_t)((it & 1) * 256);
#pragma unroll
      for (int kb = 0; kb < kKBlocks; ++kb) {
        const uint32_t a_src = base + kb * kABytes;
        const uint32_t b_src = base + kKBlocks * kABytes + kb * kBBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_desc(a_src + k * 32, 16, 1024, 2), bd = make_desc(b_src + k * 32, 16, 1024, 2);
          if (GROUP == 2) umma_tf32_pair(tmem_d, ad, bd, idesc, (kb | k) ? 1u : 0u);
          else umma_tf32(tmem_d, ad, bd, idesc, (kb | k) ? 1u : 0u);
        }
      }
      if (GROUP == 2) {      // arrive on the leader's barrier only
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)1) : "memory");
      } else {
        umma_commit(bar);
      }
    }
    for (int it = max(0, iters - 2); it < iters; ++it) mbar_wait((it & 1) ? bar1 : bar0, (uint32_t)(it >> 1) & 1u);
    tc_fence_after();
  }
  tc_fence_before();
  if (GROUP == 2) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    if (GROUP == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <int GROUP>
int run_probe(int iters, int reps, float* tflops, float* ms_out) {
  constexpr int kBBytes = GROUP == 2 ? kBBytesPair : kBBytesSingle;
  constexpr int smem = kKBlocks * (kABytes + kBBytes) + 64 + 1024;
  auto kern = tf32_mma_peak_kernel<GROUP>;
  NLPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kNumSMs, 1, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = nullptr;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = GROUP; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  NLPS_CUDA(cudaEventCreate(&e0));
  NLPS_CUDA(cudaEventCreate(&e1));
  NLPS_CUDA(cudaLaunchKernelEx(&cfg, kern, iters));          // warm-up
  NLPS_CUDA(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    NLPS_CUDA(cudaEventRecord(e0, nullptr));
    NLPS_CUDA(cudaLaunchKernelEx(&cfg, kern, iters));
    NLPS_CUDA(cudaEventRecord(e1, nullptr));
    NLPS_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    NLPS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  count_launch(reps + 1);
  // per group and issuing unit: M x 256 x (kKBlocks * 32) MACs; cta_group::1 has one unit per SM, ::2 one per SM pair with M = 256
  const double units = GROUP == 2 ? kNumSMs / 2 : kNumSMs;
  const double flop = units * (double)iters * 2.0 * (GROUP == 2 ? 256.0 : 128.0) * 256.0 * (kKBlocks * 32.0);
  if (tflops) *tflops = (float)(flop / (best * 1e-3) / 1e12);
  if (ms_out) *ms_out = best;
  return 0;
}
}  // namespace

int tf32_mma_peak(int cta_group, int iters, int reps, float* tflops, float* ms) {
  NLPS_REQUIRE(cta_group == 1 || cta_group == 2, "tf32_mma_peak: cta_group must be 1 or 2");
  NLPS_REQUIRE(iters >= 2 && reps >= 1, "tf32_mma_peak: iters >= 2, reps >= 1");
  return cta_group == 2 ? run_probe<2>(iters, reps, tflops, ms) : run_probe<1>(iters, reps, tflops, ms);
}
}  // namespace nlps
