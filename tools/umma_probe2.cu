// Probe 2: can a K-major tf32 operand be read from a SWIZZLE_128B_ATOM_32B image (layout type 1), so that one TMA
// delivery serves both the K-major and the MN-major use of a tile?
// (derived from umma_probe.cu) which (LBO, SBO, k-step) encoding makes tcgen05.mma kind::tf32 read an MN-major SWIZZLE_128B
// operand that TMA wrote as 32-float-wide column slabs?  Prints the max abs error per candidate.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t lt = 2) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (lt << 61);
}

struct Cand { int a_mn, b_mn; uint32_t lbo, sbo, kstep; int lt; int a_k32, b_k32; int nk; uint32_t kadv; };

// A: MN-major source is [K=32][128] floats; K-major source is [128][K=32].  Same for B (N=128).
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map_a_mn, const __grid_constant__ CUtensorMap map_a_k,
                                            const __grid_constant__ CUtensorMap map_b_mn, const __grid_constant__ CUtensorMap map_b_k,
                                            const __grid_constant__ CUtensorMap map_a_k32, const __grid_constant__ CUtensorMap map_b_k32,
                                            Cand c, float* out, float* smem_dump) {
  extern __shared__ uint8_t raw[];
  const uint32_t rawa = smem_u32(raw);
  const uint32_t base = (rawa + 1023u) & ~1023u;
  uint8_t* bp = raw + (base - rawa);
  const uint32_t a_s = base, b_s = base + 16384, bar = base + 32768, bar2 = bar + 8, slot = bar + 16;
  volatile uint32_t* slot_p = reinterpret_cast<volatile uint32_t*>(bp + 32768 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot_p;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 32768);
    if (c.a_mn) { for (int j = 0; j < 4; ++j) tma_load_2d(a_s + j * 4096, &map_a_mn, bar, 32 * j, 0); }
    else tma_load_2d(a_s, c.a_k32 ? &map_a_k32 : &map_a_k, bar, 0, 0);
    if (c.b_mn) { for (int j = 0; j < 4; ++j) tma_load_2d(b_s + j * 4096, &map_b_mn, bar, 32 * j, 0); }
    else tma_load_2d(b_s, c.b_k32 ? &map_b_k32 : &map_b_k, bar, 0, 0);
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < c.nk; ++k) {
      const uint64_t ad = c.a_mn ? make_desc(a_s + k * c.kstep, c.lbo, c.sbo, c.lt) : make_desc(a_s + k * (c.a_k32 ? c.kadv : 32), 16, 1024, c.a_k32 ? 1 : 2);
      const uint64_t bd = c.b_mn ? make_desc(b_s + k * c.kstep, c.lbo, c.sbo, c.lt) : make_desc(b_s + k * (c.b_k32 ? c.kadv : 32), 16, 1024, c.b_k32 ? 1 : 2);
      const uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2) : "memory");
    mbar_wait(bar2, 0);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (smem_dump) for (int i = threadIdx.x; i < 8192; i += 128) smem_dump[i] = reinterpret_cast<float*>(bp)[i];
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + cc * 32;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 128 + cc * 32 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

CUtensorMap mk(EncodeFn fn, float* p, uint64_t inner, uint64_t outer, uint32_t bi, uint32_t bo, CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer}; cuuint64_t str[1] = {inner * 4}; cuuint32_t box[2] = {bi, bo}; cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  const int M = 128, N = 128, K = 32;
  std::vector<float> a_mn(K * M), a_k(M * K), b_mn(K * N), b_k(N * K), ref(M * N);
  srand(1);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { float v = (float)((rand() % 7) - 3); a_mn[k * M + m] = v; a_k[m * K + k] = v; }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { float v = (float)((rand() % 5) - 2); b_mn[k * N + n] = v; b_k[n * K + k] = v; }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += a_k[m * K + k] * b_k[n * K + k]; ref[m * N + n] = s; }
  float *da_mn, *da_k, *db_mn, *db_k, *dout, *ddump;
  CK(cudaMalloc(&da_mn, 4 * K * M)); CK(cudaMalloc(&da_k, 4 * K * M)); CK(cudaMalloc(&db_mn, 4 * K * N)); CK(cudaMalloc(&db_k, 4 * K * N));
  CK(cudaMalloc(&dout, 4 * M * N)); CK(cudaMalloc(&ddump, 4 * 8192));
  CK(cudaMemcpy(da_mn, a_mn.data(), 4 * K * M, cudaMemcpyHostToDevice)); CK(cudaMemcpy(da_k, a_k.data(), 4 * K * M, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db_mn, b_mn.data(), 4 * K * N, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db_k, b_k.data(), 4 * K * N, cudaMemcpyHostToDevice));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn fn = (EncodeFn)fp;
  CUtensorMap ma_mn = mk(fn, da_mn, M, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), ma_k = mk(fn, da_k, K, M, 32, 128), mb_mn = mk(fn, db_mn, N, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), mb_k = mk(fn, db_k, K, N, 32, 128);
  CUtensorMap ma_k32 = mk(fn, da_k, K, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), mb_k32 = mk(fn, db_k, K, N, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
  std::vector<Cand> cands = {
      {0, 0, 0, 0, 0, 2, 0, 0, 4, 32},
      {0, 0, 0, 0, 0, 2, 1, 0, 1, 32}, {0, 0, 0, 0, 0, 2, 0, 1, 1, 32},
      {0, 0, 0, 0, 0, 2, 1, 0, 4, 32}, {0, 0, 0, 0, 0, 2, 0, 1, 4, 32}, {0, 0, 0, 0, 0, 2, 1, 1, 4, 32},
      {1, 0, 4096, 512, 1024, 1, 0, 1, 4, 32}, {0, 1, 4096, 512, 1024, 1, 1, 0, 4, 32},
  };
  std::vector<float> out(M * N), dump(8192);
  for (size_t ci = 0; ci < cands.size(); ++ci) {
    if (only >= 0 && (int)ci != only) continue;
    auto& c = cands[ci];
    CK(cudaMemset(dout, 0xff, 4 * M * N));
    probe<<<1, 128, 40000>>>(ma_mn, ma_k, mb_mn, mb_k, ma_k32, mb_k32, c, dout, ddump);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cand a_mn=%d b_mn=%d lbo=%u sbo=%u kstep=%u : CUDA error %s\n", c.a_mn, c.b_mn, c.lbo, c.sbo, c.kstep, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(out.data(), dout, 4 * M * N, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dump.data(), ddump, 4 * 8192, cudaMemcpyDeviceToHost));
    if (c.nk < 4) for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < 8 * c.nk; ++k) s += a_k[m * K + k] * b_k[n * K + k]; ref[m * N + n] = s; }
    double maxerr = 0, sumabs = 0; int nz = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs((double)out[i] - ref[i])); sumabs += fabs(out[i]); nz += out[i] != 0; }
    double dsum = 0; for (int i = 0; i < 4096; ++i) dsum += fabs(dump[i]);
    printf("cand nk=%d a_k32=%d b_k32=%d lt=%d a_mn=%d b_mn=%d lbo=%5u sbo=%5u kstep=%4u : max_err=%g  sum|out|=%g nonzero=%d  sum|smemA|=%g  out[0..3]=%g %g %g %g ref=%g %g %g %g\n", c.nk, c.a_k32, c.b_k32, c.lt, c.a_mn, c.b_mn, c.lbo, c.sbo,
           c.kstep, maxerr, sumabs, nz, dsum, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    if (c.a_mn == 1 && c.b_mn == 0 && c.lbo == 4096 && c.sbo == 512) {
      // check the TMA image of A: logical (m,k) expected at slab m/32, row k, chunk ((m%32)/4) ^ (k%8)
      int bad = 0;
      for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
        int off = (m / 32) * 1024 + k * 32 + ((((m % 32) / 8) ^ (k % 4)) * 8) + (m % 8);
        if (dump[off] != a_mn[k * M + m]) ++bad;
      }
      printf("   TMA image of MN-major A matches the assumed slab/row/swizzle layout: %s (%d mismatches)\n", bad ? "NO" : "yes", bad);
    }
  }
  return 0;
}
