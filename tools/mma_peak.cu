// Microbenchmark: peak rate of legacy mma.sync m16n8k8 TF32 (and m16n8k16 BF16) on this GPU.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void __launch_bounds__(256) k_tf32(float* out, int iters) {
  float c[8][4] = {};
  uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b0 = 0x3f800000u, b1 = 0x3f800000u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  if (s == 12345.f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_bf16(float* out, int iters) {
  float c[8][4] = {};
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  if (s == 12345.f) out[0] = s;
}
int main() {
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 4;
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_tf32<<<grid, 256>>>(d, iters); else k_bf16<<<grid, 256>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)grid * 8 * iters * 8;
    double flops = mmas * (which == 0 ? 2048.0 : 4096.0);
    printf("%s: %.3f ms  %.1f TFLOP/s  (%.2f mma/cycle/SM at 1.9 GHz)\n", which == 0 ? "mma.sync m16n8k8 tf32" : "mma.sync m16n8k16 bf16", ms, flops / ms / 1e9,
           mmas / (ms * 1e-3) / 148 / 1.9e9);
  }
  return 0;
}
