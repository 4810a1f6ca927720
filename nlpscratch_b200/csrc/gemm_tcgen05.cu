// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = op(A)[M,K] * op(B)[K,N]  with fused epilogue.
//
//   * operands are FP32 in HBM, fed by TMA (128-byte-swizzled boxes, 128 B = 32 floats wide) into a
//     4..6 stage shared-memory ring; `tcgen05.mma.cta_group::1.kind::tf32` (M=128, N=128|256, K=8)
//     accumulates FP32 in TMEM; two accumulator buffers let the epilogue of tile i overlap the
//     MMAs of tile i+1.
//   * both operand majors are native: a row-major [rows,K] operand is K-major, a row-major
//     [K,rows] operand is MN-major (instruction-descriptor transpose bits), so forward (X@W),
//     dgrad (dY@W^T) and wgrad (X^T@dY) all read the tensors where they lie -- no transposes.
//   * warp roles: warp0 = TMA producer, warp1 = TMEM owner + single-thread MMA issuer,
//     warps2-5 = epilogue (tcgen05.ld 32x32b -> swizzled smem transpose -> coalesced float4 stores
//     with bias / ReLU / ReLU-gate / Philox dropout / residual fused).
//   * persistent: grid = min(#work items, #SMs); work item = (split-K slice, m tile, n tile), n
//     fastest: the ~148 tiles in flight then cover a few m-tiles times ALL n-tiles, so each
//     activation tile is fetched from HBM once and the (small) weight matrix stays in L2.
//     Small-MN/large-K products (weight gradients) are split along K into a workspace and reduced
//     in a fixed order (deterministic); the split count is chosen to fill whole waves of 148 CTAs.
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>

namespace nlps {
namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int BK = 32;                 // floats per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 8;              // kind::tf32: 32 bytes of K per instruction
constexpr int kEpiWarps = 8;             // one warp per TMEM lane quarter (8 = two per quarter, splitting columns)
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kEpiStageFloats = 32 * 32;   // per epilogue warp: 32x32 block, 16-B chunks XOR-swizzled by row

template <int BN>
struct TileCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 4;
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kEpiBytes = kEpiWarps * kEpiStageFloats * 4;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;
};

struct KernelArgs {
  GemmArgs g;
  int num_m_tiles, num_n_tiles, splits, kblocks_total, kblocks_per_split;
  float* ws;            // split-K workspace [splits][M][ws_ld] or null
  int64_t ws_ld;
  int vec_ok;           // float4 epilogue allowed (alignment of C / residual / gate / N)
  int direct_ok;        // 32-B aligned output/fused operands: epilogue stores rows straight from registers
  int x3_kb;            // 3xTF32 by K-concatenation: k-blocks per operand segment (0 = plain product); operands are [Ah | Al]
  int x3_cb;            // and [Bh ; Bl] along K.  The virtual K axis is cut into chunks of x3_cb k-blocks; chunk c walks
                        // (Al,Bh), (Ah,Bl), (Ah,Bh) over its k-blocks -- see x3_coords
  int store_hint;       // 1: the direct epilogue stores carry L2::evict_first (outputs larger than the L2 can keep anyway)
  int debug;            // NLPS_GEMM_DEBUG bits: 1 = skip global stores, 2 = skip the whole epilogue body, 4 = .release.cluster
                        // TMEM-buffer hand-over, 8 = releasing cluster barrier at the end (the two pre-fix forms, for A/B)
};

// 3xTF32 by K-concatenation: K coordinates (in floats) of the A and B boxes of virtual k-block kb.
// The tensor core accumulates with truncation (measured: the error of a kind::tf32 chain grows linearly with its length,
// ~1e-8 relative per k-block, even when the addends are tiny), so an FP32-faithful product keeps its chains short and its
// small terms first: the virtual axis is cut into chunks of `cb` k-blocks of the original K; chunk c is one split-K item
// that accumulates Al*Bh, then Ah*Bl (while the accumulator is still ~2^-11 of its final size), then Ah*Bh over the chunk's
// k-blocks; the chunk partials are summed with round-to-nearest FP32 adds by the split-K reducer.
__device__ __forceinline__ void x3_coords(int kb, int seg_kb, int cb, int& a_k, int& b_k) {
  const int c = kb / (3 * cb), j = kb - c * 3 * cb;
  const int nb = min(cb, seg_kb - c * cb);
  const int seg = j / nb, r = c * cb + (j - seg * nb);
  a_k = ((seg == 0 ? seg_kb : 0) + r) * BK;      // Al, Ah, Ah
  b_k = ((seg == 1 ? seg_kb : 0) + r) * BK;      // Bh, Bl, Bh
}

// ---- epilogue store of one 32x32 fp32 block, coalesced ---------------------------------
// The warp re-reads the block it just transposed through smem so that one store instruction covers
// 4 rows x 128 contiguous bytes.  The epilogue warps run alone on their SM sub-partitions, so the
// code is organised for instruction-level parallelism: all 8 smem reads, then all global reads of a
// fused operand, then all 8 stores, with the feature tests hoisted out of the loops.
__device__ __forceinline__ void f4_add(float4& v, const float4& t) { v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }

__device__ __forceinline__ void store_block(const KernelArgs& ka, const float* stg, int lane, int row0, int col0,
                                            int split) {
  const GemmArgs& g = ka.g;
  const int r_in = lane >> 3, c4 = (lane & 7) * 4;
  const int col = col0 + c4;
  float4 v[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + r_in;
    v[it] = *reinterpret_cast<const float4*>(stg + r * 32 + (((lane & 7) ^ (r & 7)) << 2));
  }
  if (ka.ws != nullptr) {        // split-K partial: raw accumulators, the epilogue runs in the reducer
    if (col < g.N) {
      float* dst = ka.ws + ((int64_t)split * g.M + row0 + r_in) * ka.ws_ld + col;   // ws_ld % 4 == 0, padded
#pragma unroll
      for (int it = 0; it < 8; ++it)
        if (row0 + it * 4 + r_in < g.M) *reinterpret_cast<float4*>(dst + (int64_t)it * 4 * ka.ws_ld) = v[it];
    }
    return;
  }
  const bool interior = (row0 + 32 <= g.M) && (col0 + 32 <= g.N);
  if (ka.vec_ok && interior) {
    const int64_t row = row0 + r_in;
    if (g.gate) {
      const float* src = g.gate + row * g.ldg + col;
      float4 t[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) t[it] = *reinterpret_cast<const float4*>(src + (int64_t)it * 4 * g.ldg);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        v[it].x = t[it].x > 0.f ? v[it].x : 0.f; v[it].y = t[it].y > 0.f ? v[it].y : 0.f;
        v[it].z = t[it].z > 0.f ? v[it].z : 0.f; v[it].w = t[it].w > 0.f ? v[it].w : 0.f;
      }
    }
    if (g.residual) {
      const float* src = g.residual + row * g.ldr + col;
      float4 t[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) t[it] = *reinterpret_cast<const float4*>(src + (int64_t)it * 4 * g.ldr);
#pragma unroll
      for (int it = 0; it < 8; ++it) f4_add(v[it], t[it]);
    }
    float* dst = g.C + row * g.ldc + col;
    if (g.accumulate) {
      float4 t[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) t[it] = *reinterpret_cast<const float4*>(dst + (int64_t)it * 4 * g.ldc);
#pragma unroll
      for (int it = 0; it < 8; ++it) f4_add(v[it], t[it]);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) *reinterpret_cast<float4*>(dst + (int64_t)it * 4 * g.ldc) = v[it];
    return;
  }
  // edge blocks / unaligned outputs: element-wise with bounds checks
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int row = row0 + it * 4 + r_in;
    if (row >= g.M) continue;
    const float vv[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (col + j >= g.N) break;
      float o = vv[j];
      if (g.gate) o = g.gate[(int64_t)row * g.ldg + col + j] > 0.f ? o : 0.f;
      if (g.residual) o += g.residual[(int64_t)row * g.ldr + col + j];
      float* dst = g.C + (int64_t)row * g.ldc + col + j;
      if (g.accumulate) o += *dst;
      *dst = o;
    }
  }
}

// Phase 1 of the epilogue, on the accumulator row a thread just read from TMEM (32 consecutive
// columns in registers): bias, ReLU (+ bit-packed mask out), bit-packed ReLU gate, inverted dropout.
// Everything here needs no extra global traffic beyond a broadcast bias load and one mask word.
__device__ __forceinline__ void epilogue_row(const GemmArgs& g, uint32_t (&v)[32], int row, int col0) {
  const bool full = col0 + 32 <= g.N;
  if (g.bias) {
    if (full && ((reinterpret_cast<uintptr_t>(g.bias) & 15u) == 0)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = *reinterpret_cast<const float4*>(g.bias + col0 + 4 * i);
        v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + b.x);
        v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + b.y);
        v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + b.z);
        v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + b.w);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (col0 + c < g.N) v[c] = __float_as_uint(__uint_as_float(v[c]) + g.bias[col0 + c]);
    }
  }
  if (g.relu) {
    uint32_t bits = 0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float f = fmaxf(__uint_as_float(v[c]), 0.f);
      v[c] = __float_as_uint(f);
      bits |= (f > 0.f ? 1u : 0u) << c;
    }
    if (g.relu_bits_out && row < g.M) g.relu_bits_out[(int64_t)row * g.ld_bits + (col0 >> 5)] = bits;
  }
  if (g.gate_bits) {
    const uint32_t bits = row < g.M ? g.gate_bits[(int64_t)row * g.ld_bits + (col0 >> 5)] : 0u;
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = ((bits >> c) & 1u) ? v[c] : 0u;
  }
  if (g.drop.p > 0.f) {
    if (g.N % 8 == 0) {                       // col0 % 32 == 0, so row*N + col0 is a multiple of 8
      const uint64_t idx8 = ((uint64_t)row * (uint64_t)g.N + (uint64_t)col0) >> 3;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 r = dropout_rand8(g.drop, idx8 + k);
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          const int c = 8 * k + f;
          v[c] = dropout_keep(g.drop, r, f) ? __float_as_uint(__uint_as_float(v[c]) * g.drop.inv_keep) : 0u;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (col0 + c < g.N)
          v[c] = __float_as_uint(dropout_one(g.drop, (uint64_t)row * (uint64_t)g.N + (uint64_t)(col0 + c), __uint_as_float(v[c])));
    }
  }
}

// Column sums of a warp's 32x32 block (thread = row, v = its 32 columns): recursive halving over the lanes, 31 shuffles;
// lane j ends up with the total of column j.  Fixed order => deterministic.
__device__ __forceinline__ float block_colsum32(const uint32_t (&v)[32], bool row_live, int lane) {
  float t[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) t[i] = row_live ? __uint_as_float(v[i]) : 0.f;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? t[i] : t[i + off];
      const float keep = upper ? t[i + off] : t[i];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return t[0];
}

// Phase 2 without the shared-memory transpose: the thread owns 32 consecutive output floats of one row
// (128 B) and moves them with four 256-bit accesses (LDG/STG.E.ENL2.256).  Every access is a full 32-B
// sector, and the 2 x 128 KB per tile of staging traffic -- which competes with TMA fills and UMMA
// operand reads for the 128 B/clk of shared-memory bandwidth -- disappears.
__device__ __forceinline__ void ld_v8(const float* p, float (&r)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st_v8(float* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// The same store with an L2 eviction hint: the output tiles of a product are written once and not read again by it, so their
// dirty lines should be the first to leave the 126 MB L2, not the operands.  Same-box A/B, three interleaved runs each: c4
// 13.22 -> 13.16 ms/step, c3 1.93 -> 1.90, c2 0.915 -> 0.909, c5 equal within noise; hinting only the outputs above / below
// 96 MB measured the same as hinting all.  NLPS_GEMM_STORE_HINT=0 restores plain stores (2 / 3: only large / only small
// outputs); NLPS_GEMM_DEBUG bit 128 = evict_first + no L1 allocation, measured equal.
__device__ __forceinline__ void st_v8_ef(float* p, const uint32_t* v) {
  asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void st_v8_naef(float* p, const uint32_t* v) {
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void store_row_direct(const KernelArgs& ka, uint32_t (&v)[32], int row, int col0, int split) {
  const GemmArgs& g = ka.g;
  if (ka.ws != nullptr) {
    float* dst = ka.ws + ((int64_t)split * g.M + row) * ka.ws_ld + col0;
#pragma unroll
    for (int i = 0; i < 4; ++i) st_v8(dst + 8 * i, v + 8 * i);
    return;
  }
  if (g.gate) {
    const float* src = g.gate + (int64_t)row * g.ldg + col0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t[8];
      ld_v8(src + 8 * i, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * i + j] = t[j] > 0.f ? v[8 * i + j] : 0u;
    }
  }
  if (g.residual) {
    const float* src = g.residual + (int64_t)row * g.ldr + col0;
    float t[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_v8(src + 8 * i, t[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * i + j] = __float_as_uint(__uint_as_float(v[8 * i + j]) + t[i][j]);
  }
  float* dst = g.C + (int64_t)row * g.ldc + col0;
  if (g.accumulate) {
    float t[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_v8(dst + 8 * i, t[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * i + j] = __float_as_uint(__uint_as_float(v[8 * i + j]) + t[i][j]);
  }
  if (ka.store_hint || (ka.debug & 64)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) st_v8_ef(dst + 8 * i, v + 8 * i);
  } else if (ka.debug & 128) {
#pragma unroll
    for (int i = 0; i < 4; ++i) st_v8_naef(dst + 8 * i, v + 8 * i);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) st_v8(dst + 8 * i, v + 8 * i);
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ KernelArgs ka) {
  using Cfg = TileCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // SWIZZLE_128B atoms need 1024-B alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t epi_base = base + Cfg::kStages * Cfg::kStageBytes;
  float* epi_ptr = reinterpret_cast<float*>(base_ptr + Cfg::kStages * Cfg::kStageBytes);
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  // barrier layout (8 B each): full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::kStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::kStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::kStages * Cfg::kStageBytes + Cfg::kEpiBytes +
                                           8 * (2 * Cfg::kStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmArgs& g = ka.g;
  const int tiles = ka.num_m_tiles * ka.num_n_tiles;
  const int items = tiles * ka.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int split = item / tiles;
        const int t = item - split * tiles;
        const int m0 = (t / ka.num_n_tiles) * BM;
        const int n0 = (t % ka.num_n_tiles) * BN;
        const int kb0 = split * ka.kblocks_per_split;
        const int kb1 = min(ka.kblocks_total, kb0 + ka.kblocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = base + stage * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + Cfg::kABytes;
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          int ka0 = kb * BK, kb0k = ka0;
          if (ka.x3_kb) x3_coords(kb, ka.x3_kb, ka.x3_cb, ka0, kb0k);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(a_dst + j * 4096, &map_a, full_bar(stage), m0 + 32 * j, ka0);
          } else {
            tma_load_2d(a_dst, &map_a, full_bar(stage), ka0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_2d(b_dst + j * 4096, &map_b, full_bar(stage), n0 + 32 * j, kb0k);
          } else {
            tma_load_2d(b_dst, &map_b, full_bar(stage), kb0k, n0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // instruction descriptor: D=f32 (bits4-5=1), A=B=tf32 (2 at bits 7-9 / 10-12), majors, N>>3, M>>4
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                 ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0; int local = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++local) {
        const int split = item / tiles;
        const int kb0 = split * ka.kblocks_per_split;
        const int kb1 = min(ka.kblocks_total, kb0 + ka.kblocks_per_split);
        const int buf = local & 1;
        const uint32_t aphase = (uint32_t)(local >> 1) & 1u;
        mbar_wait(tempty_bar(buf), aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_src = base + stage * Cfg::kStageBytes;
          const uint32_t b_src = a_src + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: rows of 128 B, 8-row groups 1024 B apart (SBO), +32 B per K=8 step.
            // MN-major: 32-float column slabs of (BK x 128 B) = 4096 B (LBO), 4-row K groups 512 B
            //           apart (SBO), +1024 B per K=8 step.
            const uint64_t ad = A_MN ? make_desc(a_src + k * 1024, 4096, 512, 1) : make_desc(a_src + k * 32, 16, 1024, 2);
            const uint64_t bd = B_MN ? make_desc(b_src + k * 1024, 4096, 512, 1) : make_desc(b_src + k * 32, 16, 1024, 2);
            umma_tf32(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));        // smem slot free once these MMAs retire
          if (kb == kb1 - 1) umma_commit(tfull_bar(buf));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;            // which share of the tile's columns (0 when kEpiWarps == 4)
    constexpr int kChunksPerWarp = BN / 32 / (kEpiWarps / 4);
    float* stg = epi_ptr + (warp - 2) * kEpiStageFloats;
    int local = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++local) {
      const int split = item / tiles;
      const int t = item - split * tiles;
      const int m0 = (t / ka.num_n_tiles) * BM;
      const int n0 = (t % ka.num_n_tiles) * BN;
      const int buf = local & 1;
      const uint32_t aphase = (uint32_t)(local >> 1) & 1u;
      mbar_wait(tfull_bar(buf), aphase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
      const bool rows_live = (m0 + q * 32) < g.M;
      const int c_begin = half * kChunksPerWarp;
      int c_end = min(c_begin + kChunksPerWarp, (g.N - n0 + 31) / 32);    // live chunks only
      if (!rows_live || (ka.debug & 2)) c_end = c_begin;
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(trow + (uint32_t)(c * 32), v);
        if (ka.ws == nullptr) epilogue_row(g, v, m0 + q * 32 + lane, n0 + c * 32);
        if (c == c_end - 1) {      // last TMEM read of this warp's share: hand the buffer back before storing
          tc_fence_before();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        if (ka.direct_ok && (m0 + q * 32 + 32 <= g.M) && (n0 + c * 32 + 32 <= g.N)) {      // warp-uniform
          if (!(ka.debug & 1)) store_row_direct(ka, v, m0 + q * 32 + lane, n0 + c * 32, split);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((i ^ (lane & 7)) << 2)) =
                make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                            __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          __syncwarp();
          if (!(ka.debug & 1)) store_block(ka, stg, lane, m0 + q * 32, n0 + c * 32, split);
          __syncwarp();
        }
      }
      if (c_end <= c_begin) {
        tc_fence_before();
        if (lane == 0) mbar_arrive(tempty_bar(buf));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =========================================================================================
// CTA-pair variant (cta_group::2): two CTAs on one TPC compute a 256x256 tile.  CTA r loads A rows
// [m0+128r, +128) and B rows/columns [n0+128r, +128); the leader CTA issues one M=256,N=256 MMA per K=8
// that reads A and B from BOTH shared memories, so each SM moves 32 KB per k-block through smem
// instead of 48 KB -- shared-memory bandwidth, not the tensor pipe, is what bounds the FP32-operand
// (TF32) GEMM.  Each CTA keeps its own 128 accumulator rows in its TMEM and runs its own epilogue.
constexpr int kPairStages = 6;
constexpr int kPairThreads = kThreads + 32;                      // + warp 10: the second TMA producer (B operand)
constexpr int kPairStageBytes = 2 * BM * BK * 4;                 // A half + B half, 16 KB each
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kEpiWarps * kEpiStageFloats * 4 + 256 + 1024;

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ KernelArgs ka) {
  constexpr int BN = 256, BMP = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t epi_base = base + kPairStages * kPairStageBytes;
  float* epi_ptr = reinterpret_cast<float*>(base_ptr + kPairStages * kPairStageBytes);
  constexpr int kEpiBytes = kEpiWarps * kEpiStageFloats * 4;
  const uint32_t bar_base = epi_base + kEpiBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPairStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kPairStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kPairStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kPairStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + kPairStages * kPairStageBytes + kEpiBytes + 8 * (2 * kPairStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const GemmArgs& g = ka.g;
  const int tiles = ka.num_m_tiles * ka.num_n_tiles;
  const int items = tiles * ka.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kPairStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();                      // both CTAs' barriers and TMEM exist before anyone signals
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0 || warp == 2 + kEpiWarps) {
    // ===================== TMA producers (two per CTA: warp 0 loads A, warp 10 loads B) =====================
    // An MN-major operand needs four 32-row boxes per k-block (the 128-B swizzle span holds 32 floats), so a weight-gradient
    // product (both operands MN-major) issues 8 TMA instructions per stage; one thread issuing all of them could not keep
    // up with the 512-clk k-block of the MMA (measured main-loop ceiling 555 TF against 750 for the 2- and 5-box products).
    if (lane == 0) {
      const bool load_a = warp == 0;
      int stage = 0; uint32_t phase = 0;
      for (int item = pair; item < items; item += num_pairs) {
        const int split = item / tiles;
        const int t = item - split * tiles;
        const int m0 = (t / ka.num_n_tiles) * BMP + (int)rank * BM;
        const int n0 = (t % ka.num_n_tiles) * BN + (int)rank * 128;
        const int kb0 = split * ka.kblocks_per_split;
        const int kb1 = min(ka.kblocks_total, kb0 + ka.kblocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = base + stage * kPairStageBytes;
          const uint32_t b_dst = a_dst + BM * BK * 4;
          const uint32_t lead_full = mapa_rank(full_bar(stage), 0);
          int ka0 = kb * BK, kb0k = ka0;
          if (ka.x3_kb) x3_coords(kb, ka.x3_kb, ka.x3_cb, ka0, kb0k);
          if (load_a) {
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * kPairStageBytes);   // both CTAs' A and B bytes land here
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < 4; ++j) tma_load_2d_pair(a_dst + j * 4096, &map_a, lead_full, m0 + 32 * j, ka0);
            } else {
              tma_load_2d_pair(a_dst, &map_a, lead_full, ka0, m0);
            }
          } else {
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < 4; ++j) tma_load_2d_pair(b_dst + j * 4096, &map_b, lead_full, n0 + 32 * j, kb0k);
            } else {
              tma_load_2d_pair(b_dst, &map_b, lead_full, kb0k, n0);
            }
          }
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer: one thread of the leader CTA =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                 ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BMP >> 4) << 24);
      int stage = 0; uint32_t phase = 0; int local = 0;
      for (int item = pair; item < items; item += num_pairs, ++local) {
        const int split = item / tiles;
        const int kb0 = split * ka.kblocks_per_split;
        const int kb1 = min(ka.kblocks_total, kb0 + ka.kblocks_per_split);
        const int buf = local & 1;
        const uint32_t aphase = (uint32_t)(local >> 1) & 1u;
        mbar_wait(tempty_bar(buf), aphase ^ 1u);       // both CTAs' epilogues drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_src = base + stage * kPairStageBytes;
          const uint32_t b_src = a_src + BM * BK * 4;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = A_MN ? make_desc(a_src + k * 1024, 4096, 512, 1) : make_desc(a_src + k * 32, 16, 1024, 2);
            const uint64_t bd = B_MN ? make_desc(b_src + k * 1024, 4096, 512, 1) : make_desc(b_src + k * 32, 16, 1024, 2);
            umma_tf32_pair(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(empty_bar(stage));           // frees the slot in both CTAs
          if (kb == kb1 - 1) umma_commit_pair(tfull_bar(buf));
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (each CTA stores its own 128 rows) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kChunksPerWarp = BN / 32 / (kEpiWarps / 4);
    float* stg = epi_ptr + (warp - 2) * kEpiStageFloats;
    int local = 0;
    // rowdot operand: this thread's 512 B of the second matrix for a tile are requested into L2 one tile ahead (the
    // epilogue of a K = 512 product has no slack for a DRAM round trip per chunk)
    auto rowdot_prefetch = [&](int it) {
      const bool want_dot = g.rowdot_out != nullptr && !(ka.debug & 16);
      // the same lead for the residual rows measured SLOWER (same-box A/B at c4: +0.15 ms/step); opt-in for A/B runs only
      const bool want_res = g.residual != nullptr && ka.ws == nullptr && (ka.debug & 32);
      if ((!want_dot && !want_res) || it >= items) return;
      const int tt = it % tiles;
      const int row = (tt / ka.num_n_tiles) * BMP + (int)rank * BM + q * 32 + lane;
      const int col = (tt % ka.num_n_tiles) * BN + half * kChunksPerWarp * 32;
      if (row >= g.M) return;
      if (want_dot) {
        const float* src = g.rowdot_in + (int64_t)row * g.ld_rowdot + col;
#pragma unroll
        for (int i = 0; i < kChunksPerWarp; ++i)
          if (col + i * 32 < g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + i * 32));
      }
      if (want_res) {
        const float* src = g.residual + (int64_t)row * g.ldr + col;
#pragma unroll
        for (int i = 0; i < kChunksPerWarp; ++i)
          if (col + i * 32 < g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + i * 32));
      }
    };
    rowdot_prefetch(pair);
    for (int item = pair; item < items; item += num_pairs, ++local) {
      const int split = item / tiles;
      const int t = item - split * tiles;
      const int m0 = (t / ka.num_n_tiles) * BMP + (int)rank * BM;
      const int n0 = (t % ka.num_n_tiles) * BN;
      const int buf = local & 1;
      const uint32_t aphase = (uint32_t)(local >> 1) & 1u;
      rowdot_prefetch(item + num_pairs);
      mbar_wait(tfull_bar(buf), aphase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
      const uint32_t lead_tempty = mapa_rank(tempty_bar(buf), 0);
      const bool rows_live = (m0 + q * 32) < g.M;
      const int c_begin = half * kChunksPerWarp;
      int c_end = min(c_begin + kChunksPerWarp, (g.N - n0 + 31) / 32);
      if (!rows_live || (ka.debug & 2)) c_end = c_begin;
      float rowdot_carry = 0.f;
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(trow + (uint32_t)(c * 32), v);
        if (ka.ws == nullptr) epilogue_row(g, v, m0 + q * 32 + lane, n0 + c * 32);
        if (g.rowstat_out != nullptr && m0 + q * 32 + lane < g.M) {   // softmax statistics of this row's 32 columns
          const int col0 = n0 + c * 32;
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < g.N) mx = fmaxf(mx, __uint_as_float(v[j]));
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < g.N) sum += exp2f((__uint_as_float(v[j]) - mx) * 1.4426950408889634f);
          g.rowstat_out[(int64_t)(m0 + q * 32 + lane) * ((g.N + 31) >> 5) + (col0 >> 5)] = make_float2(mx, sum);
        }
        if (g.rowdot_out != nullptr) {       // blocked row dots with a second matrix (delta = dO . O of the attention backward)
          const int row = m0 + q * 32 + lane;
          float part = 0.f;
          if (row < g.M) {
            const float* src = g.rowdot_in + (int64_t)row * g.ld_rowdot + n0 + c * 32;
            float t[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i) ld_v8(src + 8 * i, t[i]);
            float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) ps[i] = fmaf(__uint_as_float(v[8 * i + j]), t[i][j], ps[i]);
            part = (ps[0] + ps[1]) + (ps[2] + ps[3]);
          }
          // a 64-wide block is two consecutive chunks of the same thread (chunk ranges start even, hold an even count)
          if (g.rowdot_w == 64 && (c & 1) == 0) rowdot_carry = part;
          else if (row < g.M) {
            const float tot = g.rowdot_w == 64 ? rowdot_carry + part : part;
            const int bi = row / g.rowdot_S, si = row - bi * g.rowdot_S;
            g.rowdot_out[((int64_t)bi * (g.N / g.rowdot_w) + (n0 + c * 32) / g.rowdot_w) * g.rowdot_S + si] = tot;
          }
        }
        if (g.colsum32_out != nullptr) {     // the epilogue warps have slack: the bias-gradient partials ride along
          const float tot = block_colsum32(v, m0 + q * 32 + lane < g.M, lane);
          const int col = n0 + c * 32 + lane;
          if (col < g.N) g.colsum32_out[(int64_t)((m0 + q * 32) >> 5) * g.N + col] = tot;
        }
        if (c == c_end - 1) {
          tc_fence_before();
          if (lane == 0) { if (ka.debug & 4) mbar_arrive_cluster_release(lead_tempty); else mbar_arrive_cluster(lead_tempty); }
        }
        if (ka.direct_ok && (m0 + q * 32 + 32 <= g.M) && (n0 + c * 32 + 32 <= g.N)) {      // warp-uniform
          if (!(ka.debug & 1)) store_row_direct(ka, v, m0 + q * 32 + lane, n0 + c * 32, split);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((i ^ (lane & 7)) << 2)) =
                make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                            __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          __syncwarp();
          if (!(ka.debug & 1)) store_block(ka, stg, lane, m0 + q * 32, n0 + c * 32, split);
          __syncwarp();
        }
      }
      if (c_end <= c_begin) {
        tc_fence_before();
        if (lane == 0) { if (ka.debug & 4) mbar_arrive_cluster_release(lead_tempty); else mbar_arrive_cluster(lead_tempty); }
      }
    }
  }
  tc_fence_before();
  if (ka.debug & 8) cluster_sync_all(); else cluster_sync_relaxed();   // nobody leaves while the peer may still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---- split-K reducer (fixed summation order => deterministic) --------------------------
__global__ void splitk_reduce_kernel(const KernelArgs ka) {
  const GemmArgs& g = ka.g;
  const int64_t total = (int64_t)g.M * g.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / g.N), col = (int)(i - (int64_t)row * g.N);
    float acc = 0.f;
    for (int s = 0; s < ka.splits; ++s) acc += ka.ws[((int64_t)s * g.M + row) * ka.ws_ld + col];
    float o = gemm_epilogue_one(g, acc, row, col);
    float* dst = g.C + (int64_t)row * g.ldc + col;
    if (g.accumulate) o += *dst;
    *dst = o;
  }
}

// weight-gradient case (no fused epilogue operands, 16-B aligned rows): float4 lanes, all split loads of a thread in
// flight before the adds; same split order as the scalar kernel => same bits
__global__ void __launch_bounds__(256) splitk_reduce_vec_kernel(const KernelArgs ka) {
  const GemmArgs& g = ka.g;
  const int n4 = g.N >> 2;
  const int64_t total = (int64_t)g.M * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / n4), col = (int)(i - (int64_t)row * n4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = ka.ws + (int64_t)row * ka.ws_ld + col;
    const int64_t sstride = (int64_t)g.M * ka.ws_ld;
    int s = 0;
    for (; s + 4 <= ka.splits; s += 4) {
      const float4 a = *reinterpret_cast<const float4*>(src + (s + 0) * sstride);
      const float4 b = *reinterpret_cast<const float4*>(src + (s + 1) * sstride);
      const float4 c = *reinterpret_cast<const float4*>(src + (s + 2) * sstride);
      const float4 d = *reinterpret_cast<const float4*>(src + (s + 3) * sstride);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
      acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
      acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
    }
    for (; s < ka.splits; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(src + s * sstride);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    float4* dst = reinterpret_cast<float4*>(g.C + (int64_t)row * g.ldc + col);
    if (g.accumulate) { const float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    *dst = acc;
  }
}

// reducer WITH the fused epilogue (bias -> ReLU -> float gate -> dropout -> residual -> accumulate) on float4 lanes: the
// forward / dgrad products of the 3xTF32 mode whose K exceeds one chunk end here.  Same order of operations per element as
// gemm_epilogue_one, same split order as the scalar kernel => same bits.
__global__ void __launch_bounds__(256) splitk_reduce_epi_vec_kernel(const KernelArgs ka) {
  const GemmArgs& g = ka.g;
  const int n4 = g.N >> 2;
  const int64_t total = (int64_t)g.M * n4;
  const int64_t sstride = (int64_t)g.M * ka.ws_ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / n4), col = (int)(i - (int64_t)row * n4) * 4;
    const float* src = ka.ws + (int64_t)row * ka.ws_ld + col;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 2 <= ka.splits; s += 2) {
      const float4 a = *reinterpret_cast<const float4*>(src + (s + 0) * sstride);
      const float4 b = *reinterpret_cast<const float4*>(src + (s + 1) * sstride);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    for (; s < ka.splits; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(src + s * sstride);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    if (g.bias) { const float4 b = *reinterpret_cast<const float4*>(g.bias + col); acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w; }
    if (g.relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    if (g.gate) {
      const float4 t = *reinterpret_cast<const float4*>(g.gate + (int64_t)row * g.ldg + col);
      acc.x = t.x > 0.f ? acc.x : 0.f; acc.y = t.y > 0.f ? acc.y : 0.f; acc.z = t.z > 0.f ? acc.z : 0.f; acc.w = t.w > 0.f ? acc.w : 0.f;
    }
    if (g.drop.p > 0.f) dropout_four(g.drop, (uint64_t)row * (uint64_t)g.N + (uint64_t)col, acc);
    if (g.residual) { const float4 t = *reinterpret_cast<const float4*>(g.residual + (int64_t)row * g.ldr + col); acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
    float4* dst = reinterpret_cast<float4*>(g.C + (int64_t)row * g.ldc + col);
    if (g.accumulate) { const float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    *dst = acc;
  }
}

// 3xTF32 operand: hi = round-to-nearest TF32 of x (low 13 mantissa bits zero), lo = x - hi (exact in FP32), written side by
// side ALONG K into one buffer with segment stride Kp (a multiple of the 32-float k-block; the gap is zero-filled):
//   k_cols != 0 (K is the contiguous axis, source rows x K):  out[r][k] = hi, out[r][Kp + k] = lo,        ldo >= 2*Kp
//   k_cols == 0 (K is the row axis, source K x cols):          out[k][c] = hi, out[Kp + k][c] = lo,        ldo >= cols
__global__ void split_concat_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols, int k_cols, int Kp,
                                    int64_t ldo, float* __restrict__ out) {
  const int R = k_cols ? rows : Kp, Cn = k_cols ? Kp : (int)ldo;
  const int64_t total = (int64_t)R * Cn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Cn), c = (int)(i - (int64_t)r * Cn);
    const float x = (r < rows && c < cols) ? src[(int64_t)r * ld + c] : 0.f;
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
    const float h = __uint_as_float(hb);
    out[(int64_t)r * ldo + c] = h;
    if (k_cols) out[(int64_t)r * ldo + Kp + c] = x - h;
    else out[((int64_t)Kp + r) * ldo + c] = x - h;
  }
}

// float4 form: source rows 16-B aligned (ld % 4 == 0, cols % 4 == 0); Cn / ldo / Kp are multiples of 4 by construction
__global__ void __launch_bounds__(256) split_concat_vec_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols,
                                                               int k_cols, int Kp, int64_t ldo, float* __restrict__ out) {
  const int R = k_cols ? rows : Kp, C4 = (k_cols ? Kp : (int)ldo) >> 2;
  const int64_t total = (int64_t)R * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C4), c = (int)(i - (int64_t)r * C4) << 2;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows && c < cols) x = *reinterpret_cast<const float4*>(src + (int64_t)r * ld + c);
    uint32_t h0, h1, h2, h3;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0) : "f"(x.x));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1) : "f"(x.y));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h2) : "f"(x.z));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h3) : "f"(x.w));
    const float4 hi = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
    const float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
    *reinterpret_cast<float4*>(out + (int64_t)r * ldo + c) = hi;
    if (k_cols) *reinterpret_cast<float4*>(out + (int64_t)r * ldo + Kp + c) = lo;
    else *reinterpret_cast<float4*>(out + ((int64_t)Kp + r) * ldo + c) = lo;
  }
}

// ---- host side ------------------------------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Split-K workspace: one persistent buffer per (device, stream), grown on demand.  GEMMs on one
// stream are ordered, so consecutive calls can share it; no allocator call on the hot path.
struct SplitWs { float* ptr = nullptr; size_t bytes = 0; };
int splitk_workspace(cudaStream_t st, size_t bytes, float** out) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, SplitWs> table;
  int dev = 0;
  NLPS_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  SplitWs& w = table[{dev, st}];
  if (w.bytes < bytes) {
    NLPS_CUDA(cudaStreamSynchronize(st));
    if (w.ptr) NLPS_CUDA(cudaFree(w.ptr));
    bump_alloc_epoch();
    w.ptr = nullptr; w.bytes = 0;
    const size_t want = std::max(bytes, (size_t)64 << 20);
    NLPS_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.ptr), want));
    w.bytes = want;
  }
  *out = w.ptr;
  return 0;
}

// 3xTF32 operand workspace: like the split-K workspace, one grow-only buffer per (device, stream) -- no allocator call on
// the hot path, so a 3xTF32 step can be captured into a CUDA graph (a re-allocation bumps the epoch and drops graphs)
int x3_workspace(cudaStream_t st, size_t bytes, float** out) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, SplitWs> table;
  int dev = 0;
  NLPS_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  SplitWs& w = table[{dev, st}];
  if (w.bytes < bytes) {
    NLPS_CUDA(cudaStreamSynchronize(st));
    if (w.ptr) NLPS_CUDA(cudaFree(w.ptr));
    bump_alloc_epoch();
    w.ptr = nullptr; w.bytes = 0;
    const size_t want = std::max(bytes + bytes / 4, (size_t)64 << 20);
    NLPS_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.ptr), want));
    w.bytes = want;
  }
  *out = w.ptr;
  return 0;
}

template <int BN>
int launch(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const KernelArgs& ka, int grid) {
  using Cfg = TileCfg<BN>;
  const bool amn = ka.g.trans_a, bmn = !ka.g.trans_b;
#define NLPS_GEMM_CASE(AM, BMN)                                                                              \
  {                                                                                                          \
    auto kern = gemm_tf32_kernel<BN, AM, BMN>;                                                               \
    static bool attr_set = false;                                                                            \
    if (!attr_set) {                                                                                         \
      NLPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));   \
      attr_set = true;                                                                                       \
    }                                                                                                        \
    kern<<<grid, kThreads, Cfg::kSmemBytes, st>>>(ma, mb, ka);                                               \
  }
  if (amn && bmn) NLPS_GEMM_CASE(true, true)
  else if (amn) NLPS_GEMM_CASE(true, false)
  else if (bmn) NLPS_GEMM_CASE(false, true)
  else NLPS_GEMM_CASE(false, false)
#undef NLPS_GEMM_CASE
  NLPS_LAUNCH_CHECK();
  return 0;
}

template <bool AM, bool BMN>
int launch_pair_one(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const KernelArgs& ka, int grid) {
  auto kern = gemm_tf32_pair_kernel<AM, BMN>;
  static bool attr_set = false;
  if (!attr_set) {
    NLPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    attr_set = true;
  }
  kern<<<grid, kPairThreads, kPairSmemBytes, st>>>(ma, mb, ka);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int launch_pair(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const KernelArgs& ka, int grid) {
  const bool amn = ka.g.trans_a, bmn = !ka.g.trans_b;
  if (amn && bmn) return launch_pair_one<true, true>(st, ma, mb, ka, grid);
  if (amn) return launch_pair_one<true, false>(st, ma, mb, ka, grid);
  if (bmn) return launch_pair_one<false, true>(st, ma, mb, ka, grid);
  return launch_pair_one<false, false>(st, ma, mb, ka, grid);
}

int pair_mode() {                 // NLPS_GEMM_PAIR=0 disables the CTA-pair kernel (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("NLPS_GEMM_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

// x3 != 0: g.A / g.B are K-concatenated hi|lo operands (split_concat_kernel) with segment stride ceil(K/32)*32 and g.K is
// the K of the original product; the kernel walks 3 * ceil(K/32) virtual k-blocks (x3_coords).
int gemm_tf32_impl(cudaStream_t st, const GemmArgs& g, bool round_tf32, bool x3 = false) {
  if (g.M <= 0 || g.N <= 0) return 0;
  NLPS_REQUIRE(g.K > 0, "gemm_tf32: K must be positive");
  NLPS_REQUIRE(gemm_tf32_supported(g), "gemm_tf32: operands not TMA-addressable (need 16-B aligned bases, ld %% 4 == 0)");
  // tile shape: CTA pairs on 256x256 tiles when both extents allow it, else 128x256 / 128x128 single CTAs
  const bool pair = pair_mode() && g.M > 128 && g.N > 128;
  const int BN = g.N > 128 ? 256 : 128;
  const int BMT = pair ? 256 : BM;
  const int units = pair ? kNumSMs / 2 : kNumSMs;          // schedulable units per wave
  KernelArgs ka{};
  ka.g = g;
  ka.num_m_tiles = (int)ceil_div(g.M, BMT);
  ka.num_n_tiles = (int)ceil_div(g.N, BN);
  const int seg_kb = (int)ceil_div(g.K, BK);
  ka.x3_kb = x3 ? seg_kb : 0;
  ka.kblocks_total = x3 ? 3 * seg_kb : seg_kb;
  const uint64_t k_extent = x3 ? (uint64_t)2 * seg_kb * BK : (uint64_t)g.K;     // K extent of the tensor maps
  const int tiles = ka.num_m_tiles * ka.num_n_tiles;
  int splits = 1;
  ka.x3_cb = 0;
  if (x3) {
    // one split-K item per chunk of x3_cb original k-blocks.  Default 16 = 512 floats of K (64 accumulation steps per
    // chain: the K = 512 products of the step need no split at all and keep their fused in-kernel epilogue; measured
    // norm-wise error < 2e-6 per product); NLPS_X3_CHUNK overrides
    static int cb_env = -1;
    if (cb_env < 0) { const char* e = getenv("NLPS_X3_CHUNK"); cb_env = e ? std::max(1, atoi(e)) : 16; }
    ka.x3_cb = std::min(cb_env, seg_kb);
    splits = (int)ceil_div(seg_kb, ka.x3_cb);
  } else if (tiles < 2 * units && ka.kblocks_total >= 64) {
    // weight-gradient shaped (few tiles, long K): pick the split that minimises
    //   waves(s) * (k-blocks per split + fixed per-tile cost)   with whole waves of CTAs / CTA pairs
    double best = 1e30;
    for (int s = 1; s <= 32 && ka.kblocks_total / s >= 8; ++s) {
      const int64_t waves = ceil_div((int64_t)tiles * s, units);
      const double t = (double)waves * ((double)ceil_div(ka.kblocks_total, s) + 24.0);
      if (t < best - 1e-9) { best = t; splits = s; }
    }
  }
  ka.kblocks_per_split = x3 ? 3 * ka.x3_cb : (int)ceil_div(ka.kblocks_total, splits);
  splits = (int)ceil_div(ka.kblocks_total, ka.kblocks_per_split);
  ka.splits = splits;
  ka.vec_ok = (aligned16(g.C) && g.ldc % 4 == 0 &&
               (!g.residual || (aligned16(g.residual) && g.ldr % 4 == 0)) &&
               (!g.gate || (aligned16(g.gate) && g.ldg % 4 == 0))) ? 1 : 0;
  NLPS_REQUIRE(!(g.relu_bits_out || g.gate_bits) || (g.N % 32 == 0 && splits == 1),
               "gemm_tf32: bit-packed ReLU masks need N %% 32 == 0 and no split-K");
  NLPS_REQUIRE((g.colsum32_out == nullptr && g.rowstat_out == nullptr && g.rowdot_out == nullptr) ||
                   (pair && splits == 1 && !g.gate && !g.residual && !g.accumulate),
               "gemm_tf32: colsum32_out / rowstat_out / rowdot_out need the CTA-pair kernel without split-K, float gate, residual or accumulate");
  NLPS_REQUIRE(g.rowdot_out == nullptr || gemm_tf32_rowdot_fusable(g), "gemm_tf32: rowdot_out is not fusable for this product");
  ka.ws = nullptr;
  ka.ws_ld = round_up(g.N, 8);
  {
    auto a32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
    static int nodirect = -1;
    if (nodirect < 0) { const char* e = getenv("NLPS_GEMM_STAGED"); nodirect = (e && e[0] == '1') ? 1 : 0; }
    ka.direct_ok = (!nodirect && a32(g.C) && g.ldc % 8 == 0 && (!g.residual || (a32(g.residual) && g.ldr % 8 == 0)) &&
                    (!g.gate || (a32(g.gate) && g.ldg % 8 == 0))) ? 1 : 0;
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("NLPS_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
    ka.debug = dbg;
    static int hint = -1;      // 0 none, 1 every direct store, 2 only outputs >= 96 MB, 3 only outputs < 96 MB
    if (hint < 0) { const char* e = getenv("NLPS_GEMM_STORE_HINT"); hint = e ? atoi(e) : 1; }
    const bool big = (int64_t)g.M * g.N * 4 >= ((int64_t)96 << 20);
    ka.store_hint = (hint == 1 || (hint == 2 && big) || (hint == 3 && !big)) ? 1 : 0;
  }
  if (splits > 1) NLPS_TRY(splitk_workspace(st, (size_t)splits * g.M * ka.ws_ld * sizeof(float), &ka.ws));
  CUtensorMap ma, mb;
  const uint32_t b_rows = pair ? 128u : (uint32_t)BN;      // a CTA of a pair loads half of the B tile
  if (g.trans_a) NLPS_TRY(make_map(&ma, g.A, (uint64_t)g.M, k_extent, g.lda, 32, BK, round_tf32, true));
  else NLPS_TRY(make_map(&ma, g.A, k_extent, (uint64_t)g.M, g.lda, BK, BM, round_tf32, false));
  if (g.trans_b) NLPS_TRY(make_map(&mb, g.B, k_extent, (uint64_t)g.N, g.ldb, BK, b_rows, round_tf32, false));
  else NLPS_TRY(make_map(&mb, g.B, (uint64_t)g.N, k_extent, g.ldb, 32, BK, round_tf32, true));
  const int items = tiles * splits;
  count_kind(pair ? KIND_GEMM_PAIR : KIND_GEMM_TC_SINGLE);
  if (pair) {
    NLPS_TRY(launch_pair(st, ma, mb, ka, 2 * std::min(items, units)));
  } else {
    const int grid = std::min(items, kNumSMs);
    if (BN == 256) NLPS_TRY(launch<256>(st, ma, mb, ka, grid));
    else NLPS_TRY(launch<128>(st, ma, mb, ka, grid));
  }
  if (splits > 1) {
    const int64_t total = (int64_t)g.M * g.N;
    const bool plain = !g.bias && !g.relu && !g.gate && !g.gate_bits && !g.relu_bits_out && g.drop.p <= 0.f && !g.residual &&
                       g.N % 4 == 0 && ka.ws_ld % 4 == 0 && g.ldc % 4 == 0 && aligned16(g.C) && aligned16(ka.ws);
    const bool vec_epi = !g.gate_bits && !g.relu_bits_out && g.N % 4 == 0 && ka.ws_ld % 4 == 0 && g.ldc % 4 == 0 && aligned16(g.C) &&
                         aligned16(ka.ws) && (!g.bias || aligned16(g.bias)) && (!g.gate || (aligned16(g.gate) && g.ldg % 4 == 0)) &&
                         (!g.residual || (aligned16(g.residual) && g.ldr % 4 == 0));
    if (plain) {
      const int rgrid = (int)std::min<int64_t>(ceil_div(total / 4, 256), kNumSMs * 8);
      splitk_reduce_vec_kernel<<<rgrid, 256, 0, st>>>(ka);
    } else if (vec_epi) {
      const int rgrid = (int)std::min<int64_t>(ceil_div(total / 4, 256), kNumSMs * 16);
      splitk_reduce_epi_vec_kernel<<<rgrid, 256, 0, st>>>(ka);
    } else {
      const int rgrid = (int)std::min<int64_t>(ceil_div(total, 256), kNumSMs * 8);
      splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(ka);
    }
    NLPS_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace

bool gemm_tf32_supported(const GemmArgs& g) {
  return aligned16(g.A) && aligned16(g.B) && g.lda % 4 == 0 && g.ldb % 4 == 0 && g.M > 0 && g.N > 0 && g.K > 0;
}

int gemm_tf32(cudaStream_t st, const GemmArgs& g) { return gemm_tf32_impl(st, g, /*round_tf32=*/true); }

// mirrors the tile / split choice of gemm_tf32_impl
bool gemm_tf32_colsum_fusable(const GemmArgs& g) {
  if (!gemm_tf32_supported(g) || !(pair_mode() && g.M > 128 && g.N > 128)) return false;
  if (g.gate || g.residual || g.accumulate) return false;
  const int64_t tiles = ceil_div(g.M, 256) * ceil_div(g.N, 256);
  const int64_t kblocks = ceil_div(g.K, BK);
  return !(tiles < 2 * (kNumSMs / 2) && kblocks >= 64);      // no split-K
}

bool gemm_tf32_rowdot_fusable(const GemmArgs& g) {
  if (!gemm_tf32_colsum_fusable(g)) return false;
  if (g.rowdot_w != 32 && g.rowdot_w != 64) return false;
  if (g.N % g.rowdot_w != 0 || g.rowdot_S <= 0 || g.M % g.rowdot_S != 0) return false;
  return g.rowdot_in != nullptr && (reinterpret_cast<uintptr_t>(g.rowdot_in) & 31u) == 0 && g.ld_rowdot % 8 == 0;
}

// 3xTF32: A = Ah + Al, B = Bh + Bl with Ah, Bh exactly representable in TF32; C = Al*Bh + Ah*Bl + Ah*Bh, dropping only
// Al*Bl (~2^-22 relative).  ONE tensor-core launch over a virtual K axis three segments long: the operands are written once
// as [Ah | Al] and [Bh ; Bl] (split_concat_kernel) and the kernel's TMA producer pairs the segments chunk by chunk
// (x3_coords); chunks of 512 K are split-K items whose partials the reducer adds in FP32 (round to nearest) before the fused
// epilogue (bias / ReLU / gate / dropout / residual) -- no allocator call (CUDA-graph capturable).
int gemm_tf32x3(cudaStream_t st, const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0) return 0;
  const int Kp = (int)ceil_div(g.K, BK) * BK;
  const int64_t lda2 = g.trans_a ? round_up(g.M, 4) : 2 * (int64_t)Kp;
  const int64_t ldb2 = g.trans_b ? 2 * (int64_t)Kp : round_up(g.N, 4);
  const size_t na = (size_t)(g.trans_a ? 2 * (int64_t)Kp : (int64_t)g.M) * lda2;
  const size_t nb = (size_t)(g.trans_b ? (int64_t)g.N : 2 * (int64_t)Kp) * ldb2;
  const size_t na_al = (na + 63) / 64 * 64;                      // keep B2 256-B aligned
  float* buf = nullptr;
  NLPS_TRY(x3_workspace(st, (na_al + nb) * sizeof(float), &buf));
  float *a2 = buf, *b2 = buf + na_al;
  auto split = [&](const float* src, int64_t ld, int rows, int cols, int k_cols, int64_t ldo, float* out) {
    const int64_t n_out = (int64_t)(k_cols ? rows : Kp) * (k_cols ? Kp : ldo);
    if (aligned16(src) && ld % 4 == 0 && cols % 4 == 0) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_out / 4, 256), kNumSMs * 16));
      split_concat_vec_kernel<<<grid, 256, 0, st>>>(src, ld, rows, cols, k_cols, Kp, ldo, out);
    } else {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_out, 256), kNumSMs * 8));
      split_concat_kernel<<<grid, 256, 0, st>>>(src, ld, rows, cols, k_cols, Kp, ldo, out);
    }
  };
  if (g.trans_a) split(g.A, g.lda, g.K, g.M, 0, lda2, a2); else split(g.A, g.lda, g.M, g.K, 1, lda2, a2);
  NLPS_LAUNCH_CHECK();
  if (g.trans_b) split(g.B, g.ldb, g.N, g.K, 1, ldb2, b2); else split(g.B, g.ldb, g.K, g.N, 0, ldb2, b2);
  NLPS_LAUNCH_CHECK();
  GemmArgs p = g;
  p.A = a2; p.lda = lda2; p.B = b2; p.ldb = ldb2;
  return gemm_tf32_impl(st, p, false, /*x3=*/true);
}

int gemm_dispatch(cudaStream_t st, int precision, const GemmArgs& g) {
  if (precision == 0 || !gemm_tf32_supported(g)) { count_kind(KIND_GEMM_SIMT); return gemm_fp32(st, g); }
  if (precision == 2) { count_kind(KIND_GEMM_X3); return gemm_tf32x3(st, g); }
  return gemm_tf32(st, g);
}

}  // namespace nlps
