// Fused causal + key-padding attention BACKWARD on tcgen05 with the A operands in tensor memory
// ("TS" form of tcgen05.mma), head_dim 32 or 64.  Replaces transformer.py:427-441 (recomputed P@V,
// softmax_backward, dQ/dK/dV) together with the masks of :116-118; same outputs as attention.cu.
//
// Why TS: with FP32 (TF32) operands an SS-mode M=128,N=64 MMA reads 6 KB of shared memory per 32-cycle
// instruction -- more than the 128 B/clk the SM has -- and the P / dS tiles have to bounce through shared
// memory.  Here every A operand lives in TMEM:
//   * the CTA's resident tile (K and V in the dK/dV kernel, Q and dO in the dQ kernel) is copied
//     smem -> TMEM once in the prologue;
//   * the softmax threads overwrite S with P (and dP with dS) in place in TMEM (tcgen05.st), FA4 style.
// Shared memory then only holds the streamed B operands, NS stages deep, and the tensor pipe, the TMA
// and the 8 softmax warps run concurrently: S/dP accumulators are double-buffered in TMEM so the MMA
// warp issues the products of block it+1 / it+2 while the threads work on block it.
//
//   dK/dV kernel: CTA = (batch, head, 128 keys); thread (r, half) owns key r and 32 of the 64 queries of a
//     block.  S^T = K Q^T and dP^T = V dO^T (M = keys), P^T and dS^T = P^T o (dP^T - delta) in place,
//     dV += P^T dO and dK += dS^T Q accumulate in TMEM over the whole loop.  No atomics.
//   dQ kernel:    CTA = (batch, head, 128 queries); S = Q K^T, dP = dO V^T, dS in place, dQ += dS K.
#include "tc_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>

namespace nlps {
namespace {
using namespace tc;

constexpr int kThreads = 448;   // warp0 TMA + lse/delta staging, warp1 MMA + TMEM owner, warps 2-9 softmax, warps 10-13 re-swizzle
constexpr int kSwzWarps = 4;
constexpr int kPsThreads = 480;  // persistent backward kernels: + warp 14, a second MMA issuer (see attn_bwd_dkdv_ps)
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kTraceSlots = 4096;

struct Dims {
  int B, S, h, E;
  float scale_log2;   // (1/sqrt(d)) * log2(e)
  int trace;
};

__device__ long long g_ts_trace[kTraceSlots];
__device__ __forceinline__ void trace(bool on, int slot) { if (on && slot < kTraceSlots) g_ts_trace[slot] = clock64(); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// pins the first USE of a value whose load was issued earlier: the empty volatile asm cannot move above the
// barrier waits around it, so the scoreboard wait for the load lands here and not at the top of the kernel
__device__ __forceinline__ int use_here(int v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ uint32_t low_bits32(int n) {   // n lowest bits set, n clamped to [0,32]
  return n <= 0 ? 0u : (n >= 32 ? ~0u : ((1u << n) - 1u));
}

// one lane of a converged warp (lets the compiler issue tcgen05.mma without a per-instruction election loop)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// packed FP32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100): halves the issue slots of the per-element math
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) { return pack2(__float_as_uint(lo), __float_as_uint(hi)); }
__device__ __forceinline__ void unpack2(uint64_t v, uint32_t& lo, uint32_t& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// round-to-nearest (ties away) onto the TF32 grid for finite values: the tensor core ignores the low 13
// mantissa bits, so adding half a TF32 ulp to the bit pattern is all cvt.rna.tf32 would do
__device__ __forceinline__ uint32_t rna_tf32_bits(uint32_t bits) { return bits + 0x1000u; }

// P (optional) and dS for two adjacent columns.  s2/dp2: raw scores and dP; nl2: -lse*log2e; ndl2: -delta
template <bool MASKED, bool WANT_P>
__device__ __forceinline__ void p_ds_pair(uint32_t& s_lo, uint32_t& s_hi, uint32_t& d_lo, uint32_t& d_hi, uint64_t scale2,
                                          uint64_t nl2, uint64_t ndl2, uint32_t vis, int c) {
  uint32_t x_lo, x_hi;
  unpack2(ffma2(pack2(s_lo, s_hi), scale2, nl2), x_lo, x_hi);
  float e0 = fast_exp2(__uint_as_float(x_lo)), e1 = fast_exp2(__uint_as_float(x_hi));
  if (MASKED) {
    e0 = (vis & (1u << c)) ? e0 : 0.f;
    e1 = (vis & (2u << c)) ? e1 : 0.f;
  }
  uint32_t r_lo, r_hi;
  unpack2(fmul2(pack2f(e0, e1), fadd2(pack2(d_lo, d_hi), ndl2)), r_lo, r_hi);
  if (WANT_P) {
    s_lo = rna_tf32_bits(__float_as_uint(e0));
    s_hi = rna_tf32_bits(__float_as_uint(e1));
  }
  d_lo = rna_tf32_bits(r_lo);
  d_hi = rna_tf32_bits(r_hi);
}

// [128 rows][D floats] tile in shared memory, 16-byte chunks XOR-swizzled by (row & 7): thread-per-row writes and
// warp-per-row reads are both bank-conflict free
template <int D>
__device__ __forceinline__ void tile_put_row32(uint8_t* tile, int r, int col0, const uint32_t (&v)[32], float scale) {
  uint8_t* row = tile + r * (D * 4);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int chunk = (col0 >> 2) + j;
    *reinterpret_cast<float4*>(row + ((chunk ^ (r & 7)) << 4)) =
        make_float4(__uint_as_float(v[4 * j]) * scale, __uint_as_float(v[4 * j + 1]) * scale,
                    __uint_as_float(v[4 * j + 2]) * scale, __uint_as_float(v[4 * j + 3]) * scale);
  }
}
// rows [0, n_rows) of the tile -> global rows of stride ld floats, one coalesced row per warp instruction;
// NW warps share the rows, four loads in flight per warp
template <int D, int NW = 8>
__device__ __forceinline__ void tile_store_rows(const uint8_t* tile, float* out, int64_t ld, int n_rows, int w, int lane) {
  for (int r0 = w; r0 < n_rows; r0 += 4 * NW) {
    if (D == 64) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = min(r0 + u * NW, 127);
        v[u] = *reinterpret_cast<const float2*>(tile + r * (D * 4) + (((lane >> 1) ^ (r & 7)) << 4) + ((lane & 1) << 3));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r0 + u * NW < n_rows) *reinterpret_cast<float2*>(out + (r0 + u * NW) * ld + 2 * lane) = v[u];
    } else {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = min(r0 + u * NW, 127);
        v[u] = *reinterpret_cast<const float*>(tile + r * (D * 4) + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r0 + u * NW < n_rows) out[(r0 + u * NW) * ld + lane] = v[u];
    }
  }
}

// one 128-byte row (32 floats) of a K-major SWIZZLE_128B tile [rows][128 B] -> registers
__device__ __forceinline__ void load_swizzled_row(const uint8_t* tile, int r, uint32_t (&v)[32]) {
  const uint8_t* row = tile + r * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 w = *reinterpret_cast<const uint4*>(row + ((j ^ (r & 7)) << 4));
    v[4 * j] = w.x; v[4 * j + 1] = w.y; v[4 * j + 2] = w.z; v[4 * j + 3] = w.w;
  }
}

// A streamed tile is needed in both operand majors (contracted over head_dim for S / dP, over its rows for
// dK / dV / dQ).  For FP32 the two shared-memory images hold the same row-major rows of 128 B and differ only
// in the swizzle: K-major SWIZZLE_128B stores 16-B chunk i of row r at i ^ (r & 7), the MN-major
// SWIZZLE_128B_BASE32B image stores 32-B chunk j at j ^ (r & 3).  Two warps permute the chunks of the landed
// K-major image into the second image -- the L2 -> SM traffic of the backward kernels (their bound) halves.
__device__ __forceinline__ void reswizzle_rows(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n_rows, int w, int lane) {
  const int i = lane & 7;                                      // logical 16-B chunk of the row
  for (int r0 = w * 4 + (lane >> 3); r0 < n_rows; r0 += 32 * kSwzWarps) {   // 8 loads in flight, then 8 stores
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = min(r0 + 4 * kSwzWarps * u, n_rows - 1);
      v[u] = *reinterpret_cast<const uint4*>(src + r * 128 + ((i ^ (r & 7)) << 4));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + 4 * kSwzWarps * u;
      if (r < n_rows) *reinterpret_cast<uint4*>(dst + r * 128 + (((((i >> 1) ^ (r & 3)) << 1) | (i & 1)) << 4)) = v[u];
    }
  }
}

template <int D, int NS>
struct Cfg {
  static constexpr int kAtoms = D / 32;                      // 128-byte atoms per row
  static constexpr int kBlk = kAtoms * 64 * 128;             // a streamed 64-row operand tile
  static constexpr int kStage = 2 * kBlk;                    // two tiles per stage (also = one resident 128-row tile)
  static constexpr int kCtl = 2048;                          // barriers (256 B) + lse/delta stages (2 x 2 x 64 floats)
  static constexpr int kSmemDkdv = 2 * NS * kStage + kCtl + 1024;
  static constexpr int kSmemDq = NS * kStage + NS * kBlk + kCtl + 1024;
  static constexpr int kTmemCols = 512;
};

// ============================================================================================
// dK / dV
// ============================================================================================
template <int D, int NS>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_dkdv_ts(const __grid_constant__ CUtensorMap map_res,   // qkv, box {32,128}, K-major
                                                              const __grid_constant__ CUtensorMap map_qk,    // qkv, box {32,64}, K-major
                                                              const __grid_constant__ CUtensorMap map_qmn,   // qkv, box {32,64}, MN-major
                                                              const __grid_constant__ CUtensorMap map_dok,   // dctx, box {32,64}, K-major
                                                              const __grid_constant__ CUtensorMap map_domn,  // dctx, box {32,64}, MN-major
                                                              const Dims p, const int32_t* __restrict__ x,
                                                              const int32_t* __restrict__ fnp,
                                                              const float* __restrict__ lse,
                                                              const float* __restrict__ delta,
                                                              float* __restrict__ dqkv) {
  using C = Cfg<D, NS>;
  static_assert(NS >= 2, "the prologue parks K and V in the first two MN stages");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t kst = base;                                  // NS stages of [Q K-major | dO K-major]
  const uint32_t mst = base + NS * C::kStage;                 // NS stages of [Q MN-major | dO MN-major]
  uint8_t* mst_ptr = base_ptr + NS * C::kStage;
  const uint32_t bar = base + 2 * NS * C::kStage;
  uint8_t* ctl_ptr = base_ptr + 2 * NS * C::kStage;
  // barriers
  auto kfull = [&](int s) { return bar + 8u * s; };
  auto kfree = [&](int s) { return bar + 8u * (4 + s); };
  auto mfull = [&](int s) { return bar + 8u * (8 + s); };
  auto mfree = [&](int s) { return bar + 8u * (12 + s); };
  auto sdp_full = [&](int b2) { return bar + 8u * (16 + b2); };
  auto pds_full = [&](int b2) { return bar + 8u * (18 + b2); };
  auto ld_full = [&](int b2) { return bar + 8u * (20 + b2); };
  const uint32_t kv_full = bar + 8u * 22, kv_done = bar + 8u * 23, acc_full = bar + 8u * 24, tmem_slot = bar + 8u * 25;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(ctl_ptr + 8 * 25);
  float* lse_s = reinterpret_cast<float*>(ctl_ptr + 256);     // [2][64]  -lse*log2e
  float* delta_s = lse_s + 128;                               // [2][64]  -delta

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y;
  const int k0 = ((int)gridDim.x - 1 - (int)blockIdx.x) * 128;   // heavy (early-key) tiles are scheduled first
  // first_real is requested here and first consumed after the causal blocks (global loads take thousands of
  // cycles while TMA traffic saturates the L2): query blocks [first_causal, n_q) first, then the fully-masked
  // leading blocks [0, n_deg) whose rows attend to every key
  const int first_real = __ldg(fnp + b);
  const int row_base = b * p.S;
  const int n_q = (p.S + 63) / 64;
  const int first_causal = k0 / 64;                               // first query block with q >= some key here
  const int n_causal = n_q - first_causal;
  auto more_blocks = [&](int it) { return it < n_causal || it - n_causal < min(min((use_here(first_real) + 63) / 64, n_q), first_causal); };
  auto block_of = [&](int it) { return it < n_causal ? first_causal + it : it - n_causal; };
  const bool tr = p.trace && blockIdx.x == gridDim.x - 1 && blockIdx.y == 0 && blockIdx.z == gridDim.z / 2;
  trace(tr && threadIdx.x == 0, 0);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(kfull(s), 1); mbar_init(kfree(s), 1 + kSwzWarps); mbar_init(mfull(s), kSwzWarps); mbar_init(mfree(s), 1); }
    for (int b2 = 0; b2 < 2; ++b2) { mbar_init(sdp_full(b2), 1); mbar_init(pds_full(b2), 256); mbar_init(ld_full(b2), 32); }
    mbar_init(kv_full, 1); mbar_init(kv_done, 256); mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    // the resident K and V tiles start moving before the TMEM allocation / setup barrier
    mbar_expect_tx(kv_full, 2 * C::kStage);
#pragma unroll
    for (int a = 0; a < C::kAtoms; ++a) {
      tma_load_2d(mst + a * (128 * 128), &map_res, kv_full, p.E + head * D + a * 32, row_base + k0);
      tma_load_2d(mst + C::kStage + a * (128 * 128), &map_res, kv_full, 2 * p.E + head * D + a * 32, row_base + k0);
    }
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t t_k = tmem, t_v = tmem + D, t_st = tmem + 2 * D, t_dpt = t_st + 128, t_dv = t_dpt + 128, t_dk = t_dv + D;

  if (warp == 0) {
    // ---------------- producer: TMA + lse/delta staging ----------------
    const int64_t so = ((int64_t)b * p.h + head) * p.S;
    // lse/delta of a block are fetched one iteration ahead so that their latency never stalls the TMA lane
    // (raw values only: any arithmetic on them here would make the warp wait for the loads right away)
    float nl_r[2], nd_r[2];
    auto fetch_ld = [&](int it) {
      const int q0 = block_of(it) * 64;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int qi = min(q0 + lane + 32 * j, p.S - 1);
        nl_r[j] = __ldg(lse + so + qi);
        nd_r[j] = __ldg(delta + so + qi);
      }
    };
    fetch_ld(0);
    for (int it = 0; more_blocks(it); ++it) {
      const int s = it % NS;
      const uint32_t free_par = (uint32_t)((it / NS) & 1) ^ 1u;
      const int q0 = block_of(it) * 64;
      if (lane == 0) {
        const int qrow = row_base + q0;
        mbar_wait(kfree(s), free_par);
        trace(tr, 100 + it * 2);
        mbar_expect_tx(kfull(s), C::kStage);
#pragma unroll
        for (int a = 0; a < C::kAtoms; ++a) {
          tma_load_2d(kst + s * C::kStage + a * (64 * 128), &map_qk, kfull(s), head * D + a * 32, qrow);
          tma_load_2d(kst + s * C::kStage + C::kBlk + a * (64 * 128), &map_dok, kfull(s), head * D + a * 32, qrow);
        }
      }
      __syncwarp();
      // lse / delta of the 64 queries of this block (2 per lane), exp2 domain, negated for the FFMA
      const int b2 = it & 1;
      if (it >= 2) mbar_wait(pds_full(b2), (uint32_t)(((it - 2) >> 1) & 1));   // threads are done with this stage
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        lse_s[b2 * 64 + lane + 32 * j] = -nl_r[j] * kLog2e;     // columns past the sequence end are masked by vis
        delta_s[b2 * 64 + lane + 32 * j] = -nd_r[j];
      }
      mbar_arrive(ld_full(b2));
      if (more_blocks(it + 1)) fetch_ld(it + 1);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the warp stays converged, one elected lane issues ----------------
    constexpr uint32_t idesc_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_a = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto issue_sdp = [&](int it) {        // S^T = K Q^T ; dP^T = V dO^T   (A = K / V in TMEM)
      const int s = it % NS, b2 = it & 1;
      mbar_wait(kfull(s), (uint32_t)((it / NS) & 1));
      tc_fence_after();
      trace(tr && lane == 0, 200 + it * 4);
      if (elect_one_sync()) {
        const uint64_t qk = make_desc(kst + s * C::kStage, 16, 1024, 2), dok = make_desc(kst + s * C::kStage + C::kBlk, 16, 1024, 2);
#pragma unroll
        for (int kk = 0; kk < D / 8; ++kk) {
          const uint64_t bo = (uint64_t)(((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4);
          umma_tf32_ts(t_st + b2 * 64, t_k + kk * 8, qk + bo, idesc_t, kk > 0 ? 1u : 0u);
          umma_tf32_ts(t_dpt + b2 * 64, t_v + kk * 8, dok + bo, idesc_t, kk > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full(b2));
        umma_commit(kfree(s));
      }
      __syncwarp();
      trace(tr && lane == 0, 201 + it * 4);
    };
    mbar_wait(kv_done, 0);
    tc_fence_after();
    issue_sdp(0);
    if (more_blocks(1)) issue_sdp(1);
    for (int it = 0; more_blocks(it); ++it) {
      const int s = it % NS, b2 = it & 1;
      mbar_wait(mfull(s), (uint32_t)((it / NS) & 1));
      trace(tr && lane == 0, 500 + it);
      mbar_wait(pds_full(b2), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      trace(tr && lane == 0, 202 + it * 4);
      if (elect_one_sync()) {
        const uint64_t qmn = make_desc(mst + s * C::kStage, 64 * 128, 512, 1), domn = make_desc(mst + s * C::kStage + C::kBlk, 64 * 128, 512, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {      // dV += P^T dO ; dK += dS^T Q   (K = 64 queries, A = P^T / dS^T in TMEM)
          const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
          umma_tf32_ts(t_dv, t_st + b2 * 64 + kk * 8, domn + (uint64_t)(kk * 64), idesc_a, acc);
          umma_tf32_ts(t_dk, t_dpt + b2 * 64 + kk * 8, qmn + (uint64_t)(kk * 64), idesc_a, acc);
        }
        umma_commit(mfree(s));
      }
      __syncwarp();
      trace(tr && lane == 0, 203 + it * 4);
      if (more_blocks(it + 2)) issue_sdp(it + 2);  // reuses the S^T/dP^T buffer just consumed: same-thread MMAs retire in order
    }
    if (elect_one_sync()) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 10) {
    // ---------------- re-swizzle warps: MN-major images of the landed Q / dO tiles ----------------
    const int w2 = warp - 10;
    mbar_wait(kv_done, 0);                             // K and V have left the first two MN stages
    for (int it = 0; more_blocks(it); ++it) {
      const int s = it % NS;
      const bool trc = tr && w2 == 0 && lane == 0 && it < 16;
      trace(trc, 600 + it * 4);
      mbar_wait(mfree(s), (uint32_t)((it / NS) & 1) ^ 1u);
      trace(trc, 601 + it * 4);
      mbar_wait(kfull(s), (uint32_t)((it / NS) & 1));
      trace(trc, 602 + it * 4);
      reswizzle_rows(base_ptr + s * C::kStage, mst_ptr + s * C::kStage, 2 * C::kAtoms * 64, w2, lane);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(mfull(s)); mbar_arrive(kfree(s)); }
      trace(trc, 603 + it * 4);
    }
  } else {
    // ---------------- softmax threads: (key row r, query half) ----------------
    const int cw = warp - 2;
    const int q = warp & 3;                         // TMEM lane quarter this warp may touch
    const int half = cw >> 2;
    const int r = q * 32 + lane;
    const int key = k0 + r;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const bool key_ok = key < p.S;
    const bool key_pad = key_ok ? (__ldg(x + (int64_t)b * p.S + key) == 0) : true;
    const bool trs = tr && threadIdx.x == 64;

    // prologue: resident K and V tiles smem -> TMEM (32 columns per thread)
    mbar_wait(kv_full, 0);
    {
      uint32_t v[32];
      if (D == 64) {
        load_swizzled_row(mst_ptr + half * (128 * 128), r, v);
        tmem_st32(t_k + lane_addr + half * 32, v);
        load_swizzled_row(mst_ptr + C::kStage + half * (128 * 128), r, v);
        tmem_st32(t_v + lane_addr + half * 32, v);
      } else {
        load_swizzled_row(mst_ptr + half * C::kStage, r, v);
        tmem_st32((half ? t_v : t_k) + lane_addr, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(kv_done);
    }

    for (int it = 0; more_blocks(it); ++it) {
      const int b2 = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      const int qb = block_of(it) * 64 + half * 32;   // first query of this thread's 32 columns
      // visibility of (query qb+c, this key): causal & not PAD, or the query is a fully-masked row
      const uint32_t range_bits = low_bits32(p.S - qb);
      const uint32_t causal = (key_ok && !key_pad) ? ~low_bits32(key - qb) : 0u;
      const uint32_t degq = key_ok ? low_bits32(first_real - qb) : 0u;
      const uint32_t vis = (causal | degq) & range_bits;
      trace(trs, 300 + it * 3);
      mbar_wait(ld_full(b2), ph);
      mbar_wait(sdp_full(b2), ph);
      tc_fence_after();
      trace(trs, 301 + it * 3);
      uint32_t sv[32], dv[32];
      const uint32_t a_s = t_st + b2 * 64 + half * 32 + lane_addr, a_d = t_dpt + b2 * 64 + half * 32 + lane_addr;
      tmem_ld32_nowait(a_s, sv);
      tmem_ld32_nowait(a_d, dv);
      tmem_ld_wait();
      const float4* l4 = reinterpret_cast<const float4*>(lse_s + b2 * 64 + half * 32);
      const float4* d4 = reinterpret_cast<const float4*>(delta_s + b2 * 64 + half * 32);
      const uint64_t scale2 = pack2f(p.scale_log2, p.scale_log2);
      if (__all_sync(0xffffffffu, vis == ~0u)) {      // interior block: no mask tests
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 nl = l4[i], nd = d4[i];
          p_ds_pair<false, true>(sv[4 * i], sv[4 * i + 1], dv[4 * i], dv[4 * i + 1], scale2, pack2f(nl.x, nl.y), pack2f(nd.x, nd.y), vis, 4 * i);
          p_ds_pair<false, true>(sv[4 * i + 2], sv[4 * i + 3], dv[4 * i + 2], dv[4 * i + 3], scale2, pack2f(nl.z, nl.w), pack2f(nd.z, nd.w), vis, 4 * i + 2);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 nl = l4[i], nd = d4[i];
          p_ds_pair<true, true>(sv[4 * i], sv[4 * i + 1], dv[4 * i], dv[4 * i + 1], scale2, pack2f(nl.x, nl.y), pack2f(nd.x, nd.y), vis, 4 * i);
          p_ds_pair<true, true>(sv[4 * i + 2], sv[4 * i + 3], dv[4 * i + 2], dv[4 * i + 3], scale2, pack2f(nl.z, nl.w), pack2f(nd.z, nd.w), vis, 4 * i + 2);
        }
      }
      tmem_st32(a_s, sv);
      tmem_st32(a_d, dv);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(pds_full(b2));
      trace(trs, 302 + it * 3);
      trace(tr && lane == 0, 400 + it * 8 + cw);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    trace(trs, 1);
    {
      // accumulators -> swizzled smem tiles (the operand stages are idle now) -> coalesced row stores
      const float scale = p.scale_log2 * (1.f / kLog2e);
      uint8_t* dv_tile = base_ptr;
      uint8_t* dk_tile = base_ptr + 128 * D * 4;
      uint32_t a[32];
      if (D == 64) {
        tmem_ld32_nowait(t_dv + lane_addr + half * 32, a);
        tmem_ld_wait();
        tile_put_row32<D>(dv_tile, r, half * 32, a, 1.f);
        tmem_ld32_nowait(t_dk + lane_addr + half * 32, a);
        tmem_ld_wait();
        tile_put_row32<D>(dk_tile, r, half * 32, a, scale);
      } else {
        tmem_ld32_nowait((half ? t_dk : t_dv) + lane_addr, a);
        tmem_ld_wait();
        tile_put_row32<D>(half ? dk_tile : dv_tile, r, 0, a, half ? scale : 1.f);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int n_rows = min(128, p.S - k0);
      float* out = dqkv + ((int64_t)b * p.S + k0) * 3 * p.E + head * D;
      tile_store_rows<D>(dv_tile, out + 2 * p.E, 3 * (int64_t)p.E, n_rows, cw, lane);
      tile_store_rows<D>(dk_tile, out + p.E, 3 * (int64_t)p.E, n_rows, cw, lane);
    }
    trace(trs, 2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

// ============================================================================================
// dQ
// ============================================================================================
template <int D, int NS>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_dq_ts(const __grid_constant__ CUtensorMap map_qres,   // qkv, box {32,128}, K-major
                                                            const __grid_constant__ CUtensorMap map_dores,  // dctx, box {32,128}, K-major
                                                            const __grid_constant__ CUtensorMap map_kk,     // qkv, box {32,64}, K-major
                                                            const __grid_constant__ CUtensorMap map_kmn,    // qkv, box {32,64}, MN-major
                                                            const Dims p, const int32_t* __restrict__ x,
                                                            const int32_t* __restrict__ fnp,
                                                            const float* __restrict__ lse,
                                                            const float* __restrict__ delta,
                                                            float* __restrict__ dqkv) {
  using C = Cfg<D, NS>;
  static_assert(NS >= 2, "the prologue parks Q and dO in the first two K-major stages");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t kst = base;                                  // NS stages of [K K-major | V K-major]
  const uint32_t mst = base + NS * C::kStage;                 // NS stages of [K MN-major]
  const uint32_t bar = mst + NS * C::kBlk;
  uint8_t* ctl_ptr = base_ptr + NS * C::kStage + NS * C::kBlk;
  auto kfull = [&](int s) { return bar + 8u * s; };
  auto kfree = [&](int s) { return bar + 8u * (4 + s); };
  auto mfull = [&](int s) { return bar + 8u * (8 + s); };
  auto mfree = [&](int s) { return bar + 8u * (12 + s); };
  auto sdp_full = [&](int b2) { return bar + 8u * (16 + b2); };
  auto ds_full = [&](int b2) { return bar + 8u * (18 + b2); };
  const uint32_t res_full = bar + 8u * 22, res_done = bar + 8u * 23, acc_full = bar + 8u * 24, tmem_slot = bar + 8u * 25;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(ctl_ptr + 8 * 25);
  unsigned long long* padmask_s = reinterpret_cast<unsigned long long*>(ctl_ptr + 256);   // up to 64 key blocks
  constexpr int kMaxPadBlocks = 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y;
  const int q0 = ((int)gridDim.x - 1 - (int)blockIdx.x) * 128;    // heavy (late-query) tiles first
  const int first_real = __ldg(fnp + b);           // first consumed after the causal blocks (see the dK/dV kernel)
  const int n_all = (p.S + 63) / 64;
  const int n_causal = min(n_all, (q0 + 127) / 64 + 1);
  auto more_blocks = [&](int kb) { return kb < n_causal || (q0 < use_here(first_real) && kb < n_all); };
  const int row_base = b * p.S;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(kfull(s), 1); mbar_init(kfree(s), 1 + kSwzWarps); mbar_init(mfull(s), kSwzWarps); mbar_init(mfree(s), 1); }
    for (int b2 = 0; b2 < 2; ++b2) { mbar_init(sdp_full(b2), 1); mbar_init(ds_full(b2), 256); }
    mbar_init(res_full, 1); mbar_init(res_done, 256); mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    mbar_expect_tx(res_full, 2 * C::kStage);
#pragma unroll
    for (int a = 0; a < C::kAtoms; ++a) {
      tma_load_2d(kst + a * (128 * 128), &map_qres, res_full, head * D + a * 32, row_base + q0);
      tma_load_2d(kst + C::kStage + a * (128 * 128), &map_dores, res_full, head * D + a * 32, row_base + q0);
    }
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t t_q = tmem, t_do = tmem + D, t_s = tmem + 2 * D, t_dp = t_s + 128, t_dq = t_dp + 128;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; more_blocks(kb); ++kb) {
        const int s = kb % NS;
        const uint32_t free_par = (uint32_t)((kb / NS) & 1) ^ 1u;
        const int krow = row_base + kb * 64;
        if (kb == 0) mbar_wait(res_done, 0);          // Q and dO have left the first two K-major stages
        mbar_wait(kfree(s), free_par);
        mbar_expect_tx(kfull(s), C::kStage);
#pragma unroll
        for (int a = 0; a < C::kAtoms; ++a) {
          tma_load_2d(kst + s * C::kStage + a * (64 * 128), &map_kk, kfull(s), p.E + head * D + a * 32, krow);
          tma_load_2d(kst + s * C::kStage + C::kBlk + a * (64 * 128), &map_kk, kfull(s), 2 * p.E + head * D + a * 32, krow);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_a = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto issue_sdp = [&](int kb) {        // S = Q K^T ; dP = dO V^T   (A = Q / dO in TMEM)
      const int s = kb % NS, b2 = kb & 1;
      mbar_wait(kfull(s), (uint32_t)((kb / NS) & 1));
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t kk_d = make_desc(kst + s * C::kStage, 16, 1024, 2), vk_d = make_desc(kst + s * C::kStage + C::kBlk, 16, 1024, 2);
#pragma unroll
        for (int kk = 0; kk < D / 8; ++kk) {
          const uint64_t bo = (uint64_t)(((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4);
          umma_tf32_ts(t_s + b2 * 64, t_q + kk * 8, kk_d + bo, idesc_t, kk > 0 ? 1u : 0u);
          umma_tf32_ts(t_dp + b2 * 64, t_do + kk * 8, vk_d + bo, idesc_t, kk > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full(b2));
        umma_commit(kfree(s));
      }
      __syncwarp();
    };
    mbar_wait(res_done, 0);
    tc_fence_after();
    issue_sdp(0);
    if (more_blocks(1)) issue_sdp(1);
    for (int kb = 0; more_blocks(kb); ++kb) {
      const int s = kb % NS, b2 = kb & 1;
      mbar_wait(mfull(s), (uint32_t)((kb / NS) & 1));
      mbar_wait(ds_full(b2), (uint32_t)((kb >> 1) & 1));
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t kmn = make_desc(mst + s * C::kBlk, 64 * 128, 512, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)        // dQ += dS K   (K = 64 keys, A = dS in TMEM)
          umma_tf32_ts(t_dq, t_dp + b2 * 64 + kk * 8, kmn + (uint64_t)(kk * 64), idesc_a, (kb > 0 || kk > 0) ? 1u : 0u);
        umma_commit(mfree(s));
      }
      __syncwarp();
      if (more_blocks(kb + 2)) issue_sdp(kb + 2);
    }
    if (elect_one_sync()) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 10) {
    // re-swizzle warps: MN-major image of the landed K tile (B operand of dQ += dS K)
    const int w2 = warp - 10;
    uint8_t* mst_ptr = base_ptr + NS * C::kStage;
    for (int kb = 0; more_blocks(kb); ++kb) {
      const int s = kb % NS;
      mbar_wait(mfree(s), (uint32_t)((kb / NS) & 1) ^ 1u);
      mbar_wait(kfull(s), (uint32_t)((kb / NS) & 1));
      reswizzle_rows(base_ptr + s * C::kStage, mst_ptr + s * C::kBlk, C::kAtoms * 64, w2, lane);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(mfull(s)); mbar_arrive(kfree(s)); }
    }
  } else {
    const int cw = warp - 2;
    const int q = warp & 3;
    const int half = cw >> 2;
    const int r = q * 32 + lane;
    const int row = q0 + r;
    const bool row_ok = row < p.S;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int64_t so = ((int64_t)b * p.h + head) * p.S;
    const float nls = row_ok ? -lse[so + row] * kLog2e : 0.f;
    const float ndl = row_ok ? -delta[so + row] : 0.f;

    // PAD bits of every key block of this sequence, built once (two ballots per block)
    const bool pads_staged = n_all <= kMaxPadBlocks;
    if (pads_staged) {                                // 32 keys per ballot, all loads of a warp in flight together
      uint32_t* pad32 = reinterpret_cast<uint32_t*>(padmask_s);
      const int n_words = 2 * n_all;
      for (int w0 = cw; w0 < n_words; w0 += 32) {
        int32_t id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int key = (w0 + 8 * u) * 32 + lane;
          id[u] = key < p.S ? __ldg(x + (int64_t)b * p.S + key) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t bits = __ballot_sync(0xffffffffu, id[u] == 0);
          if (lane == 0 && w0 + 8 * u < n_words) pad32[w0 + 8 * u] = bits;
        }
      }
    }
    // prologue: resident Q and dO tiles smem -> TMEM
    mbar_wait(res_full, 0);
    {
      uint32_t v[32];
      if (D == 64) {
        load_swizzled_row(base_ptr + half * (128 * 128), r, v);
        tmem_st32(t_q + lane_addr + half * 32, v);
        load_swizzled_row(base_ptr + C::kStage + half * (128 * 128), r, v);
        tmem_st32(t_do + lane_addr + half * 32, v);
      } else {
        load_swizzled_row(base_ptr + half * C::kStage, r, v);
        tmem_st32((half ? t_do : t_q) + lane_addr, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(res_done);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // padmask_s visible to all softmax warps

    const bool deg = row < first_real;
    for (int kb = 0; more_blocks(kb); ++kb) {
      const int b2 = kb & 1;
      const uint32_t ph = (uint32_t)((kb >> 1) & 1);
      const int kbase = kb * 64 + half * 32;          // first key of this thread's 32 columns
      uint32_t padbits;
      if (pads_staged) {
        padbits = (uint32_t)(padmask_s[kb] >> (half * 32));
      } else {
        const int ka = kbase + lane;
        padbits = __ballot_sync(0xffffffffu, ka < p.S ? (x[(int64_t)b * p.S + ka] == 0) : true);
      }
      const uint32_t range_bits = low_bits32(p.S - kbase);
      uint32_t vis;
      if (!row_ok) vis = 0u;
      else if (deg) vis = range_bits;
      else vis = low_bits32(row - kbase + 1) & ~padbits & range_bits;
      mbar_wait(sdp_full(b2), ph);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      const uint32_t a_s = t_s + b2 * 64 + half * 32 + lane_addr, a_d = t_dp + b2 * 64 + half * 32 + lane_addr;
      tmem_ld32_nowait(a_s, sv);
      tmem_ld32_nowait(a_d, dv);
      tmem_ld_wait();
      const uint64_t scale2 = pack2f(p.scale_log2, p.scale_log2), nl2 = pack2f(nls, nls), ndl2 = pack2f(ndl, ndl);
      if (__all_sync(0xffffffffu, vis == ~0u)) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) p_ds_pair<false, false>(sv[c], sv[c + 1], dv[c], dv[c + 1], scale2, nl2, ndl2, vis, c);
      } else {
#pragma unroll
        for (int c = 0; c < 32; c += 2) p_ds_pair<true, false>(sv[c], sv[c + 1], dv[c], dv[c + 1], scale2, nl2, ndl2, vis, c);
      }
      tmem_st32(a_d, dv);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ds_full(b2));
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const float scale = p.scale_log2 * (1.f / kLog2e);
    uint8_t* dq_tile = base_ptr;                      // the operand stages are idle now
    if (D == 64 || half == 0) {
      uint32_t a[32];
      tmem_ld32_nowait(t_dq + lane_addr + (D == 64 ? half * 32 : 0), a);
      tmem_ld_wait();
      tile_put_row32<D>(dq_tile, r, D == 64 ? half * 32 : 0, a, scale);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    tile_store_rows<D>(dq_tile, dqkv + ((int64_t)b * p.S + q0) * 3 * p.E + head * D, 3 * (int64_t)p.E, min(128, p.S - q0), cw, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}


// ============================================================================================
// Persistent dK / dV: one CTA per SM walks a heaviest-first list of (batch, head, 128-key) tiles.
// Same per-block pipeline as attn_bwd_dkdv_ts; what changes is the tile boundary, which cost ~40 % of a
// 5-block tile there (K/V load -> TMEM copy in front, accumulator drain behind, nothing overlapped):
//   * K and V of tile i+1 are fetched into their own 64 KB buffer while tile i runs and are copied to TMEM
//     BEFORE tile i's accumulators are drained, so the MMA warp forms S^T / dP^T of the next tile's first
//     two query blocks while the threads store dV / dK;
//   * the producer / re-swizzle warps simply keep going: stage and barrier parities run on a global
//     block counter, the per-tile block count travels through a small ring in shared memory;
//   * dV / dK leave as 256-bit row stores straight from registers (full 32-B sectors, no staging tile).
// ============================================================================================
// Draw order of the persistent kernels.  Tiles of one (batch, head) pair share their streamed blocks, so they should
// run at the same time (one HBM read, L2 hits for the rest); heavy tiles should be drawn first so that the tail of the
// kernel is made of light ones.  Compromise: bands of `band` consecutive tile ranks (rank 0 = heaviest); band-major,
// then pair-major, then rank inside the band.
__device__ __forceinline__ void banded_tile(int t, int n_rank, int n_bh, int band, int& bh, int& rank) {
  const int full = (n_rank / band) * band;                    // ranks covered by complete bands
  if (t < full * n_bh) {
    const int per = band * n_bh, bi = t / per, rem = t - bi * per;
    bh = rem / band;
    rank = bi * band + (rem - bh * band);
  } else {
    const int tail = n_rank - full, rem = t - full * n_bh;
    bh = rem / tail;
    rank = full + (rem - bh * tail);
  }
}

template <int D>
struct PsCfg {
  static constexpr int NSK = 3;                              // K-major [Q | dO] stages (filled by TMA, latency-critical)
  static constexpr int NSM = 2;                              // MN-major stages (filled by the re-swizzle warps)
  static constexpr int kAtoms = D / 32;
  static constexpr int kBlk = kAtoms * 64 * 128;
  static constexpr int kStage = 2 * kBlk;                    // [Q | dO] block pair, also one resident 128-row tile
  static constexpr int kKv = 2 * kStage;                     // K and V of the next tile
  static constexpr int kCtl = 2048;
  static constexpr int kSmemDkdv = (NSK + NSM) * kStage + kKv + kCtl + 1024;   // 227 KB for head_dim 64: the SM's limit
  static constexpr int kTmemCols = 512;
};

// 32 floats of row r -> one 128-byte atom of a TMA SWIZZLE_128B tile [128 rows][128 B]
__device__ __forceinline__ void put_row_sw128(uint8_t* atom, int r, const uint32_t (&v)[32], float scale) {
  uint8_t* row = atom + r * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(row + ((j ^ (r & 7)) << 4)) =
        make_float4(__uint_as_float(v[4 * j]) * scale, __uint_as_float(v[4 * j + 1]) * scale,
                    __uint_as_float(v[4 * j + 2]) * scale, __uint_as_float(v[4 * j + 3]) * scale);
}

template <int D>
__global__ void __launch_bounds__(kPsThreads, 1) attn_bwd_dkdv_ps(const __grid_constant__ CUtensorMap map_res,   // qkv, box {32,128}, K-major
                                                              const __grid_constant__ CUtensorMap map_qk,    // qkv, box {32,64}, K-major
                                                              const __grid_constant__ CUtensorMap map_dok,   // dctx, box {32,64}, K-major
                                                              const __grid_constant__ CUtensorMap map_out,   // dqkv as (3E, S, B), box {32,128,1}
                                                              const Dims p, const int n_tiles, int* __restrict__ sched, const int32_t* __restrict__ x,
                                                              const int32_t* __restrict__ fnp,
                                                              const float* __restrict__ lse,
                                                              const float* __restrict__ delta,
                                                              float* __restrict__ dqkv) {
  using C = PsCfg<D>;
  constexpr int NSK = C::NSK, NSM = C::NSM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t kst = base;                                  // NSK stages of [Q K-major | dO K-major]
  uint8_t* mst_ptr = base_ptr + NSK * C::kStage;              // NSM stages of [Q MN-major | dO MN-major]
  const uint32_t mst = base + NSK * C::kStage;
  const uint32_t kvb = base + (NSK + NSM) * C::kStage;        // K | V of the upcoming tile
  uint8_t* kvb_ptr = base_ptr + (NSK + NSM) * C::kStage;
  const uint32_t bar = kvb + C::kKv;
  uint8_t* ctl_ptr = kvb_ptr + C::kKv;
  auto kfull = [&](int s) { return bar + 8u * s; };
  auto kfree = [&](int s) { return bar + 8u * (4 + s); };
  auto mfull = [&](int s) { return bar + 8u * (8 + s); };
  auto mfree = [&](int s) { return bar + 8u * (12 + s); };
  auto sdp_full = [&](int b2) { return bar + 8u * (16 + b2); };
  auto pds_full = [&](int b2) { return bar + 8u * (18 + b2); };
  auto ld_full = [&](int b2) { return bar + 8u * (20 + b2); };
  const uint32_t kv_full = bar + 8u * 22, kv_done = bar + 8u * 23, acc_full = bar + 8u * 24, st_done = bar + 8u * 25, tmem_slot = bar + 8u * 26;
  auto sfree = [&](int b2) { return bar + 8u * (27 + b2); };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(ctl_ptr + 8 * 26);
  float* lse_s = reinterpret_cast<float*>(ctl_ptr + 256);     // [2][64]  -lse*log2e
  float* delta_s = lse_s + 128;                               // [2][64]  -delta
  volatile int* meta_s = reinterpret_cast<volatile int*>(ctl_ptr + 1280);   // ring of 4 x {tile or -1, block count, first_real, -}

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_bh = p.B * p.h;
  const int n_q = (p.S + 63) / 64;
  struct Tile { int b, head, k0; };
  const int n_kt = (p.S + 127) / 128;
  // Tiles are drawn from a global counter (sched[0]) in (batch*head)-major order, heaviest key tile of each pair first:
  // the CTAs running at any moment work on neighbouring (batch, head) pairs, so the Q / dO blocks that the key tiles of
  // one pair share are read from HBM once and found in L2 afterwards, and the dynamic draw balances the SMs.
  const int band = n_kt <= 8 ? 2 : 4;                         // see banded_tile
  auto tile_of = [&](int t) {
    Tile tl;
    int bh, kt;
    banded_tile(t, n_kt, n_bh, band, bh, kt);                 // key tile 0 has the most query blocks: first
    tl.k0 = kt * 128;
    tl.b = bh / p.h;
    tl.head = bh - tl.b * p.h;
    return tl;
  };
  // blocks of a tile: the causal ones [first_causal, n_q) first, then the fully-masked leading blocks [0, n_deg)
  auto n_blocks = [&](const Tile& tl, int first_real) {
    const int first_causal = tl.k0 / 64;
    return (n_q - first_causal) + min(min((first_real + 63) / 64, n_q), first_causal);
  };
  auto block_of = [&](const Tile& tl, int it) {
    const int first_causal = tl.k0 / 64, n_causal = n_q - first_causal;
    return it < n_causal ? first_causal + it : it - n_causal;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSK; ++s) { mbar_init(kfull(s), 1); mbar_init(kfree(s), 1 + kSwzWarps); }
    for (int s = 0; s < NSM; ++s) { mbar_init(mfull(s), kSwzWarps); mbar_init(mfree(s), 1); }
    for (int b2 = 0; b2 < 2; ++b2) { mbar_init(sdp_full(b2), 1); mbar_init(pds_full(b2), 256); mbar_init(ld_full(b2), 32); }
    mbar_init(kv_full, 1); mbar_init(kv_done, 256); mbar_init(acc_full, 1); mbar_init(st_done, 1);
    mbar_init(sfree(0), 1); mbar_init(sfree(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t t_k = tmem, t_v = tmem + D, t_st = tmem + 2 * D, t_dpt = t_st + 128, t_dv = t_dpt + 128, t_dk = t_dv + D;

  if (warp == 0) {
    // ---------------- producer: TMA + lse/delta staging + tile metadata ----------------
    auto issue_kv = [&](const Tile& tl, int idx, int first_real, int t) {   // lane 0 only
      meta_s[(idx & 3) * 4] = t;
      meta_s[(idx & 3) * 4 + 1] = n_blocks(tl, first_real);
      meta_s[(idx & 3) * 4 + 2] = first_real;
      mbar_expect_tx(kv_full, C::kKv);                                   // (release: the ring entry is visible to every waiter)
#pragma unroll
      for (int a = 0; a < C::kAtoms; ++a) {
        tma_load_2d(kvb + a * (128 * 128), &map_res, kv_full, p.E + tl.head * D + a * 32, tl.b * p.S + tl.k0);
        tma_load_2d(kvb + C::kStage + a * (128 * 128), &map_res, kv_full, 2 * p.E + tl.head * D + a * 32, tl.b * p.S + tl.k0);
      }
    };
    float nl_r[2], nd_r[2];
    auto fetch_ld = [&](const Tile& tl, int it) {                        // raw values only (see attn_bwd_dkdv_ts)
      const int64_t so = ((int64_t)tl.b * p.h + tl.head) * p.S;
      const int q0 = block_of(tl, it) * 64;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int qi = min(q0 + lane + 32 * j, p.S - 1);
        nl_r[j] = __ldg(lse + so + qi);
        nd_r[j] = __ldg(delta + so + qi);
      }
    };
    auto draw = [&]() {                                      // next tile index for this CTA (warp-uniform), -1 = none left
      int t = 0;
      if (lane == 0) t = atomicAdd(sched, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      return t < n_tiles ? t : -1;
    };
    auto publish_stop = [&](int idx) {                       // lane 0: ring entry "no more tiles", completes the phase
      meta_s[(idx & 3) * 4] = -1;
      mbar_arrive(kv_full);
    };
    int t = draw();
    Tile tl = tile_of(max(t, 0));
    int first_real = t >= 0 ? __ldg(fnp + tl.b) : 0;
    if (lane == 0) { if (t >= 0) issue_kv(tl, 0, first_real, t); else publish_stop(0); }
    if (t >= 0) fetch_ld(tl, 0);
    int g = 0;
    for (int i = 0; t >= 0; ++i) {
      const int n_it = n_blocks(tl, first_real);
      const int t_next = draw();
      Tile tn = tl;
      int fr_next = 0;
      if (t_next >= 0) { tn = tile_of(t_next); fr_next = __ldg(fnp + tn.b); }
      for (int it = 0; it < n_it; ++it, ++g) {
        const int s = g % NSK;
        if (lane == 0) {
          const int qrow = tl.b * p.S + block_of(tl, it) * 64;
          mbar_wait(kfree(s), (uint32_t)((g / NSK) & 1) ^ 1u);
          mbar_expect_tx(kfull(s), C::kStage);
#pragma unroll
          for (int a = 0; a < C::kAtoms; ++a) {
            tma_load_2d(kst + s * C::kStage + a * (64 * 128), &map_qk, kfull(s), tl.head * D + a * 32, qrow);
            tma_load_2d(kst + s * C::kStage + C::kBlk + a * (64 * 128), &map_dok, kfull(s), tl.head * D + a * 32, qrow);
          }
        }
        __syncwarp();
        const int b2 = g & 1;
        if (g >= 2) mbar_wait(pds_full(b2), (uint32_t)(((g - 2) >> 1) & 1));   // threads are done with this lse/delta stage
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          lse_s[b2 * 64 + lane + 32 * j] = -nl_r[j] * kLog2e;
          delta_s[b2 * 64 + lane + 32 * j] = -nd_r[j];
        }
        mbar_arrive(ld_full(b2));
        if (it + 1 < n_it) fetch_ld(tl, it + 1);
        else if (t_next >= 0) fetch_ld(tn, 0);
        // The next tile is published only AFTER this block's lse/delta stage: st_done(i-1) is signalled by a thread that
        // has finished block 0 of this tile, which needs that stage (a one-block tile would deadlock the other way round).
        if (it == min(1, n_it - 1)) {
          if (lane == 0) {
            mbar_wait(kv_done, (uint32_t)(i & 1));             // this tile's K / V have left the buffer ...
            if (i > 0) mbar_wait(st_done, (uint32_t)((i - 1) & 1));   // ... and so has the previous tile's dV / dK (it doubles as their store staging)
            if (t_next >= 0) issue_kv(tn, i + 1, fr_next, t_next); else publish_stop(i + 1);
          }
          __syncwarp();
        }
      }
      tl = tn;
      first_real = fr_next;
      t = t_next;
    }
    // the CTA that draws the last "none left" puts the counters back for the next launch
    if (lane == 0 && atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }
  } else if (warp == 1) {
    // ---------------- MMA issuer A: S^T = K Q^T ; dP^T = V dO^T (A = K / V in TMEM) ----------------
    // Two issuing warps: with one, the products of block it+2 queued behind the wait for the threads of block it and
    // ~800 clk of every block were lost to that serialisation.  Issue order across the two warps is not program order,
    // so the reuse of an S^T/dP^T buffer is guarded by completion: warp B commits sfree(b2) behind its products.
    constexpr uint32_t idesc_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int g = 0;
    for (int i = 0;; ++i) {
      mbar_wait(kv_full, (uint32_t)(i & 1));                   // (acquire) ring entry of tile i
      if (meta_s[(i & 3) * 4] < 0) break;
      mbar_wait(kv_done, (uint32_t)(i & 1));                   // K and V of this tile are in TMEM
      tc_fence_after();
      const int n_it = meta_s[(i & 3) * 4 + 1];
      for (int it = 0; it < n_it; ++it, ++g) {
        const int s = g % NSK, b2 = g & 1;
        if (g >= 2) mbar_wait(sfree(b2), (uint32_t)(((g >> 1) - 1) & 1));   // dV / dK products of block g-2 have retired
        mbar_wait(kfull(s), (uint32_t)((g / NSK) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t qk = make_desc(kst + s * C::kStage, 16, 1024, 2), dok = make_desc(kst + s * C::kStage + C::kBlk, 16, 1024, 2);
#pragma unroll
          for (int kk = 0; kk < D / 8; ++kk) {
            const uint64_t bo = (uint64_t)(((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4);
            umma_tf32_ts(t_st + b2 * 64, t_k + kk * 8, qk + bo, idesc_t, kk > 0 ? 1u : 0u);
            umma_tf32_ts(t_dpt + b2 * 64, t_v + kk * 8, dok + bo, idesc_t, kk > 0 ? 1u : 0u);
          }
          umma_commit(sdp_full(b2));
          umma_commit(kfree(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 14) {
    // ---------------- MMA issuer B: dV += P^T dO ; dK += dS^T Q (A = P^T / dS^T in TMEM, K = 64 queries) ----------------
    constexpr uint32_t idesc_a = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int g = 0;
    for (int i = 0;; ++i) {
      mbar_wait(kv_full, (uint32_t)(i & 1));
      if (meta_s[(i & 3) * 4] < 0) break;
      const int n_it = meta_s[(i & 3) * 4 + 1];
      for (int it = 0; it < n_it; ++it, ++g) {
        const int s = g % NSM, b2 = g & 1;
        mbar_wait(mfull(s), (uint32_t)((g / NSM) & 1));
        mbar_wait(pds_full(b2), (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t qmn = make_desc(mst + s * C::kStage, 64 * 128, 512, 1), domn = make_desc(mst + s * C::kStage + C::kBlk, 64 * 128, 512, 1);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
            umma_tf32_ts(t_dv, t_st + b2 * 64 + kk * 8, domn + (uint64_t)(kk * 64), idesc_a, acc);
            umma_tf32_ts(t_dk, t_dpt + b2 * 64 + kk * 8, qmn + (uint64_t)(kk * 64), idesc_a, acc);
          }
          umma_commit(mfree(s));
          umma_commit(sfree(b2));
          if (it == n_it - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 10) {
    // ---------------- re-swizzle warps ----------------
    const int w4 = warp - 10;
    int g = 0;
    for (int i = 0;; ++i) {
      mbar_wait(kv_full, (uint32_t)(i & 1));                   // (acquire) this tile's ring entry is visible
      if (meta_s[(i & 3) * 4] < 0) break;
      const int n_it = meta_s[(i & 3) * 4 + 1];
      for (int it = 0; it < n_it; ++it, ++g) {
        const int sk = g % NSK, sm = g % NSM;
        mbar_wait(mfree(sm), (uint32_t)((g / NSM) & 1) ^ 1u);
        mbar_wait(kfull(sk), (uint32_t)((g / NSK) & 1));
        reswizzle_rows(base_ptr + sk * C::kStage, mst_ptr + sm * C::kStage, 2 * C::kAtoms * 64, w4, lane);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(mfull(sm)); mbar_arrive(kfree(sk)); }
      }
    }
  } else {
    // ---------------- softmax threads: (key row r, query half) ----------------
    const int cw = warp - 2;
    const int q = warp & 3;
    const int half = cw >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint64_t scale2 = pack2f(p.scale_log2, p.scale_log2);
    const float scale = p.scale_log2 * (1.f / kLog2e);
    const bool storer = cw == 0 && lane == 0;

    auto stage_kv = [&](int idx) {                             // resident K and V: buffer -> TMEM (32 columns per thread); false = no tile
      mbar_wait(kv_full, (uint32_t)(idx & 1));
      if (meta_s[(idx & 3) * 4] < 0) return false;
      uint32_t v[32];
      if (D == 64) {
        load_swizzled_row(kvb_ptr + half * (128 * 128), r, v);
        tmem_st32(t_k + lane_addr + half * 32, v);
        load_swizzled_row(kvb_ptr + C::kStage + half * (128 * 128), r, v);
        tmem_st32(t_v + lane_addr + half * 32, v);
      } else {
        load_swizzled_row(kvb_ptr + half * C::kStage, r, v);
        tmem_st32((half ? t_v : t_k) + lane_addr, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(kv_done);
      return true;
    };
    auto key_id_of = [&](const Tile& tl) { return (tl.k0 + r < p.S) ? __ldg(x + (int64_t)tl.b * p.S + tl.k0 + r) : 0; };

    const bool trs = p.trace && blockIdx.x == 5 && cw == 0 && lane == 0;
    trace(trs, 0);
    bool have = stage_kv(0);
    Tile tl = tile_of(have ? meta_s[0] : 0);
    int key_id = have ? key_id_of(tl) : 0;
    int g0 = 0;
    for (int i = 0; have; ++i) {
      const int n_it = meta_s[(i & 3) * 4 + 1], first_real = meta_s[(i & 3) * 4 + 2];
      const int key = tl.k0 + r;
      const bool key_ok = key < p.S;
      const bool key_pad = key_id == 0;
      trace(trs && i < 16, 700 + i * 8);
      if (trs && i < 16) g_ts_trace[707 + i * 8] = n_it;
      for (int it = 0; it < n_it; ++it) {
        const int g = g0 + it, b2 = g & 1;
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        const int qb = block_of(tl, it) * 64 + half * 32;     // first query of this thread's 32 columns
        const uint32_t range_bits = low_bits32(p.S - qb);
        const uint32_t causal = (key_ok && !key_pad) ? ~low_bits32(key - qb) : 0u;
        const uint32_t degq = key_ok ? low_bits32(first_real - qb) : 0u;
        const uint32_t vis = (causal | degq) & range_bits;
        mbar_wait(ld_full(b2), ph);
        mbar_wait(sdp_full(b2), ph);
        tc_fence_after();
        if (it == 0) trace(trs && i < 16, 701 + i * 8);
        uint32_t sv[32], dv[32];
        const uint32_t a_s = t_st + b2 * 64 + half * 32 + lane_addr, a_d = t_dpt + b2 * 64 + half * 32 + lane_addr;
        tmem_ld32_nowait(a_s, sv);
        tmem_ld32_nowait(a_d, dv);
        tmem_ld_wait();
        const float4* l4 = reinterpret_cast<const float4*>(lse_s + b2 * 64 + half * 32);
        const float4* d4 = reinterpret_cast<const float4*>(delta_s + b2 * 64 + half * 32);
        if (__all_sync(0xffffffffu, vis == ~0u)) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 nl = l4[c], nd = d4[c];
            p_ds_pair<false, true>(sv[4 * c], sv[4 * c + 1], dv[4 * c], dv[4 * c + 1], scale2, pack2f(nl.x, nl.y), pack2f(nd.x, nd.y), vis, 4 * c);
            p_ds_pair<false, true>(sv[4 * c + 2], sv[4 * c + 3], dv[4 * c + 2], dv[4 * c + 3], scale2, pack2f(nl.z, nl.w), pack2f(nd.z, nd.w), vis, 4 * c + 2);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 nl = l4[c], nd = d4[c];
            p_ds_pair<true, true>(sv[4 * c], sv[4 * c + 1], dv[4 * c], dv[4 * c + 1], scale2, pack2f(nl.x, nl.y), pack2f(nd.x, nd.y), vis, 4 * c);
            p_ds_pair<true, true>(sv[4 * c + 2], sv[4 * c + 3], dv[4 * c + 2], dv[4 * c + 3], scale2, pack2f(nl.z, nl.w), pack2f(nd.z, nd.w), vis, 4 * c + 2);
          }
        }
        tmem_st32(a_s, sv);
        tmem_st32(a_d, dv);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(pds_full(b2));
        if (it == 0 && i > 0 && storer) {                      // the previous tile's bulk store has had a whole block to read its staging
          bulk_wait_read0();
          mbar_arrive(st_done);
        }
      }
      g0 += n_it;
      trace(trs && i < 16, 702 + i * 8);
      // every S^T / dP^T product of this tile has retired (its results were read above): K / V of the next tile
      // go to TMEM now, so the tensor pipe has work while this tile's accumulators drain
      have = stage_kv(i + 1);
      Tile tn = tl;
      int key_id_next = 0;
      if (have) {                                              // requested now, consumed after the drain
        tn = tile_of(meta_s[((i + 1) & 3) * 4]);
        key_id_next = key_id_of(tn);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");          // every thread has left the K/V buffer: it becomes the store staging
      trace(trs && i < 16, 703 + i * 8);
      mbar_wait(acc_full, (uint32_t)(i & 1));
      tc_fence_after();
      trace(trs && i < 16, 704 + i * 8);
      {
        // accumulators -> SWIZZLE_128B tiles in shared memory -> asynchronous TMA stores (rows past the end of the
        // sequence are clipped by the 3-D tensor map); the threads go straight on to the next tile
        uint8_t* dv_tile = kvb_ptr;
        uint8_t* dk_tile = kvb_ptr + C::kStage;
        uint32_t a[32];
        if (D == 64) {
          tmem_ld32_nowait(t_dv + lane_addr + half * 32, a);
          tmem_ld_wait();
          put_row_sw128(dv_tile + half * (128 * 128), r, a, 1.f);
          tmem_ld32_nowait(t_dk + lane_addr + half * 32, a);
          tmem_ld_wait();
          put_row_sw128(dk_tile + half * (128 * 128), r, a, scale);
        } else {
          tmem_ld32_nowait((half ? t_dk : t_dv) + lane_addr, a);
          tmem_ld_wait();
          put_row_sw128(half ? dk_tile : dv_tile, r, a, half ? scale : 1.f);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (storer) {
#pragma unroll
          for (int at = 0; at < C::kAtoms; ++at) {
            tma_store_3d(&map_out, kvb + at * (128 * 128), 2 * p.E + tl.head * D + at * 32, tl.k0, tl.b);
            tma_store_3d(&map_out, kvb + C::kStage + at * (128 * 128), p.E + tl.head * D + at * 32, tl.k0, tl.b);
          }
          bulk_commit();
        }
      }
      tc_fence_before();                                       // the drain is ordered before the next tile's pds_full arrivals
      trace(trs && i < 16, 705 + i * 8);
      tl = tn;
      key_id = key_id_next;
    }
    if (storer) bulk_wait_read0();                             // the last tile's staging must outlive its store
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}


// ============================================================================================
// Persistent dQ: same boundary treatment as attn_bwd_dkdv_ps.  Tile = (batch, head, 128 queries), latest
// (heaviest) query tiles first.  Resident Q / dO of the next tile are prefetched to their own buffer, the
// PAD words and lse / delta of the next tile are requested a whole tile ahead, dQ leaves by TMA store.
// ============================================================================================
template <int D>
struct PsDqCfg {
  static constexpr int NSK = 3;                              // [K | V] K-major stages (TMA)
  static constexpr int NSM = 2;                              // K MN-major stages (re-swizzle warps)
  static constexpr int kAtoms = D / 32;
  static constexpr int kBlk = kAtoms * 64 * 128;
  static constexpr int kStage = 2 * kBlk;                    // [K | V] block pair, also one resident 128-row tile
  static constexpr int kRes = 2 * kStage;                    // Q and dO of the next tile; later the dQ store staging
  static constexpr int kCtl = 2048;                          // barriers 256 B | pad masks 2 x 64 x 8 B | ring
  static constexpr int kSmem = NSK * kStage + NSM * kBlk + kRes + kCtl + 1024;
  static constexpr int kTmemCols = 512;
};

template <int D>
__global__ void __launch_bounds__(kPsThreads, 1) attn_bwd_dq_ps(const __grid_constant__ CUtensorMap map_qres,   // qkv, box {32,128}, K-major
                                                            const __grid_constant__ CUtensorMap map_dores,  // dctx, box {32,128}, K-major
                                                            const __grid_constant__ CUtensorMap map_kk,     // qkv, box {32,64}, K-major
                                                            const __grid_constant__ CUtensorMap map_out,    // dqkv as (3E, S, B), box {32,128,1}
                                                            const Dims p, const int n_tiles, int* __restrict__ sched, const int32_t* __restrict__ x,
                                                            const int32_t* __restrict__ fnp,
                                                            const float* __restrict__ lse,
                                                            const float* __restrict__ delta) {
  using C = PsDqCfg<D>;
  constexpr int NSK = C::NSK, NSM = C::NSM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t kst = base;                                  // NSK stages of [K K-major | V K-major]
  const uint32_t mst = base + NSK * C::kStage;                // NSM stages of [K MN-major]
  uint8_t* mst_ptr = base_ptr + NSK * C::kStage;
  const uint32_t resb = mst + NSM * C::kBlk;                  // Q | dO of the upcoming tile
  uint8_t* res_ptr = mst_ptr + NSM * C::kBlk;
  const uint32_t bar = resb + C::kRes;
  uint8_t* ctl_ptr = res_ptr + C::kRes;
  auto kfull = [&](int s) { return bar + 8u * s; };
  auto kfree = [&](int s) { return bar + 8u * (4 + s); };
  auto mfull = [&](int s) { return bar + 8u * (8 + s); };
  auto mfree = [&](int s) { return bar + 8u * (12 + s); };
  auto sdp_full = [&](int b2) { return bar + 8u * (16 + b2); };
  auto ds_full = [&](int b2) { return bar + 8u * (18 + b2); };
  const uint32_t res_full = bar + 8u * 22, res_done = bar + 8u * 23, acc_full = bar + 8u * 24, st_done = bar + 8u * 25, tmem_slot = bar + 8u * 26;
  auto sfree = [&](int b2) { return bar + 8u * (27 + b2); };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(ctl_ptr + 8 * 26);
  uint32_t* pad32 = reinterpret_cast<uint32_t*>(ctl_ptr + 256);             // [2][128] words: 32 keys each
  volatile int* meta_s = reinterpret_cast<volatile int*>(ctl_ptr + 256 + 1024);   // ring of 4 x {block count, first_real}
  constexpr int kMaxPadWords = 128;                            // S <= 4096 keeps the staged PAD words; longer: per-block ballots

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_bh = p.B * p.h;
  const int n_qt = (p.S + 127) / 128;
  const int n_all = (p.S + 63) / 64;
  const int n_words = 2 * n_all;
  struct Tile { int b, head, q0; };
  const int band = n_qt <= 8 ? 2 : 4;
  auto tile_of = [&](int t) {
    Tile tl;
    int bh, k;
    banded_tile(t, n_qt, n_bh, band, bh, k);                  // drawn from sched[0]: see attn_bwd_dkdv_ps
    tl.q0 = (n_qt - 1 - k) * 128;                             // latest query tiles have the most key blocks: first
    tl.b = bh / p.h;
    tl.head = bh - tl.b * p.h;
    return tl;
  };
  auto n_blocks = [&](const Tile& tl, int first_real) {
    return tl.q0 < first_real ? n_all : min(n_all, (tl.q0 + 127) / 64 + 1);
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSK; ++s) { mbar_init(kfull(s), 1); mbar_init(kfree(s), 1 + kSwzWarps); }
    for (int s = 0; s < NSM; ++s) { mbar_init(mfull(s), kSwzWarps); mbar_init(mfree(s), 1); }
    for (int b2 = 0; b2 < 2; ++b2) { mbar_init(sdp_full(b2), 1); mbar_init(ds_full(b2), 256); }
    mbar_init(res_full, 1); mbar_init(res_done, 256); mbar_init(acc_full, 1); mbar_init(st_done, 1);
    mbar_init(sfree(0), 1); mbar_init(sfree(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t t_q = tmem, t_do = tmem + D, t_s = tmem + 2 * D, t_dp = t_s + 128, t_dq = t_dp + 128;

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane == 0) {
      auto issue_res = [&](const Tile& tl, int idx, int first_real, int t) {
        meta_s[(idx & 3) * 4] = t;
        meta_s[(idx & 3) * 4 + 1] = n_blocks(tl, first_real);
        meta_s[(idx & 3) * 4 + 2] = first_real;
        mbar_expect_tx(res_full, C::kRes);                     // (release: the ring entry is visible to every waiter)
#pragma unroll
        for (int a = 0; a < C::kAtoms; ++a) {
          tma_load_2d(resb + a * (128 * 128), &map_qres, res_full, tl.head * D + a * 32, tl.b * p.S + tl.q0);
          tma_load_2d(resb + C::kStage + a * (128 * 128), &map_dores, res_full, tl.head * D + a * 32, tl.b * p.S + tl.q0);
        }
      };
      auto publish_stop = [&](int idx) {
        meta_s[(idx & 3) * 4] = -1;
        mbar_arrive(res_full);
      };
      auto draw = [&]() { const int t = atomicAdd(sched, 1); return t < n_tiles ? t : -1; };
      int t = draw();
      Tile tl = tile_of(max(t, 0));
      int first_real = t >= 0 ? __ldg(fnp + tl.b) : 0;
      if (t >= 0) issue_res(tl, 0, first_real, t); else publish_stop(0);
      int g = 0;
      for (int i = 0; t >= 0; ++i) {
        const int n_kv = n_blocks(tl, first_real);
        const int t_next = draw();
        Tile tn = tl;
        int fr_next = 0;
        if (t_next >= 0) { tn = tile_of(t_next); fr_next = __ldg(fnp + tn.b); }
        for (int kb = 0; kb < n_kv; ++kb, ++g) {
          const int s = g % NSK;
          const int krow = tl.b * p.S + kb * 64;
          mbar_wait(kfree(s), (uint32_t)((g / NSK) & 1) ^ 1u);
          mbar_expect_tx(kfull(s), C::kStage);
#pragma unroll
          for (int a = 0; a < C::kAtoms; ++a) {
            tma_load_2d(kst + s * C::kStage + a * (64 * 128), &map_kk, kfull(s), p.E + tl.head * D + a * 32, krow);
            tma_load_2d(kst + s * C::kStage + C::kBlk + a * (64 * 128), &map_kk, kfull(s), 2 * p.E + tl.head * D + a * 32, krow);
          }
          if (kb == min(1, n_kv - 1)) {
            mbar_wait(res_done, (uint32_t)(i & 1));            // this tile's Q / dO have left the buffer ...
            if (i > 0) mbar_wait(st_done, (uint32_t)((i - 1) & 1));   // ... and so has the previous tile's dQ staging
            if (t_next >= 0) issue_res(tn, i + 1, fr_next, t_next); else publish_stop(i + 1);
          }
        }
        tl = tn;
        first_real = fr_next;
        t = t_next;
      }
      if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }   // counters ready for the next launch
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer A: S = Q K^T ; dP = dO V^T (A = Q / dO in TMEM); see attn_bwd_dkdv_ps ----------------
    constexpr uint32_t idesc_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int g = 0;
    for (int i = 0;; ++i) {
      mbar_wait(res_full, (uint32_t)(i & 1));                  // (acquire) ring entry of tile i
      if (meta_s[(i & 3) * 4] < 0) break;
      mbar_wait(res_done, (uint32_t)(i & 1));                  // Q and dO of this tile are in TMEM
      tc_fence_after();
      const int n_kv = meta_s[(i & 3) * 4 + 1];
      for (int kb = 0; kb < n_kv; ++kb, ++g) {
        const int s = g % NSK, b2 = g & 1;
        if (g >= 2) mbar_wait(sfree(b2), (uint32_t)(((g >> 1) - 1) & 1));   // the dQ product of block g-2 has retired
        mbar_wait(kfull(s), (uint32_t)((g / NSK) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t kk_d = make_desc(kst + s * C::kStage, 16, 1024, 2), vk_d = make_desc(kst + s * C::kStage + C::kBlk, 16, 1024, 2);
#pragma unroll
          for (int kk = 0; kk < D / 8; ++kk) {
            const uint64_t bo = (uint64_t)(((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4);
            umma_tf32_ts(t_s + b2 * 64, t_q + kk * 8, kk_d + bo, idesc_t, kk > 0 ? 1u : 0u);
            umma_tf32_ts(t_dp + b2 * 64, t_do + kk * 8, vk_d + bo, idesc_t, kk > 0 ? 1u : 0u);
          }
          umma_commit(sdp_full(b2));
          umma_commit(kfree(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 14) {
    // ---------------- MMA issuer B: dQ += dS K (A = dS in TMEM, K = 64 keys) ----------------
    constexpr uint32_t idesc_a = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int g = 0;
    for (int i = 0;; ++i) {
      mbar_wait(res_full, (uint32_t)(i & 1));
      if (meta_s[(i & 3) * 4] < 0) break;
      const int n_kv = meta_s[(i & 3) * 4 + 1];
      for (int kb = 0; kb < n_kv; ++kb, ++g) {
        const int s = g % NSM, b2 = g & 1;
        mbar_wait(mfull(s), (uint32_t)((g / NSM) & 1));
        mbar_wait(ds_full(b2), (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t kmn = make_desc(mst + s * C::kBlk, 64 * 128, 512, 1);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_tf32_ts(t_dq, t_dp + b2 * 64 + kk * 8, kmn + (uint64_t)(kk * 64), idesc_a, (kb > 0 || kk > 0) ? 1u : 0u);
          umma_commit(mfree(s));
          umma_commit(sfree(b2));
          if (kb == n_kv - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 10) {
    // ---------------- re-swizzle warps: MN-major image of the landed K tile ----------------
    const int w4 = warp - 10;
    int g = 0;
    for (int i = 0;; ++i) {
      mbar_wait(res_full, (uint32_t)(i & 1));                  // (acquire) this tile's ring entry is visible
      if (meta_s[(i & 3) * 4] < 0) break;
      const int n_kv = meta_s[(i & 3) * 4 + 1];
      for (int kb = 0; kb < n_kv; ++kb, ++g) {
        const int sk = g % NSK, sm = g % NSM;
        mbar_wait(mfree(sm), (uint32_t)((g / NSM) & 1) ^ 1u);
        mbar_wait(kfull(sk), (uint32_t)((g / NSK) & 1));
        reswizzle_rows(base_ptr + sk * C::kStage, mst_ptr + sm * C::kBlk, C::kAtoms * 64, w4, lane);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(mfull(sm)); mbar_arrive(kfree(sk)); }
      }
    }
  } else {
    // ---------------- softmax threads: (query row r, key half) ----------------
    const int cw = warp - 2;
    const int q = warp & 3;
    const int half = cw >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint64_t scale2 = pack2f(p.scale_log2, p.scale_log2);
    const float scale = p.scale_log2 * (1.f / kLog2e);
    const bool storer = cw == 0 && lane == 0;
    const bool pads_staged = n_words <= kMaxPadWords;

    auto stage_res = [&](int idx) {                            // resident Q and dO: buffer -> TMEM (32 columns per thread); false = no tile
      mbar_wait(res_full, (uint32_t)(idx & 1));
      if (meta_s[(idx & 3) * 4] < 0) return false;
      uint32_t v[32];
      if (D == 64) {
        load_swizzled_row(res_ptr + half * (128 * 128), r, v);
        tmem_st32(t_q + lane_addr + half * 32, v);
        load_swizzled_row(res_ptr + C::kStage + half * (128 * 128), r, v);
        tmem_st32(t_do + lane_addr + half * 32, v);
      } else {
        load_swizzled_row(res_ptr + half * C::kStage, r, v);
        tmem_st32((half ? t_do : t_q) + lane_addr, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(res_done);
      return true;
    };
    // PAD words of a sequence: warp cw owns words cw, cw+8, ... ; up to 4 of them ride in registers a whole tile ahead
    int32_t pad_id[4];
    auto fetch_pads = [&](int b) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int key = (cw + 8 * u) * 32 + lane;
        pad_id[u] = key < p.S ? __ldg(x + (int64_t)b * p.S + key) : 0;
      }
    };
    auto store_pads = [&](int b, int idx) {
      uint32_t* dst = pad32 + (idx & 1) * kMaxPadWords;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t bits = __ballot_sync(0xffffffffu, pad_id[u] == 0);
        if (lane == 0 && cw + 8 * u < n_words) dst[cw + 8 * u] = bits;
      }
      for (int w = cw + 32; w < n_words; w += 8) {            // sequences longer than 1024: the rest, synchronously
        const int key = w * 32 + lane;
        const uint32_t bits = __ballot_sync(0xffffffffu, key < p.S ? (__ldg(x + (int64_t)b * p.S + key) == 0) : true);
        if (lane == 0) dst[w] = bits;
      }
    };

    float lse_r = 0.f, dl_r = 0.f;
    auto fetch_row = [&](const Tile& tl) {                     // this thread's lse / delta of a tile
      const int row = tl.q0 + r;
      const int64_t so = ((int64_t)tl.b * p.h + tl.head) * p.S;
      lse_r = row < p.S ? __ldg(lse + so + row) : 0.f;
      dl_r = row < p.S ? __ldg(delta + so + row) : 0.f;
    };
    bool have = stage_res(0);
    Tile tl = tile_of(have ? meta_s[0] : 0);
    if (have) {
      fetch_row(tl);
      if (pads_staged) { fetch_pads(tl.b); store_pads(tl.b, 0); }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");            // PAD words visible
    int g0 = 0;
    for (int i = 0; have; ++i) {
      const int n_kv = meta_s[(i & 3) * 4 + 1], first_real = meta_s[(i & 3) * 4 + 2];
      const int row = tl.q0 + r;
      const bool row_ok = row < p.S;
      const bool deg = row < first_real;
      const float nls = -lse_r * kLog2e, ndl = -dl_r;
      const uint64_t nl2 = pack2f(nls, nls), ndl2 = pack2f(ndl, ndl);
      const uint32_t* pads = pad32 + (i & 1) * kMaxPadWords;
      for (int kb = 0; kb < n_kv; ++kb) {
        const int g = g0 + kb, b2 = g & 1;
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        const int kbase = kb * 64 + half * 32;                // first key of this thread's 32 columns
        uint32_t padbits;
        if (pads_staged) {
          padbits = pads[2 * kb + half];
        } else {
          const int ka = kbase + lane;
          padbits = __ballot_sync(0xffffffffu, ka < p.S ? (x[(int64_t)tl.b * p.S + ka] == 0) : true);
        }
        const uint32_t range_bits = low_bits32(p.S - kbase);
        uint32_t vis;
        if (!row_ok) vis = 0u;
        else if (deg) vis = range_bits;
        else vis = low_bits32(row - kbase + 1) & ~padbits & range_bits;
        mbar_wait(sdp_full(b2), ph);
        tc_fence_after();
        uint32_t sv[32], dv[32];
        const uint32_t a_s = t_s + b2 * 64 + half * 32 + lane_addr, a_d = t_dp + b2 * 64 + half * 32 + lane_addr;
        tmem_ld32_nowait(a_s, sv);
        tmem_ld32_nowait(a_d, dv);
        tmem_ld_wait();
        if (__all_sync(0xffffffffu, vis == ~0u)) {
#pragma unroll
          for (int c = 0; c < 32; c += 2) p_ds_pair<false, false>(sv[c], sv[c + 1], dv[c], dv[c + 1], scale2, nl2, ndl2, vis, c);
        } else {
#pragma unroll
          for (int c = 0; c < 32; c += 2) p_ds_pair<true, false>(sv[c], sv[c + 1], dv[c], dv[c + 1], scale2, nl2, ndl2, vis, c);
        }
        tmem_st32(a_d, dv);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(ds_full(b2));
        if (kb == 0 && i > 0 && storer) {                      // the previous tile's bulk store has had a whole block to read its staging
          bulk_wait_read0();
          mbar_arrive(st_done);
        }
      }
      g0 += n_kv;
      // every S / dP product of this tile has retired: the next tile's Q / dO go to TMEM before the drain
      have = stage_res(i + 1);
      Tile tn = tl;
      if (have) {                                              // the next tile's per-row values and PAD words: requested now, used after the drain
        tn = tile_of(meta_s[((i + 1) & 3) * 4]);
        fetch_row(tn);
        if (pads_staged) fetch_pads(tn.b);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");          // everyone has left the Q/dO buffer: it becomes the store staging
      mbar_wait(acc_full, (uint32_t)(i & 1));
      tc_fence_after();
      if (D == 64 || half == 0) {
        uint32_t a[32];
        tmem_ld32_nowait(t_dq + lane_addr + (D == 64 ? half * 32 : 0), a);
        tmem_ld_wait();
        put_row_sw128(res_ptr + (D == 64 ? half : 0) * (128 * 128), r, a, scale);
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (storer) {
#pragma unroll
        for (int at = 0; at < C::kAtoms; ++at)
          tma_store_3d(&map_out, resb + at * (128 * 128), tl.head * D + at * 32, tl.q0, tl.b);
        bulk_commit();
      }
      tc_fence_before();                                       // the drain is ordered before the next tile's ds_full arrivals
      if (have && pads_staged) store_pads(tn.b, i + 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");          // PAD words of the next tile visible
      tl = tn;
    }
    if (storer) bulk_wait_read0();                             // the last tile's staging must outlive its store
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

// two ints per (device, stream) and kernel: next tile, CTAs finished -- zero between launches (the kernels restore them)
int sched_counters(cudaStream_t st, int which, int** out) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, int*> table;
  int dev = 0;
  NLPS_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int*& ptr = table[{dev, st}];
  if (ptr == nullptr) {
    NLPS_CUDA(cudaMalloc(reinterpret_cast<void**>(&ptr), 64));
    NLPS_CUDA(cudaMemset(ptr, 0, 64));
  }
  *out = ptr + 2 * which;
  return 0;
}

template <int D>
int launch_bwd(cudaStream_t st, int B, int S, int h, const float* qkv, const int32_t* x, const int32_t* fnp,
               const float* dctx, const float* lse, const float* delta, float* dqkv) {
  constexpr int NS = 3;
  using C = Cfg<D, NS>;
  const int E = h * D;
  const uint64_t rows = (uint64_t)B * S;
  CUtensorMap res, qk, qmn, dok, domn, dores;
  NLPS_TRY(make_map(&res, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, 128, true, false));
  NLPS_TRY(make_map(&qk, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, 64, true, false));
  NLPS_TRY(make_map(&qmn, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, 64, true, true));
  NLPS_TRY(make_map(&dok, dctx, (uint64_t)E, rows, (int64_t)E, 32, 64, true, false));
  NLPS_TRY(make_map(&domn, dctx, (uint64_t)E, rows, (int64_t)E, 32, 64, true, true));
  NLPS_TRY(make_map(&dores, dctx, (uint64_t)E, rows, (int64_t)E, 32, 128, true, false));
  static bool once = false;
  if (!once) {
    NLPS_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_ts<D, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemDkdv));
    NLPS_CUDA(cudaFuncSetAttribute(attn_bwd_dq_ts<D, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemDq));
    once = true;
  }
  static int trace_on = -1;
  if (trace_on < 0) { const char* e = getenv("NLPS_ATTN_TRACE"); trace_on = (e && e[0] == '2') ? 1 : 0; }
  Dims p{B, S, h, E, (float)(1.0 / std::sqrt((double)D)) * kLog2e, trace_on};
  dim3 grid((unsigned)ceil_div(S, 128), (unsigned)h, (unsigned)B);
  static int persistent = -1;
  if (persistent < 0) { const char* e = getenv("NLPS_ATTN_PS"); persistent = (e && e[0] == '0') ? 0 : 1; }
  const int64_t n_tiles = ceil_div(S, 128) * (int64_t)B * h;
  if (persistent && (reinterpret_cast<uintptr_t>(dqkv) & 31u) == 0 && n_tiles < (1ll << 30)) {
    static bool once_ps = false;
    if (!once_ps) {
      NLPS_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_ps<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, PsCfg<D>::kSmemDkdv));
      once_ps = true;
    }
    CUtensorMap out3;
    NLPS_TRY(make_map3(&out3, dqkv, 3 * (uint64_t)E, (uint64_t)S, (uint64_t)B, 3 * (int64_t)E, (int64_t)S * 3 * E, 32, 128));
    int* sched = nullptr;
    NLPS_TRY(sched_counters(st, 0, &sched));
    attn_bwd_dkdv_ps<D><<<(unsigned)std::min<int64_t>(n_tiles, kNumSMs), kPsThreads, PsCfg<D>::kSmemDkdv, st>>>(
        res, qk, dok, out3, p, (int)n_tiles, sched, x, fnp, lse, delta, dqkv);
    if (trace_on) {
      static int printed_ps = 0;
      if (printed_ps++ == 3) {
        static long long hb[kTraceSlots];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(hb, g_ts_trace, sizeof(hb));
        const long long t0 = hb[0];
        printf("attn_bwd_dkdv_ps trace, CTA 5 (cycles since the threads started)\n");
        for (int i = 0; i < 14; ++i)
          printf("tile %2d (%lld blocks) | loop top %7lld first S^T ready %7lld loop done %7lld next K/V in TMEM %7lld acc_full %7lld drained %7lld\n", i,
                 hb[707 + i * 8], hb[700 + i * 8] - t0, hb[701 + i * 8] - t0, hb[702 + i * 8] - t0, hb[703 + i * 8] - t0, hb[704 + i * 8] - t0, hb[705 + i * 8] - t0);
        fflush(stdout);
      }
    }
  } else {
    attn_bwd_dkdv_ts<D, NS><<<grid, kThreads, C::kSmemDkdv, st>>>(res, qk, qmn, dok, domn, p, x, fnp, lse, delta, dqkv);
  }
  NLPS_LAUNCH_CHECK();
  if (persistent && (reinterpret_cast<uintptr_t>(dqkv) & 31u) == 0 && n_tiles < (1ll << 30)) {
    static bool once_dq = false;
    if (!once_dq) {
      NLPS_CUDA(cudaFuncSetAttribute(attn_bwd_dq_ps<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, PsDqCfg<D>::kSmem));
      once_dq = true;
    }
    CUtensorMap out3;
    NLPS_TRY(make_map3(&out3, dqkv, 3 * (uint64_t)E, (uint64_t)S, (uint64_t)B, 3 * (int64_t)E, (int64_t)S * 3 * E, 32, 128));
    int* sched = nullptr;
    NLPS_TRY(sched_counters(st, 1, &sched));
    attn_bwd_dq_ps<D><<<(unsigned)std::min<int64_t>(n_tiles, kNumSMs), kPsThreads, PsDqCfg<D>::kSmem, st>>>(
        res, dores, qk, out3, p, (int)n_tiles, sched, x, fnp, lse, delta);
  } else {
    attn_bwd_dq_ts<D, NS><<<grid, kThreads, C::kSmemDq, st>>>(res, dores, qk, qmn, p, x, fnp, lse, delta, dqkv);
  }
  NLPS_LAUNCH_CHECK();
  if (trace_on) {
    static int printed = 0;
    if (printed++ == 3) {
      static long long hbuf[kTraceSlots];
      cudaStreamSynchronize(st);
      cudaMemcpyFromSymbol(hbuf, g_ts_trace, sizeof(hbuf));
      const long long t0 = hbuf[0];
      printf("attn_bwd_dkdv_ts trace, key tile 0 (cycles since CTA start): accumulators ready %lld, stores done %lld\n", hbuf[1] - t0, hbuf[2] - t0);
      const int n_it = std::min(8, (S + 63) / 64);
      for (int it = 0; it < n_it; ++it)
        printf("it %d | TMA issue %6lld | MMA sdp start %6lld issued %6lld acc start %6lld issued %6lld | thread top %6lld sdp ready %6lld P/dS stored %6lld\n",
               it, hbuf[100 + it * 2] - t0, hbuf[200 + it * 4] - t0, hbuf[201 + it * 4] - t0,
               hbuf[202 + it * 4] - t0, hbuf[203 + it * 4] - t0, hbuf[300 + it * 3] - t0, hbuf[301 + it * 3] - t0, hbuf[302 + it * 3] - t0);
      for (int it = 0; it < n_it; ++it)
        printf("it %d | re-swizzle warp: top %6lld mfree %6lld kfull %6lld done %6lld\n", it, hbuf[600 + it * 4] - t0, hbuf[601 + it * 4] - t0,
               hbuf[602 + it * 4] - t0, hbuf[603 + it * 4] - t0);
      for (int it = 0; it < n_it; ++it) {
        printf("it %d | mfull passed %6lld | per-warp P/dS stored:", it, hbuf[500 + it] - t0);
        for (int w = 0; w < 8; ++w) printf(" %6lld", hbuf[400 + it * 8 + w] - t0);
        printf("\n");
      }
      fflush(stdout);
    }
  }
  return 0;
}


// ============================================================================================
// Forward (transformer.py:116-141).  One CTA per (batch, head, 128 queries), TWO CTAs per SM (256 TMEM
// columns, ~99 KB of shared memory each): while one CTA is in its prologue / epilogue or waits for the
// tensor pipe, the other one's softmax warps run.
//   Q is copied smem -> TMEM once and is the A operand of S = Q K^T; thread r of the 4 softmax warps owns
//   query row r (= TMEM lane r): it reads its 64 scores, does the online max/sum with no cross-thread
//   traffic, overwrites S with P in TMEM (the A operand of O_blk = P V), and folds O_blk(j-1) into its
//   register accumulator only AFTER P(j) is handed over -- S is double-buffered, so the thread never
//   waits for the P V product it has just enabled.
// ============================================================================================
constexpr int kFwdThreads = 192;   // warp0 TMA, warp1 MMA + TMEM owner, warps 2-5 softmax

template <int D, int NS>
struct FwdCfg {
  static constexpr int kAtoms = D / 32;
  static constexpr int kBlk = kAtoms * 64 * 128;             // K (K-major) or V (MN-major) block of 64 keys
  static constexpr int kStage = 2 * kBlk;
  static constexpr int kTile = 128 * D * 4;                  // Q image, later the normalised output tile
  static constexpr int kCtl = 1024;                          // barriers (256 B) | pad masks 64 x 8 B
  static constexpr int kSmem = NS * kStage + kTile + kCtl + 1024;
  static constexpr int kTmemCols = 256;
};

template <int D, int NS>
__global__ void __launch_bounds__(kFwdThreads, 2) attn_fwd_ts(const __grid_constant__ CUtensorMap map_q,    // qkv, box {32,128}, K-major
                                                            const __grid_constant__ CUtensorMap map_k,    // qkv, box {32,64}, K-major
                                                            const __grid_constant__ CUtensorMap map_v,    // qkv, box {32,64}, MN-major
                                                            const Dims p, const int32_t* __restrict__ x,
                                                            const int32_t* __restrict__ fnp, float* __restrict__ ctx,
                                                            float* __restrict__ lse) {
  using C = FwdCfg<D, NS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t kv_s = base;                                 // NS stages of [K block | V block]
  const uint32_t q_s = base + NS * C::kStage;
  uint8_t* q_ptr = base_ptr + NS * C::kStage;
  uint8_t* ctl_ptr = q_ptr + C::kTile;
  const uint32_t bar = q_s + C::kTile;
  auto kfull = [&](int s) { return bar + 8u * s; };
  auto kfree = [&](int s) { return bar + 8u * (4 + s); };
  auto vfull = [&](int s) { return bar + 8u * (8 + s); };
  auto vfree = [&](int s) { return bar + 8u * (12 + s); };
  auto s_full = [&](int b2) { return bar + 8u * (16 + b2); };
  auto p_full = [&](int b2) { return bar + 8u * (18 + b2); };
  const uint32_t o_full = bar + 8u * 20, o_read = bar + 8u * 21, q_full = bar + 8u * 22, q_done = bar + 8u * 23, tmem_slot = bar + 8u * 24;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(ctl_ptr + 8 * 24);
  unsigned long long* padmask_s = reinterpret_cast<unsigned long long*>(ctl_ptr + 256);   // [64]
  constexpr int kMaxPadBlocks = 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y;
  const int q0 = ((int)gridDim.x - 1 - (int)blockIdx.x) * 128;     // heavy (late-query) tiles are scheduled first
  // Global loads take thousands of cycles while the TMA traffic of 296 CTAs saturates the L2: first_real is
  // requested here and first consumed after the causal blocks (a fully-masked leading tile appends the rest).
  const int first_real = __ldg(fnp + b);
  const int n_all = (p.S + 63) / 64;
  const int n_causal = min(n_all, (q0 + 127) / 64 + 1);
  auto more_blocks = [&](int kb) { return kb < n_causal || (q0 < use_here(first_real) && kb < n_all); };
  const int row_base = b * p.S;
  const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == gridDim.z / 2;   // a mid-grid CTA (warm caches)
  trace(tr && threadIdx.x == 0, 0);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(kfull(s), 1); mbar_init(kfree(s), 1); mbar_init(vfull(s), 1); mbar_init(vfree(s), 1); }
    for (int b2 = 0; b2 < 2; ++b2) { mbar_init(s_full(b2), 1); mbar_init(p_full(b2), 128); }
    mbar_init(o_full, 1); mbar_init(o_read, 128); mbar_init(q_full, 1); mbar_init(q_done, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    trace(tr, 9);
    mbar_expect_tx(q_full, C::kTile);                         // Q starts moving before the TMEM allocation / setup barrier
#pragma unroll
    for (int a = 0; a < C::kAtoms; ++a)
      tma_load_2d(q_s + a * (128 * 128), &map_q, q_full, head * D + a * 32, row_base + q0);
    trace(tr, 10);
  }
  if (warp == 1) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    }
    tmem_alloc(tmem_slot, C::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  trace(tr && threadIdx.x == 64, 4);
  const uint32_t t_q = tmem, t_o = tmem + D + 128;
  auto t_s = [&](int b2) { return tmem + D + 64 * b2; };

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane == 0) {
      for (int kb = 0; more_blocks(kb); ++kb) {
        const int s = kb % NS;
        const uint32_t free_par = (uint32_t)((kb / NS) & 1) ^ 1u;
        const int krow = row_base + kb * 64;
        mbar_wait(kfree(s), free_par);
        trace(tr && kb < 16, 100 + kb * 2);
        mbar_expect_tx(kfull(s), C::kBlk);
#pragma unroll
        for (int a = 0; a < C::kAtoms; ++a)
          tma_load_2d(kv_s + s * C::kStage + a * (64 * 128), &map_k, kfull(s), p.E + head * D + a * 32, krow);
        mbar_wait(vfree(s), free_par);
        trace(tr && kb < 16, 101 + kb * 2);
        mbar_expect_tx(vfull(s), C::kBlk);
#pragma unroll
        for (int a = 0; a < C::kAtoms; ++a)
          tma_load_2d(kv_s + s * C::kStage + C::kBlk + a * (64 * 128), &map_v, vfull(s), 2 * p.E + head * D + a * 32, krow);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer (converged warp, elected lane) ----------------
    constexpr uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto issue_s = [&](int j) {            // S = Q K^T (A = Q in TMEM)
      const int s = j % NS, b2 = j & 1;
      mbar_wait(kfull(s), (uint32_t)((j / NS) & 1));
      tc_fence_after();
      trace(tr && lane == 0 && j < 16, 200 + j * 4);
      if (elect_one_sync()) {
        const uint64_t kd = make_desc(kv_s + s * C::kStage, 16, 1024, 2);
#pragma unroll
        for (int kk = 0; kk < D / 8; ++kk)
          umma_tf32_ts(t_s(b2), t_q + kk * 8, kd + (uint64_t)(((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(s_full(b2));
        umma_commit(kfree(s));
      }
      __syncwarp();
      trace(tr && lane == 0 && j < 16, 201 + j * 4);
    };
    mbar_wait(q_done, 0);
    tc_fence_after();
    issue_s(0);
    if (more_blocks(1)) issue_s(1);
    for (int j = 0; more_blocks(j); ++j) {
      const int s = j % NS, b2 = j & 1;
      mbar_wait(vfull(s), (uint32_t)((j / NS) & 1));
      mbar_wait(p_full(b2), (uint32_t)((j >> 1) & 1));
      if (j > 0) mbar_wait(o_read, (uint32_t)((j - 1) & 1));       // the previous O block has been folded into registers
      tc_fence_after();
      trace(tr && lane == 0 && j < 16, 202 + j * 4);
      if (elect_one_sync()) {
        const uint64_t vd = make_desc(kv_s + s * C::kStage + C::kBlk, 64 * 128, 512, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)          // O_blk = P V (A = P in TMEM, K = 64 keys)
          umma_tf32_ts(t_o, t_s(b2) + kk * 8, vd + (uint64_t)(kk * 64), idesc_o, kk > 0 ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(vfree(s));
      }
      __syncwarp();
      trace(tr && lane == 0 && j < 16, 203 + j * 4);
      if (more_blocks(j + 2)) issue_s(j + 2);    // overwrites the P just consumed: same-thread MMAs retire in order
    }
  } else {
    // ---------------- softmax warps: thread == query row ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int row = q0 + r;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint64_t scale2 = pack2f(p.scale_log2, p.scale_log2);
    const bool pads_staged = n_all <= kMaxPadBlocks;
    const bool trs = tr && q == 2 && lane == 0;
    // PAD bits of every key of this sequence, built once: 32 keys per ballot, all loads of a warp in flight together
    if (pads_staged) {
      uint32_t* pad32 = reinterpret_cast<uint32_t*>(padmask_s);
      const int n_words = 2 * n_all;                     // words [2kb, 2kb+1] = keys of block kb
      for (int w0 = warp - 2; w0 < n_words; w0 += 16) {
        int32_t id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int key = (w0 + 4 * u) * 32 + lane;
          id[u] = key < p.S ? __ldg(x + (int64_t)b * p.S + key) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t bits = __ballot_sync(0xffffffffu, id[u] == 0);
          if (lane == 0 && w0 + 4 * u < n_words) pad32[w0 + 4 * u] = bits;
        }
      }
    }
    // Q image -> TMEM (this thread's row)
    trace(trs, 5);
    mbar_wait(q_full, 0);
    trace(trs, 6);
#pragma unroll
    for (int a = 0; a < C::kAtoms; ++a) {
      uint32_t v[32];
      load_swizzled_row(q_ptr + a * (128 * 128), r, v);
      tmem_st32(t_q + lane_addr + a * 32, v);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(q_done);
    asm volatile("bar.sync 1, 128;" ::: "memory");          // pad masks visible; every row of the Q image has been read
    trace(trs, 1);

    float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
    uint32_t o[D];
#pragma unroll
    for (int c = 0; c < D; ++c) o[c] = 0u;
    auto fold = [&](int j, float alpha) {                    // o = o * alpha + O_blk(j)
      const uint64_t alpha2 = pack2f(alpha, alpha);
      mbar_wait(o_full, (uint32_t)(j & 1));
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc) {
        uint32_t ov[32];
        tmem_ld32_nowait(t_o + lane_addr + cc * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 2)
          unpack2(ffma2(pack2(o[cc * 32 + c], o[cc * 32 + c + 1]), alpha2, pack2(ov[c], ov[c + 1])), o[cc * 32 + c], o[cc * 32 + c + 1]);
      }
      tc_fence_before();
      mbar_arrive(o_read);
    };
    const bool deg = row < first_real;
    int n_kv = 0;
    for (int j = 0; more_blocks(j); ++j) {
      n_kv = j + 1;
      const int b2 = j & 1;
      const int k0 = j * 64;
      unsigned long long padbits;
      if (pads_staged) {
        padbits = padmask_s[j];
      } else {
        const int ka = k0 + lane, kb2 = k0 + 32 + lane;
        const bool pa = ka < p.S ? (x[(int64_t)b * p.S + ka] == 0) : true;
        const bool pb = kb2 < p.S ? (x[(int64_t)b * p.S + kb2] == 0) : true;
        padbits = (unsigned long long)__ballot_sync(0xffffffffu, pa) | ((unsigned long long)__ballot_sync(0xffffffffu, pb) << 32);
      }
      const int in_range = min(64, p.S - k0);
      const unsigned long long range_bits = in_range >= 64 ? ~0ull : ((1ull << in_range) - 1ull);
      unsigned long long vis;
      if (row >= p.S) vis = 0ull;
      else if (deg) vis = range_bits;
      else {
        const int upto = row - k0;
        const unsigned long long causal = upto < 0 ? 0ull : (upto >= 63 ? ~0ull : ((1ull << (upto + 1)) - 1ull));
        vis = causal & ~padbits & range_bits;
      }
      trace(trs && j < 16, 300 + j * 4);
      mbar_wait(s_full(b2), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      trace(trs && j < 16, 301 + j * 4);
      uint32_t s0[32], s1[32];
      tmem_ld32_nowait(t_s(b2) + lane_addr, s0);
      tmem_ld32_nowait(t_s(b2) + lane_addr + 32, s1);
      tmem_ld_wait();
      const bool all_vis = __all_sync(0xffffffffu, vis == ~0ull);
      float mx = -INFINITY;
      if (all_vis) {
#pragma unroll
        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[c]), __uint_as_float(s1[c])));
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (vis & (1ull << c)) mx = fmaxf(mx, __uint_as_float(s0[c]));
          if (vis & (1ull << (c + 32))) mx = fmaxf(mx, __uint_as_float(s1[c]));
        }
      }
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;      // row with no visible key yet
      const float alpha = fast_exp2(m - m_use);                   // exp2(-inf) = 0 when m was -inf
      const uint64_t negm2 = pack2f(-m_use, -m_use);
      uint64_t rsum2 = 0ull;
      const uint32_t vlo = (uint32_t)vis, vhi = (uint32_t)(vis >> 32);
      if (all_vis) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          uint32_t xa, xb, ya, yb;
          unpack2(ffma2(pack2(s0[c], s0[c + 1]), scale2, negm2), xa, xb);
          unpack2(ffma2(pack2(s1[c], s1[c + 1]), scale2, negm2), ya, yb);
          const float e0 = fast_exp2(__uint_as_float(xa)), e1 = fast_exp2(__uint_as_float(xb));
          const float f0 = fast_exp2(__uint_as_float(ya)), f1 = fast_exp2(__uint_as_float(yb));
          rsum2 = fadd2(rsum2, fadd2(pack2f(e0, e1), pack2f(f0, f1)));
          s0[c] = rna_tf32_bits(__float_as_uint(e0)); s0[c + 1] = rna_tf32_bits(__float_as_uint(e1));
          s1[c] = rna_tf32_bits(__float_as_uint(f0)); s1[c + 1] = rna_tf32_bits(__float_as_uint(f1));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          uint32_t xa, xb, ya, yb;
          unpack2(ffma2(pack2(s0[c], s0[c + 1]), scale2, negm2), xa, xb);
          unpack2(ffma2(pack2(s1[c], s1[c + 1]), scale2, negm2), ya, yb);
          const float e0 = (vlo & (1u << c)) ? fast_exp2(__uint_as_float(xa)) : 0.f, e1 = (vlo & (2u << c)) ? fast_exp2(__uint_as_float(xb)) : 0.f;
          const float f0 = (vhi & (1u << c)) ? fast_exp2(__uint_as_float(ya)) : 0.f, f1 = (vhi & (2u << c)) ? fast_exp2(__uint_as_float(yb)) : 0.f;
          rsum2 = fadd2(rsum2, fadd2(pack2f(e0, e1), pack2f(f0, f1)));
          s0[c] = rna_tf32_bits(__float_as_uint(e0)); s0[c + 1] = rna_tf32_bits(__float_as_uint(e1));
          s1[c] = rna_tf32_bits(__float_as_uint(f0)); s1[c + 1] = rna_tf32_bits(__float_as_uint(f1));
        }
      }
      tmem_st32(t_s(b2) + lane_addr, s0);
      tmem_st32(t_s(b2) + lane_addr + 32, s1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(b2));
      trace(trs && j < 16, 302 + j * 4);
      uint32_t ra, rb;
      unpack2(rsum2, ra, rb);
      l = l * alpha + (__uint_as_float(ra) + __uint_as_float(rb));
      m = m_new;
      if (j > 0) fold(j - 1, alpha_prev);        // the product this thread enabled one block ago has long retired
      alpha_prev = alpha;
      trace(trs && j < 16, 303 + j * 4);
    }
    fold(n_kv - 1, alpha_prev);
    trace(trs, 2);

    // normalise -> swizzled smem tile (the Q image is dead) -> coalesced row stores
    {
      const float inv = 1.f / l;
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc) {
        uint32_t part[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) part[c] = o[cc * 32 + c];
        tile_put_row32<D>(q_ptr, r, cc * 32, part, inv);
      }
      // natural-log LSE of the scaled scores (m, l are in the exp2 domain)
      if (row < p.S) lse[((int64_t)b * p.h + head) * p.S + row] = (m + log2f(l)) * (1.f / kLog2e);
      trace(trs, 7);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      trace(trs, 8);
      tile_store_rows<D, 4>(q_ptr, ctx + ((int64_t)b * p.S + q0) * p.E + head * D, (int64_t)p.E, min(128, p.S - q0), warp - 2, lane);
    }
    trace(trs, 3);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}


// ============================================================================================
// Persistent forward: attn_fwd_ts's block pipeline, two CTAs per SM, each walking tiles drawn from the global
// counter in the banded order of the backward kernels.  The Q image of the next tile is prefetched into the Q
// buffer while the current tile runs (Q itself lives in TMEM), copied to TMEM before the current output is
// drained, and the buffer then stages the normalised O tile for an asynchronous TMA store.
// ============================================================================================
template <int D, int NS>
__global__ void __launch_bounds__(kFwdThreads, 2) attn_fwd_ps(const __grid_constant__ CUtensorMap map_q,    // qkv, box {32,128}, K-major
                                                            const __grid_constant__ CUtensorMap map_k,    // qkv, box {32,64}, K-major
                                                            const __grid_constant__ CUtensorMap map_v,    // qkv, box {32,64}, MN-major
                                                            const __grid_constant__ CUtensorMap map_out,  // ctx as (E, S, B), box {32,128,1}
                                                            const Dims p, const int n_tiles, int* __restrict__ sched,
                                                            const int32_t* __restrict__ x,
                                                            const int32_t* __restrict__ fnp, float* __restrict__ lse) {
  using C = FwdCfg<D, NS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t kv_s = base;                                 // NS stages of [K block | V block]
  const uint32_t q_s = base + NS * C::kStage;                 // Q image of the upcoming tile, later the O staging
  uint8_t* q_ptr = base_ptr + NS * C::kStage;
  uint8_t* ctl_ptr = q_ptr + C::kTile;
  const uint32_t bar = q_s + C::kTile;
  auto kfull = [&](int s) { return bar + 8u * s; };
  auto kfree = [&](int s) { return bar + 8u * (2 + s); };
  auto vfull = [&](int s) { return bar + 8u * (4 + s); };
  auto vfree = [&](int s) { return bar + 8u * (6 + s); };
  auto s_full = [&](int b2) { return bar + 8u * (8 + b2); };
  auto p_full = [&](int b2) { return bar + 8u * (10 + b2); };
  const uint32_t o_full = bar + 8u * 12, o_read = bar + 8u * 13, q_full = bar + 8u * 14, q_done = bar + 8u * 15, st_done = bar + 8u * 16,
                 tmem_slot = bar + 8u * 17;
  static_assert(NS <= 2, "barrier table holds two K/V stages");
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(ctl_ptr + 8 * 17);
  uint32_t* pad32 = reinterpret_cast<uint32_t*>(ctl_ptr + 192);               // [2][64] words: 32 keys each (S <= 2048 staged)
  volatile int* meta_s = reinterpret_cast<volatile int*>(ctl_ptr + 192 + 512); // ring of 4 x {tile or -1, block count, first_real, -}
  constexpr int kMaxPadWords = 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_bh = p.B * p.h;
  const int n_qt = (p.S + 127) / 128;
  const int n_all = (p.S + 63) / 64;
  const int n_words = 2 * n_all;
  const int band = n_qt <= 8 ? 2 : 4;
  struct Tile { int b, head, q0; };
  auto tile_of = [&](int t) {
    Tile tl;
    int bh, k;
    banded_tile(t, n_qt, n_bh, band, bh, k);
    tl.q0 = (n_qt - 1 - k) * 128;                             // latest query tiles have the most key blocks: first
    tl.b = bh / p.h;
    tl.head = bh - tl.b * p.h;
    return tl;
  };
  auto n_blocks = [&](const Tile& tl, int first_real) {
    return tl.q0 < first_real ? n_all : min(n_all, (tl.q0 + 127) / 64 + 1);
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(kfull(s), 1); mbar_init(kfree(s), 1); mbar_init(vfull(s), 1); mbar_init(vfree(s), 1); }
    for (int b2 = 0; b2 < 2; ++b2) { mbar_init(s_full(b2), 1); mbar_init(p_full(b2), 128); }
    mbar_init(o_full, 1); mbar_init(o_read, 128); mbar_init(q_full, 1); mbar_init(q_done, 128); mbar_init(st_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t t_q = tmem, t_o = tmem + D + 128;
  auto t_s = [&](int b2) { return tmem + D + 64 * b2; };

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane == 0) {
      auto issue_q = [&](const Tile& tl, int idx, int first_real, int t) {
        meta_s[(idx & 3) * 4] = t;
        meta_s[(idx & 3) * 4 + 1] = n_blocks(tl, first_real);
        meta_s[(idx & 3) * 4 + 2] = first_real;
        mbar_expect_tx(q_full, C::kTile);                      // (release: the ring entry is visible to every waiter)
#pragma unroll
        for (int a = 0; a < C::kAtoms; ++a)
          tma_load_2d(q_s + a * (128 * 128), &map_q, q_full, tl.head * D + a * 32, tl.b * p.S + tl.q0);
      };
      auto publish_stop = [&](int idx) {
        meta_s[(idx & 3) * 4] = -1;
        mbar_arrive(q_full);
      };
      auto draw = [&]() { const int t = atomicAdd(sched, 1); return t < n_tiles ? t : -1; };
      int t = draw();
      Tile tl = tile_of(max(t, 0));
      int first_real = t >= 0 ? __ldg(fnp + tl.b) : 0;
      if (t >= 0) issue_q(tl, 0, first_real, t); else publish_stop(0);
      int g = 0;
      for (int i = 0; t >= 0; ++i) {
        const int n_kv = n_blocks(tl, first_real);
        const int t_next = draw();
        Tile tn = tl;
        int fr_next = 0;
        if (t_next >= 0) { tn = tile_of(t_next); fr_next = __ldg(fnp + tn.b); }
        for (int kb = 0; kb < n_kv; ++kb, ++g) {
          const int s = g % NS;
          const uint32_t free_par = (uint32_t)((g / NS) & 1) ^ 1u;
          const int krow = tl.b * p.S + kb * 64;
          mbar_wait(kfree(s), free_par);
          mbar_expect_tx(kfull(s), C::kBlk);
#pragma unroll
          for (int a = 0; a < C::kAtoms; ++a)
            tma_load_2d(kv_s + s * C::kStage + a * (64 * 128), &map_k, kfull(s), p.E + tl.head * D + a * 32, krow);
          mbar_wait(vfree(s), free_par);
          mbar_expect_tx(vfull(s), C::kBlk);
#pragma unroll
          for (int a = 0; a < C::kAtoms; ++a)
            tma_load_2d(kv_s + s * C::kStage + C::kBlk + a * (64 * 128), &map_v, vfull(s), 2 * p.E + tl.head * D + a * 32, krow);
          if (kb == min(1, n_kv - 1)) {
            mbar_wait(q_done, (uint32_t)(i & 1));              // this tile's Q image has left the buffer ...
            if (i > 0) mbar_wait(st_done, (uint32_t)((i - 1) & 1));   // ... and so has the previous tile's O staging
            if (t_next >= 0) issue_q(tn, i + 1, fr_next, t_next); else publish_stop(i + 1);
          }
        }
        tl = tn;
        first_real = fr_next;
        t = t_next;
      }
      if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }   // counters ready for the next launch
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer (converged warp, elected lane) ----------------
    constexpr uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int g0 = 0;
    for (int i = 0;; ++i) {
      mbar_wait(q_full, (uint32_t)(i & 1));                    // (acquire) ring entry of tile i
      if (meta_s[(i & 3) * 4] < 0) break;
      mbar_wait(q_done, (uint32_t)(i & 1));                    // Q of this tile is in TMEM
      tc_fence_after();
      const int n_kv = meta_s[(i & 3) * 4 + 1];
      auto issue_s = [&](int j) {            // S = Q K^T (A = Q in TMEM)
        const int g = g0 + j, s = g % NS, b2 = g & 1;
        mbar_wait(kfull(s), (uint32_t)((g / NS) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t kd = make_desc(kv_s + s * C::kStage, 16, 1024, 2);
#pragma unroll
          for (int kk = 0; kk < D / 8; ++kk)
            umma_tf32_ts(t_s(b2), t_q + kk * 8, kd + (uint64_t)(((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4), idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(s_full(b2));
          umma_commit(kfree(s));
        }
        __syncwarp();
      };
      issue_s(0);
      if (n_kv > 1) issue_s(1);
      for (int j = 0; j < n_kv; ++j) {
        const int g = g0 + j, s = g % NS, b2 = g & 1;
        mbar_wait(vfull(s), (uint32_t)((g / NS) & 1));
        mbar_wait(p_full(b2), (uint32_t)((g >> 1) & 1));
        if (g > 0) mbar_wait(o_read, (uint32_t)((g - 1) & 1));      // the previous O block has been folded into registers
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t vd = make_desc(kv_s + s * C::kStage + C::kBlk, 64 * 128, 512, 1);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)          // O_blk = P V (A = P in TMEM, K = 64 keys)
            umma_tf32_ts(t_o, t_s(b2) + kk * 8, vd + (uint64_t)(kk * 64), idesc_o, kk > 0 ? 1u : 0u);
          umma_commit(o_full);
          umma_commit(vfree(s));
        }
        __syncwarp();
        if (j + 2 < n_kv) issue_s(j + 2);
      }
      g0 += n_kv;
    }
  } else {
    // ---------------- softmax warps: thread == query row ----------------
    const int cw = warp - 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint64_t scale2 = pack2f(p.scale_log2, p.scale_log2);
    const bool pads_staged = n_words <= kMaxPadWords;
    const bool storer = cw == 0 && lane == 0;

    auto stage_q = [&](int idx) {                              // Q image -> TMEM (this thread's row); false = no tile
      mbar_wait(q_full, (uint32_t)(idx & 1));
      if (meta_s[(idx & 3) * 4] < 0) return false;
#pragma unroll
      for (int a = 0; a < C::kAtoms; ++a) {
        uint32_t v[32];
        load_swizzled_row(q_ptr + a * (128 * 128), r, v);
        tmem_st32(t_q + lane_addr + a * 32, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(q_done);
      return true;
    };
    // PAD words of a sequence: warp cw owns words cw, cw+4, ...; up to 4 of them ride in registers across a drain
    int32_t pad_id[4];
    auto fetch_pads = [&](int b) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int key = (cw + 4 * u) * 32 + lane;
        pad_id[u] = key < p.S ? __ldg(x + (int64_t)b * p.S + key) : 0;
      }
    };
    auto store_pads = [&](int b, int idx) {
      uint32_t* dst = pad32 + (idx & 1) * kMaxPadWords;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t bits = __ballot_sync(0xffffffffu, pad_id[u] == 0);
        if (lane == 0 && cw + 4 * u < n_words) dst[cw + 4 * u] = bits;
      }
      for (int w = cw + 16; w < n_words; w += 4) {            // sequences longer than 512: the rest, synchronously
        const int key = w * 32 + lane;
        const uint32_t bits = __ballot_sync(0xffffffffu, key < p.S ? (__ldg(x + (int64_t)b * p.S + key) == 0) : true);
        if (lane == 0) dst[w] = bits;
      }
    };

    bool have = stage_q(0);
    Tile tl = tile_of(have ? meta_s[0] : 0);
    if (have && pads_staged) { fetch_pads(tl.b); store_pads(tl.b, 0); }
    asm volatile("bar.sync 1, 128;" ::: "memory");            // PAD words visible
    int g0 = 0;
    for (int i = 0; have; ++i) {
      const int n_kv = meta_s[(i & 3) * 4 + 1], first_real = meta_s[(i & 3) * 4 + 2];
      const int row = tl.q0 + r;
      const bool deg = row < first_real;
      const uint32_t* pads = pad32 + (i & 1) * kMaxPadWords;
      float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
      uint32_t o[D];
#pragma unroll
      for (int c = 0; c < D; ++c) o[c] = 0u;
      auto fold = [&](int g, float alpha) {                    // o = o * alpha + O_blk(g)
        const uint64_t alpha2 = pack2f(alpha, alpha);
        mbar_wait(o_full, (uint32_t)(g & 1));
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < D / 32; ++cc) {
          uint32_t ov[32];
          tmem_ld32_nowait(t_o + lane_addr + cc * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; c += 2)
            unpack2(ffma2(pack2(o[cc * 32 + c], o[cc * 32 + c + 1]), alpha2, pack2(ov[c], ov[c + 1])), o[cc * 32 + c], o[cc * 32 + c + 1]);
        }
        tc_fence_before();
        mbar_arrive(o_read);
      };
      for (int j = 0; j < n_kv; ++j) {
        const int g = g0 + j, b2 = g & 1;
        const int k0 = j * 64;
        unsigned long long padbits;
        if (pads_staged) {
          padbits = (unsigned long long)pads[2 * j] | ((unsigned long long)pads[2 * j + 1] << 32);
        } else {
          const int ka = k0 + lane, kb2 = k0 + 32 + lane;
          const bool pa = ka < p.S ? (x[(int64_t)tl.b * p.S + ka] == 0) : true;
          const bool pb = kb2 < p.S ? (x[(int64_t)tl.b * p.S + kb2] == 0) : true;
          padbits = (unsigned long long)__ballot_sync(0xffffffffu, pa) | ((unsigned long long)__ballot_sync(0xffffffffu, pb) << 32);
        }
        const int in_range = min(64, p.S - k0);
        const unsigned long long range_bits = in_range >= 64 ? ~0ull : ((1ull << in_range) - 1ull);
        unsigned long long vis;
        if (row >= p.S) vis = 0ull;
        else if (deg) vis = range_bits;
        else {
          const int upto = row - k0;
          const unsigned long long causal = upto < 0 ? 0ull : (upto >= 63 ? ~0ull : ((1ull << (upto + 1)) - 1ull));
          vis = causal & ~padbits & range_bits;
        }
        mbar_wait(s_full(b2), (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        uint32_t s0[32], s1[32];
        tmem_ld32_nowait(t_s(b2) + lane_addr, s0);
        tmem_ld32_nowait(t_s(b2) + lane_addr + 32, s1);
        tmem_ld_wait();
        const bool all_vis = __all_sync(0xffffffffu, vis == ~0ull);
        float mx = -INFINITY;
        if (all_vis) {
#pragma unroll
          for (int c = 0; c < 32; ++c) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[c]), __uint_as_float(s1[c])));
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (vis & (1ull << c)) mx = fmaxf(mx, __uint_as_float(s0[c]));
            if (vis & (1ull << (c + 32))) mx = fmaxf(mx, __uint_as_float(s1[c]));
          }
        }
        const float m_new = fmaxf(m, mx * p.scale_log2);
        const float m_use = m_new == -INFINITY ? 0.f : m_new;      // row with no visible key yet
        const float alpha = fast_exp2(m - m_use);                   // exp2(-inf) = 0 when m was -inf
        const uint64_t negm2 = pack2f(-m_use, -m_use);
        uint64_t rsum2 = 0ull;
        const uint32_t vlo = (uint32_t)vis, vhi = (uint32_t)(vis >> 32);
        if (all_vis) {
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            uint32_t xa, xb, ya, yb;
            unpack2(ffma2(pack2(s0[c], s0[c + 1]), scale2, negm2), xa, xb);
            unpack2(ffma2(pack2(s1[c], s1[c + 1]), scale2, negm2), ya, yb);
            const float e0 = fast_exp2(__uint_as_float(xa)), e1 = fast_exp2(__uint_as_float(xb));
            const float f0 = fast_exp2(__uint_as_float(ya)), f1 = fast_exp2(__uint_as_float(yb));
            rsum2 = fadd2(rsum2, fadd2(pack2f(e0, e1), pack2f(f0, f1)));
            s0[c] = rna_tf32_bits(__float_as_uint(e0)); s0[c + 1] = rna_tf32_bits(__float_as_uint(e1));
            s1[c] = rna_tf32_bits(__float_as_uint(f0)); s1[c + 1] = rna_tf32_bits(__float_as_uint(f1));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            uint32_t xa, xb, ya, yb;
            unpack2(ffma2(pack2(s0[c], s0[c + 1]), scale2, negm2), xa, xb);
            unpack2(ffma2(pack2(s1[c], s1[c + 1]), scale2, negm2), ya, yb);
            const float e0 = (vlo & (1u << c)) ? fast_exp2(__uint_as_float(xa)) : 0.f, e1 = (vlo & (2u << c)) ? fast_exp2(__uint_as_float(xb)) : 0.f;
            const float f0 = (vhi & (1u << c)) ? fast_exp2(__uint_as_float(ya)) : 0.f, f1 = (vhi & (2u << c)) ? fast_exp2(__uint_as_float(yb)) : 0.f;
            rsum2 = fadd2(rsum2, fadd2(pack2f(e0, e1), pack2f(f0, f1)));
            s0[c] = rna_tf32_bits(__float_as_uint(e0)); s0[c + 1] = rna_tf32_bits(__float_as_uint(e1));
            s1[c] = rna_tf32_bits(__float_as_uint(f0)); s1[c + 1] = rna_tf32_bits(__float_as_uint(f1));
          }
        }
        tmem_st32(t_s(b2) + lane_addr, s0);
        tmem_st32(t_s(b2) + lane_addr + 32, s1);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(b2));
        if (j == 0 && i > 0 && storer) {                       // the previous tile's bulk store has had a whole block to read its staging
          bulk_wait_read0();
          mbar_arrive(st_done);
        }
        uint32_t ra, rb;
        unpack2(rsum2, ra, rb);
        l = l * alpha + (__uint_as_float(ra) + __uint_as_float(rb));
        m = m_new;
        if (j > 0) fold(g - 1, alpha_prev);        // the product this thread enabled one block ago has long retired
        alpha_prev = alpha;
      }
      fold(g0 + n_kv - 1, alpha_prev);
      g0 += n_kv;
      // natural-log LSE of the scaled scores (m, l are in the exp2 domain)
      if (row < p.S) lse[((int64_t)tl.b * p.h + tl.head) * p.S + row] = (m + log2f(l)) * (1.f / kLog2e);

      // every S product of this tile has retired: the next tile's Q goes to TMEM before this tile's output is drained
      have = stage_q(i + 1);
      Tile tn = tl;
      if (have) {
        tn = tile_of(meta_s[((i + 1) & 3) * 4]);
        if (pads_staged) fetch_pads(tn.b);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");          // every row of the Q image has been read: the buffer becomes the O staging
      {
        const float inv = 1.f / l;
#pragma unroll
        for (int cc = 0; cc < D / 32; ++cc) {
          uint32_t part[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) part[c] = o[cc * 32 + c];
          put_row_sw128(q_ptr + cc * (128 * 128), r, part, inv);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (storer) {
#pragma unroll
          for (int at = 0; at < C::kAtoms; ++at)
            tma_store_3d(&map_out, q_s + at * (128 * 128), tl.head * D + at * 32, tl.q0, tl.b);
          bulk_commit();
        }
      }
      if (have && pads_staged) store_pads(tn.b, i + 1);
      asm volatile("bar.sync 1, 128;" ::: "memory");          // PAD words of the next tile visible
      tl = tn;
    }
    if (storer) bulk_wait_read0();                             // the last tile's staging must outlive its store
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

template <int D>
int launch_fwd(cudaStream_t st, int B, int S, int h, const float* qkv, const int32_t* x, const int32_t* fnp, float* ctx,
               float* lse) {
  constexpr int NS = 2;
  using C = FwdCfg<D, NS>;
  const int E = h * D;
  CUtensorMap mq, mk, mv;
  const uint64_t rows = (uint64_t)B * S;
  NLPS_TRY(make_map(&mq, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, 128, true, false));
  NLPS_TRY(make_map(&mk, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, 64, true, false));
  NLPS_TRY(make_map(&mv, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, 64, true, true));
  static bool once = false;
  if (!once) {
    NLPS_CUDA(cudaFuncSetAttribute(attn_fwd_ts<D, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    once = true;
  }
  static int trace_on = -1;
  if (trace_on < 0) { const char* e = getenv("NLPS_ATTN_TRACE"); trace_on = (e && e[0] == '1') ? 1 : 0; }
  Dims p{B, S, h, E, (float)(1.0 / std::sqrt((double)D)) * kLog2e, trace_on};
  dim3 grid((unsigned)ceil_div(S, 128), (unsigned)h, (unsigned)B);
  static int persistent = -1;
  if (persistent < 0) { const char* e = getenv("NLPS_ATTN_PS"); persistent = (e && e[0] == '0') ? 0 : 1; }
  const int64_t n_tiles = ceil_div(S, 128) * (int64_t)B * h;
  if (persistent && !trace_on && (reinterpret_cast<uintptr_t>(ctx) & 15u) == 0 && n_tiles < (1ll << 30)) {
    static bool once_ps = false;
    if (!once_ps) {
      NLPS_CUDA(cudaFuncSetAttribute(attn_fwd_ps<D, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
      once_ps = true;
    }
    CUtensorMap out3;
    NLPS_TRY(make_map3(&out3, ctx, (uint64_t)E, (uint64_t)S, (uint64_t)B, (int64_t)E, (int64_t)S * E, 32, 128));
    int* sched = nullptr;
    NLPS_TRY(sched_counters(st, 2, &sched));
    attn_fwd_ps<D, NS><<<(unsigned)std::min<int64_t>(n_tiles, 2 * kNumSMs), kFwdThreads, C::kSmem, st>>>(
        mq, mk, mv, out3, p, (int)n_tiles, sched, x, fnp, lse);
  } else {
    attn_fwd_ts<D, NS><<<grid, kFwdThreads, C::kSmem, st>>>(mq, mk, mv, p, x, fnp, ctx, lse);
  }
  NLPS_LAUNCH_CHECK();
  if (trace_on) {
    static int printed = 0;
    if (printed++ == 3) {
      static long long hbuf[kTraceSlots];
      cudaStreamSynchronize(st);
      cudaMemcpyFromSymbol(hbuf, g_ts_trace, sizeof(hbuf));
      const long long t0 = hbuf[0];
      printf("attn_fwd_ts trace, last query tile of (b0,h0) (cycles since CTA start): Q TMA issue %lld..%lld, setup sync %lld, pad masks built %lld, Q landed %lld, Q in TMEM %lld, "
             "loop done %lld, tile in smem %lld, barrier %lld, stores done %lld\n",
             hbuf[9] - t0, hbuf[10] - t0, hbuf[4] - t0, hbuf[5] - t0, hbuf[6] - t0, hbuf[1] - t0, hbuf[2] - t0, hbuf[7] - t0, hbuf[8] - t0, hbuf[3] - t0);
      const int n = std::min(8, (S + 63) / 64);
      for (int j = 0; j < n; ++j)
        printf("kb %2d | TMA K %6lld V %6lld | MMA S start %6lld issued %6lld PV start %6lld issued %6lld | thread top %6lld s_full %6lld P stored %6lld prev O folded %6lld\n",
               j, hbuf[100 + j * 2] - t0, hbuf[101 + j * 2] - t0, hbuf[200 + j * 4] - t0, hbuf[201 + j * 4] - t0, hbuf[202 + j * 4] - t0,
               hbuf[203 + j * 4] - t0, hbuf[300 + j * 4] - t0, hbuf[301 + j * 4] - t0, hbuf[302 + j * 4] - t0, hbuf[303 + j * 4] - t0);
      fflush(stdout);
    }
  }
  return 0;
}

}  // namespace

int attention_backward_ts(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                          const int32_t* fnp, const float* dctx, const float* lse, const float* delta, float* dqkv) {
  if (d == 64) return launch_bwd<64>(st, B, S, h, qkv, x, fnp, dctx, lse, delta, dqkv);
  return launch_bwd<32>(st, B, S, h, qkv, x, fnp, dctx, lse, delta, dqkv);
}

int attention_forward_ts(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                         const int32_t* fnp, float* ctx, float* lse) {
  if (d == 64) return launch_fwd<64>(st, B, S, h, qkv, x, fnp, ctx, lse);
  return launch_fwd<32>(st, B, S, h, qkv, x, fnp, ctx, lse);
}

}  // namespace nlps
