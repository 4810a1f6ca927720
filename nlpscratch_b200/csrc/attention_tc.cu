// Fused causal + key-padding attention on tcgen05 tensor cores (kind::tf32, accumulators in TMEM),
// forward and backward, for head_dim 32 or 64.  Same semantics and outputs as attention.cu /
// attention_mma.cu (context + log-sum-exp; dQ|dK|dV packed; -1e9 additive-mask behaviour including
// fully-masked rows); every output element has one owner thread (no atomics, deterministic).
//
// Forward, one CTA per (batch, head, 128-query tile), 2 CTAs per SM:
//   TMA      : Q tile (128 x d, K-major, SWIZZLE_128B) once; per 64-key block the K tile (K-major) and
//              the V tile (MN-major, 128B swizzle with 32-B atoms) -- all three are boxes of the same
//              packed (B*S, 3E) qkv matrix, the head split is a column coordinate, never a copy.
//   tcgen05  : S = Q K^T (M=128, N=64, K=d) into TMEM; thread r of the 4 softmax warps owns query row r
//              (= TMEM lane r): tcgen05.ld its 64 scores, mask, online max/sum (no cross-thread
//              reduction at all), write P (TF32-rounded) into smem in the K-major SWIZZLE_128B layout
//              the tensor core reads; O_blk = P V (M=128, N=d, K=64) into TMEM; the thread folds O_blk
//              into its register accumulator with the running rescale.
//   Two CTAs per SM interleave: one runs its softmax while the other's MMAs and TMA are in flight.
#include "tc_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace nlps {
namespace {
using namespace tc;

constexpr int BQ = 128;         // query rows per CTA == TMEM lanes
constexpr int BKV = 64;         // keys per block
constexpr int kThreads = 192;   // warp0 TMA, warp1 MMA + TMEM owner, warps2-5 softmax
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kMaxPadBlocks = 64;

struct Dims {
  int B, S, h, E;
  float scale_log2;   // (1/sqrt(d)) * log2(e)
  int trace;
};

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

// NLPS_ATTN_TRACE=1: CTA (3,0,0) of the forward kernel records clock64() at its pipeline events
__device__ long long g_attn_trace[4096];
__device__ __forceinline__ void trace(bool on, int slot) { if (on) g_attn_trace[slot] = clock64(); }

template <int D>
struct FwdCfg {
  static constexpr int kAtoms = D / 32;                 // 128-byte K atoms per row
  static constexpr int kQBytes = kAtoms * BQ * 128;     // [atom][128 rows][128 B]
  static constexpr int kKBytes = kAtoms * BKV * 128;    // [atom][64 rows][128 B]
  static constexpr int kVBytes = kAtoms * BKV * 128;    // [slab of 32 dims][64 key rows][128 B]
  static constexpr int kPBytes = (BKV / 32) * BQ * 128; // [atom of 32 keys][128 rows][128 B]
  static constexpr int kBarBytes = 1024;                // barriers (128 B) + PAD bit masks of up to 64 key blocks
  static constexpr int kSmemBytes = kQBytes + kKBytes + kVBytes + kPBytes + kBarBytes + 1024;
  static constexpr int kTmemCols = 128;                 // S: 64 columns, O block: D columns
};

template <int D>
__global__ void __launch_bounds__(kThreads, 2) attn_fwd_tc(const __grid_constant__ CUtensorMap map_q,
                                                        const __grid_constant__ CUtensorMap map_k,
                                                        const __grid_constant__ CUtensorMap map_v, const Dims p,
                                                        const int32_t* __restrict__ x,
                                                        const int32_t* __restrict__ fnp, float* __restrict__ ctx,
                                                        float* __restrict__ lse) {
  using Cfg = FwdCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t q_s = base, k_s = q_s + Cfg::kQBytes, v_s = k_s + Cfg::kKBytes, p_s = v_s + Cfg::kVBytes;
  uint8_t* p_ptr = base_ptr + Cfg::kQBytes + Cfg::kKBytes + Cfg::kVBytes;
  const uint32_t bar = p_s + Cfg::kPBytes;
  // K and V have separate full/empty barriers: K(j+1) streams in as soon as S(j) has been issued (during
  // softmax j and P V j), V(j+1) as soon as P V j retires -- TMA latency hidden without a second buffer.
  const uint32_t q_full = bar, k_full = bar + 8, k_empty = bar + 16, s_full = bar + 24, p_full = bar + 32,
                 o_full = bar + 40, o_read = bar + 48, v_full = bar + 56, v_empty = bar + 64, tmem_slot = bar + 72;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(p_ptr + Cfg::kPBytes + 72);
  unsigned long long* padmask_s = reinterpret_cast<unsigned long long*>(p_ptr + Cfg::kPBytes + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * BQ;
  const int first_real = fnp[b];
  const bool deg_block = q0 < first_real;
  const int n_all = (p.S + BKV - 1) / BKV;
  const int n_kv = deg_block ? n_all : min(n_all, (q0 + BQ - 1) / BKV + 1);
  const int row_base = b * p.S;                       // first row of this sequence in the (B*S, 3E) matrix
  const bool tr = p.trace && blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0;
  trace(tr && threadIdx.x == 0, 0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(s_full, 1);
    mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_read, 128); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tmem_s = tmem, tmem_o = tmem + 64;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
      for (int a = 0; a < Cfg::kAtoms; ++a)
        tma_load_2d(q_s + a * (BQ * 128), &map_q, q_full, head * D + a * 32, row_base + q0);
      for (int kb = 0; kb < n_kv; ++kb) {
        const int krow = row_base + kb * BKV;
        mbar_wait(k_empty, (uint32_t)(kb & 1) ^ 1u);
        trace(tr, 100 + kb * 2);
        mbar_expect_tx(k_full, Cfg::kKBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_2d(k_s + a * (BKV * 128), &map_k, k_full, p.E + head * D + a * 32, krow);
        mbar_wait(v_empty, (uint32_t)(kb & 1) ^ 1u);
        trace(tr, 101 + kb * 2);
        mbar_expect_tx(v_full, Cfg::kVBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_2d(v_s + a * (BKV * 128), &map_v, v_full, 2 * p.E + head * D + a * 32, krow);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // S = Q K^T : A, B K-major.  O = P V : A K-major (P in smem), B MN-major (V).
      constexpr uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      mbar_wait(q_full, 0);
      for (int kb = 0; kb < n_kv; ++kb) {
        const uint32_t ph = (uint32_t)(kb & 1);
        mbar_wait(k_full, ph);
        trace(tr, 200 + kb * 4);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 8; ++kk) {
          const uint64_t ad = make_desc(q_s + (kk >> 2) * (BQ * 128) + (kk & 3) * 32, 16, 1024, 2);
          const uint64_t bd = make_desc(k_s + (kk >> 2) * (BKV * 128) + (kk & 3) * 32, 16, 1024, 2);
          umma_tf32(tmem_s, ad, bd, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(k_empty);                          // K tile free once S retires
        trace(tr, 201 + kb * 4);
        mbar_wait(v_full, ph);
        trace(tr, 202 + kb * 4);
        mbar_wait(p_full, ph);                         // P written (and S consumed) by all 128 rows
        if (kb > 0) mbar_wait(o_read, (uint32_t)((kb - 1) & 1));   // previous O block folded into registers
        trace(tr, 203 + kb * 4);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BKV / 8; ++kk) {
          const uint64_t ad = make_desc(p_s + (kk >> 2) * (BQ * 128) + (kk & 3) * 32, 16, 1024, 2);
          const uint64_t bd = make_desc(v_s + kk * 1024, BKV * 128, 512, 1);
          umma_tf32(tmem_o, ad, bd, idesc_o, kk > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(v_empty);
      }
    }
    __syncwarp();
  } else {
    // ---- softmax / accumulate: thread == query row ----
    const int q = warp & 3;
    const int r = q * 32 + lane;                      // row inside the tile == TMEM lane
    const int row = q0 + r;
    const bool deg = row < first_real;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float o[D];
#pragma unroll
    for (int c = 0; c < D; ++c) o[c] = 0.f;
    // PAD bits of every key block of this sequence, built once by the four softmax warps
    // (two ballots per block) instead of a dependent global load at the top of every iteration
    if (n_kv <= kMaxPadBlocks) {
      for (int kb = q; kb < n_kv; kb += 4) {
        const int ka = kb * BKV + lane, kb2 = ka + 32;
        const bool pa = ka < p.S ? (x[(int64_t)b * p.S + ka] == 0) : true;
        const bool pb = kb2 < p.S ? (x[(int64_t)b * p.S + kb2] == 0) : true;
        const unsigned long long bits = (unsigned long long)__ballot_sync(0xffffffffu, pa) |
                                        ((unsigned long long)__ballot_sync(0xffffffffu, pb) << 32);
        if (lane == 0) padmask_s[kb] = bits;
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
    }
    for (int kb = 0; kb < n_kv; ++kb) {
      const uint32_t ph = (uint32_t)(kb & 1);
      const int k0 = kb * BKV;
      // visibility bits of this row for the 64 keys of the block
      unsigned long long padbits;
      if (n_kv <= kMaxPadBlocks) {
        padbits = padmask_s[kb];
      } else {
        const int ka = k0 + lane, kb2 = k0 + 32 + lane;
        const bool pa = ka < p.S ? (x[(int64_t)b * p.S + ka] == 0) : true;
        const bool pb = kb2 < p.S ? (x[(int64_t)b * p.S + kb2] == 0) : true;
        padbits = (unsigned long long)__ballot_sync(0xffffffffu, pa) |
                  ((unsigned long long)__ballot_sync(0xffffffffu, pb) << 32);
      }
      const int in_range = min(BKV, p.S - k0);        // >= 1
      const unsigned long long range_bits = in_range >= 64 ? ~0ull : ((1ull << in_range) - 1ull);
      unsigned long long vis;
      if (row >= p.S) vis = 0ull;
      else if (deg) vis = range_bits;
      else {
        const int upto = row - k0;                    // keys k0..row are causal-visible
        const unsigned long long causal = upto < 0 ? 0ull : (upto >= 63 ? ~0ull : ((1ull << (upto + 1)) - 1ull));
        vis = causal & ~padbits & range_bits;
      }
      const bool trs = tr && threadIdx.x == 64;
      trace(trs, 300 + kb * 5);
      mbar_wait(s_full, ph);
      trace(trs, 301 + kb * 5);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(tmem_s + lane_addr, s0);
      tmem_ld_32x32(tmem_s + lane_addr + 32, s1);
      // running max in the exp2 domain: scores are multiplied by scale*log2(e) > 0, so the max of the raw
      // scores decides; p = exp2(fma(s, scale_log2, -m)) is one FFMA + one MUFU per element
      const bool all_vis = __all_sync(0xffffffffu, vis == ~0ull);      // interior block: no mask tests at all
      float mx = -INFINITY;
      if (all_vis) {
#pragma unroll
        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[c]), __uint_as_float(s1[c])));
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if ((vis >> c) & 1ull) mx = fmaxf(mx, __uint_as_float(s0[c]));
          if ((vis >> (c + 32)) & 1ull) mx = fmaxf(mx, __uint_as_float(s1[c]));
        }
      }
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;      // row with no visible key yet
      const float alpha = fast_exp2(m - m_use);                   // exp2(-inf) = 0 when m was -inf
      const float neg_m = -m_use;
      float rsum = 0.f;
      // P row -> smem, K-major SWIZZLE_128B: [atom of 32 keys][row][8 chunks of 16 B, chunk ^= row & 7]
      uint8_t* prow = p_ptr + r * 128;
      if (all_vis) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = i * 4 + j;
            pv[j] = fast_exp2(fmaf(__uint_as_float(c < 32 ? s0[c & 31] : s1[c & 31]), p.scale_log2, neg_m));
            rsum += pv[j];
          }
          uint4 w = make_uint4(to_tf32(pv[0]), to_tf32(pv[1]), to_tf32(pv[2]), to_tf32(pv[3]));
          *reinterpret_cast<uint4*>(prow + (i >> 3) * (BQ * 128) + (((i & 7) ^ (r & 7)) << 4)) = w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = i * 4 + j;
            const float e = fast_exp2(fmaf(__uint_as_float(c < 32 ? s0[c & 31] : s1[c & 31]), p.scale_log2, neg_m));
            pv[j] = ((vis >> c) & 1ull) ? e : 0.f;
            rsum += pv[j];
          }
          uint4 w = make_uint4(to_tf32(pv[0]), to_tf32(pv[1]), to_tf32(pv[2]), to_tf32(pv[3]));
          *reinterpret_cast<uint4*>(prow + (i >> 3) * (BQ * 128) + (((i & 7) ^ (r & 7)) << 4)) = w;
        }
      }
      l = l * alpha + rsum;
      m = m_new;
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
      trace(trs, 302 + kb * 5);
#pragma unroll
      for (int c = 0; c < D; ++c) o[c] *= alpha;
      mbar_wait(o_full, ph);
      trace(trs, 303 + kb * 5);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc) {
        uint32_t ov[32];
        tmem_ld_32x32(tmem_o + lane_addr + cc * 32, ov);
#pragma unroll
        for (int c = 0; c < 32; ++c) o[cc * 32 + c] += __uint_as_float(ov[c]);
      }
      tc_fence_before();
      mbar_arrive(o_read);
      trace(trs, 304 + kb * 5);
    }
    trace(tr && threadIdx.x == 64, 1);
    if (row < p.S) {
      const float inv = 1.f / l;
      float4* dst = reinterpret_cast<float4*>(ctx + ((int64_t)b * p.S + row) * p.E + head * D);
#pragma unroll
      for (int c = 0; c < D / 4; ++c) dst[c] = make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
      // natural-log LSE of the scaled scores: m and l are in the exp2 domain
      lse[((int64_t)b * p.h + head) * p.S + row] = (m + log2f(l)) * (1.f / kLog2e);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::kTmemCols);
  }
}

template <int D>
int launch_fwd(cudaStream_t st, int B, int S, int h, const float* qkv, const int32_t* x, const int32_t* fnp, float* ctx,
               float* lse) {
  using Cfg = FwdCfg<D>;
  const int E = h * D;
  CUtensorMap mq, mk, mv;
  const uint64_t rows = (uint64_t)B * S;
  NLPS_TRY(make_map(&mq, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, BQ, true, false));
  NLPS_TRY(make_map(&mk, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, BKV, true, false));
  NLPS_TRY(make_map(&mv, qkv, 3 * (uint64_t)E, rows, 3 * (int64_t)E, 32, BKV, true, true));
  static bool once = false;
  if (!once) {
    NLPS_CUDA(cudaFuncSetAttribute(attn_fwd_tc<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    once = true;
  }
  static int trace_on = -1;
  if (trace_on < 0) { const char* e = getenv("NLPS_ATTN_TRACE"); trace_on = (e && e[0] == '1') ? 1 : 0; }
  Dims p{B, S, h, E, (float)(1.0 / std::sqrt((double)D)) * kLog2e, trace_on};
  dim3 grid((unsigned)ceil_div(S, BQ), (unsigned)h, (unsigned)B);
  attn_fwd_tc<D><<<grid, kThreads, Cfg::kSmemBytes, st>>>(mq, mk, mv, p, x, fnp, ctx, lse);
  NLPS_LAUNCH_CHECK();
  if (trace_on) {
    static int printed = 0;
    if (printed++ == 3) {
      long long h[4096];
      cudaStreamSynchronize(st);
      cudaMemcpyFromSymbol(h, g_attn_trace, sizeof(h));
      const long long t0 = h[0];
      printf("attn_fwd trace (cycles since CTA start), CTA end of loop at %lld\n", h[1] - t0);
      for (int kb = 0; kb < 8; ++kb)
        printf("kb %d | TMA K issue %6lld V issue %6lld | MMA k_full %6lld S issued %6lld v_full %6lld p+o ready %6lld | softmax wait_s %6lld s_full %6lld P arrived %6lld o_full %6lld o done %6lld\n",
               kb, h[100 + kb * 2] - t0, h[101 + kb * 2] - t0, h[200 + kb * 4] - t0, h[201 + kb * 4] - t0, h[202 + kb * 4] - t0,
               h[203 + kb * 4] - t0, h[300 + kb * 5] - t0, h[301 + kb * 5] - t0, h[302 + kb * 5] - t0, h[303 + kb * 5] - t0, h[304 + kb * 5] - t0);
      fflush(stdout);
    }
  }
  return 0;
}


// =========================================================================================
// Backward.  Recomputes P from Q, K and the stored log-sum-exp; delta_i = dO_i . O_i comes from
// attn_delta_kernel (attention.cu).  Two kernels so that every output has exactly one owner:
//
//   dK/dV kernel: CTA = (batch, head, 128 keys), thread r of the compute warps owns key r.  Per block of
//     64 queries the tensor core forms the TRANSPOSED tiles S^T = K Q^T and dP^T = V dO^T (M = keys), the
//     thread turns its two TMEM rows into P^T and dS^T = P^T o (dP^T - delta) and writes them to smem as
//     K-major A operands; then dV += P^T dO and dK += dS^T Q accumulate in TMEM over the whole loop.
//     Q and dO are needed in both operand majors ([N=q][K=dim] for the first pair of products,
//     [K=q][N=dim] for the second): TMA delivers each tile twice, once per swizzle mode.
//   dQ kernel: CTA = (batch, head, 128 queries), thread r owns query r: S = Q K^T, dP = dO V^T,
//     dS -> smem, dQ += dS K accumulates in TMEM.
template <int D>
struct BwdCfg {
  static constexpr int kAtoms = D / 32;
  static constexpr int kResBytes = kAtoms * 128 * 128;       // a resident 128-row operand tile
  static constexpr int kBlkBytes = kAtoms * 64 * 128;        // a streamed 64-row operand tile
  static constexpr int kPBytes = 2 * 128 * 128;              // [2 atoms of 32][128 rows][128 B]
  static constexpr int kBarBytes = 128 + 2 * 64 * 4;         // barriers + lse/delta of the query block
  static constexpr int kSmemDkdv = 2 * kResBytes + 4 * kBlkBytes + 2 * kPBytes + kBarBytes + 1024;
  static constexpr int kSmemDq = 2 * kResBytes + 3 * kBlkBytes + kPBytes + kBarBytes + 1024;
  static constexpr int kTmemCols = 256;
};

// write one 64-wide row of an A operand (K-major, SWIZZLE_128B, [atom of 32][128 rows][128 B])
__device__ __forceinline__ void store_a_row(uint8_t* tile, int r, const float (&v)[64]) {
  uint8_t* row = tile + r * 128;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint4 w = make_uint4(to_tf32(v[4 * i]), to_tf32(v[4 * i + 1]), to_tf32(v[4 * i + 2]), to_tf32(v[4 * i + 3]));
    *reinterpret_cast<uint4*>(row + (i >> 3) * (128 * 128) + (((i & 7) ^ (r & 7)) << 4)) = w;
  }
}

__device__ __forceinline__ unsigned long long low_bits(int n) {   // n lowest bits set, n clamped to [0,64]
  return n <= 0 ? 0ull : (n >= 64 ? ~0ull : ((1ull << n) - 1ull));
}

template <int D>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_dkdv_tc(const __grid_constant__ CUtensorMap map_res,   // qkv, box {32,128}, K-major
                                                              const __grid_constant__ CUtensorMap map_qk,    // qkv, box {32,64}, K-major
                                                              const __grid_constant__ CUtensorMap map_qmn,   // qkv, box {32,64}, MN-major
                                                              const __grid_constant__ CUtensorMap map_dok,   // dctx, box {32,64}, K-major
                                                              const __grid_constant__ CUtensorMap map_domn,  // dctx, box {32,64}, MN-major
                                                              const Dims p, const int32_t* __restrict__ x,
                                                              const int32_t* __restrict__ fnp,
                                                              const float* __restrict__ lse,
                                                              const float* __restrict__ delta,
                                                              float* __restrict__ dqkv) {
  using Cfg = BwdCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t k_s = base, v_s = k_s + Cfg::kResBytes;
  const uint32_t qk_s = v_s + Cfg::kResBytes, dok_s = qk_s + Cfg::kBlkBytes, qmn_s = dok_s + Cfg::kBlkBytes,
                 domn_s = qmn_s + Cfg::kBlkBytes;
  const uint32_t pt_s = domn_s + Cfg::kBlkBytes, dst_s = pt_s + Cfg::kPBytes;
  const int pt_off = 2 * Cfg::kResBytes + 4 * Cfg::kBlkBytes;
  uint8_t* pt_ptr = base_ptr + pt_off;
  uint8_t* dst_ptr = pt_ptr + Cfg::kPBytes;
  const uint32_t bar = dst_s + Cfg::kPBytes;
  uint8_t* bar_ptr = dst_ptr + Cfg::kPBytes;
  // two operand groups with their own barriers: the K-major copies (first pair of products) are refilled
  // while the threads work on P^T/dS^T, the MN-major copies (second pair) while the next S^T is formed
  const uint32_t res_full = bar, blk_full = bar + 8, blk_empty = bar + 16, sdp_full = bar + 24, pds_full = bar + 32,
                 acc_full = bar + 40, mn_full = bar + 48, mn_empty = bar + 56, tmem_slot = bar + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bar_ptr + 64);
  float* lse_s = reinterpret_cast<float*>(bar_ptr + 128);
  float* delta_s = lse_s + 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y, k0 = blockIdx.x * 128;
  const int first_real = fnp[b];
  const int row_base = b * p.S;
  const int n_q = (p.S + 63) / 64;
  const int first_causal = k0 / 64;                               // first query block with q >= some key here
  const int n_deg = min(min((first_real + 63) / 64, n_q), first_causal);
  const int n_it = n_deg + (n_q - first_causal);
  auto block_of = [&](int it) { return it < n_deg ? it : it - n_deg + first_causal; };

  if (warp == 0 && lane == 0) {
    mbar_init(res_full, 1); mbar_init(blk_full, 1); mbar_init(blk_empty, 1); mbar_init(sdp_full, 1);
    mbar_init(pds_full, 128); mbar_init(acc_full, 1); mbar_init(mn_full, 1); mbar_init(mn_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t t_st = tmem, t_dpt = tmem + 64, t_dv = tmem + 128, t_dk = tmem + 128 + D;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(res_full, 2 * Cfg::kResBytes);
#pragma unroll
      for (int a = 0; a < Cfg::kAtoms; ++a) {
        tma_load_2d(k_s + a * (128 * 128), &map_res, res_full, p.E + head * D + a * 32, row_base + k0);
        tma_load_2d(v_s + a * (128 * 128), &map_res, res_full, 2 * p.E + head * D + a * 32, row_base + k0);
      }
      for (int it = 0; it < n_it; ++it) {
        const int qrow = row_base + block_of(it) * 64;
        mbar_wait(blk_empty, (uint32_t)(it & 1) ^ 1u);
        mbar_expect_tx(blk_full, 2 * Cfg::kBlkBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a) {
          tma_load_2d(qk_s + a * (64 * 128), &map_qk, blk_full, head * D + a * 32, qrow);
          tma_load_2d(dok_s + a * (64 * 128), &map_dok, blk_full, head * D + a * 32, qrow);
        }
        mbar_wait(mn_empty, (uint32_t)(it & 1) ^ 1u);
        mbar_expect_tx(mn_full, 2 * Cfg::kBlkBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a) {
          tma_load_2d(qmn_s + a * (64 * 128), &map_qmn, mn_full, head * D + a * 32, qrow);
          tma_load_2d(domn_s + a * (64 * 128), &map_domn, mn_full, head * D + a * 32, qrow);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_a = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(res_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        mbar_wait(blk_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 8; ++kk) {      // S^T = K Q^T ; dP^T = V dO^T
          const uint32_t ao = (kk >> 2) * (128 * 128) + (kk & 3) * 32, bo = (kk >> 2) * (64 * 128) + (kk & 3) * 32;
          umma_tf32(t_st, make_desc(k_s + ao, 16, 1024, 2), make_desc(qk_s + bo, 16, 1024, 2), idesc_t, kk > 0 ? 1u : 0u);
          umma_tf32(t_dpt, make_desc(v_s + ao, 16, 1024, 2), make_desc(dok_s + bo, 16, 1024, 2), idesc_t, kk > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full);
        umma_commit(blk_empty);                    // K-major copies free once S^T / dP^T retire
        mbar_wait(mn_full, ph);
        mbar_wait(pds_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {          // dV += P^T dO ; dK += dS^T Q   (K = 64 queries)
          const uint32_t ao = (kk >> 2) * (128 * 128) + (kk & 3) * 32;
          const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
          umma_tf32(t_dv, make_desc(pt_s + ao, 16, 1024, 2), make_desc(domn_s + kk * 1024, 64 * 128, 512, 1), idesc_a, acc);
          umma_tf32(t_dk, make_desc(dst_s + ao, 16, 1024, 2), make_desc(qmn_s + kk * 1024, 64 * 128, 512, 1), idesc_a, acc);
        }
        umma_commit(mn_empty);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                    // key row inside the tile == TMEM lane
    const int key = k0 + r;
    const int ctid = threadIdx.x - 64;              // 0..127 among the compute threads
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const bool key_ok = key < p.S;
    const bool key_pad = key_ok ? (x[(int64_t)b * p.S + key] == 0) : true;
    const int64_t so = ((int64_t)b * p.h + head) * p.S;
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int q0 = block_of(it) * 64;
      // stage lse / delta of the 64 queries (natural-log lse -> exp2 domain)
      asm volatile("bar.sync 1, 128;" ::: "memory");          // previous iteration finished reading them
      if (ctid < 64) lse_s[ctid] = (q0 + ctid < p.S) ? lse[so + q0 + ctid] * kLog2e : 0.f;
      else delta_s[ctid - 64] = (q0 + ctid - 64 < p.S) ? delta[so + q0 + ctid - 64] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // visibility of (query q0+c, this key): causal & not PAD, or the query is a fully-masked row
      const unsigned long long range_bits = low_bits(p.S - q0);
      const unsigned long long causal = (key_ok && !key_pad) ? ~low_bits(key - q0) : 0ull;
      const unsigned long long degq = key_ok ? low_bits(first_real - q0) : 0ull;
      const unsigned long long vis = (causal | degq) & range_bits;
      mbar_wait(sdp_full, ph);
      tc_fence_after();
      float pt[64], dst[64];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(t_st + lane_addr + half * 32, sv);
        tmem_ld_32x32(t_dpt + lane_addr + half * 32, dv);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int cc = half * 32 + c;
          const float pv = ((vis >> cc) & 1ull) ? fast_exp2(__uint_as_float(sv[c]) * p.scale_log2 - lse_s[cc]) : 0.f;
          pt[cc] = pv;
          dst[cc] = pv * (__uint_as_float(dv[c]) - delta_s[cc]);
        }
      }
      store_a_row(pt_ptr, r, pt);
      store_a_row(dst_ptr, r, dst);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(pds_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    {
      // tcgen05.ld is warp-collective: every lane loads, only rows inside the sequence store
      const float scale = p.scale_log2 * (1.f / kLog2e);
      float* out = dqkv + ((int64_t)b * p.S + key) * 3 * p.E + head * D;
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc) {
        uint32_t a[32], c2[32];
        tmem_ld_32x32(t_dv + lane_addr + cc * 32, a);
        tmem_ld_32x32(t_dk + lane_addr + cc * 32, c2);
        if (key_ok) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<float4*>(out + 2 * p.E + cc * 32 + 4 * c) =
                make_float4(__uint_as_float(a[4 * c]), __uint_as_float(a[4 * c + 1]), __uint_as_float(a[4 * c + 2]), __uint_as_float(a[4 * c + 3]));
            *reinterpret_cast<float4*>(out + p.E + cc * 32 + 4 * c) =
                make_float4(__uint_as_float(c2[4 * c]) * scale, __uint_as_float(c2[4 * c + 1]) * scale,
                            __uint_as_float(c2[4 * c + 2]) * scale, __uint_as_float(c2[4 * c + 3]) * scale);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::kTmemCols);
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_dq_tc(const __grid_constant__ CUtensorMap map_qres,   // qkv, box {32,128}, K-major
                                                            const __grid_constant__ CUtensorMap map_dores,  // dctx, box {32,128}, K-major
                                                            const __grid_constant__ CUtensorMap map_kk,     // qkv, box {32,64}, K-major
                                                            const __grid_constant__ CUtensorMap map_kmn,    // qkv, box {32,64}, MN-major
                                                            const Dims p, const int32_t* __restrict__ x,
                                                            const int32_t* __restrict__ fnp,
                                                            const float* __restrict__ lse,
                                                            const float* __restrict__ delta,
                                                            float* __restrict__ dqkv) {
  using Cfg = BwdCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t q_s = base, do_s = q_s + Cfg::kResBytes;
  const uint32_t kk_s = do_s + Cfg::kResBytes, vk_s = kk_s + Cfg::kBlkBytes, kmn_s = vk_s + Cfg::kBlkBytes;
  const uint32_t ds_s = kmn_s + Cfg::kBlkBytes;
  uint8_t* ds_ptr = base_ptr + 2 * Cfg::kResBytes + 3 * Cfg::kBlkBytes;
  const uint32_t bar = ds_s + Cfg::kPBytes;
  uint8_t* bar_ptr = ds_ptr + Cfg::kPBytes;
  const uint32_t res_full = bar, blk_full = bar + 8, blk_empty = bar + 16, sdp_full = bar + 24, ds_full = bar + 32,
                 acc_full = bar + 40, mn_full = bar + 48, mn_empty = bar + 56, tmem_slot = bar + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bar_ptr + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * BQ;
  const int first_real = fnp[b];
  const bool deg_block = q0 < first_real;
  const int n_all = (p.S + BKV - 1) / BKV;
  const int n_kv = deg_block ? n_all : min(n_all, (q0 + BQ - 1) / BKV + 1);
  const int row_base = b * p.S;

  if (warp == 0 && lane == 0) {
    mbar_init(res_full, 1); mbar_init(blk_full, 1); mbar_init(blk_empty, 1); mbar_init(sdp_full, 1);
    mbar_init(ds_full, 128); mbar_init(acc_full, 1); mbar_init(mn_full, 1); mbar_init(mn_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t t_s = tmem, t_dp = tmem + 64, t_dq = tmem + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(res_full, 2 * Cfg::kResBytes);
#pragma unroll
      for (int a = 0; a < Cfg::kAtoms; ++a) {
        tma_load_2d(q_s + a * (128 * 128), &map_qres, res_full, head * D + a * 32, row_base + q0);
        tma_load_2d(do_s + a * (128 * 128), &map_dores, res_full, head * D + a * 32, row_base + q0);
      }
      for (int kb = 0; kb < n_kv; ++kb) {
        const int krow = row_base + kb * BKV;
        mbar_wait(blk_empty, (uint32_t)(kb & 1) ^ 1u);
        mbar_expect_tx(blk_full, 2 * Cfg::kBlkBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a) {
          tma_load_2d(kk_s + a * (64 * 128), &map_kk, blk_full, p.E + head * D + a * 32, krow);
          tma_load_2d(vk_s + a * (64 * 128), &map_kk, blk_full, 2 * p.E + head * D + a * 32, krow);
        }
        mbar_wait(mn_empty, (uint32_t)(kb & 1) ^ 1u);
        mbar_expect_tx(mn_full, Cfg::kBlkBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_2d(kmn_s + a * (64 * 128), &map_kmn, mn_full, p.E + head * D + a * 32, krow);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_a = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(res_full, 0);
      for (int kb = 0; kb < n_kv; ++kb) {
        const uint32_t ph = (uint32_t)(kb & 1);
        mbar_wait(blk_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 8; ++kk) {      // S = Q K^T ; dP = dO V^T
          const uint32_t ao = (kk >> 2) * (128 * 128) + (kk & 3) * 32, bo = (kk >> 2) * (64 * 128) + (kk & 3) * 32;
          umma_tf32(t_s, make_desc(q_s + ao, 16, 1024, 2), make_desc(kk_s + bo, 16, 1024, 2), idesc_t, kk > 0 ? 1u : 0u);
          umma_tf32(t_dp, make_desc(do_s + ao, 16, 1024, 2), make_desc(vk_s + bo, 16, 1024, 2), idesc_t, kk > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full);
        umma_commit(blk_empty);                    // K-major K/V copies free once S / dP retire
        mbar_wait(mn_full, ph);
        mbar_wait(ds_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {          // dQ += dS K   (K = 64 keys)
          const uint32_t ao = (kk >> 2) * (128 * 128) + (kk & 3) * 32;
          umma_tf32(t_dq, make_desc(ds_s + ao, 16, 1024, 2), make_desc(kmn_s + kk * 1024, 64 * 128, 512, 1), idesc_a,
                    (kb > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(mn_empty);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int row = q0 + r;
    const bool deg = row < first_real;
    const bool row_ok = row < p.S;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int64_t so = ((int64_t)b * p.h + head) * p.S;
    const float ls = row_ok ? lse[so + row] * kLog2e : 0.f;
    const float dl = row_ok ? delta[so + row] : 0.f;
    for (int kb = 0; kb < n_kv; ++kb) {
      const uint32_t ph = (uint32_t)(kb & 1);
      const int k0 = kb * BKV;
      const int ka = k0 + lane, kb2 = k0 + 32 + lane;
      const bool pa = ka < p.S ? (x[(int64_t)b * p.S + ka] == 0) : true;
      const bool pb = kb2 < p.S ? (x[(int64_t)b * p.S + kb2] == 0) : true;
      const unsigned long long padbits = (unsigned long long)__ballot_sync(0xffffffffu, pa) |
                                         ((unsigned long long)__ballot_sync(0xffffffffu, pb) << 32);
      const unsigned long long range_bits = low_bits(p.S - k0);
      unsigned long long vis;
      if (!row_ok) vis = 0ull;
      else if (deg) vis = range_bits;
      else vis = low_bits(row - k0 + 1) & ~padbits & range_bits;
      mbar_wait(sdp_full, ph);
      tc_fence_after();
      float ds[64];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(t_s + lane_addr + half * 32, sv);
        tmem_ld_32x32(t_dp + lane_addr + half * 32, dv);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int cc = half * 32 + c;
          const float pv = ((vis >> cc) & 1ull) ? fast_exp2(__uint_as_float(sv[c]) * p.scale_log2 - ls) : 0.f;
          ds[cc] = pv * (__uint_as_float(dv[c]) - dl);
        }
      }
      store_a_row(ds_ptr, r, ds);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const float scale = p.scale_log2 * (1.f / kLog2e);
    float* out = dqkv + ((int64_t)b * p.S + row) * 3 * p.E + head * D;
#pragma unroll
    for (int cc = 0; cc < D / 32; ++cc) {
      uint32_t a[32];
      tmem_ld_32x32(t_dq + lane_addr + cc * 32, a);
      if (row_ok) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(out + cc * 32 + 4 * c) =
              make_float4(__uint_as_float(a[4 * c]) * scale, __uint_as_float(a[4 * c + 1]) * scale,
                          __uint_as_float(a[4 * c + 2]) * scale, __uint_as_float(a[4 * c + 3]) * scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::kTmemCols);
  }
}

template <int D>
int launch_bwd(cudaStream_t st, int B, int S, int h, const float* qkv, const int32_t* x, const int32_t* fnp,
               const float* dctx, const float* lse, const float* delta, float* dqkv) {
  using Cfg = BwdCfg<D>;
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
    NLPS_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_tc<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemDkdv));
    NLPS_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemDq));
    once = true;
  }
  Dims p{B, S, h, E, (float)(1.0 / std::sqrt((double)D)) * kLog2e, 0};
  dim3 grid((unsigned)ceil_div(S, 128), (unsigned)h, (unsigned)B);
  attn_bwd_dkdv_tc<D><<<grid, kThreads, Cfg::kSmemDkdv, st>>>(res, qk, qmn, dok, domn, p, x, fnp, lse, delta, dqkv);
  NLPS_LAUNCH_CHECK();
  attn_bwd_dq_tc<D><<<grid, kThreads, Cfg::kSmemDq, st>>>(res, dores, qk, qmn, p, x, fnp, lse, delta, dqkv);
  NLPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool attention_tc_supported(int d, int E, const float* qkv) {
  return (d == 32 || d == 64) && (E % 4 == 0) && ((reinterpret_cast<uintptr_t>(qkv) & 15u) == 0);
}

int attention_forward_tc(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                         const int32_t* fnp, float* ctx, float* lse) {
  if (d == 64) return launch_fwd<64>(st, B, S, h, qkv, x, fnp, ctx, lse);
  return launch_fwd<32>(st, B, S, h, qkv, x, fnp, ctx, lse);
}

int attention_backward_tc(cudaStream_t st, int B, int S, int h, int d, const float* qkv, const int32_t* x,
                          const int32_t* fnp, const float* dctx, const float* lse, const float* delta, float* dqkv) {
  if (d == 64) return launch_bwd<64>(st, B, S, h, qkv, x, fnp, dctx, lse, delta, dqkv);
  return launch_bwd<32>(st, B, S, h, qkv, x, fnp, dctx, lse, delta, dqkv);
}

}  // namespace nlps
