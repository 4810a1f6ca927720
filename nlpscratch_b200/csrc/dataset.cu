// Dataset construction on the device: the step in front of the training loop.
//   prefix_dataset  : generate_sequence_data / generate_bpe_sequence_data (Transformer/transformer_train.py:47-60,
//                     transformer_train_bpe.py:52-65) with pad_sequence (transformer.py:575-582): every proper prefix
//                     ids[:i] of every sentence becomes one right-padded row of X (the LAST max_len ids when the
//                     prefix is longer), Y holds ids[i].  Integer work, bit-exact; bound by the HBM writes of X.
//   expand_tokens   : BPEEncoder.encode_ids over a batch (bpe_encoder.py:225-249): word-type ids -> subword ids
//                     through the per-type table the host builds once (the reference's own _cache), plus BOS/EOS.
#include "common.cuh"
#include <algorithm>

namespace nlps {
namespace {

// largest s in [0, n) with off[s] <= v   (off is non-decreasing, off[0] == 0 <= v)
__device__ __forceinline__ int upper_row(const int64_t* __restrict__ off, int n, int64_t v) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= v) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// one warp per output row; 128-bit stores when max_len % 4 == 0
__global__ void __launch_bounds__(256) prefix_rows_kernel(int n_sent, int64_t n_rows, int max_len, int32_t pad,
                                                         const int32_t* __restrict__ ids, const int64_t* __restrict__ sent_off,
                                                         const int64_t* __restrict__ row_off, int32_t* __restrict__ X,
                                                         int32_t* __restrict__ Y, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
    const int s = upper_row(row_off, n_sent, row);
    const int64_t i = row - row_off[s] + 1;            // prefix length, 1 .. len-1
    const int32_t* src = ids + sent_off[s];
    const int64_t skip = i > max_len ? i - max_len : 0; // keep the most recent context
    const int64_t keep = i - skip;
    int32_t* dst = X + row * max_len;
    if (vec) {
      for (int c = lane * 4; c < max_len; c += 128) {
        int4 v;
        v.x = c < keep ? src[skip + c] : pad;
        v.y = c + 1 < keep ? src[skip + c + 1] : pad;
        v.z = c + 2 < keep ? src[skip + c + 2] : pad;
        v.w = c + 3 < keep ? src[skip + c + 3] : pad;
        *reinterpret_cast<int4*>(dst + c) = v;
      }
    } else {
      for (int c = lane; c < max_len; c += 32) dst[c] = c < keep ? src[skip + c] : pad;
    }
    if (lane == 0) Y[row] = src[i];
  }
}

// per-sentence output length of encode_ids: sum of the table lengths of its words (+BOS, +EOS)
__global__ void __launch_bounds__(256) expand_len_kernel(int n_sent, const int32_t* __restrict__ words,
                                                        const int64_t* __restrict__ sent_off, const int32_t* __restrict__ tab_off,
                                                        int extra, int32_t* __restrict__ out_len) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_sent; s += warps) {
    int32_t acc = 0;
    for (int64_t w = sent_off[s] + lane; w < sent_off[s + 1]; w += 32) {
      const int32_t t = words[w];
      acc += tab_off[t + 1] - tab_off[t];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out_len[s] = acc + extra;
  }
}

// exclusive scan int32 -> int64 offsets (n+1 entries), single CTA (sentence counts are small)
__global__ void __launch_bounds__(1024) scan64_kernel(int n, const int32_t* __restrict__ len, int64_t* __restrict__ off) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = i < n ? len[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int64_t before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - v;
    if (i < n) off[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = carry;
}

// one warp per sentence: words in order, each word's subword ids copied by the lanes
__global__ void __launch_bounds__(256) expand_fill_kernel(int n_sent, const int32_t* __restrict__ words,
                                                         const int64_t* __restrict__ sent_off, const int32_t* __restrict__ tab_off,
                                                         const int32_t* __restrict__ tab_ids, int32_t bos, int32_t eos,
                                                         const int64_t* __restrict__ out_off, int32_t* __restrict__ out_ids) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_sent; s += warps) {
    int64_t pos = out_off[s];
    if (bos >= 0) { if (lane == 0) out_ids[pos] = bos; ++pos; }
    for (int64_t w0 = sent_off[s]; w0 < sent_off[s + 1]; w0 += 32) {
      // lane l owns word w0+l: its length, then a warp scan gives every word's position
      const int64_t w = w0 + lane;
      const int32_t t = w < sent_off[s + 1] ? words[w] : -1;
      const int32_t b = t >= 0 ? tab_off[t] : 0, n = t >= 0 ? tab_off[t + 1] - b : 0;
      int32_t incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int64_t mine = pos + incl - n;
      for (int k = 0; k < n; ++k) out_ids[mine + k] = tab_ids[b + k];     // subword lists are 1-10 ids long
      pos += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (eos >= 0 && lane == 0) out_ids[pos] = eos;
  }
}

}  // namespace

int prefix_dataset(cudaStream_t st, int n_sent, int64_t n_rows, int max_len, int32_t pad, const int32_t* ids,
                   const int64_t* sent_off, const int64_t* row_off, int32_t* X, int32_t* Y) {
  if (n_rows <= 0) return 0;
  NLPS_REQUIRE(n_sent > 0 && max_len > 0, "prefix_dataset: n_sent=%d max_len=%d", n_sent, max_len);
  const int vec = (max_len % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0) ? 1 : 0;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)kNumSMs * 8);
  prefix_rows_kernel<<<grid, 256, 0, st>>>(n_sent, n_rows, max_len, pad, ids, sent_off, row_off, X, Y, vec);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int expand_tokens_offsets(cudaStream_t st, int n_sent, const int32_t* words, const int64_t* sent_off, const int32_t* tab_off,
                          int32_t bos, int32_t eos, int32_t* len_scratch, int64_t* out_off) {
  NLPS_REQUIRE(n_sent > 0, "expand_tokens: n_sent=%d", n_sent);
  const int extra = (bos >= 0 ? 1 : 0) + (eos >= 0 ? 1 : 0);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n_sent, 8), (int64_t)kNumSMs * 8);
  expand_len_kernel<<<grid, 256, 0, st>>>(n_sent, words, sent_off, tab_off, extra, len_scratch);
  NLPS_LAUNCH_CHECK();
  scan64_kernel<<<1, 1024, 0, st>>>(n_sent, len_scratch, out_off);
  NLPS_LAUNCH_CHECK();
  return 0;
}

int expand_tokens_fill(cudaStream_t st, int n_sent, const int32_t* words, const int64_t* sent_off, const int32_t* tab_off,
                       const int32_t* tab_ids, int32_t bos, int32_t eos, const int64_t* out_off, int32_t* out_ids) {
  NLPS_REQUIRE(n_sent > 0, "expand_tokens: n_sent=%d", n_sent);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n_sent, 8), (int64_t)kNumSMs * 8);
  expand_fill_kernel<<<grid, 256, 0, st>>>(n_sent, words, sent_off, tab_off, tab_ids, bos, eos, out_off, out_ids);
  NLPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace nlps
