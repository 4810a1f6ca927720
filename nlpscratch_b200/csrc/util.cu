#include "common.cuh"
#include <atomic>
namespace nlps {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count(bool reset) {
  return reset ? g_launches.exchange(0, std::memory_order_relaxed) : g_launches.load(std::memory_order_relaxed);
}

// which implementation served a call (tests assert that the tensor-core kernels -- not a fallback -- ran)
static std::atomic<long long> g_kinds[KIND_COUNT];
void count_kind(int kind) { if (kind >= 0 && kind < KIND_COUNT) g_kinds[kind].fetch_add(1, std::memory_order_relaxed); }
long long kind_count(int kind, bool reset) {
  if (kind < 0 || kind >= KIND_COUNT) return 0;
  return reset ? g_kinds[kind].exchange(0, std::memory_order_relaxed) : g_kinds[kind].load(std::memory_order_relaxed);
}

// Bumped whenever a device buffer that a captured CUDA graph may point into is re-allocated (split-K workspace,
// model workspaces, activation caches): graphs captured under an older epoch are dropped instead of replayed.
static std::atomic<unsigned long long> g_alloc_epoch{1};
unsigned long long alloc_epoch() { return g_alloc_epoch.load(std::memory_order_relaxed); }
void bump_alloc_epoch() { g_alloc_epoch.fetch_add(1, std::memory_order_relaxed); }

// cudaMallocAsync scratch (split-K workspaces, 3xTF32 splits) must not go back to the driver at every
// synchronisation: keep the default pool's memory cached.
void ensure_mempool_config() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[dev] = true;
}
}  // namespace nlps
