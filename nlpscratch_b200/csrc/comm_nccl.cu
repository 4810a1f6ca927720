// NCCL owned by the C library (SURVEY 8b `nlps_comm_init`): the data-parallel gradient exchange of the training step.
//
// The reference has no distributed code; under data parallelism the step needs exactly one exchange -- the sum of the flat
// gradient buffer over ranks.  libnccl is bound at run time with dlopen/dlsym (prototypes restated below from
// NCCL's public, ABI-stable C interface) so that libnlps_b200.so keeps loading on machines without NCCL or without a GPU
// (the CPU tests check every symbol) and so that a process that has already loaded an NCCL (torch's bundled one) shares it.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>

namespace nlps {
namespace {

typedef struct { char internal[128]; } NcclUniqueId;     // NCCL_UNIQUE_ID_BYTES
typedef void* NcclComm;
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

struct NcclApi {
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*CommCount)(NcclComm, int*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* handle = nullptr;
  bool ok = false;
};

NcclApi* api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return;
#define NLPS_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym))
    NLPS_SYM(GetVersion, "ncclGetVersion");
    NLPS_SYM(GetUniqueId, "ncclGetUniqueId");
    NLPS_SYM(CommInitRank, "ncclCommInitRank");
    NLPS_SYM(CommDestroy, "ncclCommDestroy");
    NLPS_SYM(CommCount, "ncclCommCount");
    NLPS_SYM(AllReduce, "ncclAllReduce");
    NLPS_SYM(Broadcast, "ncclBroadcast");
    NLPS_SYM(GetErrorString, "ncclGetErrorString");
#undef NLPS_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Broadcast && a.GetErrorString;
  });
  return &a;
}

#define NLPS_NCCL(call)                                                                                   \
  do {                                                                                                    \
    const int r_ = (call);                                                                                \
    if (r_ != 0) {                                                                                        \
      ::nlps::set_error("%s failed: %s (%s:%d)", #call, api()->GetErrorString(r_), __FILE__, __LINE__); \
      return 1;                                                                                           \
    }                                                                                                     \
  } while (0)

}  // namespace

int comm_available(int* version) {
  NcclApi* a = api();
  if (version) { *version = 0; if (a->ok && a->GetVersion) a->GetVersion(version); }
  return a->ok ? 1 : 0;
}

int comm_unique_id(void* out128) {
  NcclApi* a = api();
  NLPS_REQUIRE(a->ok, "libnccl.so.2 could not be loaded: no data-parallel exchange on this machine");
  NcclUniqueId id;
  NLPS_NCCL(a->GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int comm_create(const void* id128, int rank, int world, void** comm_out) {
  NcclApi* a = api();
  NLPS_REQUIRE(a->ok, "libnccl.so.2 could not be loaded: no data-parallel exchange on this machine");
  NLPS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_create: rank %d of %d", rank, world);
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm c = nullptr;
  NLPS_NCCL(a->CommInitRank(&c, world, id, rank));
  int n = 0;
  if (a->CommCount) NLPS_NCCL(a->CommCount(c, &n));
  NLPS_REQUIRE(!a->CommCount || n == world, "communicator has %d ranks, expected %d", n, world);
  *comm_out = c;
  return 0;
}

int comm_destroy(void* comm) {
  if (comm) NLPS_NCCL(api()->CommDestroy(comm));
  return 0;
}

int comm_allreduce_sum(void* comm, float* buf, size_t count, cudaStream_t st) {
  if (count == 0) return 0;
  NLPS_NCCL(api()->AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, comm, st));
  count_launch();
  return 0;
}

int comm_broadcast(void* comm, float* buf, size_t count, int root, cudaStream_t st) {
  if (count == 0) return 0;
  NLPS_NCCL(api()->Broadcast(buf, buf, count, kNcclFloat32, root, comm, st));
  count_launch();
  return 0;
}

}  // namespace nlps
