// sd_nccl.cpp -- the communicator behind sd_comm: NCCL, dlopen'ed on first use so that libsnappygpu.so itself
// loads on a box without NCCL (and picks up the copy a host process such as PyTorch has already loaded).
//
// Role in the reference: the Exchange between partial and final aggregation that SnappyStrategies plans
// (core/.../SnappyStrategies.scala:566-604) and that CollectAggregateExec short-cuts on the driver
// (core/.../aggregate/CollectAggregateExec.scala:67-121).  Here: ONE ncclAllGather of every partition's
// partial rows over NVLink / NVSwitch; the merge happens on every rank.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "sd_host.h"

namespace sd {

namespace {
// the handful of NCCL entry points used, declared here (stable C ABI since NCCL 2.x) to avoid a build dependency
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
constexpr int kNcclUint8 = 1;   // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_p) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_p, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("SD_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { if (nm && *nm && (n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break; }
    if (!n.lib) { n.why = "libnccl.so.2 not found (set SD_NCCL_LIB)"; return; }
#define NC(field, sym) *(void**)(&n.field) = dlsym(n.lib, sym); if (!n.field) { n.why = sym " missing"; return; }
    NC(GetUniqueId, "ncclGetUniqueId") NC(CommInitRank, "ncclCommInitRank") NC(CommDestroy, "ncclCommDestroy")
    NC(AllGather, "ncclAllGather") NC(GetErrorString, "ncclGetErrorString")
#undef NC
    n.ok = true;
  });
  return n;
}
int nccl_fail(const char* what, int rc) {
  return set_error(SD_ERR_CUDA, "%s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
}
}  // namespace

int comm_unique_id(void* out128) {
  Nccl& n = nccl();
  if (!n.ok) return set_error(SD_ERR_CUDA, "NCCL unavailable: %s", n.why.c_str());
  ncclUniqueId_t id;
  int rc = n.GetUniqueId(&id);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(out128, &id, 128);
  return 0;
}

int comm_init(const void* id128, int rank, int world, void** out) {
  Nccl& n = nccl();
  if (!n.ok) return set_error(SD_ERR_CUDA, "NCCL unavailable: %s", n.why.c_str());
  ncclUniqueId_t id;
  memcpy(&id, id128, 128);
  ncclComm_p c = nullptr;
  int rc = n.CommInitRank(&c, world, id, rank);
  if (rc) return nccl_fail("ncclCommInitRank", rc);
  *out = c;
  return 0;
}

void comm_destroy(void* c) { if (c && nccl().ok) nccl().CommDestroy(c); }

int comm_all_gather_bytes(void* c, const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t stream) {
  int rc = nccl().AllGather(d_send, d_recv, bytes_per_rank, kNcclUint8, c, stream);
  if (rc) return nccl_fail("ncclAllGather", rc);
  return 0;
}

}  // namespace sd
