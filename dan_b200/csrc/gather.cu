// The one exchange step of the path: every rank's NMS kernel has written its zero padded keep lists and
// per-image counts into ONE fixed-capacity slab; dan_gather_detections moves all slabs with one ncclAllGather ENQUEUED
// ON THE CALLER'S STREAM right after the NMS kernel, so the collective is part of the step's CUDA graph and costs no host
// time per step.  The reference has no collective to mirror (in-graph towers, tf_replicate_model_fn.py:458-501).
//
// NCCL is bound at run time (dlopen): inside a PyTorch process the already loaded libnccl.so.2 is reused, so there is one
// NCCL in the process; the library itself has no link-time dependency on NCCL and loads on machines without it.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace dan {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
// ncclConfig_t as of NCCL 2.18 (size, magic, version first: NCCL copies `size` bytes over its own defaults, so an older,
// shorter layout stays valid with newer libraries)
struct NcclConfig218 {
  size_t size;
  unsigned int magic;
  unsigned int version;
  int blocking;
  int cgaClusterSize;
  int minCTAs;
  int maxCTAs;
  const char* netName;
  int splitShare;
};
typedef int (*CommInitRankConfigFn)(NcclComm*, int, NcclUniqueId, int, NcclConfig218*);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);

struct NcclApi {
  void* handle;
  GetUniqueIdFn get_unique_id;
  CommInitRankFn comm_init_rank;
  CommInitRankConfigFn comm_init_rank_config;
  CommDestroyFn comm_destroy;
  AllGatherFn all_gather;
  GetErrorStringFn error_string;
  GetVersionFn get_version;
};

static NcclApi g_nccl = {};

static int nccl_bind(const char* path) {
  if (g_nccl.handle != nullptr) return DAN_OK;
  void* h = nullptr;
  if (path != nullptr && path[0] != 0) {
    h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  } else {
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);            // the copy the process already uses (PyTorch's)
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  }
  DAN_REQUIRE(h != nullptr, DAN_ERR_UNSUPPORTED, "cannot load NCCL (%s): %s", path && path[0] ? path : "libnccl.so.2", dlerror());
  NcclApi a = {};
  a.handle = h;
  a.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  a.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  a.comm_init_rank_config = (CommInitRankConfigFn)dlsym(h, "ncclCommInitRankConfig");     // NCCL >= 2.14 (optional)
  a.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  a.all_gather = (AllGatherFn)dlsym(h, "ncclAllGather");
  a.error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
  a.get_version = (GetVersionFn)dlsym(h, "ncclGetVersion");
  DAN_REQUIRE(a.get_unique_id && a.comm_init_rank && a.comm_destroy && a.all_gather && a.error_string, DAN_ERR_UNSUPPORTED,
              "the loaded NCCL lacks a required symbol");
  g_nccl = a;
  return DAN_OK;
}

static int nccl_fail(int rc, const char* what) {
  set_error("NCCL error %d (%s) at %s", rc, g_nccl.error_string ? g_nccl.error_string(rc) : "?", what);
  return DAN_ERR_CUDA;
}

}  // namespace dan

using namespace dan;

extern "C" {

int dan_nccl_load(const char* path) { return nccl_bind(path); }

int dan_nccl_version(void) {
  if (nccl_bind(nullptr) != DAN_OK || g_nccl.get_version == nullptr) return 0;
  int v = 0;
  return g_nccl.get_version(&v) == 0 ? v : 0;
}

int dan_comm_unique_id(void* out_id128) {
  DAN_REQUIRE(out_id128 != nullptr, DAN_ERR_INVALID_ARGUMENT, "out_id128 is NULL");
  int rc = nccl_bind(nullptr);
  if (rc != DAN_OK) return rc;
  NcclUniqueId id;
  const int nrc = g_nccl.get_unique_id(&id);
  if (nrc != 0) return nccl_fail(nrc, "ncclGetUniqueId");
  memcpy(out_id128, &id, sizeof(id));
  return DAN_OK;
}

int dan_comm_init_ctas(const void* id128, int32_t rank, int32_t world_size, int32_t max_ctas, void** out_comm) {
  DAN_REQUIRE(id128 != nullptr && out_comm != nullptr, DAN_ERR_INVALID_ARGUMENT, "NULL pointer");
  DAN_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, DAN_ERR_INVALID_ARGUMENT, "rank %d of %d", rank, world_size);
  int rc = nccl_bind(nullptr);
  if (rc != DAN_OK) return rc;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm comm = nullptr;
  if (max_ctas > 0 && g_nccl.comm_init_rank_config != nullptr) {
    NcclConfig218 cfg;
    cfg.size = sizeof(NcclConfig218);
    cfg.magic = 0xcafebeefu;
    cfg.version = 21800;                     // NCCL_VERSION(2, 18, 0)
    cfg.blocking = cfg.cgaClusterSize = cfg.splitShare = (int)0x80000000;      // NCCL_CONFIG_UNDEF_INT
    cfg.netName = nullptr;
    cfg.minCTAs = 1;
    cfg.maxCTAs = max_ctas;
    const int nrc = g_nccl.comm_init_rank_config(&comm, world_size, id, rank, &cfg);
    if (nrc != 0) return nccl_fail(nrc, "ncclCommInitRankConfig");
  } else {
    const int nrc = g_nccl.comm_init_rank(&comm, world_size, id, rank);
    if (nrc != 0) return nccl_fail(nrc, "ncclCommInitRank");
  }
  *out_comm = comm;
  return DAN_OK;
}

// NCCL's own choice of CTAs.  (Measured at 8 ranks x 8 steps in flight: with max_ctas = 1 a 480 KB all-gather takes ~570 us and
// the step rate drops 3x; with NCCL's default the CTAs of the collectives hold SMs while they wait for the slowest rank,
// 67 vs 47 us per step.  The peer exchange of dan_postprocess_batch_peers has neither cost.)
int dan_comm_init(const void* id128, int32_t rank, int32_t world_size, void** out_comm) {
  return dan_comm_init_ctas(id128, rank, world_size, 0, out_comm);
}

int dan_comm_destroy(void* comm) {
  if (comm == nullptr) return DAN_OK;
  DAN_REQUIRE(g_nccl.handle != nullptr, DAN_ERR_INVALID_ARGUMENT, "NCCL is not loaded");
  const int nrc = g_nccl.comm_destroy(comm);
  return nrc == 0 ? DAN_OK : nccl_fail(nrc, "ncclCommDestroy");
}

int dan_gather_detections(void* comm, const void* send_slab, void* recv_slabs, size_t slab_bytes, void* stream) {
  DAN_REQUIRE(comm != nullptr, DAN_ERR_INVALID_ARGUMENT, "comm is NULL");
  DAN_REQUIRE(send_slab != nullptr && recv_slabs != nullptr, DAN_ERR_INVALID_ARGUMENT, "NULL slab pointer");
  DAN_REQUIRE(g_nccl.handle != nullptr, DAN_ERR_INVALID_ARGUMENT, "NCCL is not loaded");
  if (slab_bytes == 0) return DAN_OK;
  const int nrc = g_nccl.all_gather(send_slab, recv_slabs, slab_bytes, /*ncclInt8*/ 0, comm, (cudaStream_t)stream);
  return nrc == 0 ? DAN_OK : nccl_fail(nrc, "ncclAllGather");
}

}  // extern "C"
