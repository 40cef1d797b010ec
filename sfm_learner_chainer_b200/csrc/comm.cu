// comm.cu -- the one collective of the sharded path: the all-reduce of the five loss partials (SURVEY 8(e);
// the reference analogue is Chainer's MultiprocessParallelUpdater reduce, config_utils.py:123-126).  Two forms:
//   SfmPeer  (bottom of this file) the sum is done INSIDE the epilogue kernel over NVLink peer memory (CUDA IPC mapped
//            slot arrays, plain peer stores + flags; fused_loss.cu: epilogue_losses) -- no extra launch;
//   SfmComm  an NCCL all-reduce enqueued behind the step (the library-call baseline).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy the process already holds, e.g. the one bundled
// with torch, or the one named with sfm_nccl_set_library) so that libsfmloss.so itself has no link-time NCCL
// dependency and still loads on a machine without it.  Only the five entry points below are used, with the
// enum values of nccl.h 2.x (ncclFloat = 7, ncclSum = 0, ncclUniqueId = 128 opaque bytes).
// The all-reduce is enqueued on the caller's stream and is capturable in a CUDA graph, so the per-step call
// costs no host work when the step is replayed as a graph.
#include <dlfcn.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>

#include "common.cuh"
#include "kernels.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[SFM_NCCL_UNIQUE_ID_BYTES]; } ncclUniqueId;
typedef int ncclResult_t;

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;
std::string g_nccl_path;

int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return 0;
  void* h = nullptr;
  if (!g_nccl_path.empty()) h = dlopen(g_nccl_path.c_str(), RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy this process already mapped (torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    sfm_set_error("NCCL is not available: %s (name the library with sfm_nccl_set_library)", dlerror());
    return SFM_E_UNSUPPORTED;
  }
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString) {
    sfm_set_error("libnccl is missing one of ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllReduce");
    return SFM_E_UNSUPPORTED;
  }
  g_nccl = a;
  return 0;
}

int nccl_check(ncclResult_t r, const char* what) {
  if (r == 0) return 0;
  sfm_set_error("%s failed: NCCL error %d (%s)", what, (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return SFM_E_COMM;
}

}  // namespace

struct SfmComm {
  ncclComm_t comm;
  int nranks, rank;
};

extern "C" int sfm_nccl_set_library(const char* path) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) { sfm_set_error("sfm_nccl_set_library: NCCL is already loaded"); return SFM_E_UNSUPPORTED; }
  g_nccl_path = path ? path : "";
  return 0;
}

extern "C" int sfm_nccl_version(void) {
  if (nccl_load()) return 0;
  int v = 0;
  if (g_nccl.GetVersion && g_nccl.GetVersion(&v) == 0) return v;
  return 0;
}

extern "C" int sfm_comm_unique_id(void* id_out) {
  if (!id_out) { sfm_set_error("sfm_comm_unique_id: id_out is NULL"); return SFM_E_NULL_POINTER; }
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  if ((rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId"))) return rc;
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

extern "C" int sfm_comm_create(const void* unique_id, int nranks, int rank, SfmComm** comm_out) {
  if (!unique_id || !comm_out) { sfm_set_error("sfm_comm_create: null pointer"); return SFM_E_NULL_POINTER; }
  if (nranks < 1 || rank < 0 || rank >= nranks) { sfm_set_error("sfm_comm_create: rank %d outside a world of %d", rank, nranks); return SFM_E_INVALID_DESC; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { sfm_set_error("no CUDA device: libsfmloss has no CPU fallback"); return SFM_E_NO_DEVICE; }
  int rc = nccl_load();
  if (rc) return rc;
  SfmComm* c = new (std::nothrow) SfmComm();
  if (!c) { sfm_set_error("out of host memory"); return SFM_E_INVALID_DESC; }
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  c->nranks = nranks;
  c->rank = rank;
  if ((rc = nccl_check(g_nccl.CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank"))) { delete c; return rc; }
  *comm_out = c;
  return 0;
}

extern "C" int sfm_comm_destroy(SfmComm* comm) {
  if (!comm) return 0;
  int rc = 0;
  if (g_nccl.CommDestroy) rc = nccl_check(g_nccl.CommDestroy(comm->comm), "ncclCommDestroy");
  delete comm;
  return rc;
}

extern "C" int sfm_allreduce_partials(SfmComm* comm, float* losses, int count, void* stream) {
  if (!comm || !losses) { sfm_set_error("sfm_allreduce_partials: null pointer"); return SFM_E_NULL_POINTER; }
  if (count < 1) { sfm_set_error("sfm_allreduce_partials: count %d < 1", count); return SFM_E_INVALID_SHAPE; }
  return nccl_check(g_nccl.AllReduce(losses, losses, (size_t)count, /*ncclFloat*/ 7, /*ncclSum*/ 0, comm->comm, (cudaStream_t)stream),
                    "ncclAllReduce");
}


// ------------------------------------------------------------------------------------------------
// SfmPeer: slot arrays for the in-kernel cross-GPU sum (one process per GPU, one node, NVLink / NVSwitch)
// ------------------------------------------------------------------------------------------------
struct SfmPeer {
  int nranks, rank;
  char* local;                       // [2][nranks] SfmPeerSlot + the step counter
  void* mapped[SFM_MAX_PEERS];       // peers' arrays as mapped here (mapped[rank] == local)
  bool connected;
  SfmPeerDev dev;
};

static size_t peer_bytes(int nranks) { return sfm_align_up((size_t)2 * nranks * sizeof(SfmPeerSlot), 256) + 256; }

extern "C" int sfm_peer_create(int nranks, int rank, SfmPeer** peer_out, void* ipc_handle_out) {
  if (!peer_out || !ipc_handle_out) { sfm_set_error("sfm_peer_create: null pointer"); return SFM_E_NULL_POINTER; }
  if (nranks < 1 || nranks > SFM_MAX_PEERS || rank < 0 || rank >= nranks) {
    sfm_set_error("sfm_peer_create: rank %d outside a world of %d (1..%d ranks)", rank, nranks, SFM_MAX_PEERS);
    return SFM_E_INVALID_DESC;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == SFM_IPC_HANDLE_BYTES, "SFM_IPC_HANDLE_BYTES");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { sfm_set_error("no CUDA device: libsfmloss has no CPU fallback"); return SFM_E_NO_DEVICE; }
  SfmPeer* q = new (std::nothrow) SfmPeer();
  if (!q) { sfm_set_error("out of host memory"); return SFM_E_INVALID_DESC; }
  q->nranks = nranks; q->rank = rank; q->connected = false;
  for (int r = 0; r < SFM_MAX_PEERS; ++r) q->mapped[r] = nullptr;
  cudaError_t e = cudaMalloc((void**)&q->local, peer_bytes(nranks));          // plain cudaMalloc: exportable with CUDA IPC
  if (e == cudaSuccess) e = cudaMemset(q->local, 0, peer_bytes(nranks));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, q->local);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    sfm_set_error("sfm_peer_create: %s", cudaGetErrorString(e));
    if (q->local) cudaFree(q->local);
    delete q;
    return (int)e;
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *peer_out = q;
  return 0;
}

extern "C" int sfm_peer_connect(SfmPeer* q, const void* all_handles) {
  if (!q || !all_handles) { sfm_set_error("sfm_peer_connect: null pointer"); return SFM_E_NULL_POINTER; }
  if (q->connected) { sfm_set_error("sfm_peer_connect: already connected"); return SFM_E_UNSUPPORTED; }
  const char* hs = (const char*)all_handles;
  for (int r = 0; r < q->nranks; ++r) {
    if (r == q->rank) { q->mapped[r] = q->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)r * sizeof(h), sizeof(h));
    const cudaError_t e = cudaIpcOpenMemHandle(&q->mapped[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      sfm_set_error("sfm_peer_connect: cudaIpcOpenMemHandle of rank %d failed: %s (the ranks must be processes on GPUs of one node "
                    "with peer access)", r, cudaGetErrorString(e));
      return (int)e;
    }
  }
  q->dev.nranks = q->nranks;
  q->dev.rank = q->rank;
  q->dev.counter = (unsigned*)(q->local + sfm_align_up((size_t)2 * q->nranks * sizeof(SfmPeerSlot), 256));
  for (int r = 0; r < SFM_MAX_PEERS; ++r) q->dev.slots[r] = (SfmPeerSlot*)q->mapped[r < q->nranks ? r : q->rank];
  q->connected = true;
  return 0;
}

extern "C" int sfm_peer_destroy(SfmPeer* q) {
  if (!q) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < q->nranks; ++r)
    if (r != q->rank && q->mapped[r]) cudaIpcCloseMemHandle(q->mapped[r]);
  if (q->local) cudaFree(q->local);
  delete q;
  return 0;
}

const SfmPeerDev* sfm_peer_dev(const SfmPeer* q) { return (q && q->connected) ? &q->dev : nullptr; }
