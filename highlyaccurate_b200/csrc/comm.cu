// Multi-GPU entry points of libha_b200.so: the one collective of the path, an all-gather of the
// final poses.  The reference has no distributed code (train_kitti.py:526-529
// uses cuda:0 only); samples are independent, so ranks own contiguous batch shards and exchange
// nothing until the end.
//
// NCCL is bound at run time (dlopen), from the copy already loaded in the process when there is
// one (torch ships its own libnccl.so.2 and a communicator must come from the library that uses
// it), so libha_b200.so has no link-time dependency on NCCL and single-GPU consumers never touch it.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace ha {
namespace {

typedef struct { char internal[HA_COMM_ID_BYTES]; } NcclId;     // == ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
typedef void* NcclComm;
constexpr int kNcclSuccess = 0, kNcclFloat32 = 7;                 // ncclResult_t / ncclDataType_t values of nccl.h

struct NcclApi {
  int (*GetUniqueId)(NcclId*);
  int (*CommInitRank)(NcclComm*, int, NcclId, int);
  int (*CommDestroy)(NcclComm);
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
  const char* (*GetErrorString)(int);
  bool ok;
};

const NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    memset(&a, 0, sizeof(a));
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);    // the copy the process already uses (torch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclId, int)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(h, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(dlsym(h, "ncclAllGather"));
    a.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather;
    return a;
  }();
  return api;
}

int comm_fail(const char* what, int rc) {
  char msg[256];
  const NcclApi& n = nccl();
  snprintf(msg, sizeof(msg), "%s: %s", what, !n.ok ? "libnccl.so.2 could not be loaded" : (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
  set_error_text(msg);
  return HA_ECOMM;
}

}  // namespace
}  // namespace ha

extern "C" int ha_comm_unique_id(void* id128_host) {
  if (!id128_host) return HA_EINVAL;
  const ha::NcclApi& n = ha::nccl();
  if (!n.ok) return ha::comm_fail("ha_comm_unique_id", 0);
  ha::NcclId id;
  const int rc = n.GetUniqueId(&id);
  if (rc != ha::kNcclSuccess) return ha::comm_fail("ncclGetUniqueId", rc);
  memcpy(id128_host, &id, HA_COMM_ID_BYTES);
  return HA_OK;
}

extern "C" int ha_comm_init(void** comm, int world, int rank, const void* id128_host, int device) {
  if (!comm || !id128_host || world < 1 || rank < 0 || rank >= world) return HA_EINVAL;
  const ha::NcclApi& n = ha::nccl();
  if (!n.ok) return ha::comm_fail("ha_comm_init", 0);
  HA_CUDA_TRY(cudaSetDevice(device));
  ha::NcclId id;
  memcpy(&id, id128_host, HA_COMM_ID_BYTES);
  ha::NcclComm c = nullptr;
  const int rc = n.CommInitRank(&c, world, id, rank);
  if (rc != ha::kNcclSuccess) return ha::comm_fail("ncclCommInitRank", rc);
  *comm = c;
  return HA_OK;
}

extern "C" int ha_comm_destroy(void* comm) {
  if (!comm) return HA_EINVAL;
  const ha::NcclApi& n = ha::nccl();
  if (!n.ok) return ha::comm_fail("ha_comm_destroy", 0);
  const int rc = n.CommDestroy(comm);
  return rc == ha::kNcclSuccess ? HA_OK : ha::comm_fail("ncclCommDestroy", rc);
}

extern "C" int ha_pose_allgather(void* comm, const float* local, float* all, int n_local, void* stream) {
  if (!comm || !local || !all || n_local <= 0) return HA_EINVAL;
  const ha::NcclApi& n = ha::nccl();
  if (!n.ok) return ha::comm_fail("ha_pose_allgather", 0);
  const int rc = n.AllGather(local, all, (size_t)n_local * 3, ha::kNcclFloat32, comm, reinterpret_cast<cudaStream_t>(stream));
  return rc == ha::kNcclSuccess ? HA_OK : ha::comm_fail("ncclAllGather", rc);
}
