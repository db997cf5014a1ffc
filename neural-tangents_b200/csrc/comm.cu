// ntk_comm_*: the multi-GPU plumbing of libntk_b200.so -- NCCL over NVLink / NVSwitch, one rank per GPU.
//
// The Gram matrix shards by rows with no exchange inside the computation: inputs are broadcast
// once, result slabs are all-gathered, nothing is reduced.  The reference does the same with `pmap` over x1
// rows and x2 replicated (`_src/batching.py:505-644`).  libnccl is resolved with dlopen on first use: the
// library loads (and every single-GPU entry point works) on hosts without NCCL, and a process that already
// carries a libnccl.so.2 (e.g. the one bundled with another framework) shares it instead of loading a second.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"

using namespace ntk;

struct ntk_comm {
  ncclComm_t comm = nullptr;
  ntk_context_t* ctx = nullptr;
  int rank = 0, world = 1;
};

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("NTK_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so",
                           "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (a.handle) break;
      a.error = dlerror();
    }
    if (!a.handle) return;
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(a.handle, name);
      if (!p) {
        ok = false;
        a.error = std::string("missing symbol ") + name;
      }
      return p;
    };
    a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.Broadcast = (decltype(a.Broadcast))sym("ncclBroadcast");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    if (!ok) {
      dlclose(a.handle);
      a.handle = nullptr;
    }
  });
  return a;
}

int need_nccl() {
  if (api().handle) return NTK_OK;
  return fail(NTK_EUNSUPPORTED, "NCCL is not available (%s); multi-GPU entry points need libnccl.so.2",
              api().error.c_str());
}

#define NTK_NCCL(expr)                                                                            \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess)                                                                        \
      return fail(NTK_ECUDA, "%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #expr, api().GetErrorString(_r)); \
  } while (0)

}  // namespace

extern "C" {

int ntk_comm_nccl_version(int* version) {
  if (!version) return fail(NTK_EINVAL, "version is NULL");
  NTK_TRY(need_nccl());
  NTK_NCCL(api().GetVersion(version));
  return NTK_OK;
}

int ntk_comm_unique_id(void* id_out) {
  static_assert(NTK_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "NTK_COMM_ID_BYTES must match ncclUniqueId");
  if (!id_out) return fail(NTK_EINVAL, "id_out is NULL");
  NTK_TRY(need_nccl());
  ncclUniqueId id;
  NTK_NCCL(api().GetUniqueId(&id));
  memcpy(id_out, id.internal, NCCL_UNIQUE_ID_BYTES);
  return NTK_OK;
}

int ntk_comm_create(ntk_context_t* ctx, const void* id, int32_t rank, int32_t world, ntk_comm_t** out) {
  if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return fail(NTK_EINVAL, "bad arguments");
  NTK_TRY(need_nccl());
  NTK_CUDA(cudaSetDevice(ntk_context_device(ctx)));
  ncclUniqueId uid;
  memcpy(uid.internal, id, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t c = nullptr;
  NTK_NCCL(api().CommInitRank(&c, world, uid, rank));
  ntk_comm* cm = new ntk_comm();
  cm->comm = c;
  cm->ctx = ctx;
  cm->rank = rank;
  cm->world = world;
  *out = cm;
  return NTK_OK;
}

void ntk_comm_destroy(ntk_comm_t* comm) {
  if (!comm) return;
  if (comm->comm && api().handle) {
    cudaSetDevice(ntk_context_device(comm->ctx));
    ntk_context_synchronize(comm->ctx);
    api().CommDestroy(comm->comm);
  }
  delete comm;
}

int ntk_comm_rank(const ntk_comm_t* comm) { return comm ? comm->rank : -1; }
int ntk_comm_world(const ntk_comm_t* comm) { return comm ? comm->world : -1; }

int ntk_comm_broadcast(ntk_comm_t* comm, void* dev_buf, size_t bytes, int32_t root) {
  if (!comm || !dev_buf || root < 0 || root >= comm->world) return fail(NTK_EINVAL, "bad arguments");
  if (bytes == 0) return NTK_OK;
  NTK_CUDA(cudaSetDevice(ntk_context_device(comm->ctx)));
  NTK_NCCL(api().Broadcast(dev_buf, dev_buf, bytes, ncclChar, root, comm->comm,
                           (cudaStream_t)ntk_context_stream(comm->ctx)));
  return NTK_OK;
}

int ntk_comm_all_gather(ntk_comm_t* comm, const void* send_dev, void* recv_dev, size_t bytes_per_rank) {
  if (!comm || !send_dev || !recv_dev) return fail(NTK_EINVAL, "bad arguments");
  if (bytes_per_rank == 0) return NTK_OK;
  NTK_CUDA(cudaSetDevice(ntk_context_device(comm->ctx)));
  NTK_NCCL(api().AllGather(send_dev, recv_dev, bytes_per_rank, ncclChar, comm->comm,
                           (cudaStream_t)ntk_context_stream(comm->ctx)));
  return NTK_OK;
}

}  // extern "C"
