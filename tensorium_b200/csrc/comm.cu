// comm.cu — the one exchange step of the path: the all-reduce of the update buffers between TNNet.backward and
// TNNet.update when the batch is sharded over several GPUs.  The reference is single-device
// ("todo ... multi GPU", source/nConvolutionLayer.pas:472,498), so these entry points have no counterpart in
// source/nncuda.pas; they exist so that the FreePascal host — one process, or one TNNCuda instance + host thread, per
// GPU — can train data-parallel without any Python: rank 0 calls tnn_comm_unique_id, ships the 128 bytes to the
// other ranks by whatever means the host has (a file, a socket, MPI), every rank calls tnn_comm_init_rank, and
// TNNet.update calls tnn_allreduce_sum on the flat bucket that aliases every layer's weight_updates / bias_updates /
// scale_updates before the SGD kernels (source/nnet.pas:371-403, source/nConvolutionLayer.pas:1294-1330).
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"): a process that already carries an NCCL (torch.distributed in
// bench.py) shares that one instead of loading a second copy, and hosts that never train multi-GPU need no NCCL at
// all.  The collective is enqueued on the context's stream, so it is ordered after the backward kernels and before the
// update kernels like any other operator — and it can be captured into the step's CUDA graph
// (tnn_graph_begin_capture ... tnn_graph_end_capture).
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclSum = 0, ncclMax = 2 };

struct NcclApi {
  void *lib;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int *);
};

NcclApi *nccl() {
  static NcclApi api;
  static int state = 0;  // 0 untried, 1 ok, -1 failed
  if (state == 0) {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    state = -1;
    if (api.lib) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
      api.Broadcast = (decltype(api.Broadcast))dlsym(api.lib, "ncclBroadcast");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
      api.GetVersion = (decltype(api.GetVersion))dlsym(api.lib, "ncclGetVersion");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Broadcast && api.GetErrorString)
        state = 1;
    }
  }
  return state == 1 ? &api : nullptr;
}

}  // namespace

struct tnn_comm {
  ncclComm_t comm;
  int nranks, rank, device;
};

#define TNN_NCCL(api, call)                                                                         \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != 0) {                                                                                 \
      tnn_set_error("%s:%d %s failed: %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r__)); \
      return 7;                                                                                     \
    }                                                                                               \
  } while (0)

extern "C" {

int32_t tnn_comm_unique_id(void *id, int64_t bytes) {
  NcclApi *api = nccl();
  TNN_REQUIRE(api != nullptr, "NCCL (libnccl.so.2) could not be loaded");
  TNN_REQUIRE(id != nullptr && bytes >= (int64_t)sizeof(ncclUniqueId), "the unique id needs a 128-byte buffer");
  ncclUniqueId u;
  TNN_NCCL(api, api->GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return 0;
}

int32_t tnn_comm_init_rank(tnn_context *ctx, int32_t nranks, int32_t rank, const void *id, tnn_comm **comm) {
  TNN_CHECK_CTX(ctx);
  NcclApi *api = nccl();
  TNN_REQUIRE(api != nullptr, "NCCL (libnccl.so.2) could not be loaded");
  TNN_REQUIRE(comm != nullptr && id != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, "bad communicator arguments");
  *comm = nullptr;
  TNN_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  tnn_comm *c = new tnn_comm();
  c->nranks = nranks;
  c->rank = rank;
  c->device = ctx->device;
  ncclResult_t r = api->CommInitRank(&c->comm, nranks, u, rank);
  if (r != 0) {
    tnn_set_error("ncclCommInitRank(%d of %d) failed: %s", rank, nranks, api->GetErrorString(r));
    delete c;
    return 7;
  }
  *comm = c;
  return 0;
}

int32_t tnn_comm_destroy(tnn_comm *comm) {
  if (!comm) return 0;
  NcclApi *api = nccl();
  if (api) api->CommDestroy(comm->comm);
  delete comm;
  return 0;
}

int32_t tnn_comm_info(tnn_comm *comm, int32_t *nranks, int32_t *rank, int32_t *nccl_version) {
  TNN_REQUIRE(comm != nullptr, "null communicator");
  if (nranks) *nranks = comm->nranks;
  if (rank) *rank = comm->rank;
  if (nccl_version) {
    int v = 0;
    NcclApi *api = nccl();
    if (api && api->GetVersion) api->GetVersion(&v);
    *nccl_version = v;
  }
  return 0;
}

int32_t tnn_allreduce_sum(tnn_context *ctx, tnn_comm *comm, float *buf, int64_t offset, int64_t n) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(comm != nullptr && comm->device == ctx->device, "communicator belongs to another device");
  if (n <= 0) return 0;
  NcclApi *api = nccl();
  TNN_REQUIRE(api != nullptr, "NCCL (libnccl.so.2) could not be loaded");
  TNN_NCCL(api, api->AllReduce(buf + offset, buf + offset, (size_t)n, ncclFloat32, ncclSum, comm->comm, ctx->stream));
  ctx->launches++;
  return 0;
}

int32_t tnn_broadcast(tnn_context *ctx, tnn_comm *comm, float *buf, int64_t offset, int64_t n, int32_t root) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(comm != nullptr && comm->device == ctx->device, "communicator belongs to another device");
  TNN_REQUIRE(root >= 0 && root < comm->nranks, "bad root");
  if (n <= 0 || comm->nranks == 1) return 0;
  NcclApi *api = nccl();
  TNN_REQUIRE(api != nullptr, "NCCL (libnccl.so.2) could not be loaded");
  TNN_NCCL(api, api->Broadcast(buf + offset, buf + offset, (size_t)n, ncclFloat32, root, comm->comm, ctx->stream));
  ctx->launches++;
  return 0;
}

}  // extern "C"
