// cml_comm.cu -- the collective of the hot path inside the C ABI (cml_set_comm, cml_allreduce_counts):
// one ncclAllReduce(sum, fp64) of the reduce buffer [count slots | sum ln P | sum w ln P | n_zero] per EM iteration,
// enqueued on the context's stream between the E-step kernels and the M-step kernels -- no host synchronisation and no
// host callback in between.  The reference has no counterpart (it is single process, single thread); the data-parallel
// identity is "examples are independent given the weights" (cached_derivs.h:69-75, forest-em.hpp:573-578).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the loader returns the copy a host process such as PyTorch has
// already mapped): the library has no link-time NCCL dependency and single-GPU users never touch it.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include <nccl.h>

#include "cml_ctx.cuh"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.handle) {
      api.err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    auto sym = [&](const char* s) {
      void* p = dlsym(api.handle, s);
      if (!p && api.err.empty()) api.err = std::string("libnccl: missing symbol ") + s;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return api;
}

}  // namespace

// 128-byte rendezvous token of a new communicator (rank 0 creates it, the launcher ships it to the other ranks)
extern "C" int cml_comm_unique_id(unsigned char out[CML_COMM_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) == CML_COMM_ID_BYTES, "ncclUniqueId size");
  NcclApi& n = nccl();
  if (!n.err.empty() || !out) return CML_ERR_CUDA;
  ncclUniqueId id;
  if (n.GetUniqueId(&id) != ncclSuccess) return CML_ERR_CUDA;
  std::memcpy(out, &id, sizeof(id));
  return CML_OK;
}

static int comm_fail(cml_ctx* ctx, ncclResult_t r, const char* what) {
  ctx->err = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
  return CML_ERR_CUDA;
}

// one process per GPU (torchrun, MPI, ...): every rank calls this with the same token
extern "C" int cml_comm_init_rank(cml_ctx* ctx, int n_ranks, int rank, const unsigned char id[CML_COMM_ID_BYTES]) {
  if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return CML_ERR_ARG;
  NcclApi& n = nccl();
  CML_REQUIRE(n.err.empty(), CML_ERR_CUDA, n.err);
  cudaSetDevice(ctx->device);
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t c;
  const ncclResult_t r = n.CommInitRank(&c, n_ranks, uid, rank);
  if (r != ncclSuccess) return comm_fail(ctx, r, "ncclCommInitRank");
  if (ctx->comm && ctx->own_comm) n.CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = c;
  ctx->own_comm = true;
  ctx->comm_rank = rank;
  ctx->comm_size = n_ranks;
  ctx->graph_dirty = true;
  return CML_OK;
}

// adopt an existing ncclComm_t (e.g. one made with ncclCommInitAll by a single-process multi-GPU host)
extern "C" int cml_set_comm(cml_ctx* ctx, void* nccl_comm, int rank, int n_ranks) {
  if (!ctx) return CML_ERR_ARG;
  if (ctx->comm && ctx->own_comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nccl_comm;
  ctx->own_comm = false;
  ctx->comm_rank = rank;
  ctx->comm_size = nccl_comm ? n_ranks : 1;
  ctx->graph_dirty = true;
  return CML_OK;
}

// single process, one context per device: communicators for all of them at once (ncclCommInitAll)
extern "C" int cml_comm_init_all(cml_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return CML_ERR_ARG;
  NcclApi& api = nccl();
  if (!api.err.empty()) {
    ctxs[0]->err = api.err;
    return CML_ERR_CUDA;
  }
  std::vector<int> dev(n);
  std::vector<ncclComm_t> comms(n);
  for (int i = 0; i < n; ++i) dev[i] = ctxs[i]->device;
  const ncclResult_t r = api.CommInitAll(comms.data(), n, dev.data());
  if (r != ncclSuccess) return comm_fail(ctxs[0], r, "ncclCommInitAll");
  for (int i = 0; i < n; ++i) {
    cml_ctx* ctx = ctxs[i];
    if (ctx->comm && ctx->own_comm) api.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = comms[i];
    ctx->own_comm = true;
    ctx->comm_rank = i;
    ctx->comm_size = n;
    ctx->graph_dirty = true;
  }
  return CML_OK;
}

extern "C" int cml_comm_info(cml_ctx* ctx, int* rank, int* n_ranks) {
  if (!ctx) return CML_ERR_ARG;
  if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
  if (n_ranks) *n_ranks = ctx->comm ? ctx->comm_size : 1;
  return CML_OK;
}

void cml_comm_release(cml_ctx* ctx) {  // from cml_destroy
  if (ctx->comm && ctx->own_comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
}

// sum the reduce buffer over the ranks, in place, stream-ordered after the E-step kernels of this context.
// A context without a communicator (single GPU) returns at once.
extern "C" int cml_allreduce_counts(cml_ctx* ctx) {
  if (!ctx) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model, CML_ERR_STATE, "cml_set_model first");
  if (!ctx->comm || ctx->comm_size <= 1) return CML_OK;
  cudaSetDevice(ctx->device);
  const ncclResult_t r = nccl().AllReduce(ctx->reduce, ctx->reduce, ctx->reduce_n, ncclDouble, ncclSum,
                                          (ncclComm_t)ctx->comm, ctx->stream);
  if (r != ncclSuccess) return comm_fail(ctx, r, "ncclAllReduce");
  ++ctx->collectives;
  return CML_OK;
}

// sum n doubles at a device address over the ranks (corpus statistics before training; forest contexts reuse it)
extern "C" int cml_allreduce_buffer(cml_ctx* ctx, void* device_ptr, uint64_t n) {
  if (!ctx || !device_ptr) return CML_ERR_ARG;
  if (!ctx->comm || ctx->comm_size <= 1) return CML_OK;
  cudaSetDevice(ctx->device);
  const ncclResult_t r =
      nccl().AllReduce(device_ptr, device_ptr, n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream);
  if (r != ncclSuccess) return comm_fail(ctx, r, "ncclAllReduce");
  ++ctx->collectives;
  return CML_OK;
}

extern "C" uint64_t cml_collective_count(cml_ctx* ctx) { return ctx ? ctx->collectives : 0; }

// ---- the same collective for forest contexts (cml_forest.cu): generic helpers on a bare communicator ---------------
int cml_nccl_init_rank(void** comm, int n_ranks, int rank, const unsigned char* id, std::string& err) {
  NcclApi& n = nccl();
  if (!n.err.empty()) {
    err = n.err;
    return CML_ERR_CUDA;
  }
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t c;
  const ncclResult_t r = n.CommInitRank(&c, n_ranks, uid, rank);
  if (r != ncclSuccess) {
    err = std::string("ncclCommInitRank: ") + n.GetErrorString(r);
    return CML_ERR_CUDA;
  }
  *comm = c;
  return CML_OK;
}
int cml_nccl_allreduce(void* comm, double* p, uint64_t count, cudaStream_t s, std::string& err) {
  const ncclResult_t r = nccl().AllReduce(p, p, count, ncclDouble, ncclSum, (ncclComm_t)comm, s);
  if (r != ncclSuccess) {
    err = std::string("ncclAllReduce: ") + nccl().GetErrorString(r);
    return CML_ERR_CUDA;
  }
  return CML_OK;
}
void cml_nccl_destroy(void* comm) {
  if (comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)comm);
}
