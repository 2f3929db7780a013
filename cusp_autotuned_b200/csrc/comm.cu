// comm.cu — NCCL communicator owned by the handle: halo exchange of x / p with
// the two neighbouring ranks (ncclSend/ncclRecv over NVLink 5 / NVSwitch) and
// the scalar all-reduces of CG.  NCCL is resolved at run time with dlopen so the
// library has no link-time dependency and shares the NCCL that the host
// process (e.g. torch.distributed) already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include "comm.h"

namespace b200sp {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi &nccl() {
  static NcclApi api;
  if (api.lib) return api;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    api.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process?
    if (api.lib) break;
  }
  if (!api.lib)
    for (const char *nm : names) {
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
  if (!api.lib) return api;
#define SYM(field, name) *(void **)(&api.field) = dlsym(api.lib, name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Send && api.Recv &&
           api.GroupStart && api.GroupEnd && api.GetErrorString;
  return api;
}

#define B200SP_NCCL(h, expr)                                                                   \
  do {                                                                                         \
    ncclResult_t _r = (expr);                                                                  \
    if (_r != ncclSuccess)                                                                     \
      return b200sp::set_error((h), B200SP_COMM_ERROR, "%s failed: %s", #expr,                 \
                               b200sp::nccl().GetErrorString(_r));                             \
  } while (0)

b200sp_status comm_allreduce_sum(b200sp_handle h, cudaStream_t st, void *dev, int count, bool is_double) {
  if (h->world <= 1) return B200SP_OK;
  B200SP_NCCL(h, nccl().AllReduce(dev, dev, (size_t)count, is_double ? ncclDouble : ncclFloat, ncclSum,
                                  (ncclComm_t)h->nccl_comm, st));
  h->launches++;
  return B200SP_OK;
}

template <typename T>
__global__ void sqrt_inplace_kernel(T *v) {
  *v = (T)sqrt((double)*v);
}

void comm_sqrt_inplace(b200sp_handle h, cudaStream_t st, void *dev, bool is_double) {
  if (is_double)
    sqrt_inplace_kernel<double><<<1, 1, 0, st>>>((double *)dev);
  else
    sqrt_inplace_kernel<float><<<1, 1, 0, st>>>((float *)dev);
  h->launches++;
}

b200sp_status comm_halo_exchange(b200sp_handle h, cudaStream_t st, void *window, i64 n, i64 halo_lo,
                                 i64 halo_hi, size_t elem) {
  if (h->world <= 1) return B200SP_OK;
  B200SP_REQUIRE(h, halo_lo <= n && halo_hi <= n, "halo larger than the local block");
  char *w = reinterpret_cast<char *>(window);
  char *local = w + (size_t)halo_lo * elem;
  ncclComm_t comm = (ncclComm_t)h->nccl_comm;
  B200SP_NCCL(h, nccl().GroupStart());
  if (h->rank > 0 && halo_lo > 0) {
    B200SP_NCCL(h, nccl().Send(local, (size_t)halo_lo * elem, ncclChar, h->rank - 1, comm, st));
    B200SP_NCCL(h, nccl().Recv(w, (size_t)halo_lo * elem, ncclChar, h->rank - 1, comm, st));
  }
  if (h->rank < h->world - 1 && halo_hi > 0) {
    B200SP_NCCL(h, nccl().Send(local + (size_t)(n - halo_hi) * elem, (size_t)halo_hi * elem, ncclChar,
                               h->rank + 1, comm, st));
    B200SP_NCCL(h, nccl().Recv(local + (size_t)n * elem, (size_t)halo_hi * elem, ncclChar, h->rank + 1, comm,
                               st));
  }
  B200SP_NCCL(h, nccl().GroupEnd());
  h->launches++;
  return B200SP_OK;
}

}  // namespace b200sp

extern "C" {

b200sp_status b200sp_comm_unique_id(void *id128) {
  if (!id128) return b200sp::set_error(nullptr, B200SP_INVALID_INPUT, "comm_unique_id: null buffer");
  auto &api = b200sp::nccl();
  if (!api.ok) return b200sp::set_error(nullptr, B200SP_COMM_ERROR, "NCCL library not found (libnccl.so.2)");
  static_assert(sizeof(ncclUniqueId) == B200SP_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  ncclResult_t r = api.GetUniqueId(&id);
  if (r != ncclSuccess)
    return b200sp::set_error(nullptr, B200SP_COMM_ERROR, "ncclGetUniqueId: %s", api.GetErrorString(r));
  memcpy(id128, &id, sizeof(id));
  return B200SP_OK;
}

b200sp_status b200sp_comm_init(b200sp_handle h, const void *id128, int world_size, int rank) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, id128 && world_size >= 1 && rank >= 0 && rank < world_size, "comm_init: bad arguments");
  auto &api = b200sp::nccl();
  if (!api.ok) return b200sp::set_error(h, B200SP_COMM_ERROR, "NCCL library not found (libnccl.so.2)");
  if (h->nccl_comm) {
    api.CommDestroy((ncclComm_t)h->nccl_comm);
    h->nccl_comm = nullptr;
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  B200SP_NCCL(h, api.CommInitRank(&comm, world_size, id, rank));
  h->nccl_comm = comm;
  h->world = world_size;
  h->rank = rank;
  return B200SP_OK;
}

b200sp_status b200sp_comm_destroy(b200sp_handle h) {
  B200SP_CHECK_HANDLE(h);
  if (h->nccl_comm) {
    b200sp::nccl().CommDestroy((ncclComm_t)h->nccl_comm);
    h->nccl_comm = nullptr;
  }
  h->world = 1;
  h->rank = 0;
  return B200SP_OK;
}
}
