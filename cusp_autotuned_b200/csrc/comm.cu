// comm.cu — NCCL communicator owned by the handle: halo exchange of x / p with
// the two neighbouring ranks (ncclSend/ncclRecv over NVLink 5 / NVSwitch) and
// the scalar all-reduces of CG.  NCCL is resolved at run time with dlopen so the
// library has no link-time dependency and shares the NCCL that the host
// process (e.g. torch.distributed) already loaded.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include "comm.h"

namespace b200sp {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
  SYM(AllGather, "ncclAllGather");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.Send && api.Recv &&
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


// ---------------------------------------------------------------------------
// NVLink peer-memory path: CUDA-IPC mapped mailboxes + neighbour workspaces
// ---------------------------------------------------------------------------
b200sp_status comm_allgather_host(b200sp_handle h, cudaStream_t st, const void *host_in, size_t bytes,
                                  void *host_out) {
  B200SP_REQUIRE(h, h->nccl_comm && bytes > 0 && bytes <= 256, "allgather: bad arguments");
  b200sp_status s = ensure_scratch(h, (size_t)(h->world + 1) * 256);
  if (s != B200SP_OK) return s;
  char *d = reinterpret_cast<char *>(h->scratch);
  B200SP_CUDA(h, cudaMemcpyAsync(d, host_in, bytes, cudaMemcpyHostToDevice, st));
  B200SP_NCCL(h, nccl().AllGather(d, d + 256, bytes, ncclChar, (ncclComm_t)h->nccl_comm, st));
  B200SP_CUDA(h, cudaMemcpyAsync(host_out, d + 256, bytes * (size_t)h->world, cudaMemcpyDeviceToHost, st));
  B200SP_CUDA(h, cudaStreamSynchronize(st));
  return B200SP_OK;
}

static bool all_ranks_ok(b200sp_handle h, cudaStream_t st, int mine) {
  int flags[P2P_MAX_WORLD];
  if (comm_allgather_host(h, st, &mine, sizeof(int), flags) != B200SP_OK) return false;
  for (int r = 0; r < h->world; ++r)
    if (!flags[r]) return false;
  return true;
}

static void p2p_teardown(b200sp_handle h) {
  for (int r = 0; r < P2P_MAX_WORLD; ++r) {
    if (h->peer_mail[r] && h->peer_mail[r] != h->mail) cudaIpcCloseMemHandle(h->peer_mail[r]);
    h->peer_mail[r] = nullptr;
  }
  for (int k = 0; k < 2; ++k) {
    if (h->nbr_ws[k]) cudaIpcCloseMemHandle(h->nbr_ws[k]);
    h->nbr_ws[k] = nullptr;
    memset(h->nbr_ws_handle[k], 0, 64);
  }
  for (int k = 0; k < 2; ++k) {
    if (h->nbr_stage[k]) cudaIpcCloseMemHandle(h->nbr_stage[k]);
    h->nbr_stage[k] = nullptr;
  }
  for (int r = 0; r < P2P_MAX_WORLD; ++r) {
    if (h->peer_gather[r] && h->peer_gather[r] != h->gather_stage) cudaIpcCloseMemHandle(h->peer_gather[r]);
    h->peer_gather[r] = nullptr;
  }
  if (h->gather_stage) cudaFree(h->gather_stage);
  h->gather_stage = nullptr;
  h->gather_slice_cap = 0;
  h->gather_epoch = 0;
  if (h->halo_stage) cudaFree(h->halo_stage);
  h->halo_stage = nullptr;
  if (h->mail) cudaFree(h->mail);
  h->mail = nullptr;
  h->p2p_ok = false;
  h->xchg_epoch = 0;
  cudaGetLastError();
}

// collective, called from b200sp_comm_init: every rank exports its mailbox and maps
// everybody else's.  Any failure on any rank leaves the whole job on the NCCL path.
static void p2p_setup(b200sp_handle h) {
  h->p2p_ok = false;
  const char *off = getenv("B200SP_DISABLE_P2P");
  int want = !(off && off[0] && off[0] != '0') && h->world > 1 && h->world <= P2P_MAX_WORLD;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  struct Info {
    unsigned char handle[64];
    unsigned char stage_handle[64];
    int ok;
  } mine;
  memset(&mine, 0, sizeof(mine));
  if (want) {
    if (cudaMalloc(&h->mail, 4096) == cudaSuccess && cudaMemset(h->mail, 0, 4096) == cudaSuccess &&
        cudaMalloc(&h->halo_stage, 4 * P2P_STAGE_SIDE) == cudaSuccess) {
      cudaIpcMemHandle_t hd, hs;
      if (cudaIpcGetMemHandle(&hd, h->mail) == cudaSuccess && cudaIpcGetMemHandle(&hs, h->halo_stage) == cudaSuccess) {
        memcpy(mine.handle, &hd, 64);
        memcpy(mine.stage_handle, &hs, 64);
        mine.ok = 1;
      }
    }
    cudaGetLastError();
  }
  Info all[P2P_MAX_WORLD];
  if (h->world > P2P_MAX_WORLD || comm_allgather_host(h, nullptr, &mine, sizeof(Info), all) != B200SP_OK) {
    p2p_teardown(h);
    return;
  }
  int ok = 1;
  for (int r = 0; r < h->world; ++r) ok = ok && all[r].ok;
  if (ok) {
    for (int r = 0; r < h->world && ok; ++r) {
      if (r == h->rank) {
        h->peer_mail[r] = h->mail;
        continue;
      }
      cudaIpcMemHandle_t hd;
      memcpy(&hd, all[r].handle, 64);
      if (cudaIpcOpenMemHandle(&h->peer_mail[r], hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        h->peer_mail[r] = nullptr;
        ok = 0;
        cudaGetLastError();
      }
      const int side = (r == h->rank - 1) ? 0 : ((r == h->rank + 1) ? 1 : -1);
      if (ok && side >= 0) {
        memcpy(&hd, all[r].stage_handle, 64);
        if (cudaIpcOpenMemHandle(&h->nbr_stage[side], hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          h->nbr_stage[side] = nullptr;
          ok = 0;
          cudaGetLastError();
        }
      }
    }
  }
  if (!all_ranks_ok(h, nullptr, ok)) {
    p2p_teardown(h);
    return;
  }
  h->p2p_ok = true;
}


// ---- one-kernel halo exchange through peer memory ---------------------------------
// grid of XCHG_CTAS co-resident CTAs (the stream is otherwise idle when it runs):
//   A  copy my first halo_lo / last halo_hi local elements into the staging buffers of
//      rank-1 / rank+1 (16-byte stores over NVLink);
//   B  last CTA to finish A publishes xchg_flag = epoch in both neighbours' mailboxes,
//      waits until both neighbours have published theirs here, then raises xchg_go;
//   C  every CTA waits for xchg_go and copies the staged planes into my window.
// Staging is double-buffered by epoch parity: a neighbour can be at most one exchange
// ahead (its next-but-one push needs my next push, which follows my copy-out in stream
// order).
constexpr int XCHG_CTAS = 64;
constexpr int XCHG_BLOCK = 256;

__device__ __forceinline__ void copy_bytes16(char *dst, const char *src, size_t bytes, size_t tid, size_t nthreads) {
  // both pointers 16-byte aligned when bytes % 16 == 0 in our layouts; otherwise byte-wise tail
  const size_t n16 = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) ? 0 : bytes / 16;
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst);
  for (size_t i = tid; i < n16; i += nthreads) d4[i] = s4[i];
  for (size_t i = n16 * 16 + tid; i < bytes; i += nthreads) dst[i] = src[i];
}

__global__ void __launch_bounds__(XCHG_BLOCK) halo_xchg_kernel(char *window, size_t local_off, size_t n_bytes,
                                                               size_t lo_bytes, size_t hi_bytes, char *stage_mine,
                                                               char *stage_lo_nbr, char *stage_hi_nbr, Mailbox *mine,
                                                               Mailbox *mail_lo_nbr, Mailbox *mail_hi_nbr,
                                                               unsigned int *ticket, unsigned long long epoch) {
  const size_t tid = (size_t)blockIdx.x * XCHG_BLOCK + threadIdx.x, nth = (size_t)gridDim.x * XCHG_BLOCK;
  const size_t par = (size_t)(epoch & 1) * 2 * P2P_STAGE_SIDE;
  char *local = window + local_off;
  // A: my low edge is rank-1's upper halo (its slot "from rank+1"), my high edge rank+1's lower halo
  if (stage_lo_nbr) copy_bytes16(stage_lo_nbr + par + P2P_STAGE_SIDE, local, lo_bytes, tid, nth);
  if (stage_hi_nbr) copy_bytes16(stage_hi_nbr + par, local + n_bytes - hi_bytes, hi_bytes, tid, nth);
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence_system();
    unsigned long long *f;
    if (mail_lo_nbr) {
      f = &mail_lo_nbr->xchg_flag[1];
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    }
    if (mail_hi_nbr) {
      f = &mail_hi_nbr->xchg_flag[0];
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    }
    for (int side = 0; side < 2; ++side) {
      if (!(side == 0 ? mail_lo_nbr : mail_hi_nbr)) continue;
      unsigned long long v;
      SpinGuard guard;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&mine->xchg_flag[side]) : "memory");
      } while (v < epoch && !guard.expired(mine));  // monotonic epochs: a neighbour may be one exchange ahead
    }
    *ticket = 0;
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&mine->xchg_go), "l"(epoch) : "memory");
  }
  // C
  if (threadIdx.x == 0) {
    unsigned long long v;
    SpinGuard guard;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&mine->xchg_go) : "memory");
    } while (v < epoch && !guard.expired(mine));
  }
  __syncthreads();
  if (mail_lo_nbr) copy_bytes16(window, stage_mine + par, lo_bytes, tid, nth);
  if (mail_hi_nbr) copy_bytes16(local + n_bytes, stage_mine + par + P2P_STAGE_SIDE, hi_bytes, tid, nth);
}

b200sp_status comm_halo_exchange_auto(b200sp_handle h, cudaStream_t st, void *window, i64 n, i64 halo_lo,
                                      i64 halo_hi, size_t elem) {
  if (h->world <= 1) return B200SP_OK;
  const bool has_lo = h->rank > 0 && halo_lo > 0, has_hi = h->rank < h->world - 1 && halo_hi > 0;
  const size_t lo_bytes = (size_t)halo_lo * elem, hi_bytes = (size_t)halo_hi * elem;
  // symmetric sizes make this decision identical on both ends of every link
  if (!h->p2p_ok || lo_bytes > P2P_STAGE_SIDE || hi_bytes > P2P_STAGE_SIDE)
    return comm_halo_exchange(h, st, window, n, halo_lo, halo_hi, elem);
  B200SP_REQUIRE(h, halo_lo <= n && halo_hi <= n, "halo larger than the local block");
  const unsigned long long epoch = ++h->xchg_epoch;
  if (!has_lo && !has_hi) return B200SP_OK;
  Mailbox *mine = reinterpret_cast<Mailbox *>(h->mail);
  halo_xchg_kernel<<<XCHG_CTAS, XCHG_BLOCK, 0, st>>>(
      reinterpret_cast<char *>(window), (size_t)halo_lo * elem, (size_t)n * elem, has_lo ? lo_bytes : 0,
      has_hi ? hi_bytes : 0, reinterpret_cast<char *>(h->halo_stage),
      has_lo ? reinterpret_cast<char *>(h->nbr_stage[0]) : nullptr,
      has_hi ? reinterpret_cast<char *>(h->nbr_stage[1]) : nullptr, mine,
      has_lo ? reinterpret_cast<Mailbox *>(h->peer_mail[h->rank - 1]) : nullptr,
      has_hi ? reinterpret_cast<Mailbox *>(h->peer_mail[h->rank + 1]) : nullptr, h->red_counters + 2, epoch);
  B200SP_LAUNCH_CHECK(h, "halo_xchg_kernel");
  return B200SP_OK;
}

bool comm_fused_xchg_prepare(b200sp_handle h, void *window, i64 n, i64 halo_lo, i64 halo_hi, size_t elem,
                             FusedXchg *xc) {
  memset(xc, 0, sizeof(*xc));
  if (h->world <= 1 || !h->p2p_ok) return false;
  const bool has_lo = h->rank > 0 && halo_lo > 0, has_hi = h->rank < h->world - 1 && halo_hi > 0;
  const size_t lo_bytes = (size_t)halo_lo * elem, hi_bytes = (size_t)halo_hi * elem;
  if (!has_lo && !has_hi) return false;
  if (lo_bytes > P2P_STAGE_SIDE || hi_bytes > P2P_STAGE_SIDE || halo_lo > n || halo_hi > n) return false;
  // halo regions must not share a 128-byte line with the local part: interior tiles read
  // x through the non-coherent path while the copy-out is still in flight
  const uintptr_t w = reinterpret_cast<uintptr_t>(window);
  if ((w & 127) || (lo_bytes & 127) || (((size_t)(halo_lo + n) * elem) & 127)) return false;
  xc->enabled = 1;
  xc->window = reinterpret_cast<char *>(window);
  xc->local_off = lo_bytes;
  xc->n_bytes = (size_t)n * elem;
  xc->lo_bytes = has_lo ? lo_bytes : 0;
  xc->hi_bytes = has_hi ? hi_bytes : 0;
  xc->stage_mine = reinterpret_cast<char *>(h->halo_stage);
  xc->stage_lo_nbr = has_lo ? reinterpret_cast<char *>(h->nbr_stage[0]) : nullptr;
  xc->stage_hi_nbr = has_hi ? reinterpret_cast<char *>(h->nbr_stage[1]) : nullptr;
  xc->mine = reinterpret_cast<Mailbox *>(h->mail);
  xc->mail_lo_nbr = has_lo ? reinterpret_cast<Mailbox *>(h->peer_mail[h->rank - 1]) : nullptr;
  xc->mail_hi_nbr = has_hi ? reinterpret_cast<Mailbox *>(h->peer_mail[h->rank + 1]) : nullptr;
  xc->tickets = h->red_counters + 4;
  xc->epoch = ++h->xchg_epoch;
  xc->wait_lo = xc->wait_hi = &xc->mine->xchg_go;  // raised locally when both planes are in the window
  xc->wait_epoch = xc->epoch;
  return true;
}

// ---- all-gather of x slices through peer memory ------------------------------------
// copy with congruent misalignment: dst and src have the same address modulo 16
__device__ __forceinline__ void copy_congruent(char *dst, const char *src, size_t bytes, size_t tid, size_t nth) {
  size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (head > bytes) head = bytes;
  for (size_t i = tid; i < head; i += nth) dst[i] = src[i];
  const size_t n16 = (bytes - head) / 16;
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
  // 4 independent 16-byte loads in flight per thread: a peer read over NVLink takes microseconds
  size_t i = tid;
  for (; i + 3 * nth < n16; i += 4 * nth) {
    const uint4 a = s4[i], b = s4[i + nth], c = s4[i + 2 * nth], d = s4[i + 3 * nth];
    d4[i] = a;
    d4[i + nth] = b;
    d4[i + 2 * nth] = c;
    d4[i + 3 * nth] = d;
  }
  for (; i < n16; i += nth) d4[i] = s4[i];
  for (size_t j = head + n16 * 16 + tid; j < bytes; j += nth) dst[j] = src[j];
}

struct GatherArgs {
  char *x_full;
  char *stage_mine;
  char *stage_peer[P2P_MAX_WORLD];
  Mailbox *mail[P2P_MAX_WORLD];
  unsigned long long off[P2P_MAX_WORLD + 1];  // byte offsets of the slices in x_full
  size_t par_off;                             // byte offset of this epoch's half of a staging buffer
  unsigned int *ticket;
  unsigned long long epoch;
  int world, rank;
};

constexpr int GATHER_BLOCK = 256;

// Grid: at most 2 CTAs per SM (co-resident: CTAs spin on flags raised by other GPUs).
//   A  my slice -> my staging buffer (same misalignment as in x_full); the CTA that finishes last
//      publishes gather_flag[rank] = epoch in every peer's mailbox;
//   B  for each peer in rotated order: wait for its flag here, pull its slice into x_full.
__global__ void __launch_bounds__(GATHER_BLOCK) allgather_pull_kernel(GatherArgs g) {
  const size_t tid = (size_t)blockIdx.x * GATHER_BLOCK + threadIdx.x, nth = (size_t)gridDim.x * GATHER_BLOCK;
  {
    const size_t o = g.off[g.rank], b = g.off[g.rank + 1] - o;
    copy_congruent(g.stage_mine + g.par_off + (o & 15), g.x_full + o, b, tid, nth);
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(g.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last && threadIdx.x < g.world && (int)threadIdx.x != g.rank) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&g.mail[threadIdx.x]->gather_flag[g.rank]), "l"(g.epoch)
                 : "memory");
  }
  if (is_last && threadIdx.x == 0) *g.ticket = 0;
  Mailbox *mine = g.mail[g.rank];
  for (int k = 1; k < g.world; ++k) {
    const int r = (g.rank + k) % g.world;
    if (threadIdx.x == 0) {
      unsigned long long v;
      SpinGuard guard;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&mine->gather_flag[r]) : "memory");
      } while (v < g.epoch && !guard.expired(mine));
    }
    __syncthreads();
    const size_t o = g.off[r], b = g.off[r + 1] - o;
    copy_congruent(g.x_full + o, g.stage_peer[r] + g.par_off + (o & 15), b, tid, nth);
  }
}

// Push variant (default): NVLink stores are posted, loads pay a round trip per request, so the
// slices travel as stores.  Every staging buffer mirrors the whole of x (same byte offsets).
//   A  my slice -> every peer's staging buffer, rotated peer order; the CTA that finishes last
//      publishes gather_flag[rank] = epoch in every peer's mailbox;
//   B  for each peer: wait for its flag here, copy its slice from MY staging buffer into x_full
//      (local copy).
// Reuse of a staging half two epochs later is safe: to get there a rank has passed B of the epoch in
// between, i.e. has seen every peer's flag of that epoch, which a peer raises only after its
// previous kernel (and with it its copy-out of the older data) has completed.
__global__ void __launch_bounds__(GATHER_BLOCK) allgather_push_kernel(GatherArgs g) {
  const size_t tid = (size_t)blockIdx.x * GATHER_BLOCK + threadIdx.x, nth = (size_t)gridDim.x * GATHER_BLOCK;
  {
    const size_t o = g.off[g.rank], b = g.off[g.rank + 1] - o;
    for (int k = 1; k < g.world; ++k) {
      const int r = (g.rank + k) % g.world;
      copy_congruent(g.stage_peer[r] + g.par_off + o, g.x_full + o, b, tid, nth);
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(g.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last && threadIdx.x < g.world && (int)threadIdx.x != g.rank) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&g.mail[threadIdx.x]->gather_flag[g.rank]), "l"(g.epoch)
                 : "memory");
  }
  if (is_last && threadIdx.x == 0) *g.ticket = 0;
  Mailbox *mine = g.mail[g.rank];
  for (int k = 1; k < g.world; ++k) {
    const int r = (g.rank + g.world - k) % g.world;  // the peer that pushed to me first comes first
    if (threadIdx.x == 0) {
      unsigned long long v;
      SpinGuard guard;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&mine->gather_flag[r]) : "memory");
      } while (v < g.epoch && !guard.expired(mine));
    }
    __syncthreads();
    const size_t o = g.off[r], b = g.off[r + 1] - o;
    copy_congruent(g.x_full + o, g.stage_mine + g.par_off + o, b, tid, nth);
  }
}

// (re)allocate and IPC-share the staging buffers so that a slice of `need` bytes fits; collective
static b200sp_status gather_stage_ensure(b200sp_handle h, cudaStream_t st, size_t need) {
  if (h->gather_stage && h->gather_slice_cap >= need) return B200SP_OK;
  B200SP_CUDA(h, cudaStreamSynchronize(st));
  for (int r = 0; r < P2P_MAX_WORLD; ++r) {
    if (h->peer_gather[r] && h->peer_gather[r] != h->gather_stage) cudaIpcCloseMemHandle(h->peer_gather[r]);
    h->peer_gather[r] = nullptr;
  }
  struct Info {
    unsigned char handle[64];
    int ok, pad;
  } mine, all[P2P_MAX_WORLD];
  memset(&mine, 0, sizeof(mine));
  // every rank must have closed its mappings of the old buffers before anybody frees one
  if (!all_ranks_ok(h, st, 1)) return set_error(h, B200SP_COMM_ERROR, "allgather: staging re-allocation handshake failed");
  if (h->gather_stage) cudaFree(h->gather_stage);
  h->gather_stage = nullptr;
  const size_t cap = ((need + 16 + 255) & ~(size_t)255);
  if (cudaMalloc(&h->gather_stage, 2 * cap) == cudaSuccess) {
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, h->gather_stage) == cudaSuccess) {
      memcpy(mine.handle, &hd, 64);
      mine.ok = 1;
    }
  }
  cudaGetLastError();
  b200sp_status s = comm_allgather_host(h, st, &mine, sizeof(Info), all);
  if (s != B200SP_OK) return s;
  int ok = 1;
  for (int r = 0; r < h->world; ++r) ok = ok && all[r].ok;
  for (int r = 0; r < h->world && ok; ++r) {
    if (r == h->rank) {
      h->peer_gather[r] = h->gather_stage;
      continue;
    }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, all[r].handle, 64);
    if (cudaIpcOpenMemHandle(&h->peer_gather[r], hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      h->peer_gather[r] = nullptr;
      ok = 0;
      cudaGetLastError();
    }
  }
  if (!all_ranks_ok(h, st, ok)) {
    h->gather_slice_cap = 0;
    return set_error(h, B200SP_COMM_ERROR, "allgather: mapping the peers' staging buffers failed on some rank");
  }
  h->gather_slice_cap = cap - 16;
  return B200SP_OK;
}

b200sp_status comm_allgather_slices(b200sp_handle h, cudaStream_t st, void *x_full, const int64_t *slice_offsets,
                                    size_t elem) {
  if (h->world <= 1) return B200SP_OK;
  B200SP_REQUIRE(h, x_full && slice_offsets, "allgather: null argument");
  // the peer-memory kernels move 16-byte words relative to x_full; refusing here (on every path, so that
  // all ranks agree) beats a misaligned-address fault inside the kernel
  B200SP_REQUIRE(h, aligned16(x_full), "allgather: x_full must be 16-byte aligned");
  size_t max_slice = 0;
  for (int r = 0; r < h->world; ++r) {
    B200SP_REQUIRE(h, slice_offsets[r + 1] >= slice_offsets[r], "allgather: slice offsets must ascend");
    const size_t b = (size_t)(slice_offsets[r + 1] - slice_offsets[r]) * elem;
    max_slice = b > max_slice ? b : max_slice;
  }
  char *x = reinterpret_cast<char *>(x_full);
  if (h->p2p_ok) {
    // B200SP_GATHER_PULL=1: the pull protocol (peers read each other's staged slice); default: push
    static const bool pull = [] {
      const char *e = getenv("B200SP_GATHER_PULL");
      return e && e[0] && e[0] != '0';
    }();
    const size_t total = (size_t)slice_offsets[h->world] * elem;
    b200sp_status s = gather_stage_ensure(h, st, pull ? max_slice : total);
    if (s != B200SP_OK) return s;
    GatherArgs g;
    memset(&g, 0, sizeof(g));
    g.x_full = x;
    g.stage_mine = reinterpret_cast<char *>(h->gather_stage);
    for (int r = 0; r < h->world; ++r) {
      g.stage_peer[r] = reinterpret_cast<char *>(h->peer_gather[r]);
      g.mail[r] = reinterpret_cast<Mailbox *>(h->peer_mail[r]);
      g.off[r] = (unsigned long long)slice_offsets[r] * elem;
    }
    g.off[h->world] = (unsigned long long)slice_offsets[h->world] * elem;
    g.epoch = ++h->gather_epoch;
    g.par_off = (size_t)(g.epoch & 1) * (h->gather_slice_cap + 16);
    g.ticket = h->red_counters + 6;
    g.world = h->world;
    g.rank = h->rank;
    if (pull) {
      allgather_pull_kernel<<<h->num_sms * 2, GATHER_BLOCK, 0, st>>>(g);
      B200SP_LAUNCH_CHECK(h, "allgather_pull_kernel");
    } else {
      allgather_push_kernel<<<h->num_sms * 2, GATHER_BLOCK, 0, st>>>(g);
      B200SP_LAUNCH_CHECK(h, "allgather_push_kernel");
    }
    return B200SP_OK;
  }
  ncclComm_t comm = (ncclComm_t)h->nccl_comm;
  const size_t mine_b = (size_t)(slice_offsets[h->rank + 1] - slice_offsets[h->rank]) * elem;
  B200SP_NCCL(h, nccl().GroupStart());
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) continue;
    const size_t b = (size_t)(slice_offsets[r + 1] - slice_offsets[r]) * elem;
    if (mine_b) B200SP_NCCL(h, nccl().Send(x + (size_t)slice_offsets[h->rank] * elem, mine_b, ncclChar, r, comm, st));
    if (b) B200SP_NCCL(h, nccl().Recv(x + (size_t)slice_offsets[r] * elem, b, ncclChar, r, comm, st));
  }
  B200SP_NCCL(h, nccl().GroupEnd());
  h->launches++;
  return B200SP_OK;
}

P2PView comm_p2p_view(b200sp_handle h) {
  P2PView v;
  memset(&v, 0, sizeof(v));
  for (int r = 0; r < h->world && r < P2P_MAX_WORLD; ++r) v.peer[r] = reinterpret_cast<Mailbox *>(h->peer_mail[r]);
  v.mine = reinterpret_cast<Mailbox *>(h->mail);
  v.world = h->world;
  v.rank = h->rank;
  return v;
}

b200sp_status comm_next_solve_id(b200sp_handle h, cudaStream_t st, unsigned long long *id) {
  unsigned long long mine = h->solve_id, all[P2P_MAX_WORLD];
  b200sp_status s = comm_allgather_host(h, st, &mine, sizeof(mine), all);
  if (s != B200SP_OK) return s;
  unsigned long long m = 0;
  for (int r = 0; r < h->world; ++r) m = all[r] > m ? all[r] : m;
  h->solve_id = m + 1;
  *id = h->solve_id;
  return B200SP_OK;
}

b200sp_status comm_p2p_map_windows(b200sp_handle h, cudaStream_t st, void *ws_base, size_t pwin_offset, i64 n,
                                   i64 halo_lo, i64 halo_hi, size_t elem, void **dst_lo, void **dst_hi) {
  B200SP_REQUIRE(h, h->p2p_ok && ws_base && dst_lo && dst_hi, "p2p_map_windows: peer path not available");
  struct WInfo {
    unsigned char handle[64];
    unsigned long long pwin_offset;
    long long n, halo_lo, halo_hi;
    int ok, pad;
  } mine, all[P2P_MAX_WORLD];
  memset(&mine, 0, sizeof(mine));
  cudaIpcMemHandle_t hd;
  if (cudaIpcGetMemHandle(&hd, ws_base) == cudaSuccess) {
    memcpy(mine.handle, &hd, 64);
    mine.ok = 1;
  } else {
    cudaGetLastError();
  }
  mine.pwin_offset = pwin_offset;
  mine.n = n;
  mine.halo_lo = halo_lo;
  mine.halo_hi = halo_hi;
  b200sp_status s = comm_allgather_host(h, st, &mine, sizeof(WInfo), all);
  if (s != B200SP_OK) return s;
  int ok = 1;
  for (int r = 0; r < h->world; ++r) ok = ok && all[r].ok;
  *dst_lo = *dst_hi = nullptr;
  for (int side = 0; side < 2 && ok; ++side) {
    const int nb = side == 0 ? h->rank - 1 : h->rank + 1;
    const i64 my_halo = side == 0 ? halo_lo : halo_hi;
    if (nb < 0 || nb >= h->world || my_halo == 0) continue;
    // halos are symmetric: what I send to a neighbour has the size of what I receive from it
    const long long nb_halo = side == 0 ? all[nb].halo_hi : all[nb].halo_lo;
    if (nb_halo != my_halo) {
      ok = 0;
      break;
    }
    if (!h->nbr_ws[side] || memcmp(h->nbr_ws_handle[side], all[nb].handle, 64) != 0) {
      if (h->nbr_ws[side]) cudaIpcCloseMemHandle(h->nbr_ws[side]);
      h->nbr_ws[side] = nullptr;
      cudaIpcMemHandle_t nh;
      memcpy(&nh, all[nb].handle, 64);
      if (cudaIpcOpenMemHandle(&h->nbr_ws[side], nh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        h->nbr_ws[side] = nullptr;
        cudaGetLastError();
        ok = 0;
        break;
      }
      memcpy(h->nbr_ws_handle[side], all[nb].handle, 64);
    }
    char *base = reinterpret_cast<char *>(h->nbr_ws[side]) + all[nb].pwin_offset;
    if (side == 0)
      *dst_lo = base + (size_t)(all[nb].halo_lo + all[nb].n) * elem;  // the neighbour's halo_hi region
    else
      *dst_hi = base;  // the neighbour's halo_lo region
  }
  if (!all_ranks_ok(h, st, ok))
    return set_error(h, B200SP_COMM_ERROR, "p2p: mapping the neighbours' CG workspaces failed on some rank");
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
  B200SP_REQUIRE(h, world_size <= b200sp::P2P_MAX_WORLD, "comm_init: at most 16 ranks (the GPUs of one node)");
  auto &api = b200sp::nccl();
  if (!api.ok) return b200sp::set_error(h, B200SP_COMM_ERROR, "NCCL library not found (libnccl.so.2)");
  if (h->nccl_comm) {
    b200sp::p2p_teardown(h);
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
  b200sp::p2p_setup(h);  // NVLink peer-memory mailboxes; silently stays on NCCL if unavailable
  return B200SP_OK;
}

int b200sp_comm_p2p_enabled(b200sp_handle h) { return (h && h->p2p_ok) ? 1 : 0; }

int64_t b200sp_comm_timeouts(b200sp_handle h, b200sp_stream stream) {
  if (!h || !h->p2p_ok || !h->mail) return 0;
  unsigned long long t = 0;
  b200sp::Mailbox *m = reinterpret_cast<b200sp::Mailbox *>(h->mail);
  if (cudaMemcpyAsync(&t, &m->timeouts, sizeof(t), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
      cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (int64_t)t;
}

b200sp_status b200sp_comm_destroy(b200sp_handle h) {
  B200SP_CHECK_HANDLE(h);
  if (h->nccl_comm) {
    b200sp::p2p_teardown(h);
    b200sp::nccl().CommDestroy((ncclComm_t)h->nccl_comm);
    h->nccl_comm = nullptr;
  }
  h->world = 1;
  h->rank = 0;
  return B200SP_OK;
}
}
