// cg.cu — conjugate gradients with device-resident scalars, and the generic
// SpMV dispatch on a b200sp_matrix descriptor.
//
// Replaces cusp::krylov::cg with the identity preconditioner
// (cusp/krylov/detail/cg.inl:35-107) + cusp::monitor (cusp/detail/monitor.inl:
// 107-111,178-208).  The reference runs, per iteration, 1 SpMV + dotc + axpy +
// axpy + copy + dotc + axpby + nrm2 = 8 passes, 18*N*sizeof(V) bytes of vector
// traffic and 3 host round trips.  Here one iteration is
//   K1  y = A p  with  <y,p>  fused into the SpMV epilogue           (matrix + 2N)
//   K2  alpha = rz/<y,p>; x += alpha p; r -= alpha y; rz' = <r,r>    (6N)
//       last CTA: beta = rz'/rz, ||r|| = sqrt(rz'), residual log, stop flag
//   K3  p = r + beta p                                               (3N)
// i.e. matrix + 11N, no host round trip: the host only polls a flag every
// `check_interval` iterations; kernels launched after convergence see the flag
// and return immediately, so x holds exactly the iterate at which the
// reference's monitor would have stopped.  Every element update uses the
// reference's expression (alpha*p + x, (-alpha)*y + r, z + beta*p with z == r).
//
// Opt-in (B200SP_CG_FUSE=1), DIA operators through the bulk kernel: a two-kernel iteration
//   K1' y = A p  where  p = r + beta p_old  is rebuilt at every gather and stored once per row into the other of two
//       p buffers, <y,p> fused                                          (matrix + r + p_old + p + y = matrix + 4N)
//   K2  as above, reading the p that K1' stored                          (6N)
// i.e. matrix + 10N: the direction update costs no pass of its own, and between GPUs r's edge planes travel instead
// of p's.  Same expressions, same order: the iterates are bit-identical to the three-kernel form (tests).  Measured on
// B200 it loses, hence off by default: poisson7pt 512^3 fp64 3.03 ms / iteration against 2.95 (256^3, ncu: K1' 277 us
// for 1.47 GB of DRAM traffic = 5.3 TB/s, where K1 + K3 take 187 + 53 us for 1.53 GB; the doubled gathers add 40 % to
// the instruction count of a kernel whose issue slots are half used and whose warps wait on the long scoreboard, and
// neither more warps (256 x 1: 3.31 ms), a third stage (3.10) nor runs of consecutive tiles per CTA for L1 reuse
// (3.08 - 3.25) buy it back; profiles/r03_cg_fused_direction.md).
//
// Multi-GPU (row-block partition): halo planes of p are exchanged
// with ncclSend/ncclRecv before K1 and the two scalars are all-reduced with
// NCCL between the kernels; the scalar step then runs as its own 1-thread kernel.
#include "common.cuh"
#include "comm.h"

#include <cooperative_groups.h>

namespace b200sp {

template <typename T>
b200sp_status spmv_csr(b200sp_handle, cudaStream_t, i64, i64, i64, const int *, const int *, const T *,
                       const T *, T *, int, const b200sp_cfg *, const T *, T *);
template <typename T>
b200sp_status spmv_ell(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *, const T *, const int *,
                       const T *, T *, int, const b200sp_cfg *, const T *, T *);
template <typename T>
b200sp_status spmv_dia(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *, const T *, const T *,
                       T *, int, const b200sp_cfg *, const T *, T *);
template <typename T>
b200sp_status spmv_coo(b200sp_handle, cudaStream_t, i64, i64, i64, const int *, const int *, const T *,
                       const T *, T *, int, const b200sp_cfg *);
template <typename T>
b200sp_status spmv_hyb(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *, const T *, i64,
                       const int *, const int *, const T *, const T *, T *, int, const b200sp_cfg *,
                       const b200sp_cfg *);
template <typename T, int MODE>
b200sp_status reduce(b200sp_handle, cudaStream_t, i64, const T *, const T *, T *, T *);
int tune_lookup_on(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, b200sp_cfg *cfg);  // api.cu
template <typename T>
b200sp_status spmv_dia_xchg(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *, const T *, const T *,
                            T *, int, const b200sp_cfg *, const T *, T *, const FusedXchg *, int *);
bool dia_can_fuse_xchg(i64 rows, i64 ndiag, i64 pitch, const void *vals, size_t elem, const b200sp_cfg *cfg);
bool dia_can_fuse_direction(i64 rows, i64 ndiag, i64 pitch, const void *vals, size_t elem, const b200sp_cfg *cfg);
template <typename T>
b200sp_status spmv_dia_fused_direction(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *, const T *, T *,
                                       const b200sp_cfg *, const T *, const T *, T *, const T *, const int *, int, i64, i64,
                                       i64, T *, const FusedXchg *);

// y = A x (+ optional fused <y, dotv>).  Formats whose kernel has no fused
// epilogue (COO / HYB) get a separate deterministic dot kernel.
template <typename T>
b200sp_status spmv_any(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, const T *x, T *y,
                       int accumulate, const b200sp_cfg *cfg, const T *dotv, T *dot_result) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A != nullptr, "spmv: null matrix descriptor");
  const T *vals = reinterpret_cast<const T *>(A->values);
  b200sp_cfg cached;
  if (!cfg && tune_lookup_on(h, st, A, &cached)) cfg = &cached;
  b200sp_status s;
  switch (A->format) {
    case B200SP_FMT_CSR:
      return spmv_csr<T>(h, st, A->num_rows, A->num_cols, A->num_entries, A->row_offsets,
                         A->column_indices, vals, x, y, accumulate, cfg, dotv, dot_result);
    case B200SP_FMT_ELL:
      return spmv_ell<T>(h, st, A->num_rows, A->num_cols, A->num_cols_per_row, A->pitch,
                         A->column_indices, vals, nullptr, x, y, accumulate, cfg, dotv, dot_result);
    case B200SP_FMT_ELLR:
      B200SP_REQUIRE(h, A->row_offsets != nullptr, "ellr: row_lengths (row_offsets field) is null");
      return spmv_ell<T>(h, st, A->num_rows, A->num_cols, A->num_cols_per_row, A->pitch,
                         A->column_indices, vals, A->row_offsets, x, y, accumulate, cfg, dotv, dot_result);
    case B200SP_FMT_DIA:
      return spmv_dia<T>(h, st, A->num_rows, A->num_cols, A->num_cols_per_row, A->pitch,
                         A->diagonal_offsets, vals, x, y, accumulate, cfg, dotv, dot_result);
    case B200SP_FMT_COO:
      s = spmv_coo<T>(h, st, A->num_rows, A->num_cols, A->num_entries, A->row_indices, A->column_indices,
                      vals, x, y, accumulate, cfg);
      break;
    case B200SP_FMT_HYB:
      s = spmv_hyb<T>(h, st, A->num_rows, A->num_cols, A->num_cols_per_row, A->pitch, A->column_indices,
                      vals, A->coo_num_entries, A->coo_row_indices, A->coo_column_indices,
                      reinterpret_cast<const T *>(A->coo_values), x, y, accumulate, cfg, nullptr);
      break;
    default:
      return set_error(h, B200SP_INVALID_INPUT, "spmv: unknown format %d", (int)A->format);
  }
  if (s != B200SP_OK) return s;
  if (dotv) return reduce<T, 0>(h, st, A->num_rows, y, dotv, dot_result, nullptr);
  return B200SP_OK;
}

template b200sp_status spmv_any<float>(b200sp_handle, cudaStream_t, const b200sp_matrix *, const float *,
                                       float *, int, const b200sp_cfg *, const float *, float *);
template b200sp_status spmv_any<double>(b200sp_handle, cudaStream_t, const b200sp_matrix *, const double *,
                                        double *, int, const b200sp_cfg *, const double *, double *);

// ---------------------------------------------------------------------------
// CG state in device memory
// ---------------------------------------------------------------------------
template <typename T>
struct CgState {
  T rz;      // <r,z> of the current iterate (z == r)
  T yp;      // <A p, p>
  T rz_new;  // scratch for the reduction
  T beta;
  T tol;     // absolute + relative*||b||
  T bnorm;
  T rnorm;
  int iter;
  int limit;
  int done;       // 1 once the monitor says finished
  int converged;  // rnorm <= tol
  int nres;       // residuals recorded
};

// the monitor step (cusp/detail/monitor.inl:178-208): record ||r||, decide
template <typename T>
__device__ __forceinline__ void monitor_step(CgState<T> *S, double *residuals) {
  const T rn = (T)sqrt((double)S->rz);
  S->rnorm = rn;
  residuals[S->nres++] = (double)rn;
  if (rn <= S->tol) {
    S->converged = 1;
    S->done = 1;
  } else if (S->iter >= S->limit) {
    S->done = 1;
  }
}

constexpr int CG_BLOCK = 256;
constexpr int CG_UNROLL = 4;

static inline i64 cg_grid(b200sp_handle h, i64 n) {
  i64 g = ceil_div(n, (i64)CG_BLOCK * CG_UNROLL);
  const i64 cap = (i64)h->num_sms * 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : g;
}

// r = b - y (axpby(b,y,r,1,-1)); p = r (z = M r = r; p = z); rz = <r,r>
template <typename T, bool DIST>
__global__ void __launch_bounds__(CG_BLOCK) cg_init_kernel(i64 n, const T *b, const T *y, T *r, T *p,
                                                           CgState<T> *S, T *partials, unsigned int *ticket,
                                                           double *residuals) {
  __shared__ T s_red[32];
  T acc = T(0);
  const i64 stride = (i64)gridDim.x * CG_BLOCK;
  for (i64 i = (i64)blockIdx.x * CG_BLOCK + threadIdx.x; i < n; i += stride) {
    const T ri = T(1) * b[i] + T(-1) * y[i];
    r[i] = ri;
    p[i] = ri;
    acc = acc + ri * ri;
  }
  T bs = block_sum<CG_BLOCK>(acc, s_red);
  grid_reduce_finish<CG_BLOCK>(bs, partials, ticket, s_red, [&](T total) {
    S->rz = total;
    if (!DIST) monitor_step(S, residuals);
  });
}

// x += alpha p ; r -= alpha y ; rz_new = <r,r> ; then the scalar step
// Launched with programmatic stream serialization (see pdl_wait in common.cuh): the first batch of p, x, r — none of
// which the preceding SpMV writes — is requested before the wait, so those loads and the launch latency overlap the
// SpMV's tail; y and the scalars are read after it.
template <typename T, bool DIST>
__global__ void __launch_bounds__(CG_BLOCK) cg_update_kernel(i64 n, const T *p, const T *y, T *x, T *r,
                                                             CgState<T> *S, T *partials,
                                                             unsigned int *ticket, double *residuals) {
  __shared__ T s_red[32];
  const i64 stride = (i64)gridDim.x * CG_BLOCK;
  i64 i = (i64)blockIdx.x * CG_BLOCK + threadIdx.x;
  T pv[CG_UNROLL], yv[CG_UNROLL], xv[CG_UNROLL], rv[CG_UNROLL];
  const bool first = i + (CG_UNROLL - 1) * stride < n;
  if (first) {
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      pv[u] = p[i + u * stride];
      xv[u] = x[i + u * stride];
      rv[u] = r[i + u * stride];
    }
  }
  pdl_wait();
  pdl_trigger();
  if (S->done) return;
  const T alpha = S->rz / S->yp;
  const T nalpha = -alpha;
  T acc = T(0);
  if (first) {
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) yv[u] = y[i + u * stride];
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      x[i + u * stride] = alpha * pv[u] + xv[u];
      const T rn = nalpha * yv[u] + rv[u];
      r[i + u * stride] = rn;
      acc = acc + rn * rn;
    }
    i += CG_UNROLL * stride;
  }
  for (; i + (CG_UNROLL - 1) * stride < n; i += CG_UNROLL * stride) {
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      pv[u] = p[i + u * stride];
      yv[u] = y[i + u * stride];
      xv[u] = x[i + u * stride];
      rv[u] = r[i + u * stride];
    }
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      x[i + u * stride] = alpha * pv[u] + xv[u];
      const T rn = nalpha * yv[u] + rv[u];
      r[i + u * stride] = rn;
      acc = acc + rn * rn;
    }
  }
  for (; i < n; i += stride) {
    x[i] = alpha * p[i] + x[i];
    const T rn = nalpha * y[i] + r[i];
    r[i] = rn;
    acc = acc + rn * rn;
  }
  T bs = block_sum<CG_BLOCK>(acc, s_red);
  grid_reduce_finish<CG_BLOCK>(bs, partials, ticket, s_red, [&](T total) {
    S->rz_new = total;
    if (!DIST) {
      S->beta = total / S->rz;
      S->rz = total;
      S->iter += 1;
      monitor_step(S, residuals);
    }
  });
}

// distributed: after the all-reduce of rz_new
template <typename T>
__global__ void cg_scalar_step_kernel(CgState<T> *S, double *residuals, int first) {
  if (first) {
    monitor_step(S, residuals);
    return;
  }
  if (S->done) return;
  S->beta = S->rz_new / S->rz;
  S->rz = S->rz_new;
  S->iter += 1;
  monitor_step(S, residuals);
}

// p = z + beta p  (axpby(z,p,p,1,beta), z == r).  p is not written by the preceding update kernel: its loads are
// issued before the dependency wait.
template <typename T>
__global__ void __launch_bounds__(CG_BLOCK) cg_direction_kernel(i64 n, const T *r, T *p, const CgState<T> *S) {
  const i64 base = (i64)blockIdx.x * (CG_BLOCK * CG_UNROLL) + threadIdx.x;
  T rv[CG_UNROLL], pv[CG_UNROLL];
#pragma unroll
  for (int u = 0; u < CG_UNROLL; ++u) {
    const i64 i = base + (i64)u * CG_BLOCK;
    pv[u] = i < n ? p[i] : T(0);
  }
  pdl_wait();
  pdl_trigger();
  if (S->done) return;
  const T beta = S->beta;
#pragma unroll
  for (int u = 0; u < CG_UNROLL; ++u) {
    const i64 i = base + (i64)u * CG_BLOCK;
    rv[u] = i < n ? r[i] : T(0);
  }
#pragma unroll
  for (int u = 0; u < CG_UNROLL; ++u) {
    const i64 i = base + (i64)u * CG_BLOCK;
    if (i < n) p[i] = T(1) * rv[u] + beta * pv[u];
  }
}


// ---------------------------------------------------------------------------
// NVLink peer-memory variants (row-block partitioned CG, comm.h: Mailbox / P2PView)
//
// One CG iteration on N GPUs is the same three kernels as on one GPU, and the
// exchange steps ride inside them as stores into peer memory:
//   K1  y = A p, local <y,p>                                  (unchanged SpMV kernel)
//   K2  CTA 0 stores the local <y,p> into slot[0][rank] of EVERY rank's mailbox; every
//       CTA spins until all N slots carry this iteration's epoch and adds them in rank
//       order (bit-identical alpha everywhere); x += alpha p; r -= alpha y; the last CTA
//       stores the local <r,r> into slot[1][rank] of every mailbox
//   K3  every CTA sums slot[1][*] -> beta; p = r + beta p, and the threads that own the
//       first / last halo plane also store it straight into the neighbour's p window
//       over NVLink; the last CTA publishes halo_flag = epoch to both neighbours,
//       performs the monitor step, and leaves only when both neighbours' planes have
//       landed here, so the next K1 may read its halos without any further sync.
// No NCCL call, no extra launch, no host round trip per iteration.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// lane r < world: publish `v` tagged `tag` in rank r's mailbox (two atomic 8-byte words)
__device__ __forceinline__ void p2p_publish(const P2PView &c, int ch, double v, unsigned tag, int r) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  MailSlot *s = &c.peer[r]->slot[ch][c.rank];
  st_relaxed_sys(&s->w[0], ((unsigned long long)tag << 32) | (bits & 0xffffffffull));
  st_relaxed_sys(&s->w[1], ((unsigned long long)tag << 32) | (bits >> 32));
}
// first warp of a CTA: lane r polls rank r's slot, then the values are added in rank
// order (every rank forms the bit-identical sum).  Result valid in every lane.
template <typename T>
__device__ __forceinline__ T p2p_sum_warp(const P2PView &c, int ch, unsigned tag) {
  const int lane = threadIdx.x & 31;
  double mine = 0.0;
  if (lane < c.world) {
    const MailSlot *s = &c.mine->slot[ch][lane];
    unsigned long long w0, w1;
    SpinGuard guard;
    do {
      w0 = ld_relaxed_sys(&s->w[0]);
      w1 = ld_relaxed_sys(&s->w[1]);
    } while (((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) && !guard.expired(c.mine));
    mine = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
  }
  T sum = T(0);
  for (int r = 0; r < c.world; ++r) sum = sum + (T)__shfl_sync(0xffffffffu, mine, r);
  return sum;
}

// trace slots per iteration (B200SP_CG_TRACE): globaltimer stamps of CTA 0 / the finishing CTA
enum { TR_K2_ENTER = 0, TR_K2_POLLED = 1, TR_K2_END = 2, TR_K3_ENTER = 3, TR_K3_POLLED = 4, TR_K3_END = 5, TR_SLOTS = 8 };

template <typename T>
__global__ void __launch_bounds__(CG_BLOCK) cg_update_p2p_kernel(i64 n, const T *p, const T *y, T *x, T *r,
                                                                 CgState<T> *S, T *partials, unsigned int *ticket,
                                                                 P2PView c, unsigned tag, unsigned long long *trace) {
  __shared__ T s_red[32];
  __shared__ T s_yp;
  __shared__ double s_pub;
  __shared__ int s_fin;
  // before the dependency wait: the first batch of p, x, r (the preceding SpMV writes none of them)
  const i64 stride = (i64)gridDim.x * CG_BLOCK;
  i64 i = (i64)blockIdx.x * CG_BLOCK + threadIdx.x;
  T pv[CG_UNROLL], yv[CG_UNROLL], xv[CG_UNROLL], rv[CG_UNROLL];
  const bool first = i + (CG_UNROLL - 1) * stride < n;
  if (first) {
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      pv[u] = p[i + u * stride];
      xv[u] = x[i + u * stride];
      rv[u] = r[i + u * stride];
    }
  }
  pdl_wait();
  pdl_trigger();
  if (S->done) return;
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[TR_K2_ENTER] = global_timer_ns();
  if (threadIdx.x == 0) s_fin = 0;
  if (blockIdx.x == 0 && threadIdx.x < c.world) p2p_publish(c, 0, (double)S->yp, tag, threadIdx.x);
  if (first) {  // y of the first batch is on its way while the slots are polled
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) yv[u] = y[i + u * stride];
  }
  if (threadIdx.x < 32) {
    const T t = p2p_sum_warp<T>(c, 0, tag);
    if (threadIdx.x == 0) s_yp = t;
  }
  __syncthreads();
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[TR_K2_POLLED] = global_timer_ns();
  const T alpha = S->rz / s_yp;
  const T nalpha = -alpha;
  T acc = T(0);
  if (first) {
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      x[i + u * stride] = alpha * pv[u] + xv[u];
      const T rn = nalpha * yv[u] + rv[u];
      r[i + u * stride] = rn;
      acc = acc + rn * rn;
    }
    i += CG_UNROLL * stride;
  }
  for (; i + (CG_UNROLL - 1) * stride < n; i += CG_UNROLL * stride) {
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      pv[u] = p[i + u * stride];
      yv[u] = y[i + u * stride];
      xv[u] = x[i + u * stride];
      rv[u] = r[i + u * stride];
    }
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      x[i + u * stride] = alpha * pv[u] + xv[u];
      const T rn = nalpha * yv[u] + rv[u];
      r[i + u * stride] = rn;
      acc = acc + rn * rn;
    }
  }
  for (; i < n; i += stride) {
    x[i] = alpha * p[i] + x[i];
    const T rn = nalpha * y[i] + r[i];
    r[i] = rn;
    acc = acc + rn * rn;
  }
  T bs = block_sum<CG_BLOCK>(acc, s_red);
  grid_reduce_finish<CG_BLOCK>(bs, partials, ticket, s_red, [&](T total) {
    S->rz_new = total;  // local part; the global sum is formed by every CTA of K3
    s_pub = (double)total;
    s_fin = 1;
  });
  __syncthreads();
  if (s_fin && threadIdx.x < c.world) p2p_publish(c, 1, s_pub, tag, threadIdx.x);  // one lane per peer
  if (trace && s_fin && threadIdx.x == 0) trace[TR_K2_END] = global_timer_ns();
}

template <typename T>
__global__ void __launch_bounds__(CG_BLOCK) cg_direction_p2p_kernel(i64 n, const T *r, T *p, CgState<T> *S,
                                                                    unsigned int *ticket, double *residuals,
                                                                    P2PView c, unsigned tag,
                                                                    unsigned long long epoch, i64 halo_lo,
                                                                    i64 halo_hi, T *dst_lo, T *dst_hi,
                                                                    int defer_wait, unsigned long long *trace) {
  __shared__ T s_rz;
  __shared__ bool is_last;
  // before the dependency wait: the first batch of p (the preceding update kernel does not write p)
  const i64 stride = (i64)gridDim.x * CG_BLOCK;
  i64 i = (i64)blockIdx.x * CG_BLOCK + threadIdx.x;
  T rv[CG_UNROLL], pv[CG_UNROLL];
  const bool first = i + (CG_UNROLL - 1) * stride < n;
  if (first) {
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) pv[u] = p[i + u * stride];
  }
  pdl_wait();
  pdl_trigger();
  if (S->done) {
    // the solve is over, but a next K1 that waits for this epoch may already be queued on a neighbour
    if (defer_wait && blockIdx.x == 0 && threadIdx.x == 0) {
      if (dst_lo) st_release_sys(&c.peer[c.rank - 1]->halo_flag[1], epoch);
      if (dst_hi) st_release_sys(&c.peer[c.rank + 1]->halo_flag[0], epoch);
    }
    return;
  }
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[TR_K3_ENTER] = global_timer_ns();
  if (first) {  // r of the first batch is on its way while the slots are polled
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) rv[u] = r[i + u * stride];
  }
  if (threadIdx.x < 32) {
    const T t = p2p_sum_warp<T>(c, 1, tag);
    if (threadIdx.x == 0) s_rz = t;
  }
  __syncthreads();
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[TR_K3_POLLED] = global_timer_ns();
  const T rz_new = s_rz;
  const T beta = rz_new / S->rz;
  const i64 hi_begin = n - halo_hi;
  bool remote = false;
  for (bool pre = first; i + (CG_UNROLL - 1) * stride < n; i += CG_UNROLL * stride, pre = false) {
    if (!pre) {
#pragma unroll
      for (int u = 0; u < CG_UNROLL; ++u) {
        rv[u] = r[i + u * stride];
        pv[u] = p[i + u * stride];
      }
    }
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      const i64 j = i + u * stride;
      const T v = T(1) * rv[u] + beta * pv[u];
      p[j] = v;
      if (dst_lo && j < halo_lo) {  // rank-1's upper halo, straight over NVLink
        dst_lo[j] = v;
        remote = true;
      }
      if (dst_hi && j >= hi_begin) {  // rank+1's lower halo
        dst_hi[j - hi_begin] = v;
        remote = true;
      }
    }
  }
  for (; i < n; i += stride) {
    const T v = T(1) * r[i] + beta * p[i];
    p[i] = v;
    if (dst_lo && i < halo_lo) {
      dst_lo[i] = v;
      remote = true;
    }
    if (dst_hi && i >= hi_begin) {
      dst_hi[i - hi_begin] = v;
      remote = true;
    }
  }
  // last CTA: publish the halo epoch, do the monitor step, wait for the neighbours' planes.
  // Only threads that stored into peer memory pay for a system-scope fence.
  if (remote) __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence_system();
    if (dst_lo) st_release_sys(&c.peer[c.rank - 1]->halo_flag[1], epoch);  // I am its rank+1
    if (dst_hi) st_release_sys(&c.peer[c.rank + 1]->halo_flag[0], epoch);  // I am its rank-1
    S->beta = beta;
    S->rz = rz_new;
    S->iter += 1;
    monitor_step(S, residuals);
    // defer_wait: the next K1 (DIA bulk kernel) visits its halo-reading tiles last and waits
    // for these flags there, so the arrival latency hides behind its interior tiles
    SpinGuard guard;
    if (dst_lo && !defer_wait)
      while (ld_acquire_sys(&c.mine->halo_flag[0]) < epoch && !guard.expired(c.mine)) {
      }
    if (dst_hi && !defer_wait)
      while (ld_acquire_sys(&c.mine->halo_flag[1]) < epoch && !guard.expired(c.mine)) {
      }
    *ticket = 0;
    __threadfence();
    if (trace) trace[TR_K3_END] = global_timer_ns();
  }
}

// Two-kernel iteration on the peer-memory path (K1' = dia_bulk_kernel<..., FUSED>): the update kernel also carries
// what the direction kernel did between GPUs.  The threads that own the first / last plane store the new r straight
// into the neighbour's r window (K1' rebuilds the halo rows of p from it: p never travels); the CTA that finishes the
// local <r,r> publishes it to every rank, raises the neighbours' halo flags, waits for the other ranks' parts (they
// finish within the skew of the ranks), and does the scalar + monitor step, so the next K1' reads beta and the stop
// flag from S like on one GPU.
template <typename T>
__global__ void __launch_bounds__(CG_BLOCK) cg_update_push_p2p_kernel(i64 n, const T *p, const T *y, T *x, T *r,
                                                                      CgState<T> *S, T *partials, unsigned int *ticket,
                                                                      double *residuals, P2PView c, unsigned tag,
                                                                      unsigned long long epoch, i64 halo_lo, i64 halo_hi,
                                                                      T *dst_lo, T *dst_hi, unsigned long long *trace) {
  __shared__ T s_red[32];
  __shared__ T s_yp;
  __shared__ double s_pub;
  __shared__ int s_fin;
  if (S->done) return;
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[TR_K2_ENTER] = global_timer_ns();
  const i64 stride = (i64)gridDim.x * CG_BLOCK;
  i64 i = (i64)blockIdx.x * CG_BLOCK + threadIdx.x;
  T pv[CG_UNROLL], yv[CG_UNROLL], xv[CG_UNROLL], rv[CG_UNROLL];
  const bool first = i + (CG_UNROLL - 1) * stride < n;
  if (threadIdx.x == 0) s_fin = 0;
  if (blockIdx.x == 0 && threadIdx.x < c.world) p2p_publish(c, 0, (double)S->yp, tag, threadIdx.x);
  if (first) {  // the first batch is on its way while the slots are polled
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      pv[u] = p[i + u * stride];
      xv[u] = x[i + u * stride];
      rv[u] = r[i + u * stride];
      yv[u] = y[i + u * stride];
    }
  }
  if (threadIdx.x < 32) {
    const T t = p2p_sum_warp<T>(c, 0, tag);
    if (threadIdx.x == 0) s_yp = t;
  }
  __syncthreads();
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[TR_K2_POLLED] = global_timer_ns();
  const T alpha = S->rz / s_yp;
  const T nalpha = -alpha;
  const i64 hi_begin = n - halo_hi;
  T acc = T(0);
  bool remote = false;
  auto store_r = [&](i64 j, T rn) {
    r[j] = rn;
    if (dst_lo && j < halo_lo) {  // rank-1's upper halo, straight over NVLink
      dst_lo[j] = rn;
      remote = true;
    }
    if (dst_hi && j >= hi_begin) {  // rank+1's lower halo
      dst_hi[j - hi_begin] = rn;
      remote = true;
    }
  };
  for (bool pre = first; i + (CG_UNROLL - 1) * stride < n; i += CG_UNROLL * stride, pre = false) {
    if (!pre) {
#pragma unroll
      for (int u = 0; u < CG_UNROLL; ++u) {
        pv[u] = p[i + u * stride];
        yv[u] = y[i + u * stride];
        xv[u] = x[i + u * stride];
        rv[u] = r[i + u * stride];
      }
    }
#pragma unroll
    for (int u = 0; u < CG_UNROLL; ++u) {
      x[i + u * stride] = alpha * pv[u] + xv[u];
      const T rn = nalpha * yv[u] + rv[u];
      store_r(i + u * stride, rn);
      acc = acc + rn * rn;
    }
  }
  for (; i < n; i += stride) {
    x[i] = alpha * p[i] + x[i];
    const T rn = nalpha * y[i] + r[i];
    store_r(i, rn);
    acc = acc + rn * rn;
  }
  if (remote) __threadfence_system();  // only the threads that stored into peer memory pay for it
  T bs = block_sum<CG_BLOCK>(acc, s_red);
  grid_reduce_finish<CG_BLOCK>(bs, partials, ticket, s_red, [&](T total) {
    S->rz_new = total;  // local part
    s_pub = (double)total;
    s_fin = 1;
  });
  __syncthreads();
  if (!s_fin) return;
  // the CTA that took the last ticket: every CTA's stores (local and peer) are ordered before its ticket
  if (threadIdx.x < c.world) p2p_publish(c, 1, s_pub, tag, threadIdx.x);  // one lane per peer
  if (threadIdx.x == 0) {  // the thread that saw the last ticket
    __threadfence_system();
    if (dst_lo) st_release_sys(&c.peer[c.rank - 1]->halo_flag[1], epoch);  // I am its rank+1
    if (dst_hi) st_release_sys(&c.peer[c.rank + 1]->halo_flag[0], epoch);  // I am its rank-1
  }
  if (threadIdx.x < 32) {
    __syncwarp();
    const T rz_new = p2p_sum_warp<T>(c, 1, tag);
    if (threadIdx.x == 0) {
      S->beta = rz_new / S->rz;
      S->rz = rz_new;
      S->iter += 1;
      monitor_step(S, residuals);
      __threadfence();
      if (trace) {
        const unsigned long long t = global_timer_ns();
        trace[TR_K2_END] = trace[TR_K3_ENTER] = trace[TR_K3_POLLED] = trace[TR_K3_END] = t;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Small systems (CSR, one GPU): the whole solve in ONE persistent cooperative kernel.
//
// poisson5pt 512^2 is 21 MB of matrix and 2 MB per vector: launch by launch an iteration costs 25 us, replayed from a
// CUDA graph 19 us, of which the kernels' own work is a fraction.  Here every CTA (one of 1024 threads per SM) owns a
// contiguous block of rows and keeps its slice of the matrix (values, columns, offsets) and its rows of x, r, p, y in
// shared memory for the whole solve; per iteration only r and p pass through L2 (the neighbours' gathers) and two grid
// barriers are paid: one behind each dot product.  The direction update needs no barrier of its own: the product
// rebuilds  p[j] = r[j] + beta p_old[j]  where it gathers (both already visible behind the <r,r> barrier), the
// two-kernel idea of B200SP_CG_FUSE, which does pay here because it saves a barrier rather than a pass over HBM.
// Same expressions and the same per-row order as the three-kernel form; the dot products are grouped by CTA
// (tolerance-level difference, like between any two grid sizes).  The stopping rule runs on the device
// (monitor_step's arithmetic, evaluated identically by every CTA); the host is not involved until the end.
// A slice that does not fit in shared memory raises *fallback and the solve takes the ordinary path.
// ---------------------------------------------------------------------------
constexpr int CGS_BLOCK = 1024;

template <typename T>
struct CgSmallArgs {
  i64 n, rows_per_cta;
  const int *Ap, *Aj;
  const T *Ax;
  T *x;
  const T *b;
  T *r_g, *p_g0, *p_g1;  // global mirrors of r and of the two direction buffers (n each)
  T *part_a, *part_b;    // gridDim.x partial sums each: the two alternating reductions of an iteration
  CgState<T> *S;
  double *residuals;
  int *fallback;
  unsigned smem_bytes;
};

// Grid-wide sum behind a grid barrier: every CTA stores its partial, cooperative_groups' grid barrier, then every CTA
// adds all partials itself in index order (the same order everywhere).  Two hand-rolled replacements that fold the
// barrier into the reduction were measured on B200 (148 CTAs, poisson5pt 512^2) and lost: every CTA polling every
// CTA's tagged partial (mailbox words as on the multi-GPU path) 16.4 us per iteration — 148 x 148 polls on 19 L2
// lines; a ticket whose last taker sums and publishes one tagged total that everybody polls 16.3 us; this form 10.4 us.
template <typename T>
__device__ __forceinline__ T cgs_allreduce(cooperative_groups::grid_group &grid, T *partials, T block_value /* thread 0 */,
                                           T *s_red) {
  if (threadIdx.x == 0) __stcg(partials + blockIdx.x, block_value);
  __threadfence();
  grid.sync();
  if (threadIdx.x < 32) {
    T v = T(0);
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) v = v + __ldcg(partials + i);
    v = warp_sum(v);
    if (threadIdx.x == 0) s_red[0] = v;
  }
  __syncthreads();
  const T t = s_red[0];
  __syncthreads();
  return t;
}

template <typename T>
__global__ void __launch_bounds__(CGS_BLOCK, 1) cg_small_csr_kernel(CgSmallArgs<T> a) {
  namespace cgx = cooperative_groups;
  cgx::grid_group grid = cgx::this_grid();
  extern __shared__ __align__(16) unsigned char cgs_smem[];
  __shared__ T s_red[32];
  const int tid = threadIdx.x;
  const i64 r0 = min(a.n, (i64)blockIdx.x * a.rows_per_cta), r1 = min(a.n, r0 + a.rows_per_cta);
  const int rc = (int)(r1 - r0);
  const int e0 = rc > 0 ? a.Ap[r0] : 0, e1 = rc > 0 ? a.Ap[r1] : 0;
  const int nz = e1 - e0;
  // shared memory: Ax[nz] x[rc] r[rc] p[rc] y[rc] (T) | Aj[nz] off[rc + 1] (int)
  const size_t need = ((size_t)nz + 4 * (size_t)rc) * sizeof(T) + ((size_t)nz + rc + 1) * sizeof(int) + 16;
  if (need > a.smem_bytes && tid == 0) atomicExch(a.fallback, 1);
  __threadfence();
  grid.sync();
  if (__ldcg(a.fallback)) return;
  T *s_ax = reinterpret_cast<T *>(cgs_smem);
  T *s_x = s_ax + nz, *s_r = s_x + rc, *s_p = s_r + rc, *s_y = s_p + rc;
  int *s_aj = reinterpret_cast<int *>(s_y + rc);
  int *s_off = s_aj + nz;
  for (int k = tid; k < nz; k += CGS_BLOCK) {
    s_ax[k] = a.Ax[e0 + k];
    s_aj[k] = a.Aj[e0 + k];
  }
  for (int i = tid; i <= rc; i += CGS_BLOCK) s_off[i] = (rc > 0 ? a.Ap[r0 + i] : 0) - e0;
  __syncthreads();

  CgState<T> *S = a.S;
  const T tol = S->tol;
  const int limit = S->limit;
  int iter = 0, nres = 0, converged = 0;
  // cusp::monitor::finished (monitor.inl:178-208), evaluated by every CTA; CTA 0 keeps the log
  T rnorm = T(0);
  auto finished = [&](T rz) -> bool {
    rnorm = (T)sqrt((double)rz);
    if (blockIdx.x == 0 && tid == 0) a.residuals[nres] = (double)rnorm;
    ++nres;
    if (rnorm <= tol) {
      converged = 1;
      return true;
    }
    return iter >= limit;
  };

  // y = A x0 ; r = b - y ; rz = <r,r>   (cg.inl:63-72)
  T acc = T(0);
  for (int i = tid; i < rc; i += CGS_BLOCK) {
    T sum = T(0);
    for (int k = s_off[i]; k < s_off[i + 1]; ++k) sum = sum + s_ax[k] * a.x[s_aj[k]];
    const T ri = T(1) * a.b[r0 + i] + T(-1) * sum;
    s_x[i] = a.x[r0 + i];
    s_r[i] = ri;
    s_p[i] = ri;
    __stcg(a.r_g + r0 + i, ri);
    acc = acc + ri * ri;
  }
  T bs = block_sum<CGS_BLOCK>(acc, s_red);
  T rz = cgs_allreduce(grid, a.part_b, bs, s_red);
  bool done = finished(rz);

  T beta = T(0);
  const T *p_old = a.r_g;  // first iteration: beta = 0 and p_old = r, so p = r
  T *p_new = a.p_g0;
  while (!done) {
    // y = A p with p rebuilt at the gathers; the own rows' p kept in shared memory and mirrored for the neighbours
    acc = T(0);
    for (int i = tid; i < rc; i += CGS_BLOCK) {
      const T pn = T(1) * s_r[i] + beta * s_p[i];
      T sum = T(0);
      for (int k = s_off[i]; k < s_off[i + 1]; ++k) {
        const int j = s_aj[k];
        const T pj = T(1) * __ldcg(a.r_g + j) + beta * __ldcg(p_old + j);
        sum = sum + s_ax[k] * pj;
      }
      s_p[i] = pn;
      __stcg(p_new + r0 + i, pn);
      s_y[i] = sum;
      acc = acc + sum * pn;
    }
    bs = block_sum<CGS_BLOCK>(acc, s_red);
    const T yp = cgs_allreduce(grid, a.part_a, bs, s_red);
    // x += alpha p ; r -= alpha y ; <r,r>
    const T alpha = rz / yp;
    const T nalpha = -alpha;
    acc = T(0);
    for (int i = tid; i < rc; i += CGS_BLOCK) {
      s_x[i] = alpha * s_p[i] + s_x[i];
      const T rn = nalpha * s_y[i] + s_r[i];
      s_r[i] = rn;
      __stcg(a.r_g + r0 + i, rn);
      acc = acc + rn * rn;
    }
    bs = block_sum<CGS_BLOCK>(acc, s_red);
    const T rz_new = cgs_allreduce(grid, a.part_b, bs, s_red);
    beta = rz_new / rz;
    rz = rz_new;
    ++iter;
    done = finished(rz);
    p_old = p_new;
    p_new = (p_new == a.p_g0) ? a.p_g1 : a.p_g0;
  }
  for (int i = tid; i < rc; i += CGS_BLOCK) a.x[r0 + i] = s_x[i];
  if (blockIdx.x == 0 && tid == 0) {
    S->rz = rz;
    S->beta = beta;
    S->rnorm = rnorm;
    S->iter = iter;
    S->nres = nres;
    S->converged = converged;
    S->done = 1;
  }
}

template <typename T>
__global__ void cg_setup_state_kernel(CgState<T> *S, const T *bnorm, double rel, double abs_tol, int limit) {
  S->bnorm = *bnorm;
  // monitor::tolerance(): absolute + relative*b_norm evaluated in Real
  S->tol = (T)abs_tol + (T)rel * (*bnorm);
  S->iter = 0;
  S->limit = limit;
  S->done = 0;
  S->converged = 0;
  S->nres = 0;
  S->beta = T(0);
  S->yp = T(1);
  S->rz = T(0);
  S->rz_new = T(0);
  S->rnorm = T(0);
}

// kernels that must not run once `done` is set are gated inside; the SpMV is
// not gated (its output y is scratch) — it is simply not launched after the
// host has seen the flag.
template <typename T>
static b200sp_status cg_impl(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, const b200sp_halo *halo,
                             T *x, const T *b, const b200sp_cg_params *params, const b200sp_cfg *cfg,
                             b200sp_cg_result *result, double *residuals_host) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A && x && b && result, "cg: null argument");
  const bool dist = halo != nullptr;
  const i64 n = A->num_rows;
  const i64 halo_lo = dist ? halo->halo_lo : 0, halo_hi = dist ? halo->halo_hi : 0;
  B200SP_REQUIRE(h, dist || A->num_rows == A->num_cols, "cg: matrix must be square");
  B200SP_REQUIRE(h, !dist || A->num_cols == n + halo_lo + halo_hi, "cg: num_cols != halo_lo+local+halo_hi");
  B200SP_REQUIRE(h, !dist || h->nccl_comm, "cg: distributed call without b200sp_comm_init");
  b200sp_cg_params prm = params ? *params : b200sp_cg_params{500, 1e-5, 0.0, 0};
  if (prm.check_interval <= 0) prm.check_interval = 16;
  B200SP_REQUIRE(h, prm.iteration_limit >= 0 && prm.iteration_limit < (1ll << 30), "cg: bad iteration limit");

  // the two-kernel iteration (direction update folded into the product): DIA through the bulk kernel, on one GPU or on
  // the peer-memory path with line-aligned halos
  b200sp_cfg cached_cfg;
  const b200sp_cfg *use_cfg = cfg;
  if (!use_cfg && tune_lookup_on(h, st, A, &cached_cfg)) use_cfg = &cached_cfg;
  const char *fuse_env = getenv("B200SP_CG_FUSE");
  const bool p2p = dist && h->p2p_ok && h->world > 1;
  const bool halo_lines = ((size_t)halo_lo * sizeof(T)) % 128 == 0 && ((size_t)(halo_lo + n) * sizeof(T)) % 128 == 0;
  const bool fuse_dir = fuse_env && fuse_env[0] == '1' && n > 0 && A->format == B200SP_FMT_DIA &&
                        dia_can_fuse_direction(A->num_rows, A->num_cols_per_row, A->pitch, A->values, sizeof(T), use_cfg) &&
                        (!dist || (p2p && halo_lines));
  // workspace: y (n) | r window | p window [| second p window], windows = halo_lo + n + halo_hi, each 128-byte aligned
  const size_t win = (size_t)(halo_lo + n + halo_hi);
  auto lines = [](size_t elems) { return (elems * sizeof(T) + 127) / 128 * 128; };
  const size_t off_r = lines((size_t)n), off_p0 = off_r + lines(win), off_p1 = off_p0 + lines(win);
  // one persistent cooperative kernel for small CSR systems on one GPU (cg_small_csr_kernel); B200SP_CG_PERSISTENT=0
  // turns it off.  The size test is the necessary condition (the mean slice fits); the kernel checks every slice.
  const char *pers_env = getenv("B200SP_CG_PERSISTENT");
  const i64 pers_grid = h->num_sms < 1024 ? h->num_sms : 1024;
  const size_t pers_smem = (size_t)h->max_smem_optin > 1024 ? (size_t)h->max_smem_optin - 1024 : 0;
  const bool persistent = !(pers_env && pers_env[0] == '0') && !dist && A->format == B200SP_FMT_CSR && n > 0 &&
                          A->num_entries < (1ll << 31) &&
                          ((size_t)A->num_entries * (sizeof(T) + 4) + (size_t)n * (4 * sizeof(T) + 4)) / (size_t)pers_grid +
                                  4096 <= pers_smem;
  const size_t need = off_p1 + ((fuse_dir || persistent) ? lines(win) : 0) + 256;
  if (h->cg_ws_bytes < need) {
    if (h->cg_ws) cudaFree(h->cg_ws);
    h->cg_ws = nullptr;
    h->cg_ws_bytes = 0;
    if (cudaMalloc(&h->cg_ws, need) != cudaSuccess) {
      cudaGetLastError();
      return set_error(h, B200SP_ALLOC_FAILED, "cg: cannot allocate %zu B workspace", need);
    }
    h->cg_ws_bytes = need;
  }
  const size_t nres_cap = (size_t)prm.iteration_limit + 2;
  if (h->cg_residuals_cap < nres_cap) {
    if (h->cg_residuals) cudaFree(h->cg_residuals);
    h->cg_residuals = nullptr;
    h->cg_residuals_cap = 0;
    if (cudaMalloc(&h->cg_residuals, nres_cap * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      return set_error(h, B200SP_ALLOC_FAILED, "cg: cannot allocate residual log");
    }
    h->cg_residuals_cap = nres_cap;
  }
  char *ws = reinterpret_cast<char *>(h->cg_ws);
  T *y = reinterpret_cast<T *>(ws);
  T *rwin = reinterpret_cast<T *>(ws + off_r);  // [halo_lo | n | halo_hi]; the halos are used by the two-kernel iteration only
  T *r = rwin + halo_lo;
  T *pwin = reinterpret_cast<T *>(ws + off_p0);  // [halo_lo | n | halo_hi]
  T *p = pwin + halo_lo;
  T *pwin_b = reinterpret_cast<T *>(ws + off_p1);  // second p window (two-kernel iteration)
  CgState<T> *S = reinterpret_cast<CgState<T> *>(h->dev_scalars);
  T *bn = reinterpret_cast<T *>(h->dev_scalars + 32);
  T *partials = reinterpret_cast<T *>(h->red_partials);
  unsigned int *ticket = h->red_counters;
  double *res = h->cg_residuals;
  const i64 g = cg_grid(h, n);
  b200sp_status s;

  // ||b||  (monitor constructor, monitor.inl:26-45)
  if (dist) {
    s = reduce<T, 0>(h, st, n, b, b, bn, nullptr);
    if (s != B200SP_OK) return s;
    s = comm_allreduce_sum(h, st, bn, 1, sizeof(T) == 8);
    if (s != B200SP_OK) return s;
    comm_sqrt_inplace(h, st, bn, sizeof(T) == 8);
  } else {
    s = reduce<T, 1>(h, st, n, b, nullptr, bn, nullptr);
    if (s != B200SP_OK) return s;
  }
  cg_setup_state_kernel<T><<<1, 1, 0, st>>>(S, bn, prm.relative_tolerance, prm.absolute_tolerance,
                                            (int)prm.iteration_limit);
  B200SP_LAUNCH_CHECK(h, "cg_setup_state_kernel");

  CgState<T> *hs = reinterpret_cast<CgState<T> *>(h->pinned_scalars);
  bool solved = false;
  if (persistent) {
    int *fallback = reinterpret_cast<int *>(h->red_counters + 12);
    B200SP_CUDA(h, cudaMemsetAsync(fallback, 0, sizeof(int), st));
    CgSmallArgs<T> ka;
    ka.n = n;
    ka.rows_per_cta = ceil_div(n, pers_grid);
    ka.Ap = A->row_offsets;
    ka.Aj = A->column_indices;
    ka.Ax = reinterpret_cast<const T *>(A->values);
    ka.x = x;
    ka.b = b;
    ka.r_g = r;
    ka.p_g0 = p;
    ka.p_g1 = pwin_b + halo_lo;
    ka.part_a = partials;
    ka.part_b = partials + 1024;
    ka.S = S;
    ka.residuals = res;
    ka.fallback = fallback;
    ka.smem_bytes = (unsigned)pers_smem;
    auto kern = cg_small_csr_kernel<T>;
    int resident = 0;
    bool ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pers_smem) == cudaSuccess &&
              cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, CGS_BLOCK, pers_smem) == cudaSuccess && resident >= 1;
    if (ok) {
      void *kargs[] = {&ka};
      ok = cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)pers_grid), dim3(CGS_BLOCK), kargs, pers_smem, st) ==
           cudaSuccess;
    }
    if (!ok) {
      cudaGetLastError();  // not launchable here (MPS / MIG limits, co-residency): the ordinary path below
    } else {
      h->launches++;
      int fb = 0;
      B200SP_CUDA(h, cudaMemcpyAsync(hs, S, sizeof(CgState<T>), cudaMemcpyDeviceToHost, st));
      B200SP_CUDA(h, cudaMemcpyAsync(&fb, fallback, sizeof(int), cudaMemcpyDeviceToHost, st));
      B200SP_CUDA(h, cudaStreamSynchronize(st));
      solved = fb == 0 && hs->done;
    }
  }

  // y = A x0 ; r = b - y ; p = r ; rz = <r,r>
  if (solved) {
    // nothing to set up: the persistent kernel has run the whole solve
  } else if (dist) {
    // x0 needs its own halo: reuse the p window as staging
    B200SP_CUDA(h, cudaMemcpyAsync(p, x, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    s = comm_halo_exchange(h, st, pwin, n, halo_lo, halo_hi, sizeof(T));
    if (s != B200SP_OK) return s;
    s = spmv_any<T>(h, st, A, pwin, y, 0, cfg, nullptr, nullptr);
  } else {
    s = spmv_any<T>(h, st, A, x, y, 0, cfg, nullptr, nullptr);
  }
  if (s != B200SP_OK) return s;
  if (solved) {
  } else if (dist) {
    cg_init_kernel<T, true><<<(unsigned)g, CG_BLOCK, 0, st>>>(n, b, y, r, p, S, partials, ticket, res);
    B200SP_LAUNCH_CHECK(h, "cg_init_kernel");
    s = comm_allreduce_sum(h, st, &S->rz, 1, sizeof(T) == 8);
    if (s != B200SP_OK) return s;
    cg_scalar_step_kernel<T><<<1, 1, 0, st>>>(S, res, 1);
    B200SP_LAUNCH_CHECK(h, "cg_scalar_step_kernel");
  } else {
    cg_init_kernel<T, false><<<(unsigned)g, CG_BLOCK, 0, st>>>(n, b, y, r, p, S, partials, ticket, res);
    B200SP_LAUNCH_CHECK(h, "cg_init_kernel");
  }

  auto poll = [&]() -> b200sp_status {
    B200SP_CUDA(h, cudaMemcpyAsync(hs, S, sizeof(CgState<T>), cudaMemcpyDeviceToHost, st));
    B200SP_CUDA(h, cudaStreamSynchronize(st));
    return B200SP_OK;
  };
  if (!solved) {
    s = poll();
    if (s != B200SP_OK) return s;
  }

  const i64 gdir = ceil_div(n, (i64)CG_BLOCK * CG_UNROLL);

  // NVLink peer-memory path: map the neighbours' p windows, agree on a solve id, and
  // exchange the halos of p_0 once with NCCL; from then on the iteration is NCCL-free.
  P2PView view;
  T *dst_lo = nullptr, *dst_hi = nullptr;
  unsigned long long solve = 0, kiter = 0;
  // DIA through the bulk kernel: K3 does not wait for the neighbours' planes; the next K1 does,
  // just before the tiles that read them (visited last)
  const bool defer = p2p && A->format == B200SP_FMT_DIA &&
                     dia_can_fuse_xchg(A->num_rows, A->num_cols_per_row, A->pitch, A->values, sizeof(T), use_cfg) &&
                     halo_lines && (reinterpret_cast<uintptr_t>(pwin) & 127) == 0;
  if (p2p) {
    void *dl = nullptr, *dh = nullptr;
    // the window the neighbours store into: p (three-kernel form) or r (two-kernel form)
    T *xwin = fuse_dir ? rwin : pwin;
    s = comm_p2p_map_windows(h, st, h->cg_ws, (size_t)((char *)xwin - (char *)h->cg_ws), n, halo_lo, halo_hi,
                             sizeof(T), &dl, &dh);
    if (s != B200SP_OK) return s;
    dst_lo = reinterpret_cast<T *>(dl);
    dst_hi = reinterpret_cast<T *>(dh);
    s = comm_next_solve_id(h, st, &solve);
    if (s != B200SP_OK) return s;
    view = comm_p2p_view(h);
    s = comm_halo_exchange(h, st, fuse_dir ? rwin : pwin, n, halo_lo, halo_hi, sizeof(T));  // halos of p_0 = r_0
    if (s != B200SP_OK) return s;
  }
  // Programmatic dependent launch between the three kernels of an iteration (opt-in: B200SP_CG_PDL=1):
  // every kernel requests what its predecessor does not write before griddepcontrol.wait, so launch latency and
  // first-batch load latency overlap the predecessor's tail.  The DIA bulk kernel joins in through h->pdl_spmv
  // (its producer lane streams matrix slabs into the ring while the direction kernel is still finishing).
  // Off by default — measured, it loses: 512^3 on one B200 3.14 ms / iteration with PDL against 2.95 without; two GPUs
  // with 64 planes each (the per-GPU size of the 8-GPU job) 0.428 against 0.420 ms.  The globaltimer trace
  // (B200SP_CG_TRACE, tools/cg_trace.py, profiles/r03_cg_trace.md) shows why: the K2 -> K3 gap does shrink (5.2 -> 2.3 us)
  // but the successor's early-resident CTAs slow the predecessor's tail by more than that (K3 body 72 -> 76 us, K1 + 3 us).
  const char *pdl_env = getenv("B200SP_CG_PDL");
  const bool pdl = pdl_env && pdl_env[0] != '0';
  struct PdlScope {  // the flag must not leak into products outside this solve
    b200sp_handle h;
    ~PdlScope() { h->pdl_spmv = false; }
  } pdl_scope{h};
  h->pdl_spmv = pdl;
  // B200SP_CG_TRACE=<path>: globaltimer stamps per iteration (poll waits, kernel ends) -> <path>.<rank>
  const char *trace_path = getenv("B200SP_CG_TRACE");
  unsigned long long *trace = nullptr;
  const size_t trace_words = trace_path ? ((size_t)prm.iteration_limit + 2) * TR_SLOTS : 0;
  if (trace_words) {
    B200SP_CUDA(h, cudaMalloc(&trace, trace_words * sizeof(unsigned long long)));
    B200SP_CUDA(h, cudaMemsetAsync(trace, 0, trace_words * sizeof(unsigned long long), st));
  }
  struct TraceScope {
    unsigned long long *&p;
    ~TraceScope() {
      if (p) cudaFree(p);
    }
  } trace_scope{trace};
  // Small systems are launch-bound (poisson5pt 512^2: ~5 us of kernels per iteration against ~25 us of launch calls):
  // one GPU, no per-iteration kernel arguments -> the `check_interval` iterations between two host polls are captured
  // once in a CUDA graph and replayed (B200SP_CG_GRAPH=0 / 1 overrides the size rule).  Capture happens on a stream
  // of the handle (the caller's may be the legacy default stream, which cannot be captured), ordered behind and
  // before the caller's stream with events.
  // two-kernel iteration: which p window holds the direction of the previous iteration
  int pcur = 0;       // K1' reads pwins[pcur] as p_old and stores p into pwins[pcur ^ 1]
  int first_dir = 1;  // first iteration: p = r
  T *pwins[2] = {pwin, pwin_b};
  auto fused_product = [&](cudaStream_t fst, const FusedXchg *xw) -> b200sp_status {
    b200sp_status fs = spmv_dia_fused_direction<T>(h, fst, A->num_rows, A->num_cols, A->num_cols_per_row, A->pitch,
                                                   A->diagonal_offsets, reinterpret_cast<const T *>(A->values), y, use_cfg, rwin,
                                                   pwins[pcur], pwins[pcur ^ 1], &S->beta, &S->done, first_dir, halo_lo, halo_lo,
                                                   halo_hi, &S->yp, xw);
    pcur ^= 1;  // pwins[pcur] now holds the current direction
    first_dir = 0;
    return fs;
  };
  const char *graph_env = getenv("B200SP_CG_GRAPH");
  bool use_graph = !solved && !dist && (graph_env ? graph_env[0] != '0' : n <= ((i64)1 << 22));
  int graph_iters = prm.check_interval;
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t run_st = st;
  if (use_graph) {
    if (!h->graph_stream) {
      cudaStream_t gs;
      if (cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking) == cudaSuccess) h->graph_stream = gs;
    }
    if (!h->graph_event) {
      cudaEvent_t ev;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) h->graph_event = ev;
    }
    use_graph = h->graph_stream && h->graph_event;
  }
  if (use_graph) {
    cudaStream_t gs = (cudaStream_t)h->graph_stream;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      const uint64_t launches_before = h->launches;
      if (fuse_dir) {
        // the first iteration (p = r) runs outside the graph; the graph holds an even number of iterations so that
        // every replay starts with the two p windows in the same roles
        const int pairs = (prm.check_interval + 1) / 2;
        graph_iters = 2 * pairs;
        first_dir = 0;
        pcur = 1;  // after the uncaptured first iteration the direction lives in pwins[1]
        for (int k = 0; k < graph_iters && ok; ++k) {
          ok = fused_product(gs, nullptr) == B200SP_OK;
          ok = ok && launch_kernel_pdl(cg_update_kernel<T, false>, dim3((unsigned)g), dim3(CG_BLOCK), 0, gs, false, n,
                                       (const T *)(pwins[pcur] + halo_lo), (const T *)y, x, r, S, partials, ticket,
                                       res) == cudaSuccess;
        }
        first_dir = 1;
        pcur = 0;
      }
      for (int k = 0; !fuse_dir && k < prm.check_interval && ok; ++k) {
        ok = spmv_any<T>(h, gs, A, p, y, 0, cfg, p, &S->yp) == B200SP_OK;
        ok = ok && launch_kernel_pdl(cg_update_kernel<T, false>, dim3((unsigned)g), dim3(CG_BLOCK), 0, gs, pdl, n, (const T *)p,
                                     (const T *)y, x, r, S, partials, ticket, res) == cudaSuccess;
        ok = ok && launch_kernel_pdl(cg_direction_kernel<T>, dim3((unsigned)gdir), dim3(CG_BLOCK), 0, gs, pdl, n, (const T *)r, p,
                                     (const CgState<T> *)S) == cudaSuccess;
      }
      h->launches = launches_before;  // counted per replay below
      ok = (cudaStreamEndCapture(gs, &graph) == cudaSuccess) && ok && graph;
    }
    if (ok) ok = cudaGraphInstantiate(&graph_exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      graph_exec = nullptr;
      use_graph = false;
    } else {
      // everything queued so far on the caller's stream precedes the first replay
      B200SP_CUDA(h, cudaEventRecord((cudaEvent_t)h->graph_event, st));
      B200SP_CUDA(h, cudaStreamWaitEvent(gs, (cudaEvent_t)h->graph_event, 0));
      run_st = gs;
    }
  }
  struct GraphScope {
    cudaGraphExec_t &e;
    ~GraphScope() {
      if (e) cudaGraphExecDestroy(e);
    }
  } graph_scope{graph_exec};
  auto poll_on = [&](cudaStream_t ps) -> b200sp_status {
    B200SP_CUDA(h, cudaMemcpyAsync(hs, S, sizeof(CgState<T>), cudaMemcpyDeviceToHost, ps));
    B200SP_CUDA(h, cudaStreamSynchronize(ps));
    return B200SP_OK;
  };
  if (use_graph && fuse_dir && !hs->done) {  // the uncaptured first iteration of the two-kernel form
    s = fused_product(run_st, nullptr);
    if (s != B200SP_OK) return s;
    B200SP_CUDA(h, launch_kernel_pdl(cg_update_kernel<T, false>, dim3((unsigned)g), dim3(CG_BLOCK), 0, run_st, false, n,
                                     (const T *)(pwins[pcur] + halo_lo), (const T *)y, x, r, S, partials, ticket, res));
    h->launches++;
  }
  while (use_graph && !hs->done) {
    B200SP_CUDA(h, cudaGraphLaunch(graph_exec, run_st));
    h->launches += (fuse_dir ? 2 : 3) * (uint64_t)graph_iters;
    s = poll_on(run_st);
    if (s != B200SP_OK) return s;
  }
  while (!hs->done) {
    for (int k = 0; k < prm.check_interval; ++k) {
      if (p2p && fuse_dir) {
        const unsigned long long prev_epoch = (solve << 32) | kiter;
        const unsigned long long epoch = (solve << 32) | (++kiter);
        const unsigned tag = (unsigned)((solve << 20) + kiter);
        FusedXchg xw;
        memset(&xw, 0, sizeof(xw));
        if (kiter > 1) {  // K1' waits for the planes of r that the neighbours' previous update kernel stored here
          xw.enabled = 2;
          xw.mine = view.mine;
          xw.lo_bytes = dst_lo ? (size_t)halo_lo * sizeof(T) : 0;
          xw.hi_bytes = dst_hi ? (size_t)halo_hi * sizeof(T) : 0;
          xw.wait_lo = &view.mine->halo_flag[0];
          xw.wait_hi = &view.mine->halo_flag[1];
          xw.wait_epoch = prev_epoch;
        }
        s = fused_product(st, kiter > 1 ? &xw : nullptr);
        if (s != B200SP_OK) return s;
        unsigned long long *tr = (trace && kiter < (unsigned long long)prm.iteration_limit + 2) ? trace + kiter * TR_SLOTS : nullptr;
        B200SP_CUDA(h, launch_kernel_pdl(cg_update_push_p2p_kernel<T>, dim3((unsigned)g), dim3(CG_BLOCK), 0, st, false, n,
                                         (const T *)(pwins[pcur] + halo_lo), (const T *)y, x, r, S, partials, ticket, res, view,
                                         tag, epoch, halo_lo, halo_hi, dst_lo, dst_hi, tr));
        h->launches++;
        continue;
      }
      if (p2p) {
        const unsigned long long prev_epoch = (solve << 32) | kiter;
        const unsigned long long epoch = (solve << 32) | (++kiter);
        const unsigned tag = (unsigned)((solve << 20) + kiter);  // differs from whatever the slot held before
        if (defer && kiter > 1) {
          // K1 waits for the planes of p that the neighbours' previous K3 stored here
          FusedXchg xw;
          memset(&xw, 0, sizeof(xw));
          xw.enabled = 2;
          xw.mine = view.mine;
          xw.lo_bytes = dst_lo ? (size_t)halo_lo * sizeof(T) : 0;
          xw.hi_bytes = dst_hi ? (size_t)halo_hi * sizeof(T) : 0;
          xw.wait_lo = &view.mine->halo_flag[0];
          xw.wait_hi = &view.mine->halo_flag[1];
          xw.wait_epoch = prev_epoch;
          int fused = 0;
          s = spmv_dia_xchg<T>(h, st, A->num_rows, A->num_cols, A->num_cols_per_row, A->pitch, A->diagonal_offsets,
                               reinterpret_cast<const T *>(A->values), pwin, y, 0, use_cfg, p, &S->yp, &xw, &fused);
          if (s == B200SP_OK && !fused) s = set_error(h, B200SP_COMM_ERROR, "cg_dist: deferred halo wait was not launched");
        } else {
          s = spmv_any<T>(h, st, A, pwin, y, 0, cfg, p, &S->yp);
        }
        if (s != B200SP_OK) return s;
        unsigned long long *tr = (trace && kiter < (unsigned long long)prm.iteration_limit + 2) ? trace + kiter * TR_SLOTS : nullptr;
        B200SP_CUDA(h, launch_kernel_pdl(cg_update_p2p_kernel<T>, dim3((unsigned)g), dim3(CG_BLOCK), 0, st, pdl, n, p, y, x,
                                         r, S, partials, ticket, view, tag, tr));
        h->launches++;
        B200SP_CUDA(h, launch_kernel_pdl(cg_direction_p2p_kernel<T>, dim3((unsigned)g), dim3(CG_BLOCK), 0, st, pdl, n, r, p,
                                         S, ticket + 1, res, view, tag, epoch, halo_lo, halo_hi, dst_lo, dst_hi,
                                         defer ? 1 : 0, tr));
        h->launches++;
        continue;
      }
      if (dist) {
        s = comm_halo_exchange(h, st, pwin, n, halo_lo, halo_hi, sizeof(T));
        if (s != B200SP_OK) return s;
        s = spmv_any<T>(h, st, A, pwin, y, 0, cfg, p, &S->yp);
        if (s != B200SP_OK) return s;
        s = comm_allreduce_sum(h, st, &S->yp, 1, sizeof(T) == 8);
        if (s != B200SP_OK) return s;
        cg_update_kernel<T, true><<<(unsigned)g, CG_BLOCK, 0, st>>>(n, p, y, x, r, S, partials, ticket, res);
        B200SP_LAUNCH_CHECK(h, "cg_update_kernel");
        s = comm_allreduce_sum(h, st, &S->rz_new, 1, sizeof(T) == 8);
        if (s != B200SP_OK) return s;
        cg_scalar_step_kernel<T><<<1, 1, 0, st>>>(S, res, 0);
        B200SP_LAUNCH_CHECK(h, "cg_scalar_step_kernel");
      } else if (fuse_dir) {
        s = fused_product(st, nullptr);
        if (s != B200SP_OK) return s;
        B200SP_CUDA(h, launch_kernel_pdl(cg_update_kernel<T, false>, dim3((unsigned)g), dim3(CG_BLOCK), 0, st, false, n,
                                         (const T *)(pwins[pcur] + halo_lo), (const T *)y, x, r, S, partials, ticket, res));
        h->launches++;
        continue;
      } else {
        s = spmv_any<T>(h, st, A, p, y, 0, cfg, p, &S->yp);
        if (s != B200SP_OK) return s;
        B200SP_CUDA(h, launch_kernel_pdl(cg_update_kernel<T, false>, dim3((unsigned)g), dim3(CG_BLOCK), 0, st, pdl, n,
                                         (const T *)p, (const T *)y, x, r, S, partials, ticket, res));
        h->launches++;
      }
      B200SP_CUDA(h, launch_kernel_pdl(cg_direction_kernel<T>, dim3((unsigned)gdir), dim3(CG_BLOCK), 0, st, pdl && !dist, n,
                                       (const T *)r, p, (const CgState<T> *)S));
      h->launches++;
    }
    s = poll();
    if (s != B200SP_OK) return s;
  }

  h->pdl_spmv = false;
  if (trace && trace_path) {
    std::vector<unsigned long long> tr(trace_words);
    B200SP_CUDA(h, cudaMemcpyAsync(tr.data(), trace, trace_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    B200SP_CUDA(h, cudaStreamSynchronize(st));
    char path[1024];
    snprintf(path, sizeof(path), "%s.%d", trace_path, h->rank);
    if (FILE *f = fopen(path, "w")) {
      fprintf(f, "# iter k2_enter k2_polled k2_end k3_enter k3_polled k3_end (globaltimer ns, this GPU)\n");
      for (unsigned long long it = 1; it <= kiter && it < (unsigned long long)prm.iteration_limit + 2; ++it) {
        fprintf(f, "%llu", it);
        for (int q = 0; q < 6; ++q) fprintf(f, " %llu", tr[it * TR_SLOTS + q]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  if (p2p) {
    unsigned long long timeouts = 0;
    B200SP_CUDA(h, cudaMemcpyAsync(&timeouts, &view.mine->timeouts, sizeof(timeouts), cudaMemcpyDeviceToHost, st));
    B200SP_CUDA(h, cudaStreamSynchronize(st));
    if (timeouts)
      return set_error(h, B200SP_COMM_ERROR, "cg_dist: a peer never published its data (%llu spin time-outs)", timeouts);
  }
  result->iteration_count = hs->iter;
  result->converged = hs->converged;
  result->residual_norm = (double)hs->rnorm;
  result->b_norm = (double)hs->bnorm;
  result->num_residuals = hs->nres;
  if (residuals_host && hs->nres > 0) {
    B200SP_CUDA(h, cudaMemcpyAsync(residuals_host, res, (size_t)hs->nres * sizeof(double),
                                   cudaMemcpyDeviceToHost, st));
    B200SP_CUDA(h, cudaStreamSynchronize(st));
  }
  return B200SP_OK;
}

}  // namespace b200sp

extern "C" {

b200sp_status b200sp_spmv(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A, const void *x,
                          void *y, int accumulate, const b200sp_cfg *cfg) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A != nullptr, "spmv: null matrix descriptor");
  if (A->dtype == B200SP_F32)
    return b200sp::spmv_any<float>(h, (cudaStream_t)stream, A, (const float *)x, (float *)y, accumulate,
                                   cfg, nullptr, nullptr);
  if (A->dtype == B200SP_F64)
    return b200sp::spmv_any<double>(h, (cudaStream_t)stream, A, (const double *)x, (double *)y, accumulate,
                                    cfg, nullptr, nullptr);
  return b200sp::set_error(h, B200SP_INVALID_INPUT, "spmv: unknown dtype %d", (int)A->dtype);
}

b200sp_status b200sp_cg(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A, void *x,
                        const void *b, const b200sp_cg_params *params, const b200sp_cfg *spmv_cfg,
                        b200sp_cg_result *result, double *residuals_host) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A != nullptr, "cg: null matrix descriptor");
  if (A->dtype == B200SP_F32)
    return b200sp::cg_impl<float>(h, (cudaStream_t)stream, A, nullptr, (float *)x, (const float *)b, params,
                                  spmv_cfg, result, residuals_host);
  if (A->dtype == B200SP_F64)
    return b200sp::cg_impl<double>(h, (cudaStream_t)stream, A, nullptr, (double *)x, (const double *)b,
                                   params, spmv_cfg, result, residuals_host);
  return b200sp::set_error(h, B200SP_INVALID_INPUT, "cg: unknown dtype %d", (int)A->dtype);
}

b200sp_status b200sp_cg_dist(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A_local,
                             const b200sp_halo *halo, void *x_local, const void *b_local,
                             const b200sp_cg_params *params, const b200sp_cfg *spmv_cfg,
                             b200sp_cg_result *result, double *residuals_host) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A_local != nullptr && halo != nullptr, "cg_dist: null argument");
  if (A_local->dtype == B200SP_F32)
    return b200sp::cg_impl<float>(h, (cudaStream_t)stream, A_local, halo, (float *)x_local,
                                  (const float *)b_local, params, spmv_cfg, result, residuals_host);
  if (A_local->dtype == B200SP_F64)
    return b200sp::cg_impl<double>(h, (cudaStream_t)stream, A_local, halo, (double *)x_local,
                                   (const double *)b_local, params, spmv_cfg, result, residuals_host);
  return b200sp::set_error(h, B200SP_INVALID_INPUT, "cg_dist: unknown dtype %d", (int)A_local->dtype);
}

b200sp_status b200sp_spmv_dist(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A_local,
                               const b200sp_halo *halo, void *x_window, void *y_local,
                               const b200sp_cfg *cfg) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A_local && halo && x_window && y_local, "spmv_dist: null argument");
  B200SP_REQUIRE(h, h->nccl_comm, "spmv_dist: b200sp_comm_init has not been called");
  const size_t elem = A_local->dtype == B200SP_F64 ? 8 : 4;
  // DIA through the bulk kernel: the exchange rides inside the product (peer memory,
  // idle producer lanes) and is hidden behind the interior tiles
  if (A_local->format == B200SP_FMT_DIA && h->world > 1 && h->p2p_ok) {
    b200sp_cfg cached;
    const b200sp_cfg *use = cfg;
    if (!use && b200sp::tune_lookup_on(h, (cudaStream_t)stream, A_local, &cached)) use = &cached;
    b200sp::FusedXchg xc;
    if (b200sp::dia_can_fuse_xchg(A_local->num_rows, A_local->num_cols_per_row, A_local->pitch, A_local->values, elem,
                                  use) &&
        b200sp::comm_fused_xchg_prepare(h, x_window, A_local->num_rows, halo->halo_lo, halo->halo_hi, elem, &xc)) {
      int fused = 0;
      b200sp_status fs;
      if (elem == 8)
        fs = b200sp::spmv_dia_xchg<double>(h, (cudaStream_t)stream, A_local->num_rows, A_local->num_cols,
                                           A_local->num_cols_per_row, A_local->pitch, A_local->diagonal_offsets,
                                           (const double *)A_local->values, (const double *)x_window,
                                           (double *)y_local, 0, use, nullptr, nullptr, &xc, &fused);
      else
        fs = b200sp::spmv_dia_xchg<float>(h, (cudaStream_t)stream, A_local->num_rows, A_local->num_cols,
                                          A_local->num_cols_per_row, A_local->pitch, A_local->diagonal_offsets,
                                          (const float *)A_local->values, (const float *)x_window,
                                          (float *)y_local, 0, use, nullptr, nullptr, &xc, &fused);
      if (fs != B200SP_OK) return fs;
      if (!fused) return b200sp::set_error(h, B200SP_COMM_ERROR, "spmv_dist: fused halo exchange was not launched");
      return B200SP_OK;
    }
  }
  b200sp_status s = b200sp::comm_halo_exchange_auto(h, (cudaStream_t)stream, x_window, A_local->num_rows,
                                                    halo->halo_lo, halo->halo_hi, elem);
  if (s != B200SP_OK) return s;
  return b200sp_spmv(h, stream, A_local, x_window, y_local, 0, cfg);
}

b200sp_status b200sp_spmv_dist_gather(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A_local,
                                      const int64_t *slice_offsets, void *x_full, void *y_local,
                                      const b200sp_cfg *cfg) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A_local && slice_offsets && x_full && y_local, "spmv_dist_gather: null argument");
  B200SP_REQUIRE(h, h->nccl_comm || h->world == 1, "spmv_dist_gather: b200sp_comm_init has not been called");
  B200SP_REQUIRE(h, slice_offsets[0] == 0 && slice_offsets[h->world] == A_local->num_cols,
                 "spmv_dist_gather: slice offsets must cover [0, num_cols)");
  const size_t elem = A_local->dtype == B200SP_F64 ? 8 : 4;
  b200sp_status s = b200sp::comm_allgather_slices(h, (cudaStream_t)stream, x_full, slice_offsets, elem);
  if (s != B200SP_OK) return s;
  return b200sp_spmv(h, stream, A_local, x_full, y_local, 0, cfg);
}
}
