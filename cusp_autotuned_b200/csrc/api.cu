// api.cu — handle lifecycle, error reporting, host-buffer SpMV and the tuner.
//
// The tuner replaces the KTT glue of the reference: cusp::ktt::{multiply,tune,
// reset_tuning} (cusp/ktt/detail/ktt.inl:83-142), cusp::system::cuda::ktt::
// {multiply,tune} (cusp/system/cuda/ktt/multiply.h:56-153) and the per-format
// parameter spaces (cusp/system/cuda/ktt/{csr,ell,dia,coo}_multiply.h).  KTT
// JIT-compiles every configuration with NVRTC; here every point of the space is
// a precompiled template instantiation, so a tuning step costs one launch.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "comm.h"

static char g_global_err[1024] = "no error";

namespace b200sp {

b200sp_status set_error(b200sp_handle h, b200sp_status s, const char *fmt, ...) {
  char *dst = h ? h->err : g_global_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 1024, fmt, ap);
  va_end(ap);
  return s;
}

b200sp_status ensure_scratch(b200sp_handle h, size_t bytes) {
  if (h->scratch_bytes >= bytes) return B200SP_OK;
  if (h->scratch) cudaFree(h->scratch);  // implicit device sync: in-flight users are done
  h->scratch = nullptr;
  h->scratch_bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  if (cudaMalloc(&h->scratch, want) != cudaSuccess) {
    cudaGetLastError();
    return set_error(h, B200SP_ALLOC_FAILED, "cannot allocate %zu B of scratch", want);
  }
  h->scratch_bytes = want;
  return B200SP_OK;
}

template <typename T>
b200sp_status spmv_any(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, const T *x, T *y,
                       int accumulate, const b200sp_cfg *cfg, const T *dotv, T *dot_result);

static i64 stored_entries(const b200sp_matrix *A) {
  switch (A->format) {
    case B200SP_FMT_ELL:
    case B200SP_FMT_ELLR:
    case B200SP_FMT_DIA:
      return A->num_entries > 0 ? A->num_entries : A->num_cols_per_row * A->num_rows;
    case B200SP_FMT_HYB:
      return A->num_cols_per_row * A->num_rows + A->coo_num_entries;
    default:
      return A->num_entries;
  }
}

static int ilog2(i64 v) {
  int l = 0;
  while (v > 1) {
    v >>= 1;
    ++l;
  }
  return l;
}

int csr_structure_class(b200sp_handle h, cudaStream_t st, i64 rows, i64 nnz, const int *Ap, const int *Aj);  // spmv_csr.cu
int coo_structure_class(b200sp_handle h, cudaStream_t st, i64 nnz, const int *Aj, size_t elem);                 // spmv_coo.cu

static b200sp_tune_key make_key(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A) {
  b200sp_tune_key k;
  k.format = (int)A->format;
  k.dtype = (int)A->dtype;
  k.rows_log2 = ilog2(A->num_rows > 0 ? A->num_rows : 1);
  const i64 rows = A->num_rows > 0 ? A->num_rows : 1;
  k.nnz_per_row_log2 = ilog2((stored_entries(A) + rows - 1) / rows);
  const size_t elem = A->dtype == B200SP_F64 ? 8 : 4;
  k.structure = 0;  // the probes are cached per array address: one analysis kernel per matrix, then a map lookup
  if (A->format == B200SP_FMT_CSR)
    k.structure = csr_structure_class(h, st, A->num_rows, A->num_entries, A->row_offsets, A->column_indices);
  else if (A->format == B200SP_FMT_COO)
    k.structure = coo_structure_class(h, st, A->num_entries, A->column_indices, elem);
  else if (A->format == B200SP_FMT_HYB)
    k.structure = coo_structure_class(h, st, A->coo_num_entries, A->coo_column_indices, elem);
  return k;
}

int tune_lookup_on(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, b200sp_cfg *cfg) {
  if (!h || !A || h->tune_cache.empty()) return 0;
  auto it = h->tune_cache.find(make_key(h, st, A));
  if (it == h->tune_cache.end() || !it->second.has_best) return 0;
  // during dynamic tuning the cached "best so far" is only used once the space is exhausted
  if (it->second.next_index < b200sp_cfg_space(A->format, A->dtype, nullptr, 0)) return 0;
  if (cfg) *cfg = it->second.best;
  return 1;
}

// ---- configuration spaces -----------------------------------------------------
static void push(std::vector<b200sp_cfg> &v, int kernel, int block, int tpr, int unroll, int stages, int cps) {
  b200sp_cfg c{};
  c.kernel = kernel;
  c.block_size = block;
  c.threads_per_row = tpr;
  c.unroll = unroll;
  c.vector_width = 1;
  c.tile_rows = (kernel == B200SP_K_ELL_BULK) ? block * unroll : 0;
  c.stages = stages;
  c.ctas_per_sm = cps;
  v.push_back(c);
}

static std::vector<b200sp_cfg> cfg_space_vec(b200sp_format f, b200sp_dtype dt) {
  std::vector<b200sp_cfg> v;
  const int blocks[3] = {128, 256, 512};
  switch (f) {
    case B200SP_FMT_CSR:
      for (int b : blocks)
        for (int tpr : {1, 2, 4, 8, 16, 32})
          for (int u : {1, 2, 4}) push(v, B200SP_K_CSR_VECTOR, b, tpr, u, 0, 0);
      for (int b : blocks)
        for (int u : {4, 8, 16}) {
          if (b == 512 && u == 16) continue;
          push(v, B200SP_K_CSR_STREAM, b, 0, u, 0, 0);
        }
      {
        const int bu[5][2] = {{128, 4}, {128, 8}, {128, 16}, {256, 4}, {256, 8}};
        for (auto &p : bu)
          for (int st : {2, 3, 4})
            for (int cps : {2, 4, 6}) push(v, B200SP_K_CSR_RING, p[0], 0, p[1], st, cps);
        push(v, B200SP_K_CSR_BALANCED, 128, 0, 7, 0, 0);
        for (int u : {5, 7, 9}) push(v, B200SP_K_CSR_BALANCED, 256, 0, u, 0, 0);
      }
      break;
    case B200SP_FMT_ELL:
    case B200SP_FMT_ELLR:
    case B200SP_FMT_HYB:
    case B200SP_FMT_DIA: {
      const int k_ldg = (f == B200SP_FMT_DIA) ? B200SP_K_DIA_LDG : B200SP_K_ELL_LDG;
      const int k_bulk = (f == B200SP_FMT_DIA) ? B200SP_K_DIA_BULK : B200SP_K_ELL_BULK;
      for (int b : blocks)
        for (int u : {1, 2, 4}) push(v, k_ldg, b, 0, u, 0, 0);
      const int bu[6][2] = {{128, 2}, {128, 4}, {128, 8}, {256, 1}, {256, 2}, {256, 4}};
      for (auto &p : bu)
        for (int st : {2, 3, 4})
          for (int cps : {1, 2, 4}) push(v, k_bulk, p[0], 0, p[1], st, cps);
      break;
    }
    case B200SP_FMT_COO:
      for (int b : blocks)
        for (int u : {5, 7, 9, 11}) {
          if (b == 512 && u > 7) continue;
          push(v, B200SP_K_COO_SEGSCAN, b, 0, u, 0, 0);
        }
      {
        const int bu[8][2] = {{128, 7}, {128, 9}, {128, 11}, {256, 5}, {256, 7}, {256, 9}, {512, 5}, {512, 7}};
        for (auto &p : bu)
          for (int st : {2, 3})
            for (int cps : {2, 4, 6}) push(v, B200SP_K_COO_RING, p[0], 0, p[1], st, cps);
      }
      for (int vw : {4, 8})  // K_COO_WARP: entries per lane per load x units per warp tile
        for (int u : {1, 2, 4}) {
          if (dt == B200SP_F64 && vw == 8 && u == 4) continue;
          push(v, B200SP_K_COO_WARP, 256, 0, u, 0, 0);
          v.back().vector_width = vw;
        }
      break;
  }
  return v;
}

// ---- validation ---------------------------------------------------------------
template <typename T>
__global__ void absmax_kernel(i64 n, const T *a, unsigned int *out) {
  float m = 0.f;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf((float)a[i]));
  atomicMax(out, __float_as_uint(m));
}

template <typename T>
__global__ void relerr_kernel(i64 n, const T *y, const T *ref, const unsigned int *scale_bits,
                              unsigned int *out) {
  const double floor_ = 1e-6 * (double)__uint_as_float(*scale_bits) + 1e-300;
  float m = 0.f;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const double d = fabs((double)y[i] - (double)ref[i]);
    const double den = fmax(fabs((double)ref[i]), floor_);
    float e = (float)(d / den);
    if (!(d == d)) e = 3.0e38f;  // NaN
    m = fmaxf(m, e);
  }
  atomicMax(out, __float_as_uint(m));
}

template <typename T>
static b200sp_status max_rel_error(b200sp_handle h, cudaStream_t st, i64 n, const T *y, const T *ref,
                                   double *out) {
  unsigned int *bits = h->red_counters + 8;  // [8]=scale, [9]=err
  B200SP_CUDA(h, cudaMemsetAsync(bits, 0, 2 * sizeof(unsigned int), st));
  if (n > 0) {
    const unsigned g = (unsigned)(ceil_div(n, 256) < 1184 ? ceil_div(n, 256) : 1184);
    absmax_kernel<T><<<g, 256, 0, st>>>(n, ref, bits);
    relerr_kernel<T><<<g, 256, 0, st>>>(n, y, ref, bits, bits + 1);
    h->launches += 2;
  }
  unsigned int hb[2];
  B200SP_CUDA(h, cudaMemcpyAsync(hb, bits, sizeof(hb), cudaMemcpyDeviceToHost, st));
  B200SP_CUDA(h, cudaStreamSynchronize(st));
  float e;
  memcpy(&e, &hb[1], 4);
  *out = (double)e;
  return B200SP_OK;
}

static b200sp_status tune_events(b200sp_handle h) {
  if (h->tune_events.size() == 2) return B200SP_OK;
  cudaEvent_t a, b;
  B200SP_CUDA(h, cudaEventCreate(&a));
  B200SP_CUDA(h, cudaEventCreate(&b));
  h->tune_events.push_back(a);
  h->tune_events.push_back(b);
  return B200SP_OK;
}

template <typename T>
static b200sp_status tune_impl(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, const T *x, T *y,
                               const T *yref_in, double tol, int repeats, const int64_t *order, int64_t order_len,
                               b200sp_tune_callback callback, void *user, b200sp_tune_result *results,
                               int64_t capacity, int64_t *num_results, b200sp_cfg *best) {
  if (repeats <= 0) repeats = 5;
  if (tol <= 0) tol = sizeof(T) == 4 ? 1e-5 : 1e-12;
  b200sp_status s = tune_events(h);
  if (s != B200SP_OK) return s;
  cudaEvent_t e0 = (cudaEvent_t)h->tune_events[0], e1 = (cudaEvent_t)h->tune_events[1];
  const i64 n = A->num_rows;
  std::vector<b200sp_cfg> space = cfg_space_vec(A->format, A->dtype);

  // reference output: caller's, or the engine-default configuration's
  T *yref = nullptr;
  const T *ref = yref_in;
  if (!ref) {
    B200SP_CUDA(h, cudaMalloc(&yref, (size_t)(n > 0 ? n : 1) * sizeof(T)));
    b200sp_cfg dflt{};
    s = spmv_any<T>(h, st, A, x, yref, 0, &dflt, nullptr, nullptr);
    if (s != B200SP_OK) {
      cudaFree(yref);
      return s;
    }
    ref = yref;
  }

  float best_ms = 1e30f;
  b200sp_cfg best_cfg{};
  bool have = false;
  i64 count = 0;
  // the searcher's order (cuda/ktt/multiply.h:129-133 SetSearcher): a permutation / subset of the space, or the
  // space's own order; the stop condition is consulted after every configuration (Tune(kernel, stop_condition))
  const i64 visits = order ? order_len : (i64)space.size();
  bool stopped = false;
  for (i64 vi = 0; vi < visits && !stopped; ++vi) {
    const i64 ci = order ? order[vi] : vi;
    if (ci < 0 || ci >= (i64)space.size()) continue;
    const b200sp_cfg &c = space[(size_t)ci];
    b200sp_tune_result r{};
    r.cfg = c;
    char saved[1024];
    memcpy(saved, h->err, sizeof(saved));
    s = spmv_any<T>(h, st, A, x, y, 0, &c, nullptr, nullptr);
    cudaError_t ce = (s == B200SP_OK) ? cudaStreamSynchronize(st) : cudaSuccess;
    if (s == B200SP_INVALID_INPUT || s == B200SP_NOT_IMPLEMENTED) {
      r.status = B200SP_TUNE_UNSUPPORTED;
      memcpy(h->err, saved, sizeof(saved));
    } else if (s != B200SP_OK || ce != cudaSuccess) {
      r.status = B200SP_TUNE_LAUNCH_FAILED;
      cudaGetLastError();
    } else {
      s = max_rel_error<T>(h, st, n, y, ref, &r.max_rel_error);
      if (s != B200SP_OK) {
        if (yref) cudaFree(yref);
        return s;
      }
      if (r.max_rel_error > tol) {
        r.status = B200SP_TUNE_VALIDATION_FAILED;
      } else {
        cudaEventRecord(e0, st);
        for (int k = 0; k < repeats; ++k) spmv_any<T>(h, st, A, x, y, 0, &c, nullptr, nullptr);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        r.milliseconds = ms / repeats;
        r.status = B200SP_TUNE_OK;
        if (r.milliseconds < best_ms) {
          best_ms = r.milliseconds;
          best_cfg = c;
          have = true;
        }
      }
    }
    if (results && count < capacity) results[count] = r;
    ++count;
    if (callback && callback(&r, user)) stopped = true;
  }
  if (num_results) *num_results = count;
  if (yref) cudaFree(yref);
  if (!have) return set_error(h, B200SP_CUDA_ERROR, "tune: no configuration produced a valid result");
  b200sp_tune_entry &e = h->tune_cache[make_key(h, st, A)];
  e.best = best_cfg;
  e.has_best = true;
  e.best_ms = best_ms;
  e.next_index = (i64)space.size();  // dynamic tuning (b200sp_tune_step) starts from the winner of this search
  if (best) *best = best_cfg;
  // leave y = A x computed by the winner
  return spmv_any<T>(h, st, A, x, y, 0, &best_cfg, nullptr, nullptr);
}

template <typename T>
static b200sp_status tune_step_impl(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, const T *x,
                                    T *y, b200sp_tune_result *result) {
  std::vector<b200sp_cfg> space = cfg_space_vec(A->format, A->dtype);
  b200sp_tune_entry &e = h->tune_cache[make_key(h, st, A)];
  b200sp_status s;
  while (e.next_index < (i64)space.size()) {
    const b200sp_cfg c = space[(size_t)e.next_index++];
    s = tune_events(h);
    if (s != B200SP_OK) return s;
    cudaEvent_t e0 = (cudaEvent_t)h->tune_events[0], e1 = (cudaEvent_t)h->tune_events[1];
    char saved[1024];
    memcpy(saved, h->err, sizeof(saved));
    cudaEventRecord(e0, st);
    s = spmv_any<T>(h, st, A, x, y, 0, &c, nullptr, nullptr);
    if (s == B200SP_INVALID_INPUT || s == B200SP_NOT_IMPLEMENTED) {  // not runnable here: next point
      memcpy(h->err, saved, sizeof(saved));
      continue;
    }
    if (s != B200SP_OK) return s;
    cudaEventRecord(e1, st);
    B200SP_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (!e.has_best || ms < e.best_ms) {
      e.best = c;
      e.best_ms = ms;
      e.has_best = true;
    }
    if (result) {
      result->cfg = c;
      result->status = B200SP_TUNE_OK;
      result->milliseconds = ms;
      result->max_rel_error = 0.0;
    }
    return B200SP_OK;
  }
  // space exhausted: steady state, no synchronisation
  b200sp_cfg c = e.has_best ? e.best : b200sp_cfg{};
  s = spmv_any<T>(h, st, A, x, y, 0, &c, nullptr, nullptr);
  if (result) {
    result->cfg = c;
    result->status = s == B200SP_OK ? B200SP_TUNE_OK : B200SP_TUNE_LAUNCH_FAILED;
    result->milliseconds = e.best_ms;
    result->max_rel_error = 0.0;
  }
  return s;
}

}  // namespace b200sp

extern "C" {

int b200sp_version(void) { return B200SP_VERSION; }

const char *b200sp_status_string(b200sp_status s) {
  switch (s) {
    case B200SP_OK: return "ok";
    case B200SP_INVALID_INPUT: return "invalid input";
    case B200SP_CUDA_ERROR: return "CUDA error";
    case B200SP_NOT_IMPLEMENTED: return "not implemented";
    case B200SP_ALLOC_FAILED: return "allocation failed";
    case B200SP_COMM_ERROR: return "communication error";
  }
  return "unknown status";
}

const char *b200sp_last_error_string(b200sp_handle h) { return h ? h->err : g_global_err; }

uint64_t b200sp_launch_count(b200sp_handle h) { return h ? h->launches : 0; }

b200sp_status b200sp_create(b200sp_handle *out) {
  if (!out) return b200sp::set_error(nullptr, B200SP_INVALID_INPUT, "create: null output pointer");
  *out = nullptr;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return b200sp::set_error(nullptr, B200SP_CUDA_ERROR,
                             "no usable CUDA device (%s): libb200sp has no CPU fallback",
                             cudaGetErrorString(e));
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return b200sp::set_error(nullptr, B200SP_CUDA_ERROR, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  }
  if (prop.major != 10)
    return b200sp::set_error(nullptr, B200SP_CUDA_ERROR,
                             "device %d is sm_%d%d; libb200sp contains sm_100a code only", dev, prop.major,
                             prop.minor);
  b200sp_context *h = new b200sp_context();
  strcpy(h->err, "no error");
  h->device = dev;
  h->num_sms = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  h->l2_bytes = (size_t)prop.l2CacheSize;
  h->max_persist_l2 = prop.persistingL2CacheMaxSize;
  bool ok = cudaMalloc(&h->red_partials, (size_t)RED_MAX_PARTIALS * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&h->red_counters, 16 * sizeof(unsigned int)) == cudaSuccess &&
            cudaMalloc(&h->dev_scalars, 64 * sizeof(double)) == cudaSuccess &&
            cudaMallocHost(&h->pinned_scalars, 64 * sizeof(double)) == cudaSuccess &&
            cudaMemset(h->red_counters, 0, 16 * sizeof(unsigned int)) == cudaSuccess &&
            cudaMemset(h->dev_scalars, 0, 64 * sizeof(double)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    b200sp_destroy(h);
    return b200sp::set_error(nullptr, B200SP_ALLOC_FAILED, "create: cannot allocate handle workspace");
  }
  *out = h;
  return B200SP_OK;
}

b200sp_status b200sp_destroy(b200sp_handle h) {
  if (!h) return B200SP_OK;
  if (h->nccl_comm) b200sp_comm_destroy(h);
  if (h->scratch) cudaFree(h->scratch);
  if (h->red_partials) cudaFree(h->red_partials);
  if (h->red_counters) cudaFree(h->red_counters);
  if (h->dev_scalars) cudaFree(h->dev_scalars);
  if (h->pinned_scalars) cudaFreeHost(h->pinned_scalars);
  for (void *e : h->pipe_events) cudaEventDestroy((cudaEvent_t)e);
  if (h->copy_in_stream) cudaStreamDestroy((cudaStream_t)h->copy_in_stream);
  if (h->copy_out_stream) cudaStreamDestroy((cudaStream_t)h->copy_out_stream);
  if (h->stage_x) cudaFree(h->stage_x);
  if (h->stage_y) cudaFree(h->stage_y);
  if (h->cg_ws) cudaFree(h->cg_ws);
  if (h->cg_residuals) cudaFree(h->cg_residuals);
  for (void *ev : h->tune_events) cudaEventDestroy((cudaEvent_t)ev);
  for (void *ev : h->coo_choice_events) cudaEventDestroy((cudaEvent_t)ev);
  if (h->graph_event) cudaEventDestroy((cudaEvent_t)h->graph_event);
  if (h->graph_stream) cudaStreamDestroy((cudaStream_t)h->graph_stream);
  delete h;
  return B200SP_OK;
}

b200sp_status b200sp_set_l2_persist(b200sp_handle h, b200sp_stream stream, const void *ptr, size_t bytes) {
  B200SP_CHECK_HANDLE(h);
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (bytes == 0 || ptr == nullptr) {
    attr.accessPolicyWindow.base_ptr = nullptr;
    attr.accessPolicyWindow.num_bytes = 0;
    attr.accessPolicyWindow.hitRatio = 0.f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    B200SP_CUDA(h, cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    B200SP_CUDA(h, cudaCtxResetPersistingL2Cache());
    return B200SP_OK;
  }
  int max_window = 0;
  B200SP_CUDA(h, cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device));
  size_t carve = bytes < (size_t)h->max_persist_l2 ? bytes : (size_t)h->max_persist_l2;
  if (carve == 0) return b200sp::set_error(h, B200SP_NOT_IMPLEMENTED, "device has no persisting L2");
  B200SP_CUDA(h, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
  const size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
  attr.accessPolicyWindow.base_ptr = const_cast<void *>(ptr);
  attr.accessPolicyWindow.num_bytes = win;
  attr.accessPolicyWindow.hitRatio = (float)((double)carve / (double)win > 1.0 ? 1.0 : (double)carve / (double)win);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  B200SP_CUDA(h, cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  return B200SP_OK;
}
}  // extern "C"

namespace b200sp {

// DIA matrices are banded: rows [r0, r1) only read x[r0 + min_off, r1 + max_off).  The host
// path therefore runs as a pipeline over row chunks — x pieces go up on one copy stream, the
// chunk products run on the caller's stream as soon as their columns have landed, y chunks go
// down on a second copy stream — so the two PCIe directions and the kernels overlap instead of
// running back to back.  Same kernels, same per-row arithmetic: results are bit-identical to
// the one-shot path.  *done = 0 when the matrix does not qualify (caller falls back).
//
// Partitioned form (halo != nullptr, b200sp_spmv_dist_host): A is a rank's row block in window
// coordinates [halo_lo | local | halo_hi], x_host is the rank's LOCAL slice.  The two edge pieces
// go up first, the halo planes are exchanged with the neighbours (peer memory / NCCL) while the
// interior pieces are still in flight, and every chunk waits only for the pieces it reads.
static b200sp_status spmv_host_pipelined_dia(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A,
                                             const void *x_host, void *y_host, const b200sp_cfg *cfg,
                                             const b200sp_halo *halo, int *done) {
  *done = 0;
  const i64 K = A->num_cols_per_row, rows = A->num_rows, cols = A->num_cols;
  const size_t elem = A->dtype == B200SP_F64 ? 8 : 4;
  const i64 hlo = halo ? halo->halo_lo : 0;
  const i64 local_cols = halo ? rows : cols;  // columns that come from x_host, window columns [hlo, hlo + local_cols)
  if (K <= 0 || K > 1024 || rows < (1 << 20)) return B200SP_OK;
  std::vector<int> off((size_t)K);
  B200SP_CUDA(h, cudaMemcpyAsync(off.data(), A->diagonal_offsets, (size_t)K * sizeof(int), cudaMemcpyDeviceToHost, st));
  B200SP_CUDA(h, cudaStreamSynchronize(st));
  i64 lo = 0, up = 0;
  for (int o : off) {
    lo = o < -lo ? -(i64)o : lo;  // lo = max(0, -min_off)
    up = o > up ? (i64)o : up;    // up = max(0, max_off)
  }
  int want_chunks = 16;
  if (const char *e = getenv("B200SP_HOST_CHUNKS")) want_chunks = atoi(e) >= 3 ? atoi(e) : want_chunks;
  i64 chunk = ((ceil_div(rows, (i64)want_chunks) + 1023) / 1024) * 1024;
  if (chunk < lo + 1024 || chunk < up + 1024) return B200SP_OK;  // band wider than a chunk: nothing to overlap
  if (ceil_div(rows, chunk) < 3) return B200SP_OK;
  // Chunk boundaries (rows; the x pieces use the same boundaries as local columns, the last piece runs to local_cols).
  // Uniform: chunks that ramp from the smallest size the band allows at both ends to the uniform size in the middle
  // (so that the D2H direction starts after 0.04 instead of 0.35 ms) were measured and bought nothing — 3.417 against
  // 3.417 ms, 3.38 against 3.32 with y stored straight to host: the steady state, not the fill, sets the time
  // (tools/host_pipe_trace.py, tools/pcie_probe.py: 38 - 40 GB/s each way once kernels run between the copies,
  // 48 - 49 GB/s for the same copy pattern without them).
  std::vector<i64> bnd;
  for (i64 b = 0; b < rows; b += chunk) bnd.push_back(b);
  bnd.push_back(rows);
  const int nch = (int)bnd.size() - 1;
  const int npieces = nch;  // piece p = local columns [bnd[p], bnd[p+1]), the last one runs to local_cols
  auto piece_of = [&](i64 col) {
    int p = (int)(std::upper_bound(bnd.begin(), bnd.end(), col) - bnd.begin()) - 1;
    return p < 0 ? 0 : (p > npieces - 1 ? npieces - 1 : p);
  };

  if (!h->copy_in_stream) {
    cudaStream_t a, b;
    B200SP_CUDA(h, cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    B200SP_CUDA(h, cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    h->copy_in_stream = a;
    h->copy_out_stream = b;
  }
  const size_t need_events = (size_t)2 * nch + 2;
  const char *trace_env = getenv("B200SP_HOST_TRACE");  // 1: print the pipeline's timeline (ms since its start) to stderr
  const bool trace = trace_env && trace_env[0] == '1';
  if (trace && !h->pipe_events_timed) {  // events so far were created without timing: replace them once
    for (void *e : h->pipe_events) cudaEventDestroy((cudaEvent_t)e);
    h->pipe_events.clear();
    h->pipe_events_timed = true;
  }
  while (h->pipe_events.size() < need_events) {
    cudaEvent_t e;
    B200SP_CUDA(h, cudaEventCreateWithFlags(&e, h->pipe_events_timed ? cudaEventDefault : cudaEventDisableTiming));
    h->pipe_events.push_back(e);
  }
  cudaStream_t cin = (cudaStream_t)h->copy_in_stream, cout = (cudaStream_t)h->copy_out_stream;
  auto ev = [&](size_t i) { return (cudaEvent_t)h->pipe_events[i]; };

  // diagonal offsets shifted by `lo` for chunks that address x through a window starting at r0 - lo
  b200sp_status s = ensure_scratch(h, (size_t)K * sizeof(int));
  if (s != B200SP_OK) return s;
  std::vector<int> off_shift((size_t)K);
  for (i64 k = 0; k < K; ++k) off_shift[(size_t)k] = off[(size_t)k] + (int)lo;
  int *d_off_shift = reinterpret_cast<int *>(h->scratch);
  B200SP_CUDA(h, cudaMemcpyAsync(d_off_shift, off_shift.data(), (size_t)K * sizeof(int), cudaMemcpyHostToDevice, st));

  char *dx = reinterpret_cast<char *>(h->stage_x), *dy = reinterpret_cast<char *>(h->stage_y);
  const char *hx = reinterpret_cast<const char *>(x_host);
  char *hy = reinterpret_cast<char *>(y_host);
  // Pinned (device-mapped) y: the chunk products store y straight into host memory — posted PCIe writes from the
  // SMs — instead of into a staging buffer that a second DMA stream then copies down.  One stream fewer, no
  // per-chunk D2H launch, the last chunk's copy-out no longer trails the pipeline: 3.41 -> 3.31 ms per step on one GPU.
  // With several GPUs behind one host the SM-issued writes lose to the copy engines (2 GPUs: 4.54 against 4.09 ms), so
  // the partitioned form keeps the staged y.  B200SP_HOST_Y_DIRECT=0 / 1 forces either; pageable y is always staged.
  char *y_direct = nullptr;
  {
    const char *e = getenv("B200SP_HOST_Y_DIRECT");
    cudaPointerAttributes pa;
    const bool want = e ? e[0] == '1' : halo == nullptr;  // default: one GPU only (see above)
    if (want && cudaPointerGetAttributes(&pa, y_host) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
        pa.devicePointer != nullptr)
      y_direct = reinterpret_cast<char *>(pa.devicePointer);
    cudaGetLastError();
  }
  B200SP_CUDA(h, cudaEventRecord(ev(0), st));  // staging buffers are free once prior work on `st` is done
  B200SP_CUDA(h, cudaStreamWaitEvent(cin, ev(0), 0));
  B200SP_CUDA(h, cudaStreamWaitEvent(cout, ev(0), 0));
  // upload order: partitioned -> both edge pieces first (the neighbours need them), then the interior
  std::vector<int> order, ord_of((size_t)npieces);
  order.push_back(0);
  if (halo) order.push_back(npieces - 1);
  for (int p = 1; p < npieces; ++p)
    if (!(halo && p == npieces - 1)) order.push_back(p);
  for (int k = 0; k < npieces; ++k) {
    const int p = order[(size_t)k];
    ord_of[(size_t)p] = k;
    const i64 c0 = bnd[(size_t)p], c1 = (p == npieces - 1) ? local_cols : std::min(local_cols, bnd[(size_t)p + 1]);
    if (c1 > c0)
      B200SP_CUDA(h, cudaMemcpyAsync(dx + (size_t)(hlo + c0) * elem, hx + (size_t)c0 * elem, (size_t)(c1 - c0) * elem,
                                     cudaMemcpyHostToDevice, cin));
    B200SP_CUDA(h, cudaEventRecord(ev(1 + (size_t)k), cin));
  }
  if (halo) {
    B200SP_CUDA(h, cudaStreamWaitEvent(st, ev(1 + 1), 0));  // both edge pieces (stream order: piece 0, then the last)
    s = comm_halo_exchange_auto(h, st, dx, rows, halo->halo_lo, halo->halo_hi, elem);
    if (s != B200SP_OK) return s;
  }
  for (int c = 0; c < nch; ++c) {
    const i64 r0 = bnd[(size_t)c], r1 = bnd[(size_t)c + 1];
    // window columns the chunk reads -> local pieces -> the one uploaded last
    i64 first_col = r0 - lo - hlo, last_col = std::min(cols, r1 + up) - 1 - hlo;  // in local coordinates
    if (first_col < 0) first_col = 0;
    if (last_col > local_cols - 1) last_col = local_cols - 1;
    int p_lo = piece_of(first_col), p_hi = piece_of(last_col);
    if (p_lo > p_hi) p_lo = p_hi;
    int wait_ord = 0;
    for (int p = p_lo; p <= p_hi; ++p) wait_ord = std::max(wait_ord, ord_of[(size_t)p]);
    B200SP_CUDA(h, cudaStreamWaitEvent(st, ev(1 + (size_t)wait_ord), 0));
    b200sp_matrix sub = *A;
    sub.num_rows = r1 - r0;
    sub.values = reinterpret_cast<const char *>(A->values) + (size_t)r0 * elem;
    const void *xp;
    if (c == 0) {
      xp = dx;  // original offsets, x from column 0
    } else {
      sub.diagonal_offsets = d_off_shift;
      sub.num_cols = cols - (r0 - lo);
      xp = dx + (size_t)(r0 - lo) * elem;
    }
    sub.num_entries = 0;
    s = b200sp_spmv(h, (b200sp_stream)st, &sub, xp, (y_direct ? y_direct : dy) + (size_t)r0 * elem, 0, cfg);
    if (s != B200SP_OK) return s;
    if (y_direct) continue;
    B200SP_CUDA(h, cudaEventRecord(ev(1 + (size_t)nch + (size_t)c), st));
    B200SP_CUDA(h, cudaStreamWaitEvent(cout, ev(1 + (size_t)nch + (size_t)c), 0));
    B200SP_CUDA(h, cudaMemcpyAsync(hy + (size_t)r0 * elem, dy + (size_t)r0 * elem, (size_t)(r1 - r0) * elem,
                                   cudaMemcpyDeviceToHost, cout));
  }
  B200SP_CUDA(h, cudaEventRecord(ev(1 + 2 * (size_t)nch), cout));
  B200SP_CUDA(h, cudaStreamWaitEvent(st, ev(1 + 2 * (size_t)nch), 0));
  B200SP_CUDA(h, cudaStreamSynchronize(st));
  if (trace) {
    B200SP_CUDA(h, cudaStreamSynchronize(cin));
    B200SP_CUDA(h, cudaStreamSynchronize(cout));
    fprintf(stderr, "[b200sp host pipeline] %d chunks%s; x piece up / product done%s (ms since start)\n", nch,
            y_direct ? ", y stored straight to host" : "", y_direct ? "" : " (the D2H copy follows it)");
    for (int c = 0; c < nch; ++c) {
      float up_ms = 0.f, pr_ms = 0.f;
      cudaEventElapsedTime(&up_ms, ev(0), ev(1 + (size_t)ord_of[(size_t)c]));
      if (!y_direct) cudaEventElapsedTime(&pr_ms, ev(0), ev(1 + (size_t)nch + (size_t)c));
      fprintf(stderr, "  chunk %2d rows %9lld  x up %.3f  product %.3f\n", c, (long long)(bnd[(size_t)c + 1] - bnd[(size_t)c]), up_ms, pr_ms);
    }
    float end_ms = 0.f;
    if (!y_direct) cudaEventElapsedTime(&end_ms, ev(0), ev(1 + 2 * (size_t)nch));
    fprintf(stderr, "  last D2H done %.3f\n", end_ms);
    cudaGetLastError();
  }
  *done = 1;
  return B200SP_OK;
}

static b200sp_status ensure_host_staging(b200sp_handle h, size_t xb, size_t yb) {
  if (h->stage_x_bytes < xb) {
    if (h->stage_x) cudaFree(h->stage_x);
    h->stage_x = nullptr;
    h->stage_x_bytes = 0;
    if (cudaMalloc(&h->stage_x, xb ? xb : 1) != cudaSuccess) {
      cudaGetLastError();
      return set_error(h, B200SP_ALLOC_FAILED, "spmv_host: cannot stage x (%zu B)", xb);
    }
    h->stage_x_bytes = xb;
  }
  if (h->stage_y_bytes < yb) {
    if (h->stage_y) cudaFree(h->stage_y);
    h->stage_y = nullptr;
    h->stage_y_bytes = 0;
    if (cudaMalloc(&h->stage_y, yb ? yb : 1) != cudaSuccess) {
      cudaGetLastError();
      return set_error(h, B200SP_ALLOC_FAILED, "spmv_host: cannot stage y (%zu B)", yb);
    }
    h->stage_y_bytes = yb;
  }
  return B200SP_OK;
}

}  // namespace b200sp

extern "C" {

b200sp_status b200sp_spmv_host(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A,
                               const void *x_host, void *y_host, int accumulate, const b200sp_cfg *cfg) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A && x_host && y_host, "spmv_host: null argument");
  const size_t elem = A->dtype == B200SP_F64 ? 8 : 4;
  const size_t xb = (size_t)A->num_cols * elem, yb = (size_t)A->num_rows * elem;
  cudaStream_t st = (cudaStream_t)stream;
  b200sp_status s = b200sp::ensure_host_staging(h, xb, yb);
  if (s != B200SP_OK) return s;
  if (!accumulate && A->format == B200SP_FMT_DIA && !getenv("B200SP_HOST_ONE_SHOT")) {
    int done = 0;
    b200sp_status ps = b200sp::spmv_host_pipelined_dia(h, st, A, x_host, y_host, cfg, nullptr, &done);
    if (ps != B200SP_OK || done) return ps;
  }
  B200SP_CUDA(h, cudaMemcpyAsync(h->stage_x, x_host, xb, cudaMemcpyHostToDevice, st));
  if (accumulate) B200SP_CUDA(h, cudaMemcpyAsync(h->stage_y, y_host, yb, cudaMemcpyHostToDevice, st));
  s = b200sp_spmv(h, stream, A, h->stage_x, h->stage_y, accumulate, cfg);
  if (s != B200SP_OK) return s;
  B200SP_CUDA(h, cudaMemcpyAsync(y_host, h->stage_y, yb, cudaMemcpyDeviceToHost, st));
  B200SP_CUDA(h, cudaStreamSynchronize(st));
  return B200SP_OK;
}

b200sp_status b200sp_spmv_dist_host(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A_local,
                                    const b200sp_halo *halo, const void *x_host_local, void *y_host_local,
                                    const b200sp_cfg *cfg) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A_local && halo && x_host_local && y_host_local, "spmv_dist_host: null argument");
  B200SP_REQUIRE(h, h->nccl_comm, "spmv_dist_host: b200sp_comm_init has not been called");
  B200SP_REQUIRE(h, A_local->num_cols == A_local->num_rows + halo->halo_lo + halo->halo_hi,
                 "spmv_dist_host: num_cols != halo_lo + local + halo_hi");
  const size_t elem = A_local->dtype == B200SP_F64 ? 8 : 4;
  const size_t xb = (size_t)A_local->num_cols * elem, yb = (size_t)A_local->num_rows * elem;
  cudaStream_t st = (cudaStream_t)stream;
  b200sp_status s = b200sp::ensure_host_staging(h, xb, yb);
  if (s != B200SP_OK) return s;
  // Both paths exchange the halos with comm_halo_exchange_auto exactly once, so ranks whose blocks
  // qualify for the pipeline and ranks whose blocks do not still pair up.
  if (A_local->format == B200SP_FMT_DIA && !getenv("B200SP_HOST_ONE_SHOT")) {
    int done = 0;
    b200sp_status ps = b200sp::spmv_host_pipelined_dia(h, st, A_local, x_host_local, y_host_local, cfg, halo, &done);
    if (ps != B200SP_OK || done) return ps;
  }
  char *dx = reinterpret_cast<char *>(h->stage_x);
  B200SP_CUDA(h, cudaMemcpyAsync(dx + (size_t)halo->halo_lo * elem, x_host_local, yb, cudaMemcpyHostToDevice, st));
  s = b200sp::comm_halo_exchange_auto(h, st, dx, A_local->num_rows, halo->halo_lo, halo->halo_hi, elem);
  if (s != B200SP_OK) return s;
  s = b200sp_spmv(h, stream, A_local, dx, h->stage_y, 0, cfg);
  if (s != B200SP_OK) return s;
  B200SP_CUDA(h, cudaMemcpyAsync(y_host_local, h->stage_y, yb, cudaMemcpyDeviceToHost, st));
  B200SP_CUDA(h, cudaStreamSynchronize(st));
  return B200SP_OK;
}

/* ---- captured products (small, launch-bound systems) ---------------------------------------------------- */
struct b200sp_graph_s {
  cudaGraphExec_t exec;
  int count;
};

b200sp_status b200sp_spmv_graph_create(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A, const void *x,
                                       void *y, int accumulate, const b200sp_cfg *cfg, int count, b200sp_graph *out) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A && x && y && out && count >= 1 && count <= 4096, "spmv_graph_create: bad arguments");
  *out = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  // one plain product first: structure analysis, scratch growth and tuning-cache lookups happen outside the capture
  b200sp_status s = b200sp_spmv(h, stream, A, x, y, accumulate, cfg);
  if (s != B200SP_OK) return s;
  if (!h->graph_stream) {
    cudaStream_t gs;
    B200SP_CUDA(h, cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
    h->graph_stream = gs;
  }
  B200SP_CUDA(h, cudaStreamSynchronize(st));
  cudaStream_t gs = (cudaStream_t)h->graph_stream;
  cudaGraph_t graph = nullptr;
  B200SP_CUDA(h, cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
  const uint64_t launches_before = h->launches;
  for (int k = 0; k < count && s == B200SP_OK; ++k) s = b200sp_spmv(h, (b200sp_stream)gs, A, x, y, accumulate, cfg);
  const uint64_t per_replay = h->launches - launches_before;
  h->launches = launches_before;
  cudaError_t ce = cudaStreamEndCapture(gs, &graph);
  if (s != B200SP_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return s != B200SP_OK ? s : b200sp::set_error(h, B200SP_CUDA_ERROR, "spmv_graph_create: capture failed: %s", cudaGetErrorString(ce));
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    return b200sp::set_error(h, B200SP_CUDA_ERROR, "spmv_graph_create: cudaGraphInstantiate: %s", cudaGetErrorString(ce));
  }
  b200sp_graph g = new b200sp_graph_s();
  g->exec = exec;
  g->count = (int)per_replay;
  *out = g;
  return B200SP_OK;
}

b200sp_status b200sp_graph_launch(b200sp_handle h, b200sp_stream stream, b200sp_graph graph) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, graph != nullptr, "graph_launch: null graph");
  B200SP_CUDA(h, cudaGraphLaunch(graph->exec, (cudaStream_t)stream));
  h->launches += (uint64_t)graph->count;
  return B200SP_OK;
}

b200sp_status b200sp_graph_destroy(b200sp_handle h, b200sp_graph graph) {
  B200SP_CHECK_HANDLE(h);
  if (!graph) return B200SP_OK;
  cudaGraphExecDestroy(graph->exec);
  delete graph;
  return B200SP_OK;
}

int64_t b200sp_cfg_space(b200sp_format format, b200sp_dtype dtype, b200sp_cfg *out, int64_t capacity) {
  std::vector<b200sp_cfg> v = b200sp::cfg_space_vec(format, dtype);
  if (out)
    for (int64_t i = 0; i < (int64_t)v.size() && i < capacity; ++i) out[i] = v[(size_t)i];
  return (int64_t)v.size();
}

b200sp_status b200sp_tune_ex(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A, const void *x,
                             void *y, const void *y_reference, double tol, int repeats, const int64_t *order,
                             int64_t order_len, b200sp_tune_callback callback, void *user,
                             b200sp_tune_result *results, int64_t capacity, int64_t *num_results,
                             b200sp_cfg *best) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A && x && y, "tune: null argument");
  B200SP_REQUIRE(h, order_len >= 0 && (order || order_len == 0), "tune: bad search order");
  if (A->dtype == B200SP_F32)
    return b200sp::tune_impl<float>(h, (cudaStream_t)stream, A, (const float *)x, (float *)y,
                                    (const float *)y_reference, tol, repeats, order, order_len, callback, user, results,
                                    capacity, num_results, best);
  if (A->dtype == B200SP_F64)
    return b200sp::tune_impl<double>(h, (cudaStream_t)stream, A, (const double *)x, (double *)y,
                                     (const double *)y_reference, tol, repeats, order, order_len, callback, user, results,
                                     capacity, num_results, best);
  return b200sp::set_error(h, B200SP_INVALID_INPUT, "tune: unknown dtype");
}

b200sp_status b200sp_tune(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A, const void *x,
                          void *y, const void *y_reference, double tol, int repeats,
                          b200sp_tune_result *results, int64_t capacity, int64_t *num_results,
                          b200sp_cfg *best) {
  return b200sp_tune_ex(h, stream, A, x, y, y_reference, tol, repeats, nullptr, 0, nullptr, nullptr, results, capacity,
                        num_results, best);
}

b200sp_status b200sp_tune_step(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A, const void *x,
                               void *y, b200sp_tune_result *result) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A && x && y, "tune_step: null argument");
  if (A->dtype == B200SP_F32)
    return b200sp::tune_step_impl<float>(h, (cudaStream_t)stream, A, (const float *)x, (float *)y, result);
  if (A->dtype == B200SP_F64)
    return b200sp::tune_step_impl<double>(h, (cudaStream_t)stream, A, (const double *)x, (double *)y, result);
  return b200sp::set_error(h, B200SP_INVALID_INPUT, "tune_step: unknown dtype");
}

b200sp_status b200sp_tune_reset(b200sp_handle h, const b200sp_matrix *A) {
  B200SP_CHECK_HANDLE(h);
  if (!A)
    h->tune_cache.clear();
  else
    h->tune_cache.erase(b200sp::make_key(h, (cudaStream_t)0, A));
  return B200SP_OK;
}

int b200sp_tune_lookup(b200sp_handle h, const b200sp_matrix *A, b200sp_cfg *cfg) {
  return b200sp::tune_lookup_on(h, (cudaStream_t)0, A, cfg);
}

b200sp_status b200sp_tune_save(b200sp_handle h, const char *path) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, path, "tune_save: null path");
  FILE *f = fopen(path, "w");
  if (!f) return b200sp::set_error(h, B200SP_INVALID_INPUT, "tune_save: cannot open %s", path);
  fprintf(f, "# b200sp tuning cache v2: format dtype rows_log2 nnz_per_row_log2 structure kernel block tpr unroll vec "
             "tile stages ctas_per_sm ms\n");
  for (auto &kv : h->tune_cache) {
    if (!kv.second.has_best) continue;
    const b200sp_cfg &c = kv.second.best;
    fprintf(f, "%d %d %d %d %d %d %d %d %d %d %d %d %d %.6f\n", kv.first.format, kv.first.dtype,
            kv.first.rows_log2, kv.first.nnz_per_row_log2, kv.first.structure, c.kernel, c.block_size, c.threads_per_row,
            c.unroll, c.vector_width, c.tile_rows, c.stages, c.ctas_per_sm, kv.second.best_ms);
  }
  fclose(f);
  return B200SP_OK;
}

b200sp_status b200sp_tune_load(b200sp_handle h, const char *path) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, path, "tune_load: null path");
  FILE *f = fopen(path, "r");
  if (!f) return b200sp::set_error(h, B200SP_INVALID_INPUT, "tune_load: cannot open %s", path);
  char line[512];
  while (fgets(line, sizeof(line), f)) {
    if (line[0] == '#') continue;
    b200sp_tune_key k;
    b200sp_cfg c{};
    float ms = 0.f;
    if (sscanf(line, "%d %d %d %d %d %d %d %d %d %d %d %d %d %f", &k.format, &k.dtype, &k.rows_log2,
               &k.nnz_per_row_log2, &k.structure, &c.kernel, &c.block_size, &c.threads_per_row, &c.unroll,
               &c.vector_width, &c.tile_rows, &c.stages, &c.ctas_per_sm, &ms) != 14)
      continue;
    b200sp_tune_entry &e = h->tune_cache[k];
    e.best = c;
    e.has_best = true;
    e.best_ms = ms;
    e.next_index = 1 << 30;
  }
  fclose(f);
  return B200SP_OK;
}
}
