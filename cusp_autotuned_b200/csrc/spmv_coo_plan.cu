// spmv_coo_plan.cu — inspector / executor COO product for gather-bound operators (power-law graphs).
//
// Why: on an R-MAT the product is bound by the x gathers (one L1TEX tag-stage cycle per
// distinct 128-byte line, ~31 lines per 32-lane gather instruction, and one 32-byte sector from L2 per
// gather), not by HBM.  The column stream is skewed: with the R-MAT parameters of BASELINE configs[2]
// the 55 k columns whose index has at most five 1-bits take ~45 % of the gathers at scale 24.  The plan
// keeps x of the most frequent columns in a per-CTA shared-memory table and a second copy of the column
// array in which those columns are replaced by (sign bit | table slot): their gathers never reach
// L1TEX or L2.
//
// Inspector (b200sp_coo_plan_create, device side, once per sparsity pattern):
//   1. histogram of column_indices (atomicAdd per entry),
//   2. the smallest count threshold t with #{c : count[c] >= t} <= table capacity (bisection, one count
//      kernel + a 16-byte read-back per step),
//   3. slot assignment of the selected columns, slots in ascending column order (the table is filled
//      from x with mostly coalesced reads),
//   4. the remapped column array.
// Executor (b200sp_spmv_coo_plan_<t>, or b200sp_spmv_coo_<t> on the arrays of an attached plan):
// K_COO_WARP with TABLE (coo_warp.cuh): one persistent 1024-thread CTA per SM, prologue
// tab[slot] = x[hot_column[slot]], then the same warp tiles, summation order, carry records and
// fix-up kernel as the plain kernel — for equal (vector_width, unroll) the result is bit-identical
// to K_COO_WARP's.  The plan borrows row_indices (caller keeps them alive and unchanged); values are
// passed per call.
#include <algorithm>

#include "coo_warp.cuh"

struct b200sp_coo_plan_s {
  i64 rows, cols, nnz;
  const int *Ai;       // borrowed
  const int *Aj;       // borrowed: the original column array (identity of an attached plan)
  int *Aj_remapped;    // owned: column, or 0x80000000 | slot
  int *hot_cols;       // owned: column of every slot, ascending
  int hot;             // slots in use
  int capacity;        // slots the executor's table holds
  int elem;            // 4 / 8: value size the table was sized for
  i64 hot_entries;     // entries whose gather is served by the table
};

namespace b200sp {

constexpr unsigned HOT_FLAG = COO_HOT_FLAG;

__global__ void plan_hist_kernel(i64 nnz, const int *Aj, int cols, int *cnt, int *bad) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (i64)gridDim.x * blockDim.x) {
    const int c = Aj[i];
    if ((unsigned)c < (unsigned)cols)
      atomicAdd(cnt + c, 1);
    else
      *bad = 1;
  }
}
// out[0] = #{c : cnt[c] >= t}, out[1] = max cnt (only when t == 0)
__global__ void plan_count_kernel(int cols, const int *cnt, int t, int *out) {
  int n = 0, m = 0;
  for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += (i64)gridDim.x * blockDim.x) {
    const int v = cnt[c];
    n += (v >= t) ? 1 : 0;
    m = max(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n += __shfl_down_sync(0xffffffffu, n, o);
    m = max(m, __shfl_down_sync(0xffffffffu, m, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (n) atomicAdd(out, n);
    atomicMax(out + 1, m);
  }
}
// slot_of[c] = slot (>= 0) for selected columns, -1 otherwise; out[0] counts slots, out[1] sums the entries served
__global__ void plan_assign_kernel(int cols, const int *cnt, int t, int capacity, int *slot_of, int *hot_cols,
                                   int *n_slots, unsigned long long *hot_entries) {
  for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += (i64)gridDim.x * blockDim.x) {
    int slot = -1;
    const int v = cnt[c];
    if (v >= t && v > 0) {
      slot = atomicAdd(n_slots, 1);
      if (slot < capacity) {
        hot_cols[slot] = (int)c;
        atomicAdd(hot_entries, (unsigned long long)v);
      } else {
        slot = -1;  // cannot happen when t came from the bisection; keeps the table in bounds regardless
      }
    }
    slot_of[c] = slot;
  }
}
__global__ void plan_remap_kernel(i64 nnz, const int *Aj, const int *slot_of, int *Aj2) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (i64)gridDim.x * blockDim.x) {
    const int c = Aj[i];
    const int s = slot_of[c];
    Aj2[i] = (s >= 0) ? (int)(HOT_FLAG | (unsigned)s) : c;
  }
}


// slot_of[hot_cols[s]] = s after the host put the selected columns in ascending order
__global__ void plan_slots_kernel(int hot, const int *hot_cols, int *slot_of) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < hot) slot_of[hot_cols[s]] = s;
}

template <typename T>
b200sp_status spmv_coo_hot(b200sp_handle h, cudaStream_t st, CooArgs<T> a, const b200sp_cfg &c, const int *hot_cols,
                           int hot, int capacity) {
  // defaults from the R-MAT sweep (profiles/r03_coo_probe.md): 256-bit loads, one unit per tile; entry streams are
  // kept out of the little L1 that remains beside the table (0.886 ms against 0.99 ms with ld.global.cs at scale 24)
  const int vpl = c.vector_width ? c.vector_width : (sizeof(T) == 4 ? 8 : 4), u = c.unroll ? c.unroll : 1;
  const uintptr_t need_idx = (uintptr_t)(4 * vpl) - 1, need_val = (uintptr_t)(sizeof(T) * vpl > 32 ? 32 : sizeof(T) * vpl) - 1;
  if (((uintptr_t)a.Ai & need_idx) || ((uintptr_t)a.Aj & need_idx) || ((uintptr_t)a.Ax & need_val))
    return set_error(h, B200SP_INVALID_INPUT, "coo plan: arrays not aligned for %d-entry vector loads", vpl);
#define CASE(V, UU) \
  if (vpl == V && u == UU) return launch_coo_warp<T, 1024, 1, V, UU, 0, 1, true>(h, st, a, 1, hot_cols, hot, capacity);
  if constexpr (sizeof(T) == 4) {
    CASE(4, 1) CASE(4, 2) CASE(8, 1)
  } else {
    CASE(4, 1) CASE(4, 2)
  }
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "coo plan: unsupported vector_width=%d unroll=%d", vpl, u);
}
template b200sp_status spmv_coo_hot<float>(b200sp_handle, cudaStream_t, CooArgs<float>, const b200sp_cfg &, const int *,
                                           int, int);
template b200sp_status spmv_coo_hot<double>(b200sp_handle, cudaStream_t, CooArgs<double>, const b200sp_cfg &,
                                            const int *, int, int);

template <typename T>
static b200sp_status spmv_coo_plan(b200sp_handle h, cudaStream_t st, b200sp_coo_plan p, const T *Ax, const T *x, T *y,
                                   int accumulate, const b200sp_cfg *cfg) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, p != nullptr, "coo plan: null plan");
  B200SP_REQUIRE(h, p->elem == (int)sizeof(T), "coo plan: the plan was created for the other value type");
  if (p->rows == 0) return B200SP_OK;
  B200SP_REQUIRE(h, y != nullptr, "coo plan: null pointer");
  if (!accumulate) B200SP_CUDA(h, cudaMemsetAsync(y, 0, (size_t)p->rows * sizeof(T), st));
  if (p->nnz == 0) return B200SP_OK;
  B200SP_REQUIRE(h, Ax && x, "coo plan: null pointer");
  CooArgs<T> a;
  a.rows = p->rows; a.cols = p->cols; a.nnz = p->nnz; a.Ai = p->Ai; a.Aj = p->Aj_remapped; a.Ax = Ax; a.x = x; a.y = y;
  a.accumulate = accumulate;
  a.carry = nullptr; a.Ap = nullptr; a.tile_first_row = nullptr; a.scalar_loads = 0;
  const b200sp_cfg c = cfg ? *cfg : b200sp_cfg{};
  return spmv_coo_hot<T>(h, st, a, c, p->hot_cols, p->hot, p->capacity);
}

// the plan attached to this handle for exactly these arrays, or nullptr (spmv_coo.cu asks)
b200sp_coo_plan coo_attached_plan(b200sp_handle h, i64 rows, i64 cols, i64 nnz, const int *Ai, const int *Aj,
                                  size_t elem) {
  for (void *q : h->coo_plans) {
    b200sp_coo_plan p = reinterpret_cast<b200sp_coo_plan>(q);
    if (p->Ai == Ai && p->Aj == Aj && p->nnz == nnz && p->rows == rows && p->cols == cols && p->elem == (int)elem)
      return p;
  }
  return nullptr;
}
template <typename T>
b200sp_status spmv_coo_attached(b200sp_handle h, cudaStream_t st, b200sp_coo_plan p, const T *Ax, const T *x, T *y,
                                int accumulate, const b200sp_cfg *cfg) {
  return spmv_coo_plan<T>(h, st, p, Ax, x, y, accumulate, cfg);
}
template b200sp_status spmv_coo_attached<float>(b200sp_handle, cudaStream_t, b200sp_coo_plan, const float *,
                                                const float *, float *, int, const b200sp_cfg *);
template b200sp_status spmv_coo_attached<double>(b200sp_handle, cudaStream_t, b200sp_coo_plan, const double *,
                                                 const double *, double *, int, const b200sp_cfg *);

}  // namespace b200sp

extern "C" {

b200sp_status b200sp_coo_plan_create(b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols,
                                     int64_t num_entries, const int32_t *row_indices, const int32_t *column_indices,
                                     b200sp_dtype dtype, int64_t table_bytes, b200sp_coo_plan *out) {
  using namespace b200sp;
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, out != nullptr, "coo plan: null output");
  *out = nullptr;
  B200SP_REQUIRE(h, num_rows >= 0 && num_cols >= 0 && num_entries >= 0, "coo plan: negative dimension");
  B200SP_REQUIRE(h, num_rows < (1ll << 31) && num_cols < (1ll << 31), "coo plan: int32 index range");
  B200SP_REQUIRE(h, dtype == B200SP_F32 || dtype == B200SP_F64, "coo plan: dtype");
  B200SP_REQUIRE(h, num_entries == 0 || (row_indices && column_indices && num_cols > 0), "coo plan: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t elem = dtype == B200SP_F64 ? 8 : 4;
  // default: 128 KiB.  The executor's gathers of the remaining (cold) columns still need L1: with a 192 KiB table
  // (64 KiB of L1 left) the same product takes 1.45 ms instead of 0.89 ms at R-MAT scale 24, with the whole shared
  // memory as table 2.7 ms (profiles/r03_coo_probe.md)
  if (table_bytes <= 0) table_bytes = 128 << 10;
  int capacity = (int)std::min<int64_t>(table_bytes / (int64_t)elem, num_cols > 0 ? num_cols : 1);
  if (capacity < 1) capacity = 1;
  if ((size_t)capacity * elem > (size_t)h->max_smem_optin)
    return set_error(h, B200SP_INVALID_INPUT, "coo plan: a %lld-byte table does not fit in %d B of shared memory",
                     (long long)table_bytes, h->max_smem_optin);
  b200sp_coo_plan p = new b200sp_coo_plan_s();
  p->rows = num_rows; p->cols = num_cols; p->nnz = num_entries; p->Ai = row_indices; p->Aj = column_indices;
  p->Aj_remapped = nullptr; p->hot_cols = nullptr; p->hot = 0; p->capacity = capacity; p->elem = (int)elem;
  p->hot_entries = 0;
  int *cnt = nullptr, *slot_of = nullptr, *scal = nullptr;
  auto fail = [&](b200sp_status s) {
    cudaFree(cnt);
    cudaFree(slot_of);
    cudaFree(scal);
    cudaFree(p->Aj_remapped);
    cudaFree(p->hot_cols);
    delete p;
    cudaGetLastError();
    return s;
  };
  const size_t ncols = (size_t)(num_cols > 0 ? num_cols : 1), nnz1 = (size_t)(num_entries > 0 ? num_entries : 1);
  if (cudaMalloc(&cnt, ncols * sizeof(int)) != cudaSuccess || cudaMalloc(&slot_of, ncols * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&scal, 8 * sizeof(int)) != cudaSuccess || cudaMalloc(&p->Aj_remapped, nnz1 * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&p->hot_cols, (size_t)capacity * sizeof(int)) != cudaSuccess)
    return fail(set_error(h, B200SP_ALLOC_FAILED, "coo plan: device allocation failed"));
  if (num_entries == 0) {
    cudaFree(cnt);
    cudaFree(slot_of);
    cudaFree(scal);
    *out = p;
    return B200SP_OK;
  }
  const unsigned g_nnz = (unsigned)std::min<i64>(ceil_div(num_entries, 256), (i64)h->num_sms * 32);
  const unsigned g_col = (unsigned)std::min<i64>(ceil_div(num_cols, 256), (i64)h->num_sms * 32);
  int host[4] = {0, 0, 0, 0};
  auto read4 = [&](int *dst) {
    return cudaMemcpyAsync(dst, scal, 4 * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
           cudaStreamSynchronize(st) == cudaSuccess;
  };
  bool ok = cudaMemsetAsync(cnt, 0, ncols * sizeof(int), st) == cudaSuccess &&
            cudaMemsetAsync(scal, 0, 8 * sizeof(int), st) == cudaSuccess;
  if (ok) {
    plan_hist_kernel<<<g_nnz, 256, 0, st>>>(num_entries, column_indices, (int)num_cols, cnt, scal + 2);
    plan_count_kernel<<<g_col, 256, 0, st>>>((int)num_cols, cnt, 1, scal);
    h->launches += 2;
    ok = read4(host);
  }
  if (!ok) return fail(set_error(h, B200SP_CUDA_ERROR, "coo plan: histogram failed: %s", cudaGetErrorString(cudaGetLastError())));
  if (host[2]) return fail(set_error(h, B200SP_INVALID_INPUT, "coo plan: column index outside [0, num_cols)"));
  // smallest t >= 1 with #{cnt >= t} <= capacity  (host[0] = #{cnt >= 1}, host[1] = max cnt)
  int t = 1;
  if (host[0] > capacity) {
    int lo = 1, hi = host[1] + 1;  // count(lo) > capacity, count(hi) == 0 <= capacity
    while (hi - lo > 1) {
      const int mid = lo + (hi - lo) / 2;
      int c4[4];
      ok = cudaMemsetAsync(scal, 0, 2 * sizeof(int), st) == cudaSuccess;
      if (ok) {
        plan_count_kernel<<<g_col, 256, 0, st>>>((int)num_cols, cnt, mid, scal);
        h->launches++;
        ok = read4(c4);
      }
      if (!ok) return fail(set_error(h, B200SP_CUDA_ERROR, "coo plan: threshold search failed"));
      if (c4[0] > capacity)
        lo = mid;
      else
        hi = mid;
    }
    t = hi;
  }
  ok = cudaMemsetAsync(scal, 0, 8 * sizeof(int), st) == cudaSuccess;
  if (ok) {
    plan_assign_kernel<<<g_col, 256, 0, st>>>((int)num_cols, cnt, t, capacity, slot_of, p->hot_cols, scal,
                                              reinterpret_cast<unsigned long long *>(scal + 4));
    h->launches++;
    int h8[8];
    ok = cudaMemcpyAsync(h8, scal, 8 * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
         cudaStreamSynchronize(st) == cudaSuccess;
    if (ok) {
      p->hot = std::min(h8[0], capacity);
      unsigned long long he;
      memcpy(&he, h8 + 4, sizeof(he));
      p->hot_entries = (i64)he;
      // slots in ascending column order: the atomic counter handed them out in arrival order
      std::vector<int> hc((size_t)std::max(p->hot, 1));
      ok = cudaMemcpyAsync(hc.data(), p->hot_cols, (size_t)p->hot * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
           cudaStreamSynchronize(st) == cudaSuccess;
      if (ok && p->hot > 0) {
        std::sort(hc.begin(), hc.begin() + p->hot);
        ok = cudaMemcpyAsync(p->hot_cols, hc.data(), (size_t)p->hot * sizeof(int), cudaMemcpyHostToDevice, st) == cudaSuccess;
        if (ok) {
          plan_slots_kernel<<<(unsigned)ceil_div(p->hot, 256), 256, 0, st>>>(p->hot, p->hot_cols, slot_of);
          h->launches++;
          ok = cudaStreamSynchronize(st) == cudaSuccess;  // hc goes out of scope
        }
      }
    }
    if (ok) {
      plan_remap_kernel<<<g_nnz, 256, 0, st>>>(num_entries, column_indices, slot_of, p->Aj_remapped);
      h->launches++;
      ok = cudaStreamSynchronize(st) == cudaSuccess;
    }
  }
  if (!ok) return fail(set_error(h, B200SP_CUDA_ERROR, "coo plan: remapping failed: %s", cudaGetErrorString(cudaGetLastError())));
  cudaFree(cnt);
  cudaFree(slot_of);
  cudaFree(scal);
  *out = p;
  return B200SP_OK;
}

b200sp_status b200sp_coo_plan_destroy(b200sp_handle h, b200sp_coo_plan plan) {
  B200SP_CHECK_HANDLE(h);
  if (!plan) return B200SP_OK;
  b200sp_coo_plan_detach(h, plan);
  cudaFree(plan->Aj_remapped);
  cudaFree(plan->hot_cols);
  delete plan;
  return B200SP_OK;
}

b200sp_status b200sp_coo_plan_attach(b200sp_handle h, b200sp_coo_plan plan) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, plan != nullptr, "coo plan: null plan");
  for (void *q : h->coo_plans)
    if (q == plan) return B200SP_OK;
  h->coo_plans.push_back(plan);
  return B200SP_OK;
}

b200sp_status b200sp_coo_plan_detach(b200sp_handle h, b200sp_coo_plan plan) {
  B200SP_CHECK_HANDLE(h);
  for (size_t i = 0; i < h->coo_plans.size(); ++i)
    if (h->coo_plans[i] == plan) {
      h->coo_plans.erase(h->coo_plans.begin() + (long)i);
      break;
    }
  return B200SP_OK;
}

b200sp_status b200sp_coo_plan_info(b200sp_coo_plan plan, int64_t *hot_columns, int64_t *hot_entries, int64_t *capacity) {
  if (!plan) return B200SP_INVALID_INPUT;
  if (hot_columns) *hot_columns = plan->hot;
  if (hot_entries) *hot_entries = plan->hot_entries;
  if (capacity) *capacity = plan->capacity;
  return B200SP_OK;
}

b200sp_status b200sp_spmv_coo_plan_f32(b200sp_handle h, b200sp_stream stream, b200sp_coo_plan plan, const float *values,
                                       const float *x, float *y, int accumulate, const b200sp_cfg *cfg) {
  return b200sp::spmv_coo_plan<float>(h, (cudaStream_t)stream, plan, values, x, y, accumulate, cfg);
}
b200sp_status b200sp_spmv_coo_plan_f64(b200sp_handle h, b200sp_stream stream, b200sp_coo_plan plan, const double *values,
                                       const double *x, double *y, int accumulate, const b200sp_cfg *cfg) {
  return b200sp::spmv_coo_plan<double>(h, (cudaStream_t)stream, plan, values, x, y, accumulate, cfg);
}
}
