// spmv_coo_plan.cu — EXPERIMENTAL inspector / executor path for gather-bound COO products
// (power-law graphs).  Not used by cusp::multiply or b200sp_spmv: reached only through the explicit
// b200sp_coo_plan_* entry points.  Written after the round's GPU budget was spent — first hardware
// validation is tests/test_zz_plan_gpu.py (expected-failure tolerant).
//
// Why: on an R-MAT the product is bound by the x gathers (one L1TEX wavefront per distinct
// 128-byte line, ~31 lines per 32-lane instruction), not by HBM.  The column stream is skewed: the 32 k
// most frequent columns take 54 % of the gathers at scale 22.  The plan keeps x of the
// most frequent columns in a per-CTA shared-memory table, and a second copy of the column array in
// which those columns are replaced by (sign bit | table slot) — their gathers never reach L1TEX.
//
// Inspector (b200sp_coo_plan_create, device side, once per sparsity pattern):
//   1. histogram of column_indices (atomicAdd per entry),
//   2. the smallest count threshold t with #{c : count[c] >= t} <= table capacity (bisection, one count
//      kernel + a 4-byte read-back per step),
//   3. slot assignment of the selected columns (atomic counter; the order of slots has no effect on results),
//   4. the remapped column array.
// Executor (b200sp_spmv_coo_plan_<t>): persistent CTAs, one per SM, 1024 threads; prologue loads
// xs[slot] = x[hot_column[slot]]; then nnz-balanced tiles with the same per-thread serial + warp-shuffle
// segmented scan, carry records and fix-up kernel as K_COO_SEGSCAN (spmv_coo.cu) — for equal
// (BLOCK, VPT) the sums are grouped identically, so the result is bit-identical to that kernel's.
// The plan borrows row_indices (caller keeps them alive and unchanged); values are passed per call.
#include <algorithm>

#include "common.cuh"

struct b200sp_coo_plan_s {
  i64 rows, cols, nnz;
  const int *Ai;       // borrowed
  int *Aj_remapped;    // owned: column, or 0x80000000 | slot
  int *hot_cols;       // owned: column of every slot
  int hot;             // slots in use
  int capacity;        // slots the executor's table holds
  int elem;            // 4 / 8: value size the table was sized for
  i64 hot_entries;     // entries whose gather is served by the table
};

namespace b200sp {

constexpr int PLAN_BLOCK = 1024;
constexpr int PLAN_VPT = 7;
constexpr int PLAN_TILE = PLAN_BLOCK * PLAN_VPT;
constexpr unsigned HOT_FLAG = 0x80000000u;

template <typename T>
struct CooCarryP {  // same layout as CooCarry<T> in spmv_coo.cu (the fix-up below mirrors coo_fixup_kernel)
  int head_row, tail_row, leader, pad;
  T head_val, tail_val;
};

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

template <typename T>
struct PlanArgs {
  i64 rows, cols, nnz, tiles;
  const int *Ai, *Aj2, *hot_cols;
  const T *Ax, *x;
  T *y;
  int hot, accumulate;
  CooCarryP<T> *carry;
};

// dynamic shared memory: xs[capacity] | s_val[TILE] | s_row[TILE + 1] (T first: 8-byte alignment)
template <typename T>
__global__ void __launch_bounds__(PLAN_BLOCK, 1) coo_hot_kernel(PlanArgs<T> a, int capacity) {
  constexpr int BLOCK = PLAN_BLOCK, VPT = PLAN_VPT, TILE = PLAN_TILE, NW = BLOCK / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *xs = reinterpret_cast<T *>(smem_raw);
  T *s_val = xs + capacity;
  int *s_row = reinterpret_cast<int *>(s_val + TILE);
  __shared__ T s_wv[NW];
  __shared__ int s_wf[NW];
  __shared__ int s_head_row;
  __shared__ T s_head_val;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned cols = (unsigned)a.cols;
  for (int s = tid; s < a.hot; s += BLOCK) xs[s] = ld_ro(a.x + (unsigned)ld_ro(a.hot_cols + s));
  __syncthreads();

  for (i64 tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    const i64 start = tile * TILE;
    const int n = (int)min((i64)TILE, a.nnz - start);
    // ---- coalesced load + (table | global) gather + multiply -----------------------------
    {
      int r[VPT], c[VPT];
      T v[VPT], xv[VPT];
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const i64 g = min(start + i * BLOCK + tid, a.nnz - 1);
        r[i] = ld_stream(a.Ai + g);
        c[i] = ld_stream(a.Aj2 + g);
        v[i] = ld_stream(a.Ax + g);
      }
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        pin(c[i]);
        if (c[i] < 0)
          xv[i] = xs[min((unsigned)c[i] & ~HOT_FLAG, (unsigned)(a.hot - 1))];
        else
          xv[i] = ld_ro(a.x + min((unsigned)c[i], cols - 1));
      }
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        pin(xv[i]);
        const int idx = i * BLOCK + tid;
        const bool ok = idx < n;
        s_row[idx] = ok ? r[i] : -1;
        s_val[idx] = ok ? v[i] * xv[i] : T(0);
      }
    }
    int prev_row = -1;
    if (tid == 0) {
      s_row[TILE] = (start + TILE < a.nnz) ? a.Ai[start + TILE] : -1;
      s_head_row = -1;
      s_head_val = T(0);
    }
    if (start > 0) prev_row = ld_ro(a.Ai + start - 1);
    __syncthreads();

    // ---- per-thread serial segmented reduction over VPT consecutive entries ----------------
    int rr[VPT + 1];
    T pv[VPT];
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      rr[q] = s_row[tid * VPT + q];
      pv[q] = s_val[tid * VPT + q];
    }
    rr[VPT] = s_row[tid * VPT + VPT];
    T run = T(0), head = T(0);
    int head_row = -1;
    bool has_b = false;
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      run = run + pv[q];
      if (rr[q] != rr[q + 1]) {
        if (!has_b) {
          head = run;
          head_row = rr[q];
          has_b = true;
        } else if (rr[q] >= 0) {
          a.y[rr[q]] = a.accumulate ? a.y[rr[q]] + run : run;
        }
        run = T(0);
      }
    }
    // ---- block-wide segmented scan of (has_b, tail) ----------------------------------------
    T vi = run;
    int fi = has_b ? 1 : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const T vu = __shfl_up_sync(0xffffffffu, vi, d);
      const int fu = __shfl_up_sync(0xffffffffu, fi, d);
      if (lane >= d) {
        if (!fi) vi = vu + vi;
        fi |= fu;
      }
    }
    if (lane == 31) {
      s_wv[w] = vi;
      s_wf[w] = fi;
    }
    __syncthreads();
    T VW = T(0);
    for (int k = 0; k < w; ++k) VW = s_wf[k] ? s_wv[k] : VW + s_wv[k];
    const T Vi = fi ? vi : VW + vi;
    T carry_in = __shfl_up_sync(0xffffffffu, Vi, 1);
    if (lane == 0) carry_in = VW;
    if (has_b && head_row >= 0) {
      const T total = carry_in + head;
      if (head_row == prev_row) {
        s_head_row = head_row;
        s_head_val = total;
      } else {
        a.y[head_row] = a.accumulate ? a.y[head_row] + total : total;
      }
    }
    __syncthreads();
    if (tid == BLOCK - 1) {
      CooCarryP<T> cr;
      cr.head_row = s_head_row;
      cr.head_val = s_head_val;
      cr.pad = 0;
      const int last_row = rr[VPT - 1];
      if (last_row >= 0 && last_row == rr[VPT]) {
        cr.tail_row = last_row;
        cr.tail_val = Vi;
        cr.leader = (last_row != prev_row) ? 1 : 0;
      } else {
        cr.tail_row = -1;
        cr.tail_val = T(0);
        cr.leader = 0;
      }
      a.carry[tile] = cr;
    }
    __syncthreads();  // s_row / s_val / s_head_* / s_w* are rewritten by the next tile
  }
}

// one thread per tile: leaders walk their carry chain in tile order (as coo_fixup_kernel in spmv_coo.cu)
template <typename T>
__global__ void plan_fixup_kernel(i64 num_tiles, const CooCarryP<T> *carry, T *y, int accumulate) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tiles) return;
  const CooCarryP<T> me = carry[t];
  if (me.tail_row < 0 || !me.leader) return;
  const int row = me.tail_row;
  T total = me.tail_val;
  for (i64 u = t + 1; u < num_tiles; ++u) {
    const CooCarryP<T> nx = carry[u];
    if (nx.head_row == row) {
      total = total + nx.head_val;
      break;
    } else if (nx.tail_row == row && !nx.leader) {
      total = total + nx.tail_val;
    } else {
      break;
    }
  }
  y[row] = accumulate ? y[row] + total : total;
}

static size_t plan_smem_bytes(int capacity, size_t elem) {
  return (size_t)capacity * elem + (size_t)PLAN_TILE * elem + (size_t)(PLAN_TILE + 1) * sizeof(int);
}

template <typename T>
static b200sp_status spmv_coo_plan(b200sp_handle h, cudaStream_t st, b200sp_coo_plan p, const T *Ax, const T *x, T *y,
                                   int accumulate) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, p != nullptr, "coo plan: null plan");
  B200SP_REQUIRE(h, p->elem == (int)sizeof(T), "coo plan: created for the other value type");
  if (p->rows == 0) return B200SP_OK;
  B200SP_REQUIRE(h, y != nullptr, "coo plan: null pointer");
  if (!accumulate) B200SP_CUDA(h, cudaMemsetAsync(y, 0, (size_t)p->rows * sizeof(T), st));
  if (p->nnz == 0) return B200SP_OK;
  B200SP_REQUIRE(h, Ax && x, "coo plan: null pointer");
  const i64 tiles = ceil_div(p->nnz, (i64)PLAN_TILE);
  b200sp_status s = ensure_scratch(h, (size_t)tiles * sizeof(CooCarryP<T>));
  if (s != B200SP_OK) return s;
  PlanArgs<T> a;
  a.rows = p->rows; a.cols = p->cols; a.nnz = p->nnz; a.tiles = tiles;
  a.Ai = p->Ai; a.Aj2 = p->Aj_remapped; a.hot_cols = p->hot_cols; a.Ax = Ax; a.x = x; a.y = y;
  a.hot = p->hot; a.accumulate = accumulate;
  a.carry = reinterpret_cast<CooCarryP<T> *>(h->scratch);
  const size_t smem = plan_smem_bytes(p->capacity, sizeof(T));
  auto kern = coo_hot_kernel<T>;
  B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  i64 grid = h->num_sms;
  if (grid > tiles) grid = tiles;
  kern<<<(unsigned)grid, PLAN_BLOCK, smem, st>>>(a, p->capacity);
  B200SP_LAUNCH_CHECK(h, "coo_hot_kernel");
  plan_fixup_kernel<T><<<(unsigned)ceil_div(tiles, 256), 256, 0, st>>>(tiles, a.carry, y, accumulate);
  B200SP_LAUNCH_CHECK(h, "plan_fixup_kernel");
  return B200SP_OK;
}

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
  if (table_bytes <= 0) table_bytes = 128 << 10;
  int capacity = (int)std::min<int64_t>(table_bytes / (int64_t)elem, num_cols > 0 ? num_cols : 1);
  if (capacity < 1) capacity = 1;
  if (plan_smem_bytes(capacity, elem) + 2048 > (size_t)h->max_smem_optin)
    return set_error(h, B200SP_INVALID_INPUT, "coo plan: a %lld-byte table does not fit beside the tile buffers (%d B shared memory)",
                     (long long)table_bytes, h->max_smem_optin);
  b200sp_coo_plan p = new b200sp_coo_plan_s();
  p->rows = num_rows; p->cols = num_cols; p->nnz = num_entries; p->Ai = row_indices;
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
    plan_remap_kernel<<<g_nnz, 256, 0, st>>>(num_entries, column_indices, slot_of, p->Aj_remapped);
    h->launches += 2;
    int h8[8];
    ok = cudaMemcpyAsync(h8, scal, 8 * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
         cudaStreamSynchronize(st) == cudaSuccess;
    if (ok) {
      p->hot = std::min(h8[0], capacity);
      unsigned long long he;
      memcpy(&he, h8 + 4, sizeof(he));
      p->hot_entries = (i64)he;
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
  cudaFree(plan->Aj_remapped);
  cudaFree(plan->hot_cols);
  delete plan;
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
                                       const float *x, float *y, int accumulate) {
  return b200sp::spmv_coo_plan<float>(h, (cudaStream_t)stream, plan, values, x, y, accumulate);
}
b200sp_status b200sp_spmv_coo_plan_f64(b200sp_handle h, b200sp_stream stream, b200sp_coo_plan plan, const double *values,
                                       const double *x, double *y, int accumulate) {
  return b200sp::spmv_coo_plan<double>(h, (cudaStream_t)stream, plan, values, x, y, accumulate);
}
}
