// spmv_ell.cu — ELL / ELL-R SpMV for sm_100a.
//
// Replaces spmv_ell_kernel (cusp/system/cuda/detail/multiply/ell_spmv.h:47-93)
// and the KTT kernels ktt_ell_kernel / ktt_ellr_kernel
// (cusp/system/cuda/ktt/kernels/ell_kernel.h:86-213).  Semantics follow the host
// loop cusp/system/detail/sequential/multiply/ell_spmv.h:34-76:
//     y[i] = init(y[i]);  for n ascending: if (col(i,n) != -1) y[i] += val(i,n)*x[col(i,n)]
// One thread owns a row and walks the slots in ascending n, so with -fmad=false
// the result is bit-identical to that loop.
//
// Layout (HBM): column_indices / values are column-major with pitch: slot (row,n)
// at [n*pitch + row]; padding is col = -1, val = 0 (cusp/ell_matrix.h:129).
//
//  K_ELL_LDG : RPT rows per thread strided by BLOCK (coalesced), 8 slots unrolled
//              -> 8*RPT index loads, 8*RPT value loads and 8*RPT x gathers in
//              flight per thread.  ld.global.cs for the slabs, ld.global.nc for x.
//  K_ELL_BULK: persistent CTAs, producer lane stages [KC slots x R rows] of both
//              slabs with cp.async.bulk (UBLKCP) + mbarrier ring, L2 evict-first.
//
// Algorithmic bytes / row: K*(4+sizeof(T)) + sizeof(T) x + sizeof(T) y.
#include "common.cuh"

namespace b200sp {

template <typename T, int MODE>
b200sp_status reduce(b200sp_handle, cudaStream_t, i64, const T *, const T *, T *, T *);  // blas1.cu

template <typename T>
struct EllArgs {
  i64 rows, cols, pitch;
  int K;
  const int *cidx;
  const T *vals;
  const int *row_lengths;  // ELL-R, may be null
  const T *x;
  T *y;
  int accumulate;
  const T *dotv;
  T *dot_partials;
  unsigned int *dot_ticket;
  T *dot_result;
};

constexpr int ELL_DU = 8;

template <typename T, int BLOCK, int RPT>
__global__ void __launch_bounds__(BLOCK) ell_ldg_kernel(EllArgs<T> a) {
  __shared__ T s_red[32];
  const unsigned rows = (unsigned)a.rows, cols = (unsigned)a.cols;
  const unsigned base = blockIdx.x * (unsigned)(BLOCK * RPT) + threadIdx.x;
  T acc[RPT];
  unsigned rc[RPT];
  int len[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const unsigned r = base + i * BLOCK;
    rc[i] = min(r, rows - 1);
    acc[i] = (a.accumulate && r < rows) ? a.y[r] : T(0);
    len[i] = a.row_lengths ? ld_ro(a.row_lengths + rc[i]) : a.K;
  }

  int k0 = 0;
  for (; k0 + ELL_DU <= a.K; k0 += ELL_DU) {
    int c[ELL_DU][RPT];
    T v[ELL_DU][RPT], xv[ELL_DU][RPT];
#pragma unroll
    for (int u = 0; u < ELL_DU; ++u) {
      const i64 so = (i64)(k0 + u) * a.pitch;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        c[u][i] = ld_stream(a.cidx + so + rc[i]);
        v[u][i] = ld_stream(a.vals + so + rc[i]);
      }
    }
#pragma unroll
    for (int u = 0; u < ELL_DU; ++u)
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        pin(c[u][i]);
        xv[u][i] = ld_ro(a.x + min((unsigned)max(c[u][i], 0), cols - 1));
      }
#pragma unroll
    for (int u = 0; u < ELL_DU; ++u)
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        pin(xv[u][i]);
        const T t = acc[i] + v[u][i] * xv[u][i];
        acc[i] = (c[u][i] != -1 && k0 + u < len[i]) ? t : acc[i];
      }
  }
  for (; k0 < a.K; ++k0) {
    const i64 so = (i64)k0 * a.pitch;
    int c[RPT];
    T v[RPT], xv[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      c[i] = ld_stream(a.cidx + so + rc[i]);
      v[i] = ld_stream(a.vals + so + rc[i]);
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) xv[i] = ld_ro(a.x + min((unsigned)max(c[i], 0), cols - 1));
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const T t = acc[i] + v[i] * xv[i];
      acc[i] = (c[i] != -1 && k0 < len[i]) ? t : acc[i];
    }
  }

  T dsum = 0;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const unsigned r = base + i * BLOCK;
    if (r < rows) {
      a.y[r] = acc[i];
      if (a.dotv) dsum = dsum + acc[i] * ld_ro(a.dotv + r);
    }
  }
  if (a.dotv) {
    T bs = block_sum<BLOCK>(dsum, s_red);
    grid_reduce_finish<BLOCK>(bs, a.dot_partials, a.dot_ticket, s_red,
                              [&](T total) { *a.dot_result = total; });
  }
}

// ---------------------------------------------------------------------------
// bulk-async staged variant
// ---------------------------------------------------------------------------
constexpr int ELL_KC = 4;  // slots per stage

template <typename T, int BLOCK, int RPT>
__global__ void __launch_bounds__(BLOCK + 32) ell_bulk_kernel(EllArgs<T> a, int stages, i64 num_tiles) {
  constexpr int R = BLOCK * RPT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: vals [stages][KC][R] T | cidx [stages][KC][R] int | full[stages] | empty[stages]
  T *s_vals = reinterpret_cast<T *>(smem_raw);
  int *s_cidx = reinterpret_cast<int *>(smem_raw + (size_t)stages * ELL_KC * R * sizeof(T));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stages * ELL_KC * R * (sizeof(T) + sizeof(int)));
  uint64_t *empty = full + stages;
  __shared__ T s_red[32];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], BLOCK / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int nchunks = (a.K + ELL_KC - 1) / ELL_KC;
  const unsigned rows = (unsigned)a.rows, cols = (unsigned)a.cols;
  T dsum = 0;

  if (tid >= BLOCK) {
    if (tid == BLOCK) {
      const uint64_t pol = l2_policy_evict_first();
      int s = 0;
      uint32_t ph = 0;
      for (i64 tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const i64 r0 = tile * R;
        if (r0 + R > a.rows) continue;
        for (int c = 0; c < nchunks; ++c) {
          const int k0 = c * ELL_KC;
          const int kc = min(ELL_KC, a.K - k0);
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], (uint32_t)(kc * R * (sizeof(T) + sizeof(int))));
          for (int u = 0; u < kc; ++u) {
            const i64 so = (i64)(k0 + u) * a.pitch + r0;
            bulk_g2s(s_cidx + ((size_t)s * ELL_KC + u) * R, a.cidx + so, (uint32_t)(R * sizeof(int)),
                     &full[s], pol);
            bulk_g2s(s_vals + ((size_t)s * ELL_KC + u) * R, a.vals + so, (uint32_t)(R * sizeof(T)),
                     &full[s], pol);
          }
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    int s = 0;
    uint32_t ph = 0;
    const int lane = tid & 31;
    for (i64 tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const unsigned r0 = (unsigned)(tile * R);
      if ((i64)r0 + R > a.rows) {
        // ragged last tile: same arithmetic straight from global
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const unsigned r = r0 + tid + i * BLOCK;
          if (r >= rows) continue;
          T s_acc = a.accumulate ? a.y[r] : T(0);
          const int len = a.row_lengths ? a.row_lengths[r] : a.K;
          for (int k = 0; k < a.K; ++k) {
            const int cc = ld_stream(a.cidx + (i64)k * a.pitch + r);
            if (cc != -1 && k < len)
              s_acc = s_acc + ld_stream(a.vals + (i64)k * a.pitch + r) * ld_ro(a.x + min((unsigned)max(cc, 0), cols - 1));
          }
          a.y[r] = s_acc;
          if (a.dotv) dsum = dsum + s_acc * ld_ro(a.dotv + r);
        }
        continue;
      }
      T acc[RPT];
      int len[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        acc[i] = a.accumulate ? a.y[r0 + tid + i * BLOCK] : T(0);
        len[i] = a.row_lengths ? ld_ro(a.row_lengths + r0 + tid + i * BLOCK) : a.K;
      }
      for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * ELL_KC;
        const int kc = min(ELL_KC, a.K - k0);
        mbar_wait(&full[s], ph);
        const T *sv = s_vals + (size_t)s * ELL_KC * R;
        const int *sc = s_cidx + (size_t)s * ELL_KC * R;
        int cc[ELL_KC][RPT];
        T xv[ELL_KC][RPT];
#pragma unroll
        for (int u = 0; u < ELL_KC; ++u)
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            cc[u][i] = (u < kc) ? sc[u * R + tid + i * BLOCK] : -1;
            xv[u][i] = ld_ro(a.x + min((unsigned)max(cc[u][i], 0), cols - 1));
          }
#pragma unroll
        for (int u = 0; u < ELL_KC; ++u)
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            pin(xv[u][i]);
            const T vv = (u < kc) ? sv[u * R + tid + i * BLOCK] : T(0);
            const T t = acc[i] + vv * xv[u][i];
            acc[i] = (cc[u][i] != -1 && k0 + u < len[i]) ? t : acc[i];
          }
#pragma unroll
        for (int i = 0; i < RPT; ++i) consume_before_release(acc[i]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const unsigned r = r0 + tid + i * BLOCK;
        a.y[r] = acc[i];
        if (a.dotv) dsum = dsum + acc[i] * ld_ro(a.dotv + r);
      }
    }
  }
  if (a.dotv) {
    if (tid >= BLOCK) dsum = 0;
    T bs = block_sum<BLOCK + 32>(dsum, s_red);
    grid_reduce_finish<BLOCK + 32>(bs, a.dot_partials, a.dot_ticket, s_red,
                                   [&](T total) { *a.dot_result = total; });
  }
}

// row_lengths for ELL-R (cusp/ktt/detail/ellr_matrix.inl:16-52): number of
// leading slots with a non-negative column index.
__global__ void ell_row_lengths_kernel(i64 rows, int K, i64 pitch, const int *cidx, int *out) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int n = 0;
  while (n < K && cidx[(i64)n * pitch + r] >= 0) ++n;
  out[r] = n;
}

// ---------------------------------------------------------------------------
template <typename T, int BLOCK, int RPT>
static b200sp_status launch_ldg(b200sp_handle h, cudaStream_t st, EllArgs<T> a) {
  const i64 grid = ceil_div(a.rows, (i64)BLOCK * RPT);
  // the fused <y, dotv> epilogue keeps one partial per CTA: beyond the workspace the dot runs as its own
  // deterministic reduction after the product
  const T *late_dotv = nullptr;
  if (a.dotv && grid > RED_MAX_PARTIALS) {
    late_dotv = a.dotv;
    a.dotv = nullptr;
  }
  ell_ldg_kernel<T, BLOCK, RPT><<<(unsigned)grid, BLOCK, 0, st>>>(a);
  B200SP_LAUNCH_CHECK(h, "ell_ldg_kernel");
  if (late_dotv) return reduce<T, 0>(h, st, a.rows, a.y, late_dotv, a.dot_result, nullptr);
  return B200SP_OK;
}

template <typename T>
static b200sp_status dispatch_ldg(b200sp_handle h, cudaStream_t st, const EllArgs<T> &a, int block, int rpt) {
#define CASE(B, R) \
  if (block == B && rpt == R) return launch_ldg<T, B, R>(h, st, a);
  CASE(128, 1) CASE(128, 2) CASE(128, 4) CASE(256, 1) CASE(256, 2) CASE(256, 4) CASE(512, 1)
  CASE(512, 2) CASE(512, 4)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "ell ldg: unsupported block_size=%d unroll=%d", block, rpt);
}

template <typename T, int BLOCK, int RPT>
static b200sp_status launch_bulk(b200sp_handle h, cudaStream_t st, EllArgs<T> a, i64 num_tiles, int stages,
                                 int ctas_per_sm) {
  constexpr int R = BLOCK * RPT;
  auto kern = ell_bulk_kernel<T, BLOCK, RPT>;
  size_t smem = (size_t)stages * ELL_KC * R * (sizeof(T) + sizeof(int)) + 2 * stages * sizeof(uint64_t) + 16;
  if (smem > (size_t)h->max_smem_optin)
    return set_error(h, B200SP_INVALID_INPUT, "ell bulk: %zu B smem exceeds %d", smem, h->max_smem_optin);
  B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent grid: no more CTAs than are resident at once (a second wave would start
  // only after a first-wave CTA has finished all of its tiles)
  int resident = 0;
  B200SP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BLOCK + 32, smem));
  if (resident < 1) return set_error(h, B200SP_INVALID_INPUT, "ell bulk: configuration does not fit on an SM");
  i64 grid = (i64)h->num_sms * (ctas_per_sm < resident ? ctas_per_sm : resident);
  if (grid > num_tiles) grid = num_tiles;
  kern<<<(unsigned)grid, BLOCK + 32, smem, st>>>(a, stages, num_tiles);
  B200SP_LAUNCH_CHECK(h, "ell_bulk_kernel");
  return B200SP_OK;
}

template <typename T>
static b200sp_status dispatch_bulk(b200sp_handle h, cudaStream_t st, const EllArgs<T> &a, i64 num_tiles,
                                   int block, int rpt, int stages, int cps) {
#define CASE(B, R) \
  if (block == B && rpt == R) return launch_bulk<T, B, R>(h, st, a, num_tiles, stages, cps);
  CASE(128, 2) CASE(128, 4) CASE(128, 8) CASE(256, 1) CASE(256, 2) CASE(256, 4)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "ell bulk: unsupported block_size=%d unroll=%d", block, rpt);
}

static void ell_defaults(b200sp_cfg &c, size_t elem) {
  // round-1 sweep on B200, poisson7pt 256^3 (profiles/r01_probe_256.md)
  if (c.kernel == 0) c.kernel = B200SP_K_ELL_BULK;
  if (c.block_size == 0) c.block_size = 256;
  if (c.unroll == 0) c.unroll = (c.kernel == B200SP_K_ELL_BULK) ? ((elem == 4) ? 2 : 1) : 4;
  if (c.stages == 0) c.stages = 3;
  if (c.ctas_per_sm == 0) c.ctas_per_sm = 4;
}

template <typename T>
b200sp_status spmv_ell(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 K, i64 pitch,
                       const int *cidx, const T *vals, const int *row_lengths, const T *x, T *y,
                       int accumulate, const b200sp_cfg *cfg, const T *dotv, T *dot_result) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, rows >= 0 && cols >= 0 && K >= 0, "ell: negative dimension");
  B200SP_REQUIRE(h, rows < (1ll << 31) && cols < (1ll << 31), "ell: int32 index range");
  B200SP_REQUIRE(h, pitch >= rows, "ell: pitch < num_rows");
  if (rows == 0) {
    if (dot_result) B200SP_CUDA(h, cudaMemsetAsync(dot_result, 0, sizeof(T), st));
    return B200SP_OK;
  }
  B200SP_REQUIRE(h, y != nullptr && (K == 0 || (cidx && vals && x)), "ell: null pointer");
  B200SP_REQUIRE(h, cols > 0 || K == 0, "ell: num_cols == 0 with stored slots");

  b200sp_cfg c = cfg ? *cfg : b200sp_cfg{};
  ell_defaults(c, sizeof(T));

  EllArgs<T> a;
  a.rows = rows; a.cols = cols; a.pitch = pitch; a.K = (int)K;
  a.cidx = cidx; a.vals = vals; a.row_lengths = row_lengths; a.x = x; a.y = y;
  a.accumulate = accumulate; a.dotv = dotv; a.dot_result = dot_result;
  a.dot_partials = reinterpret_cast<T *>(h->red_partials);
  a.dot_ticket = h->red_counters;

  if (c.kernel == B200SP_K_ELL_BULK) {
    const int R = c.block_size * c.unroll;
    const bool ok = (pitch * sizeof(int)) % 16 == 0 && (pitch * sizeof(T)) % 16 == 0 && aligned16(vals) &&
                    aligned16(cidx) && ((size_t)R * sizeof(int)) % 16 == 0 && K > 0;
    if (ok)
      return dispatch_bulk<T>(h, st, a, ceil_div(rows, (i64)R), c.block_size, c.unroll, c.stages, c.ctas_per_sm);
    c.kernel = B200SP_K_ELL_LDG;  // layout not bulk-copyable: same results via LDG
    if (c.unroll > 4) c.unroll = 4;
  }
  if (c.kernel != B200SP_K_ELL_LDG)
    return set_error(h, B200SP_INVALID_INPUT, "ell: unknown kernel id %d", c.kernel);
  return dispatch_ldg<T>(h, st, a, c.block_size, c.unroll);
}

template b200sp_status spmv_ell<float>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                       const float *, const int *, const float *, float *, int,
                                       const b200sp_cfg *, const float *, float *);
template b200sp_status spmv_ell<double>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                        const double *, const int *, const double *, double *, int,
                                        const b200sp_cfg *, const double *, double *);

}  // namespace b200sp

extern "C" {
#define DEF(T, sfx)                                                                              \
  b200sp_status b200sp_spmv_ell_##sfx(b200sp_handle h, b200sp_stream stream, int64_t num_rows,   \
                                      int64_t num_cols, int64_t num_cols_per_row, int64_t pitch, \
                                      const int32_t *column_indices, const T *values, const T *x, \
                                      T *y, int accumulate, const b200sp_cfg *cfg) {             \
    return b200sp::spmv_ell<T>(h, (cudaStream_t)stream, num_rows, num_cols, num_cols_per_row,    \
                               pitch, column_indices, values, nullptr, x, y, accumulate, cfg,    \
                               nullptr, nullptr);                                                \
  }                                                                                              \
  b200sp_status b200sp_spmv_ellr_##sfx(b200sp_handle h, b200sp_stream stream, int64_t num_rows,  \
                                       int64_t num_cols, int64_t num_cols_per_row, int64_t pitch, \
                                       const int32_t *column_indices, const T *values,           \
                                       const int32_t *row_lengths, const T *x, T *y,             \
                                       int accumulate, const b200sp_cfg *cfg) {                  \
    if (h && !row_lengths)                                                                       \
      return b200sp::set_error(h, B200SP_INVALID_INPUT, "ellr: null row_lengths");               \
    return b200sp::spmv_ell<T>(h, (cudaStream_t)stream, num_rows, num_cols, num_cols_per_row,    \
                               pitch, column_indices, values, row_lengths, x, y, accumulate,     \
                               cfg, nullptr, nullptr);                                           \
  }
DEF(float, f32)
DEF(double, f64)
#undef DEF

b200sp_status b200sp_ell_row_lengths(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                     int64_t num_cols_per_row, int64_t pitch,
                                     const int32_t *column_indices, int32_t *row_lengths) {
  B200SP_CHECK_HANDLE(h);
  if (num_rows == 0) return B200SP_OK;
  B200SP_REQUIRE(h, row_lengths && (num_cols_per_row == 0 || column_indices), "ell_row_lengths: null pointer");
  b200sp::ell_row_lengths_kernel<<<(unsigned)b200sp::ceil_div(num_rows, 256), 256, 0, (cudaStream_t)stream>>>(
      num_rows, (int)num_cols_per_row, pitch, column_indices, row_lengths);
  B200SP_LAUNCH_CHECK(h, "ell_row_lengths_kernel");
  return B200SP_OK;
}
}
