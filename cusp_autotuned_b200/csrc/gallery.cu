// gallery.cu — device-side builders for the Poisson test/bench operators.
//
// cusp::gallery::poisson5pt / poisson7pt (cusp/gallery/detail/poisson.inl:29-47,
// 75-96) build a DIA matrix with generate_matrix_from_stencil
// (cusp/gallery/detail/stencil.inl:143-206: row = ix + nx*(iy + ny*iz), one
// diagonal per stencil point, value where the neighbour is inside the grid and 0
// elsewhere) and then cusp::convert it:
//   DIA -> CSR/COO  drops zeros, row-major, diagonals in ascending offset
//                   (generic/conversions/dia_to_other.h:61-161)
//   DIA -> ELL      K = #diagonals, pitch = DIA pitch, col = -1 where the value
//                   is 0, then every row is stably left-packed
//                   (dia_to_other.h:163-251 — two thrust::stable_partition calls
//                   PER ROW, unusable at 1.7e7 rows)
// These kernels write the same arrays directly, for any contiguous block of
// rows (row-partitioned operators), one thread per row.
#include "common.cuh"

namespace b200sp {

struct Grid {
  i64 nx, ny, nz;
  int stencil;  // 5 or 7
};

__host__ __device__ inline int stencil_points(const Grid &g) { return g.stencil; }

// d-th stencil point in the reference's (ascending offset) order
__host__ __device__ inline void stencil_point(const Grid &g, int d, int &dx, int &dy, int &dz) {
  dx = dy = dz = 0;
  if (g.stencil == 7) {
    switch (d) {
      case 0: dz = -1; break;
      case 1: dy = -1; break;
      case 2: dx = -1; break;
      case 3: break;
      case 4: dx = 1; break;
      case 5: dy = 1; break;
      default: dz = 1; break;
    }
  } else {
    switch (d) {
      case 0: dy = -1; break;
      case 1: dx = -1; break;
      case 2: break;
      case 3: dx = 1; break;
      default: dy = 1; break;
    }
  }
}

__host__ __device__ inline i64 stencil_offset(const Grid &g, int d) {
  int dx, dy, dz;
  stencil_point(g, d, dx, dy, dz);
  return (i64)dx + g.nx * ((i64)dy + g.ny * (i64)dz);
}

__host__ __device__ inline bool stencil_inside(const Grid &g, i64 row, int d) {
  int dx, dy, dz;
  stencil_point(g, d, dx, dy, dz);
  const i64 ix = row % g.nx + dx;
  const i64 iy = (row / g.nx) % g.ny + dy;
  const i64 iz = row / (g.nx * g.ny) + dz;
  return ix >= 0 && ix < g.nx && iy >= 0 && iy < g.ny && iz >= 0 && iz < g.nz;
}

template <typename T>
__host__ __device__ inline T stencil_value(const Grid &g, int d) {
  const int centre = g.stencil == 7 ? 3 : 2;
  return d == centre ? T(g.stencil - 1) : T(-1);
}

// number of stored entries in global rows [0, m)
__host__ __device__ inline i64 poisson_prefix(const Grid &g, i64 m) {
  auto first = [](i64 Q, i64 n) { return (Q + n - 1) / n; };  // #{q<Q : q%n==0}
  auto last = [](i64 Q, i64 n) { return Q / n; };             // #{q<Q : q%n==n-1}
  i64 missing = first(m, g.nx) + last(m, g.nx);
  const i64 Q = m / g.nx, rem = m % g.nx;
  missing += g.nx * first(Q, g.ny) + ((Q % g.ny == 0) ? rem : 0);
  missing += g.nx * last(Q, g.ny) + ((Q % g.ny == g.ny - 1) ? rem : 0);
  if (g.stencil == 7) {
    const i64 plane = g.nx * g.ny;
    const i64 P = m / plane, rem2 = m % plane;
    missing += plane * first(P, g.nz) + ((P % g.nz == 0) ? rem2 : 0);
    missing += plane * last(P, g.nz) + ((P % g.nz == g.nz - 1) ? rem2 : 0);
  }
  return (i64)g.stencil * m - missing;
}

template <typename T>
__global__ void poisson_dia_kernel(Grid g, i64 row_begin, i64 nrows, i64 col_shift, i64 pitch, int *offs,
                                   T *vals) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0)
    for (int d = 0; d < g.stencil; ++d) offs[d] = (int)(stencil_offset(g, d) + row_begin - col_shift);
  if (r >= nrows) return;
  for (int d = 0; d < g.stencil; ++d)
    vals[(i64)d * pitch + r] = stencil_inside(g, row_begin + r, d) ? stencil_value<T>(g, d) : T(0);
}

template <typename T>
__global__ void poisson_ell_kernel(Grid g, i64 row_begin, i64 nrows, i64 col_shift, i64 pitch, int *cidx,
                                   T *vals) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const i64 row = row_begin + r;
  int k = 0;
  for (int d = 0; d < g.stencil; ++d)
    if (stencil_inside(g, row, d)) {
      cidx[(i64)k * pitch + r] = (int)(row + stencil_offset(g, d) - col_shift);
      vals[(i64)k * pitch + r] = stencil_value<T>(g, d);
      ++k;
    }
  for (; k < g.stencil; ++k) {
    cidx[(i64)k * pitch + r] = -1;
    vals[(i64)k * pitch + r] = T(0);
  }
}

__global__ void poisson_csr_offsets_kernel(Grid g, i64 row_begin, i64 nrows, int *Ap) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nrows) return;
  Ap[r] = (int)(poisson_prefix(g, row_begin + r) - poisson_prefix(g, row_begin));
}

template <typename T>
__global__ void poisson_csr_kernel(Grid g, i64 row_begin, i64 nrows, i64 col_shift, const int *Ap, int *Aj,
                                   T *Ax) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const i64 row = row_begin + r;
  i64 j = Ap[r];
  for (int d = 0; d < g.stencil; ++d)
    if (stencil_inside(g, row, d)) {
      Aj[j] = (int)(row + stencil_offset(g, d) - col_shift);
      Ax[j] = stencil_value<T>(g, d);
      ++j;
    }
}

static b200sp_status check_grid(b200sp_handle h, int stencil, i64 nx, i64 ny, i64 nz, i64 row_begin,
                                i64 nrows, Grid &g) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, stencil == 5 || stencil == 7, "poisson: stencil must be 5 or 7");
  B200SP_REQUIRE(h, nx > 0 && ny > 0 && nz > 0, "poisson: empty grid");
  B200SP_REQUIRE(h, stencil == 7 || nz == 1, "poisson: 5-point stencil is 2-D (nz must be 1)");
  B200SP_REQUIRE(h, nx * ny * nz < (1ll << 31), "poisson: int32 index range");
  B200SP_REQUIRE(h, row_begin >= 0 && nrows >= 0 && row_begin + nrows <= nx * ny * nz, "poisson: bad row block");
  g.nx = nx; g.ny = ny; g.nz = nz; g.stencil = stencil;
  return B200SP_OK;
}

}  // namespace b200sp

extern "C" {

int64_t b200sp_poisson_num_entries(int stencil, int64_t nx, int64_t ny, int64_t nz, int64_t row_begin,
                                   int64_t num_rows) {
  if (!(stencil == 5 || stencil == 7) || nx <= 0 || ny <= 0 || nz <= 0) return -1;
  b200sp::Grid g{nx, ny, nz, stencil};
  return b200sp::poisson_prefix(g, row_begin + num_rows) - b200sp::poisson_prefix(g, row_begin);
}

b200sp_status b200sp_poisson_csr_offsets(b200sp_handle h, b200sp_stream s, int stencil, int64_t nx,
                                         int64_t ny, int64_t nz, int64_t row_begin, int64_t num_rows,
                                         int32_t *row_offsets) {
  b200sp::Grid g;
  b200sp_status st = b200sp::check_grid(h, stencil, nx, ny, nz, row_begin, num_rows, g);
  if (st != B200SP_OK) return st;
  B200SP_REQUIRE(h, row_offsets, "poisson: null pointer");
  b200sp::poisson_csr_offsets_kernel<<<(unsigned)b200sp::ceil_div(num_rows + 1, 256), 256, 0, (cudaStream_t)s>>>(
      g, row_begin, num_rows, row_offsets);
  B200SP_LAUNCH_CHECK(h, "poisson_csr_offsets_kernel");
  return B200SP_OK;
}

#define DEF(T, sfx)                                                                                 \
  b200sp_status b200sp_poisson_dia_##sfx(b200sp_handle h, b200sp_stream s, int stencil, int64_t nx, \
                                         int64_t ny, int64_t nz, int64_t row_begin,                 \
                                         int64_t num_rows, int64_t col_shift, int64_t pitch,        \
                                         int32_t *diagonal_offsets, T *values) {                    \
    b200sp::Grid g;                                                                                 \
    b200sp_status st = b200sp::check_grid(h, stencil, nx, ny, nz, row_begin, num_rows, g);          \
    if (st != B200SP_OK) return st;                                                                 \
    B200SP_REQUIRE(h, diagonal_offsets && values && pitch >= num_rows, "poisson_dia: bad arguments"); \
    b200sp::poisson_dia_kernel<T><<<(unsigned)b200sp::ceil_div(num_rows > 0 ? num_rows : 1, 256), 256, 0, \
                                    (cudaStream_t)s>>>(g, row_begin, num_rows, col_shift, pitch,    \
                                                       diagonal_offsets, values);                   \
    B200SP_LAUNCH_CHECK(h, "poisson_dia_kernel");                                                   \
    return B200SP_OK;                                                                               \
  }                                                                                                 \
  b200sp_status b200sp_poisson_ell_##sfx(b200sp_handle h, b200sp_stream s, int stencil, int64_t nx, \
                                         int64_t ny, int64_t nz, int64_t row_begin,                 \
                                         int64_t num_rows, int64_t col_shift, int64_t pitch,        \
                                         int32_t *column_indices, T *values) {                      \
    b200sp::Grid g;                                                                                 \
    b200sp_status st = b200sp::check_grid(h, stencil, nx, ny, nz, row_begin, num_rows, g);          \
    if (st != B200SP_OK) return st;                                                                 \
    if (num_rows == 0) return B200SP_OK;                                                            \
    B200SP_REQUIRE(h, column_indices && values && pitch >= num_rows, "poisson_ell: bad arguments"); \
    b200sp::poisson_ell_kernel<T><<<(unsigned)b200sp::ceil_div(num_rows, 256), 256, 0,              \
                                    (cudaStream_t)s>>>(g, row_begin, num_rows, col_shift, pitch,    \
                                                       column_indices, values);                     \
    B200SP_LAUNCH_CHECK(h, "poisson_ell_kernel");                                                   \
    return B200SP_OK;                                                                               \
  }                                                                                                 \
  b200sp_status b200sp_poisson_csr_##sfx(b200sp_handle h, b200sp_stream s, int stencil, int64_t nx, \
                                         int64_t ny, int64_t nz, int64_t row_begin,                 \
                                         int64_t num_rows, int64_t col_shift,                       \
                                         const int32_t *row_offsets, int32_t *column_indices,       \
                                         T *values) {                                               \
    b200sp::Grid g;                                                                                 \
    b200sp_status st = b200sp::check_grid(h, stencil, nx, ny, nz, row_begin, num_rows, g);          \
    if (st != B200SP_OK) return st;                                                                 \
    if (num_rows == 0) return B200SP_OK;                                                            \
    B200SP_REQUIRE(h, row_offsets && column_indices && values, "poisson_csr: null pointer");        \
    b200sp::poisson_csr_kernel<T><<<(unsigned)b200sp::ceil_div(num_rows, 256), 256, 0,              \
                                    (cudaStream_t)s>>>(g, row_begin, num_rows, col_shift,           \
                                                       row_offsets, column_indices, values);        \
    B200SP_LAUNCH_CHECK(h, "poisson_csr_kernel");                                                   \
    return B200SP_OK;                                                                               \
  }
DEF(float, f32)
DEF(double, f64)
#undef DEF
}
