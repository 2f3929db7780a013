// cusp/convert.h — cusp::convert(src, dst) and cusp::copy(src, dst) between
// formats and memory spaces, following the reference's rules bit for bit
// (reference: cusp/system/detail/generic/convert.inl:40-110 and
// generic/conversions/{coo,csr,dia,ell,hyb}_to_other.h):
//   COO -> CSR   row_indices must be sorted; indices_to_offsets; entry order kept
//   CSR -> COO   offsets_to_indices
//   DIA -> COO/CSR  row-major scan, entries with value == 0 dropped,
//                   diagonals in ascending offset          (dia_to_other.h:61-161)
//   DIA -> ELL   K = #diagonals, pitch = DIA pitch, zero slots become padding,
//                every row left-packed stably              (dia_to_other.h:163-251)
//   ELL -> COO/CSR  row-major scan, value == 0 dropped     (ell_to_other.h:52-140)
//   CSR -> ELL   k-th entry of row i -> slot (i,k); pad col -1 / val 0;
//                pitch = round_up(rows, 32); K = max row length; refuses a
//                fill-in > 3x when the slab exceeds 1e6 slots
//                num_entries = nnz - count(values == 0)   (csr_to_other.h:155-227)
//   CSR -> DIA   occupied diagonals ascending, pitch = round_up(rows, 32),
//                same fill-in guard                        (csr_to_other.h:73-153)
//   CSR -> HYB   K = compute_optimal_entries_per_row(relative_speed 3,
//                breakeven 4096); first K entries of a row to ELL, the rest to
//                COO in CSR order   (csr_to_other.h:229-306, format_utils.inl:281-321,
//                                    detail/functional.inl:114-132)
//   anything else goes through COO/CSR                     (convert.inl:53-70)
// Conversions are setup-time operations.  Device containers with 32-bit indices and float /
// double values convert on the device through the engine's conversion kernels, every source
// x destination pair (b200sp_{dia,ell,hyb}_to_csr_offsets/_fill, b200sp_dia_to_ell,
// b200sp_csr_to_{ell,coo_tail,dia}, b200sp_{offsets_to_indices,indices_to_offsets},
// b200sp_csr_convert_query — csrc/convert.cu), same layouts bit for bit; host
// containers, other index / value types and dense (array2d) ends run on the host.  The device
// builders for the benchmark operators are in cusp/gallery/poisson.h.
#pragma once
#include <algorithm>
#include <vector>

#include "coo_matrix.h"
#include "csr_matrix.h"
#include "detail/descriptor.h"
#include "dia_matrix.h"
#include "ell_matrix.h"
#include "hyb_matrix.h"

namespace cusp {
namespace detail {

// host CSR image: the hub every conversion goes through
template <typename I, typename V>
struct host_csr {
  size_t rows = 0, cols = 0;
  std::vector<I> offsets, columns;
  std::vector<V> values;
  size_t nnz() const { return columns.size(); }
};

template <typename I, typename V, typename M>
void gather(const M &A, host_csr<I, V> &H, coo_format) {
  H.rows = A.num_rows;
  H.cols = A.num_cols;
  auto ri = to_host_vector(A.row_indices);
  auto ci = to_host_vector(A.column_indices);
  auto vv = to_host_vector(A.values);
  const size_t n = A.num_entries;
  H.offsets.assign(H.rows + 1, 0);
  size_t k = 0;  // indices_to_offsets: offsets[i] = #indices < i
  for (size_t i = 0; i <= H.rows; ++i) {
    while (k < n && (size_t)ri[k] < i) ++k;
    H.offsets[i] = (I)k;
  }
  H.columns.assign(ci.begin(), ci.begin() + n);
  H.values.assign(vv.begin(), vv.begin() + n);
}
template <typename I, typename V, typename M>
void gather(const M &A, host_csr<I, V> &H, csr_format) {
  H.rows = A.num_rows;
  H.cols = A.num_cols;
  auto ro = to_host_vector(A.row_offsets);
  auto ci = to_host_vector(A.column_indices);
  auto vv = to_host_vector(A.values);
  H.offsets.assign(ro.begin(), ro.end());
  if (H.offsets.size() != H.rows + 1) H.offsets.resize(H.rows + 1, H.offsets.empty() ? 0 : H.offsets.back());
  H.columns.assign(ci.begin(), ci.begin() + A.num_entries);
  H.values.assign(vv.begin(), vv.begin() + A.num_entries);
}
template <typename I, typename V, typename M>
void gather(const M &A, host_csr<I, V> &H, dia_format) {
  H.rows = A.num_rows;
  H.cols = A.num_cols;
  auto offs = to_host_vector(A.diagonal_offsets);
  auto vv = to_host_vector(A.values.values);
  const size_t pitch = A.values.pitch, nd = offs.size();
  H.offsets.assign(H.rows + 1, 0);
  for (size_t i = 0; i < H.rows; ++i) {
    for (size_t d = 0; d < nd; ++d) {
      const auto v = vv[d * pitch + i];
      if (v != 0) {
        H.columns.push_back((I)((long long)i + offs[d]));
        H.values.push_back((V)v);
      }
    }
    H.offsets[i + 1] = (I)H.columns.size();
  }
}
template <typename I, typename V, typename M>
void gather(const M &A, host_csr<I, V> &H, ell_format) {
  H.rows = A.num_rows;
  H.cols = A.num_cols;
  auto ci = to_host_vector(A.column_indices.values);
  auto vv = to_host_vector(A.values.values);
  const size_t pitch = A.column_indices.pitch, K = A.column_indices.num_cols;
  H.offsets.assign(H.rows + 1, 0);
  for (size_t i = 0; i < H.rows; ++i) {
    for (size_t k = 0; k < K; ++k) {
      const auto v = vv[k * pitch + i];
      if (v != 0) {
        H.columns.push_back((I)ci[k * pitch + i]);
        H.values.push_back((V)v);
      }
    }
    H.offsets[i + 1] = (I)H.columns.size();
  }
}
// HYB (hyb_to_other.h:45-56 + cusp/detail/coo_matrix.inl:269-341): the ELL slots whose COLUMN is valid, merged by
// (row, column) with the COO entries, ties ELL first
template <typename I, typename V, typename M>
void gather(const M &A, host_csr<I, V> &H, hyb_format) {
  host_csr<I, V> C;
  gather(A.coo, C, coo_format());
  auto ec = to_host_vector(A.ell.column_indices.values);
  auto ev = to_host_vector(A.ell.values.values);
  const size_t pitch = A.ell.column_indices.pitch, K = A.ell.column_indices.num_cols;
  H.rows = A.num_rows;
  H.cols = A.num_cols;
  H.offsets.assign(H.rows + 1, 0);
  for (size_t i = 0; i < H.rows; ++i) {
    I j = C.offsets[i];
    const I jend = C.offsets[i + 1];
    for (size_t k = 0; k < K; ++k) {
      const I c = (I)ec[k * pitch + i];
      for (; j < jend && C.columns[j] < c; ++j) {
        H.columns.push_back(C.columns[j]);
        H.values.push_back(C.values[j]);
      }
      if (c != (I)-1) {
        H.columns.push_back(c);
        H.values.push_back((V)ev[k * pitch + i]);
      }
    }
    for (; j < jend; ++j) {
      H.columns.push_back(C.columns[j]);
      H.values.push_back(C.values[j]);
    }
    H.offsets[i + 1] = (I)H.columns.size();
  }
}
template <typename I, typename V, typename M>
void gather(const M &A, host_csr<I, V> &H, array2d_format) {
  H.rows = A.num_rows;
  H.cols = A.num_cols;
  auto vv = to_host_vector(A.values);
  typedef orient<typename M::orientation> O;
  H.offsets.assign(H.rows + 1, 0);
  for (size_t i = 0; i < H.rows; ++i) {
    for (size_t j = 0; j < H.cols; ++j) {
      const auto v = vv[O::index(i, j, A.pitch)];
      if (v != 0) {
        H.columns.push_back((I)j);
        H.values.push_back((V)v);
      }
    }
    H.offsets[i + 1] = (I)H.columns.size();
  }
}

template <typename T, typename Array>
void upload(const std::vector<T> &h, Array &dst) {
  dst.resize(h.size());
  raw_copy<T, host_memory, typename Array::memory_space>(h.data(), raw_ptr(dst), h.size());
}

template <typename I>
size_t max_entries_per_row(const std::vector<I> &offsets) {
  size_t m = 0;
  for (size_t i = 0; i + 1 < offsets.size(); ++i) m = std::max<size_t>(m, offsets[i + 1] - offsets[i]);
  return m;
}

// cusp::compute_optimal_entries_per_row (generic/format_utils.inl:281-321) with
// speed_threshold_functor (detail/functional.inl:114-132)
template <typename I>
size_t optimal_entries_per_row(const std::vector<I> &offsets, float relative_speed = 3.0f,
                               size_t breakeven_threshold = 4096) {
  const size_t rows = offsets.size() - 1;
  const size_t maxc = max_entries_per_row(offsets);
  std::vector<size_t> cum(maxc + 1, 0);  // cum[k] = #rows with length <= k
  for (size_t i = 0; i < rows; ++i) cum[offsets[i + 1] - offsets[i]]++;
  for (size_t k = 1; k <= maxc; ++k) cum[k] += cum[k - 1];
  for (size_t k = 0; k < maxc; ++k) {
    const size_t longer = rows - cum[k];
    if (relative_speed * (float)longer < (float)rows || longer < breakeven_threshold) return k;
  }
  return maxc;
}

inline void fill_guard(size_t slots_per_row, size_t rows, size_t nnz, const char *what) {
  const float size = float(slots_per_row) * float(rows);
  const float fill_ratio = size / std::max(1.0f, float(nnz));
  if (3.0f < fill_ratio && size > 1e6f) throw cusp::format_conversion_exception(what);
}

template <typename I, typename V, typename M>
void scatter(const host_csr<I, V> &H, M &B, csr_format) {
  B.resize(H.rows, H.cols, H.nnz());
  upload(H.offsets, B.row_offsets);
  upload(H.columns, B.column_indices);
  upload(H.values, B.values);
}
template <typename I, typename V, typename M>
void scatter(const host_csr<I, V> &H, M &B, coo_format) {
  std::vector<I> ri(H.nnz());
  for (size_t i = 0; i < H.rows; ++i)
    for (I k = H.offsets[i]; k < H.offsets[i + 1]; ++k) ri[k] = (I)i;
  B.resize(H.rows, H.cols, H.nnz());
  upload(ri, B.row_indices);
  upload(H.columns, B.column_indices);
  upload(H.values, B.values);
}
template <typename I, typename V>
void ell_slabs(const host_csr<I, V> &H, size_t K, size_t pitch, std::vector<I> &ci, std::vector<V> &vv,
               std::vector<I> *coo_r, std::vector<I> *coo_c, std::vector<V> *coo_v) {
  ci.assign(K * pitch, (I)-1);
  vv.assign(K * pitch, V(0));
  for (size_t i = 0; i < H.rows; ++i)
    for (I jj = H.offsets[i]; jj < H.offsets[i + 1]; ++jj) {
      const size_t k = (size_t)(jj - H.offsets[i]);
      if (k < K) {
        ci[k * pitch + i] = H.columns[jj];
        vv[k * pitch + i] = H.values[jj];
      } else if (coo_r) {
        coo_r->push_back((I)i);
        coo_c->push_back(H.columns[jj]);
        coo_v->push_back(H.values[jj]);
      }
    }
}
template <typename I, typename V, typename M>
void scatter(const host_csr<I, V> &H, M &B, ell_format, size_t K = 0, size_t alignment = 32) {
  if (H.nnz() == 0) {
    B.resize(H.rows, H.cols, 0, K);
    return;
  }
  if (K == 0) {
    K = max_entries_per_row(H.offsets);
    fill_guard(K, H.rows, H.nnz(), "ell_matrix fill-in would exceed maximum tolerance");
  }
  const size_t zeros = (size_t)std::count(H.values.begin(), H.values.end(), V(0));
  B.resize(H.rows, H.cols, H.nnz() - zeros, K, alignment);
  std::vector<I> ci;
  std::vector<V> vv;
  ell_slabs<I, V>(H, K, B.column_indices.pitch, ci, vv, nullptr, nullptr, nullptr);
  upload(ci, B.column_indices.values);
  upload(vv, B.values.values);
}
template <typename I, typename V, typename M>
void scatter(const host_csr<I, V> &H, M &B, hyb_format, size_t K = (size_t)-1, size_t alignment = 32) {
  if (H.nnz() == 0) {
    B.resize(H.rows, H.cols, 0, 0, K == (size_t)-1 ? 0 : K);
    return;
  }
  if (K == (size_t)-1) K = optimal_entries_per_row(H.offsets);
  std::vector<I> ci, cr, cc;
  std::vector<V> vv, cv;
  const size_t pitch = round_up(H.rows, alignment);
  ell_slabs<I, V>(H, K, pitch, ci, vv, &cr, &cc, &cv);
  B.resize(H.rows, H.cols, H.nnz() - cr.size(), cr.size(), K, alignment);
  upload(ci, B.ell.column_indices.values);
  upload(vv, B.ell.values.values);
  upload(cr, B.coo.row_indices);
  upload(cc, B.coo.column_indices);
  upload(cv, B.coo.values);
}
template <typename I, typename V, typename M>
void scatter(const host_csr<I, V> &H, M &B, dia_format, size_t alignment = 32) {
  std::vector<char> occ(H.rows + H.cols, 0);
  for (size_t i = 0; i < H.rows; ++i)
    for (I jj = H.offsets[i]; jj < H.offsets[i + 1]; ++jj) occ[(size_t)((long long)H.columns[jj] - (long long)i + (long long)H.rows)] = 1;
  std::vector<long long> map(H.rows + H.cols, -1);
  std::vector<I> offs;
  for (size_t k = 0; k < occ.size(); ++k)
    if (occ[k]) {
      map[k] = (long long)offs.size();
      offs.push_back((I)((long long)k - (long long)H.rows));
    }
  fill_guard(offs.size(), H.rows, H.nnz(), "dia_matrix fill-in would exceed maximum tolerance");
  B.resize(H.rows, H.cols, H.nnz(), offs.size(), alignment);
  const size_t pitch = B.values.pitch;
  std::vector<V> vv(pitch * offs.size(), V(0));
  for (size_t i = 0; i < H.rows; ++i)
    for (I jj = H.offsets[i]; jj < H.offsets[i + 1]; ++jj)
      vv[(size_t)map[(size_t)((long long)H.columns[jj] - (long long)i + (long long)H.rows)] * pitch + i] = H.values[jj];
  upload(offs, B.diagonal_offsets);
  upload(vv, B.values.values);
}
template <typename I, typename V, typename M>
void scatter(const host_csr<I, V> &H, M &B, array2d_format) {
  B.resize(H.rows, H.cols);
  typedef orient<typename M::orientation> O;
  std::vector<typename M::value_type> vv(B.values.size(), 0);
  for (size_t i = 0; i < H.rows; ++i)
    for (I jj = H.offsets[i]; jj < H.offsets[i + 1]; ++jj) vv[O::index(i, (size_t)H.columns[jj], B.pitch)] += H.values[jj];
  upload(vv, B.values);
}

// same format: element-wise copy of every array (cusp::copy)
template <typename S, typename D>
void copy_same(const S &src, D &dst, coo_format) {
  dst.resize(src.num_rows, src.num_cols, src.num_entries);
  dst.row_indices = src.row_indices;
  dst.column_indices = src.column_indices;
  dst.values = src.values;
}
template <typename S, typename D>
void copy_same(const S &src, D &dst, csr_format) {
  dst.resize(src.num_rows, src.num_cols, src.num_entries);
  dst.row_offsets = src.row_offsets;
  dst.column_indices = src.column_indices;
  dst.values = src.values;
}
template <typename S, typename D>
void copy_same(const S &src, D &dst, dia_format) {
  dst.resize(src.num_rows, src.num_cols, src.num_entries, src.diagonal_offsets.size());
  dst.diagonal_offsets = src.diagonal_offsets;
  dst.values.resize(src.values.num_rows, src.values.num_cols, src.values.pitch);
  dst.values.values = src.values.values;
}
template <typename S, typename D>
void copy_same(const S &src, D &dst, ell_format) {
  dst.resize(src.num_rows, src.num_cols, src.num_entries, src.column_indices.num_cols);
  dst.column_indices.resize(src.column_indices.num_rows, src.column_indices.num_cols, src.column_indices.pitch);
  dst.column_indices.values = src.column_indices.values;
  dst.values.resize(src.values.num_rows, src.values.num_cols, src.values.pitch);
  dst.values.values = src.values.values;
}
template <typename S, typename D>
void copy_same(const S &src, D &dst, hyb_format) {
  copy_same(src.ell, dst.ell, ell_format());
  copy_same(src.coo, dst.coo, coo_format());
  dst.num_rows = src.num_rows;
  dst.num_cols = src.num_cols;
  dst.num_entries = src.num_entries;
}
template <typename S, typename D>
void copy_same(const S &src, D &dst, array2d_format) {
  dst.resize(src.num_rows, src.num_cols, src.pitch);
  dst.values = src.values;
}
template <typename S, typename D>
void copy_same(const S &src, D &dst, array1d_format) {
  dst = src;
}

template <typename S, typename D, typename F>
void convert_dispatch(const S &src, D &dst, F, F) {
  copy_same(src, dst, F());
}
// DIA -> ELL keeps the DIA pitch and K = #diagonals (dia_to_other.h:163-251)
template <typename S, typename D>
void convert_dispatch(const S &src, D &dst, dia_format, ell_format) {
  typedef typename D::index_type I;
  typedef typename D::value_type V;
  auto offs = to_host_vector(src.diagonal_offsets);
  auto vv = to_host_vector(src.values.values);
  const size_t pitch = src.values.pitch, nd = offs.size(), rows = src.num_rows;
  std::vector<I> ci(nd * pitch, (I)-1);
  std::vector<V> ev(nd * pitch, V(0));
  for (size_t i = 0; i < rows; ++i) {
    size_t k = 0;
    for (size_t d = 0; d < nd; ++d) {
      const auto v = vv[d * pitch + i];
      if (v != 0) {
        ci[k * pitch + i] = (I)((long long)i + offs[d]);
        ev[k * pitch + i] = (V)v;
        ++k;
      }
    }
  }
  dst.resize(src.num_rows, src.num_cols, src.num_entries, nd);
  dst.column_indices.resize(rows, nd, pitch);
  dst.values.resize(rows, nd, pitch);
  upload(ci, dst.column_indices.values);
  upload(ev, dst.values.values);
}
// ---- device -> device through the conversion kernels ---------------------------------
inline b200sp_status csr_to_ell_(int64_t r, int64_t K, int64_t p, const int *ro, const int *ci, const float *v, int *ec,
                                 float *ev) {
  return b200sp_csr_to_ell_f32(engine(), current_stream(), r, K, p, ro, ci, v, ec, ev);
}
inline b200sp_status csr_to_ell_(int64_t r, int64_t K, int64_t p, const int *ro, const int *ci, const double *v,
                                 int *ec, double *ev) {
  return b200sp_csr_to_ell_f64(engine(), current_stream(), r, K, p, ro, ci, v, ec, ev);
}
inline b200sp_status csr_to_tail_(int64_t r, int64_t K, const int *ro, const int *ci, const float *v, int *tr, int *tc,
                                  float *tv) {
  return b200sp_csr_to_coo_tail_f32(engine(), current_stream(), r, K, ro, ci, v, tr, tc, tv);
}
inline b200sp_status csr_to_tail_(int64_t r, int64_t K, const int *ro, const int *ci, const double *v, int *tr,
                                  int *tc, double *tv) {
  return b200sp_csr_to_coo_tail_f64(engine(), current_stream(), r, K, ro, ci, v, tr, tc, tv);
}
inline b200sp_status csr_to_dia_(int64_t r, int64_t c, int64_t nd, int64_t p, const int *ro, const int *ci,
                                 const float *v, int *off, float *dv) {
  return b200sp_csr_to_dia_f32(engine(), current_stream(), r, c, nd, p, ro, ci, v, off, dv);
}
inline b200sp_status csr_to_dia_(int64_t r, int64_t c, int64_t nd, int64_t p, const int *ro, const int *ci,
                                 const double *v, int *off, double *dv) {
  return b200sp_csr_to_dia_f64(engine(), current_stream(), r, c, nd, p, ro, ci, v, off, dv);
}
inline b200sp_status count_zeros_(int64_t n, const float *v, int64_t *c) {
  return b200sp_count_zeros_f32(engine(), current_stream(), n, v, c);
}
inline b200sp_status count_zeros_(int64_t n, const double *v, int64_t *c) {
  return b200sp_count_zeros_f64(engine(), current_stream(), n, v, c);
}

// both ends live on the device, same value type, the ABI's index / value types
template <typename S, typename D>
struct device_convertible
    : std::integral_constant<bool, abi_matrix<S>::value && abi_matrix<D>::value &&
                                       std::is_same<typename S::value_type, typename D::value_type>::value> {};

template <typename S>
b200sp_convert_info csr_query(const S &src, bool diagonals) {
  b200sp_convert_info info;
  check(b200sp_csr_convert_query(engine(), current_stream(), (int64_t)src.num_rows, (int64_t)src.num_cols,
                                 (int64_t)src.num_entries, raw_ptr(src.row_offsets),
                                 diagonals ? raw_ptr(src.column_indices) : nullptr, 3.0f, 4096, &info));
  return info;
}

template <typename S, typename D>
void device_convert(const S &src, D &dst, csr_format, coo_format) {
  dst.resize(src.num_rows, src.num_cols, src.num_entries);
  check(b200sp_offsets_to_indices(engine(), current_stream(), (int64_t)src.num_rows, raw_ptr(src.row_offsets),
                                  raw_ptr(dst.row_indices)));
  dst.column_indices = src.column_indices;
  dst.values = src.values;
}
template <typename S, typename D>
void device_convert(const S &src, D &dst, coo_format, csr_format) {
  dst.resize(src.num_rows, src.num_cols, src.num_entries);
  check(b200sp_indices_to_offsets(engine(), current_stream(), (int64_t)src.num_rows, (int64_t)src.num_entries,
                                  raw_ptr(src.row_indices), raw_ptr(dst.row_offsets)));
  dst.column_indices = src.column_indices;
  dst.values = src.values;
}
template <typename S, typename D>
void device_convert(const S &src, D &dst, csr_format, ell_format, size_t K = 0, size_t alignment = 32) {
  if (src.num_entries == 0) {
    dst.resize(src.num_rows, src.num_cols, 0, K);
    return;
  }
  if (K == 0) {
    K = (size_t)csr_query(src, false).max_entries_per_row;
    fill_guard(K, src.num_rows, src.num_entries, "ell_matrix fill-in would exceed maximum tolerance");
  }
  int64_t zeros = 0;
  check(count_zeros_((int64_t)src.num_entries, raw_ptr(src.values), &zeros));
  dst.resize(src.num_rows, src.num_cols, src.num_entries - (size_t)zeros, K, alignment);
  check(csr_to_ell_((int64_t)src.num_rows, (int64_t)K, (int64_t)dst.column_indices.pitch, raw_ptr(src.row_offsets),
                    raw_ptr(src.column_indices), raw_ptr(src.values), raw_ptr(dst.column_indices.values),
                    raw_ptr(dst.values.values)));
}
template <typename S, typename D>
void device_convert(const S &src, D &dst, csr_format, hyb_format, size_t K = (size_t)-1, size_t alignment = 32) {
  if (src.num_entries == 0) {
    dst.resize(src.num_rows, src.num_cols, 0, 0, K == (size_t)-1 ? 0 : K);
    return;
  }
  size_t tail;
  if (K == (size_t)-1) {
    const b200sp_convert_info info = csr_query(src, false);
    K = (size_t)info.hyb_entries_per_row;
    tail = (size_t)info.hyb_coo_entries;
  } else {  // explicit width: count the tail on the host side of the offsets (set-up time)
    auto ro = to_host_vector(src.row_offsets);
    tail = 0;
    for (size_t i = 0; i + 1 < ro.size(); ++i) tail += (size_t)std::max<long long>((long long)ro[i + 1] - ro[i] - (long long)K, 0);
  }
  dst.resize(src.num_rows, src.num_cols, src.num_entries - tail, tail, K, alignment);
  check(csr_to_ell_((int64_t)src.num_rows, (int64_t)K, (int64_t)dst.ell.column_indices.pitch, raw_ptr(src.row_offsets),
                    raw_ptr(src.column_indices), raw_ptr(src.values), raw_ptr(dst.ell.column_indices.values),
                    raw_ptr(dst.ell.values.values)));
  check(csr_to_tail_((int64_t)src.num_rows, (int64_t)K, raw_ptr(src.row_offsets), raw_ptr(src.column_indices),
                     raw_ptr(src.values), raw_ptr(dst.coo.row_indices), raw_ptr(dst.coo.column_indices),
                     raw_ptr(dst.coo.values)));
}
template <typename S, typename D>
void device_convert(const S &src, D &dst, csr_format, dia_format, size_t alignment = 32) {
  if (src.num_entries == 0) {
    dst.resize(src.num_rows, src.num_cols, 0, 0);
    return;
  }
  const size_t nd = (size_t)csr_query(src, true).num_diagonals;
  fill_guard(nd, src.num_rows, src.num_entries, "dia_matrix fill-in would exceed maximum tolerance");
  dst.resize(src.num_rows, src.num_cols, src.num_entries, nd, alignment);
  check(csr_to_dia_((int64_t)src.num_rows, (int64_t)src.num_cols, (int64_t)nd, (int64_t)dst.values.pitch,
                    raw_ptr(src.row_offsets), raw_ptr(src.column_indices), raw_ptr(src.values),
                    raw_ptr(dst.diagonal_offsets), raw_ptr(dst.values.values)));
}
// COO source: to CSR on the device, then as above
template <typename S, typename D, typename F2>
void device_convert(const S &src, D &dst, coo_format, F2) {
  cusp::csr_matrix<typename S::index_type, typename S::value_type, device_memory> csr;
  device_convert(src, csr, coo_format(), csr_format());
  device_convert(csr, dst, csr_format(), F2());
}

// ---- DIA / ELL / HYB sources: to CSR on the device (b200sp_{dia,ell,hyb}_to_csr_offsets / _fill:
// generic/conversions/{dia,ell,hyb}_to_other.h), then on to the destination like any CSR ----
#define CUSP_B200_TO_CSR_WRAPPERS(T, sfx)                                                                              \
  inline b200sp_status dia_offs_(int64_t r, int64_t nd, int64_t p, const T *v, int *ro, int64_t *n) {                  \
    return b200sp_dia_to_csr_offsets_##sfx(engine(), current_stream(), r, nd, p, v, ro, n);                            \
  }                                                                                                                    \
  inline b200sp_status dia_fill_(int64_t r, int64_t nd, int64_t p, const int *off, const T *v, const int *ro, int *cj, \
                                 T *cv) {                                                                              \
    return b200sp_dia_to_csr_fill_##sfx(engine(), current_stream(), r, nd, p, off, v, ro, cj, cv);                     \
  }                                                                                                                    \
  inline b200sp_status ell_offs_(int64_t r, int64_t K, int64_t p, const int *ec, const T *ev, int *ro, int64_t *n) {   \
    return b200sp_ell_to_csr_offsets_##sfx(engine(), current_stream(), r, K, p, ec, ev, ro, n);                        \
  }                                                                                                                    \
  inline b200sp_status ell_fill_(int64_t r, int64_t K, int64_t p, const int *ec, const T *ev, const int *ro, int *cj,  \
                                 T *cv) {                                                                              \
    return b200sp_ell_to_csr_fill_##sfx(engine(), current_stream(), r, K, p, ec, ev, ro, cj, cv);                      \
  }                                                                                                                    \
  inline b200sp_status hyb_offs_(int64_t r, int64_t K, int64_t p, const int *ec, const T *ev, int64_t cn,              \
                                 const int *ci, int *ro, int64_t *n) {                                                 \
    return b200sp_hyb_to_csr_offsets_##sfx(engine(), current_stream(), r, K, p, ec, ev, cn, ci, ro, n);                \
  }                                                                                                                    \
  inline b200sp_status hyb_fill_(int64_t r, int64_t K, int64_t p, const int *ec, const T *ev, int64_t cn,              \
                                 const int *ci, const int *cc, const T *cvv, const int *ro, int *cj, T *cv) {          \
    return b200sp_hyb_to_csr_fill_##sfx(engine(), current_stream(), r, K, p, ec, ev, cn, ci, cc, cvv, ro, cj, cv);     \
  }                                                                                                                    \
  inline b200sp_status dia_to_ell_(int64_t r, int64_t nd, int64_t p, const int *off, const T *v, int *ec, T *ev) {     \
    return b200sp_dia_to_ell_##sfx(engine(), current_stream(), r, nd, p, off, v, ec, ev);                              \
  }
CUSP_B200_TO_CSR_WRAPPERS(float, f32)
CUSP_B200_TO_CSR_WRAPPERS(double, f64)
#undef CUSP_B200_TO_CSR_WRAPPERS

template <typename S, typename D>
void device_convert(const S &src, D &dst, dia_format, csr_format) {
  dst.resize(src.num_rows, src.num_cols, 0);
  int64_t n = 0;
  check(dia_offs_((int64_t)src.num_rows, (int64_t)src.values.num_cols, (int64_t)src.values.pitch,
                  raw_ptr(src.values.values), raw_ptr(dst.row_offsets), &n));
  dst.column_indices.resize((size_t)n);  // row_offsets stays as computed
  dst.values.resize((size_t)n);
  dst.num_entries = (size_t)n;
  check(dia_fill_((int64_t)src.num_rows, (int64_t)src.values.num_cols, (int64_t)src.values.pitch,
                  raw_ptr(src.diagonal_offsets), raw_ptr(src.values.values), raw_ptr(dst.row_offsets),
                  raw_ptr(dst.column_indices), raw_ptr(dst.values)));
}
template <typename S, typename D>
void device_convert(const S &src, D &dst, ell_format, csr_format) {
  dst.resize(src.num_rows, src.num_cols, 0);
  const int64_t K = (int64_t)src.column_indices.num_cols, p = (int64_t)src.column_indices.pitch;
  int64_t n = 0;
  check(ell_offs_((int64_t)src.num_rows, K, p, raw_ptr(src.column_indices.values), raw_ptr(src.values.values),
                  raw_ptr(dst.row_offsets), &n));
  dst.column_indices.resize((size_t)n);  // row_offsets stays as computed
  dst.values.resize((size_t)n);
  dst.num_entries = (size_t)n;
  check(ell_fill_((int64_t)src.num_rows, K, p, raw_ptr(src.column_indices.values), raw_ptr(src.values.values),
                  raw_ptr(dst.row_offsets), raw_ptr(dst.column_indices), raw_ptr(dst.values)));
}
template <typename S, typename D>
void device_convert(const S &src, D &dst, hyb_format, csr_format) {
  dst.resize(src.num_rows, src.num_cols, 0);
  const int64_t K = (int64_t)src.ell.column_indices.num_cols, p = (int64_t)src.ell.column_indices.pitch;
  const int64_t cn = (int64_t)src.coo.num_entries;
  int64_t n = 0;
  check(hyb_offs_((int64_t)src.num_rows, K, p, raw_ptr(src.ell.column_indices.values), raw_ptr(src.ell.values.values), cn,
                  raw_ptr(src.coo.row_indices), raw_ptr(dst.row_offsets), &n));
  dst.column_indices.resize((size_t)n);  // row_offsets stays as computed
  dst.values.resize((size_t)n);
  dst.num_entries = (size_t)n;
  check(hyb_fill_((int64_t)src.num_rows, K, p, raw_ptr(src.ell.column_indices.values), raw_ptr(src.ell.values.values), cn,
                  raw_ptr(src.coo.row_indices), raw_ptr(src.coo.column_indices), raw_ptr(src.coo.values),
                  raw_ptr(dst.row_offsets), raw_ptr(dst.column_indices), raw_ptr(dst.values)));
}
// DIA -> ELL: the fork's direct rule (dia_to_other.h:163-251): K = #diagonals, pitch = the DIA pitch
template <typename S, typename D>
void device_convert(const S &src, D &dst, dia_format, ell_format) {
  dst.resize(src.num_rows, src.num_cols, src.num_entries, src.values.num_cols, 1);
  if (dst.column_indices.pitch != src.values.pitch) {  // keep the source's pitch like the reference does
    dst.column_indices.resize(src.num_rows, src.values.num_cols, src.values.pitch);
    dst.values.resize(src.num_rows, src.values.num_cols, src.values.pitch);
  }
  check(dia_to_ell_((int64_t)src.num_rows, (int64_t)src.values.num_cols, (int64_t)src.values.pitch,
                    raw_ptr(src.diagonal_offsets), raw_ptr(src.values.values), raw_ptr(dst.column_indices.values),
                    raw_ptr(dst.values.values)));
}
// ELL -> HYB: the ELL part is the matrix itself, the COO part is empty (ell_to_other.h:145-163)
template <typename S, typename D>
void device_convert(const S &src, D &dst, ell_format, hyb_format) {
  dst.resize(src.num_rows, src.num_cols, src.num_entries, 0, src.column_indices.num_cols);
  dst.ell = src;
}
// every other pair with a slab source: through CSR
template <typename S, typename D, typename F1, typename F2>
void device_convert_via_csr(const S &src, D &dst, F1, F2) {
  cusp::csr_matrix<typename S::index_type, typename S::value_type, device_memory> csr;
  device_convert(src, csr, F1(), csr_format());
  device_convert(csr, dst, csr_format(), F2());
}
template <typename S, typename D>
void device_convert(const S &src, D &dst, dia_format, coo_format) { device_convert_via_csr(src, dst, dia_format(), coo_format()); }
template <typename S, typename D>
void device_convert(const S &src, D &dst, dia_format, hyb_format) { device_convert_via_csr(src, dst, dia_format(), hyb_format()); }
template <typename S, typename D>
void device_convert(const S &src, D &dst, ell_format, coo_format) { device_convert_via_csr(src, dst, ell_format(), coo_format()); }
template <typename S, typename D>
void device_convert(const S &src, D &dst, ell_format, dia_format) { device_convert_via_csr(src, dst, ell_format(), dia_format()); }
template <typename S, typename D>
void device_convert(const S &src, D &dst, hyb_format, coo_format) { device_convert_via_csr(src, dst, hyb_format(), coo_format()); }
template <typename S, typename D>
void device_convert(const S &src, D &dst, hyb_format, ell_format) { device_convert_via_csr(src, dst, hyb_format(), ell_format()); }
template <typename S, typename D>
void device_convert(const S &src, D &dst, hyb_format, dia_format) { device_convert_via_csr(src, dst, hyb_format(), dia_format()); }

template <typename F1, typename F2>
struct device_path : std::false_type {};
template <>
struct device_path<csr_format, coo_format> : std::true_type {};
template <>
struct device_path<csr_format, ell_format> : std::true_type {};
template <>
struct device_path<csr_format, hyb_format> : std::true_type {};
template <>
struct device_path<csr_format, dia_format> : std::true_type {};
template <>
struct device_path<coo_format, csr_format> : std::true_type {};
template <>
struct device_path<coo_format, ell_format> : std::true_type {};
template <>
struct device_path<coo_format, hyb_format> : std::true_type {};
template <>
struct device_path<coo_format, dia_format> : std::true_type {};
#define CUSP_B200_DEVICE_PATH(F1, F2) \
  template <>                         \
  struct device_path<F1, F2> : std::true_type {};
CUSP_B200_DEVICE_PATH(dia_format, csr_format)
CUSP_B200_DEVICE_PATH(dia_format, coo_format)
CUSP_B200_DEVICE_PATH(dia_format, ell_format)
CUSP_B200_DEVICE_PATH(dia_format, hyb_format)
CUSP_B200_DEVICE_PATH(ell_format, csr_format)
CUSP_B200_DEVICE_PATH(ell_format, coo_format)
CUSP_B200_DEVICE_PATH(ell_format, dia_format)
CUSP_B200_DEVICE_PATH(ell_format, hyb_format)
CUSP_B200_DEVICE_PATH(hyb_format, csr_format)
CUSP_B200_DEVICE_PATH(hyb_format, coo_format)
CUSP_B200_DEVICE_PATH(hyb_format, ell_format)
CUSP_B200_DEVICE_PATH(hyb_format, dia_format)
#undef CUSP_B200_DEVICE_PATH

template <typename S, typename D, typename F1, typename F2>
void convert_generic(const S &src, D &dst, F1, F2, std::true_type) {
  device_convert(src, dst, F1(), F2());
}
template <typename S, typename D, typename F1, typename F2>
void convert_generic(const S &src, D &dst, F1, F2, std::false_type) {
  host_csr<typename D::index_type, typename D::value_type> H;
  gather(src, H, F1());
  scatter(H, dst, F2());
}

template <typename S, typename D, typename F1, typename F2>
void convert_dispatch(const S &src, D &dst, F1, F2) {
  convert_generic(src, dst, F1(), F2(),
                  std::integral_constant<bool, device_path<F1, F2>::value && device_convertible<S, D>::value>());
}
template <typename S, typename D, typename F1, typename F2>
void convert_dispatch_host_only(const S &src, D &dst, F1, F2) {
  host_csr<typename D::index_type, typename D::value_type> H;
  gather(src, H, F1());
  scatter(H, dst, F2());
}
// dense destinations have no index_type
template <typename S, typename D, typename F1>
void convert_dispatch(const S &src, D &dst, F1, array2d_format) {
  host_csr<typename S::index_type, typename D::value_type> H;
  gather(src, H, F1());
  scatter(H, dst, array2d_format());
}
template <typename S, typename D>
void convert_dispatch(const S &src, D &dst, array2d_format, array2d_format) {
  copy_same(src, dst, array2d_format());
}

}  // namespace detail

template <typename SourceType, typename DestinationType>
void convert(const SourceType &src, DestinationType &dst) {
  detail::convert_dispatch(src, dst, typename SourceType::format(), typename DestinationType::format());
}
namespace detail {
namespace adl_default {
template <typename P, typename SourceType, typename DestinationType>
void convert(cusp::execution_policy<P> &, const SourceType &src, DestinationType &dst) {
  cusp::convert(src, dst);
}
}  // namespace adl_default
}  // namespace detail
// leading execution policy: dispatched on the derived policy (cusp/memory.h: derived_cast;
// testing/convert.cu:662-690, TestConvertDispatch)
template <typename P, typename SourceType, typename DestinationType>
void convert(const execution_policy<P> &exec, const SourceType &src, DestinationType &dst) {
  using detail::adl_default::convert;
  detail::stream_scope<P> on_stream(detail::derived_cast(exec));
  convert(detail::derived_cast(exec), src, dst);
}

// cusp::copy: same format, any memory spaces (cusp/copy.h)
template <typename SourceType, typename DestinationType>
void copy(const SourceType &src, DestinationType &dst) {
  detail::copy_same(src, dst, typename SourceType::format());
}

// explicit CSR -> ELL / HYB with a chosen width (csr_to_other.h:155-306 optional arguments)
template <typename SourceType, typename DestinationType>
void convert(const SourceType &src, DestinationType &dst, size_t num_entries_per_row, size_t alignment = 32) {
  detail::host_csr<typename DestinationType::index_type, typename DestinationType::value_type> H;
  detail::gather(src, H, typename SourceType::format());
  detail::scatter(H, dst, typename DestinationType::format(), num_entries_per_row, alignment);
}

// cusp/format_utils.h entry points used by the tests
template <typename Array>
size_t compute_max_entries_per_row(const Array &row_offsets) {
  return detail::max_entries_per_row(detail::to_host_vector(row_offsets));
}
template <typename Array>
size_t compute_optimal_entries_per_row(const Array &row_offsets, float relative_speed = 3.0f,
                                       size_t breakeven_threshold = 4096) {
  return detail::optimal_entries_per_row(detail::to_host_vector(row_offsets), relative_speed, breakeven_threshold);
}

}  // namespace cusp
