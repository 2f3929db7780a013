// cusp::convert — the reference's testing/convert.cu: the 4x4 / 7-entry example
// hard-coded in every format (:63-200), every source x destination pair in both
// memory spaces compared through the dense image (:225-385), and the exact
// array layouts of CSR -> DIA / ELL / HYB (:405-497); plus format_utils
// (testing/format_utils.cu:13-75).
#include <cusp/array2d.h>
#include <cusp/convert.h>
#include <cusp/format_utils.h>

#include "check.h"

static const int X = -1;  // ell_matrix::invalid_index

template <typename M>
void init_csr(M &csr) {
  csr.resize(4, 4, 7);
  const int ro[5] = {0, 2, 3, 6, 7};
  const int ci[7] = {0, 1, 2, 0, 2, 3, 1};
  const float v[7] = {10.25f, 11.00f, 12.50f, 13.75f, 14.00f, 15.25f, 16.50f};
  for (int i = 0; i < 5; ++i) csr.row_offsets[i] = ro[i];
  for (int i = 0; i < 7; ++i) { csr.column_indices[i] = ci[i]; csr.values[i] = v[i]; }
}
template <typename M>
void init_coo(M &coo) {
  coo.resize(4, 4, 7);
  const int ri[7] = {0, 0, 1, 2, 2, 2, 3};
  const int ci[7] = {0, 1, 2, 0, 2, 3, 1};
  const float v[7] = {10.25f, 11.00f, 12.50f, 13.75f, 14.00f, 15.25f, 16.50f};
  for (int i = 0; i < 7; ++i) { coo.row_indices[i] = ri[i]; coo.column_indices[i] = ci[i]; coo.values[i] = v[i]; }
}
template <typename M>
void init_dia(M &dia) {
  dia.resize(4, 4, 7, 3, 1);
  const int off[3] = {-2, 0, 1};
  const float v[12] = {0, 0, 13.75f, 16.50f, 10.25f, 0, 14.00f, 0, 11.00f, 12.50f, 15.25f, 0};
  for (int i = 0; i < 3; ++i) dia.diagonal_offsets[i] = off[i];
  for (int i = 0; i < 12; ++i) dia.values.values[i] = v[i];
}
template <typename M>
void init_ell(M &ell) {
  ell.resize(4, 4, 7, 3, 1);
  const int ci[12] = {0, 2, 0, 1, 1, X, 2, X, X, X, 3, X};
  const float v[12] = {10.25f, 12.50f, 13.75f, 16.50f, 11.00f, 0, 14.00f, 0, 0, 0, 15.25f, 0};
  for (int i = 0; i < 12; ++i) { ell.column_indices.values[i] = ci[i]; ell.values.values[i] = v[i]; }
}
template <typename M>
void init_hyb(M &hyb) {
  hyb.resize(4, 4, 4, 3, 1, 1);
  const int eci[4] = {0, 2, 0, 1};
  const float ev[4] = {10.25f, 12.50f, 13.75f, 16.50f};
  for (int i = 0; i < 4; ++i) { hyb.ell.column_indices.values[i] = eci[i]; hyb.ell.values.values[i] = ev[i]; }
  const int ri[3] = {0, 2, 2}, ci[3] = {1, 2, 3};
  const float v[3] = {11.00f, 14.00f, 15.25f};
  for (int i = 0; i < 3; ++i) { hyb.coo.row_indices[i] = ri[i]; hyb.coo.column_indices[i] = ci[i]; hyb.coo.values[i] = v[i]; }
}
template <typename S>
void init(cusp::csr_matrix<int, float, S> &m) { init_csr(m); }
template <typename S>
void init(cusp::coo_matrix<int, float, S> &m) { init_coo(m); }
template <typename S>
void init(cusp::dia_matrix<int, float, S> &m) { init_dia(m); }
template <typename S>
void init(cusp::ell_matrix<int, float, S> &m) { init_ell(m); }
template <typename S>
void init(cusp::hyb_matrix<int, float, S> &m) { init_hyb(m); }

static cusp::array2d<float, cusp::host_memory> dense_image() {
  cusp::array2d<float, cusp::host_memory> D(4, 4, 0.0f);
  D(0, 0) = 10.25f; D(0, 1) = 11.00f; D(1, 2) = 12.50f; D(2, 0) = 13.75f;
  D(2, 2) = 14.00f; D(2, 3) = 15.25f; D(3, 1) = 16.50f;
  return D;
}

template <typename Src, typename Dst>
void convert_pair() {
  Src src;
  init(src);
  Dst dst;
  cusp::convert(src, dst);
  ASSERT_EQUAL(dst.num_rows, (size_t)4);
  ASSERT_EQUAL(dst.num_cols, (size_t)4);
  ASSERT_EQUAL(dst.num_entries, (size_t)7);
  cusp::array2d<float, cusp::host_memory> image(dst);
  ASSERT_TRUE(image == dense_image());
  Dst via_ctor(src);  // converting constructor and assignment
  cusp::array2d<float, cusp::host_memory> image2(via_ctor);
  ASSERT_TRUE(image2 == dense_image());
}

template <typename Src, typename DstSpace>
void convert_from() {
  convert_pair<Src, cusp::coo_matrix<int, float, DstSpace>>();
  convert_pair<Src, cusp::csr_matrix<int, float, DstSpace>>();
  convert_pair<Src, cusp::dia_matrix<int, float, DstSpace>>();
  convert_pair<Src, cusp::ell_matrix<int, float, DstSpace>>();
  convert_pair<Src, cusp::hyb_matrix<int, float, DstSpace>>();
}
template <typename SrcSpace, typename DstSpace>
void convert_all() {
  convert_from<cusp::coo_matrix<int, float, SrcSpace>, DstSpace>();
  convert_from<cusp::csr_matrix<int, float, SrcSpace>, DstSpace>();
  convert_from<cusp::dia_matrix<int, float, SrcSpace>, DstSpace>();
  convert_from<cusp::ell_matrix<int, float, SrcSpace>, DstSpace>();
  convert_from<cusp::hyb_matrix<int, float, SrcSpace>, DstSpace>();
}
void TestConvertHostToHost() { convert_all<cusp::host_memory, cusp::host_memory>(); }
TEST_HOST(TestConvertHostToHost)
void TestConvertAcrossSpaces() {
  convert_all<cusp::host_memory, cusp::device_memory>();
  convert_all<cusp::device_memory, cusp::host_memory>();
  convert_all<cusp::device_memory, cusp::device_memory>();
}
TEST_DEVICE(TestConvertAcrossSpaces)

// exact layouts (testing/convert.cu:405-497)
template <typename MemorySpace>
void TestConvertExactLayouts() {
  cusp::csr_matrix<int, float, MemorySpace> csr;
  init_csr(csr);
  {
    cusp::dia_matrix<int, float, cusp::host_memory> want;
    init_dia(want);
    cusp::dia_matrix<int, float, MemorySpace> dia;
    cusp::detail::host_csr<int, float> H;
    cusp::detail::gather(csr, H, cusp::csr_format());
    cusp::detail::scatter(H, dia, cusp::dia_format(), 1);  // alignment 1 -> pitch 4
    ASSERT_EQUAL(dia.diagonal_offsets, want.diagonal_offsets);
    ASSERT_EQUAL(dia.values.pitch, (size_t)4);
    ASSERT_EQUAL(dia.values.values, want.values.values);
    cusp::dia_matrix<int, float, MemorySpace> dia32(csr);  // default alignment 32
    ASSERT_EQUAL(dia32.values.pitch, (size_t)32);
    ASSERT_EQUAL(dia32.diagonal_offsets, want.diagonal_offsets);
  }
  {
    cusp::ell_matrix<int, float, cusp::host_memory> want;
    init_ell(want);
    cusp::ell_matrix<int, float, MemorySpace> ell;
    cusp::convert(csr, ell, 3, 1);
    ASSERT_EQUAL(ell.column_indices.values, want.column_indices.values);
    ASSERT_EQUAL(ell.values.values, want.values.values);
    cusp::ell_matrix<int, float, MemorySpace> ell32(csr);
    ASSERT_EQUAL(ell32.column_indices.pitch, (size_t)32);
    ASSERT_EQUAL(ell32.column_indices.num_cols, (size_t)3);
  }
  {
    cusp::hyb_matrix<int, float, cusp::host_memory> want;
    init_hyb(want);
    cusp::hyb_matrix<int, float, MemorySpace> hyb;
    cusp::convert(csr, hyb, 1, 1);
    ASSERT_EQUAL(hyb.ell.column_indices.values, want.ell.column_indices.values);
    ASSERT_EQUAL(hyb.ell.values.values, want.ell.values.values);
    ASSERT_EQUAL(hyb.coo.row_indices, want.coo.row_indices);
    ASSERT_EQUAL(hyb.coo.column_indices, want.coo.column_indices);
    ASSERT_EQUAL(hyb.coo.values, want.coo.values);
  }
}
TEST_HOST_DEVICE(TestConvertExactLayouts)

template <typename MemorySpace>
void TestFormatUtils() {
  const int off[8] = {0, 0, 0, 1, 1, 2, 5, 10};
  const int idx[10] = {2, 4, 5, 5, 5, 6, 6, 6, 6, 6};
  cusp::array1d<int, MemorySpace> offsets(off, off + 8), indices(idx, idx + 10);
  cusp::array1d<int, MemorySpace> got_idx(10), got_off(8);
  cusp::offsets_to_indices(offsets, got_idx);
  cusp::indices_to_offsets(indices, got_off);
  ASSERT_EQUAL(got_idx, indices);
  ASSERT_EQUAL(got_off, offsets);
  ASSERT_EQUAL(cusp::compute_max_entries_per_row(offsets), (size_t)5);
  cusp::csr_matrix<int, float, MemorySpace> csr;
  init_csr(csr);
  cusp::array1d<float, MemorySpace> d;
  cusp::extract_diagonal(csr, d);
  ASSERT_EQUAL(d.size(), (size_t)4);
  ASSERT_EQUAL(d[0], 10.25f); ASSERT_EQUAL(d[1], 0.0f); ASSERT_EQUAL(d[2], 14.0f); ASSERT_EQUAL(d[3], 0.0f);
}
TEST_HOST_DEVICE(TestFormatUtils)

// COO sorting helpers (cusp/detail/coo_matrix.inl:95-127)
template <typename MemorySpace>
void TestCooSort() {
  cusp::coo_matrix<int, float, MemorySpace> A(3, 3, 4);
  const int r[4] = {2, 0, 2, 1}, c[4] = {1, 2, 0, 1};
  for (int i = 0; i < 4; ++i) { A.row_indices[i] = r[i]; A.column_indices[i] = c[i]; A.values[i] = (float)i; }
  ASSERT_EQUAL(A.is_sorted_by_row(), false);
  A.sort_by_row();
  ASSERT_EQUAL(A.is_sorted_by_row(), true);
  ASSERT_EQUAL(A.values[1], 3.0f);  // stable: row 2 keeps (col 1, col 0) order
  ASSERT_EQUAL(A.column_indices[2], 1);
  ASSERT_EQUAL(A.is_sorted_by_row_and_column(), false);
  A.sort_by_row_and_column();
  ASSERT_EQUAL(A.is_sorted_by_row_and_column(), true);
  ASSERT_EQUAL(A.column_indices[2], 0);
}
TEST_HOST_DEVICE(TestCooSort)

// fill-in guard (csr_to_other.h:118-121,196-199)
void TestConvertFillInGuard() {
  const size_t n = 600000;
  cusp::csr_matrix<int, float, cusp::host_memory> A(n, n, n + 3);
  // row 0 has 4 entries, every other row 1: ELL would need 4 n slots > 3 x nnz and > 1e6
  A.row_offsets[0] = 0;
  for (size_t i = 0; i < n; ++i) A.row_offsets[i + 1] = (int)(i + 4);
  for (size_t k = 0; k < 4; ++k) { A.column_indices[k] = (int)k; A.values[k] = 1; }
  for (size_t i = 1; i < n; ++i) { A.column_indices[i + 3] = (int)i; A.values[i + 3] = 1; }
  cusp::ell_matrix<int, float, cusp::host_memory> ell;
  ASSERT_THROWS(cusp::convert(A, ell), cusp::format_conversion_exception);
  cusp::hyb_matrix<int, float, cusp::host_memory> hyb(A);  // HYB: K = 1, 3 COO entries
  ASSERT_EQUAL(hyb.ell.column_indices.num_cols, (size_t)1);
  ASSERT_EQUAL(hyb.coo.num_entries, (size_t)3);
}
TEST_HOST(TestConvertFillInGuard)
