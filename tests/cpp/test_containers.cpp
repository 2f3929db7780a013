// Container API of the drop-in headers (SURVEY §8b "API surface the compatibility headers must keep"):
// testing/{csr,coo,dia,ell,hyb}_matrix.cu restated — shape constructors and the sizes of the public arrays
// (pitch = round_up(rows, alignment) for ELL / DIA), copy construction, resize, swap, rebind across
// memory spaces, COO sort helpers — for host_memory and device_memory.
#include <cusp/array2d.h>
#include <cusp/coo_matrix.h>
#include <cusp/csr_matrix.h>
#include <cusp/dia_matrix.h>
#include <cusp/ell_matrix.h>
#include <cusp/hyb_matrix.h>

#include "check.h"

template <class Space>
void TestCsrMatrixContainer() {  // testing/csr_matrix.cu:4-140
  cusp::csr_matrix<int, float, Space> matrix(3, 2, 6);
  ASSERT_EQUAL(matrix.num_rows, (size_t)3);
  ASSERT_EQUAL(matrix.num_cols, (size_t)2);
  ASSERT_EQUAL(matrix.num_entries, (size_t)6);
  ASSERT_EQUAL(matrix.row_offsets.size(), (size_t)4);
  ASSERT_EQUAL(matrix.column_indices.size(), (size_t)6);
  ASSERT_EQUAL(matrix.values.size(), (size_t)6);
  for (int r = 0; r <= 3; ++r) matrix.row_offsets[r] = 2 * r;
  for (int n = 0; n < 6; ++n) {
    matrix.column_indices[n] = n % 2;
    matrix.values[n] = (float)n;
  }
  cusp::csr_matrix<int, float, Space> copy_of_matrix(matrix);
  ASSERT_EQUAL(copy_of_matrix.num_rows, (size_t)3);
  ASSERT_EQUAL(copy_of_matrix.num_entries, (size_t)6);
  ASSERT_EQUAL(copy_of_matrix.row_offsets, matrix.row_offsets);
  ASSERT_EQUAL(copy_of_matrix.column_indices, matrix.column_indices);
  ASSERT_EQUAL(copy_of_matrix.values, matrix.values);

  cusp::csr_matrix<int, float, Space> resized;
  resized.resize(3, 2, 6);
  ASSERT_EQUAL(resized.num_rows, (size_t)3);
  ASSERT_EQUAL(resized.row_offsets.size(), (size_t)4);
  ASSERT_EQUAL(resized.values.size(), (size_t)6);

  cusp::csr_matrix<int, float, Space> A(1, 2, 2), B(3, 1, 3);
  A.row_offsets[0] = 0; A.row_offsets[1] = 2;
  A.column_indices[0] = 0; A.values[0] = 0; A.column_indices[1] = 1; A.values[1] = 1;
  for (int r = 0; r <= 3; ++r) B.row_offsets[r] = r;
  for (int n = 0; n < 3; ++n) { B.column_indices[n] = 0; B.values[n] = (float)n; }
  cusp::csr_matrix<int, float, Space> A_copy(A), B_copy(B);
  A.swap(B);
  ASSERT_EQUAL(A.num_rows, (size_t)3);
  ASSERT_EQUAL(A.num_cols, (size_t)1);
  ASSERT_EQUAL(A.num_entries, (size_t)3);
  ASSERT_EQUAL(A.row_offsets, B_copy.row_offsets);
  ASSERT_EQUAL(A.values, B_copy.values);
  ASSERT_EQUAL(B.num_rows, (size_t)1);
  ASSERT_EQUAL(B.column_indices, A_copy.column_indices);
}
TEST_HOST_DEVICE(TestCsrMatrixContainer)

template <class Space>
void TestCooMatrixContainer() {  // testing/coo_matrix.cu
  cusp::coo_matrix<int, float, Space> matrix(3, 2, 6);
  ASSERT_EQUAL(matrix.row_indices.size(), (size_t)6);
  ASSERT_EQUAL(matrix.column_indices.size(), (size_t)6);
  ASSERT_EQUAL(matrix.values.size(), (size_t)6);
  const int rows[6] = {2, 0, 1, 0, 2, 1}, cols[6] = {1, 1, 0, 0, 0, 1};
  for (int n = 0; n < 6; ++n) {
    matrix.row_indices[n] = rows[n];
    matrix.column_indices[n] = cols[n];
    matrix.values[n] = (float)n;
  }
  ASSERT_EQUAL(matrix.is_sorted_by_row(), false);
  cusp::coo_matrix<int, float, Space> by_row(matrix);
  by_row.sort_by_row();
  ASSERT_EQUAL(by_row.is_sorted_by_row(), true);
  cusp::coo_matrix<int, float, Space> by_both(matrix);
  by_both.sort_by_row_and_column();
  ASSERT_EQUAL(by_both.is_sorted_by_row_and_column(), true);
  const float want[6] = {3, 1, 2, 5, 4, 0};  // (0,0) (0,1) (1,0) (1,1) (2,0) (2,1)
  for (int n = 0; n < 6; ++n) ASSERT_EQUAL((float)by_both.values[n], want[n]);
  cusp::coo_matrix<int, float, Space> other(1, 1, 1);
  other.row_indices[0] = 0; other.column_indices[0] = 0; other.values[0] = 9;
  other.swap(by_both);
  ASSERT_EQUAL(other.num_entries, (size_t)6);
  ASSERT_EQUAL(by_both.num_entries, (size_t)1);
  ASSERT_EQUAL((float)by_both.values[0], 9.0f);
  matrix.resize(5, 4, 2);
  ASSERT_EQUAL(matrix.num_rows, (size_t)5);
  ASSERT_EQUAL(matrix.values.size(), (size_t)2);
}
TEST_HOST_DEVICE(TestCooMatrixContainer)

template <class Space>
void TestEllMatrixContainer() {  // testing/ell_matrix.cu:4-120
  cusp::ell_matrix<int, float, Space> matrix(3, 2, 6, 2, 4);
  ASSERT_EQUAL(matrix.num_rows, (size_t)3);
  ASSERT_EQUAL(matrix.num_entries, (size_t)6);
  ASSERT_EQUAL(matrix.column_indices.num_cols, (size_t)2);
  ASSERT_EQUAL(matrix.column_indices.num_rows, (size_t)3);
  ASSERT_EQUAL(matrix.column_indices.pitch, (size_t)4);
  ASSERT_EQUAL(matrix.column_indices.num_entries, (size_t)6);
  ASSERT_EQUAL(matrix.values.num_cols, (size_t)2);
  ASSERT_EQUAL(matrix.values.pitch, (size_t)4);
  for (int n = 0; n < 8; ++n) {
    matrix.column_indices.values[n] = n / 4;
    matrix.values.values[n] = (float)n;
  }
  cusp::ell_matrix<int, float, Space> copy_of_matrix(matrix);
  ASSERT_EQUAL(copy_of_matrix.column_indices.pitch, (size_t)4);
  ASSERT_EQUAL(copy_of_matrix.column_indices.values, matrix.column_indices.values);
  ASSERT_EQUAL(copy_of_matrix.values.values, matrix.values.values);
  typedef cusp::ell_matrix<int, float, Space> Ell;
  ASSERT_EQUAL((int)Ell::invalid_index, -1);
  cusp::ell_matrix<int, float, Space> dflt(70, 9, 100, 3);  // default alignment 32
  ASSERT_EQUAL(dflt.values.pitch, (size_t)96);
  cusp::ell_matrix<int, float, Space> A(1, 2, 2, 2, 1), B(3, 1, 3, 1, 1);
  A.swap(B);
  ASSERT_EQUAL(A.num_rows, (size_t)3);
  ASSERT_EQUAL(A.values.num_cols, (size_t)1);
  ASSERT_EQUAL(B.values.num_cols, (size_t)2);
  matrix.resize(5, 5, 9, 2, 8);
  ASSERT_EQUAL(matrix.values.pitch, (size_t)8);
  ASSERT_EQUAL(matrix.column_indices.num_rows, (size_t)5);
}
TEST_HOST_DEVICE(TestEllMatrixContainer)

template <class Space>
void TestDiaMatrixContainer() {  // testing/dia_matrix.cu:4-120
  cusp::dia_matrix<int, float, Space> matrix(4, 5, 7, 3, 8);
  ASSERT_EQUAL(matrix.num_rows, (size_t)4);
  ASSERT_EQUAL(matrix.num_cols, (size_t)5);
  ASSERT_EQUAL(matrix.num_entries, (size_t)7);
  ASSERT_EQUAL(matrix.diagonal_offsets.size(), (size_t)3);
  ASSERT_EQUAL(matrix.values.num_rows, (size_t)4);
  ASSERT_EQUAL(matrix.values.num_cols, (size_t)3);
  ASSERT_EQUAL(matrix.values.pitch, (size_t)8);
  cusp::dia_matrix<int, float, Space> packed(4, 5, 7, 3, 1);
  ASSERT_EQUAL(packed.values.pitch, (size_t)4);
  packed.diagonal_offsets[0] = -2; packed.diagonal_offsets[1] = 0; packed.diagonal_offsets[2] = 1;
  const float v[12] = {0, 0, 13, 16, 10, 0, 14, 0, 11, 12, 15, 0};
  for (int n = 0; n < 12; ++n) packed.values.values[n] = v[n];
  cusp::dia_matrix<int, float, Space> copy_of_matrix(packed);
  ASSERT_EQUAL(copy_of_matrix.diagonal_offsets, packed.diagonal_offsets);
  ASSERT_EQUAL(copy_of_matrix.values.values, packed.values.values);
  cusp::dia_matrix<int, float, Space> other(1, 1, 1, 1, 1);
  other.swap(packed);
  ASSERT_EQUAL(other.num_rows, (size_t)4);
  ASSERT_EQUAL(packed.num_rows, (size_t)1);
  matrix.resize(6, 6, 10, 2);  // the 4-argument resize takes pitch = num_rows (detail/dia_matrix.inl:53-59)
  ASSERT_EQUAL(matrix.values.pitch, (size_t)6);
  ASSERT_EQUAL(matrix.diagonal_offsets.size(), (size_t)2);
  matrix.resize(6, 6, 10, 2, 32);
  ASSERT_EQUAL(matrix.values.pitch, (size_t)32);
}
TEST_HOST_DEVICE(TestDiaMatrixContainer)

template <class Space>
void TestHybMatrixContainer() {  // testing/hyb_matrix.cu:4-34
  cusp::hyb_matrix<int, float, Space> matrix(10, 10, 42, 13, 5, 16);
  ASSERT_EQUAL(matrix.num_rows, (size_t)10);
  ASSERT_EQUAL(matrix.num_entries, (size_t)55);
  ASSERT_EQUAL(matrix.ell.num_entries, (size_t)42);
  ASSERT_EQUAL(matrix.ell.column_indices.num_rows, (size_t)10);
  ASSERT_EQUAL(matrix.ell.column_indices.num_cols, (size_t)5);
  ASSERT_EQUAL(matrix.ell.column_indices.pitch, (size_t)16);
  ASSERT_EQUAL(matrix.ell.values.pitch, (size_t)16);
  ASSERT_EQUAL(matrix.coo.num_rows, (size_t)10);
  ASSERT_EQUAL(matrix.coo.num_entries, (size_t)13);
  ASSERT_EQUAL(matrix.coo.row_indices.size(), (size_t)13);
  ASSERT_EQUAL(matrix.coo.values.size(), (size_t)13);
  cusp::hyb_matrix<int, float, Space> copy_of_matrix(matrix);
  ASSERT_EQUAL(copy_of_matrix.num_entries, (size_t)55);
  ASSERT_EQUAL(copy_of_matrix.ell.values.pitch, (size_t)16);
  cusp::hyb_matrix<int, float, Space> other(3, 4, 5, 3, 2, 1);
  other.swap(matrix);
  ASSERT_EQUAL(other.num_rows, (size_t)10);
  ASSERT_EQUAL(matrix.num_rows, (size_t)3);
  ASSERT_EQUAL(matrix.coo.num_entries, (size_t)3);
}
TEST_HOST_DEVICE(TestHybMatrixContainer)

// rebind<Space>::type and construction across memory spaces (testing/*_matrix.cu: Test*MatrixRebind)
void TestMatrixRebind() {
  typedef cusp::csr_matrix<int, float, cusp::host_memory> HostCsr;
  typedef HostCsr::rebind<cusp::device_memory>::type DeviceCsr;
  HostCsr h_csr(10, 10, 100);
  for (int r = 0; r <= 10; ++r) h_csr.row_offsets[r] = 10 * r;
  for (int n = 0; n < 100; ++n) { h_csr.column_indices[n] = n % 10; h_csr.values[n] = (float)n; }
  DeviceCsr d_csr(h_csr);
  ASSERT_EQUAL(h_csr.num_entries, d_csr.num_entries);
  ASSERT_EQUAL(d_csr.values == h_csr.values, true);
  typedef cusp::ell_matrix<int, double, cusp::host_memory>::rebind<cusp::device_memory>::type DeviceEll;
  cusp::ell_matrix<int, double, cusp::host_memory> h_ell(h_csr);
  DeviceEll d_ell(h_ell);
  ASSERT_EQUAL(d_ell.values.pitch, h_ell.values.pitch);
  ASSERT_EQUAL(d_ell.values.values == h_ell.values.values, true);
  typedef cusp::hyb_matrix<int, float, cusp::host_memory>::rebind<cusp::device_memory>::type DeviceHyb;
  cusp::hyb_matrix<int, float, cusp::host_memory> h_hyb(h_csr);
  DeviceHyb d_hyb(h_hyb);
  ASSERT_EQUAL(d_hyb.num_entries, h_hyb.num_entries);
}
TEST_DEVICE(TestMatrixRebind)

// testing/{csr,coo,ell,dia,hyb}_matrix_view.cu — views alias the container's arrays: built from parts, from a
// matrix (View V = M; make_*_matrix_view(M)), from another view, from a const matrix; writes go through
template <class Space>
void TestMatrixViews() {
  typedef typename cusp::array1d<int, Space>::iterator IndexIterator;
  typedef typename cusp::array1d<float, Space>::iterator ValueIterator;
  typedef cusp::array1d_view<IndexIterator> IndexView;
  typedef cusp::array1d_view<ValueIterator> ValueView;
  {  // csr_matrix_view.cu:7-194
    typedef cusp::csr_matrix<int, float, Space> Matrix;
    typedef cusp::csr_matrix_view<IndexView, IndexView, ValueView> View;
    Matrix M(3, 2, 6);
    View V(3, 2, 6, cusp::make_array1d_view(M.row_offsets.begin(), M.row_offsets.end()),
           cusp::make_array1d_view(M.column_indices.begin(), M.column_indices.end()),
           cusp::make_array1d_view(M.values.begin(), M.values.end()));
    ASSERT_EQUAL(V.num_rows, (size_t)3);
    ASSERT_EQUAL(V.num_entries, (size_t)6);
    ASSERT_TRUE(V.row_offsets.begin() == M.row_offsets.begin() && V.values.end() == M.values.end());
    View W(M);
    ASSERT_TRUE(W.column_indices.begin() == M.column_indices.begin());
    View A = M;  // assignment form
    View B = A;
    ASSERT_TRUE(B.values.begin() == M.values.begin() && B.num_cols == 2);
    View P = cusp::make_csr_matrix_view(3, 2, 6, cusp::make_array1d_view(M.row_offsets),
                                        cusp::make_array1d_view(M.column_indices), cusp::make_array1d_view(M.values));
    P.row_offsets[0] = 0;
    P.column_indices[0] = 1;
    P.values[0] = 2;
    ASSERT_EQUAL((int)M.column_indices[0], 1);
    ASSERT_EQUAL((float)M.values[0], 2.0f);
    View Q = cusp::make_csr_matrix_view(M);
    View R = cusp::make_csr_matrix_view(Q);
    ASSERT_TRUE(R.values.begin() == M.values.begin());
    const Matrix CM(3, 2, 6);
    ASSERT_EQUAL(cusp::make_csr_matrix_view(CM).num_entries, (size_t)6);
    ASSERT_TRUE(cusp::make_csr_matrix_view(CM).values.begin() == CM.values.begin());
  }
  {  // coo_matrix_view.cu
    typedef cusp::coo_matrix<int, float, Space> Matrix;
    typedef cusp::coo_matrix_view<IndexView, IndexView, ValueView> View;
    Matrix M(3, 2, 6);
    View V(3, 2, 6, cusp::make_array1d_view(M.row_indices), cusp::make_array1d_view(M.column_indices),
           cusp::make_array1d_view(M.values));
    ASSERT_TRUE(V.row_indices.begin() == M.row_indices.begin() && V.num_entries == 6);
    View W = cusp::make_coo_matrix_view(M);
    W.values[5] = 9.0f;
    ASSERT_EQUAL((float)M.values[5], 9.0f);
    const Matrix CM(3, 2, 6);
    ASSERT_TRUE(cusp::make_coo_matrix_view(CM).row_indices.begin() == CM.row_indices.begin());
  }
  {  // ell_matrix_view.cu / dia_matrix_view.cu / hyb_matrix_view.cu: from a matrix, writes alias
    cusp::ell_matrix<int, float, Space> E(3, 2, 6, 2, 4);
    auto EV = cusp::make_ell_matrix_view(E);
    ASSERT_EQUAL(EV.num_entries, (size_t)6);
    ASSERT_EQUAL(EV.values.pitch, (size_t)4);
    EV.values.values[1] = 3.0f;
    ASSERT_EQUAL((float)E.values.values[1], 3.0f);
    cusp::dia_matrix<int, float, Space> D(4, 5, 7, 3, 8);
    auto DV = cusp::make_dia_matrix_view(D);
    DV.diagonal_offsets[2] = 1;
    ASSERT_EQUAL((int)D.diagonal_offsets[2], 1);
    ASSERT_EQUAL(DV.values.pitch, (size_t)8);
    cusp::hyb_matrix<int, float, Space> H(10, 10, 42, 13, 5, 16);
    auto HV = cusp::make_hyb_matrix_view(H);
    ASSERT_EQUAL(HV.num_entries, (size_t)55);
    HV.coo.values[12] = 4.0f;
    ASSERT_EQUAL((float)H.coo.values[12], 4.0f);
    typename cusp::hyb_matrix<int, float, Space>::view HV2(H);
    ASSERT_EQUAL(HV2.ell.values.pitch, (size_t)16);
  }
}
static void TestMatrixViewsHost() { TestMatrixViews<cusp::host_memory>(); }
TEST_HOST(TestMatrixViewsHost)

// testing/format.cu and testing/memory.cu — compile-time traits
void TestFormatAndMemoryTraits() {
  typedef cusp::array1d<float, cusp::host_memory> A1D;
  typedef cusp::array2d<float, cusp::host_memory> A2D;
  typedef cusp::coo_matrix<int, float, cusp::host_memory> COO;
  typedef cusp::csr_matrix<int, float, cusp::device_memory> CSR;
  typedef cusp::dia_matrix<int, float, cusp::host_memory> DIA;
  typedef cusp::ell_matrix<int, float, cusp::device_memory> ELL;
  typedef cusp::hyb_matrix<int, float, cusp::host_memory> HYB;
  static_assert(std::is_same<A1D::format, cusp::array1d_format>::value &&
                    !std::is_convertible<A1D::format, cusp::sparse_format>::value &&
                    std::is_convertible<A1D::format, cusp::dense_format>::value &&
                    std::is_convertible<A1D::format, cusp::known_format>::value, "array1d format");
  static_assert(std::is_same<A2D::format, cusp::array2d_format>::value &&
                    std::is_convertible<A2D::format, cusp::dense_format>::value, "array2d format");
  static_assert(std::is_same<COO::format, cusp::coo_format>::value && std::is_same<CSR::format, cusp::csr_format>::value &&
                    std::is_same<DIA::format, cusp::dia_format>::value && std::is_same<ELL::format, cusp::ell_format>::value &&
                    std::is_same<HYB::format, cusp::hyb_format>::value, "sparse formats");
  static_assert(std::is_convertible<HYB::format, cusp::sparse_format>::value &&
                    !std::is_convertible<HYB::format, cusp::dense_format>::value &&
                    std::is_convertible<ELL::format, cusp::known_format>::value, "sparse format hierarchy");
  static_assert(std::is_same<CSR::memory_space, cusp::device_memory>::value &&
                    std::is_same<CSR::view::memory_space, cusp::device_memory>::value, "memory_space of containers and views");
  typedef cusp::host_memory H;
  typedef cusp::device_memory D;
  typedef cusp::any_memory A;
  static_assert(std::is_same<cusp::minimum_space<H, H>::type, H>::value && std::is_same<cusp::minimum_space<H, A>::type, H>::value &&
                    std::is_same<cusp::minimum_space<A, H>::type, H>::value && std::is_same<cusp::minimum_space<D, D>::type, D>::value &&
                    std::is_same<cusp::minimum_space<D, A>::type, D>::value && std::is_same<cusp::minimum_space<A, D>::type, D>::value &&
                    std::is_same<cusp::minimum_space<A, A>::type, A>::value, "minimum_space of two");
  static_assert(std::is_same<cusp::minimum_space<H, H, A>::type, H>::value && std::is_same<cusp::minimum_space<A, H, A>::type, H>::value &&
                    std::is_same<cusp::minimum_space<D, A, A>::type, D>::value && std::is_same<cusp::minimum_space<A, A, A>::type, A>::value &&
                    std::is_same<cusp::minimum_space<H, A, H>::type, H>::value && std::is_same<cusp::minimum_space<A, D, D>::type, D>::value,
                "minimum_space of three");
  ASSERT_TRUE(true);
}
TEST_HOST(TestFormatAndMemoryTraits)
