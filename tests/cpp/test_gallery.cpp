// cusp::gallery — dense golden images of the reference's testing/poisson.cu:6-93,
// bit-identical arrays between the host pipeline (stencil -> DIA -> convert) and
// the engine's device builders, gallery::random determinism, and the banded
// generators of cusp/ktt/matrix_generation.h.
#include <cusp/array2d.h>
#include <cusp/gallery/poisson.h>
#include <cusp/gallery/random.h>
#include <cusp/ktt/matrix_generation.h>

#include "check.h"

template <typename MemorySpace, size_t N>
void expect_image(const cusp::array2d<float, MemorySpace> &got, const float (&want)[N][N]) {
  ASSERT_EQUAL(got.num_rows, N);
  ASSERT_EQUAL(got.num_cols, N);
  cusp::array2d<float, cusp::host_memory> h(got);
  for (size_t i = 0; i < N; ++i)
    for (size_t j = 0; j < N; ++j) ASSERT_EQUAL(h(i, j), want[i][j]);
}

template <typename MemorySpace>
void TestPoissonGoldenImages() {
  const float p5[6][6] = {{4, -1, -1, 0, 0, 0}, {-1, 4, 0, -1, 0, 0}, {-1, 0, 4, -1, -1, 0},
                          {0, -1, -1, 4, 0, -1}, {0, 0, -1, 0, 4, -1}, {0, 0, 0, -1, -1, 4}};
  const float p9[6][6] = {{8, -1, -1, -1, 0, 0}, {-1, 8, -1, -1, 0, 0}, {-1, -1, 8, -1, -1, -1},
                          {-1, -1, -1, 8, -1, -1}, {0, 0, -1, -1, 8, -1}, {0, 0, -1, -1, -1, 8}};
  const float p7[8][8] = {{6, -1, -1, 0, -1, 0, 0, 0}, {-1, 6, 0, -1, 0, -1, 0, 0}, {-1, 0, 6, -1, 0, 0, -1, 0},
                          {0, -1, -1, 6, 0, 0, 0, -1}, {-1, 0, 0, 0, 6, -1, -1, 0}, {0, -1, 0, 0, -1, 6, 0, -1},
                          {0, 0, -1, 0, -1, 0, 6, -1}, {0, 0, 0, -1, 0, -1, -1, 6}};
  float p27[8][8];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) p27[i][j] = i == j ? 26.0f : -1.0f;
  {
    cusp::csr_matrix<int, float, MemorySpace> A;
    cusp::gallery::poisson5pt(A, 2, 3);
    ASSERT_EQUAL(A.num_entries, (size_t)20);
    expect_image(cusp::array2d<float, MemorySpace>(A), p5);
  }
  {
    cusp::dia_matrix<int, float, MemorySpace> A;
    cusp::gallery::poisson9pt(A, 2, 3);
    expect_image(cusp::array2d<float, MemorySpace>(A), p9);
  }
  {
    cusp::ell_matrix<int, float, MemorySpace> A;
    cusp::gallery::poisson7pt(A, 2, 2, 2);
    expect_image(cusp::array2d<float, MemorySpace>(A), p7);
  }
  {
    cusp::coo_matrix<int, float, MemorySpace> A;
    cusp::gallery::poisson27pt(A, 2, 2, 2);
    expect_image(cusp::array2d<float, MemorySpace>(A), p27);
  }
}
TEST_HOST_DEVICE(TestPoissonGoldenImages)

// device builders == host pipeline, array for array
template <typename V>
void device_equals_host(size_t nx, size_t ny, size_t nz) {
  {
    cusp::dia_matrix<int, V, cusp::host_memory> h;
    cusp::dia_matrix<int, V, cusp::device_memory> d;
    if (nz) { cusp::gallery::poisson7pt(h, nx, ny, nz); cusp::gallery::poisson7pt(d, nx, ny, nz); }
    else { cusp::gallery::poisson5pt(h, nx, ny); cusp::gallery::poisson5pt(d, nx, ny); }
    ASSERT_EQUAL(d.num_entries, h.num_entries);
    ASSERT_EQUAL(d.values.pitch, h.values.pitch);
    ASSERT_EQUAL(d.diagonal_offsets, h.diagonal_offsets);
    ASSERT_EQUAL(d.values.values, h.values.values);
  }
  {
    cusp::ell_matrix<int, V, cusp::host_memory> h;
    cusp::ell_matrix<int, V, cusp::device_memory> d;
    if (nz) { cusp::gallery::poisson7pt(h, nx, ny, nz); cusp::gallery::poisson7pt(d, nx, ny, nz); }
    else { cusp::gallery::poisson5pt(h, nx, ny); cusp::gallery::poisson5pt(d, nx, ny); }
    ASSERT_EQUAL(d.num_entries, h.num_entries);
    ASSERT_EQUAL(d.column_indices.pitch, h.column_indices.pitch);
    ASSERT_EQUAL(d.column_indices.num_cols, h.column_indices.num_cols);
    ASSERT_EQUAL(d.column_indices.values, h.column_indices.values);
    ASSERT_EQUAL(d.values.values, h.values.values);
  }
  {
    cusp::csr_matrix<int, V, cusp::host_memory> h;
    cusp::csr_matrix<int, V, cusp::device_memory> d;
    if (nz) { cusp::gallery::poisson7pt(h, nx, ny, nz); cusp::gallery::poisson7pt(d, nx, ny, nz); }
    else { cusp::gallery::poisson5pt(h, nx, ny); cusp::gallery::poisson5pt(d, nx, ny); }
    ASSERT_EQUAL(d.num_entries, h.num_entries);
    ASSERT_EQUAL(d.row_offsets, h.row_offsets);
    ASSERT_EQUAL(d.column_indices, h.column_indices);
    ASSERT_EQUAL(d.values, h.values);
  }
}
void TestPoissonDeviceBuilders() {
  device_equals_host<float>(7, 5, 0);
  device_equals_host<double>(33, 17, 0);
  device_equals_host<float>(5, 4, 3);
  device_equals_host<double>(17, 9, 11);
  device_equals_host<double>(1, 1, 1);
}
TEST_DEVICE(TestPoissonDeviceBuilders)

void TestGalleryRandom() {
  cusp::coo_matrix<int, float, cusp::host_memory> A, B;
  cusp::gallery::random(A, 50, 40, 300);
  cusp::gallery::random(B, 50, 40, 300);
  ASSERT_TRUE(A.num_entries <= 300 && A.num_entries > 200);
  ASSERT_EQUAL(A.row_indices, B.row_indices);
  ASSERT_EQUAL(A.column_indices, B.column_indices);
  ASSERT_EQUAL(A.is_sorted_by_row_and_column(), true);
  for (size_t k = 1; k < A.num_entries; ++k)
    ASSERT_TRUE(A.row_indices[k] != A.row_indices[k - 1] || A.column_indices[k] != A.column_indices[k - 1]);
  for (size_t k = 0; k < A.num_entries; ++k) ASSERT_EQUAL(A.values[k], 1.0f);
}
TEST_HOST(TestGalleryRandom)

void TestDiagonalGenerators() {
  auto A = cusp::ktt::make_diagonal_symmetric_matrix(8, 6, 2, 3);  // first = -2*3/2 = -3: offsets -3, -1, 1
  ASSERT_EQUAL(A.diagonal_offsets.size(), (size_t)3);
  ASSERT_EQUAL(A.diagonal_offsets[0], -3); ASSERT_EQUAL(A.diagonal_offsets[1], -1); ASSERT_EQUAL(A.diagonal_offsets[2], 1);
  ASSERT_EQUAL(A.values.pitch, (size_t)8);
  ASSERT_EQUAL(A.num_entries, (size_t)(5 + 6 + 5));
  ASSERT_EQUAL(A.values(2, 0), 0.0f); ASSERT_EQUAL(A.values(3, 0), 1.0f); ASSERT_EQUAL(A.values(7, 0), 1.0f);
  ASSERT_EQUAL(A.values(0, 1), 0.0f); ASSERT_EQUAL(A.values(1, 1), 1.0f); ASSERT_EQUAL(A.values(6, 1), 1.0f);
  ASSERT_EQUAL(A.values(7, 1), 0.0f);
  ASSERT_EQUAL(A.values(0, 2), 1.0f); ASSERT_EQUAL(A.values(4, 2), 1.0f); ASSERT_EQUAL(A.values(5, 2), 0.0f);
  ASSERT_THROWS(cusp::ktt::make_diagonal_symmetric_matrix(4, 4, 1, 10), std::runtime_error);
}
TEST_HOST(TestDiagonalGenerators)
