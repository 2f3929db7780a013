// SURVEY §8(f) row 4 — cusp::io MatrixMarket, restating testing/matrix_market.cu:13-330.  The
// reference reads ../data/test/*.mtx (not shipped with the fork); the same files are written
// here from their published contents (NIST MatrixMarket format examples the expected dense
// images in that test correspond to).
#include <cusp/array1d.h>
#include <cusp/array2d.h>
#include <cusp/coo_matrix.h>
#include <cusp/csr_matrix.h>
#include <cusp/ell_matrix.h>
#include <cusp/gallery/poisson.h>
#include <cusp/hyb_matrix.h>
#include <cusp/io/matrix_market.h>
#include <cusp/multiply.h>

#include <cstdio>
#include <sstream>

#include "check.h"

static const char random_file_name[] = "test_93298409283221.mtx";

static const char coordinate_real_general[] =
    "%%MatrixMarket matrix coordinate real general\n"
    "%=================================================================================\n"
    "% a 5x5 sparse matrix with 8 nonzeros\n"
    "%=================================================================================\n"
    "  5  5  8\n"
    "    1     1   1.000e+00\n"
    "    2     2   1.050e+01\n"
    "    3     3   2.500e-01\n"
    "    1     4   6.000e+00\n"
    "    4     2   2.505e+02\n"
    "    4     4  -2.500e+02\n"
    "    4     5   3.875e+01\n"
    "    5     5   1.200e+01\n";
static const char coordinate_pattern_symmetric[] =
    "%%MatrixMarket matrix coordinate pattern symmetric\r\n"
    "% CRLF line ends on purpose\r\n"
    "5 5 7\r\n"
    "1 1\r\n2 2\r\n3 3\r\n4 2\r\n4 4\r\n5 4\r\n5 5\r\n";
static const char array_real_general[] =
    "%%MatrixMarket matrix array real general\n"
    "% column-major\n"
    "4 3\n"
    "1.0\n2.0\n3.0\n4.0\n5.0\n6.0\n7.0\n8.0\n9.0\n10.0\n11.0\n12.0\n";

static cusp::array2d<float, cusp::host_memory> expected_real_general() {
  cusp::array2d<float, cusp::host_memory> E(5, 5, 0.0f);
  E(0, 0) = 1.000e+00f; E(0, 3) = 6.000e+00f; E(1, 1) = 1.050e+01f; E(2, 2) = 2.500e-01f;
  E(3, 1) = 2.505e+02f; E(3, 3) = -2.500e+02f; E(3, 4) = 3.875e+01f; E(4, 4) = 1.200e+01f;
  return E;
}

template <class MemorySpace>
void TestReadWriteMarketFileRealArray1d() {  // matrix_market.cu:13-34
  cusp::array1d<float, cusp::host_memory> a(5);
  a[0] = 10; a[1] = 0; a[2] = 20; a[3] = 0; a[4] = 30;
  cusp::io::write_matrix_market_file(a, random_file_name);
  cusp::array1d<float, MemorySpace> b;
  cusp::io::read_matrix_market_file(b, random_file_name);
  remove(random_file_name);
  ASSERT_EQUAL(a == b, true);
}
TEST_HOST_DEVICE(TestReadWriteMarketFileRealArray1d)

void TestReadMatrixMarketCoordinateRealGeneral() {  // :60-101
  cusp::coo_matrix<int, float, cusp::host_memory> coo;
  std::istringstream in(coordinate_real_general);
  cusp::io::read_matrix_market_stream(coo, in);
  ASSERT_EQUAL(coo.num_entries, (size_t)8);
  ASSERT_TRUE(coo.is_sorted_by_row_and_column());
  cusp::array2d<float, cusp::host_memory> D(coo);
  ASSERT_EQUAL(D == expected_real_general(), true);
}
TEST_HOST(TestReadMatrixMarketCoordinateRealGeneral)

void TestReadMatrixMarketCoordinatePatternSymmetric() {  // :146-185
  cusp::coo_matrix<int, float, cusp::host_memory> coo;
  std::istringstream in(coordinate_pattern_symmetric);
  cusp::io::read_matrix_market_stream(coo, in);
  ASSERT_EQUAL(coo.num_entries, (size_t)9);  // 5 diagonal + 2 mirrored pairs
  cusp::array2d<float, cusp::host_memory> D(coo);
  cusp::array2d<float, cusp::host_memory> E(5, 5, 0.0f);
  for (int i = 0; i < 5; ++i) E(i, i) = 1.0f;
  E(1, 3) = E(3, 1) = 1.0f;
  E(3, 4) = E(4, 3) = 1.0f;
  ASSERT_EQUAL(D == E, true);
}
TEST_HOST(TestReadMatrixMarketCoordinatePatternSymmetric)

void TestReadMatrixMarketArrayRealGeneral() {  // :187-213
  cusp::coo_matrix<int, float, cusp::host_memory> coo;
  std::istringstream in(array_real_general);
  cusp::io::read_matrix_market_stream(coo, in);
  cusp::array2d<float, cusp::host_memory> D(coo);
  cusp::array2d<float, cusp::host_memory> E(4, 3);
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 4; ++i) E(i, j) = (float)(1 + i + 4 * j);
  ASSERT_EQUAL(D == E, true);
}
TEST_HOST(TestReadMatrixMarketArrayRealGeneral)

template <class MemorySpace>
void TestReadMatrixMarketFileToCsrMatrix() {  // :215-255
  {
    std::ofstream f(random_file_name);
    f << coordinate_real_general;
  }
  cusp::csr_matrix<int, float, MemorySpace> csr;
  cusp::io::read_matrix_market_file(csr, random_file_name);
  remove(random_file_name);
  cusp::array2d<float, cusp::host_memory> D(csr);
  ASSERT_EQUAL(D == expected_real_general(), true);
}
TEST_HOST_DEVICE(TestReadMatrixMarketFileToCsrMatrix)

template <class MemorySpace>
void TestWriteMatrixMarketFileCoordinateRealGeneral() {  // :257-291
  cusp::array2d<float, cusp::host_memory> E(4, 3, 0.0f);
  E(0, 0) = 1.000e+00f; E(1, 1) = 1.050e+01f; E(2, 2) = 2.500e-01f; E(3, 1) = 2.505e+02f;
  cusp::coo_matrix<int, float, MemorySpace> coo(E);
  cusp::io::write_matrix_market_file(coo, random_file_name);
  cusp::io::read_matrix_market_file(coo, random_file_name);
  remove(random_file_name);
  cusp::array2d<float, cusp::host_memory> D(coo);
  ASSERT_EQUAL(D == E, true);
}
TEST_HOST_DEVICE(TestWriteMatrixMarketFileCoordinateRealGeneral)

// every sparse format and both value types survive write -> read bit for bit (max_digits10), and
// the product of the re-read matrix equals the product of the original in that format
template <class MemorySpace>
void TestMatrixMarketRoundTripAllFormats() {
  cusp::csr_matrix<int, double, cusp::host_memory> P;
  cusp::gallery::poisson5pt(P, 9, 7);
  for (size_t n = 0; n < P.num_entries; ++n) P.values[n] = (double)P.values[n] * (1.0 / 3.0 + 1e-9 * (double)n);
  cusp::array1d<double, MemorySpace> x(P.num_cols);
  for (size_t i = 0; i < P.num_cols; ++i) x[i] = 0.25 + (double)(i % 5);
  cusp::array1d<double, MemorySpace> y0(P.num_rows), y1(P.num_rows);
  auto roundtrip = [&](auto A) {
    A = P;
    std::stringstream ss;
    cusp::io::write_matrix_market_stream(A, ss);
    decltype(A) B;
    cusp::io::read_matrix_market_stream(B, ss);
    cusp::csr_matrix<int, double, cusp::host_memory> Q(B);
    ASSERT_EQUAL(Q.row_offsets == P.row_offsets, true);
    ASSERT_EQUAL(Q.column_indices == P.column_indices, true);
    ASSERT_EQUAL(Q.values == P.values, true);
    cusp::multiply(A, x, y0);  // same format, same kernel: identical arrays give identical bits
    cusp::multiply(B, x, y1);
    ASSERT_EQUAL(y0 == y1, true);
  };
  roundtrip(cusp::coo_matrix<int, double, MemorySpace>());
  roundtrip(cusp::csr_matrix<int, double, MemorySpace>());
  roundtrip(cusp::ell_matrix<int, double, MemorySpace>());
  roundtrip(cusp::hyb_matrix<int, double, MemorySpace>());
}
TEST_HOST_DEVICE(TestMatrixMarketRoundTripAllFormats)

void TestMatrixMarketErrors() {  // matrix_market.inl:80-98, 203-227, 271-279
  auto read = [](const char *text) {
    cusp::coo_matrix<int, float, cusp::host_memory> coo;
    std::istringstream in(text);
    cusp::io::read_matrix_market_stream(coo, in);
  };
  ASSERT_THROWS(read("%MatrixMarket matrix coordinate real general\n1 1 0\n"), cusp::io_exception);
  ASSERT_THROWS(read("%%MatrixMarket matrix banded real general\n1 1 0\n"), cusp::io_exception);
  ASSERT_THROWS(read("%%MatrixMarket matrix coordinate quaternion general\n1 1 0\n"), cusp::io_exception);
  ASSERT_THROWS(read("%%MatrixMarket matrix coordinate real general\n2 2\n"), cusp::io_exception);
  ASSERT_THROWS(read("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n"), cusp::io_exception);  // EOF
  ASSERT_THROWS(read("%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 1.0\n"), cusp::io_exception);
  ASSERT_THROWS(read("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 3 1.0\n"), cusp::io_exception);
  ASSERT_THROWS(read("%%MatrixMarket matrix coordinate real hermitian\n2 2 1\n1 1 1.0\n"), cusp::not_implemented_exception);
  ASSERT_THROWS(read("%%MatrixMarket matrix coordinate real skew-symmetric\n2 2 1\n2 1 1.0\n"),
                cusp::not_implemented_exception);
  ASSERT_THROWS(read("%%MatrixMarket matrix array pattern general\n1 1\n"), cusp::not_implemented_exception);
  cusp::coo_matrix<int, float, cusp::host_memory> coo;
  ASSERT_THROWS(cusp::io::read_matrix_market_file(coo, "/nonexistent/dir/file.mtx"), cusp::io_exception);
}
TEST_HOST(TestMatrixMarketErrors)

// cusp::print (cusp/detail/print.inl:33-140): same text as the reference for arrays and every sparse format
#include <cusp/print.h>
void TestPrint() {
  cusp::array1d<float, cusp::host_memory> a(2);
  a[0] = 1.5f;
  a[1] = -2.0f;
  std::ostringstream s1;
  cusp::print(a, s1);
  // " " + setw(8) of "(" + value + ")": eight blanks before the parenthesis (print.inl:36)
  ASSERT_EQUAL(s1.str(), std::string("array1d <2>\n        (1.5)\n        (-2)\n"));
  cusp::array2d<float, cusp::host_memory> D(2, 2, 0.0f);
  D(0, 1) = 3.0f;
  D(1, 0) = 4.25f;
  cusp::csr_matrix<int, float, cusp::host_memory> A(D);
  std::ostringstream s2, s3;
  cusp::print(A, s2);
  cusp::coo_matrix<int, float, cusp::host_memory> C(A);
  cusp::print(C, s3);
  ASSERT_EQUAL(s2.str(), s3.str());  // every sparse format prints through its COO image
  ASSERT_TRUE(s2.str().find("sparse matrix <2, 2> with 2 entries\n") == 0);
  ASSERT_TRUE(s2.str().find("(4.25)") != std::string::npos);
  std::ostringstream s4;
  cusp::print(D, s4);
  ASSERT_TRUE(s4.str().find("array2d <2, 2>\n") == 0);
}
TEST_HOST(TestPrint)
