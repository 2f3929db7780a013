// Views over raw device pointers — the reference's examples/Views/cg_raw.cu:27-107:
// cudaMalloc'd COO arrays wrapped in device pointers -> array1d_view ->
// coo_matrix_view, solved in place by cusp::krylov::cg.  Also csr/ell/dia views
// built with make_*_matrix_view and cusp::cuda::par.on(stream).
#include <cuda_runtime.h>

#include <cusp/array1d.h>
#include <cusp/coo_matrix.h>
#include <cusp/csr_matrix.h>
#include <cusp/gallery/poisson.h>
#include <cusp/krylov/cg.h>
#include <cusp/monitor.h>
#include <cusp/multiply.h>

#include "check.h"

void TestCgRawPointers() {
  int host_I[10] = {0, 0, 1, 1, 1, 2, 2, 2, 3, 3};
  int host_J[10] = {0, 1, 0, 1, 2, 1, 2, 3, 2, 3};
  float host_V[10] = {2, -1, -1, 2, -1, -1, 2, -1, -1, 2};
  float host_x[4] = {0, 0, 0, 0};
  float host_b[4] = {1, 2, 2, 1};
  int *device_I, *device_J;
  float *device_V, *device_x, *device_b;
  cudaMalloc(&device_I, 10 * sizeof(int));
  cudaMalloc(&device_J, 10 * sizeof(int));
  cudaMalloc(&device_V, 10 * sizeof(float));
  cudaMalloc(&device_x, 4 * sizeof(float));
  cudaMalloc(&device_b, 4 * sizeof(float));
  cudaMemcpy(device_I, host_I, 10 * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(device_J, host_J, 10 * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(device_V, host_V, 10 * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(device_x, host_x, 4 * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(device_b, host_b, 4 * sizeof(float), cudaMemcpyHostToDevice);

  cusp::device_ptr<int> wrapped_device_I(device_I), wrapped_device_J(device_J);
  cusp::device_ptr<float> wrapped_device_V(device_V), wrapped_device_x(device_x), wrapped_device_b(device_b);
  typedef cusp::array1d_view<cusp::device_ptr<int>> DeviceIndexArrayView;
  typedef cusp::array1d_view<cusp::device_ptr<float>> DeviceValueArrayView;
  DeviceIndexArrayView row_indices(wrapped_device_I, wrapped_device_I + 10);
  DeviceIndexArrayView column_indices(wrapped_device_J, wrapped_device_J + 10);
  DeviceValueArrayView values(wrapped_device_V, wrapped_device_V + 10);
  DeviceValueArrayView x(wrapped_device_x, wrapped_device_x + 4);
  DeviceValueArrayView b(wrapped_device_b, wrapped_device_b + 4);
  typedef cusp::coo_matrix_view<DeviceIndexArrayView, DeviceIndexArrayView, DeviceValueArrayView> DeviceView;
  DeviceView A(4, 4, 10, row_indices, column_indices, values);

  cusp::monitor<float> monitor(b, 100, 1e-5, 0, false);
  cusp::krylov::cg(A, x, b, monitor);
  cudaMemcpy(host_x, device_x, 4 * sizeof(float), cudaMemcpyDeviceToHost);
  ASSERT_TRUE(monitor.converged());
  // tridiag(-1,2,-1) x = (1,2,2,1)  ->  x = (3,5,5,3)
  const float expect[4] = {3, 5, 5, 3};
  for (int i = 0; i < 4; ++i) ASSERT_NEAR(host_x[i], expect[i], 1e-4);
  cudaFree(device_I); cudaFree(device_J); cudaFree(device_V); cudaFree(device_x); cudaFree(device_b);
}
TEST_DEVICE(TestCgRawPointers)

void TestHostRawPointerViews() {
  int I[10] = {0, 0, 1, 1, 1, 2, 2, 2, 3, 3};
  int J[10] = {0, 1, 0, 1, 2, 1, 2, 3, 2, 3};
  float V[10] = {2, -1, -1, 2, -1, -1, 2, -1, -1, 2};
  float xs[4] = {1, 2, 3, 4}, ys[4] = {9, 9, 9, 9};
  typedef cusp::array1d_view<int *> IV;
  typedef cusp::array1d_view<float *> VV;
  cusp::coo_matrix_view<IV, IV, VV> A(4, 4, 10, IV(I, I + 10), IV(J, J + 10), VV(V, V + 10));
  VV x(xs, xs + 4), y(ys, ys + 4);
  cusp::multiply(A, x, y);
  ASSERT_EQUAL(ys[0], 0.0f); ASSERT_EQUAL(ys[1], 0.0f); ASSERT_EQUAL(ys[2], 0.0f); ASSERT_EQUAL(ys[3], 5.0f);
}
TEST_HOST(TestHostRawPointerViews)

void TestMakeViewsAndStream() {
  cusp::csr_matrix<int, double, cusp::device_memory> A;
  cusp::gallery::poisson5pt(A, 9, 7);
  cusp::array1d<double, cusp::device_memory> x(A.num_cols, 1.0), y(A.num_rows, -1.0), y2(A.num_rows, -1.0);
  auto V = cusp::make_csr_matrix_view(A.num_rows, A.num_cols, A.num_entries, cusp::make_array1d_view(A.row_offsets),
                                      cusp::make_array1d_view(A.column_indices), cusp::make_array1d_view(A.values));
  cusp::multiply(V, x, y);
  cudaStream_t s;
  cudaStreamCreate(&s);
  cusp::multiply(cusp::cuda::par.on(s), A, x, y2);
  // the policy carries the stream for that one call only: afterwards the thread is back on its own stream
  ASSERT_TRUE(cusp::detail::current_stream() == nullptr);
  cudaStreamSynchronize(s);
  cudaStreamDestroy(s);
  cusp::multiply(A, x, y2);  // on the default stream again: must not touch the destroyed one
  ASSERT_EQUAL(y, y2);
  cusp::array1d<double, cusp::host_memory> h(y);
  ASSERT_EQUAL(h[0], 2.0);   // corner row: 4 - 1 - 1
  ASSERT_EQUAL(h[10], 0.0);  // interior row
}
TEST_DEVICE(TestMakeViewsAndStream)

// generator arrays: testing/array1d_view.cu:449-458 and testing/random.cu:13-112
template <class MemorySpace>
void TestGeneratorArrays() {
  cusp::counting_array<int> W(4, 5);
  ASSERT_EQUAL(W.size(), (size_t)4);
  ASSERT_EQUAL((int)W[0], 5);
  ASSERT_EQUAL((int)W[3], 8);
  cusp::constant_array<int> X(200, 5);
  ASSERT_EQUAL((int)X[0], 5);
  ASSERT_EQUAL((int)X[3], 5);
  ASSERT_EQUAL((int)X[199], 5);
  cusp::array1d<int, MemorySpace> Wd(W);
  ASSERT_EQUAL(Wd == W, true);
  // random integers: every nibble of the raw value is uniform within 5 %
  const size_t n = 123456;
  {
    cusp::random_array<int> random(n);
    size_t counts[8][16] = {{0}};
    for (size_t i = 0; i < n; i++) {
      const unsigned long long raw = (unsigned int)(int)random[i];
      for (size_t nibble = 0; nibble < 8; nibble++) counts[nibble][(raw >> (4 * nibble)) % 16]++;
    }
    size_t lo = n, hi = 0;
    for (auto &row : counts)
      for (size_t c : row) {
        lo = std::min(lo, c);
        hi = std::max(hi, c);
      }
    ASSERT_TRUE(lo >= (size_t)(0.95 * (n / 16)) && hi <= (size_t)(1.05 * (n / 16)));
    cusp::array1d<int, cusp::host_memory> h(random);
    cusp::array1d<int, MemorySpace> d(random);
    ASSERT_EQUAL(h == d, true);
  }
  // random reals: in [0, 1), 32 buckets uniform within 5 %
  {
    cusp::random_array<double> random(n);
    cusp::random_array<float> randomf(n);
    size_t b64[32] = {0}, b32[32] = {0};
    for (size_t i = 0; i < n; i++) {
      const double v = random[i];
      const float f = randomf[i];
      ASSERT_TRUE(0.0 <= v && v < 1.0 && 0.0f <= f && f <= 1.0f);
      b64[(size_t)(v * 32.0)]++;
      b32[std::min<size_t>(31, (size_t)(f * 32.0f))]++;
    }
    for (int k = 0; k < 32; ++k) {
      ASSERT_TRUE(b64[k] >= (size_t)(0.95 * (n / 32)) && b64[k] <= (size_t)(1.05 * (n / 32)));
      ASSERT_TRUE(b32[k] >= (size_t)(0.95 * (n / 32)) && b32[k] <= (size_t)(1.05 * (n / 32)));
    }
    cusp::array1d<double, MemorySpace> d(random);
    ASSERT_EQUAL(d == random, true);
    ASSERT_TRUE(!(cusp::random_array<double>(16, 1) == cusp::random_array<double>(16, 2)));  // the seed matters
  }
  // known answers of the published 64-bit integer hash (T. Wang), seed 0
  ASSERT_EQUAL(cusp::detail::random_hash64(0, 0), 0x77cfa1eef01bca90ull);
  ASSERT_EQUAL(cusp::detail::random_hash64(1, 0), 0x5bca7c69b794f8ceull);
}
TEST_HOST_DEVICE(TestGeneratorArrays)

// testing/linear_operator.cu:5-47 — linear_operator carries shape and types; identity_operator copies;
// a user operator (unknown_format) is applied through operator() by cusp::multiply and works as the A of cg
template <class MemorySpace>
struct scale_by_two : cusp::linear_operator<float, MemorySpace> {
  scale_by_two(int n) : cusp::linear_operator<float, MemorySpace>(n, n) {}
  template <typename V1, typename V2>
  void operator()(const V1 &x, V2 &y) const {
    cusp::blas::axpby(x, x, y, 1.0f, 1.0f);
  }
};
template <class MemorySpace>
void TestLinearOperators() {
  typedef cusp::linear_operator<float, MemorySpace, long> LinearOperator;
  LinearOperator A(4, 3);
  ASSERT_EQUAL(A.num_rows, (size_t)4);
  ASSERT_EQUAL(A.num_cols, (size_t)3);
  static_assert(std::is_same<typename LinearOperator::value_type, float>::value, "value_type");
  static_assert(std::is_same<typename LinearOperator::index_type, long>::value, "index_type");
  cusp::array1d<float, MemorySpace> x(4), y(4);
  x[0] = 7.0f; y[0] = 0.0f; x[1] = 5.0f; y[1] = -2.0f; x[2] = 4.0f; y[2] = 0.0f; x[3] = -3.0f; y[3] = 5.0f;
  cusp::identity_operator<float, MemorySpace> I(4, 4);
  I(x, y);
  ASSERT_EQUAL((float)y[0], 7.0f);
  ASSERT_EQUAL((float)y[3], -3.0f);
  scale_by_two<MemorySpace> S(4);
  cusp::multiply(S, x, y);  // generic/multiply.inl:59-73: unknown_format -> A(x, y)
  ASSERT_EQUAL((float)y[0], 14.0f);
  ASSERT_EQUAL((float)y[3], -6.0f);
  cusp::array1d<float, MemorySpace> b(4, 3.0f), s(4, 0.0f);
  cusp::monitor<float> monitor(b, 10, 1e-6);
  cusp::krylov::cg(S, s, b, monitor);  // 2 s = b
  ASSERT_TRUE(monitor.converged());
  ASSERT_EQUAL((float)s[2], 1.5f);
}
static void TestLinearOperatorsHost() { TestLinearOperators<cusp::host_memory>(); }
TEST_HOST(TestLinearOperatorsHost)

// testing/array1d.cu:10-190 — array1d as a container: push_back, constructors (size, fill, other array /
// std::vector / iterator range), assignment and equality across array types
template <typename MemorySpace>
void TestArray1dContainer() {
  cusp::array1d<int, MemorySpace> a(4);
  ASSERT_EQUAL(a.size(), (size_t)4);
  for (int i = 0; i < 4; ++i) a[i] = i;
  a.push_back(4);
  ASSERT_EQUAL(a.size(), (size_t)5);
  for (int i = 0; i < 5; ++i) ASSERT_EQUAL((int)a[i], i);

  cusp::array1d<int, cusp::host_memory> h(a);
  ASSERT_EQUAL(h.size(), (size_t)5);
  ASSERT_EQUAL((int)h[4], 4);
  const cusp::array1d<int, cusp::host_memory> ch(2, 10);
  ASSERT_EQUAL((int)ch[1], 10);
  cusp::array1d<int, MemorySpace> from_const(ch);
  ASSERT_EQUAL((int)from_const[0], 10);

  std::vector<int> v(2, 10);
  cusp::array1d<int, MemorySpace> av(v);
  ASSERT_EQUAL(av.size(), (size_t)2);
  ASSERT_EQUAL((int)av[1], 10);
  cusp::array1d<int, MemorySpace> ar(v.begin(), v.end());
  ASSERT_EQUAL((int)ar[0], 10);
  cusp::array1d<int, MemorySpace> assigned;
  assigned = v;
  ASSERT_EQUAL(assigned.size(), (size_t)2);
  assigned = a;
  ASSERT_EQUAL(assigned.size(), (size_t)5);
  assigned = ch;
  ASSERT_EQUAL((int)assigned[1], 10);

  cusp::array1d<int, MemorySpace> A(2);
  A[0] = 10;
  A[1] = 20;
  cusp::array1d<int, cusp::host_memory> hh(A.begin(), A.end());
  std::vector<int> vv(2);
  vv[0] = 10;
  vv[1] = 20;
  ASSERT_EQUAL(A == hh, true);
  ASSERT_EQUAL(A == vv, true);
  hh.push_back(30);
  vv.push_back(30);
  ASSERT_EQUAL(A != hh, true);
  ASSERT_EQUAL(A != vv, true);
  // resize / reserve / swap
  A.resize(3);
  ASSERT_EQUAL(A.size(), (size_t)3);
  ASSERT_EQUAL((int)A[1], 20);
  cusp::array1d<int, MemorySpace> B(1, 7);
  A.swap(B);
  ASSERT_EQUAL(A.size(), (size_t)1);
  ASSERT_EQUAL((int)A[0], 7);
  ASSERT_EQUAL(B.size(), (size_t)3);
}
static void TestArray1dContainerHost() { TestArray1dContainer<cusp::host_memory>(); }
TEST_HOST(TestArray1dContainerHost)

// testing/array2d.cu:100-298 — storage order, pitch, resize, swap, mixed orientations
template <class Space>
void TestArray2dContainer() {
  cusp::array2d<float, Space, cusp::row_major> A(2, 3);
  const float vals[6] = {10, 20, 30, 40, 50, 60};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) A(i, j) = vals[3 * i + j];
  for (int n = 0; n < 6; ++n) ASSERT_EQUAL((float)A.values[n], vals[n]);
  A.resize(2, 3, 4);  // non-trivial pitch
  cusp::blas::fill(A.values, 0.0f);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) A(i, j) = vals[3 * i + j];
  const float padded[8] = {10, 20, 30, 0, 40, 50, 60, 0};
  for (int n = 0; n < 8; ++n) ASSERT_EQUAL((float)A.values[n], padded[n]);

  cusp::array2d<float, Space, cusp::column_major> C(2, 3);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) C(i, j) = vals[3 * i + j];
  const float cm[6] = {10, 40, 20, 50, 30, 60};
  for (int n = 0; n < 6; ++n) ASSERT_EQUAL((float)C.values[n], cm[n]);

  // mixed orientations: assignment converts the storage order (array2d.cu:200-229)
  cusp::array2d<float, Space, cusp::row_major> R(2, 3);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) R(i, j) = vals[3 * i + j];
  cusp::array2d<float, Space, cusp::column_major> C2;
  C2 = R;
  cusp::array2d<float, Space, cusp::row_major> R2(C2);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) {
      ASSERT_EQUAL((float)C2(i, j), vals[3 * i + j]);
      ASSERT_EQUAL((float)R2(i, j), vals[3 * i + j]);
    }
  ASSERT_EQUAL(C2 == R, true);

  cusp::array2d<float, Space> Z;
  Z.resize(3, 2);
  ASSERT_EQUAL(Z.pitch, (size_t)2);
  ASSERT_EQUAL(Z.num_entries, (size_t)6);
  ASSERT_EQUAL(Z.values.size(), (size_t)6);
  Z.resize(3, 2, 4);
  ASSERT_EQUAL(Z.pitch, (size_t)4);
  ASSERT_EQUAL(Z.values.size(), (size_t)12);
  ASSERT_THROWS(Z.resize(3, 2, 1), cusp::invalid_input_exception);

  cusp::array2d<float, Space> P(2, 2, 1.0f), Q(3, 1, 2.0f);
  cusp::array2d<float, Space> P_copy(P), Q_copy(Q);
  P.swap(Q);
  ASSERT_EQUAL(P.num_rows, Q_copy.num_rows);
  ASSERT_EQUAL(P.values == Q_copy.values, true);
  ASSERT_EQUAL(Q.num_cols, P_copy.num_cols);
  ASSERT_EQUAL(Q.values == P_copy.values, true);
}
static void TestArray2dContainerHost() { TestArray2dContainer<cusp::host_memory>(); }
TEST_HOST(TestArray2dContainerHost)

// testing/array1d_view.cu:12-564 — views over containers: aliasing, make_array1d_view, re-seating by
// assignment, resize within the capacity, capacity taken from the container, equality, subarray
template <typename MemorySpace>
void TestArray1dViewSemantics() {
  typedef cusp::array1d<int, MemorySpace> Array;
  typedef typename Array::iterator Iterator;
  typedef typename Array::const_iterator ConstIterator;
  typedef cusp::array1d_view<Iterator> View;
  {
    Array A(4);
    A[0] = 10; A[1] = 20; A[2] = 30; A[3] = 40;
    View V(A.begin(), A.end());
    ASSERT_EQUAL(V.size(), (size_t)4);
    ASSERT_EQUAL(V.capacity(), (size_t)4);
    ASSERT_EQUAL((int)V[2], 30);
    ASSERT_TRUE(V.begin() == A.begin() && V.end() == A.end());
    V[1] = 17;
    ASSERT_EQUAL((int)A[1], 17);
    const View CV(A);  // a const view still writes through
    CV[2] = 33;
    ASSERT_EQUAL((int)A[2], 33);
    const Array CA(4, 10);
    cusp::array1d_view<ConstIterator> R(CA.begin(), CA.end());
    ASSERT_EQUAL((int)R[3], 10);
    View M = cusp::make_array1d_view(A.begin(), A.end());
    ASSERT_TRUE(M.begin() == A.begin());
    View M2 = cusp::make_array1d_view(A);
    M2[0] = 5;
    ASSERT_EQUAL((int)A[0], 5);
  }
  {  // assignment re-seats the view (array1d_view.cu:305-343)
    Array A(4), B(8);
    View V(A.begin(), A.end());
    V = View(B);
    ASSERT_EQUAL(V.size(), (size_t)8);
    ASSERT_TRUE(V.begin() == B.begin() && V.end() == B.end());
    const View W = View(V);
    ASSERT_EQUAL(W.capacity(), (size_t)8);
  }
  {  // resize within the capacity (array1d_view.cu:345-383)
    Array A(4);
    View V(A.begin(), A.end());
    V.resize(3);
    ASSERT_EQUAL(V.size(), (size_t)3);
    ASSERT_EQUAL(V.capacity(), (size_t)4);
    ASSERT_TRUE(V.end() == A.begin() + 3);
    V.resize(4);
    ASSERT_EQUAL(V.size(), (size_t)4);
    ASSERT_THROWS(V.resize(5), cusp::not_implemented_exception);
    View W = V;
    V.resize(2);
    ASSERT_EQUAL(W.size(), (size_t)4);
  }
  {  // the capacity comes from the container (array1d_view.cu:414-435)
    Array A(4);
    A.resize(2);
    View V = View(A);
    ASSERT_EQUAL(V.size(), (size_t)2);
    ASSERT_EQUAL(V.capacity(), (size_t)4);
  }
  {  // equality (array1d_view.cu:490-530)
    Array A(2), B(3);
    A[0] = 10; A[1] = 20; B[0] = 10; B[1] = 20; B[2] = 30;
    View V(A), W(B);
    ASSERT_TRUE(A == V && V == A && V == V && !(A != V) && !(V != A));
    ASSERT_TRUE(!(V == B) && !(B == V) && !(V == W) && V != B && B != V && V != W);
    W.resize(2);
    ASSERT_TRUE(V == W && !(V != W));
  }
  {  // subarray (array1d_view.cu:532-564)
    Array A(4);
    A[0] = 10; A[1] = 20; A[2] = 30; A[3] = 40;
    View V = A.subarray(1, 3);
    ASSERT_EQUAL(V.size(), (size_t)3);
    ASSERT_TRUE(V.begin() == A.begin() + 1 && V.end() == A.begin() + 4);
    View W = V.subarray(0, 1);
    ASSERT_EQUAL(W.size(), (size_t)1);
    ASSERT_EQUAL((int)W[0], 20);
  }
}
static void TestArray1dViewSemanticsHost() { TestArray1dViewSemantics<cusp::host_memory>(); }
TEST_HOST(TestArray1dViewSemanticsHost)
