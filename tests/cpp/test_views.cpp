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
  cudaStreamSynchronize(s);
  cusp::cuda::par.on(0);
  cudaStreamDestroy(s);
  ASSERT_EQUAL(y, y2);
  cusp::array1d<double, cusp::host_memory> h(y);
  ASSERT_EQUAL(h[0], 2.0);   // corner row: 4 - 1 - 1
  ASSERT_EQUAL(h[10], 0.0);  // interior row
}
TEST_DEVICE(TestMakeViewsAndStream)
