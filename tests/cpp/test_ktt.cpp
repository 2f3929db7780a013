// cusp::ktt — the reference's testing/ktt.cu: every configuration of the tuning
// space for dia / ell / ellr <int,float,device_memory> on the five small
// matrices and the three all-ones banded matrices (:214-282), validated against
// the non-tuned device path (:173-196), failing on any invalid result (:84-140);
// extended to csr / coo / hyb (the reference has no test for its CSR/COO KTT
// kernels), dynamic tuning, explicit configurations and reset_tuning.
#include <cstring>

#include <cusp/array2d.h>
#include <cusp/coo_matrix.h>
#include <cusp/csr_matrix.h>
#include <cusp/dia_matrix.h>
#include <cusp/ell_matrix.h>
#include <cusp/hyb_matrix.h>
#include <cusp/gallery/poisson.h>
#include <cusp/ktt/ellr_matrix.h>
#include <cusp/ktt/ktt.h>
#include <cusp/ktt/matrix_generation.h>
#include <cusp/multiply.h>

#include "check.h"

struct UnitTestStopCondition : ::ktt::StopCondition {
  bool IsFulfilled() const override { return failed_ || explored_ == total_; }
  void Initialize(const uint64_t configurationsCount) override {
    total_ = configurationsCount;
    explored_ = 0;
    failed_ = false;
  }
  void Update(const ::ktt::KernelResult &result) override {
    failed_ = failed_ || !result.IsValid();
    explored_++;
  }
  std::string GetStatusString() const override {
    if (failed_) return "Encountered failing configuration";
    return "No failing configuration encountered. Explored configurations: " + std::to_string(explored_) + " / " +
           std::to_string(total_);
  }
  bool failed_ = false;
  uint64_t total_ = 0, explored_ = 0;
};

static void assert_tuning_results_valid(const std::vector<::ktt::KernelResult> &results, const std::string &arg_name) {
  ASSERT_TRUE(!results.empty());
  for (const auto &result : results) {
    if (result.IsValid()) continue;
    std::string reason;
    switch (result.GetStatus()) {
      case ::ktt::ResultStatus::Ok: continue;
      case ::ktt::ResultStatus::CompilationFailed: reason = "CompilationFailed"; break;
      case ::ktt::ResultStatus::ComputationFailed: reason = "ComputationFailed"; break;
      case ::ktt::ResultStatus::DeviceLimitsExceeded: reason = "DeviceLimitsExceeded"; break;
      case ::ktt::ResultStatus::ValidationFailed: reason = "ValidationFailed"; break;
    }
    std::string conf;
    for (auto parameter : result.GetConfiguration().GetPairs()) conf += "  " + parameter.GetString() + "\n";
    CHECK_FAIL(result.GetKernelName() << ": " << reason << " on matrix " << arg_name << " in configuration:\n" << conf);
  }
}

template <typename SparseMatrixType, typename TestMatrixType>
void CheckAllConfigurations(const TestMatrixType &test_matrix, const std::string &arg_name) {
  using ValueType = typename SparseMatrixType::value_type;
  using DeviceTestMatrix = typename TestMatrixType::template rebind<cusp::device_memory>::type;

  cusp::array1d<ValueType, cusp::host_memory> host_x(test_matrix.num_cols);
  for (size_t i = 0; i < host_x.size(); i++) host_x[i] = i % 10;
  cusp::array1d<ValueType, cusp::device_memory> device_x = host_x;
  cusp::array1d<ValueType, cusp::host_memory> reference_y(test_matrix.num_rows, 10);

  DeviceTestMatrix device_matrix = test_matrix;
  {
    cusp::array1d<ValueType, cusp::device_memory> y(test_matrix.num_rows, 10);
    cusp::ktt::disable();
    cusp::multiply(device_matrix, device_x, y);
    cusp::ktt::enable();
    reference_y = y;
  }
  // the non-tuned device result itself equals the host product (exact: small integers)
  {
    cusp::array1d<ValueType, cusp::host_memory> host_y(test_matrix.num_rows, 10);
    cusp::multiply(test_matrix, host_x, host_y);
    ASSERT_EQUAL(host_y, reference_y);
  }

  SparseMatrixType A = device_matrix;
  cusp::array1d<ValueType, cusp::host_memory> host_y(A.num_rows, 10);
  cusp::array1d<ValueType, cusp::device_memory> device_y = host_y;

  ::ktt::ReferenceComputation reference_computation = [&](void *raw_buffer) {
    std::memcpy(raw_buffer, (void *)reference_y.data(), sizeof(ValueType) * reference_y.size());
  };
  std::ostringstream log;
  cusp::ktt::get_tuner().SetLoggingTarget(log);
  auto stop = std::make_unique<UnitTestStopCondition>();
  auto results = cusp::ktt::tune(A, device_x, device_y, reference_computation, std::move(stop));
  cusp::ktt::get_tuner().SetLoggingTarget(std::cerr);
  assert_tuning_results_valid(results, arg_name);
  // the space was explored completely (minus points whose smem ring cannot hold this matrix's K)
  b200sp_matrix d = cusp::detail::describe(A);
  const int64_t space = b200sp_cfg_space(d.format, d.dtype, nullptr, 0);
  ASSERT_TRUE((int64_t)results.size() <= space && (int64_t)results.size() * 2 > space);
  // y holds the product computed by the winner
  ASSERT_EQUAL(device_y, reference_y);
}
#define CHECK_ALL_CONFIGURATIONS(MatrixTypeUnderTest, input_matrix) \
  CheckAllConfigurations<MatrixTypeUnderTest>(input_matrix, #input_matrix)

template <class TestMatrix>
void TestKttSparseMatrixVectorMultiply() {
  using ValueType = typename TestMatrix::value_type;
  using IndexType = typename TestMatrix::index_type;
  typedef cusp::array2d<ValueType, cusp::host_memory> Dense;
  typedef cusp::coo_matrix<IndexType, ValueType, cusp::host_memory> Coo;

  Dense A(5, 4);
  const ValueType a[5][4] = {{13, 80, 0, 0}, {0, 27, 0, 0}, {55, 0, 24, 42}, {0, 69, 0, 83}, {0, 0, 27, 0}};
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 4; ++j) A(i, j) = a[i][j];
  Coo A_coo = A;
  CHECK_ALL_CONFIGURATIONS(TestMatrix, A_coo);

  Dense B(2, 4);
  const ValueType b[2][4] = {{0, 2, 3, 4}, {5, 0, 0, 8}};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 4; ++j) B(i, j) = b[i][j];
  Coo B_coo = B;
  CHECK_ALL_CONFIGURATIONS(TestMatrix, B_coo);

  Dense C(2, 2);
  C(0, 0) = 0; C(0, 1) = 0; C(1, 0) = 3; C(1, 1) = 5;
  Coo C_coo = C;
  CHECK_ALL_CONFIGURATIONS(TestMatrix, C_coo);

  Dense D(2, 1);
  D(0, 0) = 2; D(1, 0) = 3;
  Coo D_coo = D;
  CHECK_ALL_CONFIGURATIONS(TestMatrix, D_coo);

  Dense F(2, 3);
  F(0, 0) = 0; F(0, 1) = 1.5; F(0, 2) = 3.0; F(1, 0) = 0.5; F(1, 1) = 0; F(1, 2) = 0;
  Coo F_coo = F;
  CHECK_ALL_CONFIGURATIONS(TestMatrix, F_coo);
}

template <class TestMatrix>
void TestKttBanded() {
  auto G_dia = cusp::ktt::make_diagonal_symmetric_matrix(4096, 4096, 1, 1024);
  CHECK_ALL_CONFIGURATIONS(TestMatrix, G_dia);
  auto H_dia = cusp::ktt::make_diagonal_symmetric_matrix(4096, 2048, 1, 1024);
  CHECK_ALL_CONFIGURATIONS(TestMatrix, H_dia);
  auto I_dia = cusp::ktt::make_diagonal_symmetric_matrix(2048, 4096, 1, 1024);
  CHECK_ALL_CONFIGURATIONS(TestMatrix, I_dia);
}

#define KTT_CASE(fn, Name, ...)                                  \
  static void fn##Name() { fn<__VA_ARGS__>(); }                  \
  static check::registrar reg_##fn##Name(#fn "<" #Name ">", true, fn##Name);
typedef cusp::dia_matrix<int, float, cusp::device_memory> DiaF;
typedef cusp::ell_matrix<int, float, cusp::device_memory> EllF;
typedef cusp::ktt::ellr_matrix<int, float, cusp::device_memory> EllrF;
typedef cusp::csr_matrix<int, float, cusp::device_memory> CsrF;
typedef cusp::coo_matrix<int, float, cusp::device_memory> CooF;
typedef cusp::hyb_matrix<int, float, cusp::device_memory> HybF;
typedef cusp::dia_matrix<int, double, cusp::device_memory> DiaD;
typedef cusp::csr_matrix<int, double, cusp::device_memory> CsrD;
KTT_CASE(TestKttSparseMatrixVectorMultiply, Dia, DiaF)
KTT_CASE(TestKttSparseMatrixVectorMultiply, Ell, EllF)
KTT_CASE(TestKttSparseMatrixVectorMultiply, Ellr, EllrF)
KTT_CASE(TestKttSparseMatrixVectorMultiply, Csr, CsrF)
KTT_CASE(TestKttSparseMatrixVectorMultiply, Coo, CooF)
KTT_CASE(TestKttSparseMatrixVectorMultiply, Hyb, HybF)
KTT_CASE(TestKttSparseMatrixVectorMultiply, DiaF64, DiaD)
KTT_CASE(TestKttSparseMatrixVectorMultiply, CsrF64, CsrD)
KTT_CASE(TestKttBanded, Dia, DiaF)
KTT_CASE(TestKttBanded, Ell, EllF)
KTT_CASE(TestKttBanded, Ellr, EllrF)
KTT_CASE(TestKttBanded, Csr, CsrF)

// dynamic tuning: every call runs one new configuration, then the best one
void TestKttDynamicTuning() {
  auto host = cusp::ktt::make_diagonal_symmetric_matrix(3000, 3000, 3, 9);
  DiaF A = host;
  cusp::array1d<float, cusp::device_memory> x(A.num_cols, 1.0f), y(A.num_rows, 0.0f), want(A.num_rows, 0.0f);
  cusp::ktt::disable();
  cusp::multiply(A, x, want);
  cusp::ktt::enable();
  cusp::ktt::reset_tuning(A, x, y);
  b200sp_matrix d = cusp::detail::describe(A);
  const int64_t space = b200sp_cfg_space(d.format, d.dtype, nullptr, 0);
  std::vector<std::string> seen;
  for (int64_t i = 0; i < space + 3; ++i) {
    cusp::blas::fill(y, -1.0f);
    ::ktt::KernelResult r = cusp::ktt::multiply(A, x, y);
    ASSERT_TRUE(r.IsValid());
    ASSERT_EQUAL(y, want);
    const std::string conf = r.GetConfiguration().GetString();
    if (i < space) {
      for (const std::string &s : seen) ASSERT_TRUE(s != conf);  // a new point every step
      seen.push_back(conf);
    } else {
      bool known = false;
      for (const std::string &s : seen) known = known || s == conf;
      ASSERT_TRUE(known);  // exhausted: the best known configuration
    }
  }
  // an explicit configuration
  ::ktt::KernelConfiguration conf =
      cusp::ktt::get_tuner().CreateConfiguration(0, {{"KERNEL", B200SP_K_DIA_LDG}, {"BLOCK_SIZE", 256}, {"UNROLL", 4}});
  cusp::blas::fill(y, -1.0f);
  ::ktt::KernelResult r = cusp::ktt::multiply(A, x, y, conf);
  ASSERT_TRUE(r.IsValid());
  ASSERT_EQUAL(y, want);
  ASSERT_THROWS(::ktt::KernelConfiguration({{"NO_SUCH_PARAMETER", 1}}), std::runtime_error);
}
TEST_DEVICE(TestKttDynamicTuning)

// Searcher and stop condition take effect DURING the search (cuda/ktt/multiply.h:106-153: SetSearcher + Tune(kernel,
// stop_condition)): a ConfigurationCount(5) budget runs five configurations, a RandomSearcher visits the space in another
// order than the DeterministicSearcher, and whatever was visited is valid and leaves y = A x.
void TestKttSearcherAndStopCondition() {
  cusp::csr_matrix<int, float, cusp::host_memory> Ah;
  cusp::gallery::poisson5pt(Ah, 64, 48);
  cusp::csr_matrix<int, float, cusp::device_memory> A(Ah);
  cusp::array1d<float, cusp::host_memory> xh(A.num_cols), yh(A.num_rows, 0.0f);
  for (size_t i = 0; i < xh.size(); ++i) xh[i] = (float)((i % 10) + 1);
  cusp::multiply(Ah, xh, yh);
  cusp::array1d<float, cusp::device_memory> x(xh), y(A.num_rows, 0.0f);
  std::ostringstream log;
  cusp::ktt::get_tuner().SetLoggingTarget(log);
  auto r5 = cusp::ktt::tune(A, x, y, std::nullopt, std::make_unique<::ktt::ConfigurationCount>(5));
  ASSERT_EQUAL(r5.size(), (size_t)5);
  ASSERT_EQUAL(y, yh);
  cusp::ktt::reset_tuning(A, x, y);
  auto det = cusp::ktt::tune(A, x, y, std::nullopt, std::make_unique<::ktt::ConfigurationCount>(12),
                             std::make_unique<::ktt::DeterministicSearcher>());
  cusp::ktt::reset_tuning(A, x, y);
  auto rnd = cusp::ktt::tune(A, x, y, std::nullopt, std::make_unique<::ktt::ConfigurationCount>(12),
                             std::make_unique<::ktt::RandomSearcher>(7));
  cusp::ktt::get_tuner().SetLoggingTarget(std::cerr);
  ASSERT_EQUAL(det.size(), (size_t)12);
  ASSERT_EQUAL(rnd.size(), (size_t)12);
  for (size_t i = 0; i < 5; ++i)  // the deterministic order is the space's: the first five are the budgeted run's
    ASSERT_TRUE(det[i].GetConfiguration().GetString() == r5[i].GetConfiguration().GetString());
  size_t different = 0;
  for (size_t i = 0; i < 12; ++i) {
    ASSERT_TRUE(rnd[i].IsValid());
    if (rnd[i].GetConfiguration().GetString() != det[i].GetConfiguration().GetString()) ++different;
  }
  ASSERT_TRUE(different >= 6);
  ASSERT_EQUAL(y, yh);
  ASSERT_TRUE(log.str().find("Explored configurations: 12 / 12") != std::string::npos);
  cusp::ktt::reset_tuning(A, x, y);
  // the other KTT stop conditions: a fraction of the space, a time budget that is already spent after the first
  // configuration, a duration target every valid configuration meets
  auto all = cusp::ktt::tune(A, x, y);
  cusp::ktt::reset_tuning(A, x, y);
  auto quarter = cusp::ktt::tune(A, x, y, std::nullopt, std::make_unique<::ktt::ConfigurationFraction>(0.25));
  ASSERT_TRUE(quarter.size() >= all.size() / 4 && quarter.size() <= all.size() / 4 + 1);
  cusp::ktt::reset_tuning(A, x, y);
  auto timed = cusp::ktt::tune(A, x, y, std::nullopt, std::make_unique<::ktt::TuningDuration>(0.0));
  ASSERT_TRUE(timed.size() <= 1);  // a budget that is spent before (or right after) the first configuration
  cusp::ktt::reset_tuning(A, x, y);
  auto fast = cusp::ktt::tune(A, x, y, std::nullopt, std::make_unique<::ktt::ConfigurationDuration>(1000.0));
  ASSERT_EQUAL(fast.size(), (size_t)1);
  ASSERT_EQUAL(y, yh);
  cusp::ktt::reset_tuning(A, x, y);
}
TEST_DEVICE(TestKttSearcherAndStopCondition)

// the stop conditions by themselves (host logic)
void TestKttStopConditionsHost() {
  b200sp_tune_result ok{};
  ok.status = B200SP_TUNE_OK;
  ok.milliseconds = 2.0;
  b200sp_tune_result bad = ok;
  bad.status = B200SP_TUNE_VALIDATION_FAILED;
  bad.milliseconds = 0.1;
  ::ktt::ConfigurationFraction f(0.5);
  f.Initialize(4);
  ASSERT_TRUE(!f.IsFulfilled());
  f.Update(::ktt::KernelResult("k", ok));
  ASSERT_TRUE(!f.IsFulfilled());
  f.Update(::ktt::KernelResult("k", ok));
  ASSERT_TRUE(f.IsFulfilled());
  ::ktt::ConfigurationDuration d(1.0);
  d.Initialize(10);
  d.Update(::ktt::KernelResult("k", bad));  // invalid results do not count
  ASSERT_TRUE(!d.IsFulfilled());
  d.Update(::ktt::KernelResult("k", ok));  // 2 ms > 1 ms
  ASSERT_TRUE(!d.IsFulfilled());
  ok.milliseconds = 0.5;
  d.Update(::ktt::KernelResult("k", ok));
  ASSERT_TRUE(d.IsFulfilled());
  ::ktt::TuningDuration t(3600.0);
  t.Initialize(10);
  ASSERT_TRUE(!t.IsFulfilled());
  ::ktt::ConfigurationCount c(3);
  c.Initialize(2);  // clamped to the size of the space
  c.Update(::ktt::KernelResult("k", ok));
  c.Update(::ktt::KernelResult("k", ok));
  ASSERT_TRUE(c.IsFulfilled());
}
TEST_HOST(TestKttStopConditionsHost)

// tune() on a COO matrix with scattered, skewed columns also inspects the column stream (hot-column plan): whether or
// not the plan wins the timing, cusp::multiply stays exact before, with and after it, and reset_tuning drops it.
void TestKttCooPlanLifecycle() {
  const int rows = 1 << 18;
  cusp::coo_matrix<int, float, cusp::host_memory> Ah;
  {
    // skewed columns: column = (hash % rows) >> (hash % 9): half of the gathers land in the low column range
    const size_t per_row = 20;
    Ah.resize(rows, rows, (size_t)rows * per_row);
    size_t k = 0;
    for (int i = 0; i < rows; ++i) {
      int prev = -1;
      for (size_t q = 0; q < per_row; ++q) {
        unsigned long long hsh = ((unsigned long long)i * 1315423911ull + q * 2654435761ull) ^ (q << 17);
        int c = (int)((hsh % (unsigned long long)rows) >> (hsh % 9));
        if (c <= prev) c = prev + 1;  // ascending within the row, no duplicates
        if (c >= rows) break;
        prev = c;
        Ah.row_indices[k] = i;
        Ah.column_indices[k] = c;
        Ah.values[k] = (float)((int)((hsh >> 7) % 5) - 2);
        ++k;
      }
    }
    Ah.row_indices.resize(k);
    Ah.column_indices.resize(k);
    Ah.values.resize(k);
    Ah.num_entries = k;
  }
  cusp::coo_matrix<int, float, cusp::device_memory> A(Ah);
  cusp::array1d<float, cusp::host_memory> xh(rows), yh(rows, 0.0f);
  for (int i = 0; i < rows; ++i) xh[i] = (float)((i % 7) - 3);
  cusp::multiply(Ah, xh, yh);
  cusp::array1d<float, cusp::device_memory> x(xh), y(rows, 5.0f);
  cusp::multiply(A, x, y);
  ASSERT_EQUAL(y, yh);
  std::ostringstream log;
  cusp::ktt::get_tuner().SetLoggingTarget(log);
  auto results = cusp::ktt::tune(A, x, y, std::nullopt, std::make_unique<::ktt::ConfigurationCount>(6));
  cusp::ktt::get_tuner().SetLoggingTarget(std::cerr);
  ASSERT_EQUAL(results.size(), (size_t)6);
  ASSERT_TRUE(cusp::ktt::detail::coo_plans().size() <= 1);
  const bool attached = cusp::ktt::detail::coo_plans().size() == 1;
  ASSERT_TRUE(attached == (log.str().find("coo plan attached") != std::string::npos));
  y = cusp::array1d<float, cusp::device_memory>(rows, -1.0f);
  cusp::multiply(A, x, y);  // through the attached plan when it won
  ASSERT_EQUAL(y, yh);
  cusp::ktt::reset_tuning(A, x, y);
  ASSERT_EQUAL(cusp::ktt::detail::coo_plans().size(), (size_t)0);
  cusp::multiply(A, x, y);
  ASSERT_EQUAL(y, yh);
  // the same operator as HYB with one ELL column (what the reference's split rule gives power-law graphs): the plan
  // goes on the COO tail
  cusp::hyb_matrix<int, float, cusp::host_memory> Hh;
  cusp::convert(Ah, Hh, (size_t)1);
  cusp::hyb_matrix<int, float, cusp::device_memory> H(Hh);
  ASSERT_TRUE(H.coo.num_entries >= ((size_t)1 << 22));
  y = cusp::array1d<float, cusp::device_memory>(rows, 3.0f);
  cusp::multiply(H, x, y);
  ASSERT_EQUAL(y, yh);
  cusp::ktt::tune(H, x, y, std::nullopt, std::make_unique<::ktt::ConfigurationCount>(4));
  ASSERT_TRUE(cusp::ktt::detail::coo_plans().size() <= 1);
  y = cusp::array1d<float, cusp::device_memory>(rows, -2.0f);
  cusp::multiply(H, x, y);
  ASSERT_EQUAL(y, yh);
  cusp::ktt::reset_tuning(H, x, y);
  ASSERT_EQUAL(cusp::ktt::detail::coo_plans().size(), (size_t)0);
  cusp::multiply(H, x, y);
  ASSERT_EQUAL(y, yh);
}
TEST_DEVICE(TestKttCooPlanLifecycle)
