// cusp/ktt/ktt.h — the fork's autotuning API
// (reference: cusp/ktt/ktt.h:14-127, cusp/ktt/detail/ktt.inl:20-142,
// cusp/system/cuda/ktt/multiply.h:27-153).
//
//   enable() / disable()          the switch read by plain cusp::multiply for ELL / DIA
//   get_tuner()                   process-wide ::ktt::Tuner stand-in (logging, CreateConfiguration)
//   multiply(A, x, y)             one step of dynamic tuning            -> b200sp_tune_step
//   multiply(A, x, y, conf)       run exactly `conf`                    -> b200sp_spmv(cfg)
//   tune(A, x, y[, ref, stop, searcher])  offline tuning with validation -> b200sp_tune_ex; for COO matrices it also
//                                 inspects the column stream and attaches a hot-column plan when that wins
//                                 (b200sp_coo_plan_*: tuning is tied to the matrix it ran on, like the reference's)
//   reset_tuning(A, x, y)         forget results for A's kernel         -> b200sp_tune_reset (+ drop A's plan)
//
// Every configuration is a precompiled sm_100a instantiation: a tuning step costs
// a launch, not an NVRTC compile.  Works for csr / coo / ell / dia / hyb / ellr
// matrices in device_memory (the reference's KTT glue covers csr, coo, ell, dia, ellr).
#pragma once
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "../array1d.h"
#include "../detail/descriptor.h"
#include "ktt_types.h"
#include "state.h"

namespace cusp {

namespace system {
namespace cuda {
namespace ktt {
// cuda/ktt/kernel.h: what get_kernel() hands back (main.cu:448,510 read kernel_id)
struct kernel_context {
  ::ktt::KernelId kernel_id = 0;
  std::string name;
};
inline const char *format_name(b200sp_format f) {
  switch (f) {
    case B200SP_FMT_CSR: return "csr_spmv";
    case B200SP_FMT_ELL: return "ell_spmv";
    case B200SP_FMT_DIA: return "dia_spmv";
    case B200SP_FMT_COO: return "coo_spmv";
    case B200SP_FMT_HYB: return "hyb_spmv";
    default: return "ellr_spmv";
  }
}
template <typename Matrix, typename V1, typename V2>
kernel_context get_kernel(::ktt::Tuner &, const Matrix &A, const V1 &, const V2 &) {
  b200sp_matrix d = cusp::detail::describe(A);
  kernel_context k;
  k.kernel_id = (::ktt::KernelId)d.format * 2 + (::ktt::KernelId)d.dtype;
  k.name = format_name(d.format);
  return k;
}
}  // namespace ktt
}  // namespace cuda
}  // namespace system

namespace ktt {

inline ::ktt::Tuner &get_tuner() {
  static ::ktt::Tuner tuner;
  return tuner;
}

namespace detail {
using cusp::detail::check;
using cusp::detail::current_stream;
using cusp::detail::describe;
using cusp::detail::engine;
using cusp::detail::raw_ptr;

template <typename Matrix, typename V1, typename V2>
void require_device(const Matrix &A, const V1 &x, const V2 &y) {
  static_assert(cusp::detail::abi_matrix<Matrix>::value,
                "cusp::ktt: device_memory sparse matrix with 32-bit indices and float/double values required");
  static_assert(std::is_same<typename V1::memory_space, cusp::device_memory>::value &&
                    std::is_same<typename V2::memory_space, cusp::device_memory>::value,
                "cusp::ktt: x and y must be device arrays");
  if (A.num_cols != x.size() || A.num_rows != y.size())
    throw cusp::invalid_input_exception("cusp::ktt: matrix and vector dimensions do not match");
}
}  // namespace detail

namespace detail {
// hot-column plans attached by tune() for COO matrices (and the COO part of HYB), by column-array address
inline std::map<const void *, b200sp_coo_plan> &coo_plans() {
  static std::map<const void *, b200sp_coo_plan> plans;
  return plans;
}
inline void drop_coo_plan(const void *column_indices) {
  auto it = coo_plans().find(column_indices);
  if (it == coo_plans().end()) return;
  b200sp_coo_plan_destroy(engine(), it->second);  // detaches too
  coo_plans().erase(it);
}
// Tuning a matrix whose product is gather-bound also inspects its column stream: a plan (b200sp_coo_plan_create) is
// built, timed against the best configuration just found and attached to the engine when it wins — from then on
// cusp::multiply on exactly this matrix keeps the hot columns of x in shared memory.  The plan refers to the
// matrix's index arrays by address: reset_tuning (or tuning again) before changing the sparsity pattern in place.
template <typename T>
void try_attach_coo_plan(int64_t rows, int64_t cols, int64_t nnz, const int *Ai, const int *Aj, const T *Ax, const T *x, T *y,
                         const b200sp_matrix &d) {
  drop_coo_plan(Aj);
  if (nnz < (int64_t)1 << 22) return;  // small products are not gather-bound
  b200sp_coo_plan plan = nullptr;
  const b200sp_dtype dt = std::is_same<T, float>::value ? B200SP_F32 : B200SP_F64;
  if (b200sp_coo_plan_create(engine(), current_stream(), rows, cols, nnz, Ai, Aj, dt, 0, &plan) != B200SP_OK || !plan) return;
  cudaStream_t st = (cudaStream_t)current_stream();
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventCreate(&e2);
  const int reps = 3;
  b200sp_spmv(engine(), current_stream(), &d, x, y, 0, nullptr);  // warm (cached winner)
  cudaEventRecord(e0, st);
  for (int k = 0; k < reps; ++k) b200sp_spmv(engine(), current_stream(), &d, x, y, 0, nullptr);
  cudaEventRecord(e1, st);
  b200sp_status ps = B200SP_OK;
  for (int k = 0; k < reps && ps == B200SP_OK; ++k)
    ps = std::is_same<T, float>::value
             ? b200sp_spmv_coo_plan_f32(engine(), current_stream(), plan, (const float *)Ax, (const float *)x, (float *)y, 0, nullptr)
             : b200sp_spmv_coo_plan_f64(engine(), current_stream(), plan, (const double *)Ax, (const double *)x, (double *)y, 0, nullptr);
  cudaEventRecord(e2, st);
  cudaEventSynchronize(e2);
  float ms_plain = 0.f, ms_plan = 0.f;
  cudaEventElapsedTime(&ms_plain, e0, e1);
  cudaEventElapsedTime(&ms_plan, e1, e2);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaEventDestroy(e2);
  if (ps == B200SP_OK && ms_plan < 0.97f * ms_plain && b200sp_coo_plan_attach(engine(), plan) == B200SP_OK) {
    coo_plans()[Aj] = plan;
    get_tuner().log() << "coo plan attached: " << ms_plan / reps << " ms against " << ms_plain / reps << " ms" << std::endl;
  } else {
    b200sp_coo_plan_destroy(engine(), plan);
  }
}
template <typename Matrix, typename V1, typename V2>
void after_tune(const Matrix &A, const V1 &x, V2 &y, const b200sp_matrix &d, cusp::coo_format) {
  try_attach_coo_plan((int64_t)A.num_rows, (int64_t)A.num_cols, (int64_t)A.num_entries, raw_ptr(A.row_indices),
                      raw_ptr(A.column_indices), raw_ptr(A.values), raw_ptr(x), raw_ptr(y), d);
}
// hyb_matrix: the plan goes on the COO tail (spmv_hyb reaches it through the tail's column array); timed on the tail alone
template <typename Matrix, typename V1, typename V2>
void after_tune(const Matrix &A, const V1 &x, V2 &y, const b200sp_matrix &, cusp::hyb_format) {
  if ((int64_t)A.coo.num_entries < ((int64_t)1 << 22)) {
    drop_coo_plan(raw_ptr(A.coo.column_indices));
    return;
  }
  const b200sp_matrix dc = describe(A.coo);
  try_attach_coo_plan((int64_t)A.num_rows, (int64_t)A.num_cols, (int64_t)A.coo.num_entries, raw_ptr(A.coo.row_indices),
                      raw_ptr(A.coo.column_indices), raw_ptr(A.coo.values), raw_ptr(x), raw_ptr(y), dc);
}
template <typename Matrix, typename V1, typename V2, typename Format>
void after_tune(const Matrix &, const V1 &, V2 &, const b200sp_matrix &, Format) {}
template <typename Matrix>
void before_reset(const Matrix &A, cusp::coo_format) {
  drop_coo_plan(raw_ptr(A.column_indices));
}
template <typename Matrix>
void before_reset(const Matrix &A, cusp::hyb_format) {
  drop_coo_plan(raw_ptr(A.coo.column_indices));
}
template <typename Matrix, typename Format>
void before_reset(const Matrix &, Format) {}
}  // namespace detail

// one step of dynamic autotuning (cuda/ktt/multiply.h:56-77)
template <typename Matrix, typename V1, typename V2>
::ktt::KernelResult multiply(const Matrix &A, const V1 &x, V2 &y) {
  detail::require_device(A, x, y);
  b200sp_matrix d = detail::describe(A);
  b200sp_tune_result r;
  detail::check(
      b200sp_tune_step(detail::engine(), detail::current_stream(), &d, detail::raw_ptr(x), detail::raw_ptr(y), &r));
  return ::ktt::KernelResult(cusp::system::cuda::ktt::format_name(d.format), r);
}

// run one given configuration (cuda/ktt/multiply.h:79-104); timed with CUDA events
template <typename Matrix, typename V1, typename V2>
::ktt::KernelResult multiply(const Matrix &A, const V1 &x, V2 &y, const ::ktt::KernelConfiguration &configuration,
                             bool run_with_profiling = false) {
  (void)run_with_profiling;  // CUPTI counters: use ncu on the precompiled kernels instead
  detail::require_device(A, x, y);
  b200sp_matrix d = detail::describe(A);
  cudaStream_t s = (cudaStream_t)detail::current_stream();
  cudaEvent_t e0, e1;
  cusp::detail::cuda_check(cudaEventCreate(&e0), "cudaEventCreate");
  cusp::detail::cuda_check(cudaEventCreate(&e1), "cudaEventCreate");
  cudaEventRecord(e0, s);
  b200sp_status st =
      b200sp_spmv(detail::engine(), detail::current_stream(), &d, detail::raw_ptr(x), detail::raw_ptr(y), 0,
                  &configuration.cfg());
  cudaEventRecord(e1, s);
  cudaEventSynchronize(e1);
  b200sp_tune_result r;
  r.cfg = configuration.cfg();
  r.max_rel_error = 0;
  r.milliseconds = 0;
  cudaEventElapsedTime(&r.milliseconds, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  r.status = st == B200SP_OK ? B200SP_TUNE_OK : (st == B200SP_INVALID_INPUT ? B200SP_TUNE_UNSUPPORTED
                                                                            : B200SP_TUNE_LAUNCH_FAILED);
  return ::ktt::KernelResult(cusp::system::cuda::ktt::format_name(d.format), r);
}

// offline tuning over the whole space, each configuration validated
// (cuda/ktt/multiply.h:106-153).  reference_computation fills a host buffer of
// y.size() values with the expected result; without it the engine-default
// configuration's output is the reference.
template <typename Matrix, typename V1, typename V2>
std::vector<::ktt::KernelResult> tune(const Matrix &A, const V1 &x, V2 &y,
                                      std::optional<::ktt::ReferenceComputation> reference_computation = std::nullopt,
                                      std::unique_ptr<::ktt::StopCondition> stop_condition = nullptr,
                                      std::unique_ptr<::ktt::Searcher> searcher = nullptr) {
  typedef typename V2::value_type T;
  detail::require_device(A, x, y);
  b200sp_matrix d = detail::describe(A);
  const int64_t space = b200sp_cfg_space(d.format, d.dtype, nullptr, 0);
  std::vector<b200sp_tune_result> raw((size_t)std::max<int64_t>(space, 1));
  cusp::array1d<T, cusp::device_memory> y_ref;
  const void *ref_ptr = nullptr;
  if (reference_computation) {
    std::vector<T> host_ref(y.size());
    (*reference_computation)((void *)host_ref.data());
    y_ref = cusp::array1d<T, cusp::device_memory>(host_ref.begin(), host_ref.end());
    ref_ptr = detail::raw_ptr(y_ref);
  }
  // the searcher fixes the order of the visit, the stop condition is consulted after every configuration while the
  // search runs (cuda/ktt/multiply.h:129-146: SetSearcher, Tune(kernel, stop_condition))
  std::vector<int64_t> order = searcher ? searcher->Order(space) : std::vector<int64_t>();
  struct Live {
    ::ktt::StopCondition *stop;
    std::vector<::ktt::KernelResult> *results;
    const char *name;
  };
  std::vector<::ktt::KernelResult> results;
  const char *name = cusp::system::cuda::ktt::format_name(d.format);
  Live live{stop_condition.get(), &results, name};
  if (stop_condition) stop_condition->Initialize((uint64_t)(searcher ? order.size() : (size_t)space));
  auto on_result = [](const b200sp_tune_result *r, void *user) -> int {
    Live *L = static_cast<Live *>(user);
    // points whose resources (smem ring = stages x K x tile) do not fit this matrix are outside its space, the way KTT
    // constraints drop configurations before tuning: they are neither reported nor counted
    if (r->status == B200SP_TUNE_UNSUPPORTED) return 0;
    L->results->emplace_back(L->name, *r);
    if (!L->stop) return 0;
    L->stop->Update(L->results->back());
    return L->stop->IsFulfilled() ? 1 : 0;
  };
  int64_t n = 0;
  b200sp_cfg best;
  const double tol = std::is_same<T, float>::value ? 1e-5 : 1e-12;  // north-star parity bound
  if (!(stop_condition && stop_condition->IsFulfilled()))
    detail::check(b200sp_tune_ex(detail::engine(), detail::current_stream(), &d, detail::raw_ptr(x), detail::raw_ptr(y),
                                 ref_ptr, tol, 3, searcher ? order.data() : nullptr, searcher ? (int64_t)order.size() : 0,
                                 on_result, &live, raw.data(), (int64_t)raw.size(), &n, &best));
  if (stop_condition) get_tuner().log() << stop_condition->GetStatusString() << std::endl;
  detail::after_tune(A, x, y, d, typename Matrix::format());
  return results;
}

template <typename MatrixType, typename V1, typename V2>
void reset_tuning(const MatrixType &A, const V1 &, V2 &) {
  detail::before_reset(A, typename MatrixType::format());
  b200sp_matrix d = detail::describe(A);
  detail::check(b200sp_tune_reset(detail::engine(), &d));
}

}  // namespace ktt
}  // namespace cusp
