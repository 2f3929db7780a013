// cusp/ktt/ktt_types.h — the part of the KTT v2.1 vocabulary that appears in the
// cusp::ktt signatures (reference: cusp/ktt/ktt.h:35-101 — ::ktt::Tuner,
// KernelResult, KernelConfiguration, ReferenceComputation, StopCondition,
// Searcher).  KTT itself (NVRTC JIT, CUPTI, searchers) is replaced by the
// engine's own tuner behind b200sp_tune*/b200sp_cfg_space; these classes carry
// its results in the shapes user code already reads
// (testing/ktt.cu:47-140: IsValid, GetStatus, GetKernelName,
// GetConfiguration().GetPairs()[i].GetString(); StopCondition overrides).
// Define CUSP_B200_USE_REAL_KTT to compile against a real <Ktt.h> instead.
#pragma once
#ifndef CUSP_B200_USE_REAL_KTT
#include <chrono>
#include <cstdint>
#include <functional>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../b200sp.h"

namespace ktt {

typedef uint64_t KernelId;
typedef uint64_t KernelDefinitionId;
typedef uint64_t ArgumentId;
typedef uint64_t Nanoseconds;

enum class ResultStatus { Ok, CompilationFailed, ComputationFailed, DeviceLimitsExceeded, ValidationFailed };
enum class LoggingLevel { Off, Error, Warning, Info, Debug };
enum class TimeUnit { Nanoseconds, Microseconds, Milliseconds, Seconds };
enum class ComputeApi { OpenCL, CUDA, Vulkan };

typedef std::function<void(void *)> ReferenceComputation;

class ParameterPair {
 public:
  ParameterPair() : value_(0) {}
  ParameterPair(const std::string &name, uint64_t value) : name_(name), value_(value) {}
  const std::string &GetName() const { return name_; }
  uint64_t GetValue() const { return value_; }
  uint64_t GetValueUint() const { return value_; }
  int64_t GetValueInt() const { return (int64_t)value_; }
  double GetValueDouble() const { return (double)value_; }
  std::string GetString() const { return name_ + ": " + std::to_string(value_); }
  std::string GetValueString() const { return std::to_string(value_); }

 private:
  std::string name_;
  uint64_t value_;
};

// one point of the engine's tuning space (b200sp_cfg) seen as KTT parameter pairs
class KernelConfiguration {
 public:
  KernelConfiguration() { cfg_ = b200sp_cfg(); }
  explicit KernelConfiguration(const b200sp_cfg &c) : cfg_(c) {}
  explicit KernelConfiguration(const std::vector<ParameterPair> &pairs) {
    cfg_ = b200sp_cfg();
    for (const ParameterPair &p : pairs) set(p.GetName(), (int)p.GetValue());
  }
  std::vector<ParameterPair> GetPairs() const {
    return {ParameterPair("KERNEL", (uint64_t)cfg_.kernel),
            ParameterPair("BLOCK_SIZE", (uint64_t)cfg_.block_size),
            ParameterPair("THREADS_PER_ROW", (uint64_t)cfg_.threads_per_row),
            ParameterPair("UNROLL", (uint64_t)cfg_.unroll),
            ParameterPair("VECTOR_WIDTH", (uint64_t)cfg_.vector_width),
            ParameterPair("TILE_ROWS", (uint64_t)cfg_.tile_rows),
            ParameterPair("STAGES", (uint64_t)cfg_.stages),
            ParameterPair("CTAS_PER_SM", (uint64_t)cfg_.ctas_per_sm)};
  }
  bool IsValid() const { return true; }
  std::string GetString() const {
    std::string s;
    for (const ParameterPair &p : GetPairs()) s += (s.empty() ? "" : ", ") + p.GetString();
    return s;
  }
  const b200sp_cfg &cfg() const { return cfg_; }

 private:
  void set(const std::string &n, int v) {
    if (n == "KERNEL") cfg_.kernel = v;
    else if (n == "BLOCK_SIZE") cfg_.block_size = v;
    else if (n == "THREADS_PER_ROW") cfg_.threads_per_row = v;
    else if (n == "UNROLL") cfg_.unroll = v;
    else if (n == "VECTOR_WIDTH") cfg_.vector_width = v;
    else if (n == "TILE_ROWS") cfg_.tile_rows = v;
    else if (n == "STAGES") cfg_.stages = v;
    else if (n == "CTAS_PER_SM") cfg_.ctas_per_sm = v;
    else throw std::runtime_error("Unknown tuning parameter " + n);  // cuda/ktt/utils.h:108-127
  }
  b200sp_cfg cfg_;
};

class KernelResult {
 public:
  KernelResult() : status_(ResultStatus::ComputationFailed), duration_ns_(0), max_rel_error_(0) {}
  KernelResult(const std::string &kernel_name, const b200sp_tune_result &r)
      : name_(kernel_name), config_(r.cfg), duration_ns_((Nanoseconds)((double)r.milliseconds * 1e6)),
        max_rel_error_(r.max_rel_error) {
    switch (r.status) {
      case B200SP_TUNE_OK: status_ = ResultStatus::Ok; break;
      case B200SP_TUNE_LAUNCH_FAILED: status_ = ResultStatus::ComputationFailed; break;
      case B200SP_TUNE_VALIDATION_FAILED: status_ = ResultStatus::ValidationFailed; break;
      default: status_ = ResultStatus::DeviceLimitsExceeded; break;
    }
  }
  const std::string &GetKernelName() const { return name_; }
  const KernelConfiguration &GetConfiguration() const { return config_; }
  ResultStatus GetStatus() const { return status_; }
  bool IsValid() const { return status_ == ResultStatus::Ok; }
  Nanoseconds GetKernelDuration() const { return duration_ns_; }
  Nanoseconds GetTotalDuration() const { return duration_ns_; }
  Nanoseconds GetKernelOverhead() const { return 0; }
  bool HasRemainingProfilingRuns() const { return false; }
  double GetMaxRelativeError() const { return max_rel_error_; }

 private:
  std::string name_;
  KernelConfiguration config_;
  ResultStatus status_;
  Nanoseconds duration_ns_;
  double max_rel_error_;
};

class StopCondition {
 public:
  virtual ~StopCondition() = default;
  virtual bool IsFulfilled() const = 0;
  virtual void Initialize(const uint64_t configurationsCount) = 0;
  virtual void Update(const KernelResult &result) = 0;
  virtual std::string GetStatusString() const = 0;
};

class ConfigurationCount : public StopCondition {
 public:
  explicit ConfigurationCount(uint64_t count) : target_(count), seen_(0) {}
  bool IsFulfilled() const override { return seen_ >= target_; }
  void Initialize(const uint64_t total) override { seen_ = 0; if (target_ > total) target_ = total; }
  void Update(const KernelResult &) override { ++seen_; }
  std::string GetStatusString() const override {
    return "Explored configurations: " + std::to_string(seen_) + " / " + std::to_string(target_);
  }

 private:
  uint64_t target_, seen_;
};

// explored / total >= fraction
class ConfigurationFraction : public StopCondition {
 public:
  explicit ConfigurationFraction(double fraction) : fraction_(fraction < 0.0 ? 0.0 : (fraction > 1.0 ? 1.0 : fraction)), total_(0), seen_(0) {}
  bool IsFulfilled() const override { return total_ > 0 && (double)seen_ / (double)total_ >= fraction_; }
  void Initialize(const uint64_t total) override { total_ = total; seen_ = 0; }
  void Update(const KernelResult &) override { ++seen_; }
  std::string GetStatusString() const override {
    return "Explored configurations: " + std::to_string(seen_) + " / " + std::to_string(total_) + ", target fraction " +
           std::to_string(fraction_);
  }

 private:
  double fraction_;
  uint64_t total_, seen_;
};

// a valid configuration whose kernel time is at most `duration` (in milliseconds) has been found
class ConfigurationDuration : public StopCondition {
 public:
  explicit ConfigurationDuration(double duration_ms) : target_ms_(duration_ms), best_ms_(-1.0) {}
  bool IsFulfilled() const override { return best_ms_ >= 0.0 && best_ms_ <= target_ms_; }
  void Initialize(const uint64_t) override { best_ms_ = -1.0; }
  void Update(const KernelResult &r) override {
    if (!r.IsValid()) return;
    const double ms = (double)r.GetKernelDuration() * 1e-6;
    if (best_ms_ < 0.0 || ms < best_ms_) best_ms_ = ms;
  }
  std::string GetStatusString() const override {
    return "Best duration: " + (best_ms_ < 0.0 ? std::string("none") : std::to_string(best_ms_)) + " ms, target " +
           std::to_string(target_ms_) + " ms";
  }

 private:
  double target_ms_, best_ms_;
};

// wall-clock budget for the whole search, in seconds
class TuningDuration : public StopCondition {
 public:
  explicit TuningDuration(double seconds) : target_s_(seconds), start_(std::chrono::steady_clock::now()) {}
  bool IsFulfilled() const override { return elapsed() >= target_s_; }
  void Initialize(const uint64_t) override { start_ = std::chrono::steady_clock::now(); }
  void Update(const KernelResult &) override {}
  std::string GetStatusString() const override {
    return "Tuning time: " + std::to_string(elapsed()) + " / " + std::to_string(target_s_) + " s";
  }

 private:
  double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count(); }
  double target_s_;
  std::chrono::steady_clock::time_point start_;
};

// Searchers decide the ORDER in which cusp::ktt::tune visits the space (b200sp_tune_ex takes it as an index list):
// DeterministicSearcher = the space's own order, RandomSearcher = a uniformly random permutation (optionally seeded).
// A user searcher overrides Order().
class Searcher {
 public:
  virtual ~Searcher() = default;
  virtual std::vector<int64_t> Order(int64_t space_size) const {
    std::vector<int64_t> o((size_t)space_size);
    for (int64_t i = 0; i < space_size; ++i) o[(size_t)i] = i;
    return o;
  }
};
class DeterministicSearcher : public Searcher {};
class RandomSearcher : public Searcher {
 public:
  explicit RandomSearcher(uint64_t seed = 0) : seed_(seed) {}
  std::vector<int64_t> Order(int64_t space_size) const override {
    std::vector<int64_t> o = Searcher::Order(space_size);
    uint64_t s = seed_ ? seed_ : 0x9e3779b97f4a7c15ull;
    for (int64_t i = space_size - 1; i > 0; --i) {  // Fisher-Yates with splitmix64
      s += 0x9e3779b97f4a7c15ull;
      uint64_t z = s;
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      z ^= z >> 31;
      std::swap(o[(size_t)i], o[(size_t)(z % (uint64_t)(i + 1))]);
    }
    return o;
  }

 private:
  uint64_t seed_;
};

// stand-in for ::ktt::Tuner: logging controls + configuration factory.  The
// tuning state itself lives in the engine handle (b200sp_tune_*).
class Tuner {
 public:
  Tuner() : log_(&std::cerr), level_(LoggingLevel::Info) {}
  void SetLoggingTarget(std::ostream &os) { log_ = &os; }
  static void SetLoggingLevel(LoggingLevel) {}
  void SetTimeUnit(TimeUnit) {}
  void SetSearcher(KernelId, std::unique_ptr<Searcher>) {}
  void SetCompilerOptions(const std::string &) {}
  KernelConfiguration CreateConfiguration(KernelId, const std::vector<ParameterPair> &pairs) const {
    return KernelConfiguration(pairs);
  }
  std::ostream &log() { return *log_; }
  // ClearData(kernel_id) is routed to b200sp_tune_reset by cusp::ktt::reset_tuning

 private:
  std::ostream *log_;
  LoggingLevel level_;
};

}  // namespace ktt
#else
#include <Ktt.h>
#endif
