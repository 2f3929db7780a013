// cusp/detail/engine.h — the one place where the compatibility headers meet the
// C ABI (include/b200sp.h).  A thread-local engine handle is created on first
// use; every status other than B200SP_OK becomes a cusp exception
// (B200SP_INVALID_INPUT -> cusp::invalid_input_exception, else
// cusp::runtime_exception), mirroring cusp/exception.h.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../b200sp.h"
#include "../exception.h"
#include "../memory.h"

namespace cusp {
namespace detail {

struct engine_holder {
  b200sp_handle h = nullptr;
  ~engine_holder() {
    if (h) b200sp_destroy(h);
  }
};

inline b200sp_handle engine() {
  static thread_local engine_holder holder;
  if (!holder.h) {
    b200sp_status s = b200sp_create(&holder.h);
    if (s != B200SP_OK) throw cusp::runtime_exception(std::string("b200sp_create: ") + b200sp_last_error_string(nullptr));
  }
  return holder.h;
}

inline void check(b200sp_status s) {
  if (s == B200SP_OK) return;
  std::string msg = b200sp_last_error_string(engine());
  if (s == B200SP_INVALID_INPUT) throw cusp::invalid_input_exception(msg);
  throw cusp::runtime_exception(msg);
}

inline void cuda_check(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw cusp::runtime_exception(std::string(what) + ": " + cudaGetErrorString(e));
}

// stream used by the default device policy: the legacy default stream, like
// cusp::device_memory == cusp::cuda::par (cusp/system/cuda/detail/par.h:36-54)
inline b200sp_stream &current_stream() {
  static thread_local b200sp_stream s = nullptr;
  return s;
}

template <typename T>
struct dtype_of;
template <>
struct dtype_of<float> {
  static const b200sp_dtype value = B200SP_F32;
};
template <>
struct dtype_of<double> {
  static const b200sp_dtype value = B200SP_F64;
};

}  // namespace detail

namespace cuda {
// cusp::cuda::par.on(stream) returns a policy VALUE that carries the stream (like the reference's
// thrust::cuda::par.on): only the call it is passed to runs on that stream — the entry points that take a leading
// policy install it for the duration of the call (detail::stream_scope) and restore the thread's previous stream.
struct stream_policy : cusp::execution_policy<stream_policy> {
  cudaStream_t stream = nullptr;
};
struct par_t : cusp::execution_policy<par_t> {
  stream_policy on(cudaStream_t s) const {
    stream_policy p;
    p.stream = s;
    return p;
  }
};
static const par_t par{};
}  // namespace cuda

namespace detail {
template <typename Policy>
struct stream_scope {
  explicit stream_scope(const Policy &) {}
};
template <>
struct stream_scope<cusp::cuda::stream_policy> {
  b200sp_stream saved;
  explicit stream_scope(const cusp::cuda::stream_policy &p) : saved(current_stream()) {
    current_stream() = (b200sp_stream)p.stream;
  }
  ~stream_scope() { current_stream() = saved; }
  stream_scope(const stream_scope &) = delete;
  stream_scope &operator=(const stream_scope &) = delete;
};
}  // namespace detail
}  // namespace cusp
