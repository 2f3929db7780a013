// cusp/ktt/state.h — the process-wide "autotuning enabled" switch
// (reference: cusp/ktt/detail/ktt.inl:20-21 `inline bool is_enabled = true`,
// toggled by cusp::ktt::enable() / disable(), read by the ELL/DIA hook in
// cusp/system/detail/generic/multiply.inl:141-154).
#pragma once

namespace cusp {
namespace ktt {
namespace detail {
inline bool &enabled_flag() {
  static bool flag = true;
  return flag;
}
inline bool is_enabled() { return enabled_flag(); }
}  // namespace detail

inline void disable() { detail::enabled_flag() = false; }
inline void enable() { detail::enabled_flag() = true; }

}  // namespace ktt
}  // namespace cusp
