// cusp/memory.h — memory-space and format tags (cusp/memory.h, cusp/format.h).
// As in the reference, the memory-space tags double as execution policies
// (cusp/iterator/detail/device_system_tag.h:29: device_memory is the cuda `par_t`
// policy): every algorithm has an overload with a leading policy argument, and
// user policies derive from cusp::execution_policy<Derived>.
#pragma once

namespace cusp {

template <typename DerivedPolicy>
struct execution_policy {};

struct host_memory : execution_policy<host_memory> {};
struct device_memory : execution_policy<device_memory> {};

struct known_format {};
struct unknown_format {};
struct dense_format : known_format {};
struct array1d_format : dense_format {};
struct array2d_format : dense_format {};
struct sparse_format : known_format {};
struct coo_format : sparse_format {};
struct csr_format : sparse_format {};
struct dia_format : sparse_format {};
struct ell_format : sparse_format {};
struct hyb_format : sparse_format {};

struct row_major {};
struct column_major {};

}  // namespace cusp
