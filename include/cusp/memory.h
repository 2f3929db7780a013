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

// compatible with every space (cusp/memory.h: any_memory) and the combination rule used by algorithms
// that mix operands (cusp/memory.h: minimum_space<T1,T2,T3>; mixing host and device has no type)
struct any_memory : execution_policy<any_memory> {};
namespace detail {
template <typename T1, typename T2>
struct minimum_space2 {};
template <typename T>
struct minimum_space2<T, T> { typedef T type; };
template <typename T>
struct minimum_space2<T, any_memory> { typedef T type; };
template <typename T>
struct minimum_space2<any_memory, T> { typedef T type; };
template <>
struct minimum_space2<any_memory, any_memory> { typedef any_memory type; };
}  // namespace detail
template <typename T1, typename T2 = any_memory, typename T3 = any_memory>
struct minimum_space {
  typedef typename detail::minimum_space2<typename detail::minimum_space2<T1, T2>::type, T3>::type type;
};

namespace detail {
// Dispatch on the derived policy, as thrust/cusp do (cusp/detail/multiply.inl:27-46: the public
// entry point calls `multiply(derived_cast(exec), ...)` unqualified): a user policy
//     struct my_system : cusp::execution_policy<my_system> { ... };
// with its own overload  `void multiply(my_system&, const A&, const B&, C&)`  in its namespace is
// reached through argument-dependent lookup; without such an overload the call lands on the
// library's implementation in `adl_default` (taken by `execution_policy<P>&`, a worse match than a
// user's exact `my_system&`, a better one than the public `const execution_policy<P>&` entry).
template <typename P>
P &derived_cast(const execution_policy<P> &exec) {
  return const_cast<P &>(static_cast<const P &>(exec));
}
}  // namespace detail

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
