// cusp/array1d.h — cusp::array1d<T,MemorySpace>, cusp::array1d_view<Iterator>
// (reference: cusp/array1d.h:98-517), written without Thrust: host arrays sit on
// std::vector, device arrays on cudaMalloc.  Device iterators are tagged
// pointers (cusp::device_ptr<T>, the role thrust::device_ptr<T> plays in the
// reference); element access on device arrays goes through a proxy reference
// (one cudaMemcpy per access, as thrust::device_reference).  When Thrust headers
// are on the include path, thrust::device_ptr<T> is accepted wherever a
// cusp::device_ptr<T> is (examples/Views/cg_raw.cu wraps raw pointers that way).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <iterator>
#include <type_traits>
#include <vector>

#include "detail/engine.h"
#include "memory.h"

#if !defined(CUSP_B200_NO_THRUST) && defined(__has_include)
#if __has_include(<thrust/device_ptr.h>) && defined(__CUDACC__)
#include <thrust/device_ptr.h>
#include <thrust/device_reference.h>
#define CUSP_B200_HAVE_THRUST 1
#endif
#endif

namespace cusp {

template <typename T, typename MemorySpace>
class array1d;
template <typename Iterator>
class array1d_view;

#ifdef CUSP_B200_HAVE_THRUST
// Compiled by nvcc with Thrust at hand (what a user of the reference has): device arrays iterate with
// thrust::device_ptr and hand out thrust::device_reference, exactly like the reference's containers, so
// user code that runs Thrust algorithms on them (thrust::fill(x.begin(), x.end(), v)) or takes
// thrust::raw_pointer_cast(&x[0]) compiles unchanged (examples/Solvers/gmres.cu, LinearOperator/stencil.cu).
template <typename T>
using device_ptr = thrust::device_ptr<T>;
using thrust::device_pointer_cast;
using thrust::raw_pointer_cast;
namespace detail {
template <typename T>
using device_reference = thrust::device_reference<T>;
template <typename T>
device_reference<T> make_device_reference(T *p) {
  return device_reference<T>(thrust::device_ptr<T>(p));
}
}  // namespace detail
#else
template <typename T>
class device_ptr;

namespace detail {

template <typename T>
class device_reference {
 public:
  typedef typename std::remove_const<T>::type value_type;
  explicit device_reference(T *p) : p_(p) {}
  operator value_type() const {
    value_type v;
    cuda_check(cudaMemcpy(&v, p_, sizeof(T), cudaMemcpyDeviceToHost), "device_reference read");
    return v;
  }
  device_reference &operator=(const value_type &v) {
    cuda_check(cudaMemcpy(const_cast<value_type *>(p_), &v, sizeof(T), cudaMemcpyHostToDevice),
               "device_reference write");
    return *this;
  }
  device_reference &operator=(const device_reference &o) { return *this = (value_type)o; }
  device_ptr<T> operator&() const;

 private:
  T *p_;
};

}  // namespace detail

// tagged device pointer: the iterator of device arrays and views
template <typename T>
class device_ptr {
 public:
  typedef typename std::remove_const<T>::type value_type;
  typedef std::ptrdiff_t difference_type;
  typedef T *pointer;
  typedef detail::device_reference<T> reference;
  typedef std::random_access_iterator_tag iterator_category;

  device_ptr() : p_(nullptr) {}
  explicit device_ptr(T *p) : p_(p) {}
  template <typename U, typename = typename std::enable_if<std::is_convertible<U *, T *>::value>::type>
  device_ptr(const device_ptr<U> &o) : p_(o.get()) {}
  T *get() const { return p_; }
  reference operator*() const { return reference(p_); }
  reference operator[](std::ptrdiff_t i) const { return reference(p_ + i); }
  device_ptr operator+(std::ptrdiff_t n) const { return device_ptr(p_ + n); }
  device_ptr operator-(std::ptrdiff_t n) const { return device_ptr(p_ - n); }
  std::ptrdiff_t operator-(const device_ptr &o) const { return p_ - o.p_; }
  device_ptr &operator++() { ++p_; return *this; }
  device_ptr operator++(int) { device_ptr t(*this); ++p_; return t; }
  device_ptr &operator--() { --p_; return *this; }
  device_ptr &operator+=(std::ptrdiff_t n) { p_ += n; return *this; }
  bool operator==(const device_ptr &o) const { return p_ == o.p_; }
  bool operator!=(const device_ptr &o) const { return p_ != o.p_; }
  bool operator<(const device_ptr &o) const { return p_ < o.p_; }

 private:
  T *p_;
};

template <typename T>
device_ptr<T> device_pointer_cast(T *p) {
  return device_ptr<T>(p);
}
template <typename T>
T *raw_pointer_cast(const device_ptr<T> &p) {
  return p.get();
}
template <typename T>
T *raw_pointer_cast(T *p) {
  return p;
}

namespace detail {

template <typename T>
device_ptr<T> device_reference<T>::operator&() const {
  return device_ptr<T>(p_);
}
template <typename T>
device_reference<T> make_device_reference(T *p) {
  return device_reference<T>(p);
}
}  // namespace detail
#endif  // CUSP_B200_HAVE_THRUST

namespace detail {

// iterator -> (value type, memory space, raw pointer).  Raw pointers are host
// iterators, exactly like in the reference (thrust's host system).
template <typename Iterator>
struct iter_traits;
template <typename T>
struct iter_traits<T *> {
  typedef typename std::remove_const<T>::type value_type;
  typedef T element_type;
  typedef host_memory memory_space;
  static T *raw(T *p) { return p; }
};
template <typename T>
struct iter_traits<cusp::device_ptr<T>> {
  typedef typename std::remove_const<T>::type value_type;
  typedef T element_type;
  typedef device_memory memory_space;
  static T *raw(const cusp::device_ptr<T> &p) { return p.get(); }
};
// (with Thrust, cusp::device_ptr IS thrust::device_ptr: the specialization above covers views over
//  thrust::device_ptr built from raw cudaMalloc pointers, examples/Views/cg_raw.cu)

template <typename Space1, typename Space2>
struct copy_kind;
template <>
struct copy_kind<host_memory, host_memory> {
  static const cudaMemcpyKind value = cudaMemcpyHostToHost;
};
template <>
struct copy_kind<host_memory, device_memory> {  // src host -> dst device
  static const cudaMemcpyKind value = cudaMemcpyHostToDevice;
};
template <>
struct copy_kind<device_memory, host_memory> {
  static const cudaMemcpyKind value = cudaMemcpyDeviceToHost;
};
template <>
struct copy_kind<device_memory, device_memory> {
  static const cudaMemcpyKind value = cudaMemcpyDeviceToDevice;
};

template <typename T, typename SrcSpace, typename DstSpace>
inline void raw_copy(const T *src, T *dst, size_t n) {
  if (n == 0) return;
  if (std::is_same<SrcSpace, host_memory>::value && std::is_same<DstSpace, host_memory>::value)
    std::memcpy(dst, src, n * sizeof(T));
  else
    cuda_check(cudaMemcpy(dst, src, n * sizeof(T), copy_kind<SrcSpace, DstSpace>::value), "cusp array copy");
}

// converting copy through the host (different value types)
template <typename T, typename U, typename SrcSpace, typename DstSpace>
inline void convert_copy(const U *src, T *dst, size_t n) {
  if (n == 0) return;
  std::vector<U> hs(n);
  raw_copy<U, SrcSpace, host_memory>(src, hs.data(), n);
  std::vector<T> hd(n);
  for (size_t i = 0; i < n; ++i) hd[i] = static_cast<T>(hs[i]);
  raw_copy<T, host_memory, DstSpace>(hd.data(), dst, n);
}

template <typename T, typename U, typename SrcSpace, typename DstSpace>
inline void any_copy(const U *src, T *dst, size_t n) {
  typedef typename std::remove_const<U>::type UU;
  if (std::is_same<UU, T>::value)
    raw_copy<T, SrcSpace, DstSpace>(reinterpret_cast<const T *>(src), dst, n);
  else
    convert_copy<T, UU, SrcSpace, DstSpace>(src, dst, n);
}

// raw pointer of any array1d / array1d_view (possibly const)
template <typename Array>
auto raw_ptr(Array &a) -> decltype(iter_traits<decltype(a.begin())>::raw(a.begin())) {
  return iter_traits<decltype(a.begin())>::raw(a.begin());
}

template <typename Array>
std::vector<typename Array::value_type> to_host_vector(const Array &a) {
  std::vector<typename Array::value_type> h(a.size());
  raw_copy<typename Array::value_type, typename Array::memory_space, host_memory>(raw_ptr(a), h.data(), a.size());
  return h;
}

}  // namespace detail

// ---------------------------------------------------------------------------
// host array
// ---------------------------------------------------------------------------
template <typename T>
class array1d<T, host_memory> {
 public:
  typedef T value_type;
  typedef host_memory memory_space;
  typedef array1d_format format;
  typedef T *iterator;
  typedef const T *const_iterator;
  typedef T *pointer;
  typedef T &reference;
  typedef size_t size_type;
  typedef array1d container;
  typedef array1d_view<iterator> view;
  typedef array1d_view<const_iterator> const_view;
  template <typename Space>
  struct rebind {
    typedef array1d<T, Space> type;
  };

  array1d() {}
  explicit array1d(size_t n) : v_(n) {}
  array1d(size_t n, const T &value) : v_(n, value) {}
  array1d(const array1d &o) : v_(o.v_) {}
  template <typename U, typename Space>
  array1d(const array1d<U, Space> &o) {
    assign_from(o);
  }
  template <typename It>
  array1d(const array1d_view<It> &o) {
    assign_from(o);
  }
  template <typename It, typename = typename std::enable_if<!std::is_integral<It>::value>::type>
  array1d(It first, It last) : v_(first, last) {}
  // std::vector interoperability (testing/array1d.cu:66-140)
  template <typename U, typename Alloc>
  array1d(const std::vector<U, Alloc> &v) : v_(v.begin(), v.end()) {}
  template <typename U, typename Alloc>
  array1d &operator=(const std::vector<U, Alloc> &v) {
    v_.assign(v.begin(), v.end());
    return *this;
  }

  array1d &operator=(const array1d &o) {
    v_ = o.v_;
    return *this;
  }
  template <typename U, typename Space>
  array1d &operator=(const array1d<U, Space> &o) {
    assign_from(o);
    return *this;
  }
  template <typename It>
  array1d &operator=(const array1d_view<It> &o) {
    assign_from(o);
    return *this;
  }

  size_t size() const { return v_.size(); }
  size_t capacity() const { return v_.capacity(); }
  bool empty() const { return v_.empty(); }
  void resize(size_t n) { v_.resize(n); }
  void resize(size_t n, const T &value) { v_.resize(n, value); }
  void reserve(size_t n) { v_.reserve(n); }
  void push_back(const T &x) { v_.push_back(x); }
  void clear() { v_.clear(); }
  void swap(array1d &o) { v_.swap(o.v_); }
  void assign(size_t n, const T &value) { v_.assign(n, value); }
  T &operator[](size_t i) { return v_[i]; }
  const T &operator[](size_t i) const { return v_[i]; }
  T &front() { return v_.front(); }
  T &back() { return v_.back(); }
  const T &back() const { return v_.back(); }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }
  iterator begin() { return v_.data(); }
  iterator end() { return v_.data() + v_.size(); }
  const_iterator begin() const { return v_.data(); }
  const_iterator end() const { return v_.data() + v_.size(); }
  view subarray(size_t start, size_t num) { return view(begin() + start, begin() + start + num); }
  const_view subarray(size_t start, size_t num) const {
    return const_view(begin() + start, begin() + start + num);
  }

 private:
  template <typename A>
  void assign_from(const A &o) {
    v_.resize(o.size());
    detail::any_copy<T, typename A::value_type, typename A::memory_space, host_memory>(detail::raw_ptr(o),
                                                                                       v_.data(), o.size());
  }
  std::vector<T> v_;
};

// ---------------------------------------------------------------------------
// device array
// ---------------------------------------------------------------------------
template <typename T>
class array1d<T, device_memory> {
 public:
  typedef T value_type;
  typedef device_memory memory_space;
  typedef array1d_format format;
  typedef device_ptr<T> iterator;
  typedef device_ptr<const T> const_iterator;
  typedef device_ptr<T> pointer;
  typedef detail::device_reference<T> reference;
#ifdef CUSP_B200_HAVE_THRUST
  typedef detail::device_reference<const T> const_reference;  // &x[0] of a const array is a device_ptr<const T>
#else
  typedef T const_reference;
#endif
  typedef size_t size_type;
  typedef array1d container;
  typedef array1d_view<iterator> view;
  typedef array1d_view<const_iterator> const_view;
  template <typename Space>
  struct rebind {
    typedef array1d<T, Space> type;
  };

  array1d() {}
  explicit array1d(size_t n) {
    resize(n);
    if (n_) detail::cuda_check(cudaMemset(p_, 0, n_ * sizeof(T)), "cusp::array1d<device> zero fill");
  }
  array1d(size_t n, const T &value) {
    resize(n);
    fill_range(0, n_, value);
  }
  array1d(const array1d &o) { assign_from(o); }
  array1d(array1d &&o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = nullptr; o.n_ = o.cap_ = 0; }
  template <typename U, typename Space>
  array1d(const array1d<U, Space> &o) {
    assign_from(o);
  }
  template <typename It>
  array1d(const array1d_view<It> &o) {
    assign_from(o);
  }
  template <typename It, typename = typename std::enable_if<!std::is_integral<It>::value>::type>
  array1d(It first, It last) {
    std::vector<T> h(first, last);
    resize(h.size());
    detail::raw_copy<T, host_memory, device_memory>(h.data(), p_, n_);
  }
  // std::vector interoperability (testing/array1d.cu:66-140)
  template <typename U, typename Alloc>
  array1d(const std::vector<U, Alloc> &v) {
    std::vector<T> h(v.begin(), v.end());
    resize(h.size());
    detail::raw_copy<T, host_memory, device_memory>(h.data(), p_, n_);
  }
  template <typename U, typename Alloc>
  array1d &operator=(const std::vector<U, Alloc> &v) {
    std::vector<T> h(v.begin(), v.end());
    resize(h.size());
    detail::raw_copy<T, host_memory, device_memory>(h.data(), p_, n_);
    return *this;
  }
  ~array1d() { release(); }

  array1d &operator=(const array1d &o) {
    if (this != &o) assign_from(o);
    return *this;
  }
  array1d &operator=(array1d &&o) noexcept {
    swap(o);
    return *this;
  }
  template <typename U, typename Space>
  array1d &operator=(const array1d<U, Space> &o) {
    assign_from(o);
    return *this;
  }
  template <typename It>
  array1d &operator=(const array1d_view<It> &o) {
    assign_from(o);
    return *this;
  }

  size_t size() const { return n_; }
  size_t capacity() const { return cap_; }
  bool empty() const { return n_ == 0; }
  void reserve(size_t n) {
    if (n <= cap_) return;
    T *np = nullptr;
    detail::cuda_check(cudaMalloc(&np, n * sizeof(T)), "cusp::array1d<device> allocation");
    if (n_) detail::raw_copy<T, device_memory, device_memory>(p_, np, n_);
    if (p_) cudaFree(p_);
    p_ = np;
    cap_ = n;
  }
  void resize(size_t n) {
    reserve(n);
    n_ = n;
  }
  void resize(size_t n, const T &value) {
    const size_t old = n_;
    resize(n);
    if (n > old) fill_range(old, n, value);
  }
  void assign(size_t n, const T &value) {
    resize(n);
    fill_range(0, n, value);
  }
  void push_back(const T &x) {
    if (n_ == cap_) reserve(cap_ ? 2 * cap_ : 16);
    detail::raw_copy<T, host_memory, device_memory>(&x, p_ + n_, 1);
    ++n_;
  }
  void clear() { n_ = 0; }
  void swap(array1d &o) {
    std::swap(p_, o.p_);
    std::swap(n_, o.n_);
    std::swap(cap_, o.cap_);
  }
  reference operator[](size_t i) { return detail::make_device_reference<T>(p_ + i); }
#ifdef CUSP_B200_HAVE_THRUST
  const_reference operator[](size_t i) const { return detail::make_device_reference<const T>(p_ + i); }
#else
  T operator[](size_t i) const { return (T)detail::device_reference<const T>(p_ + i); }
#endif
  pointer data() { return pointer(p_); }
  device_ptr<const T> data() const { return device_ptr<const T>(p_); }
  iterator begin() { return iterator(p_); }
  iterator end() { return iterator(p_ + n_); }
  const_iterator begin() const { return const_iterator(p_); }
  const_iterator end() const { return const_iterator(p_ + n_); }
  view subarray(size_t start, size_t num) { return view(begin() + start, begin() + start + num); }
  const_view subarray(size_t start, size_t num) const {
    return const_view(begin() + start, begin() + start + num);
  }

 private:
  void fill_range(size_t lo, size_t hi, const T &value) {
    if (hi <= lo) return;
    std::vector<T> h(hi - lo, value);
    detail::raw_copy<T, host_memory, device_memory>(h.data(), p_ + lo, hi - lo);
  }
  template <typename A>
  void assign_from(const A &o) {
    resize(o.size());
    detail::any_copy<T, typename A::value_type, typename A::memory_space, device_memory>(detail::raw_ptr(o), p_,
                                                                                         o.size());
  }
  void release() {
    if (p_) cudaFree(p_);
    p_ = nullptr;
    n_ = cap_ = 0;
  }
  T *p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};

// ---------------------------------------------------------------------------
// non-owning view over [first, last)   (cusp/array1d.h:361-517)
// ---------------------------------------------------------------------------
template <typename Iterator>
class array1d_view {
  typedef detail::iter_traits<Iterator> traits;

 public:
  typedef Iterator iterator;
  typedef Iterator const_iterator;
  typedef typename traits::value_type value_type;
  typedef typename traits::memory_space memory_space;
  typedef array1d_format format;
  typedef size_t size_type;
  typedef array1d<value_type, memory_space> container;
  typedef array1d_view view;
  typedef array1d_view const_view;
  typedef decltype(std::declval<Iterator>()[0]) reference;

  array1d_view() : first_(), n_(0), cap_(0) {}
  array1d_view(Iterator first, Iterator last) : first_(first), n_((size_t)(last - first)), cap_(n_) {}
  array1d_view(const array1d_view &o) = default;
  // from a container or another view whose iterator converts to ours
  template <typename Array, typename = typename std::enable_if<
                                std::is_convertible<decltype(std::declval<Array &>().begin()), Iterator>::value &&
                                !std::is_same<typename std::decay<Array>::type, array1d_view>::value>::type>
  array1d_view(Array &a) : first_(a.begin()), n_(a.size()), cap_(a.capacity()) {}  // array1d_view.cu:414-435

  // views assign element-wise (like the reference: view = array copies data)
  array1d_view &operator=(const array1d_view &o) = default;

  size_t size() const { return n_; }
  size_t capacity() const { return cap_; }
  bool empty() const { return n_ == 0; }
  Iterator begin() const { return first_; }
  Iterator end() const { return first_ + (std::ptrdiff_t)n_; }
  Iterator data() const { return first_; }
  reference operator[](size_t i) const { return first_[(std::ptrdiff_t)i]; }
  reference front() const { return first_[0]; }
  reference back() const { return first_[(std::ptrdiff_t)n_ - 1]; }
  // cusp/detail/array1d.inl: a view may shrink or grow back up to its capacity
  void resize(size_t n) {
    if (n > cap_) throw cusp::not_implemented_exception("array1d_view cannot resize() larger than capacity()");
    n_ = n;
  }
  array1d_view subarray(size_t start, size_t num) const {
    return array1d_view(first_ + (std::ptrdiff_t)start, first_ + (std::ptrdiff_t)(start + num));
  }

 private:
  Iterator first_;
  size_t n_, cap_;
};

template <typename Iterator>
array1d_view<Iterator> make_array1d_view(Iterator first, Iterator last) {
  return array1d_view<Iterator>(first, last);
}
template <typename T, typename Space>
typename array1d<T, Space>::view make_array1d_view(array1d<T, Space> &a) {
  return typename array1d<T, Space>::view(a);
}
template <typename T, typename Space>
typename array1d<T, Space>::const_view make_array1d_view(const array1d<T, Space> &a) {
  return typename array1d<T, Space>::const_view(a);
}
template <typename Iterator>
array1d_view<Iterator> make_array1d_view(const array1d_view<Iterator> &a) {
  return a;
}

// ---------------------------------------------------------------------------
// equality across memory spaces (the unit tests compare device and host arrays)
// ---------------------------------------------------------------------------
namespace detail {
template <typename A, typename B>
bool arrays_equal(const A &a, const B &b) {
  if (a.size() != b.size()) return false;
  auto ha = to_host_vector(a);
  auto hb = to_host_vector(b);
  for (size_t i = 0; i < ha.size(); ++i)
    if (!(ha[i] == hb[i])) return false;
  return true;
}
}  // namespace detail

template <typename T1, typename S1, typename T2, typename S2>
bool operator==(const array1d<T1, S1> &a, const array1d<T2, S2> &b) {
  return detail::arrays_equal(a, b);
}
template <typename T1, typename S1, typename T2, typename S2>
bool operator!=(const array1d<T1, S1> &a, const array1d<T2, S2> &b) {
  return !detail::arrays_equal(a, b);
}
template <typename T1, typename S1, typename U, typename Alloc>
bool operator==(const array1d<T1, S1> &a, const std::vector<U, Alloc> &v) {
  if (a.size() != v.size()) return false;
  auto ha = detail::to_host_vector(a);
  for (size_t i = 0; i < ha.size(); ++i)
    if (!(ha[i] == v[i])) return false;
  return true;
}
template <typename T1, typename S1, typename U, typename Alloc>
bool operator==(const std::vector<U, Alloc> &v, const array1d<T1, S1> &a) {
  return a == v;
}
template <typename T1, typename S1, typename U, typename Alloc>
bool operator!=(const array1d<T1, S1> &a, const std::vector<U, Alloc> &v) {
  return !(a == v);
}
template <typename T1, typename S1, typename U, typename Alloc>
bool operator!=(const std::vector<U, Alloc> &v, const array1d<T1, S1> &a) {
  return !(a == v);
}
template <typename T1, typename S1, typename It>
bool operator==(const array1d<T1, S1> &a, const array1d_view<It> &b) {
  return detail::arrays_equal(a, b);
}
template <typename T1, typename S1, typename It>
bool operator==(const array1d_view<It> &a, const array1d<T1, S1> &b) {
  return detail::arrays_equal(a, b);
}
template <typename It1, typename It2>
bool operator==(const array1d_view<It1> &a, const array1d_view<It2> &b) {
  return detail::arrays_equal(a, b);
}
template <typename T1, typename S1, typename It>
bool operator!=(const array1d<T1, S1> &a, const array1d_view<It> &b) {
  return !detail::arrays_equal(a, b);
}
template <typename T1, typename S1, typename It>
bool operator!=(const array1d_view<It> &a, const array1d<T1, S1> &b) {
  return !detail::arrays_equal(a, b);
}
template <typename It1, typename It2>
bool operator!=(const array1d_view<It1> &a, const array1d_view<It2> &b) {
  return !detail::arrays_equal(a, b);
}

// ---------------------------------------------------------------------------
// generator arrays (cusp/array1d.h:564-662: array1d_views over thrust::counting_iterator,
// thrust::constant_iterator and cusp::random_iterator).  Here they are host arrays filled at
// construction with the same element values — usable wherever an array1d is (copy into either
// memory space, cusp::copy, blas, operator==); O(n) memory instead of a fancy iterator.
// ---------------------------------------------------------------------------
template <typename ValueType>
class counting_array : public array1d<ValueType, host_memory> {
 public:
  typedef size_t size_type;
  counting_array(size_type size, ValueType init = ValueType(0)) : array1d<ValueType, host_memory>(size) {
    for (size_type i = 0; i < size; ++i) (*this)[i] = (ValueType)(init + (ValueType)i);
  }
};

template <typename ValueType>
class constant_array : public array1d<ValueType, host_memory> {
 public:
  typedef size_t size_type;
  constant_array(size_type size, ValueType value) : array1d<ValueType, host_memory>(size, value) {}
};

namespace detail {
// element i of cusp::random_iterator<T>(seed)  (cusp/iterator/detail/random_iterator.inl:30-128):
// a 64-bit integer hash of (i ^ seed) (the index type is ptrdiff_t, so the 64-bit variant is the one
// taken on LP64), truncated to 32 bits for value types of at most 4 bytes; integers take the hash
// as is, reals divide it by 2^32 resp. 2^64 -> [0, 1).
inline unsigned long long random_hash64(unsigned long long i, unsigned long long seed) {
  unsigned long long h = i ^ seed;
  h = ~h + (h << 21);
  h = h ^ (h >> 24);
  h = (h + (h << 3)) + (h << 8);
  h = h ^ (h >> 14);
  h = (h + (h << 2)) + (h << 4);
  h = h ^ (h >> 28);
  h = h + (h << 31);
  return h;
}
template <typename T, bool IsFloat = std::is_floating_point<T>::value, bool Wide = (sizeof(T) > 4)>
struct random_value;
template <typename T>
struct random_value<T, false, false> {
  static T of(unsigned long long h) { return (T)(unsigned int)h; }
};
template <typename T>
struct random_value<T, false, true> {
  static T of(unsigned long long h) { return (T)h; }
};
template <typename T>
struct random_value<T, true, false> {
  static T of(unsigned long long h) { return T((unsigned int)h) / (T(1u << 16) * T(1u << 16)); }
};
template <typename T>
struct random_value<T, true, true> {
  static T of(unsigned long long h) { return T(h) / (T(1ull << 32) * T(1ull << 32)); }
};
}  // namespace detail

template <typename ValueType>
class random_array : public array1d<ValueType, host_memory> {
 public:
  typedef size_t size_type;
  random_array(size_type size, size_type seed = 0) : array1d<ValueType, host_memory>(size) {
    for (size_type i = 0; i < size; ++i)
      (*this)[i] = detail::random_value<ValueType>::of(detail::random_hash64((unsigned long long)i, (unsigned long long)seed));
  }
};

}  // namespace cusp
