// cusp/array1d.h — cusp::array1d<T,MemorySpace> and array1d_view
// (cusp/array1d.h:98-517), written without Thrust: host arrays sit on
// std::vector, device arrays on cudaMalloc.  Element access on device arrays goes
// through a proxy reference (one cudaMemcpy per access, as thrust::device_reference).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <type_traits>
#include <vector>

#include "detail/engine.h"
#include "memory.h"

namespace cusp {

template <typename T, typename MemorySpace>
class array1d;
template <typename T, typename MemorySpace>
class array1d_view;

namespace detail {

template <typename T>
class device_reference {
 public:
  explicit device_reference(T *p) : p_(p) {}
  operator T() const {
    T v;
    cuda_check(cudaMemcpy(&v, p_, sizeof(T), cudaMemcpyDeviceToHost), "device_reference read");
    return v;
  }
  device_reference &operator=(const T &v) {
    cuda_check(cudaMemcpy(p_, &v, sizeof(T), cudaMemcpyHostToDevice), "device_reference write");
    return *this;
  }
  device_reference &operator=(const device_reference &o) { return *this = (T)o; }

 private:
  T *p_;
};

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
  typedef T &reference;
  typedef array1d_view<T, host_memory> view;
  typedef array1d_view<const T, host_memory> const_view;
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
    assign_from(o.data(), o.size(), Space());
  }
  template <typename U, typename Space>
  array1d(const array1d_view<U, Space> &o) {
    assign_from(o.data(), o.size(), Space());
  }
  template <typename It, typename = typename std::enable_if<!std::is_integral<It>::value>::type>
  array1d(It first, It last) : v_(first, last) {}

  array1d &operator=(const array1d &o) {
    v_ = o.v_;
    return *this;
  }
  template <typename U, typename Space>
  array1d &operator=(const array1d<U, Space> &o) {
    assign_from(o.data(), o.size(), Space());
    return *this;
  }
  template <typename U, typename Space>
  array1d &operator=(const array1d_view<U, Space> &o) {
    assign_from(o.data(), o.size(), Space());
    return *this;
  }

  size_t size() const { return v_.size(); }
  bool empty() const { return v_.empty(); }
  void resize(size_t n) { v_.resize(n); }
  void resize(size_t n, const T &value) { v_.resize(n, value); }
  void reserve(size_t n) { v_.reserve(n); }
  void push_back(const T &x) { v_.push_back(x); }
  void clear() { v_.clear(); }
  void swap(array1d &o) { v_.swap(o.v_); }
  T &operator[](size_t i) { return v_[i]; }
  const T &operator[](size_t i) const { return v_[i]; }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }
  iterator begin() { return v_.data(); }
  iterator end() { return v_.data() + v_.size(); }
  const_iterator begin() const { return v_.data(); }
  const_iterator end() const { return v_.data() + v_.size(); }
  view subarray(size_t start, size_t num) { return view(data() + start, num); }

 private:
  template <typename U, typename Space>
  void assign_from(const U *src, size_t n, Space) {
    v_.resize(n);
    typedef typename std::remove_const<U>::type UU;
    if (std::is_same<UU, T>::value)
      detail::raw_copy<T, Space, host_memory>(reinterpret_cast<const T *>(src), v_.data(), n);
    else
      detail::convert_copy<T, UU, Space, host_memory>(src, v_.data(), n);
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
  typedef T *iterator;  // raw device pointers (not dereferenceable on the host)
  typedef const T *const_iterator;
  typedef detail::device_reference<T> reference;
  typedef array1d_view<T, device_memory> view;
  typedef array1d_view<const T, device_memory> const_view;
  template <typename Space>
  struct rebind {
    typedef array1d<T, Space> type;
  };

  array1d() {}
  explicit array1d(size_t n) { resize(n); fill_bytes_zero(); }
  array1d(size_t n, const T &value) {
    resize(n);
    fill(value);
  }
  array1d(const array1d &o) { assign_from(o.data(), o.size(), device_memory()); }
  template <typename U, typename Space>
  array1d(const array1d<U, Space> &o) {
    assign_from(o.data(), o.size(), Space());
  }
  template <typename U, typename Space>
  array1d(const array1d_view<U, Space> &o) {
    assign_from(o.data(), o.size(), Space());
  }
  ~array1d() { release(); }

  array1d &operator=(const array1d &o) {
    if (this != &o) assign_from(o.data(), o.size(), device_memory());
    return *this;
  }
  template <typename U, typename Space>
  array1d &operator=(const array1d<U, Space> &o) {
    assign_from(o.data(), o.size(), Space());
    return *this;
  }
  template <typename U, typename Space>
  array1d &operator=(const array1d_view<U, Space> &o) {
    assign_from(o.data(), o.size(), Space());
    return *this;
  }

  size_t size() const { return n_; }
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
    if (n > old) {
      std::vector<T> h(n - old, value);
      detail::raw_copy<T, host_memory, device_memory>(h.data(), p_ + old, n - old);
    }
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
  reference operator[](size_t i) { return reference(p_ + i); }
  T operator[](size_t i) const { return (T)detail::device_reference<T>(p_ + i); }
  T *data() { return p_; }
  const T *data() const { return p_; }
  iterator begin() { return p_; }
  iterator end() { return p_ + n_; }
  const_iterator begin() const { return p_; }
  const_iterator end() const { return p_ + n_; }
  view subarray(size_t start, size_t num) { return view(p_ + start, num); }

 private:
  void fill(const T &value) {
    if (!n_) return;
    std::vector<T> h(n_, value);
    detail::raw_copy<T, host_memory, device_memory>(h.data(), p_, n_);
  }
  void fill_bytes_zero() {
    if (n_) detail::cuda_check(cudaMemset(p_, 0, n_ * sizeof(T)), "cusp::array1d<device> zero fill");
  }
  template <typename U, typename Space>
  void assign_from(const U *src, size_t n, Space) {
    resize(n);
    typedef typename std::remove_const<U>::type UU;
    if (std::is_same<UU, T>::value)
      detail::raw_copy<T, Space, device_memory>(reinterpret_cast<const T *>(src), p_, n);
    else
      detail::convert_copy<T, UU, Space, device_memory>(src, p_, n);
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
// non-owning view over [first, first+n) in a memory space
// (the reference's array1d_view<Iterator>; examples/Views/cg_raw.cu wraps raw
// device pointers this way)
// ---------------------------------------------------------------------------
template <typename T, typename MemorySpace>
class array1d_view {
 public:
  typedef typename std::remove_const<T>::type value_type;
  typedef MemorySpace memory_space;
  typedef array1d_format format;
  typedef T *iterator;
  typedef array1d_view view;

  array1d_view() : p_(nullptr), n_(0) {}
  array1d_view(T *first, size_t n) : p_(first), n_(n) {}
  array1d_view(T *first, T *last) : p_(first), n_((size_t)(last - first)) {}
  array1d_view(array1d<value_type, MemorySpace> &a) : p_(a.data()), n_(a.size()) {}
  array1d_view(const array1d<value_type, MemorySpace> &a) : p_(const_cast<T *>(a.data())), n_(a.size()) {}

  size_t size() const { return n_; }
  T *data() const { return p_; }
  T *begin() const { return p_; }
  T *end() const { return p_ + n_; }
  void resize(size_t n) {
    if (n > n_) throw cusp::not_implemented_exception("array1d_view cannot resize() larger than the wrapped range");
    n_ = n;
  }
  // host views index directly; device views through a proxy
  template <typename S = MemorySpace>
  typename std::enable_if<std::is_same<S, host_memory>::value, T &>::type operator[](size_t i) const {
    return p_[i];
  }
  template <typename S = MemorySpace>
  typename std::enable_if<std::is_same<S, device_memory>::value, detail::device_reference<value_type>>::type
  operator[](size_t i) const {
    return detail::device_reference<value_type>(const_cast<value_type *>(p_) + i);
  }

 private:
  T *p_;
  size_t n_;
};

template <typename T, typename Space>
array1d_view<T, Space> make_array1d_view(array1d<T, Space> &a) {
  return array1d_view<T, Space>(a);
}
template <typename Space, typename T>
array1d_view<T, Space> make_array1d_view(T *first, T *last) {
  return array1d_view<T, Space>(first, last);
}

// equality across memory spaces (testing uses ASSERT_EQUAL(device_array, host_array))
namespace detail {
template <typename A>
std::vector<typename A::value_type> to_host_vector(const A &a) {
  std::vector<typename A::value_type> h(a.size());
  raw_copy<typename A::value_type, typename A::memory_space, host_memory>(a.data(), h.data(), a.size());
  return h;
}
}  // namespace detail

template <typename T1, typename S1, typename T2, typename S2>
bool operator==(const array1d<T1, S1> &a, const array1d<T2, S2> &b) {
  if (a.size() != b.size()) return false;
  auto ha = detail::to_host_vector(a);
  auto hb = detail::to_host_vector(b);
  for (size_t i = 0; i < ha.size(); ++i)
    if (!(ha[i] == hb[i])) return false;
  return true;
}
template <typename T1, typename S1, typename T2, typename S2>
bool operator!=(const array1d<T1, S1> &a, const array1d<T2, S2> &b) {
  return !(a == b);
}

}  // namespace cusp
