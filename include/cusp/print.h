// cusp/print.h — cusp::print(p[, stream]) / cusp::print_matrix(A) (reference: cusp/print.h,
// cusp/detail/print.inl:33-140), used by the reference's examples (examples/Views/cg_raw.cu) and main.cu.
// Same text: a header line ("sparse matrix <r, c> with n entries", "array2d <r, c>", "array1d <n>") and
// one "(value)" per entry with setprecision(4) / setw as in the reference; every sparse format prints
// through its COO image; device containers are copied to the host first.
#pragma once
#include <iomanip>
#include <iostream>

#include "array1d.h"
#include "array2d.h"
#include "convert.h"
#include "coo_matrix.h"

namespace cusp {
namespace detail {

template <typename T, typename Stream>
void print_value(const T &val, Stream &s, bool newline = true) {
  s << " " << std::setprecision(4) << std::setw(8) << "(" << val << ")" << (newline ? "\n" : " ");
}

template <typename Printable, typename Stream>
void print_as(const Printable &p, Stream &s, sparse_format) {
  cusp::coo_matrix<typename Printable::index_type, typename Printable::value_type, cusp::host_memory> coo(p);
  s << "sparse matrix <" << coo.num_rows << ", " << coo.num_cols << "> with " << coo.num_entries << " entries\n";
  for (size_t n = 0; n < coo.num_entries; n++) {
    s << " " << std::setw(14) << coo.row_indices[n];
    s << " " << std::setw(14) << coo.column_indices[n];
    print_value(coo.values[n], s);
  }
}
template <typename Printable, typename Stream>
void print_as(const Printable &p, Stream &s, array2d_format) {  // containers and views, either space: element access
  s << "array2d <" << p.num_rows << ", " << p.num_cols << ">\n";
  for (size_t i = 0; i < p.num_rows; i++) {
    for (size_t j = 0; j < p.num_cols; j++) print_value((typename Printable::value_type)p(i, j), s, false);
    s << "\n";
  }
}
template <typename Printable, typename Stream>
void print_as(const Printable &p, Stream &s, array1d_format) {
  s << "array1d <" << p.size() << ">\n";
  for (size_t i = 0; i < p.size(); i++) print_value((typename Printable::value_type)p[i], s);
}

}  // namespace detail

template <typename Printable, typename Stream>
void print(const Printable &p, Stream &s) {
  typedef typename Printable::format Format;
  typedef typename std::conditional<std::is_base_of<sparse_format, Format>::value, sparse_format, Format>::type Tag;
  detail::print_as(p, s, Tag());
}
template <typename Printable>
void print(const Printable &p) {
  cusp::print(p, std::cout);
}
template <typename Matrix>
void print_matrix(const Matrix &A) {
  cusp::print(A);
}

}  // namespace cusp
