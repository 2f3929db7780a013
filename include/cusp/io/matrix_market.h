// cusp/io/matrix_market.h — MatrixMarket text I/O (reference: cusp/io/matrix_market.h,
// cusp/io/detail/matrix_market.inl:72-598).  SURVEY §8(f) row 4: real matrices (SuiteSparse)
// reach the engine through the same containers as the gallery.
//
//   read_matrix_market_file(mtx, filename) / read_matrix_market_stream(mtx, stream)
//   write_matrix_market_file(mtx, filename) / write_matrix_market_stream(mtx, stream)
//
// for any sparse container, array2d and array1d in either memory space.  Behaviour kept from
// the reference: banner "%%MatrixMarket matrix {coordinate|array} {real|integer|pattern|complex}
// {general|symmetric|hermitian|skew-symmetric}"; comment lines start with '%'; 1-based
// indices are range-checked; `pattern` entries get the value 1; `symmetric` files are expanded
// (off-diagonal entries mirrored); hermitian / skew-symmetric and pattern arrays throw
// cusp::not_implemented_exception; entries are sorted by (row, column); writers always emit
// "coordinate real general" (sparse) or "array real general" (dense, column-major order) with
// a "\t rows \t cols [\t entries]" size line.  `complex` files throw: the engine has no complex
// value type (SURVEY §8 out of scope).
//
// The parser is not the reference's line-by-line iostream extraction: the stream is read into
// one buffer and scanned with strtol/strtod — tens of millions of entries per second instead
// of operator>> per token.
#pragma once
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../array1d.h"
#include "../array2d.h"
#include "../convert.h"
#include "../coo_matrix.h"
#include "../exception.h"

namespace cusp {
namespace io {
namespace detail {

struct matrix_market_banner {
  std::string storage;   // "array" | "coordinate"
  std::string type;      // "real" | "integer" | "pattern" | "complex"
  std::string symmetry;  // "general" | "symmetric" | "hermitian" | "skew-symmetric"
};

// whole-buffer scanner
class mm_scanner {
 public:
  explicit mm_scanner(std::string text) : buf_(std::move(text)), p_(buf_.c_str()), end_(buf_.c_str() + buf_.size()) {}

  // next line (without the terminator); false at end of input
  bool line(std::string &out) {
    if (p_ >= end_) return false;
    const char *e = static_cast<const char *>(memchr(p_, '\n', (size_t)(end_ - p_)));
    const char *stop = e ? e : end_;
    out.assign(p_, stop);
    if (!out.empty() && out.back() == '\r') out.pop_back();
    p_ = e ? e + 1 : end_;
    return true;
  }
  bool integer(long long &v) {
    skip_space();
    if (p_ >= end_) return false;
    char *q;
    errno = 0;
    v = strtoll(p_, &q, 10);
    if (q == p_) return false;
    p_ = q;
    return true;
  }
  bool real(double &v) {
    skip_space();
    if (p_ >= end_) return false;
    char *q;
    v = strtod(p_, &q);
    if (q == p_) return false;
    p_ = q;
    return true;
  }

 private:
  void skip_space() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
  }
  std::string buf_;
  const char *p_, *end_;
};

inline std::vector<std::string> split(const std::string &s) {
  std::istringstream is(s);
  return std::vector<std::string>(std::istream_iterator<std::string>(is), std::istream_iterator<std::string>());
}

// matrix_market.inl:72-99
inline matrix_market_banner read_banner(mm_scanner &in) {
  std::string line;
  in.line(line);
  const std::vector<std::string> t = split(line);
  if (t.size() != 5 || t[0] != "%%MatrixMarket" || t[1] != "matrix")
    throw cusp::io_exception("invalid MatrixMarket banner");
  matrix_market_banner b{t[2], t[3], t[4]};
  if (b.storage != "array" && b.storage != "coordinate")
    throw cusp::io_exception("invalid MatrixMarket storage format [" + b.storage + "]");
  if (b.type != "complex" && b.type != "real" && b.type != "integer" && b.type != "pattern")
    throw cusp::io_exception("invalid MatrixMarket data type [" + b.type + "]");
  if (b.symmetry != "general" && b.symmetry != "symmetric" && b.symmetry != "hermitian" &&
      b.symmetry != "skew-symmetric")
    throw cusp::io_exception("invalid MatrixMarket symmetry [" + b.symmetry + "]");
  return b;
}

// first line that is not a comment, as tokens
inline std::vector<std::string> size_line(mm_scanner &in) {
  std::string line;
  do {
    if (!in.line(line)) throw cusp::io_exception("unexpected EOF while reading MatrixMarket header");
  } while (!line.empty() && line[0] == '%');
  return split(line);
}

// coordinate storage -> host COO sorted by (row, column)  (matrix_market.inl:146-290)
template <typename IndexType, typename ValueType>
void read_coordinate(cusp::coo_matrix<IndexType, ValueType, cusp::host_memory> &coo, mm_scanner &in,
                     const matrix_market_banner &banner) {
  const std::vector<std::string> t = size_line(in);
  if (t.size() != 3) throw cusp::io_exception("invalid MatrixMarket coordinate format");
  const size_t num_rows = (size_t)strtoull(t[0].c_str(), nullptr, 10), num_cols = (size_t)strtoull(t[1].c_str(), nullptr, 10),
               stored = (size_t)strtoull(t[2].c_str(), nullptr, 10);
  if (banner.type == "complex")
    throw cusp::not_implemented_exception("MatrixMarket complex data: the B200 engine has real value types only");
  if (banner.symmetry == "hermitian")
    throw cusp::not_implemented_exception("MatrixMarket I/O does not currently support hermitian matrices");
  if (banner.symmetry == "skew-symmetric")
    throw cusp::not_implemented_exception("MatrixMarket I/O does not currently support skew-symmetric matrices");
  if (num_rows > (size_t)std::numeric_limits<IndexType>::max() || num_cols > (size_t)std::numeric_limits<IndexType>::max())
    throw cusp::io_exception("MatrixMarket dimensions exceed the index type");
  const bool pattern = banner.type == "pattern", mirror = banner.symmetry == "symmetric";

  std::vector<IndexType> ri, ci;
  std::vector<ValueType> va;
  ri.reserve(mirror ? 2 * stored : stored);
  ci.reserve(ri.capacity());
  va.reserve(ri.capacity());
  for (size_t n = 0; n < stored; ++n) {
    long long r, c;
    double v = 1.0;
    if (!in.integer(r) || !in.integer(c) || (!pattern && !in.real(v)))
      throw cusp::io_exception("unexpected EOF while reading MatrixMarket entries");
    if (r < 1) throw cusp::io_exception("found invalid row index (index < 1)");
    if (c < 1) throw cusp::io_exception("found invalid column index (index < 1)");
    if ((size_t)r > num_rows) throw cusp::io_exception("found invalid row index (index > num_rows)");
    if ((size_t)c > num_cols) throw cusp::io_exception("found invalid column index (index > num_columns)");
    ri.push_back((IndexType)(r - 1));
    ci.push_back((IndexType)(c - 1));
    va.push_back((ValueType)v);
    if (mirror && r != c) {  // off-diagonal entries of a symmetric file appear on both sides
      ri.push_back((IndexType)(c - 1));
      ci.push_back((IndexType)(r - 1));
      va.push_back((ValueType)v);
    }
  }
  coo.resize(num_rows, num_cols, ri.size());
  for (size_t n = 0; n < ri.size(); ++n) {
    coo.row_indices[n] = ri[n];
    coo.column_indices[n] = ci[n];
    coo.values[n] = va[n];
  }
  coo.sort_by_row_and_column();
}

// array storage -> host array2d (file order is column-major)  (matrix_market.inl:343-417)
template <typename ValueType>
void read_array(cusp::array2d<ValueType, cusp::host_memory> &mtx, mm_scanner &in, const matrix_market_banner &banner) {
  const std::vector<std::string> t = size_line(in);
  if (t.size() != 2) throw cusp::io_exception("invalid MatrixMarket array format");
  const size_t num_rows = (size_t)strtoull(t[0].c_str(), nullptr, 10), num_cols = (size_t)strtoull(t[1].c_str(), nullptr, 10);
  if (banner.type == "pattern")
    throw cusp::not_implemented_exception("pattern array MatrixMarket format is not supported");
  if (banner.type == "complex")
    throw cusp::not_implemented_exception("MatrixMarket complex data: the B200 engine has real value types only");
  mtx.resize(num_rows, num_cols);
  for (size_t j = 0; j < num_cols; ++j)
    for (size_t i = 0; i < num_rows; ++i) {
      double v;
      if (!in.real(v)) throw cusp::io_exception("unexpected EOF while reading MatrixMarket entries");
      mtx(i, j) = (ValueType)v;
    }
  if (banner.symmetry != "general")
    throw cusp::not_implemented_exception("only general array symmetric MatrixMarket format is supported");
}

template <typename Stream>
std::string slurp(Stream &input) {
  return std::string(std::istreambuf_iterator<char>(input), std::istreambuf_iterator<char>());
}

// sparse / dense destinations
template <typename Matrix, typename Format>
void read_into(Matrix &mtx, mm_scanner &in, const matrix_market_banner &banner, Format) {
  typedef typename Matrix::value_type ValueType;
  if (banner.storage == "coordinate") {
    cusp::coo_matrix<int, ValueType, cusp::host_memory> coo;
    read_coordinate(coo, in, banner);
    cusp::convert(coo, mtx);
  } else {
    cusp::array2d<ValueType, cusp::host_memory> dense;
    read_array(dense, in, banner);
    cusp::convert(dense, mtx);
  }
}
// vectors: an n x 1 (or 1 x n) matrix in either storage  (matrix_market.inl:463-475)
template <typename Array>
void read_into(Array &a, mm_scanner &in, const matrix_market_banner &banner, cusp::array1d_format) {
  typedef typename Array::value_type ValueType;
  cusp::array2d<ValueType, cusp::host_memory> dense;
  read_into(dense, in, banner, cusp::array2d_format());
  if (dense.num_rows != 1 && dense.num_cols != 1 && dense.num_entries != 0)
    throw cusp::format_conversion_exception("MatrixMarket: a matrix with several rows and columns cannot become an array1d");
  cusp::array1d<ValueType, cusp::host_memory> h(dense.num_rows * dense.num_cols);
  for (size_t i = 0; i < dense.num_rows; ++i)
    for (size_t j = 0; j < dense.num_cols; ++j) h[i * dense.num_cols + j] = dense(i, j);
  a = h;
}

template <typename IndexType, typename ValueType, typename Stream>
void write_coordinate(const cusp::coo_matrix<IndexType, ValueType, cusp::host_memory> &coo, Stream &output) {
  std::ostringstream os;  // one buffered write; max_digits10 so that read(write(A)) == A
  os.precision(std::numeric_limits<ValueType>::max_digits10);
  os << "%%MatrixMarket matrix coordinate real general\n";
  os << "\t" << coo.num_rows << "\t" << coo.num_cols << "\t" << coo.num_entries << "\n";
  for (size_t i = 0; i < coo.num_entries; ++i)
    os << (coo.row_indices[i] + 1) << " " << (coo.column_indices[i] + 1) << " " << (ValueType)coo.values[i] << "\n";
  output << os.str();
}

template <typename Matrix, typename Stream>
void write_from(const Matrix &mtx, Stream &output, cusp::sparse_format) {
  cusp::coo_matrix<typename Matrix::index_type, typename Matrix::value_type, cusp::host_memory> coo(mtx);
  write_coordinate(coo, output);
}
template <typename Array, typename Stream>
void write_from(const Array &a, Stream &output, cusp::array1d_format) {
  typedef typename Array::value_type ValueType;
  cusp::array1d<ValueType, cusp::host_memory> h(a);
  std::ostringstream os;
  os.precision(std::numeric_limits<ValueType>::max_digits10);
  os << "%%MatrixMarket matrix array real general\n";
  os << "\t" << h.size() << "\t1\n";
  for (size_t i = 0; i < h.size(); ++i) os << (ValueType)h[i] << "\n";
  output << os.str();
}
template <typename Matrix, typename Stream>
void write_from(const Matrix &mtx, Stream &output, cusp::array2d_format) {
  typedef typename Matrix::value_type ValueType;
  cusp::array2d<ValueType, cusp::host_memory, typename Matrix::orientation> h(mtx);
  std::ostringstream os;
  os.precision(std::numeric_limits<ValueType>::max_digits10);
  os << "%%MatrixMarket matrix array real general\n";
  os << "\t" << h.num_rows << "\t" << h.num_cols << "\n";
  for (size_t j = 0; j < h.num_cols; ++j)
    for (size_t i = 0; i < h.num_rows; ++i) os << (ValueType)h(i, j) << "\n";
  output << os.str();
}

}  // namespace detail

template <typename Matrix, typename Stream>
void read_matrix_market_stream(Matrix &mtx, Stream &input) {
  detail::mm_scanner in(detail::slurp(input));
  const detail::matrix_market_banner banner = detail::read_banner(in);
  detail::read_into(mtx, in, banner, typename Matrix::format());
}

template <typename Matrix>
void read_matrix_market_file(Matrix &mtx, const std::string &filename) {
  std::ifstream file(filename.c_str(), std::ios::in | std::ios::binary);
  if (!file) throw cusp::io_exception(std::string("unable to open file \"") + filename + std::string("\" for reading"));
  read_matrix_market_stream(mtx, file);
}

template <typename Matrix, typename Stream>
void write_matrix_market_stream(const Matrix &mtx, Stream &output) {
  typedef typename Matrix::format Format;
  typedef typename std::conditional<std::is_base_of<cusp::sparse_format, Format>::value, cusp::sparse_format, Format>::type Tag;
  detail::write_from(mtx, output, Tag());
}

template <typename Matrix>
void write_matrix_market_file(const Matrix &mtx, const std::string &filename) {
  std::ofstream file(filename.c_str());
  if (!file) throw cusp::io_exception(std::string("unable to open file \"") + filename + std::string("\" for writing"));
  write_matrix_market_stream(mtx, file);
}

}  // namespace io
}  // namespace cusp
