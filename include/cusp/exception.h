// cusp/exception.h — same hierarchy as the reference (cusp/exception.h:31-84).
#pragma once
#include <stdexcept>
#include <string>

namespace cusp {

class exception : public std::exception {
 public:
  exception(const std::string &msg) : message(msg) {}
  ~exception() throw() {}
  const char *what() const throw() { return message.c_str(); }

 protected:
  std::string message;
};

class not_implemented_exception : public exception {
 public:
  template <typename M>
  not_implemented_exception(const M &msg) : exception(msg) {}
};
class io_exception : public exception {
 public:
  template <typename M>
  io_exception(const M &msg) : exception(msg) {}
};
class invalid_input_exception : public exception {
 public:
  template <typename M>
  invalid_input_exception(const M &msg) : exception(msg) {}
};
class format_exception : public exception {
 public:
  template <typename M>
  format_exception(const M &msg) : exception(msg) {}
};
class format_conversion_exception : public format_exception {
 public:
  template <typename M>
  format_conversion_exception(const M &msg) : format_exception(msg) {}
};
class runtime_exception : public exception {
 public:
  template <typename M>
  runtime_exception(const M &msg) : exception(msg) {}
};

}  // namespace cusp
