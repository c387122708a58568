// stand-in for <boost/format.hpp>: printf-like directives are replaced by the streamed arguments in order
#pragma once
#include <ostream>
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
 public:
  explicit format(const std::string &f) : fmt_(f) {}
  template <class T>
  format &operator%(const T &v) { std::ostringstream s; s << v; args_.push_back(s.str()); return *this; }
  std::string str() const {
    std::string out;
    size_t a = 0;
    for (size_t i = 0; i < fmt_.size(); i++) {
      if (fmt_[i] == '%' && i + 1 < fmt_.size()) {
        if (fmt_[i + 1] == '%') { out += '%'; i++; continue; }
        size_t j = i + 1;
        while (j < fmt_.size() && !isalpha((unsigned char)fmt_[j])) j++;
        out += a < args_.size() ? args_[a++] : std::string();
        i = j;
        continue;
      }
      out += fmt_[i];
    }
    return out;
  }
 private:
  std::string fmt_;
  std::vector<std::string> args_;
};
inline std::ostream &operator<<(std::ostream &o, const format &f) { return o << f.str(); }
}  // namespace boost
