// stand-in for <boost/lexical_cast.hpp>
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : public std::bad_cast {
  const char *what() const noexcept override { return "bad lexical cast"; }
};
template <class Target, class Source = std::string>
inline Target lexical_cast(const Source &s) {
  std::stringstream ss;
  ss << s;
  Target t;
  if (!(ss >> t)) throw bad_lexical_cast();
  char c;
  if (ss >> c) throw bad_lexical_cast();
  return t;
}
template <>
inline std::string lexical_cast<std::string, std::string>(const std::string &s) { return s; }
}  // namespace boost
