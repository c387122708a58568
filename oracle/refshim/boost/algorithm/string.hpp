// stand-in for <boost/algorithm/string.hpp>: split / is_any_of / token_compress_on
#pragma once
#include <string>
#include <vector>
namespace boost {
namespace algorithm {
struct any_of_pred { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline any_of_pred is_any_of(const char *s) { return any_of_pred{s}; }
enum token_compress_mode_type { token_compress_on, token_compress_off };
}
using algorithm::is_any_of;
using algorithm::token_compress_on;
template <class Pred>
inline std::vector<std::string> &split(std::vector<std::string> &out, const std::string &s, Pred p,
                                       algorithm::token_compress_mode_type mode = algorithm::token_compress_off) {
  out.clear();
  std::string cur;
  bool last_sep = false;
  for (size_t i = 0; i < s.size(); i++) {
    if (p(s[i])) {
      if (!(mode == algorithm::token_compress_on && last_sep)) { out.push_back(cur); cur.clear(); }
      last_sep = true;
    } else { cur += s[i]; last_sep = false; }
  }
  out.push_back(cur);
  return out;
}
}  // namespace boost
