// stand-in for <boost/tokenizer.hpp>: char_separator tokenizer that drops empty tokens
#pragma once
#include <string>
#include <vector>
namespace boost {
template <class C> class char_separator {
 public:
  explicit char_separator(const C *seps) : seps_(seps) {}
  std::basic_string<C> seps_;
};
template <class Sep> class tokenizer {
 public:
  typedef std::vector<std::string>::const_iterator iterator;
  tokenizer(const std::string &s, const Sep &sep) {
    std::string cur;
    for (size_t i = 0; i < s.size(); i++) {
      if (sep.seps_.find(s[i]) != std::string::npos) { if (!cur.empty()) toks_.push_back(cur); cur.clear(); }
      else cur += s[i];
    }
    if (!cur.empty()) toks_.push_back(cur);
  }
  iterator begin() const { return toks_.begin(); }
  iterator end() const { return toks_.end(); }
 private:
  std::vector<std::string> toks_;
};
}  // namespace boost
