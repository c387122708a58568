// stand-in for <boost/regex.hpp> on top of <regex>
#pragma once
#include <regex>
namespace boost {
using std::match_results;
using std::regex;
using std::regex_match;
using std::regex_search;
using std::smatch;
}  // namespace boost
