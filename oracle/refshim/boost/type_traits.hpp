// stand-in for <boost/type_traits.hpp>
#pragma once
#include <type_traits>
namespace boost {
using std::is_arithmetic;
using std::is_integral;
using std::is_pointer;
using std::is_same;
}  // namespace boost
