// stand-in for <boost/function.hpp>
#pragma once
#include <functional>
namespace boost {
template <class Sig> using function = std::function<Sig>;
}
