// stand-in for the reference's logger: messages are discarded, LOG_ERROR goes to stderr
#ifndef SPARSEX_INTERNALS_LOGGER_LOGGER_HPP
#define SPARSEX_INTERNALS_LOGGER_LOGGER_HPP
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
using namespace std;
namespace sparsex { namespace logging {
struct NullLog { template <class T> NullLog &operator<<(const T &) { return *this; } };
} }
#define LOG_ERROR std::cerr
#define LOG_WARNING sparsex::logging::NullLog()
#define LOG_INFO sparsex::logging::NullLog()
#define LOG_VERBOSE sparsex::logging::NullLog()
#define LOG_DEBUG sparsex::logging::NullLog()
#endif
