#pragma once
#include <boost/assign/list_of.hpp>
