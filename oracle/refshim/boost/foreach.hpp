// stand-in for <boost/foreach.hpp>: range-based for
#pragma once
#define BOOST_FOREACH(decl, range) for (decl : range)
