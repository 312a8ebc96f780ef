// oracle/boost_stub -- see boost/icl/interval_map.hpp.
#pragma once
#include <cstdlib>
#include <cstdint>
namespace boost { namespace math { namespace tools {
template <class F, class T> inline T newton_raphson_iterate(F, T, T, T, int) { abort(); }
template <class F, class T> inline T newton_raphson_iterate(F, T, T, T, int, std::uintmax_t &) { abort(); }
}}}
