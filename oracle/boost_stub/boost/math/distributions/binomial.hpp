// oracle/boost_stub -- see boost/icl/interval_map.hpp.  Only the search stage calls these; abort if reached.
#pragma once
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>
namespace boost { namespace math {
struct binomial { binomial(double, double) {} };
template <class D> struct complemented2 { D d; double q; };
template <class D> inline complemented2<D> complement(const D &d, double q) { return {d, q}; }
template <class D> inline double quantile(const complemented2<D> &) { abort(); }
template <class D> inline double quantile(const D &, double) { abort(); }
}}
