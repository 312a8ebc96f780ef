// oracle/boost_stub -- TEST INFRASTRUCTURE.  Minimal stand-ins for the Boost headers the reference's
// align-stage sources include (src/search.h:22-34, src/util.cc:17-18).  Boost is not installed here and
// nothing on the ksw_extz2 / Alignment path uses these types at run time; the stubs only let
// chain.cc / refine.cc / util.cc compile so that the reference's own Alignment class can be linked
// into oracle/_ref/libsedef_ref.so.
#pragma once
#include <algorithm>
#include <cassert>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <unordered_map>
#include <utility>
namespace boost { namespace icl {
template <class T> struct discrete_interval {
	T lo{}, hi{};
	discrete_interval() {}
	discrete_interval(T a, T b) : lo(a), hi(b) {}
	T lower() const { return lo; }
	T upper() const { return hi; }
	bool operator<(const discrete_interval &o) const { return lo < o.lo || (lo == o.lo && hi < o.hi); }
	bool operator==(const discrete_interval &o) const { return lo == o.lo && hi == o.hi; }
};
template <class K, class V> struct interval_map : std::map<discrete_interval<K>, V> {};
}}
