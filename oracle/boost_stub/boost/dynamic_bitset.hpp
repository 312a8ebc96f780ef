// oracle/boost_stub -- TEST INFRASTRUCTURE.  Stand-in for boost::dynamic_bitset, which only `sedef stats diff`
// (src/stats_main.cc:397-511) uses; it lets stats_main.cc compile so that the search-less reference binary
// oracle/_ref/sedef_ref can be linked.  Never on a product path.
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
template <class B = unsigned long> class dynamic_bitset {
	std::vector<bool> v_;
public:
	dynamic_bitset() {}
	explicit dynamic_bitset(std::size_t n) : v_(n, false) {}
	void set(std::size_t i) { v_[i] = true; }
	bool operator[](std::size_t i) const { return v_[i]; }
	std::size_t size() const { return v_.size(); }
	std::size_t count() const { std::size_t c = 0; for (bool b : v_) c += b; return c; }
	dynamic_bitset operator~() const { dynamic_bitset r(*this); r.v_.flip(); return r; }
	dynamic_bitset operator&(const dynamic_bitset &o) const
	{
		dynamic_bitset r(v_.size() < o.v_.size() ? v_.size() : o.v_.size());
		for (std::size_t i = 0; i < r.v_.size(); ++i) r.v_[i] = v_[i] && o.v_[i];
		return r;
	}
};
}
