// oracle/ref_nosearch.cc -- TEST INFRASTRUCTURE.  Link shim for the search-less reference binary oracle/_ref/sedef_ref:
// src/search.cc needs real Boost.ICL (not installed), so `sedef search` is the one command that cannot be built here
// (SURVEY.md section 8c).  search_main.cc still references search() and three counters (src/search.cc:29-31); they are
// defined here so that `sedef align bucket|generate` and `sedef stats generate` -- the unmodified reference code -- link.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "search.h"
int64_t TOTAL_ATTEMPTED = 0, JACCARD_FAILED = 0, INTERVAL_FAILED = 0;
std::vector<Hit> search(int, std::shared_ptr<Index>, std::shared_ptr<Index>, Tree &, const bool, const int, const bool, const bool)
{
	fprintf(stderr, "sedef_ref: the search stage is not part of this build (oracle/ref_nosearch.cc)\n");
	abort();
}
