// oracle/ref_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry points onto the REFERENCE's own classes (compiled from /root/reference by oracle/Makefile
// into oracle/_ref/libsedef_ref.so).  Used by tests/golden/make_golden.py to produce SEDEF-level
// golden records (CIGAR string in SEDEF's M/D/I alphabet and the Alignment error counters) and to
// validate oracle/sd_stats_port.c.  Nothing here restates reference logic: it only calls it.
#include <cstring>
#include <string>
#include "align.h"   // /root/reference/src/align.h

extern "C" {
// Alignment(fa, fb): reference src/align.cc:76-88 (align_dna + align_helper + populate_nice_alignment)
int ref_alignment(const char *fa, const char *fb, char *cigar_out, int cigar_cap,
                  int *span, int *matches, int *mismatches, int *gaps, int *gap_bases)
{
	Alignment a{std::string(fa), std::string(fb)};
	std::string c = a.cigar_string();
	if ((int)c.size() + 1 > cigar_cap) return -1;
	memcpy(cigar_out, c.c_str(), c.size() + 1);
	*span = a.span(); *matches = a.matches(); *mismatches = a.mismatches(); *gaps = a.gaps(); *gap_bases = a.gap_bases();
	return 0;
}
// Alignment(fa, fb, cigar): reference src/align.cc:90-105 ("from_cigar")
int ref_alignment_from_cigar(const char *fa, const char *fb, const char *cigar,
                             int *span, int *matches, int *mismatches, int *gaps, int *gap_bases)
{
	Alignment a{std::string(fa), std::string(fb), std::string(cigar)};
	*span = a.span(); *matches = a.matches(); *mismatches = a.mismatches(); *gaps = a.gaps(); *gap_bases = a.gap_bases();
	return 0;
}
}
