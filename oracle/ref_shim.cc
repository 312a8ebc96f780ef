// oracle/ref_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry points onto the REFERENCE's own classes (compiled from /root/reference by oracle/Makefile
// into oracle/_ref/libsedef_ref.so).  Used by tests/golden/make_golden.py to produce SEDEF-level
// golden records (CIGAR string in SEDEF's M/D/I alphabet and the Alignment error counters) and to
// validate oracle/sd_stats_port.c.  Nothing here restates reference logic: it only calls it.
#include <cstring>
#include <string>
#include "align.h"   // /root/reference/src/align.h
#include "chain.h"   // /root/reference/src/chain.h
#include "hit.h"
#include "hash.h"
#include <algorithm>
#include <memory>
#include <sstream>
#include "globals.h"

// non-static functions of /root/reference/src/chain.cc that chain.h does not declare (chain.cc:24,103)
std::vector<Anchor> generate_anchors(const std::string &query, const std::string &ref, const Hit &orig, const int kmer_size);
std::pair<std::vector<int>, std::vector<std::pair<int, bool>>> chain_anchors(std::vector<Anchor> &anchors);

// class Alignment declares `friend void test(int, char **argv);` (src/align.h) and none of the reference sources linked here
// defines it.  The shim supplies it as a READ-ONLY window onto the private column strings: argv = {fa, fb, out_a, out_b, &rc}.
void test(int cap, char **argv)
{
	Alignment a{std::string(argv[0]), std::string(argv[1])};
	int *rc = reinterpret_cast<int *>(argv[4]);
	if ((int)a.align_a.size() + 1 > cap || (int)a.align_b.size() + 1 > cap) { *rc = -1; return; }
	memcpy(argv[2], a.align_a.c_str(), a.align_a.size() + 1);
	memcpy(argv[3], a.align_b.c_str(), a.align_b.size() + 1);
	*rc = (int)a.align_a.size();
}

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
// The column strings the reference's Alignment(fa, fb) builds (populate_nice_alignment, src/align.cc:274-315; private
// members align_a / align_b, src/align.h:39): the input of the BEDPE stat loop of src/stats_main.cc:244-271.
int ref_alignment_strings(const char *fa, const char *fb, char *out_a, char *out_b, int cap)
{
	int rc = -1;
	char *argv[5] = {const_cast<char *>(fa), const_cast<char *>(fb), out_a, out_b, reinterpret_cast<char *>(&rc)};
	test(cap, argv);
	return rc;
}
// Alignment(fa, fb, cigar): reference src/align.cc:90-105 ("from_cigar")
int ref_alignment_from_cigar(const char *fa, const char *fb, const char *cigar,
                             int *span, int *matches, int *mismatches, int *gaps, int *gap_bases)
{
	Alignment a{std::string(fa), std::string(fb), std::string(cigar)};
	*span = a.span(); *matches = a.matches(); *mismatches = a.mismatches(); *gaps = a.gaps(); *gap_bases = a.gap_bases();
	return 0;
}
// fast_align(query, ref, orig, k): the reference's whole per-region align path (src/chain.cc:203-268: anchors,
// chaining, Alignment guide constructors with ksw gap fills, refine_chains with merges and +-500 bp side
// extensions) -- i.e. every call site of the hot path, untouched.  One text line per resulting hit.
int ref_fast_align(const char *query, const char *ref, int kmer_size, char *out, int cap)
{
	std::string q(query), r(ref);
	auto qp = std::make_shared<Sequence>("QRY", q);
	auto rp = std::make_shared<Sequence>("REF", r);
	Hit orig{qp, 0, (int)q.size(), rp, 0, (int)r.size()};
	std::vector<Hit> hits = fast_align(q, r, orig, kmer_size);
	std::ostringstream os;
	for (auto &h : hits)
		os << h.query_start << ' ' << h.query_end << ' ' << h.ref_start << ' ' << h.ref_end << ' ' << h.aln.cigar_string() << ' '
		   << h.aln.span() << ' ' << h.aln.matches() << ' ' << h.aln.mismatches() << ' ' << h.aln.gaps() << ' ' << h.aln.gap_bases() << '\n';
	std::string sres = os.str();
	if ((int)sres.size() + 1 > cap) return -1;
	memcpy(out, sres.c_str(), sres.size() + 1);
	return (int)hits.size();
}
// The chain wave of fast_align (src/chain.cc:211-258): anchors, chaining, the chain filter, and for every kept chain
// the reference's Alignment(query, ref, anchors, guide_idx) constructor (src/align.cc:199-270).  One line per chain:
//   start_a end_a start_b end_b cigar span matches mismatches gaps gap_bases n  q r l  q r l ...   (anchors in guide order)
int ref_chain_guides(const char *query, const char *ref, int kmer_size, char *out, int cap)
{
	std::string q(query), r(ref);
	auto qp = std::make_shared<Sequence>("QRY", q);
	auto rp = std::make_shared<Sequence>("REF", r);
	Hit orig{qp, 0, (int)q.size(), rp, 0, (int)r.size()};
	auto anchors = generate_anchors(q, r, orig, kmer_size);
	auto chains_init = chain_anchors(anchors);
	auto &bounds = chains_init.second;
	auto &chain = chains_init.first;
	std::ostringstream os;
	int n = 0;
	for (int bi = 1; bi < (int)bounds.size(); bi++) {                   // the loop of src/chain.cc:222-247, called, not changed
		bool has_u = bounds[bi].second;
		int be = bounds[bi].first, bs = bounds[bi - 1].first;
		int qlo = anchors[chain[be - 1]].q, qhi = anchors[chain[bs]].q + anchors[chain[bs]].l;
		int rlo = anchors[chain[be - 1]].r, rhi = anchors[chain[bs]].r + anchors[chain[bs]].l;
		int span = std::max(rhi - rlo, qhi - qlo);
		if ((!has_u || span < Globals::Chain::MIN_UPPERCASE_MATCH) &&
		    span < Globals::Search::MIN_READ_SIZE * (1 - Globals::Search::MAX_ERROR)) continue;
		std::vector<int> guide;
		for (int k = be - 1; k >= bs; k--) guide.emplace_back(chain[k]);
		Alignment a(q, r, anchors, guide);
		Hit h{qp, qlo, qhi, rp, rlo, rhi};
		h.aln = a;
		update_from_alignment(h);
		os << h.query_start << ' ' << h.query_end << ' ' << h.ref_start << ' ' << h.ref_end << ' ' << a.cigar_string() << ' ' << a.span() << ' '
		   << a.matches() << ' ' << a.mismatches() << ' ' << a.gaps() << ' ' << a.gap_bases() << ' ' << guide.size();
		for (int g : guide) os << ' ' << anchors[g].q << ' ' << anchors[g].r << ' ' << anchors[g].l;
		os << '\n';
		++n;
	}
	std::string sres = os.str();
	if ((int)sres.size() + 1 > cap) return -1;
	memcpy(out, sres.c_str(), sres.size() + 1);
	return n;
}
// One region as the region-level driver sees it: the reference's own anchors and (filtered) chains for a seed hit `orig`
// (same chromosome or not, with its start coordinates), followed by what the reference's fast_align makes of the same region.
//   "A q r l has_u"              one line per anchor (index = line order)
//   "C n i0 i1 ..."              one line per chain that passes the filter of src/chain.cc:222-247: anchor indices in query order
//   "H qs qe rs re cigar span matches mismatches gaps gap_bases"      one line per hit fast_align returns
int ref_region(const char *query, const char *ref, int kmer_size, int same_chr, int orig_qs, int orig_rs, char *out, int cap)
{
	std::string q(query), r(ref);
	auto qp = std::make_shared<Sequence>(same_chr ? "CHR" : "QRY", q);
	auto rp = std::make_shared<Sequence>(same_chr ? "CHR" : "REF", r);
	Hit orig{qp, orig_qs, orig_qs + (int)q.size(), rp, orig_rs, orig_rs + (int)r.size()};
	auto anchors = generate_anchors(q, r, orig, kmer_size);
	auto chains_init = chain_anchors(anchors);
	auto &bounds = chains_init.second;
	auto &chain = chains_init.first;
	std::ostringstream os;
	for (auto &a : anchors) os << "A " << a.q << ' ' << a.r << ' ' << a.l << ' ' << (a.has_u ? 1 : 0) << '\n';
	for (int bi = 1; bi < (int)bounds.size(); bi++) {                   // the loop of src/chain.cc:222-247, called, not changed
		bool has_u = bounds[bi].second;
		int be = bounds[bi].first, bs = bounds[bi - 1].first;
		int qlo = anchors[chain[be - 1]].q, qhi = anchors[chain[bs]].q + anchors[chain[bs]].l;
		int rlo = anchors[chain[be - 1]].r, rhi = anchors[chain[bs]].r + anchors[chain[bs]].l;
		int span = std::max(rhi - rlo, qhi - qlo);
		if ((!has_u || span < Globals::Chain::MIN_UPPERCASE_MATCH) &&
		    span < Globals::Search::MIN_READ_SIZE * (1 - Globals::Search::MAX_ERROR)) continue;
		os << "C " << (be - bs);
		for (int k = be - 1; k >= bs; k--) os << ' ' << chain[k];
		os << '\n';
	}
	std::vector<Hit> hits = fast_align(q, r, orig, kmer_size);
	for (auto &h : hits)
		os << "H " << h.query_start << ' ' << h.query_end << ' ' << h.ref_start << ' ' << h.ref_end << ' ' << h.aln.cigar_string() << ' '
		   << h.aln.span() << ' ' << h.aln.matches() << ' ' << h.aln.mismatches() << ' ' << h.aln.gaps() << ' ' << h.aln.gap_bases() << '\n';
	std::string sres = os.str();
	if ((int)sres.size() + 1 > cap) return -1;
	memcpy(out, sres.c_str(), sres.size() + 1);
	return (int)hits.size();
}
// generate_anchors alone (src/chain.cc:24-101): one line "q r l has_u" per anchor, in the order the reference emits them.
int ref_region_anchors(const char *query, const char *ref, int kmer_size, int same_chr, int orig_qs, int orig_rs, char *out, int cap)
{
	std::string q(query), r(ref);
	auto qp = std::make_shared<Sequence>(same_chr ? "CHR" : "QRY", q);
	auto rp = std::make_shared<Sequence>(same_chr ? "CHR" : "REF", r);
	Hit orig{qp, orig_qs, orig_qs + (int)q.size(), rp, orig_rs, orig_rs + (int)r.size()};
	auto anchors = generate_anchors(q, r, orig, kmer_size);
	std::ostringstream os;
	for (auto &a : anchors) os << a.q << ' ' << a.r << ' ' << a.l << ' ' << (a.has_u ? 1 : 0) << '\n';
	std::string sres = os.str();
	if ((int)sres.size() + 1 > cap) return -1;
	memcpy(out, sres.c_str(), sres.size() + 1);
	return (int)anchors.size();
}
// The final constructor of the refine wave: Alignment(qstr, rstr, vector<Hit> guide, side) (src/align.cc:107-197:
// gap fills between consecutive hits, +-side extensions with trim_front / trim_back, src/align.cc:343-456), run on a
// guide made of the reference's own chain alignments: the chains of the chain wave, sorted, greedily thinned to a
// strictly co-linear, non-overlapping sequence (what refine_chains' merges guarantee, src/refine.cc:164-183).
// Line 1: the result (start_a end_a start_b end_b cigar span matches mismatches gaps gap_bases);
// following lines: the guide hits (query_start query_end ref_start ref_end cigar).
int ref_hit_guide(const char *query, const char *ref, int kmer_size, int side, char *out, int cap)
{
	std::string q(query), r(ref);
	auto qp = std::make_shared<Sequence>("QRY", q);
	auto rp = std::make_shared<Sequence>("REF", r);
	Hit orig{qp, 0, (int)q.size(), rp, 0, (int)r.size()};
	auto anchors = generate_anchors(q, r, orig, kmer_size);
	auto chains_init = chain_anchors(anchors);
	auto &bounds = chains_init.second;
	auto &chain = chains_init.first;
	std::vector<Hit> hits;
	for (int bi = 1; bi < (int)bounds.size(); bi++) {
		int be = bounds[bi].first, bs = bounds[bi - 1].first;
		std::vector<int> guide;
		for (int k = be - 1; k >= bs; k--) guide.emplace_back(chain[k]);
		if (guide.empty()) continue;
		Hit h{qp, 0, 0, rp, 0, 0};
		h.aln = Alignment(q, r, anchors, guide);
		update_from_alignment(h);
		hits.push_back(h);
	}
	std::sort(hits.begin(), hits.end());
	std::vector<Hit> guide;
	for (auto &h : hits)
		if (guide.empty() || (h.query_start >= guide.back().query_end && h.ref_start >= guide.back().ref_end)) guide.push_back(h);
	if (guide.empty()) return 0;
	Alignment res(q, r, guide, side);
	Hit hr{qp, 0, 0, rp, 0, 0};
	hr.aln = res;
	update_from_alignment(hr);
	std::ostringstream os;
	os << hr.query_start << ' ' << hr.query_end << ' ' << hr.ref_start << ' ' << hr.ref_end << ' ' << res.cigar_string() << ' ' << res.span() << ' '
	   << res.matches() << ' ' << res.mismatches() << ' ' << res.gaps() << ' ' << res.gap_bases() << '\n';
	for (auto &h : guide)
		os << h.query_start << ' ' << h.query_end << ' ' << h.ref_start << ' ' << h.ref_end << ' ' << h.aln.cigar_string() << '\n';
	std::string sres = os.str();
	if ((int)sres.size() + 1 > cap) return -1;
	memcpy(out, sres.c_str(), sres.size() + 1);
	return (int)guide.size();
}
// Alignment::merge (src/align.cc:505-610) on overlapping pairs of the reference's own chain alignments (what
// refine_chains does at src/refine.cc:171-173).  Per merged pair three lines: "P ..." prev before, "C ..." cur before,
// "M ..." result; each: start_a end_a start_b end_b cigar (+ span matches mismatches gaps gap_bases for M).
int ref_merge_pairs(const char *query, const char *ref, int kmer_size, char *out, int cap)
{
	std::string q(query), r(ref);
	auto qp = std::make_shared<Sequence>("QRY", q);
	auto rp = std::make_shared<Sequence>("REF", r);
	Hit orig{qp, 0, (int)q.size(), rp, 0, (int)r.size()};
	auto anchors = generate_anchors(q, r, orig, kmer_size);
	auto chains_init = chain_anchors(anchors);
	auto &bounds = chains_init.second;
	auto &chain = chains_init.first;
	std::vector<Hit> hits;
	for (int bi = 1; bi < (int)bounds.size(); bi++) {
		int be = bounds[bi].first, bs = bounds[bi - 1].first;
		std::vector<int> guide;
		for (int k = be - 1; k >= bs; k--) guide.emplace_back(chain[k]);
		if (guide.empty()) continue;
		Hit h{qp, 0, 0, rp, 0, 0};
		h.aln = Alignment(q, r, anchors, guide);
		update_from_alignment(h);
		hits.push_back(h);
	}
	std::sort(hits.begin(), hits.end());
	std::ostringstream os;
	int n = 0;
	auto line = [&](char tag, const Hit &h, bool full) {
		os << tag << ' ' << h.query_start << ' ' << h.query_end << ' ' << h.ref_start << ' ' << h.ref_end << ' ' << h.aln.cigar_string();
		if (full) os << ' ' << h.aln.span() << ' ' << h.aln.matches() << ' ' << h.aln.mismatches() << ' ' << h.aln.gaps() << ' ' << h.aln.gap_bases();
		os << '\n';
	};
	for (size_t i = 0; i < hits.size(); ++i)
		for (size_t j = 0; j < hits.size(); ++j) {
			if (i == j) continue;
			Hit p = hits[i], c = hits[j];
			// the preconditions merge() asserts (src/align.cc:506-508) plus strict progress, as on a refine path
			if (!(c.query_start < p.query_end || c.ref_start < p.ref_end)) continue;
			if (!(p.query_end <= c.query_end && p.ref_end <= c.ref_end)) continue;
			if (!(p.query_start < c.query_start && p.ref_start < c.ref_start)) continue;
			if (p.query_end - c.query_start > 2000 || p.ref_end - c.ref_start > 2000) continue;
			line('P', p, false); line('C', c, false);
			p.aln.merge(c.aln, q, r);
			update_from_alignment(p);
			line('M', p, true);
			if (++n >= 40) goto done;
		}
done:
	std::string sres = os.str();
	if ((int)sres.size() + 1 > cap) return -1;
	memcpy(out, sres.c_str(), sres.size() + 1);
	return n;
}
}
