// tests/cpp/align_queue_driver.cc -- exercises the C++ host layer (include/sedef_align.hpp) from C++, no Python:
// reads "fa fb [cigar]" lines on stdin, pushes them through AlignQueue / from_cigar_batch, prints one line per
// request: cigar span matches mismatches gaps gap_bases indel_a indel_b alnB matchB mismatchB ts tv upA upB upM total_error(.1f)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "sedef_align.hpp"

// mode "chains": line 1 = query region, line 2 = reference region, then one line per chain: n  q r l  q r l ...
static int run_chains()
{
	std::string q, r, line;
	std::getline(std::cin, q); std::getline(std::cin, r);
	std::vector<sedef_b200::Anchor> anchors;
	std::vector<sedef_b200::ChainGuide> chains;
	std::vector<std::vector<int>> idx;
	while (std::getline(std::cin, line)) {
		std::istringstream is(line);
		int n; if (!(is >> n)) continue;
		std::vector<int> g;
		for (int k = 0; k < n; ++k) { sedef_b200::Anchor a{}; is >> a.q >> a.r >> a.l; g.push_back((int)anchors.size()); anchors.push_back(a); }
		idx.push_back(g);
	}
	for (auto &g : idx) chains.push_back({&q, &r, &anchors, g});
	std::vector<sedef_b200::GuidedAlignment> res;
	try { res = sedef_b200::align_chains_batch(chains); }
	catch (const std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); return 2; }
	for (auto &a : res)
		printf("%d %d %d %d %s %d %d %d %d %d\n", a.start_a, a.end_a, a.start_b, a.end_b, a.cigar_string().c_str(), a.span(), a.matches(),
		       a.mismatches(), a.gaps(), a.gap_bases());
	return 0;
}

// mode "regions": any number of regions, each: "R same_chr orig_qs orig_rs" / query / reference / "A q r l" anchor lines /
// "C n i0 i1 ..." chain lines / "E".  ALL regions go through ONE refine_regions_batch call.  Output per region: "R k" then one
// "H qs qe rs re cigar span matches mismatches gaps gap_bases" line per refined hit; last line "S rounds batch_calls requests".
static int run_regions()
{
	struct Reg { std::string q, r; std::vector<sedef_b200::Anchor> anchors; sedef_b200::RegionTask t; };
	std::vector<Reg *> regs;
	std::string line;
	Reg *cur = nullptr;
	while (std::getline(std::cin, line)) {
		if (line.empty()) continue;
		std::istringstream is(line);
		char tag; is >> tag;
		if (tag == 'R') {
			cur = new Reg(); regs.push_back(cur);
			int sc; is >> sc >> cur->t.orig_query_start >> cur->t.orig_ref_start; cur->t.same_chr = sc != 0;
			std::getline(std::cin, cur->q); std::getline(std::cin, cur->r);
		} else if (tag == 'A') { sedef_b200::Anchor a{}; is >> a.q >> a.r >> a.l >> a.has_u; cur->anchors.push_back(a); }
		else if (tag == 'C') { int n; is >> n; std::vector<int> g(n); for (int &x : g) is >> x; cur->t.guides.push_back(g); }
	}
	std::vector<sedef_b200::RegionTask> tasks;
	for (Reg *r : regs) { r->t.qstr = &r->q; r->t.rstr = &r->r; r->t.anchors = &r->anchors; tasks.push_back(r->t); }
	sedef_b200::RefineStats st;
	std::vector<std::vector<sedef_b200::GuidedAlignment>> res;
	double best_ms = 1e30;
	const char *reps_env = getenv("REGIONS_REPS");
	const int reps = reps_env ? atoi(reps_env) : 1;
	try {
		for (int rep = 0; rep < reps; ++rep) {
			auto t0 = std::chrono::steady_clock::now();
			res = sedef_b200::refine_regions_batch(tasks, sedef_b200::AlignParams(), &st);
			best_ms = std::min(best_ms, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
		}
	}
	catch (const std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); return 2; }
	fprintf(stderr, "refine_regions_batch: %zu regions, best of %d: %.2f ms\n", tasks.size(), reps, best_ms);
	for (size_t k = 0; k < res.size(); ++k) {
		printf("R %zu\n", k);
		for (auto &a : res[k])
			printf("H %d %d %d %d %s %d %d %d %d %d\n", a.start_a, a.end_a, a.start_b, a.end_b, a.cigar_string().c_str(), a.span(), a.matches(),
			       a.mismatches(), a.gaps(), a.gap_bases());
	}
	printf("S %d %lld %lld\n", st.rounds, st.batch_calls, st.ksw_requests);
	return 0;
}

// mode "fastalign": any number of regions, each: "R same_chr orig_qs orig_rs" / query / reference.  ALL regions go through ONE
// fast_align_batch call (anchors on the GPU, chaining on the host, the region-level driver).  Output as in mode "regions".
static int run_fastalign()
{
	struct Reg { std::string q, r; sedef_b200::RegionSeed s; };
	std::vector<Reg *> regs;
	std::string line;
	while (std::getline(std::cin, line)) {
		if (line.empty() || line[0] != 'R') continue;
		std::istringstream is(line);
		char tag; int sc;
		Reg *cur = new Reg(); regs.push_back(cur);
		is >> tag >> sc >> cur->s.orig_query_start >> cur->s.orig_ref_start; cur->s.same_chr = sc != 0;
		std::getline(std::cin, cur->q); std::getline(std::cin, cur->r);
	}
	std::vector<sedef_b200::RegionSeed> seeds;
	for (Reg *r : regs) { r->s.qstr = &r->q; r->s.rstr = &r->r; seeds.push_back(r->s); }
	sedef_b200::RefineStats st;
	std::vector<std::vector<sedef_b200::GuidedAlignment>> res;
	double best_ms = 1e30;
	const char *reps_env = getenv("REGIONS_REPS");
	const int reps = reps_env ? atoi(reps_env) : 1;
	try {
		for (int rep = 0; rep < reps; ++rep) {
			auto t0 = std::chrono::steady_clock::now();
			res = sedef_b200::fast_align_batch(seeds, 11, sedef_b200::AlignParams(), &st);
			best_ms = std::min(best_ms, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
		}
	}
	catch (const std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); return 2; }
	fprintf(stderr, "fast_align_batch: %zu regions, best of %d: %.2f ms\n", seeds.size(), reps, best_ms);
	for (size_t k = 0; k < res.size(); ++k) {
		printf("R %zu\n", k);
		for (auto &a : res[k])
			printf("H %d %d %d %d %s %d %d %d %d %d\n", a.start_a, a.end_a, a.start_b, a.end_b, a.cigar_string().c_str(), a.span(), a.matches(),
			       a.mismatches(), a.gaps(), a.gap_bases());
	}
	printf("S %d %lld %lld\n", st.rounds, st.batch_calls, st.ksw_requests);
	return 0;
}

// mode "hitguide": line 1 = query region, line 2 = reference region, line 3 = side, then one guide hit per line:
// query_start query_end ref_start ref_end cigar
static int run_hitguide()
{
	std::string q, r, line;
	std::getline(std::cin, q); std::getline(std::cin, r); std::getline(std::cin, line);
	sedef_b200::HitGuide hg; hg.qstr = &q; hg.rstr = &r; hg.side = atoi(line.c_str());
	while (std::getline(std::cin, line)) {
		std::istringstream is(line);
		sedef_b200::GuidedAlignment g; std::string cig;
		if (!(is >> g.start_a >> g.end_a >> g.start_b >> g.end_b >> cig)) continue;
		int num = 0;
		for (char ch : cig) { if (ch >= '0' && ch <= '9') num = 10 * num + (ch - '0'); else { g.cigar.push_back({ch, num}); num = 0; } }
		hg.guide.push_back(g);
	}
	std::vector<sedef_b200::GuidedAlignment> res;
	try { res = sedef_b200::align_hit_guides_batch({hg}); }
	catch (const std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); return 2; }
	for (auto &a : res)
		printf("%d %d %d %d %s %d %d %d %d %d\n", a.start_a, a.end_a, a.start_b, a.end_b, a.cigar_string().c_str(), a.span(), a.matches(),
		       a.mismatches(), a.gaps(), a.gap_bases());
	return 0;
}

static sedef_b200::GuidedAlignment parse_hit(std::istringstream &is)
{
	sedef_b200::GuidedAlignment g; std::string cig;
	is >> g.start_a >> g.end_a >> g.start_b >> g.end_b >> cig;
	int num = 0;
	for (char ch : cig) { if (ch >= '0' && ch <= '9') num = 10 * num + (ch - '0'); else { g.cigar.push_back({ch, num}); num = 0; } }
	return g;
}
// mode "merge": line 1 = query region, line 2 = reference region, then pairs of lines "P ..." / "C ..." (hit format)
static int run_merge()
{
	std::string q, r, line;
	std::getline(std::cin, q); std::getline(std::cin, r);
	std::vector<sedef_b200::MergeRequest> reqs;
	sedef_b200::GuidedAlignment prev;
	while (std::getline(std::cin, line)) {
		std::istringstream is(line);
		char tag; if (!(is >> tag)) continue;
		if (tag == 'P') prev = parse_hit(is);
		else if (tag == 'C') reqs.push_back({prev, parse_hit(is), &q, &r});
	}
	std::vector<sedef_b200::GuidedAlignment> res;
	try { res = sedef_b200::merge_batch(reqs); }
	catch (const std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); return 2; }
	for (auto &a : res)
		printf("M %d %d %d %d %s %d %d %d %d %d\n", a.start_a, a.end_a, a.start_b, a.end_b, a.cigar_string().c_str(), a.span(), a.matches(),
		       a.mismatches(), a.gaps(), a.gap_bases());
	return 0;
}

int main(int argc, char **argv)
{
	if (argc > 1 && std::string(argv[1]) == "regions") return run_regions();
	if (argc > 1 && std::string(argv[1]) == "fastalign") return run_fastalign();
	if (argc > 1 && std::string(argv[1]) == "merge") return run_merge();
	if (argc > 1 && std::string(argv[1]) == "chains") return run_chains();
	if (argc > 1 && std::string(argv[1]) == "hitguide") return run_hitguide();
	const bool from_cigar = argc > 1 && std::string(argv[1]) == "from_cigar";
	std::vector<std::pair<std::string, std::string>> pairs;
	std::vector<std::string> cigars;
	std::string line;
	while (std::getline(std::cin, line)) {
		std::istringstream is(line);
		std::string a, b, c;
		if (!(is >> a >> b)) continue;
		is >> c;
		pairs.emplace_back(a, b); cigars.push_back(c);
	}
	std::vector<sedef_b200::Alignment> res;
	try {
		if (from_cigar) res = sedef_b200::from_cigar_batch(pairs, cigars);
		else {
			sedef_b200::AlignQueue q;
			for (auto &p : pairs) q.push(p.first, p.second);
			res = q.flush();
		}
	} catch (const std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); return 2; }
	for (auto &r : res) {
		const sd_stats_t &s = r.stats;
		printf("%s %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %.1f\n", r.cigar_string().c_str(), s.span, s.matches, s.mismatches, s.gaps,
		       s.gap_bases, s.indel_a, s.indel_b, s.alnB, s.matchB, s.mismatchB, s.transitionsB, s.transversionsB, s.uppercaseA,
		       s.uppercaseB, s.uppercaseMatches, r.total_error());
	}
	return 0;
}
