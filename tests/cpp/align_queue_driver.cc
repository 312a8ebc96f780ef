// tests/cpp/align_queue_driver.cc -- exercises the C++ host layer (include/sedef_align.hpp) from C++, no Python:
// reads "fa fb [cigar]" lines on stdin, pushes them through AlignQueue / from_cigar_batch, prints one line per
// request: cigar span matches mismatches gaps gap_bases indel_a indel_b alnB matchB mismatchB ts tv upA upB upM total_error(.1f)
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "sedef_align.hpp"

int main(int argc, char **argv)
{
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
