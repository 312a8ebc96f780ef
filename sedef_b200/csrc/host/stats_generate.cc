// stats_generate.cc -- `sedef stats generate`: the SD report (final.bed) for a whole *.aligned.bed at once.
//
// Mirrors (argument meaning, results, text format; not code) of the reference:
//   stats                      src/stats_main.cc:338-395   read the aligned hits, order them, header, process() each
//   process                    src/stats_main.cc:213-336   Alignment(fa, fb, cigar), split, the BEDPE stat loop, the
//                                                           floating-point columns, the filters, one 34-column line per piece
//   split_alignment / subhit   src/stats_main.cc:32-211    pieces at assembly gaps (>= 100 N columns) and, with --max-ok-gap, at
//                              / gap_split                  large gaps; each piece is re-trimmed (trim_back, trim_front)
// The shape of the work changes: every piece of every hit goes through ONE statistics-from-CIGAR call on the GPU
// (sd_stats_from_cigar_batch_flat: populate_nice_alignment's counters and the BEDPE stat loop, rows a15-a17 of SURVEY.md section 8),
// the floating-point columns are derived on the host from those integers (a18).  What stays on the host is the control logic of
// the splitting: scanning the columns for N runs and the trims of the (rare) split pieces.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <fstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#include "../../../include/sedef_align.hpp"

namespace sedef_b200 {

namespace {

const int kMinAssemblyGap = 100;       // Globals::Stats::MIN_ASSEMBLY_GAP_SIZE (src/globals.h:101)
const int kBigOverlap = 100;           // Globals::Stats::BIG_OVERLAP_THRESHOLD (src/globals.h:102)
const int kMinRead = 900;              // Globals::Chain::Refine::MIN_READ (src/globals.h:84)

inline char up(char c) { return (c >= 'a' && c <= 'z') ? char(c - 32) : c; }
inline bool ceq(char x, char y)        // src/align.cc:29-35: case-insensitive, N never equals anything
{
	if (x == '-' || y == '-') return false;
	if (up(x) == 'N' || up(y) == 'N') return false;
	return up(x) == up(y);
}

// An alignment with its column strings, as far as the splitting needs it.  The reference keeps three strings (align_a, alignment,
// align_b) that are only rebuilt by populate_nice_alignment -- a trim that clears the alignment returns BEFORE rebuilding them, and
// what process() then measures is the stale strings -- so the strings are state here too, not a function of the CIGAR.
struct SAln {
	std::string a, b;
	int start_a = 0, end_a = 0, start_b = 0, end_b = 0;
	std::deque<std::pair<char, int>> cigar;
	std::string col_a, col_b;          // align_a / align_b; column i matches ("|") iff both are bases and ceq
	int matches = 0, mismatches = 0, gaps = 0, gap_bases = 0;        // AlignmentError

	int span() const { return (int)col_a.size(); }
	bool bar(int i) const { return ceq(col_a[i], col_b[i]); }        // alignment[i] == '|'
	void populate()                                                  // src/align.cc:274-315
	{
		col_a.clear(); col_b.clear();
		size_t ia = 0, ib = 0;
		for (auto &c : cigar)
			for (int k = 0; k < c.second; ++k) {
				if (c.first != 'D') { if (ib >= b.size()) throw std::runtime_error("CIGAR overruns the reference sequence"); col_b.push_back(b[ib++]); }
				else col_b.push_back('-');
				if (c.first != 'I') { if (ia >= a.size()) throw std::runtime_error("CIGAR overruns the query sequence"); col_a.push_back(a[ia++]); }
				else col_a.push_back('-');
			}
		matches = mismatches = gaps = gap_bases = 0;
		for (auto &c : cigar) if (c.first != 'M') { ++gaps; gap_bases += c.second; }
		for (size_t i = 0; i < col_a.size(); ++i)
			if (col_a[i] != '-' && col_b[i] != '-') { if (ceq(col_a[i], col_b[i])) ++matches; else ++mismatches; }
	}
	void cigar_from_columns()                                        // cigar_from_alignment, src/align.cc:479-501
	{
		cigar.clear();
		int sz = 0; char op = 0;
		for (size_t i = 0; i < col_a.size(); ++i) {
			const char top = col_a[i] == '-' ? 'I' : (col_b[i] == '-' ? 'D' : 'M');
			if (op != top) { if (op) cigar.push_back({op, sz}); op = top; sz = 0; }
			++sz;
		}
		cigar.push_back({op, sz});
	}
	int col_score(int i, int nb, const AlignParams &p) const          // score of column i; nb: the neighbour that decides "gap opens here"
	{
		if (bar(i)) return p.match;
		if (col_a[i] != '-' && col_b[i] != '-') return p.mismatch;
		int s = 0;
		if (nb < 0 || nb >= span() || (col_a[i] == '-' && col_a[nb] != '-') || (col_b[i] == '-' && col_b[nb] != '-')) s += -p.gap_open;
		return s + -p.gap_extend;
	}
	void trim_back(const AlignParams &p)                              // src/align.cc:400-456   ABCD -> AB--
	{
		int max_score = 0, max_i = -1, score = 0;
		for (int i = 0; i < span(); ++i) {
			score += col_score(i, i - 1, p);
			if (score >= max_score) { max_score = score; max_i = i; }
		}
		if (max_i == -1) { a.clear(); b.clear(); end_a = start_a; end_b = start_b; cigar.clear(); return; }      // (the strings stay)
		++max_i;
		end_a = start_a; end_b = start_b;
		for (int ci = 0, cur = 0; ci < (int)cigar.size(); ++ci) {
			if (cigar[ci].second + cur >= max_i) {
				const int need = max_i - cur;
				cigar[ci].second = need;
				while ((int)cigar.size() - 1 > ci) cigar.pop_back();
				end_a += need; end_b += need;
				break;
			}
			cur += cigar[ci].second;
			if (cigar[ci].first == 'M') { end_a += cigar[ci].second; end_b += cigar[ci].second; }
			else if (cigar[ci].first == 'I') end_b += cigar[ci].second;
			else end_a += cigar[ci].second;
		}
		a = substr_checked(a, start_a, end_a - start_a); b = substr_checked(b, start_b, end_b - start_b);
		populate();
	}
	void trim_front(const AlignParams &p)                             // src/align.cc:343-398   ABCD -> --CD
	{
		int max_score = 0, max_i = (int)a.size(), score = 0;
		for (int i = span() - 1; i >= 0; --i) {
			score += col_score(i, i + 1, p);
			if (score >= max_score) { max_score = score; max_i = i; }
		}
		if (max_i == (int)a.size()) { a.clear(); b.clear(); start_a = end_a; start_b = end_b; cigar.clear(); return; }
		for (int ci = 0, cur = 0; ci < (int)cigar.size(); ++ci) {
			if (cigar[ci].second + cur > max_i) {
				const int need = max_i - cur;
				cigar[ci].second -= need;
				for (int cj = 0; cj < ci; ++cj) cigar.pop_front();
				start_a += need; start_b += need;
				break;
			}
			cur += cigar[ci].second;
			if (cigar[ci].first == 'M') { start_a += cigar[ci].second; start_b += cigar[ci].second; }
			else if (cigar[ci].first == 'I') start_b += cigar[ci].second;
			else start_a += cigar[ci].second;
		}
		a = substr_checked(a, start_a, end_a - start_a); b = substr_checked(b, start_b, end_b - start_b);
		populate();
	}
	static std::string substr_checked(const std::string &s, int pos, int len)
	{
		if (pos < 0 || pos > (int)s.size()) throw std::runtime_error("trim: coordinates outside the sequence (the reference's substr would throw)");
		return s.substr((size_t)pos, (size_t)std::max(0, len));
	}
	double gap_error() const { return 100.0 * gap_bases / double(matches + gap_bases + mismatches); }
};

struct SHit { BedHit h; SAln aln; };

// subhit (src/stats_main.cc:32-85): columns [start, end) of hin as a hit of its own, re-trimmed
bool subhit(const SHit &hin, int start, int end, SHit &out, const AlignParams &p)
{
	const int n = hin.aln.span();
	if (end >= n) end = n;
	if (start >= end) return false;
	out = hin;
	int sa = 0, la = 0, sb = 0, lb = 0;
	for (int i = 0; i < end; ++i) {
		if (out.aln.col_a[i] != '-') { if (i < start) ++sa; else ++la; }
		if (out.aln.col_b[i] != '-') { if (i < start) ++sb; else ++lb; }
	}
	out.aln.col_a = out.aln.col_a.substr(start, end - start);
	out.aln.col_b = out.aln.col_b.substr(start, end - start);
	out.aln.a = SAln::substr_checked(out.aln.a, sa, la); out.aln.start_a = 0; out.aln.end_a = la;
	out.aln.b = SAln::substr_checked(out.aln.b, sb, lb); out.aln.start_b = 0; out.aln.end_b = lb;
	out.aln.cigar_from_columns();
	out.aln.trim_back(p);
	out.aln.trim_front(p);
	out.h.query_start += sa; out.h.query_end = out.h.query_start + la;
	if (out.h.ref_rc) { out.h.ref_start = out.h.ref_end - (lb + sb); out.h.ref_end = out.h.ref_end - sb; }
	else { out.h.ref_start += sb; out.h.ref_end = out.h.ref_start + lb; }
	return true;
}

// gap_split (src/stats_main.cc:87-157): with --max-ok-gap, cut at the largest gap that is at least MIN_SPLIT_SIZE away from both
// ends and whose share of the alignment reaches the threshold; recursively
void gap_split(const SHit &h, const StatsParams &sp, const AlignParams &p, std::vector<SHit> &out)
{
	struct Gap { int start_a, start_b, len_a, len_b, start, len; };
	std::vector<Gap> gaps;
	Gap g{h.aln.start_a, h.aln.start_b, 0, 0, 0, 0};
	for (auto &c : h.aln.cigar) {
		if (c.second && c.first != 'M') {
			if (c.first != 'D') { g.len_a = 0; g.len_b = c.second; }
			else { g.len_b = 0; g.len_a = c.second; }
			g.len = c.second;
			gaps.push_back(g);
		}
		if (c.first != 'D') g.start_b += c.second;
		if (c.first != 'I') g.start_a += c.second;
		g.start += c.second;
	}
	// (std::sort in the reference: the order of equally long gaps is unspecified there; stable here)
	std::stable_sort(gaps.begin(), gaps.end(), [](const Gap &x, const Gap &y) { return x.len > y.len; });
	if (sp.max_ok_gap > -1)
		for (auto &gp : gaps) {
			if (gp.start_a - h.aln.start_a < sp.min_split_size || gp.start_b - h.aln.start_b < sp.min_split_size) continue;
			if (h.aln.end_a - (gp.start_a + gp.len_a) < sp.min_split_size || h.aln.end_b - (gp.start_b + gp.len_b) < sp.min_split_size) continue;
			const double g_score = 100.0 * gp.len / double(h.aln.matches + h.aln.gap_bases + h.aln.mismatches);
			if (g_score >= sp.max_ok_gap) {
				SHit hh;
				if (subhit(h, 0, gp.start, hh, p)) gap_split(hh, sp, p, out);
				if (subhit(h, gp.start + gp.len, h.aln.span(), hh, p)) gap_split(hh, sp, p, out);
				return;
			}
		}
	out.push_back(h);
}

// split_alignment (src/stats_main.cc:159-211)
std::vector<SHit> split_alignment(const SHit &h, const StatsParams &sp, const AlignParams &p)
{
	std::vector<SHit> hits;
	int prev_an = 0, prev_bn = 0, hit_begin = 0;
	SHit hh;
	for (int i = 0; i < h.aln.span(); ++i) {
		if (up(h.aln.col_a[i]) == 'N') ++prev_an;
		else {
			if (prev_an >= kMinAssemblyGap) { if (subhit(h, hit_begin, i - prev_an, hh, p)) hits.push_back(hh); hit_begin = i; }
			prev_an = 0;
		}
		if (up(h.aln.col_b[i]) == 'N') ++prev_bn;
		else {
			if (prev_bn >= kMinAssemblyGap) { if (subhit(h, hit_begin, i - prev_bn, hh, p)) hits.push_back(hh); hit_begin = i; }
			prev_bn = 0;
		}
	}
	if (!hit_begin) hits.push_back(h);
	else if (subhit(h, hit_begin, h.aln.span(), hh, p)) hits.push_back(hh);
	std::vector<SHit> fin;
	for (auto &x : hits) gap_split(x, sp, p, fin);
	return fin;
}

std::string fmt_g(double v)                // fmt 4 "{}" of a double (extern/format.h:2964-3075): sign, nan / inf spelled out, else "%g"
{
	std::string s;
	if (std::signbit(v)) { s = "-"; v = -v; }
	if (std::isnan(v)) return s + "nan";
	if (std::isinf(v)) return s + "inf";
	char buf[64];
	snprintf(buf, sizeof buf, "%g", v);
	return s + buf;
}

} // namespace

const char *stats_header()
{
	return "#chr1\tstart1\tend1\tchr2\tstart2\tend2\tname\tscore\tstrand1\tstrand2\tmax_len\taln_len\tcomment\t"
	       "indel_a\tindel_b\talnB\tmatchB\tmismatchB\ttransitionsB\ttransversions\tfracMatch\tfracMatchIndel\tjck\tk2K\t"
	       "aln_gaps\tuppercaseA\tuppercaseB\tuppercaseMatches\taln_matches\taln_mismatches\taln_gaps\taln_gap_bases\tcigar\tfilter_score";
}

namespace {
// the first half of stats() / process() (src/stats_main.cc:213-231,338-374): read and order the aligned hits, cut the sequences,
// Alignment(fa, fb, cigar), split -- the pieces of every hit that are long enough to be measured.  Host control logic only.
std::vector<std::vector<SHit>> stats_pieces(const std::string &ref_path, const std::string &bed_path, const StatsParams &sp, const AlignParams &p,
                                            long long *n_hits)
{
	FastaFile fr(ref_path);
	std::ifstream fin(bed_path.c_str());
	if (!fin.is_open()) throw std::runtime_error("BED file " + bed_path + " does not exist");
	// ---- read + order (src/stats_main.cc:345-374) ----
	struct In { BedHit h; std::string cigar; };
	std::vector<In> hits;
	std::string s;
	while (std::getline(fin, s)) {
		In in;
		in.h = BedHit::from_bed(s);
		{   // Hit::from_bed(bed, &cigar): column 13
			size_t pos = 0; int col = 0;
			while (col < 12 && (pos = s.find('\t', pos)) != std::string::npos) { ++pos; ++col; }
			if (col == 12 && pos != std::string::npos) { const size_t e = s.find('\t', pos); in.cigar = s.substr(pos, e == std::string::npos ? std::string::npos : e - pos); }
		}
		if (std::tie(in.h.query_name, in.h.query_start, in.h.query_end) > std::tie(in.h.ref_name, in.h.ref_start, in.h.ref_end)) {
			std::swap(in.h.query_name, in.h.ref_name);
			std::swap(in.h.query_start, in.h.ref_start);
			std::swap(in.h.query_end, in.h.ref_end);
			for (char &c : in.cigar) { if (c == 'I') c = 'D'; else if (c == 'D') c = 'I'; }
		}
		hits.push_back(std::move(in));
	}
	std::stable_sort(hits.begin(), hits.end(), [](const In &x, const In &y) {
		return std::tie(x.h.ref_rc, x.h.query_name, x.h.ref_name, x.h.query_start, x.h.ref_start) <
		       std::tie(y.h.ref_rc, y.h.query_name, y.h.ref_name, y.h.query_start, y.h.ref_start);
	});
	if (n_hits) *n_hits = (long long)hits.size();
	// ---- per hit: the two sequences, Alignment(fa, fb, cigar), the pieces (one hit per thread) ----
	std::vector<std::vector<SHit>> pieces(hits.size());
	std::string err;
#pragma omp parallel for schedule(dynamic, 8)
	for (long i = 0; i < (long)hits.size(); ++i) {
		try {
			SHit h;
			h.h = hits[i].h;
			h.aln.a = fr.get_sequence(h.h.query_name, h.h.query_start, &h.h.query_end);
			h.aln.b = fr.get_sequence(h.h.ref_name, h.h.ref_start, &h.h.ref_end);
			if (h.h.query_rc) h.aln.a = reverse_complement(h.aln.a);
			if (h.h.ref_rc) h.aln.b = reverse_complement(h.aln.b);
			h.aln.start_a = 0; h.aln.end_a = (int)h.aln.a.size(); h.aln.start_b = 0; h.aln.end_b = (int)h.aln.b.size();
			int num = 0;
			for (char ch : hits[i].cigar) {                          // src/align.cc:94-103
				if (ch >= '0' && ch <= '9') num = 10 * num + (ch - '0');
				else if (ch == ';') continue;
				else { h.aln.cigar.push_back({ch, num}); num = 0; }
			}
			h.aln.populate();
			std::vector<SHit> ps = split_alignment(h, sp, p);
			for (auto &x : ps)
				if (x.aln.span() >= kMinRead) pieces[i].push_back(std::move(x));
		} catch (const std::exception &e) {
#pragma omp critical
			if (err.empty()) err = e.what();
		}
	}
	if (!err.empty()) throw std::runtime_error(err);
	return pieces;
}
} // namespace

StatsGenerateCounts stats_generate(const std::string &ref_path, const std::string &bed_path, FILE *out, const StatsParams &sp, const AlignParams &p)
{
	StatsGenerateCounts cnt;
	const std::vector<std::vector<SHit>> pieces = stats_pieces(ref_path, bed_path, sp, p, &cnt.hits);
	// ---- ONE statistics call for all pieces: the pieces' own columns as (CIGAR, a, b) ----
	std::vector<GuidedAlignment> flat;
	for (const auto &v : pieces)
		for (const auto &x : v) {
			GuidedAlignment g;
			// (from the column strings, which is what process() walks -- for an ordinary piece they are the piece's a / b / cigar)
			for (char c : x.aln.col_a) if (c != '-') g.a.push_back(c);
			for (char c : x.aln.col_b) if (c != '-') g.b.push_back(c);
			SAln tmp; tmp.col_a = x.aln.col_a; tmp.col_b = x.aln.col_b; tmp.cigar_from_columns();
			g.cigar = tmp.cigar;
			flat.push_back(std::move(g));
		}
	std::vector<const GuidedAlignment *> ptrs;
	for (auto &g : flat) ptrs.push_back(&g);
	const std::vector<sd_stats_t> st = stats_of_alignments(ptrs);
	cnt.pieces = (long long)flat.size();
	// ---- filters + text (src/stats_main.cc:273-335) ----
	std::string text = stats_header();
	text += '\n';
	size_t k = 0;
	for (size_t i = 0; i < pieces.size(); ++i)
		for (const auto &x : pieces[i]) {
			const sd_stats_t &t = st[k++];
			sd_stats_t own = t;                                       // AlignmentError of the piece: the counters its last populate() left
			own.matches = x.aln.matches; own.mismatches = x.aln.mismatches; own.gaps = x.aln.gaps; own.gap_bases = x.aln.gap_bases;
			sd_stats_fp_t fp;
			sd_stats_derive_fp(&own, &fp);
			const BedHit &h = x.h;
			const bool same_chr = h.query_name == h.ref_name && h.query_rc == h.ref_rc;
			const int overlap = !same_chr ? 0 : std::max(0, std::min(h.query_end, h.ref_end) - std::max(h.query_start, h.ref_start));
			bool too_big_overlap = (h.query_end - h.query_start - overlap) < kBigOverlap || (h.ref_end - h.ref_start - overlap) < kBigOverlap;
			too_big_overlap = too_big_overlap && same_chr;
			if (!(t.uppercaseA >= sp.min_uppercase && t.uppercaseB >= sp.min_uppercase && !too_big_overlap &&
			      fp.errorScaled <= sp.max_scaled_error && t.uppercaseMatches >= sp.min_uppercase)) continue;
			BedHit o = h;
			o.name = "S"; o.comment.clear();
			Alignment al;                                            // what to_bed needs: span, the error getters, the cigar string
			al.stats = own; al.stats.span = t.span;
			al.cigar = x.aln.cigar;
			text += o.to_bed(&al, false);
			const int ints1[] = {t.indel_a, t.indel_b, t.alnB, t.matchB, t.mismatchB, t.transitionsB, t.transversionsB};
			for (int v : ints1) { text += '\t'; text += std::to_string(v); }
			const double d1[] = {fp.fracMatch, fp.fracMatchIndel, fp.jcK, fp.k2K};
			for (double v : d1) { text += '\t'; text += fmt_g(v); }
			const int ints2[] = {own.gaps, t.uppercaseA, t.uppercaseB, t.uppercaseMatches, own.matches, own.mismatches, own.gaps, own.gap_bases};
			for (int v : ints2) { text += '\t'; text += std::to_string(v); }
			text += '\t'; text += al.cigar_string();
			text += '\t'; text += fmt_g(1 - fp.errorScaled);
			text += '\n';
			++cnt.lines;
		}
	if (out) { if (fwrite(text.data(), 1, text.size(), out) != text.size()) throw std::runtime_error("write failed"); fflush(out); }
	return cnt;
}

} // namespace sedef_b200

static thread_local std::string g_stats_error;

// Host-only half of the report (no device needed; callers without C++ and the CPU tests): the pieces `stats generate` would measure,
// one line each: "query_name qs qe ref_name rs re strand_q strand_r span cigar" (coordinates, strands and CIGAR as the report prints
// them).  Returns the bytes needed (text truncated to cap), -1 on error.
extern "C" long long sedef_b200_stats_pieces(const char *ref_path, const char *bed_path, int max_ok_gap, int min_split, char *out, long long cap)
{
	try {
		sedef_b200::StatsParams sp;
		sp.max_ok_gap = max_ok_gap; sp.min_split_size = min_split;
		const auto pieces = sedef_b200::stats_pieces(ref_path, bed_path, sp, sedef_b200::AlignParams(), nullptr);
		std::string text;
		for (const auto &v : pieces)
			for (const auto &x : v) {
				sedef_b200::Alignment al; al.cigar = x.aln.cigar;
				text += x.h.query_name + "\t" + std::to_string(x.h.query_start) + "\t" + std::to_string(x.h.query_end) + "\t" + x.h.ref_name + "\t" +
				        std::to_string(x.h.ref_start) + "\t" + std::to_string(x.h.ref_end) + "\t" + (x.h.query_rc ? "-" : "+") + "\t" + (x.h.ref_rc ? "-" : "+") +
				        "\t" + std::to_string(x.aln.span()) + "\t" + al.cigar_string() + "\n";
			}
		const long long n = std::min<long long>((long long)text.size(), cap);
		if (out && n > 0) memcpy(out, text.data(), (size_t)n);
		return (long long)text.size();
	} catch (const std::exception &e) { g_stats_error = e.what(); return -1; }
}
extern "C" const char *sedef_b200_stats_generate_error(void) { return g_stats_error.c_str(); }

// `sedef stats generate [--max-ok-gap G] [--min-split S] [--uppercase U] [--max-error E] ref_path bed_path > out_path`
// (src/stats_main.cc:513-537).  counts[3] (may be NULL): hits read, pieces measured, lines written.  Returns 0 or -1.
extern "C" int sedef_b200_stats_generate(const char *ref_path, const char *bed_path, const char *out_path, int max_ok_gap, int min_split,
                                         int min_uppercase, double max_scaled_error, long long *counts)
{
	FILE *out = stdout;
	try {
		if (!ref_path || !bed_path) throw std::runtime_error("Not enough arguments to stats");
		if (out_path && strcmp(out_path, "-") != 0) {
			out = fopen(out_path, "w");
			if (!out) throw std::runtime_error(std::string("Cannot open file ") + out_path + " for writing");
		}
		sedef_b200::StatsParams sp;
		sp.max_ok_gap = max_ok_gap; sp.min_split_size = min_split; sp.min_uppercase = min_uppercase; sp.max_scaled_error = max_scaled_error;
		const sedef_b200::StatsGenerateCounts c = sedef_b200::stats_generate(ref_path, bed_path, out, sp);
		if (out != stdout) fclose(out);
		if (counts) { counts[0] = c.hits; counts[1] = c.pieces; counts[2] = c.lines; }
		return 0;
	} catch (const std::exception &e) {
		g_stats_error = e.what();
		if (out && out != stdout) fclose(out);
		return -1;
	}
}
