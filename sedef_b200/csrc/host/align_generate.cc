// align_generate.cc -- the align stage's driver: `sedef align generate` for a whole bucket file (or a directory of them) at once.
//
// Mirrors (argument meaning, results, text format and error messages; not code) of the reference:
//   generate_alignments                 src/align_main.cc:285-337   read the seed hits, cut the two regions out of the FASTA,
//                                                                    fast_align, translate the hits back, print BEDPE lines
//   bucket_alignments(path, 1, "", 0)   src/align_main.cc:211-283   the order the seed hits are processed in (complexity bins)
//   Hit::from_bed / Hit::to_bed         src/hit.cc:29-60,134-196    BED line <-> hit, the 14-column text
//   FastaIndex / FastaReference         src/fasta.cc:25-143         .fai index + get_sequence on the memory-mapped FASTA
//   rc                                  src/util.cc:43-48
// What changes is the shape of the work: the reference calls fast_align for one seed hit at a time (and inside it ksw_extz2_sse
// for one gap at a time); here ALL regions of a bucket go through fast_align_batch together (anchors on the GPU, chaining on the
// host, every alignment wave as one batched ksw_extz2 call), in groups bounded by sequence bytes.  The printed lines are
// byte-identical to the reference binary's (tests/test_align_stage.py compares whole *.aligned.bed files).
#include <errno.h>
#include <fcntl.h>
#include <glob.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../../include/sedef_align.hpp"

namespace sedef_b200 {

namespace {

inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// split() of src/util.cc:33-41: getline on a stringstream -- a trailing delimiter yields NO trailing empty field
std::vector<std::string> split_fields(const std::string &s, char delim)
{
	std::vector<std::string> out;
	std::stringstream ss(s);
	std::string item;
	while (std::getline(ss, item, delim)) out.push_back(item);
	return out;
}

std::string fmt1(double v)                                     // fmt 4 "{:.1f}" is printf's "%.1f"
{
	char buf[64];
	snprintf(buf, sizeof buf, "%.1f", v);
	return buf;
}

mode_t mode_of(const std::string &path)
{
	struct stat st;
	if (stat(path.c_str(), &st) != 0) return 0;
	return st.st_mode;
}

} // namespace

// ---- FASTA (src/fasta.cc) --------------------------------------------------------------------------------------------------
FastaFile::FastaFile(const std::string &filename)
{
	fd_ = open(filename.c_str(), O_RDONLY);
	if (fd_ < 0) throw std::runtime_error("Cannot open file " + filename);
	const std::string index_name = filename + ".fai";
	struct stat st;
	if (stat(index_name.c_str(), &st) == 0) {                  // without an index every lookup fails, like the reference
		std::ifstream fin(index_name.c_str());
		if (!fin.is_open()) throw std::runtime_error("Index file " + index_name + " does not exist");
		std::string line;
		long long linenum = 0;
		while (std::getline(fin, line)) {
			++linenum;
			const std::vector<std::string> f = split_fields(line, '\t');
			if (f.size() != 5) throw std::runtime_error("Index file " + index_name + " is malformed at line " + std::to_string(linenum));
			const std::vector<std::string> first = split_fields(f[0], ' ');
			if (first.empty()) throw std::runtime_error("Index file " + index_name + " is malformed at line " + std::to_string(linenum));
			Entry e;
			e.length = atoi(f[1].c_str()); e.offset = strtoll(f[2].c_str(), nullptr, 10);
			e.line_blen = atoi(f[3].c_str()); e.line_len = atoi(f[4].c_str());
			index_.emplace(first[0], e);                       // keyed by the first token of the name; the first entry of a name wins
		}
	}
	if (fstat(fd_, &st) == -1) { close(fd_); throw std::runtime_error("Cannot stat file " + filename); }
	size_ = (size_t)st.st_size;
	map_ = size_ ? mmap(nullptr, size_, PROT_READ, MAP_SHARED, fd_, 0) : nullptr;
	if (map_ == MAP_FAILED) { close(fd_); throw std::runtime_error("Cannot map file " + filename); }
}

FastaFile::~FastaFile()
{
	if (map_) munmap(map_, size_);
	if (fd_ >= 0) close(fd_);
}

// FastaReference::get_sequence (src/fasta.cc:106-143): [start, *end) of a chromosome, *end clamped to its length (and written
// back); the line feeds inside the byte range are dropped
std::string FastaFile::get_sequence(const std::string &name, int start, int *end) const
{
	auto it = index_.find(name);
	if (it == index_.end()) throw std::runtime_error("Chromosome " + name + " does not exist");
	const Entry &e = it->second;
	if (start < 0) start = 0;
	int length;
	if (end == nullptr || *end > e.length) {
		length = e.length - start;
		if (end != nullptr) *end = e.length;
	} else length = *end - start;
	if (length <= 0 || e.line_blen <= 0) return std::string();
	const long long newlines_before = start > 0 ? (start - 1) / e.line_blen : 0;
	const long long newlines_by_end = (start + (long long)length - 1) / e.line_blen;
	const long long seqlen = length + (newlines_by_end - newlines_before);
	long long from = e.offset + newlines_before + start;
	long long to = from + seqlen;
	if (from > (long long)size_) from = (long long)size_;
	if (to > (long long)size_) to = (long long)size_;
	std::string s;
	s.reserve((size_t)length);
	const char *p = (const char *)map_;
	for (long long i = from; i < to;) {                            // line by line; std::remove of '\n' and '\0' (src/fasta.cc:137-138)
		const char *nl = (const char *)memchr(p + i, '\n', (size_t)(to - i));
		const long long stop = nl ? (long long)(nl - p) : to;
		if (memchr(p + i, '\0', (size_t)(stop - i))) {
			for (long long k = i; k < stop; ++k) if (p[k] != '\0') s.push_back(p[k]);
		} else s.append(p + i, (size_t)(stop - i));
		i = stop + 1;
	}
	return s;
}

// ---- BED hits (src/hit.cc) -----------------------------------------------------------------------------------------------------
BedHit BedHit::from_bed(const std::string &bed)
{
	const std::vector<std::string> ss = split_fields(bed, '\t');
	if (ss.size() < 10) throw std::runtime_error("BED line with fewer than 10 columns: " + bed);
	BedHit h;
	h.query_name = ss[0]; h.query_rc = ss[8].empty() || ss[8][0] != '+';
	h.ref_name = ss[3]; h.ref_rc = ss[9].empty() || ss[9][0] != '+';
	h.query_start = atoi(ss[1].c_str()); h.query_end = atoi(ss[2].c_str());
	h.ref_start = atoi(ss[4].c_str()); h.ref_end = atoi(ss[5].c_str());
	h.name = ss[6];
	if (ss.size() >= 15) h.comment = ss[14];
	if (ss.size() >= 14) h.jaccard = atoi(ss[13].c_str());
	return h;
}

// Hit::to_bed(do_rc = false, with_cigar, fr = nullptr) (src/hit.cc:134-196)
std::string BedHit::to_bed(const Alignment *aln, bool with_cigar) const
{
	const int span = aln ? aln->span() : 0;
	std::string s;
	s.reserve(256 + (aln && with_cigar ? aln->cigar.size() * 6 : 0));
	s += query_name; s += '\t'; s += std::to_string(query_start); s += '\t'; s += std::to_string(query_end); s += '\t';
	s += ref_name; s += '\t'; s += std::to_string(ref_start); s += '\t'; s += std::to_string(ref_end); s += '\t';
	s += name; s += '\t';
	if (span) s += fmt1(aln->total_error());
	s += '\t';
	s += query_rc ? "-" : "+"; s += '\t'; s += ref_rc ? "-" : "+"; s += '\t';
	s += std::to_string(std::max(query_end - query_start, ref_end - ref_start)); s += '\t';
	s += std::to_string(span); s += '\t';
	if (with_cigar) { if (aln) s += aln->cigar_string(); s += '\t'; }
	if (span) { s += "m="; s += fmt1(aln->mismatch_error()); s += ";g="; s += fmt1(aln->gap_error()); }
	if (!comment.empty()) { s += ';'; s += comment; }
	return s;
}

std::string reverse_complement(const std::string &s)               // rc, src/util.cc:43-48 with rev_dna of src/common.h:72-93
{
	std::string r(s.size(), 'N');
	for (size_t i = 0; i < s.size(); ++i) {
		char c;
		switch (s[s.size() - 1 - i]) {
		case 'A': c = 'T'; break; case 'a': c = 't'; break;
		case 'C': c = 'G'; break; case 'c': c = 'g'; break;
		case 'G': c = 'C'; break; case 'g': c = 'c'; break;
		case 'T': c = 'A'; break; case 't': c = 'a'; break;
		default: c = 'N';
		}
		r[i] = c;
	}
	return r;
}

// the seed hits of a BED file (or of every *.bed in a directory) in the order generate_alignments processes them:
// bucket_alignments(path, 1, "", false) bins them by (int)sqrt(query span * ref span) / 1000 and walks the bins in order
std::vector<BedHit> read_schedule(const std::string &bed_path)
{
	std::vector<std::string> files;
	const mode_t mode = mode_of(bed_path);
	if (S_ISREG(mode)) files.push_back(bed_path);
	else if (S_ISDIR(mode)) {
		glob_t g;
		glob((bed_path + "/*.bed").c_str(), GLOB_TILDE, nullptr, &g);
		for (size_t i = 0; i < g.gl_pathc; ++i)
			if (S_ISREG(mode_of(g.gl_pathv[i]))) files.push_back(g.gl_pathv[i]);
		globfree(&g);
	} else throw std::runtime_error("Path " + bed_path + " is neither file nor directory");
	std::vector<BedHit> hits;
	for (const std::string &f : files) {
		std::ifstream fin(f.c_str());
		if (!fin.is_open()) throw std::runtime_error("BED file " + bed_path + " does not exist");
		std::string s;
		while (std::getline(fin, s)) hits.push_back(BedHit::from_bed(s));
	}
	auto complexity = [](const BedHit &h) {
		return (int)sqrt(double(h.query_end - h.query_start) * double(h.ref_end - h.ref_start));
	};
	int max_complexity = 0;
	for (const BedHit &h : hits) max_complexity = std::max(max_complexity, complexity(h));
	std::vector<std::vector<BedHit>> bins(max_complexity / 1000 + 1);
	for (BedHit &h : hits) {
		const int c = complexity(h) / 1000;
		bins[c < 0 ? 0 : c].push_back(std::move(h));
	}
	std::vector<BedHit> order;
	order.reserve(hits.size());
	for (auto &bin : bins)
		for (BedHit &h : bin) order.push_back(std::move(h));
	return order;
}

// generate_alignments (src/align_main.cc:285-337) for the seed hits [shard_index :: shard_count] of the schedule
GenerateStats align_generate(const std::string &ref_path, const std::string &bed_path, int kmer_size, FILE *out,
                             const AlignParams &p, int shard_index, int shard_count, size_t group_bytes)
{
	GenerateStats gs;
	const double t_start = now_ms();
	std::vector<BedHit> schedule = read_schedule(bed_path);
	if (shard_count > 1) {                                       // one process per GPU: every rank takes its residue class
		std::vector<BedHit> mine;
		for (size_t i = 0; i < schedule.size(); ++i)
			if ((int)(i % (size_t)shard_count) == shard_index) mine.push_back(std::move(schedule[i]));
		schedule.swap(mine);
	}
	FastaFile fr(ref_path);
	if (group_bytes == 0) group_bytes = (size_t)768 << 20;
	size_t at = 0;
	while (at < schedule.size()) {
		// one group of regions: cut the strings, then the whole group advances through fast_align_batch together
		std::vector<std::string> fa, fb;
		size_t bytes = 0, end = at;
		const double t0 = now_ms();
		while (end < schedule.size() && (end == at || bytes < group_bytes)) {      // the group: by the (unclamped) spans
			const BedHit &h = schedule[end];
			bytes += (size_t)std::max(0, h.query_end - h.query_start) + (size_t)std::max(0, h.ref_end - h.ref_start);
			++end;
		}
		fa.resize(end - at); fb.resize(end - at);
		std::string cut_error;
#pragma omp parallel for schedule(dynamic, 4)
		for (long i = (long)at; i < (long)end; ++i) {
			try {
				BedHit &h = schedule[i];
				fa[i - at] = fr.get_sequence(h.query_name, h.query_start, &h.query_end);
				std::string b = fr.get_sequence(h.ref_name, h.ref_start, &h.ref_end);
				fb[i - at] = h.ref_rc ? reverse_complement(b) : std::move(b);
			} catch (const std::exception &e) {
#pragma omp critical
				if (cut_error.empty()) cut_error = e.what();
			}
		}
		if (!cut_error.empty()) throw std::runtime_error(cut_error);
		bytes = 0;
		for (size_t i = 0; i < fa.size(); ++i) bytes += fa[i].size() + fb[i].size();
		std::vector<RegionSeed> seeds(end - at);
		for (size_t i = at; i < end; ++i) {
			const BedHit &h = schedule[i];
			RegionSeed &s = seeds[i - at];
			s.qstr = &fa[i - at]; s.rstr = &fb[i - at];
			s.same_chr = h.query_name == h.ref_name && h.query_rc == h.ref_rc;      // src/chain.cc:49-50, src/refine.cc:29-30
			s.orig_query_start = h.query_start; s.orig_ref_start = h.ref_start;
		}
		const double t1 = now_ms();
		RefineStats rs;
		std::vector<std::vector<GuidedAlignment>> hits = fast_align_batch(seeds, kmer_size, p, &rs);
		const double t2 = now_ms();
		gs.rounds += rs.rounds; gs.batch_calls += rs.batch_calls; gs.ksw_requests += rs.ksw_requests;
		gs.ksw_pairs += rs.ksw_pairs; gs.ksw_cells += rs.ksw_cells;
		std::vector<std::string> texts(end - at);
#pragma omp parallel for schedule(dynamic, 4)
		for (long i = (long)at; i < (long)end; ++i) {
			const BedHit &h = schedule[i];
			const std::string orig_text = h.to_bed(nullptr, true);
			std::string &text = texts[i - at];
			for (const GuidedAlignment &g : hits[i - at]) {
				BedHit hh;                                       // the refined hit in genome coordinates (src/align_main.cc:314-329)
				hh.query_name = h.query_name; hh.ref_name = h.ref_name;
				hh.query_rc = false;                             // fast_align's "QRY" sequence is never reverse-complemented
				hh.ref_rc = h.ref_rc;
				hh.query_start = g.start_a + h.query_start; hh.query_end = g.end_a + h.query_start;
				if (h.ref_rc) { hh.ref_start = h.ref_end - g.end_b; hh.ref_end = h.ref_end - g.start_b; }
				else { hh.ref_start = g.start_b + h.ref_start; hh.ref_end = g.end_b + h.ref_start; }
				text += hh.to_bed(&g, true); text += '\t'; text += orig_text; text += '\n';
			}
		}
		std::string text;
		for (size_t i = 0; i < texts.size(); ++i) { text += texts[i]; gs.hits += (long long)hits[i].size(); }
		if (out && !text.empty() && fwrite(text.data(), 1, text.size(), out) != text.size())
			throw std::runtime_error(std::string("write failed: ") + strerror(errno));
		gs.regions += (long long)(end - at); gs.region_bytes += (long long)bytes; ++gs.groups;
		gs.ms_io += (t1 - t0) + (now_ms() - t2); gs.ms_align += t2 - t1;
		at = end;
	}
	if (out) fflush(out);
	gs.ms_total = now_ms() - t_start;
	return gs;
}

} // namespace sedef_b200

// ---- C ABI ---------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_generate_error;

extern "C" const char *sedef_b200_align_generate_error(void) { return g_generate_error.c_str(); }

// `sedef align generate -k kmer_size ref_path bed_path > out_path` (src/align_main.cc:285-337, 368-373).  out_path NULL or "-":
// stdout.  stats[9] (may be NULL): regions, hits, groups, rounds, batch_calls, ksw_requests, region_bytes, ksw_pairs, ksw_cells.  Returns 0, or -1 with
// the message in sedef_b200_align_generate_error() (the reference throws its message and exits).
extern "C" int sedef_b200_align_generate(const char *ref_path, const char *bed_path, int kmer_size, const char *out_path,
                                         int shard_index, int shard_count, long long *stats, double *ms)
{
	FILE *out = stdout;
	try {
		if (!ref_path || !bed_path) throw std::runtime_error("Not enough arguments to align");
		if (out_path && strcmp(out_path, "-") != 0) {
			out = fopen(out_path, "w");
			if (!out) throw std::runtime_error(std::string("Cannot open file ") + out_path + " for writing");
		}
		const sedef_b200::GenerateStats gs = sedef_b200::align_generate(ref_path, bed_path, kmer_size, out, sedef_b200::AlignParams(),
		                                                                shard_index, shard_count < 1 ? 1 : shard_count, 0);
		if (out != stdout) fclose(out);
		if (stats) {
			stats[0] = gs.regions; stats[1] = gs.hits; stats[2] = gs.groups; stats[3] = gs.rounds;
			stats[4] = gs.batch_calls; stats[5] = gs.ksw_requests; stats[6] = gs.region_bytes;
			stats[7] = gs.ksw_pairs; stats[8] = gs.ksw_cells;
		}
		if (ms) { ms[0] = gs.ms_total; ms[1] = gs.ms_align; ms[2] = gs.ms_io; }
		return 0;
	} catch (const std::exception &e) {
		g_generate_error = e.what();
		if (out && out != stdout) fclose(out);
		return -1;
	}
}

// ---- small host-only entry points (no device needed): the text / FASTA layer on its own, for callers without C++ and the CPU tests
// FastaReference::get_sequence.  Returns the number of bases copied to out (at most cap), *end_io clamped like the reference; -1 on error.
extern "C" long long sedef_b200_fasta_fetch(const char *ref_path, const char *name, int start, int *end_io, char *out, long long cap)
{
	try {
		sedef_b200::FastaFile fr(ref_path);
		const std::string s = fr.get_sequence(name, start, end_io);
		const long long n = std::min<long long>((long long)s.size(), cap);
		if (out && n > 0) memcpy(out, s.data(), (size_t)n);
		return n;
	} catch (const std::exception &e) { g_generate_error = e.what(); return -1; }
}
// The schedule of a bucket file / directory as generate_alignments walks it: one "Hit::to_bed(false)" line per seed hit.
// Returns the number of bytes needed (text is truncated to cap); -1 on error.
extern "C" long long sedef_b200_bed_schedule(const char *bed_path, char *out, long long cap)
{
	try {
		std::string text;
		for (const sedef_b200::BedHit &h : sedef_b200::read_schedule(bed_path)) { text += h.to_bed(nullptr, true); text += '\n'; }
		const long long n = std::min<long long>((long long)text.size(), cap);
		if (out && n > 0) memcpy(out, text.data(), (size_t)n);
		return (long long)text.size();
	} catch (const std::exception &e) { g_generate_error = e.what(); return -1; }
}
// rc (src/util.cc:43-48) of n bytes
extern "C" void sedef_b200_reverse_complement(const char *in, long long n, char *out)
{
	const std::string r = sedef_b200::reverse_complement(std::string(in, (size_t)n));
	memcpy(out, r.data(), (size_t)n);
}
