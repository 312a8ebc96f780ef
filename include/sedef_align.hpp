// sedef_align.hpp -- C++ host-side mirror of SEDEF's Alignment front end for the batched ksw_extz2 engine.
//
// Mirrors (argument meaning, results and error behaviour; not code) of the reference:
//   Alignment::Alignment(fa, fb)            src/align.cc:76-88   (align_dna + align_helper + populate_nice_alignment)
//   Alignment::Alignment(fa, fb, cigar)     src/align.cc:90-105  ("from_cigar")
//   align_helper                             src/align.cc:39-68   (60 kbp chunking, ksw op -> "MDI" remap)
//   getters span/matches/.../total_error     src/align.h:79-92
//   BEDPE stat loop + fp fields of process() src/stats_main.cc:231-283,297-299
// The difference is batching: requests are queued (AlignQueue) and flushed through ONE ksw_extz2_batch call, which is
// what src/chain.cc, src/refine.cc and src/align.cc call sites do once they collect pairs per wave.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <deque>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include "ksw2_b200.h"

namespace sedef_b200 {

struct AlignParams {                       // reference defaults: src/globals.cc:25-28
	int match = 5, mismatch = -4, gap_open = 40, gap_extend = 1, bandwidth = -1;
};

class Alignment {
public:
	std::string a, b;                                  // original-case strings (fa = query, fb = target)
	std::deque<std::pair<char, int>> cigar;            // SEDEF alphabet: M, D (a only), I (b only)
	sd_stats_t stats{};                                // every integer SEDEF derives from the alignment

	int span() const { return stats.span; }
	int matches() const { return stats.matches; }
	int mismatches() const { return stats.mismatches; }
	int gap_bases() const { return stats.gap_bases; }
	int gaps() const { return stats.gaps; }
	double gap_error() const;                          // src/align.h:84-87
	double mismatch_error() const;                     // src/align.h:88-91
	double total_error() const { return mismatch_error() + gap_error(); }
	std::string cigar_string() const;                  // src/align.cc:614-621
	sd_stats_fp_t bedpe_fp() const;                    // src/stats_main.cc:273-283,297-299
};

// Globals::Align::MAX_KSW_SEQ_LEN (src/globals.h:18,54: 60 * KB with KB = 1000): align_helper's chunk length
int max_ksw_seq_len();

// Batched Alignment(fa, fb) for every pair.  Throws std::runtime_error with the engine's message on failure.
std::vector<Alignment> align_batch(const std::vector<std::pair<std::string, std::string>> &pairs, const AlignParams &p = AlignParams());
// Batched Alignment(fa, fb, cigar_string): statistics from an existing CIGAR (no DP), computed on the GPU.
std::vector<Alignment> from_cigar_batch(const std::vector<std::pair<std::string, std::string>> &pairs,
                                        const std::vector<std::string> &cigars);

// ---- the chain wave -----------------------------------------------------------------
// Mirror of Alignment::Alignment(qstr, rstr, vector<Anchor> guide, vector<int> guide_idx) (src/align.cc:199-270) for MANY
// chains at once: every gap between consecutive anchors of every chain becomes one request of a single batched
// ksw_extz2 call ("close" gaps <= 1000 x 1000 as they are, larger ones as the mi x mi prefix fill -- the reference's
// second mi x mi alignment is dead work, its result is never selected, src/align.cc:244), the CIGARs are stitched with
// append_cigar's run merging (src/align.cc:468-477) and the statistics of the stitched alignments come from one
// sd_stats_from_cigar call.
struct Anchor { int q, r, l, has_u; };                  // src/align.h:25-28
struct ChainGuide {
	const std::string *qstr, *rstr;                     // the region pair the chain lives in
	const std::vector<Anchor> *anchors;
	std::vector<int> guide_idx;                         // anchors of the chain, in query order
};
struct GuidedAlignment : Alignment {
	int start_a = 0, end_a = 0, start_b = 0, end_b = 0; // src/align.h: start_a/end_a/start_b/end_b
};
std::vector<GuidedAlignment> align_chains_batch(const std::vector<ChainGuide> &chains, const AlignParams &p = AlignParams());

// ---- the refine wave's final constructor -----------------------------------------------
// Mirror of Alignment::Alignment(qstr, rstr, vector<Hit> guide, side) (src/align.cc:107-197) for MANY guides at once: the
// gap fills between consecutive (co-linear, non-overlapping) hits and the two +-side extensions of every guide -- the
// 500 x 500 unbanded alignments that hold 80-95 % of SEDEF's ksw cells -- go through ONE batched
// ksw_extz2 call; trim_front / trim_back (src/align.cc:343-456) run on the host over the <= 1000 columns of each side
// extension; the statistics of the finished alignments come from one sd_stats_from_cigar call.
struct HitGuide {
	const std::string *qstr, *rstr;
	std::vector<GuidedAlignment> guide;                 // hits in order; only start/end coordinates and cigar are used
	int side = 500;                                     // Globals::Chain::Refine::SIDE_ALIGN
};
std::vector<GuidedAlignment> align_hit_guides_batch(const std::vector<HitGuide> &guides, const AlignParams &p = AlignParams());
// trim_front / trim_back of one alignment whose start/end are relative to its own a/b (as after Alignment(fa, fb))
void trim_front(GuidedAlignment &g, const AlignParams &p = AlignParams());
void trim_back(GuidedAlignment &g, const AlignParams &p = AlignParams());

// Alignment::merge (src/align.cc:505-610) for MANY (prev, cur) pairs: the overlap is trimmed from prev's tail and cur's
// head on the host (column walks on the CIGARs), the gap fills of all pairs go through ONE batched ksw_extz2 call.
// A refine path with several merges is processed level by level (k-th merge of every path per call).
struct MergeRequest { GuidedAlignment prev, cur; const std::string *qstr, *rstr; };
std::vector<GuidedAlignment> merge_batch(const std::vector<MergeRequest> &reqs, const AlignParams &p = AlignParams());

// ---- the region-level driver ----------------------------------------------------------------
// Everything fast_align does AFTER anchoring and chaining (src/chain.cc:249-265 + refine_chains, src/refine.cc:23-193), for MANY
// regions at once, driven as WAVES of batched ksw_extz2 calls instead of one synchronous call per gap:
//   wave 0      every chain of every region:   Alignment(query, ref, anchors, guide_idx)         (align_chains_batch)
//   host        per region: sort, the refine DP over chain alignments, path extraction            (src/refine.cc:27-118)
//   wave k      per region the next step of its current path: one Alignment::merge of two overlapping chain alignments
//               (merge_batch) or the final Alignment(qstr, rstr, guide, SIDE_ALIGN) (align_hit_guides_batch); the steps of
//               one region are sequential (a path's filters look at the hits accepted before it), regions advance together.
// Anchors and chains come from the caller (SEDEF's own generate_anchors / chain_anchors, src/chain.cc:24-199): `guides` are the
// chains that passed the filter of src/chain.cc:222-247, anchors in query order.  Results: per region the refined hits in the
// order refine_chains leaves them, bit-identical to the reference.
struct RegionTask {
	const std::string *qstr, *rstr;
	const std::vector<Anchor> *anchors;
	std::vector<std::vector<int>> guides;
	bool same_chr = false;                              // orig.query->name == orig.ref->name && same strand (src/refine.cc:29-30)
	int orig_query_start = 0, orig_ref_start = 0;       // of the seed hit the region was cut from
};
struct RefineStats {
	int rounds = 0; long long batch_calls = 0, ksw_requests = 0;      // requests: chains / merges / guide constructors
	long long ksw_pairs = 0, ksw_cells = 0;                            // ksw_extz2 pairs of the batched calls and their DP cells (unbanded: qlen x tlen)
};
std::vector<std::vector<GuidedAlignment>> refine_regions_batch(const std::vector<RegionTask> &regions, const AlignParams &p = AlignParams(),
                                                              RefineStats *stats = nullptr);

// ---- anchoring + chaining: with these the region-level driver is a complete fast_align ---------------------------------------
// generate_anchors (src/chain.cc:24-101) for many region pairs on the GPU (sedef_anchors_batch, include/ksw2_b200.h)
struct RegionSeed {
	const std::string *qstr, *rstr;                     // the two region strings (original case)
	bool same_chr = false; int orig_query_start = 0, orig_ref_start = 0;
};
std::vector<std::vector<Anchor>> anchors_batch(const std::vector<RegionSeed> &regions, int kmer_size = 11);
// chain_anchors (src/chain.cc:103-199) on the host: the chaining DP over the anchors of ONE region (sweep over anchor start / end
// events in query order, best predecessor by a range-maximum query over reference end coordinates within MAX_CHAIN_GAP), then the
// chain extraction in score order.  Returns the chains that pass the filter of src/chain.cc:222-247 (an upper-case anchor and a
// span of at least MIN_UPPERCASE_MATCH, or a span of at least 490), anchors in query order -- the `guides` of RegionTask.
std::vector<std::vector<int>> chain_anchors(const std::vector<Anchor> &anchors);
// fast_align (src/chain.cc:203-268) for many regions: anchors on the GPU, chaining on the host (one region per host thread), then
// refine_regions_batch.  Hits per region, in the order the reference returns them.
std::vector<std::vector<GuidedAlignment>> fast_align_batch(const std::vector<RegionSeed> &regions, int kmer_size = 11,
                                                           const AlignParams &p = AlignParams(), RefineStats *stats = nullptr);

// ---- the align stage's driver: `sedef align generate` (BASELINE.json configs 1/4/5) ------------------
// FastaIndex + FastaReference::get_sequence (src/fasta.cc:25-143): .fai index, memory-mapped FASTA
class FastaFile {
public:
	explicit FastaFile(const std::string &filename);     // throws "Cannot open file ...", "Index file ... is malformed at line ..."
	~FastaFile();
	FastaFile(const FastaFile &) = delete;
	FastaFile &operator=(const FastaFile &) = delete;
	// bases [start, *end) of chromosome `name`; *end is clamped to the chromosome length.  Throws "Chromosome ... does not exist".
	std::string get_sequence(const std::string &name, int start, int *end = nullptr) const;
private:
	struct Entry { int length = 0; long long offset = 0; int line_blen = 0, line_len = 0; };
	std::unordered_map<std::string, Entry> index_;
	int fd_ = -1; void *map_ = nullptr; size_t size_ = 0;
};
// Hit (src/hit.h:22-53) as far as the BED text needs it
struct BedHit {
	std::string query_name, ref_name, name, comment;
	int query_start = 0, query_end = 0, ref_start = 0, ref_end = 0, jaccard = 0;
	bool query_rc = false, ref_rc = false;
	static BedHit from_bed(const std::string &bed);      // Hit::from_bed(bed), src/hit.cc:29-60
	// Hit::to_bed(false, with_cigar) (src/hit.cc:134-196); aln == nullptr: a hit without an alignment (span 0)
	std::string to_bed(const Alignment *aln, bool with_cigar = true) const;
};
std::string reverse_complement(const std::string &s);   // rc, src/util.cc:43-48
std::vector<BedHit> read_schedule(const std::string &bed_path);     // bucket_alignments(path, 1, "", false), src/align_main.cc:211-283
struct GenerateStats {
	long long regions = 0, hits = 0, groups = 0, rounds = 0, batch_calls = 0, ksw_requests = 0, region_bytes = 0;
	long long ksw_pairs = 0, ksw_cells = 0;
	double ms_total = 0, ms_align = 0, ms_io = 0;
};
// generate_alignments (src/align_main.cc:285-337): every seed hit of `bed_path` (a bucket file or a directory of *.bed) -> regions
// out of the FASTA -> fast_align_batch on ALL of them together (groups of <= group_bytes of sequence; 0 = default) -> one line
// per refined hit, "<hit.to_bed>\t<seed.to_bed>", byte-identical to the reference binary's output.  With shard_count > 1 only the
// seed hits shard_index, shard_index + shard_count, ... of the schedule are processed (one process per GPU; outputs are merged by
// the caller, as `sedef.sh` merges its per-bucket outputs with sort).
GenerateStats align_generate(const std::string &ref_path, const std::string &bed_path, int kmer_size, FILE *out,
                             const AlignParams &p = AlignParams(), int shard_index = 0, int shard_count = 1, size_t group_bytes = 0);

// ---- the SD report: `sedef stats generate` (src/stats_main.cc:213-395,513-537) ---------------------------------------------------
// populate_nice_alignment's counters + the BEDPE stat loop of many finished alignments (SEDEF-alphabet run lists over their own
// original-case strings) in ONE statistics-from-CIGAR call on the GPU
std::vector<sd_stats_t> stats_of_alignments(const std::vector<const GuidedAlignment *> &alns);
struct StatsParams {                                   // Globals::Stats (src/globals.cc:36-39), the command's --max-ok-gap / --min-split / --uppercase / --max-error
	int max_ok_gap = -1, min_split_size = 1000, min_uppercase = 100;
	double max_scaled_error = 0.5;
};
struct StatsGenerateCounts { long long hits = 0, pieces = 0, lines = 0; };
const char *stats_header();                            // the "#chr1\tstart1..." line (src/stats_main.cc:379-386)
// stats() (src/stats_main.cc:338-395): the aligned hits of `bed_path` (28-column lines of `align generate`, after sedef.sh's
// sort | uniq) -> Alignment(fa, fb, cigar) -> pieces at assembly gaps / large gaps, re-trimmed -> statistics of ALL pieces in one
// GPU call -> filters -> the header and one 34-column line per piece, in the reference's (sequential) order.
StatsGenerateCounts stats_generate(const std::string &ref_path, const std::string &bed_path, FILE *out,
                                   const StatsParams &sp = StatsParams(), const AlignParams &p = AlignParams());

// Deferred-alignment queue: call sites push requests, the driver flushes a whole wave at once.
class AlignQueue {
public:
	explicit AlignQueue(const AlignParams &p = AlignParams()) : params_(p) {}
	// returns the ticket (index into the vector flush() returns)
	size_t push(std::string fa, std::string fb) { reqs_.emplace_back(std::move(fa), std::move(fb)); return reqs_.size() - 1; }
	size_t size() const { return reqs_.size(); }
	std::vector<Alignment> flush() { auto r = align_batch(reqs_, params_); reqs_.clear(); return r; }
private:
	AlignParams params_;
	std::vector<std::pair<std::string, std::string>> reqs_;
};

} // namespace sedef_b200

// The ksw_extz2 calls align_helper makes for one Alignment(fa, fb) with |fa| = alen, |fb| = blen (src/align.cc:46-53):
// call k aligns (fa + sp[k], qlen[k]) against (fb + sp[k], tlen[k]).  Returns the number of calls (fills at most `cap`).
extern "C" int sedef_b200_chunk_plan(int64_t alen, int64_t blen, int cap, int64_t *sp, int *qlen, int *tlen);
// chain_anchors (src/chain.cc:103-199 + the filter of :222-247) for callers without C++: anchors as (q, r, l, has_u) rows; chain k
// holds chain_len[k] anchor indices (query order), concatenated in chain_idx.  Returns the number of chains (fills at most cap_*).
extern "C" int sedef_b200_chain_anchors(int n, const int32_t *anchors4, int cap_chains, int *chain_len, int cap_idx, int *chain_idx);
// `sedef align generate -k kmer_size ref_path bed_path > out_path` (src/align_main.cc:285-337,368-373) through fast_align_batch.
// out_path NULL or "-": stdout.  stats[9] (may be NULL): regions, hits, groups, rounds, batch_calls, ksw_requests, region_bytes,
// ksw_pairs, ksw_cells;
// ms[3] (may be NULL): total, align, io.  Returns 0, or -1 with the message in sedef_b200_align_generate_error().
extern "C" int sedef_b200_align_generate(const char *ref_path, const char *bed_path, int kmer_size, const char *out_path,
                                         int shard_index, int shard_count, long long *stats, double *ms);
extern "C" const char *sedef_b200_align_generate_error(void);
// `sedef stats generate [--max-ok-gap G] [--min-split S] [--uppercase U] [--max-error E] ref_path bed_path > out_path`
// (reference defaults: -1, 1000, 100, 0.5).  counts[3] (may be NULL): hits read, pieces measured, lines written.
extern "C" int sedef_b200_stats_generate(const char *ref_path, const char *bed_path, const char *out_path, int max_ok_gap, int min_split,
                                         int min_uppercase, double max_scaled_error, long long *counts);
extern "C" const char *sedef_b200_stats_generate_error(void);
// host-only half of the report (no device needed): the pieces it would measure, one "qname qs qe rname rs re strand strand span cigar"
// line each (returns the bytes needed, text truncated to cap; -1 on error)
extern "C" long long sedef_b200_stats_pieces(const char *ref_path, const char *bed_path, int max_ok_gap, int min_split, char *out, long long cap);
// host-only pieces of the same driver (no device needed): FastaReference::get_sequence; the seed hits of a bucket file / directory
// in processing order, one Hit::to_bed(false) line each (returns the bytes needed, text truncated to cap); rc() of n bytes
extern "C" long long sedef_b200_fasta_fetch(const char *ref_path, const char *name, int start, int *end_io, char *out, long long cap);
extern "C" long long sedef_b200_bed_schedule(const char *bed_path, char *out, long long cap);
extern "C" void sedef_b200_reverse_complement(const char *in, long long n, char *out);
