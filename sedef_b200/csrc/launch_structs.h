// launch_structs.h -- the plain-data launch descriptors shared by engine.cu (host) and the kernel translation units.
#pragma once
#include <stdint.h>
#include "extz_core.cuh"
#include "../../include/ksw2_b200.h"

namespace extz {

// Device-side view of one DP launch.
struct DpLaunch {
	const PairDesc *pairs;      // [n] sorted by descending work
	PairResult *results;        // [n] (indexed like pairs)
	const uint8_t *seq;         // sequence arena (codes 0..7, one byte per base; offsets in PairDesc)
	uint8_t *tb;                // traceback arena of this wave (or nullptr when score-only)
	const uint32_t *table;      // [kTableStride * kTableStride] (s + 2(q+e)) << 24 per (target, query) symbol
	int *work_counter;          // dynamic work distribution
	int n;
	Scoring sc;
};

// Traceback + statistics launch (extz_tb.cuh).
struct TbLaunch {
	const PairDesc *pairs;
	PairResult *results;
	uint8_t *tb;                 // traceback arena of the wave
	const uint8_t *raw;          // arena of ORIGINAL-CASE bytes (same offsets as the code arena) or nullptr
	const uint8_t *seq;          // code arena (used to synthesise "ACGTN" when raw == nullptr)
	uint32_t *cigar_arena;       // compact output
	unsigned long long *cigar_cursor;
	unsigned long long cigar_capacity;
	sd_stats_t *stats;           // statistics output (see stats_by_orig), or nullptr
	int *overflow;               // set to 1 when the compact arena is too small
	int n, NS, flag;
	int packed;                  // traceback rows written by the packed kernel (extz_dp16.cuh layout)
	int spare;                   // rows carry 16 extra bytes: the codes of the class's SPARE block (extz_dp16.cuh Spare16) at [NS/2, NS/2 + 8)
	int32_t *trims;              // per pair {trim_front max_i | -1, trim_back kept columns | -1} (extz_core.cuh TrimAcc), indexed like
	                             // stats; nullptr: not wanted
	int t_match, t_mismatch, t_gapo, t_gape;   // alignment scoring of the trim scans (Globals::Align, src/align.cc:343-456)
	int stats_by_orig;           // 1: stats[] is indexed by PairDesc::orig (the caller's order), `stats` = base of the whole
	                             //    batch; 0: indexed like pairs, `stats` = base of this wave
};

// Alignment(fa, fb, cigar): SD statistics from existing CIGARs.
struct CigarStatsLaunch {
	const uint32_t *cig; const int64_t *cig_off; const int64_t *n_cig;
	const uint8_t *a; const int64_t *a_off; const int *alen;
	const uint8_t *b; const int64_t *b_off; const int *blen;
	sd_stats_t *out; int *status; int n;
};

// Result gather: PairResult (device order) -> ksw_extz_t records, ready to be copied to the host as they are.
struct GatherLaunch {
	const PairDesc *pairs;
	const PairResult *results;
	uint64_t *ez_out;            // ksw_extz_t records, 7 x 8 bytes each
	uint64_t host_cigar_base;    // HOST address the compact CIGAR arena is copied to (ez.cigar = base + 4 * cigar_off)
	int n;
	int by_orig;                 // 1: record of pair k goes to ez_out[pairs[k].orig], 0: to ez_out[k]
	int with_cigar;
};

} // namespace extz
