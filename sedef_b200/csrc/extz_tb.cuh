// extz_tb.cuh -- K3 + K4 of SURVEY.md section 2.1: device-side traceback -> CIGAR, fused with the SD
// statistics pass (Alignment::populate_nice_alignment + the BEDPE stat loop), one THREAD per pair.
//
// The walk is a sequential pointer chase of <= qlen+tlen dependent 4-bit reads per pair, so it is
// latency-bound per pair and throughput comes from having every lane of every warp on a different
// pair (10^5 pairs in flight hide the L2/HBM latency).  The reversed CIGAR is written IN PLACE
// over the pair's own, already consumed traceback rows (row bytes >= 4 per step, see below), so the
// kernel needs no scratch memory; the finished CIGAR is then copied to a compact arena whose
// cursor is a single atomicAdd per pair.
#pragma once
#include <cuda_runtime.h>
#include "extz_core.cuh"
#include "../../include/ksw2_b200.h"
#include "launch_structs.h"

namespace extz {

constexpr int kPrefetchRows = 16;

__device__ __forceinline__ int raw_byte(const TbLaunch &L, int64_t off, int idx)
{
	if (L.raw) return L.raw[off + idx];
	int c = L.seq[off + idx];
	return c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : c == 3 ? 'T' : 'N';
}

template <bool kStats>
__global__ void __launch_bounds__(128)
extz_traceback_kernel(TbLaunch L)
{
	int pi = blockIdx.x * blockDim.x + threadIdx.x;
	if (pi >= L.n) return;
	const PairDesc pd = L.pairs[pi];
	PairResult pr = L.results[pi];
	const int qlen = pd.qlen, tlen = pd.tlen, w = pd.w, NS = L.NS;
	const int T = (tlen + 15) & ~15;
	const int rowB = (NS >> 1) + (L.spare ? 16 : 0);

	StatAcc sa;
	sa.span = sa.gap_bases = sa.matches = sa.mismatches = sa.indel_a = sa.indel_b = sa.alnB = sa.matchB = 0;
	sa.mismatchB = sa.transitionsB = sa.transversionsB = sa.uppercaseA = sa.uppercaseB = sa.uppercaseMatches = 0;
	TrimAcc ta;                                                     // trim_front / trim_back scans ride on the same walk
	trim_reset(ta, L.t_match, L.t_mismatch, L.t_gapo, L.t_gape);

	int i0, j0; bool run = true;                                   // extern/ksw2_extz2_sse.cc:290-295
	if (!pr.zdropped && !(L.flag & kFlagExtzOnly)) { i0 = tlen - 1; j0 = qlen - 1; }
	else if (pr.max_t >= 0 && pr.max_q >= 0) { i0 = pr.max_t; j0 = pr.max_q; }
	else { i0 = j0 = -1; run = false; }

	uint8_t *tbp = L.tb + pd.tb_off;
	// reversed CIGAR grows downward from the end of the rows that can be visited: entry k lives at
	// end - 4(k+1).  After reading row r, rows >= r are dead, and k+1 <= r0 - r + 1 steps have
	// been pushed, so the entries never reach a live row as long as rowB >= 4.
	uint32_t *cend = (uint32_t *)(tbp + ((int64_t)(run ? i0 + j0 : 0) + 1) * rowB);
	int64_t n = 0; int32_t gaps = 0; uint32_t last = 0;             // `last` caches cend[-n]
	auto push = [&](uint32_t op, int len) {                         // ksw_push_cigar (extern/ksw2.h:98-111)
		if (n == 0 || op != (last & 0xfu)) {
			if (n) cend[-n] = last;
			++n; last = (uint32_t)len << 4 | op;
			if (op != 0) ++gaps;
		} else last += (uint32_t)len << 4;
	};

	if (run) {
		int i = i0, j = j0, state = 0;
		while (i >= 0 && j >= 0) {                                  // extern/ksw2.h:124-144
			int r = i + j;
			// the walk is a dependent chain of nibble reads; pull the rows it will need next into L2/L1 early
			// (the column moves by at most one slot per row, so row r-16 is read within 8 bytes of the current column)
			if (NS >= 512 && r >= kPrefetchRows) {                    // narrow rows (<= 128 B) are adjacent in memory already
				const int pc = (i - kPrefetchRows / 2) & (NS - 1);
				const uint8_t *pf = tbp + (int64_t)(r - kPrefetchRows) * rowB + (L.packed ? (((pc >> 5) << 4) | (pc & 15)) : (pc >> 1));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
			}
			Band b; band_of(r, qlen, tlen, w, T, false, b);
			int force = -1;
			if (i < b.st) force = 2;
			if (i > b.en) force = 1;
			uint32_t tmp = 0;
			if (force < 0) tmp = tb_fetch(tbp, NS, r, i, L.packed != 0, L.spare, b.st);
			int hstate = (tmp & 2u) ? 2 : (int)(tmp & 1u);          // which of H/E/F gave the max
			if (state == 0) state = hstate;
			else if (!((tmp >> (state + 1)) & 1u)) state = 0;       // continuation bits: E -> bit2, F -> bit3
			if (state == 0) state = hstate;
			if (force >= 0) state = force;
			if (state == 0) {
				push(0, 1);
				if (kStats) trim_col(ta, 0, stat_match_col(sa, raw_byte(L, pd.q_off, j), raw_byte(L, pd.t_off, i)));
				--i; --j;
			} else if (state == 1) {
				push(2, 1);                                          // ksw D: consumes the target
				if (kStats) { stat_tonly_col(sa, raw_byte(L, pd.t_off, i)); trim_col(ta, 1, false); }
				--i;
			} else {
				push(1, 1);                                          // ksw I: consumes the query
				if (kStats) { stat_qonly_col(sa, raw_byte(L, pd.q_off, j)); trim_col(ta, 2, false); }
				--j;
			}
		}
		if (i >= 0) {                                               // extern/ksw2.h:145
			push(2, i + 1);
			if (kStats) for (int k = i; k >= 0; --k) { stat_tonly_col(sa, raw_byte(L, pd.t_off, k)); trim_col(ta, 1, false); }
		}
		if (j >= 0) {                                               // extern/ksw2.h:146
			push(1, j + 1);
			if (kStats) for (int k = j; k >= 0; --k) { stat_qonly_col(sa, raw_byte(L, pd.q_off, k)); trim_col(ta, 2, false); }
		}
		if (n) cend[-n] = last;
	}

	// ---- copy to the compact arena: memory order [cend-n, cend) is the FORWARD cigar ----
	unsigned long long off = 0;
	if (n) {
		off = atomicAdd(L.cigar_cursor, (unsigned long long)n);
		if (off + (unsigned long long)n > L.cigar_capacity) { *L.overflow = 1; n = 0; }
	}
	const bool rev = (L.flag & kFlagRevCigar) != 0;
	for (int64_t k = 0; k < n; ++k)
		L.cigar_arena[off + k] = rev ? cend[-1 - k] : cend[-n + k];
	pr.n_cigar = (int32_t)n;
	pr.cigar_off = (int32_t)off;
	L.results[pi].n_cigar = (int32_t)n;
	L.results[pi].cigar_off = (int32_t)off;

	if (kStats && L.stats) {
		sd_stats_t s;
		s.span = sa.span; s.gaps = gaps; s.gap_bases = sa.gap_bases; s.matches = sa.matches; s.mismatches = sa.mismatches;
		s.indel_a = sa.indel_a; s.indel_b = sa.indel_b; s.alnB = sa.alnB; s.matchB = sa.matchB; s.mismatchB = sa.mismatchB;
		s.transitionsB = sa.transitionsB; s.transversionsB = sa.transversionsB;
		s.uppercaseA = sa.uppercaseA; s.uppercaseB = sa.uppercaseB; s.uppercaseMatches = sa.uppercaseMatches; s.reserved = 0;
		L.stats[L.stats_by_orig ? pd.orig : pi] = s;
		if (L.trims) {
			int32_t tf, tb2;
			trim_finish(ta, tf, tb2);
			reinterpret_cast<int2 *>(L.trims)[L.stats_by_orig ? pd.orig : pi] = make_int2(tf, tb2);
		}
	}
}

// ---- long pairs: one WARP per pair ---------------------------------------------------------------------------------
// The walk of a long pair is a chain of dependent reads, each of them a DRAM access (consecutive rows are NS/2 bytes
// apart and the arena does not fit L2): ~0.8 us per step, 80 ms for a 100k-step pair.  Here the warp stages a TILE of 32
// rows first -- lane k loads, from row r-k, the two 16-byte chunks that hold slots [i-31, i] (the column moves by at most
// one slot per row, so every nibble the next 32 rows can ask for is in there) -- and lane 0 then walks the tile out of
// shared memory.  One DRAM latency per 32 rows instead of one per step.  Same state machine, same in-place reversed
// CIGAR (writes stay in rows >= the current one; the tile's rows are already staged), same statistics.
template <bool kStats>
__global__ void __launch_bounds__(128)
extz_traceback_warp_kernel(TbLaunch L)
{
	__shared__ __align__(16) uint8_t sTile[4][32][48];             // per row: the two 16-byte chunks around the column + the spare block's codes
	const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int pi = blockIdx.x * 4 + wid;
	if (pi >= L.n) return;                                         // whole warps leave together
	const PairDesc pd = L.pairs[pi];
	const PairResult pr = L.results[pi];
	const int qlen = pd.qlen, tlen = pd.tlen, w = pd.w, NS = L.NS;
	const int T = (tlen + 15) & ~15;
	const int rowB = (NS >> 1) + (L.spare ? 16 : 0);
	const bool packed = L.packed != 0;

	StatAcc sa;
	sa.span = sa.gap_bases = sa.matches = sa.mismatches = sa.indel_a = sa.indel_b = sa.alnB = sa.matchB = 0;
	sa.mismatchB = sa.transitionsB = sa.transversionsB = sa.uppercaseA = sa.uppercaseB = sa.uppercaseMatches = 0;
	TrimAcc ta;                                                     // trim_front / trim_back scans ride on the same walk
	trim_reset(ta, L.t_match, L.t_mismatch, L.t_gapo, L.t_gape);

	int i0, j0; bool run = true;                                   // extern/ksw2_extz2_sse.cc:290-295
	if (!pr.zdropped && !(L.flag & kFlagExtzOnly)) { i0 = tlen - 1; j0 = qlen - 1; }
	else if (pr.max_t >= 0 && pr.max_q >= 0) { i0 = pr.max_t; j0 = pr.max_q; }
	else { i0 = j0 = -1; run = false; }

	uint8_t *tbp = L.tb + pd.tb_off;
	uint32_t *cend = (uint32_t *)(tbp + ((int64_t)(run ? i0 + j0 : 0) + 1) * rowB);
	int64_t n = 0; int32_t gaps = 0; uint32_t last = 0;             // lane 0 only
	auto push = [&](uint32_t op, int len) {                         // ksw_push_cigar (extern/ksw2.h:98-111)
		if (n == 0 || op != (last & 0xfu)) {
			if (n) cend[-n] = last;
			++n; last = (uint32_t)len << 4 | op;
			if (op != 0) ++gaps;
		} else last += (uint32_t)len << 4;
	};

	int i = i0, j = j0, state = 0;                                  // i, j are warp-uniform at tile boundaries
	while (run && i >= 0 && j >= 0) {
		const int r_top = i + j;
		const int c_lo = (i - 31) & (NS - 1), c_hi = i & (NS - 1);
		const int g_lo = c_lo >> 5, g_hi = c_hi >> 5;               // 32 slots per 16-byte chunk in both row layouts
		if (r_top - lane >= 0) {
			const uint4 *row = (const uint4 *)(tbp + (int64_t)(r_top - lane) * rowB);
			*(uint4 *)&sTile[wid][lane][0] = row[g_lo];
			*(uint4 *)&sTile[wid][lane][16] = row[g_hi];
			if (L.spare) *(uint4 *)&sTile[wid][lane][32] = row[NS >> 5];
		}
		__syncwarp();
		if (lane == 0) {
			while (i >= 0 && j >= 0 && i + j > r_top - 32) {          // extern/ksw2.h:124-144
				const int r = i + j;
				Band b; band_of(r, qlen, tlen, w, T, false, b);
				int force = -1;
				if (i < b.st) force = 2;
				if (i > b.en) force = 1;
				uint32_t tmp = 0;
				if (force < 0) {
					const int c = i & (NS - 1);
					int off = ((c >> 5) == g_hi ? 16 : 0) + (packed ? (c & 15) : ((c >> 1) & 15));
					int nib = packed ? ((c >> 4) & 1) : (c & 1);
					if (L.spare && i >= b.st + NS) { off = 32 + ((i & 15) >> 1); nib = i & 1; }      // the class's spare block
					tmp = (sTile[wid][r_top - r][off] >> (nib * 4)) & 0xfu;
				}
				int hstate = (tmp & 2u) ? 2 : (int)(tmp & 1u);
				if (state == 0) state = hstate;
				else if (!((tmp >> (state + 1)) & 1u)) state = 0;
				if (state == 0) state = hstate;
				if (force >= 0) state = force;
				if (state == 0) {
					push(0, 1);
					if (kStats) trim_col(ta, 0, stat_match_col(sa, raw_byte(L, pd.q_off, j), raw_byte(L, pd.t_off, i)));
					--i; --j;
				} else if (state == 1) {
					push(2, 1);
					if (kStats) { stat_tonly_col(sa, raw_byte(L, pd.t_off, i)); trim_col(ta, 1, false); }
					--i;
				} else {
					push(1, 1);
					if (kStats) { stat_qonly_col(sa, raw_byte(L, pd.q_off, j)); trim_col(ta, 2, false); }
					--j;
				}
			}
		}
		i = __shfl_sync(0xffffffffu, i, 0);
		j = __shfl_sync(0xffffffffu, j, 0);
		__syncwarp();                                                // the tile is rewritten next
	}
	if (lane == 0 && run) {
		if (i >= 0) {                                               // extern/ksw2.h:145
			push(2, i + 1);
			if (kStats) for (int k = i; k >= 0; --k) { stat_tonly_col(sa, raw_byte(L, pd.t_off, k)); trim_col(ta, 1, false); }
		}
		if (j >= 0) {                                               // extern/ksw2.h:146
			push(1, j + 1);
			if (kStats) for (int k = j; k >= 0; --k) { stat_qonly_col(sa, raw_byte(L, pd.q_off, k)); trim_col(ta, 2, false); }
		}
		if (n) cend[-n] = last;
	}
	__syncwarp();                                                   // lane 0's CIGAR stores are visible to the copy below

	// ---- copy to the compact arena (all lanes): memory order [cend-n, cend) is the FORWARD cigar ----
	unsigned long long off = 0;
	if (lane == 0 && n) {
		off = atomicAdd(L.cigar_cursor, (unsigned long long)n);
		if (off + (unsigned long long)n > L.cigar_capacity) { *L.overflow = 1; n = 0; }
	}
	n = __shfl_sync(0xffffffffu, n, 0);
	off = __shfl_sync(0xffffffffu, off, 0);
	const bool rev = (L.flag & kFlagRevCigar) != 0;
	for (int64_t k = lane; k < n; k += 32)
		L.cigar_arena[off + k] = rev ? cend[-1 - k] : cend[-n + k];
	if (lane == 0) {
		L.results[pi].n_cigar = (int32_t)n;
		L.results[pi].cigar_off = (int32_t)off;
		if (kStats && L.stats) {
			sd_stats_t s;
			s.span = sa.span; s.gaps = gaps; s.gap_bases = sa.gap_bases; s.matches = sa.matches; s.mismatches = sa.mismatches;
			s.indel_a = sa.indel_a; s.indel_b = sa.indel_b; s.alnB = sa.alnB; s.matchB = sa.matchB; s.mismatchB = sa.mismatchB;
			s.transitionsB = sa.transitionsB; s.transversionsB = sa.transversionsB;
			s.uppercaseA = sa.uppercaseA; s.uppercaseB = sa.uppercaseB; s.uppercaseMatches = sa.uppercaseMatches; s.reserved = 0;
			L.stats[L.stats_by_orig ? pd.orig : pi] = s;
		if (L.trims) {
			int32_t tf, tb2;
			trim_finish(ta, tf, tb2);
			reinterpret_cast<int2 *>(L.trims)[L.stats_by_orig ? pd.orig : pi] = make_int2(tf, tb2);
		}
		}
	}
}

// ---- Alignment(fa, fb, cigar): SD statistics from an EXISTING CIGAR (src/align.cc:90-105,274-315; the consumers are
// `sedef stats generate`, src/stats_main.cc:224, and the finished alignments of every wave of the region driver).  One WARP per
// alignment: every counter is a sum over columns, and the position of a column in the two strings is the run's start plus an
// offset -- so the lanes split the columns of each CIGAR run (coalesced reads of 32 consecutive bases per side) and the 14
// counters are warp-reduced at the end.  (One thread per alignment walked 10^4 columns serially: 3 ms for a single 20 kbp
// alignment, a fifth of the region driver's time, profiles/r02_tuning.md.) ----
__global__ void __launch_bounds__(128)
sd_stats_from_cigar_kernel(CigarStatsLaunch L)
{
	const int k = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (k >= L.n) return;                                            // whole warps leave together
	StatAcc sa;
	sa.span = sa.gap_bases = sa.matches = sa.mismatches = sa.indel_a = sa.indel_b = sa.alnB = sa.matchB = 0;
	sa.mismatchB = sa.transitionsB = sa.transversionsB = sa.uppercaseA = sa.uppercaseB = sa.uppercaseMatches = 0;
	const uint32_t *c = L.cig + L.cig_off[k];
	const uint8_t *a = L.a + L.a_off[k], *b = L.b + L.b_off[k];
	const int alen = L.alen[k], blen = L.blen[k];
	const int64_t nc = L.n_cig[k];
	int64_t ia = 0, ib = 0;                                          // warp-uniform
	int gaps = 0, bad = 0;
	for (int64_t x = 0; x < nc && !bad; ++x) {
		const uint32_t w = c[x];
		const uint32_t op = w & 0xfu; const int len = (int)(w >> 4);
		if (op != 0) ++gaps;                                         // every non-M run counts, zero-length ones too (src/align.cc:300-305)
		if (op == 0 || op >= 3) {
			// M -- or any other op letter: populate_nice_alignment (src/align.cc:283-297) consumes BOTH strings ("not D", "not I"), so
			// the column is gap-free for the match / mismatch and BEDPE counters, while the run still counts as a gap run
			int64_t ok = len;
			if (ok > alen - ia) ok = alen - ia;
			if (ok > blen - ib) ok = blen - ib;
			if (ok < 0) ok = 0;
			if (ok < len) bad = 1;                                   // the reference asserts (src/align.cc:281-282) / reads past the string
			for (int64_t y = lane; y < ok; y += 32) { stat_match_col(sa, a[ia + y], b[ib + y]); if (op != 0) sa.gap_bases++; }
			ia += len; ib += len;
		} else if (op == 1) {
			for (int64_t y = lane; y < len; y += 32) stat_qonly_col(sa, ia + y < alen ? a[ia + y] : 0);
			ia += len;
		} else {
			for (int64_t y = lane; y < len; y += 32) stat_tonly_col(sa, ib + y < blen ? b[ib + y] : 0);
			ib += len;
		}
	}
	int32_t *f = &sa.span;                                           // 14 consecutive int32 counters
#pragma unroll
	for (int i = 0; i < 14; ++i) f[i] = (int32_t)__reduce_add_sync(0xffffffffu, (uint32_t)f[i]);
	if (lane == 0) {
		sd_stats_t s;
		s.span = sa.span; s.gaps = gaps; s.gap_bases = sa.gap_bases; s.matches = sa.matches; s.mismatches = sa.mismatches;
		s.indel_a = sa.indel_a; s.indel_b = sa.indel_b; s.alnB = sa.alnB; s.matchB = sa.matchB; s.mismatchB = sa.mismatchB;
		s.transitionsB = sa.transitionsB; s.transversionsB = sa.transversionsB;
		s.uppercaseA = sa.uppercaseA; s.uppercaseB = sa.uppercaseB; s.uppercaseMatches = sa.uppercaseMatches; s.reserved = 0;
		L.out[k] = s;
		L.status[k] = bad ? -1 : 0;
	}
}

} // namespace extz
