// extz_core.cuh -- per-lane building blocks of the B200 ksw_extz2 anti-diagonal kernel.
//
// What is computed is the reference's `ksw_extz2_sse` (extern/ksw2_extz2_sse.cc:23-298): the
// Suzuki-Kasahara difference recurrence over anti-diagonals r = i + j with state (u, v, x, y, s)
// per target slot t, 16-slot block rounding of the band, exact 32-bit H tracking, z-drop.
// How it is computed is B200-specific:
//
//   * ONE PAIR PER LANE GROUP of G lanes (G = 8/16/32, a warp holds 32/G pairs); each lane owns
//     S consecutive slots -> NS = G*S live slots, addressed circularly (slot t lives at t mod NS)
//     so the window slides with the band without moving data.  All state is in REGISTERS.
//   * values are kept in the TOP BYTE of a 32-bit register (v << 24).  32-bit wrap-around is
//     then exactly the reference's int8 wrap-around, signed and unsigned 32-bit min/max/compare
//     are exactly _mm_{max,min}_ep{i,u}8 / _mm_cmpgt_epi8, and every reference vector op is ONE
//     sm_100a integer instruction (IADD3 / VIMNMX / ISETP) per cell.  This matters: fringe cells
//     of the rounded band DO wrap in the reference and (rarely) feed in-band cells, so a wider
//     non-wrapping representation is not bit-exact.  extz_dp16.cuh applies the same idea to
//     16-bit halves (value << 8, two slots per register) for the packed sm_100a instructions.
//   * neighbours: slot t needs the OLD x,v of slot t-1 -> in-lane register of slot i-1, or one
//     __shfl from the circular predecessor lane for slot 0.
//   * the 32-bit H[] row lives in SHARED memory (the only state that needs dynamic slot
//     addressing: H[en0], H[en0-1], H[st0], H[tlen-1]); lanes update their own S entries with
//     vector LDS/STS, special entries are touched by the group leader between __syncwarp()s.
//   * traceback codes are 4 bits per cell (bit0: E beats H, bit1: F beats both, bit2: E
//     continues, bit3: F continues), one coalesced store per lane per diagonal.
//
// EXTZ_HD functions (band geometry, z-drop, tie keys, statistics columns) also compile for the host: engine.cu uses
// them for planning and cell counting.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EXTZ_HD __host__ __device__ __forceinline__
#else
#define EXTZ_HD inline
#endif

namespace extz {

constexpr int kNegInf = -0x40000000;         // KSW_NEG_INF (extern/ksw2.h:6)
constexpr int kArenaSlack = 256;             // readable bytes behind the last sequence of an arena (the query window runs <= 15 bytes past a query)
constexpr int kTableStride = 8;              // score table row stride (entries); symbols must be < 8

enum : int {
	kFlagScoreOnly = 0x01, kFlagRight = 0x02, kFlagGenericSc = 0x04, kFlagApproxMax = 0x08,
	kFlagApproxDrop = 0x10, kFlagExtzOnly = 0x40, kFlagRevCigar = 0x80
};

// Scoring parameters shared by a batch (uniform).
struct Scoring {
	uint32_t q_s;          // q << 24
	uint32_t maxsc_s;      // (mat[0] + 2(q+e)) << 24       (max_sc_ of :69)
	uint32_t s0_s;         // (0 + 2(q+e)) << 24: z of a never-filled slot (s[] starts zeroed, :83)
	int q, e, qe;
	int zdrop, flag, w_in; // w as passed by the caller (<0: unbanded)
	// the same constants for the packed kernel (extz_dp16.cuh): int8 << 8 in BOTH 16-bit halves.  Precomputed on the host so
	// that the kernel can take them straight from the constant bank instead of holding them in registers.
	uint32_t q16, qeps2, maxsc2, s0_2, zr2;
};

// One pair, as the DP kernel sees it.
struct PairDesc {
	int64_t q_off;         // byte offset of query[0] in the sequence arena (no padding between sequences)
	int64_t t_off;         // byte offset of target[0]
	int64_t tb_off;        // byte offset of this pair's traceback rows in the wave's tb arena
	int32_t qlen, tlen;
	int32_t w;             // resolved band (w<0 -> max(qlen,tlen), :71)
	int32_t orig;          // index of the pair in the caller's batch
};

// Result record written by the DP kernel (consumed by the traceback kernel and the host).
struct PairResult {
	int32_t max;           // ez->max (31-bit, >= 0)
	int32_t zdropped;
	int32_t max_q, max_t, mqe, mqe_t, mte, mte_q, score;
	int32_t n_diag;        // anti-diagonals fully processed (R)
	int32_t n_cigar;       // filled by the traceback kernel
	int32_t cigar_off;     // offset (in uint32) into the compact CIGAR arena
};

// ---- band geometry of one anti-diagonal (uniform per pair) -------------------------------
struct Band {
	int st0, en0;          // true in-band slot range          (:106-109,114)
	int st, en;            // 16-rounded DP range               (:115)
	int fe;                // last slot written by the score fill (:125), clamped to T-1
};

EXTZ_HD bool band_of(int r, int qlen, int tlen, int w, int T, bool generic, Band &b)
{
	int st = 0, en = tlen - 1;
	if (st < r - qlen + 1) st = r - qlen + 1;
	if (en > r) en = r;
	if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
	if (en > ((r + w) >> 1)) en = (r + w) >> 1;
	b.st0 = st; b.en0 = en;
	if (st > en) return false;                 // band exhausted -> zdropped = 1 (:110-113)
	b.st = st & ~15; b.en = en | 15;
	int fe = generic ? en : st + (((en - st) >> 4) + 1) * 16 - 1;   // GENERIC_SC fills st0..en0 only (:140-141)
	b.fe = fe < T - 1 ? fe : T - 1;
	return true;
}

// ---- the per-cell recurrence in "top byte" arithmetic ----------------------------------------
// All operands are multiples of 2^24, so every + and - below is the reference's wrapping
// _mm_add_epi8/_mm_sub_epi8 on one byte lane; (int32_t) compares are _mm_cmpgt_epi8/_mm_max_epi8,
// uint32_t compares are _mm_max_epu8/_mm_min_epu8.
template <bool kRight, bool kCigar>
EXTZ_HD uint32_t cell(uint32_t z, uint32_t xt1, uint32_t vt1, uint32_t &U, uint32_t &V, uint32_t &X, uint32_t &Y,
                      const Scoring &sc)
{
	uint32_t a = xt1 + vt1;                                        // :36
	uint32_t ut = U;
	uint32_t b = Y + ut;                                           // :38
	uint32_t code = 0;
	if (!kCigar) {
		z = (int32_t)z > (int32_t)a ? z : a;                       // :153
	} else if (!kRight) {
		if ((int32_t)a > (int32_t)z) code |= 1u;                   // :175  d = a > z ? 1 : 0
		z = (int32_t)z > (int32_t)a ? z : a;                       // :177
		if ((int32_t)b > (int32_t)z) code |= 2u;                   // :178-179  d = b > z ? 2 : d
	} else {
		if (!((int32_t)z > (int32_t)a)) code |= 1u;                // :201  d = z > a ? 0 : 1
		z = (int32_t)z > (int32_t)a ? z : a;                       // :203
		if (!((int32_t)z > (int32_t)b)) code |= 2u;                // :204-205  d = z > b ? d : 2
	}
	z = z > b ? z : b;                                             // :41  _mm_max_epu8
	z = z < sc.maxsc_s ? z : sc.maxsc_s;                           // :42  _mm_min_epu8
	U = z - vt1;                                                   // :43
	V = z - ut;                                                    // :44
	uint32_t zq = z - sc.q_s;                                      // :45
	a -= zq;                                                       // :46
	b -= zq;                                                       // :47
	if (!kRight) {
		bool pa = (int32_t)a > 0, pb = (int32_t)b > 0;             // :187,190 (and :160-161 for score-only)
		X = pa ? a : 0u; Y = pb ? b : 0u;
		if (kCigar) { if (pa) code |= 4u; if (pb) code |= 8u; }    // :189,192
	} else {
		bool na = 0 > (int32_t)a, nb = 0 > (int32_t)b;             // :213,216
		X = na ? 0u : a; Y = nb ? 0u : b;
		if (kCigar) { if (!na) code |= 4u; if (!nb) code |= 8u; }  // :215,218
	}
	return code;
}

// ---- arg-max tie-break key (:226-258) ---------------------------------
// Among slots whose H equals the diagonal maximum the reference picks: en0 itself first, then
// the 4-lane SIMD part [st0, en1) by (lane = (t-st0)&3, then t), then the scalar tail [en1, en0)
// by t.  Smaller key wins.
EXTZ_HD uint32_t tie_key(int t, int st0, int en0)
{
	int en1 = st0 + ((en0 - st0) >> 2) * 4;
	if (t == en0) return 0u;
	if (t < en1) return (1u << 30) | ((uint32_t)((t - st0) & 3) << 28) | (uint32_t)t;
	return (2u << 30) | (uint32_t)t;
}
EXTZ_HD int tie_key_slot(uint32_t key, int en0) { return key == 0u ? en0 : (int)(key & 0x0fffffffu); }

// ---- ksw_apply_zdrop (extern/ksw2.h:161-177), is_rot = 1 -------------------------------------
struct EzState {
	int32_t max, max_t, max_q, mqe, mqe_t, mte, mte_q, score, zdropped;
};
EXTZ_HD void ez_reset(EzState &z)
{
	z.max = 0; z.max_t = z.max_q = z.mqe_t = z.mte_q = -1;
	z.mqe = z.mte = z.score = kNegInf; z.zdropped = 0;
}
// returns true when the extension must stop
EXTZ_HD bool ez_apply_zdrop(EzState &z, int32_t H, int r, int t, int zdrop, int e)
{
	if (H > z.max) {
		z.max = H; z.max_t = t; z.max_q = r - t;
	} else if (t >= z.max_t && r - t >= z.max_q) {
		int tl = t - z.max_t, ql = (r - t) - z.max_q;
		int l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && z.max - H > zdrop + l * e) { z.zdropped = 1; return true; }
	}
	return false;
}

// ---- traceback storage layout ---------------------------------------------------------------------
// Row r of a pair holds NS 4-bit codes, code of slot t at nibble (t mod NS): NS/2 bytes per row.
// Lane L of the group owns nibbles [L*S, L*S+S) = S/2 contiguous bytes, so a group writes one
// contiguous NS/2-byte segment per diagonal (S=8: one 32-bit word per lane, 128 B per warp).
EXTZ_HD int64_t tb_row_bytes(int NS) { return NS >> 1; }
//
// Packed kernel (extz_dp16.cuh): same row size; a lane owns 16 bytes, byte i of lane L holds slot 32L + i in its
// low nibble (block A) and slot 32L + 16 + i in its high nibble (block B).
// Classes with a spare block (extz_dp16.cuh Spare16): rows are NS/2 + 16 bytes; on anti-diagonals whose rounded range spans
// NS/16 + 1 blocks the top block's codes sit behind the packed part, slot st + NS + 2k in the low nibble of byte NS/2 + k.
EXTZ_HD uint32_t tb_fetch(const uint8_t *tb_pair, int NS, int64_t r, int t, bool packed = false, int spare = 0, int st = 0)
{
	const int64_t rowB = (NS >> 1) + (spare ? 16 : 0);
	if (spare && t >= st + NS) {
		uint8_t byte = tb_pair[r * rowB + (NS >> 1) + ((t & 15) >> 1)];
		return (byte >> ((t & 1) * 4)) & 0xfu;
	}
	int c = t & (NS - 1);
	int byte_idx = packed ? (((c >> 5) << 4) | (c & 15)) : (c >> 1);
	int nib = packed ? ((c >> 4) & 1) : (c & 1);
	uint8_t byte = tb_pair[r * rowB + byte_idx];
	return (byte >> (nib * 4)) & 0xfu;
}

// ---- ksw_backtrack (extern/ksw2.h:117-151, is_rot = 1) + fused SD statistics -------------------------
// Emits the CIGAR in REVERSE order (end -> start) exactly like the reference's loop does before
// its final reversal; the caller reverses unless KSW_EZ_REV_CIGAR.  `push` is called per step.
struct CigarSink {
	uint32_t *buf;      // written upward from buf[0] (reverse order)
	int64_t n;
	int32_t gaps;       // number of non-M runs
};
EXTZ_HD void cigar_push(CigarSink &c, uint32_t op, int len)      // ksw_push_cigar, extern/ksw2.h:98-111
{
	if (c.n == 0 || op != (c.buf[c.n - 1] & 0xfu)) {
		c.buf[c.n++] = (uint32_t)len << 4 | op;
		if (op != 0) c.gaps++;
	} else c.buf[c.n - 1] += (uint32_t)len << 4;
}

// SD statistics accumulator (field meaning: include/ksw2_b200.h sd_stats_t).
struct StatAcc {
	int32_t span, gap_bases, matches, mismatches, indel_a, indel_b, alnB, matchB, mismatchB,
	        transitionsB, transversionsB, uppercaseA, uppercaseB, uppercaseMatches;
};
EXTZ_HD int up(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }          // toupper, C locale
EXTZ_HD bool isup(int c) { return c >= 'A' && c <= 'Z'; }                       // isupper, C locale
// one gap-free column (a = query byte, b = target byte, original case)
// returns ceq(a, b): the column is a '|' of Alignment::alignment
EXTZ_HD bool stat_match_col(StatAcc &s, int ca, int cb)
{
	int ua = up(ca), ub = up(cb);
	s.span++; s.alnB++;
	bool ceq = !(ua == 'N' || ub == 'N') && ua == ub;               // src/align.cc:29-35
	if (ceq) s.matches++; else s.mismatches++;                       // src/align.cc:306-314
	s.matchB += (ua == ub);                                          // src/stats_main.cc:249
	s.uppercaseA += (ua != 'N' && isup(ca));                         // :250-252
	s.uppercaseB += (ub != 'N' && isup(cb));                         // :253-255
	if (ua != ub) {                                                  // :258-266
		s.mismatchB++;
		bool bpur = (ub == 'A' || ub == 'G');
		if (ua == 'A' || ua == 'G') { s.transitionsB += bpur; s.transversionsB += !bpur; }
		else { bool bpyr = (ub == 'C' || ub == 'T'); s.transitionsB += bpyr; s.transversionsB += !bpyr; }
	} else if (isup(ca) && isup(cb)) s.uppercaseMatches++;           // :267-269
	return ceq;
}
// column consuming only the query (ksw I, SEDEF 'D': align_b = '-')
EXTZ_HD void stat_qonly_col(StatAcc &s, int ca)
{
	s.span++; s.gap_bases++; s.indel_b++;
	s.uppercaseA += (up(ca) != 'N' && isup(ca));
	// matchB += a != '-' && a == b: b is '-', a is never '-' -> 0
}
// column consuming only the target (ksw D, SEDEF 'I': align_a = '-')
EXTZ_HD void stat_tonly_col(StatAcc &s, int cb)
{
	s.span++; s.gap_bases++; s.indel_a++;
	s.uppercaseB += (up(cb) != 'N' && isup(cb));
}

// ---- Alignment::trim_front / trim_back (src/align.cc:343-456) fused into the traceback walk ------------------------------
// Both are maximum-prefix scans over the alignment's columns with the ALIGNMENT scoring (match, mismatch; a gap run pays
// `open` once plus `ext` per column).  The traceback visits the columns from the LAST to the first, which is trim_front's own
// scan order.  trim_back scans forward; its prefix score up to column i is total - S(i+1), where S(j) = forward score of columns
// j..n-1 -- accumulated on the same backward walk, with the open of a gap run added when the walk LEAVES the run (forward, the
// open belongs to the run's first column) -- so the best prefix is the smallest S, at the largest j on ties.
struct TrimAcc {
	int32_t score_f, max_f, kf;      // trim_front: running score, best score, visit number of the best column (-1: never)
	int32_t S, minS, kmin;           // trim_back: S(j) for j = n - k, its minimum and where
	int32_t prev, k;                 // type of the previously visited column (0 M, 1 a-gap, 2 b-gap, -1 none), columns visited
	int32_t match, mismatch, open, ext;
};
EXTZ_HD void trim_reset(TrimAcc &t, int match, int mismatch, int gap_open, int gap_extend)
{
	t.score_f = 0; t.max_f = 0; t.kf = -1; t.S = 0; t.minS = 0x7fffffff; t.kmin = 0; t.prev = -1; t.k = 0;
	t.match = match; t.mismatch = mismatch; t.open = -gap_open; t.ext = -gap_extend;
}
// one visited column: type 0 = both bases (is_match = ceq), 1 = align_a is '-' (target only), 2 = align_b is '-' (query only)
EXTZ_HD void trim_col(TrimAcc &t, int type, bool is_match)
{
	// trim_back: the column to the right closed a gap run if this one is of another type -> its open is now known
	if (t.prev > 0 && type != t.prev) t.S += t.open;
	if (t.S < t.minS) { t.minS = t.S; t.kmin = t.k; }                  // candidate j = n - k (k = 0: keep everything)
	const int32_t own = type == 0 ? (is_match ? t.match : t.mismatch) : t.ext;
	t.S += own;
	// trim_front (src/align.cc:347-365): a gap column pays the open when the column to its right is not the same kind of gap
	t.score_f += own + ((type != 0 && type != t.prev) ? t.open : 0);
	if (t.score_f >= t.max_f) { t.max_f = t.score_f; t.kf = t.k; }
	t.prev = type; ++t.k;
}
// front: max_i of trim_front (leading columns to drop), or -1 when no suffix scores >= 0 (the reference then keeps its
//        initial max_i = a.size(), sic); back: columns trim_back keeps (max_i + 1), or -1 when no prefix scores >= 0
EXTZ_HD void trim_finish(TrimAcc &t, int32_t &front, int32_t &back)
{
	if (t.prev > 0) t.S += t.open;                                      // column 0 starts its run (i == 0)
	const int32_t n = t.k;
	front = t.kf < 0 ? -1 : n - 1 - t.kf;
	back = (n > 0 && t.S - t.minS >= 0) ? n - t.kmin : -1;
}

} // namespace extz
