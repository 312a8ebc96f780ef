// extz_dp16.cuh -- the PACKED anti-diagonal DP kernel: two cells per 32-bit register lane-op.
//
// Same recurrence, band geometry, bookkeeping and bit-exactness contract as extz_dp.cuh (which see); what changes is
// the representation.  Every u/v/x/y/z value is the reference's int8 kept in the TOP BYTE OF A 16-BIT HALF (value << 8,
// low byte 0), two slots per register.  16-bit wrap-around of a multiple of 256 is exactly the reference's int8
// wrap-around, signed / unsigned 16-bit compares are exactly _mm_cmpgt_epi8 / _mm_max_epu8 / _mm_min_epu8, and the
// sm_100a packed integer instructions (VIADD.16x2, VIMNMX.{S,U}16x2 with per-half predicate outputs) process both
// slots in ONE ALU-pipe issue slot -- the pipe that bounds this kernel (profiles/r01_int_peak.json: VIADD.16x2 and
// VIMNMX.16x2 issue at the same 64 lanes/clk/SM as their 32-bit forms).
//
// sm_100a has no packed subtract.  z - v is computed as ~(~z + v): per half ~X = -X - 1, so ~z + v = -(z - v) - 1 and
// the outer complement gives z - v with a clean (zero) low byte; z - q is folded into one add of ~z + (q + 1).
//
// Lane layout: a lane owns TWO 16-slot blocks, block A in the low halves and block B in the high halves of its 16
// state registers, i.e. register i holds slots (tA + i, tB + i).  The blocks are the 2G "virtual lanes" of the pair:
// virtual lane 2l is block A of lane l, 2l+1 its block B; virtual lane v owns the slots congruent to [16v, 16v+16)
// modulo NS = 32 G and slides by NS when the band has left it -- each block independently, so the window holds NS
// live slots like the S = 16 classes of extz_dp.cuh.  Neighbour slot i-1 is register i-1 for both halves at once; the
// carries into register 0 are the predecessor lane's block-B top (one shuffle, into the low half) and the lane's own
// block-A top (into the high half).
//
// There is no per-block "active" branch on the cell path: all 32 slots of a lane are computed on every anti-diagonal.
// Blocks outside the rounded band [st, en] compute garbage, which is harmless because
//   * a block below st is never read again (its storage is re-initialised when it slides),
//   * a block above en is zeroed (u, v, x, y) on the anti-diagonal it enters the band -- the state calloc() gave it
//     in the reference (extern/ksw2_extz2_sse.cc:83) -- while its z (the persistent s[] of :124-141, which the
//     reference's score fill does write up to 15 slots beyond en) is never touched by the cell code,
//   * their lazy-H entries sit at ~KSW_NEG_INF and their traceback nibbles are never visited.
#pragma once
#include <cuda_runtime.h>
#include "extz_dp.cuh"

namespace extz {

struct Sc16 {
	uint32_t q16;        // q << 8
	uint32_t qeps2;      // (q << 8) + 1 in both halves: ~z + qeps2 == q - z
	uint32_t maxsc2;     // max_sc_ (:69) in both halves
	uint32_t s0_2;       // z of a never-filled slot in both halves
	uint32_t zr;         // 0 held in a register: as an immediate ptxas re-materialises it (PRMT RZ) for every VIADDMNMX
};
__device__ __forceinline__ Sc16 make_sc16(const Scoring &sc)      // host-computed (engine.cu build_scoring)
{
	Sc16 s;
	s.q16 = sc.q16; s.qeps2 = sc.qeps2; s.maxsc2 = sc.maxsc2; s.s0_2 = sc.s0_2; s.zr = sc.zr2;
	return s;
}

// ---- the recurrence for the two slots of one register (:36-47 and the traceback arms :175-218) ----
// SH: bit position of this register's code byte in the code word `cw` (low nibble: block A slot, high nibble: block B).
// __vibmax_s16x2(a, b, &hi, &lo) returns max(a, b) per half and the predicates (a >= b).
// CAUTION (ptxas 12.9, tools/probes/vibmax_probe.cu): when the FIRST operand is a compile-time constant ptxas commutes
// the operands of the fused VIMNMX and the predicates come out wrong; a constant SECOND operand is fine.  Only the
// right-aligned arm uses the predicate form (all its tests are ">=" with the variable first).
template <bool kRight, bool kCigar, int SH>
__device__ __forceinline__ void cell2(uint32_t z, uint32_t xt1, uint32_t vt1, uint32_t &U, uint32_t &V, uint32_t &X, uint32_t &Y,
                                      const Sc16 &sc, uint32_t &cw)
{
	const uint32_t ut = U;
	const uint32_t a = __vadd2(xt1, vt1);                                  // :36
	const uint32_t b = __vadd2(Y, ut);                                     // :38
	bool h0 = false, l0 = false, h1 = false, l1 = false, h2 = false, l2 = false, h3 = false, l3 = false;
	uint32_t z1, m1 = 0u;
	if (!kCigar) z1 = __vmaxs2(z, a);                                      // :153
	else if (!kRight) {
		z1 = __vmaxs2(z, a);                                               // :177
		m1 = __vmaxs2(z1, b);                                              // only for the d = 2 flag below
	} else {
		z1 = __vibmax_s16x2(a, z, &h0, &l0);                               // :201-203  d = z > a ? 0 : 1   (bit = (a >= z))
		(void)__vibmax_s16x2(b, z1, &h1, &l1);                             // :204-205  d = z > b ? d : 2   (bit = (b >= z))
	}
	const uint32_t zz = __vminu2(__vmaxu2(z1, b), sc.maxsc2);              // :41-42
	const uint32_t cz = ~zz;
	U = ~__vadd2(cz, vt1);                                                 // :43  z - v(t-1)
	V = ~__vadd2(cz, ut);                                                  // :44  z - u(t)
	const uint32_t t = __vadd2(cz, sc.qeps2);                              // :45  -(z - q)
	if (!kCigar) {
		X = __viaddmax_s16x2(a, t, sc.zr);                                    // :46,160
		Y = __viaddmax_s16x2(b, t, sc.zr);                                    // :47,161
		return;
	}
	if (!kRight) {
		// Left-aligned arm.  The four code bits come out as packed FLAGS (0x0100 per half when set) of values that are
		// already there -- no predicates (8 live predicates per register pair spill; r01 line profile):
		//   d = a > z            <=> max(z, a) != z                  (:175-177)
		//   d = b > z' ? 2 : d   <=> max_s(z', b) != z'              (:178-179)
		//   x' > 0, y' > 0       <=> relu(..) != 0                   (:187-192)
		// Every operand is a multiple of 0x100 per half, so umin(v, 0x0100) is the "non-zero" flag.
		constexpr uint32_t K = 0x01000100u;
		X = __viaddmax_s16x2(a, t, sc.zr);                                    // :46,187
		Y = __viaddmax_s16x2(b, t, sc.zr);                                    // :47,190
		const uint32_t f0 = __vminu2(z1 ^ z, K), f1 = __vminu2(m1 ^ z1, K), f2 = __vminu2(X, K), f3 = __vminu2(Y, K);
		const uint32_t c = (f3 * 2u + f2) * 4u + (f1 * 2u + f0);           // per half: code nibble at bits 8..11 (FMA pipe)
		const uint32_t byte = (uint32_t)__dp4a((int)c, 0x10000100, 0);     // low-half nibble + 16 * high-half nibble
		cw = byte * (1u << SH) + cw;
		return;
	}
	const uint32_t a2 = __vadd2(a, t), b2 = __vadd2(b, t);                 // :46-47
	{
		X = __vibmax_s16x2(a2, 0u, &h2, &l2);                              // :213-215  bit = !(0 > a) = (a >= 0)
		Y = __vibmax_s16x2(b2, 0u, &h3, &l3);                              // :216-218
		if (l0) cw |= 0x01u << SH;
		if (l1) cw |= 0x02u << SH;
		if (l2) cw |= 0x04u << SH;
		if (l3) cw |= 0x08u << SH;
		if (h0) cw |= 0x10u << SH;
		if (h1) cw |= 0x20u << SH;
		if (h2) cw |= 0x40u << SH;
		if (h3) cw |= 0x80u << SH;
	}
}

// ---- per-lane state ----
struct Lane16 {
	uint32_t U[16], V[16], X[16], Y[16];   // register i: low half = slot t0[0] + i (block A), high half = slot t0[1] + i (block B)
	uint32_t Z[16];                         // s + 2(q+e), persistent / stale-aware like LaneState::Z
	uint32_t TW[8], QW[8];                  // [0..3] block A, [4..7] block B: byte i = 32 * target[t0 + i] / 4 * query[r - (t0 + i)]
	int t0[2];
};

template <int HALF>
__device__ __forceinline__ void lane16_load_seq(Lane16 &ls, const uint8_t *tseq, int tlen, const uint8_t *qseq, int r)
{
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		uint32_t tw = 0, qw = 0;
#pragma unroll
		for (int b = 0; b < 4; ++b) {
			const int t = ls.t0[HALF] + 4 * k + b;
			tw |= (t < tlen ? ld_u8(tseq + t) << 5 : 0u) << (8 * b);
			qw |= qbyte4(qseq, r - 1 - t) << (8 * b);
		}
		ls.TW[HALF * 4 + k] = tw; ls.QW[HALF * 4 + k] = qw;
	}
}
template <int HALF>
__device__ __forceinline__ void lane16_zero_state(Lane16 &ls)
{
	constexpr uint32_t keep = HALF ? 0x0000ffffu : 0xffff0000u;
#pragma unroll
	for (int i = 0; i < 16; ++i) { ls.U[i] &= keep; ls.V[i] &= keep; ls.X[i] &= keep; ls.Y[i] &= keep; }
}
template <int HALF>
__device__ __forceinline__ void lane16_reset_z(Lane16 &ls, uint32_t s0_2)
{
	constexpr uint32_t keep = HALF ? 0x0000ffffu : 0xffff0000u;
#pragma unroll
	for (int i = 0; i < 16; ++i) ls.Z[i] = (ls.Z[i] & keep) | (s0_2 & ~keep);
}

// window slide / band entry, query shift, top-row boundary and score fill for one anti-diagonal
template <int NS>
__device__ __forceinline__ void lane16_prepare(Lane16 &ls, const Band &b, int r, int last_en, const uint8_t *qseq, const uint8_t *tseq,
                                               int tlen, uint32_t table_saddr, const Sc16 &sc)
{
	// block left the rounded band: take the slots NS further up (z back to its calloc'ed value, fresh sequence bytes)
	if (ls.t0[0] + 15 < b.st) { ls.t0[0] += NS; lane16_load_seq<0>(ls, tseq, tlen, qseq, r); lane16_reset_z<0>(ls, sc.s0_2); }
	if (ls.t0[1] + 15 < b.st) { ls.t0[1] += NS; lane16_load_seq<1>(ls, tseq, tlen, qseq, r); lane16_reset_z<1>(ls, sc.s0_2); }
	// block enters the rounded band: u, v, x, y are the reference's calloc'ed zeros (whatever was computed above en is dropped)
	if (ls.t0[0] <= b.en && ls.t0[0] > last_en) lane16_zero_state<0>(ls);
	if (ls.t0[1] <= b.en && ls.t0[1] > last_en) lane16_zero_state<1>(ls);
	// the query slides past the slots by one per anti-diagonal
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const uint32_t nb = qbyte4(qseq, r - ls.t0[h]);
#pragma unroll
		for (int k = 3; k > 0; --k) ls.QW[h * 4 + k] = __funnelshift_l(ls.QW[h * 4 + k - 1], ls.QW[h * 4 + k], 8);
		ls.QW[h * 4] = (ls.QW[h * 4] << 8) | nb;
	}
	// top-row boundary (:122): y[r] = 0, u[r] = r ? q : 0 when the rounded range reaches slot r
	if (b.en >= r) {
		const uint32_t uq = r ? sc.q16 : 0u;
		const int kA = r - ls.t0[0], kB = r - ls.t0[1];
		if ((unsigned)kA < 16u) {
#pragma unroll
			for (int i = 0; i < 16; ++i) if (kA == i) { ls.Y[i] &= 0xffff0000u; ls.U[i] = (ls.U[i] & 0xffff0000u) | uq; }
		}
		if ((unsigned)kB < 16u) {
#pragma unroll
			for (int i = 0; i < 16; ++i) if (kB == i) { ls.Y[i] &= 0x0000ffffu; ls.U[i] = (ls.U[i] & 0x0000ffffu) | (uq << 16); }
		}
	}
	// score fill (:124-141): slots st0..fe get a fresh s, all others keep the stale one
	uint32_t fillmask = 0;
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		int lo = b.st0 - ls.t0[h], hi = b.fe - ls.t0[h] + 1;
		lo = lo < 0 ? 0 : (lo > 16 ? 16 : lo);
		hi = hi < 0 ? 0 : (hi > 16 ? 16 : hi);
		const uint32_t m = hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
		fillmask |= m << (16 * h);
	}
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t offA = ls.TW[k] + ls.QW[k], offB = ls.TW[4 + k] + ls.QW[4 + k];   // byte b: 32*target + 4*query <= 252
#pragma unroll
		for (int bb = 0; bb < 4; ++bb) {
			const int i = 4 * k + bb;
			const uint32_t zA = lds_u32(table_saddr + __byte_perm(offA, 0u, 0x4440 | bb));
			const uint32_t zB = lds_u32(table_saddr + __byte_perm(offB, 0u, 0x4440 | bb));
			if (fillmask & (1u << i)) ls.Z[i] = __byte_perm(ls.Z[i], zA, 0x3254);          // low half <- table entry
			if (fillmask & (1u << (16 + i))) ls.Z[i] = __byte_perm(ls.Z[i], zB, 0x5410);   // high half <- table entry
		}
	}
}

// the cells of one anti-diagonal for this lane, traceback codes, u' dump and lazy-H update (:233-255).
// Hrow / Urow point at this thread's column of the CTA-wide row arrays: row k of the thread is Hrow[k * RS]
// (RS = threads per CTA).
// H rows 2j / 2j+1 hold slots 4j..4j+3 of block A / block B; U row j holds registers 4j..4j+3.
// kApprox (KSW_EZ_APPROX_MAX): no H row at all -- the lane dumps v' next to u' (into the first four H rows, which are unused
// then) for the leader's one tracked score.
template <bool kCigar, bool kRight, int RS = 128, bool kApprox = false>
__device__ __forceinline__ int32_t lane16_cells(Lane16 &ls, const Band &b, int r, int last_st, uint32_t xin, uint32_t vin,
                                                uint4 *tb_dst, int4 *Hrow, uint4 *Urow, const Sc16 &sc)
{
	// carry into the first slot of a block that starts the rounded range (:117-121)
	uint32_t orv = 0u;
	if (ls.t0[0] == b.st || ls.t0[1] == b.st) {
		const uint32_t keep = ls.t0[0] == b.st ? 0xffff0000u : 0x0000ffffu;
		if (b.st > 0) { if (!(b.st > last_st)) { xin &= keep; vin &= keep; } }       // slot st-1 was not computed on the last diagonal
		else { xin &= keep; vin = (vin & keep) | ((r ? sc.q16 : 0u) << (ls.t0[0] == b.st ? 0 : 16)); }
		// x1 / v1 are int8_t and go through _mm_cvtsi32_si128() (:102,144-145): a carry byte >= 0x80 is sign-extended into lanes
		// 1..3, and the _mm_or_si128 of the first block (:30,34) turns x[t-1] / v[t-1] of slots st+1..st+3 into 0xff.  (x is
		// always in [0,127]; v reaches 128+ once 2(q+e) + match exceeds 127 -- never with SEDEF's scoring.)
		// Only v can trigger it: x = and(cmpgt(a, 0), a) (:187-188) is always in [0,127].
		if (vin & ~keep & 0x80008000u) orv = ~keep & 0xff00ff00u;
	}
	uint32_t cw[4] = {0u, 0u, 0u, 0u};
#define EXTZ_CELL2(ii) \
	cell2<kRight, kCigar, 8 * ((ii) & 3)>(ls.Z[ii], (ii) ? ls.X[(ii) ? (ii) - 1 : 0] : xin, \
	                                       (ii) ? (ls.V[(ii) ? (ii) - 1 : 0] | (((ii) >= 1 && (ii) <= 3) ? orv : 0u)) : vin, \
	                                       ls.U[ii], ls.V[ii], ls.X[ii], ls.Y[ii], sc, cw[(ii) >> 2]);
	// descending: register i reads the OLD x, v of register i-1
	EXTZ_CELL2(15) EXTZ_CELL2(14) EXTZ_CELL2(13) EXTZ_CELL2(12) EXTZ_CELL2(11) EXTZ_CELL2(10) EXTZ_CELL2(9) EXTZ_CELL2(8)
	EXTZ_CELL2(7) EXTZ_CELL2(6) EXTZ_CELL2(5) EXTZ_CELL2(4) EXTZ_CELL2(3) EXTZ_CELL2(2) EXTZ_CELL2(1) EXTZ_CELL2(0)
#undef EXTZ_CELL2
	if (kCigar) *tb_dst = make_uint4(cw[0], cw[1], cw[2], cw[3]);
	int32_t lane_max = kNegInf;
	if (kApprox) {
		uint4 *Vrow = reinterpret_cast<uint4 *>(Hrow);
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			Urow[j * RS] = make_uint4(ls.U[4 * j], ls.U[4 * j + 1], ls.U[4 * j + 2], ls.U[4 * j + 3]);
			Vrow[j * RS] = make_uint4(ls.V[4 * j], ls.V[4 * j + 1], ls.V[4 * j + 2], ls.V[4 * j + 3]);
		}
		return lane_max;
	}
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		// H += v8[t] (extern/ksw2_extz2_sse.cc:103,255: v8 is uint8_t*, the byte is ZERO-extended): the UNSIGNED IDP.4A against a
		// one-hot byte vector picks the top byte of a half.  (A signed pick is the same for SEDEF's 5/-4/40/1, where every v byte is
		// below 128, and wrong as soon as 2(q+e) + max score exceeds 127 -- found by the random-scoring soak, profiles/r01_soak.md.)
		Urow[j * RS] = make_uint4(ls.U[4 * j], ls.U[4 * j + 1], ls.U[4 * j + 2], ls.U[4 * j + 3]);
		int4 ha = Hrow[(2 * j) * RS], hb = Hrow[(2 * j + 1) * RS];
		ha.x = (int32_t)__dp4a(ls.V[4 * j], 0x00000100u, (uint32_t)ha.x);     hb.x = (int32_t)__dp4a(ls.V[4 * j], 0x01000000u, (uint32_t)hb.x);
		ha.y = (int32_t)__dp4a(ls.V[4 * j + 1], 0x00000100u, (uint32_t)ha.y); hb.y = (int32_t)__dp4a(ls.V[4 * j + 1], 0x01000000u, (uint32_t)hb.y);
		ha.z = (int32_t)__dp4a(ls.V[4 * j + 2], 0x00000100u, (uint32_t)ha.z); hb.z = (int32_t)__dp4a(ls.V[4 * j + 2], 0x01000000u, (uint32_t)hb.z);
		ha.w = (int32_t)__dp4a(ls.V[4 * j + 3], 0x00000100u, (uint32_t)ha.w); hb.w = (int32_t)__dp4a(ls.V[4 * j + 3], 0x01000000u, (uint32_t)hb.w);
		Hrow[(2 * j) * RS] = ha; Hrow[(2 * j + 1) * RS] = hb;
		int32_t m0 = ha.x > ha.y ? ha.x : ha.y, m1 = ha.z > ha.w ? ha.z : ha.w;
		int32_t m2 = hb.x > hb.y ? hb.x : hb.y, m3 = hb.z > hb.w ? hb.z : hb.w;
		m0 = m0 > m1 ? m0 : m1; m2 = m2 > m3 ? m2 : m3; m0 = m0 > m2 ? m0 : m2;
		lane_max = lane_max > m0 ? lane_max : m0;
	}
	return lane_max;
}

// The fast pass returns (count << 24) + sum of t.  Only "count == 1" matters, but a group can hold up to 1024 equal maxima
// (degenerate scoring: match 1, mismatch -100 makes whole anti-diagonals tie) and 257 of them would wrap to 1.  Every lane
// therefore reports a count of at most 2: the group sum stays below 2 * 32 and "exactly one" stays exact.
__device__ __forceinline__ uint32_t clamp_tie_count(uint32_t acc) { return (acc >> 24) > 1u ? (2u << 24) : acc; }

// arg-max, fast pass: (count << 24) + sum of t over this lane's slots whose lazy H equals gm
template <int RS = 128>
__device__ __forceinline__ uint32_t lane16_argmax_count(const Lane16 &ls, const int4 *Hrow, int32_t gm)
{
	uint32_t acc = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const int4 h = Hrow[k * RS];
		const uint32_t base = (1u << 24) + (uint32_t)(ls.t0[k & 1] + 4 * (k >> 1));
		if (h.x == gm) acc += base;
		if (h.y == gm) acc += base + 1;
		if (h.z == gm) acc += base + 2;
		if (h.w == gm) acc += base + 3;
	}
	return clamp_tie_count(acc);
}
// the same count over the 8 rows of ONE lane (`wl`, group-relative), split between the G lanes of the group
template <int G>
__device__ __forceinline__ uint32_t group_argmax_count(const int4 *Hgroup, int wl, int wt0a, int wt0b, int gl, int32_t gm)
{
	constexpr int RPL = G >= 8 ? 1 : 8 / G;              // rows per lane
	uint32_t acc = 0;
	if (G > 8 && gl >= 8) return 0u;
#pragma unroll
	for (int j = 0; j < RPL; ++j) {
		const int k = gl * RPL + j;
		const int4 h = Hgroup[k * 128 + wl];
		const uint32_t base = (1u << 24) + (uint32_t)(((k & 1) ? wt0b : wt0a) + 4 * (k >> 1));
		if (h.x == gm) acc += base;
		if (h.y == gm) acc += base + 1;
		if (h.z == gm) acc += base + 2;
		if (h.w == gm) acc += base + 3;
	}
	return clamp_tie_count(acc);
}
// arg-max, exact pass (only on real ties): smallest tie-break key among this lane's slots whose lazy H equals gm
// t0a / t0b: slot bases of the lane's two blocks ON THE DIAGONAL BEING SCANNED.  The pipelined kernels call this after
// prepare(r+1) has run, which may already have moved a block NS further up -- and the block that leaves the band at r+1
// can hold the maximum of diagonal r (a maximum on the last query row sits at slot st0), so ls.t0 must not be used there.
template <int RS = 128>
__device__ __forceinline__ uint32_t lane16_argmax_key(const Band &b, const int4 *Hrow, int32_t gm, int t0a, int t0b)
{
	uint32_t key = 0xffffffffu;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const int4 h = Hrow[k * RS];
		const int32_t hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			const int t = ((k & 1) ? t0b : t0a) + 4 * (k >> 1) + e;
			if (t >= b.st0 && t <= b.en0 && !(t == b.en0 && b.en0 > 0) && hv[e] == gm) {
				const uint32_t kk = tie_key(t, b.st0, b.en0);
				key = kk < key ? kk : key;
			}
		}
	}
	return key;
}

// H / u' rows of one pair inside the CTA-wide row arrays, addressed by circular slot index (for the Leader)
template <int G, int RS = 128>
struct PackedRows {
	static constexpr int kMask = G * 32 - 1;
	int32_t *H;         // (int32_t *)&sH[0][first thread of the group]
	uint32_t *Us;       // (uint32_t *)&sU[0][first thread of the group]
	__device__ __forceinline__ int32_t &h(int c) const
	{
		const int vl = c >> 4, i = c & 15;
		return H[(((((i >> 2) << 1) | (vl & 1)) * RS + (vl >> 1)) << 2) | (i & 3)];
	}
	__device__ __forceinline__ uint32_t u(int c) const          // top-byte form (value << 24), like LocalRows::u
	{
		const int vl = c >> 4, i = c & 15;
		const uint32_t w = Us[((((i >> 2) * RS) + (vl >> 1)) << 2) | (i & 3)];
		return (vl & 1) ? (w & 0xffff0000u) : (w << 16);
	}
	__device__ __forceinline__ uint32_t v(int c) const          // v' dump of the approx-max variant: same layout, in the H rows' storage
	{
		const int vl = c >> 4, i = c & 15;
		const uint32_t w = reinterpret_cast<const uint32_t *>(H)[((((i >> 2) * RS) + (vl >> 1)) << 2) | (i & 3)];
		return (vl & 1) ? (w & 0xffff0000u) : (w << 16);
	}
};

// =====================================================================================================
// packed narrow kernel: G <= 32 lanes x 32 slots per pair; the 32/G groups of a warp each run their own pair at their own
// anti-diagonal (independent groups, see below)
// =====================================================================================================
#ifndef EXTZ_MIN_BLOCKS_P
#define EXTZ_MIN_BLOCKS_P 3
#endif
// The SPARE block of the 16-lane class: a 33rd 16-slot block for the anti-diagonals on which the rounded range [st, max(en, fe)]
// spans 33 blocks -- the block ENTERING at the top while the block congruent to it (same lane, same half) at the bottom is still
// live.  Group lane j holds slot t0 + j of it in the LOW halves of five registers (same int8 << 8 form as Lane16) and its lazy H in
// a register; the slot below is the previous lane's (one shuffle each for x, v, H) or, for lane 0, the top slot of the block below
// the spare.  The cell is the same cell2<>.  When the bottom block leaves the band, the lane that owns the position takes the 16
// slots over into its registers and rows (96 shuffles, once per 32 anti-diagonals), and the block lives on as an ordinary one.
// Its traceback codes go to 8 extra bytes per row (row = NS/2 + 16 bytes).
struct Spare16 {
	uint32_t U, V, X, Y, Z;      // low half: the slot's value; high half: unused
	uint32_t TW;                 // 32 * target[t0 + j]
	int32_t H;                   // lazy H of the slot
	int t0;                      // first slot of the block, -1 while there is none
};
template <int G, bool kApprox> struct kSpareClass { static constexpr bool value = (G == 16) && !kApprox; };
template <int G, bool kCigar, bool kRight, bool kApprox = false>
__global__ void __launch_bounds__(128, EXTZ_MIN_BLOCKS_P)
extz_dp16_kernel(DpLaunch L)
{
	constexpr int NS = G * 32;
	constexpr int PPW = 32 / G;                          // pairs per warp
	constexpr unsigned FULL = 0xffffffffu;
	static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0, "G must be a power of two");
	// The 16-lane class holds 33 blocks: its 32 in the lanes' registers plus a SPARE block spread over the 16 lanes, one slot each
	// (see Spare16).  A band of w = 496..511 (BASELINE.json configs[2]: w = 500) needs 33 blocks for a quarter of the anti-diagonals
	// and would otherwise run in the 64-block class with half of the lane work outside the band.
	constexpr bool kSpare = kSpareClass<G, kApprox>::value;
	constexpr int ROWB = (NS >> 1) + (kSpare ? 16 : 0);  // bytes per traceback row

	__shared__ int4 sH[8][128];                          // lazy H: H + (q+e)*r, row k of thread x at sH[k][x]
	__shared__ uint4 sU[4][128];                         // u' of the current diagonal (for H[en0], :228), packed
	__shared__ uint32_t sTable[kTableStride * kTableStride];

	for (int i = threadIdx.x; i < kTableStride * kTableStride; i += blockDim.x) sTable[i] = (L.table[i] >> 16) * 0x00010001u;
	__syncthreads();

	const uint32_t table_saddr = (uint32_t)__cvta_generic_to_shared(sTable);
	const int lane_w = threadIdx.x & 31;
	const int gl = threadIdx.x % G;                                  // lane within group
	const int pred_lane = (gl + G - 1) & (G - 1);                    // circular predecessor (relative to group)
	int4 *Hrow = &sH[0][threadIdx.x];
	uint4 *Urow = &sU[0][threadIdx.x];
	const PackedRows<G> rows{(int32_t *)&sH[0][threadIdx.x - gl], (uint32_t *)&sU[0][threadIdx.x - gl]};
	const Scoring &sc = L.sc;
	const Sc16 sc16 = make_sc16(L.sc);
	const int qe = sc.qe;
	const bool generic = (sc.flag & kFlagGenericSc) != 0;

	// Per-GROUP state: every lane of a group holds the same values.  The 32/G groups of a warp advance INDEPENDENTLY -- each at
	// its own anti-diagonal r of its own pair; a group whose pair is finished (end of the DP, z-drop, band exhausted) hands in the
	// result and takes the next pair from the queue (pairs sorted by descending work) while the others carry on.  (Until round 2
	// the pairs of a warp ran in lock-step from a common start, and a warp was busy until its longest pair had finished: 37 % of
	// BASELINE.json configs[2]'s pairs z-drop somewhere along the way.)
	int pi = -1, qlen = 0, tlen = 0, w = 0, r = 0;        // (R = qlen + tlen - 1, T = tlen rounded up to 16 and the number of processed
	                                                      // anti-diagonals = r are recomputed where they are needed: registers are scarce here)
	const uint8_t *qseq = L.seq, *tseq = L.seq;
	uint8_t *tbp = nullptr;
	bool alive = false, exhausted = false;
	Lane16 ls;
	ls.t0[0] = ls.t0[1] = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) ls.U[i] = ls.V[i] = ls.X[i] = ls.Y[i] = ls.Z[i] = 0u;
#pragma unroll
	for (int i = 0; i < 8; ++i) ls.TW[i] = ls.QW[i] = 0u;
	Leader ld; ld.reset();
	int last_st = -1, last_en = -1;
	Spare16 sp;
	sp.U = sp.V = sp.X = sp.Y = sp.Z = sp.TW = 0u; sp.H = kNegInf; sp.t0 = -1;

	for (;;) {
		{
			const bool done = !exhausted && !(alive && r < qlen + tlen - 1);
			if (__any_sync(FULL, done)) {
				if (done && gl == 0 && pi >= 0) ld.store(&L.results[pi], r);     // r = anti-diagonals processed
				// dynamic work queue: ONE atomic per warp for all of its waiting groups (with one lane per pair that is up to 32 at once)
				const unsigned dmask = __ballot_sync(FULL, done && gl == 0);
				const int first = __ffs(dmask) - 1;
				int base = 0;
				if (lane_w == first) base = atomicAdd(L.work_counter, __popc(dmask));
				base = __shfl_sync(FULL, base, first);
				const int nxt = base + __popc(dmask & ((1u << (lane_w - gl)) - 1u));   // my group's rank among the waiting ones
				if (done) {
					pi = nxt < L.n ? nxt : -1;
					alive = pi >= 0; exhausted = !alive; r = 0;
					if (!alive) { qlen = 0; tlen = 0; }
					if (alive) {
						const PairDesc pd = L.pairs[pi];
						qlen = pd.qlen; tlen = pd.tlen; w = pd.w;
						qseq = L.seq + pd.q_off;                         // qseq[j], j in [0, qlen + 15] is read
						tseq = L.seq + pd.t_off;
						tbp = kCigar ? L.tb + pd.tb_off + gl * 16 : nullptr;
						ls.t0[0] = gl * 32; ls.t0[1] = gl * 32 + 16;
#pragma unroll
						for (int i = 0; i < 16; ++i) { ls.U[i] = ls.V[i] = ls.X[i] = ls.Y[i] = 0u; ls.Z[i] = sc16.s0_2; }   // calloc'ed arrays (:83)
						lane16_load_seq<0>(ls, tseq, tlen, qseq, 0);
						lane16_load_seq<1>(ls, tseq, tlen, qseq, 0);
#pragma unroll
						for (int k = 0; k < 8; ++k) Hrow[k * 128] = make_int4(kNegInf, kNegInf, kNegInf, kNegInf);       // :86-89
						ld.reset();
						last_st = -1; last_en = -1;
						sp.U = sp.V = sp.X = sp.Y = sp.Z = sp.TW = 0u; sp.H = kNegInf; sp.t0 = -1;
					}
				}
				__syncwarp();
				if (!__any_sync(FULL, alive)) break;
			}
		}
		{
			bool act = alive && r < qlen + tlen - 1;
			const int T = (tlen + 15) & ~15;
			Band b;
			const bool okb = band_of(r, qlen, tlen, w, T, generic, b);
			if (act && !okb) { ld.ez.zdropped = 1; alive = false; act = false; }               // :110-113 (nothing reads the flag before the store)

			// carries: OLD x,v of the slot below each block (:28-35)
			uint32_t xin, vin;
			if (G == 1) { xin = __byte_perm(ls.X[15], ls.X[15], 0x5432); vin = __byte_perm(ls.V[15], ls.V[15], 0x5432); }
			else {
				xin = __byte_perm(__shfl_sync(FULL, ls.X[15], pred_lane, G), ls.X[15], 0x5432);
				vin = __byte_perm(__shfl_sync(FULL, ls.V[15], pred_lane, G), ls.V[15], 0x5432);
			}
			if (kApprox) {
				// KSW_EZ_APPROX_MAX: no H row, no maximum, no arg-max -- the leader walks its one score over the u' / v' dumps
				if (act) {
					lane16_prepare<NS>(ls, b, r, last_en, qseq, tseq, tlen, table_saddr, sc16);
					lane16_cells<kCigar, kRight, 128, true>(ls, b, r, last_st, xin, vin, (uint4 *)(tbp + (int64_t)r * (NS >> 1)), Hrow, Urow, sc16);
				}
				__syncwarp();
				int stop = 0;
				if (act && gl == 0) stop = ld.approx(rows, b, r, qe, ls.V[0] << 16, qlen, tlen, sc.zdrop, sc.e, (sc.flag & kFlagApproxDrop) != 0);
				stop = __shfl_sync(FULL, stop, 0, G);
				if (act) { last_st = b.st; last_en = b.en; if (stop) alive = false; ++r; }
				__syncwarp();            // the leader read the dumps; the next diagonal overwrites them
				continue;
			}
			// ---- exact maximum, LEADERLESS ----------------------------------------------------------------------------------
			// The scalar bookkeeping of a pair (ez, the special H entries) is not done by one leader lane reaching into the
			// other lanes' shared-memory rows between __syncwarp()s: every lane of the group keeps an IDENTICAL copy of the scalar
			// state (same issue slots under SIMT), and the three H entries that need dynamic slot addressing -- H[en0], the old
			// H[en0-1] it is computed from, the exited H[st0-1] -- are handled by the lane that OWNS the slot, on its own rows
			// (program order, no barrier).  What crosses lanes travels in registers: the diagonal maximum (xor-reduction), H[en0]
			// (one shuffle from its owner) and the old top H of the predecessor lane (one shuffle).  The only cross-lane
			// shared-memory reads left are the arg-max passes.
			const int en_c = b.en0 & (NS - 1);
			const int en_owner = en_c >> 5;                                 // group-relative lane that owns slot en0
			// old H of my block-B top slot as the reference reads it on this diagonal: the lazy row entry, or -- if that slot has
			// left the band -- its TRUE H frozen at exit (stale read of :228), moved into the lazy domain of diagonal r-1
			int32_t htop;
			{
				// (the block may already have moved NS further up -- it slides on the diagonal its top slot exits -- hence both tests)
				const int ttop = ls.t0[1] + 15;
				htop = (ttop == ld.exit_slot || ttop - NS == ld.exit_slot) ? ld.exit_H + qe * (r - 1)
				                                                            : reinterpret_cast<const int32_t *>(Hrow)[((7 * 128) << 2) | 3];
			}
			const int32_t hcar = G == 1 ? htop : __shfl_sync(FULL, htop, pred_lane, G);
			int32_t Hen0 = kNegInf;
			int32_t lane_max = kNegInf;                                     // maximum over this lane's H ROWS (the spare's H is a register)
			auto hptr = [&](int t) -> int32_t * {                           // lazy-H entry of a slot THIS lane owns
				const int i = t & 15, half = (t >> 4) & 1;
				return reinterpret_cast<int32_t *>(Hrow) + ((((((i >> 2) << 1) | half) * 128) << 2) | (i & 3));
			};
			auto owns = [&](int t) { return G == 1 || ((t & (NS - 1)) >> 5) == gl; };
			// slot st0-1 left the band: freeze its TRUE H, drop it from the max.  (First of all: with a spare block the rows of the
			// leaving block are overwritten below.)
			if (act && r > 0 && b.st0 > ld.st0_prev) {
				const int xs = b.st0 - 1;
				ld.exit_slot = xs;
				if (owns(xs)) { int32_t *px = hptr(xs); ld.exit_H = *px - qe * (r - 1); *px = kNegInf; }
			}
			// ---- the spare block, part 1: its position in the ring has become free -> the owning lane takes it over ----------
			bool sp_on = false, sp_in = false;                              // spare in use on this diagonal / en0 lies in it
			uint32_t sxin = 0u, svin = 0u, scw = 0u; int32_t shprev = kNegInf;
			if (kSpare) {
				const bool adopt = act && sp.t0 >= 0 && sp.t0 - NS + 15 < b.st;
				if (__any_sync(FULL, adopt)) {
					const int La = (sp.t0 & (NS - 1)) >> 5, ha = (sp.t0 >> 4) & 1;
					const bool me = adopt && gl == La;
					int32_t *hr = reinterpret_cast<int32_t *>(Hrow);
#pragma unroll
					for (int i = 0; i < 16; ++i) {
						const uint32_t u = __shfl_sync(FULL, sp.U, i, 16), v = __shfl_sync(FULL, sp.V, i, 16);
						const uint32_t x = __shfl_sync(FULL, sp.X, i, 16), y = __shfl_sync(FULL, sp.Y, i, 16);
						const uint32_t z = __shfl_sync(FULL, sp.Z, i, 16);
						const int32_t h = __shfl_sync(FULL, sp.H, i, 16);
						if (me) {
							if (ha) {
								ls.U[i] = __byte_perm(ls.U[i], u, 0x5410); ls.V[i] = __byte_perm(ls.V[i], v, 0x5410);
								ls.X[i] = __byte_perm(ls.X[i], x, 0x5410); ls.Y[i] = __byte_perm(ls.Y[i], y, 0x5410);
								ls.Z[i] = __byte_perm(ls.Z[i], z, 0x5410);
							} else {
								ls.U[i] = __byte_perm(ls.U[i], u, 0x3254); ls.V[i] = __byte_perm(ls.V[i], v, 0x3254);
								ls.X[i] = __byte_perm(ls.X[i], x, 0x3254); ls.Y[i] = __byte_perm(ls.Y[i], y, 0x3254);
								ls.Z[i] = __byte_perm(ls.Z[i], z, 0x3254);
							}
							hr[((((((i >> 2) << 1) | ha) * 128) << 2) | (i & 3))] = h;
						}
					}
					if (me) {
						if (ha) { ls.t0[1] = sp.t0; lane16_load_seq<1>(ls, tseq, tlen, qseq, r); }
						else { ls.t0[0] = sp.t0; lane16_load_seq<0>(ls, tseq, tlen, qseq, r); }
					}
					if (adopt) sp.t0 = -1;
				}
			}
			if (act) lane16_prepare<NS>(ls, b, r, last_en, qseq, tseq, tlen, table_saddr, sc16);
			// ---- the spare block, part 2: activation, band entry, carries from its neighbours, top row, score fill -----------
			if (kSpare) {
				const int top = b.en > b.fe ? b.en : b.fe;                  // the score fill can run ahead of the rounded range (:125)
				if (act && sp.t0 < 0 && top >= b.st + NS) {
					sp.t0 = b.st + NS;
					sp.U = sp.V = sp.X = sp.Y = 0u; sp.Z = sc16.s0_2; sp.H = kNegInf;          // calloc'ed state (:83-89)
					const int t = sp.t0 + gl;
					sp.TW = t < tlen ? ld_u8(tseq + t) << 5 : 0u;
				}
				sp_on = act && sp.t0 >= 0;
				if (__any_sync(FULL, sp_on)) {
					if (sp_on && sp.t0 <= b.en && sp.t0 > last_en) sp.U = sp.V = sp.X = sp.Y = 0u;   // enters the rounded band
					// OLD x, v, H of slot t-1: the previous spare lane, or (lane 0) the top slot of the block below the spare
					const uint32_t xn = __shfl_up_sync(FULL, sp.X, 1, 16), vn = __shfl_up_sync(FULL, sp.V, 1, 16);
					const int32_t hn = __shfl_up_sync(FULL, sp.H, 1, 16);
					const int tp = sp.t0 - 16;
					const int Lp = (tp & (NS - 1)) >> 5, hp = (tp >> 4) & 1;
					const uint32_t cx = __shfl_sync(FULL, ls.X[15], Lp, 16), cv = __shfl_sync(FULL, ls.V[15], Lp, 16);
					const int32_t h15 = reinterpret_cast<const int32_t *>(Hrow)[(((hp ? 7 : 6) * 128) << 2) | 3];
					const int32_t ch = __shfl_sync(FULL, h15, Lp, 16);
					sxin = gl ? xn : (hp ? cx >> 16 : cx & 0xffffu);
					svin = gl ? vn : (hp ? cv >> 16 : cv & 0xffffu);
					shprev = gl ? hn : ch;
					const int t = sp.t0 + gl;
					if (sp_on && b.en >= r && t == r) { sp.Y = 0u; sp.U = r ? sc16.q16 : 0u; }   // top-row boundary (:122)
					if (sp_on && t >= b.st0 && t <= b.fe) sp.Z = lds_u32(table_saddr + sp.TW + qbyte4(qseq, r - t));   // score fill (:124-141)
				}
				sp_in = sp_on && b.en0 >= sp.t0;
			}
			int32_t tot_max = kNegInf;
			if (act) {
				// (written branch-free where it can be: the owner-only steps are predicated loads / stores on the lane's own rows)
				const bool own_en = (kSpare && sp_in) ? (gl == (b.en0 & 15)) : (en_owner == gl);
				const int ps = b.en0 - 1;
				int32_t hprev = hcar;                                       // en0 starts my block A: en0-1 is the predecessor lane's top slot
				if (kSpare && sp_in) hprev = shprev;
				else if (r > 0 && own_en && b.en0 > 0) {                    // H[en0] is recomputed from the OLD H[en0-1] (:228)
					if ((en_c & 31) != 0) hprev = (ps == ld.exit_slot) ? ld.exit_H + qe * (r - 1) : *hptr(ps);
					*hptr(b.en0) = kNegInf;                                 // the regular update below must not count for slot en0
				}
				lane_max = lane16_cells<kCigar, kRight>(ls, b, r, last_st, xin, vin, (uint4 *)(tbp + (int64_t)r * ROWB), Hrow, Urow, sc16);
				if (kSpare && sp_on) {
					cell2<kRight, kCigar, 0>(sp.Z, sxin, svin, sp.U, sp.V, sp.X, sp.Y, sc16, scw);
					sp.H = (int32_t)__dp4a(sp.V, 0x00000100u, (uint32_t)sp.H);
				}
				{
					const int i = b.en0 & 15;
					const uint32_t uw = reinterpret_cast<const uint32_t *>(Urow)[(((i >> 2) * 128) << 2) | (i & 3)];       // my own u' dump
					int32_t ub = (int32_t)(((b.en0 >> 4) & 1) ? (uw >> 24) : ((uw >> 8) & 0xffu));
					if (kSpare && sp_in) ub = (int32_t)((sp.U >> 8) & 0xffu);
					int32_t hen = hprev + ub;                                                              // :228 in the lazy domain
					if (r == 0) hen = (int32_t)((ls.V[0] >> 8) & 0xffu) - 2 * qe;                          // :259
					else if (b.en0 == 0) hen = *hptr(0);                    // en0 == 0: the regular update (:228 else-arm)
					if (own_en) {
						Hen0 = hen;
						if (kSpare && sp_in) sp.H = hen;
						else { *hptr(b.en0) = hen; lane_max = lane_max > hen ? lane_max : hen; }
					}
				}
				tot_max = lane_max;
				if (kSpare && sp_on) tot_max = tot_max > sp.H ? tot_max : sp.H;
			}
			if (kSpare && kCigar) {
				// the 16 codes of the spare block: 8 bytes behind the packed part of the row (slot 2k low nibble, 2k+1 high nibble)
				if (__any_sync(FULL, sp_on)) {
					const uint32_t nib = scw & 0xfu;
					const uint32_t o = __shfl_down_sync(FULL, nib, 1, 16);
					if (sp_on && !(gl & 1)) (tbp - gl * 16 + (NS >> 1))[(int64_t)r * ROWB + (gl >> 1)] = (uint8_t)(nib | (o << 4));   // behind the packed part of the row
				}
			}
			ld.gmax = group_max<G>(tot_max);
			ld.Hen0_lazy = G == 1 ? Hen0 : __shfl_sync(FULL, Hen0, (kSpare && sp_in) ? (b.en0 & 15) : en_owner, G);
			int need = 0;
			if (act) {
				const int32_t maxH_true = ld.gmax - qe * r;
				need = (maxH_true > ld.ez.max) || (sc.zdrop >= 0 && ld.ez.max - maxH_true > sc.zdrop);
			}
			const int32_t gm = ld.gmax;
			int max_t = b.en0;
			if (__any_sync(FULL, need)) {
				__syncwarp();                                               // the passes below read other lanes' rows
				uint32_t cnt = 0;
				if (G == 1) { if (need) cnt = lane16_argmax_count(ls, Hrow, gm); }
				else {
					// Only a lane whose own maximum reaches gm can hold the arg-max.  With ONE such lane (the rule) the G lanes
					// of the group split ITS 8 rows between them instead of every lane scanning its own 32 entries.
					const bool cand = need && lane_max == gm;
					const unsigned bal = __ballot_sync(FULL, cand);
					const unsigned gmask = G == 32 ? bal : ((bal >> (lane_w & ~(G - 1))) & ((1u << (G & 31)) - 1u));
					const int wl = gmask ? __ffs(gmask) - 1 : 0;
					const int wt0a = __shfl_sync(FULL, ls.t0[0], wl, G), wt0b = __shfl_sync(FULL, ls.t0[1], wl, G);
					if (need) {
						if (__popc(gmask) == 1) cnt = group_argmax_count<G>(Hrow - gl, wl, wt0a, wt0b, gl, gm);
						else if (gmask) cnt = lane16_argmax_count(ls, Hrow, gm);
					}
				}
				if (kSpare && need && sp_on && sp.H == gm) cnt += (1u << 24) + (uint32_t)(sp.t0 + gl);
				cnt = group_sum_u<G>(cnt);
				max_t = (int)(cnt & 0x00ffffffu);
				const int tie = need && (cnt >> 24) != 1u;
				if (__any_sync(FULL, tie)) {                                                      // real ties: exact 4-lane rule
					uint32_t key = 0xffffffffu;
					if (tie) {
						key = lane16_argmax_key(b, Hrow, gm, ls.t0[0], ls.t0[1]);
						if (kSpare && sp_on && sp.H == gm) {
							const int t = sp.t0 + gl;
							if (t >= b.st0 && t <= b.en0 && !(t == b.en0 && b.en0 > 0)) { const uint32_t kk = tie_key(t, b.st0, b.en0); key = kk < key ? kk : key; }
						}
						if (gl == 0) { uint32_t k0 = ld.en0_key(b, r); key = k0 < key ? k0 : key; }
					}
					key = group_min_u<G>(key);
					if (tie) max_t = tie_key_slot(key, b.en0);
				}
				__syncwarp();                                               // ... before the owners overwrite them on the next diagonal
			}
			// H[st0] when the diagonal ends on the last query row (:263-264): from its owner
			const bool want_q = act && (r - b.st0 == qlen - 1) && b.st0 != b.en0;
			int32_t Hst0_lazy = ld.Hen0_lazy;
			if (__any_sync(FULL, want_q)) {
				int32_t hs = 0;
				if (want_q && (G == 1 || ((b.st0 & (NS - 1)) >> 5) == gl)) {
					const int i = b.st0 & 15, half = (b.st0 >> 4) & 1;
					hs = reinterpret_cast<const int32_t *>(Hrow)[(((((i >> 2) << 1) | half) * 128) << 2) | (i & 3)];
				}
				hs = G == 1 ? hs : __shfl_sync(FULL, hs, (b.st0 & (NS - 1)) >> 5, G);
				if (want_q) Hst0_lazy = hs;
			}
			if (act) {
				const int stop = ld.fin_local(b, r, qe, max_t, Hst0_lazy, qlen, tlen, sc.zdrop, sc.e);
				last_st = b.st; last_en = b.en;
				if (stop) alive = false;
				++r;
			}
		}
	}
}

// =====================================================================================================
// packed wide kernel: one CTA of G = 64 / 128 / 256 lanes x 32 slots per pair (2048 / 4096 / 8192 live slots: unbanded gap
// fills of a few kbp).  Carries between warps and the per-diagonal reductions go through shared memory, ordering by __syncthreads;
// otherwise the per-lane code of the narrow kernel.
// =====================================================================================================
template <int G, bool kCigar, bool kRight, bool kApprox = false>
__global__ void __launch_bounds__(G, G == 128 ? 3 : 1)       // 128 lanes: 3 CTAs/SM at <= 170 registers; 64: 5 fit anyway; 256: 1
extz_dp16_wide_kernel(DpLaunch L)
{
	constexpr int NS = G * 32;
	constexpr int NW = G / 32;
	static_assert(G > 32 && G % 32 == 0 && (G & (G - 1)) == 0, "wide kernel: whole warps, power of two");

	// the H / u' rows are dynamic shared memory: 192 B per lane, 48 KB for G = 256 (beyond the static limit with the rest)
	extern __shared__ __align__(16) unsigned char extz_dyn_smem[];
	int4 (*sH)[G] = reinterpret_cast<int4 (*)[G]>(extz_dyn_smem);
	uint4 (*sU)[G] = reinterpret_cast<uint4 (*)[G]>(extz_dyn_smem + sizeof(int4) * 8 * G);
	__shared__ uint32_t sTable[kTableStride * kTableStride];
	__shared__ uint32_t sCarryX[NW], sCarryV[NW];       // OLD x,v (packed) of every warp's top register
	__shared__ int32_t sWarpMax[NW];
	__shared__ uint32_t sWarpKey[NW];
	__shared__ int sPair, sNeedArg, sStop;
	__shared__ int32_t sGmax;

	for (int i = threadIdx.x; i < kTableStride * kTableStride; i += blockDim.x) sTable[i] = (L.table[i] >> 16) * 0x00010001u;
	__syncthreads();

	const uint32_t table_saddr = (uint32_t)__cvta_generic_to_shared(sTable);
	const int gl = threadIdx.x;
	const int lane = gl & 31, wid = gl >> 5;
	int4 *Hrow = &sH[0][gl];
	uint4 *Urow = &sU[0][gl];
	const PackedRows<G, G> rows{(int32_t *)&sH[0][0], (uint32_t *)&sU[0][0]};
	const Scoring &sc = L.sc;
	const Sc16 sc16 = make_sc16(L.sc);
	const int qe = sc.qe;
	const bool generic = (sc.flag & kFlagGenericSc) != 0;

	for (;;) {
		if (gl == 0) sPair = atomicAdd(L.work_counter, 1);
		__syncthreads();
		const int pi = sPair;
		if (pi >= L.n) break;
		const PairDesc pd = L.pairs[pi];
		const int qlen = pd.qlen, tlen = pd.tlen, w = pd.w;
		const int T = (tlen + 15) & ~15;
		const uint8_t *qseq = L.seq + pd.q_off;
		const uint8_t *tseq = L.seq + pd.t_off;
		uint8_t *tbp = kCigar ? L.tb + pd.tb_off + gl * 16 : nullptr;

		Lane16 ls;
		ls.t0[0] = gl * 32; ls.t0[1] = gl * 32 + 16;
#pragma unroll
		for (int i = 0; i < 16; ++i) { ls.U[i] = ls.V[i] = ls.X[i] = ls.Y[i] = 0u; ls.Z[i] = sc16.s0_2; }
		lane16_load_seq<0>(ls, tseq, tlen, qseq, 0);
		lane16_load_seq<1>(ls, tseq, tlen, qseq, 0);
#pragma unroll
		for (int k = 0; k < 8; ++k) Hrow[k * G] = make_int4(kNegInf, kNegInf, kNegInf, kNegInf);
		Leader ld; ld.reset();
		int last_st = -1, last_en = -1, n_diag = 0, zdropped_band = 0;
		const int R = qlen + tlen - 1;
		__syncthreads();

		// Software pipeline over anti-diagonals: while the leader does the bookkeeping of diagonal r (and knocks out the H
		// entries of r+1), every other lane already runs prepare(r+1) -- window slide, query shift, score fill -- which
		// touches neither H nor the leader's state.  Two barriers per diagonal (plus two when the arg-max is needed).
		Band b;
		bool okb = band_of(0, qlen, tlen, w, T, generic, b);
		if (okb) {
			lane16_prepare<NS>(ls, b, 0, -1, qseq, tseq, tlen, table_saddr, sc16);
			if (!kApprox && gl == 0) ld.pre(rows, b, 0, qe);
		} else zdropped_band = 1;
		uint32_t xp = 0u, vp = 0u;                                    // OLD x,v of the register below (zero state at r = 0)
		__syncthreads();
		for (int r = 0; okb && r < R; ++r) {
			// cells of diagonal r
			const uint32_t xin = __byte_perm(xp, ls.X[15], 0x5432), vin = __byte_perm(vp, ls.V[15], 0x5432);
			const int32_t lane_max = lane16_cells<kCigar, kRight, G, kApprox>(ls, b, r, last_st, xin, vin, (uint4 *)(tbp + (int64_t)r * (NS >> 1)),
			                                                                  Hrow, Urow, sc16);
			if (kApprox) {
				// KSW_EZ_APPROX_MAX: the leader walks its one score over the u' / v' dumps of diagonal r while the other lanes
				// run prepare(r+1); two barriers per diagonal
				if (lane == 31) { sCarryX[wid] = ls.X[15]; sCarryV[wid] = ls.V[15]; }
				xp = __shfl_up_sync(0xffffffffu, ls.X[15], 1);
				vp = __shfl_up_sync(0xffffffffu, ls.V[15], 1);
				Band bn;
				const bool okn = (r + 1 < R) && band_of(r + 1, qlen, tlen, w, T, generic, bn);
				__syncthreads();
				if (lane == 0) { const int pw = (wid + NW - 1) % NW; xp = sCarryX[pw]; vp = sCarryV[pw]; }
				if (gl == 0) sStop = ld.approx(rows, b, r, qe, ls.V[0] << 16, qlen, tlen, sc.zdrop, sc.e, (sc.flag & kFlagApproxDrop) != 0);
				if (okn) lane16_prepare<NS>(ls, bn, r + 1, b.en, qseq, tseq, tlen, table_saddr, sc16);
				__syncthreads();
				n_diag = r + 1;
				last_st = b.st; last_en = b.en;
				if (sStop) break;
				if (r + 1 < R && !okn) { zdropped_band = 1; break; }
				b = bn;
				continue;
			}
			const int32_t wmax = __reduce_max_sync(0xffffffffu, lane_max);
			if (lane == 0) sWarpMax[wid] = wmax;
			if (lane == 31) { sCarryX[wid] = ls.X[15]; sCarryV[wid] = ls.V[15]; }      // OLD values for diagonal r+1
			xp = __shfl_up_sync(0xffffffffu, ls.X[15], 1);
			vp = __shfl_up_sync(0xffffffffu, ls.V[15], 1);
			Band bn;
			const bool okn = (r + 1 < R) && band_of(r + 1, qlen, tlen, w, T, generic, bn);
			__syncthreads();                                                                   // 1
			if (lane == 0) { const int pw = (wid + NW - 1) % NW; xp = sCarryX[pw]; vp = sCarryV[pw]; }
			if (gl == 0) {
				int32_t red = sWarpMax[0];
#pragma unroll
				for (int k = 1; k < NW; ++k) red = red > sWarpMax[k] ? red : sWarpMax[k];
				const int need = ld.mid(rows, b, r, qe, red, ls.V[0] << 16, sc.zdrop);
				sNeedArg = need; sGmax = ld.gmax;
				if (!need) {
					const int stop = ld.fin(rows, b, r, qe, b.en0, qlen, tlen, sc.zdrop, sc.e);
					sStop = stop;
					if (!stop && okn) ld.pre(rows, bn, r + 1, qe);
				}
			}
			const int t0a = ls.t0[0], t0b = ls.t0[1];                     // slot bases of diagonal r, for its arg-max pass
			if (okn) lane16_prepare<NS>(ls, bn, r + 1, b.en, qseq, tseq, tlen, table_saddr, sc16);
			__syncthreads();                                                                   // 2
			if (sNeedArg) {
				uint32_t key = lane16_argmax_key<G>(b, Hrow, sGmax, t0a, t0b);
				key = __reduce_min_sync(0xffffffffu, key);
				if (lane == 0) sWarpKey[wid] = key;
				__syncthreads();                                                               // 3
				if (gl == 0) {
					uint32_t k = ld.en0_key(b, r);
#pragma unroll
					for (int j = 0; j < NW; ++j) k = sWarpKey[j] < k ? sWarpKey[j] : k;
					const int stop = ld.fin(rows, b, r, qe, tie_key_slot(k, b.en0), qlen, tlen, sc.zdrop, sc.e);
					sStop = stop;
					if (!stop && okn) ld.pre(rows, bn, r + 1, qe);
				}
				__syncthreads();                                                               // 4
			}
			const int stop = sStop;
			n_diag = r + 1;
			last_st = b.st; last_en = b.en;
			if (stop) break;
			if (r + 1 < R && !okn) { zdropped_band = 1; break; }                               // band exhausted (:110-113)
			b = bn;
		}
		if (gl == 0) { if (zdropped_band) ld.ez.zdropped = 1; ld.store(&L.results[pi], n_diag); }
		__syncthreads();
	}
}

// =====================================================================================================
// packed cluster kernel: one thread-block CLUSTER of C CTAs x 256 lanes x 32 slots per pair (C = 2: 16384 live slots, the
// unbanded <= 10 kbp gap fills of src/align.cc:130-139; C = 4 / 8: 32768 / 65536 live slots, which covers the largest call
// the reference can make -- 60 000 x 60 000 per align_helper chunk, src/align.cc:46-53).  Every CTA keeps the H / u' rows of its own lanes; the carry
// between CTAs, the leader's accesses to arbitrary slots and the per-diagonal reductions go through DISTRIBUTED SHARED
// MEMORY (cluster.map_shared_rank), ordering by cluster.sync() -- the choreography of extz_dp_cluster_kernel with the
// packed lane code.
// =====================================================================================================
template <int C>
struct ClusterRows16 {
	static constexpr int GC = 256, kMask = C * GC * 32 - 1;
	int32_t *H; uint32_t *Us;                           // this CTA's arrays (same offset in every CTA)
	__device__ __forceinline__ int32_t &h(int c) const
	{
		cg::cluster_group cl = cg::this_cluster();
		const int gl = c >> 5, half = (c >> 4) & 1, i = c & 15;
		return cl.map_shared_rank(H, gl / GC)[(((((i >> 2) << 1) | half) * GC + (gl % GC)) << 2) | (i & 3)];
	}
	__device__ __forceinline__ uint32_t u(int c) const
	{
		cg::cluster_group cl = cg::this_cluster();
		const int gl = c >> 5, half = (c >> 4) & 1, i = c & 15;
		const uint32_t w = cl.map_shared_rank(Us, gl / GC)[((((i >> 2) * GC) + (gl % GC)) << 2) | (i & 3)];
		return half ? (w & 0xffff0000u) : (w << 16);
	}
	__device__ __forceinline__ uint32_t v(int c) const
	{
		cg::cluster_group cl = cg::this_cluster();
		const int gl = c >> 5, half = (c >> 4) & 1, i = c & 15;
		const uint32_t w = cl.map_shared_rank(reinterpret_cast<uint32_t *>(H), gl / GC)[((((i >> 2) * GC) + (gl % GC)) << 2) | (i & 3)];
		return half ? (w & 0xffff0000u) : (w << 16);
	}
};

template <int C, bool kCigar, bool kRight, bool kApprox = false>
__global__ void __launch_bounds__(256, 1)
extz_dp16_cluster_kernel(DpLaunch L)
{
	constexpr int GC = 256, G = GC * C, NS = G * 32, NW = GC / 32;
	static_assert(C == 2 || C == 4 || C == 8, "packed cluster kernel: 2, 4 or 8 CTAs (8 = the portable cluster limit)");
	cg::cluster_group cluster = cg::this_cluster();
	const int rank = (int)cluster.block_rank();

	extern __shared__ __align__(16) unsigned char extz_dyn_smem[];
	int4 (*sH)[GC] = reinterpret_cast<int4 (*)[GC]>(extz_dyn_smem);
	uint4 (*sU)[GC] = reinterpret_cast<uint4 (*)[GC]>(extz_dyn_smem + sizeof(int4) * 8 * GC);
	__shared__ uint32_t sTable[kTableStride * kTableStride];
	__shared__ uint32_t sCarryX[NW], sCarryV[NW];       // OLD x,v (packed) of every warp's top register
	__shared__ int32_t sAllMax[8 * NW];                 // [rank][warp], only CTA 0's copy is used
	__shared__ uint32_t sAllKey[8 * NW];
	__shared__ int sPair, sNeedArg, sStop;              // written into EVERY CTA's copy by the leader
	__shared__ int32_t sGmax;

	for (int i = threadIdx.x; i < kTableStride * kTableStride; i += blockDim.x) sTable[i] = (L.table[i] >> 16) * 0x00010001u;
	// cluster.sync(), not __syncthreads(): the first thing the leader does is write sPair into the OTHER CTAs' shared memory,
	// and a peer's distributed shared memory may only be touched once that CTA is known to be running (racecheck:
	// "write to a block that might not have entered yet").  It also orders the table fill inside each CTA.
	cluster.sync();

	const uint32_t table_saddr = (uint32_t)__cvta_generic_to_shared(sTable);
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int gl = rank * GC + tid;                      // lane within the pair
	const bool leader = gl == 0;
	int4 *Hrow = &sH[0][tid];
	uint4 *Urow = &sU[0][tid];
	const ClusterRows16<C> rows{(int32_t *)&sH[0][0], (uint32_t *)&sU[0][0]};
	const Scoring &sc = L.sc;
	const Sc16 sc16 = make_sc16(L.sc);
	const int qe = sc.qe;
	const bool generic = (sc.flag & kFlagGenericSc) != 0;
	int32_t *max0 = cluster.map_shared_rank(sAllMax, 0);
	uint32_t *key0 = cluster.map_shared_rank(sAllKey, 0);

	for (;;) {
		if (leader) {
			const int p = atomicAdd(L.work_counter, 1);
			for (int k = 0; k < C; ++k) *cluster.map_shared_rank(&sPair, k) = p;
		}
		cluster.sync();
		const int pi = sPair;
		if (pi >= L.n) break;
		const PairDesc pd = L.pairs[pi];
		const int qlen = pd.qlen, tlen = pd.tlen, w = pd.w;
		const int T = (tlen + 15) & ~15;
		const uint8_t *qseq = L.seq + pd.q_off;
		const uint8_t *tseq = L.seq + pd.t_off;
		uint8_t *tbp = kCigar ? L.tb + pd.tb_off + gl * 16 : nullptr;

		Lane16 ls;
		ls.t0[0] = gl * 32; ls.t0[1] = gl * 32 + 16;
#pragma unroll
		for (int i = 0; i < 16; ++i) { ls.U[i] = ls.V[i] = ls.X[i] = ls.Y[i] = 0u; ls.Z[i] = sc16.s0_2; }
		lane16_load_seq<0>(ls, tseq, tlen, qseq, 0);
		lane16_load_seq<1>(ls, tseq, tlen, qseq, 0);
#pragma unroll
		for (int k = 0; k < 8; ++k) Hrow[k * GC] = make_int4(kNegInf, kNegInf, kNegInf, kNegInf);
		Leader ld; ld.reset();
		int last_st = -1, last_en = -1, n_diag = 0, zdropped_band = 0;
		const int R = qlen + tlen - 1;
		cluster.sync();

		// software pipeline over anti-diagonals, as in extz_dp16_wide_kernel: prepare(r+1) of every lane overlaps the leader's
		// bookkeeping of diagonal r; two cluster barriers per diagonal (plus two when the arg-max is needed)
		Band b;
		bool okb = band_of(0, qlen, tlen, w, T, generic, b);
		if (okb) {
			lane16_prepare<NS>(ls, b, 0, -1, qseq, tseq, tlen, table_saddr, sc16);
			if (!kApprox && leader) ld.pre(rows, b, 0, qe);
		} else zdropped_band = 1;
		uint32_t xp = 0u, vp = 0u;
		cluster.sync();
		for (int r = 0; okb && r < R; ++r) {
			const uint32_t xin = __byte_perm(xp, ls.X[15], 0x5432), vin = __byte_perm(vp, ls.V[15], 0x5432);
			const int32_t lane_max = lane16_cells<kCigar, kRight, GC, kApprox>(ls, b, r, last_st, xin, vin, (uint4 *)(tbp + (int64_t)r * (NS >> 1)),
			                                                                   Hrow, Urow, sc16);
			if (kApprox) {
				// KSW_EZ_APPROX_MAX, as in extz_dp16_wide_kernel; the leader reads the dumps of other CTAs through DSMEM
				if (lane == 31) { sCarryX[wid] = ls.X[15]; sCarryV[wid] = ls.V[15]; }
				xp = __shfl_up_sync(0xffffffffu, ls.X[15], 1);
				vp = __shfl_up_sync(0xffffffffu, ls.V[15], 1);
				Band bn;
				const bool okn = (r + 1 < R) && band_of(r + 1, qlen, tlen, w, T, generic, bn);
				cluster.sync();
				if (lane == 0) {
					if (wid > 0) { xp = sCarryX[wid - 1]; vp = sCarryV[wid - 1]; }
					else {
						const int pr = (rank + C - 1) % C;
						xp = cluster.map_shared_rank(sCarryX, pr)[NW - 1];
						vp = cluster.map_shared_rank(sCarryV, pr)[NW - 1];
					}
				}
				if (leader) {
					const int stop = ld.approx(rows, b, r, qe, ls.V[0] << 16, qlen, tlen, sc.zdrop, sc.e, (sc.flag & kFlagApproxDrop) != 0);
					for (int k = 0; k < C; ++k) *cluster.map_shared_rank(&sStop, k) = stop;
				}
				if (okn) lane16_prepare<NS>(ls, bn, r + 1, b.en, qseq, tseq, tlen, table_saddr, sc16);
				cluster.sync();
				n_diag = r + 1;
				last_st = b.st; last_en = b.en;
				if (sStop) break;
				if (r + 1 < R && !okn) { zdropped_band = 1; break; }
				b = bn;
				continue;
			}
			const int32_t wmax = __reduce_max_sync(0xffffffffu, lane_max);
			if (lane == 0) max0[rank * NW + wid] = wmax;
			if (lane == 31) { sCarryX[wid] = ls.X[15]; sCarryV[wid] = ls.V[15]; }      // OLD values for diagonal r+1
			xp = __shfl_up_sync(0xffffffffu, ls.X[15], 1);
			vp = __shfl_up_sync(0xffffffffu, ls.V[15], 1);
			Band bn;
			const bool okn = (r + 1 < R) && band_of(r + 1, qlen, tlen, w, T, generic, bn);
			cluster.sync();                                                                    // 1
			// warp 0 of a CTA takes its carry from the last warp of the previous CTA (DSMEM)
			if (lane == 0) {
				if (wid > 0) { xp = sCarryX[wid - 1]; vp = sCarryV[wid - 1]; }
				else {
					const int pr = (rank + C - 1) % C;
					xp = cluster.map_shared_rank(sCarryX, pr)[NW - 1];
					vp = cluster.map_shared_rank(sCarryV, pr)[NW - 1];
				}
			}
			if (leader) {
				int32_t red = sAllMax[0];
				for (int k = 1; k < C * NW; ++k) red = red > sAllMax[k] ? red : sAllMax[k];
				const int need = ld.mid(rows, b, r, qe, red, ls.V[0] << 16, sc.zdrop);
				int stop = 0;
				if (!need) {
					stop = ld.fin(rows, b, r, qe, b.en0, qlen, tlen, sc.zdrop, sc.e);
					if (!stop && okn) ld.pre(rows, bn, r + 1, qe);
				}
				for (int k = 0; k < C; ++k) {
					*cluster.map_shared_rank(&sNeedArg, k) = need;
					*cluster.map_shared_rank(&sGmax, k) = ld.gmax;
					*cluster.map_shared_rank(&sStop, k) = stop;
				}
			}
			const int t0a = ls.t0[0], t0b = ls.t0[1];                     // slot bases of diagonal r, for its arg-max pass
			if (okn) lane16_prepare<NS>(ls, bn, r + 1, b.en, qseq, tseq, tlen, table_saddr, sc16);
			cluster.sync();                                                                    // 2
			if (sNeedArg) {
				uint32_t key = lane16_argmax_key<GC>(b, Hrow, sGmax, t0a, t0b);
				key = __reduce_min_sync(0xffffffffu, key);
				if (lane == 0) key0[rank * NW + wid] = key;
				cluster.sync();                                                                // 3
				if (leader) {
					uint32_t k = ld.en0_key(b, r);
					for (int j = 0; j < C * NW; ++j) k = sAllKey[j] < k ? sAllKey[j] : k;
					const int stop = ld.fin(rows, b, r, qe, tie_key_slot(k, b.en0), qlen, tlen, sc.zdrop, sc.e);
					if (!stop && okn) ld.pre(rows, bn, r + 1, qe);
					for (int kk = 0; kk < C; ++kk) *cluster.map_shared_rank(&sStop, kk) = stop;
				}
				cluster.sync();                                                                // 4
			}
			const int stop = sStop;
			n_diag = r + 1;
			last_st = b.st; last_en = b.en;
			if (stop) break;
			if (r + 1 < R && !okn) { zdropped_band = 1; break; }                               // band exhausted (:110-113)
			b = bn;
		}
		if (leader) { if (zdropped_band) ld.ez.zdropped = 1; ld.store(&L.results[pi], n_diag); }
		cluster.sync();
	}
}

} // namespace extz
