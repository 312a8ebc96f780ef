// extz_dp.cuh -- the anti-diagonal DP kernels (K1/K2 of SURVEY.md section 2.1) for sm_100a.
// See extz_core.cuh for the representation; this file is the warp / CTA choreography.
//
// Two kernels share the same per-lane code:
//   extz_dp_kernel<G,S>       G <= 32 lanes per pair (32/G pairs per warp): carries by __shfl, reductions by
//                             REDUX, ordering by __syncwarp.  NS = G*S <= 1024 live slots.
//   extz_dp_wide_kernel<G,S>  one CTA of G = 64..256 lanes per pair: carries between warps and the
//                             per-diagonal max go through shared memory, ordering by __syncthreads.
//                             NS = G*S up to 4096 live slots (unbanded gap fills of a few kbp).
#pragma once
#include <cuda_runtime.h>
#include "extz_core.cuh"
#include "launch_structs.h"

namespace extz {

// sequence symbols are masked to 0..7: a byte the caller should not have passed (the engine reports KSW_B200_ERR_ARG for the
// batch, engine.cu) must not index past the 8 x 8 score table meanwhile
__device__ __forceinline__ uint32_t ld_u8(const uint8_t *p) { return (uint32_t)__ldg(p) & 7u; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
	uint32_t v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));   // the table is read-only after the prologue
	return v;
}

// ---- shared-memory index of circular slot c (conflict-free 128-bit rows: quad j of every lane is contiguous) ----
template <int G, int S>
__device__ __forceinline__ int hidx(int c)
{
	const int lane = c / S, i = c % S;                       // S is a power of two
	return (((i >> 2) * G + lane) << 2) | (i & 3);
}

// ---- per-lane state: S consecutive slots starting at t0 (circular window, slot t lives at t mod NS) ----
template <int S>
struct LaneState {
	uint32_t U[S], V[S], X[S], Y[S];   // u, v, x, y of extern/ksw2_extz2_sse.cc:54, top-byte form
	uint32_t Z[S];                      // s + 2(q+e) as last written by the score fill (persistent: stale outside the fill range)
	uint32_t TW[S / 4];                 // target symbols of the lane's slots, one byte each, pre-scaled by 32 (table row offset)
	uint32_t QW[S / 4];                 // sliding query window: byte i = 4 * query[r - (t0 + i)] (table column offset)
	int t0;
};

// 4 * query[j], or 0 where the reference reads the zeroed tail of qr[] (j < 0, App. A.1).  j may run up to 15 bytes past
// the query (values there are never used: the slot is below the band, its score is not refilled) but stays inside the
// arena, which ends with 256 bytes of slack.
__device__ __forceinline__ uint32_t qbyte4(const uint8_t *qseq, int j)
{
	return j >= 0 ? ld_u8(qseq + j) << 2 : 0u;
}

// (re)load the lane's slots for the window position t0, as seen at the START of anti-diagonal r (before its shift)
template <int S>
__device__ __forceinline__ void lane_load_slots(LaneState<S> &ls, const uint8_t *tseq, int tlen, const uint8_t *qseq, int r,
                                                const Scoring &sc)
{
#pragma unroll
	for (int i = 0; i < S; ++i) { ls.U[i] = ls.V[i] = ls.X[i] = ls.Y[i] = 0u; ls.Z[i] = sc.s0_s; }   // calloc'ed arrays (:83)
#pragma unroll
	for (int k = 0; k < S / 4; ++k) {
		uint32_t tw = 0, qw = 0;
#pragma unroll
		for (int b = 0; b < 4; ++b) {
			const int t = ls.t0 + 4 * k + b;
			tw |= (t < tlen ? ld_u8(tseq + t) << 5 : 0u) << (8 * b);                                 // sf[] reads 0 beyond tlen (App. A.1)
			qw |= qbyte4(qseq, r - 1 - t) << (8 * b);
		}
		ls.TW[k] = tw; ls.QW[k] = qw;
	}
}

// window slide + query shift + top-row boundary + score fill for one anti-diagonal
template <int S, int NS>
__device__ __forceinline__ void lane_prepare(LaneState<S> &ls, const Band &b, int r, const uint8_t *qseq, const uint8_t *tseq,
                                             int tlen, uint32_t table_saddr, const Scoring &sc)
{
	// lanes whose slots all fell below the rounded range take the slots NS further up
	if (ls.t0 + S - 1 < b.st) { ls.t0 += NS; lane_load_slots<S>(ls, tseq, tlen, qseq, r, sc); }
	// the query slides past the slots by one per anti-diagonal: shift the byte window, inject query[r - t0]
	{
		const uint32_t nb = qbyte4(qseq, r - ls.t0);
#pragma unroll
		for (int k = S / 4 - 1; k > 0; --k) ls.QW[k] = __funnelshift_l(ls.QW[k - 1], ls.QW[k], 8);
		ls.QW[0] = (ls.QW[0] << 8) | nb;
	}
	// top-row boundary (:122): y[r] = 0, u[r] = r ? q : 0 when the rounded range reaches slot r
	if (b.en >= r) {
		const int k = r - ls.t0;
		if ((unsigned)k < (unsigned)S) {
#pragma unroll
			for (int i = 0; i < S; ++i) if (k == i) { ls.Y[i] = 0u; ls.U[i] = r ? sc.q_s : 0u; }
		}
	}
	// score fill (:124-141): slots st0..fe get a fresh s, all others keep the stale one.  Branch-free: the
	// table offsets of 4 slots come from one packed add, the in-range mask is tested bit by bit (R2P).
	int lo = b.st0 - ls.t0, hi = b.fe - ls.t0 + 1;
	lo = lo < 0 ? 0 : (lo > S ? S : lo);
	hi = hi < 0 ? 0 : (hi > S ? S : hi);
	const uint32_t fillmask = hi > lo ? ((hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u)) : 0u;
#pragma unroll
	for (int k = 0; k < S / 4; ++k) {
		const uint32_t offs = ls.TW[k] + ls.QW[k];                 // byte b: 32*target + 4*query <= 252
#pragma unroll
		for (int bb = 0; bb < 4; ++bb) {
			const int i = 4 * k + bb;
			const uint32_t off = __byte_perm(offs, 0u, 0x4440 | bb);
			const uint32_t zn = lds_u32(table_saddr + off);
			if (fillmask & (1u << i)) ls.Z[i] = zn;
		}
	}
}

// the cells of one anti-diagonal for this lane (:149-220), traceback codes, u' dump and lazy-H update (:233-255).
// Returns the maximum of the lane's updated lazy-H entries (slot en0 was knocked out by the leader beforehand).
// G / lane: lanes that share the H rows (one CTA) and this lane's index among them; GT / lane_tb: lanes of the whole pair
// (a cluster spans several CTAs) and this lane's index among those -- they differ only in the cluster kernel.
template <int G, int S, bool kCigar, bool kRight, int GT = G>
__device__ __forceinline__ int32_t lane_cells(LaneState<S> &ls, const Band &b, int r, int last_st, int lane, uint32_t xin, uint32_t vin,
                                              uint8_t *tbp, int32_t *H, uint32_t *Us, const Scoring &sc, int lane_tb = -1)
{
	constexpr int NS = GT * S;
	if (lane_tb < 0) lane_tb = lane;
	constexpr int NSUB = (S + 15) / 16;                 // 16-slot sub-blocks per lane (1 unless S == 32)
	constexpr int SUBW = S < 16 ? S : 16;
	int32_t lane_max = kNegInf;
#pragma unroll
	for (int sb = NSUB - 1; sb >= 0; --sb) {
		const int tb0 = ls.t0 + sb * 16;
		const bool active = (tb0 >= b.st) && (tb0 <= b.en);
		if (active) {
			// carry into the first slot of this 16-block: OLD x,v of slot t-1 (:28-35), or the boundary values when
			// the block is the first of the rounded range (:117-121)
			uint32_t xc = (sb == 0) ? xin : ls.X[sb * 16 - 1];
			uint32_t vc = (sb == 0) ? vin : ls.V[sb * 16 - 1];
			uint32_t orv = 0u;
			if (tb0 == b.st) {
				if (b.st > 0) { if (!(b.st > last_st)) xc = vc = 0u; }      // slot st-1 was not computed on the last diagonal
				else { xc = 0u; vc = r ? sc.q_s : 0u; }
				// x1 / v1 are int8_t and go through _mm_cvtsi32_si128() (:102,144-145): a carry byte >= 0x80 is sign-extended into
				// lanes 1..3, and the _mm_or_si128 of the first block (:30,34) turns x[t-1] / v[t-1] of slots st+1..st+3 into 0xff
				// (x is always in [0,127]; v reaches 128+ once 2(q+e) + match exceeds 127 -- never with SEDEF's scoring)
				// only v can trigger it: x = and(cmpgt(a, 0), a) is always in [0,127]
				if (vc & 0x80000000u) orv = 0xff000000u;
			}
			uint32_t codes = 0;
#pragma unroll
			for (int ii = SUBW - 1; ii >= 0; --ii) {        // descending: slot i reads the OLD x,v of slot i-1
				const int i = sb * 16 + ii;
				uint32_t xt1 = (ii == 0) ? xc : ls.X[i - 1];
				uint32_t vt1 = (ii == 0) ? vc : (ls.V[i - 1] | ((ii <= 3) ? orv : 0u));
				uint32_t c = cell<kRight, kCigar>(ls.Z[i], xt1, vt1, ls.U[i], ls.V[i], ls.X[i], ls.Y[i], sc);
				if (kCigar) codes |= c << ((ii & 7) * 4);
				if (kCigar && (ii & 7) == 0) {
					// 8 codes = one 32-bit word; S == 4 packs 4 codes into 16 bits
					const int c0 = lane_tb * S + i;                                // == (t0 + i) mod NS
					uint8_t *dst = tbp + (int64_t)r * (NS >> 1) + (c0 >> 1);
					if (S >= 8) *(uint32_t *)dst = codes; else *(uint16_t *)dst = (uint16_t)codes;
					codes = 0;
				}
			}
#pragma unroll
			for (int ii = 0; ii < SUBW; ii += 4) {
				const int i = sb * 16 + ii;
				const int h0 = (((i >> 2) * G + lane) << 2);                       // hidx of the lane's quad
				*(uint4 *)&Us[h0] = make_uint4(ls.U[i], ls.U[i + 1], ls.U[i + 2], ls.U[i + 3]);
				int4 h = *(int4 *)&H[h0];
				h.x += (int32_t)(ls.V[i] >> 24); h.y += (int32_t)(ls.V[i + 1] >> 24);
				h.z += (int32_t)(ls.V[i + 2] >> 24); h.w += (int32_t)(ls.V[i + 3] >> 24);
				*(int4 *)&H[h0] = h;
				int32_t m01 = h.x > h.y ? h.x : h.y, m23 = h.z > h.w ? h.z : h.w;
				int32_t m = m01 > m23 ? m01 : m23;
				lane_max = lane_max > m ? lane_max : m;
			}
		}
	}
	return lane_max;
}

// arg-max, fast pass: (count << 24) + sum of t over this lane's slots whose lazy H equals gm.  Exited and not yet
// entered slots hold ~KSW_NEG_INF and can never match; slot en0 (register value = knocked-out) is added by the leader.
template <int G, int S>
__device__ __forceinline__ uint32_t lane_argmax_count(const LaneState<S> &ls, int lane, const int32_t *H, int32_t gm)
{
	uint32_t acc = 0;
#pragma unroll
	for (int j = 0; j < S / 4; ++j) {
		const int4 h = *(const int4 *)&H[((j * G + lane) << 2)];
#ifdef EXTZ_OPT_QUAD
		const int32_t m01 = h.x > h.y ? h.x : h.y, m23 = h.z > h.w ? h.z : h.w;
		if ((m01 > m23 ? m01 : m23) != gm) continue;             // only the quad holding the diagonal maximum is examined
#endif
		const uint32_t base = (1u << 24) + (uint32_t)(ls.t0 + 4 * j);
		if (h.x == gm) acc += base;
		if (h.y == gm) acc += base + 1;
		if (h.z == gm) acc += base + 2;
		if (h.w == gm) acc += base + 3;
	}
	// only "count == 1" matters; clamp to 2 so that the group sum of up to 1024 equal maxima cannot wrap to 1 (extz_dp16.cuh)
	return (acc >> 24) > 1u ? (2u << 24) : acc;
}

// arg-max, exact pass (only on real ties): smallest tie-break key among this lane's slots whose lazy H equals gm
template <int G, int S>
__device__ __forceinline__ uint32_t lane_argmax_key(const LaneState<S> &ls, const Band &b, int lane, const int32_t *H, int32_t gm)
{
	uint32_t key = 0xffffffffu;
#pragma unroll
	for (int i = 0; i < S; ++i) {
		int t = ls.t0 + i;
		if (t >= b.st0 && t <= b.en0 && !(t == b.en0 && b.en0 > 0) && H[(((i >> 2) * G + lane) << 2) | (i & 3)] == gm) {
			uint32_t k = tie_key(t, b.st0, b.en0);
			key = k < key ? k : key;
		}
	}
	return key;
}

// H / u' rows in the CTA's own shared memory (narrow and wide kernels)
template <int G, int S>
struct LocalRows {
	static constexpr int kMask = G * S - 1;
	int32_t *H; uint32_t *Us;
	__device__ __forceinline__ int32_t &h(int c) const { return H[hidx<G, S>(c)]; }
	__device__ __forceinline__ uint32_t u(int c) const { return Us[hidx<G, S>(c)]; }
};

// ---- the scalar per-pair bookkeeping done by the group leader (:222-267) -----------------------------------
struct Leader {
	EzState ez;
	int st0_prev;
	int exit_slot; int32_t exit_H;          // most recently exited slot and its TRUE H (stale reads of :228)
	int32_t Hprev_true;                     // H[en0-1] of the previous diagonal
	int32_t Hen0_lazy, gmax;
	int32_t H0, last_H0_t;                  // KSW_EZ_APPROX_MAX: the one tracked score and its slot (:51,268-284)

	__device__ __forceinline__ void reset() { ez_reset(ez); st0_prev = 0; exit_slot = -2; exit_H = kNegInf; Hprev_true = kNegInf; Hen0_lazy = gmax = kNegInf; H0 = 0; last_H0_t = 0; }

	// KSW_EZ_APPROX_MAX (:268-284): instead of the exact maximum over the anti-diagonal, ONE score H0 is walked along the
	// locally better of the two moves (stay on slot t: + v[t] - qe; step to slot t+1: + u[t+1] - qe).  Every slot it reads is
	// in band: last_H0_t never exceeds en0 (en0 is non-decreasing) and falls at most one slot behind st0, where the third arm
	// steps it back to st0.  mqe / mte are never updated; max / z-drop only with KSW_EZ_APPROX_DROP.  `A` additionally gives
	// a.v(c), the v' of the current diagonal.  Returns stop.
	template <class A>
	__device__ __forceinline__ int approx(const A &acc, const Band &b, int r, int qe, uint32_t v0_r0, int qlen, int tlen, int zdrop, int e, bool drop)
	{
		if (r > 0) {
			const bool in0 = last_H0_t >= b.st0 && last_H0_t <= b.en0;
			const bool in1 = last_H0_t + 1 >= b.st0 && last_H0_t + 1 <= b.en0;
			if (in0 && in1) {
				const int32_t d0 = (int32_t)(acc.v(last_H0_t & A::kMask) >> 24) - qe;
				const int32_t d1 = (int32_t)(acc.u((last_H0_t + 1) & A::kMask) >> 24) - qe;
				if (d0 > d1) H0 += d0;
				else { H0 += d1; ++last_H0_t; }
			} else if (in0) H0 += (int32_t)(acc.v(last_H0_t & A::kMask) >> 24) - qe;
			else { ++last_H0_t; H0 += (int32_t)(acc.u(last_H0_t & A::kMask) >> 24) - qe; }
			if (drop && ez_apply_zdrop(ez, H0, r, last_H0_t, zdrop, e)) return 1;
		} else { H0 = (int32_t)(v0_r0 >> 24) - 2 * qe; last_H0_t = 0; }
		if (r == qlen + tlen - 2 && b.en0 == tlen - 1) ez.score = H0;
		return 0;
	}

	// `A` gives access to the H / u' rows by circular slot index: a.h(c) (int32_t&), a.u(c) (uint32_t), A::kMask.
	// before the cells: remember/knock out slots that must not take part in the regular H update
	template <class A>
	__device__ __forceinline__ void pre(const A &acc, const Band &b, int r, int qe)
	{
		Hprev_true = kNegInf;
		if (r == 0) return;
		if (b.st0 > st0_prev) {                              // slot st0-1 left the band: keep its TRUE H, drop it from the max
			int xs = b.st0 - 1;
			int32_t &hx = acc.h(xs & A::kMask);
			exit_slot = xs; exit_H = hx - qe * (r - 1);
			hx = kNegInf;
		}
		if (b.en0 > 0) {
			int ps = b.en0 - 1;
			Hprev_true = (ps == exit_slot) ? exit_H : acc.h(ps & A::kMask) - qe * (r - 1);
			acc.h(b.en0 & A::kMask) = kNegInf;               // the regular update must not count for slot en0
		}
	}
	// after the cells: H[en0] (:228 / :259), diagonal max in the lazy domain; returns need_arg
	template <class A>
	__device__ __forceinline__ int mid(const A &acc, const Band &b, int r, int qe, int32_t reduced_max, uint32_t v0_r0, int zdrop)
	{
		gmax = reduced_max;
		if (r == 0) { Hen0_lazy = (int32_t)(v0_r0 >> 24) - 2 * qe; acc.h(0) = Hen0_lazy; gmax = Hen0_lazy; }   // :259
		else if (b.en0 > 0) {
			Hen0_lazy = Hprev_true + (int32_t)(acc.u(b.en0 & A::kMask) >> 24) - qe + qe * r;                  // :228, true -> lazy
			acc.h(b.en0 & A::kMask) = Hen0_lazy;
			gmax = gmax > Hen0_lazy ? gmax : Hen0_lazy;
		} else Hen0_lazy = acc.h(0);                          // en0 == 0: regular update (:228 else-arm)
		int32_t maxH_true = gmax - qe * r;
		// ksw_apply_zdrop (extern/ksw2.h:161-177) only looks at the arg-max slot when the maximum improves, or when
		// max - H > zdrop (+ l*e >= 0) can fire; in every other case it has no effect, so the arg-max is not needed
		return (maxH_true > ez.max) || (zdrop >= 0 && ez.max - maxH_true > zdrop);
	}
	__device__ __forceinline__ uint32_t en0_key(const Band &b, int r) const
	{
		return (Hen0_lazy == gmax && (r == 0 || b.en0 > 0)) ? 0u : 0xffffffffu;   // slot en0 wins every tie (:229-231)
	}
	// end scores, z-drop (:261-267); returns stop
	template <class A>
	__device__ __forceinline__ int fin(const A &acc, const Band &b, int r, int qe, int max_t, int qlen, int tlen, int zdrop, int e)
	{
		const int R = qlen + tlen - 1;
		int32_t Hen0_true = Hen0_lazy - qe * r;
		int32_t maxH_true = gmax - qe * r;
		if (b.en0 == tlen - 1 && Hen0_true > ez.mte) { ez.mte = Hen0_true; ez.mte_q = r - b.en; }          // :261-262
		if (r - b.st0 == qlen - 1) {                                                                      // :263-264
			int32_t Hst0 = (b.st0 == b.en0) ? Hen0_true : acc.h(b.st0 & A::kMask) - qe * r;
			if (Hst0 > ez.mqe) { ez.mqe = Hst0; ez.mqe_t = b.st0; }
		}
		int stop = 0;
		if (ez_apply_zdrop(ez, maxH_true, r, max_t, zdrop, e)) stop = 1;                                  // :265
		else if (r == R - 1 && b.en0 == tlen - 1) ez.score = Hen0_true;                                   // :266-267
		st0_prev = b.st0;
		return stop;
	}
	// end scores and z-drop as in fin(), for the leaderless kernels: every lane of the group runs it on identical inputs.
	// Hst0_lazy: lazy H[st0] (only read on diagonals that end on the last query row).
	__device__ __forceinline__ int fin_local(const Band &b, int r, int qe, int max_t, int32_t Hst0_lazy, int qlen, int tlen, int zdrop, int e)
	{
		const int R = qlen + tlen - 1;
		const int32_t Hen0_true = Hen0_lazy - qe * r;
		const int32_t maxH_true = gmax - qe * r;
		if (b.en0 == tlen - 1 && Hen0_true > ez.mte) { ez.mte = Hen0_true; ez.mte_q = r - b.en; }          // :261-262
		if (r - b.st0 == qlen - 1) {                                                                      // :263-264
			const int32_t Hst0 = (b.st0 == b.en0) ? Hen0_true : Hst0_lazy - qe * r;
			if (Hst0 > ez.mqe) { ez.mqe = Hst0; ez.mqe_t = b.st0; }
		}
		int stop = 0;
		if (ez_apply_zdrop(ez, maxH_true, r, max_t, zdrop, e)) stop = 1;                                  // :265
		else if (r == R - 1 && b.en0 == tlen - 1) ez.score = Hen0_true;                                   // :266-267
		st0_prev = b.st0;
		return stop;
	}
	__device__ __forceinline__ void store(PairResult *out, int n_diag) const
	{
		PairResult pr;
		pr.max = ez.max; pr.zdropped = ez.zdropped; pr.max_q = ez.max_q; pr.max_t = ez.max_t;
		pr.mqe = ez.mqe; pr.mqe_t = ez.mqe_t; pr.mte = ez.mte; pr.mte_q = ez.mte_q; pr.score = ez.score;
		pr.n_diag = n_diag; pr.n_cigar = 0; pr.cigar_off = 0;
		*out = pr;
	}
};

// ---- group collectives (all 32 lanes of the warp execute them; groups are G-aligned lane ranges) ----
// (G == 16: two full-warp REDUX with the other half masked to the identity -- two independent instructions instead of a chain of
// four dependent shuffle + op steps)
template <int G> __device__ __forceinline__ int32_t group_max(int32_t v)
{
	if (G == 32) return __reduce_max_sync(0xffffffffu, v);
	if (G == 16) {
		const bool hi = (threadIdx.x & 16) != 0;
		const int32_t a = __reduce_max_sync(0xffffffffu, hi ? (int32_t)0x80000000 : v), b = __reduce_max_sync(0xffffffffu, hi ? v : (int32_t)0x80000000);
		return hi ? b : a;
	}
#pragma unroll
	for (int d = 1; d < G; d <<= 1) { int32_t o = __shfl_xor_sync(0xffffffffu, v, d); v = v > o ? v : o; }
	return v;
}
template <int G> __device__ __forceinline__ uint32_t group_min_u(uint32_t v)
{
	if (G == 32) return __reduce_min_sync(0xffffffffu, v);
	if (G == 16) {
		const bool hi = (threadIdx.x & 16) != 0;
		const uint32_t a = __reduce_min_sync(0xffffffffu, hi ? 0xffffffffu : v), b = __reduce_min_sync(0xffffffffu, hi ? v : 0xffffffffu);
		return hi ? b : a;
	}
#pragma unroll
	for (int d = 1; d < G; d <<= 1) { uint32_t o = __shfl_xor_sync(0xffffffffu, v, d); v = v < o ? v : o; }
	return v;
}
template <int G> __device__ __forceinline__ uint32_t group_sum_u(uint32_t v)
{
	if (G == 32) return __reduce_add_sync(0xffffffffu, v);
	if (G == 16) {
		const bool hi = (threadIdx.x & 16) != 0;
		const uint32_t a = __reduce_add_sync(0xffffffffu, hi ? 0u : v), b = __reduce_add_sync(0xffffffffu, hi ? v : 0u);
		return hi ? b : a;
	}
#pragma unroll
	for (int d = 1; d < G; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
	return v;
}

// =====================================================================================================
// narrow kernel: G <= 32 lanes per pair, the 32/G pairs of a warp advance in lock-step
// =====================================================================================================
// Occupancy targets found by measurement on B200 (profiles/r01_tuning.md): 16 slots per lane want 3 CTAs/SM
// (168 registers, no spills): 341 GCUPS on config 2, against 293 (227 regs, 2 CTAs) and 335 (128 regs, spills).
#ifndef EXTZ_MIN_BLOCKS
#define EXTZ_MIN_BLOCKS 4
#endif
#ifndef EXTZ_MIN_BLOCKS16
#define EXTZ_MIN_BLOCKS16 3
#endif
template <int G, int S, bool kCigar, bool kRight>
__global__ void __launch_bounds__(128, (S <= 8 ? EXTZ_MIN_BLOCKS : (S == 16 ? EXTZ_MIN_BLOCKS16 : 1)))
extz_dp_kernel(DpLaunch L)
{
	constexpr int NS = G * S;
	constexpr int GROUPS_PER_BLOCK = 128 / G;
	constexpr int PPW = 32 / G;                          // pairs per warp
	constexpr unsigned FULL = 0xffffffffu;
	static_assert(G <= 32 && (NS & (NS - 1)) == 0, "NS must be a power of two");
	static_assert(S % 4 == 0 && (16 % S == 0 || S % 16 == 0), "lane slots must tile 16-slot blocks");

	__shared__ __align__(16) int32_t sH[GROUPS_PER_BLOCK][NS];        // lazy H: H + (q+e)*r   (:222-259)
	__shared__ __align__(16) uint32_t sU[GROUPS_PER_BLOCK][NS];       // u' of the current diagonal (for H[en0], :228)
	__shared__ uint32_t sTable[kTableStride * kTableStride];

	for (int i = threadIdx.x; i < kTableStride * kTableStride; i += blockDim.x) sTable[i] = L.table[i];
	__syncthreads();

	const uint32_t table_saddr = (uint32_t)__cvta_generic_to_shared(sTable);
	const int lane_w = threadIdx.x & 31;
	const int gl = threadIdx.x % G;                                  // lane within group
	const int gidx = threadIdx.x / G;                                // group within block
	const int pred_lane = (gl + G - 1) & (G - 1);                    // circular predecessor (relative to group)
	int32_t *H = sH[gidx];
	uint32_t *Us = sU[gidx];
	const LocalRows<G, S> rows{H, Us};
	const Scoring sc = L.sc;
	const int qe = sc.qe;
	const bool generic = (sc.flag & kFlagGenericSc) != 0;

	for (;;) {
		int base = 0;
		if (lane_w == 0) base = atomicAdd(L.work_counter, PPW);      // dynamic work queue, pairs sorted by descending work
		base = __shfl_sync(FULL, base, 0);
		if (base >= L.n) break;
		const int pi = base + lane_w / G;
		bool alive = pi < L.n;
		const PairDesc pd = L.pairs[alive ? pi : base];
		const int qlen = pd.qlen, tlen = pd.tlen, w = pd.w;
		const int T = (tlen + 15) & ~15;
		const uint8_t *qseq = L.seq + pd.q_off;                      // qseq[j], j in [0, qlen + 15] is read
		const uint8_t *tseq = L.seq + pd.t_off;
		uint8_t *tbp = kCigar ? L.tb + pd.tb_off : nullptr;
		const int R = alive ? qlen + tlen - 1 : 0;
		int maxR = R;
#pragma unroll
		for (int d = G; d < 32; d <<= 1) { int o = __shfl_xor_sync(FULL, maxR, d); maxR = maxR > o ? maxR : o; }

		LaneState<S> ls;
		ls.t0 = gl * S;
		lane_load_slots<S>(ls, tseq, tlen, qseq, 0, sc);
#pragma unroll
		for (int j = 0; j < S / 4; ++j) *(int4 *)&H[(j * G + gl) << 2] = make_int4(kNegInf, kNegInf, kNegInf, kNegInf);   // :86-89
		__syncwarp();

		Leader ld; ld.reset();                                       // meaningful in the leader lane only
		int last_st = -1, n_diag = 0, zdropped_band = 0;

		for (int r = 0; r < maxR; ++r) {
			bool act = alive && r < R;
			Band b;
			const bool okb = band_of(r, qlen, tlen, w, T, generic, b);
			if (act && !okb) { zdropped_band = 1; alive = false; act = false; }                // :110-113

			// carry from the circular predecessor: OLD x,v of its top slot (:28-35)
			const uint32_t xin = __shfl_sync(FULL, ls.X[S - 1], pred_lane, G);
			const uint32_t vin = __shfl_sync(FULL, ls.V[S - 1], pred_lane, G);
			if (act) {
				lane_prepare<S, NS>(ls, b, r, qseq, tseq, tlen, table_saddr, sc);
				if (gl == 0) ld.pre(rows, b, r, qe);
			}
			__syncwarp();
			int32_t lane_max = kNegInf;
			if (act) lane_max = lane_cells<G, S, kCigar, kRight>(ls, b, r, last_st, gl, xin, vin, tbp, H, Us, sc);
			__syncwarp();
			const int32_t red = group_max<G>(lane_max);
			int need = 0;
			if (act && gl == 0) need = ld.mid(rows, b, r, qe, red, ls.V[0], sc.zdrop);
			__syncwarp();
			need = __shfl_sync(FULL, need, 0, G);
			int max_t = b.en0;
			if (__any_sync(FULL, need)) {
				const int32_t gm = __shfl_sync(FULL, ld.gmax, 0, G);
				uint32_t cnt = 0;
				if (need) cnt = lane_argmax_count<G, S>(ls, gl, H, gm);
				cnt = group_sum_u<G>(cnt);
				max_t = (int)(cnt & 0x00ffffffu);
				const int tie = need && (cnt >> 24) != 1u;
				if (__any_sync(FULL, tie)) {                                                      // real ties: exact 4-lane rule
					uint32_t key = 0xffffffffu;
					if (tie) {
						key = lane_argmax_key<G, S>(ls, b, gl, H, gm);
						if (gl == 0) { uint32_t k0 = ld.en0_key(b, r); key = k0 < key ? k0 : key; }
					}
					key = group_min_u<G>(key);
					if (tie) max_t = tie_key_slot(key, b.en0);
				}
			}
			int stop = 0;
			if (act && gl == 0) stop = ld.fin(rows, b, r, qe, max_t, qlen, tlen, sc.zdrop, sc.e);
			stop = __shfl_sync(FULL, stop, 0, G);
			if (act) { n_diag = r + 1; last_st = b.st; if (stop) alive = false; }
			__syncwarp();            // the arg-max passes read H; the next diagonal's leader writes it (racecheck-clean ordering)
		}
		if (pi < L.n && gl == 0) { if (zdropped_band) ld.ez.zdropped = 1; ld.store(&L.results[pi], n_diag); }
		__syncwarp();
	}
}

// =====================================================================================================
// wide kernel: one CTA (G = blockDim.x = 64..256 lanes) per pair
// =====================================================================================================
template <int G, int S, bool kCigar, bool kRight>
__global__ void __launch_bounds__(G)
extz_dp_wide_kernel(DpLaunch L)
{
	constexpr int NS = G * S;
	constexpr int NW = G / 32;
	static_assert(G > 32 && G % 32 == 0 && (NS & (NS - 1)) == 0, "wide kernel: whole warps, NS power of two");
	static_assert(S % 4 == 0 && (16 % S == 0 || S % 16 == 0), "lane slots must tile 16-slot blocks");

	__shared__ __align__(16) int32_t H[NS];
	__shared__ __align__(16) uint32_t Us[NS];
	__shared__ uint32_t sTable[kTableStride * kTableStride];
	__shared__ uint32_t sCarryX[NW], sCarryV[NW];       // OLD x,v of every warp's top slot
	__shared__ int32_t sWarpMax[NW];
	__shared__ uint32_t sWarpKey[NW];
	__shared__ int sPair, sNeedArg, sStop;
	__shared__ int32_t sGmax;

	for (int i = threadIdx.x; i < kTableStride * kTableStride; i += blockDim.x) sTable[i] = L.table[i];
	__syncthreads();

	const uint32_t table_saddr = (uint32_t)__cvta_generic_to_shared(sTable);
	const LocalRows<G, S> rows{H, Us};
	const int gl = threadIdx.x;
	const int lane = gl & 31, wid = gl >> 5;
	const Scoring sc = L.sc;
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
		uint8_t *tbp = kCigar ? L.tb + pd.tb_off : nullptr;

		LaneState<S> ls;
		ls.t0 = gl * S;
		lane_load_slots<S>(ls, tseq, tlen, qseq, 0, sc);
#pragma unroll
		for (int j = 0; j < S / 4; ++j) *(int4 *)&H[(j * G + gl) << 2] = make_int4(kNegInf, kNegInf, kNegInf, kNegInf);
		Leader ld; ld.reset();
		int last_st = -1, n_diag = 0, zdropped_band = 0;
		const int R = qlen + tlen - 1;
		__syncthreads();

		for (int r = 0; r < R; ++r) {
			Band b;
			if (!band_of(r, qlen, tlen, w, T, generic, b)) { zdropped_band = 1; break; }

			// phase 1: publish the OLD top slot of every warp; the leader prepares H (nobody else touches H now)
			uint32_t xin = __shfl_up_sync(0xffffffffu, ls.X[S - 1], 1);
			uint32_t vin = __shfl_up_sync(0xffffffffu, ls.V[S - 1], 1);
			if (lane == 31) { sCarryX[wid] = ls.X[S - 1]; sCarryV[wid] = ls.V[S - 1]; }
			if (gl == 0) ld.pre(rows, b, r, qe);
			__syncthreads();                                                                   // A
			// phase 2: cells
			if (lane == 0) { int pw = (wid + NW - 1) % NW; xin = sCarryX[pw]; vin = sCarryV[pw]; }
			lane_prepare<S, NS>(ls, b, r, qseq, tseq, tlen, table_saddr, sc);
			int32_t lane_max = lane_cells<G, S, kCigar, kRight>(ls, b, r, last_st, gl, xin, vin, tbp, H, Us, sc);
			int32_t wmax = __reduce_max_sync(0xffffffffu, lane_max);
			if (lane == 0) sWarpMax[wid] = wmax;
			__syncthreads();                                                                   // B
			// phase 3: leader
			if (gl == 0) {
				int32_t red = sWarpMax[0];
#pragma unroll
				for (int k = 1; k < NW; ++k) red = red > sWarpMax[k] ? red : sWarpMax[k];
				int need = ld.mid(rows, b, r, qe, red, ls.V[0], sc.zdrop);
				sNeedArg = need; sGmax = ld.gmax;
				if (!need) sStop = ld.fin(rows, b, r, qe, b.en0, qlen, tlen, sc.zdrop, sc.e);
			}
			__syncthreads();                                                                   // C
			if (sNeedArg) {
				uint32_t key = lane_argmax_key<G, S>(ls, b, gl, H, sGmax);
				key = __reduce_min_sync(0xffffffffu, key);
				if (lane == 0) sWarpKey[wid] = key;
				__syncthreads();                                                               // D
				if (gl == 0) {
					uint32_t k = ld.en0_key(b, r);
#pragma unroll
					for (int j = 0; j < NW; ++j) k = sWarpKey[j] < k ? sWarpKey[j] : k;
					sStop = ld.fin(rows, b, r, qe, tie_key_slot(k, b.en0), qlen, tlen, sc.zdrop, sc.e);
				}
				__syncthreads();                                                               // E
			}
			const int stop = sStop;
			n_diag = r + 1;
			last_st = b.st;
			if (stop) break;
		}
		if (gl == 0) { if (zdropped_band) ld.ez.zdropped = 1; ld.store(&L.results[pi], n_diag); }
		__syncthreads();
	}
}

// =====================================================================================================
// cluster kernel: one thread-block CLUSTER of C CTAs (256 lanes each) per pair -- 8192 / 16384 live slots, for the
// unbanded multi-kbp gap fills SEDEF issues (src/align.cc:130-139 with MAX_GAP = 10 kbp, src/refine.cc:77).
// Every CTA keeps the H / u' rows of its own lanes in its shared memory; the carries between CTAs, the leader's
// accesses to arbitrary slots and the per-diagonal reductions go through DISTRIBUTED SHARED MEMORY
// (cluster.map_shared_rank), ordering by cluster.sync().
// =====================================================================================================
} // namespace extz
#include <cooperative_groups.h>
namespace extz {
namespace cg = cooperative_groups;

template <int C, int S>
struct ClusterRows {                                    // H / u' rows spread over the CTAs of the cluster
	static constexpr int GC = 256, NSC = GC * S, kMask = C * NSC - 1;
	int32_t *H; uint32_t *Us;                           // this CTA's arrays (same offset in every CTA)
	__device__ __forceinline__ int32_t &h(int c) const
	{
		cg::cluster_group cl = cg::this_cluster();
		return cl.map_shared_rank(H, c / NSC)[hidx<GC, S>(c % NSC)];
	}
	__device__ __forceinline__ uint32_t u(int c) const
	{
		cg::cluster_group cl = cg::this_cluster();
		return cl.map_shared_rank(Us, c / NSC)[hidx<GC, S>(c % NSC)];
	}
};

template <int C, int S, bool kCigar, bool kRight>
__global__ void __launch_bounds__(256, 1)
extz_dp_cluster_kernel(DpLaunch L)
{
	constexpr int GC = 256, G = GC * C, NS = G * S, NW = GC / 32;
	static_assert(S == 16 && (C == 2 || C == 4 || C == 8), "cluster kernel: 16 slots per lane, 2/4/8 CTAs");
	cg::cluster_group cluster = cg::this_cluster();
	const int rank = (int)cluster.block_rank();

	__shared__ __align__(16) int32_t H[GC * S];
	__shared__ __align__(16) uint32_t Us[GC * S];
	__shared__ uint32_t sTable[kTableStride * kTableStride];
	__shared__ uint32_t sCarryX[NW], sCarryV[NW];       // OLD x,v of every warp's top slot (read by the next warp / next CTA)
	__shared__ int32_t sAllMax[8 * NW];                 // [rank][warp], only CTA 0's copy is used
	__shared__ uint32_t sAllKey[8 * NW];
	__shared__ int sPair, sNeedArg, sStop;              // written into EVERY CTA's copy by the leader
	__shared__ int32_t sGmax;

	for (int i = threadIdx.x; i < kTableStride * kTableStride; i += blockDim.x) sTable[i] = L.table[i];
	// cluster.sync(), not __syncthreads(): the first thing the leader does is write sPair into the OTHER CTAs' shared memory,
	// and a peer's distributed shared memory may only be touched once that CTA is known to be running (racecheck:
	// "write to a block that might not have entered yet").  It also orders the table fill inside each CTA.
	cluster.sync();

	const uint32_t table_saddr = (uint32_t)__cvta_generic_to_shared(sTable);
	const ClusterRows<C, S> rows{H, Us};
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int gl = rank * GC + tid;                      // lane within the pair
	const bool leader = gl == 0;
	const Scoring sc = L.sc;
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
		uint8_t *tbp = kCigar ? L.tb + pd.tb_off : nullptr;

		LaneState<S> ls;
		ls.t0 = gl * S;
		lane_load_slots<S>(ls, tseq, tlen, qseq, 0, sc);
#pragma unroll
		for (int j = 0; j < S / 4; ++j) *(int4 *)&H[(j * GC + tid) << 2] = make_int4(kNegInf, kNegInf, kNegInf, kNegInf);
		Leader ld; ld.reset();
		int last_st = -1, n_diag = 0, zdropped_band = 0;
		const int R = qlen + tlen - 1;
		cluster.sync();

		for (int r = 0; r < R; ++r) {
			Band b;
			if (!band_of(r, qlen, tlen, w, T, generic, b)) { zdropped_band = 1; break; }

			// phase 1: publish the OLD top slot of every warp; the leader prepares H (nobody else touches H now)
			uint32_t xin = __shfl_up_sync(0xffffffffu, ls.X[S - 1], 1);
			uint32_t vin = __shfl_up_sync(0xffffffffu, ls.V[S - 1], 1);
			if (lane == 31) { sCarryX[wid] = ls.X[S - 1]; sCarryV[wid] = ls.V[S - 1]; }
			if (leader) ld.pre(rows, b, r, qe);
			cluster.sync();                                                                    // A
			// phase 2: cells; warp 0 of a CTA takes its carry from the last warp of the previous CTA (DSMEM)
			if (lane == 0) {
				if (wid > 0) { xin = sCarryX[wid - 1]; vin = sCarryV[wid - 1]; }
				else {
					const int pr = (rank + C - 1) % C;
					xin = cluster.map_shared_rank(sCarryX, pr)[NW - 1];
					vin = cluster.map_shared_rank(sCarryV, pr)[NW - 1];
				}
			}
			lane_prepare<S, NS>(ls, b, r, qseq, tseq, tlen, table_saddr, sc);
			int32_t lane_max = lane_cells<GC, S, kCigar, kRight, G>(ls, b, r, last_st, tid, xin, vin, tbp, H, Us, sc, gl);
			int32_t wmax = __reduce_max_sync(0xffffffffu, lane_max);
			if (lane == 0) max0[rank * NW + wid] = wmax;
			cluster.sync();                                                                    // B
			// phase 3: leader
			if (leader) {
				int32_t red = sAllMax[0];
				for (int k = 1; k < C * NW; ++k) red = red > sAllMax[k] ? red : sAllMax[k];
				const int need = ld.mid(rows, b, r, qe, red, ls.V[0], sc.zdrop);
				int stop = 0;
				if (!need) stop = ld.fin(rows, b, r, qe, b.en0, qlen, tlen, sc.zdrop, sc.e);
				for (int k = 0; k < C; ++k) {
					*cluster.map_shared_rank(&sNeedArg, k) = need;
					*cluster.map_shared_rank(&sGmax, k) = ld.gmax;
					*cluster.map_shared_rank(&sStop, k) = stop;
				}
			}
			cluster.sync();                                                                    // C
			if (sNeedArg) {
				uint32_t key = lane_argmax_key<GC, S>(ls, b, tid, H, sGmax);
				key = __reduce_min_sync(0xffffffffu, key);
				if (lane == 0) key0[rank * NW + wid] = key;
				cluster.sync();                                                                // D
				if (leader) {
					uint32_t k = ld.en0_key(b, r);
					for (int j = 0; j < C * NW; ++j) k = sAllKey[j] < k ? sAllKey[j] : k;
					const int stop = ld.fin(rows, b, r, qe, tie_key_slot(k, b.en0), qlen, tlen, sc.zdrop, sc.e);
					for (int kk = 0; kk < C; ++kk) *cluster.map_shared_rank(&sStop, kk) = stop;
				}
				cluster.sync();                                                                // E
			}
			const int stop = sStop;
			n_diag = r + 1;
			last_st = b.st;
			if (stop) break;
		}
		if (leader) { if (zdropped_band) ld.ez.zdropped = 1; ld.store(&L.results[pi], n_diag); }
		cluster.sync();
	}
}

} // namespace extz
