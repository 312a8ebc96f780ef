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

namespace extz {

// Device-side view of one launch.
struct DpLaunch {
	const PairDesc *pairs;      // [n] sorted by descending work
	PairResult *results;        // [n] (indexed like pairs)
	const uint8_t *seq;         // packed sequence arena (codes)
	uint8_t *tb;                // traceback arena of this wave (or nullptr when score-only)
	const uint32_t *table;      // [kTableStride * kTableStride] (s + 2(q+e)) << 24 per (target, query) symbol
	int *work_counter;          // dynamic work distribution
	int n;
	Scoring sc;
};

__device__ __forceinline__ uint32_t ld_u8(const uint8_t *p) { return (uint32_t)__ldg(p); }

// ---- per-lane state: S consecutive slots starting at t0 (circular window, slot t lives at t mod NS) ----
template <int S>
struct LaneState {
	uint32_t U[S], V[S], X[S], Y[S];   // u, v, x, y of extern/ksw2_extz2_sse.cc:54, top-byte form
	uint32_t Z[S];                      // s + 2(q+e) as last written by the score fill (persistent: stale outside the fill range)
	uint32_t TC[S];                     // byte offset of the slot's target symbol row in the score table
	int t0;
};

template <int S>
__device__ __forceinline__ void lane_load_slots(LaneState<S> &ls, const uint8_t *tseq, int tlen, const Scoring &sc)
{
#pragma unroll
	for (int i = 0; i < S; ++i) {
		ls.U[i] = ls.V[i] = ls.X[i] = ls.Y[i] = 0u; ls.Z[i] = sc.s0_s;                  // calloc'ed arrays (:83)
		int t = ls.t0 + i;
		ls.TC[i] = (t < tlen ? ld_u8(tseq + t) : 0u) * (kTableStride * 4);          // sf[] reads 0 beyond tlen (App. A.1)
	}
}

// window slide + top-row boundary + score fill for one anti-diagonal
template <int S, int NS>
__device__ __forceinline__ void lane_prepare(LaneState<S> &ls, const Band &b, int r, const uint8_t *qseq, const uint8_t *tseq,
                                             int tlen, const uint32_t *sTable, const Scoring &sc)
{
	// lanes whose slots all fell below the rounded range take the slots NS further up
	if (ls.t0 + S - 1 < b.st) { ls.t0 += NS; lane_load_slots<S>(ls, tseq, tlen, sc); }
	// top-row boundary (:122): y[r] = 0, u[r] = r ? q : 0 when the rounded range reaches slot r
	if (b.en >= r) {
		int k = r - ls.t0;
#pragma unroll
		for (int i = 0; i < S; ++i) if (k == i) { ls.Y[i] = 0u; ls.U[i] = r ? sc.q_s : 0u; }
	}
	// score fill (:124-141): slots st0..fe get a fresh s, all others keep the stale one
	const uint8_t *qp = qseq + (r - ls.t0);                      // query[r - t] for slot t = t0 + i is qp[-i]
	const int lo = b.st0 - ls.t0, hi = b.fe - ls.t0;
#pragma unroll
	for (int i = 0; i < S; ++i) {
		if (i >= lo && i <= hi) {
			uint32_t qc = ld_u8(qp - i);
			ls.Z[i] = *(const uint32_t *)((const char *)sTable + ls.TC[i] + qc * 4);
		}
	}
}

// the cells of one anti-diagonal for this lane (:149-220), traceback codes, u' dump and lazy-H update (:233-255).
// Returns the maximum of the lane's updated lazy-H entries (slot en0 was knocked out by the leader beforehand).
template <int S, int NS, bool kCigar, bool kRight>
__device__ __forceinline__ int32_t lane_cells(LaneState<S> &ls, const Band &b, int r, int last_st, uint32_t xin, uint32_t vin,
                                              uint8_t *tbp, int32_t *H, uint32_t *Us, const Scoring &sc)
{
	constexpr int MASK = NS - 1;
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
			if (tb0 == b.st) {
				if (b.st > 0) { if (!(b.st > last_st)) xc = vc = 0u; }      // slot st-1 was not computed on the last diagonal
				else { xc = 0u; vc = r ? sc.q_s : 0u; }
			}
			uint32_t codes = 0;
#pragma unroll
			for (int ii = SUBW - 1; ii >= 0; --ii) {        // descending: slot i reads the OLD x,v of slot i-1
				const int i = sb * 16 + ii;
				uint32_t xt1 = (ii == 0) ? xc : ls.X[i - 1];
				uint32_t vt1 = (ii == 0) ? vc : ls.V[i - 1];
				uint32_t c = cell<kRight, kCigar>(ls.Z[i], xt1, vt1, ls.U[i], ls.V[i], ls.X[i], ls.Y[i], sc);
				if (kCigar) codes |= c << ((ii & 7) * 4);
				if (kCigar && (ii & 7) == 0) {
					// 8 codes = one 32-bit word; S == 4 packs 4 codes into 16 bits
					int c0 = (ls.t0 + i) & MASK;
					uint8_t *dst = tbp + (int64_t)r * (NS >> 1) + (c0 >> 1);
					if (S >= 8) *(uint32_t *)dst = codes; else *(uint16_t *)dst = (uint16_t)codes;
					codes = 0;
				}
			}
#pragma unroll
			for (int ii = 0; ii < SUBW; ii += 4) {
				const int i = sb * 16 + ii;
				const int c0 = (ls.t0 + i) & MASK;
				*(uint4 *)&Us[c0] = make_uint4(ls.U[i], ls.U[i + 1], ls.U[i + 2], ls.U[i + 3]);
				int4 h = *(int4 *)&H[c0];
				h.x += (int32_t)(ls.V[i] >> 24); h.y += (int32_t)(ls.V[i + 1] >> 24);
				h.z += (int32_t)(ls.V[i + 2] >> 24); h.w += (int32_t)(ls.V[i + 3] >> 24);
				*(int4 *)&H[c0] = h;
				int32_t m01 = h.x > h.y ? h.x : h.y, m23 = h.z > h.w ? h.z : h.w;
				int32_t m = m01 > m23 ? m01 : m23;
				lane_max = lane_max > m ? lane_max : m;
			}
		}
	}
	return lane_max;
}

// smallest tie-break key among this lane's slots whose lazy H equals gm (slot en0 is handled by the leader)
template <int S, int NS>
__device__ __forceinline__ uint32_t lane_argmax_key(const LaneState<S> &ls, const Band &b, const int32_t *H, int32_t gm)
{
	constexpr int MASK = NS - 1;
	uint32_t key = 0xffffffffu;
#pragma unroll
	for (int i = 0; i < S; ++i) {
		int t = ls.t0 + i;
		if (t >= b.st0 && t <= b.en0 && !(t == b.en0 && b.en0 > 0) && H[t & MASK] == gm) {
			uint32_t k = tie_key(t, b.st0, b.en0);
			key = k < key ? k : key;
		}
	}
	return key;
}

// ---- the scalar per-pair bookkeeping done by the group leader (:222-267) -----------------------------------
struct Leader {
	EzState ez;
	int st0_prev;
	int exit_slot; int32_t exit_H;          // most recently exited slot and its TRUE H (stale reads of :228)
	int32_t Hprev_true;                     // H[en0-1] of the previous diagonal
	int32_t Hen0_lazy, gmax;

	__device__ __forceinline__ void reset() { ez_reset(ez); st0_prev = 0; exit_slot = -2; exit_H = kNegInf; Hprev_true = kNegInf; Hen0_lazy = gmax = kNegInf; }

	// before the cells: remember/knock out slots that must not take part in the regular H update
	template <int MASK>
	__device__ __forceinline__ void pre(int32_t *H, const Band &b, int r, int qe)
	{
		Hprev_true = kNegInf;
		if (r == 0) return;
		if (b.st0 > st0_prev) {                              // slot st0-1 left the band: keep its TRUE H, drop it from the max
			int xs = b.st0 - 1;
			exit_slot = xs; exit_H = H[xs & MASK] - qe * (r - 1);
			H[xs & MASK] = kNegInf;
		}
		if (b.en0 > 0) {
			int ps = b.en0 - 1;
			Hprev_true = (ps == exit_slot) ? exit_H : H[ps & MASK] - qe * (r - 1);
			H[b.en0 & MASK] = kNegInf;                       // the regular update must not count for slot en0
		}
	}
	// after the cells: H[en0] (:228 / :259), diagonal max in the lazy domain; returns need_arg
	template <int MASK>
	__device__ __forceinline__ int mid(int32_t *H, const uint32_t *Us, const Band &b, int r, int qe, int32_t reduced_max,
	                                   uint32_t v0_r0, int zdrop)
	{
		gmax = reduced_max;
		if (r == 0) { Hen0_lazy = (int32_t)(v0_r0 >> 24) - 2 * qe; H[0] = Hen0_lazy; gmax = Hen0_lazy; }   // :259
		else if (b.en0 > 0) {
			Hen0_lazy = Hprev_true + (int32_t)(Us[b.en0 & MASK] >> 24) - qe + qe * r;                       // :228, true -> lazy
			H[b.en0 & MASK] = Hen0_lazy;
			gmax = gmax > Hen0_lazy ? gmax : Hen0_lazy;
		} else Hen0_lazy = H[0];                              // en0 == 0: regular update (:228 else-arm)
		int32_t maxH_true = gmax - qe * r;
		return (maxH_true > ez.max) || (zdrop >= 0);
	}
	__device__ __forceinline__ uint32_t en0_key(const Band &b, int r) const
	{
		return (Hen0_lazy == gmax && (r == 0 || b.en0 > 0)) ? 0u : 0xffffffffu;   // slot en0 wins every tie (:229-231)
	}
	// end scores, z-drop (:261-267); returns stop
	template <int MASK>
	__device__ __forceinline__ int fin(const int32_t *H, const Band &b, int r, int qe, int max_t, int qlen, int tlen,
	                                   int zdrop, int e)
	{
		const int R = qlen + tlen - 1;
		int32_t Hen0_true = Hen0_lazy - qe * r;
		int32_t maxH_true = gmax - qe * r;
		if (b.en0 == tlen - 1 && Hen0_true > ez.mte) { ez.mte = Hen0_true; ez.mte_q = r - b.en; }          // :261-262
		if (r - b.st0 == qlen - 1) {                                                                      // :263-264
			int32_t Hst0 = (b.st0 == b.en0) ? Hen0_true : H[b.st0 & MASK] - qe * r;
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

// =====================================================================================================
// narrow kernel: G <= 32 lanes per pair
// =====================================================================================================
template <int G, int S, bool kCigar, bool kRight>
__global__ void __launch_bounds__(128)
extz_dp_kernel(DpLaunch L)
{
	constexpr int NS = G * S;
	constexpr int MASK = NS - 1;
	constexpr int GROUPS_PER_BLOCK = 128 / G;
	static_assert(G <= 32 && (NS & MASK) == 0, "NS must be a power of two");
	static_assert(S % 4 == 0 && (16 % S == 0 || S % 16 == 0), "lane slots must tile 16-slot blocks");

	__shared__ int32_t sH[GROUPS_PER_BLOCK][NS];        // lazy H: H + (q+e)*r   (:222-259)
	__shared__ uint32_t sU[GROUPS_PER_BLOCK][NS];       // u' of the current diagonal (for H[en0], :228)
	__shared__ uint32_t sTable[kTableStride * kTableStride];

	for (int i = threadIdx.x; i < kTableStride * kTableStride; i += blockDim.x) sTable[i] = L.table[i];
	__syncthreads();

	const int lane_w = threadIdx.x & 31;
	const int gl = threadIdx.x % G;                                  // lane within group
	const int gidx = threadIdx.x / G;                                // group within block
	const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane_w & ~(G - 1)));
	const int pred_lane = (gl + G - 1) & (G - 1);                    // circular predecessor (relative to group)
	int32_t *H = sH[gidx];
	uint32_t *Us = sU[gidx];
	const Scoring sc = L.sc;
	const int qe = sc.qe;
	const bool generic = (sc.flag & kFlagGenericSc) != 0;

	for (;;) {
		int pi = 0;
		if (gl == 0) pi = atomicAdd(L.work_counter, 1);              // dynamic work queue, pairs sorted by descending work
		pi = __shfl_sync(gmask, pi, 0, G);
		if (pi >= L.n) break;
		const PairDesc pd = L.pairs[pi];
		const int qlen = pd.qlen, tlen = pd.tlen, w = pd.w;
		const int T = (tlen + 15) & ~15;
		const uint8_t *qseq = L.seq + pd.q_off;                      // qseq[j], j in [-kQPadL, qlen) readable
		const uint8_t *tseq = L.seq + pd.t_off;
		uint8_t *tbp = kCigar ? L.tb + pd.tb_off : nullptr;

		LaneState<S> ls;
		ls.t0 = gl * S;
		lane_load_slots<S>(ls, tseq, tlen, sc);
#pragma unroll
		for (int i = 0; i < S; ++i) H[gl * S + i] = kNegInf;         // :86-89
		__syncwarp(gmask);

		Leader ld; ld.reset();                                       // meaningful in the leader lane only
		int last_st = -1, n_diag = 0, zdropped_band = 0;
		const int R = qlen + tlen - 1;

		for (int r = 0; r < R; ++r) {
			Band b;
			if (!band_of(r, qlen, tlen, w, T, generic, b)) { zdropped_band = 1; break; }      // :110-113

			// carry from the circular predecessor: OLD x,v of its top slot (:28-35,117-121)
			uint32_t xin = __shfl_sync(gmask, ls.X[S - 1], pred_lane, G);
			uint32_t vin = __shfl_sync(gmask, ls.V[S - 1], pred_lane, G);
			lane_prepare<S, NS>(ls, b, r, qseq, tseq, tlen, sTable, sc);
			if (gl == 0) ld.pre<MASK>(H, b, r, qe);
			__syncwarp(gmask);

			int32_t lane_max = lane_cells<S, NS, kCigar, kRight>(ls, b, r, last_st, xin, vin, tbp, H, Us, sc);
			__syncwarp(gmask);

			int32_t red = __reduce_max_sync(gmask, lane_max);
			int need_arg = 0;
			if (gl == 0) need_arg = ld.mid<MASK>(H, Us, b, r, qe, red, ls.V[0], sc.zdrop);
			need_arg = __shfl_sync(gmask, need_arg, 0, G);
			int max_t = b.en0;
			if (need_arg) {
				int32_t gm = __shfl_sync(gmask, ld.gmax, 0, G);
				uint32_t key = lane_argmax_key<S, NS>(ls, b, H, gm);
				if (gl == 0) { uint32_t k0 = ld.en0_key(b, r); key = k0 < key ? k0 : key; }
				key = __reduce_min_sync(gmask, key);
				max_t = tie_key_slot(key, b.en0);
			}
			int stop = 0;
			if (gl == 0) stop = ld.fin<MASK>(H, b, r, qe, max_t, qlen, tlen, sc.zdrop, sc.e);
			stop = __shfl_sync(gmask, stop, 0, G);
			n_diag = r + 1;
			last_st = b.st;
			if (stop) break;
		}
		if (gl == 0) { if (zdropped_band) ld.ez.zdropped = 1; ld.store(&L.results[pi], n_diag); }
		__syncwarp(gmask);
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
	constexpr int MASK = NS - 1;
	constexpr int NW = G / 32;
	static_assert(G > 32 && G % 32 == 0 && (NS & MASK) == 0, "wide kernel: whole warps, NS power of two");
	static_assert(S % 4 == 0 && (16 % S == 0 || S % 16 == 0), "lane slots must tile 16-slot blocks");

	__shared__ int32_t H[NS];
	__shared__ uint32_t Us[NS];
	__shared__ uint32_t sTable[kTableStride * kTableStride];
	__shared__ uint32_t sCarryX[NW], sCarryV[NW];       // OLD x,v of every warp's top slot
	__shared__ int32_t sWarpMax[NW];
	__shared__ uint32_t sWarpKey[NW];
	__shared__ int sPair, sNeedArg, sStop;
	__shared__ int32_t sGmax;

	for (int i = threadIdx.x; i < kTableStride * kTableStride; i += blockDim.x) sTable[i] = L.table[i];
	__syncthreads();

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
		lane_load_slots<S>(ls, tseq, tlen, sc);
#pragma unroll
		for (int i = 0; i < S; ++i) H[gl * S + i] = kNegInf;
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
			if (gl == 0) ld.pre<MASK>(H, b, r, qe);
			__syncthreads();                                                                   // A
			// phase 2: cells
			if (lane == 0) { int pw = (wid + NW - 1) % NW; xin = sCarryX[pw]; vin = sCarryV[pw]; }
			lane_prepare<S, NS>(ls, b, r, qseq, tseq, tlen, sTable, sc);
			int32_t lane_max = lane_cells<S, NS, kCigar, kRight>(ls, b, r, last_st, xin, vin, tbp, H, Us, sc);
			int32_t wmax = __reduce_max_sync(0xffffffffu, lane_max);
			if (lane == 0) sWarpMax[wid] = wmax;
			__syncthreads();                                                                   // B
			// phase 3: leader
			if (gl == 0) {
				int32_t red = sWarpMax[0];
#pragma unroll
				for (int k = 1; k < NW; ++k) red = red > sWarpMax[k] ? red : sWarpMax[k];
				int need = ld.mid<MASK>(H, Us, b, r, qe, red, ls.V[0], sc.zdrop);
				sNeedArg = need; sGmax = ld.gmax;
				if (!need) sStop = ld.fin<MASK>(H, b, r, qe, b.en0, qlen, tlen, sc.zdrop, sc.e);
			}
			__syncthreads();                                                                   // C
			if (sNeedArg) {
				uint32_t key = lane_argmax_key<S, NS>(ls, b, H, sGmax);
				key = __reduce_min_sync(0xffffffffu, key);
				if (lane == 0) sWarpKey[wid] = key;
				__syncthreads();                                                               // D
				if (gl == 0) {
					uint32_t k = ld.en0_key(b, r);
#pragma unroll
					for (int j = 0; j < NW; ++j) k = sWarpKey[j] < k ? sWarpKey[j] : k;
					sStop = ld.fin<MASK>(H, b, r, qe, tie_key_slot(k, b.en0), qlen, tlen, sc.zdrop, sc.e);
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

} // namespace extz
