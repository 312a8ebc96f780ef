// extz_dp.cuh -- the anti-diagonal DP kernel (K1/K2 of SURVEY.md section 2.1) for sm_100a.
// See extz_core.cuh for the representation; this file is the warp-level choreography.
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

// One group of G lanes aligns one pair; each lane owns S consecutive slots (NS = G*S live slots).
// kCigar: write traceback codes.  kRight: KSW_EZ_RIGHT tie rules.
template <int G, int S, bool kCigar, bool kRight>
__global__ void __launch_bounds__(128)
extz_dp_kernel(DpLaunch L)
{
	constexpr int NS = G * S;
	constexpr int MASK = NS - 1;
	constexpr int GROUPS_PER_BLOCK = 128 / G;
	constexpr int NSUB = (S + 15) / 16;                 // 16-slot sub-blocks per lane (1 unless S == 32)
	constexpr int SUBW = S < 16 ? S : 16;
	static_assert((NS & MASK) == 0, "NS must be a power of two");
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
		// ---- fetch a pair (group leader) ----
		int pi = 0;
		if (gl == 0) pi = atomicAdd(L.work_counter, 1);
		pi = __shfl_sync(gmask, pi, 0, G);
		if (pi >= L.n) break;
		const PairDesc pd = L.pairs[pi];
		const int qlen = pd.qlen, tlen = pd.tlen, w = pd.w;
		const int T = (tlen + 15) & ~15;
		const uint8_t *qseq = L.seq + pd.q_off;                      // qseq[j], j in [-kQPadL, qlen) readable
		const uint8_t *tseq = L.seq + pd.t_off;
		uint8_t *tbp = kCigar ? L.tb + pd.tb_off : nullptr;

		// ---- per-lane state ----
		uint32_t U[S], V[S], X[S], Y[S], Z[S], TC[S];
		int t0 = gl * S;
#pragma unroll
		for (int i = 0; i < S; ++i) {
			U[i] = V[i] = X[i] = Y[i] = 0u; Z[i] = sc.s0_s;
			int t = t0 + i;
			TC[i] = (t < tlen ? ld_u8(tseq + t) : 0u) * (kTableStride * 4);   // byte offset of the table row
			H[gl * S + i] = kNegInf;
		}
		__syncwarp(gmask);

		EzState ez; ez_reset(ez);                                    // meaningful in the leader lane only
		int last_st = -1, st0_prev = 0;
		int exit_slot = -2; int32_t exit_H = kNegInf;                // most recently exited slot and its TRUE H
		int n_diag = 0;
		const int R = qlen + tlen - 1;

		for (int r = 0; r < R; ++r) {
			Band b;
			if (!band_of(r, qlen, tlen, w, T, generic, b)) { ez.zdropped = 1; break; }          // :110-113

			// ---- carry from the circular predecessor: OLD x,v of its top slot (:28-35,117-121) ----
			uint32_t xin = __shfl_sync(gmask, X[S - 1], pred_lane, G);
			uint32_t vin = __shfl_sync(gmask, V[S - 1], pred_lane, G);
			if (t0 == b.st) {
				if (b.st > 0) { if (!(b.st > last_st)) xin = vin = 0u; }
				else { xin = 0u; vin = r ? sc.q_s : 0u; }
			}
			// ---- slide the circular window: lanes whose slots all fell below st take slots +NS ----
			if (t0 + S - 1 < b.st) {
				t0 += NS;
#pragma unroll
				for (int i = 0; i < S; ++i) {
					U[i] = V[i] = X[i] = Y[i] = 0u; Z[i] = sc.s0_s;
					int t = t0 + i;
					TC[i] = (t < tlen ? ld_u8(tseq + t) : 0u) * (kTableStride * 4);
				}
			}
			// ---- top-row boundary (:122): y[r] = 0, u[r] = r ? q : 0 when the rounded range reaches slot r ----
			if (b.en >= r) {
				int k = r - t0;
#pragma unroll
				for (int i = 0; i < S; ++i) if (k == i) { Y[i] = 0u; U[i] = r ? sc.q_s : 0u; }
			}
			// ---- score fill (:124-138): slots st0..fe get a fresh s; others keep the stale one ----
			{
				const uint8_t *qp = qseq + (r - t0);                  // query[r - t] for slot t = t0 + i is qp[-i]
				const int lo = b.st0 - t0, hi = b.fe - t0;
#pragma unroll
				for (int i = 0; i < S; ++i) {
					if (i >= lo && i <= hi) {
						uint32_t qc = ld_u8(qp - i);
						Z[i] = *(const uint32_t *)((const char *)sTable + TC[i] + qc * 4);
					}
				}
			}
			// ---- H pre-phase (leader): stale reads and knock-outs (:228) ----
			int32_t Hprev_true = kNegInf;
			if (gl == 0 && r > 0) {
				if (b.st0 > st0_prev) {                                // slot st0-1 left the band: remember its TRUE H
					int xs = b.st0 - 1;
					exit_slot = xs; exit_H = H[xs & MASK] - qe * (r - 1);
					H[xs & MASK] = kNegInf;
				}
				if (b.en0 > 0) {
					int ps = b.en0 - 1;
					Hprev_true = (ps == exit_slot) ? exit_H : H[ps & MASK] - qe * (r - 1);
					H[b.en0 & MASK] = kNegInf;                         // the regular update below must not count for slot en0
				}
			}
			__syncwarp(gmask);

			// ---- the cells (:172-194 / :198-220 / :149-168) ----
			int32_t lane_max = kNegInf;
#pragma unroll
			for (int sb = NSUB - 1; sb >= 0; --sb) {
				const int tb0 = t0 + sb * 16;
				const bool active = (tb0 >= b.st) && (tb0 <= b.en);
				if (active) {
					uint32_t codes = 0;
#pragma unroll
					for (int ii = SUBW - 1; ii >= 0; --ii) {
						const int i = sb * 16 + ii;
						uint32_t xt1 = (i == 0) ? xin : X[i - 1];
						uint32_t vt1 = (i == 0) ? vin : V[i - 1];
						uint32_t c = cell<kRight, kCigar>(Z[i], xt1, vt1, U[i], V[i], X[i], Y[i], sc);
						if (kCigar) codes |= c << ((ii & 7) * 4);
						if (kCigar && (ii & 7) == 0) {
							// 8 codes = one 32-bit word; S == 4 packs 4 codes into 16 bits
							int c0 = (t0 + i) & MASK;
							uint8_t *dst = tbp + (int64_t)r * (NS >> 1) + (c0 >> 1);
							if (S >= 8) *(uint32_t *)dst = codes; else *(uint16_t *)dst = (uint16_t)codes;
							codes = 0;
						}
					}
					// dump u' (needed for H[en0]) and advance the lazy H row: H[t] += v[t]  (:233-239,255)
#pragma unroll
					for (int ii = 0; ii < SUBW; ii += 4) {
						const int i = sb * 16 + ii;
						const int c0 = (t0 + i) & MASK;
						*(uint4 *)&Us[c0] = make_uint4(U[i], U[i + 1], U[i + 2], U[i + 3]);
						int4 h = *(int4 *)&H[c0];
						h.x += (int32_t)(V[i] >> 24); h.y += (int32_t)(V[i + 1] >> 24);
						h.z += (int32_t)(V[i + 2] >> 24); h.w += (int32_t)(V[i + 3] >> 24);
						*(int4 *)&H[c0] = h;
						int32_t m01 = h.x > h.y ? h.x : h.y, m23 = h.z > h.w ? h.z : h.w;
						int32_t m = m01 > m23 ? m01 : m23;
						lane_max = lane_max > m ? lane_max : m;
					}
				}
			}
			__syncwarp(gmask);

			// ---- exact max, end scores, z-drop (:222-267) ----
			int32_t gmax = __reduce_max_sync(gmask, lane_max);      // lazy domain; excludes slot en0 when r>0 && en0>0
			int32_t Hen0_lazy = kNegInf;
			int need_arg = 0;
			// r == 0: H[0] = v8[0] - 2(q+e) (:259).  Slot 0 is lane 0 / register 0 at r == 0.
			if (r == 0) {
				if (gl == 0) { Hen0_lazy = (int32_t)(V[0] >> 24) - 2 * qe; H[0] = Hen0_lazy; gmax = Hen0_lazy; }
			} else if (gl == 0) {
				if (b.en0 > 0) {
					Hen0_lazy = Hprev_true + (int32_t)(Us[b.en0 & MASK] >> 24) - qe + qe * r;   // true -> lazy
					H[b.en0 & MASK] = Hen0_lazy;
					gmax = gmax > Hen0_lazy ? gmax : Hen0_lazy;
				} else {
					Hen0_lazy = H[0];                                  // en0 == 0: regular update (:228 else-arm)
				}
			}
			int32_t maxH_true = gmax - qe * r;                       // valid in the leader
			if (gl == 0) need_arg = (maxH_true > ez.max) || (sc.zdrop >= 0);
			need_arg = __shfl_sync(gmask, need_arg, 0, G);
			int max_t = b.en0;
			if (need_arg) {
				int32_t gm = __shfl_sync(gmask, gmax, 0, G);
				int32_t he = __shfl_sync(gmask, Hen0_lazy, 0, G);
				uint32_t key = 0xffffffffu;
				if (he == gm && (r == 0 || b.en0 > 0)) key = 0u;       // slot en0 wins every tie (:229-231)
#pragma unroll
				for (int i = 0; i < S; ++i) {
					int t = t0 + i;
					if (t >= b.st0 && t <= b.en0 && !(t == b.en0 && b.en0 > 0) && H[t & MASK] == gm) {
						uint32_t k = tie_key(t, b.st0, b.en0);
						key = k < key ? k : key;
					}
				}
				key = __reduce_min_sync(gmask, key);
				max_t = tie_key_slot(key, b.en0);
			}
			int stop = 0;
			if (gl == 0) {
				int32_t Hen0_true = Hen0_lazy - qe * r;
				if (b.en0 == tlen - 1 && Hen0_true > ez.mte) { ez.mte = Hen0_true; ez.mte_q = r - b.en; }     // :261-262
				if (r - b.st0 == qlen - 1) {                                                                  // :263-264
					int32_t Hst0 = (b.st0 == b.en0) ? Hen0_true : H[b.st0 & MASK] - qe * r;
					if (Hst0 > ez.mqe) { ez.mqe = Hst0; ez.mqe_t = b.st0; }
				}
				if (ez_apply_zdrop(ez, maxH_true, r, max_t, sc.zdrop, sc.e)) stop = 1;                        // :265
				else if (r == R - 1 && b.en0 == tlen - 1) ez.score = Hen0_true;                               // :266-267
			}
			stop = __shfl_sync(gmask, stop, 0, G);
			n_diag = r + 1;
			last_st = b.st; st0_prev = b.st0;
			if (stop) break;
		}
		if (gl == 0) {
			PairResult pr;
			pr.max = ez.max; pr.zdropped = ez.zdropped; pr.max_q = ez.max_q; pr.max_t = ez.max_t;
			pr.mqe = ez.mqe; pr.mqe_t = ez.mqe_t; pr.mte = ez.mte; pr.mte_q = ez.mte_q; pr.score = ez.score;
			pr.n_diag = n_diag; pr.n_cigar = 0; pr.cigar_off = 0;
			L.results[pi] = pr;
		}
		__syncwarp(gmask);
	}
}

} // namespace extz
