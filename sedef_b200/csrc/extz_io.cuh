// extz_io.cuh -- the small kernels either side of the DP: input preparation and result gather.
//
//   encode_kernel         original-case ASCII -> align_dna codes (src/common.h:58-70,91: ACGT/acgt -> 0..3, everything else
//                         -> 4), elementwise over the whole sequence arena.  With it a caller uploads ONE byte per base (the
//                         bytes SEDEF's Alignment(fa, fb) holds, src/align.cc:76-87) instead of codes + original case.
//   check_symbols_kernel  caller-supplied codes must be < limit (8, or m with KSW_EZ_GENERIC_SC); sets *flag otherwise (the DP
//                         kernels mask symbols to 0..7 meanwhile, so a bad batch is reported, never out of bounds).
//   gather_kernel         PairResult (device order) -> ksw_extz_t records in the CALLER's order with host CIGAR pointers
//                         already in place, so that the host only copies the array (no per-pair work, no malloc).
//   fill_reset_kernel     ksw_reset_extz (extern/ksw2.h:153-159) for every record (pairs the DP never sees: empty ones).
#pragma once
#include <cuda_runtime.h>
#include "launch_structs.h"

namespace extz {

__device__ __forceinline__ uint32_t align_dna4(uint32_t w)
{
	// four bytes at once: fold case (& 0xDF), then A C G T -> 0 1 2 3, everything else -> 4
	uint32_t out = 0;
#pragma unroll
	for (int b = 0; b < 4; ++b) {
		const uint32_t c = (w >> (8 * b)) & 0xdfu;
		const uint32_t v = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
		out |= v << (8 * b);
	}
	return out;
}
// NOTE on "& 0xDF": it maps 'a'..'z' to 'A'..'Z' and leaves 'A'..'Z' alone, but it also maps e.g. 0x21 '!' -> 0x01; none of
// those images is one of A/C/G/T except from their own lower/upper-case letters: c & 0xDF == 'A' (0x41) iff c is 0x41 or
// 0x61.  Bytes >= 0x80 index past the reference's 128-entry table (undefined there); here they encode as 4.

__global__ void __launch_bounds__(256) encode_kernel(const uint8_t *__restrict__ raw, uint8_t *__restrict__ codes, size_t nbytes)
{
	// nbytes is a multiple of 16 (arena sizes are), both pointers are 256-byte aligned
	const size_t nvec = nbytes >> 4;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
		const uint4 v = __ldg((const uint4 *)raw + i);
		((uint4 *)codes)[i] = make_uint4(align_dna4(v.x), align_dna4(v.y), align_dna4(v.z), align_dna4(v.w));
	}
}

// one warp per pair (grid-stride): only bytes a pair references are looked at -- a dense upload may carry bytes of the
// caller's buffers that belong to no pair
__global__ void __launch_bounds__(256) check_symbols_kernel(const PairDesc *__restrict__ pairs, int n, const uint8_t *__restrict__ codes,
                                                            uint32_t limit, int *flag)
{
	const int lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	uint32_t bad = 0;
	for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
		const PairDesc pd = pairs[k];
		const uint8_t *q = codes + pd.q_off, *t = codes + pd.t_off;
		for (int i = lane; i < pd.qlen; i += 32) bad |= (uint32_t)(__ldg(q + i) >= limit);
		for (int i = lane; i < pd.tlen; i += 32) bad |= (uint32_t)(__ldg(t + i) >= limit);
	}
	if (__any_sync(0xffffffffu, bad) && lane == 0) *flag = 1;
}

__device__ __forceinline__ void store_reset_record(uint64_t *rec)
{
	// ksw_extz_t: {max:31, zdropped:1 | max_q} {max_t | mqe} {mqe_t | mte} {mte_q | score} cigar m_cigar n_cigar
	const uint64_t neg = (uint32_t)kNegInf, m1 = 0xffffffffull;
	rec[0] = 0ull | (m1 << 32);               // max = 0, zdropped = 0, max_q = -1
	rec[1] = m1 | (neg << 32);                // max_t = -1, mqe = NEG_INF
	rec[2] = m1 | (neg << 32);                // mqe_t = -1, mte = NEG_INF
	rec[3] = m1 | (neg << 32);                // mte_q = -1, score = NEG_INF
	rec[4] = 0; rec[5] = 0; rec[6] = 0;       // cigar = NULL, m_cigar = n_cigar = 0
}

__global__ void __launch_bounds__(256) fill_reset_kernel(uint64_t *ez_out, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) store_reset_record(ez_out + 7 * (size_t)i);
}

__global__ void __launch_bounds__(256) gather_kernel(GatherLaunch L)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= L.n) return;
	const PairResult r = L.results[k];
	uint64_t *rec = L.ez_out + 7 * (size_t)(L.by_orig ? L.pairs[k].orig : k);
	auto lo32 = [](int32_t v) { return (uint64_t)(uint32_t)v; };
	rec[0] = (lo32(r.max) & 0x7fffffffull) | ((uint64_t)(r.zdropped ? 1u : 0u) << 31) | (lo32(r.max_q) << 32);
	rec[1] = lo32(r.max_t) | (lo32(r.mqe) << 32);
	rec[2] = lo32(r.mqe_t) | (lo32(r.mte) << 32);
	rec[3] = lo32(r.mte_q) | (lo32(r.score) << 32);
	if (L.with_cigar && r.n_cigar > 0) {
		int64_t cap = 4;                                  // capacity ksw_push_cigar would have grown to (extern/ksw2.h:101-105)
		while (cap < r.n_cigar) cap <<= 1;
		rec[4] = L.host_cigar_base + 4ull * (uint64_t)(uint32_t)r.cigar_off;
		rec[5] = (uint64_t)cap;
		rec[6] = (uint64_t)r.n_cigar;
	} else { rec[4] = 0; rec[5] = 0; rec[6] = 0; }
}

} // namespace extz
