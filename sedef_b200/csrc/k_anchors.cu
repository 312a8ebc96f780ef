// k_anchors.cu -- SURVEY.md section 8 f3: SEDEF's generate_anchors (reference src/chain.cc:24-101) for MANY region pairs on the
// GPU.  What the reference computes, restated: on every diagonal of the (query, ref) dot plot, every maximal run of equal bases
// (case-insensitive, an N on either side ends a run) of length >= k yields ONE anchor -- it starts at the first position of the
// run whose k-mer is "eligible" (k-mers that occur >= 1000 times in the reference region are skipped, chain.cc:59-61) and
// extends to the end of the run (the reference extends forwards only and remembers the end per diagonal in `slide[]`).  With the
// same-chromosome rule (chain.cc:66-68) whole diagonals within k of the main diagonal are dropped.  has_u = "any upper-case
// base in the match" (a bool the reference accumulates with +=, chain.cc:74,84).
//
// How: the reference's hash-map-of-lists over the ref k-mers (the churn SURVEY 3.2 measures at 28-33 % of fast_align) becomes
// one open-addressing table per region in HBM (key, occurrence count, head of a position list threaded through `next[]`), built
// by one kernel; a second kernel looks every query k-mer up, and for every (q, r) hit decides locally whether it STARTS an
// anchor: the hit one step down the diagonal must not belong to the same run with an eligible k-mer.  Starts extend forwards
// and emit.  Two passes (count, fill) size the output exactly.  Anchors come back sorted by (q, r) like the reference emits them.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/ksw2_b200.h"

namespace {

constexpr uint32_t kEmpty = 0xffffffffu;
constexpr int kMaxOcc = 1000;                         // chain.cc:60: `it->second.size() >= 1000` -> skip

struct AnchorLaunch {
	const uint8_t *q, *r;                             // original-case bytes, flat
	const int64_t *qoff, *roff, *rbase;               // rbase[i]: first entry of region i in next[]
	const int *qlen, *rlen;
	const uint8_t *same_chr; const int64_t *oqs, *ors;
	uint32_t *keys; int32_t *count, *head, *next;
	const int64_t *tbl_off; const int *tbl_cap;       // table of region i: [tbl_off[i], tbl_off[i] + tbl_cap[i]), capacity a power of two
	unsigned long long *n_anchors;                    // per region
	sedef_anchor_t *out; const int64_t *out_off;      // pass 1
	int k, n, pass;
};

__device__ __forceinline__ int up(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }
__device__ __forceinline__ uint32_t hash_dna(int c)   // src/common.h:50-60: ACGT/acgt -> 0..3, everything else -> 0
{
	const int u = up(c);
	return u == 'C' ? 1u : u == 'G' ? 2u : u == 'T' ? 3u : 0u;
}
// k-mer starting at s: key, or kEmpty when it holds an N (chain.cc:31-38: last_n)
__device__ __forceinline__ uint32_t kmer_key(const uint8_t *s, int k)
{
	uint32_t h = 0; bool bad = false;
	for (int i = 0; i < k; ++i) { const int c = s[i]; bad |= up(c) == 'N'; h = (h << 2) | hash_dna(c); }
	return bad ? kEmpty : h;
}
__device__ __forceinline__ uint32_t slot_of(uint32_t key, int cap) { return (key * 0x9E3779B1u) & (uint32_t)(cap - 1); }

__global__ void __launch_bounds__(256) anchor_build_kernel(AnchorLaunch L)
{
	const int reg = blockIdx.x;
	const uint8_t *ref = L.r + L.roff[reg];
	const int rlen = L.rlen[reg], cap = L.tbl_cap[reg];
	uint32_t *keys = L.keys + L.tbl_off[reg];
	int32_t *count = L.count + L.tbl_off[reg], *head = L.head + L.tbl_off[reg], *next = L.next + L.rbase[reg];
	for (int p = threadIdx.x; p + L.k <= rlen; p += blockDim.x) {
		const uint32_t key = kmer_key(ref + p, L.k);
		if (key == kEmpty) continue;
		uint32_t s = slot_of(key, cap);
		for (;;) {
			const uint32_t old = atomicCAS(&keys[s], kEmpty, key);
			if (old == kEmpty || old == key) break;
			s = (s + 1) & (uint32_t)(cap - 1);
		}
		atomicAdd(&count[s], 1);
		next[p] = atomicExch(&head[s], p);
	}
}

// slot of `key` in the region's table, or -1
__device__ __forceinline__ int lookup(const uint32_t *keys, int cap, uint32_t key)
{
	uint32_t s = slot_of(key, cap);
	for (;;) {
		const uint32_t kk = keys[s];
		if (kk == key) return (int)s;
		if (kk == kEmpty) return -1;
		s = (s + 1) & (uint32_t)(cap - 1);
	}
}

__global__ void __launch_bounds__(256) anchor_find_kernel(AnchorLaunch L)
{
	const int reg = blockIdx.x;
	const uint8_t *qry = L.q + L.qoff[reg], *ref = L.r + L.roff[reg];
	const int qlen = L.qlen[reg], rlen = L.rlen[reg], cap = L.tbl_cap[reg], k = L.k;
	const uint32_t *keys = L.keys + L.tbl_off[reg];
	const int32_t *count = L.count + L.tbl_off[reg], *head = L.head + L.tbl_off[reg], *next = L.next + L.rbase[reg];
	const bool same = L.same_chr && L.same_chr[reg];
	const int64_t oqs = L.oqs ? L.oqs[reg] : 0, ors = L.ors ? L.ors[reg] : 0;
	auto eq = [&](int qi, int ri) {                   // bases equal and neither an N: the pair of positions continues a run
		const int a = up(qry[qi]), b = up(ref[ri]);
		return a == b && a != 'N';
	};
	auto capped = [&](int qi) {                       // k-mer at query position qi occurs >= 1000 times in the reference region
		const uint32_t key = kmer_key(qry + qi, k);
		if (key == kEmpty) return true;
		const int s = lookup(keys, cap, key);
		return s < 0 || count[s] >= kMaxOcc;
	};
	for (int q = threadIdx.x; q + k <= qlen; q += blockDim.x) {
		const uint32_t key = kmer_key(qry + q, k);
		if (key == kEmpty) continue;
		const int s = lookup(keys, cap, key);
		if (s < 0 || count[s] >= kMaxOcc) continue;
		for (int r = head[s]; r >= 0; r = next[r]) {
			if (same) {                                                                   // chain.cc:66-68
				const int64_t d = ors + r - (oqs + q);
				if ((d < 0 ? -d : d) <= k) continue;
			}
			// equal hashes are equal bases for A/C/G/T; any other letter hashes like 'A' (hash_dna) and must really match
			bool ok = true;
			for (int i = 0; i < k && ok; ++i) ok = eq(q + i, r + i);
			if (!ok) continue;
			// does an EARLIER position of the same run carry an eligible k-mer?  then that one started the anchor (slide[])
			bool start = true;
			for (int j = 1; q - j >= 0 && r - j >= 0 && eq(q - j, r - j); ++j)
				if (!capped(q - j)) { start = false; break; }
			if (!start) continue;
			int len = k; bool has_u = false;
			for (int i = 0; i < k; ++i) has_u |= (qry[q + i] >= 'A' && qry[q + i] <= 'Z') || (ref[r + i] >= 'A' && ref[r + i] <= 'Z');
			while (q + len < qlen && r + len < rlen && eq(q + len, r + len)) {
				has_u |= (qry[q + len] >= 'A' && qry[q + len] <= 'Z') || (ref[r + len] >= 'A' && ref[r + len] <= 'Z');
				++len;
			}
			const unsigned long long at = atomicAdd(&L.n_anchors[reg], 1ull);
			if (L.pass == 1) L.out[L.out_off[reg] + (int64_t)at] = sedef_anchor_t{q, r, len, has_u ? 1 : 0};
		}
	}
}

// scratch buffers persist between calls (cudaMalloc / cudaFree of tens of MB cost more than the kernels): slot k of the
// per-process scratch grows on demand; one call at a time (mutex)
struct Scratch {
	void *p[12] = {nullptr}; size_t cap[12] = {0}; int dev = -1;
	void *get(int k, size_t bytes)
	{
		int cur = 0; cudaGetDevice(&cur);
		if (cur != dev) { for (int i = 0; i < 12; ++i) { if (p[i]) cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; } dev = cur; }
		if (bytes <= cap[k]) return p[k];
		if (p[k]) cudaFree(p[k]);
		const size_t want = bytes + bytes / 4 + 256;
		if (cudaMalloc(&p[k], want) != cudaSuccess) { cudaGetLastError(); p[k] = nullptr; cap[k] = 0; return nullptr; }
		cap[k] = want; return p[k];
	}
};
Scratch g_scratch;
std::mutex g_scratch_mu;
struct Dev {
	void *p = nullptr; int slot = 0;
	bool alloc(size_t bytes) { p = g_scratch.get(slot, bytes ? bytes : 16); return p != nullptr; }
};

} // namespace

static inline double wall_ms_() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

extern "C" int sedef_anchors_batch(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                                   const int *rlen, const int64_t *roff, const uint8_t *rbuf, int kmer_size,
                                   const uint8_t *same_chr, const int64_t *orig_query_start, const int64_t *orig_ref_start,
                                   sedef_anchor_t **anchors_out, int64_t *anchor_off)
{
	if (n < 0 || !anchors_out || !anchor_off || (n > 0 && (!qlen || !qoff || !qbuf || !rlen || !roff || !rbuf))) return KSW_B200_ERR_ARG;
	if (kmer_size < 1 || kmer_size > 15) return KSW_B200_ERR_ARG;
	*anchors_out = nullptr;
	for (int i = 0; i <= n; ++i) anchor_off[i] = 0;
	if (n == 0) return KSW_B200_OK;
	int ndev = ksw_b200_num_devices();
	if (ndev <= 0 && (ndev = ksw_b200_init(0, 0)) <= 0) return ndev < 0 ? ndev : KSW_B200_ERR_NO_DEVICE;
	int64_t qtot = 0, rtot = 0, ttot = 0;
	std::vector<int64_t> rbase(n), tbl_off(n);
	std::vector<int> tbl_cap(n);
	for (int i = 0; i < n; ++i) {
		if (qlen[i] < 0 || rlen[i] < 0) return KSW_B200_ERR_ARG;
		qtot = std::max<int64_t>(qtot, qoff[i] + qlen[i]);
		rtot = std::max<int64_t>(rtot, roff[i] + rlen[i]);
		rbase[i] = i ? rbase[i - 1] + rlen[i - 1] : 0;
		int cap = 64;
		while (cap < 2 * rlen[i]) cap <<= 1;
		tbl_cap[i] = cap; tbl_off[i] = ttot; ttot += cap;
	}
	const int64_t ntot = rbase[n - 1] + rlen[n - 1];
	std::lock_guard<std::mutex> lk(g_scratch_mu);
	const double t_begin = wall_ms_();
	Dev dq, dr, dmeta, dkeys, dcount, dhead, dnext, dna, dout;
	dq.slot = 0; dr.slot = 1; dmeta.slot = 2; dkeys.slot = 3; dcount.slot = 4; dhead.slot = 5; dnext.slot = 6; dna.slot = 7; dout.slot = 8;
	const size_t meta_bytes = (size_t)n * (8 * 6 + 4 * 3 + 1) + 256;
	if (!dq.alloc(qtot + 16) || !dr.alloc(rtot + 16) || !dmeta.alloc(meta_bytes) || !dkeys.alloc(ttot * 4) || !dcount.alloc(ttot * 4) ||
	    !dhead.alloc(ttot * 4) || !dnext.alloc(ntot * 4) || !dna.alloc((size_t)n * 8)) { cudaGetLastError(); return KSW_B200_ERR_NOMEM; }
	// meta layout: int64 qoff, roff, rbase, tbl_off, oqs, ors, out_off ... built below
	char *mp = (char *)dmeta.p;
	int64_t *m_qoff = (int64_t *)mp, *m_roff = m_qoff + n, *m_rbase = m_roff + n, *m_toff = m_rbase + n, *m_oqs = m_toff + n, *m_ors = m_oqs + n;
	int *m_qlen = (int *)(m_ors + n), *m_rlen = m_qlen + n, *m_cap = m_rlen + n;
	uint8_t *m_same = (uint8_t *)(m_cap + n);
	std::vector<int64_t> zeros(n, 0);
	std::vector<uint8_t> same(n, 0);
	if (same_chr) same.assign(same_chr, same_chr + n);
	bool ok = cudaMemcpy(dq.p, qbuf, qtot, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(dr.p, rbuf, rtot, cudaMemcpyHostToDevice) == cudaSuccess &&
	          cudaMemcpy(m_qoff, qoff, n * 8, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(m_roff, roff, n * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
	          cudaMemcpy(m_rbase, rbase.data(), n * 8, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(m_toff, tbl_off.data(), n * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
	          cudaMemcpy(m_oqs, orig_query_start ? orig_query_start : zeros.data(), n * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
	          cudaMemcpy(m_ors, orig_ref_start ? orig_ref_start : zeros.data(), n * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
	          cudaMemcpy(m_qlen, qlen, n * 4, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(m_rlen, rlen, n * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
	          cudaMemcpy(m_cap, tbl_cap.data(), n * 4, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(m_same, same.data(), n, cudaMemcpyHostToDevice) == cudaSuccess &&
	          cudaMemset(dkeys.p, 0xff, ttot * 4) == cudaSuccess && cudaMemset(dcount.p, 0, ttot * 4) == cudaSuccess &&
	          cudaMemset(dhead.p, 0xff, ttot * 4) == cudaSuccess && cudaMemset(dna.p, 0, (size_t)n * 8) == cudaSuccess;
	if (!ok) { cudaGetLastError(); return KSW_B200_ERR_CUDA; }
	cudaDeviceSynchronize();
	const double t_up = wall_ms_();
	AnchorLaunch L{};
	L.q = (const uint8_t *)dq.p; L.r = (const uint8_t *)dr.p; L.qoff = m_qoff; L.roff = m_roff; L.rbase = m_rbase; L.qlen = m_qlen; L.rlen = m_rlen;
	L.same_chr = m_same; L.oqs = m_oqs; L.ors = m_ors;
	L.keys = (uint32_t *)dkeys.p; L.count = (int32_t *)dcount.p; L.head = (int32_t *)dhead.p; L.next = (int32_t *)dnext.p;
	L.tbl_off = m_toff; L.tbl_cap = m_cap; L.n_anchors = (unsigned long long *)dna.p; L.k = kmer_size; L.n = n; L.pass = 0;
	anchor_build_kernel<<<n, 256>>>(L);
	anchor_find_kernel<<<n, 256>>>(L);                                            // pass 0: count
	std::vector<unsigned long long> cnt(n);
	if (cudaMemcpy(cnt.data(), dna.p, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return KSW_B200_ERR_CUDA; }
	const double t_count = wall_ms_();
	for (int i = 0; i < n; ++i) anchor_off[i + 1] = anchor_off[i] + (int64_t)cnt[i];
	const int64_t total = anchor_off[n];
	sedef_anchor_t *host = (sedef_anchor_t *)malloc(std::max<int64_t>(1, total) * sizeof(sedef_anchor_t));
	if (!host) return KSW_B200_ERR_NOMEM;
	Dev doff; doff.slot = 9;
	if (!dout.alloc((size_t)std::max<int64_t>(1, total) * sizeof(sedef_anchor_t)) || !doff.alloc((size_t)(n + 1) * 8)) { free(host); cudaGetLastError(); return KSW_B200_ERR_NOMEM; }
	ok = cudaMemcpy(doff.p, anchor_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemset(dna.p, 0, (size_t)n * 8) == cudaSuccess;
	L.out = (sedef_anchor_t *)dout.p; L.out_off = (const int64_t *)doff.p; L.pass = 1;
	if (ok) anchor_find_kernel<<<n, 256>>>(L);                                    // pass 1: fill
	ok = ok && cudaMemcpy(host, dout.p, (size_t)total * sizeof(sedef_anchor_t), cudaMemcpyDeviceToHost) == cudaSuccess;
	if (!ok) { free(host); cudaGetLastError(); return KSW_B200_ERR_CUDA; }
	const double t_fill = wall_ms_();
	// the reference emits anchors by query position, and for one query position by reference position (chain.cc:50-64)
#pragma omp parallel for schedule(dynamic, 8)
	for (int i = 0; i < n; ++i)
		std::sort(host + anchor_off[i], host + anchor_off[i + 1], [](const sedef_anchor_t &a, const sedef_anchor_t &b) { return a.q != b.q ? a.q < b.q : a.r < b.r; });
	if (getenv("SEDEF_B200_TRACE"))
		fprintf(stderr, "[anchors] %d regions, %.1f MB up: upload+clear %.1f ms, build+count %.1f ms, fill+download %.1f ms, sort %.1f ms, %lld anchors\n",
		        n, (qtot + rtot) / 1e6, t_up - t_begin, t_count - t_up, t_fill - t_count, wall_ms_() - t_fill, (long long)total);
	*anchors_out = host;
	return KSW_B200_OK;
}
