// sedef_align.cc -- see include/sedef_align.hpp.  Host C++ on top of the C ABI; compiled into libsedef_b200.so.
#include "../../../include/sedef_align.hpp"
#include <algorithm>
#include <functional>
#include <memory>
#include <mutex>
#include <tuple>
#include <unordered_map>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace sedef_b200 {

static const int kMaxKswSeqLen = 60 * 1000;            // Globals::Align::MAX_KSW_SEQ_LEN = 60 * KB, and KB is 1000 (src/globals.h:18,54)

double Alignment::gap_error() const
{
	return 100.0 * stats.gap_bases / double(stats.matches + stats.gap_bases + stats.mismatches);   // pct(), src/common.h:99
}
double Alignment::mismatch_error() const
{
	return 100.0 * stats.mismatches / double(stats.matches + stats.gap_bases + stats.mismatches);
}
std::string Alignment::cigar_string() const
{
	std::string s;
	for (auto &c : cigar)
		if (c.second) { s += std::to_string(c.second); s += c.first; }               // zero-length runs are not printed
	return s;
}
sd_stats_fp_t Alignment::bedpe_fp() const
{
	sd_stats_fp_t o;
	sd_stats_derive_fp(&stats, &o);
	return o;
}

static inline double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// ksw_extz2 pairs and in-band cells of the batched calls issued on this thread (every SEDEF call is unbanded: cells = qlen x tlen)
static thread_local long long tl_ksw_pairs = 0, tl_ksw_cells = 0;
static inline bool region_trace() { static const bool on = getenv("SEDEF_B200_TRACE") != nullptr; return on; }   // developer aid: phase times on stderr

static void add_stats(sd_stats_t &d, const sd_stats_t &s)
{
	int32_t *dp = reinterpret_cast<int32_t *>(&d);
	const int32_t *sp = reinterpret_cast<const int32_t *>(&s);
	for (size_t k = 0; k < sizeof(sd_stats_t) / sizeof(int32_t); ++k) dp[k] += sp[k];
}

struct StrPair { const std::string *a, *b; };
static std::vector<sd_stats_t> stats_of(const std::vector<StrPair> &pairs, const std::vector<const std::deque<std::pair<char, int>> *> &cigars);

int max_ksw_seq_len() { return kMaxKswSeqLen; }

// align_helper's chunk loop (src/align.cc:46-53): for (SP = 0; SP < min(|t|, |q|); SP += MAX) one ksw call on
// (q + SP, min(MAX, |q| - SP)) x (t + SP, min(MAX, |t| - SP))
struct Chunk { size_t sp; int la, lb; };
static inline void chunk_plan(size_t alen, size_t blen, std::vector<Chunk> &out)
{
	const size_t n = std::min(alen, blen);
	for (size_t sp = 0; sp < n; sp += kMaxKswSeqLen)
		out.push_back({sp, (int)std::min<size_t>(kMaxKswSeqLen, alen - sp), (int)std::min<size_t>(kMaxKswSeqLen, blen - sp)});
}

// device-side results of the trim scans of one request: valid only for requests aligned in a single ksw call
struct TrimScan { bool valid = false; int front = -1, back = -1; };
static std::vector<Alignment> align_batch_impl(const std::vector<std::pair<std::string, std::string>> &pairs, const AlignParams &p,
                                               std::vector<TrimScan> *trims);
std::vector<Alignment> align_batch(const std::vector<std::pair<std::string, std::string>> &pairs, const AlignParams &p)
{
	return align_batch_impl(pairs, p, nullptr);
}
static std::vector<Alignment> align_batch_impl(const std::vector<std::pair<std::string, std::string>> &pairs, const AlignParams &p,
                                               std::vector<TrimScan> *trims)
{
	// align_helper's matrix (src/align.cc:41-44)
	const int8_t a = (int8_t)p.match, b = p.mismatch < 0 ? (int8_t)p.mismatch : (int8_t)(-p.mismatch);
	const int8_t mat[25] = {a, b, b, b, 0, b, a, b, b, 0, b, b, a, b, 0, b, b, b, a, 0, 0, 0, 0, 0, 0};
	// one flat buffer per side holding the ORIGINAL-CASE bytes (align_dna runs on the device); pairs longer than
	// MAX_KSW_SEQ_LEN are chunked with the same offset on both strings (src/align.cc:46-53)
	std::vector<int> ql, tl, owner;
	std::vector<int64_t> qo, to;
	std::vector<Chunk> chunks;
	size_t qtot = 0, ttot = 0;
	for (size_t i = 0; i < pairs.size(); ++i) {
		const size_t c0 = chunks.size();
		chunk_plan(pairs[i].first.size(), pairs[i].second.size(), chunks);
		for (size_t c = c0; c < chunks.size(); ++c) {
			owner.push_back((int)i);
			ql.push_back(chunks[c].la); tl.push_back(chunks[c].lb);
			qo.push_back((int64_t)qtot); to.push_back((int64_t)ttot);
			qtot += chunks[c].la; ttot += chunks[c].lb;
		}
	}
	std::vector<uint8_t> qraw(qtot + 1), traw(ttot + 1);                                   // never pass null buffers
#pragma omp parallel for schedule(static) if (chunks.size() >= 512)
	for (long k = 0; k < (long)chunks.size(); ++k) {
		memcpy(&qraw[qo[k]], pairs[owner[k]].first.data() + chunks[k].sp, ql[k]);
		memcpy(&traw[to[k]], pairs[owner[k]].second.data() + chunks[k].sp, tl[k]);
	}
	const int n = (int)owner.size();
	{
		long long cells = 0;
		for (int k = 0; k < n; ++k) cells += (long long)ql[k] * tl[k];
		tl_ksw_pairs += n; tl_ksw_cells += cells;
	}
	ksw_b200_result_t *res = nullptr;
	const double t_call = wall_ms();
	int rc = ksw_extz2_batch_arena(n, ql.data(), qo.data(), nullptr, tl.data(), to.data(), nullptr, 5, mat,
	                               (int8_t)p.gap_open, (int8_t)p.gap_extend, p.bandwidth, -1, 0, 1, qraw.data(), traw.data(), &res);
	if (rc) throw std::runtime_error(std::string("ksw_extz2_batch_arena: ") + ksw_b200_strerror(rc) + " -- " + ksw_b200_last_error());
	if (region_trace()) fprintf(stderr, "[regions]     ksw_extz2_batch_arena: %d pairs, %.1f ms\n", n, wall_ms() - t_call);
	const ksw_extz_t *ez = ksw_b200_result_ez(res);
	const sd_stats_t *st = ksw_b200_result_stats(res);
	if (trims) {
		const int32_t *tr = ksw_b200_result_trims(res);
		trims->assign(pairs.size(), TrimScan());
		for (int k = 0; k < n; ++k) {
			const bool single = (k == 0 || owner[k - 1] != owner[k]) && (k + 1 == n || owner[k + 1] != owner[k]);
			if (single && tr) (*trims)[owner[k]] = TrimScan{true, tr[2 * k], tr[2 * k + 1]};
		}
	}
	std::vector<Alignment> out(pairs.size());
	// The statistics of a chunked pair are the sum of its chunks' statistics as long as every chunk but the last one is
	// aligned end to end: the reference walks the CONCATENATED cigar from (0, 0) (populate_nice_alignment), so a chunk cut
	// short by a band break (user-set bandwidth) shifts every later column.  Those pairs are re-walked below.
	std::vector<char> rewalk(pairs.size(), 0);
	std::vector<int> first(pairs.size() + 1, n);                                           // first chunk of every pair
	for (int k = n - 1; k >= 0; --k) first[owner[k]] = k;
	for (long i = (long)pairs.size() - 1; i >= 0; --i) if (first[i] == n || first[i] > first[i + 1]) first[i] = first[i + 1];   // pairs without chunks
#pragma omp parallel for schedule(static) if (pairs.size() >= 512)
	for (long i = 0; i < (long)pairs.size(); ++i) {
		Alignment &al = out[i];
		al.a = pairs[i].first; al.b = pairs[i].second;
		for (int k = first[i]; k < first[i + 1]; ++k) {
			for (int64_t c = 0; c < ez[k].n_cigar; ++c) {
				const int idx = ez[k].cigar[c] & 0xf, len = (int)(ez[k].cigar[c] >> 4);
				if (idx < 3) al.cigar.push_back({"MDI"[idx], len});                          // src/align.cc:58-63
			}
			add_stats(al.stats, st[k]);
			if (ez[k].zdropped && k + 1 < first[i + 1]) rewalk[i] = 1;
		}
	}
	ksw_b200_result_free(res);
	std::vector<StrPair> rp;
	std::vector<const std::deque<std::pair<char, int>> *> rc_;
	std::vector<size_t> ri;
	for (size_t i = 0; i < pairs.size(); ++i) if (rewalk[i]) { rp.push_back({&pairs[i].first, &pairs[i].second}); rc_.push_back(&out[i].cigar); ri.push_back(i); }
	if (!ri.empty()) {
		std::vector<sd_stats_t> rs = stats_of(rp, rc_);
		for (size_t k = 0; k < ri.size(); ++k) out[ri[k]].stats = rs[k];
	}
	return out;
}

// statistics of alignments given as SEDEF-alphabet run lists (populate_nice_alignment, src/align.cc:274-315).
// Zero-length runs are kept: the reference counts every non-M run in `gaps`, also the ('\0', 0) run that
// cigar_from_alignment leaves for an alignment trimmed to nothing (src/align.cc:300-305,479-501).
static std::vector<sd_stats_t> stats_of(const std::vector<StrPair> &pairs, const std::vector<const std::deque<std::pair<char, int>> *> &cigars)
{
	const int n = (int)pairs.size();
	const double t_pack = wall_ms();
	std::vector<int64_t> coff(n), cn(n), ao(n), bo(n);
	std::vector<int> al(n), bl(n);
	int64_t ctot = 0, atot = 0, btot = 0;
	for (int i = 0; i < n; ++i) {
		coff[i] = ctot; cn[i] = (int64_t)cigars[i]->size(); ctot += cn[i];
		ao[i] = atot; bo[i] = btot;
		al[i] = (int)pairs[i].a->size(); bl[i] = (int)pairs[i].b->size();
		atot += al[i]; btot += bl[i];
	}
	// (uninitialised: the parallel copies below are the first touch of the pages)
	std::unique_ptr<uint32_t[]> cbuf(new uint32_t[ctot + 1]);
	std::unique_ptr<uint8_t[]> abuf(new uint8_t[atot + 1]), bbuf(new uint8_t[btot + 1]);
#pragma omp parallel for schedule(dynamic, 16) if (n >= 64)
	for (int i = 0; i < n; ++i) {
		int64_t c = coff[i];
		for (auto &run : *cigars[i]) {
			const char ch = run.first;
			uint32_t op = ch == 'M' ? 0u : (ch == 'D' ? 1u : (ch == 'I' ? 2u : 3u));   // SEDEF 'D' = a only = ksw I; 3 = any other letter
			cbuf[c++] = (uint32_t)run.second << 4 | op;
		}
		memcpy(&abuf[ao[i]], pairs[i].a->data(), al[i]);
		memcpy(&bbuf[bo[i]], pairs[i].b->data(), bl[i]);
	}
	std::vector<sd_stats_t> st(n);
	std::vector<int> status(n);
	const double t_call = wall_ms();
	int rc = sd_stats_from_cigar_batch_flat(n, coff.data(), cn.data(), cbuf.get(), al.data(), ao.data(), abuf.get(),
	                                       bl.data(), bo.data(), bbuf.get(), st.data(), status.data());
	if (region_trace()) fprintf(stderr, "[regions]     sd_stats_from_cigar: %d alignments, %.1f MB, pack %.1f ms, call %.1f ms\n", n, (atot + btot) / 1e6,
	                            t_call - t_pack, wall_ms() - t_call);
	if (rc) throw std::runtime_error(std::string("sd_stats_from_cigar_batch_flat: ") + ksw_b200_strerror(rc) + " -- " + ksw_b200_last_error());
	for (int i = 0; i < n; ++i)
		if (status[i]) throw std::runtime_error("CIGAR " + std::to_string(i) + " overruns a sequence (the reference asserts, src/align.cc:281-282)");
	return st;
}

std::vector<Alignment> from_cigar_batch(const std::vector<std::pair<std::string, std::string>> &pairs,
                                        const std::vector<std::string> &cigars)
{
	if (pairs.size() != cigars.size()) throw std::runtime_error("from_cigar_batch: size mismatch");
	const int n = (int)pairs.size();
	std::vector<Alignment> out(n);
	std::vector<std::deque<std::pair<char, int>>> runs(n);
	for (int i = 0; i < n; ++i) {
		out[i].a = pairs[i].first; out[i].b = pairs[i].second;
		int num = 0;
		for (char ch : cigars[i]) {                                                          // src/align.cc:94-103
			if (ch >= '0' && ch <= '9') num = 10 * num + (ch - '0');
			else if (ch == ';') continue;
			else { runs[i].push_back({ch, num}); num = 0; }
		}
		out[i].cigar = runs[i];
	}
	std::vector<StrPair> sp(n);
	std::vector<const std::deque<std::pair<char, int>> *> cp(n);
	for (int i = 0; i < n; ++i) { sp[i] = {&pairs[i].first, &pairs[i].second}; cp[i] = &runs[i]; }
	std::vector<sd_stats_t> st = stats_of(sp, cp);
	for (int i = 0; i < n; ++i) out[i].stats = st[i];
	return out;
}

std::vector<sd_stats_t> stats_of_alignments(const std::vector<const GuidedAlignment *> &alns)
{
	std::vector<StrPair> sp(alns.size());
	std::vector<const std::deque<std::pair<char, int>> *> cp(alns.size());
	for (size_t i = 0; i < alns.size(); ++i) { sp[i] = {&alns[i]->a, &alns[i]->b}; cp[i] = &alns[i]->cigar; }
	if (alns.empty()) return {};
	return stats_of(sp, cp);
}

// append_cigar (src/align.cc:468-477): merge the first appended run into the last one when the ops are equal
static void append_cigar(std::deque<std::pair<char, int>> &cigar, const std::deque<std::pair<char, int>> &app)
{
	if (app.empty()) return;
	if (!cigar.empty() && cigar.back().first == app.front().first) {
		cigar.back().second += app.front().second;
		cigar.insert(cigar.end(), std::next(app.begin()), app.end());
	} else cigar.insert(cigar.end(), app.begin(), app.end());
}

// Requests given as WINDOWS of the region strings: the strings go up once (one flat buffer per side, every distinct region
// string copied once), a request is four numbers, and the results are read straight out of the arena -- no std::string, no
// Alignment object and no deque per request.  A region's chain wave is hundreds of gap fills of a few bases each (SURVEY 3.2:
// the median call is <= 32 bp), so the per-request host cost is what the wave costs.
namespace {
// The region strings of the fast_align_batch call in flight on this thread, copied ONCE into two page-locked flat buffers: the
// anchors kernel reads them from there, and the chain wave's window requests reference them by offset instead of concatenating
// the strings a second time.  The buffers are grow-only and reused from call to call (cudaHostAlloc of a gigabyte costs more than
// the stage); one call at a time owns them, a concurrent caller simply works without.
struct RegionArena {
	char *q = nullptr, *r = nullptr;
	size_t qcap = 0, rcap = 0;
	int64_t qtot = 0, rtot = 0;
	std::unordered_map<const std::string *, int64_t> qoff, roff;
	std::mutex mu;
	bool reserve(char *&buf, size_t &cap, size_t need)
	{
		if (need <= cap) return true;
		if (buf) ksw_b200_host_free(buf);
		cap = need + need / 4 + 4096;
		buf = (char *)ksw_b200_host_alloc(cap);
		if (!buf) cap = 0;
		return buf != nullptr;
	}
};
RegionArena g_region_arena;
thread_local RegionArena *tl_arena = nullptr;
} // namespace

namespace {
struct WindowBatch {
	std::vector<const std::string *> qstrs, tstrs;       // distinct region strings, in order of first use
	std::vector<int64_t> qbase, tbase;                   // their offsets in the flat buffers
	std::unique_ptr<char[]> qbuf, tbuf;                  // uninitialised; fill() copies the region strings in (first touch in parallel)
	const char *qptr = nullptr, *tptr = nullptr;         // the flat buffers the requests refer to: qbuf / tbuf, or the call's RegionArena
	std::vector<int> ql, tl;
	std::vector<int64_t> qo, to;
	ksw_b200_result_t *res = nullptr;
	const ksw_extz_t *ez = nullptr;
	~WindowBatch() { if (res) ksw_b200_result_free(res); }
	// base offset of a region string (registering it on first use); call serially.  The bytes are copied by fill() afterwards.
	int64_t base_of(const std::string *s, std::vector<const std::string *> &strs, std::vector<int64_t> &bases, int64_t &total, const std::string *&last, int64_t &last_base)
	{
		if (s == last) return last_base;
		for (size_t k = strs.size(); k-- > 0;) if (strs[k] == s) { last = s; return last_base = bases[k]; }
		strs.push_back(s); bases.push_back(total); total += (int64_t)s->size();
		last = s; return last_base = bases.back();
	}
	int64_t qtotal = 0, ttotal = 0;
	void fill()                                          // one allocation per side, the region strings copied in parallel
	{
		qbuf.reset(new char[(size_t)qtotal + 1]); tbuf.reset(new char[(size_t)ttotal + 1]);
		qbuf[(size_t)qtotal] = tbuf[(size_t)ttotal] = '\0';
		qptr = qbuf.get(); tptr = tbuf.get();
#pragma omp parallel for schedule(dynamic, 8)
		for (long k = 0; k < (long)(qstrs.size() + tstrs.size()); ++k) {
			if (k < (long)qstrs.size()) memcpy(&qbuf[(size_t)qbase[k]], qstrs[k]->data(), qstrs[k]->size());
			else { const size_t j = (size_t)k - qstrs.size(); memcpy(&tbuf[(size_t)tbase[j]], tstrs[j]->data(), tstrs[j]->size()); }
		}
	}
	void run(const AlignParams &p)
	{
		const int8_t a = (int8_t)p.match, b = p.mismatch < 0 ? (int8_t)p.mismatch : (int8_t)(-p.mismatch);
		const int8_t mat[25] = {a, b, b, b, 0, b, a, b, b, 0, b, b, a, b, 0, b, b, b, a, 0, 0, 0, 0, 0, 0};
		const double t0 = wall_ms();
		{
			long long cells = 0;
			for (size_t k = 0; k < ql.size(); ++k) cells += (long long)ql[k] * tl[k];
			tl_ksw_pairs += (long long)ql.size(); tl_ksw_cells += cells;
		}
		int rc = ksw_extz2_batch_arena((int)ql.size(), ql.data(), qo.data(), nullptr, tl.data(), to.data(), nullptr, 5, mat,
		                               (int8_t)p.gap_open, (int8_t)p.gap_extend, p.bandwidth, -1, 0, 0, (const uint8_t *)qptr, (const uint8_t *)tptr, &res);
		if (rc) throw std::runtime_error(std::string("ksw_extz2_batch_arena: ") + ksw_b200_strerror(rc) + " -- " + ksw_b200_last_error());
		if (region_trace()) fprintf(stderr, "[regions]     ksw_extz2_batch_arena (windows): %zu pairs, %.1f ms\n", ql.size(), wall_ms() - t0);
		ez = ksw_b200_result_ez(res);
	}
};
// append the raw ksw CIGAR of one result in SEDEF's alphabet (src/align.cc:58-63), merging equal neighbours like append_cigar
inline void append_raw(std::deque<std::pair<char, int>> &cigar, const ksw_extz_t &z)
{
	for (int64_t c = 0; c < z.n_cigar; ++c) {
		const int idx = z.cigar[c] & 0xf, len = (int)(z.cigar[c] >> 4);
		if (idx >= 3) continue;
		const char op = "MDI"[idx];
		if (c == 0 && !cigar.empty() && cigar.back().first == op) cigar.back().second += len;
		else cigar.push_back({op, len});
	}
}
} // namespace

std::vector<GuidedAlignment> align_chains_batch(const std::vector<ChainGuide> &chains, const AlignParams &p)
{
	const long nc = (long)chains.size();
	// pass 1a: how many gap fills does every chain need?  (a fill is needed where both sequences have bases between two anchors)
	std::vector<size_t> fill_first(nc + 1, 0);
#pragma omp parallel for schedule(static) if (nc >= 64)
	for (long ci = 0; ci < nc; ++ci) {
		const ChainGuide &cg = chains[ci];
		const std::vector<Anchor> &g = *cg.anchors;
		size_t cnt = 0;
		for (size_t k = 1; k < cg.guide_idx.size(); ++k) {
			const Anchor &pv = g[cg.guide_idx[k - 1]], &cu = g[cg.guide_idx[k]];
			cnt += (cu.q - (pv.q + pv.l)) != 0 && (cu.r - (pv.r + pv.l)) != 0;
		}
		fill_first[ci + 1] = cnt;
	}
	for (long ci = 0; ci < nc; ++ci) fill_first[ci + 1] += fill_first[ci];
	// the region strings, once each
	WindowBatch wb;
	std::vector<int64_t> cq(nc), ct(nc);
	bool in_arena = tl_arena != nullptr;
	if (in_arena) {
		const std::string *lq = nullptr, *lt = nullptr; int64_t lqb = 0, ltb = 0;
		for (long ci = 0; ci < nc && in_arena; ++ci) {
			if (chains[ci].qstr != lq) { auto it = tl_arena->qoff.find(chains[ci].qstr); if (it == tl_arena->qoff.end()) in_arena = false; else { lq = chains[ci].qstr; lqb = it->second; } }
			if (chains[ci].rstr != lt) { auto it = tl_arena->roff.find(chains[ci].rstr); if (it == tl_arena->roff.end()) in_arena = false; else { lt = chains[ci].rstr; ltb = it->second; } }
			cq[ci] = lqb; ct[ci] = ltb;
		}
		if (in_arena) { wb.qptr = tl_arena->q; wb.tptr = tl_arena->r; }
	}
	if (!in_arena) {
		const std::string *lq = nullptr, *lt = nullptr; int64_t lqb = 0, ltb = 0;
		for (long ci = 0; ci < nc; ++ci) {
			cq[ci] = wb.base_of(chains[ci].qstr, wb.qstrs, wb.qbase, wb.qtotal, lq, lqb);
			ct[ci] = wb.base_of(chains[ci].rstr, wb.tstrs, wb.tbase, wb.ttotal, lt, ltb);
		}
		wb.fill();
	}
	// pass 1b: the requests, as windows
	const size_t nf = fill_first[nc];
	wb.ql.resize(nf); wb.tl.resize(nf); wb.qo.resize(nf); wb.to.resize(nf);
	std::vector<int> tail_len(nf, 0);                                            // > 0: the fill covers mi x mi and a gap of this length follows
	std::vector<char> tail_op(nf, 0);
#pragma omp parallel for schedule(static) if (nc >= 64)
	for (long ci = 0; ci < nc; ++ci) {
		const ChainGuide &cg = chains[ci];
		const std::vector<Anchor> &g = *cg.anchors;
		size_t f = fill_first[ci];
		for (size_t k = 1; k < cg.guide_idx.size(); ++k) {
			const Anchor &pv = g[cg.guide_idx[k - 1]], &cu = g[cg.guide_idx[k]];
			const int qpe = pv.q + pv.l, rpe = pv.r + pv.l;
			const int qgap = cu.q - qpe, rgap = cu.r - rpe;
			if (!(qgap && rgap)) continue;
			int la = qgap, lb = rgap;
			if (!(qgap <= 1000 && rgap <= 1000)) {                                  // src/align.cc:237-246: ma1 (mi x mi, then the rest as a gap) is always taken
				const int ma = std::max(qgap, rgap), mi = std::min(qgap, rgap);
				la = lb = mi; tail_op[f] = qgap == mi ? 'I' : 'D'; tail_len[f] = ma - mi;
			}
			wb.ql[f] = la; wb.tl[f] = lb; wb.qo[f] = cq[ci] + qpe; wb.to[f] = ct[ci] + rpe;
			++f;
		}
	}
	bool chunked = false;
	for (size_t f = 0; f < nf && !chunked; ++f) chunked = wb.ql[f] > kMaxKswSeqLen || wb.tl[f] > kMaxKswSeqLen;
	if (chunked) throw std::runtime_error("align_chains_batch: a gap fill longer than MAX_KSW_SEQ_LEN (chains come with gaps of at most MAX_CHAIN_GAP)");
	wb.run(p);                                                                   // ONE batched ksw_extz2 call
	// pass 2: stitch (chains are independent)
	std::vector<GuidedAlignment> out(nc);
	std::vector<StrPair> finals(nc);
	std::vector<const std::deque<std::pair<char, int>> *> final_cigars(nc);
#pragma omp parallel for schedule(dynamic, 16) if (nc >= 64)
	for (long ci = 0; ci < nc; ++ci) {
		const ChainGuide &cg = chains[ci];
		GuidedAlignment &al = out[ci];
		size_t fpos = fill_first[ci];
		finals[ci] = {&al.a, &al.b}; final_cigars[ci] = &al.cigar;
		if (cg.guide_idx.empty()) continue;                                             // src/align.cc:202-205
		const std::vector<Anchor> &g = *cg.anchors;
		const Anchor &a0 = g[cg.guide_idx[0]];
		al.start_a = a0.q; al.end_a = a0.q + a0.l; al.start_b = a0.r; al.end_b = a0.r + a0.l;
		al.cigar = {{'M', a0.l}};
		for (size_t k = 1; k < cg.guide_idx.size(); ++k) {
			const Anchor &pv = g[cg.guide_idx[k - 1]], &cu = g[cg.guide_idx[k]];
			const int qgap = cu.q - (pv.q + pv.l), rgap = cu.r - (pv.r + pv.l);
			al.end_a = cu.q + cu.l; al.end_b = cu.r + cu.l;
			if (qgap && rgap) {
				// append_cigar(cigar, fill [+ tail]) (src/align.cc:468-477): only the FIRST appended run may merge
				if (wb.ez[fpos].n_cigar > 0) {
					append_raw(al.cigar, wb.ez[fpos]);
					if (tail_op[fpos]) al.cigar.push_back({tail_op[fpos], tail_len[fpos]});
				} else if (tail_op[fpos]) append_cigar(al.cigar, {{tail_op[fpos], tail_len[fpos]}});
				++fpos;
			} else if (qgap) append_cigar(al.cigar, {{'D', qgap}});
			else if (rgap) append_cigar(al.cigar, {{'I', rgap}});
			append_cigar(al.cigar, {{'M', cu.l}});
		}
		al.a = cg.qstr->substr(al.start_a, al.end_a - al.start_a);
		al.b = cg.rstr->substr(al.start_b, al.end_b - al.start_b);
	}
	// populate_nice_alignment for every stitched alignment: one statistics-from-CIGAR call
	std::vector<sd_stats_t> st = stats_of(finals, final_cigars);
	for (long ci = 0; ci < nc; ++ci) out[ci].stats = st[ci];
	return out;
}

// prepend_cigar (src/align.cc:458-466)
static void prepend_cigar(std::deque<std::pair<char, int>> &cigar, const std::deque<std::pair<char, int>> &app)
{
	if (app.empty()) return;
	if (!cigar.empty() && cigar.front().first == app.back().first) {
		cigar.front().second += app.back().second;
		cigar.insert(cigar.begin(), app.begin(), app.begin() + (app.size() - 1));
	} else cigar.insert(cigar.begin(), app.begin(), app.end());
}

static inline bool ceq(char x, char y)                                  // src/align.cc:29-35
{
	auto up = [](char c) { return (c >= 'a' && c <= 'z') ? char(c - 32) : c; };
	if (x == '-' || y == '-') return false;
	if (up(x) == 'N' || up(y) == 'N') return false;
	return up(x) == up(y);
}

// column classes of an alignment: 0 = '|' (M and ceq), 1 = mismatch (both bases), 2 = a is '-' (SEDEF 'I'), 3 = b is '-' ('D')
static std::vector<uint8_t> columns_of(const GuidedAlignment &g)
{
	std::vector<uint8_t> col;
	size_t ia = 0, ib = 0;
	for (auto &c : g.cigar)
		for (int k = 0; k < c.second; ++k) {
			if (c.first == 'M') { col.push_back(ceq(g.a[ia], g.b[ib]) ? 0 : 1); ++ia; ++ib; }
			else if (c.first == 'I') { col.push_back(2); ++ib; }
			else { col.push_back(3); ++ia; }
		}
	return col;
}
static void clear_alignment(GuidedAlignment &g) { g.a.clear(); g.b.clear(); g.cigar.clear(); }

// maximum-suffix scan of trim_front over the columns (src/align.cc:345-365): max_i, or -1 when no suffix scores >= 0
static int scan_trim_front(const GuidedAlignment &g, const AlignParams &p)
{
	const std::vector<uint8_t> col = columns_of(g);
	const int n = (int)col.size();
	int max_score = 0, max_i = -1, score = 0;
	for (int i = n - 1; i >= 0; --i) {
		if (col[i] == 0) score += p.match;
		else if (col[i] == 1) score += p.mismatch;
		else {
			if (i == n - 1 || (col[i] == 2 && col[i + 1] != 2) || (col[i] == 3 && col[i + 1] != 3)) score += -p.gap_open;
			score += -p.gap_extend;
		}
		if (score >= max_score) { max_score = score; max_i = i; }
	}
	return max_i;
}
// maximum-prefix scan of trim_back (src/align.cc:402-420): columns kept (max_i + 1), or -1 when no prefix scores >= 0
static int scan_trim_back(const GuidedAlignment &g, const AlignParams &p)
{
	const std::vector<uint8_t> col = columns_of(g);
	const int n = (int)col.size();
	int max_score = 0, max_i = -1, score = 0;
	for (int i = 0; i < n; ++i) {
		if (col[i] == 0) score += p.match;
		else if (col[i] == 1) score += p.mismatch;
		else {
			if (i == 0 || (col[i] == 2 && col[i - 1] != 2) || (col[i] == 3 && col[i - 1] != 3)) score += -p.gap_open;
			score += -p.gap_extend;
		}
		if (score >= max_score) { max_score = score; max_i = i; }
	}
	return max_i < 0 ? -1 : max_i + 1;
}
// the CIGAR surgery of trim_front for a given scan result (src/align.cc:366-397)
static void apply_trim_front(GuidedAlignment &g, int scan)
{
	const int max_i = scan < 0 ? (int)g.a.size() : scan;              // (sic) the reference initialises max_i with the SEQUENCE length
	if (max_i == (int)g.a.size()) { clear_alignment(g); g.start_a = g.end_a; g.start_b = g.end_b; return; }
	for (int ci = 0, cur_len = 0; ci < (int)g.cigar.size(); ++ci) {
		if (g.cigar[ci].second + cur_len > max_i) {
			const int need = max_i - cur_len;
			g.cigar[ci].second -= need;
			for (int cj = 0; cj < ci; ++cj) g.cigar.pop_front();
			g.start_a += need; g.start_b += need;
			break;
		}
		cur_len += g.cigar[ci].second;
		if (g.cigar[ci].first == 'M') { g.start_a += g.cigar[ci].second; g.start_b += g.cigar[ci].second; }
		else if (g.cigar[ci].first == 'I') g.start_b += g.cigar[ci].second;
		else g.start_a += g.cigar[ci].second;
	}
	g.a = g.a.substr(g.start_a, g.end_a - g.start_a);
	g.b = g.b.substr(g.start_b, g.end_b - g.start_b);
}
// the CIGAR surgery of trim_back for a given scan result (src/align.cc:421-455)
static void apply_trim_back(GuidedAlignment &g, int keep)
{
	if (keep < 0) { clear_alignment(g); g.end_a = g.start_a; g.end_b = g.start_b; return; }
	const int max_i = keep;
	g.end_a = g.start_a; g.end_b = g.start_b;
	for (int ci = 0, cur_len = 0; ci < (int)g.cigar.size(); ++ci) {
		if (g.cigar[ci].second + cur_len >= max_i) {
			const int need = max_i - cur_len;
			g.cigar[ci].second = need;
			while ((int)g.cigar.size() - 1 > ci) g.cigar.pop_back();
			g.end_a += need; g.end_b += need;
			break;
		}
		cur_len += g.cigar[ci].second;
		if (g.cigar[ci].first == 'M') { g.end_a += g.cigar[ci].second; g.end_b += g.cigar[ci].second; }
		else if (g.cigar[ci].first == 'I') g.end_b += g.cigar[ci].second;
		else g.end_a += g.cigar[ci].second;
	}
	g.a = g.a.substr(g.start_a, g.end_a - g.start_a);
	g.b = g.b.substr(g.start_b, g.end_b - g.start_b);
}
void trim_front(GuidedAlignment &g, const AlignParams &p) { apply_trim_front(g, scan_trim_front(g, p)); }   // src/align.cc:343-398   ABCD -> --CD
void trim_back(GuidedAlignment &g, const AlignParams &p) { apply_trim_back(g, scan_trim_back(g, p)); }      // src/align.cc:400-456   ABCD -> AB--

std::vector<GuidedAlignment> align_hit_guides_batch(const std::vector<HitGuide> &guides, const AlignParams &p)
{
	enum { FILL = 0, LEFT = 1, RIGHT = 2 };
	struct Req { size_t guide; int kind; char tail_op; int tail_len; };
	std::vector<std::pair<std::string, std::string>> reqs;
	std::vector<Req> meta;
	std::vector<size_t> req_first(guides.size() + 1, 0);                         // first request of every guide
	for (size_t gi = 0; gi < guides.size(); ++gi) {
		const HitGuide &hg = guides[gi];
		req_first[gi] = meta.size();
		if (hg.guide.empty()) continue;
		const std::string &qstr = *hg.qstr, &rstr = *hg.rstr;
		for (size_t k = 1; k < hg.guide.size(); ++k) {
			const GuidedAlignment &pv = hg.guide[k - 1], &cu = hg.guide[k];
			const int qpe = pv.end_a, rpe = pv.end_b, qs = cu.start_a, rs = cu.start_b;
			const int qgap = qs - qpe, rgap = rs - rpe;
			if (qgap && rgap) {
				if (qgap <= 1000 && rgap <= 1000) { reqs.emplace_back(qstr.substr(qpe, qgap), rstr.substr(rpe, rgap)); meta.push_back({gi, FILL, 0, 0}); }
				else {
					const int ma = std::max(qgap, rgap), mi = std::min(qgap, rgap);
					reqs.emplace_back(qstr.substr(qpe, mi), rstr.substr(rpe, mi));
					meta.push_back({gi, FILL, qgap == mi ? 'I' : 'D', ma - mi});
				}
			}
		}
		if (hg.side) {                                                                   // src/align.cc:153-186
			const int qlo = hg.guide.front().start_a, rlo = hg.guide.front().start_b;
			const int qhi = hg.guide.back().end_a, rhi = hg.guide.back().end_b;
			const int qlo_n = std::max(0, qlo - hg.side), rlo_n = std::max(0, rlo - hg.side);
			if (qlo - qlo_n && rlo - rlo_n) { reqs.emplace_back(qstr.substr(qlo_n, qlo - qlo_n), rstr.substr(rlo_n, rlo - rlo_n)); meta.push_back({gi, LEFT, 0, 0}); }
			const int qhi_n = std::min(qhi + hg.side, (int)qstr.size()), rhi_n = std::min(rhi + hg.side, (int)rstr.size());
			if (qhi_n - qhi && rhi_n - rhi) { reqs.emplace_back(qstr.substr(qhi, qhi_n - qhi), rstr.substr(rhi, rhi_n - rhi)); meta.push_back({gi, RIGHT, 0, 0}); }
		}
	}
	// ONE batched ksw_extz2 call; the trim scans of the side extensions come back with it (computed on the traceback walk)
	req_first[guides.size()] = meta.size();
	const double t_req = wall_ms();
	std::vector<TrimScan> scans;
	std::vector<Alignment> done = align_batch_impl(reqs, p, &scans);
	const double t_aln = wall_ms();
	std::vector<GuidedAlignment> out(guides.size());
	std::vector<StrPair> finals(guides.size());
	std::vector<const std::deque<std::pair<char, int>> *> final_cigars(guides.size());
#pragma omp parallel for schedule(dynamic, 8) if (guides.size() >= 32)
	for (long gi = 0; gi < (long)guides.size(); ++gi) {
		const HitGuide &hg = guides[gi];
		GuidedAlignment &al = out[gi];
		size_t pos = req_first[gi];
		finals[gi] = {&al.a, &al.b}; final_cigars[gi] = &al.cigar;
		if (hg.guide.empty()) continue;
		const std::string &qstr = *hg.qstr, &rstr = *hg.rstr;
		al.cigar = hg.guide.front().cigar;
		for (size_t k = 1; k < hg.guide.size(); ++k) {
			const GuidedAlignment &pv = hg.guide[k - 1], &cu = hg.guide[k];
			const int qgap = cu.start_a - pv.end_a, rgap = cu.start_b - pv.end_b;
			if (qgap && rgap) {
				std::deque<std::pair<char, int>> gc = done[pos].cigar;
				if (meta[pos].tail_op) gc.push_back({meta[pos].tail_op, meta[pos].tail_len});
				append_cigar(al.cigar, gc);
				++pos;
			} else if (qgap) append_cigar(al.cigar, {{'D', qgap}});
			else if (rgap) append_cigar(al.cigar, {{'I', rgap}});
			append_cigar(al.cigar, cu.cigar);
		}
		int qlo = hg.guide.front().start_a, rlo = hg.guide.front().start_b;
		int qhi = hg.guide.back().end_a, rhi = hg.guide.back().end_b;
		if (hg.side) {
			if (pos < meta.size() && meta[pos].guide == (size_t)gi && meta[pos].kind == LEFT) {
				GuidedAlignment gap; gap.a = done[pos].a; gap.b = done[pos].b; gap.cigar = done[pos].cigar;
				gap.start_a = gap.start_b = 0; gap.end_a = (int)gap.a.size(); gap.end_b = (int)gap.b.size();
				apply_trim_front(gap, scans[pos].valid ? scans[pos].front : scan_trim_front(gap, p));
				qlo -= gap.end_a - gap.start_a; rlo -= gap.end_b - gap.start_b;
				prepend_cigar(al.cigar, gap.cigar);
				++pos;
			}
			if (pos < meta.size() && meta[pos].guide == (size_t)gi && meta[pos].kind == RIGHT) {
				GuidedAlignment gap; gap.a = done[pos].a; gap.b = done[pos].b; gap.cigar = done[pos].cigar;
				gap.start_a = gap.start_b = 0; gap.end_a = (int)gap.a.size(); gap.end_b = (int)gap.b.size();
				apply_trim_back(gap, scans[pos].valid ? scans[pos].back : scan_trim_back(gap, p));
				qhi += gap.end_a; rhi += gap.end_b;
				append_cigar(al.cigar, gap.cigar);
				++pos;
			}
		}
		al.start_a = qlo; al.end_a = qhi; al.start_b = rlo; al.end_b = rhi;
		al.a = qstr.substr(qlo, qhi - qlo); al.b = rstr.substr(rlo, rhi - rlo);
	}
	const double t_st = wall_ms();
	std::vector<sd_stats_t> st = stats_of(finals, final_cigars);
	if (region_trace()) fprintf(stderr, "[regions]   guides: %zu requests, align %.1f ms, stitch %.1f ms, statistics %.1f ms\n", reqs.size(), t_aln - t_req, t_st - t_aln, wall_ms() - t_st);
	for (size_t gi = 0; gi < guides.size(); ++gi) out[gi].stats = st[gi];
	return out;
}

// ---- merge ---------------------------------------------------------------------------------------------------------
static std::vector<char> expand_ops(const std::deque<std::pair<char, int>> &cigar)
{
	std::vector<char> ops;
	for (auto &c : cigar) ops.insert(ops.end(), c.second, c.first);
	return ops;
}
// cigar_from_alignment (src/align.cc:479-501): an empty alignment yields the single run ('\0', 0)
static std::deque<std::pair<char, int>> cigar_from_ops(const std::vector<char> &ops)
{
	std::deque<std::pair<char, int>> cigar;
	int sz = 0; char op = 0;
	for (char top : ops) {
		if (op != top) { if (op) cigar.push_back({op, sz}); op = top; sz = 0; }
		sz++;
	}
	cigar.push_back({op, sz});
	return cigar;
}
// cut columns from the tail of `ops` until `lim` bases of a (by_a) / b have been removed; returns removed (q, r)
static void cut_tail(std::vector<char> &ops, int lim, bool by_a, int &q, int &r)
{
	q = 0; r = 0;
	while (!ops.empty() && (by_a ? q : r) < lim) {
		const char op = ops.back(); ops.pop_back();
		if (op != 'I') q++;                                              // align_a[i] != '-'
		if (op != 'D') r++;                                              // align_b[i] != '-'
	}
}
static void cut_head(std::vector<char> &ops, int lim, bool by_a, int &q, int &r)
{
	q = 0; r = 0;
	size_t i = 0;
	for (; i < ops.size() && (by_a ? q : r) < lim; ++i) {
		if (ops[i] != 'I') q++;
		if (ops[i] != 'D') r++;
	}
	ops.erase(ops.begin(), ops.begin() + i);
}

std::vector<GuidedAlignment> merge_batch(const std::vector<MergeRequest> &reqs, const AlignParams &p)
{
	struct Work { GuidedAlignment prev, cur; int fill = -1; char tail_op = 0; int tail_len = 0; int qgap = 0, rgap = 0; std::string fa, fb; bool want = false; };
	std::vector<Work> work(reqs.size());
	std::vector<std::pair<std::string, std::string>> fills;
#pragma omp parallel for schedule(dynamic, 8) if (reqs.size() >= 32)
	for (long k = 0; k < (long)reqs.size(); ++k) {
		Work &w = work[k];
		w.prev = reqs[k].prev; w.cur = reqs[k].cur;
		const std::string &qstr = *reqs[k].qstr, &rstr = *reqs[k].rstr;
		std::vector<char> po = expand_ops(w.prev.cigar), co = expand_ops(w.cur.cigar);
		int q, r;
		int trim = w.prev.end_a - w.cur.start_a;                                           // src/align.cc:510-538
		cut_tail(po, trim, true, q, r); w.prev.end_a -= q; w.prev.end_b -= r;
		cut_head(co, trim, true, q, r); w.cur.start_a += q; w.cur.start_b += r;
		trim = w.prev.end_b - w.cur.start_b;                                               // src/align.cc:540-568
		cut_tail(po, trim, false, q, r); w.prev.end_a -= q; w.prev.end_b -= r;
		cut_head(co, trim, false, q, r); w.cur.start_a += q; w.cur.start_b += r;
		w.prev.cigar = cigar_from_ops(po); w.cur.cigar = cigar_from_ops(co);               // src/align.cc:570-571
		w.qgap = w.cur.start_a - w.prev.end_a; w.rgap = w.cur.start_b - w.prev.end_b;
		if (w.qgap && w.rgap) {                                                            // src/align.cc:579-594
			w.want = true;
			if (w.qgap <= 1000 && w.rgap <= 1000) { w.fa = qstr.substr(w.prev.end_a, w.qgap); w.fb = rstr.substr(w.prev.end_b, w.rgap); }
			else {
				const int ma = std::max(w.qgap, w.rgap), mi = std::min(w.qgap, w.rgap);
				w.fa = qstr.substr(w.prev.end_a, mi); w.fb = rstr.substr(w.prev.end_b, mi);
				w.tail_op = w.qgap == mi ? 'I' : 'D'; w.tail_len = ma - mi;
			}
		}
	}
	for (size_t k = 0; k < reqs.size(); ++k)
		if (work[k].want) { work[k].fill = (int)fills.size(); fills.emplace_back(std::move(work[k].fa), std::move(work[k].fb)); }
	std::vector<Alignment> done = align_batch(fills, p);                                  // ONE batched ksw_extz2 call
	std::vector<GuidedAlignment> out(reqs.size());
	std::vector<StrPair> finals(reqs.size());
	std::vector<const std::deque<std::pair<char, int>> *> final_cigars(reqs.size());
#pragma omp parallel for schedule(dynamic, 8) if (reqs.size() >= 32)
	for (long k = 0; k < (long)reqs.size(); ++k) {
		Work &w = work[k];
		GuidedAlignment &al = out[k];
		finals[k] = {&al.a, &al.b}; final_cigars[k] = &al.cigar;
		al = w.prev;
		if (w.fill >= 0) {
			std::deque<std::pair<char, int>> gc = done[w.fill].cigar;
			if (w.tail_op) gc.push_back({w.tail_op, w.tail_len});
			append_cigar(al.cigar, gc);
		} else if (w.qgap) append_cigar(al.cigar, {{'D', w.qgap}});
		else if (w.rgap) append_cigar(al.cigar, {{'I', w.rgap}});
		al.end_a = w.cur.end_a; al.end_b = w.cur.end_b;
		append_cigar(al.cigar, w.cur.cigar);
		al.a = reqs[k].qstr->substr(al.start_a, al.end_a - al.start_a);
		al.b = reqs[k].rstr->substr(al.start_b, al.end_b - al.start_b);
	}
	std::vector<sd_stats_t> st = stats_of(finals, final_cigars);
	for (size_t k = 0; k < reqs.size(); ++k) out[k].stats = st[k];
	return out;
}

// ---- region-level driver ------------------------------------------------------------------------------------------
namespace {
// Globals::Chain::Refine (src/globals.h:80-86)
const double kRefMatch = 10, kRefMismatch = 1, kRefGap = 0.5, kRefGapOpen = 100;
const int kRefMinRead = 900, kRefSideAlign = 500, kRefMaxGap = 10 * 1000;

struct RegionState {
	std::vector<GuidedAlignment> anc;                   // chain alignments, sorted like refine_chains sorts its hits
	std::vector<int> prev;
	std::vector<std::pair<int, int>> order;             // (dp, index), descending: the `maxes` set of src/refine.cc:39
	size_t mpos = 0;
	std::vector<char> used;
	std::vector<GuidedAlignment> accepted;
	std::deque<int> path;
	size_t pi = 0;
	int prev_idx = -1;
	std::vector<GuidedAlignment> guide;
	enum Phase { NEXT_PATH, WALK, WAIT_MERGE, WAIT_GUIDE, DONE } phase = NEXT_PATH;
};

// a guide / merge request needs the coordinates, the CIGAR and the counters of a chain alignment, not its two strings
inline GuidedAlignment light(const GuidedAlignment &g)
{
	GuidedAlignment o;
	o.cigar = g.cigar; o.stats = g.stats;
	o.start_a = g.start_a; o.end_a = g.end_a; o.start_b = g.start_b; o.end_b = g.end_b;
	return o;
}

inline bool hit_less(const GuidedAlignment &a, const GuidedAlignment &b)        // Hit::operator< (src/hit.h:44-47)
{
	return std::tie(a.start_a, a.end_a, a.start_b, a.end_b) < std::tie(b.start_a, b.end_a, b.start_b, b.end_b);
}

// the dynamic program of refine_chains over one region's chain alignments (src/refine.cc:27-98)
void refine_dp(const RegionTask &t, RegionState &st)
{
	std::sort(st.anc.begin(), st.anc.end(), hit_less);
	const int n = (int)st.anc.size();
	std::vector<int> score(n);
	for (int i = 0; i < n; ++i)
		score[i] = (int)(+kRefMatch * st.anc[i].matches() - kRefMismatch * st.anc[i].mismatches() - kRefGap * st.anc[i].gap_bases());
	std::vector<int> dp(n, 0);
	st.prev.assign(n, -1);
	st.used.assign(n, 0);
	for (int ai = 0; ai < n; ++ai) {
		const GuidedAlignment &c = st.anc[ai];
		if (t.same_chr) {
			const int qlo = c.start_a, qhi = c.end_a, rlo = c.start_b, rhi = c.end_b;
			const int qo = std::max(0, std::min(t.orig_query_start + qhi, t.orig_ref_start + rhi) - std::max(t.orig_query_start + qlo, t.orig_ref_start + rlo));
			if ((rhi - rlo) - qo < kRefSideAlign && (qhi - qlo) - qo < kRefSideAlign) continue;       // no gap between
		}
		dp[ai] = score[ai];
		for (int aj = ai - 1; aj >= 0; --aj) {
			const GuidedAlignment &pv = st.anc[aj];
			int cqs = c.start_a; if (cqs < pv.end_a) cqs = pv.end_a;
			int crs = c.start_b; if (crs < pv.end_b) crs = pv.end_b;
			if (pv.end_a >= c.end_a || pv.end_b >= c.end_b) continue;
			if (pv.start_b >= c.start_b) continue;
			const int ma = std::max(cqs - pv.end_a, crs - pv.end_b), mi = std::min(cqs - pv.end_a, crs - pv.end_b);
			if (ma >= kRefMaxGap) continue;
			if (t.same_chr) {
				const int qlo = pv.end_a, qhi = cqs, rlo = pv.end_b, rhi = crs;
				const int qo = std::max(0, std::min(t.orig_query_start + qhi, t.orig_ref_start + rhi) - std::max(t.orig_query_start + qlo, t.orig_ref_start + rlo));
				if (qo >= 1) continue;
			}
			const int mis = (int)(kRefMismatch * mi), gap = (int)(kRefGapOpen + kRefGap * (ma - mi));
			const int sco = dp[aj] + score[ai] - mis - gap;
			if (sco >= dp[ai]) { dp[ai] = sco; st.prev[ai] = aj; }
		}
		st.order.push_back({dp[ai], ai});
	}
	std::sort(st.order.begin(), st.order.end(), std::greater<std::pair<int, int>>());    // set<..., greater<>>: keys are unique (index)
}

// advance one region until it needs an alignment (returns true and fills the request) or is done
struct Request { int kind; MergeRequest merge; HitGuide hg; };     // kind 1 = merge, 2 = final guide constructor
bool advance(const RegionTask &t, RegionState &st, Request &rq)
{
	for (;;) {
		if (st.phase == RegionState::DONE) return false;
		if (st.phase == RegionState::NEXT_PATH) {
			if (st.mpos >= st.order.size() || st.order[st.mpos].first == 0) { st.phase = RegionState::DONE; return false; }   // src/refine.cc:104-105
			int maxi = st.order[st.mpos++].second;
			if (st.used[maxi]) continue;
			st.path.clear();
			while (maxi != -1 && !st.used[maxi]) { st.path.push_front(maxi); st.used[maxi] = 1; maxi = st.prev[maxi]; }
			const int qlo = st.anc[st.path.front()].start_a, qhi = st.anc[st.path.back()].end_a;
			const int rlo = st.anc[st.path.front()].start_b, rhi = st.anc[st.path.back()].end_b;
			int est_size = st.anc[st.path[0]].span();
			for (size_t i = 1; i < st.path.size(); ++i) {
				est_size += st.anc[st.path[i]].span();
				est_size += std::max(st.anc[st.path[i]].start_a - st.anc[st.path[i - 1]].end_a, st.anc[st.path[i]].start_b - st.anc[st.path[i - 1]].end_b);
			}
			if (est_size < kRefMinRead - kRefSideAlign) continue;                                        // src/refine.cc:142-146
			bool overlap = false;
			for (auto &h : st.accepted) {                                                                 // src/refine.cc:148-160
				const int qo = std::max(0, std::min(qhi, h.end_a) - std::max(qlo, h.start_a));
				const int ro = std::max(0, std::min(rhi, h.end_b) - std::max(rlo, h.start_b));
				if (qhi - qlo - qo < kRefSideAlign && rhi - rlo - ro < kRefSideAlign) { overlap = true; break; }
			}
			if (overlap) continue;
			st.guide.clear();
			st.prev_idx = st.path[0]; st.pi = 1;
			st.phase = RegionState::WALK;
		}
		if (st.phase == RegionState::WALK) {                                                              // src/refine.cc:167-179
			while (st.pi < st.path.size()) {
				const GuidedAlignment &cur = st.anc[st.path[st.pi]];
				const GuidedAlignment &pv = st.anc[st.prev_idx];
				if (cur.start_a < pv.end_a || cur.start_b < pv.end_b) {
					rq.kind = 1; rq.merge = MergeRequest{light(pv), light(cur), t.qstr, t.rstr};
					st.phase = RegionState::WAIT_MERGE;
					return true;
				}
				st.guide.push_back(light(pv));
				st.prev_idx = st.path[st.pi++];
			}
			st.guide.push_back(light(st.anc[st.prev_idx]));
			rq.kind = 2; rq.hg = HitGuide{t.qstr, t.rstr, std::move(st.guide), kRefSideAlign};
			st.guide.clear();
			st.phase = RegionState::WAIT_GUIDE;
			return true;
		}
		return false;   // WAIT_*: the caller has not delivered the result yet
	}
}
} // namespace

std::vector<std::vector<GuidedAlignment>> refine_regions_batch(const std::vector<RegionTask> &regions, const AlignParams &p, RefineStats *stats)
{
	RefineStats rs;
	const long long pairs0 = tl_ksw_pairs, cells0 = tl_ksw_cells;
	const double t_begin = wall_ms();
	// wave 0: every chain of every region through one batched call
	std::vector<ChainGuide> chains;
	std::vector<size_t> owner;
	for (size_t ri = 0; ri < regions.size(); ++ri)
		for (auto &g : regions[ri].guides) { chains.push_back(ChainGuide{regions[ri].qstr, regions[ri].rstr, regions[ri].anchors, g}); owner.push_back(ri); }
	std::vector<GuidedAlignment> wave0 = align_chains_batch(chains, p);
	rs.batch_calls += 2; rs.ksw_requests += (long long)chains.size(); rs.rounds = 1;
	if (region_trace()) fprintf(stderr, "[regions] chain wave: %zu chains, %.1f ms\n", chains.size(), wall_ms() - t_begin);
	std::vector<RegionState> st(regions.size());
	for (size_t k = 0; k < wave0.size(); ++k) st[owner[k]].anc.push_back(std::move(wave0[k]));
#pragma omp parallel for schedule(dynamic, 4) if (regions.size() >= 16)
	for (long ri = 0; ri < (long)regions.size(); ++ri) refine_dp(regions[ri], st[ri]);
	// waves 1..: every region contributes the next step of its current path
	for (;;) {
		std::vector<MergeRequest> merges; std::vector<size_t> merge_owner;
		std::vector<HitGuide> guides; std::vector<size_t> guide_owner;
		std::vector<Request> rqs(regions.size());
		std::vector<char> has(regions.size(), 0);
#pragma omp parallel for schedule(dynamic, 4) if (regions.size() >= 16)
		for (long ri = 0; ri < (long)regions.size(); ++ri) has[ri] = advance(regions[ri], st[ri], rqs[ri]) ? 1 : 0;
		for (size_t ri = 0; ri < regions.size(); ++ri) {
			if (!has[ri]) continue;
			if (rqs[ri].kind == 1) { merges.push_back(std::move(rqs[ri].merge)); merge_owner.push_back(ri); }
			else { guides.push_back(std::move(rqs[ri].hg)); guide_owner.push_back(ri); }
		}
		if (merges.empty() && guides.empty()) break;
		++rs.rounds;
		const double t_round = wall_ms();
		if (!merges.empty()) {
			std::vector<GuidedAlignment> done = merge_batch(merges, p);
			rs.batch_calls += 2; rs.ksw_requests += (long long)merges.size();
			for (size_t k = 0; k < done.size(); ++k) {
				RegionState &s = st[merge_owner[k]];
				s.anc[s.prev_idx] = std::move(done[k]);                                                   // prev->aln.merge(...); update_from_alignment(*prev)
				++s.pi;
				s.phase = RegionState::WALK;
			}
		}
		if (!guides.empty()) {
			std::vector<GuidedAlignment> done = align_hit_guides_batch(guides, p);
			rs.batch_calls += 2; rs.ksw_requests += (long long)guides.size();
			for (size_t k = 0; k < done.size(); ++k) {
				RegionState &s = st[guide_owner[k]];
				if (done[k].span() >= kRefMinRead) s.accepted.push_back(std::move(done[k]));                  // src/refine.cc:186-191
				s.phase = RegionState::NEXT_PATH;
			}
		}
		if (region_trace()) fprintf(stderr, "[regions] round %d: %zu merges, %zu guides, %.1f ms\n", rs.rounds, merges.size(), guides.size(), wall_ms() - t_round);
	}
	std::vector<std::vector<GuidedAlignment>> out(regions.size());
	for (size_t ri = 0; ri < regions.size(); ++ri) out[ri] = std::move(st[ri].accepted);
	rs.ksw_pairs = tl_ksw_pairs - pairs0; rs.ksw_cells = tl_ksw_cells - cells0;
	if (stats) *stats = rs;
	return out;
}

// ---- anchoring + chaining -----------------------------------------------------------------------------------------------
std::vector<std::vector<Anchor>> anchors_batch(const std::vector<RegionSeed> &regions, int kmer_size)
{
	const int n = (int)regions.size();
	std::vector<int> ql(n), rl(n);
	std::vector<int64_t> qo(n), ro(n), oq(n), orr(n);
	std::vector<uint8_t> same(n);
	int64_t qtot = 0, rtot = 0;
	for (int i = 0; i < n; ++i) {
		ql[i] = (int)regions[i].qstr->size(); rl[i] = (int)regions[i].rstr->size();
		qo[i] = qtot; ro[i] = rtot; qtot += ql[i]; rtot += rl[i];
		same[i] = regions[i].same_chr; oq[i] = regions[i].orig_query_start; orr[i] = regions[i].orig_ref_start;
	}
	// the flat buffers: the call's page-locked RegionArena (fast_align_batch), or two plain ones
	std::unique_ptr<char[]> qown, rown;
	char *qbuf = nullptr, *rbuf = nullptr;
	RegionArena *ar = tl_arena;
	if (ar && ar->reserve(ar->q, ar->qcap, (size_t)qtot + 16) && ar->reserve(ar->r, ar->rcap, (size_t)rtot + 16)) {
		qbuf = ar->q; rbuf = ar->r; ar->qtot = qtot; ar->rtot = rtot;
		ar->qoff.clear(); ar->roff.clear();
		ar->qoff.reserve((size_t)n * 2); ar->roff.reserve((size_t)n * 2);
		for (int i = 0; i < n; ++i) { ar->qoff.emplace(regions[i].qstr, qo[i]); ar->roff.emplace(regions[i].rstr, ro[i]); }   // (a string used twice: its first copy)
	} else {
		if (ar) { ar->qoff.clear(); ar->roff.clear(); }
		qown.reset(new char[(size_t)qtot + 1]); rown.reset(new char[(size_t)rtot + 1]);              // uninitialised: first touch in parallel
		qbuf = qown.get(); rbuf = rown.get();
	}
	qbuf[(size_t)qtot] = rbuf[(size_t)rtot] = '\0';
#pragma omp parallel for schedule(dynamic, 16)
	for (int i = 0; i < n; ++i) {
		memcpy(qbuf + qo[i], regions[i].qstr->data(), (size_t)ql[i]);
		memcpy(rbuf + ro[i], regions[i].rstr->data(), (size_t)rl[i]);
	}
	sedef_anchor_t *flat = nullptr;
	std::vector<int64_t> off(n + 1, 0);
	int rc = sedef_anchors_batch(n, ql.data(), qo.data(), (const uint8_t *)qbuf, rl.data(), ro.data(), (const uint8_t *)rbuf, kmer_size,
	                             same.data(), oq.data(), orr.data(), &flat, off.data());
	if (rc) throw std::runtime_error(std::string("sedef_anchors_batch: ") + ksw_b200_strerror(rc) + " -- " + ksw_b200_last_error());
	std::vector<std::vector<Anchor>> out(n);
#pragma omp parallel for schedule(dynamic, 16)
	for (int i = 0; i < n; ++i) {
		out[i].reserve((size_t)(off[i + 1] - off[i]));
		for (int64_t k = off[i]; k < off[i + 1]; ++k) out[i].push_back(Anchor{flat[k].q, flat[k].r, flat[k].l, flat[k].has_u});
	}
	free(flat);
	return out;
}

namespace {
// Globals::Chain / Globals::Search (src/globals.h:71-75, src/globals.cc:20-30)
const int kMatchChainScore = 4, kMaxChainGap = 210, kMinUppercaseMatch = 90;
const double kMinChainSpan = 700 * (1 - 0.30);             // MIN_READ_SIZE * (1 - MAX_ERROR), evaluated in double like the reference

// The reference's range structure (src/segment.h, src/segment.tpp): a static binary tree over the points sorted by key; every
// node additionally holds the best ACTIVE point of its subtree that no ancestor holds (a priority search tree).  Ties are part
// of the contract -- chains are extracted by following `prev[]` -- so the same rules are kept: on activation an equal score
// displaces the resident point (>=), a range query prefers the LEFT subtree on equal scores (>=), a removal promotes the right
// child's point only when it is strictly better.
struct PointTree {
	typedef std::pair<int, int> Key;
	struct Node { int best = -1, leaf = -1; Key hi; };     // best: node index of the held leaf; leaf: index into pts (-1: internal)
	struct Pt { Key x; int score, pos; };
	static const int kMin = INT_MIN;
	std::vector<Node> t;
	std::vector<Pt> &pts;
	explicit PointTree(std::vector<Pt> &p) : pts(p)
	{
		std::sort(pts.begin(), pts.end(), [](const Pt &a, const Pt &b) { return a.x < b.x; });
		int size = 1;
		while (size < (int)pts.size()) size <<= 1;
		if (pts.size() <= 1) size = 1;
		t.resize((size_t)size << 1);
		int next = 0;
		build(0, 0, (int)pts.size(), next);
	}
	void build(int i, int s, int e, int &next)
	{
		if (i >= (int)t.size() || s >= e) return;
		if (s + 1 == e) { t[i].leaf = next; t[i].hi = pts[next].x; pts[next].score = kMin; ++next; return; }
		const int mid = (s + e + 1) / 2;
		build(2 * i + 1, s, mid, next);
		build(2 * i + 2, mid, e, next);
		t[i].hi = t[2 * i + 1 + (2 * i + 2 < (int)t.size())].hi;
	}
	int find_leaf(const Key &q) const
	{
		int i = 0;
		while (i < (int)t.size() && (t[i].leaf == -1 || q != pts[t[i].leaf].x)) i = 2 * i + 1 + (q > t[2 * i + 1].hi);
		return i;
	}
	int score_of(int node) const { return pts[t[node].leaf].score; }
	void activate(const Key &q, int score)
	{
		int carry = find_leaf(q);
		pts[t[carry].leaf].score = score;
		for (int i = 0; i < (int)t.size();) {
			if (t[i].best == -1 || score_of(carry) >= score_of(t[i].best)) std::swap(t[i].best, carry);
			if (carry == -1) break;
			i = 2 * i + 1 + (pts[t[carry].leaf].x > t[2 * i + 1].hi);
		}
	}
	void deactivate(const Key &q)
	{
		int gone = find_leaf(q);
		pts[t[gone].leaf].score = kMin;
		for (int i = 0; i < (int)t.size();) {
			if (t[i].best == -1) break;
			if (t[i].best == gone) {
				if (t[i].leaf != -1) t[i].best = -1;
				else {
					const int l = 2 * i + 1, r = 2 * i + 2;
					if (r < (int)t.size() && t[r].best != -1 && (t[l].best == -1 || score_of(t[r].best) > score_of(t[l].best))) { t[i].best = gone = t[r].best; i = r; }
					else { t[i].best = gone = t[l].best; i = l; }
				}
			} else i = 2 * i + 1 + (q > t[2 * i + 1].hi);
		}
	}
	int query(const Key &lo, const Key &hi, int i) const      // node index of the best point with lo <= key <= hi, or -1
	{
		if (i >= (int)t.size()) return -1;
		if (t[i].leaf != -1) return (lo <= pts[t[i].leaf].x && pts[t[i].leaf].x <= hi) ? i : -1;
		const int b = t[i].best;
		if (b == -1) return -1;
		if (lo <= pts[t[b].leaf].x && pts[t[b].leaf].x <= hi) return b;
		if (hi <= t[2 * i + 1].hi) return query(lo, hi, 2 * i + 1);
		if (lo > t[2 * i + 1].hi) return query(lo, hi, 2 * i + 2);
		const int m1 = query(lo, hi, 2 * i + 1), m2 = query(lo, hi, 2 * i + 2);
		if (m1 == -1) return m2;
		if (m2 == -1) return m1;
		return score_of(m1) >= score_of(m2) ? m1 : m2;
	}
	int query(const Key &lo, const Key &hi) const { const int i = query(lo, hi, 0); return i == -1 ? -1 : t[i].leaf; }
};
} // namespace

std::vector<std::vector<int>> chain_anchors(const std::vector<Anchor> &anchors)
{
	const int n = (int)anchors.size();
	std::vector<std::vector<int>> guides;
	if (n == 0) return guides;
	struct Ev { std::pair<int, int> x; };
	std::vector<Ev> xs; xs.reserve(2 * (size_t)n);
	std::vector<PointTree::Pt> ys; ys.reserve(n);
	int max_q = 0, max_r = 0;
	for (int i = 0; i < n; ++i) {
		const Anchor &a = anchors[i];
		xs.push_back({{a.q, i}}); xs.push_back({{a.q + a.l, i}});
		ys.push_back({{a.r + a.l - 1, i}, PointTree::kMin, i});
		max_q = std::max(max_q, a.q + a.l); max_r = std::max(max_r, a.r + a.l);
	}
	std::sort(xs.begin(), xs.end(), [](const Ev &a, const Ev &b) { return a.x < b.x; });
	PointTree tree(ys);
	std::vector<int> prev(n, -1);
	std::vector<std::pair<int, int>> dp(n);
	for (int i = 0; i < n; ++i) dp[i] = {0, i};
	size_t bound = 0;
	for (size_t xi = 0; xi < xs.size(); ++xi) {
		const int i = xs[xi].x.second;
		const Anchor &a = anchors[i];
		if (xs[xi].x.first == a.q) {                                                      // the anchor starts: pick its predecessor
			while (bound < xi) {                                                        // retire end points that are too far behind
				const int t = xs[bound].x.second;
				if (xs[bound].x.first == anchors[t].q + anchors[t].l) {
					if (a.q - (anchors[t].q + anchors[t].l) <= kMaxChainGap) break;
					tree.deactivate({anchors[t].r + anchors[t].l - 1, t});
				}
				++bound;
			}
			const int w = kMatchChainScore * a.has_u + (kMatchChainScore / 2) * (a.l - a.has_u);
			int j = tree.query({a.r - kMaxChainGap, 0}, {a.r - 1, n});
			if (j != -1 && ys[j].score != PointTree::kMin) {
				j = ys[j].pos;
				const Anchor &pv = anchors[j];
				const int gap = a.q - (pv.q + pv.l) + a.r - (pv.r + pv.l);
				if (w + dp[j].first - gap > 0) { dp[i].first = w + dp[j].first - gap; prev[i] = j; }
				else dp[i].first = w;
			} else dp[i].first = w;
		} else {                                                                            // the anchor ends: it becomes a candidate
			const int gap = max_q + 1 - (a.q + a.l) + max_r + 1 - (a.r + a.l);
			tree.activate({a.r + a.l - 1, i}, dp[i].first - gap);
		}
	}
	std::sort(dp.begin(), dp.end(), std::greater<std::pair<int, int>>());
	std::vector<char> used(n, 0);
	for (auto &m : dp) {                                                                    // chains in score order (chain.cc:181-197)
		int at = m.second;
		if (used[at]) continue;
		std::vector<int> path;                                                              // from the chain's LAST anchor back to its first
		int has_u = 0;
		while (at != -1 && !used[at]) { path.push_back(at); has_u += anchors[at].has_u; used[at] = 1; at = prev[at]; }
		// the filter of fast_align (chain.cc:222-247)
		const Anchor &first = anchors[path.back()], &last = anchors[path.front()];
		const int span = std::max(last.r + last.l - first.r, last.q + last.l - first.q);
		if ((!(has_u != 0) || span < kMinUppercaseMatch) && span < kMinChainSpan) continue;
		guides.emplace_back(path.rbegin(), path.rend());
	}
	return guides;
}

std::vector<std::vector<GuidedAlignment>> fast_align_batch(const std::vector<RegionSeed> &regions, int kmer_size, const AlignParams &p,
                                                           RefineStats *stats)
{
	const double t0 = wall_ms();
	// the call's region arena (page-locked flat copies of the region strings, shared by the anchors kernel and the chain wave)
	struct ArenaGuard {
		bool own = false;
		ArenaGuard() { if (!tl_arena && g_region_arena.mu.try_lock()) { own = true; tl_arena = &g_region_arena; g_region_arena.qoff.clear(); g_region_arena.roff.clear(); } }
		~ArenaGuard() { if (own) { g_region_arena.qoff.clear(); g_region_arena.roff.clear(); tl_arena = nullptr; g_region_arena.mu.unlock(); } }
	} arena_guard;
	std::vector<std::vector<Anchor>> anchors = anchors_batch(regions, kmer_size);
	const double t1 = wall_ms();
	std::vector<RegionTask> tasks(regions.size());
#pragma omp parallel for schedule(dynamic, 1)
	for (long ri = 0; ri < (long)regions.size(); ++ri) {
		RegionTask &t = tasks[ri];
		t.qstr = regions[ri].qstr; t.rstr = regions[ri].rstr; t.anchors = &anchors[ri];
		t.same_chr = regions[ri].same_chr; t.orig_query_start = regions[ri].orig_query_start; t.orig_ref_start = regions[ri].orig_ref_start;
		t.guides = chain_anchors(anchors[ri]);
	}
	if (region_trace()) fprintf(stderr, "[regions] anchors %.1f ms, chaining %.1f ms\n", t1 - t0, wall_ms() - t1);
	// (Cutting large region sets into groups whose wave sequences run concurrently from several host threads was tried and is
	// SLOWER -- 4000 regions: 2300 ms in 8 groups against 1322 ms in one -- the groups' persistent DP grids and their host
	// threads get in each other's way; profiles/r02_tuning.md.)
	return refine_regions_batch(tasks, p, stats);
}

} // namespace sedef_b200

// chains of one region for callers without C++ (and the CPU parity test): anchors as (q, r, l, has_u) rows; chain k holds
// chain_len[k] anchor indices, concatenated in chain_idx.  Returns the number of chains (fills at most cap_* entries).
extern "C" int sedef_b200_chain_anchors(int n, const int32_t *anchors4, int cap_chains, int *chain_len, int cap_idx, int *chain_idx)
{
	std::vector<sedef_b200::Anchor> a(n < 0 ? 0 : n);
	for (int i = 0; i < n; ++i) a[i] = sedef_b200::Anchor{anchors4[4 * i], anchors4[4 * i + 1], anchors4[4 * i + 2], anchors4[4 * i + 3]};
	std::vector<std::vector<int>> g = sedef_b200::chain_anchors(a);
	int pos = 0;
	for (size_t k = 0; k < g.size(); ++k) {
		if ((int)k < cap_chains) chain_len[k] = (int)g[k].size();
		for (int v : g[k]) { if (pos < cap_idx) chain_idx[pos] = v; ++pos; }
	}
	return (int)g.size();
}

// C view of the chunk plan, for hosts that drive the C ABI themselves (and for the parity test of the chunk arithmetic)
extern "C" int sedef_b200_chunk_plan(int64_t alen, int64_t blen, int cap, int64_t *sp, int *qlen, int *tlen)
{
	std::vector<sedef_b200::Chunk> c;
	if (alen < 0 || blen < 0) return -1;
	sedef_b200::chunk_plan((size_t)alen, (size_t)blen, c);
	for (size_t k = 0; k < c.size() && (int)k < cap; ++k) { sp[k] = (int64_t)c[k].sp; qlen[k] = c[k].la; tlen[k] = c[k].lb; }
	return (int)c.size();
}
