// sedef_align.cc -- see include/sedef_align.hpp.  Host C++ on top of the C ABI; compiled into libsedef_b200.so.
#include "../../../include/sedef_align.hpp"
#include <algorithm>
#include <cstdlib>
#include <stdexcept>

namespace sedef_b200 {

static const int kMaxKswSeqLen = 60 * 1024;            // Globals::Align::MAX_KSW_SEQ_LEN (src/globals.h:54)

static inline uint8_t align_dna(char c)                // src/common.h:58-70,91
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}

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
	for (auto &c : cigar) { s += std::to_string(c.second); s += c.first; }
	return s;
}
sd_stats_fp_t Alignment::bedpe_fp() const
{
	sd_stats_fp_t o;
	sd_stats_derive_fp(&stats, &o);
	return o;
}

static void add_stats(sd_stats_t &d, const sd_stats_t &s)
{
	int32_t *dp = reinterpret_cast<int32_t *>(&d);
	const int32_t *sp = reinterpret_cast<const int32_t *>(&s);
	for (size_t k = 0; k < sizeof(sd_stats_t) / sizeof(int32_t); ++k) dp[k] += sp[k];
}

std::vector<Alignment> align_batch(const std::vector<std::pair<std::string, std::string>> &pairs, const AlignParams &p)
{
	// align_helper's matrix (src/align.cc:41-44)
	const int8_t a = (int8_t)p.match, b = p.mismatch < 0 ? (int8_t)p.mismatch : (int8_t)(-p.mismatch);
	const int8_t mat[25] = {a, b, b, b, 0, b, a, b, b, 0, b, b, a, b, 0, b, b, b, a, 0, 0, 0, 0, 0, 0};
	// one flat buffer per side; pairs longer than MAX_KSW_SEQ_LEN are chunked with the same offset on both (src/align.cc:46-53)
	std::vector<int> ql, tl, owner;
	std::vector<int64_t> qo, to;
	std::vector<uint8_t> qcodes, tcodes, qraw, traw;
	for (size_t i = 0; i < pairs.size(); ++i) {
		const std::string &fa = pairs[i].first, &fb = pairs[i].second;
		const size_t n = std::min(fa.size(), fb.size());
		for (size_t sp = 0; sp < n; sp += kMaxKswSeqLen) {
			const size_t la = std::min<size_t>(kMaxKswSeqLen, fa.size() - sp), lb = std::min<size_t>(kMaxKswSeqLen, fb.size() - sp);
			owner.push_back((int)i);
			ql.push_back((int)la); tl.push_back((int)lb);
			qo.push_back((int64_t)qcodes.size()); to.push_back((int64_t)tcodes.size());
			for (size_t k = 0; k < la; ++k) { qraw.push_back((uint8_t)fa[sp + k]); qcodes.push_back(align_dna(fa[sp + k])); }
			for (size_t k = 0; k < lb; ++k) { traw.push_back((uint8_t)fb[sp + k]); tcodes.push_back(align_dna(fb[sp + k])); }
		}
	}
	const int n = (int)owner.size();
	std::vector<ksw_extz_t> ez(n);
	std::vector<sd_stats_t> st(n);
	qcodes.push_back(0); tcodes.push_back(0); qraw.push_back(0); traw.push_back(0);       // never pass null buffers
	int rc = ksw_extz2_batch_flat(n, ql.data(), qo.data(), qcodes.data(), tl.data(), to.data(), tcodes.data(), 5, mat,
	                              (int8_t)p.gap_open, (int8_t)p.gap_extend, p.bandwidth, -1, 0, ez.data(), st.data(),
	                              qraw.data(), traw.data());
	if (rc) throw std::runtime_error(std::string("ksw_extz2_batch_flat: ") + ksw_b200_strerror(rc) + " -- " + ksw_b200_last_error());
	std::vector<Alignment> out(pairs.size());
	for (size_t i = 0; i < pairs.size(); ++i) { out[i].a = pairs[i].first; out[i].b = pairs[i].second; }
	for (int k = 0; k < n; ++k) {
		Alignment &al = out[owner[k]];
		for (int64_t c = 0; c < ez[k].n_cigar; ++c) {
			const int idx = ez[k].cigar[c] & 0xf, len = (int)(ez[k].cigar[c] >> 4);
			if (idx < 3) al.cigar.push_back({"MDI"[idx], len});                              // src/align.cc:58-63
		}
		add_stats(al.stats, st[k]);
		free(ez[k].cigar);
	}
	return out;
}

std::vector<Alignment> from_cigar_batch(const std::vector<std::pair<std::string, std::string>> &pairs,
                                        const std::vector<std::string> &cigars)
{
	if (pairs.size() != cigars.size()) throw std::runtime_error("from_cigar_batch: size mismatch");
	const int n = (int)pairs.size();
	std::vector<Alignment> out(n);
	std::vector<int64_t> coff(n), cn(n), ao(n), bo(n);
	std::vector<int> al(n), bl(n);
	std::vector<uint32_t> cbuf;
	std::vector<uint8_t> abuf, bbuf;
	for (int i = 0; i < n; ++i) {
		out[i].a = pairs[i].first; out[i].b = pairs[i].second;
		coff[i] = (int64_t)cbuf.size();
		int num = 0;
		for (char ch : cigars[i]) {                                                          // src/align.cc:94-103
			if (ch >= '0' && ch <= '9') num = 10 * num + (ch - '0');
			else if (ch == ';') continue;
			else {
				out[i].cigar.push_back({ch, num});
				const uint32_t op = ch == 'M' ? 0u : (ch == 'D' ? 1u : (ch == 'I' ? 2u : 3u));   // SEDEF 'D' = a only = ksw I
				cbuf.push_back((uint32_t)num << 4 | op);
				num = 0;
			}
		}
		cn[i] = (int64_t)cbuf.size() - coff[i];
		ao[i] = (int64_t)abuf.size(); bo[i] = (int64_t)bbuf.size();
		al[i] = (int)pairs[i].first.size(); bl[i] = (int)pairs[i].second.size();
		abuf.insert(abuf.end(), pairs[i].first.begin(), pairs[i].first.end());
		bbuf.insert(bbuf.end(), pairs[i].second.begin(), pairs[i].second.end());
	}
	cbuf.push_back(0); abuf.push_back(0); bbuf.push_back(0);
	std::vector<sd_stats_t> st(n);
	std::vector<int> status(n);
	int rc = sd_stats_from_cigar_batch_flat(n, coff.data(), cn.data(), cbuf.data(), al.data(), ao.data(), abuf.data(),
	                                       bl.data(), bo.data(), bbuf.data(), st.data(), status.data());
	if (rc) throw std::runtime_error(std::string("sd_stats_from_cigar_batch_flat: ") + ksw_b200_strerror(rc) + " -- " + ksw_b200_last_error());
	for (int i = 0; i < n; ++i) {
		if (status[i]) throw std::runtime_error("from_cigar_batch: CIGAR " + std::to_string(i) + " overruns a sequence (the reference asserts, src/align.cc:281-282)");
		out[i].stats = st[i];
	}
	return out;
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

std::vector<GuidedAlignment> align_chains_batch(const std::vector<ChainGuide> &chains, const AlignParams &p)
{
	// pass 1: walk every chain, queue its gap fills
	struct Fill { size_t chain; size_t step; char tail_op; int tail_len; };   // fill result is spliced in at `step`
	std::vector<std::pair<std::string, std::string>> reqs;
	std::vector<Fill> fills;
	for (size_t ci = 0; ci < chains.size(); ++ci) {
		const ChainGuide &cg = chains[ci];
		const std::vector<Anchor> &g = *cg.anchors;
		for (size_t k = 1; k < cg.guide_idx.size(); ++k) {
			const Anchor &pv = g[cg.guide_idx[k - 1]], &cu = g[cg.guide_idx[k]];
			const int qpe = pv.q + pv.l, rpe = pv.r + pv.l, qs = cu.q, rs = cu.r;
			const int qgap = qs - qpe, rgap = rs - rpe;
			if (qgap && rgap) {
				if (qgap <= 1000 && rgap <= 1000) {                                     // "close" hits, src/align.cc:233-236
					reqs.emplace_back(cg.qstr->substr(qpe, qgap), cg.rstr->substr(rpe, rgap));
					fills.push_back({ci, k, 0, 0});
				} else {                                                                // src/align.cc:237-246: ma1 is always taken
					const int ma = std::max(qgap, rgap), mi = std::min(qgap, rgap);
					reqs.emplace_back(cg.qstr->substr(qpe, mi), cg.rstr->substr(rpe, mi));
					fills.push_back({ci, k, qgap == mi ? 'I' : 'D', ma - mi});
				}
			}
		}
	}
	std::vector<Alignment> filled = align_batch(reqs, p);                              // ONE batched ksw_extz2 call
	// pass 2: stitch
	std::vector<GuidedAlignment> out(chains.size());
	std::vector<std::pair<std::string, std::string>> finals(chains.size());
	std::vector<std::string> final_cigars(chains.size());
	size_t fpos = 0;
	for (size_t ci = 0; ci < chains.size(); ++ci) {
		const ChainGuide &cg = chains[ci];
		GuidedAlignment &al = out[ci];
		if (cg.guide_idx.empty()) continue;                                             // src/align.cc:202-205
		const std::vector<Anchor> &g = *cg.anchors;
		const Anchor &a0 = g[cg.guide_idx[0]];
		al.start_a = a0.q; al.end_a = a0.q + a0.l; al.start_b = a0.r; al.end_b = a0.r + a0.l;
		al.cigar = {{'M', a0.l}};
		for (size_t k = 1; k < cg.guide_idx.size(); ++k) {
			const Anchor &pv = g[cg.guide_idx[k - 1]], &cu = g[cg.guide_idx[k]];
			const int qpe = pv.q + pv.l, rpe = pv.r + pv.l, qs = cu.q, rs = cu.r;
			const int qgap = qs - qpe, rgap = rs - rpe;
			al.end_a = cu.q + cu.l; al.end_b = cu.r + cu.l;
			if (qgap && rgap) {
				const Fill &f = fills[fpos];
				std::deque<std::pair<char, int>> gc = filled[fpos].cigar;
				if (f.tail_op) gc.push_back({f.tail_op, f.tail_len});
				append_cigar(al.cigar, gc);
				++fpos;
			} else if (qgap) append_cigar(al.cigar, {{'D', qgap}});
			else if (rgap) append_cigar(al.cigar, {{'I', rgap}});
			append_cigar(al.cigar, {{'M', cu.l}});
		}
		al.a = cg.qstr->substr(al.start_a, al.end_a - al.start_a);
		al.b = cg.rstr->substr(al.start_b, al.end_b - al.start_b);
		finals[ci] = {al.a, al.b};
		final_cigars[ci] = al.cigar_string();
	}
	// populate_nice_alignment for every stitched alignment: one statistics-from-CIGAR call
	std::vector<Alignment> st = from_cigar_batch(finals, final_cigars);
	for (size_t ci = 0; ci < chains.size(); ++ci) out[ci].stats = st[ci].stats;
	return out;
}

} // namespace sedef_b200
