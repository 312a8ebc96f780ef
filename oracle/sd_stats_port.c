/*
 * oracle/sd_stats_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar C restatement of the per-alignment statistics SEDEF derives from a CIGAR:
 *   - Alignment::populate_nice_alignment()        /root/reference/src/align.cc:274-315
 *     with ceq()                                   /root/reference/src/align.cc:29-35
 *   - the BEDPE stat loop of process()             /root/reference/src/stats_main.cc:231-271
 *   - align_helper()'s ksw-op -> "MDI" remap        /root/reference/src/align.cc:58-63
 *   - align_dna()                                   /root/reference/src/common.h:58-70,91
 *
 * Pinning: the known-answer record of SURVEY.md Appendix B.3 (Alignment(seq1,seq2) of
 * python/simulations.py:6-7) in tests/golden/ksw2_kat.json, plus records produced by the real
 * reference classes through oracle/ref_shim.cc (oracle/_ref/libsedef_ref.so) in
 * tests/golden/sd_stats_golden.json (generator: tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <string.h>
#include <ctype.h>
#include "../include/ksw2_b200.h"

/* src/common.h:58-70 (dna_align_lookup: ACGT/acgt -> 0..3, everything else 4) */
uint8_t oracle_align_dna(uint8_t c)
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}

/* src/align.cc:29-35 */
static int port_ceq(int a, int b)
{
	if (a == '-' || b == '-') return 0;
	if (toupper(a) == 'N' || toupper(b) == 'N') return 0;
	return toupper(a) == toupper(b);
}

/*
 * cigar: raw ksw ops ((len<<4)|op, op 0=M, 1=I (query only), 2=D (target only)) in forward
 * order.  a = query original-case bytes, b = target original-case bytes.
 * Op 3 stands for "any other op letter" of an Alignment(fa, fb, cigar) string: populate_nice_alignment treats it as
 * not-M (a gap run), not-D and not-I (consumes both strings), src/align.cc:283-305.  (ksw_extz2 itself never emits
 * ops >= 3, and align_helper would drop them, src/align.cc:61.)
 * Returns 0, or -1 if the CIGAR overruns a sequence on an M column (the reference asserts) or on an op-3 column
 * (where the reference reads past the string).
 */
int oracle_sd_stats(const uint32_t *cigar, int64_t n_cigar, const uint8_t *a, int alen,
                    const uint8_t *b, int blen, sd_stats_t *s)
{
	int64_t k; int ia = 0, ib = 0;
	memset(s, 0, sizeof(*s));
	for (k = 0; k < n_cigar; ++k) {
		int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4), i;
		char c;
		c = op < 3 ? "MDI"[op] : 'X';               /* ksw I -> 'D' (a only), ksw D -> 'I' (b only) */
		if (c != 'M') { s->gaps++; s->gap_bases += len; }      /* src/align.cc:300-305 */
		for (i = 0; i < len; ++i) {
			int ca, cb, ua, ub;
			if ((c == 'M' || c == 'X') && (ia >= alen || ib >= blen)) return -1;
			cb = (c != 'D') ? (ib < blen ? b[ib] : 0) : '-';
			ca = (c != 'I') ? (ia < alen ? a[ia] : 0) : '-';
			if (c != 'D') ib++;
			if (c != 'I') ia++;
			s->span++;
			/* populate_nice_alignment second loop, src/align.cc:306-314 */
			if (ca != '-' && cb != '-') {
				if (port_ceq(ca, cb)) s->matches++; else s->mismatches++;
			}
			/* stat loop, src/stats_main.cc:244-271 */
			ua = toupper(ca); ub = toupper(cb);
			s->indel_a += ua == '-';
			s->indel_b += ub == '-';
			s->matchB += ua != '-' && ua == ub;
			s->uppercaseA += (ca != '-' && toupper(ca) != 'N' && isupper(ca));
			s->uppercaseB += (cb != '-' && toupper(cb) != 'N' && isupper(cb));
			if (ua != '-' && ub != '-') {
				s->alnB += 1;
				if (ua != ub) {
					s->mismatchB += 1;
					if (ua == 'A' || ua == 'G') {
						s->transitionsB += ub == 'A' || ub == 'G';
						s->transversionsB += !(ub == 'A' || ub == 'G');
					} else {
						s->transitionsB += ub == 'C' || ub == 'T';
						s->transversionsB += !(ub == 'C' || ub == 'T');
					}
				} else if (isupper(ca) && isupper(cb)) {
					s->uppercaseMatches++;
				}
			}
		}
	}
	return 0;
}
