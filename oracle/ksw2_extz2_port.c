/*
 * oracle/ksw2_extz2_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain scalar C restatement of the reference's `ksw_extz2_sse`
 * (/root/reference/extern/ksw2_extz2_sse.cc:23-298, SSE4.1 code path) and of the ksw2 helpers it
 * uses (/root/reference/extern/ksw2.h:98-177).  It exists so that tests can check the CUDA path
 * (and the compiled reference in oracle/_ref/) against a readable executable specification.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product library never links or calls it.
 *
 * Pinning: this port is validated bit-for-bit against the compiled reference
 * (oracle/_ref/libksw2_ref.so, built by oracle/Makefile from the untouched reference source)
 * by tests/test_oracle.py (differential fuzz over lengths, bands, z-drop and every flag) and
 * against the known-answer vectors of SURVEY.md Appendix B.3 / tests/golden/ksw2_kat.json.
 *
 * The restatement deliberately reproduces the quirks that make banded results depend on the
 * 16-lane SSE block structure:
 *   - one zeroed allocation [u|v|x|y|s|sf|qr|slack], indexed by slot t, persistent across
 *     anti-diagonals, including the out-of-range reads/writes of the 16-byte score fill;
 *   - DP over the 16-rounded slot range, score fill over 16-byte chunks from st0;
 *   - neighbour reads use the previous diagonal's values;
 *   - int8 wrap-around arithmetic with the signed/unsigned min/max mix of the SSE4.1 path;
 *   - the 4-lane arg-max tie-break, mte_q from the rounded `en`, H[en0] from old H[en0-1].
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/ksw2_b200.h"

static inline int8_t s8(int v) { return (int8_t)(uint8_t)(v & 0xff); }          /* wrap like _mm_*_epi8 */
static inline int8_t max_s8(int8_t a, int8_t b) { return a > b ? a : b; }       /* _mm_max_epi8 */
static inline int8_t max_u8(int8_t a, int8_t b) { return (uint8_t)a > (uint8_t)b ? a : b; } /* _mm_max_epu8 */
static inline int8_t min_u8(int8_t a, int8_t b) { return (uint8_t)a < (uint8_t)b ? a : b; } /* _mm_min_epu8 */

/* extern/ksw2.h:153-159 */
static void port_reset_extz(ksw_extz_t *ez)
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0; ez->score = ez->mqe = ez->mte = KSW_NEG_INF;
	ez->n_cigar = 0; ez->m_cigar = 0; ez->zdropped = 0;
	ez->cigar = 0;
}

/* extern/ksw2.h:161-177 (is_rot == 1: a = r, b = t) */
static int port_apply_zdrop(ksw_extz_t *ez, int32_t H, int r, int t, int zdrop, int8_t e)
{
	if (H > (int32_t)ez->max) {
		ez->max = H; ez->max_t = t; ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l;
		l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && (int32_t)ez->max - H > zdrop + l * e) {
			ez->zdropped = 1;
			return 1;
		}
	}
	return 0;
}

/* extern/ksw2.h:98-111 */
static uint32_t *port_push_cigar(int64_t *n_cigar, int64_t *m_cigar, uint32_t *cigar, uint32_t op, int len)
{
	if (*n_cigar == 0 || op != (cigar[(*n_cigar) - 1] & 0xf)) {
		if (*n_cigar == *m_cigar) {
			if (*m_cigar > 2 * 1024LL * 1024LL * 1024LL) *m_cigar = *m_cigar + 1024LL * 1024LL * 1024LL;
			else *m_cigar = *m_cigar ? (*m_cigar) << 1 : 4;
			cigar = (uint32_t *)realloc(cigar, (size_t)(*m_cigar) << 2);
			if (!cigar) abort();
		}
		cigar[(*n_cigar)++] = (uint32_t)len << 4 | op;
	} else cigar[(*n_cigar) - 1] += (uint32_t)len << 4;
	return cigar;
}

/* extern/ksw2.h:117-151, specialised to is_rot = 1, with_N = 0 */
static void port_backtrack(int is_rev, const uint8_t *p, const int *off, const int *off_end, int n_col,
                           int i0, int j0, int64_t *m_cigar_, int64_t *n_cigar_, uint32_t **cigar_)
{
	int64_t n_cigar = 0, m_cigar = *m_cigar_, i = i0, j = j0, r, state = 0;
	uint32_t *cigar = *cigar_, tmp;
	while (i >= 0 && j >= 0) {
		int force_state = -1;
		r = i + j;
		if (i < off[r]) force_state = 2;
		if (off_end && i > off_end[r]) force_state = 1;
		tmp = force_state < 0 ? p[r * n_col + i - off[r]] : 0;
		if (state == 0) state = tmp & 7;
		else if (!(tmp >> (state + 2) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (force_state >= 0) state = force_state;
		if (state == 0) { cigar = port_push_cigar(&n_cigar, &m_cigar, cigar, 0, 1); --i; --j; }
		else if (state == 1 || state == 3) { cigar = port_push_cigar(&n_cigar, &m_cigar, cigar, 2, 1); --i; }
		else { cigar = port_push_cigar(&n_cigar, &m_cigar, cigar, 1, 1); --j; }
	}
	if (i >= 0) cigar = port_push_cigar(&n_cigar, &m_cigar, cigar, 2, (int)(i + 1));
	if (j >= 0) cigar = port_push_cigar(&n_cigar, &m_cigar, cigar, 1, (int)(j + 1));
	if (!is_rev)
		for (i = 0; i < n_cigar >> 1; ++i) {
			tmp = cigar[i]; cigar[i] = cigar[n_cigar - 1 - i]; cigar[n_cigar - 1 - i] = tmp;
		}
	*m_cigar_ = m_cigar; *n_cigar_ = n_cigar; *cigar_ = cigar;
}

/* Counters that tests read to learn whether a pair ever left the "no clamp / no wrap" domain the
 * CUDA kernel's 16-bit lanes assume ("exactness domain"). */
typedef struct {
	int64_t n_clamp;      /* cells where max(z,a,b) exceeded max_sc (the _mm_min_epu8 was binding) */
	int64_t n_wrap;       /* cells where an int8 add/sub left [-128,127] before wrapping */
	int64_t n_diag;       /* anti-diagonals processed */
	int64_t n_cells;      /* in-band cells (en0-st0+1 summed over processed diagonals) */
} oracle_diag_t;

static oracle_diag_t g_last_diag;
void oracle_last_diag(oracle_diag_t *d) { *d = g_last_diag; }

#define WRAPCHK(v) do { int v_ = (v); if (v_ < -128 || v_ > 127) ++dg.n_wrap; } while (0)

void oracle_ksw_extz2(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                      int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                      ksw_extz_t *ez)
{
	int r, t, qe = q + e, n_col_, *off = 0, *off_end = 0, tlen_, qlen_, last_st, last_en, max_sc, min_sc;
	int with_cigar = !(flag & KSW_EZ_SCORE_ONLY), approx_max = !!(flag & KSW_EZ_APPROX_MAX);
	int32_t *H = 0, H0 = 0, last_H0_t = 0;
	uint8_t *mem, *u8, *v8, *x8, *y8, *s8a, *sf, *qr, *p = 0;
	int8_t qe2, maxsc8;
	size_t T, mem_bytes;
	oracle_diag_t dg = {0, 0, 0, 0};
	(void)km;

	port_reset_extz(ez);
	g_last_diag = dg;
	if (m <= 0 || qlen <= 0 || tlen <= 0) return;                                  /* :57 */

	qe2 = s8((q + e) * 2);
	maxsc8 = s8(mat[0] + (q + e) * 2);
	if (w < 0) w = tlen > qlen ? tlen : qlen;                                       /* :71 */
	tlen_ = (tlen + 15) / 16;
	n_col_ = qlen < tlen ? qlen : tlen;
	n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;                     /* :74-75 */
	qlen_ = (qlen + 15) / 16;
	for (t = 1, max_sc = mat[0], min_sc = mat[1]; t < m * m; ++t) {
		max_sc = max_sc > mat[t] ? max_sc : mat[t];
		min_sc = min_sc < mat[t] ? min_sc : mat[t];
	}
	if (-min_sc > 2 * (q + e)) return;                                              /* :81 */

	T = (size_t)tlen_ * 16;
	mem_bytes = ((size_t)tlen_ * 6 + qlen_ + 1) * 16;                               /* :83 */
	mem = (uint8_t *)calloc(mem_bytes + 64, 1);  /* +64: the reference's reads stay inside its own block; slack only guards this port */
	u8 = mem; v8 = u8 + T; x8 = v8 + T; y8 = x8 + T; s8a = y8 + T; sf = s8a + T; qr = sf + T; /* :84-85 */
	if (!approx_max) {
		H = (int32_t *)malloc(T * 4);
		for (t = 0; t < (int)T; ++t) H[t] = KSW_NEG_INF;
	}
	if (with_cigar) {
		p = (uint8_t *)malloc(((uint64_t)(qlen + tlen - 1) * n_col_ + 1) * 16);
		off = (int *)malloc((uint64_t)(qlen + tlen - 1) * sizeof(int) * 2);
		off_end = off + qlen + tlen - 1;
	}
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];                         /* :97 */
	memcpy(sf, target, tlen);                                                        /* :98 */

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		int8_t x1, v1;
		uint8_t *qrr = qr + (qlen - 1 - r);
		/* band boundaries :106-115 */
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) { ez->zdropped = 1; break; }
		st0 = st; en0 = en;
		st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
		++dg.n_diag; dg.n_cells += en0 - st0 + 1;
		/* boundary conditions :117-122 */
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) { x1 = (int8_t)x8[st - 1]; v1 = (int8_t)v8[st - 1]; }
			else x1 = v1 = 0;
		} else { x1 = 0; v1 = r ? q : 0; }
		if (en >= r) { y8[r] = 0; u8[r] = r ? (uint8_t)q : 0; }
		/* score fill :124-142 */
		if (!(flag & KSW_EZ_GENERIC_SC)) {
			for (t = st0; t <= en0; t += 16) {
				int k;
				uint8_t tmp[16];
				for (k = 0; k < 16; ++k) {             /* loadu sf[t..], qrr[t..] BEFORE the store (they may alias s) */
					uint8_t sq = sf[t + k], sq2 = qrr[t + k];
					int wild = (sq == (uint8_t)(m - 1)) || (sq2 == (uint8_t)(m - 1));
					tmp[k] = wild ? 0 : (uint8_t)(sq == sq2 ? mat[0] : mat[1]);
				}
				memcpy(s8a + t, tmp, 16);              /* storeu: may spill past s into sf[0..] */
			}
		} else {
			for (t = st0; t <= en0; ++t) s8a[t] = (uint8_t)mat[sf[t] * m + qrr[t]];
		}
		/* core loop :144-221 */
		if (with_cigar) { off[r] = st; off_end[r] = en; }
		{
			int8_t px = x1, pv = v1;
			uint8_t *pr = with_cigar ? p + ((size_t)r * n_col_ - st / 16) * 16 : 0;
			for (t = st; t <= en; ++t) {
				int8_t z, a, b, xt1, vt1, ut, d, zc;
				xt1 = px; vt1 = pv; px = (int8_t)x8[t]; pv = (int8_t)v8[t];
				/* :102,144-145: x1 / v1 are int8_t and go through _mm_cvtsi32_si128(), so a carry byte >= 0x80 is sign-extended
				 * into lanes 1..3 of x1_ / v1_, and the _mm_or_si128 of the FIRST block (:30,34) turns x[t-1] / v[t-1] of slots
				 * st+1..st+3 into 0xff.  x is always in [0,127]; v reaches 128+ once 2(q+e) + match exceeds 127. */
				if (t > st && t <= st + 3) {
					if (x1 < 0) xt1 = (int8_t)0xff;
					if (v1 < 0) vt1 = (int8_t)0xff;
				}
				WRAPCHK((int8_t)s8a[t] + qe2);  z = s8((int8_t)s8a[t] + qe2);
				WRAPCHK(xt1 + vt1);             a = s8(xt1 + vt1);
				ut = (int8_t)u8[t];
				WRAPCHK((int8_t)y8[t] + ut);    b = s8((int8_t)y8[t] + ut);
				if (!with_cigar) {
					z = max_s8(z, a);
					d = 0;
				} else if (!(flag & KSW_EZ_RIGHT)) {
					d = a > z ? 1 : 0;
					z = max_s8(z, a);
					d = b > z ? 2 : d;
				} else {
					d = z > a ? 0 : 1;
					z = max_s8(z, a);
					d = z > b ? d : 2;
				}
				z = max_u8(z, b);
				zc = min_u8(z, maxsc8);
				if (zc != z) ++dg.n_clamp;
				z = zc;
				WRAPCHK(z - vt1); u8[t] = (uint8_t)s8(z - vt1);
				WRAPCHK(z - ut);  v8[t] = (uint8_t)s8(z - ut);
				WRAPCHK(z - q);   z = s8(z - q);
				WRAPCHK(a - z);   a = s8(a - z);
				WRAPCHK(b - z);   b = s8(b - z);
				if (!with_cigar) {
					x8[t] = (uint8_t)max_s8(a, 0);
					y8[t] = (uint8_t)max_s8(b, 0);
				} else if (!(flag & KSW_EZ_RIGHT)) {
					x8[t] = a > 0 ? (uint8_t)a : 0; d |= a > 0 ? 0x08 : 0;
					y8[t] = b > 0 ? (uint8_t)b : 0; d |= b > 0 ? 0x10 : 0;
					pr[t] = (uint8_t)d;
				} else {
					x8[t] = 0 > a ? 0 : (uint8_t)a; d |= 0 > a ? 0 : 0x08;
					y8[t] = 0 > b ? 0 : (uint8_t)b; d |= 0 > b ? 0 : 0x10;
					pr[t] = (uint8_t)d;
				}
			}
		}
		if (!approx_max) {                                                          /* :222-267 */
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u8[en0] - qe : H[en0] + v8[en0] - qe;
				max_t = en0;
				for (i = 0; i < 4; ++i) { HH[i] = max_H; tt[i] = max_t; }
				for (t = st0; t < en1; t += 4) {
					for (i = 0; i < 4; ++i) {
						int32_t H1 = H[t + i] + (int32_t)v8[t + i] - qe;
						H[t + i] = H1;
						if (H1 > HH[i]) { HH[i] = H1; tt[i] = t; }
					}
				}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) { max_H = HH[i]; max_t = tt[i] + i; }
				for (; t < en0; ++t) {
					H[t] += (int32_t)v8[t] - qe;
					if (H[t] > max_H) { max_H = H[t]; max_t = t; }
				}
			} else { H[0] = v8[0] - qe - qe; max_H = H[0]; max_t = 0; }
			if (en0 == tlen - 1 && H[en0] > ez->mte) { ez->mte = H[en0]; ez->mte_q = r - en; }
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) { ez->mqe = H[st0]; ez->mqe_t = st0; }
			if (port_apply_zdrop(ez, max_H, r, max_t, zdrop, e)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else {                                                                     /* :268-284 */
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v8[last_H0_t] - qe;
					int32_t d1 = u8[last_H0_t + 1] - qe;
					if (d0 > d1) H0 += d0;
					else { H0 += d1; ++last_H0_t; }
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v8[last_H0_t] - qe;
				} else {
					++last_H0_t; H0 += u8[last_H0_t] - qe;
				}
				if ((flag & KSW_EZ_APPROX_DROP) && port_apply_zdrop(ez, H0, r, last_H0_t, zdrop, e)) break;
			} else { H0 = v8[0] - qe - qe; last_H0_t = 0; }
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
		last_st = st; last_en = en;
	}
	free(mem);
	if (!approx_max) free(H);
	if (with_cigar) {                                                                /* :290-297 */
		int rev_cigar = !!(flag & KSW_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & KSW_EZ_EXTZ_ONLY))
			port_backtrack(rev_cigar, p, off, off_end, n_col_ * 16, tlen - 1, qlen - 1, &ez->m_cigar, &ez->n_cigar, &ez->cigar);
		else if (ez->max_t >= 0 && ez->max_q >= 0)
			port_backtrack(rev_cigar, p, off, off_end, n_col_ * 16, ez->max_t, ez->max_q, &ez->m_cigar, &ez->n_cigar, &ez->cigar);
		free(p); free(off);
	}
	g_last_diag = dg;
}

/* In-band cell count when every anti-diagonal is processed. */
int64_t oracle_count_cells(int qlen, int tlen, int w)
{
	int64_t c = 0; int r;
	if (qlen <= 0 || tlen <= 0) return 0;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	for (r = 0; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) break;
		c += en - st + 1;
	}
	return c;
}
