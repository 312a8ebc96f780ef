/* oracle/ksw_record.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * A recording stand-in for ksw_extz2_sse: linked under the UNMODIFIED reference objects (oracle/_ref/libsedef_ref_rec.so) it
 * lets a CPU test read the exact (qlen, tlen, pointer) sequence of kernel calls the reference's align_helper makes
 * (src/align.cc:46-57) -- i.e. its 60 kbp chunk arithmetic -- without running a 60 000 x 60 000 DP.  It aligns nothing:
 * every call returns the reset record. */
#include <stdint.h>
#include <string.h>
#include "../include/ksw2_b200.h"

#define REC_CAP 256
static int rec_n;
static int rec_qlen[REC_CAP], rec_tlen[REC_CAP], rec_w[REC_CAP], rec_zdrop[REC_CAP], rec_flag[REC_CAP];
static const uint8_t *rec_q[REC_CAP], *rec_t[REC_CAP];

void ksw_extz2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                   int8_t q, int8_t e, int w, int zdrop, int flag, ksw_extz_t *ez)
{
	(void)km; (void)m; (void)mat; (void)q; (void)e;
	if (rec_n < REC_CAP) {
		rec_qlen[rec_n] = qlen; rec_tlen[rec_n] = tlen; rec_q[rec_n] = query; rec_t[rec_n] = target;
		rec_w[rec_n] = w; rec_zdrop[rec_n] = zdrop; rec_flag[rec_n] = flag;
	}
	++rec_n;
	memset(ez, 0, sizeof(*ez));
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->score = ez->mqe = ez->mte = KSW_NEG_INF;
}
void ksw_record_reset(void) { rec_n = 0; }
int ksw_record_count(void) { return rec_n; }
/* call k: lengths, and the distance of its query / target pointers from those of call 0 (the chunk offset SP) */
int ksw_record_get(int k, int *qlen, int *tlen, int64_t *q_off, int64_t *t_off, int *w, int *zdrop, int *flag)
{
	if (k < 0 || k >= rec_n || k >= REC_CAP) return -1;
	*qlen = rec_qlen[k]; *tlen = rec_tlen[k];
	*q_off = (int64_t)(rec_q[k] - rec_q[0]); *t_off = (int64_t)(rec_t[k] - rec_t[0]);
	*w = rec_w[k]; *zdrop = rec_zdrop[k]; *flag = rec_flag[k];
	return 0;
}
