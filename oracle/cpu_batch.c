/*
 * oracle/cpu_batch.c -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * OpenMP batch driver used (a) by tests to run many pairs through the compiled reference or the
 * scalar port in one call and (b) by bench.py's cpu_baseline / `--impl reference` legs to time the
 * reference's own ksw_extz2_sse on the GPU box's host cores (the reference
 * has no threaded ksw2 path, so the driver is `#pragma omp parallel for schedule(dynamic)` over
 * pairs, one call + free(ez.cigar) per pair, exactly as align_helper does per call,
 * /root/reference/src/align.cc:49-65).
 *
 * Compiled twice by oracle/Makefile:
 *   -DKSW_FN=ksw_extz2_sse      -> oracle/_ref/libksw2_ref.so   (reference object linked in)
 *   -DKSW_FN=oracle_ksw_extz2   -> oracle/liboracle_port.so     (scalar port)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/ksw2_b200.h"

void KSW_FN(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
            const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag, ksw_extz_t *ez);

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

/* Run n pairs.  If ez != NULL the records (with malloc'd cigars) are returned to the caller,
 * else cigars are freed immediately (timing mode).  Returns wall seconds. */
double cpu_batch_run(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                     const int *tlen, const int64_t *toff, const uint8_t *tbuf,
                     int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                     ksw_extz_t *ez, int nthreads)
{
	double t0;
	int i;
#ifdef _OPENMP
	if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
	t0 = now_s();
#pragma omp parallel for schedule(dynamic, 4)
	for (i = 0; i < n; ++i) {
		ksw_extz_t loc;
		KSW_FN(0, qlen[i], qbuf + qoff[i], tlen[i], tbuf + toff[i], m, mat, q, e, w, zdrop, flag, &loc);
		if (ez) ez[i] = loc;
		else free(loc.cigar);
	}
	return now_s() - t0;
}

int cpu_batch_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

void cpu_batch_free_cigars(int n, ksw_extz_t *ez)
{
	int i;
	for (i = 0; i < n; ++i) { free(ez[i].cigar); ez[i].cigar = 0; }
}
