/*
 * oracle/ksw_redirect.c -- TEST INFRASTRUCTURE.
 * The one-line binding of INTEGRATION.md section 1, as a link-time shim: the reference's own align-stage objects
 * (align.cc, chain.cc, refine.cc ... compiled unmodified from /root/reference) call `ksw_extz2_sse`
 * (src/align.cc:49); this definition forwards every call to the product library's drop-in `ksw_extz2_b200`.
 * Linked into oracle/_ref/libsedef_ref_b200.so INSTEAD of the reference's extern/ksw2_extz2_sse.cc, so that tests
 * can run the reference's fast_align() with the CUDA kernel underneath and demand identical hits and CIGARs.
 */
#include <stdint.h>
#include "../include/ksw2_b200.h"

static long g_calls = 0;
long ksw_redirect_calls(void) { return g_calls; }

void ksw_extz2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                   int8_t q, int8_t e, int w, int zdrop, int flag, ksw_extz_t *ez)
{
	++g_calls;
	ksw_extz2_b200(km, qlen, query, tlen, target, m, mat, q, e, w, zdrop, flag, ez);
}
