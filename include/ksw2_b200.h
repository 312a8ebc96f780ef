/*
 * ksw2_b200.h -- C ABI of the B200-native batched banded-alignment engine.
 *
 * This is the drop-in boundary for SEDEF's one data-parallel hot path: the ksw2
 * `ksw_extz2_sse` call made by `align_helper()` (reference src/align.cc:49-57), plus the
 * per-alignment statistics that `Alignment::populate_nice_alignment()` (src/align.cc:274-315)
 * and the BEDPE stat loop of `process()` (src/stats_main.cc:244-271) derive from the CIGAR.
 *
 * Everything here is plain C: pointers, sizes, POD structs. No torch / CUDA types.
 * All entry points need a CUDA device; there is NO CPU fallback. A call made without a
 * usable device returns KSW_B200_ERR_NO_DEVICE (batch API) or is fatal (ksw2-compatible
 * single-pair API, which has no error channel: see ksw_b200_set_fatal_handler).
 */
#ifndef KSW2_B200_H_
#define KSW2_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ksw2 call surface (reference extern/ksw2.h:6-30) ------------------------------- */

#ifndef KSW_NEG_INF
#define KSW_NEG_INF -0x40000000            /* extern/ksw2.h:6 */
#endif
#ifndef KSW_EZ_SCORE_ONLY
#define KSW_EZ_SCORE_ONLY  0x01            /* extern/ksw2.h:8  : no traceback / CIGAR */
#define KSW_EZ_RIGHT       0x02            /* extern/ksw2.h:9  : right-align gaps */
#define KSW_EZ_GENERIC_SC  0x04            /* extern/ksw2.h:10 : full m*m matrix (m <= 8) */
#define KSW_EZ_APPROX_MAX  0x08            /* extern/ksw2.h:11 : approximate max (one tracked H), no mqe/mte */
#define KSW_EZ_APPROX_DROP 0x10            /* extern/ksw2.h:12 : only meaningful with APPROX_MAX */
#define KSW_EZ_EXTZ_ONLY   0x40            /* extern/ksw2.h:13 : always trace back from (max_t,max_q) */
#define KSW_EZ_REV_CIGAR   0x80            /* extern/ksw2.h:14 : emit the CIGAR end->start */
#endif

#ifndef KSW2_H_
/* Layout-identical to the reference's result record (extern/ksw2.h:22-30, 56 bytes). */
typedef struct {
	uint32_t max:31, zdropped:1;
	int max_q, max_t;      /* max extension coordinate */
	int mqe, mqe_t;        /* max score when reaching the end of query */
	int mte, mte_q;        /* max score when reaching the end of target */
	int score;             /* max score reaching both ends; may be KSW_NEG_INF */
	uint32_t *cigar;       /* malloc()'d by the callee, free()'d by the caller (src/align.cc:65) */
	int64_t m_cigar, n_cigar;
} ksw_extz_t;
#endif

/* ---- SD statistics record ("Alignment::from_cigar" statistics) ----------------------- */
/*
 * All integer fields SEDEF derives from one alignment.  `a` is the query string (first
 * argument of Alignment(fa, fb), src/align.cc:76), `b` the target.  ksw op I(1) consumes the
 * query and is SEDEF's 'D'; ksw op D(2) consumes the target and is SEDEF's 'I'
 * (src/align.cc:58-63).  Computed on the GPU from the CIGAR and the ORIGINAL-CASE bytes.
 */
typedef struct {
	int32_t span;              /* alignment.size(): number of columns          (src/align.h:79)  */
	int32_t gaps;              /* number of non-M CIGAR runs                   (src/align.cc:300-305) */
	int32_t gap_bases;         /* sum of non-M run lengths                     (src/align.cc:300-305) */
	int32_t matches;           /* gap-free columns with ceq(a,b) (N never equal) (src/align.cc:29-35,306-314) */
	int32_t mismatches;        /* gap-free columns without ceq                 (src/align.cc:306-314) */
	int32_t indel_a;           /* columns with a == '-'                        (src/stats_main.cc:247) */
	int32_t indel_b;           /* columns with b == '-'                        (src/stats_main.cc:248) */
	int32_t alnB;              /* gap-free columns                             (src/stats_main.cc:257) */
	int32_t matchB;            /* toupper(a)==toupper(b), a not gap (N==N counts) (src/stats_main.cc:249) */
	int32_t mismatchB;         /*                                              (src/stats_main.cc:259) */
	int32_t transitionsB;      /*                                              (src/stats_main.cc:260-266) */
	int32_t transversionsB;    /*                                              (src/stats_main.cc:260-266) */
	int32_t uppercaseA;        /* non-gap, non-N, isupper(a)                   (src/stats_main.cc:250-252) */
	int32_t uppercaseB;        /*                                              (src/stats_main.cc:253-255) */
	int32_t uppercaseMatches;  /* equal columns with both bases upper-case     (src/stats_main.cc:267-269) */
	int32_t reserved;          /* pad to 64 bytes; always 0 */
} sd_stats_t;

/* Floating-point BEDPE fields, derived ON THE HOST from the integers above with the same
 * double-precision expressions as src/stats_main.cc:273-283,297-299. */
typedef struct {
	double fracMatch, fracMatchIndel, jcK, k2K, errorScaled, filter_score;
	double gap_error, mismatch_error, total_error;    /* src/align.h:84-92 */
} sd_stats_fp_t;
void sd_stats_derive_fp(const sd_stats_t *s, sd_stats_fp_t *out);

/* ---- error codes ---------------------------------------------------------------------- */
enum {
	KSW_B200_OK = 0,
	KSW_B200_ERR_NO_DEVICE   = -1,   /* no CUDA device / driver: there is no CPU fallback */
	KSW_B200_ERR_CUDA        = -2,   /* a CUDA runtime call failed (see ksw_b200_last_error) */
	KSW_B200_ERR_DOMAIN      = -3,   /* reserved: the kernels reproduce the reference's int8 wrap-around, so there is
	                                    no scoring domain restriction */
	KSW_B200_ERR_UNSUPPORTED = -4,   /* an alphabet with m > 8 */
	KSW_B200_ERR_TOO_WIDE    = -5,   /* a pair needs more live slots per anti-diagonal than the widest kernel */
	KSW_B200_ERR_NOMEM       = -6,   /* host or device allocation failed */
	KSW_B200_ERR_ARG         = -7,   /* bad argument: n<0, NULL pointers, a sequence symbol >= 8 (>= m with KSW_EZ_GENERIC_SC,
	                                    where the reference would index past mat[]) */
	KSW_B200_ERR_INEXACT     = -8    /* reserved */
};
const char *ksw_b200_strerror(int code);
const char *ksw_b200_last_error(void);   /* thread-local detail string of the last failure */

/* ---- context ---------------------------------------------------------------------------- */
/* Bind the engine to `ndev` devices starting at `first_dev` (ndev<=0: all visible devices).
 * Idempotent; the first batch call initialises lazily with (0, all).  Returns the number of
 * devices bound or a negative error code. */
int  ksw_b200_init(int first_dev, int ndev);
void ksw_b200_destroy(void);
int  ksw_b200_num_devices(void);
/* Host threads used for packing / gathering (0 = OpenMP default). */
void ksw_b200_set_host_threads(int n);
/* Upper bound on rounded live slots per anti-diagonal a pair may need (SURVEY App. C:
 * 16*n_col_, or 16*ceil(tlen/16) if smaller). */
int  ksw_b200_max_slots(void);

/* ---- page-locked host memory ---------------------------------------------------------------- */
/* Sequence buffers that live in page-locked memory are copied to the device IN PLACE (no host pass over the bytes);
 * pageable buffers are staged through the engine's own pinned buffers.  Allocate inputs here, or page-lock existing
 * memory (e.g. the genome the caller already holds) once. */
void *ksw_b200_host_alloc(size_t bytes);
void  ksw_b200_host_free(void *p);
int   ksw_b200_host_register(void *p, size_t bytes);
int   ksw_b200_host_unregister(void *p);

/* ---- single pair: identical signature and ownership to ksw_extz2_sse -------------------- */
/* Replaces extern/ksw2.h:50 / extern/ksw2_extz2_sse.cc:23.  `km` is ignored, as in the
 * reference build (HAVE_KALLOC undefined, extern/ksw2.h:88-96).
 * ksw_extz2_sse has no error channel, so neither has this: a request the engine cannot serve (no device, a pair wider
 * than ksw_b200_max_slots()) is FATAL -- reported on stderr, then abort() -- unless a handler is installed; if the handler
 * returns, `ez` holds the reset record. */
typedef void (*ksw_b200_fatal_fn)(int code, const char *detail);
void ksw_b200_set_fatal_handler(ksw_b200_fatal_fn fn);     /* NULL restores the default (abort) */
void ksw_extz2_b200(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                    int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                    ksw_extz_t *ez);

/* ---- batch ------------------------------------------------------------------------------- */
/*
 * n independent pairs with shared scoring parameters.  Pairs are bucketed by live-slot width
 * and length, sharded over the bound devices by a length-balanced greedy (LPT) partition on
 * in-band cell counts, and gathered by original index.  ez[i] is fully overwritten exactly
 * as ksw_extz2_sse would; ez[i].cigar is malloc()'d (caller frees) unless SCORE_ONLY.
 * stats (optional, may be NULL): per-pair SD statistics; needs q_raw/t_raw = original-case
 * ASCII bytes of the same lengths (if NULL the encoded bytes are decoded as "ACGTN").
 * Pairs with qlen<=0 || tlen<=0 get the reset record (extern/ksw2_extz2_sse.cc:56-57).
 * Returns KSW_B200_OK or an error code; on error no ez[i].cigar is left allocated.
 */
int ksw_extz2_batch(int n, const int *qlen, const uint8_t *const *query,
                    const int *tlen, const uint8_t *const *target,
                    int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                    ksw_extz_t *ez, sd_stats_t *stats,
                    const uint8_t *const *q_raw, const uint8_t *const *t_raw);

/* Sequences as ORIGINAL-CASE BYTES ONLY: pass query = target = NULL (qbuf = tbuf = NULL in the flat forms) together with
 * q_raw / t_raw and m = 5.  The codes are then derived ON THE DEVICE with SEDEF's align_dna (src/common.h:58-70,91:
 * A C G T, either case -> 0..3, anything else -> 4), which is what Alignment(fa, fb) does before it calls ksw2
 * (src/align.cc:79-82) -- one byte per base crosses PCIe instead of two. */

/* H2D / D2H bytes and kernel launches of the last one-shot batch call made by this thread. */
void ksw_b200_last_call_io(int64_t *h2d, int64_t *d2h, int *launches);

/* Convenience for batch callers: free() every ez[i].cigar of a result array and clear the fields. */
void ksw_b200_free_cigars(ksw_extz_t *ez, int n);

/* Same, with all sequences in two flat arrays (offsets in bytes; q_raw/t_raw share the
 * offsets).  This is the zero-gather form the align-stage driver and the Python wrapper use. */
int ksw_extz2_batch_flat(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                         const int *tlen, const int64_t *toff, const uint8_t *tbuf,
                         int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                         ksw_extz_t *ez, sd_stats_t *stats,
                         const uint8_t *q_raw_buf, const uint8_t *t_raw_buf);

/* ---- arena output ("one arena + offsets with an explicit free") ------------------------------ */
/* Same computation as ksw_extz2_batch_flat; the results stay in ONE page-locked arena owned by `*out`:
 * ksw_b200_result_ez() is the ksw_extz_t array in the caller's order, every record exactly as ksw_extz2_sse fills it
 * (m_cigar = the capacity ksw_push_cigar would have grown to), except that ez[i].cigar points INTO the arena -- do not
 * free() or realloc() it; ksw_b200_result_free() releases everything at once.  The device writes the final records, so the
 * host side of a call is three DMA copies and no per-pair work (no malloc, no gather loop). */
typedef struct ksw_b200_result ksw_b200_result_t;
int ksw_extz2_batch_arena(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                          const int *tlen, const int64_t *toff, const uint8_t *tbuf,
                          int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                          int want_stats, const uint8_t *q_raw_buf, const uint8_t *t_raw_buf,
                          ksw_b200_result_t **out);
const ksw_extz_t *ksw_b200_result_ez(const ksw_b200_result_t *r);        /* [n] */
const sd_stats_t *ksw_b200_result_stats(const ksw_b200_result_t *r);     /* [n], or NULL when not requested / SCORE_ONLY */
/* Per pair two ints computed on the same traceback walk as the statistics: the maximum-suffix / maximum-prefix scans of
 * Alignment::trim_front / trim_back (src/align.cc:343-456) with the ALIGNMENT scoring (mat[0], mat[1], gap open, gap extend):
 *   [2i]   trim_front's max_i = leading alignment columns to drop; -1 when no suffix scores >= 0 (the reference then keeps its
 *          initial max_i = a.size(), sic)
 *   [2i+1] columns trim_back keeps (its max_i + 1); -1 when no prefix scores >= 0 (the alignment is cleared)
 * NULL when statistics were not requested. */
const int32_t *ksw_b200_result_trims(const ksw_b200_result_t *r);
int  ksw_b200_result_count(const ksw_b200_result_t *r);
void ksw_b200_result_io(const ksw_b200_result_t *r, int64_t *h2d, int64_t *d2h, int *launches);
void ksw_b200_result_free(ksw_b200_result_t *r);
/* Position-independent copy into CALLER-OWNED buffers (e.g. a shared-memory segment: the host-side gather of a
 * one-process-per-GPU deployment).  Record i goes to ez_dst[index ? index[i] : i] (its statistics likewise, stats_dst may be
 * NULL); CIGAR words are appended to cigar_dst in the arena's order and ez.cigar holds the WORD OFFSET cigar_base + position
 * instead of a pointer.  Returns the CIGAR words written, or a negative error code (cigar_cap too small: nothing is copied). */
int64_t ksw_b200_result_export(const ksw_b200_result_t *r, ksw_extz_t *ez_dst, sd_stats_t *stats_dst,
                               uint32_t *cigar_dst, int64_t cigar_cap, int64_t cigar_base, const int64_t *index);

/* ---- Alignment(fa, fb, cigar): statistics from existing CIGARs ------------------------------ */
/* Replaces src/align.cc:90-105 + populate_nice_alignment + the stat loop for `sedef stats generate`
 * (src/stats_main.cc:224,244-271): n alignments, raw ksw ops ((len<<4)|op, 0=M, 1=I query only, 2=D target only)
 * in one flat buffer, a/b = original-case bytes.  status[i] = -1 when the CIGAR overruns a sequence on an
 * M column (the reference asserts, src/align.cc:281-282), else 0. */
int sd_stats_from_cigar_batch_flat(int n, const int64_t *cig_off, const int64_t *n_cigar, const uint32_t *cig_buf,
                                   const int *alen, const int64_t *aoff, const uint8_t *abuf,
                                   const int *blen, const int64_t *boff, const uint8_t *bbuf,
                                   sd_stats_t *out, int *status);

/* ---- anchors ------------------------------------------------------------- */
/* SEDEF's generate_anchors (src/chain.cc:24-101) for n region pairs at once, on the GPU: the maximal exact-match runs
 * (case-insensitive; N ends a run) of length >= kmer_size on every diagonal, each started at its first k-mer that occurs fewer
 * than 1000 times in the reference region, with the same-chromosome diagonal exclusion (same_chr[i] != 0: diagonals within
 * kmer_size of orig_ref_start[i] + r == orig_query_start[i] + q are dropped).  q/r buffers hold ORIGINAL-CASE bytes.
 * Output: one malloc()'d array (caller free()s) with the anchors of region i at [anchor_off[i], anchor_off[i+1]), in the order
 * the reference emits them (by q, then r); has_u = 1 when the match holds an upper-case base (Anchor, src/align.h:25-28). */
typedef struct { int32_t q, r, l, has_u; } sedef_anchor_t;
int sedef_anchors_batch(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                        const int *rlen, const int64_t *roff, const uint8_t *rbuf, int kmer_size,
                        const uint8_t *same_chr, const int64_t *orig_query_start, const int64_t *orig_ref_start,
                        sedef_anchor_t **anchors_out, int64_t *anchor_off /* [n + 1] */);

/* ---- resident batches (measurement + pipelined callers) -------------------------------- */
/*
 * A resident batch keeps the encoded inputs in HBM so that repeated runs time the device
 * path alone.  upload = pack + H2D; run = DP + traceback + stats kernels on the device(s)
 * (returns device time in ms measured with CUDA events on the launching streams, max over
 * devices); fetch = D2H + gather into ez/stats.
 */
typedef struct ksw_b200_batch ksw_b200_batch_t;
ksw_b200_batch_t *ksw_b200_batch_upload(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                                        const int *tlen, const int64_t *toff, const uint8_t *tbuf,
                                        int8_t m, const int8_t *mat, int8_t q, int8_t e,
                                        int w, int zdrop, int flag,
                                        const uint8_t *q_raw_buf, const uint8_t *t_raw_buf, int *err);
int  ksw_b200_batch_run(ksw_b200_batch_t *b, float *device_ms);
int  ksw_b200_batch_fetch(ksw_b200_batch_t *b, ksw_extz_t *ez, sd_stats_t *stats);
int  ksw_b200_batch_fetch_arena(ksw_b200_batch_t *b, int want_stats, ksw_b200_result_t **out);   /* no per-pair malloc */
/* number of kernel launches issued by the last run, and per-kernel-class device ms */
int  ksw_b200_batch_launches(const ksw_b200_batch_t *b);
int  ksw_b200_batch_kernel_ms(const ksw_b200_batch_t *b, float *dp_ms, float *tb_ms, float *aux_ms);
int64_t ksw_b200_batch_cells(const ksw_b200_batch_t *b);   /* host-side in-band cell count (all diagonals) */
/* bytes copied host->device by upload(+run) and device->host by the last fetch */
int  ksw_b200_batch_io_bytes(const ksw_b200_batch_t *b, int64_t *h2d, int64_t *d2h);
/* host-side phase times of this batch in ms: {plan, pack, h2d wait, d2h, gather} */
int  ksw_b200_batch_host_ms(const ksw_b200_batch_t *b, double *out5);
/* the fused SD-statistics pass is on by default for CIGAR runs; 0 switches it off for this batch */
void ksw_b200_batch_set_stats(ksw_b200_batch_t *b, int on);
void ksw_b200_batch_free(ksw_b200_batch_t *b);

/* In-band DP cells of one pair if every anti-diagonal is processed
 * (sum over r of en0-st0+1, extern/ksw2_extz2_sse.cc:105-109). */
int64_t ksw_b200_count_cells(int qlen, int tlen, int w);

#ifdef __cplusplus
}
#endif
#endif /* KSW2_B200_H_ */
