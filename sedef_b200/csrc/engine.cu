// engine.cu -- host side of libsedef_b200.so: the C ABI of include/ksw2_b200.h.
//
// Responsibilities (all host C++, no torch):
//   * validate + plan: per pair compute the rounded live-slot width (16*n_col_ of
//     extern/ksw2_extz2_sse.cc:74-75, or T if smaller), pick the narrowest kernel class that holds
//     it, count in-band cells, sort each class by descending work;
//   * shard over the bound devices with a length-balanced greedy (LPT) partition on cells;
//   * pack sequences into one pinned arena per device ([16 zero bytes | query | pad][target | pad]),
//     one H2D copy; optional second arena with the original-case bytes for the SD statistics;
//   * run: per class, in waves bounded by traceback memory, launch the DP kernel (persistent
//     grid, dynamic work queue) and the traceback+stats kernel on the device stream;
//   * fetch: D2H of compact records + compact CIGAR arena, gather by original index into
//     ksw_extz_t / sd_stats_t exactly as ksw_extz2_sse would have filled them.
// There is no CPU fallback: without a device every entry point fails.
#include <cuda_runtime.h>
#include <omp.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>
#include <condition_variable>

#include "../../include/ksw2_b200.h"
#include "kernels.h"

using namespace extz;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg) { g_last_error = msg; return code; }
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
	return fail(KSW_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } } while (0)

extern "C" const char *ksw_b200_strerror(int code)
{
	switch (code) {
	case KSW_B200_OK: return "ok";
	case KSW_B200_ERR_NO_DEVICE: return "no usable CUDA device (this engine has no CPU fallback)";
	case KSW_B200_ERR_CUDA: return "CUDA runtime error";
	case KSW_B200_ERR_DOMAIN: return "scoring parameters outside the supported domain";
	case KSW_B200_ERR_UNSUPPORTED: return "unsupported flag or alphabet (m > 8)";
	case KSW_B200_ERR_TOO_WIDE: return "pair needs more live slots per anti-diagonal than the widest kernel";
	case KSW_B200_ERR_NOMEM: return "out of memory";
	case KSW_B200_ERR_ARG: return "bad argument";
	case KSW_B200_ERR_INEXACT: return "inexact";
	default: return "unknown error";
	}
}
extern "C" const char *ksw_b200_last_error(void) { return g_last_error.c_str(); }

// ------------------------------------------------------------------------------------------------
// kernel classes
// ------------------------------------------------------------------------------------------------
// A class is a live-slot capacity: 32, 64, ..., 16384 (a pair needs min(16*ceil(tlen/16), 16*n_col_) slots).  Every class has
// two implementations:
//   * packed (default, extz_dp16.cuh; two slots per register, NS / 32 lanes per pair):
//       NS <= 1024           extz_dp16_kernel<NS/32>          128-thread CTAs, 32 / (NS/32) pairs per warp, each group at its own anti-diagonal;
//                                                             the 512-slot class holds 528 (a spare block spread over its 16 lanes)
//       NS = 2048/4096/8192  extz_dp16_wide_kernel<NS/32>     one CTA of 64 / 128 / 256 lanes per pair
//       NS = 16384 ... 65536 extz_dp16_cluster_kernel<C>      one cluster of C = 2 / 4 / 8 CTAs x 256 lanes per pair (DSMEM); 65536
//                                                             live slots hold the largest call the reference makes (60 000 per chunk)
//   * one slot per register (KSW_B200_PACKED=0, extz_dp.cuh; the table below: G lanes x S slots), kept for A/B runs and
//     under test: narrow (G <= 32), CTA-wide (one CTA of G lanes), cluster (`cluster` CTAs x 256 lanes).
struct KClass { int G, S; bool wide; int cluster; };
static const KClass kClasses[] = { {2, 16, false, 0}, {4, 16, false, 0}, {8, 16, false, 0}, {16, 16, false, 0}, {32, 16, false, 0},
                                   {32, 32, false, 0}, {64, 16, true, 0}, {128, 16, true, 0}, {256, 16, true, 0},
                                   {512, 16, true, 2}, {1024, 16, true, 4},
                                   {2048, 16, true, 0}, {4096, 16, true, 0} };       // packed only: clusters of 4 / 8 CTAs (32768 / 65536 slots)
static const int kNumClasses = sizeof(kClasses) / sizeof(kClasses[0]);
static inline bool class_packed(int c);
// all classes are ordered by capacity; the last two exist in packed form only
static inline int num_sized_classes() { return class_packed(0) ? kNumClasses : kNumClasses - 2; }
#define kNumSizedClasses num_sized_classes()
static inline int class_ns(int c) { return kClasses[c].G * kClasses[c].S; }
static inline bool packed_enabled()
{
	static const bool on = [] { const char *e = getenv("KSW_B200_PACKED"); return !(e && e[0] == '0'); }();
	return on;
}
static inline bool class_packed(int c) { return packed_enabled(); }                              // every class has a packed kernel
static inline bool class_packed_cluster(int c) { return class_packed(c) && class_ns(c) > 8192; } // 2 CTAs x 256 lanes x 32 slots
static inline int class_cluster(int c) { return class_packed(c) ? (class_ns(c) > 8192 ? class_ns(c) / 8192 : 0) : kClasses[c].cluster; }
static inline bool class_packed_wide(int c) { return class_packed(c) && class_ns(c) > 1024 && class_ns(c) <= 8192; }    // one CTA of NS/32 lanes per pair
// one-slot lanes with S == 32 switch whole-lane (two 16-blocks at once), which costs 16 slots of window (extz_dp.cuh)
// the packed 16-lane class (512 slots in registers) holds a 33rd block spread over its lanes (extz_dp16.cuh Spare16) -- not in the
// KSW_EZ_APPROX_MAX variant
static inline bool class_spare(int c, bool approx) { return class_packed(c) && class_ns(c) == 512 && !approx; }
static inline int class_capacity(int c, bool approx)
{
	if (class_spare(c, approx)) return class_ns(c) + 16;
	return (kClasses[c].S > 16 && !class_packed(c)) ? class_ns(c) - 16 : class_ns(c);
}
static inline size_t class_row_bytes(int c, bool approx) { return (size_t)class_ns(c) / 2 + (class_spare(c, approx) ? 16 : 0); }
static inline int class_pairs_per_block(int c)
{
	if (class_packed(c)) return class_ns(c) > 1024 ? 1 : 128 * 32 / class_ns(c);
	return kClasses[c].wide ? 1 : 128 / kClasses[c].G;
}

// grid: CTAs for narrow / wide classes, clusters for cluster classes
static cudaError_t launch_dp(int c, const DpLaunch &L, bool cigar, bool right, bool approx, int grid, cudaStream_t st)
{
	if (class_packed_cluster(c)) return k_dp16_cluster_dispatch(class_cluster(c), L, cigar, right, approx, grid, st, nullptr);
	if (class_cluster(c)) return k_dp_cluster_dispatch(class_cluster(c), L, cigar, right, grid, st, nullptr);
	if (class_packed_wide(c)) return k_dp16_wide_launch(class_ns(c) / 32, L, cigar, right, approx, grid, st);
	if (class_packed(c)) return k_dp16_launch(class_ns(c) / 32, L, cigar, right, approx, grid, st);
	return k_dp_launch(c, L, cigar, right, grid, st);
}
// resident CTAs per SM (narrow / wide) or co-resident clusters on the whole device (cluster classes), on the CURRENT device
static int dp_occupancy_query(int c, bool cigar, bool right, bool approx)
{
	if (class_cluster(c)) {
		int n = 0; DpLaunch dummy = {};
		cudaError_t e = class_packed_cluster(c) ? k_dp16_cluster_dispatch(class_cluster(c), dummy, cigar, right, approx, 1, nullptr, &n)
		                                        : k_dp_cluster_dispatch(class_cluster(c), dummy, cigar, right, 1, nullptr, &n);
		if (e != cudaSuccess) { cudaGetLastError(); return 0; }
		return n;
	}
	if (class_packed_wide(c)) return k_dp16_wide_occupancy(class_ns(c) / 32, cigar, right, approx);
	if (class_packed(c)) return k_dp16_occupancy(class_ns(c) / 32, cigar, right, approx);
	return k_dp_occupancy(c, cigar, right);
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevCtx {
	int dev = -1;
	int sms = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t tb_stream = nullptr;   // highest priority: a finished chunk's traceback must not queue behind the persistent
	                                    // DP CTAs of the chunks launched after it (one-shot pipeline)
	cudaStream_t wave_streams[8] = {};  // kernel classes of one batch run side by side (launch_sub)
	size_t tb_budget = 0;          // bytes of traceback memory one wave may use
	int occ[kNumClasses][6];       // cached occupancy per (class, {score-only, cigar-left, cigar-right} x {exact, approx max}); -1 = not asked yet
	DevCtx() { for (auto &row : occ) for (int &v : row) v = -1; }
};
// occupancy of class c on device dc (which must be current): asked once per device
static int dp_occupancy(DevCtx &dc, int c, bool cigar, bool right, bool approx)
{
	int &slot = dc.occ[c][(cigar ? (right ? 2 : 1) : 0) + (approx ? 3 : 0)];
	if (slot < 0) slot = dp_occupancy_query(c, cigar, right, approx);
	return slot;
}
static std::mutex g_mu;
static std::vector<DevCtx> g_devs;
// host threads used for packing / gathering (0: OpenMP default).  Launchers such as torchrun export
// OMP_NUM_THREADS=1, which would serialise the host side of every rank; callers can override it here.
static int g_host_threads = 0;
static inline int host_threads() { return g_host_threads > 0 ? g_host_threads : omp_get_max_threads(); }
// threads worth waking for a loop over n items of ~`grain` items per thread-millisecond: the host side of a batch is a few
// hundred microseconds of work, and an OpenMP team that is larger than the work only adds wake-up latency and spinning
// (with one process per GPU the ranks share the cores)
static inline int threads_for(int64_t n, int64_t grain) { return (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), n / grain)); }
extern "C" void ksw_b200_set_host_threads(int n) { g_host_threads = n > 0 ? n : 0; }

extern "C" int ksw_b200_init(int first_dev, int ndev)
{
	std::lock_guard<std::mutex> lk(g_mu);
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0)
		return fail(KSW_B200_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
	if (first_dev < 0) first_dev = 0;
	if (ndev <= 0 || first_dev + ndev > count) ndev = count - first_dev;
	if (ndev <= 0) return fail(KSW_B200_ERR_NO_DEVICE, "no device in the requested range");
	if ((int)g_devs.size() == ndev && g_devs[0].dev == first_dev) return ndev;
	for (auto &d : g_devs) { cudaSetDevice(d.dev); if (d.stream) cudaStreamDestroy(d.stream); if (d.tb_stream) cudaStreamDestroy(d.tb_stream); for (auto &ws : d.wave_streams) if (ws) cudaStreamDestroy(ws); }
	g_devs.clear();
	for (int i = 0; i < ndev; ++i) {
		DevCtx d; d.dev = first_dev + i;
		CUDA_TRY(cudaSetDevice(d.dev));
		cudaDeviceProp p; CUDA_TRY(cudaGetDeviceProperties(&p, d.dev));
		d.sms = p.multiProcessorCount;
		CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
		int prio_least = 0, prio_greatest = 0;
		CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
		CUDA_TRY(cudaStreamCreateWithPriority(&d.tb_stream, cudaStreamNonBlocking, prio_greatest));
		for (auto &ws : d.wave_streams) CUDA_TRY(cudaStreamCreateWithFlags(&ws, cudaStreamNonBlocking));
		size_t fr = 0, tot = 0; CUDA_TRY(cudaMemGetInfo(&fr, &tot));
		const char *env = getenv("KSW_B200_TB_BUDGET_MB");
		// A wave must hold enough pairs to fill the machine with SIMILAR work (3 CTAs x 4 warps per SM, up to 32 pairs per warp) --
		// with 10-50 kbp pairs (16 MB of rows each) a 48 GB wave was 2000 pairs for 3552 resident groups and ran at 1.7 warps per
		// scheduler; three quarters of the free memory (137 GB on a 180 GB B200) brought BASELINE configs[2] from 333 to 448 GCUPS
		// (profiles/r02_tuning.md).  Memory is only taken as far as a wave needs it.
		d.tb_budget = env ? (size_t)atoll(env) << 20 : std::min<size_t>(fr / 4 * 3, (size_t)140 << 30);
		g_devs.push_back(d);
	}
	return ndev;
}
static void pool_clear();
extern "C" void ksw_b200_destroy(void)
{
	pool_clear();
	std::lock_guard<std::mutex> lk(g_mu);
	for (auto &d : g_devs) { cudaSetDevice(d.dev); if (d.stream) cudaStreamDestroy(d.stream); if (d.tb_stream) cudaStreamDestroy(d.tb_stream); for (auto &ws : d.wave_streams) if (ws) cudaStreamDestroy(ws); }
	g_devs.clear();
}
extern "C" int ksw_b200_num_devices(void) { return (int)g_devs.size(); }

// ---- page-locked host memory for callers: sequences that live in pinned memory are copied to the device in place ----
extern "C" void *ksw_b200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}
extern "C" void ksw_b200_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int ksw_b200_host_register(void *p, size_t bytes)
{
	if (!p || !bytes) return fail(KSW_B200_ERR_ARG, "null range");
	CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
	return KSW_B200_OK;
}
extern "C" int ksw_b200_host_unregister(void *p)
{
	if (!p) return KSW_B200_OK;
	CUDA_TRY(cudaHostUnregister(p));
	return KSW_B200_OK;
}
// true when [p, p + bytes) can be the source / destination of an asynchronous DMA copy as it is
static bool is_pinned(const void *p)
{
	if (!p) return false;
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
}
extern "C" int ksw_b200_max_slots(void) { return class_capacity(kNumSizedClasses - 1, false); }

static int ensure_init()
{
	if (!g_devs.empty()) return (int)g_devs.size();
	return ksw_b200_init(0, 0);
}

// ------------------------------------------------------------------------------------------------
// planning helpers
// ------------------------------------------------------------------------------------------------
extern "C" int64_t ksw_b200_count_cells(int qlen, int tlen, int w)
{
	if (qlen <= 0 || tlen <= 0) return 0;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	// closed form would need care with the >>1 rounding; the loop is O(q+t) and only runs at plan time
	int64_t c = 0;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
		if (en > ((r + w) >> 1)) en = (r + w) >> 1;
		if (st > en) break;
		c += en - st + 1;
	}
	return c;
}
// cheap estimate used for load balancing only
static inline int64_t est_cells(int qlen, int tlen, int w)
{
	int64_t mn = std::min(qlen, tlen);
	int64_t width = std::min<int64_t>(mn, (int64_t)w + 1);
	return width * ((int64_t)qlen + tlen - 1);
}

// Buffers are recycled through a small process-wide pool: cudaMalloc / cudaMallocHost of hundreds of MB cost
// tens of milliseconds, which would dominate a one-shot ksw_extz2_batch call (the align-stage driver issues
// one batch per wave, back to back).
struct PoolEntry { void *p; size_t cap; int dev; bool pinned; };
static std::mutex g_pool_mu;
static std::vector<PoolEntry> g_pool;
static void *pool_take(size_t bytes, int dev, bool pinned, size_t *cap_out)
{
	std::lock_guard<std::mutex> lk(g_pool_mu);
	int best = -1;
	for (int i = 0; i < (int)g_pool.size(); ++i)
		if (g_pool[i].pinned == pinned && (pinned || g_pool[i].dev == dev) && g_pool[i].cap >= bytes &&
		    (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = i;
	if (best < 0) return nullptr;
	void *p = g_pool[best].p; *cap_out = g_pool[best].cap;
	g_pool.erase(g_pool.begin() + best);
	return p;
}
static void pool_give(void *p, size_t cap, int dev, bool pinned)
{
	std::lock_guard<std::mutex> lk(g_pool_mu);
	g_pool.push_back({p, cap, dev, pinned});
}
static void pool_clear()
{
	std::lock_guard<std::mutex> lk(g_pool_mu);
	for (auto &e : g_pool) { if (e.pinned) cudaFreeHost(e.p); else { cudaSetDevice(e.dev); cudaFree(e.p); } }
	g_pool.clear();
}
struct DevBuf {
	void *p = nullptr; size_t cap = 0; int dev = 0;
	int ensure(size_t bytes) {
		if (bytes <= cap) return 0;
		release();
		cudaGetDevice(&dev);
		if ((p = pool_take(bytes, dev, false, &cap))) return 0;
		size_t want = bytes + bytes / 8 + 256;
		if (cudaMalloc(&p, want) != cudaSuccess) {
			cudaGetLastError(); pool_clear();
			if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
			want = bytes;
		}
		cap = want; return 0;
	}
	void release() { if (p) pool_give(p, cap, dev, false); p = nullptr; cap = 0; }
};
struct PinBuf {
	void *p = nullptr; size_t cap = 0;
	int ensure(size_t bytes) {
		if (bytes <= cap) return 0;
		release();
		if ((p = pool_take(bytes, 0, true, &cap))) return 0;
		size_t want = bytes + bytes / 8 + 256;
		if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
		cap = want; return 0;
	}
	void release() { if (p) pool_give(p, cap, 0, true); p = nullptr; cap = 0; }
};

struct Wave { int cls; int first, count; size_t tb_bytes; size_t tb_base = 0; };   // tb_base: offset of the wave's rows in the arena (concurrent waves)

// one device's share of a batch
struct SubBatch {
	DevCtx *dc = nullptr;
	std::vector<PairDesc> pairs;        // grouped by class, each class sorted by descending work
	std::vector<Wave> waves;
	int class_first[kNumClasses + 1] = {0};
	size_t arena_bytes = 0;             // bytes of one sequence plane on the device (codes; same size for the original-case plane)
	PinBuf h_stage[2], h_recs, h_stats, h_trims, h_pairs;
	DevBuf d_arena, d_raw, d_pairs, d_results, d_tb, d_cigar, d_stats, d_trims, d_misc, d_table, d_ez;
	size_t cigar_cap = 0;               // entries
	unsigned long long cigar_used = 0;
	cudaStream_t stream = nullptr;      // every batch owns a stream so that the H2D of one batch overlaps the kernels of another
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	float dp_ms = 0, tb_ms = 0, total_ms = 0;
	int launches = 0, aux_launches = 0;
	int check_limit = 0;                // > 0: caller-supplied codes must be below this (checked on the device, first run only)
	size_t h2d_seq_bytes = 0, h2d_desc_bytes = 0, d2h_bytes = 0;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> dp_ev, tb_ev;   // per-wave kernel timing events of the last launch
	bool launched = false, touched = false;
	bool concurrent = false;            // the waves of this launch run side by side, each on its own stream and arena slice
	std::vector<int> ev_wave;           // wave index of every dp_ev / tb_ev entry
	void drop_events()
	{
		for (auto &p : dp_ev) cudaEventDestroy(p.first);
		for (auto &p : tb_ev) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
		dp_ev.clear(); tb_ev.clear();
	}
	void release() {
		// kernels and copies of this batch may still be in flight on an error path: the buffers go back to a process-wide pool,
		// so wait for the stream first (the traceback stream's work is ordered before it by events)
		if (stream && touched) cudaStreamSynchronize(stream);
		if (dc && touched && concurrent) for (auto &ws : dc->wave_streams) if (ws) cudaStreamSynchronize(ws);   // (joined into `stream` unless a launch failed half-way)
		drop_events();
		for (auto &b : h_stage) b.release();
		h_recs.release(); h_stats.release(); h_trims.release(); h_pairs.release();
		d_arena.release(); d_raw.release(); d_pairs.release(); d_results.release(); d_tb.release();
		d_cigar.release(); d_stats.release(); d_trims.release(); d_misc.release(); d_table.release(); d_ez.release();
		for (auto &e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
		if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
	}
};

struct ksw_b200_batch {
	int n = 0;
	int8_t m = 0; int8_t q = 0, e = 0; int w = 0, zdrop = 0, flag = 0;
	int8_t mat[64];
	bool early_out = false;             // -min_sc > 2(q+e): every pair returns the reset record (:81)
	bool want_stats = false, have_raw = false;
	bool encode_on_device = false;      // the caller passed original-case bytes only: codes = align_dna(bytes), computed on the device
	int tb_budget_div = 1;              // batches in flight at once (one-shot pipeline): each gets this share of the traceback budget
	Scoring sc;
	uint32_t table[kTableStride * kTableStride];
	std::vector<SubBatch> subs;
	std::vector<uint8_t> is_empty;      // pairs with qlen<=0 || tlen<=0 (reset record)
	int n_empty = 0;
	int64_t cells = -1;                 // exact in-band cell count, computed on first request
	std::vector<int> cq, ct;            // lengths kept for the lazy cell count
	double t_plan = 0, t_pack = 0, t_h2d = 0, t_d2h = 0, t_gather = 0;   // host-side phase times (ms)
};

// Results of a batch in ONE arena: ksw_extz_t records in the caller's order whose `cigar` pointers point
// into page-locked buffers owned by this object, and the statistics records.  Nothing here is malloc()'d per pair.
struct ksw_b200_result {
	int n = 0;
	bool has_stats = false;
	PinBuf ez, stats, trims;            // trims: {trim_front max_i | -1, trim_back kept columns | -1} per pair, with the statistics
	std::vector<PinBuf> cigars;         // one compact CIGAR arena per (chunk, device)
	int64_t h2d = 0, d2h = 0;
	int launches = 0;
	void release() { ez.release(); stats.release(); trims.release(); for (auto &c : cigars) c.release(); cigars.clear(); }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int build_scoring(ksw_b200_batch &B)
{
	const int m = B.m, q = B.q, e = B.e;
	if (m <= 0) return 0;   // every pair is reset (:57)
	if (m > kTableStride) return fail(KSW_B200_ERR_UNSUPPORTED, "alphabet size m > 8 is not supported");
	if ((B.flag & KSW_EZ_APPROX_MAX) && !packed_enabled())
		return fail(KSW_B200_ERR_UNSUPPORTED, "KSW_EZ_APPROX_MAX is implemented by the packed kernels only (unset KSW_B200_PACKED=0)");
	int max_sc = B.mat[0], min_sc = B.mat[1];
	for (int t = 1; t < m * m; ++t) { max_sc = std::max<int>(max_sc, B.mat[t]); min_sc = std::min<int>(min_sc, B.mat[t]); }
	B.early_out = (-min_sc > 2 * (q + e));                                          // :81
	const int qe2 = (q + e) * 2;
	auto top = [](int v) { return (uint32_t)(uint8_t)(int8_t)v << 24; };            // int8 truncation like _mm_set1_epi8
	B.sc.q_s = top(q); B.sc.maxsc_s = top(B.mat[0] + qe2); B.sc.s0_s = top(0 + qe2);
	B.sc.q = q; B.sc.e = e; B.sc.qe = q + e; B.sc.zdrop = B.zdrop; B.sc.flag = B.flag; B.sc.w_in = B.w;
	B.sc.q16 = B.sc.q_s >> 16;
	B.sc.qeps2 = B.sc.q16 * 0x00010001u + 0x00010001u;          // ~z + qeps2 == q - z per half
	B.sc.maxsc2 = (B.sc.maxsc_s >> 16) * 0x00010001u;
	B.sc.s0_2 = (B.sc.s0_s >> 16) * 0x00010001u;
	B.sc.zr2 = 0u;                                               // a zero the compiler cannot fold (see extz_dp16.cuh)
	for (int a = 0; a < kTableStride; ++a)
		for (int b = 0; b < kTableStride; ++b) {
			int s;
			if (B.flag & KSW_EZ_GENERIC_SC) s = (a < m && b < m) ? B.mat[a * m + b] : 0;     // :141
			else s = (a == m - 1 || b == m - 1) ? 0 : (a == b ? B.mat[0] : B.mat[1]);        // :129-136
			B.table[a * kTableStride + b] = top((int8_t)s + (int8_t)qe2);                    // :27 z = s + qe2 (wrapping)
		}
	return 0;
}

// live-slot requirement of one pair
static inline int slots_needed(int qlen, int tlen, int w)
{
	int T = (tlen + 15) / 16 * 16;
	int n_col = std::min(qlen, tlen);
	n_col = (std::min(n_col, w + 1) + 15) / 16 + 1;                                 // :74-75
	return std::min(T, n_col * 16);
}

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------
// A persistent helper thread for the one-shot pipeline's producer stage: a fresh std::thread per call would also build
// (and tear down) a fresh OpenMP team for the packing loops, ~3 ms on a 32-thread host.
class PipelineWorker {
	std::thread th_;
	std::mutex mu_;
	std::condition_variable cv_;
	std::function<void()> job_;
	bool has_job_ = false, running_ = false, stop_ = false;
	void loop()
	{
		for (;;) {
			std::function<void()> job;
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_.wait(lk, [&] { return has_job_ || stop_; });
				if (stop_ && !has_job_) return;
				job.swap(job_); has_job_ = false; running_ = true;
			}
			job();
			{ std::lock_guard<std::mutex> lk(mu_); running_ = false; }
			cv_.notify_all();
		}
	}
public:
	std::mutex claim;                                   // one call at a time owns the worker (others spawn their own thread)
	~PipelineWorker() { { std::lock_guard<std::mutex> lk(mu_); stop_ = true; } cv_.notify_all(); if (th_.joinable()) th_.join(); }
	void start(std::function<void()> f)
	{
		std::lock_guard<std::mutex> lk(mu_);
		if (!th_.joinable()) th_ = std::thread([this] { loop(); });
		job_ = std::move(f); has_job_ = true;
		cv_.notify_all();
	}
	void join() { std::unique_lock<std::mutex> lk(mu_); cv_.wait(lk, [&] { return !has_job_ && !running_; }); }
};
static PipelineWorker g_worker;

// One sequence plane (codes or original-case bytes) of one device's pairs goes to the device as
//   [ query bytes | pad to 256 | target bytes | kArenaSlack ]
// in one of two ways:
//   dense   the pairs reference (almost) every byte of a contiguous range of the caller's flat buffers (the normal case:
//           a batch built back to back, or windows of one genome): the two ranges are copied AS THEY ARE -- straight from
//           the caller's memory when it is page-locked (no host pass over the sequences at all), through a pinned staging
//           buffer in 8 MB blocks otherwise -- and a pair's offset is its caller offset minus the range start;
//   sparse  otherwise every pair is packed back to back into the staging buffer first.
struct PlanePlan {
	bool dense = false;
	int64_t qlo = 0, tlo = 0;           // start of the referenced ranges in the caller's buffers (dense)
	size_t qspan = 0, tspan = 0;        // bytes per side on the device
	size_t tbase = 0;                   // device offset of the target side
};
static int copy_plane(SubBatch &sb, const PlanePlan &pl, int slot, void *d_dst, const uint8_t *qsrc, const uint8_t *tsrc,
                      const int64_t *qoff, const int64_t *toff, cudaStream_t st, double *t_pack)
{
	uint8_t *dst = (uint8_t *)d_dst;
	if (pl.dense) {
		const uint8_t *src[2] = {qsrc + pl.qlo, tsrc + pl.tlo};
		const size_t len[2] = {pl.qspan, pl.tspan}, dofs[2] = {0, pl.tbase};
		for (int side = 0; side < 2; ++side) {
			if (!len[side]) continue;
			if (is_pinned(src[side])) { CUDA_TRY(cudaMemcpyAsync(dst + dofs[side], src[side], len[side], cudaMemcpyHostToDevice, st)); continue; }
			const double t0 = now_ms();
			if (sb.h_stage[slot].ensure(sb.arena_bytes)) return fail(KSW_B200_ERR_NOMEM, "pinned staging allocation failed");
			uint8_t *stage = (uint8_t *)sb.h_stage[slot].p + dofs[side];
			const size_t blk = (size_t)8 << 20;
			for (size_t o = 0; o < len[side]; o += blk) {                       // the DMA of block k overlaps the memcpy of block k+1
				const size_t nb = std::min(blk, len[side] - o);
				const int nt = (int)std::max<size_t>(1, std::min<size_t>(host_threads(), nb >> 18));
#pragma omp parallel for num_threads(nt) schedule(static)
				for (int t = 0; t < nt; ++t) {
					const size_t a = nb * t / nt, b = nb * (t + 1) / nt;
					memcpy(stage + o + a, src[side] + o + a, b - a);
				}
				CUDA_TRY(cudaMemcpyAsync(dst + dofs[side] + o, stage + o, nb, cudaMemcpyHostToDevice, st));
			}
			*t_pack += now_ms() - t0;
		}
		return 0;
	}
	const double t0 = now_ms();
	if (sb.h_stage[slot].ensure(sb.arena_bytes)) return fail(KSW_B200_ERR_NOMEM, "pinned staging allocation failed");
	uint8_t *stage = (uint8_t *)sb.h_stage[slot].p;
	const int64_t np = (int64_t)sb.pairs.size();
#pragma omp parallel for num_threads(host_threads()) schedule(static)
	for (int64_t k = 0; k < np; ++k) {
		const PairDesc &pd = sb.pairs[k];
		memcpy(stage + pd.q_off, qsrc + qoff[pd.orig], pd.qlen);
		memcpy(stage + pd.t_off, tsrc + toff[pd.orig], pd.tlen);
	}
	*t_pack += now_ms() - t0;
	CUDA_TRY(cudaMemcpyAsync(dst, stage, pl.tbase + pl.tspan, cudaMemcpyHostToDevice, st));
	return 0;
}

static thread_local bool tl_async_upload = false;      // set by the one-shot pipeline around its chunk uploads
static thread_local int tl_pipeline_depth = 1;
extern "C" ksw_b200_batch_t *ksw_b200_batch_upload(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                                                   const int *tlen, const int64_t *toff, const uint8_t *tbuf,
                                                   int8_t m, const int8_t *mat, int8_t q, int8_t e,
                                                   int w, int zdrop, int flag,
                                                   const uint8_t *q_raw_buf, const uint8_t *t_raw_buf, int *err)
{
	auto bail = [&](int code) -> ksw_b200_batch_t * { if (err) *err = code; return nullptr; };
	if (err) *err = KSW_B200_OK;
	const bool have_codes = qbuf && tbuf, have_raw = q_raw_buf && t_raw_buf;
	if (n < 0 || (n > 0 && (!qlen || !tlen || !qoff || !toff || (!have_codes && !have_raw))) || (m > 0 && !mat))
		return bail(fail(KSW_B200_ERR_ARG, "null pointer or negative count"));
	if (!have_codes && m > 0 && m != 5)
		return bail(fail(KSW_B200_ERR_ARG, "sequences given as original-case bytes only are encoded with align_dna (A C G T other -> 0..4): m must be 5"));
	int nd = ensure_init();
	if (nd <= 0) return bail(nd);

	ksw_b200_batch *B = new ksw_b200_batch();
	B->n = n; B->m = m; B->q = q; B->e = e; B->w = w; B->zdrop = zdrop; B->flag = flag;
	memset(B->mat, 0, sizeof(B->mat));
	if (m > 0 && m <= kTableStride) memcpy(B->mat, mat, (size_t)m * m);
	B->have_raw = have_raw;
	B->encode_on_device = !have_codes;
	B->tb_budget_div = std::max(1, tl_pipeline_depth);
	B->want_stats = !(flag & KSW_EZ_SCORE_ONLY) && !getenv("KSW_B200_NO_STATS");   // the fused K4 pass is on by default
	int rc = build_scoring(*B);
	if (rc) { delete B; return bail(rc); }
	B->is_empty.assign(n, 0);
	B->subs.resize(nd);
	for (int d = 0; d < nd; ++d) B->subs[d].dc = &g_devs[d];
	auto destroy = [&](int code) -> ksw_b200_batch_t * {
		for (auto &s : B->subs) { if (s.dc) cudaSetDevice(s.dc->dev); s.release(); }
		delete B; return bail(code);
	};

	const double t_start = now_ms();
	// ---- classify + LPT partition ----
	struct Item { int cls; int64_t work; };
	std::vector<Item> items(n);
	const bool skip_s32 = getenv("KSW_B200_SKIP_S32") != nullptr;                       // A/B: 32 lanes x 32 slots vs 64-lane CTA
	const bool approx_batch = (flag & KSW_EZ_APPROX_MAX) != 0;
	int too_wide = -1;
#pragma omp parallel for num_threads(threads_for(n, 32768)) schedule(static)
	for (int i = 0; i < n; ++i) {
		items[i].cls = -1; items[i].work = 0;
		if (m <= 0 || qlen[i] <= 0 || tlen[i] <= 0 || B->early_out) continue;           // empty record (:57,81)
		int wi = w < 0 ? std::max(qlen[i], tlen[i]) : std::min(w, std::max(qlen[i], tlen[i]));
		int need = slots_needed(qlen[i], tlen[i], wi);
		int c = 0;
		while (c < kNumSizedClasses && class_capacity(c, approx_batch) < need) ++c;
		if (c == 5 && skip_s32) c = 6;
		if (c == kNumSizedClasses) {
#pragma omp critical
			{ if (too_wide < 0 || i < too_wide) too_wide = i; }
			continue;
		}
		items[i].cls = c; items[i].work = est_cells(qlen[i], tlen[i], wi);
	}
	if (too_wide >= 0) {
		const int i = too_wide;
		int wi = w < 0 ? std::max(qlen[i], tlen[i]) : std::min(w, std::max(qlen[i], tlen[i]));
		delete B;
		return bail(fail(KSW_B200_ERR_TOO_WIDE, "pair " + std::to_string(i) + " needs " + std::to_string(slots_needed(qlen[i], tlen[i], wi)) +
		                 " live slots; widest kernel holds " + std::to_string(class_capacity(kNumSizedClasses - 1, false))));
	}
	std::vector<int> order; order.reserve(n);
	for (int i = 0; i < n; ++i) { if (items[i].cls < 0) { B->is_empty[i] = 1; ++B->n_empty; } else order.push_back(i); }
	std::vector<std::vector<int>> per_dev(nd);
	if (nd == 1) {
		// one device: the order the kernels want directly -- by class, descending work inside a class (the device-side
		// dynamic queue then hands out the long pairs first).  The order only steers load balance (results are gathered by
		// original index), so the work is quantised to 20 bits and the keys go through a stable LSD radix sort: O(n), no
		// comparison sort on the producer thread's critical path.
		int64_t max_work = 1;
		for (int i : order) max_work = std::max(max_work, items[i].work);
		int shift = 0;
		while ((max_work >> shift) >= (1 << 20)) ++shift;
		const size_t no = order.size();
		std::vector<uint32_t> key(no), key2(no);
		std::vector<int> order2(no);
		for (size_t k = 0; k < no; ++k) {
			const Item &it = items[order[k]];
			key[k] = ((uint32_t)it.cls << 20) | (uint32_t)((1 << 20) - 1 - (it.work >> shift));
		}
		for (int pass = 0; pass < 2; ++pass) {                                         // 2 x 12 bits
			const int sh = 12 * pass;
			uint32_t cnt[4097] = {0};
			for (size_t k = 0; k < no; ++k) ++cnt[((key[k] >> sh) & 4095u) + 1];
			for (int b = 0; b < 4096; ++b) cnt[b + 1] += cnt[b];
			for (size_t k = 0; k < no; ++k) { const uint32_t d = cnt[(key[k] >> sh) & 4095u]++; key2[d] = key[k]; order2[d] = order[k]; }
			key.swap(key2); order.swap(order2);
		}
		per_dev[0].swap(order);
	} else {
		std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return items[a].work > items[b].work; });
		std::vector<int64_t> load(nd, 0);
		for (int oi : order) {                                                          // greedy LPT
			int best = 0;
			for (int d = 1; d < nd; ++d) if (load[d] < load[best]) best = d;
			load[best] += items[oi].work + 1;
			per_dev[best].push_back(oi);
		}
	}

	B->t_plan = now_ms() - t_start;
	// ---- per device: group by class (descending work inside), lay the sequence planes out, copy ----
	for (int d = 0; d < nd; ++d) {
		SubBatch &sb = B->subs[d];
		auto &lst = per_dev[d];
		if (nd > 1) std::stable_sort(lst.begin(), lst.end(), [&](int a, int b) { return items[a].cls < items[b].cls; });
		sb.pairs.resize(lst.size());
		for (int c = 0; c <= kNumClasses; ++c) {
			int k = 0;
			while (k < (int)lst.size() && items[lst[k]].cls < c) ++k;
			sb.class_first[c] = k;                                                  // first pair of class >= c
		}
		if (lst.empty()) continue;
		// the ranges of the caller's buffers this device's pairs reference
		const int64_t np = (int64_t)lst.size();
		int64_t qlo = INT64_MAX, qhi = 0, tlo = INT64_MAX, thi = 0, qsum = 0, tsum = 0;
#pragma omp parallel for num_threads(threads_for(np, 65536)) schedule(static) reduction(min : qlo, tlo) reduction(max : qhi, thi) reduction(+ : qsum, tsum)
		for (int64_t k = 0; k < np; ++k) {
			const int i = lst[k];
			qlo = std::min(qlo, qoff[i]); qhi = std::max(qhi, qoff[i] + qlen[i]); qsum += qlen[i];
			tlo = std::min(tlo, toff[i]); thi = std::max(thi, toff[i] + tlen[i]); tsum += tlen[i];
		}
		PlanePlan pl;
		pl.dense = qlo >= 0 && tlo >= 0 && !getenv("KSW_B200_FORCE_SPARSE") &&
		           (uint64_t)(qhi - qlo) <= (uint64_t)qsum + (uint64_t)qsum / 2 + 65536 &&
		           (uint64_t)(thi - tlo) <= (uint64_t)tsum + (uint64_t)tsum / 2 + 65536;
		if (pl.dense) { pl.qlo = qlo; pl.tlo = tlo; pl.qspan = (size_t)(qhi - qlo); pl.tspan = (size_t)(thi - tlo); }
		else { pl.qspan = (size_t)qsum; pl.tspan = (size_t)tsum; }
		pl.tbase = align_up(pl.qspan, 256);
		sb.arena_bytes = align_up(pl.tbase + pl.tspan + kArenaSlack, 256);
		if (pl.dense) {
#pragma omp parallel for num_threads(threads_for(np, 8192)) schedule(static)
			for (int64_t k = 0; k < np; ++k) {              // a gather over the caller's arrays in kernel order: worth several threads
				PairDesc &pd = sb.pairs[k];
				const int i = lst[k];
				pd.qlen = qlen[i]; pd.tlen = tlen[i];
				pd.w = w < 0 ? std::max(qlen[i], tlen[i]) : std::min(w, std::max(qlen[i], tlen[i]));
				pd.orig = i; pd.tb_off = 0;
				pd.q_off = qoff[i] - qlo; pd.t_off = (int64_t)pl.tbase + (toff[i] - tlo);
			}
		} else {
			size_t qpos = 0, tpos = pl.tbase;
			for (int64_t k = 0; k < np; ++k) {
				PairDesc &pd = sb.pairs[k];
				const int i = lst[k];
				pd.qlen = qlen[i]; pd.tlen = tlen[i];
				pd.w = w < 0 ? std::max(qlen[i], tlen[i]) : std::min(w, std::max(qlen[i], tlen[i]));
				pd.orig = i; pd.tb_off = 0;
				pd.q_off = (int64_t)qpos; pd.t_off = (int64_t)tpos; qpos += qlen[i]; tpos += tlen[i];
			}
		}
		if (cudaSetDevice(sb.dc->dev) != cudaSuccess) return destroy(fail(KSW_B200_ERR_CUDA, "cudaSetDevice"));
		if (sb.d_arena.ensure(sb.arena_bytes) || (B->have_raw && sb.d_raw.ensure(sb.arena_bytes)))
			return destroy(fail(KSW_B200_ERR_NOMEM, "sequence arena allocation failed"));
		if (!sb.stream && cudaStreamCreateWithFlags(&sb.stream, cudaStreamNonBlocking) != cudaSuccess) sb.stream = nullptr;
		cudaStream_t st = sb.stream ? sb.stream : sb.dc->stream;
		sb.touched = true;
		const size_t pbytes = sb.pairs.size() * sizeof(PairDesc);
		bool ok = !sb.d_pairs.ensure(pbytes) && !sb.d_results.ensure(sb.pairs.size() * sizeof(PairResult)) &&
		          !sb.d_table.ensure(sizeof(B->table)) && !sb.d_misc.ensure(4096);
		if (!ok) return destroy(fail(KSW_B200_ERR_NOMEM, "descriptor allocation failed"));
		// misc: [0..2047] work counters (one per wave), [2048] cigar cursor (u64), [2112] CIGAR arena overflow, [2116] bad symbol
		if (cudaMemsetAsync(sb.d_misc.p, 0, 4096, st) != cudaSuccess) return destroy(fail(KSW_B200_ERR_CUDA, "cudaMemsetAsync"));
		// the slack behind the last sequence is read (never used): keep it defined
		ok = cudaMemsetAsync((char *)sb.d_arena.p + sb.arena_bytes - kArenaSlack - 256, 0, kArenaSlack + 256, st) == cudaSuccess;
		if (ok && B->have_raw) ok = cudaMemsetAsync((char *)sb.d_raw.p + sb.arena_bytes - kArenaSlack - 256, 0, kArenaSlack + 256, st) == cudaSuccess;
		if (!ok) return destroy(fail(KSW_B200_ERR_CUDA, "cudaMemsetAsync"));
		sb.h2d_seq_bytes = 0;
		if (have_codes) {
			if ((rc = copy_plane(sb, pl, 0, sb.d_arena.p, qbuf, tbuf, qoff, toff, st, &B->t_pack))) return destroy(rc);
			sb.h2d_seq_bytes += pl.qspan + pl.tspan;
			const int limit = (flag & KSW_EZ_GENERIC_SC) ? std::min<int>(m, kTableStride) : kTableStride;
			sb.check_limit = limit;                             // checked per pair once the descriptors are on the device (launch_sub)
		}
		if (B->have_raw) {
			if ((rc = copy_plane(sb, pl, 1, sb.d_raw.p, q_raw_buf, t_raw_buf, qoff, toff, st, &B->t_pack))) return destroy(rc);
			sb.h2d_seq_bytes += pl.qspan + pl.tspan;
			if (!have_codes) {
				if (k_encode_launch((const uint8_t *)sb.d_raw.p, (uint8_t *)sb.d_arena.p, sb.arena_bytes, st) != cudaSuccess)
					return destroy(fail(KSW_B200_ERR_CUDA, "encode launch failed"));
				++sb.aux_launches;
			}
		}
		ok = cudaMemcpyAsync(sb.d_table.p, B->table, sizeof(B->table), cudaMemcpyHostToDevice, st) == cudaSuccess;
		sb.h2d_seq_bytes += sizeof(B->table);
		for (auto &e2 : sb.ev) if (ok) ok = cudaEventCreate(&e2) == cudaSuccess;
		if (!ok) return destroy(fail(KSW_B200_ERR_CUDA, std::string("upload failed: ") + cudaGetErrorString(cudaGetLastError())));
	}
	B->cq.assign(qlen, qlen + n); B->ct.assign(tlen, tlen + n);
	if (!tl_async_upload) {           // resident API: the inputs are in HBM when this returns (the one-shot pipeline does not wait)
		const double t0 = now_ms();
		for (auto &sb : B->subs) if (!sb.pairs.empty()) { cudaSetDevice(sb.dc->dev); cudaStreamSynchronize(sb.stream ? sb.stream : sb.dc->stream); }
		B->t_h2d = now_ms() - t0;
	}
	return B;
}

extern "C" int64_t ksw_b200_batch_cells(const ksw_b200_batch_t *cb)
{
	ksw_b200_batch_t *b = const_cast<ksw_b200_batch_t *>(cb);
	if (!b) return 0;
	if (b->cells < 0) {                                                             // exact count, O(sum of lengths), parallel
		int64_t cells = 0;
#pragma omp parallel for num_threads(host_threads()) schedule(dynamic, 256) reduction(+ : cells)
		for (int i = 0; i < b->n; ++i)
			if (!b->is_empty[i]) cells += ksw_b200_count_cells(b->cq[i], b->ct[i], b->w);
		b->cells = cells;
	}
	return b->cells;
}
extern "C" int ksw_b200_batch_host_ms(const ksw_b200_batch_t *b, double *out5)
{
	if (!b || !out5) return KSW_B200_ERR_ARG;
	out5[0] = b->t_plan; out5[1] = b->t_pack; out5[2] = b->t_h2d; out5[3] = b->t_d2h; out5[4] = b->t_gather;
	return 0;
}
extern "C" int ksw_b200_batch_io_bytes(const ksw_b200_batch_t *b, int64_t *h2d, int64_t *d2h)
{
	if (!b) return KSW_B200_ERR_ARG;
	int64_t a = 0, c = 0;
	for (auto &sb : b->subs) { a += (int64_t)(sb.h2d_seq_bytes + sb.h2d_desc_bytes); c += (int64_t)sb.d2h_bytes; }
	if (h2d) *h2d = a;
	if (d2h) *d2h = c;
	return 0;
}
extern "C" void ksw_b200_batch_set_stats(ksw_b200_batch_t *b, int on) { if (b) b->want_stats = on && !(b->flag & KSW_EZ_SCORE_ONLY); }

// ------------------------------------------------------------------------------------------------
// run
// ------------------------------------------------------------------------------------------------
static int plan_waves(ksw_b200_batch &B, SubBatch &sb, bool cigar)
{
	sb.waves.clear();
	size_t max_tb = 0, max_cig = 0;
	const size_t budget = std::max<size_t>(sb.dc->tb_budget / (size_t)B.tb_budget_div, (size_t)1 << 20);
	for (int c = 0; c < kNumClasses; ++c) {
		int first = sb.class_first[c], last = sb.class_first[c + 1];
		if (first >= last) continue;
		const size_t rowB = class_row_bytes(c, (B.flag & KSW_EZ_APPROX_MAX) != 0);
		int k = first;
		while (k < last) {
			Wave wv{c, k, 0, 0};
			size_t cig = 0;
			while (k < last) {
				PairDesc &pd = sb.pairs[k];
				size_t bytes = cigar ? align_up(((size_t)pd.qlen + pd.tlen) * rowB, 256) : 0;
				if (wv.count > 0 && wv.tb_bytes + bytes > budget) break;
				pd.tb_off = (int64_t)wv.tb_bytes;
				wv.tb_bytes += bytes;
				cig += (size_t)pd.qlen + pd.tlen + 2;
				++wv.count; ++k;
			}
			max_tb = std::max(max_tb, wv.tb_bytes);
			max_cig += cig;
			sb.waves.push_back(wv);
		}
	}
	// Waves of different classes are independent.  When all of them fit the budget together they get disjoint slices of the arena
	// and run SIDE BY SIDE (launch_sub): a mixed batch -- SEDEF's waves hold everything from 30-base gap fills to a few multi-kbp
	// unbanded fills that keep one CTA or cluster busy for tens of milliseconds -- then takes as long as its slowest class instead
	// of the sum of all classes.
	size_t total_tb = 0;
	for (Wave &wv : sb.waves) { wv.tb_base = 0; total_tb += align_up(wv.tb_bytes, 256); }
	static const bool no_conc = getenv("KSW_B200_SERIAL_WAVES") != nullptr;                // A/B aid
	sb.concurrent = !no_conc && sb.waves.size() > 1 && sb.waves.size() <= 32 && total_tb <= budget;
	if (sb.concurrent) {
		size_t base = 0;
		for (Wave &wv : sb.waves) { wv.tb_base = base; base += align_up(wv.tb_bytes, 256); }
		max_tb = total_tb;
	}
	if (cigar) {
		if (sb.d_tb.ensure(max_tb + 256)) return fail(KSW_B200_ERR_NOMEM, "traceback arena allocation failed");
		// compact CIGAR arena: offsets are int32 in PairResult
		if (max_cig > 0x7fffffffull) max_cig = 0x7fffffffull;
		sb.cigar_cap = max_cig;
		if (sb.d_cigar.ensure(max_cig * 4 + 256)) return fail(KSW_B200_ERR_NOMEM, "CIGAR arena allocation failed");
	}
	return 0;
}

// the statistics records live in the caller's order when one device holds the whole batch (the traceback kernel scatters
// them by PairDesc::orig and the host copies the array as it is); with several devices each keeps its own pairs' order
static inline bool by_orig(const ksw_b200_batch &B) { return B.subs.size() == 1; }

// asynchronous part of a run: plan waves and enqueue every kernel of this device on the batch's stream
static int launch_sub(ksw_b200_batch &B, SubBatch &sb)
{
	sb.launched = false;
	if (sb.pairs.empty()) { sb.total_ms = sb.dp_ms = sb.tb_ms = 0; sb.launches = 0; return 0; }
	const bool cigar = !(B.flag & KSW_EZ_SCORE_ONLY);
	const bool right = (B.flag & KSW_EZ_RIGHT) != 0;
	const bool approx = (B.flag & KSW_EZ_APPROX_MAX) != 0;
	CUDA_TRY(cudaSetDevice(sb.dc->dev));
	cudaStream_t st = sb.stream ? sb.stream : sb.dc->stream;
	sb.touched = true;
	int rc = plan_waves(B, sb, cigar);
	if (rc) return rc;
	if (sb.waves.size() > 500) return fail(KSW_B200_ERR_NOMEM, "too many waves; raise KSW_B200_TB_BUDGET_MB");
	const bool stats_on = cigar && B.want_stats;
	const size_t n_stats = by_orig(B) ? (size_t)B.n : sb.pairs.size();
	if (stats_on) {
		if (sb.d_stats.ensure(n_stats * sizeof(sd_stats_t)) || sb.d_trims.ensure(n_stats * 8)) return fail(KSW_B200_ERR_NOMEM, "stats allocation failed");
		if (by_orig(B) && B.n_empty) {                                          // pairs no kernel sees
			CUDA_TRY(cudaMemsetAsync(sb.d_stats.p, 0, n_stats * sizeof(sd_stats_t), st));
			CUDA_TRY(cudaMemsetAsync(sb.d_trims.p, 0xff, n_stats * 8, st));
		}
	}
	// descriptors carry tb offsets -> (re)upload
	// (staged through pinned memory: a pageable source would make this call wait for the sequence copies queued before it)
	if (sb.h_pairs.ensure(std::max<size_t>(256, sb.pairs.size() * sizeof(PairDesc)))) return fail(KSW_B200_ERR_NOMEM, "pinned descriptor allocation failed");
	memcpy(sb.h_pairs.p, sb.pairs.data(), sb.pairs.size() * sizeof(PairDesc));
	CUDA_TRY(cudaMemcpyAsync(sb.d_pairs.p, sb.h_pairs.p, sb.pairs.size() * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
	sb.h2d_desc_bytes = sb.pairs.size() * sizeof(PairDesc);
	int *d_counters = (int *)sb.d_misc.p;
	unsigned long long *d_cursor = (unsigned long long *)((char *)sb.d_misc.p + 2048);
	int *d_overflow = (int *)((char *)sb.d_misc.p + 2048 + 64);
	CUDA_TRY(cudaMemsetAsync(sb.d_misc.p, 0, 2048 + 68, st));            // counters, cursor, overflow (not the bad-symbol flag of the upload)
	if (sb.check_limit > 0) {
		CUDA_TRY(k_check_symbols_launch((const PairDesc *)sb.d_pairs.p, (int)sb.pairs.size(), (const uint8_t *)sb.d_arena.p, sb.check_limit,
		                                (int *)((char *)sb.d_misc.p + 2048 + 68), st));
		sb.check_limit = 0; ++sb.aux_launches;
	}
	sb.launches = 0; sb.dp_ms = sb.tb_ms = 0;
	CUDA_TRY(cudaEventRecord(sb.ev[0], st));
	sb.drop_events();
	auto &dp_ev = sb.dp_ev; auto &tb_ev = sb.tb_ev;
	sb.ev_wave.clear();
	const bool conc = sb.concurrent;
	cudaEvent_t ev_fork = nullptr;
	if (conc) { CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming)); CUDA_TRY(cudaEventRecord(ev_fork, st)); }
	const int nw = (int)sb.waves.size();
	for (int wi = 0; wi < nw; ++wi) {
		// concurrent waves: the widest classes first -- their pairs are the latency-bound ones, they should own their SMs from the start
		const int wave_no = conc ? nw - 1 - wi : wi;
		const Wave &wv = sb.waves[wave_no];
		cudaStream_t ws = conc ? sb.dc->wave_streams[wi % 8] : st;
		if (conc) CUDA_TRY(cudaStreamWaitEvent(ws, ev_fork, 0));
		const int c = wv.cls;
		DpLaunch L;
		L.pairs = (const PairDesc *)sb.d_pairs.p + wv.first;
		L.results = (PairResult *)sb.d_results.p + wv.first;
		L.seq = (const uint8_t *)sb.d_arena.p;
		L.tb = (uint8_t *)sb.d_tb.p + wv.tb_base;
		L.table = (const uint32_t *)sb.d_table.p;
		L.work_counter = d_counters + wave_no;
		L.n = wv.count; L.sc = B.sc;
		int occ = dp_occupancy(*sb.dc, c, cigar, right, approx);
		if (occ <= 0) return fail(KSW_B200_ERR_CUDA, "DP kernel cannot be resident (occupancy 0)");
		const int groups_per_block = class_pairs_per_block(c);
		int grid = class_cluster(c) ? std::min(wv.count, occ)                                 // clusters, one pair each
		                               : std::min((wv.count + groups_per_block - 1) / groups_per_block, sb.dc->sms * occ);
		cudaEvent_t a, b2, c2;
		CUDA_TRY(cudaEventCreate(&a)); CUDA_TRY(cudaEventCreate(&b2)); CUDA_TRY(cudaEventCreate(&c2));
		dp_ev.push_back({a, b2}); tb_ev.push_back({b2, c2}); sb.ev_wave.push_back(wave_no);
		CUDA_TRY(cudaEventRecord(a, ws));
		CUDA_TRY(launch_dp(c, L, cigar, right, approx, grid, ws));
		CUDA_TRY(cudaEventRecord(b2, ws));
		++sb.launches;
		if (cigar) {
			TbLaunch TL;
			TL.pairs = L.pairs; TL.results = L.results; TL.tb = L.tb;
			TL.raw = B.have_raw ? (const uint8_t *)sb.d_raw.p : nullptr;
			TL.seq = L.seq;
			TL.cigar_arena = (uint32_t *)sb.d_cigar.p; TL.cigar_cursor = d_cursor; TL.cigar_capacity = sb.cigar_cap;
			TL.stats_by_orig = by_orig(B) ? 1 : 0;
			TL.stats = stats_on ? (sd_stats_t *)sb.d_stats.p + (TL.stats_by_orig ? 0 : wv.first) : nullptr;
			TL.trims = stats_on ? (int32_t *)sb.d_trims.p + 2 * (size_t)(TL.stats_by_orig ? 0 : wv.first) : nullptr;
			TL.t_match = B.mat[0]; TL.t_mismatch = B.mat[1]; TL.t_gapo = B.q; TL.t_gape = B.e;
			TL.overflow = d_overflow;
			TL.n = wv.count; TL.NS = class_ns(c); TL.flag = B.flag; TL.packed = class_packed(c) ? 1 : 0;
			TL.spare = class_spare(c, approx) ? 1 : 0;
			cudaStream_t tbs = conc ? ws : sb.dc->tb_stream;
			if (!conc) CUDA_TRY(cudaStreamWaitEvent(tbs, b2, 0));
			// long walks are latency chains: one warp per pair with staged row tiles; short ones: one thread per pair
			int64_t steps = 0;
			for (int k = wv.first; k < wv.first + wv.count; ++k) steps += (int64_t)sb.pairs[k].qlen + sb.pairs[k].tlen;
			static const int warp_min_steps = [] { const char *e = getenv("KSW_B200_TB_WARP_MIN"); return e ? atoi(e) : 6000; }();
			CUDA_TRY(k_traceback_launch(TL, steps / std::max(1, wv.count) >= warp_min_steps, stats_on, tbs));
			++sb.launches;
			CUDA_TRY(cudaEventRecord(c2, tbs));
			CUDA_TRY(cudaStreamWaitEvent(st, c2, 0));             // serial: the next wave reuses the traceback arena; concurrent: the join
		} else {
			CUDA_TRY(cudaEventRecord(c2, ws));
			if (conc) CUDA_TRY(cudaStreamWaitEvent(st, c2, 0));
		}
	}
	if (ev_fork) cudaEventDestroy(ev_fork);
	CUDA_TRY(cudaEventRecord(sb.ev[1], st));
	sb.launched = true;
	return 0;
}

// blocking part: wait for the stream, read the timers, the CIGAR cursor and the error flags
static int finish_sub(ksw_b200_batch &B, SubBatch &sb)
{
	if (!sb.launched) return 0;
	sb.launched = false;
	CUDA_TRY(cudaSetDevice(sb.dc->dev));
	cudaStream_t st = sb.stream ? sb.stream : sb.dc->stream;
	// cursor (8 bytes), overflow and bad-symbol flags in one small copy behind the kernels
	struct Tail { unsigned long long cursor; int pad[14]; int overflow; int badsym; } tail;
	static_assert(sizeof(Tail) == 72, "layout of the misc block");
	if (sb.h_pairs.cap < sizeof(Tail)) return fail(KSW_B200_ERR_NOMEM, "pinned buffer");
	CUDA_TRY(cudaMemcpyAsync(sb.h_pairs.p, (char *)sb.d_misc.p + 2048, sizeof(Tail), cudaMemcpyDeviceToHost, st));   // descriptors were consumed
	CUDA_TRY(cudaStreamSynchronize(st));
	memcpy(&tail, sb.h_pairs.p, sizeof(Tail));
	CUDA_TRY(cudaEventElapsedTime(&sb.total_ms, sb.ev[0], sb.ev[1]));
	static const bool wave_trace = [] { const char *e = getenv("KSW_B200_TRACE"); return e && atoi(e) >= 2; }();   // developer aid
	for (size_t k = 0; k < sb.dp_ev.size(); ++k) {
		float ms = 0, ms2 = 0;
		cudaEventElapsedTime(&ms, sb.dp_ev[k].first, sb.dp_ev[k].second); sb.dp_ms += ms;
		if (k < sb.tb_ev.size()) { cudaEventElapsedTime(&ms2, sb.tb_ev[k].first, sb.tb_ev[k].second); sb.tb_ms += ms2; }
		if (wave_trace && k < sb.ev_wave.size()) {
			const Wave &wv = sb.waves[sb.ev_wave[k]];
			fprintf(stderr, "[ksw_b200]   wave %d%s: class of %d slots, %d pairs, DP %.2f ms, traceback %.2f ms\n", sb.ev_wave[k], sb.concurrent ? " (concurrent)" : "",
			        class_ns(wv.cls), wv.count, ms, ms2);
		}
	}
	sb.drop_events();
	sb.cigar_used = tail.cursor;
	if (tail.badsym) return fail(KSW_B200_ERR_ARG, (B.flag & KSW_EZ_GENERIC_SC) ? "sequence symbol >= m" : "sequence symbol >= 8");
	if (tail.overflow) return fail(KSW_B200_ERR_NOMEM, "compact CIGAR arena overflow");
	return 0;
}

extern "C" int ksw_b200_batch_run(ksw_b200_batch_t *B, float *device_ms)
{
	if (!B) return fail(KSW_B200_ERR_ARG, "null batch");
	float worst = 0;
	// waves of different devices overlap: launches are asynchronous, finish_sub only blocks on its own stream
	int rc = 0;
	for (auto &sb : B->subs) { rc = launch_sub(*B, sb); if (rc) break; }                 // every device gets its work first
	for (auto &sb : B->subs) {
		// after a failure the devices that did launch are still waited for (their buffers may be released next)
		int rc2 = finish_sub(*B, sb);
		if (!rc) rc = rc2;
		worst = std::max(worst, sb.total_ms);
	}
	if (rc) return rc;
	if (device_ms) *device_ms = worst;
	return KSW_B200_OK;
}

extern "C" int ksw_b200_batch_launches(const ksw_b200_batch_t *B)
{
	int n = 0;
	if (B) for (auto &sb : B->subs) n += sb.launches;
	return n;
}
extern "C" int ksw_b200_batch_kernel_ms(const ksw_b200_batch_t *B, float *dp_ms, float *tb_ms, float *aux_ms)
{
	if (!B) return KSW_B200_ERR_ARG;
	float d = 0, t = 0, tot = 0;
	for (auto &sb : B->subs) { d = std::max(d, sb.dp_ms); t = std::max(t, sb.tb_ms); tot = std::max(tot, sb.total_ms); }
	if (dp_ms) *dp_ms = d;
	if (tb_ms) *tb_ms = t;
	if (aux_ms) *aux_ms = std::max(0.f, tot - d - t);
	return 0;
}

// ------------------------------------------------------------------------------------------------
// fetch
// ------------------------------------------------------------------------------------------------
static void reset_ez(ksw_extz_t *ez)                                              // extern/ksw2.h:153-159
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0; ez->score = ez->mqe = ez->mte = KSW_NEG_INF;
	ez->n_cigar = 0; ez->m_cigar = 0; ez->zdropped = 0;
	ez->cigar = 0;
}

// Results of a finished batch into records [s0, s0 + B.n) of the arena R (whose ez / stats arrays hold R.n records).
//   one device    the gather kernel writes final ksw_extz_t records in the CALLER's order -- host CIGAR pointers included --
//                 and the host side is three DMA copies into page-locked memory: no per-pair host work at all;
//   n devices     every device gathers its own pairs (device order), the host permutes the records into place.
static int fetch_into(ksw_b200_batch &B, ksw_b200_result &R, int s0)
{
	const bool cigar = !(B.flag & KSW_EZ_SCORE_ONLY);
	const bool stats_on = cigar && B.want_stats && R.has_stats;
	B.t_d2h = B.t_gather = 0;
	ksw_extz_t *ez = (ksw_extz_t *)R.ez.p + s0;
	sd_stats_t *stats = R.has_stats ? (sd_stats_t *)R.stats.p + s0 : nullptr;
	int32_t *trims = R.has_stats ? (int32_t *)R.trims.p + 2 * (size_t)s0 : nullptr;
	const bool one = by_orig(B);
	bool any = false;
	for (auto &sb : B.subs) any = any || !sb.pairs.empty();
	if (!one || !any) {                                                           // records no device will write
		const double t0 = now_ms();
		if (!any || B.n_empty) for (int i = 0; i < B.n; ++i) if (!any || B.is_empty[i]) reset_ez(&ez[i]);
		if (stats && (!any || B.n_empty || !stats_on)) { memset(stats, 0, sizeof(sd_stats_t) * (size_t)B.n); memset(trims, 0xff, 8 * (size_t)B.n); }
		B.t_gather += now_ms() - t0;
	} else if (stats && !stats_on) { memset(stats, 0, sizeof(sd_stats_t) * (size_t)B.n); memset(trims, 0xff, 8 * (size_t)B.n); }
	const double t_f0 = now_ms();
	for (auto &sb : B.subs) {
		if (sb.pairs.empty()) continue;
		CUDA_TRY(cudaSetDevice(sb.dc->dev));
		cudaStream_t st = sb.stream ? sb.stream : sb.dc->stream;
		const size_t np = sb.pairs.size(), nrec = one ? (size_t)B.n : np;
		if (sb.d_ez.ensure(nrec * sizeof(ksw_extz_t))) return fail(KSW_B200_ERR_NOMEM, "record buffer allocation failed");
		uint64_t host_base = 0;
		if (cigar && sb.cigar_used) {
			R.cigars.emplace_back();
			if (R.cigars.back().ensure(sb.cigar_used * 4)) return fail(KSW_B200_ERR_NOMEM, "pinned CIGAR buffer");
			host_base = (uint64_t)(uintptr_t)R.cigars.back().p;
			CUDA_TRY(cudaMemcpyAsync(R.cigars.back().p, sb.d_cigar.p, sb.cigar_used * 4, cudaMemcpyDeviceToHost, st));
		}
		if (one && B.n_empty) { CUDA_TRY(k_fill_reset_launch((uint64_t *)sb.d_ez.p, B.n, st)); ++sb.aux_launches; }
		GatherLaunch G;
		G.pairs = (const PairDesc *)sb.d_pairs.p; G.results = (const PairResult *)sb.d_results.p;
		G.ez_out = (uint64_t *)sb.d_ez.p; G.host_cigar_base = host_base; G.n = (int)np; G.by_orig = one ? 1 : 0; G.with_cigar = cigar ? 1 : 0;
		CUDA_TRY(k_gather_launch(G, st));
		++sb.aux_launches;
		sb.d2h_bytes = nrec * sizeof(ksw_extz_t) + (cigar ? sb.cigar_used * 4 : 0);
		if (one) {
			CUDA_TRY(cudaMemcpyAsync(ez, sb.d_ez.p, nrec * sizeof(ksw_extz_t), cudaMemcpyDeviceToHost, st));
			if (stats_on) {
				CUDA_TRY(cudaMemcpyAsync(stats, sb.d_stats.p, nrec * sizeof(sd_stats_t), cudaMemcpyDeviceToHost, st));
				CUDA_TRY(cudaMemcpyAsync(trims, sb.d_trims.p, nrec * 8, cudaMemcpyDeviceToHost, st));
				sb.d2h_bytes += nrec * (sizeof(sd_stats_t) + 8);
			}
		} else {
			if (sb.h_recs.ensure(np * sizeof(ksw_extz_t))) return fail(KSW_B200_ERR_NOMEM, "pinned record buffer");
			CUDA_TRY(cudaMemcpyAsync(sb.h_recs.p, sb.d_ez.p, np * sizeof(ksw_extz_t), cudaMemcpyDeviceToHost, st));
			if (stats_on) {
				if (sb.h_stats.ensure(np * sizeof(sd_stats_t)) || sb.h_trims.ensure(np * 8)) return fail(KSW_B200_ERR_NOMEM, "pinned stats buffer");
				CUDA_TRY(cudaMemcpyAsync(sb.h_stats.p, sb.d_stats.p, np * sizeof(sd_stats_t), cudaMemcpyDeviceToHost, st));
				CUDA_TRY(cudaMemcpyAsync(sb.h_trims.p, sb.d_trims.p, np * 8, cudaMemcpyDeviceToHost, st));
				sb.d2h_bytes += np * (sizeof(sd_stats_t) + 8);
			}
		}
	}
	for (auto &sb : B.subs) {
		if (sb.pairs.empty()) continue;
		CUDA_TRY(cudaSetDevice(sb.dc->dev));
		CUDA_TRY(cudaStreamSynchronize(sb.stream ? sb.stream : sb.dc->stream));
	}
	B.t_d2h += now_ms() - t_f0;
	if (!one) {
		const double t_g0 = now_ms();
		for (auto &sb : B.subs) {
			const int64_t np = (int64_t)sb.pairs.size();
			const ksw_extz_t *rec = (const ksw_extz_t *)sb.h_recs.p;
			const sd_stats_t *hst = (const sd_stats_t *)sb.h_stats.p;
			const int64_t *htr = (const int64_t *)sb.h_trims.p;
#pragma omp parallel for num_threads(host_threads()) schedule(static) if (np >= 4096)
			for (int64_t k = 0; k < np; ++k) {
				const int i = sb.pairs[k].orig;
				ez[i] = rec[k];
				if (stats_on) { stats[i] = hst[k]; ((int64_t *)trims)[i] = htr[k]; }
			}
		}
		B.t_gather += now_ms() - t_g0;
	}
	return KSW_B200_OK;
}

static ksw_b200_result *result_new(int n, bool with_stats, int *err)
{
	ksw_b200_result *R = new ksw_b200_result();
	R->n = n; R->has_stats = with_stats;
	if (R->ez.ensure(std::max<size_t>(1, (size_t)n) * sizeof(ksw_extz_t)) ||
	    (with_stats && (R->stats.ensure(std::max<size_t>(1, (size_t)n) * sizeof(sd_stats_t)) || R->trims.ensure(std::max<size_t>(1, (size_t)n) * 8)))) {
		R->release(); delete R;
		if (err) *err = fail(KSW_B200_ERR_NOMEM, "pinned result arena allocation failed");
		return nullptr;
	}
	return R;
}
extern "C" void ksw_b200_result_free(ksw_b200_result_t *R) { if (R) { R->release(); delete R; } }
extern "C" const ksw_extz_t *ksw_b200_result_ez(const ksw_b200_result_t *R) { return R ? (const ksw_extz_t *)R->ez.p : nullptr; }
extern "C" const sd_stats_t *ksw_b200_result_stats(const ksw_b200_result_t *R) { return (R && R->has_stats) ? (const sd_stats_t *)R->stats.p : nullptr; }
extern "C" const int32_t *ksw_b200_result_trims(const ksw_b200_result_t *R) { return (R && R->has_stats) ? (const int32_t *)R->trims.p : nullptr; }
extern "C" int ksw_b200_result_count(const ksw_b200_result_t *R) { return R ? R->n : 0; }
extern "C" void ksw_b200_result_io(const ksw_b200_result_t *R, int64_t *h2d, int64_t *d2h, int *launches)
{
	if (h2d) *h2d = R ? R->h2d : 0;
	if (d2h) *d2h = R ? R->d2h : 0;
	if (launches) *launches = R ? R->launches : 0;
}

// Position-independent copy of an arena into CALLER-OWNED buffers (e.g. a shared-memory segment another process reads: the
// gather step of a one-process-per-GPU deployment).  Record i goes to ez_dst[index ? index[i] : i] (and its statistics
// likewise); its CIGAR words are appended to cigar_dst in the arena's order and `cigar` becomes the WORD OFFSET
// cigar_base + (position in cigar_dst) instead of a pointer.  Returns the number of CIGAR words written, or a negative
// error code (KSW_B200_ERR_NOMEM when cigar_cap words are not enough).
extern "C" int64_t ksw_b200_result_export(const ksw_b200_result_t *R, ksw_extz_t *ez_dst, sd_stats_t *stats_dst,
                                          uint32_t *cigar_dst, int64_t cigar_cap, int64_t cigar_base, const int64_t *index)
{
	if (!R || (R->n > 0 && !ez_dst)) return fail(KSW_B200_ERR_ARG, "null argument");
	const int n = R->n;
	const ksw_extz_t *src = (const ksw_extz_t *)R->ez.p;
	const sd_stats_t *sst = R->has_stats ? (const sd_stats_t *)R->stats.p : nullptr;
	const int nt = threads_for(n, 8192);
	std::vector<int64_t> part(nt + 1, 0);
#pragma omp parallel num_threads(nt)
	{
		const int t = omp_get_thread_num(), T = omp_get_num_threads();
		const int lo = (int)((int64_t)n * t / T), hi = (int)((int64_t)n * (t + 1) / T);
		int64_t words = 0;
		for (int i = lo; i < hi; ++i) words += src[i].n_cigar;
		part[t + 1] = words;
#pragma omp barrier
#pragma omp single
		for (int k = 0; k < T; ++k) part[k + 1] += part[k];
		int64_t pos = part[t];
		if (part[T] <= cigar_cap || !cigar_dst)
			for (int i = lo; i < hi; ++i) {
				const int64_t d = index ? index[i] : i;
				ksw_extz_t z = src[i];
				if (z.n_cigar > 0 && cigar_dst) memcpy(cigar_dst + pos, src[i].cigar, (size_t)z.n_cigar * 4);
				z.cigar = (uint32_t *)(uintptr_t)(z.n_cigar > 0 ? cigar_base + pos : 0);
				pos += z.n_cigar;
				ez_dst[d] = z;
				if (stats_dst) { if (sst) stats_dst[d] = sst[i]; else memset(&stats_dst[d], 0, sizeof(sd_stats_t)); }
			}
	}
	int64_t total = 0;
	for (int i = 0; i < n; ++i) total += src[i].n_cigar;
	if (cigar_dst && total > cigar_cap) return fail(KSW_B200_ERR_NOMEM, "cigar_cap too small");
	return total;
}

// arena records -> the ksw2 ownership contract: every CIGAR in its own malloc() block (caller free()s, src/align.cc:65)
static int copy_out_malloc(const ksw_extz_t *src, const sd_stats_t *src_stats, int n, ksw_extz_t *ez, sd_stats_t *stats)
{
	int nomem = 0;
#pragma omp parallel for num_threads(host_threads()) schedule(static) reduction(| : nomem) if (n >= 2048)
	for (int i = 0; i < n; ++i) {
		ez[i] = src[i];
		if (src[i].n_cigar > 0) {
			ez[i].cigar = (uint32_t *)malloc((size_t)src[i].m_cigar << 2);
			if (!ez[i].cigar) { nomem |= 1; ez[i].n_cigar = ez[i].m_cigar = 0; continue; }
			memcpy(ez[i].cigar, src[i].cigar, (size_t)src[i].n_cigar * 4);
		}
	}
	if (stats) { if (src_stats) memcpy(stats, src_stats, sizeof(sd_stats_t) * (size_t)n); else memset(stats, 0, sizeof(sd_stats_t) * (size_t)n); }
	return nomem ? fail(KSW_B200_ERR_NOMEM, "malloc of a CIGAR failed") : KSW_B200_OK;
}

extern "C" void ksw_b200_free_cigars(ksw_extz_t *ez, int n)
{
	if (!ez) return;
	// serial on purpose: the blocks were malloc'ed by the gather threads, so a parallel loop frees into foreign glibc arenas
	// and contends on their locks (measured on the 32-thread host: 3.2 ms serial vs 4.7 ms parallel for 100k CIGARs)
	for (int i = 0; i < n; ++i) { free(ez[i].cigar); ez[i].cigar = nullptr; ez[i].n_cigar = ez[i].m_cigar = 0; }
}

extern "C" int ksw_b200_batch_fetch(ksw_b200_batch_t *B, ksw_extz_t *ez, sd_stats_t *stats)
{
	if (!B || (B->n > 0 && !ez)) return fail(KSW_B200_ERR_ARG, "null argument");
	for (int i = 0; i < B->n; ++i) reset_ez(&ez[i]);
	int err = 0;
	ksw_b200_result *R = result_new(B->n, stats != nullptr, &err);
	if (!R) return err;
	int rc = fetch_into(*B, *R, 0);
	if (rc == 0) {
		const double t0 = now_ms();
		rc = copy_out_malloc((const ksw_extz_t *)R->ez.p, R->has_stats ? (const sd_stats_t *)R->stats.p : nullptr, B->n, ez, stats);
		B->t_gather += now_ms() - t0;
		if (rc) ksw_b200_free_cigars(ez, B->n);
	}
	ksw_b200_result_free(R);
	return rc;
}
// the same without the per-pair malloc: records and CIGARs stay in the arena `*out` (ksw_b200_result_free)
extern "C" int ksw_b200_batch_fetch_arena(ksw_b200_batch_t *B, int want_stats, ksw_b200_result_t **out)
{
	if (!B || !out) return fail(KSW_B200_ERR_ARG, "null argument");
	*out = nullptr;
	int err = 0;
	ksw_b200_result *R = result_new(B->n, want_stats != 0, &err);
	if (!R) return err;
	int rc = fetch_into(*B, *R, 0);
	if (rc) { ksw_b200_result_free(R); return rc; }
	ksw_b200_batch_io_bytes(B, &R->h2d, &R->d2h);
	R->launches = ksw_b200_batch_launches(B);
	*out = R;
	return KSW_B200_OK;
}

extern "C" void ksw_b200_batch_free(ksw_b200_batch_t *B)
{
	if (!B) return;
	for (auto &sb : B->subs) { if (sb.dc) cudaSetDevice(sb.dc->dev); sb.release(); }
	delete B;
}

// ------------------------------------------------------------------------------------------------
// one-shot entry points
// ------------------------------------------------------------------------------------------------
// I/O accounting of the last one-shot call of this thread
static thread_local int64_t g_last_h2d = 0, g_last_d2h = 0;
static thread_local int g_last_launches = 0;
extern "C" void ksw_b200_last_call_io(int64_t *h2d, int64_t *d2h, int *launches)
{
	if (h2d) *h2d = g_last_h2d;
	if (d2h) *d2h = g_last_d2h;
	if (launches) *launches = g_last_launches;
}

struct BatchArgs {
	int n; const int *qlen; const int64_t *qoff; const uint8_t *qbuf; const int *tlen; const int64_t *toff; const uint8_t *tbuf;
	int8_t m; const int8_t *mat; int8_t q, e; int w, zdrop, flag; const uint8_t *q_raw_buf, *t_raw_buf;
};

// One-shot batch into the arena R.  Large batches are cut into chunks that flow through a software pipeline: a producer
// thread plans chunk k+1, queues its copies and LAUNCHES its kernels (own stream) while this thread waits for chunk k and
// fetches its results -- H2D, kernels and D2H overlap.  `on_chunk(s0, cnt)` runs after records [s0, s0 + cnt) are final.
static int batch_core(const BatchArgs &A, ksw_b200_result &R, const std::function<int(int, int)> &on_chunk)
{
	const int n = A.n;
	const char *env = getenv("KSW_B200_CHUNK_PAIRS");
	const int chunk_pairs = env ? std::max(1, atoi(env)) : 25000;
	// chunk sizes ramp up (1 : 2 : 4 : 4 ...) so that the first H2D copy -- the only one nothing can hide -- is short
	int nchunks = std::max(1, std::min(16, n / chunk_pairs + (n >= 2 * chunk_pairs ? 2 : 0)));
	if (!env && nchunks > 1) {
		// A chunk also has to be worth a set of kernel launches.  Every chunk runs one DP + one traceback kernel PER CLASS,
		// each with a latency floor of one pair (1-2 ms) and a tail, so chunks of mid-size pairs that carry little work
		// under-fill the GPU (100k pairs of <= 250 bp: 18 ms resident, 45 ms in six chunks; profiles/r01_tuning.md).  Keep
		// at least ~2.5 G cells per chunk.
		const int step = std::max(1, n / 512);
		int64_t cells = 0; int cnt = 0;
		for (int i = 0; i < n; i += step, ++cnt) {
			const int ql = std::max(0, A.qlen[i]), tl = std::max(0, A.tlen[i]);
			cells += est_cells(ql, tl, A.w < 0 ? std::max(ql, tl) : std::min(A.w, std::max(ql, tl)));
		}
		const double avg = cnt ? (double)cells / cnt : 0.0;
		if (avg >= 2000.0) nchunks = std::max(1, std::min(nchunks, (int)(avg * n / 2.5e9 + 0.5)));
		// tiny pairs: the kernels need ~5 ns per pair and the host 40-50 ns to plan and describe it (the producer thread is
		// the bottleneck), plus ~0.4 ms of fixed cost per chunk (stream, events, small copies): chunks of >= 50k pairs
		else nchunks = std::max(1, std::min(nchunks, n / 50000));
	}
	std::vector<int> start(nchunks + 1, 0);
	{
		std::vector<double> wgt(nchunks, 4.0);
		if (nchunks >= 3) { wgt[0] = 1.0; wgt[1] = 2.0; }
		if (nchunks >= 5) wgt[nchunks - 1] = 2.0;                      // a short last chunk keeps the un-overlapped D2H small
		double tot = 0; for (double x : wgt) tot += x;
		double acc = 0;
		for (int c = 0; c < nchunks; ++c) { acc += wgt[c]; start[c + 1] = (int)((double)n * acc / tot + 0.5); }
		start[nchunks] = n;
	}

	const bool trace = getenv("KSW_B200_TRACE") != nullptr;                // developer aid: per-chunk timeline on stderr
	const double t_call = now_ms();
	auto mark = [&](const char *what, int c) { if (trace) fprintf(stderr, "[ksw_b200] %8.2f ms  chunk %d  %s\n", now_ms() - t_call, c, what); };
	auto upload_chunk = [&](int c, int *err) -> ksw_b200_batch_t * {
		const int s0 = start[c], cnt = start[c + 1] - s0;
		tl_async_upload = nchunks > 1;
		tl_pipeline_depth = nchunks > 1 ? 3 : 1;
		struct Reset { ~Reset() { tl_async_upload = false; tl_pipeline_depth = 1; } } reset;
		ksw_b200_batch_t *B = ksw_b200_batch_upload(cnt, A.qlen + s0, A.qoff + s0, A.qbuf, A.tlen + s0, A.toff + s0, A.tbuf, A.m, A.mat, A.q, A.e,
		                                           A.w, A.zdrop, A.flag, A.q_raw_buf, A.t_raw_buf, err);
		if (B && !R.has_stats) B->want_stats = false;
		return B;
	};
	auto consume = [&](ksw_b200_batch_t *B, int c, bool launched) -> int {
		const int s0 = start[c];
		int rc = 0;
		mark("wait", c);
		if (!launched) rc = ksw_b200_batch_run(B, nullptr);
		else for (auto &sb : B->subs) { int rc2 = finish_sub(*B, sb); if (!rc) rc = rc2; }
		mark("kernels done", c);
		if (rc == 0) rc = fetch_into(*B, R, s0);
		mark("fetched", c);
		int64_t a = 0, d = 0; ksw_b200_batch_io_bytes(B, &a, &d);
		R.h2d += a; R.d2h += d; R.launches += ksw_b200_batch_launches(B);
		if (trace) fprintf(stderr, "[ksw_b200]             chunk %d  d2h %.2f ms, gather %.2f ms\n", c, B->t_d2h, B->t_gather);
		ksw_b200_batch_free(B);
		if (rc == 0 && on_chunk) rc = on_chunk(s0, start[c + 1] - s0);
		mark("handed over", c);
		return rc;
	};
	if (nchunks == 1) {
		int err = 0;
		ksw_b200_batch_t *B = upload_chunk(0, &err);
		if (!B) return err;
		return consume(B, 0, false);
	}
	// producer / consumer over chunks
	std::mutex mu; std::condition_variable cv;
	std::vector<ksw_b200_batch_t *> ready(nchunks, nullptr);
	std::vector<int> errs(nchunks, 0), done(nchunks, 0);
	int consumed = 0; bool abort_all = false;
	std::string producer_error;
	auto producer_fn = [&] {
		for (int c = 0; c < nchunks; ++c) {
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&] { return abort_all || c - consumed <= 2; });       // at most two chunks ahead (bounds pinned + traceback memory)
				if (abort_all) return;
			}
			int err = 0;
			mark("upload begin", c);
			ksw_b200_batch_t *B = upload_chunk(c, &err);
			mark("upload end", c);
			if (trace && B) fprintf(stderr, "[ksw_b200]             chunk %d  plan %.2f ms, pack %.2f ms, %d pairs\n", c, B->t_plan, B->t_pack, B->n);
			if (B) {                                                        // enqueue the kernels right behind the H2D copy
				for (auto &sb : B->subs) { err = launch_sub(*B, sb); if (err) break; }
				if (err) { std::string keep = g_last_error; ksw_b200_batch_free(B); B = nullptr; g_last_error = keep; }
			}
			mark("launched", c);
			{
				std::lock_guard<std::mutex> lk(mu);
				ready[c] = B; errs[c] = err; done[c] = 1;
				if (!B) producer_error = g_last_error;
			}
			cv.notify_all();
			if (!B) return;
		}
	};
	const bool own_worker = g_worker.claim.try_lock();
	std::thread producer;
	if (own_worker) g_worker.start(producer_fn); else producer = std::thread(producer_fn);
	int rc = 0;
	std::string first_error;
	for (int c = 0; c < nchunks && rc == 0; ++c) {
		ksw_b200_batch_t *B = nullptr;
		{
			std::unique_lock<std::mutex> lk(mu);
			cv.wait(lk, [&] { return done[c] != 0; });
			B = ready[c];
			if (!B) { rc = errs[c]; g_last_error = producer_error; }
		}
		if (B) rc = consume(B, c, true);
		if (rc) first_error = g_last_error;
		{
			std::lock_guard<std::mutex> lk(mu);
			consumed = c + 1;
			if (rc) abort_all = true;
		}
		cv.notify_all();
	}
	{ std::lock_guard<std::mutex> lk(mu); abort_all = abort_all || rc != 0; }
	cv.notify_all();
	if (own_worker) { g_worker.join(); g_worker.claim.unlock(); } else producer.join();
	if (rc) {                                                           // chunks launched but not consumed: wait for them, then release
		for (int c = consumed; c < nchunks; ++c) if (ready[c]) ksw_b200_batch_free(ready[c]);
		g_last_error = first_error;
	}
	return rc;
}

static int check_batch_args(const BatchArgs &A)
{
	const bool have_codes = A.qbuf && A.tbuf, have_raw = A.q_raw_buf && A.t_raw_buf;
	if (A.n < 0 || (A.n > 0 && (!A.qlen || !A.tlen || !A.qoff || !A.toff || (!have_codes && !have_raw))) || (A.m > 0 && !A.mat))
		return fail(KSW_B200_ERR_ARG, "null pointer or negative count");
	return 0;
}

extern "C" int ksw_extz2_batch_arena(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                                     const int *tlen, const int64_t *toff, const uint8_t *tbuf,
                                     int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                                     int want_stats, const uint8_t *q_raw_buf, const uint8_t *t_raw_buf,
                                     ksw_b200_result_t **out)
{
	g_last_h2d = g_last_d2h = 0; g_last_launches = 0;
	if (!out) return fail(KSW_B200_ERR_ARG, "null result pointer");
	*out = nullptr;
	const BatchArgs A{n, qlen, qoff, qbuf, tlen, toff, tbuf, m, mat, q, e, w, zdrop, flag, q_raw_buf, t_raw_buf};
	int rc = check_batch_args(A);
	if (rc) return rc;
	int nd = ensure_init();
	if (nd <= 0) return nd;
	int err = 0;
	ksw_b200_result *R = result_new(n, want_stats && !(flag & KSW_EZ_SCORE_ONLY), &err);
	if (!R) return err;
	rc = batch_core(A, *R, nullptr);
	g_last_h2d = R->h2d; g_last_d2h = R->d2h; g_last_launches = R->launches;
	if (rc) { std::string keep = g_last_error; ksw_b200_result_free(R); g_last_error = keep; return rc; }
	*out = R;
	return KSW_B200_OK;
}

extern "C" int ksw_extz2_batch_flat(int n, const int *qlen, const int64_t *qoff, const uint8_t *qbuf,
                                    const int *tlen, const int64_t *toff, const uint8_t *tbuf,
                                    int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                                    ksw_extz_t *ez, sd_stats_t *stats,
                                    const uint8_t *q_raw_buf, const uint8_t *t_raw_buf)
{
	g_last_h2d = g_last_d2h = 0; g_last_launches = 0;
	const BatchArgs A{n, qlen, qoff, qbuf, tlen, toff, tbuf, m, mat, q, e, w, zdrop, flag, q_raw_buf, t_raw_buf};
	int rc = check_batch_args(A);
	if (rc) return rc;
	if (n > 0 && !ez) return fail(KSW_B200_ERR_ARG, "null result array");
	// ez[] is the caller's storage and may be uninitialised (the ksw_extz2_sse contract: fully overwritten): reset it FIRST so
	// that the error paths below only ever free() pointers this call allocated
	for (int i = 0; i < n; ++i) reset_ez(&ez[i]);
	if (stats) memset(stats, 0, sizeof(sd_stats_t) * (size_t)n);
	int nd = ensure_init();
	if (nd <= 0) return nd;
	int err = 0;
	ksw_b200_result *R = result_new(n, stats != nullptr && !(flag & KSW_EZ_SCORE_ONLY), &err);
	if (!R) return err;
	// every finished chunk is converted to the ksw2 ownership contract (one malloc() per CIGAR) while the next one computes
	rc = batch_core(A, *R, [&](int s0, int cnt) {
		return copy_out_malloc((const ksw_extz_t *)R->ez.p + s0, R->has_stats ? (const sd_stats_t *)R->stats.p + s0 : nullptr, cnt,
		                       ez + s0, (stats && R->has_stats) ? stats + s0 : nullptr);
	});
	g_last_h2d = R->h2d; g_last_d2h = R->d2h; g_last_launches = R->launches;
	std::string keep = g_last_error;
	ksw_b200_result_free(R);
	if (rc) { ksw_b200_free_cigars(ez, n); for (int i = 0; i < n; ++i) reset_ez(&ez[i]); g_last_error = keep; }
	return rc;
}

extern "C" int ksw_extz2_batch(int n, const int *qlen, const uint8_t *const *query,
                               const int *tlen, const uint8_t *const *target,
                               int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                               ksw_extz_t *ez, sd_stats_t *stats,
                               const uint8_t *const *q_raw, const uint8_t *const *t_raw)
{
	if (n < 0 || (n > 0 && (!qlen || !tlen || !query || !target))) return fail(KSW_B200_ERR_ARG, "null pointer or negative count");
	// gather into flat buffers (offsets are relative to the lowest pointer would be fragile; copy instead)
	std::vector<int64_t> qoff(n), toff(n);
	int64_t qs = 0, ts = 0;
	for (int i = 0; i < n; ++i) { qoff[i] = qs; toff[i] = ts; qs += std::max(0, qlen[i]); ts += std::max(0, tlen[i]); }
	std::vector<uint8_t> qb(qs + 1), tb(ts + 1), qr, tr;
	const bool raw = q_raw && t_raw;
	if (raw) { qr.resize(qs + 1); tr.resize(ts + 1); }
	for (int i = 0; i < n; ++i) {
		if (qlen[i] > 0) { memcpy(&qb[qoff[i]], query[i], qlen[i]); if (raw) memcpy(&qr[qoff[i]], q_raw[i], qlen[i]); }
		if (tlen[i] > 0) { memcpy(&tb[toff[i]], target[i], tlen[i]); if (raw) memcpy(&tr[toff[i]], t_raw[i], tlen[i]); }
	}
	return ksw_extz2_batch_flat(n, qlen, qoff.data(), qb.data(), tlen, toff.data(), tb.data(), m, mat, q, e, w, zdrop, flag,
	                            ez, stats, raw ? qr.data() : nullptr, raw ? tr.data() : nullptr);
}

// ksw_extz2_sse has no error channel (it exit()s when kmalloc fails); neither has its replacement.  A request the engine
// cannot serve -- no device, a pair wider than the widest kernel -- must not come back looking like "no alignment", so the
// default is to report and abort(); an integrator can install a handler instead (which may longjmp / throw / log and return,
// in which case the caller sees the reset record).
static void default_fatal(int code, const char *msg)
{
	fprintf(stderr, "[ksw_extz2_b200] fatal: %s: %s\n", ksw_b200_strerror(code), msg);
	abort();
}
static ksw_b200_fatal_fn g_fatal = default_fatal;
extern "C" void ksw_b200_set_fatal_handler(ksw_b200_fatal_fn fn) { g_fatal = fn ? fn : default_fatal; }

extern "C" void ksw_extz2_b200(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                               int8_t m, const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int flag,
                               ksw_extz_t *ez)
{
	(void)km;
	reset_ez(ez);
	if (m <= 0 || qlen <= 0 || tlen <= 0) return;                                   // :57
	int64_t zero = 0;
	int rc = ksw_extz2_batch_flat(1, &qlen, &zero, query, &tlen, &zero, target, m, mat, q, e, w, zdrop, flag, ez, nullptr, nullptr, nullptr);
	if (rc) {
		reset_ez(ez);
		g_fatal(rc, ksw_b200_last_error());
	}
}

// ------------------------------------------------------------------------------------------------
// Alignment(fa, fb, cigar): statistics from existing CIGARs
// ------------------------------------------------------------------------------------------------
extern "C" int sd_stats_from_cigar_batch_flat(int n, const int64_t *cig_off, const int64_t *n_cigar, const uint32_t *cig_buf,
                                              const int *alen, const int64_t *aoff, const uint8_t *abuf,
                                              const int *blen, const int64_t *boff, const uint8_t *bbuf,
                                              sd_stats_t *out, int *status)
{
	if (n < 0 || (n > 0 && (!cig_off || !n_cigar || !cig_buf || !alen || !aoff || !abuf || !blen || !boff || !bbuf || !out || !status)))
		return fail(KSW_B200_ERR_ARG, "null pointer or negative count");
	if (n == 0) return 0;
	int nd = ensure_init();
	if (nd <= 0) return nd;
	DevCtx &dc = g_devs[0];
	CUDA_TRY(cudaSetDevice(dc.dev));
	size_t cig_total = 0, a_total = 0, b_total = 0;
	for (int i = 0; i < n; ++i) {
		cig_total = std::max<size_t>(cig_total, (size_t)(cig_off[i] + n_cigar[i]));
		a_total = std::max<size_t>(a_total, (size_t)(aoff[i] + std::max(0, alen[i])));
		b_total = std::max<size_t>(b_total, (size_t)(boff[i] + std::max(0, blen[i])));
	}
	DevBuf d_cig, d_a, d_b, d_meta, d_out;
	const size_t meta = (size_t)n * (4 * sizeof(int64_t) + 2 * sizeof(int));
	if (d_cig.ensure(cig_total * 4 + 16) || d_a.ensure(a_total + 16) || d_b.ensure(b_total + 16) || d_meta.ensure(meta + 64) ||
	    d_out.ensure((size_t)n * (sizeof(sd_stats_t) + sizeof(int)) + 64)) {
		d_cig.release(); d_a.release(); d_b.release(); d_meta.release(); d_out.release();
		return fail(KSW_B200_ERR_NOMEM, "device allocation failed");
	}
	cudaStream_t st = dc.stream;
	char *mp = (char *)d_meta.p;
	int64_t *m_coff = (int64_t *)mp, *m_cn = m_coff + n, *m_ao = m_cn + n, *m_bo = m_ao + n;
	int *m_al = (int *)(m_bo + n), *m_bl = m_al + n;
	bool ok = cudaMemcpyAsync(d_cig.p, cig_buf, cig_total * 4, cudaMemcpyHostToDevice, st) == cudaSuccess &&
	          cudaMemcpyAsync(d_a.p, abuf, a_total, cudaMemcpyHostToDevice, st) == cudaSuccess &&
	          cudaMemcpyAsync(d_b.p, bbuf, b_total, cudaMemcpyHostToDevice, st) == cudaSuccess &&
	          cudaMemcpyAsync(m_coff, cig_off, n * sizeof(int64_t), cudaMemcpyHostToDevice, st) == cudaSuccess &&
	          cudaMemcpyAsync(m_cn, n_cigar, n * sizeof(int64_t), cudaMemcpyHostToDevice, st) == cudaSuccess &&
	          cudaMemcpyAsync(m_ao, aoff, n * sizeof(int64_t), cudaMemcpyHostToDevice, st) == cudaSuccess &&
	          cudaMemcpyAsync(m_bo, boff, n * sizeof(int64_t), cudaMemcpyHostToDevice, st) == cudaSuccess &&
	          cudaMemcpyAsync(m_al, alen, n * sizeof(int), cudaMemcpyHostToDevice, st) == cudaSuccess &&
	          cudaMemcpyAsync(m_bl, blen, n * sizeof(int), cudaMemcpyHostToDevice, st) == cudaSuccess;
	sd_stats_t *d_stats = (sd_stats_t *)d_out.p;
	int *d_status = (int *)(d_stats + n);
	if (ok) {
		CigarStatsLaunch L{(const uint32_t *)d_cig.p, m_coff, m_cn, (const uint8_t *)d_a.p, m_ao, m_al, (const uint8_t *)d_b.p, m_bo, m_bl,
		                   d_stats, d_status, n};
		ok = k_stats_from_cigar_launch(L, st) == cudaSuccess &&
		     cudaMemcpyAsync(out, d_stats, (size_t)n * sizeof(sd_stats_t), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
		     cudaMemcpyAsync(status, d_status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
		     cudaStreamSynchronize(st) == cudaSuccess;
	}
	std::string msg = ok ? "" : cudaGetErrorString(cudaGetLastError());
	d_cig.release(); d_a.release(); d_b.release(); d_meta.release(); d_out.release();
	return ok ? 0 : fail(KSW_B200_ERR_CUDA, "sd_stats_from_cigar: " + msg);
}

// ------------------------------------------------------------------------------------------------
// host-side floating point fields (src/stats_main.cc:273-283,297-299; src/align.h:84-92)
// ------------------------------------------------------------------------------------------------
extern "C" void sd_stats_derive_fp(const sd_stats_t *s, sd_stats_fp_t *o)
{
	o->fracMatch = double(s->matchB) / (s->alnB);
	o->fracMatchIndel = double(s->matchB) / (s->span);
	double jcp = double(s->mismatchB) / (s->alnB);
	o->jcK = -0.75 * log(1.0 - 4.0 / 3 * jcp);
	double p = double(s->transitionsB) / (s->alnB);
	double q = double(s->transversionsB) / (s->alnB);
	double w1 = 1.0 / (1 - 2.0 * p - q);
	double w2 = 1.0 / (1 - 2.0 * q);
	o->k2K = 0.5 * log(w1) + 0.25 * log(w2);
	o->errorScaled = (s->gaps + s->mismatches) / double(s->gaps + s->mismatches + s->matches);
	o->filter_score = 1 - o->errorScaled;
	double tot = s->matches + s->gap_bases + s->mismatches;
	o->gap_error = 100.0 * s->gap_bases / tot;                                       // src/common.h:99 pct()
	o->mismatch_error = 100.0 * s->mismatches / tot;
	o->total_error = o->mismatch_error + o->gap_error;
}
