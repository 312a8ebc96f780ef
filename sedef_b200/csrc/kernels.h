// kernels.h -- host-callable launchers of the CUDA kernels, one translation unit per kernel family (k_*.cu) so that the
// library builds in parallel.  engine.cu holds no device code; it picks a kernel class and calls these.
//
// Function attributes (dynamic shared-memory opt-in) are per DEVICE: every launcher / occupancy query that needs one sets
// it on the current device right before use (cheap, and correct for in-process multi-device batches).
#pragma once
#include <cuda_runtime.h>
#include "extz_core.cuh"
#include "launch_structs.h"

namespace extz {

// ---- packed kernels (extz_dp16.cuh) -----------------------------------------------------------------------------------
// `approx`: the KSW_EZ_APPROX_MAX variant (one tracked score instead of the exact maximum; no H row)
// narrow: G in {1,2,4,8,16,32} lanes x 32 slots per pair, 128-thread CTAs
cudaError_t k_dp16_launch(int G, const DpLaunch &L, bool cigar, bool right, bool approx, int grid, cudaStream_t st);
int k_dp16_occupancy(int G, bool cigar, bool right, bool approx);
// CTA-wide: G in {64,128,256} lanes per pair
cudaError_t k_dp16_wide_launch(int G, const DpLaunch &L, bool cigar, bool right, bool approx, int grid, cudaStream_t st);
int k_dp16_wide_occupancy(int G, bool cigar, bool right, bool approx);
// cluster of C = 2 / 4 / 8 CTAs x 256 lanes; max_clusters != nullptr: occupancy query only
cudaError_t k_dp16_cluster_dispatch(int C, const DpLaunch &L, bool cigar, bool right, bool approx, int nclusters, cudaStream_t st, int *max_clusters);

// ---- one-slot kernels (extz_dp.cuh; KSW_B200_PACKED=0 A/B path) ----------------------------------------------------------
// c: index into the one-slot class table of engine.cu (0..8); cluster classes go through k_dp_cluster_dispatch
cudaError_t k_dp_launch(int c, const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st);
int k_dp_occupancy(int c, bool cigar, bool right);
cudaError_t k_dp_cluster_dispatch(int C, const DpLaunch &L, bool cigar, bool right, int nclusters, cudaStream_t st, int *max_clusters);

// ---- traceback + statistics (extz_tb.cuh) ---------------------------------------------------------------------------------
cudaError_t k_traceback_launch(const TbLaunch &L, bool warp_per_pair, bool stats, cudaStream_t st);
cudaError_t k_stats_from_cigar_launch(const CigarStatsLaunch &L, cudaStream_t st);

// ---- device-side input preparation and output gather (extz_io.cuh) --------------------------------------------------------
cudaError_t k_encode_launch(const uint8_t *raw, uint8_t *codes, size_t nbytes, cudaStream_t st);
cudaError_t k_check_symbols_launch(const PairDesc *pairs, int n, const uint8_t *codes, int limit, int *flag, cudaStream_t st);
cudaError_t k_fill_reset_launch(uint64_t *ez_out, int n, cudaStream_t st);
cudaError_t k_gather_launch(const GatherLaunch &L, cudaStream_t st);

} // namespace extz
