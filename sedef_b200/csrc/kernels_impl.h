// kernels_impl.h -- per-instantiation launchers behind kernels.h.  The templates are DEFINED in k_*.cuh and explicitly
// instantiated in the k_*.cu translation units; k_dispatch.cu only sees these declarations.
#pragma once
#include "kernels.h"

namespace extz {

template <int G, bool A> cudaError_t dp16_launch_g(const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st);
template <int G, bool A> int dp16_occupancy_g(bool cigar, bool right);
template <int G, bool A> cudaError_t dp16_wide_launch_g(const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st);
template <int G, bool A> int dp16_wide_occupancy_g(bool cigar, bool right);
template <bool A> cudaError_t dp16_cluster_dispatch_a(int C, const DpLaunch &L, bool cigar, bool right, int nclusters, cudaStream_t st, int *max_clusters);

template <int G, int S, bool W> cudaError_t dp1_launch_gs(const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st);
template <int G, int S, bool W> int dp1_occupancy_gs(bool cigar, bool right);
template <int C> cudaError_t dp1_cluster_dispatch_c(const DpLaunch &L, bool cigar, bool right, int nclusters, cudaStream_t st, int *max_clusters);

} // namespace extz
