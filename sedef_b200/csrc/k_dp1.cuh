// k_dp1.cuh -- launchers of the one-slot-per-register kernels of extz_dp.cuh (the KSW_B200_PACKED=0 A/B path)
#pragma once
#include "kernels_impl.h"
#include "extz_dp.cuh"

namespace extz {

template <int G, int S, bool W, bool C, bool R> struct KSel;
template <int G, int S, bool C, bool R> struct KSel<G, S, false, C, R> { static constexpr auto fn = extz_dp_kernel<G, S, C, R>; };
template <int G, int S, bool C, bool R> struct KSel<G, S, true, C, R> { static constexpr auto fn = extz_dp_wide_kernel<G, S, C, R>; };

template <int G, int S, bool W>
cudaError_t dp1_launch_gs(const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st)
{
	constexpr int threads = W ? G : 128;
	if (cigar) {
		if (right) KSel<G, S, W, true, true>::fn<<<grid, threads, 0, st>>>(L);
		else       KSel<G, S, W, true, false>::fn<<<grid, threads, 0, st>>>(L);
	} else       KSel<G, S, W, false, false>::fn<<<grid, threads, 0, st>>>(L);
	return cudaGetLastError();
}
template <int G, int S, bool W>
int dp1_occupancy_gs(bool cigar, bool right)
{
	constexpr int threads = W ? G : 128;
	int nb = 0;
	if (cigar) {
		if (right) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, KSel<G, S, W, true, true>::fn, threads, 0);
		else       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, KSel<G, S, W, true, false>::fn, threads, 0);
	} else       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, KSel<G, S, W, false, false>::fn, threads, 0);
	return nb;
}
#define EXTZ_INSTANTIATE_DP1(G, S, W) \
	template cudaError_t dp1_launch_gs<G, S, W>(const DpLaunch &, bool, bool, int, cudaStream_t); \
	template int dp1_occupancy_gs<G, S, W>(bool, bool);

// cluster kernels: launched with a cluster dimension attribute; "occupancy" = co-resident clusters on the device
template <int C, bool CG, bool R>
static cudaError_t cluster_launch_one(const DpLaunch &L, int nclusters, cudaStream_t st, int *max_clusters)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)(nclusters * C)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	if (max_clusters) {
		cfg.gridDim = dim3((unsigned)C);
		return cudaOccupancyMaxActiveClusters(max_clusters, extz_dp_cluster_kernel<C, 16, CG, R>, &cfg);
	}
	return cudaLaunchKernelEx(&cfg, extz_dp_cluster_kernel<C, 16, CG, R>, L);
}
template <int C>
cudaError_t dp1_cluster_dispatch_c(const DpLaunch &L, bool cigar, bool right, int nclusters, cudaStream_t st, int *max_clusters)
{
	if (cigar) return right ? cluster_launch_one<C, true, true>(L, nclusters, st, max_clusters)
	                        : cluster_launch_one<C, true, false>(L, nclusters, st, max_clusters);
	return cluster_launch_one<C, false, false>(L, nclusters, st, max_clusters);
}

} // namespace extz
