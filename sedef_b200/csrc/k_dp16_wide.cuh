// k_dp16_wide.cuh -- launchers of the packed CTA-wide kernels (64 / 128 / 256 lanes per pair) and of the packed cluster kernel
// (2 CTAs x 256 lanes); definitions, instantiated for the exact-max and the approx-max variant in k_dp16_wide.cu / k_dp16a_wide.cu
#pragma once
#include "kernels_impl.h"
#include "extz_dp16.cuh"

namespace extz {

// 192 B of dynamic shared memory per lane (H and u' rows).  The opt-in attribute is per device: set on every call.
template <int G, bool C, bool R, bool A>
static cudaError_t dp16_wide_prepare()
{
	return cudaFuncSetAttribute(extz_dp16_wide_kernel<G, C, R, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, G * 192);
}
template <int G, bool A>
cudaError_t dp16_wide_launch_g(const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st)
{
	const size_t dyn = (size_t)G * 192;
	cudaError_t e;
	if (cigar) {
		if (right) { if ((e = dp16_wide_prepare<G, true, true, A>()) != cudaSuccess) return e; extz_dp16_wide_kernel<G, true, true, A><<<grid, G, dyn, st>>>(L); }
		else       { if ((e = dp16_wide_prepare<G, true, false, A>()) != cudaSuccess) return e; extz_dp16_wide_kernel<G, true, false, A><<<grid, G, dyn, st>>>(L); }
	} else         { if ((e = dp16_wide_prepare<G, false, false, A>()) != cudaSuccess) return e; extz_dp16_wide_kernel<G, false, false, A><<<grid, G, dyn, st>>>(L); }
	return cudaGetLastError();
}
template <int G, bool A>
int dp16_wide_occupancy_g(bool cigar, bool right)
{
	int nb = 0;
	const size_t dyn = (size_t)G * 192;
	if (cigar) {
		if (right) { dp16_wide_prepare<G, true, true, A>(); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_wide_kernel<G, true, true, A>, G, dyn); }
		else       { dp16_wide_prepare<G, true, false, A>(); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_wide_kernel<G, true, false, A>, G, dyn); }
	} else         { dp16_wide_prepare<G, false, false, A>(); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_wide_kernel<G, false, false, A>, G, dyn); }
	return nb;
}
// packed cluster kernels (C = 2 / 4 / 8 CTAs x 256 lanes x 32 slots = 16384 / 32768 / 65536 live slots): 48 KB of dynamic shared memory per CTA
template <int C, bool CG, bool R, bool A>
static cudaError_t cluster16_launch_one(const DpLaunch &L, int nclusters, cudaStream_t st, int *max_clusters)
{
	const size_t dyn = 256 * 192;
	cudaError_t e = cudaFuncSetAttribute(extz_dp16_cluster_kernel<C, CG, R, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
	if (e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)(nclusters * C)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = dyn; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	if (max_clusters) {
		cfg.gridDim = dim3((unsigned)C);
		return cudaOccupancyMaxActiveClusters(max_clusters, extz_dp16_cluster_kernel<C, CG, R, A>, &cfg);
	}
	return cudaLaunchKernelEx(&cfg, extz_dp16_cluster_kernel<C, CG, R, A>, L);
}
template <int C, bool A>
static cudaError_t dp16_cluster_dispatch_ca(const DpLaunch &L, bool cigar, bool right, int nclusters, cudaStream_t st, int *max_clusters)
{
	if (cigar) return right ? cluster16_launch_one<C, true, true, A>(L, nclusters, st, max_clusters)
	                        : cluster16_launch_one<C, true, false, A>(L, nclusters, st, max_clusters);
	return cluster16_launch_one<C, false, false, A>(L, nclusters, st, max_clusters);
}
template <bool A>
cudaError_t dp16_cluster_dispatch_a(int C, const DpLaunch &L, bool cigar, bool right, int nclusters, cudaStream_t st, int *max_clusters)
{
	if (C == 2) return dp16_cluster_dispatch_ca<2, A>(L, cigar, right, nclusters, st, max_clusters);
	if (C == 4) return dp16_cluster_dispatch_ca<4, A>(L, cigar, right, nclusters, st, max_clusters);
	if (C == 8) return dp16_cluster_dispatch_ca<8, A>(L, cigar, right, nclusters, st, max_clusters);
	return cudaErrorInvalidValue;
}
#define EXTZ_INSTANTIATE_DP16_WIDE(A) \
	template cudaError_t dp16_wide_launch_g<64, A>(const DpLaunch &, bool, bool, int, cudaStream_t); \
	template cudaError_t dp16_wide_launch_g<128, A>(const DpLaunch &, bool, bool, int, cudaStream_t); \
	template cudaError_t dp16_wide_launch_g<256, A>(const DpLaunch &, bool, bool, int, cudaStream_t); \
	template int dp16_wide_occupancy_g<64, A>(bool, bool); \
	template int dp16_wide_occupancy_g<128, A>(bool, bool); \
	template int dp16_wide_occupancy_g<256, A>(bool, bool); \
	template cudaError_t dp16_cluster_dispatch_a<A>(int, const DpLaunch &, bool, bool, int, cudaStream_t, int *);

} // namespace extz
