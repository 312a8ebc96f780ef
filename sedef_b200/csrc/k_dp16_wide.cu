// packed CTA-wide kernels (64 / 128 / 256 lanes per pair) and the packed cluster kernel (2 CTAs x 256 lanes)
#include "kernels_impl.h"
#include "extz_dp16.cuh"

namespace extz {

// 192 B of dynamic shared memory per lane (H and u' rows).  The opt-in attribute is per device: set on every call.
template <int G, bool C, bool R>
static cudaError_t dp16_wide_prepare()
{
	return cudaFuncSetAttribute(extz_dp16_wide_kernel<G, C, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, G * 192);
}
template <int G>
cudaError_t dp16_wide_launch_g(const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st)
{
	const size_t dyn = (size_t)G * 192;
	cudaError_t e;
	if (cigar) {
		if (right) { if ((e = dp16_wide_prepare<G, true, true>()) != cudaSuccess) return e; extz_dp16_wide_kernel<G, true, true><<<grid, G, dyn, st>>>(L); }
		else       { if ((e = dp16_wide_prepare<G, true, false>()) != cudaSuccess) return e; extz_dp16_wide_kernel<G, true, false><<<grid, G, dyn, st>>>(L); }
	} else         { if ((e = dp16_wide_prepare<G, false, false>()) != cudaSuccess) return e; extz_dp16_wide_kernel<G, false, false><<<grid, G, dyn, st>>>(L); }
	return cudaGetLastError();
}
template <int G>
int dp16_wide_occupancy_g(bool cigar, bool right)
{
	int nb = 0;
	const size_t dyn = (size_t)G * 192;
	if (cigar) {
		if (right) { dp16_wide_prepare<G, true, true>(); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_wide_kernel<G, true, true>, G, dyn); }
		else       { dp16_wide_prepare<G, true, false>(); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_wide_kernel<G, true, false>, G, dyn); }
	} else         { dp16_wide_prepare<G, false, false>(); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_wide_kernel<G, false, false>, G, dyn); }
	return nb;
}
template cudaError_t dp16_wide_launch_g<64>(const DpLaunch &, bool, bool, int, cudaStream_t);
template cudaError_t dp16_wide_launch_g<128>(const DpLaunch &, bool, bool, int, cudaStream_t);
template cudaError_t dp16_wide_launch_g<256>(const DpLaunch &, bool, bool, int, cudaStream_t);
template int dp16_wide_occupancy_g<64>(bool, bool);
template int dp16_wide_occupancy_g<128>(bool, bool);
template int dp16_wide_occupancy_g<256>(bool, bool);

// packed cluster kernel (2 CTAs x 256 lanes x 32 slots = 16384 live slots): 48 KB of dynamic shared memory per CTA
template <bool CG, bool R>
static cudaError_t cluster16_launch_one(const DpLaunch &L, int nclusters, cudaStream_t st, int *max_clusters)
{
	constexpr int C = 2;
	const size_t dyn = 256 * 192;
	cudaError_t e = cudaFuncSetAttribute(extz_dp16_cluster_kernel<C, CG, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
	if (e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)(nclusters * C)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = dyn; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	if (max_clusters) {
		cfg.gridDim = dim3((unsigned)C);
		return cudaOccupancyMaxActiveClusters(max_clusters, extz_dp16_cluster_kernel<C, CG, R>, &cfg);
	}
	return cudaLaunchKernelEx(&cfg, extz_dp16_cluster_kernel<C, CG, R>, L);
}
cudaError_t k_dp16_cluster_dispatch(const DpLaunch &L, bool cigar, bool right, int nclusters, cudaStream_t st, int *max_clusters)
{
	if (cigar) return right ? cluster16_launch_one<true, true>(L, nclusters, st, max_clusters)
	                        : cluster16_launch_one<true, false>(L, nclusters, st, max_clusters);
	return cluster16_launch_one<false, false>(L, nclusters, st, max_clusters);
}

} // namespace extz
