// k_dp16_narrow.cuh -- launchers of the packed narrow kernel extz_dp16_kernel<G> (definitions; instantiated per G in k_dp16_n*.cu)
#pragma once
#include "kernels_impl.h"
#include "extz_dp16.cuh"

namespace extz {

template <int G, bool A>
cudaError_t dp16_launch_g(const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st)
{
	if (cigar) {
		if (right) extz_dp16_kernel<G, true, true, A><<<grid, 128, 0, st>>>(L);
		else       extz_dp16_kernel<G, true, false, A><<<grid, 128, 0, st>>>(L);
	} else       extz_dp16_kernel<G, false, false, A><<<grid, 128, 0, st>>>(L);
	return cudaGetLastError();
}
template <int G, bool A>
int dp16_occupancy_g(bool cigar, bool right)
{
	int nb = 0;
	if (cigar) {
		if (right) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_kernel<G, true, true, A>, 128, 0);
		else       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_kernel<G, true, false, A>, 128, 0);
	} else       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extz_dp16_kernel<G, false, false, A>, 128, 0);
	return nb;
}
#define EXTZ_INSTANTIATE_DP16(G, A) \
	template cudaError_t dp16_launch_g<G, A>(const DpLaunch &, bool, bool, int, cudaStream_t); \
	template int dp16_occupancy_g<G, A>(bool, bool);

} // namespace extz
