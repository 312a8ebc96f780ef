// run-time dispatch from (class parameters) to the explicitly instantiated launchers of the k_*.cu translation units
#include "kernels_impl.h"

namespace extz {

cudaError_t k_dp16_launch(int G, const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st)
{
	switch (G) {
	case 1: return dp16_launch_g<1>(L, cigar, right, grid, st);
	case 2: return dp16_launch_g<2>(L, cigar, right, grid, st);
	case 4: return dp16_launch_g<4>(L, cigar, right, grid, st);
	case 8: return dp16_launch_g<8>(L, cigar, right, grid, st);
	case 16: return dp16_launch_g<16>(L, cigar, right, grid, st);
	case 32: return dp16_launch_g<32>(L, cigar, right, grid, st);
	}
	return cudaErrorInvalidValue;
}
int k_dp16_occupancy(int G, bool cigar, bool right)
{
	switch (G) {
	case 1: return dp16_occupancy_g<1>(cigar, right);
	case 2: return dp16_occupancy_g<2>(cigar, right);
	case 4: return dp16_occupancy_g<4>(cigar, right);
	case 8: return dp16_occupancy_g<8>(cigar, right);
	case 16: return dp16_occupancy_g<16>(cigar, right);
	case 32: return dp16_occupancy_g<32>(cigar, right);
	}
	return 0;
}
cudaError_t k_dp16_wide_launch(int G, const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st)
{
	switch (G) {
	case 64: return dp16_wide_launch_g<64>(L, cigar, right, grid, st);
	case 128: return dp16_wide_launch_g<128>(L, cigar, right, grid, st);
	case 256: return dp16_wide_launch_g<256>(L, cigar, right, grid, st);
	}
	return cudaErrorInvalidValue;
}
int k_dp16_wide_occupancy(int G, bool cigar, bool right)
{
	switch (G) {
	case 64: return dp16_wide_occupancy_g<64>(cigar, right);
	case 128: return dp16_wide_occupancy_g<128>(cigar, right);
	case 256: return dp16_wide_occupancy_g<256>(cigar, right);
	}
	return 0;
}
#define EXTZ_FOR_CLASS(c, CALL) \
	switch (c) { \
	case 0: return CALL(2, 16, false); case 1: return CALL(4, 16, false); case 2: return CALL(8, 16, false); \
	case 3: return CALL(16, 16, false); case 4: return CALL(32, 16, false); case 5: return CALL(32, 32, false); \
	case 6: return CALL(64, 16, true); case 7: return CALL(128, 16, true); case 8: return CALL(256, 16, true); }
cudaError_t k_dp_launch(int c, const DpLaunch &L, bool cigar, bool right, int grid, cudaStream_t st)
{
#define EXTZ_CALL(G, S, W) dp1_launch_gs<G, S, W>(L, cigar, right, grid, st)
	EXTZ_FOR_CLASS(c, EXTZ_CALL)
#undef EXTZ_CALL
	return cudaErrorInvalidValue;
}
int k_dp_occupancy(int c, bool cigar, bool right)
{
#define EXTZ_CALL(G, S, W) dp1_occupancy_gs<G, S, W>(cigar, right)
	EXTZ_FOR_CLASS(c, EXTZ_CALL)
#undef EXTZ_CALL
	return 0;
}
cudaError_t k_dp_cluster_dispatch(int C, const DpLaunch &L, bool cigar, bool right, int nclusters, cudaStream_t st, int *max_clusters)
{
	if (C == 2) return dp1_cluster_dispatch_c<2>(L, cigar, right, nclusters, st, max_clusters);
	if (C == 4) return dp1_cluster_dispatch_c<4>(L, cigar, right, nclusters, st, max_clusters);
	return cudaErrorInvalidValue;
}

} // namespace extz
