// run-time dispatch from (class parameters) to the explicitly instantiated launchers of the k_*.cu translation units
#include "kernels_impl.h"

namespace extz {

#define EXTZ_G_SWITCH(G, CALL) \
	switch (G) { case 1: return CALL(1); case 2: return CALL(2); case 4: return CALL(4); case 8: return CALL(8); case 16: return CALL(16); case 32: return CALL(32); }
#define EXTZ_GW_SWITCH(G, CALL) \
	switch (G) { case 64: return CALL(64); case 128: return CALL(128); case 256: return CALL(256); }
cudaError_t k_dp16_launch(int G, const DpLaunch &L, bool cigar, bool right, bool approx, int grid, cudaStream_t st)
{
#define EXTZ_CALL(g) (approx ? dp16_launch_g<g, true>(L, cigar, right, grid, st) : dp16_launch_g<g, false>(L, cigar, right, grid, st))
	EXTZ_G_SWITCH(G, EXTZ_CALL)
#undef EXTZ_CALL
	return cudaErrorInvalidValue;
}
int k_dp16_occupancy(int G, bool cigar, bool right, bool approx)
{
#define EXTZ_CALL(g) (approx ? dp16_occupancy_g<g, true>(cigar, right) : dp16_occupancy_g<g, false>(cigar, right))
	EXTZ_G_SWITCH(G, EXTZ_CALL)
#undef EXTZ_CALL
	return 0;
}
cudaError_t k_dp16_wide_launch(int G, const DpLaunch &L, bool cigar, bool right, bool approx, int grid, cudaStream_t st)
{
#define EXTZ_CALL(g) (approx ? dp16_wide_launch_g<g, true>(L, cigar, right, grid, st) : dp16_wide_launch_g<g, false>(L, cigar, right, grid, st))
	EXTZ_GW_SWITCH(G, EXTZ_CALL)
#undef EXTZ_CALL
	return cudaErrorInvalidValue;
}
int k_dp16_wide_occupancy(int G, bool cigar, bool right, bool approx)
{
#define EXTZ_CALL(g) (approx ? dp16_wide_occupancy_g<g, true>(cigar, right) : dp16_wide_occupancy_g<g, false>(cigar, right))
	EXTZ_GW_SWITCH(G, EXTZ_CALL)
#undef EXTZ_CALL
	return 0;
}
cudaError_t k_dp16_cluster_dispatch(int C, const DpLaunch &L, bool cigar, bool right, bool approx, int nclusters, cudaStream_t st, int *max_clusters)
{
	return approx ? dp16_cluster_dispatch_a<true>(C, L, cigar, right, nclusters, st, max_clusters)
	              : dp16_cluster_dispatch_a<false>(C, L, cigar, right, nclusters, st, max_clusters);
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
