// one-slot CTA-wide kernels (64 / 128 / 256 lanes) and the one-slot cluster kernels (2 / 4 CTAs x 256 lanes)
#include "k_dp1.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP1(64, 16, true)
EXTZ_INSTANTIATE_DP1(128, 16, true)
EXTZ_INSTANTIATE_DP1(256, 16, true)
template cudaError_t dp1_cluster_dispatch_c<2>(const DpLaunch &, bool, bool, int, cudaStream_t, int *);
template cudaError_t dp1_cluster_dispatch_c<4>(const DpLaunch &, bool, bool, int, cudaStream_t, int *);
}
