// one-slot narrow kernels: classes (16,16) (32,16) (32,32)
#include "k_dp1.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP1(16, 16, false)
EXTZ_INSTANTIATE_DP1(32, 16, false)
EXTZ_INSTANTIATE_DP1(32, 32, false)
}
