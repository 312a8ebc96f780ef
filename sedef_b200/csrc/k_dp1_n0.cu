// one-slot narrow kernels: classes (2,16) (4,16) (8,16)
#include "k_dp1.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP1(2, 16, false)
EXTZ_INSTANTIATE_DP1(4, 16, false)
EXTZ_INSTANTIATE_DP1(8, 16, false)
}
