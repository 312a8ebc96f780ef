// packed narrow kernel, KSW_EZ_APPROX_MAX variant, G = 16 and 32 lanes per pair (see k_dp16_narrow.cuh)
#include "k_dp16_narrow.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP16(16, true)
EXTZ_INSTANTIATE_DP16(32, true)
}
