// packed narrow kernel, KSW_EZ_APPROX_MAX variant, G = 1 and 2 lanes per pair (see k_dp16_narrow.cuh)
#include "k_dp16_narrow.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP16(1, true)
EXTZ_INSTANTIATE_DP16(2, true)
}
