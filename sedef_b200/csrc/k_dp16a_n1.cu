// packed narrow kernel, KSW_EZ_APPROX_MAX variant, G = 4 and 8 lanes per pair (see k_dp16_narrow.cuh)
#include "k_dp16_narrow.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP16(4, true)
EXTZ_INSTANTIATE_DP16(8, true)
}
