// packed narrow kernel, G = 16 and 32 lanes per pair (see k_dp16_narrow.cuh)
#include "k_dp16_narrow.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP16(16, false)
EXTZ_INSTANTIATE_DP16(32, false)
}
