// packed narrow kernel, G = 4 and 8 lanes per pair (see k_dp16_narrow.cuh)
#include "k_dp16_narrow.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP16(4, false)
EXTZ_INSTANTIATE_DP16(8, false)
}
