// packed CTA-wide + cluster kernels, exact-max variant (see k_dp16_wide.cuh)
#include "k_dp16_wide.cuh"
namespace extz {
EXTZ_INSTANTIATE_DP16_WIDE(false)
}
