// fp16-split tcgen05 assign kernel instantiations for d = 16 (see pq_assign_h_kernel.cuh)
#include "pq_assign_h_kernel.cuh"
namespace equss {
namespace tch {
#ifndef EQUSS_D16_STF
#define EQUSS_D16_STF 8
#define EQUSS_D16_LAG 4
#endif
EQUSS_TCH_DISPATCH(16, 2, 6, 4, EQUSS_D16_STF, EQUSS_D16_LAG)
}  // namespace tch
}  // namespace equss
