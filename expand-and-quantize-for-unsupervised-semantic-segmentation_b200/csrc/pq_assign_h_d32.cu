// fp16-split tcgen05 assign kernel instantiations for d = 32 (see pq_assign_h_kernel.cuh)
#include "pq_assign_h_kernel.cuh"
namespace equss {
namespace tch {
#ifndef EQUSS_D32_STF
#define EQUSS_D32_STF 6
#define EQUSS_D32_LAG 4
#endif
EQUSS_TCH_DISPATCH(32, 1, 4, 3, EQUSS_D32_STF, EQUSS_D32_LAG)
}  // namespace tch
}  // namespace equss
