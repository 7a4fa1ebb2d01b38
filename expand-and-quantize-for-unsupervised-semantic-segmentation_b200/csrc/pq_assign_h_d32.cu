// fp16-split tcgen05 assign kernel instantiations for d = 32 (see pq_assign_h_kernel.cuh)
#include "pq_assign_h_kernel.cuh"
namespace equss {
namespace tch {
EQUSS_TCH_DISPATCH(32, 1, 4, 3, 6, 3)
}  // namespace tch
}  // namespace equss
