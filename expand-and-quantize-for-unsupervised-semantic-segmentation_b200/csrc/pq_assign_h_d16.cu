// fp16-split tcgen05 assign kernel instantiations for d = 16 (see pq_assign_h_kernel.cuh)
#include "pq_assign_h_kernel.cuh"
namespace equss {
namespace tch {
EQUSS_TCH_DISPATCH(16, 2, 6, 4, 8, 4)
}  // namespace tch
}  // namespace equss
