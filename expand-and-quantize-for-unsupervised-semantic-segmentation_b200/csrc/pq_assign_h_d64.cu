// fp16-split tcgen05 assign kernel instantiations for d = 64 (see pq_assign_h_kernel.cuh)
#include "pq_assign_h_kernel.cuh"
namespace equss {
namespace tch {
EQUSS_TCH_DISPATCH(64, 1, 2, 2, 0, 1)
}  // namespace tch
}  // namespace equss
