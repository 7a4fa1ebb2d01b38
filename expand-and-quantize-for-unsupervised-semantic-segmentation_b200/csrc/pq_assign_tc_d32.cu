// tcgen05 assign kernel instantiations for d = 32 (see pq_assign_tc_kernel.cuh)
#include "pq_assign_tc_kernel.cuh"
namespace equss {
namespace tc {
EQUSS_TC_DISPATCH(32, 256, 4, 2)
}  // namespace tc
}  // namespace equss
