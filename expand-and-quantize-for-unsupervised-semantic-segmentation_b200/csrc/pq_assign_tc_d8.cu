// tcgen05 assign kernel instantiations for d = 8 (see pq_assign_tc_kernel.cuh)
#include "pq_assign_tc_kernel.cuh"
namespace equss {
namespace tc {
EQUSS_TC_DISPATCH(8, 256, 8, 4)
}  // namespace tc
}  // namespace equss
