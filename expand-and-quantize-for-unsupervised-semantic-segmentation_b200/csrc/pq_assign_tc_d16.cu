// tcgen05 assign kernel instantiations for d = 16 (see pq_assign_tc_kernel.cuh)
#include "pq_assign_tc_kernel.cuh"
namespace equss {
namespace tc {
EQUSS_TC_DISPATCH(16, 256, 6, 4)
}  // namespace tc
}  // namespace equss
