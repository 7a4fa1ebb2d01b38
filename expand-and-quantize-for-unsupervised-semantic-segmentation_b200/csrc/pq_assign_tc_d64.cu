// tcgen05 assign kernel instantiations for d = 64 (see pq_assign_tc_kernel.cuh)
#include "pq_assign_tc_kernel.cuh"
namespace equss {
namespace tc {
EQUSS_TC_DISPATCH(64, 128, 2, 1)
}  // namespace tc
}  // namespace equss
