// placeholder until the tcgen05 kernel lands
#include "equss_common.cuh"
#include "pq_assign.h"
namespace equss {
bool assign_tc_supported(const equss_zdesc*, int, int, int, int, bool) { return false; }
int64_t assign_tc_workspace_bytes(int64_t, int, int, int) { return 0; }
int assign_tc_launch(const float*, const equss_zdesc*, const float*, const float*, int, int, int, int,
                     const float*, const float*, int32_t*, void*, int64_t, cudaStream_t) {
  set_error("tcgen05 assign kernel not built");
  return EQUSS_ERR_UNSUPPORTED;
}
}
