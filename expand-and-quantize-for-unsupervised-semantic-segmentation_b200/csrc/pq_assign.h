// pq_assign.h -- internal interface between the assign dispatcher (pq_assign_simt.cu) and the
// tcgen05 kernel (pq_assign_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/equss_b200.h"

namespace equss {
bool assign_tc_supported(const equss_zdesc* zd, int M, int K, int d, int norm_mode, bool want_margin);
int64_t assign_tc_workspace_bytes(int64_t n_pixels, int M, int K, int d);
int assign_tc_launch(const float* z, const equss_zdesc* zd, const float* codebook_norm, const float* cnorm2,
                     int M, int K, int d, int norm_mode, const float* norm_a, const float* norm_b,
                     int32_t* idx_out, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
// fp16-split kernel (pq_assign_h.cu): l2 rows, d in {16, 32, 64}
bool assign_tch_supported(const equss_zdesc* zd, int M, int K, int d, int norm_mode, bool want_margin);
int64_t assign_tch_workspace_bytes(int64_t n_pixels, int M, int K, int d);
bool assign_tch_fusable(const equss_zdesc* zd, int M, int K, int d, int norm_mode);
// gather_src / out / sqerr non-null: the kernel also runs K3 (gather + straight-through value + squared error)
int assign_tch_launch(const float* z, const equss_zdesc* zd, const float* codebook_norm, const float* cnorm2,
                      int M, int K, int d, int32_t* idx_out, void* workspace, int64_t workspace_bytes,
                      cudaStream_t stream, const float* gather_src = nullptr, float* out = nullptr,
                      double* sqerr = nullptr);
}  // namespace equss
