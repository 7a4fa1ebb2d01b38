// stego_corr.cu -- feature-correlation term of the STEGO correspondence loss, the step next to the PQ head
// in the training loop.  Reference: model/loss.py:647-700 (STEGOLoss.helper)
//     fd = einsum("nchw,ncij->nhwij", normalize(f1, dim=1, eps=1e-10), normalize(f2, dim=1, eps=1e-10))     (no grad)
//     pointwise:  old = fd.mean();  fd -= fd.mean([3, 4], keepdim=True);  fd = fd - fd.mean() + old
// on the S x S = 11 x 11 sampled feature maps: per image a [S^2 x C] x [C x S^2] contraction with C = 384 / 768 backbone
// channels, six eager kernels plus three full-tensor reductions in the reference.  Here: ONE kernel forms the cosine
// similarities (norms accumulated in the same channel loop), subtracts each row's mean and accumulates the two global
// sums the final correction needs; the caller adds the scalar (old_mean - mean_after) (one element-wise op).
#include "equss_common.cuh"

namespace equss {

constexpr int kCorrRows = 11;     // rows (p) of the S^2 x S^2 matrix per block

// f1, f2: [n][C][P] (P = S*S positions, contiguous); fd: [n][P][P]; sums: fp64 [2] += (sum fd before centering, sum after)
__global__ void __launch_bounds__(128)
stego_feature_corr_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int C, int P, int pointwise,
                          float* __restrict__ fd, double* __restrict__ sums) {
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * kCorrRows;
  const int q = threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* a = f1 + (long long)n * C * P;
  const float* b = f2 + (long long)n * C * P;
  __shared__ float s_a[kCorrRows];
  __shared__ float s_red[4][kCorrRows];
  __shared__ double s_d[4][2];
  float acc[kCorrRows], na[kCorrRows];
#pragma unroll
  for (int i = 0; i < kCorrRows; ++i) { acc[i] = 0.f; na[i] = 0.f; }
  float nb = 0.f;
  for (int c = 0; c < C; ++c) {
    __syncthreads();
    if (threadIdx.x < kCorrRows) s_a[threadIdx.x] = (p0 + (int)threadIdx.x < P) ? __ldg(a + (long long)c * P + p0 + threadIdx.x) : 0.f;
    __syncthreads();
    const float bv = (q < P) ? __ldg(b + (long long)c * P + q) : 0.f;
    nb = fmaf(bv, bv, nb);
#pragma unroll
    for (int i = 0; i < kCorrRows; ++i) {
      const float av = s_a[i];
      acc[i] = fmaf(av, bv, acc[i]);
      na[i] = fmaf(av, av, na[i]);
    }
  }
  const float inb = 1.f / fmaxf(sqrtf(nb), 1e-10f);
  float v[kCorrRows];
#pragma unroll
  for (int i = 0; i < kCorrRows; ++i) v[i] = (q < P) ? acc[i] * (1.f / fmaxf(sqrtf(na[i]), 1e-10f)) * inb : 0.f;
  // row means over q
  float rs[kCorrRows];
#pragma unroll
  for (int i = 0; i < kCorrRows; ++i) rs[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kCorrRows; ++i) s_red[warp][i] = rs[i];
  }
  __syncthreads();
  double before = 0.0, after = 0.0;
#pragma unroll
  for (int i = 0; i < kCorrRows; ++i) {
    const float mean = ((s_red[0][i] + s_red[1][i]) + (s_red[2][i] + s_red[3][i])) / (float)P;
    const bool live = q < P && p0 + i < P;
    const float out = pointwise ? v[i] - mean : v[i];
    if (live) {
      fd[((long long)n * P + p0 + i) * P + q] = out;
      before += (double)v[i];
      after += (double)out;
    }
  }
  before = warp_sum_d(before); after = warp_sum_d(after);
  if (lane == 0) { s_d[warp][0] = before; s_d[warp][1] = after; }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(sums, (s_d[0][0] + s_d[1][0]) + (s_d[2][0] + s_d[3][0]));
    atomicAdd(sums + 1, (s_d[0][1] + s_d[1][1]) + (s_d[2][1] + s_d[3][1]));
  }
}

}  // namespace equss

using namespace equss;

extern "C" int equss_stego_feature_corr(const float* f1, const float* f2, int n, int C, int P, int pointwise, float* fd,
                                        double* sums, void* stream) {
  EQUSS_REQUIRE(f1 && f2 && fd && sums, EQUSS_ERR_INVALID_ARG, "equss_stego_feature_corr: null pointer");
  EQUSS_REQUIRE(n > 0 && n <= 65535 && C > 0 && P > 0, EQUSS_ERR_INVALID_ARG, "equss_stego_feature_corr: bad shape n=%d C=%d P=%d", n, C, P);
  EQUSS_REQUIRE(P <= 128, EQUSS_ERR_UNSUPPORTED, "equss_stego_feature_corr: %d sampled positions > 128 (feature_samples <= 11)", P);
  dim3 grid((unsigned)((P + kCorrRows - 1) / kCorrRows), (unsigned)n);
  stego_feature_corr_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(f1, f2, C, P, pointwise, fd, sums);
  EQUSS_LAUNCH_OK("stego_feature_corr_kernel");
  return EQUSS_OK;
}
