// pq_assign_simt.cu -- exact fp32 CUDA-core distance + argmin (K1) and the soft assignment (K2).
//
// This is the always-available, shape-agnostic assign kernel: it validates the tcgen05 kernel
// (pq_assign_tc.cu) bit for bit -- both use the same canonical z_norm, the same sequential fma dot
// product and the reference's association order (sum z^2 + sum c^2) - 2*dot, first minimal index wins
// (model/quantizer.py:457-467) -- and it serves the shapes the tensor-core kernel does not cover.
#include <cstdlib>
#include "equss_common.cuh"
#include "pq_assign.h"

namespace equss {

// pq_prob.cu: register-tiled soft assignment (K <= 256, d in {8,16,32,64})
bool distance_prob_tiled_supported(int K, int d, const float* prob);
int distance_prob_tiled_launch(const float* z, const equss_zdesc* zd, const float* cb, const float* cn2, int M, int K,
                               int d, int mode, const float* na, const float* nb, float temperature, float* prob,
                               cudaStream_t st);

template <int DT>
struct AssignRow {
  float x[DT > 0 ? DT : kMaxD];
};

// grid = (pixel chunks, M), block = 128 threads, one pixel per thread.
// Codebook of subspace m is staged in shared memory in chunks of KC codes.
template <int DT>
__global__ void __launch_bounds__(128)
assign_simt_kernel(const float* __restrict__ z, ZView zv, const float* __restrict__ cb,
                   const float* __restrict__ cn2, int K, int d, int KC, int mode,
                   const float* __restrict__ na, const float* __restrict__ nb,
                   int32_t* __restrict__ idx_out, float* __restrict__ margin_out) {
  extern __shared__ __align__(16) float s_cb[];  // [KC][d] then [KC] cn2
  float* s_cn2 = s_cb + (size_t)KC * d;
  const int m = blockIdx.y;
  const int dd = DT > 0 ? DT : d;
  const float* cbm = cb + (long long)m * K * d;
  const float* cn2m = cn2 + (long long)m * K;
  const long long n_iter = (zv.n_pixels + blockDim.x - 1) / blockDim.x;
  for (long long it = blockIdx.x; it < n_iter; it += gridDim.x) {
    const long long n = it * blockDim.x + threadIdx.x;
    const bool live = n < zv.n_pixels;
    AssignRow<DT> row;
    float zn2 = 0.f;
    if (live) {
      long long base = pixel_base(zv, n) + (long long)m * d * zv.stride_c;
      for (int j = 0; j < dd; ++j) row.x[j] = __ldg(z + base + j * zv.stride_c);
      RowNorm r = row_norm_generic(mode, dd, [&](int j) { return row.x[j]; });
      for (int j = 0; j < dd; ++j) {
        float v = row.x[j];
        if (mode == EQUSS_NORM_AFFINE) v = (v - __ldg(na + m * d + j)) / __ldg(nb + m * d + j);
        else v = apply_norm(v, r, mode);
        row.x[j] = v;
      }
      zn2 = canonical_sumsq(dd, [&](int j) { return row.x[j]; });
    }
    float best = INFINITY, second = INFINITY;
    int bi = 0;
    for (int k0 = 0; k0 < K; k0 += KC) {
      const int kc = min(KC, K - k0);
      __syncthreads();
      for (int i = threadIdx.x; i < kc * d; i += blockDim.x) s_cb[i] = __ldg(cbm + (long long)k0 * d + i);
      for (int i = threadIdx.x; i < kc; i += blockDim.x) s_cn2[i] = __ldg(cn2m + k0 + i);
      __syncthreads();
      if (live) {
        for (int k = 0; k < kc; ++k) {
          const float* c = s_cb + k * d;
          float dot = 0.f;
          if (DT > 0 && (DT % 4) == 0) {
#pragma unroll
            for (int j = 0; j < dd; j += 4) {
              float4 c4 = *reinterpret_cast<const float4*>(c + j);
              dot = fmaf(row.x[j], c4.x, dot);
              dot = fmaf(row.x[j + 1], c4.y, dot);
              dot = fmaf(row.x[j + 2], c4.z, dot);
              dot = fmaf(row.x[j + 3], c4.w, dot);
            }
          } else {
            for (int j = 0; j < dd; ++j) dot = fmaf(row.x[j], c[j], dot);
          }
          float dist = ref_distance(zn2, s_cn2[k], dot);
          if (dist < best) { second = best; best = dist; bi = k0 + k; }
          else if (dist < second) { second = dist; }
        }
      }
    }
    if (live) {
      idx_out[(long long)m * zv.n_pixels + n] = bi;
      if (margin_out) {
        float denom = fmaxf(fabsf(best), 1e-30f);
        margin_out[(long long)m * zv.n_pixels + n] = (second - best) / denom;
      }
    }
  }
}

// K2: one warp per (pixel, subspace) row; lanes stride over the K codes so the N x (M*K) output is
// written with fully coalesced 128-byte rows.  KPL = ceil(K/32) distances stay in registers.
template <int KPL>
__global__ void __launch_bounds__(256)
distance_prob_kernel(const float* __restrict__ z, ZView zv, const float* __restrict__ cb,
                     const float* __restrict__ cn2, int M, int K, int d, int mode,
                     const float* __restrict__ na, const float* __restrict__ nb, float temperature,
                     float* __restrict__ prob) {
  extern __shared__ __align__(16) float s_mem[];  // [K][d+1] codebook (padded), [K] cn2, [warps][d] rows
  const int ldc = d + 1;
  float* s_cb = s_mem;
  float* s_cn2 = s_cb + (size_t)K * ldc;
  float* s_rows = s_cn2 + K;
  const int m = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < K * d; i += blockDim.x) {
    int k = i / d, j = i - k * d;
    s_cb[k * ldc + j] = __ldg(cb + (long long)m * K * d + i);
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) s_cn2[i] = __ldg(cn2 + (long long)m * K + i);
  __syncthreads();
  float* myrow = s_rows + warp * d;
  for (long long n = (long long)blockIdx.x * nwarps + warp; n < zv.n_pixels; n += (long long)gridDim.x * nwarps) {
    long long base = pixel_base(zv, n) + (long long)m * d * zv.stride_c;
    __syncwarp();
    for (int j = lane; j < d; j += 32) myrow[j] = __ldg(z + base + j * zv.stride_c);
    __syncwarp();
    RowNorm r = row_norm_generic(mode, d, [&](int j) { return myrow[j]; });
    __syncwarp();
    for (int j = lane; j < d; j += 32) {
      float v = myrow[j];
      if (mode == EQUSS_NORM_AFFINE) v = (v - __ldg(na + m * d + j)) / __ldg(nb + m * d + j);
      else v = apply_norm(v, r, mode);
      myrow[j] = v;
    }
    __syncwarp();
    float zn2 = canonical_sumsq(d, [&](int j) { return myrow[j]; });
    float x[KPL];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      int k = lane + 32 * i;
      float v = -INFINITY;
      if (k < K) {
        const float* c = s_cb + k * ldc;
        float dot = 0.f;
        for (int j = 0; j < d; ++j) dot = fmaf(myrow[j], c[j], dot);
        v = -ref_distance(zn2, s_cn2[k], dot) / temperature;   // softmax(-d / ts) (dino_pqgo.py:655)
      }
      x[i] = v;
      mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      float e = (lane + 32 * i < K) ? expf(x[i] - mx) : 0.f;
      x[i] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    float* out = prob + n * (long long)M * K + (long long)m * K;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      int k = lane + 32 * i;
      if (k < K) __stcs(out + k, x[i] / sum);
    }
  }
}

static int launch_assign_simt(const float* z, const equss_zdesc* zd, const float* cb, const float* cn2,
                              int M, int K, int d, int mode, const float* na, const float* nb,
                              int32_t* idx_out, float* margin_out, cudaStream_t st) {
  ZView zv = make_view(zd);
  // codes per shared-memory chunk: as many as fit in 96 KB
  int KC = K;
  const size_t cap = 96 * 1024;
  while ((size_t)KC * (d + 1) * sizeof(float) > cap) KC = (KC + 1) / 2;
  size_t smem = (size_t)KC * (d + 1) * sizeof(float);
  long long bx = (zd->n_pixels + 127) / 128;
  long long capx = (long long)num_sms() * 16 / M + 1;
  if (bx > capx) bx = capx;
  dim3 grid((unsigned)bx, (unsigned)M);
#define EQUSS_ASSIGN_CASE(DTV)                                                                        \
  {                                                                                                   \
    if (smem > 48 * 1024)                                                                             \
      EQUSS_CUDA_OK(cudaFuncSetAttribute(assign_simt_kernel<DTV>,                                     \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    assign_simt_kernel<DTV><<<grid, 128, smem, st>>>(z, zv, cb, cn2, K, d, KC, mode, na, nb, idx_out, \
                                                     margin_out);                                     \
  }
  switch (d) {
    case 8: EQUSS_ASSIGN_CASE(8); break;
    case 16: EQUSS_ASSIGN_CASE(16); break;
    case 32: EQUSS_ASSIGN_CASE(32); break;
    case 64: EQUSS_ASSIGN_CASE(64); break;
    default: EQUSS_ASSIGN_CASE(0); break;
  }
#undef EQUSS_ASSIGN_CASE
  EQUSS_LAUNCH_OK("assign_simt_kernel");
  return EQUSS_OK;
}

}  // namespace equss

using namespace equss;

extern "C" int64_t equss_pq_assign_workspace_bytes(int64_t n_pixels, int M, int K, int d, int algo) {
  if (algo == EQUSS_ASSIGN_SIMT) return 0;
  const int64_t a = assign_tc_workspace_bytes(n_pixels, M, K, d), b = assign_tch_workspace_bytes(n_pixels, M, K, d);
  return a > b ? a : b;
}

extern "C" int equss_pq_assign(const float* z, const equss_zdesc* zd, const float* codebook_norm,
                               const float* cnorm2, int M, int K, int d, int norm_mode,
                               const float* norm_a, const float* norm_b, int32_t* idx_out,
                               float* margin_out, void* workspace, int64_t workspace_bytes, int algo,
                               void* stream) {
  if (zd && zd->n_pixels == 0) return EQUSS_OK;   // empty tensors have null data pointers
  EQUSS_REQUIRE(z && zd && codebook_norm && cnorm2 && idx_out, EQUSS_ERR_INVALID_ARG, "equss_pq_assign: null pointer");
  int rc = validate_zdesc(zd, M, d); if (rc) return rc;
  EQUSS_REQUIRE(K > 0, EQUSS_ERR_INVALID_ARG, "equss_pq_assign: K=%d", K);
  EQUSS_REQUIRE(norm_mode >= EQUSS_NORM_NONE && norm_mode <= EQUSS_NORM_AFFINE, EQUSS_ERR_INVALID_ARG,
                "Unsupported normalize type %d", norm_mode);
  EQUSS_REQUIRE(norm_mode != EQUSS_NORM_AFFINE || (norm_a && norm_b), EQUSS_ERR_INVALID_ARG,
                "EQUSS_NORM_AFFINE needs norm_a and norm_b");
  EQUSS_REQUIRE(algo >= EQUSS_ASSIGN_AUTO && algo <= EQUSS_ASSIGN_TCGEN05_TF32, EQUSS_ERR_INVALID_ARG, "bad algo %d", algo);
  if (zd->n_pixels == 0) return EQUSS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = assign_tc_supported(zd, M, K, d, norm_mode, margin_out != nullptr);
  const bool tch_ok = assign_tch_supported(zd, M, K, d, norm_mode, margin_out != nullptr);
  if ((algo == EQUSS_ASSIGN_TCGEN05 || algo == EQUSS_ASSIGN_TCGEN05_TF32) && !tc_ok) {
    set_error("equss_pq_assign: shape (layout=%d M=%d K=%d d=%d norm=%d) is not supported by the tcgen05 kernel",
              zd->layout, M, K, d, norm_mode);
    return EQUSS_ERR_UNSUPPORTED;
  }
  if (tch_ok && (algo == EQUSS_ASSIGN_AUTO || algo == EQUSS_ASSIGN_TCGEN05)) {
    return assign_tch_launch(z, zd, codebook_norm, cnorm2, M, K, d, idx_out, workspace, workspace_bytes, st);
  }
  if (tc_ok && algo != EQUSS_ASSIGN_SIMT) {
    return assign_tc_launch(z, zd, codebook_norm, cnorm2, M, K, d, norm_mode, norm_a, norm_b, idx_out,
                            workspace, workspace_bytes, st);
  }
  return launch_assign_simt(z, zd, codebook_norm, cnorm2, M, K, d, norm_mode, norm_a, norm_b, idx_out,
                            margin_out, st);
}

extern "C" int equss_pq_assign_gather_supported(const equss_zdesc* zd, int M, int K, int d, int norm_mode) {
  return (zd && zd->n_pixels > 0 && validate_zdesc(zd, M, d) == EQUSS_OK && assign_tch_fusable(zd, M, K, d, norm_mode)) ? 1 : 0;
}

extern "C" int equss_pq_assign_gather(const float* z, const equss_zdesc* zd, const float* codebook_norm,
                                      const float* cnorm2, const float* gather_src, int M, int K, int d,
                                      int norm_mode, int32_t* idx_out, float* out, double* sqerr, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
  if (zd && zd->n_pixels == 0) return EQUSS_OK;
  EQUSS_REQUIRE(z && zd && codebook_norm && cnorm2 && gather_src && idx_out && out && sqerr, EQUSS_ERR_INVALID_ARG,
                "equss_pq_assign_gather: null pointer");
  int rc = validate_zdesc(zd, M, d); if (rc) return rc;
  EQUSS_REQUIRE(equss_pq_assign_gather_supported(zd, M, K, d, norm_mode), EQUSS_ERR_UNSUPPORTED,
                "equss_pq_assign_gather: needs l2 rows, d in {16,32}, K <= 256, flat or NCHW layout "
                "(layout=%d M=%d K=%d d=%d norm=%d); call equss_pq_assign + equss_pq_gather_loss instead",
                zd->layout, M, K, d, norm_mode);
  return assign_tch_launch(z, zd, codebook_norm, cnorm2, M, K, d, idx_out, workspace, workspace_bytes,
                           (cudaStream_t)stream, gather_src, out, sqerr);
}

extern "C" int equss_pq_distance_prob(const float* z, const equss_zdesc* zd, const float* codebook_norm,
                                      const float* cnorm2, int M, int K, int d, int norm_mode,
                                      const float* norm_a, const float* norm_b, float temperature,
                                      float* prob, void* stream) {
  if (zd && zd->n_pixels == 0) return EQUSS_OK;
  EQUSS_REQUIRE(z && zd && codebook_norm && cnorm2 && prob, EQUSS_ERR_INVALID_ARG, "equss_pq_distance_prob: null pointer");
  int rc = validate_zdesc(zd, M, d); if (rc) return rc;
  EQUSS_REQUIRE(K > 0 && K <= 2048, EQUSS_ERR_UNSUPPORTED, "equss_pq_distance_prob: K=%d outside (0,2048]", K);
  EQUSS_REQUIRE(norm_mode >= EQUSS_NORM_NONE && norm_mode <= EQUSS_NORM_AFFINE, EQUSS_ERR_INVALID_ARG,
                "Unsupported normalize type %d", norm_mode);
  EQUSS_REQUIRE(norm_mode != EQUSS_NORM_AFFINE || (norm_a && norm_b), EQUSS_ERR_INVALID_ARG,
                "EQUSS_NORM_AFFINE needs norm_a and norm_b");
  EQUSS_REQUIRE(temperature != 0.f, EQUSS_ERR_INVALID_ARG, "temperature must be non-zero");
  if (zd->n_pixels == 0) return EQUSS_OK;
  if (distance_prob_tiled_supported(K, d, prob) && getenv("EQUSS_PROB_WARP") == nullptr)
    return distance_prob_tiled_launch(z, zd, codebook_norm, cnorm2, M, K, d, norm_mode, norm_a, norm_b, temperature, prob,
                                      (cudaStream_t)stream);
  const int threads = 256, nwarps = threads / 32;
  size_t smem = ((size_t)K * (d + 1) + K + (size_t)nwarps * d) * sizeof(float);
  EQUSS_REQUIRE(smem <= 200 * 1024, EQUSS_ERR_UNSUPPORTED,
                "equss_pq_distance_prob: codebook K=%d d=%d does not fit shared memory", K, d);
  ZView zv = make_view(zd);
  long long bx = (zd->n_pixels + nwarps - 1) / nwarps;
  long long capx = (long long)num_sms() * 8 / M + 1;
  if (bx > capx) bx = capx;
  dim3 grid((unsigned)bx, (unsigned)M);
  cudaStream_t st = (cudaStream_t)stream;
  const int kpl = (K + 31) / 32;
#define EQUSS_PROB_CASE(KPLV)                                                                             \
  {                                                                                                       \
    if (smem > 48 * 1024)                                                                                 \
      EQUSS_CUDA_OK(cudaFuncSetAttribute(distance_prob_kernel<KPLV>,                                      \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    distance_prob_kernel<KPLV><<<grid, threads, smem, st>>>(z, zv, codebook_norm, cnorm2, M, K, d,        \
                                                            norm_mode, norm_a, norm_b, temperature, prob); \
  }
  if (kpl <= 1) EQUSS_PROB_CASE(1)
  else if (kpl <= 2) EQUSS_PROB_CASE(2)
  else if (kpl <= 4) EQUSS_PROB_CASE(4)
  else if (kpl <= 8) EQUSS_PROB_CASE(8)
  else if (kpl <= 16) EQUSS_PROB_CASE(16)
  else if (kpl <= 32) EQUSS_PROB_CASE(32)
  else EQUSS_PROB_CASE(64)
#undef EQUSS_PROB_CASE
  EQUSS_LAUNCH_OK("distance_prob_kernel");
  return EQUSS_OK;
}
