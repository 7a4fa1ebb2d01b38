// pq_stats.cu -- reductions around the PQ head that the reference evaluates on materialised tensors.
//
//  * soft_stats_kernel: the two consumers of distance_prob = softmax(-distance / T) in
//    model/dino_new_vq.py:447-450 -- JSDLoss (model/loss.py:508-525) between the two halves of the batch and
//    EntropyLoss (model/loss.py:490-505) of the first half's mean assignment -- WITHOUT writing the N x (K*M)
//    probabilities (3.4 GB at the cocostuff27 shape).  Same register tiling as distance_prob_tiled_kernel (pq_prob.cu);
//    a tile holds 32 pixel PAIRS (pixel n of the first half, pixel n + N/2 of the second) arranged so that both
//    members of a pair land in the same thread, whose 4 x 16 micro-tile then yields the pair's KL terms directly.
//  * channel_moments kernels (SURVEY 8e "z_trainable normalisation stats"): per-channel sum and sum of squares of the
//    activations in one pass (model/quantizer.py:433-434 makes two passes per subspace), fp64 accumulators.
#include "equss_common.cuh"

namespace equss {

constexpr int kStatRows = 64, kStatCodes = 256, kStatPairs = 32;

template <int D>
__global__ void __launch_bounds__(256, 2)
soft_stats_kernel(const float* __restrict__ z, ZView zv, const float* __restrict__ cb, const float* __restrict__ cn2,
                  int M, int K, int mode, const float* __restrict__ na, const float* __restrict__ nb, float temperature,
                  double* __restrict__ kl_sum, double* __restrict__ psum, long long tiles_per_block) {
  extern __shared__ __align__(16) float s_mem[];
  float* s_c = s_mem;                              // [D][256] codebook, transposed
  float* s_cn2 = s_c + D * kStatCodes;             // [256]
  float* s_z = s_cn2 + kStatCodes;                 // [D][64] normalised rows, transposed
  float* s_zn2 = s_z + D * kStatRows;              // [64]
  float* s_p = s_zn2 + kStatRows;                  // [256] block-level sum of first-half probabilities
  __shared__ double s_kl[8];
  const int m = blockIdx.y;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long half = zv.n_pixels / 2;
  for (int i = tid; i < kStatCodes * D; i += 256) {
    const int k = i / D, j = i - k * D;
    s_c[j * kStatCodes + k] = (k < K) ? __ldg(cb + ((long long)m * K + k) * D + j) : 0.f;
  }
  for (int k = tid; k < kStatCodes; k += 256) {
    s_cn2[k] = (k < K) ? __ldg(cn2 + (long long)m * K + k) : INFINITY;
    s_p[k] = 0.f;
  }
  const long long n_tiles = (half + kStatPairs - 1) / kStatPairs;
  long long t0 = (long long)blockIdx.x * tiles_per_block, t1 = t0 + tiles_per_block;
  if (t1 > n_tiles) t1 = n_tiles;
  const float inv_t = 1.f / temperature;
  float pacc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) pacc[c] = 0.f;
  double kl_thread = 0.0;
  for (long long t = t0; t < t1; ++t) {
    __syncthreads();
    if (tid < kStatRows) {
      // row `tid` of the tile: rows 4*y + {0,1} are pixels 2y, 2y+1 of the first half, rows 4*y + {2,3} their partners
      const int y = tid >> 2, rr = tid & 3;
      const long long a = t * kStatPairs + y * 2 + (rr & 1);
      const long long n = (rr < 2) ? a : half + a;
      float x[D];
      if (a < half) {
        const long long base = pixel_base(zv, n) + (long long)m * D * zv.stride_c;
#pragma unroll
        for (int j = 0; j < D; ++j) x[j] = __ldg(z + base + j * zv.stride_c);
        const RowNorm r = row_norm_generic(mode, D, [&](int j) { return x[j]; });
#pragma unroll
        for (int j = 0; j < D; ++j) {
          float v = x[j];
          if (mode == EQUSS_NORM_AFFINE) v = (v - __ldg(na + m * D + j)) / __ldg(nb + m * D + j);
          else v = apply_norm(v, r, mode);
          x[j] = v;
        }
        s_zn2[tid] = canonical_sumsq(D, [&](int j) { return x[j]; });
      } else {
#pragma unroll
        for (int j = 0; j < D; ++j) x[j] = 0.f;
        s_zn2[tid] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < D; ++j) s_z[j * kStatRows + tid] = x[j];
    }
    __syncthreads();
    float acc[4][16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[r][c] = 0.f;
#pragma unroll 4
    for (int j = 0; j < D; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(s_z + j * kStatRows + ty * 4);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(s_c + j * kStatCodes + q * 64 + tx * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][4 * q + 0] = fmaf(av[r], b.x, acc[r][4 * q + 0]);
          acc[r][4 * q + 1] = fmaf(av[r], b.y, acc[r][4 * q + 1]);
          acc[r][4 * q + 2] = fmaf(av[r], b.z, acc[r][4 * q + 2]);
          acc[r][4 * q + 3] = fmaf(av[r], b.w, acc[r][4 * q + 3]);
        }
      }
    }
    float c2[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 c = *reinterpret_cast<const float4*>(s_cn2 + q * 64 + tx * 4);
      c2[4 * q] = c.x; c2[4 * q + 1] = c.y; c2[4 * q + 2] = c.z; c2[4 * q + 3] = c.w;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float zn2 = s_zn2[ty * 4 + r];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        acc[r][c] = -ref_distance(zn2, c2[c], acc[r][c]) * inv_t;
        mx = fmaxf(mx, acc[r][c]);
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) { acc[r][c] = __expf(acc[r][c] - mx); sum += acc[r][c]; }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[r][c] *= inv;
    }
    // pairs (row 0, row 2) and (row 1, row 3):  sum_k (p+e) log((p+e)/mix) + (q+e) log((q+e)/mix),  mix = (p+q+e)/2
    float kl = 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long a = t * kStatPairs + ty * 2 + r;
      if (a < half) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int k = (c >> 2) * 64 + tx * 4 + (c & 3);
          if (k < K) {
            const float p = acc[r][c], q = acc[r + 2][c];
            const float pe = p + 1e-6f, qe = q + 1e-6f;
            const float rmix = 2.f / (p + q + 1e-6f);
            kl += pe * logf(pe * rmix) + qe * logf(qe * rmix);
            pacc[c] += p;
          }
        }
      }
    }
    kl_thread += (double)kl;
  }
  // block reductions: KL -> one double atomic per block; probabilities -> shared float bins -> one atomic per code
#pragma unroll
  for (int c = 0; c < 16; ++c) atomicAdd(&s_p[(c >> 2) * 64 + tx * 4 + (c & 3)], pacc[c]);
  kl_thread = warp_sum_d(kl_thread);
  if ((tid & 31) == 0) s_kl[tid >> 5] = kl_thread;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += s_kl[i];
    atomicAdd(kl_sum + m, s);
  }
  for (int k = tid; k < K; k += 256) atomicAdd(psum + (long long)m * K + k, (double)s_p[k]);
}

template <int D>
static int launch_soft_stats(const float* z, const ZView& zv, const float* cb, const float* cn2, int M, int K, int mode,
                             const float* na, const float* nb, float temperature, double* kl_sum, double* psum,
                             cudaStream_t st) {
  const size_t smem = (size_t)(D * kStatCodes + kStatCodes + D * kStatRows + kStatRows + kStatCodes) * sizeof(float);
  EQUSS_CUDA_OK(cudaFuncSetAttribute(soft_stats_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_tiles = (zv.n_pixels / 2 + kStatPairs - 1) / kStatPairs;
  long long bx = (long long)num_sms() * 4 / M + 1;
  if (bx > n_tiles) bx = n_tiles;
  const long long tpb = (n_tiles + bx - 1) / bx;
  bx = (n_tiles + tpb - 1) / tpb;
  dim3 grid((unsigned)bx, (unsigned)M);
  soft_stats_kernel<D><<<grid, 256, smem, st>>>(z, zv, cb, cn2, M, K, mode, na, nb, temperature, kl_sum, psum, tpb);
  EQUSS_LAUNCH_OK("soft_stats_kernel");
  return EQUSS_OK;
}

// ---- per-channel moments -----------------------------------------------------------------------------------------
// flat rows (channel stride 1): thread = channel, block strides over a slab of pixels; coalesced along channels
__global__ void __launch_bounds__(256)
channel_moments_flat_kernel(const float* __restrict__ z, ZView zv, int D, double* __restrict__ out, long long rows_per_block) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= D) return;
  long long n0 = (long long)blockIdx.x * rows_per_block, n1 = n0 + rows_per_block;
  if (n1 > zv.n_pixels) n1 = zv.n_pixels;
  double s = 0.0, ss = 0.0;
  for (long long n = n0; n < n1; n += 64) {          // fp32 partials over 64 rows, fp64 across them
    float a = 0.f, b = 0.f;
    const long long e = (n + 64 < n1) ? n + 64 : n1;
    for (long long i = n; i < e; ++i) {
      const float v = __ldg(z + pixel_base(zv, i) + c);
      a += v; b = fmaf(v, v, b);
    }
    s += a; ss += b;
  }
  atomicAdd(out + c, s);
  atomicAdd(out + D + c, ss);
}

// NCHW (pixel stride 1): one warp per (image, channel) plane
__global__ void __launch_bounds__(256)
channel_moments_nchw_kernel(const float* __restrict__ z, ZView zv, int D, long long planes, double* __restrict__ out) {
  const long long plane = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (plane >= planes) return;
  const int lane = threadIdx.x & 31;
  const long long b = plane / D;
  const int c = (int)(plane - b * D);
  const float* p = z + b * zv.stride_b + (long long)c * zv.stride_c;
  double s = 0.0, ss = 0.0;
  for (long long i0 = 0; i0 < zv.hw; i0 += 32 * 16) {
    float a = 0.f, q = 0.f;
#pragma unroll 4
    for (int u = 0; u < 16; ++u) {
      const long long i = i0 + u * 32 + lane;
      if (i < zv.hw) { const float v = __ldg(p + i * zv.stride_s); a += v; q = fmaf(v, v, q); }
    }
    s += a; ss += q;
  }
  s = warp_sum_d(s); ss = warp_sum_d(ss);
  if (lane == 0) { atomicAdd(out + c, s); atomicAdd(out + D + c, ss); }
}

}  // namespace equss

using namespace equss;

extern "C" int equss_pq_soft_stats_supported(int K, int d) {
  return (K > 0 && K <= kStatCodes && (d == 8 || d == 16 || d == 32 || d == 64)) ? 1 : 0;
}

extern "C" int equss_pq_soft_stats(const float* z, const equss_zdesc* zd, const float* codebook_norm, const float* cnorm2,
                                   int M, int K, int d, int norm_mode, const float* norm_a, const float* norm_b,
                                   float temperature, double* kl_sum, double* prob_sum, void* stream) {
  EQUSS_REQUIRE(z && zd && codebook_norm && cnorm2 && kl_sum && prob_sum, EQUSS_ERR_INVALID_ARG, "equss_pq_soft_stats: null pointer");
  int rc = validate_zdesc(zd, M, d); if (rc) return rc;
  EQUSS_REQUIRE(zd->n_pixels % 2 == 0, EQUSS_ERR_INVALID_ARG,
                "equss_pq_soft_stats: %lld rows cannot be split into two halves (model/dino_new_vq.py:447)", (long long)zd->n_pixels);
  EQUSS_REQUIRE(equss_pq_soft_stats_supported(K, d), EQUSS_ERR_UNSUPPORTED,
                "equss_pq_soft_stats: needs K <= 256 and d in {8,16,32,64} (K=%d d=%d)", K, d);
  EQUSS_REQUIRE(norm_mode >= EQUSS_NORM_NONE && norm_mode <= EQUSS_NORM_AFFINE, EQUSS_ERR_INVALID_ARG,
                "Unsupported normalize type %d", norm_mode);
  EQUSS_REQUIRE(norm_mode != EQUSS_NORM_AFFINE || (norm_a && norm_b), EQUSS_ERR_INVALID_ARG, "EQUSS_NORM_AFFINE needs norm_a and norm_b");
  EQUSS_REQUIRE(temperature != 0.f, EQUSS_ERR_INVALID_ARG, "temperature must be non-zero");
  if (zd->n_pixels == 0) return EQUSS_OK;
  const ZView zv = make_view(zd);
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 8: return launch_soft_stats<8>(z, zv, codebook_norm, cnorm2, M, K, norm_mode, norm_a, norm_b, temperature, kl_sum, prob_sum, st);
    case 16: return launch_soft_stats<16>(z, zv, codebook_norm, cnorm2, M, K, norm_mode, norm_a, norm_b, temperature, kl_sum, prob_sum, st);
    case 32: return launch_soft_stats<32>(z, zv, codebook_norm, cnorm2, M, K, norm_mode, norm_a, norm_b, temperature, kl_sum, prob_sum, st);
    default: return launch_soft_stats<64>(z, zv, codebook_norm, cnorm2, M, K, norm_mode, norm_a, norm_b, temperature, kl_sum, prob_sum, st);
  }
}

extern "C" int equss_channel_moments(const float* z, const equss_zdesc* zd, double* sums, void* stream) {
  EQUSS_REQUIRE(zd && sums, EQUSS_ERR_INVALID_ARG, "equss_channel_moments: null pointer");
  if (zd->n_pixels == 0) return EQUSS_OK;
  EQUSS_REQUIRE(z, EQUSS_ERR_INVALID_ARG, "equss_channel_moments: null pointer");
  EQUSS_REQUIRE(zd->dim > 0 && zd->hw > 0 && zd->n_pixels % zd->hw == 0, EQUSS_ERR_INVALID_ARG, "equss_channel_moments: bad descriptor");
  const ZView zv = make_view(zd);
  const int D = zd->dim;
  cudaStream_t st = (cudaStream_t)stream;
  if (zd->stride_c == 1) {
    long long bx = (long long)num_sms() * 8;
    const long long by = (D + 255) / 256;
    bx = bx / by + 1;
    long long rpb = (zd->n_pixels + bx - 1) / bx;
    rpb = (rpb + 63) / 64 * 64;
    bx = (zd->n_pixels + rpb - 1) / rpb;
    dim3 grid((unsigned)bx, (unsigned)by);
    channel_moments_flat_kernel<<<grid, 256, 0, st>>>(z, zv, D, sums, rpb);
    EQUSS_LAUNCH_OK("channel_moments_flat_kernel");
  } else {
    const long long planes = (zd->n_pixels / zd->hw) * D;
    channel_moments_nchw_kernel<<<(unsigned)((planes + 7) / 8), 256, 0, st>>>(z, zv, D, planes, sums);
    EQUSS_LAUNCH_OK("channel_moments_nchw_kernel");
  }
  return EQUSS_OK;
}
