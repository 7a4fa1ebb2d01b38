// pq_prob.cu -- K2 soft assignment softmax(-distance / T) (model/quantizer.py:468,609; dino_pqgo.py:655), register-tiled.
//
// The N x (M*K) output (3.4 GB at the cocostuff27 shape) makes this an HBM-write-bound op; the first kernel (one warp
// per row, two shared-memory loads per FMA) ran at 0.4 TB/s.  Here a block owns 64 pixels x all K <= 256 codes of one
// subspace: the codebook sits transposed in shared memory ([d][256], loaded once per block), the normalised rows
// transposed as [d][64], and each of the 16 x 16 threads accumulates a 4 x 16 micro-tile (5 LDS.128 per 64 FMA) with
// the canonical sequential-fma dot product.  The row softmax (max, exp, sum over the 16 threads that share a row) uses
// xor-shuffles; rows are written as 16-byte pieces that form 256-byte contiguous runs per warp instruction.
#include "equss_common.cuh"

namespace equss {

constexpr int kProbRows = 64, kProbCodes = 256;

template <int D>
__global__ void __launch_bounds__(256, 2)
distance_prob_tiled_kernel(const float* __restrict__ z, ZView zv, const float* __restrict__ cb, const float* __restrict__ cn2,
                           int M, int K, int mode, const float* __restrict__ na, const float* __restrict__ nb,
                           float temperature, float* __restrict__ prob, long long tiles_per_block) {
  extern __shared__ __align__(16) float s_mem[];
  float* s_c = s_mem;                              // [D][256] codebook, transposed
  float* s_cn2 = s_c + D * kProbCodes;             // [256]
  float* s_z = s_cn2 + kProbCodes;                 // [D][64] normalised rows, transposed
  float* s_zn2 = s_z + D * kProbRows;              // [64]
  const int m = blockIdx.y;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // codebook of this subspace; codes past K get +inf norm: distance +inf, probability 0
  for (int i = tid; i < kProbCodes * D; i += 256) {
    const int k = i / D, j = i - k * D;
    s_c[j * kProbCodes + k] = (k < K) ? __ldg(cb + ((long long)m * K + k) * D + j) : 0.f;
  }
  for (int k = tid; k < kProbCodes; k += 256) s_cn2[k] = (k < K) ? __ldg(cn2 + (long long)m * K + k) : INFINITY;
  const long long n_tiles = (zv.n_pixels + kProbRows - 1) / kProbRows;
  long long t0 = (long long)blockIdx.x * tiles_per_block, t1 = t0 + tiles_per_block;
  if (t1 > n_tiles) t1 = n_tiles;
  // raw rows of the next tile are fetched while the current tile is being computed (threads 0..63, D <= 32)
  constexpr bool kPrefetch = (D <= 32);
  float xn[kPrefetch ? D : 1];
  auto fetch = [&](long long t, float* x) {
    const long long n = t * kProbRows + tid;
    if (n < zv.n_pixels) {
      const long long base = pixel_base(zv, n) + (long long)m * D * zv.stride_c;
#pragma unroll
      for (int j = 0; j < D; ++j) x[j] = __ldg(z + base + j * zv.stride_c);
    } else {
#pragma unroll
      for (int j = 0; j < D; ++j) x[j] = 0.f;
    }
  };
  if (kPrefetch && tid < kProbRows && t0 < t1) fetch(t0, xn);
  const float inv_t = 1.f / temperature;
  for (long long t = t0; t < t1; ++t) {
    __syncthreads();                               // previous tile's readers are done (and the codebook is in place)
    if (tid < kProbRows) {
      const long long n = t * kProbRows + tid;
      float x[D];
      if (kPrefetch) {
#pragma unroll
        for (int j = 0; j < D; ++j) x[j] = xn[j];
      } else {
        fetch(t, x);
      }
      if (n < zv.n_pixels) {
        if (mode == EQUSS_NORM_L2 && (D % 4) == 0) {
          // branch-free, bit-identical to sqrtf / IEEE division (equss_common.cuh)
          float g[D / 4];
#pragma unroll
          for (int i = 0; i < D / 4; ++i) g[i] = group_sumsq(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          const float denom = l2_denom_fast(butterfly_array<D / 4>(g));
#pragma unroll
          for (int i = 0; i < D / 4; ++i) {
            const float4 q = div4_fast(make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]), denom);
            x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
          }
        } else {
          const RowNorm r = row_norm_generic(mode, D, [&](int j) { return x[j]; });
#pragma unroll
          for (int j = 0; j < D; ++j) {
            float v = x[j];
            if (mode == EQUSS_NORM_AFFINE) v = (v - __ldg(na + m * D + j)) / __ldg(nb + m * D + j);
            else v = apply_norm(v, r, mode);
            x[j] = v;
          }
        }
        s_zn2[tid] = canonical_sumsq(D, [&](int j) { return x[j]; });
      } else {
        s_zn2[tid] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < D; ++j) s_z[j * kProbRows + tid] = x[j];
      if (kPrefetch && t + 1 < t1) fetch(t + 1, xn);
    }
    __syncthreads();
    float acc[4][16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[r][c] = 0.f;
#pragma unroll 4
    for (int j = 0; j < D; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(s_z + j * kProbRows + ty * 4);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(s_c + j * kProbCodes + q * 64 + tx * 4);
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
      const int row = ty * 4 + r;
      const long long n = t * kProbRows + row;
      const float zn2 = s_zn2[row];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        acc[r][c] = -ref_distance(zn2, c2[c], acc[r][c]) * inv_t;            // softmax(-d / ts) (dino_pqgo.py:655)
        mx = fmaxf(mx, acc[r][c]);
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) { acc[r][c] = __expf(acc[r][c] - mx); sum += acc[r][c]; }   // rel. error <= |x| 2^-23
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      if (n < zv.n_pixels) {
        float* o = prob + (n * M + m) * (long long)K;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = q * 64 + tx * 4;
          if (k + 3 < K && (K & 3) == 0) {
            __stcs(reinterpret_cast<float4*>(o + k),
                   make_float4(acc[r][4 * q] * inv, acc[r][4 * q + 1] * inv, acc[r][4 * q + 2] * inv, acc[r][4 * q + 3] * inv));
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (k + e < K) o[k + e] = acc[r][4 * q + e] * inv;
          }
        }
      }
    }
  }
}

template <int D>
static int launch_tiled(const float* z, const ZView& zv, const float* cb, const float* cn2, int M, int K, int mode,
                        const float* na, const float* nb, float temperature, float* prob, cudaStream_t st) {
  const size_t smem = (size_t)(D * kProbCodes + kProbCodes + D * kProbRows + kProbRows) * sizeof(float);
  EQUSS_CUDA_OK(cudaFuncSetAttribute(distance_prob_tiled_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_tiles = (zv.n_pixels + kProbRows - 1) / kProbRows;
  long long bx = (long long)num_sms() * 4 / M + 1;          // ~4 blocks per SM in total: the codebook load amortises
  if (bx > n_tiles) bx = n_tiles;
  const long long tpb = (n_tiles + bx - 1) / bx;
  bx = (n_tiles + tpb - 1) / tpb;
  dim3 grid((unsigned)bx, (unsigned)M);
  distance_prob_tiled_kernel<D><<<grid, 256, smem, st>>>(z, zv, cb, cn2, M, K, mode, na, nb, temperature, prob, tpb);
  EQUSS_LAUNCH_OK("distance_prob_tiled_kernel");
  return EQUSS_OK;
}

bool distance_prob_tiled_supported(int K, int d, const float* prob) {
  return K <= kProbCodes && (d == 8 || d == 16 || d == 32 || d == 64) && !((uintptr_t)prob & 15);
}

int distance_prob_tiled_launch(const float* z, const equss_zdesc* zd, const float* cb, const float* cn2, int M, int K,
                               int d, int mode, const float* na, const float* nb, float temperature, float* prob,
                               cudaStream_t st) {
  const ZView zv = make_view(zd);
  switch (d) {
    case 8: return launch_tiled<8>(z, zv, cb, cn2, M, K, mode, na, nb, temperature, prob, st);
    case 16: return launch_tiled<16>(z, zv, cb, cn2, M, K, mode, na, nb, temperature, prob, st);
    case 32: return launch_tiled<32>(z, zv, cb, cn2, M, K, mode, na, nb, temperature, prob, st);
    default: return launch_tiled<64>(z, zv, cb, cn2, M, K, mode, na, nb, temperature, prob, st);
  }
}

}  // namespace equss
