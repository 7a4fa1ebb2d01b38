// probe_logits_tc.cu -- K8 step 1 on the tensor cores: token-resolution probe logits
//     logits[b*hw + s][j] = sum_c feat[b][c][s] * w[j][c] + bias[j]          (model/evaluator.py:67,98-100)
// as a tcgen05 / TMEM GEMM  [128 pixels x D] x [D x 64]  per tile, split-tf32 (hi.hi + lo.hi + hi.lo) so the
// result carries fp32-level accuracy (the argmax that follows is audited against the fp32 reference).
//
// The op is HBM-bound (it reads the feature map once: 4*D bytes per token, 2*D*C flops); the SIMT version was
// bound by fp32 FMA issue instead.  Structure (persistent CTAs, 10 warps):
//   warp 4      producer: TMA boxes of the NCHW feature map (32 channels x 128 pixels, pixel-contiguous) into a
//               ring of raw stages + bulk copies of the matching 32-channel slice of the weight image
//   warps 6-9   convert: raw[channel][pixel] -> (hi, lo) tf32 pieces in the UMMA K-major core-matrix layout
//   warp 5      MMA issuer: 12 tcgen05.mma.kind::tf32 (128 x 64 x 8) per 32-channel stage
//   warps 0-3   epilogue: TMEM -> registers, + bias, 16-byte stores of the token's C_pad logits
// Accumulation: the tensor core's fp32 accumulate truncates, which over the D/8 = 128 dependent steps of one
// output drifts by ~1e-5 of the logit scale.  The dominant hi.hi products are therefore accumulated in TMEM for
// only kPromote stages (64 channels) at a time; the epilogue warps add each partial to fp32 registers with
// round-to-nearest adds (ping-pong TMEM buffers, so the tensor core never waits).  The small lo.hi / hi.lo terms
// (2^-11 of the result) keep their own whole-tile accumulator, where the drift is irrelevant.
#include <cuda.h>
#include "equss_common.cuh"
#include "equss_tcgen05.cuh"

namespace equss {
namespace ptc {

using namespace ::equss::ptx;

constexpr int kTileM = 128;        // pixels per tile
constexpr int kN = 64;             // accumulator columns (probe channels, zero-padded)
constexpr int kKC = 32;            // feature channels per pipeline stage
constexpr int kThreads = 320;
constexpr int kEpiWarp0 = 0, kProducerWarp = 4, kMmaWarp = 5, kConvWarp0 = 6;
constexpr int kRawStages = 4, kABufs = 2, kBBufs = 3;
constexpr int kPromote = 2;        // stages per TMEM partial of the hi.hi products
constexpr int kRawBytes = kKC * kTileM * 4;                 // 16 KB
constexpr int kAPiece = kTileM * kKC * 4;                   // 16 KB: one of (hi, lo)
constexpr int kABytes = 2 * kAPiece;
constexpr int kBPiece = kN * kKC * 4;                       // 8 KB
constexpr int kBBytes = 2 * kBPiece;                        // per 32-channel slice: hi then lo
constexpr int kALBO = 128, kASBO = (kKC / 4) * kALBO;       // K-major, no swizzle: 8 chunks of 16 B per row
constexpr int kBLBO = 128, kBSBO = (kKC / 4) * kBLBO;
constexpr int kSmem = 128 + kRawStages * kRawBytes + kABufs * kABytes + kBBufs * kBBytes + 512;

__host__ __device__ constexpr uint32_t make_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

struct Params {
  int B, D, hw, c_total, c_pad;
  int tiles_per_image, n_tiles, n_kc;
  const uint8_t* image;      // [D/32][hi 8 KB | lo 8 KB]
  const float* bias;
  float* logits;
};

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Weight image: slice c (32 channels) = [hi | lo], each 64 rows (probe channel j) x 32 channels in the UMMA
// K-major core-matrix layout: 16-byte chunk q of row j at (j/8)*SBO + q*128 + (j%8)*16.
__global__ void __launch_bounds__(256)
build_probe_image_kernel(const float* __restrict__ wmat_t, int D, int c_pad, uint8_t* __restrict__ image) {
  const int c = blockIdx.x;                       // slice
  uint8_t* img = image + (size_t)c * kBBytes;
  for (int i = threadIdx.x; i < kN * (kKC / 4); i += blockDim.x) {
    const int j = i / (kKC / 4), q = i % (kKC / 4);
    float v[4], hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ch = c * kKC + 4 * q + e;
      v[e] = (j < c_pad && ch < D) ? wmat_t[(size_t)ch * c_pad + j] : 0.f;
      // hi: round to nearest tf32 (the tensor core ignores the low 13 bits); lo: exact remainder
      uint32_t u = __float_as_uint(v[e]);
      u += 0x00000FFFu + ((u >> 13) & 1u);
      hi[e] = __uint_as_float(u & 0xFFFFE000u);
      if (!(fabsf(v[e]) < 3.0e38f)) hi[e] = v[e];
      lo[e] = v[e] - hi[e];
    }
    uint8_t* p = img + (size_t)(j / 8) * kBSBO + q * kBLBO + (j % 8) * 16;
    *reinterpret_cast<float4*>(p) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(p + kBPiece) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
probe_logits_tc_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* s_raw = smem;
  uint8_t* s_a = s_raw + kRawStages * kRawBytes;
  uint8_t* s_b = s_a + kABufs * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + kBBufs * kBBytes);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + kRawStages;
  uint64_t* a_full = raw_empty + kRawStages;
  uint64_t* a_empty = a_full + kABufs;
  uint64_t* b_full = a_empty + kABufs;
  uint64_t* b_empty = b_full + kBBufs;
  uint64_t* m_full = b_empty + kBBufs;      // [2] main (hi.hi) partial ready
  uint64_t* m_empty = m_full + 2;           // [2]
  uint64_t* s_full = m_empty + 2;           // [2] small-term accumulator of a tile ready
  uint64_t* s_empty = s_full + 2;           // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_kc = p.n_kc;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRawStages; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, 4); }
    for (int i = 0; i < kABufs; ++i) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < kBBufs; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(m_full + i, 1); mbar_init(m_empty + i, 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, 4); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<4 * kN>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      int g = 0;
      for (int it = 0; it < n_my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int b = tile / p.tiles_per_image;
        const int s0 = (tile - b * p.tiles_per_image) * kTileM;
        for (int c = 0; c < n_kc; ++c, ++g) {
          const int rs = g % kRawStages, bs = g % kBBufs;
          mbar_wait(b_empty + bs, ((g / kBBufs) & 1) ^ 1, 11);
          mbar_expect_tx(b_full + bs, kBBytes);
          bulk_load_1d(s_b + bs * kBBytes, p.image + (size_t)c * kBBytes, kBBytes, b_full + bs);
          mbar_wait(raw_empty + rs, ((g / kRawStages) & 1) ^ 1, 10);
          mbar_expect_tx(raw_full + rs, kRawBytes);
          tma_load_3d(s_raw + rs * kRawBytes, &tmap, s0, c * kKC, b, raw_full + rs);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    constexpr uint32_t IDESC = make_idesc_tf32(kN);
    const uint32_t a_hi = (uint32_t)((kASBO >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t b_hi = (uint32_t)((kBSBO >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t a_lo0 = (smem_u32(s_a) >> 4) | ((uint32_t)(kALBO >> 4) << 16);
    const uint32_t b_lo0 = (smem_u32(s_b) >> 4) | ((uint32_t)(kBLBO >> 4) << 16);
    int g = 0, G = 0;                       // global stage / promotion-group counters
    for (int it = 0; it < n_my_tiles; ++it) {
      const int sb = it & 1;
      mbar_wait(s_empty + sb, ((it >> 1) & 1) ^ 1, 22);
      const uint32_t d_small = tmem_base + (uint32_t)((2 + sb) * kN);
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int as = g % kABufs, bs = g % kBBufs;
        const int mb = G & 1;
        const bool group_first = (c % kPromote) == 0;
        const bool group_last = (c % kPromote) == kPromote - 1 || c == n_kc - 1;
        if (group_first) mbar_wait(m_empty + mb, ((G >> 1) & 1) ^ 1, 23);
        mbar_wait(a_full + as, (g / kABufs) & 1, 20);
        mbar_wait(b_full + bs, (g / kBBufs) & 1, 21);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_lo = a_lo0 + (uint32_t)(as * (kABytes >> 4));
          const uint32_t b_lo = b_lo0 + (uint32_t)(bs * (kBBytes >> 4));
          const uint32_t d_main = tmem_base + (uint32_t)(mb * kN);
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            const uint32_t a_off = (part == 1) ? (kAPiece >> 4) : 0;     // x_lo for the middle product
            const uint32_t b_off = (part == 2) ? (kBPiece >> 4) : 0;     // w_lo for the last product
#pragma unroll
            for (int kk = 0; kk < kKC / 8; ++kk) {
              const uint32_t acc = (part == 0) ? ((!group_first || kk > 0) ? 1u : 0u)
                                               : ((c > 0 || part > 1 || kk > 0) ? 1u : 0u);
              umma_tf32(part == 0 ? d_main : d_small, desc_from(a_lo + a_off + (uint32_t)(2 * kk * (kALBO >> 4)), a_hi),
                        desc_from(b_lo + b_off + (uint32_t)(2 * kk * (kBLBO >> 4)), b_hi), IDESC, acc);
            }
          }
          umma_commit(a_empty + as);
          umma_commit(b_empty + bs);
          if (group_last) umma_commit(m_full + mb);
          if (c == n_kc - 1) umma_commit(s_full + sb);
        }
        __syncwarp();
        if (group_last) ++G;
      }
    }
  } else if (warp >= kConvWarp0) {
    const int row = threadIdx.x - kConvWarp0 * 32;     // pixel row of the tile
    int g = 0;
    for (int it = 0; it < n_my_tiles; ++it) {
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int rs = g % kRawStages, as = g % kABufs;
        mbar_wait(a_empty + as, ((g / kABufs) & 1) ^ 1, 30);
        mbar_wait(raw_full + rs, (g / kRawStages) & 1, 31);
        const float* raw = reinterpret_cast<const float*>(s_raw + rs * kRawBytes);
        float x[kKC];
#pragma unroll
        for (int j = 0; j < kKC; ++j) x[j] = raw[j * kTileM + row];
        uint8_t* rowp = s_a + as * kABytes + (row / 8) * kASBO + (row % 8) * 16;
#pragma unroll
        for (int q = 0; q < kKC / 4; ++q) {
          float4 hi, lo;
          hi.x = tf32_trunc(x[4 * q]); hi.y = tf32_trunc(x[4 * q + 1]); hi.z = tf32_trunc(x[4 * q + 2]); hi.w = tf32_trunc(x[4 * q + 3]);
          lo.x = x[4 * q] - hi.x; lo.y = x[4 * q + 1] - hi.y; lo.z = x[4 * q + 2] - hi.z; lo.w = x[4 * q + 3] - hi.w;
          *reinterpret_cast<float4*>(rowp + q * kALBO) = hi;
          *reinterpret_cast<float4*>(rowp + kAPiece + q * kALBO) = lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(raw_empty + rs); mbar_arrive(a_full + as); }
      }
    }
  } else {
    // epilogue warps 0-3: TMEM lane quarter = warp
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int n4 = p.c_pad >> 2;
    const int n_groups = (n_kc + kPromote - 1) / kPromote;
    int G = 0;
    for (int it = 0; it < n_my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / p.tiles_per_image;
      const int s = (tile - b * p.tiles_per_image) * kTileM + row;
      const int sb = it & 1;
      float acc[kN];
#pragma unroll
      for (int j = 0; j < kN; ++j) acc[j] = 0.f;
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int gi = 0; gi < n_groups; ++gi, ++G) {
        const int mb = G & 1;
        mbar_wait(m_full + mb, (G >> 1) & 1, 40);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_base + (uint32_t)(mb * kN), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
        tmem_ld32(lane_base + (uint32_t)(mb * kN + 32), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(m_empty + mb);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] += __uint_as_float(v[j]);
      }
      {
        mbar_wait(s_full + sb, (it >> 1) & 1, 41);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_base + (uint32_t)((2 + sb) * kN), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
        tmem_ld32(lane_base + (uint32_t)((2 + sb) * kN + 32), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + sb);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] += __uint_as_float(v[j]);
      }
      if (s < p.hw) {
        float* o = p.logits + ((long long)b * p.hw + s) * p.c_pad;
#pragma unroll
        for (int j4 = 0; j4 < kN / 4; ++j4) {
          if (j4 >= n4) break;
          float r[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * j4 + e;
            r[e] = acc[j];
            if (p.bias && j < p.c_total) r[e] += __ldg(p.bias + j);
          }
          *reinterpret_cast<float4*>(o + 4 * j4) = make_float4(r[0], r[1], r[2], r[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<4 * kN>(tmem_base);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (PFN_encodeTiled)ptr;
  return fn;
}

}  // namespace ptc
}  // namespace equss

using namespace equss;

extern "C" int64_t equss_probe_image_bytes(int D, int c_total) {
  if (D <= 0 || c_total <= 0 || D % ptc::kKC != 0 || ((c_total + 3) & ~3) > ptc::kN) return 0;
  return (int64_t)(D / ptc::kKC) * ptc::kBBytes;
}

extern "C" int equss_probe_build_image(const float* wmat_t, int D, int c_total, void* image, void* stream) {
  EQUSS_REQUIRE(wmat_t && image, EQUSS_ERR_INVALID_ARG, "equss_probe_build_image: null pointer");
  EQUSS_REQUIRE(equss_probe_image_bytes(D, c_total) > 0, EQUSS_ERR_UNSUPPORTED,
                "equss_probe_build_image: D=%d must be a multiple of %d and C_pad <= %d", D, ptc::kKC, ptc::kN);
  EQUSS_REQUIRE(!((uintptr_t)image & 15), EQUSS_ERR_INVALID_ARG, "equss_probe_build_image: image must be 16-byte aligned");
  ptc::build_probe_image_kernel<<<D / ptc::kKC, 256, 0, (cudaStream_t)stream>>>(wmat_t, D, (c_total + 3) & ~3, (uint8_t*)image);
  EQUSS_LAUNCH_OK("build_probe_image_kernel");
  return EQUSS_OK;
}

extern "C" int equss_probe_logits_tc_supported(int D, int h, int w, int c_total) {
  return equss_probe_image_bytes(D, c_total) > 0 && ((h * w) % 4) == 0;
}

extern "C" int equss_probe_logits_tc(const float* feat, int B, int D, int h, int w, const void* image, const float* bias,
                                     int c_total, float* logits, void* stream) {
  using namespace ptc;
  EQUSS_REQUIRE(feat && image && logits, EQUSS_ERR_INVALID_ARG, "equss_probe_logits_tc: null pointer");
  EQUSS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0 && c_total > 0, EQUSS_ERR_INVALID_ARG,
                "equss_probe_logits_tc: bad shape B=%d D=%d h=%d w=%d C=%d", B, D, h, w, c_total);
  EQUSS_REQUIRE(equss_probe_logits_tc_supported(D, h, w, c_total), EQUSS_ERR_UNSUPPORTED,
                "equss_probe_logits_tc: needs D %% %d == 0, C_pad <= %d, h*w %% 4 == 0", kKC, kN);
  EQUSS_REQUIRE(!((uintptr_t)feat & 15) && !((uintptr_t)logits & 15) && !((uintptr_t)image & 15), EQUSS_ERR_INVALID_ARG,
                "equss_probe_logits_tc: pointers must be 16-byte aligned");
  PFN_encodeTiled encode = get_encode_fn();
  EQUSS_REQUIRE(encode != nullptr, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int hw = h * w;
  CUtensorMap tmap;
  cuuint64_t gdim[3] = {(cuuint64_t)hw, (cuuint64_t)D, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)hw * 4, (cuuint64_t)D * hw * 4};
  cuuint32_t box[3] = {(cuuint32_t)kTileM, (cuuint32_t)kKC, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)feat, gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EQUSS_REQUIRE(cr == CUDA_SUCCESS, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)cr);
  Params p;
  p.B = B; p.D = D; p.hw = hw; p.c_total = c_total; p.c_pad = (c_total + 3) & ~3;
  p.tiles_per_image = (hw + kTileM - 1) / kTileM;
  p.n_tiles = B * p.tiles_per_image;
  p.n_kc = D / kKC;
  p.image = (const uint8_t*)image; p.bias = bias; p.logits = logits;
  int grid = num_sms();
  if (p.n_tiles < grid) grid = p.n_tiles;
  EQUSS_CUDA_OK(cudaFuncSetAttribute(probe_logits_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  probe_logits_tc_kernel<<<grid, kThreads, kSmem, (cudaStream_t)stream>>>(tmap, p);
  EQUSS_LAUNCH_OK("probe_logits_tc_kernel");
  return EQUSS_OK;
}
