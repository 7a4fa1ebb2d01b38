// head_gemm_tc.cu -- the expansion head in front of the PQ quantiser:
//     code = cluster1(x) + cluster2(x),  cluster1 = Conv1x1(C -> D),  cluster2 = Conv1x1(C -> C), ReLU, Conv1x1(C -> D)
// (model/dino_pqgo.py:104-112,127-128; model/blocks/module.py:20-44) as two launches of one tcgen05 / TMEM GEMM
//     out[r][o] = act( sum_k A[r][k] * W[o][k] + bias[o] ),   r = pixel, 128 x 256 output tiles,
//   1. h    = relu(W2 x + b2)                          A = x (NCHW, read in place),           W = W2      [C][C]
//   2. code = [W1 | W3] [x ; h] + (b1 + b3)            A = x (NCHW) then h (flat, K-major),   W = [W1|W3] [D][2C]
// The second launch is the sum of the two branches as ONE contraction over 2C channels, so `code` is written once, in
// the flat (pixel, channel) layout the PQ kernels prefer (the reference permutes NCHW -> NHWC there,
// model/dino_pqgo.py:583-584).
//
// Arithmetic: split-tf32 (hi.hi + lo.hi + hi.lo, fp32 accumulate) -- the reference's CPU path is an fp32 GEMM, its GPU
// path is cuDNN TF32 (1e-3); this kernel stays within ~1e-5 of the fp64 result (tests/test_gpu_head.py).  The raw
// operand tiles ARE the hi operands (the tensor core ignores the low 13 mantissa bits of an fp32 word); eight convert
// warps write lo = x - trunc(x) element-wise at identical (swizzled) offsets, which is layout-agnostic: an NCHW
// activation tile arrives through a 4-D tensor map (32 pixels | channel | 32-pixel block | image) in the MN-major
// SWIZZLE_128B_BASE32B layout, a flat activation or weight tile through a 2-D map in the K-major SWIZZLE_64B layout.
// The hi.hi products and the small lo.hi / hi.lo products have separate TMEM accumulators (the tensor core's fp32
// accumulate truncates; the 2^-11-sized terms stay out of the main accumulator); the epilogue adds them, the bias and
// the activation and streams the row.  Structure as in knn_tc.cu: persistent CTAs, producer warp, two issuer warps,
// eight convert warps, four epilogue warps, four 48 KB stages (16 channels each).
#include <cuda.h>
#include <cstdlib>
#include "equss_common.cuh"
#include "equss_tcgen05.cuh"

namespace equss {
namespace headtc {

using namespace ::equss::ptx;

constexpr int kBM = 128, kBN = 256, kKC = 16;        // pixels / output channels per tile, input channels per stage
// What bounds the kernel is the L2 -> shared-memory stream, not the tensor core: with the MMAs and the conversion
// switched off (EQUSS_HEAD_DEBUG=7) it still takes 63 % of its time, ~6 TB/s of operand tiles chip-wide, and a deeper
// raw ring (7 x 24 KB in flight instead of 4) changed nothing -- throughput-, not latency-limited.  The tile shape
// (128 x 256 at fp32 operand width) sets the bytes per flop; 2-CTA MMAs sharing the weight tile are the next step.
constexpr int kStages = 4;                           // 48 KB each
constexpr int kThreads = 32 * (4 + 1 + 2 + 8);       // 4 epilogue, producer, 2 MMA issuers, 8 convert warps
constexpr int kEpiWarp0 = 0, kProducerWarp = 4, kMmaWarp = 5, kConvWarp0 = 7;
constexpr int kRowB = kKC * 4;                       // bytes per K-major row of a stage = the swizzle span (64 B)
constexpr int kARaw = kBM * kRowB, kBRaw = kBN * kRowB;  // bytes: 8 KB, 16 KB
constexpr int kStageBytes = 2 * kARaw + 2 * kBRaw;   // raw A | lo A | raw W | lo W = 48 KB
constexpr int kSmem = 1024 + kStages * kStageBytes + 256;

// kind::tf32, fp32 accumulate, B K-major, M = 128, N; bit 15 = A MN-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int N, bool a_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major ? (1u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

struct Params {
  int n_images, hw, tiles_per_image;    // row tile (b, t) = pixels [t*128, t*128+128) of image b (flat input: one "image")
  int n_out, n_tiles_n, bn;             // output channels, column tiles, columns per tile (256, or 192 when that divides n_out better)
  int kc1, kc2;                         // kKC-channel stages taken from source 1 / source 2
  int a1_nchw;                          // source 1: 0 = flat rows (K-major tiles); NCHW (MN-major tiles): 1 = one 4-D box per
                                        // tile (hw % 32 == 0), 2 = one 3-D box per 32-pixel block (ragged hw, zero-filled)
  int relu;
  const float* bias;                    // [n_out] or null
  float* out;                           // [n_images*hw][out_ld]
  long long out_ld;
  int debug;                            // EQUSS_HEAD_DEBUG timing experiments: 1 = no MMAs, 2 = no small-term MMAs, 4 = no conversion
};

__global__ void __launch_bounds__(kThreads, 1)
head_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_a2,
                    const __grid_constant__ CUtensorMap tmap_w, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* raw_full = bars;                  // [kStages] both TMA boxes landed
  uint64_t* lo_full = raw_full + kStages;     // [kStages] the eight convert warps wrote the lo tiles
  uint64_t* st_empty = lo_full + kStages;     // [kStages] both issuers' MMAs completed
  uint64_t* acc_full = st_empty + kStages;    // [1] tile finished (both issuers)
  uint64_t* acc_empty = acc_full + 1;         // [1] epilogue drained the accumulators
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const long long total = (long long)p.n_images * p.tiles_per_image * p.n_tiles_n;
  const int n_my = (int)((total - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int n_kc = p.kc1 + p.kc2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(raw_full + i, 1); mbar_init(lo_full + i, 8); mbar_init(st_empty + i, 2); }
    mbar_init(acc_full, 2);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);

  // tile t (row-major: the n_tiles_n column tiles of a pixel tile are adjacent, so concurrent CTAs share its
  // activation tiles in L2 and the weights stay L2-resident)
  auto decode = [&](long long t, int& b, int& ti, int& bn) {
    const long long bm = t / p.n_tiles_n;
    bn = (int)(t - bm * p.n_tiles_n);
    b = (int)(bm / p.tiles_per_image);
    ti = (int)(bm - (long long)b * p.tiles_per_image);
  };

  if (warp == kProducerWarp) {
    {   // the whole warp walks the loop (uniform values -> uniform registers), one elected lane issues the copies
      int g = 0;
      for (int it = 0; it < n_my; ++it) {
        int b, ti, bn;
        decode(blockIdx.x + (long long)it * gridDim.x, b, ti, bn);
        const int row0 = b * p.hw + ti * kBM;            // first pixel row of the tile in the flat view
        for (int c = 0; c < n_kc; ++c, ++g) {
          const int st = g % kStages;
          uint8_t* sp = smem + st * kStageBytes;
          mbar_wait(st_empty + st, ((g / kStages) & 1) ^ 1, 10);
          if (elect_one()) {
          mbar_expect_tx(raw_full + st, (uint32_t)(kARaw + p.bn * kRowB));
          if (c < p.kc1) {
            if (p.a1_nchw == 1) {
              tma_load_4d(sp, &tmap_a1, 0, c * kKC, ti * (kBM / 32), b, raw_full + st);
            } else if (p.a1_nchw == 2) {
#pragma unroll
              for (int blk = 0; blk < kBM / 32; ++blk)      // pixels past hw are zero-filled by the TMA unit
                tma_load_3d(sp + blk * (kKC * 128), &tmap_a1, ti * kBM + blk * 32, c * kKC, b, raw_full + st);
            } else {
              tma_load_2d(sp, &tmap_a1, c * kKC, row0, raw_full + st);
            }
          } else {
            tma_load_2d(sp, &tmap_a2, (c - p.kc1) * kKC, row0, raw_full + st);
          }
          tma_load_2d(sp + 2 * kARaw, &tmap_w, c * kKC, bn * p.bn, raw_full + st);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
    // issuer 0: a_hi.w_hi -> main accumulator (columns 0..255); issuer 1: a_lo.w_hi + a_hi.w_lo -> small (256..511)
    const int part = warp - kMmaWarp;
    const uint32_t IDESC_K = make_idesc_tf32(p.bn, false), IDESC_MN = make_idesc_tf32(p.bn, true);
    // K-major SWIZZLE_64B: rows of 64 B (16 channels), 8-row groups 512 B apart (SBO); a K step of 8 floats advances 32 B.
    // MN-major SWIZZLE_128B_BASE32B (NCHW activations): 32 pixels per 128-byte row, 4 channels per 512-byte atom (SBO),
    // 32-pixel blocks kKC * 128 B apart (LBO); a K step of 8 channels advances 1024 B.
    const uint32_t k_hi = (uint32_t)(((8u * kRowB) >> 4) & 0x3FFF) | (1u << 14) | (4u << 29);
    const uint32_t mn_hi = (uint32_t)((512u >> 4) & 0x3FFF) | (1u << 14) | (1u << 29);
    const uint32_t base = smem_u32(smem);
    int g = 0;
    for (int it = 0; it < n_my; ++it) {
      mbar_wait(acc_empty, (it & 1) ^ 1, 22);
      const uint32_t d_addr = tmem_base + (uint32_t)(part * kBN);
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int st = g % kStages;
        if (part == 0) mbar_wait(raw_full + st, (g / kStages) & 1, 20);
        else mbar_wait(lo_full + st, (g / kStages) & 1, 21);          // implies raw_full
        tc_fence_after();
        if (elect_one()) {   // elected lane + uniform operands: UTCHMMA issues from uniform registers
          const bool mn = p.a1_nchw && c < p.kc1;
          const uint32_t sa = base + (uint32_t)(st * kStageBytes);
          const uint32_t a_lbo = mn ? (((uint32_t)(kKC * 128) >> 4) << 16) : (1u << 16);
          const uint32_t a_raw = (sa >> 4) | a_lbo, a_lo = ((sa + kARaw) >> 4) | a_lbo;
          const uint32_t a_step = mn ? (1024u >> 4) : 2u, a_hi = mn ? mn_hi : k_hi;
          const uint32_t idesc = mn ? IDESC_MN : IDESC_K;
          const uint32_t b_raw = ((sa + 2 * kARaw) >> 4) | (1u << 16), b_lo = ((sa + 2 * kARaw + kBRaw) >> 4) | (1u << 16);
          if (p.debug & 1) {
          } else if (part == 0) {
#pragma unroll
            for (int kk = 0; kk < kKC / 8; ++kk)
              umma_tf32(d_addr, desc_from(a_raw + a_step * kk, a_hi), desc_from(b_raw + 2 * kk, k_hi), idesc, (c > 0 || kk > 0) ? 1u : 0u);
          } else if (p.debug & 2) {
          } else {
#pragma unroll
            for (int kk = 0; kk < kKC / 8; ++kk)
              umma_tf32(d_addr, desc_from(a_lo + a_step * kk, a_hi), desc_from(b_raw + 2 * kk, k_hi), idesc, (c > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < kKC / 8; ++kk)
              umma_tf32(d_addr, desc_from(a_raw + a_step * kk, a_hi), desc_from(b_lo + 2 * kk, k_hi), idesc, 1u);
          }
          umma_commit(st_empty + st);
          if (c == n_kc - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kConvWarp0) {
    const int ct = threadIdx.x - kConvWarp0 * 32;      // 0..255
    int g = 0;
    for (int it = 0; it < n_my; ++it) {
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int st = g % kStages;
        mbar_wait(raw_full + st, (g / kStages) & 1, 31);
        uint8_t* sp = smem + st * kStageBytes;
        // A: kARaw / 16 float4 (NA per thread), W: kBRaw / 16 (NB per thread); lo tile = raw tile + kARaw / + kBRaw
        constexpr int NA = kARaw / 16 / 256, NB = kBRaw / 16 / 256;
        float4 va[NA], vb[NB];
#pragma unroll
        for (int u = 0; u < NA; ++u) va[u] = *reinterpret_cast<const float4*>(sp + (u * 256 + ct) * 16);
#pragma unroll
        for (int u = 0; u < NB; ++u) vb[u] = *reinterpret_cast<const float4*>(sp + 2 * kARaw + (u * 256 + ct) * 16);
#pragma unroll
        for (int u = 0; u < NA; ++u) {
          if (p.debug & 4) break;
          float4 l;
          l.x = va[u].x - tf32_trunc(va[u].x); l.y = va[u].y - tf32_trunc(va[u].y);
          l.z = va[u].z - tf32_trunc(va[u].z); l.w = va[u].w - tf32_trunc(va[u].w);
          *reinterpret_cast<float4*>(sp + kARaw + (u * 256 + ct) * 16) = l;
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          if (p.debug & 4) break;
          float4 l;
          l.x = vb[u].x - tf32_trunc(vb[u].x); l.y = vb[u].y - tf32_trunc(vb[u].y);
          l.z = vb[u].z - tf32_trunc(vb[u].z); l.w = vb[u].w - tf32_trunc(vb[u].w);
          *reinterpret_cast<float4*>(sp + 2 * kARaw + kBRaw + (u * 256 + ct) * 16) = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(lo_full + st);
      }
    }
  } else {
    // epilogue warps 0-3: TMEM lane quarter = warp, one pixel row per thread
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool vec_ok = (p.out_ld & 3) == 0 && (p.n_out & 3) == 0;
    for (int it = 0; it < n_my; ++it) {
      int b, ti, bn;
      decode(blockIdx.x + (long long)it * gridDim.x, b, ti, bn);
      const int s = ti * kBM + row;
      const bool live = s < p.hw;
      const int c0 = bn * p.bn;
      float* orow = p.out + ((long long)b * p.hw + s) * p.out_ld + c0;
      mbar_wait(acc_full, it & 1, 40);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < p.bn / 32; ++ch) {
        uint32_t vm[32], vs[32];
        tmem_ld32(lane_base + (uint32_t)(ch * 32), vm);
        tmem_ld32(lane_base + (uint32_t)(kBN + ch * 32), vs);
        tmem_ld_wait();
        const int cc = c0 + ch * 32;
        if (live && cc < p.n_out) {
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(vm[j]) + __uint_as_float(vs[j]);
            if (p.bias && cc + j < p.n_out) v += __ldg(p.bias + cc + j);
            r[j] = p.relu ? fmaxf(v, 0.f) : v;
          }
          float* o = orow + ch * 32;
          if (cc + 32 <= p.n_out && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) __stcs(reinterpret_cast<float4*>(o + j), make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cc + j < p.n_out) o[j] = r[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
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

}  // namespace headtc
}  // namespace equss

using namespace equss;

// out[r][o] = act( sum_{k<C1} A1[r][k] W[o][k] + sum_{k<C2} A2[r][k] W[o][C1+k] + bias[o] ),  r = b*hw + s.
//   a1: NCHW [B][C1][hw] when a1_nchw (hw % 32 == 0) else flat [B*hw][C1];  a2: flat [B*hw][C2] or null (C2 = 0)
//   w: [n_out][C1+C2] row-major; out: [B*hw][out_ld] with out_ld >= n_out
extern "C" int equss_head_gemm_supported(int C1, int C2, int hw, int a1_nchw) {
  if (C1 <= 0 || C1 % headtc::kKC != 0 || C2 < 0 || C2 % headtc::kKC != 0) return 0;
  if (a1_nchw && (hw % 4) != 0) return 0;          // channel stride must be a multiple of 16 bytes
  return 1;
}

extern "C" int equss_head_gemm(const float* a1, int a1_nchw, int C1, const float* a2, int C2, int B, int hw,
                               const float* w, const float* bias, int n_out, int relu, float* out, int64_t out_ld,
                               void* stream) {
  using namespace headtc;
  if (B == 0 || hw == 0) return EQUSS_OK;
  EQUSS_REQUIRE(a1 && w && out && (a2 || C2 == 0), EQUSS_ERR_INVALID_ARG, "equss_head_gemm: null pointer");
  EQUSS_REQUIRE(B > 0 && hw > 0 && n_out > 0 && out_ld >= n_out, EQUSS_ERR_INVALID_ARG,
                "equss_head_gemm: bad shape B=%d hw=%d n_out=%d out_ld=%lld", B, hw, n_out, (long long)out_ld);
  EQUSS_REQUIRE(equss_head_gemm_supported(C1, C2, hw, a1_nchw), EQUSS_ERR_UNSUPPORTED,
                "equss_head_gemm: needs C1 %% %d == 0, C2 %% %d == 0 and, for NCHW input, h*w %% 4 == 0 (C1=%d C2=%d hw=%d)",
                kKC, kKC, C1, C2, hw);
  EQUSS_REQUIRE(!((uintptr_t)a1 & 15) && !((uintptr_t)a2 & 15) && !((uintptr_t)w & 15) && !((uintptr_t)out & 15),
                EQUSS_ERR_INVALID_ARG, "equss_head_gemm: pointers must be 16-byte aligned");
  PFN_encodeTiled encode = get_encode_fn();
  EQUSS_REQUIRE(encode != nullptr, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const long long rows = (long long)B * hw;
  const int K = C1 + C2;
  auto make2d = [&](CUtensorMap* tm, const float* base, long long nrows, int ncols, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)ncols, (cuuint64_t)nrows};
    cuuint64_t gstr[1] = {(cuuint64_t)ncols * 4};
    cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap ta1, ta2, tw;
  CUresult c1;
  int a1_mode = a1_nchw ? ((hw % 32) == 0 ? 1 : 2) : 0;
  if (a1_mode == 2) {
    // ragged token grid (e.g. 28 x 28): pixel | channel | image, one box of 32 pixels x 16 channels per block
    cuuint64_t gdim[3] = {(cuuint64_t)hw, (cuuint64_t)C1, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)hw * 4, (cuuint64_t)C1 * hw * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)kKC, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    c1 = encode(&ta1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a1, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (a1_mode == 1) {
    // dims: 32 pixels (one 128-byte swizzle row) | channel | 32-pixel block | image -> smem [block][channel][32 px]
    cuuint64_t gdim[4] = {32, (cuuint64_t)C1, (cuuint64_t)(hw / 32), (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)hw * 4, 128, (cuuint64_t)C1 * hw * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)kKC, (cuuint32_t)(kBM / 32), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    c1 = encode(&ta1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a1, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    c1 = make2d(&ta1, a1, rows, C1, kBM);
  }
  CUresult c2 = (C2 > 0) ? make2d(&ta2, a2, rows, C2, kBM) : CUDA_SUCCESS;
  if (C2 == 0) ta2 = ta1;
  // 192-column tiles when they waste fewer columns than 256-column tiles (e.g. n_out = 384: 2 x 192 instead of 2 x 256)
  const int waste256 = (n_out + 255) / 256 * 256 - n_out, waste192 = (n_out + 191) / 192 * 192 - n_out;
  const int bn = (waste192 < waste256) ? 192 : 256;
  CUresult c3 = make2d(&tw, w, n_out, K, bn);
  EQUSS_REQUIRE(c1 == CUDA_SUCCESS && c2 == CUDA_SUCCESS && c3 == CUDA_SUCCESS, EQUSS_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d, %d, %d)", (int)c1, (int)c2, (int)c3);
  Params p;
  p.n_images = B; p.hw = hw; p.tiles_per_image = (hw + kBM - 1) / kBM;
  p.bn = bn;
  p.n_out = n_out; p.n_tiles_n = (n_out + p.bn - 1) / p.bn;
  p.kc1 = C1 / kKC; p.kc2 = C2 / kKC;
  p.a1_nchw = a1_mode; p.relu = relu ? 1 : 0;
  p.bias = bias; p.out = out; p.out_ld = out_ld;
  p.debug = getenv("EQUSS_HEAD_DEBUG") ? atoi(getenv("EQUSS_HEAD_DEBUG")) : 0;
  const long long total = (long long)p.n_images * p.tiles_per_image * p.n_tiles_n;
  int grid = num_sms();
  if (total < grid) grid = (int)total;
  EQUSS_CUDA_OK(cudaFuncSetAttribute(head_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  head_gemm_tc_kernel<<<grid, kThreads, kSmem, (cudaStream_t)stream>>>(ta1, ta2, tw, p);
  EQUSS_LAUNCH_OK("head_gemm_tc_kernel");
  return EQUSS_OK;
}
