// probe_logits_tc.cu -- K8 step 1 on the tensor cores: token-resolution probe logits
//     logits[b*hw + s][j] = sum_c feat[b][c][s] * w[j][c] + bias[j]          (model/evaluator.py:67,98-100)
// as a tcgen05 / TMEM GEMM  [128 pixels x D] x [D x 64]  per tile, split-tf32 (hi.hi + lo.hi + hi.lo) so the
// result carries fp32-level accuracy (the argmax that follows is audited against the fp32 reference).
//
// The op is HBM-bound (it reads the feature map once: 4*D bytes per token, 2*D*C flops); the SIMT version was
// bound by fp32 FMA issue instead.  Structure (persistent CTAs, 13 warps, stages of 32 channels; the raw ring is 8
// deep -- 128 KB of loads in flight per SM is what it takes to cover the HBM latency -- while the lo and weight
// rings, which only decouple convert / L2 from the tensor core, are 2 and 3 deep):
//   warp 4      producer: per stage one TMA box of the NCHW feature map (32 channels x 4 x 32 pixels, 128-byte
//               rows, 32-byte-atom swizzle); warp 10: bulk copies of the matching slice of the weight image.
//               The raw tile IS the hi operand:
//               it lands in the canonical MN-major SWIZZLE_128B_BASE32B layout and the tensor core ignores the low 13
//               mantissa bits of an fp32 word (= truncation to tf32)
//   warps 6-9   convert: lo = x - trunc_tf32(x), element-wise at identical (swizzled) offsets into the lo tile
//   warps 5-7   MMA issuers, one per product (hi.hi / lo.hi / hi.lo): 4 tcgen05.mma.kind::tf32 (128 x 64 x 8) each
//               per stage, A MN-major, B K-major.  A 128x64x8 MMA executes in 48 cycles but costs ~130 cycles of
//               descriptor set-up to issue from one thread, so a single issuer would bound the kernel.
//   warps 0-3   epilogue: TMEM partial sums -> fp32 registers, + bias, 16-byte stores of the token's C_pad logits
// Accumulation: the tensor core's fp32 accumulate truncates, which over the D/8 = 128 dependent steps of one
// output drifts by ~1e-5 of the logit scale.  The dominant hi.hi products are therefore accumulated in TMEM for
// only kPromote stages (64 channels) at a time; the epilogue warps add each partial to fp32 registers with
// round-to-nearest adds (ping-pong TMEM buffers, so the tensor core never waits).  The small lo.hi / hi.lo terms
// (2^-11 of the result) keep their own whole-tile accumulator, where the drift is irrelevant.
#include <cuda.h>
#include <cstdlib>
#include "equss_common.cuh"
#include "equss_tcgen05.cuh"

namespace equss {
namespace ptc {

using namespace ::equss::ptx;

constexpr int kTileM = 128;        // pixels per tile
constexpr int kN = 64;             // accumulator columns (probe channels, zero-padded)
constexpr int kKC = 32;            // feature channels per pipeline stage
constexpr int kThreads = 416;
constexpr int kEpiWarp0 = 0, kProducerWarp = 4, kMmaWarp = 5 /* 5,6,7: one issuer per product */, kConvWarp0 = 8, kBProducerWarp = 12;
constexpr int kRawStages = 8, kLoBufs = 2, kBBufs = 3;
constexpr int kPromote = 2;        // stages per TMEM partial of the hi.hi products
constexpr int kAPiece = kTileM * kKC * 4;                   // 16 KB: raw (= hi) or lo tile
constexpr int kBPiece = kN * kKC * 4;                       // 8 KB
constexpr int kBBytes = 2 * kBPiece;                        // per 32-channel slice: hi then lo
// A (MN-major fp32/tf32 operands use the SWIZZLE_128B_BASE32B layout = TMA's 128B_ATOM_32B: 32-byte chunks of a
// 128-byte row XOR-ed with the row index mod 4): 32 pixels contiguous (128 B), 4 channels per 512-byte swizzle atom
// (SBO), the four 32-pixel blocks of a tile 4096 B apart (LBO); one MMA (K = 8) spans two atoms = 1024 B.
// B (K-major, no swizzle): 8 chunks of 16 B per row.
constexpr int kALBO = kKC * 128, kASBO = 512;
constexpr int kBLBO = 128, kBSBO = (kKC / 4) * kBLBO;
constexpr int kSmem = 1024 + (kRawStages + kLoBufs) * kAPiece + kBBufs * kBBytes + 512;

// kind::tf32, fp32 accumulate, A MN-major (bit 15), B K-major, M = 128, N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

struct Params {
  int B, D, hw, c_total, c_pad;
  int tiles_per_image, n_tiles, n_kc;
  const uint8_t* image;      // [D/32][hi 8 KB | lo 8 KB]
  const float* bias;
  float* logits;
  int debug;                 // EQUSS_PROBE_DEBUG: 1 = issue no MMAs, 2 = also skip the lo conversion, 4 = skip promotions (timing experiments)
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
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // swizzle atoms are 1024-byte aligned
  uint8_t* s_raw = smem;                                  // [kRawStages][16 KB]  raw = hi tiles
  uint8_t* s_lo = s_raw + kRawStages * kAPiece;           // [kLoBufs][16 KB]
  uint8_t* s_b = s_lo + kLoBufs * kAPiece;                // [kBBufs][hi 8 KB | lo 8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + kBBufs * kBBytes);
  uint64_t* raw_full = bars;                    // [kRawStages] TMA box landed
  uint64_t* raw_empty = raw_full + kRawStages;  // [kRawStages] the stage's readers (two MMA issuers, convert) are done
  uint64_t* lo_full = raw_empty + kRawStages;   // [kLoBufs] convert warps wrote the lo tile
  uint64_t* lo_empty = lo_full + kLoBufs;       // [kLoBufs] MMAs completed
  uint64_t* b_full = lo_empty + kLoBufs;        // [kBBufs]
  uint64_t* b_empty = b_full + kBBufs;          // [kBBufs]
  uint64_t* m_full = b_empty + kBBufs;          // [2] main (hi.hi) partial ready
  uint64_t* m_empty = m_full + 2;               // [2]
  uint64_t* s_full = m_empty + 2;               // [2] small-term accumulator of a tile ready
  uint64_t* s_empty = s_full + 2;               // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int n_my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_kc = p.n_kc;

  if (threadIdx.x == 0) {
    // a raw stage is free again when the hi.hi / hi.lo issuers (tcgen05.commit) AND the four convert warps (which
    // read it to build the lo tile) are done with it
    for (int i = 0; i < kRawStages; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, 2 + 4); }
    for (int i = 0; i < kLoBufs; ++i) { mbar_init(lo_full + i, 4); mbar_init(lo_empty + i, 1); }         // issuer 1
    for (int i = 0; i < kBBufs; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 3); }
    for (int i = 0; i < 2; ++i) { mbar_init(m_full + i, 1); mbar_init(m_empty + i, 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(s_full + i, 2); mbar_init(s_empty + i, 4); }                 // issuers 1, 2
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);

  if (warp == kProducerWarp) {
    {   // the whole warp walks the loop (uniform values -> uniform registers), one elected lane issues the copy
      int g = 0;
      for (int it = 0; it < n_my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int b = tile / p.tiles_per_image;
        const int blk0 = (tile - b * p.tiles_per_image) * (kTileM / 32);      // first 32-pixel block of the tile
        for (int c = 0; c < n_kc; ++c, ++g) {
          const int rs = g % kRawStages;
          mbar_wait(raw_empty + rs, ((g / kRawStages) & 1) ^ 1, 10);
          if (elect_one()) {
            mbar_expect_tx(raw_full + rs, kAPiece);
            tma_load_4d(s_raw + rs * kAPiece, &tmap, 0, c * kKC, blk0, b, raw_full + rs);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kBProducerWarp) {
    if (lane == 0) {
      int g = 0;
      for (int it = 0; it < n_my_tiles; ++it) {
        for (int c = 0; c < n_kc; ++c, ++g) {
          const int bs = g % kBBufs;
          mbar_wait(b_empty + bs, ((g / kBBufs) & 1) ^ 1, 11);
          if (p.debug & 8) { mbar_arrive(b_full + bs); continue; }
          mbar_expect_tx(b_full + bs, kBBytes);
          bulk_load_1d(s_b + bs * kBBytes, p.image + (size_t)c * kBBytes, kBBytes, b_full + bs);
        }
      }
    }
  } else if (warp >= kMmaWarp && warp < kMmaWarp + 3) {
    // issuer `part`: 0 = x_hi.w_hi -> main partials (promoted), 1 = x_lo.w_hi, 2 = x_hi.w_lo -> small accumulators
    const int part = warp - kMmaWarp;
    constexpr uint32_t IDESC = make_idesc_tf32(kN);
    const uint32_t a_hi = (uint32_t)((kASBO >> 4) & 0x3FFF) | (1u << 14) | (1u << 29);     // SBO, version 1, SWIZZLE_128B_BASE32B
    const uint32_t b_hi = (uint32_t)((kBSBO >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t a_lo0 = (smem_u32(part == 1 ? s_lo : s_raw) >> 4) | ((uint32_t)(kALBO >> 4) << 16);
    const uint32_t b_lo0 = ((smem_u32(s_b) + (part == 2 ? kBPiece : 0)) >> 4) | ((uint32_t)(kBLBO >> 4) << 16);
    int g = 0, G = 0;                       // global stage / promotion-group counters
    for (int it = 0; it < n_my_tiles; ++it) {
      const int sb = it & 1;
      if (part > 0) mbar_wait(s_empty + sb, ((it >> 1) & 1) ^ 1, 22);
      const uint32_t d_small = tmem_base + (uint32_t)((2 * part + sb) * kN);      // columns 128.. (part 1), 256.. (part 2)
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int rs = g % kRawStages, ls = g % kLoBufs, bs = g % kBBufs;
        const int mb = G & 1;
        const bool group_first = (c % kPromote) == 0;
        const bool group_last = (c % kPromote) == kPromote - 1 || c == n_kc - 1;
        if (part == 0 && group_first) mbar_wait(m_empty + mb, ((G >> 1) & 1) ^ 1, 23);
        if (part == 1) mbar_wait(lo_full + ls, (g / kLoBufs) & 1, 20);
        else mbar_wait(raw_full + rs, (g / kRawStages) & 1, 24);
        mbar_wait(b_full + bs, (g / kBBufs) & 1, 21);
        tc_fence_after();
        if (elect_one()) {   // elected lane + uniform operands: UTCHMMA issues from uniform registers
          const uint32_t a_lo = a_lo0 + (uint32_t)((part == 1 ? ls : rs) * (kAPiece >> 4));
          const uint32_t b_lo = b_lo0 + (uint32_t)(bs * (kBBytes >> 4));
          const uint32_t d_addr = (part == 0) ? tmem_base + (uint32_t)(mb * kN) : d_small;
#pragma unroll
          for (int kk = 0; kk < kKC / 8; ++kk) {
            if (p.debug & 1) break;
            const uint32_t acc = (part == 0) ? ((!group_first || kk > 0) ? 1u : 0u) : ((c > 0 || kk > 0) ? 1u : 0u);
            umma_tf32(d_addr, desc_from(a_lo + (uint32_t)(kk * (1024 >> 4)), a_hi),
                      desc_from(b_lo + (uint32_t)(2 * kk * (kBLBO >> 4)), b_hi), IDESC, acc);
          }
          if (part == 1) umma_commit(lo_empty + ls); else umma_commit(raw_empty + rs);
          umma_commit(b_empty + bs);
          if (part == 0 && group_last) umma_commit(m_full + mb);
          if (part > 0 && c == n_kc - 1) umma_commit(s_full + sb);
        }
        __syncwarp();
        if (group_last) ++G;
      }
    }
  } else if (warp >= kConvWarp0 && warp < kConvWarp0 + 4) {
    const int ct = threadIdx.x - kConvWarp0 * 32;     // 0..127
    int g = 0;
    for (int it = 0; it < n_my_tiles; ++it) {
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int rs = g % kRawStages, ls = g % kLoBufs;
        mbar_wait(lo_empty + ls, ((g / kLoBufs) & 1) ^ 1, 30);
        mbar_wait(raw_full + rs, (g / kRawStages) & 1, 31);
        const uint8_t* raw = s_raw + rs * kAPiece;
        uint8_t* lo = s_lo + ls * kAPiece;
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(raw + (u * 128 + ct) * 16);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (p.debug & 2) break;
          float4 l;
          l.x = v[u].x - tf32_trunc(v[u].x); l.y = v[u].y - tf32_trunc(v[u].y);
          l.z = v[u].z - tf32_trunc(v[u].z); l.w = v[u].w - tf32_trunc(v[u].w);
          *reinterpret_cast<float4*>(lo + (u * 128 + ct) * 16) = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(lo_full + ls); mbar_arrive(raw_empty + rs); }
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
#pragma unroll
        for (int part = 1; part <= 2; ++part) {
#pragma unroll
          for (int hcol = 0; hcol < 2; ++hcol) {
            tmem_ld32(lane_base + (uint32_t)((2 * part + sb) * kN + 32 * hcol), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[32 * hcol + j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + sb);
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
  return equss_probe_image_bytes(D, c_total) > 0 && ((h * w) % 32) == 0;
}

extern "C" int equss_probe_logits_tc(const float* feat, int B, int D, int h, int w, const void* image, const float* bias,
                                     int c_total, float* logits, void* stream) {
  using namespace ptc;
  EQUSS_REQUIRE(feat && image && logits, EQUSS_ERR_INVALID_ARG, "equss_probe_logits_tc: null pointer");
  EQUSS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0 && c_total > 0, EQUSS_ERR_INVALID_ARG,
                "equss_probe_logits_tc: bad shape B=%d D=%d h=%d w=%d C=%d", B, D, h, w, c_total);
  EQUSS_REQUIRE(equss_probe_logits_tc_supported(D, h, w, c_total), EQUSS_ERR_UNSUPPORTED,
                "equss_probe_logits_tc: needs D %% %d == 0, C_pad <= %d, h*w %% 32 == 0", kKC, kN);
  EQUSS_REQUIRE(!((uintptr_t)feat & 15) && !((uintptr_t)logits & 15) && !((uintptr_t)image & 15), EQUSS_ERR_INVALID_ARG,
                "equss_probe_logits_tc: pointers must be 16-byte aligned");
  PFN_encodeTiled encode = get_encode_fn();
  EQUSS_REQUIRE(encode != nullptr, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int hw = h * w;
  // dims: 32 pixels (one 128-byte swizzle row) | channel | 32-pixel block | image -> smem [block][channel][32 px]
  CUtensorMap tmap;
  cuuint64_t gdim[4] = {32, (cuuint64_t)D, (cuuint64_t)(hw / 32), (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)hw * 4, 128, (cuuint64_t)D * hw * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)kKC, (cuuint32_t)(kTileM / 32), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)feat, gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EQUSS_REQUIRE(cr == CUDA_SUCCESS, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)cr);
  Params p;
  p.B = B; p.D = D; p.hw = hw; p.c_total = c_total; p.c_pad = (c_total + 3) & ~3;
  p.tiles_per_image = (hw + kTileM - 1) / kTileM;
  p.n_tiles = B * p.tiles_per_image;
  p.n_kc = D / kKC;
  p.image = (const uint8_t*)image; p.bias = bias; p.logits = logits;
  p.debug = getenv("EQUSS_PROBE_DEBUG") ? atoi(getenv("EQUSS_PROBE_DEBUG")) : 0;
  int grid = num_sms();
  if (p.n_tiles < grid) grid = p.n_tiles;
  EQUSS_CUDA_OK(cudaFuncSetAttribute(probe_logits_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  probe_logits_tc_kernel<<<grid, kThreads, kSmem, (cudaStream_t)stream>>>(tmap, p);
  EQUSS_LAUNCH_OK("probe_logits_tc_kernel");
  return EQUSS_OK;
}
