// pq_assign_h_kernel.cuh -- K1 for l2-normalised rows on the 5th-generation tensor cores, second design:
// kind::f16 MMAs on a two-piece fp16 split of both operands, float-domain argmax epilogue, exact fp32
// re-score of rows whose top-2 gap is inside the error bound (model/quantizer.py:457-467).
//
// Arithmetic
//   argmin_k (|z|^2 + |c_k|^2) - 2<z,c_k>  ==  argmax_k  s_k := <z,c_k> - |c_k|^2/2       (z = l2-normalised row)
//   x = z (|x_i| <= 1),  y = beta*c  with beta a power of two such that max|y| in [1/2, 1]  (per subspace)
//   x = x1 + x2 + O(2^-22|x| + 2^-25),  x1 = fp16(x), x2 = fp16(x - x1);  same for y
//   beta*s_k ~= x1.y1 + x2.y1 + x1.y2 + 1*(b1+b2+b3),   b = -beta*|c_k|^2/2 in three fp16 pieces
//   -> 3*d/16 + 1 MMAs of K = 16 per tile (the split-tf32 kernel needs 3*d/8 + 1 of K = 8: 7 vs 4 at d = 16,
//   25 vs 13 at d = 64), operand tiles half the bytes.  The normalisation feeding the GEMM is a single
//   rsqrt + multiply; only the exact path uses the canonical division.
//
// Exact argmin from an approximate GEMM
//   The accumulators are compared as plain floats.  Two orthogonal column partitions -- 16 running class
//   maxima (column mod 16) and the top-2 of the 16-column group maxima -- give the winning column and the exact
//   runner-up score at 1.3 ALU-pipe operations per element (FMNMX3).  A row whose gap exceeds the slot's
//   tolerance (2^-16 R + absolute floor, R >= max|beta s|; the GEMM error is below ~1e-6 R) has a certain
//   winner.  The others (~3e-4 of the rows) are re-scored in the reference's fp32 arithmetic -- canonical z_norm,
//   sequential fma dot, (sum z^2 + sum c^2) - 2 dot, first minimal index -- from global memory, i.e. the same
//   code path as the SIMT kernel, so both kernels return identical indices.
//
// Structure (one persistent CTA per SM, 15 warps): warps 0-3 convert (raw tile -> normalise -> fp16 split ->
// UMMA K-major core-matrix layout), 4-11 epilogue (two groups of four, alternating tiles, TMEM lane quarter =
// warp % 4), 12 TMA producer, 13/14 MMA issuers (one per accumulator half).  A CTA walks a contiguous range of
// units ordered (subspace group, code chunk, pixel tile, subspace in group): the G = 128/(4d) subspaces that
// share a 128-byte line of a flat row are processed back to back, so the second half of every line is an L2 hit,
// and the G operand images stay resident in shared memory.  Codebooks with K > NC are split into chunks whose
// winners are merged with a 64-bit atomicMax on (score, index); cross-chunk near-ties are appended to a list and
// resolved by a full exact scan (one warp per listed row) after the main kernel.
//
// Fused gather (FUSE, single-chunk codebooks, d <= 32): the convert pass leaves the canonical z_norm in the raw stage and
// the epilogue thread that found a row's winner runs K3 for it right away (gather + straight-through value + squared
// error, model/quantizer.py:474,514,534-536), then releases the stage.  The activation tensor is read from HBM once for
// assign + gather instead of twice.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include "equss_common.cuh"
#include "equss_tcgen05.cuh"
#include "pq_assign.h"

namespace equss {
namespace tch {

using namespace ::equss::ptx;
#ifdef EQUSS_DEBUG_WATCHDOG      // debug builds: the printf watchdog names the barrier a warp is stuck on (costs the
#define mbar_wait_nc mbar_wait   // flow-sensitive register allocation, see mbar_wait_nc)
#endif

constexpr int kTileM = 128;
#ifndef EQUSS_EPI_GROUPS
#define EQUSS_EPI_GROUPS 3
#endif
// Register file split by warp group (setmaxnreg): 24 warps are launched with 80 registers each (the CTA's pool is what
// the launch allocated: 24 x 80 = 1920 warp-registers, NOT the SM's 2048 -- asking for more deadlocks in TRY_ALLOC);
// producer / MMA issuers / idle warp give back 40 each, epilogue and convert groups grow to 88.
#if !defined(EQUSS_NO_SETMAXNREG) && !defined(EQUSS_SETMAXNREG)
#define EQUSS_SETMAXNREG
#endif
// Budgets (warp-registers: 12 x epilogue + 8 x convert + 4 x misc <= 1920).  d = 16: the convert pass needs < 50 registers,
// which gives the epilogue (stream + tournaments + fused gather) 96 and no spills; wider subspaces keep 88 / 88.
#ifndef EQUSS_REGS_MISC
#define EQUSS_REGS_MISC 40
#endif
#ifdef EQUSS_REGS_EPI
template <int D> struct RegPlan { static constexpr int epi = EQUSS_REGS_EPI, conv = EQUSS_REGS_CONV; };
#else
template <int D> struct RegPlan { static constexpr int epi = (D == 16) ? 96 : 88, conv = (D == 16) ? 72 : 88; };
#endif
constexpr int kEpiGroups = EQUSS_EPI_GROUPS;     // epilogue groups of four warps, rotating over the units
// A parity wait cannot skip a phase, so every accumulator barrier must always be waited on by the same consumer:
// the t_full / t_empty barriers form a ring of lcm(2 TMEM buffers, kEpiGroups) unit slots (x 2 halves); slot i % R
// belongs to one TMEM buffer and one epilogue group.
constexpr int kTSlots = (kEpiGroups % 2 == 0) ? kEpiGroups : 2 * kEpiGroups;
#ifndef EQUSS_CONV_GROUPS
#define EQUSS_CONV_GROUPS 2
#endif
constexpr int kConvGroups = EQUSS_CONV_GROUPS;   // convert groups of four warps, rotating over the units like the epilogue groups
#ifdef EQUSS_SETMAXNREG   // register file split by warp group (epilogue / convert / the rest): needs complete warp groups
constexpr int kThreads = 32 * (4 * kEpiGroups + 4 * kConvGroups + 4);
#else
constexpr int kThreads = 32 * (4 * kEpiGroups + 4 * kConvGroups + 1 + 2);
#endif
// The SMSP arbiter prefers the highest warp id among eligible warps: the convert warps (the pipeline's critical
// stage) sit above the epilogue warps (ALU-pipe heavy, plenty of slack).  Epilogue and convert groups start at a
// multiple of four, so warp % 4 is the TMEM lane quarter / the SMSP.
constexpr int kEpiWarp0 = 0, kConvWarp0 = 4 * kEpiGroups, kProducerWarp = kConvWarp0 + 4 * kConvGroups, kMmaWarp = kProducerWarp + 1;
constexpr float kTolRel = 1.52587890625e-5f;     // 2^-16
#ifndef EQUSS_EPI_SLEEP_NS
#define EQUSS_EPI_SLEEP_NS 64
#endif
#ifndef EQUSS_PROD_SLEEP_NS
#define EQUSS_PROD_SLEEP_NS 200
#endif

#ifdef EQUSS_TRACE   // scripts/trace_assign.cu: per-unit clock64 stamps of CTA 0 (pipeline timeline)
__device__ long long g_trace[256 * 12];
#define EQUSS_TR(slot, i) do { if (blockIdx.x == 0 && (i) < 256 && lane == 0) g_trace[(i) * 12 + (slot)] = clock64(); } while (0)
__device__ __forceinline__ long long equss_globaltimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define EQUSS_TRG(slot, i) do { if (blockIdx.x == 0 && (i) < 256 && lane == 0) g_trace[(i) * 12 + (slot)] = equss_globaltimer(); } while (0)
#else
#define EQUSS_TR(slot, i) do { } while (0)
#define EQUSS_TRG(slot, i) do { } while (0)
#endif

__host__ __device__ constexpr int kch(int D) { return 2 * (D / 8) + 2; }           // 16-byte K chunks per row
__host__ __device__ constexpr int b_sbo(int D) { return kch(D) * 128; }
__host__ __device__ constexpr int b_bytes(int D, int NC) { return (NC / 8) * b_sbo(D) + 128; }   // + trailer
// A operand: stride between K chunks chosen so that one convert store instruction (8-byte pieces of 32/(D/4)
// rows x D/4 lanes) touches every bank at most twice
__host__ __device__ constexpr int a_lbo(int D) { return D == 16 ? 128 : D == 32 ? 160 : 144; }
__host__ __device__ constexpr int a_sbo(int D) { return kch(D) * a_lbo(D); }
__host__ __device__ constexpr int a_bytes(int D) { return (kTileM / 8) * a_sbo(D); }
__host__ __device__ constexpr int raw_bytes(int D) { return kTileM * D * 4; }
__host__ __device__ constexpr int align_up(int x, int a) { return (x + a - 1) / a * a; }
// Epilogue gather: for d = 16 the gather source of the slot's G subspaces ([NC][D] fp32 each) is kept in shared memory,
// double-buffered by slot parity, so a row's codeword is one shared-memory read away -- for the row-major layout only
// (192 -> 175 us at C2): the channel-major path reads whole 64-byte rows per thread, which lands every quarter-warp on
// two bank groups (4-way conflicts) and loses to the L2 loads (195 vs 189 us)
__host__ __device__ constexpr int gtab_bytes(int D, int NC, int G, bool nchw) { return (D == 16 && !nchw) ? 2 * G * NC * D * 4 : 0; }
__host__ __device__ constexpr int smem_bytes(int D, int NC, int G, int stages, int a_bufs, bool fuse = false, bool nchw = false) {
  return 128 + G * align_up(b_bytes(D, NC), 128) + a_bufs * align_up(a_bytes(D), 128) + stages * raw_bytes(D) +
         (fuse ? gtab_bytes(D, NC, G, nchw) : 0) + 1024;
}
// kind::f16 instruction descriptor: fp32 accumulate, fp16 A/B, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t make_idesc(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

struct Params {
  long long n_pixels, hw;
  int M, K, D, NC, nchunks, G;
  int tiles_per_image;
  long long n_tiles;
  const float* z;
  ZView zv;
  const float* cb;            // codebook_norm [M][K][d]
  const float* cn2;           // [M][K]
  const uint8_t* images;      // [M/G][nchunks][G] operand images
  int img_bytes;
  int32_t* idx_out;
  unsigned long long* merged; // nchunks > 1: [M][N] (sortable score << 32 | index), zero-initialised
  uint32_t* flag_list;        // nchunks > 1: [M*N] rows (m*N + n) with a cross-chunk near-tie (duplicates allowed)
  unsigned int* flag_count;   // nchunks > 1: number of entries in flag_list, zero-initialised
  // fused gather (FUSE kernels only)
  const float* gsrc;          // gather source [M][K][d]
  float* out;                 // same strides as z
  double* sqerr;              // [M], zeroed by build_image_kernel
  int use_gtab;               // epilogue gather reads the codewords from the shared-memory table (slots long enough)
};

// Walks the CTA's contiguous unit range (sslot-major, then tile, then subspace-in-group) without divisions.
struct UnitIter {
  int sslot, tile, g, sg, chunk, n_tiles, nchunks, G;
  int img, timg, tpi;          // NCHW: image of the tile, tile within the image, tiles per image (flat: one "image")
  __device__ __forceinline__ void init(long long u0, int n_tiles_, int nchunks_, int G_, int tpi_ = 0) {
    n_tiles = n_tiles_; nchunks = nchunks_; G = G_;
    const long long per = (long long)n_tiles_ * G_;
    sslot = (int)(u0 / per);
    const int rem = (int)(u0 - (long long)sslot * per);
    tile = rem / G_; g = rem - tile * G_;
    sg = sslot / nchunks_; chunk = sslot - sg * nchunks_;
    tpi = tpi_ > 0 ? tpi_ : n_tiles_;
    img = tile / tpi; timg = tile - img * tpi;
  }
  __device__ __forceinline__ void next() {
    if (++g == G) {
      g = 0;
      if (++timg == tpi) { timg = 0; ++img; }
      if (++tile == n_tiles) {
        tile = 0; img = 0; timg = 0; ++sslot;
        if (++chunk == nchunks) { chunk = 0; ++sg; }
      }
    }
  }
  __device__ __forceinline__ void advance(int k) {
#pragma unroll
    for (int j = 0; j < k; ++j) next();
  }
  __device__ __forceinline__ int m() const { return sg * G + g; }
};

__device__ __forceinline__ unsigned int sortable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unsortable(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// One 32-column chunk of accumulators: 16 running class maxima (column mod 16) and the maxima of the chunk's two
// 16-column groups -- 31 FMNMX3/FMNMX for 32 elements, nothing else in the streaming part of the epilogue.
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[32], float (&cls)[16], float& gm_a, float& gm_b) {
#pragma unroll
  for (int r = 0; r < 16; ++r) cls[r] = max3f(cls[r], __uint_as_float(v[r]), __uint_as_float(v[r + 16]));
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const uint32_t* k = v + 16 * g;
#define F(i) __uint_as_float(k[i])
    float t0 = max3f(F(0), F(1), F(2)), t1 = max3f(F(3), F(4), F(5)), t2 = max3f(F(6), F(7), F(8));
    float t3 = max3f(F(9), F(10), F(11)), t4 = max3f(F(12), F(13), F(14));
    const float gm = fmaxf(max3f(t0, t1, t2), max3f(t3, t4, F(15)));
#undef F
    if (g == 0) gm_a = gm; else gm_b = gm;
  }
}

// Knock-out tournament over 16 values: index of the maximum (the lower index on exactly equal values), the maximum
// and the exact runner-up = the best of the four losers on the winner's path (the second best always loses to the
// best directly).  51 ALU operations; an equality scan plus a masked maximum takes 80.
__device__ __forceinline__ void tournament16(const float (&v)[16], int& win, float& best, float& runner) {
  float ta[8], tb[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) ta[j] = fmaxf(v[2 * j], v[2 * j + 1]);
#pragma unroll
  for (int j = 0; j < 4; ++j) tb[j] = fmaxf(ta[2 * j], ta[2 * j + 1]);
  const float tc0 = fmaxf(tb[0], tb[1]), tc1 = fmaxf(tb[2], tb[3]);
  const bool p3 = tc1 > tc0;
  const float l3 = fminf(tc0, tc1);
  const float sb0 = p3 ? tb[2] : tb[0], sb1 = p3 ? tb[3] : tb[1];
  const bool p2 = sb1 > sb0;
  const float l2 = fminf(sb0, sb1);
  float sa[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) sa[j] = p3 ? ta[4 + j] : ta[j];
  const float sa0 = p2 ? sa[2] : sa[0], sa1 = p2 ? sa[3] : sa[1];
  const bool p1 = sa1 > sa0;
  const float l1 = fminf(sa0, sa1);
  float sc8[8], sc4[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) sc8[j] = p3 ? v[8 + j] : v[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) sc4[j] = p2 ? sc8[4 + j] : sc8[j];
  const float sc0 = p1 ? sc4[2] : sc4[0], sc1 = p1 ? sc4[3] : sc4[1];
  const bool p0 = sc1 > sc0;
  const float l0 = fminf(sc0, sc1);
  win = (p3 ? 8 : 0) + (p2 ? 4 : 0) + (p1 ? 2 : 0) + (p0 ? 1 : 0);
  best = fmaxf(tc0, tc1);
  runner = fmaxf(max3f(l3, l2, l1), l0);
}

// Exact fp32 re-score of the chunk-local columns {r + 16 j : bit r of cmask set, column < kvalid} of row n,
// subspace m, by the whole warp: every lane builds the canonical z_norm from global memory (the SIMT kernel's
// arithmetic), lanes split the candidate columns, the minimum (lowest column on ties) is combined by shuffles.
// Returns the chunk-local column to all lanes.
template <int D>
__device__ __noinline__ int exact_rescore_warp(const Params& p, int m, long long n, int k0, int kvalid, uint32_t cmask, int lane) {
  constexpr int LPS = D / 4;
  // candidate of this lane (first round) -- its loads are issued together with the row's, one memory round trip
  const int ncand = __popc(cmask) * 16;
  const float* cbm = p.cb + ((long long)m * p.K + k0) * D;
  const float* cn2m = p.cn2 + (long long)m * p.K + k0;
  int c = lane;
  int col = (c < ncand) ? (int)__fns(cmask, 0, (c >> 4) + 1) + 16 * (c & 15) : kvalid;
  constexpr bool kPrefetch = (D <= 32);      // d = 64: the row alone fills the register file
  float4 cv[LPS];
  float c2 = 0.f;
  if (kPrefetch && col < kvalid) {
    const float4* c4 = reinterpret_cast<const float4*>(cbm + (long long)col * D);
#pragma unroll
    for (int q = 0; q < LPS; ++q) cv[q] = __ldg(c4 + q);
    c2 = __ldg(cn2m + col);
  }
  float x[D];
  const long long base = pixel_base(p.zv, n) + (long long)m * D * p.zv.stride_c;
#pragma unroll
  for (int j = 0; j < D; ++j) x[j] = __ldg(p.z + base + j * p.zv.stride_c);
  float gsum[LPS];
#pragma unroll
  for (int q = 0; q < LPS; ++q) gsum[q] = group_sumsq(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
  const RowNorm rn = l2_from_sumsq(butterfly_array<LPS>(gsum));
#pragma unroll
  for (int j = 0; j < D; ++j) x[j] = x[j] / rn.denom;
#pragma unroll
  for (int q = 0; q < LPS; ++q) gsum[q] = group_sumsq(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
  const float zn2 = butterfly_array<LPS>(gsum);
  float best = INFINITY;
  int best_col = 0x7fffffff;
#pragma unroll 1
  while (true) {
    if (col < kvalid) {
      if (!kPrefetch) {
        const float4* c4 = reinterpret_cast<const float4*>(cbm + (long long)col * D);
#pragma unroll
        for (int q = 0; q < LPS; ++q) cv[q] = __ldg(c4 + q);
        c2 = __ldg(cn2m + col);
      }
      float dot = 0.f;
#pragma unroll
      for (int q = 0; q < LPS; ++q) {
        dot = fmaf(x[4 * q], cv[q].x, dot);
        dot = fmaf(x[4 * q + 1], cv[q].y, dot);
        dot = fmaf(x[4 * q + 2], cv[q].z, dot);
        dot = fmaf(x[4 * q + 3], cv[q].w, dot);
      }
      const float dd = ref_distance(zn2, c2, dot);
      if (dd < best || (dd == best && col < best_col)) { best = dd; best_col = col; }
    }
    c += 32;
    if (c >= ncand) break;
    col = (int)__fns(cmask, 0, (c >> 4) + 1) + 16 * (c & 15);
    if (kPrefetch && col < kvalid) {
      const float4* c4 = reinterpret_cast<const float4*>(cbm + (long long)col * D);
#pragma unroll
      for (int q = 0; q < LPS; ++q) cv[q] = __ldg(c4 + q);
      c2 = __ldg(cn2m + col);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, s);
    const int oc = __shfl_xor_sync(0xffffffffu, best_col, s);
    if (ob < best || (ob == best && oc < best_col)) { best = ob; best_col = oc; }
  }
  return best_col == 0x7fffffff ? 0 : best_col;
}

template <int D, int NC, int G, int STAGES, int ABUFS, bool NCHW, bool FUSE, int LAG>
__global__ void __launch_bounds__(kThreads, 1)
assign_f16x2_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  constexpr int SBO = b_sbo(D);
  constexpr int ASBO = a_sbo(D), ALBO = a_lbo(D);
  constexpr int B_BYTES = align_up(b_bytes(D, NC), 128);
  constexpr int A_BYTES = align_up(a_bytes(D), 128);
  constexpr int RAW_BYTES = raw_bytes(D);
  constexpr int TMEM_COLS = (2 * NC <= 32) ? 32 : (2 * NC <= 64) ? 64 : (2 * NC <= 128) ? 128 : (2 * NC <= 256) ? 256 : 512;
  constexpr int LPS = D / 4;                   // lanes per pixel row in the flat convert
  constexpr int C8 = D / 8;                    // 16-byte chunks per operand piece
  constexpr int HALVES = (NC >= 64) ? 2 : 1;   // accumulator halves with their own barriers / issuer warps
#ifdef EQUSS_DBG_CONV_UNFUSED          // timing experiment: the convert pass of the unfused kernel inside the fused one
  constexpr bool CFUSE = false;
#else
  constexpr bool CFUSE = FUSE;
#endif
  constexpr int NH = NC / HALVES;
  constexpr uint32_t IDESC = make_idesc(NH);
  static_assert(NC % 32 == 0 && NC <= 256, "NC must be a multiple of 32, at most 256");
  static_assert(D == 16 || D == 32 || D == 64, "D must be 16, 32 or 64");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* s_b = smem;                                   // [G][B_BYTES]
  uint8_t* s_a = s_b + G * B_BYTES;                      // [ABUFS][A_BYTES]
  uint8_t* s_rawt = s_a + ABUFS * A_BYTES;               // [STAGES][RAW_BYTES]
  float* s_gt = reinterpret_cast<float*>(s_rawt + STAGES * RAW_BYTES);        // FUSE, row-major d = 16: gather table [2][G][NC][D]
  constexpr int GTAB = FUSE ? gtab_bytes(D, NC, G, NCHW) : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rawt + STAGES * RAW_BYTES + GTAB);
  uint64_t* raw_full = bars;                    // [STAGES]
  uint64_t* raw_empty = raw_full + STAGES;      // [STAGES]
  uint64_t* a_full = raw_empty + STAGES;        // [ABUFS]
  uint64_t* a_empty = a_full + ABUFS;           // [ABUFS]   arrived by tcgen05.commit: the MMAs have read A (and B)
  uint64_t* t_full = a_empty + ABUFS;           // [kTSlots][2]  (unit slot, half)
  uint64_t* t_empty = t_full + 2 * kTSlots;     // [kTSlots][2]
  uint64_t* b_full = t_empty + 2 * kTSlots;     // [1]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(b_full + 1);

  // the warp index goes through a shuffle so that ptxas knows it is warp-uniform: the role branches below are then
  // uniform branches and what the tcgen05 / TMA instructions consume can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  const long long total_units = (long long)p.M * p.nchunks * p.n_tiles;
  const long long u0 = total_units * blockIdx.x / gridDim.x;
  const long long u1 = total_units * (blockIdx.x + 1) / gridDim.x;
  const int n_units = (int)(u1 - u0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, 4); }
    for (int i = 0; i < ABUFS; ++i) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, HALVES); }
    for (int i = 0; i < 2 * kTSlots; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<TMEM_COLS>(s_tmem);
  // constant bias chunks of the A operand: fp16 (1,1,1,0,0,0,0,0) and zeros
  for (int i = threadIdx.x; i < ABUFS * kTileM; i += blockDim.x) {
    const int a = i / kTileM, row = i % kTileM;
    uint8_t* rowp = s_a + a * A_BYTES + (row / 8) * ASBO + (row % 8) * 16;
    *reinterpret_cast<uint4*>(rowp + (2 * C8) * ALBO) = make_uint4(0x3C003C00u, 0x00003C00u, 0u, 0u);
    *reinterpret_cast<uint4*>(rowp + (2 * C8 + 1) * ALBO) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
  // Programmatic dependent launch: everything above (barriers, TMEM allocation, constant operand rows) ran while
  // build_image_kernel was still writing the operand images; from here on its results (images, cleared counters) are used
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // ... and the exact-rescan kernel behind this one may be scheduled as soon as SMs free up (it waits the same way)
  asm volatile("griddepcontrol.launch_dependents;");

  if (warp >= kProducerWarp) {
#ifdef EQUSS_SETMAXNREG      // (each setmaxnreg sits at the top of the branch it governs: after a merge point ptxas assumes the smallest budget)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(EQUSS_REGS_MISC));
#endif
  if (warp == kProducerWarp) {
    // ===================================== TMA producer =====================================
    {   // the whole warp walks the loop (uniform values -> uniform registers), one elected lane issues the copy
      UnitIter it; it.init(u0, (int)p.n_tiles, p.nchunks, G, NCHW ? p.tiles_per_image : 0);
      for (int i = 0; i < n_units; ++i, it.next()) {
        const int tile = it.tile, m = it.m();
        const int s = i % STAGES;
        if (EQUSS_PROD_SLEEP_NS > 0) mbar_wait_sleep(raw_empty + s, ((i / STAGES) & 1) ^ 1, EQUSS_PROD_SLEEP_NS);
        else mbar_wait_nc(raw_empty + s, ((i / STAGES) & 1) ^ 1, 10 + s);
        if (elect_one()) {
          mbar_expect_tx(raw_full + s, RAW_BYTES);
          if (!NCHW) {
            tma_load_2d(s_rawt + s * RAW_BYTES, &tmap, m * D, tile * kTileM, raw_full + s);
          } else {
            tma_load_3d(s_rawt + s * RAW_BYTES, &tmap, it.timg * kTileM, m * D, it.img, raw_full + s);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
    // ===================================== MMA issuers ======================================
    // Everything the tcgen05 instructions consume is made provably warp-uniform (the warp index and the TMEM base go
    // through a shuffle): ptxas then keeps the descriptors in uniform registers and issues UTCHMMA / UTCBAR directly.
    // With per-thread registers it wraps every one of them in an ELECT + 5 x R2UR.BROADCAST loop, ~130 cycles of
    // dependent latency per instruction -- the issuer warps, not the tensor pipe, then set the pace (clock64 traces:
    // ~1000 cycles to issue the 4 MMAs and 2 commits of a unit).
    const int h = warp - kMmaWarp;
    const uint32_t tmem_base_u = tmem_base;
    if (h < HALVES) {
      const uint32_t b_hi = (uint32_t)((SBO >> 4) & 0x3FFF) | (1u << 14);              // SBO, descriptor version 1
      const uint32_t a_hi = (uint32_t)((ASBO >> 4) & 0x3FFF) | (1u << 14);
      const uint32_t b_lo0 = ((smem_u32(s_b) + (uint32_t)(h * (NH / 8) * SBO)) >> 4) | ((uint32_t)(128 >> 4) << 16);
      const uint32_t a_lo0 = (smem_u32(s_a) >> 4) | ((uint32_t)(ALBO >> 4) << 16);
      int b_loads = 0, cur_slot = -1;
      UnitIter it; it.init(u0, (int)p.n_tiles, p.nchunks, G, NCHW ? p.tiles_per_image : 0);
      for (int i = 0; i < n_units; ++i, it.next()) {
        const int a = i % ABUFS, t = i & 1;
        const int tb = (i % kTSlots) * 2 + h;                      // this unit's accumulator-barrier slot
        if (h == 0) EQUSS_TR(7, i);
        if (it.sslot != cur_slot) {
          mbar_wait_nc(b_full, b_loads & 1, 20);
          ++b_loads;
          cur_slot = it.sslot;
        }
        mbar_wait_nc(a_full + a, (i / ABUFS) & 1, 21);
        // TMEM buffer t was last used by unit i - 2: wait until that unit's epilogue has drained this half
        if (i >= 2) mbar_wait_nc(t_empty + ((i - 2) % kTSlots) * 2 + h, ((i - 2) / kTSlots) & 1, 22);
        tc_fence_after();
        if (h == 0) EQUSS_TR(3, i);
        const uint32_t a_lo = a_lo0 + (uint32_t)(a * (A_BYTES >> 4));
        const uint32_t b_lo = b_lo0 + (uint32_t)(it.g * (B_BYTES >> 4));
        const uint32_t d_addr = tmem_base_u + (uint32_t)(t * NC + h * NH);
        if (elect_one()) {
          uint32_t acc = 0;
          // x1.y1, x2.y1, x1.y2 : K-slice kk of a piece starts 2*kk chunks into its region
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            const int a_off = (part == 1) ? C8 : 0;
            const int b_off = (part == 2) ? C8 : 0;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
              umma_f16(d_addr, desc_from(a_lo + (uint32_t)((a_off + 2 * kk) * (ALBO >> 4)), a_hi),
                       desc_from(b_lo + (uint32_t)((b_off + 2 * kk) * 8), b_hi), IDESC, acc);
              acc = 1;
            }
          }
          umma_f16(d_addr, desc_from(a_lo + (uint32_t)(2 * C8 * (ALBO >> 4)), a_hi),
                   desc_from(b_lo + (uint32_t)(2 * C8 * 8), b_hi), IDESC, 1);
          umma_commit(t_full + tb);
          umma_commit(a_empty + a);
        }
        __syncwarp();
      }
    }
  }
  } else if (warp >= kConvWarp0) {
    // ===================================== convert warps ===================================
    // kConvGroups groups of four warps; group c owns the units i with i % kConvGroups == c (conversion AND, fused,
    // the gather of the same unit), so that two units are in conversion at any time: one warp per SMSP cannot hide
    // the latency of its own shared-memory / MUFU / conversion chain.
#ifdef EQUSS_SETMAXNREG
    // (24 warps are launched with 80 registers: below that the limit is lowered, above it raised)
    if constexpr (RegPlan<D>::conv < 80) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(RegPlan<D>::conv));
    else if constexpr (RegPlan<D>::conv > 80) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(RegPlan<D>::conv));
#endif
    const int cgroup = (warp - kConvWarp0) >> 2;
    const int ct = (threadIdx.x - kConvWarp0 * 32) & 127;   // 0..127 within the group
    UnitIter it; it.init(u0, (int)p.n_tiles, p.nchunks, G, NCHW ? p.tiles_per_image : 0);
    it.advance(cgroup);
    for (int i = cgroup; i < n_units; i += kConvGroups, it.advance(kConvGroups)) {    // this group's units only
      const int a = i % ABUFS, s = i % STAGES;
      // the operand images change with the first unit of a slot: only that unit's owner loads them
      const bool new_slot = (i == 0) || (it.tile == 0 && it.g == 0);
      mbar_wait_nc(a_empty + a, ((i / ABUFS) & 1) ^ 1, 30);
      if (warp == kConvWarp0) EQUSS_TR(0, i);
      if (new_slot) {
        // every earlier MMA must have completed before the operand images are overwritten
#pragma unroll
        for (int back = 1; back < ABUFS; ++back)
          if (i >= back) mbar_wait_nc(a_empty + ((i - back) % ABUFS), ((i - back) / ABUFS) & 1, 31);
        if (ct == 0) {
          const bool tab = FUSE && GTAB > 0 && p.use_gtab;
          const uint32_t tab_bytes = (uint32_t)p.K * D * 4;       // per subspace; K <= NC
          mbar_expect_tx(b_full, (uint32_t)(G * p.img_bytes) + (tab ? (uint32_t)G * tab_bytes : 0u));
          bulk_load_1d(s_b, p.images + (size_t)it.sslot * G * p.img_bytes, (uint32_t)(G * p.img_bytes), b_full);
          if (tab) {
#pragma unroll
            for (int g = 0; g < G; ++g)
              bulk_load_1d(s_gt + ((it.sslot & 1) * G + g) * (NC * D), p.gsrc + (size_t)(it.sg * G + g) * p.K * D, tab_bytes, b_full);
          }
        }
      }
      mbar_wait_nc(raw_full + s, (i / STAGES) & 1, 33);
      if (warp == kConvWarp0) EQUSS_TR(1, i);
      uint8_t* a_tile = s_a + a * A_BYTES;
      const float* raw = reinterpret_cast<const float*>(s_rawt + s * RAW_BYTES);
      if (!NCHW) {
        // flat: raw[row][D]; LPS lanes per row, one float4 each.  All loads of a batch of passes are issued before
        // any store (the compiler cannot prove that the A tile and the raw tile do not alias, so a store between
        // two loads would serialise the passes into one long dependent chain).
        constexpr int ROWS_PER_PASS = 128 / LPS;
        constexpr int PASSES = kTileM / ROWS_PER_PASS;
        constexpr int BATCH = PASSES < 8 ? PASSES : 8;
        const int l = ct % LPS;
        const int row0 = ct / LPS;
#pragma unroll 1
        for (int pb = 0; pb < PASSES; pb += BATCH) {
          float4 v[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; ++u)
            v[u] = *reinterpret_cast<const float4*>(raw + ((pb + u) * ROWS_PER_PASS + row0) * D + l * 4);
          float ss[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; ++u) ss[u] = group_sumsq(v[u].x, v[u].y, v[u].z, v[u].w);
#pragma unroll
          for (int sft = 1; sft < LPS; sft <<= 1) {
#pragma unroll
            for (int u = 0; u < BATCH; ++u) ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], sft);
          }
          uint2 hi[BATCH], lo[BATCH];
          float4 zn4[CFUSE ? BATCH : 1];
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            float zx, zy, zz, zw;
            if constexpr (CFUSE) {      // canonical z_norm (the gather needs it bit-exact); kept in the raw stage
#ifdef EQUSS_FUSE_FASTNORM
              const float inv_ = 1.f / l2_denom_fast(ss[u]);
              const float4 zn = make_float4(v[u].x * inv_, v[u].y * inv_, v[u].z * inv_, v[u].w * inv_);
#else
              const float4 zn = div4_fast(v[u], l2_denom_fast(ss[u]));
#endif
              zx = zn.x; zy = zn.y; zz = zn.z; zw = zn.w;
              zn4[u] = zn;
            } else {
              const float inv = fminf(rsqrtf(ss[u]), 1e12f);
              zx = v[u].x * inv; zy = v[u].y * inv; zz = v[u].z * inv; zw = v[u].w * inv;
            }
            const __half2 h01 = __floats2half2_rn(zx, zy), h23 = __floats2half2_rn(zz, zw);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(zx - f01.x, zy - f01.y), l23 = __floats2half2_rn(zz - f23.x, zw - f23.y);
            hi[u].x = *reinterpret_cast<const uint32_t*>(&h01); hi[u].y = *reinterpret_cast<const uint32_t*>(&h23);
            lo[u].x = *reinterpret_cast<const uint32_t*>(&l01); lo[u].y = *reinterpret_cast<const uint32_t*>(&l23);
          }
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            const int row = (pb + u) * ROWS_PER_PASS + row0;
            uint8_t* rowp = a_tile + (row / 8) * ASBO + (row % 8) * 16 + (l >> 1) * ALBO + (l & 1) * 8;
            *reinterpret_cast<uint2*>(rowp) = hi[u];
            *reinterpret_cast<uint2*>(rowp + C8 * ALBO) = lo[u];
            if constexpr (CFUSE)
              *reinterpret_cast<float4*>(const_cast<float*>(raw) + row * D + l * 4) = zn4[u];
          }
        }
      } else {
        // NCHW: raw[channel][128 pixels]; one thread per pixel row
        const int row = ct;
        float x[D];
#pragma unroll
        for (int j = 0; j < D; ++j) x[j] = raw[j * kTileM + row];
        float inv;
        if constexpr (CFUSE) {        // canonical z_norm, written back into the raw stage for the gather
          float gsum[LPS];
#pragma unroll
          for (int g4 = 0; g4 < LPS; ++g4) gsum[g4] = group_sumsq(x[4 * g4], x[4 * g4 + 1], x[4 * g4 + 2], x[4 * g4 + 3]);
          const float denom = l2_denom_fast(butterfly_array<LPS>(gsum));
#pragma unroll
          for (int g4 = 0; g4 < LPS; ++g4) {
            const float4 zn = div4_fast(make_float4(x[4 * g4], x[4 * g4 + 1], x[4 * g4 + 2], x[4 * g4 + 3]), denom);
            x[4 * g4] = zn.x; x[4 * g4 + 1] = zn.y; x[4 * g4 + 2] = zn.z; x[4 * g4 + 3] = zn.w;
          }
#pragma unroll
          for (int j = 0; j < D; ++j) const_cast<float*>(raw)[j * kTileM + row] = x[j];
          inv = 1.f;
        } else {
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < D; ++j) ss = fmaf(x[j], x[j], ss);
          inv = fminf(rsqrtf(ss), 1e12f);
        }
        uint8_t* rowp = a_tile + (row / 8) * ASBO + (row % 8) * 16;
#pragma unroll
        for (int c = 0; c < C8; ++c) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z0 = x[8 * c + 2 * e] * inv, z1 = x[8 * c + 2 * e + 1] * inv;
            const __half2 hh = __floats2half2_rn(z0, z1);
            const float2 ff = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(z0 - ff.x, z1 - ff.y);
            hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          *reinterpret_cast<uint4*>(rowp + c * ALBO) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(rowp + (C8 + c) * ALBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async();     // generic-proxy writes of the A tile -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) { if (!FUSE) mbar_arrive(raw_empty + s); mbar_arrive(a_full + a); }
      if (warp == kConvWarp0) EQUSS_TR(2, i);
    }
  } else if (warp < 4 * kEpiGroups) {
    // ===================================== epilogue warps ===================================
#ifdef EQUSS_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(RegPlan<D>::epi));
#endif
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int egroup = (warp - kEpiWarp0) >> 2;  // 0 .. kEpiGroups-1
    const int row = q * 32 + lane;
    UnitIter it; it.init(u0, (int)p.n_tiles, p.nchunks, G, NCHW ? p.tiles_per_image : 0);
    it.advance(egroup);
    float ge_acc[G];                              // FUSE (epilogue gather): this thread's squared error per subspace of the group
#pragma unroll
    for (int g = 0; g < G; ++g) ge_acc[g] = 0.f;
    int ge_sg = -1;
    auto ge_flush = [&]() {
      if (ge_sg < 0) return;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float t = warp_sum(ge_acc[g]);
        if (lane == 0 && t != 0.f) atomicAdd(p.sqerr + ge_sg * G + g, (double)t);
        ge_acc[g] = 0.f;
      }
    };
    for (int i = egroup; i < n_units; i += kEpiGroups, it.advance(kEpiGroups)) {   // this group's units only
      const int t = i & 1;
      const int tile = it.tile, m = it.m(), chunk = it.chunk;
      // the slot's tolerance: trailer of the operand image (global memory; latency hidden by the barrier wait)
      const float tol = __ldg(reinterpret_cast<const float*>(p.images + ((size_t)it.sslot * G + it.g) * p.img_bytes +
                                                               (size_t)(NC / 8) * SBO));
      // Streaming part: class maxima and one maximum per 16-column group, all in registers (the chunk loops are fully
      // unrolled so that the group maxima have compile-time indices); the winner and the runner-up come out of two
      // tournaments below -- no per-group bookkeeping (5 ALU operations per group before) inside the stream.
      float cls[16], gm[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) { cls[r] = -INFINITY; gm[r] = -INFINITY; }
#pragma unroll
      for (int h = 0; h < HALVES; ++h) {
        const int tb = (i % kTSlots) * 2 + h;
        if (EQUSS_EPI_SLEEP_NS > 0 && h == 0) mbar_wait_sleep(t_full + tb, (i / kTSlots) & 1, EQUSS_EPI_SLEEP_NS);
        else mbar_wait_nc(t_full + tb, (i / kTSlots) & 1, 41);
        tc_fence_after();
        if (q == 0 && h == 0) EQUSS_TR(4, i);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NC + h * NH);
        if constexpr (kEpiGroups >= 3 || NH == 32) {
          // three groups: the other groups' warps cover the TMEM load latency, one 32-register buffer is enough
          uint32_t va[32];
#pragma unroll
          for (int c = 0; c < NH / 32; ++c) {
            tmem_ld32(taddr + c * 32, va);
            tmem_ld_wait();
            epi_chunk(va, cls, gm[h * (NH / 16) + 2 * c], gm[h * (NH / 16) + 2 * c + 1]);
          }
        } else {
          uint32_t va[32], vb[32];
          tmem_ld32(taddr, va);
#pragma unroll
          for (int c = 0; c < NH / 32; c += 2) {
            tmem_ld_wait();
            tmem_ld32(taddr + (c + 1) * 32, vb);
            epi_chunk(va, cls, gm[h * (NH / 16) + 2 * c], gm[h * (NH / 16) + 2 * c + 1]);
            tmem_ld_wait();
            if (c + 2 < NH / 32) tmem_ld32(taddr + (c + 2) * 32, va);
            epi_chunk(vb, cls, gm[h * (NH / 16) + 2 * c + 2], gm[h * (NH / 16) + 2 * c + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + tb);
      }

      if (q == 0) EQUSS_TR(5, i);
      // winning column = best group * 16 + best class; the runner-up is the better of the two tournaments' runners-up
      // (a column other than the winner differs from it in its class or in its group).  On exactly equal maxima the
      // runner-up equals m1, the gap is zero and the row is ambiguous, as it must be.
      int r1, g1;
      float m1, run_c, run_g;
      tournament16(cls, r1, m1, run_c);
      tournament16(gm, g1, m1, run_g);
      int best_col = g1 * 16 + r1;
      const float runner = fmaxf(run_c, run_g);
      long long n;
      bool live;
      if (!NCHW) {
        n = (long long)tile * kTileM + row;
        live = n < p.n_pixels;
      } else {
        const int sidx = it.timg * kTileM + row;
        live = sidx < p.hw;
        n = (long long)it.img * p.hw + sidx;
      }
      const int kvalid = min(NC, p.K - chunk * NC);
      const bool amb = live && (!(m1 - runner > tol) || best_col >= kvalid);
      if (amb) {
        // Ambiguous rows (~3e-4 of them) keep a provisional winner here and are listed for an exact fp32 scan by
        // rescan_flagged_kernel after this kernel (which also repairs the fused gather of rows whose winner changes).
        // Re-scoring them inline stalled the pipeline: one warp busy with dependent global loads for several
        // microseconds holds back its TMEM quarter, the accumulator ring is two deep, and the MMA issuers wait
        // (clock64 traces: ~45 % of the d = 64 kernel's time, scripts/trace_assign.cu).
        if (best_col >= kvalid) best_col = 0;
        p.flag_list[atomicAdd(p.flag_count, 1u)] = (uint32_t)((long long)m * p.n_pixels + n);
      }
      if constexpr (FUSE) {
        // K3 for this unit (model/quantizer.py:474,514,534-536): the raw stage still holds the canonical z_norm the
        // convert pass left there.  The codeword comes from the slot's shared-memory table (d = 16, long slots) or from
        // global memory.  The table was filled by the bulk copies that completed b_full before this unit's MMAs were
        // issued, i.e. long before t_full.
        const int s = i % STAGES;
        const float* raw = reinterpret_cast<const float*>(s_rawt + s * RAW_BYTES);
        const bool tab = GTAB > 0 && p.use_gtab;
        const float* tbl = s_gt + ((it.sslot & 1) * G + it.g) * (NC * D);
        const float* gsrc_m = p.gsrc + (size_t)m * p.K * D;
        const int col = live ? best_col : 0;
        if (it.sg != ge_sg) { ge_flush(); ge_sg = it.sg; }
        float e = 0.f;
        if constexpr (NCHW) {
          // one thread per pixel row (channel-major tile: lanes read / write consecutive pixels), 16 channels at a time
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 16) {
            float4 qv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              qv[u] = tab ? *reinterpret_cast<const float4*>(tbl + col * D + c0 + 4 * u)
                          : __ldg(reinterpret_cast<const float4*>(gsrc_m + (size_t)col * D + c0) + u);
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = raw[(c0 + j) * kTileM + row];
            float ov[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float d0 = qv[u].x - x[4 * u], d1 = qv[u].y - x[4 * u + 1], d2 = qv[u].z - x[4 * u + 2], d3 = qv[u].w - x[4 * u + 3];
              ov[4 * u] = x[4 * u] + d0; ov[4 * u + 1] = x[4 * u + 1] + d1;             // STE value (:536)
              ov[4 * u + 2] = x[4 * u + 2] + d2; ov[4 * u + 3] = x[4 * u + 3] + d3;
              e += group_sumsq(d0, d1, d2, d3);
            }
            if (live) {
              float* o = p.out + (long long)it.img * p.zv.stride_b + (long long)(m * D + c0) * p.zv.stride_c + (it.timg * kTileM + row);
              const uint32_t sc = (uint32_t)p.zv.stride_c;
#pragma unroll
              for (int j = 0; j < 16; ++j) __stcs(o + (size_t)((uint32_t)j * sc), ov[j]);
            }
          }
          e = live ? e : 0.f;
        } else {
          // row-major tile: LPS lanes share a row (one float4 each), 32 / LPS rows per pass -- shared-memory reads
          // without bank conflicts and 16-byte stores that fill whole sectors; the rows' winners travel by shuffle
          constexpr int RPP = 32 / LPS;
          const int qd = lane % LPS, rsub = lane / LPS;
#pragma unroll
          for (int pass = 0; pass < LPS; ++pass) {
            const int rr = pass * RPP + rsub;                          // row within the warp's 32
            const int col_r = __shfl_sync(0xffffffffu, col, rr);
            const long long n_r = (long long)tile * kTileM + q * 32 + rr;
            const float4 zn = *reinterpret_cast<const float4*>(raw + (q * 32 + rr) * D + qd * 4);
            const float4 qq = tab ? *reinterpret_cast<const float4*>(tbl + col_r * D + qd * 4)
                                  : __ldg(reinterpret_cast<const float4*>(gsrc_m + (size_t)col_r * D) + qd);
            const float d0 = qq.x - zn.x, d1 = qq.y - zn.y, d2 = qq.z - zn.z, d3 = qq.w - zn.w;
            if (n_r < p.n_pixels) {
              __stcs(reinterpret_cast<float4*>(p.out + n_r * p.zv.stride_s + m * D) + qd,
                     make_float4(zn.x + d0, zn.y + d1, zn.z + d2, zn.w + d3));              // STE value (:536)
              e += group_sumsq(d0, d1, d2, d3);
            }
          }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) ge_acc[g] += (it.g == g) ? e : 0.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(raw_empty + s);      // the stage's last reader is done
      }
      if (live) {
        if (p.merged == nullptr) {
          p.idx_out[(long long)m * p.n_pixels + n] = best_col;
        } else {
          const long long o = (long long)m * p.n_pixels + n;
          const unsigned long long mine = ((unsigned long long)sortable(m1) << 32) | (unsigned long long)(uint32_t)(chunk * NC + best_col);
          const unsigned long long old = atomicMax(p.merged + o, mine);
          if (old != 0ull) {
            const float so = unsortable((unsigned int)(old >> 32));
            if (!(fabsf(so - m1) > 2.f * tol)) p.flag_list[atomicAdd(p.flag_count, 1u)] = (uint32_t)o;
          }
        }
      }
      if (q == 0) { EQUSS_TR(6, i); if (!FUSE) EQUSS_TRG(8, i); }
    }
    if constexpr (FUSE) ge_flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int D, int NC, int G, int STAGES, int ABUFS, bool NCHW, bool FUSE, int LAG>
static int launch_instance(const CUtensorMap& tmap, const Params& p, int grid, cudaStream_t st) {
  constexpr int SMEM = smem_bytes(D, NC, G, STAGES, ABUFS, FUSE, NCHW) < 120 * 1024 ? 120 * 1024 : smem_bytes(D, NC, G, STAGES, ABUFS, FUSE, NCHW);
  static_assert(SMEM <= 227 * 1024, "shared-memory plan exceeds 227 KB");
  EQUSS_CUDA_OK(cudaFuncSetAttribute(assign_f16x2_kernel<D, NC, G, STAGES, ABUFS, NCHW, FUSE, LAG>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  EQUSS_CUDA_OK(cudaLaunchKernelEx(&cfg, assign_f16x2_kernel<D, NC, G, STAGES, ABUFS, NCHW, FUSE, LAG>, tmap, p));
  EQUSS_LAUNCH_OK("assign_f16x2_kernel");
  return EQUSS_OK;
}

// per-d dispatch over (NC, G, layout, fused gather); one translation unit per d keeps the build parallel.  GV =
// subspaces per 128-byte line of a flat row (used when M is a multiple of it; NCHW and odd M run with G = 1).
// STF: raw-ring depth of the fused kernels (STF = 0: no fused instantiation for this d); LAGV: unused (kept in the
// instantiation names the profiles refer to).
#define EQUSS_TCH_DISPATCH(DV, GV, STV, ABV, STF, LAGV)                                                       \
  int launch_tch_d##DV(int NC, int G, bool nchw, bool fuse, const CUtensorMap& tmap, const Params& p, int grid, \
                       cudaStream_t st) {                                                                    \
    EQUSS_TCH_ONE(DV, 32, GV, STV, ABV, STF, LAGV) EQUSS_TCH_ONE(DV, 256, GV, STV, ABV, STF, LAGV)            \
    set_error("tcgen05 f16x2 assign: no instantiation for d=%d NC=%d", DV, NC);                               \
    return EQUSS_ERR_UNSUPPORTED;                                                                             \
  }
#define EQUSS_TCH_ONE(DV, NCV, GV, STV, ABV, STF, LAGV)                                                       \
  if (NC == NCV) {                                                                                            \
    if constexpr (STF > 0) {                                                                                  \
      if (fuse) {                                                                                             \
        if (nchw) return launch_instance<DV, NCV, 1, (STF > 0 ? STF : 2), ABV, true, true, LAGV>(tmap, p, grid, st);     \
        if (G == GV) return launch_instance<DV, NCV, GV, (STF > 0 ? STF : 2), ABV, false, true, LAGV>(tmap, p, grid, st); \
        return launch_instance<DV, NCV, 1, (STF > 0 ? STF : 2), ABV, false, true, LAGV>(tmap, p, grid, st);              \
      }                                                                                                       \
    }                                                                                                         \
    if (fuse) { set_error("tcgen05 f16x2 assign: fused gather is not built for d=%d", DV); return EQUSS_ERR_UNSUPPORTED; } \
    if (nchw) return launch_instance<DV, NCV, 1, STV, ABV, true, false, 1>(tmap, p, grid, st);                \
    if (G == GV) return launch_instance<DV, NCV, GV, STV, ABV, false, false, 1>(tmap, p, grid, st);           \
    return launch_instance<DV, NCV, 1, STV, ABV, false, false, 1>(tmap, p, grid, st);                         \
  }

}  // namespace tch
}  // namespace equss
