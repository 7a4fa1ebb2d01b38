// knn_tc.cu -- similarity GEMM of the global-feature kNN (K11, data/precompute_knns.py:313) on the tensor cores:
//     S[r][c] = <Q[r], DB[c]>,   Q: [rows][F], DB: [n][F] fp32 row-major (both K-major operands)
// as a tcgen05 / TMEM GEMM with 128 x 256 output tiles, split-tf32 (hi.hi + lo.hi + hi.lo) for fp32-level accuracy.
//
// Both operands arrive by TMA (boxes of 32 features x 128 / 256 rows, 128-byte rows, SWIZZLE_128B) directly in the
// canonical K-major UMMA layout; the raw tiles ARE the hi operands (the tensor core ignores the low 13 mantissa bits
// of an fp32 word), eight convert warps write lo = x - trunc(x) element-wise at identical (swizzled) offsets.
// The hi.hi products (issuer warp 0) and the small lo.hi / hi.lo products (issuer warp 1) use separate TMEM
// accumulators: the tensor core's fp32 accumulate truncates, and keeping the 2^-11-sized terms out of the main
// accumulator keeps that drift at the level of the main terms alone; the epilogue adds the two with a
// round-to-nearest add and streams the tile to the similarity workspace.
// Tiles are walked column-major (all row tiles of one database tile before the next), so at any time the 148 CTAs
// share ~3 database tiles in L2 and the query chunk stays L2-resident.
#include <cuda.h>
#include <cstring>
#include <cstdlib>
#include "equss_common.cuh"
#include "equss_tcgen05.cuh"

namespace equss {
namespace knntc {

using namespace ::equss::ptx;

constexpr int kBM = 128, kBN = 256, kKC = 32;        // tile rows / columns, features per stage
constexpr int kStages = 2;
constexpr int kCand = 128;        // fused top-k: candidate slots per row; compacted to the k best when fewer than kCandReserve are free
constexpr int kCandReserve = 16;  // columns examined between two occupancy checks
constexpr int kThreads = 32 * (4 + 1 + 2 + 8);       // 4 epilogue, producer, 2 MMA issuers, 8 convert warps
constexpr int kEpiWarp0 = 0, kProducerWarp = 4, kMmaWarp = 5, kConvWarp0 = 7;
constexpr int kARaw = kBM * 128, kBRaw = kBN * 128;  // bytes: 16 KB, 32 KB
constexpr int kStageBytes = 2 * kARaw + 2 * kBRaw;   // raw A | lo A | raw B | lo B = 96 KB
constexpr int kSmem = 1024 + kStages * kStageBytes + 256;

__host__ __device__ constexpr uint32_t make_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

struct Params {
  long long rows, n;
  int F, n_kc;
  int m_tiles, n_tiles;
  float* S;
  // fused top-k (TOPK kernels): a CTA owns (row tile, column split) items and keeps the k best of every row of the item
  // in shared memory while it walks the split's column tiles; partial lists go to part_val / part_idx [rows][splits][k]
  int k, splits;
  float* part_val;
  int* part_idx;
  int debug;           // EQUSS_KNN_DEBUG (timing experiments): 1 = epilogue reads TMEM only
  uint2* cand;         // [gridDim.x][kBM][kCand] candidate buffers (value bits, column), L2-resident scratch
};

// Order in which a CTA visits its output tiles; every warp role walks the same sequence.
//   !TOPK: tiles column-major over the whole grid (t = bn * m_tiles + bm), round-robin over the CTAs.
//   TOPK : items (bm, split) round-robin over the CTAs; inside an item the split's column tiles in increasing order.
template <bool TOPK>
struct TileWalk {
  long long t, step, total;       // !TOPK: tile id; TOPK: item id
  int m_tiles, n_tiles, splits;
  int bm, bn, bn_end;
  bool first, last;               // TOPK: first / last tile of the current item
  __device__ __forceinline__ void set_item() {
    const int sp = (int)(t / m_tiles);
    bm = (int)(t - (long long)sp * m_tiles);
    bn = (int)((long long)n_tiles * sp / splits);
    bn_end = (int)((long long)n_tiles * (sp + 1) / splits);
    first = true; last = (bn + 1 >= bn_end);
  }
  __device__ __forceinline__ bool init(const Params& p) {
    m_tiles = p.m_tiles; n_tiles = p.n_tiles; splits = p.splits;
    t = blockIdx.x; step = gridDim.x;
    total = TOPK ? (long long)p.m_tiles * p.splits : (long long)p.m_tiles * p.n_tiles;
    if (t >= total) return false;
    if (TOPK) set_item();
    else { bn = (int)(t / m_tiles); bm = (int)(t - (long long)bn * m_tiles); first = last = true; }
    return true;
  }
  __device__ __forceinline__ bool next() {
    if (TOPK && bn + 1 < bn_end) { ++bn; first = false; last = (bn + 1 >= bn_end); return true; }
    t += step;
    if (t >= total) return false;
    if (TOPK) set_item();
    else { bn = (int)(t / m_tiles); bm = (int)(t - (long long)bn * m_tiles); }
    return true;
  }
};

template <bool TOPK>
__global__ void __launch_bounds__(kThreads, 1)
knn_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db, const Params p) {
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
  const int n_kc = p.n_kc;

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

  if (warp == kProducerWarp) {
    {   // the whole warp walks the loop (uniform values -> uniform registers), one elected lane issues the copies
      int g = 0;
      TileWalk<TOPK> tw;
      for (bool ok = tw.init(p); ok; ok = tw.next()) {
        const int bn = tw.bn, bm = tw.bm;
        for (int c = 0; c < n_kc; ++c, ++g) {
          const int st = g % kStages;
          uint8_t* sp = smem + st * kStageBytes;
          mbar_wait(st_empty + st, ((g / kStages) & 1) ^ 1, 10);
          if (elect_one()) {
            mbar_expect_tx(raw_full + st, kARaw + kBRaw);
            tma_load_2d(sp, &tmap_q, c * kKC, bm * kBM, raw_full + st);
            tma_load_2d(sp + 2 * kARaw, &tmap_db, c * kKC, bn * kBN, raw_full + st);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
    // issuer 0: x_hi.w_hi -> main accumulator (columns 0..255); issuer 1: x_lo.w_hi + x_hi.w_lo -> small (256..511)
    const int part = warp - kMmaWarp;
    constexpr uint32_t IDESC = make_idesc_tf32(kBN);
    // K-major SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO); a K step of 8 floats advances 32 B
    const uint32_t d_hi = (uint32_t)((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
    const uint32_t base = smem_u32(smem);
    int g = 0, it = 0;
    TileWalk<TOPK> tw;
    for (bool ok = tw.init(p); ok; ok = tw.next(), ++it) {
      mbar_wait(acc_empty, (it & 1) ^ 1, 22);
      const uint32_t d_addr = tmem_base + (uint32_t)(part * kBN);
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int st = g % kStages;
        if (part == 0) mbar_wait(raw_full + st, (g / kStages) & 1, 20);
        else mbar_wait(lo_full + st, (g / kStages) & 1, 21);          // implies raw_full
        tc_fence_after();
        if (elect_one()) {   // elected lane + uniform operands: UTCHMMA issues from uniform registers
          const uint32_t sa = base + (uint32_t)(st * kStageBytes);
          const uint32_t a_raw = (sa >> 4) | (1u << 16), a_lo = ((sa + kARaw) >> 4) | (1u << 16);
          const uint32_t b_raw = ((sa + 2 * kARaw) >> 4) | (1u << 16), b_lo = ((sa + 2 * kARaw + kBRaw) >> 4) | (1u << 16);
          if (part == 0) {
#pragma unroll
            for (int kk = 0; kk < kKC / 8; ++kk)
              umma_tf32(d_addr, desc_from(a_raw + 2 * kk, d_hi), desc_from(b_raw + 2 * kk, d_hi), IDESC, (c > 0 || kk > 0) ? 1u : 0u);
          } else {
#pragma unroll
            for (int kk = 0; kk < kKC / 8; ++kk)
              umma_tf32(d_addr, desc_from(a_lo + 2 * kk, d_hi), desc_from(b_raw + 2 * kk, d_hi), IDESC, (c > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < kKC / 8; ++kk)
              umma_tf32(d_addr, desc_from(a_raw + 2 * kk, d_hi), desc_from(b_lo + 2 * kk, d_hi), IDESC, 1u);
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
    TileWalk<TOPK> tw;
    for (bool ok = tw.init(p); ok; ok = tw.next()) {
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int st = g % kStages;
        mbar_wait(raw_full + st, (g / kStages) & 1, 31);
        uint8_t* sp = smem + st * kStageBytes;
        // A: 1024 float4 (4 per thread), B: 2048 float4 (8 per thread); lo tile = raw tile + kARaw / + kBRaw
        float4 va[4], vb[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) va[u] = *reinterpret_cast<const float4*>(sp + (u * 256 + ct) * 16);
#pragma unroll
        for (int u = 0; u < 8; ++u) vb[u] = *reinterpret_cast<const float4*>(sp + 2 * kARaw + (u * 256 + ct) * 16);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float4 l;
          l.x = va[u].x - tf32_trunc(va[u].x); l.y = va[u].y - tf32_trunc(va[u].y);
          l.z = va[u].z - tf32_trunc(va[u].z); l.w = va[u].w - tf32_trunc(va[u].w);
          *reinterpret_cast<float4*>(sp + kARaw + (u * 256 + ct) * 16) = l;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
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
    // epilogue warps 0-3: TMEM lane quarter = warp
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    // TOPK: every row of the tile owns kCand candidate slots in an L2-resident scratch buffer.  The thread that owns a
    // row (TMEM lane = row) appends every similarity above the row's threshold `thr` -- two stores and a counter, no
    // list maintenance.  When a row has fewer than kCandReserve free slots, the WARP compacts it: the 128 candidates are
    // ranked cooperatively (4 per lane, all-pairs through shuffles; larger value first, lower column on ties), the k
    // best are written back in order and the k-th becomes the new threshold.  Exact: `thr` is always the k-th best of
    // the columns seen so far or lower, so nothing that belongs to the final top-k is ever dropped.
    const int k = p.k;
    int cnt = 0;
    float thr = -INFINITY;
    uint2* my = TOPK ? p.cand + ((size_t)blockIdx.x * kBM + row) * kCand : nullptr;
    uint2* wbase = TOPK ? p.cand + ((size_t)blockIdx.x * kBM + q * 32) * kCand : nullptr;
    // compaction of the candidates of row (q*32 + src) by the whole warp; returns the new (cnt, thr) to every lane
    auto compact = [&](int src, int n_src, float& thr_out) -> int {
      uint2* buf = wbase + (size_t)src * kCand;
      // 64-bit sort keys: (order-preserving bits of the value) << 32 | ~column  -- larger key = better candidate, keys
      // are distinct, and one integer comparison replaces the (value, column) pair logic (no divergent branches)
      unsigned long long key[4];
      int rank[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int e = lane + 32 * t;
        uint2 u = make_uint2(0u, 0u);
        if (e < n_src) u = __ldcg(buf + e);
        const uint32_t sv = (u.x & 0x80000000u) ? ~u.x : (u.x | 0x80000000u);
        key[t] = (e < n_src) ? (((unsigned long long)sv << 32) | (unsigned long long)(~u.y)) : 0ull;
        rank[t] = 0;
      }
#pragma unroll
      for (int t2 = 0; t2 < 4; ++t2) {
        if (t2 * 32 >= n_src) break;                                                        // warp-uniform
#pragma unroll 8
        for (int l2 = 0; l2 < 32; ++l2) {
          const unsigned long long ok = __shfl_sync(0xffffffffu, key[t2], l2);
#pragma unroll
          for (int t = 0; t < 4; ++t) rank[t] += (ok > key[t]) ? 1 : 0;
        }
      }
      __syncwarp();
      const int keep = n_src < k ? n_src : k;
      float kth = -INFINITY;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t sv = (uint32_t)(key[t] >> 32);
        const uint32_t vb = (sv & 0x80000000u) ? (sv & 0x7fffffffu) : ~sv;
        if (lane + 32 * t < n_src && rank[t] < keep) __stcg(buf + rank[t], make_uint2(vb, ~(uint32_t)key[t]));
        if (rank[t] == k - 1 && lane + 32 * t < n_src) kth = __uint_as_float(vb);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, o));
      __syncwarp();
      thr_out = kth;                      // -inf while the row holds fewer than k candidates
      return keep;
    };
    auto compact_rows = [&](bool mine) {
      unsigned todo = __ballot_sync(0xffffffffu, mine);
      if (todo) __syncwarp();             // the owners' appends are visible to the lanes that are about to read them
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int n_src = __shfl_sync(0xffffffffu, cnt, src);
        float nthr;
        const int ncnt = compact(src, n_src, nthr);
        if (lane == src) { cnt = ncnt; thr = nthr; }
      }
    };
    TileWalk<TOPK> tw;
    for (bool ok = tw.init(p); ok; ok = tw.next(), ++it) {
      const int bn = tw.bn, bm = tw.bm;
      const long long r = (long long)bm * kBM + row;
      const long long c0 = (long long)bn * kBN;
      if (TOPK && tw.first) { cnt = 0; thr = -INFINITY; }
      mbar_wait(acc_full, it & 1, 40);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < kBN / 32; ++ch) {
        uint32_t vm[32], vs[32];
        tmem_ld32(lane_base + (uint32_t)(ch * 32), vm);
        tmem_ld32(lane_base + (uint32_t)(kBN + ch * 32), vs);
        tmem_ld_wait();
        if constexpr (TOPK) {
          const long long cb = c0 + ch * 32;
          const int nv = (r < p.rows && cb < p.n && !(p.debug & 1)) ? (int)((p.n - cb < 32) ? (p.n - cb) : 32) : 0;
#pragma unroll
          for (int half = 0; half < 32 / kCandReserve; ++half) {
#pragma unroll
            for (int jj = 0; jj < kCandReserve; ++jj) {
              const int j = half * kCandReserve + jj;
              const float v = __uint_as_float(vm[j]) + __uint_as_float(vs[j]);
              if (j < nv && v > thr) { __stcg(my + cnt, make_uint2(__float_as_uint(v), (uint32_t)(cb + j))); ++cnt; }
            }
            compact_rows(cnt > kCand - kCandReserve);
          }
        } else {
        if (r < p.rows) {
          float* o = p.S + r * p.n + c0 + ch * 32;
          if (c0 + ch * 32 + 32 <= p.n && (p.n & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              __stcs(reinterpret_cast<float4*>(o + j),
                     make_float4(__uint_as_float(vm[j]) + __uint_as_float(vs[j]), __uint_as_float(vm[j + 1]) + __uint_as_float(vs[j + 1]),
                                 __uint_as_float(vm[j + 2]) + __uint_as_float(vs[j + 2]), __uint_as_float(vm[j + 3]) + __uint_as_float(vs[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + ch * 32 + j < p.n) o[j] = __uint_as_float(vm[j]) + __uint_as_float(vs[j]);
          }
        }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      if (TOPK && tw.last) {
        // final compaction of every row, then the (at most k, ordered) survivors of (row, split) to the partial lists
        compact_rows(cnt > 0);
        if (r < p.rows) {
          const int sp = (int)(tw.t / p.m_tiles);
          float* ov = p.part_val + ((long long)r * p.splits + sp) * k;
          int* oi = p.part_idx + ((long long)r * p.splits + sp) * k;
          for (int e = 0; e < k; ++e) {
            const uint2 u = (e < cnt) ? __ldcg(my + e) : make_uint2(0xff800000u, 0xffffffffu);
            ov[e] = __uint_as_float(u.x);
            oi[e] = (int)u.y;
          }
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

}  // namespace knntc

bool knn_gemm_tc_supported(const float* Q, const float* DB, const float* S, long long n, int F) {
  (void)n;
  return F > 0 && F % knntc::kKC == 0 && !((uintptr_t)Q & 15) && !((uintptr_t)DB & 15) && !((uintptr_t)S & 15);
}

static int make_maps(const float* Q, const float* DB, long long rows, long long n, int F, CUtensorMap* tq, CUtensorMap* td) {
  using namespace knntc;
  PFN_encodeTiled encode = get_encode_fn();
  EQUSS_REQUIRE(encode != nullptr, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  auto make = [&](CUtensorMap* tm, const float* base, long long nrows, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)nrows};
    cuuint64_t gstr[1] = {(cuuint64_t)F * 4};
    cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult c1 = make(tq, Q, rows, kBM), c2 = make(td, DB, n, kBN);
  EQUSS_REQUIRE(c1 == CUDA_SUCCESS && c2 == CUDA_SUCCESS, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d, %d)", (int)c1, (int)c2);
  return EQUSS_OK;
}

int knn_gemm_tc_launch(const float* Q, const float* DB, float* S, long long rows, long long n, int F, cudaStream_t st) {
  using namespace knntc;
  CUtensorMap tq, td;
  int rc = make_maps(Q, DB, rows, n, F, &tq, &td);
  if (rc != EQUSS_OK) return rc;
  Params p;
  memset(&p, 0, sizeof(p));
  p.rows = rows; p.n = n; p.F = F; p.n_kc = F / kKC;
  p.m_tiles = (int)((rows + kBM - 1) / kBM);
  p.n_tiles = (int)((n + kBN - 1) / kBN);
  p.S = S; p.splits = 1;
  const long long total = (long long)p.m_tiles * p.n_tiles;
  int grid = num_sms();
  if (total < grid) grid = (int)total;
  EQUSS_CUDA_OK(cudaFuncSetAttribute(knn_gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  knn_gemm_tc_kernel<false><<<grid, kThreads, kSmem, st>>>(tq, td, p);
  EQUSS_LAUNCH_OK("knn_gemm_tc_kernel");
  return EQUSS_OK;
}

// Column splits of the fused top-k kernel: (row tiles x splits) items over the SMs with the least idle tail.
int knn_topk_tc_splits(long long rows, long long n) {
  using namespace knntc;
  const long long m_tiles = (rows + kBM - 1) / kBM, n_tiles = (n + kBN - 1) / kBN;
  const int sms = num_sms();
  if (getenv("EQUSS_KNN_SPLITS")) return atoi(getenv("EQUSS_KNN_SPLITS"));
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 8 && s <= n_tiles; ++s) {
    const long long items = m_tiles * s;
    const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

// GEMM with the running top-k fused into the epilogue: partial lists [rows][splits][k] (value, index); the similarity
// matrix is never written.
int64_t knn_topk_tc_scratch_bytes() { return (int64_t)num_sms() * knntc::kBM * knntc::kCand * 8; }

int knn_topk_tc_launch(const float* Q, const float* DB, long long rows, long long n, int F, int k, int splits,
                       float* part_val, int* part_idx, void* scratch, cudaStream_t st) {
  using namespace knntc;
  CUtensorMap tq, td;
  int rc = make_maps(Q, DB, rows, n, F, &tq, &td);
  if (rc != EQUSS_OK) return rc;
  Params p;
  memset(&p, 0, sizeof(p));
  p.rows = rows; p.n = n; p.F = F; p.n_kc = F / kKC;
  p.m_tiles = (int)((rows + kBM - 1) / kBM);
  p.n_tiles = (int)((n + kBN - 1) / kBN);
  p.k = k; p.splits = splits; p.part_val = part_val; p.part_idx = part_idx; p.cand = (uint2*)scratch;
  p.debug = getenv("EQUSS_KNN_DEBUG") ? atoi(getenv("EQUSS_KNN_DEBUG")) : 0;
  const long long items = (long long)p.m_tiles * splits;
  int grid = num_sms();
  if (items < grid) grid = (int)items;
  EQUSS_CUDA_OK(cudaFuncSetAttribute(knn_gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  knn_gemm_tc_kernel<true><<<grid, kThreads, kSmem, st>>>(tq, td, p);
  EQUSS_LAUNCH_OK("knn_gemm_tc_kernel<topk>");
  return EQUSS_OK;
}

}  // namespace equss
