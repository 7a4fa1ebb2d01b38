// knn_h.cu -- global-feature kNN (K11, data/precompute_knns.py:313-315) as "screen in fp16, decide in fp32":
//
//   1. knn_stats_kernel / knn_to_half_kernel: one pass for max |x| and max row norm of queries and database, one pass that
//      writes fp16 copies scaled by a power of two (max |x| <= 1: no overflow, exact scaling).
//   2. knn_screen_kernel: S~ = Qh . DBh^T on the tensor cores (tcgen05 kind::f16, fp32 accumulate in TMEM, both operands
//      by TMA in the K-major SWIZZLE_128B layout, 128 x 256 tiles, 64 features per stage, TWO accumulator buffers so the
//      MMAs of tile t+1 run under the epilogue of tile t).  No convert warps, no lo tiles, one product instead of the three
//      of the split-tf32 kernel (knn_tc.cu): 48 KB of operand reads per stage instead of 290 KB of shared-memory traffic.
//      The epilogue keeps, per query row, every column whose approximate similarity is within `eps` of the running
//      approximate k-th best (append buffer + warp-cooperative compaction as in knn_tc.cu).  eps bounds the fp16 rounding
//      of the operands: |S~ - S| <= 2^-10 |q| |d| (Cauchy-Schwarz over the element-wise relative errors 2^-11), doubled.
//      Hence every member of the true top-k is among the survivors.
//   3. knn_rescore_kernel: one warp per query: exact fp32 dot products of the survivors (a few dozen rows of the
//      database), exact top-k (larger similarity first, lower index on ties).  A row whose survivor list overflowed
//      (hundreds of database rows within 2 eps of its k-th neighbour) is scanned exhaustively instead.
//
// Results equal an exact fp32 top-k; the similarity matrix is never formed.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstring>
#include <cstdlib>
#include "equss_common.cuh"
#include "equss_tcgen05.cuh"

namespace equss {
namespace knnh {

using namespace ::equss::ptx;

constexpr int kBM = 128, kBN = 256, kKC = 64;          // tile rows / columns, features (fp16, 128 bytes) per stage
constexpr int kCand = 128;                             // candidate slots per row and split
constexpr int kCandReserve = 16;                       // columns examined between two occupancy checks
constexpr int kThreads = 32 * 6;                       // 4 epilogue warps, TMA producer, MMA issuer
constexpr int kProducerWarp = 4, kMmaWarp = 5;
constexpr int kABytes = kBM * 128, kBBytes = kBN * 128;                 // 16 KB, 32 KB

// kind::f16: fp16 A / B (K-major), fp32 accumulate, M = 128, N
__host__ __device__ constexpr uint32_t make_idesc(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

struct Stats {             // device-side, written by knn_stats_kernel (bits of non-negative floats, atomicMax-able)
  unsigned int q_absmax, q_sumsq, d_absmax, d_sumsq;
};
__device__ __forceinline__ float pow2_scale(float absmax) {
  // power of two s with absmax * s in (1/2, 1]; 1 for an all-zero matrix
  if (!(absmax > 0.f)) return 1.f;
  int e;
  frexpf(absmax, &e);                  // absmax = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, -e);
}

struct Params {
  long long rows, n;
  int F, n_kc, m_tiles, n_tiles;
  int k, splits;
  const Stats* stats;
  uint2* cand;             // [gridDim.x][kBM][kCand] working buffers (value bits, column), L2-resident
  uint2* lists;            // [rows][splits][kCand] survivors
  int* counts;             // [rows][splits]: number of survivors, or -1 = overflow (exhaustive scan in the rescore kernel)
};

// per-row max |x| and sum of squares of a [rows][F] matrix (one warp per row)
__global__ void __launch_bounds__(256)
knn_stats_kernel(const float* __restrict__ x, long long rows, int F, unsigned int* __restrict__ absmax, unsigned int* __restrict__ sumsq) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float am = 0.f, sm = 0.f;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* xr = x + r * F;
    float a = 0.f, s = 0.f;
    for (int j = lane; j < F; j += 32) { const float v = xr[j]; a = fmaxf(a, fabsf(v)); s = fmaf(v, v, s); }
    s = warp_sum(s);
    am = fmaxf(am, a); sm = fmaxf(sm, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
  if (lane == 0) { atomicMax(absmax, __float_as_uint(am)); atomicMax(sumsq, __float_as_uint(sm)); }
}

// fp16 copy scaled by the matrix's power of two
__global__ void __launch_bounds__(256)
knn_to_half_kernel(const float* __restrict__ x, long long n4, const unsigned int* __restrict__ absmax, __half* __restrict__ out) {
  const float sc = pow2_scale(__uint_as_float(*absmax));
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
    const __half2 a = __floats2half2_rn(v.x * sc, v.y * sc), b = __floats2half2_rn(v.z * sc, v.w * sc);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

// (group of RT row tiles, column split) items round-robin over the CTAs; inside an item the split's column tiles in
// increasing order
struct TileWalk {
  long long t, step, total;
  int m_items, n_tiles, splits, bmi, bn, bn_end;
  bool first, last;
  __device__ __forceinline__ void set_item() {
    const int sp = (int)(t / m_items);
    bmi = (int)(t - (long long)sp * m_items);
    bn = (int)((long long)n_tiles * sp / splits);
    bn_end = (int)((long long)n_tiles * (sp + 1) / splits);
    first = true; last = (bn + 1 >= bn_end);
  }
  __device__ __forceinline__ bool init(const Params& p, int rt) {
    m_items = (p.m_tiles + rt - 1) / rt; n_tiles = p.n_tiles; splits = p.splits;
    t = blockIdx.x; step = gridDim.x; total = (long long)m_items * p.splits;
    if (t >= total) return false;
    set_item();
    return true;
  }
  __device__ __forceinline__ bool next() {
    if (bn + 1 < bn_end) { ++bn; first = false; last = (bn + 1 >= bn_end); return true; }
    t += step;
    if (t >= total) return false;
    set_item();
    return true;
  }
  __device__ __forceinline__ int split() const { return (int)(t / m_items); }
};

// RT = query tiles of 128 rows that share one database tile per stage: RT = 1 keeps two accumulator buffers (MMAs of the
// next tile under the epilogue of this one); RT = 2 uses both 256-column accumulators for ONE stage's two products and
// moves 64 KB instead of 96 KB of operands per pair of output tiles -- the screen is bound by that stream.
template <int RT> __host__ __device__ constexpr int stage_bytes() { return RT * kABytes + kBBytes; }
template <int RT> __host__ __device__ constexpr int n_stages() { return RT == 1 ? 4 : 3; }
template <int RT> __host__ __device__ constexpr int smem_bytes() { return 1024 + n_stages<RT>() * stage_bytes<RT>() + 256; }

template <int RT>
__global__ void __launch_bounds__(kThreads, 1)
knn_screen_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db, const Params p) {
  constexpr int STAGES = n_stages<RT>(), SB = stage_bytes<RT>(), NB = 2 / RT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // swizzle atoms are 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * SB);
  uint64_t* st_full = bars;                   // [STAGES] all TMA boxes of the stage landed
  uint64_t* st_empty = st_full + STAGES;      // [STAGES] the stage's MMAs completed
  uint64_t* acc_full = st_empty + STAGES;     // [NB] tile (group) finished in TMEM buffer b
  uint64_t* acc_empty = acc_full + 2;         // [NB] the epilogue drained buffer b
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_kc = p.n_kc;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(st_full + i, 1); mbar_init(st_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);

  if (warp == kProducerWarp) {
    int g = 0;
    TileWalk tw;
    for (bool ok = tw.init(p, RT); ok; ok = tw.next()) {
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int st = g % STAGES;
        uint8_t* sp = smem + st * SB;
        mbar_wait(st_empty + st, ((g / STAGES) & 1) ^ 1, 10);
        if (elect_one()) {
          mbar_expect_tx(st_full + st, SB);
#pragma unroll
          for (int rt = 0; rt < RT; ++rt)      // a row tile past the end is zero-filled by the TMA unit
            tma_load_2d(sp + rt * kABytes, &tmap_q, c * kKC, (tw.bmi * RT + rt) * kBM, st_full + st);
          tma_load_2d(sp + RT * kABytes, &tmap_db, c * kKC, tw.bn * kBN, st_full + st);
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    constexpr uint32_t IDESC = make_idesc(kBN);
    // K-major SWIZZLE_128B: rows of 128 B (64 fp16), 8-row groups 1024 B apart (SBO); a K step of 16 fp16 advances 32 B
    const uint32_t d_hi = (uint32_t)((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
    const uint32_t base = smem_u32(smem);
    int g = 0, it = 0;
    TileWalk tw;
    for (bool ok = tw.init(p, RT); ok; ok = tw.next(), ++it) {
      const int b = it % NB;
      mbar_wait(acc_empty + b, ((it / NB) & 1) ^ 1, 22);
      tc_fence_after();
      for (int c = 0; c < n_kc; ++c, ++g) {
        const int st = g % STAGES;
        mbar_wait(st_full + st, (g / STAGES) & 1, 20);
        tc_fence_after();
        const uint32_t sa = base + (uint32_t)(st * SB);
        const uint32_t b_lo = ((sa + RT * kABytes) >> 4) | (1u << 16);
        if (elect_one()) {
#pragma unroll
          for (int rt = 0; rt < RT; ++rt) {
            const uint32_t a_lo = ((sa + rt * kABytes) >> 4) | (1u << 16);
            const uint32_t d_addr = tmem_base + (uint32_t)((b * RT + rt) * kBN);
#pragma unroll
            for (int kk = 0; kk < kKC / 16; ++kk)
              umma_f16(d_addr, desc_from(a_lo + 2 * kk, d_hi), desc_from(b_lo + 2 * kk, d_hi), IDESC, (c > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(st_empty + st);
          if (c == n_kc - 1) umma_commit(acc_full + b);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue warps 0-3: TMEM lane quarter = warp ==================================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int k = p.k;
    // margin in the scaled domain: 2 * 2^-10 * max|q| * max|d| (plus the tensor core's fp32 accumulation error)
    const float sq = pow2_scale(__uint_as_float(p.stats->q_absmax)), sd = pow2_scale(__uint_as_float(p.stats->d_absmax));
    const float nrm = sqrtf(__uint_as_float(p.stats->q_sumsq)) * sq * sqrtf(__uint_as_float(p.stats->d_sumsq)) * sd;
    const float eps = nrm * (2.f * 9.765625e-4f + 1e-5f) + 1e-30f;
    int cnt[RT];
    bool overflow[RT];
    float thr[RT];                         // append threshold = (running approximate k-th best) - eps
#pragma unroll
    for (int rt = 0; rt < RT; ++rt) { cnt[rt] = 0; overflow[rt] = false; thr[rt] = -INFINITY; }
    // Compaction of the candidates of one row by the whole warp: all-pairs rank on 64-bit keys (order-preserving value
    // bits << 32 | ~column), survivors = everything within eps of the k-th best, written back in rank order.
    auto compact = [&](uint2* buf, int n_src, float& thr_out, bool& ovf_out) -> int {
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
      float kth = -INFINITY;                           // approximate k-th best (-inf while fewer than k candidates)
      float val[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t sv = (uint32_t)(key[t] >> 32);
        const uint32_t vb = (sv & 0x80000000u) ? (sv & 0x7fffffffu) : ~sv;
        val[t] = __uint_as_float(vb);
        if (rank[t] == k - 1 && lane + 32 * t < n_src) kth = val[t];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, o));
      const float cut = kth - eps;                     // -inf - eps = -inf: everything survives
      int keep = 0;
#pragma unroll
      for (int t = 0; t < 4; ++t) keep += __popc(__ballot_sync(0xffffffffu, lane + 32 * t < n_src && val[t] >= cut));
      ovf_out = keep > kCand - 2 * kCandReserve;       // cannot make room: the row is scanned exhaustively later
      if (ovf_out) keep = kCand - 2 * kCandReserve;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < n_src && rank[t] < keep) __stcg(buf + rank[t], make_uint2(__float_as_uint(val[t]), ~(uint32_t)key[t]));
      __syncwarp();
      thr_out = cut;
      return keep;
    };
    int it = 0;
    TileWalk tw;
    for (bool ok = tw.init(p, RT); ok; ok = tw.next(), ++it) {
      const int b = it % NB;
      const long long c0 = (long long)tw.bn * kBN;
      if (tw.first) {
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) { cnt[rt] = 0; thr[rt] = -INFINITY; overflow[rt] = false; }
      }
      mbar_wait(acc_full + b, (it / NB) & 1, 40);
      tc_fence_after();
#pragma unroll
      for (int rt = 0; rt < RT; ++rt) {
        const long long r = (long long)(tw.bmi * RT + rt) * kBM + row;
        uint2* my = p.cand + (((size_t)blockIdx.x * RT + rt) * kBM + row) * kCand;
        uint2* wbase = p.cand + (((size_t)blockIdx.x * RT + rt) * kBM + q * 32) * kCand;
        auto compact_rows = [&](bool mine) {
          unsigned todo = __ballot_sync(0xffffffffu, mine);
          if (todo) __syncwarp();
          while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int n_src = __shfl_sync(0xffffffffu, cnt[rt], src);
            float nthr; bool novf;
            const int ncnt = compact(wbase + (size_t)src * kCand, n_src, nthr, novf);
            if (lane == src) { cnt[rt] = ncnt; thr[rt] = nthr; overflow[rt] = overflow[rt] || novf; }
          }
        };
#pragma unroll 1
        for (int ch = 0; ch < kBN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld32(lane_base + (uint32_t)((b * RT + rt) * kBN + ch * 32), v);
          tmem_ld_wait();
          const long long cb = c0 + ch * 32;
          const int nv = (r < p.rows && cb < p.n) ? (int)((p.n - cb < 32) ? (p.n - cb) : 32) : 0;
#pragma unroll
          for (int half = 0; half < 32 / kCandReserve; ++half) {
#pragma unroll
            for (int jj = 0; jj < kCandReserve; ++jj) {
              const int j = half * kCandReserve + jj;
              const float sv = __uint_as_float(v[j]);
              if (j < nv && sv >= thr[rt]) { __stcg(my + cnt[rt], make_uint2(v[j], (uint32_t)(cb + j))); ++cnt[rt]; }
            }
            compact_rows(cnt[rt] > kCand - kCandReserve);
          }
        }
        if (tw.last) {
          compact_rows(cnt[rt] > 0);                   // final cut at (k-th best - eps)
          if (r < p.rows) {
            const int sp = tw.split();
            uint2* out = p.lists + ((size_t)r * p.splits + sp) * kCand;
            for (int e = 0; e < cnt[rt]; ++e) out[e] = __ldcg(my + e);
            p.counts[r * p.splits + sp] = overflow[rt] ? -1 : cnt[rt];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + b);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Exact decision: one warp per query row.  First the approximate k-th best over ALL splits' survivors (values only: an
// ordered insertion into a warp-held list), which tightens the cut to (global k-th - eps); then, for the survivors above
// that cut, exact fp32 dot products -- lanes split the features, four database rows in flight per step, a shuffle
// reduction each -- and an ordered insertion into the warp-held top-k (lane i = i-th best; larger similarity first,
// lower index on ties).  Rows flagged -1 by the screen are scanned over the whole database.
template <int FPL>      // features per lane (F = 32 * FPL), FPL <= 32
__global__ void __launch_bounds__(256)
knn_rescore_kernel(const float* __restrict__ Q, const float* __restrict__ DB, long long rows, long long n, int F, int k,
                   int splits, const Stats* __restrict__ stats, const uint2* __restrict__ lists, const int* __restrict__ counts,
                   long long* __restrict__ idx_out, float* __restrict__ sim_out) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float qv[FPL];
#pragma unroll
  for (int j = 0; j < FPL; ++j) qv[j] = __ldg(Q + r * F + lane + 32 * j);
  float myv = -INFINITY;
  long long myi = -1;
  auto insert = [&](float cv, long long ci) {
    const unsigned ahead = __ballot_sync(0xffffffffu, myi >= 0 && (myv > cv || (myv == cv && myi < ci)));
    const int pos = __popc(ahead);
    if (pos < k) {
      const float upv = __shfl_up_sync(0xffffffffu, myv, 1);
      const long long upi = __shfl_up_sync(0xffffffffu, myi, 1);
      if (lane > pos) { myv = upv; myi = upi; }
      else if (lane == pos) { myv = cv; myi = ci; }
    }
  };
  auto dot_row = [&](long long ci) -> float {
    const float* d = DB + ci * F;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < FPL; ++j) acc = fmaf(qv[j], __ldg(d + lane + 32 * j), acc);
    return acc;
  };
  bool exhaustive = false;
  for (int sp = 0; sp < splits; ++sp) exhaustive = exhaustive || (counts[r * splits + sp] < 0);
  if (exhaustive) {
    for (long long ci = 0; ci < n; ++ci) insert(warp_sum(dot_row(ci)), ci);
  } else {
    // approximate k-th best over all splits (sorted list of values, lane i = i-th best)
    float av = -INFINITY;
    int have = 0;
    for (int sp = 0; sp < splits; ++sp) {
      const int cnt = counts[r * splits + sp];
      const uint2* l = lists + ((size_t)r * splits + sp) * kCand;
      for (int base = 0; base < cnt; base += 32) {
        const float v0 = (base + lane < cnt) ? __uint_as_float(__ldg(&l[base + lane].x)) : -INFINITY;
        const int nb = min(32, cnt - base);
        for (int e = 0; e < nb; ++e) {
          const float cv = __shfl_sync(0xffffffffu, v0, e);
          const int pos = __popc(__ballot_sync(0xffffffffu, lane < have && av >= cv));
          if (pos < k) {
            const float up = __shfl_up_sync(0xffffffffu, av, 1);
            if (lane > pos) av = up; else if (lane == pos) av = cv;
            have = min(have + 1, k);
          }
        }
      }
    }
    const float kth = __shfl_sync(0xffffffffu, av, k - 1);            // -inf while fewer than k survivors exist
    const float sq = pow2_scale(__uint_as_float(stats->q_absmax)), sd = pow2_scale(__uint_as_float(stats->d_absmax));
    const float nrm = sqrtf(__uint_as_float(stats->q_sumsq)) * sq * sqrtf(__uint_as_float(stats->d_sumsq)) * sd;
    const float cut = kth - (nrm * (2.f * 9.765625e-4f + 1e-5f) + 1e-30f);
    for (int sp = 0; sp < splits; ++sp) {
      const int cnt = counts[r * splits + sp];
      const uint2* l = lists + ((size_t)r * splits + sp) * kCand;
      for (int base = 0; base < cnt; base += 32) {
        uint2 u = make_uint2(0xff800000u, 0u);
        if (base + lane < cnt) u = __ldg(l + base + lane);
        unsigned todo = __ballot_sync(0xffffffffu, base + lane < cnt && __uint_as_float(u.x) >= cut);
        while (todo) {
          // four survivors per step: their loads are in flight together
          long long ci[4];
          float acc[4];
          int nc = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            ci[t] = -1;
            if (todo) { const int src = __ffs(todo) - 1; todo &= todo - 1; ci[t] = (long long)__shfl_sync(0xffffffffu, u.y, src); ++nc; }
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) acc[t] = (t < nc) ? dot_row(ci[t]) : 0.f;
#pragma unroll
          for (int t = 0; t < 4; ++t) if (t < nc) insert(warp_sum(acc[t]), ci[t]);
        }
      }
    }
  }
  if (lane < k) {
    idx_out[r * k + lane] = myi;
    if (sim_out) sim_out[r * k + lane] = myv;
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

// Column splits (and, for experiments, query tiles per CTA): (row-tile groups x splits) items over the SMs with the least
// idle tail.
static void plan_for(long long rows, long long n, int* splits_out, int* rt_out) {
  const long long m_tiles = (rows + kBM - 1) / kBM, n_tiles = (n + kBN - 1) / kBN;
  const int sms = num_sms();
  double eff_rt[3] = {0.0, 0.0, 0.0};
  int split_rt[3] = {1, 1, 1};
  for (int rt = 1; rt <= 2; ++rt) {
    const long long m_items = (m_tiles + rt - 1) / rt;
    for (int s = 1; s <= 8 && s <= n_tiles; ++s) {
      const long long items = m_items * s;
      const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms) * (double)m_tiles / (double)(m_items * rt);
      if (eff > eff_rt[rt] + 1e-9) { eff_rt[rt] = eff; split_rt[rt] = s; }
    }
  }
  // measured at 50 000 x 50 000 x 768: RT = 2 is SLOWER (8.95 vs 6.69 ms): it trades the second accumulator buffer for 33 %
  // less operand traffic, and the serialised epilogue + three-stage ring cost more than the traffic saves.  Kept for study.
  int rt = 1;
  if (getenv("EQUSS_KNN_RT")) rt = (atoi(getenv("EQUSS_KNN_RT")) == 2 && m_tiles >= 2) ? 2 : 1;
  int splits = split_rt[rt];
  if (getenv("EQUSS_KNN_SPLITS")) splits = atoi(getenv("EQUSS_KNN_SPLITS"));
  *splits_out = splits; *rt_out = rt;
}

}  // namespace knnh

bool knn_screen_supported(const float* Q, const float* DB, int F, int k) {
  return F > 0 && F % 64 == 0 && F <= 1024 && k <= 32 && !((uintptr_t)Q & 15) && !((uintptr_t)DB & 15);
}

static int64_t al256(int64_t x) { return (x + 255) & ~(int64_t)255; }

int64_t knn_screen_workspace_bytes(int64_t nq, int64_t n, int F) {
  using namespace knnh;
  int splits, rt;
  plan_for(nq, n, &splits, &rt);
  return al256(nq * (int64_t)F * 2) + al256(n * (int64_t)F * 2) + 256 + al256((int64_t)num_sms() * 2 * kBM * kCand * 8) +
         al256(nq * (int64_t)splits * kCand * 8) + al256(nq * (int64_t)splits * 4) + 1024;
}

int knn_screen_topk_launch(const float* Q, long long nq, const float* DB, long long n, int F, int k, long long* idx_out,
                           float* sim_out, void* workspace, cudaStream_t st) {
  using namespace knnh;
  int splits, rt;
  plan_for(nq, n, &splits, &rt);
  uint8_t* w = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  __half* Qh = (__half*)w;                      w += al256(nq * (int64_t)F * 2);
  __half* Dh = (__half*)w;                      w += al256(n * (int64_t)F * 2);
  Stats* stats = (Stats*)w;                     w += 256;
  uint2* cand = (uint2*)w;                      w += al256((int64_t)num_sms() * 2 * kBM * kCand * 8);
  uint2* lists = (uint2*)w;                     w += al256(nq * (int64_t)splits * kCand * 8);
  int* counts = (int*)w;

  EQUSS_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(Stats), st));
  const int sgrid = num_sms() * 4;
  knn_stats_kernel<<<sgrid, 256, 0, st>>>(Q, nq, F, &stats->q_absmax, &stats->q_sumsq);
  knn_stats_kernel<<<sgrid, 256, 0, st>>>(DB, n, F, &stats->d_absmax, &stats->d_sumsq);
  EQUSS_LAUNCH_OK("knn_stats_kernel");
  knn_to_half_kernel<<<sgrid, 256, 0, st>>>(Q, nq * (long long)F / 4, &stats->q_absmax, Qh);
  knn_to_half_kernel<<<sgrid, 256, 0, st>>>(DB, n * (long long)F / 4, &stats->d_absmax, Dh);
  EQUSS_LAUNCH_OK("knn_to_half_kernel");

  PFN_encodeTiled encode = get_encode_fn();
  EQUSS_REQUIRE(encode != nullptr, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap tq, td;
  auto make = [&](CUtensorMap* tm, const __half* base, long long nrows, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)nrows};
    cuuint64_t gstr[1] = {(cuuint64_t)F * 2};
    cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult c1 = make(&tq, Qh, nq, kBM), c2 = make(&td, Dh, n, kBN);
  EQUSS_REQUIRE(c1 == CUDA_SUCCESS && c2 == CUDA_SUCCESS, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d, %d)", (int)c1, (int)c2);

  Params p;
  memset(&p, 0, sizeof(p));
  p.rows = nq; p.n = n; p.F = F; p.n_kc = F / kKC;
  p.m_tiles = (int)((nq + kBM - 1) / kBM);
  p.n_tiles = (int)((n + kBN - 1) / kBN);
  p.k = k; p.splits = splits; p.stats = stats; p.cand = cand; p.lists = lists; p.counts = counts;
  const long long items = (long long)((p.m_tiles + rt - 1) / rt) * splits;
  int grid = num_sms();
  if (items < grid) grid = (int)items;
  if (rt == 2) {
    EQUSS_CUDA_OK(cudaFuncSetAttribute(knn_screen_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<2>()));
    knn_screen_kernel<2><<<grid, kThreads, smem_bytes<2>(), st>>>(tq, td, p);
  } else {
    EQUSS_CUDA_OK(cudaFuncSetAttribute(knn_screen_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<1>()));
    knn_screen_kernel<1><<<grid, kThreads, smem_bytes<1>(), st>>>(tq, td, p);
  }
  EQUSS_LAUNCH_OK("knn_screen_kernel");

  const unsigned rgrid = (unsigned)((nq + 7) / 8);
  switch (F / 32) {
#define EQUSS_RESCORE(FPLV) case FPLV: knn_rescore_kernel<FPLV><<<rgrid, 256, 0, st>>>(Q, DB, nq, n, F, k, splits, stats, lists, counts, idx_out, sim_out); break;
    EQUSS_RESCORE(2) EQUSS_RESCORE(4) EQUSS_RESCORE(6) EQUSS_RESCORE(8) EQUSS_RESCORE(10) EQUSS_RESCORE(12) EQUSS_RESCORE(14)
    EQUSS_RESCORE(16) EQUSS_RESCORE(18) EQUSS_RESCORE(20) EQUSS_RESCORE(22) EQUSS_RESCORE(24) EQUSS_RESCORE(26) EQUSS_RESCORE(28)
    EQUSS_RESCORE(30) EQUSS_RESCORE(32)
#undef EQUSS_RESCORE
    default: set_error("knn rescore: unsupported feature count %d", F); return EQUSS_ERR_UNSUPPORTED;
  }
  EQUSS_LAUNCH_OK("knn_rescore_kernel");
  return EQUSS_OK;
}

}  // namespace equss
