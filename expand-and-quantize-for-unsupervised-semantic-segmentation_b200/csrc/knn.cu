// knn.cu -- global-feature kNN (K11): inner products of every query against the database and the k
// largest per query.  Replaces data/precompute_knns.py:313-315 (einsum "nf,mf->nm" on CPU tensors that
// materialises the full N x N similarity matrix, then torch.topk(30)).
//
// Round-1 implementation: query rows are processed in chunks; a tiled fp32 CUDA-core GEMM writes the
// chunk's similarities to a caller-provided workspace that stays L2/HBM resident, and a warp-per-row
// streaming selection keeps the running top-k in registers (lane i holds the i-th best), so the N x N
// matrix never exists.  Ordering: larger similarity first, ties broken by the lower database index.
#include <climits>
#include <cstdlib>
#include "equss_common.cuh"

namespace equss {

// knn_tc.cu: tcgen05 split-tf32 similarity GEMM (F % 32 == 0, 16-byte aligned pointers)
bool knn_gemm_tc_supported(const float* Q, const float* DB, const float* S, long long n, int F);
int knn_gemm_tc_launch(const float* Q, const float* DB, float* S, long long rows, long long n, int F, cudaStream_t st);
int knn_topk_tc_splits(long long rows, long long n);
int64_t knn_topk_tc_scratch_bytes();
bool knn_screen_supported(const float* Q, const float* DB, int F, int k);
int64_t knn_screen_workspace_bytes(int64_t nq, int64_t n, int F);
int knn_screen_topk_launch(const float* Q, long long nq, const float* DB, long long n, int F, int k, long long* idx_out,
                           float* sim_out, void* workspace, cudaStream_t st);
int knn_topk_tc_launch(const float* Q, const float* DB, long long rows, long long n, int F, int k, int splits,
                       float* part_val, int* part_idx, void* scratch, cudaStream_t st);

constexpr int KNN_BM = 128, KNN_BN = 128, KNN_BK = 8;

// S[r][c] = <Q[q0 + r], DB[c]>   for r in [0, rows), c in [0, n)
__global__ void __launch_bounds__(256)
knn_gemm_kernel(const float* __restrict__ Q, const float* __restrict__ DB, float* __restrict__ S,
                long long rows, long long n, int F) {
  __shared__ __align__(16) float As[2][KNN_BK][KNN_BM];
  __shared__ __align__(16) float Bs[2][KNN_BK][KNN_BN];
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.y * KNN_BM;
  const long long c0 = (long long)blockIdx.x * KNN_BN;
  const int lrow = tid >> 1;          // 0..127
  const int lk = (tid & 1) * 4;       // 0 or 4
  const int ty = tid >> 4, tx = tid & 15;   // 16 x 16 threads, 8 x 8 outputs each
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int buf, int k0) {
    long long ra = r0 + lrow, rb = c0 + lrow;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = k0 + lk + i;
      As[buf][lk + i][lrow] = (ra < rows && k < F) ? __ldg(Q + ra * F + k) : 0.f;
      Bs[buf][lk + i][lrow] = (rb < n && k < F) ? __ldg(DB + rb * F + k) : 0.f;
    }
  };
  const int nk = (F + KNN_BK - 1) / KNN_BK;
  load_tile(0, 0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile(buf ^ 1, (kt + 1) * KNN_BK);
#pragma unroll
    for (int k = 0; k < KNN_BK; ++k) {
      float a[8], b[8];
      const float4* a4 = reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4* b4 = reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
      float4 t0 = a4[0], t1 = a4[1], u0 = b4[0], u1 = b4[1];
      a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w; a[4] = t1.x; a[5] = t1.y; a[6] = t1.z; a[7] = t1.w;
      b[0] = u0.x; b[1] = u0.y; b[2] = u0.z; b[3] = u0.w; b[4] = u1.x; b[5] = u1.y; b[6] = u1.z; b[7] = u1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long r = r0 + ty * 8 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      long long c = c0 + tx * 8 + j;
      if (c < n) S[r * n + c] = acc[i][j];
    }
  }
}

// one warp per query row; lane i holds the i-th best (value, index)
__global__ void __launch_bounds__(256)
knn_select_kernel(const float* __restrict__ S, long long rows, long long n, int k,
                  long long* __restrict__ idx_out, float* __restrict__ sim_out) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* s = S + r * n;
  float myv = -INFINITY;
  long long myi = -1;
  float thr = -INFINITY;   // value of the k-th best once the list is full (warp-uniform)
  bool full = false;
  constexpr int U = 8;          // 8 x 32 similarities in flight per warp: the scan is load-latency bound otherwise
  for (long long base0 = 0; base0 < n; base0 += 32 * U) {
    float vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long c = base0 + 32 * u + lane;
      vv[u] = (c < n) ? __ldcs(s + c) : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long base = base0 + 32 * u;
      const long long c = base + lane;
      const float v = vv[u];
      bool cand = (c < n) && (!full || v > thr);
      unsigned mask = __ballot_sync(0xffffffffu, cand);
      while (mask) {
        int src = __ffs(mask) - 1;
        mask &= mask - 1;
        float cv = __shfl_sync(0xffffffffu, v, src);
        long long ci = base + src;
        if (full && !(cv > thr)) continue;       // the threshold may have risen since the ballot (warp-uniform test)
        // entries that rank before the candidate: larger value, or equal value (they have a lower index),
        // but never an empty slot
        unsigned ge = __ballot_sync(0xffffffffu, myi >= 0 && myv >= cv);
        int pos = __popc(ge);
        if (pos < k) {
          float upv = __shfl_up_sync(0xffffffffu, myv, 1);
          long long upi = __shfl_up_sync(0xffffffffu, myi, 1);
          if (lane > pos) { myv = upv; myi = upi; }
          else if (lane == pos) { myv = cv; myi = ci; }
          float tv = __shfl_sync(0xffffffffu, myv, k - 1);
          long long ti = __shfl_sync(0xffffffffu, myi, k - 1);
          full = ti >= 0;
          thr = tv;
        }
      }
    }
  }
  if (lane < k) {
    idx_out[r * k + lane] = myi;
    if (sim_out) sim_out[r * k + lane] = myv;
  }
}

// Merge of the fused kernel's partial lists: one warp per query row, lane i ends with the i-th best candidate.
// Order: larger similarity first, equal similarities by increasing database index (candidates arrive unordered).
__global__ void __launch_bounds__(256)
knn_merge_kernel(const float* __restrict__ pv, const int* __restrict__ pi, long long rows, int ncand, int k,
                 long long* __restrict__ idx_out, float* __restrict__ sim_out) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* v = pv + r * ncand;
  const int* ix = pi + r * ncand;
  float myv = -INFINITY;
  int myi = -1;
  for (int base = 0; base < ncand; base += 32) {
    const int c = base + lane;
    const float cv0 = (c < ncand) ? v[c] : -INFINITY;
    const int ci0 = (c < ncand) ? ix[c] : -1;
    unsigned mask = __ballot_sync(0xffffffffu, ci0 >= 0);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const float cv = __shfl_sync(0xffffffffu, cv0, src);
      const int ci = __shfl_sync(0xffffffffu, ci0, src);
      const unsigned ahead = __ballot_sync(0xffffffffu, myi >= 0 && (myv > cv || (myv == cv && myi < ci)));
      const int pos = __popc(ahead);
      if (pos < k) {
        const float upv = __shfl_up_sync(0xffffffffu, myv, 1);
        const int upi = __shfl_up_sync(0xffffffffu, myi, 1);
        if (lane > pos) { myv = upv; myi = upi; }
        else if (lane == pos) { myv = cv; myi = ci; }
      }
    }
  }
  if (lane < k) {
    idx_out[r * k + lane] = (long long)myi;
    if (sim_out) sim_out[r * k + lane] = myv;
  }
}

static long long knn_rows_per_chunk(long long nq, long long n, long long budget = 1536LL << 20) {
  // rows of similarity workspace per chunk of the unfused path (CUDA-core GEMM shapes)
  long long rc = budget / (n * 4);
  rc = (rc / KNN_BM) * KNN_BM;
  if (rc < KNN_BM) rc = KNN_BM;
  long long nq_up = ((nq + KNN_BM - 1) / KNN_BM) * KNN_BM;
  if (rc > nq_up) rc = nq_up;
  return rc;
}

static bool knn_fused_shape(int F) { return F > 0 && F % 32 == 0; }

}  // namespace equss

using namespace equss;

extern "C" int64_t equss_knn_workspace_bytes(int64_t nq, int64_t n, int F, int k) {
  if (nq <= 0 || n <= 0) return 0;
  if (F > 0 && F % 64 == 0 && F <= 1024 && k <= 32 && getenv("EQUSS_KNN_TF32") == nullptr && getenv("EQUSS_KNN_UNFUSED") == nullptr)
    return knn_screen_workspace_bytes(nq, n, F);                               // fp16 copies + survivor lists
  if (knn_fused_shape(F) && getenv("EQUSS_KNN_UNFUSED") == nullptr)            // partial top-k lists O(nq * k) + per-SM scratch
    return nq * (int64_t)knn_topk_tc_splits(nq, n) * k * 8 + knn_topk_tc_scratch_bytes() + 256;
  return knn_rows_per_chunk(nq, n) * n * 4;
}

extern "C" int equss_knn_topk(const float* queries, int64_t nq, const float* db, int64_t n, int F, int k,
                              int64_t* idx_out, float* sim_out, void* workspace, int64_t workspace_bytes,
                              void* stream) {
  EQUSS_REQUIRE(queries && db && idx_out, EQUSS_ERR_INVALID_ARG, "equss_knn_topk: null pointer");
  EQUSS_REQUIRE(nq >= 0 && n > 0 && F > 0, EQUSS_ERR_INVALID_ARG, "equss_knn_topk: bad shape nq=%lld n=%lld F=%d",
                (long long)nq, (long long)n, F);
  EQUSS_REQUIRE(k >= 1 && k <= 32, EQUSS_ERR_UNSUPPORTED, "equss_knn_topk: k=%d outside [1,32]", k);
  EQUSS_REQUIRE(k <= n, EQUSS_ERR_INVALID_ARG, "equss_knn_topk: k=%d > database size %lld", k, (long long)n);
  EQUSS_REQUIRE(n < (1LL << 31), EQUSS_ERR_UNSUPPORTED, "equss_knn_topk: database of %lld rows", (long long)n);
  if (nq == 0) return EQUSS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = !((uintptr_t)queries & 15) && !((uintptr_t)db & 15) && !((uintptr_t)workspace & 15);
  if (knn_screen_supported(queries, db, F, k) && getenv("EQUSS_KNN_TF32") == nullptr && getenv("EQUSS_KNN_UNFUSED") == nullptr &&
      getenv("EQUSS_KNN_SIMT") == nullptr) {
    // screen on the tensor cores in fp16 (one product), decide in exact fp32 on the survivors (knn_h.cu)
    const int64_t need = knn_screen_workspace_bytes(nq, n, F);
    EQUSS_REQUIRE(workspace && workspace_bytes >= need, EQUSS_ERR_INVALID_ARG,
                  "equss_knn_topk: workspace of %lld bytes needed, got %lld", (long long)need, (long long)workspace_bytes);
    return knn_screen_topk_launch(queries, nq, db, n, F, k, (long long*)idx_out, sim_out, workspace, st);
  }
  if (knn_fused_shape(F) && aligned && getenv("EQUSS_KNN_UNFUSED") == nullptr && getenv("EQUSS_KNN_SIMT") == nullptr) {
    // tcgen05 GEMM with the running top-k in its epilogue, then a merge of the per-split lists
    const int splits = knn_topk_tc_splits(nq, n);
    const int64_t lists = (nq * (int64_t)splits * k * 8 + 15) & ~(int64_t)15;
    const int64_t need = lists + knn_topk_tc_scratch_bytes();
    EQUSS_REQUIRE(workspace && workspace_bytes >= need, EQUSS_ERR_INVALID_ARG,
                  "equss_knn_topk: workspace of %lld bytes needed, got %lld", (long long)need, (long long)workspace_bytes);
    float* pv = (float*)workspace;
    int* pi = (int*)(pv + nq * (int64_t)splits * k);
    int rc2 = knn_topk_tc_launch(queries, db, nq, n, F, k, splits, pv, pi, (char*)workspace + lists, st);
    if (rc2 != EQUSS_OK) return rc2;
    knn_merge_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(pv, pi, nq, splits * k, k, (long long*)idx_out, sim_out);
    EQUSS_LAUNCH_OK("knn_merge_kernel");
    return EQUSS_OK;
  }
  // unfused path (odd feature counts / forced): chunked similarity workspace + streaming selection
  EQUSS_REQUIRE(workspace && workspace_bytes >= (int64_t)KNN_BM * n * 4, EQUSS_ERR_INVALID_ARG,
                "equss_knn_topk: workspace of at least %lld bytes needed, got %lld", (long long)(KNN_BM * n * 4),
                (long long)workspace_bytes);
  const long long rc = knn_rows_per_chunk(nq, n, workspace_bytes);
  float* S = (float*)workspace;
  const bool use_tc = knn_gemm_tc_supported(queries, db, S, n, F) && ((F * 4) % 16 == 0) && getenv("EQUSS_KNN_SIMT") == nullptr;
  for (long long q0 = 0; q0 < nq; q0 += rc) {
    long long rows = nq - q0 < rc ? nq - q0 : rc;
    if (use_tc) {
      int rc2 = knn_gemm_tc_launch(queries + q0 * F, db, S, rows, n, F, st);
      if (rc2 != EQUSS_OK) return rc2;
    } else {
      dim3 grid((unsigned)((n + KNN_BN - 1) / KNN_BN), (unsigned)((rows + KNN_BM - 1) / KNN_BM));
      knn_gemm_kernel<<<grid, 256, 0, st>>>(queries + q0 * F, db, S, rows, n, F);
      EQUSS_LAUNCH_OK("knn_gemm_kernel");
    }
    knn_select_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(S, rows, n, k, (long long*)idx_out + q0 * k,
                                                                   sim_out ? sim_out + q0 * k : nullptr);
    EQUSS_LAUNCH_OK("knn_select_kernel");
  }
  return EQUSS_OK;
}
