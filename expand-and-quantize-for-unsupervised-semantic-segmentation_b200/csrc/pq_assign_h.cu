// pq_assign_h.cu -- host side of the fp16-split tcgen05 assign kernel (see pq_assign_h_kernel.cuh for the
// design): shape planning, operand-image construction, TMA descriptor, dispatch, multi-chunk finalize.
#include "pq_assign_h_kernel.cuh"

namespace equss {
namespace tch {

int launch_tch_d16(int NC, int G, bool nchw, bool fuse, const CUtensorMap& tmap, const Params& p, int grid, cudaStream_t st);
int launch_tch_d32(int NC, int G, bool nchw, bool fuse, const CUtensorMap& tmap, const Params& p, int grid, cudaStream_t st);
int launch_tch_d64(int NC, int G, bool nchw, bool fuse, const CUtensorMap& tmap, const Params& p, int grid, cudaStream_t st);

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------
// Operand image builder: one block per (subspace, code chunk).
//   row r of the image = code k = chunk*NC + r, y = beta*c:
//     [y1 (D fp16) | y2 (D fp16) | b1 b2 b3 0 0 0 0 0 | 0 x 8],  b = -beta*|c|^2/2 in three fp16 pieces
//   in the UMMA K-major core-matrix layout (16-byte chunk j of row r at (r/8)*SBO + j*128 + (r%8)*16);
//   trailer (first float after the last 8-row group): the slot's ambiguity tolerance.
//   beta = 2^-e with max|c| = f*2^e, f in [1/2,1)  (clamped to 2^+-12) is a per-SUBSPACE constant, so scores of
//   different chunks of one subspace are comparable.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_image_kernel(const float* __restrict__ cb, const float* __restrict__ cn2, int K, int D, int NC, int nchunks, int G,
                   uint8_t* __restrict__ images, int img_bytes, unsigned int* __restrict__ flag_count,
                   double* __restrict__ sqerr, int M) {
  asm volatile("griddepcontrol.launch_dependents;");              // the assign kernel's prologue overlaps this kernel
  if (blockIdx.x == 0 && threadIdx.x == 0) *flag_count = 0u;      // list of rows for the exact rescan starts empty
  if (sqerr != nullptr && blockIdx.x == 0)                        // fused gather: the error sums start at zero
    for (int i = threadIdx.x; i < M; i += blockDim.x) sqerr[i] = 0.0;
  const int m = blockIdx.x / nchunks, c = blockIdx.x % nchunks;
  uint8_t* img = images + (size_t)(((m / G) * nchunks + c) * G + (m % G)) * img_bytes;
  const int sbo = b_sbo(D);
  const int C8 = D / 8;
  const float* cbm = cb + (size_t)m * K * D;
  const float* cn2m = cn2 + (size_t)m * K;
  __shared__ float s_max[8];
  float mx = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, cn2m[k]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int i = 0; i < 8; ++i) mx = fmaxf(mx, s_max[i]);
  const float cmax2 = mx, cmax = sqrtf(mx);
  int e = 0;
  if (cmax > 0.f && cmax < INFINITY) (void)frexpf(cmax, &e);
  e = max(-12, min(12, e));
  const float beta = ldexpf(1.f, -e);
  const int k0 = c * NC;
  for (int i = threadIdx.x; i < NC * C8; i += blockDim.x) {
    const int r = i / C8, jc = i % C8;
    const int k = k0 + r;
    float y[8];
    if (k < K) {
      const float4 v0 = *reinterpret_cast<const float4*>(cbm + (size_t)k * D + jc * 8);
      const float4 v1 = *reinterpret_cast<const float4*>(cbm + (size_t)k * D + jc * 8 + 4);
      y[0] = v0.x * beta; y[1] = v0.y * beta; y[2] = v0.z * beta; y[3] = v0.w * beta;
      y[4] = v1.x * beta; y[5] = v1.y * beta; y[6] = v1.z * beta; y[7] = v1.w * beta;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2 hh = __floats2half2_rn(y[2 * q], y[2 * q + 1]);
      const float2 ff = __half22float2(hh);
      hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[q] = pack_h2(y[2 * q] - ff.x, y[2 * q + 1] - ff.y);
    }
    uint8_t* rowp = img + (size_t)(r / 8) * sbo + (r % 8) * 16;
    *reinterpret_cast<uint4*>(rowp + (size_t)jc * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(rowp + (size_t)(C8 + jc) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  for (int r = threadIdx.x; r < NC; r += blockDim.x) {
    const int k = k0 + r;
    float b1 = -60000.f, b2 = 0.f, b3 = 0.f;          // padded columns never win
    if (k < K) {
      const float b = -0.5f * beta * cn2m[k];
      b1 = __half2float(__float2half_rn(b));
      b2 = __half2float(__float2half_rn(b - b1));
      b3 = __half2float(__float2half_rn((b - b1) - b2));
    }
    uint8_t* rowp = img + (size_t)(r / 8) * sbo + (r % 8) * 16;
    *reinterpret_cast<uint4*>(rowp + (size_t)(2 * C8) * 128) = make_uint4(pack_h2(b1, b2), pack_h2(b3, 0.f), 0u, 0u);
    *reinterpret_cast<uint4*>(rowp + (size_t)(2 * C8 + 1) * 128) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    // R bounds |beta s_k| for |z_norm| <= 1 (+ rounding); the absolute floor covers fp16 subnormal rounding of
    // the low pieces: 2^-25 per element on either side and on the bias, two scores compared
    const float R = 1.0001f * beta * cmax + 0.5f * beta * cmax2;
    float tol = kTolRel * R + 1.1920929e-7f * (2.f * sqrtf((float)D) + 1.f);
    if (!(tol >= 0.f)) tol = INFINITY;               // NaN codebook: every row takes the exact path
    float* tr = reinterpret_cast<float*>(img + (size_t)(NC / 8) * sbo);
    tr[0] = tol; tr[1] = beta; tr[2] = R; tr[3] = 0.f;
  }
}

// merge buffer -> indices (multi-chunk codebooks)
__global__ void __launch_bounds__(256)
finalize_merge_kernel(const unsigned long long* __restrict__ merged, long long total, int32_t* __restrict__ idx_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) idx_out[i] = (int32_t)(merged[i] & 0xFFFFFFFFull);
}

// Listed rows (ambiguous inside a chunk, or cross-chunk near-ties; duplicates allowed) get a full exact scan over all K
// codes: one warp per row, the row in registers, lanes stride over the codes, first minimal index wins (the SIMT
// kernel's arithmetic).  FIX: the main kernel already wrote the fused gather with the provisional winner; rows whose
// winner changes get their output row rewritten and the squared-error sum corrected (atomicExch on the index makes
// the repair happen exactly once per row even when the row is listed twice).
template <int D, bool FIX>
__global__ void __launch_bounds__(128)
rescan_flagged_kernel(const uint32_t* __restrict__ list, const unsigned int* __restrict__ count, const float* __restrict__ z,
                      ZView zv, const float* __restrict__ cb, const float* __restrict__ cn2, int K, int32_t* __restrict__ idx_out,
                      const float* __restrict__ gsrc, float* __restrict__ out, double* __restrict__ sqerr) {
  // one BLOCK per listed row: 128 threads stride over the K codes (a d = 64, K = 512 row is 128 KB of codebook reads;
  // one warp per row left those loads latency-bound), warp shuffles + a 4-entry shared stage combine the minima
  constexpr int LPS = D / 4;
  __shared__ float s_best[4];
  __shared__ int s_bi[4];
  __shared__ int s_res[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asm volatile("griddepcontrol.wait;" ::: "memory");               // launched programmatically behind the assign kernel
  const unsigned int n_list = *count;
  for (unsigned int e = blockIdx.x; e < n_list; e += gridDim.x) {
    const long long o = list[e];
    const int m = (int)(o / zv.n_pixels);
    const long long n = o - (long long)m * zv.n_pixels;
    const long long base = pixel_base(zv, n) + (long long)m * D * zv.stride_c;
    float x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = __ldg(z + base + j * zv.stride_c);
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
    int bi = 0x7fffffff;
    for (int k = threadIdx.x; k < K; k += 128) {
      const float4* c4 = reinterpret_cast<const float4*>(cb + ((long long)m * K + k) * D);
      float dot = 0.f;
#pragma unroll
      for (int q = 0; q < LPS; ++q) {
        const float4 c = __ldg(c4 + q);
        dot = fmaf(x[4 * q], c.x, dot);
        dot = fmaf(x[4 * q + 1], c.y, dot);
        dot = fmaf(x[4 * q + 2], c.z, dot);
        dot = fmaf(x[4 * q + 3], c.w, dot);
      }
      const float dist = ref_distance(zn2, __ldg(cn2 + (long long)m * K + k), dot);
      if (dist < best) { best = dist; bi = k; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, s);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __syncthreads();                       // the previous row's readers of the shared stage are done
    if (lane == 0) { s_best[warp] = best; s_bi[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w2 = 1; w2 < 4; ++w2)
        if (s_best[w2] < best || (s_best[w2] == best && s_bi[w2] < bi)) { best = s_best[w2]; bi = s_bi[w2]; }
      if (bi == 0x7fffffff) bi = 0;
      s_res[0] = bi;
      s_res[1] = FIX ? atomicExch(idx_out + o, bi) : (idx_out[o] = bi, bi);
    }
    if constexpr (FIX) {
      __syncthreads();
      bi = s_res[0];
      const int old = s_res[1];
      if (old != bi && warp == 0) {
        // K3 for this row with the exact winner (model/quantizer.py:474,514,536): lane q handles float4 piece q
        float e_new = 0.f, e_old = 0.f;
        if (lane < LPS) {
          const float4 cn = __ldg(reinterpret_cast<const float4*>(gsrc + ((long long)m * K + bi) * D) + lane);
          const float4 co = __ldg(reinterpret_cast<const float4*>(gsrc + ((long long)m * K + old) * D) + lane);
          float xq[4];
#pragma unroll
          for (int q = 0; q < LPS; ++q)
            if (q == lane) { xq[0] = x[4 * q]; xq[1] = x[4 * q + 1]; xq[2] = x[4 * q + 2]; xq[3] = x[4 * q + 3]; }
          const float d0 = cn.x - xq[0], d1 = cn.y - xq[1], d2 = cn.z - xq[2], d3 = cn.w - xq[3];
          e_new = group_sumsq(d0, d1, d2, d3);
          e_old = group_sumsq(co.x - xq[0], co.y - xq[1], co.z - xq[2], co.w - xq[3]);
          float* orow = out + base + (long long)(4 * lane) * zv.stride_c;
          orow[0] = xq[0] + d0; orow[zv.stride_c] = xq[1] + d1;
          orow[2 * zv.stride_c] = xq[2] + d2; orow[3 * zv.stride_c] = xq[3] + d3;
        }
        const float de = warp_sum(e_new) - warp_sum(e_old);
        if (lane == 0) atomicAdd(sqerr + m, (double)de);
      }
    }
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

struct Plan {
  int NC, nchunks, G;
  bool ok;
};

static Plan make_plan(int M, int K, int d, bool nchw) {
  Plan pl{0, 0, 1, false};
  if (!(d == 16 || d == 32 || d == 64)) return pl;
  pl.NC = (K <= 32) ? 32 : 256;
  pl.nchunks = (K + pl.NC - 1) / pl.NC;
  const int g = 32 / d;                      // subspaces per 128-byte line of a flat row
  pl.G = (!nchw && g > 1 && M % g == 0) ? g : 1;
  pl.ok = true;
  return pl;
}

static bool is_flat(const equss_zdesc* zd) { return zd->stride_c == 1 && zd->stride_s == zd->dim; }
static bool is_nchw(const equss_zdesc* zd) {
  return zd->stride_s == 1 && zd->stride_c == zd->hw && zd->stride_b == (int64_t)zd->dim * zd->hw;
}

}  // namespace tch

bool assign_tch_supported(const equss_zdesc* zd, int M, int K, int d, int norm_mode, bool want_margin) {
  if (want_margin || norm_mode != EQUSS_NORM_L2) return false;
  const bool flat = tch::is_flat(zd), nchw = tch::is_nchw(zd);
  if (!flat && !nchw) return false;
  if (!tch::make_plan(M, K, d, !flat).ok) return false;
  if ((int64_t)M * zd->n_pixels >= (int64_t)1 << 32) return false;   // near-tie list entries are 32-bit
  if (!flat && (zd->hw % 4) != 0) return false;      // TMA global strides must be multiples of 16 bytes
  if (!flat && zd->hw >= ((int64_t)1 << 24)) return false;   // 32-bit channel offsets inside one subspace (d * hw elements)
  return zd->n_pixels > 0;
}

int64_t assign_tch_workspace_bytes(int64_t n_pixels, int M, int K, int d) {
  tch::Plan pl = tch::make_plan(M, K, d, false);
  if (!pl.ok) return 0;
  const int64_t img = tch::align_up(tch::b_bytes(d, pl.NC), 128);
  int64_t bytes = (int64_t)M * pl.nchunks * img + 64;                                   // operand images, list counter
  if (pl.nchunks > 1) bytes += (int64_t)M * n_pixels * 8;                               // cross-chunk merge buffer
  bytes += (int64_t)M * n_pixels * 4 * (pl.nchunks > 1 ? 2 * pl.nchunks : 1);           // rows listed for the exact scan
  return bytes + 256;
}

bool assign_tch_fusable(const equss_zdesc* zd, int M, int K, int d, int norm_mode) {
  if (!assign_tch_supported(zd, M, K, d, norm_mode, false)) return false;
  return (d == 16 || d == 32) && K <= 256;          // single code chunk, raw ring deep enough for the gather lag
}

int assign_tch_launch(const float* z, const equss_zdesc* zd, const float* codebook_norm, const float* cnorm2, int M,
                      int K, int d, int32_t* idx_out, void* workspace, int64_t workspace_bytes, cudaStream_t st,
                      const float* gather_src, float* out, double* sqerr) {
  using namespace tch;
  const bool nchw = !is_flat(zd);
  const bool fuse = out != nullptr;
  Plan pl = make_plan(M, K, d, nchw);
  EQUSS_REQUIRE(!fuse || (gather_src && sqerr && assign_tch_fusable(zd, M, K, d, EQUSS_NORM_L2)), EQUSS_ERR_UNSUPPORTED,
                "tcgen05 f16x2 assign: fused gather needs d in {16,32}, K <= 256 (got d=%d K=%d)", d, K);
  EQUSS_REQUIRE(!fuse || (!((uintptr_t)out & 15) && !((uintptr_t)gather_src & 15)), EQUSS_ERR_INVALID_ARG,
                "tcgen05 f16x2 assign: out and gather_src must be 16-byte aligned");
  EQUSS_REQUIRE(pl.ok, EQUSS_ERR_UNSUPPORTED, "tcgen05 f16x2 assign: unsupported d=%d", d);
  EQUSS_REQUIRE(workspace && workspace_bytes >= assign_tch_workspace_bytes(zd->n_pixels, M, K, d), EQUSS_ERR_INVALID_ARG,
                "tcgen05 f16x2 assign: workspace of %lld bytes needed, got %lld",
                (long long)assign_tch_workspace_bytes(zd->n_pixels, M, K, d), (long long)workspace_bytes);
  EQUSS_REQUIRE(!((uintptr_t)z & 15) && !((uintptr_t)codebook_norm & 15), EQUSS_ERR_INVALID_ARG,
                "tcgen05 f16x2 assign: z and the codebook must be 16-byte aligned");
  PFN_encodeTiled encode = get_encode_fn();
  EQUSS_REQUIRE(encode != nullptr, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");

  uint8_t* ws = (uint8_t*)(((uintptr_t)workspace + 127) & ~(uintptr_t)127);
  const int img_bytes = align_up(b_bytes(d, pl.NC), 128);
  uint8_t* images = ws;
  unsigned long long* merged = nullptr;
  unsigned int* flag_count = (unsigned int*)(ws + (size_t)M * pl.nchunks * img_bytes);   // 16 bytes, then merge buffer / list
  uint32_t* flag_list;
  if (pl.nchunks > 1) {
    merged = (unsigned long long*)(flag_count + 4);
    flag_list = (uint32_t*)(merged + (size_t)M * zd->n_pixels);
    EQUSS_CUDA_OK(cudaMemsetAsync(flag_count, 0, 16 + (size_t)M * zd->n_pixels * 8, st));
  } else {
    flag_list = (uint32_t*)(flag_count + 4);          // the counter is reset by build_image_kernel
  }
  build_image_kernel<<<M * pl.nchunks, 256, 0, st>>>(codebook_norm, cnorm2, K, d, pl.NC, pl.nchunks, pl.G, images, img_bytes, flag_count,
                                                          fuse ? sqerr : nullptr, M);
  EQUSS_LAUNCH_OK("build_image_kernel");

  CUtensorMap tmap;
  CUresult cr;
  if (!nchw) {
    cuuint64_t gdim[2] = {(cuuint64_t)zd->dim, (cuuint64_t)zd->n_pixels};
    cuuint64_t gstr[1] = {(cuuint64_t)zd->dim * 4};
    cuuint32_t box[2] = {(cuuint32_t)d, (cuuint32_t)kTileM};
    cuuint32_t estr[2] = {1, 1};
    // 64-byte rows: the other half of each 128-byte line belongs to the partner subspace, which the same CTA
    // requests next (unit order), so full-line promotion is what we want when G > 1
    cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)z, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE,
                (d * 4 >= 128 || pl.G > 1) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const long long B = zd->n_pixels / zd->hw;
    cuuint64_t gdim[3] = {(cuuint64_t)zd->hw, (cuuint64_t)zd->dim, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)zd->hw * 4, (cuuint64_t)zd->dim * zd->hw * 4};
    cuuint32_t box[3] = {(cuuint32_t)kTileM, (cuuint32_t)d, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)z, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  EQUSS_REQUIRE(cr == CUDA_SUCCESS, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)cr);

  Params p;
  p.n_pixels = zd->n_pixels; p.hw = zd->hw; p.M = M; p.K = K; p.D = d; p.NC = pl.NC; p.nchunks = pl.nchunks; p.G = pl.G;
  p.tiles_per_image = (int)((zd->hw + kTileM - 1) / kTileM);
  p.n_tiles = nchw ? (zd->n_pixels / zd->hw) * p.tiles_per_image : (zd->n_pixels + kTileM - 1) / kTileM;
  p.z = z; p.zv = make_view(zd); p.cb = codebook_norm; p.cn2 = cnorm2;
  p.images = images; p.img_bytes = img_bytes; p.idx_out = idx_out; p.merged = merged; p.flag_list = flag_list; p.flag_count = flag_count;
  p.gsrc = gather_src; p.out = out; p.sqerr = sqerr;
  p.use_gtab = (fuse && (long long)p.n_tiles * pl.G >= 32 && getenv("EQUSS_NO_GTAB") == nullptr) ? 1 : 0;   // slots much longer than the raw ring
  const long long total_units = (long long)M * pl.nchunks * p.n_tiles;
  int grid = num_sms();
  if (total_units < grid) grid = (int)total_units;

  int rc;
  switch (d) {
    case 16: rc = launch_tch_d16(pl.NC, pl.G, nchw, fuse, tmap, p, grid, st); break;
    case 32: rc = launch_tch_d32(pl.NC, pl.G, nchw, fuse, tmap, p, grid, st); break;
    default: rc = launch_tch_d64(pl.NC, pl.G, nchw, fuse, tmap, p, grid, st); break;
  }
  if (rc != EQUSS_OK) return rc;
  if (merged) {
    const long long total = (long long)M * zd->n_pixels;
    finalize_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(merged, total, idx_out);
    EQUSS_LAUNCH_OK("finalize_merge_kernel");
  }
  // exact scan of the listed rows (a few per ten thousand); repairs the fused gather where the winner changes
  const int rgrid = num_sms() * 8;
  cudaLaunchConfig_t rcfg = {};
  rcfg.gridDim = dim3((unsigned)rgrid); rcfg.blockDim = dim3(128); rcfg.dynamicSmemBytes = 0; rcfg.stream = st;
  cudaLaunchAttribute rattr[1];
  rattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  rattr[0].val.programmaticStreamSerializationAllowed = (merged == nullptr) ? 1 : 0;   // behind finalize_merge: plain launch
  rcfg.attrs = rattr; rcfg.numAttrs = 1;
  const float* no_src = nullptr; float* no_out = nullptr; double* no_err = nullptr;
  const uint32_t* flag_list_c = flag_list; const unsigned int* flag_count_c = flag_count;
#define EQUSS_RESCAN(DV)                                                                                                  \
  if (fuse) { EQUSS_CUDA_OK(cudaLaunchKernelEx(&rcfg, rescan_flagged_kernel<DV, true>, flag_list_c, flag_count_c, z, p.zv, codebook_norm, cnorm2, K, idx_out, gather_src, out, sqerr)); } \
  else { EQUSS_CUDA_OK(cudaLaunchKernelEx(&rcfg, rescan_flagged_kernel<DV, false>, flag_list_c, flag_count_c, z, p.zv, codebook_norm, cnorm2, K, idx_out, no_src, no_out, no_err)); }
  if (d == 16) { EQUSS_RESCAN(16) } else if (d == 32) { EQUSS_RESCAN(32) } else { EQUSS_RESCAN(64) }
#undef EQUSS_RESCAN
  EQUSS_LAUNCH_OK("rescan_flagged_kernel");
  return EQUSS_OK;
}

}  // namespace equss
