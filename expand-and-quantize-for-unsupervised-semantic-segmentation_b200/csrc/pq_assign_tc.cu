// pq_assign_tc.cu -- host side of the tcgen05 assign kernel (see pq_assign_tc_kernel.cuh for the design):
// shape planning, B-operand image construction, TMA descriptor, dispatch to the per-d translation units.
#include "pq_assign_tc_kernel.cuh"

namespace equss {
namespace tc {

int launch_tc_d8(int NC, int norm_mode, bool nchw, const CUtensorMap& tmap, const Params& p, int grid, cudaStream_t st);
int launch_tc_d16(int NC, int norm_mode, bool nchw, const CUtensorMap& tmap, const Params& p, int grid, cudaStream_t st);
int launch_tc_d32(int NC, int norm_mode, bool nchw, const CUtensorMap& tmap, const Params& p, int grid, cudaStream_t st);
int launch_tc_d64(int NC, int norm_mode, bool nchw, const CUtensorMap& tmap, const Params& p, int grid, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// B operand image builder: one block per (subspace, code chunk)
//   row r of the image = code k = chunk*NC + r:  [c_hi (D) | c_lo (D) | h m l cn2 0 0 0 0]
//   stored in the UMMA core-matrix layout; trailer (16 B after the last group): max|c|, max|c|^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_b_image_kernel(const float* __restrict__ cb, const float* __restrict__ cn2, int K, int D, int NC, int nchunks,
                     uint8_t* __restrict__ images, int img_bytes) {
  const int m = blockIdx.x / nchunks, c = blockIdx.x % nchunks;
  uint8_t* img = images + (size_t)blockIdx.x * img_bytes;
  const int sbo = sbo_bytes(D);
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
  const float cmax2 = mx;
  const int k0 = c * NC;
  for (int i = threadIdx.x; i < NC * (D / 4); i += blockDim.x) {
    int r = i / (D / 4), jc = i % (D / 4);
    int k = k0 + r;
    int ksrc = (k < K) ? k : k0;               // padding rows duplicate the chunk's first code
    float4 v = *reinterpret_cast<const float4*>(cbm + (size_t)ksrc * D + jc * 4);
    float4 hi, lo;
    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
    uint8_t* rowp = img + (size_t)(r / 8) * sbo + (r % 8) * 16;
    *reinterpret_cast<float4*>(rowp + (size_t)jc * 128) = hi;
    *reinterpret_cast<float4*>(rowp + (size_t)(D / 4 + jc) * 128) = lo;
  }
  for (int r = threadIdx.x; r < NC; r += blockDim.x) {
    int k = k0 + r;
    int ksrc = (k < K) ? k : k0;
    float c2 = cn2m[ksrc];
    float v = -0.5f * c2;
    if (k >= K) v -= 0.00390625f * cmax2 + 1e-30f;   // padded duplicate always loses to the real code
    float h = tf32_hi(v);
    float mm = tf32_hi(v - h);
    float l = tf32_hi((v - h) - mm);
    uint8_t* rowp = img + (size_t)(r / 8) * sbo + (r % 8) * 16;
    *reinterpret_cast<float4*>(rowp + (size_t)(2 * (D / 4)) * 128) = make_float4(h, mm, l, c2);
    *reinterpret_cast<float4*>(rowp + (size_t)(2 * (D / 4) + 1) * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    float* tr = reinterpret_cast<float*>(img + (size_t)(NC / 8) * sbo);
    tr[0] = sqrtf(cmax2); tr[1] = cmax2; tr[2] = 0.f; tr[3] = 0.f;
  }
}

// merge buffer -> indices (multi-chunk codebooks)
__global__ void finalize_merge_kernel(const unsigned long long* __restrict__ merged, long long total,
                                      int32_t* __restrict__ idx_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) idx_out[i] = (int32_t)(merged[i] & 0xFFFFFFFFull);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
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
  int D, NC, stages, abufs, nchunks;
  bool ok;
};

static Plan make_plan(int K, int d) {
  Plan pl{d, 0, 0, 2, 0, false};
  if (!(d == 8 || d == 16 || d == 32 || d == 64)) return pl;
  // instantiated chunk widths: 32 (small codebooks, e.g. the cityscapes yaml's K=32) or 256 (128 for d=64)
  pl.NC = (K <= 32) ? 32 : (d == 64) ? 128 : 256;
  pl.nchunks = (K + pl.NC - 1) / pl.NC;
  pl.abufs = (d == 64) ? 1 : (d == 32) ? 2 : 4;
  pl.stages = (d == 8) ? 8 : (d == 16) ? 6 : (d == 32) ? 4 : 2;
  pl.ok = true;
  return pl;
}

}  // namespace tc

bool assign_tc_supported(const equss_zdesc* zd, int M, int K, int d, int norm_mode, bool want_margin) {
  (void)M;
  if (want_margin) return false;
  tc::Plan pl = tc::make_plan(K, d);
  if (!pl.ok) return false;
  if (norm_mode < EQUSS_NORM_NONE || norm_mode > EQUSS_NORM_AFFINE) return false;
  const bool flat = zd->stride_c == 1 && zd->stride_s == zd->dim;
  const bool nchw = zd->stride_s == 1 && zd->stride_c == zd->hw && zd->stride_b == (int64_t)zd->dim * zd->hw;
  if (!flat && !nchw) return false;
  if (nchw && (zd->hw % 4) != 0) return false;       // TMA global strides must be multiples of 16 bytes
  if (zd->n_pixels <= 0) return false;
  return true;
}

int64_t assign_tc_workspace_bytes(int64_t n_pixels, int M, int K, int d) {
  tc::Plan pl = tc::make_plan(K, d);
  if (!pl.ok) return 0;
  int64_t img = tc::align_up(tc::b_bytes(d, pl.NC), 128);
  int64_t bytes = (int64_t)M * pl.nchunks * img;
  if (pl.nchunks > 1) bytes += (int64_t)M * n_pixels * 8;
  return bytes + 256;
}

int assign_tc_launch(const float* z, const equss_zdesc* zd, const float* codebook_norm, const float* cnorm2, int M,
                     int K, int d, int norm_mode, const float* norm_a, const float* norm_b, int32_t* idx_out,
                     void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  using namespace tc;
  Plan pl = make_plan(K, d);
  EQUSS_REQUIRE(pl.ok, EQUSS_ERR_UNSUPPORTED, "tcgen05 assign: unsupported d=%d", d);
  EQUSS_REQUIRE(workspace && workspace_bytes >= assign_tc_workspace_bytes(zd->n_pixels, M, K, d), EQUSS_ERR_INVALID_ARG,
                "tcgen05 assign: workspace of %lld bytes needed, got %lld",
                (long long)assign_tc_workspace_bytes(zd->n_pixels, M, K, d), (long long)workspace_bytes);
  EQUSS_REQUIRE(!((uintptr_t)z & 15), EQUSS_ERR_INVALID_ARG, "tcgen05 assign: z must be 16-byte aligned");
  PFN_encodeTiled encode = get_encode_fn();
  EQUSS_REQUIRE(encode != nullptr, EQUSS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");

  uint8_t* ws = (uint8_t*)(((uintptr_t)workspace + 127) & ~(uintptr_t)127);
  const int img_bytes = align_up(b_bytes(d, pl.NC), 128);
  uint8_t* images = ws;
  unsigned long long* merged = nullptr;
  if (pl.nchunks > 1) {
    merged = (unsigned long long*)(ws + (size_t)M * pl.nchunks * img_bytes);
    EQUSS_CUDA_OK(cudaMemsetAsync(merged, 0xFF, (size_t)M * zd->n_pixels * 8, st));
  }
  build_b_image_kernel<<<M * pl.nchunks, 256, 0, st>>>(codebook_norm, cnorm2, K, d, pl.NC, pl.nchunks, images, img_bytes);
  EQUSS_LAUNCH_OK("build_b_image_kernel");

  const bool nchw = !(zd->stride_c == 1 && zd->stride_s == zd->dim);
  CUtensorMap tmap;
  CUresult cr;
  if (!nchw) {
    cuuint64_t gdim[2] = {(cuuint64_t)zd->dim, (cuuint64_t)zd->n_pixels};
    cuuint64_t gstr[1] = {(cuuint64_t)zd->dim * 4};
    cuuint32_t box[2] = {(cuuint32_t)d, (cuuint32_t)kTileM};
    cuuint32_t estr[2] = {1, 1};
    // rows of a box are d*4 bytes; promoting 64-byte rows to 128-byte L2 fetches doubles the DRAM traffic
    // (the other half of the line belongs to the neighbouring subspace, read much later by another CTA)
    cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)z, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, d * 4 >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
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
  p.n_pixels = zd->n_pixels; p.hw = zd->hw; p.M = M; p.K = K; p.D = d; p.NC = pl.NC; p.nchunks = pl.nchunks;
  p.nchw = nchw ? 1 : 0;
  p.tiles_per_image = (int)((zd->hw + kTileM - 1) / kTileM);
  p.n_tiles = nchw ? (zd->n_pixels / zd->hw) * p.tiles_per_image : (zd->n_pixels + kTileM - 1) / kTileM;
  p.norm_mode = norm_mode; p.na = norm_a; p.nb = norm_b;
  p.images = images; p.img_bytes = img_bytes; p.idx_out = idx_out; p.merged = merged;
  const long long total_units = (long long)M * pl.nchunks * p.n_tiles;
  int grid = num_sms();
  if (total_units < grid) grid = (int)total_units;

  int rc;
  switch (d) {
    case 8: rc = launch_tc_d8(pl.NC, norm_mode, nchw, tmap, p, grid, st); break;
    case 16: rc = launch_tc_d16(pl.NC, norm_mode, nchw, tmap, p, grid, st); break;
    case 32: rc = launch_tc_d32(pl.NC, norm_mode, nchw, tmap, p, grid, st); break;
    default: rc = launch_tc_d64(pl.NC, norm_mode, nchw, tmap, p, grid, st); break;
  }
  if (rc != EQUSS_OK) return rc;
  if (merged) {
    long long total = (long long)M * zd->n_pixels;
    finalize_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(merged, total, idx_out);
    EQUSS_LAUNCH_OK("finalize_merge_kernel");
  }
  return EQUSS_OK;
}

}  // namespace equss
