// pq_rows.cu -- HBM-bound per-row passes of the PQ head:
//   K3  gather + commitment/codebook squared error + straight-through value  (model/quantizer.py:474,514,534-536)
//   K3' backward of K3 w.r.t. z and the gathered codebook rows
//   K4  per-code counts and sums (segmented scatter-add in shared memory)    (model/quantizer.py:485-488)
//   K6  EMA codebook update                                                   (model/quantizer.py:233-254)
//
// Two mappings per pass:
//   * "flat vector" path: activations are (n, D) row-major, d/4 is a power of two <= 32.  LPS = d/4
//     consecutive lanes own one (pixel, subspace) row, one float4 each: every global access is a fully
//     coalesced 16-byte-per-lane stream; row statistics are an xor-shuffle butterfly over LPS lanes.
//   * "strided scalar" path: any strides (NCHW: consecutive threads = consecutive pixels, so each
//     channel read is coalesced) and any d <= 256; one thread owns a row.
// Both produce bit-identical z_norm (canonical association order, equss_common.cuh).
#include "equss_common.cuh"

namespace equss {

// ------------------------------------------------------------------------------------------------
// Row holder for the scalar path: DT > 0 -> row cached in registers, DT == 0 -> re-read from global.
// ------------------------------------------------------------------------------------------------
template <int DT>
struct ScalarRow {
  float x[DT > 0 ? DT : 1];
  const float* p;
  long long sc;
  int d;
  __device__ __forceinline__ void load(const float* base, long long stride_c, int d_) {
    p = base; sc = stride_c; d = d_;
    if (DT > 0) {
#pragma unroll
      for (int j = 0; j < (DT > 0 ? DT : 1); ++j) x[j] = __ldg(base + j * stride_c);
    }
  }
  __device__ __forceinline__ float raw(int j) const {
    if (DT > 0) return x[j];
    return __ldg(p + j * sc);
  }
  __device__ __forceinline__ RowNorm norm(int mode) const {
    if (DT >= 4) {
      RowNorm r; r.shift = 0.f; r.denom = 1.f;
      constexpr int G = (DT >= 4 ? DT / 4 : 1);
      if (mode == EQUSS_NORM_L2) {
        float g[G];
#pragma unroll
        for (int i = 0; i < G; ++i) g[i] = group_sumsq(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        r = l2_from_sumsq(butterfly_array<G>(g));
      } else if (mode == EQUSS_NORM_ZNORM) {
        float g[G];
#pragma unroll
        for (int i = 0; i < G; ++i) g[i] = group_sum(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        float mean = butterfly_array<G>(g) / (float)DT;
#pragma unroll
        for (int i = 0; i < G; ++i)
          g[i] = group_sumsq(x[4 * i] - mean, x[4 * i + 1] - mean, x[4 * i + 2] - mean, x[4 * i + 3] - mean);
        float stdv = sqrtf(butterfly_array<G>(g) / (float)(DT - 1));
        r.shift = mean; r.denom = stdv + kStdEps;
      }
      return r;
    } else {
      return row_norm_generic(mode, d, [&](int j) { return raw(j); });
    }
  }
};

__device__ __forceinline__ float norm_elem(float x, const RowNorm& r, int mode, const float* na,
                                           const float* nb, int ch) {
  if (mode == EQUSS_NORM_AFFINE) return (x - __ldg(na + ch)) / __ldg(nb + ch);
  return apply_norm(x, r, mode);
}

// ------------------------------------------------------------------------------------------------
// K3 flat vector path
// ------------------------------------------------------------------------------------------------
template <int LPS>
__global__ void __launch_bounds__(256)
gather_loss_flat_kernel(const float4* __restrict__ z, long long n_pixels, int D4, int M, int K,
                        const float4* __restrict__ src, const int32_t* __restrict__ idx, int mode,
                        const float* __restrict__ na, const float* __restrict__ nb,
                        float4* __restrict__ out, float4* __restrict__ znorm_out,
                        double* __restrict__ sqerr) {
  extern __shared__ float s_sq[];  // [M]
  for (int i = threadIdx.x; i < M; i += blockDim.x) s_sq[i] = 0.f;
  __syncthreads();
  const long long total = n_pixels * D4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // warp-uniform trip count: the xor-shuffles below need every lane of the warp, so the bound is
  // tested on the warp's first element and the tail is predicated instead of exiting the loop
  for (long long f0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); f0 < total; f0 += stride) {
    const long long f = f0 + (threadIdx.x & 31);
    const bool live = f < total;
    const long long fc = live ? f : total - 1;
    long long n = fc / D4;
    int c4 = (int)(fc - n * D4);
    int m = c4 / LPS;
    int l = c4 - m * LPS;
    float4 v = __ldcs(z + fc);
    RowNorm r; r.shift = 0.f; r.denom = 1.f;
    if (mode == EQUSS_NORM_L2) {
      r = l2_from_sumsq(butterfly_lanes<LPS>(group_sumsq(v.x, v.y, v.z, v.w)));
    } else if (mode == EQUSS_NORM_ZNORM) {
      constexpr int d = LPS * 4;
      float mean = butterfly_lanes<LPS>(group_sum(v.x, v.y, v.z, v.w)) / (float)d;
      float ssd = butterfly_lanes<LPS>(group_sumsq(v.x - mean, v.y - mean, v.z - mean, v.w - mean));
      r.shift = mean; r.denom = sqrtf(ssd / (float)(d - 1)) + kStdEps;
    }
    int ch = c4 * 4;
    float4 zn;
    if (mode == EQUSS_NORM_L2) {
      zn = div4_by(v, r.denom);              // correctly rounded, one reciprocal for the four lanes' elements
    } else {
      zn.x = norm_elem(v.x, r, mode, na, nb, ch);
      zn.y = norm_elem(v.y, r, mode, na, nb, ch + 1);
      zn.z = norm_elem(v.z, r, mode, na, nb, ch + 2);
      zn.w = norm_elem(v.w, r, mode, na, nb, ch + 3);
    }
    int code = __ldg(idx + (long long)m * n_pixels + n);
    float4 q = __ldg(src + ((long long)m * K + code) * LPS + l);
    float4 dq, o;
    dq.x = q.x - zn.x; dq.y = q.y - zn.y; dq.z = q.z - zn.z; dq.w = q.w - zn.w;
    o.x = zn.x + dq.x; o.y = zn.y + dq.y; o.z = zn.z + dq.z; o.w = zn.w + dq.w;   // STE value (:536)
    if (live) {
      __stcs(out + f, o);
      if (znorm_out) __stcs(znorm_out + f, zn);
    }
    float e = group_sumsq(dq.x, dq.y, dq.z, dq.w);
    e = butterfly_lanes<LPS>(e);
    if (l == 0 && live) atomicAdd(&s_sq[m], e);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    float v = s_sq[i];
    if (v != 0.f) atomicAdd(sqerr + i, (double)v);
  }
}

// ------------------------------------------------------------------------------------------------
// K3 flat fast path (l2 rows): as gather_loss_flat_kernel, four independent float4 per thread and iteration and
// branch-free correctly rounded sqrt / division, so their memory latencies and dependent chains overlap.
// The codeword load depends on the index load; to keep that second memory latency off the iteration's critical
// path the indices of iteration i+1 are fetched while iteration i is in flight (software pipeline: z, codeword and
// the next indices are all issued before the first use), and the (pixel, chunk) position of every element advances
// incrementally -- no division inside the loop.
// ------------------------------------------------------------------------------------------------
template <int LPS>
__global__ void __launch_bounds__(256, 3)
gather_loss_flat_l2_kernel(const float4* __restrict__ z, long long n_pixels, int D4, int M, int K,
                           const float4* __restrict__ src, const int32_t* __restrict__ idx,
                           float4* __restrict__ out, double* __restrict__ sqerr) {
  extern __shared__ float s_sq[];  // [M]
  for (int i = threadIdx.x; i < M; i += blockDim.x) s_sq[i] = 0.f;
  __syncthreads();
  constexpr int U = 4;
  const long long total = n_pixels * D4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long f00 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // element u of an iteration sits at f0 + u * stride; from one iteration to the next it moves by stride * U
  const long long step = stride * U;
  const long long dn = step / D4;
  const int dc = (int)(step - dn * D4);
  long long n[U];
  int c4[U], code[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long f = f00 + u * stride;
    n[u] = f / D4;
    c4[u] = (int)(f - n[u] * D4);
    code[u] = (f < total) ? __ldg(idx + (long long)(c4[u] / LPS) * n_pixels + n[u]) : 0;
  }
  // warp-uniform trip count (xor-shuffles inside); the tail is predicated.  A group of LPS lanes (one subspace of one
  // pixel) is live or dead as a whole: total and every warp's first element are multiples of LPS.
  for (long long f0 = f00; f0 - (threadIdx.x & 31) < total; f0 += step) {
    float4 v[U], q[U];
    int mm[U];
    bool live[U], first[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long f = f0 + u * stride;
      live[u] = f < total;
      mm[u] = c4[u] / LPS;
      const int ll = c4[u] - mm[u] * LPS;
      first[u] = ll == 0;
      v[u] = __ldcs(z + (live[u] ? f : 0));                                       // dead lanes read valid memory
      q[u] = __ldg(src + ((long long)mm[u] * K + code[u]) * LPS + ll);
    }
    // indices of the next iteration
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int c = c4[u] + dc;
      long long nn = n[u] + dn;
      if (c >= D4) { c -= D4; ++nn; }
      c4[u] = c; n[u] = nn;
      code[u] = (f0 + u * stride + step < total) ? __ldg(idx + (long long)(c / LPS) * n_pixels + nn) : 0;
    }
    float ss[U];
#pragma unroll
    for (int u = 0; u < U; ++u) ss[u] = group_sumsq(v[u].x, v[u].y, v[u].z, v[u].w);
#pragma unroll
    for (int sft = 1; sft < LPS; sft <<= 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], sft);
    }
    float e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 zn = div4_fast(v[u], l2_denom_fast(ss[u]));
      float4 dq, o;
      dq.x = q[u].x - zn.x; dq.y = q[u].y - zn.y; dq.z = q[u].z - zn.z; dq.w = q[u].w - zn.w;
      o.x = zn.x + dq.x; o.y = zn.y + dq.y; o.z = zn.z + dq.z; o.w = zn.w + dq.w;   // STE value (:536)
      if (live[u]) __stcs(out + (f0 + u * stride), o);
      e[u] = group_sumsq(dq.x, dq.y, dq.z, dq.w);
    }
#pragma unroll
    for (int sft = 1; sft < LPS; sft <<= 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) e[u] += __shfl_xor_sync(0xffffffffu, e[u], sft);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (first[u] && live[u]) atomicAdd(&s_sq[mm[u]], e[u]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    float v = s_sq[i];
    if (v != 0.f) atomicAdd(sqerr + i, (double)v);
  }
}

// ------------------------------------------------------------------------------------------------
// K3 strided scalar path: grid = (pixel chunks, M); thread = one pixel of subspace blockIdx.y
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(128)
gather_loss_scalar_kernel(const float* __restrict__ z, ZView zv, int M, int K, int d,
                          const float* __restrict__ src, const int32_t* __restrict__ idx, int mode,
                          const float* __restrict__ na, const float* __restrict__ nb,
                          float* __restrict__ out, float* __restrict__ znorm_out,
                          double* __restrict__ sqerr) {
  const int m = blockIdx.y;
  float local = 0.f;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < zv.n_pixels;
       n += (long long)gridDim.x * blockDim.x) {
    long long base = pixel_base(zv, n) + (long long)m * d * zv.stride_c;
    ScalarRow<DT> row;
    row.load(z + base, zv.stride_c, d);
    RowNorm r = row.norm(mode);
    int code = __ldg(idx + (long long)m * zv.n_pixels + n);
    const float* q = src + ((long long)m * K + code) * d;
    float e = 0.f;
    const int dd = DT > 0 ? DT : d;
    // error accumulated in the canonical group order so both paths agree
    float g = 0.f;
    float qr[DT > 0 ? DT : 1];
    if (DT >= 4) {            // codebook row: 16-byte loads (rows are d*4 bytes, d % 4 == 0 -> aligned)
#pragma unroll
      for (int j = 0; j < (DT >= 4 ? DT : 0); j += 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(q + j));
        qr[j] = t.x; qr[j + 1] = t.y; qr[j + 2] = t.z; qr[j + 3] = t.w;
      }
    }
#pragma unroll 4
    for (int j = 0; j < dd; ++j) {
      float zn = norm_elem(row.raw(j), r, mode, na, nb, m * d + j);
      float dq = (DT >= 4 ? qr[j] : __ldg(q + j)) - zn;
      out[base + j * zv.stride_c] = zn + dq;
      if (znorm_out) znorm_out[base + j * zv.stride_c] = zn;
      g = ((j & 3) == 0) ? dq * dq : fmaf(dq, dq, g);
      if ((j & 3) == 3 || j == dd - 1) { e += g; g = 0.f; }
    }
    local += e;
  }
  local = warp_sum(local);
  __shared__ float s_part[4];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_part[i];
    if (t != 0.f) atomicAdd(sqerr + m, (double)t);
  }
}

// ------------------------------------------------------------------------------------------------
// K3 NCHW fast path (l2 rows, d in {16, 32, 64}): one thread per (pixel, subspace), channel accesses coalesced
// across the warp's 32 consecutive pixels.  Branch-free correctly rounded sqrt / division (l2_denom_fast,
// div4_fast: bit-identical to sqrtf and `/`) so the compiler interleaves the d/4 channel groups instead of
// serialising them behind the slow-path calls of the IEEE routines.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(128)
gather_loss_nchw_l2_kernel(const float* __restrict__ z, long long hw, long long stride_b, int n_images, int M, int K,
                           const float* __restrict__ src, const int32_t* __restrict__ idx, float* __restrict__ out,
                           double* __restrict__ sqerr) {
  constexpr int G = DT / 4;
  const int m = blockIdx.y;
  const int tiles_per_image = (int)((hw + 127) / 128);
  const long long n_tiles = (long long)n_images * tiles_per_image;
  const long long n_pixels = (long long)n_images * hw;
  float local = 0.f;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long b = t / tiles_per_image;
    const long long s = (t - b * tiles_per_image) * 128 + threadIdx.x;
    if (s >= hw) continue;
    const long long base = b * stride_b + (long long)m * DT * hw + s;
    float x[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) x[j] = __ldcs(z + base + (long long)j * hw);
    const int code = __ldg(idx + (long long)m * n_pixels + b * hw + s);
    const float4* q4 = reinterpret_cast<const float4*>(src + ((long long)m * K + code) * DT);
    float g[G];
#pragma unroll
    for (int i = 0; i < G; ++i) g[i] = group_sumsq(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    const float denom = l2_denom_fast(butterfly_array<G>(g));
    float e = 0.f;
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const float4 q = __ldg(q4 + i);
      const float4 zn = div4_fast(make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]), denom);
      const float d0 = q.x - zn.x, d1 = q.y - zn.y, d2 = q.z - zn.z, d3 = q.w - zn.w;
      float* o = out + base + (long long)(4 * i) * hw;
      __stcs(o, zn.x + d0); __stcs(o + hw, zn.y + d1); __stcs(o + 2 * hw, zn.z + d2); __stcs(o + 3 * hw, zn.w + d3);
      e += group_sumsq(d0, d1, d2, d3);
    }
    local += e;
  }
  local = warp_sum(local);
  __shared__ float s_part[4];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float tsum = ((s_part[0] + s_part[1]) + s_part[2]) + s_part[3];
    if (tsum != 0.f) atomicAdd(sqerr + m, (double)tsum);
  }
}

// ------------------------------------------------------------------------------------------------
// K3' backward, scalar mapping for every layout (the row is needed twice; registers hold it).
//   g = grad_out + coef[m]*(z_norm - q);   grad_z = J^T g
//   NONE  : grad_z = g
//   L2    : grad_z = (g - z_norm * <z_norm, g>) / denom           (denom = max(||z||, eps); the
//            clamp branch has zero measure and is treated like the unclamped one, as autograd does
//            for ||z|| > eps)
//   ZNORM : y = (x-mean)/(std+eps);  grad_x = (g - mean(g) - y * sum(g*y) * std/((d-1)*(std+eps)) ... )
//            derived below; AFFINE: grad_z = g / denom[c]
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(128)
gather_loss_bwd_kernel(const float* __restrict__ z, ZView zv, int M, int K, int d,
                       const float* __restrict__ src, const int32_t* __restrict__ idx, int mode,
                       const float* __restrict__ na, const float* __restrict__ nb,
                       const float* __restrict__ grad_out, const float* __restrict__ coef,
                       float* __restrict__ grad_z, const float* __restrict__ cb_coef,
                       float* __restrict__ grad_codebook) {
  const int m = blockIdx.y;
  const float cf = coef ? __ldg(coef + m) : 0.f;
  const float cbf = (cb_coef && grad_codebook) ? __ldg(cb_coef + m) : 0.f;
  const int dd = DT > 0 ? DT : d;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < zv.n_pixels;
       n += (long long)gridDim.x * blockDim.x) {
    long long base = pixel_base(zv, n) + (long long)m * d * zv.stride_c;
    ScalarRow<DT> row;
    row.load(z + base, zv.stride_c, d);
    RowNorm r = row.norm(mode);
    int code = __ldg(idx + (long long)m * zv.n_pixels + n);
    const float* q = src + ((long long)m * K + code) * d;
    float* gc = grad_codebook ? grad_codebook + ((long long)m * K + code) * d : nullptr;
    // pass 1: reductions needed by the Jacobian
    float s_gy = 0.f, s_g = 0.f;
    for (int j = 0; j < dd; ++j) {
      float zn = norm_elem(row.raw(j), r, mode, na, nb, m * d + j);
      float dq = zn - __ldg(q + j);
      float g = (grad_out ? grad_out[base + j * zv.stride_c] : 0.f) + cf * dq;
      s_gy = fmaf(g, zn, s_gy);
      s_g += g;
      if (gc && cbf != 0.f) atomicAdd(gc + j, -cbf * dq);   // d/dq of cb_coef/2 * (q - z_norm)^2 summed
    }
    if (!grad_z) continue;
    float inv = 1.f / r.denom;
    float stdv = r.denom - kStdEps;
    for (int j = 0; j < dd; ++j) {
      float zn = norm_elem(row.raw(j), r, mode, na, nb, m * d + j);
      float dq = zn - __ldg(q + j);
      float g = (grad_out ? grad_out[base + j * zv.stride_c] : 0.f) + cf * dq;
      float gz;
      if (mode == EQUSS_NORM_NONE) {
        gz = g;
      } else if (mode == EQUSS_NORM_L2) {
        gz = (g - zn * s_gy) * inv;
      } else if (mode == EQUSS_NORM_ZNORM) {
        // y_j = (x_j - mu)/(s+e);  dy_j/dx_i = (delta_ij - 1/d)/(s+e) - (x_j-mu)(x_i-mu)/((d-1) s (s+e)^2)
        // grad_x_i = (g_i - mean(g))/(s+e) - y_i * (s+e) * sum_j(g_j y_j) / ((d-1) s (s+e)) ... simplified:
        float yi = zn;
        float term = (stdv > 0.f) ? yi * s_gy * r.denom / ((float)(dd - 1) * stdv) : 0.f;
        gz = (g - s_g / (float)dd - term) * inv;
      } else {  // AFFINE
        gz = g / __ldg(nb + m * d + j);
      }
      grad_z[base + j * zv.stride_c] = gz;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4 segmented scatter-add.  grid = (pixel chunks, M).  Shared accumulators [K][d+1] for subspace
// blockIdx.y (column d = count), flushed with one global atomic per touched entry.
// ------------------------------------------------------------------------------------------------
// fp32 atomicAdd on shared memory is a compare-and-swap loop on sm_100 (LDS + ATOMS.CAST.SPIN per element), and the
// first version of this kernel spent its time there (2.1 TB/s).  The shared accumulators are therefore laid out as
// [K][d+4] floats (16-byte aligned rows) and updated four elements at a time with the native 128-bit ATOMS.CAS.128;
// the per-code count is a separate int array with the native integer atomic.
__device__ __forceinline__ void shared_add4(float* addr, float4 v) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(addr);
  float4 old = *reinterpret_cast<const float4*>(addr);
  while (true) {
    const uint64_t o0 = ((uint64_t)__float_as_uint(old.y) << 32) | __float_as_uint(old.x);
    const uint64_t o1 = ((uint64_t)__float_as_uint(old.w) << 32) | __float_as_uint(old.z);
    const uint64_t n0 = ((uint64_t)__float_as_uint(old.y + v.y) << 32) | __float_as_uint(old.x + v.x);
    const uint64_t n1 = ((uint64_t)__float_as_uint(old.w + v.w) << 32) | __float_as_uint(old.z + v.z);
    uint64_t r0, r1;
    asm volatile("{\n .reg .b128 c, n, r;\n mov.b128 c, {%2, %3};\n mov.b128 n, {%4, %5};\n"
                 " atom.shared.cas.b128 r, [%6], c, n;\n mov.b128 {%0, %1}, r;\n}"
                 : "=l"(r0), "=l"(r1) : "l"(o0), "l"(o1), "l"(n0), "l"(n1), "r"(a) : "memory");
    if (r0 == o0 && r1 == o1) break;
    old.x = __uint_as_float((uint32_t)r0); old.y = __uint_as_float((uint32_t)(r0 >> 32));
    old.z = __uint_as_float((uint32_t)r1); old.w = __uint_as_float((uint32_t)(r1 >> 32));
  }
}

// flush shared [K][d+4] sums + int counts into the packed [K][d+1] statistics of one subspace
__device__ __forceinline__ void flush_shared_stats(const float* s_sum, const int* s_cnt, int K, int d, float* pm) {
  const int ld = d + 1, ls = d + 4;
  for (int i = threadIdx.x; i < K * ld; i += blockDim.x) {
    const int k = i / ld, j = i - k * ld;
    const float v = (j < d) ? s_sum[k * ls + j] : (float)s_cnt[k];
    if (v != 0.f) atomicAdd(pm + i, v);
  }
}

template <int LPS, int kU>
__global__ void __launch_bounds__(1024)
accumulate_flat_kernel(const float* __restrict__ z, long long n_pixels, int D, int K,
                       const int32_t* __restrict__ idx, int use_norm, int mode,
                       const float* __restrict__ na, const float* __restrict__ nb,
                       float* __restrict__ packed, long long rows_per_block) {
  constexpr int d = LPS * 4;
  constexpr int ls = d + 4;
  extern __shared__ __align__(16) float s_acc[];  // [K][d+4] sums, then int [K] counts
  int* s_cnt = reinterpret_cast<int*>(s_acc + K * ls);
  const int m = blockIdx.y;
  for (int i = threadIdx.x; i < K * ls + K; i += blockDim.x) s_acc[i] = 0.f;   // 0.f and 0 share the bit pattern
  __syncthreads();
  const long long n0 = (long long)blockIdx.x * rows_per_block;
  long long n1 = n0 + rows_per_block;
  if (n1 > n_pixels) n1 = n_pixels;
  const int rows_per_iter = blockDim.x / LPS;
  const int rl = threadIdx.x / LPS, l = threadIdx.x % LPS;
  const int32_t* idxm = idx + (long long)m * n_pixels;
  // software pipeline: the loads of trip t+1 are issued before the shared updates of trip t (the CAS loop is an
  // asm volatile with a memory clobber, so the compiler will not do this itself)
  float4 vn[kU];
  int coden[kU];
  auto fetch = [&](long long nb0) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long nr = nb0 + (long long)u * rows_per_iter + rl;
      const long long n = nr < n1 ? nr : n1 - 1;
      vn[u] = __ldcs(reinterpret_cast<const float4*>(z + n * D + (long long)m * d) + l);
      coden[u] = __ldg(idxm + n);
    }
  };
  if (n0 < n1) fetch(n0);
  for (long long nb0 = n0; nb0 < n1; nb0 += (long long)kU * rows_per_iter) {   // block-uniform trip count (shuffles inside)
    float4 v[kU];
    int code[kU];
    bool live[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      v[u] = vn[u]; code[u] = coden[u];
      live[u] = nb0 + (long long)u * rows_per_iter + rl < n1;
    }
    if (nb0 + (long long)kU * rows_per_iter < n1) fetch(nb0 + (long long)kU * rows_per_iter);
    if (use_norm && mode != EQUSS_NORM_NONE) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        RowNorm r; r.shift = 0.f; r.denom = 1.f;
        if (mode == EQUSS_NORM_L2) {
          r = l2_from_sumsq(butterfly_lanes<LPS>(group_sumsq(v[u].x, v[u].y, v[u].z, v[u].w)));
        } else if (mode == EQUSS_NORM_ZNORM) {
          float mean = butterfly_lanes<LPS>(group_sum(v[u].x, v[u].y, v[u].z, v[u].w)) / (float)d;
          float ssd = butterfly_lanes<LPS>(group_sumsq(v[u].x - mean, v[u].y - mean, v[u].z - mean, v[u].w - mean));
          r.shift = mean; r.denom = sqrtf(ssd / (float)(d - 1)) + kStdEps;
        }
        int ch = m * d + l * 4;
        v[u].x = norm_elem(v[u].x, r, mode, na, nb, ch);
        v[u].y = norm_elem(v[u].y, r, mode, na, nb, ch + 1);
        v[u].z = norm_elem(v[u].z, r, mode, na, nb, ch + 2);
        v[u].w = norm_elem(v[u].w, r, mode, na, nb, ch + 3);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (!live[u]) continue;
      shared_add4(s_acc + code[u] * ls + l * 4, v[u]);
      if (l == 0) atomicAdd(s_cnt + code[u], 1);
    }
  }
  __syncthreads();
  flush_shared_stats(s_acc, s_cnt, K, d, packed + (long long)m * K * (d + 1));
}

// Strided (NCHW) rows: one pixel per thread, the d values of the row in registers, d/4 128-bit shared updates.
template <int DT>
__global__ void __launch_bounds__(DT >= 32 ? 512 : 1024)
accumulate_rows_kernel(const float* __restrict__ z, ZView zv, int K,
                       const int32_t* __restrict__ idx, int use_norm, int mode,
                       const float* __restrict__ na, const float* __restrict__ nb,
                       float* __restrict__ packed, long long rows_per_block) {
  constexpr int d = DT, ls = DT + 4;
  extern __shared__ __align__(16) float s_acc[];
  int* s_cnt = reinterpret_cast<int*>(s_acc + K * ls);
  const int m = blockIdx.y;
  for (int i = threadIdx.x; i < K * ls + K; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const long long n0 = (long long)blockIdx.x * rows_per_block;
  long long n1 = n0 + rows_per_block;
  if (n1 > zv.n_pixels) n1 = zv.n_pixels;
  const int nm = use_norm ? mode : EQUSS_NORM_NONE;
  // the next pixel's row is loaded before this pixel's shared updates (see accumulate_flat_kernel)
  ScalarRow<DT> nxt;
  int code_n = 0;
  auto fetch = [&](long long n) {
    nxt.load(z + pixel_base(zv, n) + (long long)m * d * zv.stride_c, zv.stride_c, d);
    code_n = __ldg(idx + (long long)m * zv.n_pixels + n);
  };
  constexpr bool kPipe = DT <= 32;                  // two rows of 64 floats do not fit the register file
  if (kPipe && n0 + threadIdx.x < n1) fetch(n0 + threadIdx.x);
  for (long long n = n0 + threadIdx.x; n < n1; n += blockDim.x) {
    if (!kPipe) fetch(n);
    ScalarRow<DT> row = nxt;
    const int code = code_n;
    if (kPipe && n + blockDim.x < n1) fetch(n + blockDim.x);
    RowNorm r; r.shift = 0.f; r.denom = 1.f;
    if (nm != EQUSS_NORM_NONE) r = row.norm(nm);
    float* a = s_acc + code * ls;
#pragma unroll
    for (int q = 0; q < DT / 4; ++q) {
      float4 v;
      v.x = norm_elem(row.raw(4 * q + 0), r, nm, na, nb, m * d + 4 * q + 0);
      v.y = norm_elem(row.raw(4 * q + 1), r, nm, na, nb, m * d + 4 * q + 1);
      v.z = norm_elem(row.raw(4 * q + 2), r, nm, na, nb, m * d + 4 * q + 2);
      v.w = norm_elem(row.raw(4 * q + 3), r, nm, na, nb, m * d + 4 * q + 3);
      shared_add4(a + 4 * q, v);
    }
    atomicAdd(s_cnt + code, 1);
  }
  __syncthreads();
  flush_shared_stats(s_acc, s_cnt, K, d, packed + (long long)m * K * (d + 1));
}

template <int DT>
__global__ void __launch_bounds__(128)
accumulate_scalar_kernel(const float* __restrict__ z, ZView zv, int K, int d,
                         const int32_t* __restrict__ idx, int use_norm, int mode,
                         const float* __restrict__ na, const float* __restrict__ nb,
                         float* __restrict__ packed, long long rows_per_block, int use_smem) {
  extern __shared__ float s_acc[];  // [K][d+1] when use_smem
  const int m = blockIdx.y;
  const int ld = d + 1;
  float* pm = packed + (long long)m * K * ld;
  if (use_smem) {
    for (int i = threadIdx.x; i < K * ld; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
  }
  float* acc = use_smem ? s_acc : pm;
  const long long n0 = (long long)blockIdx.x * rows_per_block;
  long long n1 = n0 + rows_per_block;
  if (n1 > zv.n_pixels) n1 = zv.n_pixels;
  const int dd = DT > 0 ? DT : d;
  for (long long n = n0 + threadIdx.x; n < n1; n += blockDim.x) {
    long long base = pixel_base(zv, n) + (long long)m * d * zv.stride_c;
    ScalarRow<DT> row;
    row.load(z + base, zv.stride_c, d);
    RowNorm r; r.shift = 0.f; r.denom = 1.f;
    const int nm = use_norm ? mode : EQUSS_NORM_NONE;
    if (nm != EQUSS_NORM_NONE) r = row.norm(nm);
    int code = __ldg(idx + (long long)m * zv.n_pixels + n);
    float* a = acc + (long long)code * ld;
#pragma unroll 4
    for (int j = 0; j < dd; ++j) atomicAdd(a + j, norm_elem(row.raw(j), r, nm, na, nb, m * d + j));
    atomicAdd(a + d, 1.f);
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < K * ld; i += blockDim.x) {
      float v = s_acc[i];
      if (v != 0.f) atomicAdd(pm + i, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K6 EMA update: one block per subspace.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
ema_update_kernel(const float* __restrict__ packed, int K, int d, float decay, float alpha, float eps,
                  float k_eps, float* __restrict__ vq_count, float* __restrict__ weight_avg,
                  float* __restrict__ weight, float* __restrict__ exact_count,
                  int32_t* __restrict__ unused_out) {
  const int m = blockIdx.x;
  const int ld = d + 1;
  const float* pm = packed + (long long)m * K * ld;
  float* cnt = vq_count + (long long)m * K;
  __shared__ float s_red[32];
  __shared__ int s_unused[32];
  __shared__ float s_n;
  // vq_count.mul_(decay).add_(count, alpha=1-decay)   (model/quantizer.py:242)
  float part = 0.f;
  int unused = 0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float c = pm[(long long)k * ld + d];
    float v = cnt[k] * decay;
    v = v + alpha * c;
    cnt[k] = v;
    part += v;
    if (exact_count) exact_count[(long long)m * K + k] += c;
    unused += (c == 0.f);
  }
  part = warp_sum(part);
  for (int o = 16; o > 0; o >>= 1) unused += __shfl_xor_sync(0xffffffffu, unused, o);
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = part; s_unused[threadIdx.x >> 5] = unused; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f; int u = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { t += s_red[i]; u += s_unused[i]; }
    s_n = t;
    if (unused_out) unused_out[m] = u;
  }
  __syncthreads();
  const float n = s_n;
  const float denom_n = n + k_eps;   // n + num_codebook * eps, the product formed in double (:250)
  // weight_avg.mul_(decay).add_(sum, alpha=1-decay); weight = weight_avg / smoothed   (:245-254)
  for (int i = threadIdx.x; i < K * d; i += blockDim.x) {
    int k = i / d, j = i - k * d;
    long long o = (long long)m * K * d + i;
    float a = weight_avg[o] * decay;
    a = a + alpha * pm[(long long)k * ld + j];
    weight_avg[o] = a;
    float smoothed = (cnt[k] + eps) / denom_n * n;
    weight[o] = a / smoothed;
  }
}

// cnorm2[m][k] = sum_j c^2 (canonical order)
// ------------------------------------------------------------------------------------------------
// K7 usage percentiles (get_histogram_count, model/quantizer.py:15-30), one block per subspace:
// prob = count / (sum + 1), sorted descending (bitonic sort in shared memory), sequential cumulative sum,
// first rank whose cumulative usage reaches 10 / 50 / 90 %, divided by K; NaN where the reference returns None.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
usage_percentiles_kernel(const float* __restrict__ count, long long row_stride, long long k_stride, int K, int Kp,
                         float* __restrict__ out) {
  extern __shared__ float s_p[];     // [Kp]
  __shared__ float s_tot[8];
  const int m = blockIdx.x;
  const float* c = count + (long long)m * row_stride;
  float part = 0.f;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    const float v = (k < K) ? c[(long long)k * k_stride] : -1.f;      // padding sorts to the end
    s_p[k] = v;
    if (k < K) part += v;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_tot[threadIdx.x >> 5] = part;
  __syncthreads();
  float total = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += s_tot[i];
  for (int size = 2; size <= Kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < Kp / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const float a = s_p[lo], b = s_p[hi];
        if (desc ? (a < b) : (a > b)) { s_p[lo] = b; s_p[hi] = a; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float denom = total + 1.f;
    // torch.cumsum on the reference's CPU path accumulates fp32 inputs in double and rounds each prefix to fp32
    double acc = 0.0;
    float p10 = NAN, p50 = NAN, p90 = NAN;
    bool f10 = false, f50 = false, f90 = false;
    for (int i = 0; i < K; ++i) {
      acc += (double)(s_p[i] / denom);
      const float cs = (float)acc;
      if (!f10 && cs >= 0.1f) { f10 = true; p10 = (float)i / (float)K; }
      if (!f50 && cs >= 0.5f) { f50 = true; p50 = (float)i / (float)K; }
      if (!f90 && cs >= 0.9f) { f90 = true; p90 = (float)i / (float)K; break; }
    }
    out[m * 3 + 0] = p10; out[m * 3 + 1] = p50; out[m * 3 + 2] = p90;
  }
}

// K <= 1024: rank sort (every thread counts the values ahead of its own: K broadcast reads) and a parallel prefix sum
// in double instead of the 36 barrier-separated bitonic stages and the serial scan above (30 us -> ~3 us at K = 256).
// The prefix is summed in double like torch's CPU cumsum; double addition of fp32 quotients in [0, 1] is exact unless a
// term lies more than 2^29 below the running sum, so the tree order agrees with the sequential one.
constexpr int kPctThreads = 256, kPctItems = 4;
// block-wide (kPctThreads threads); out3 receives p10 / p50 / p90 from threads 0..2
__device__ __forceinline__ void block_usage_percentiles(const float* __restrict__ c, long long k_stride, int K,
                                                        float* __restrict__ out3) {
  __shared__ float s_v[kPctThreads * kPctItems];
  __shared__ float s_sorted[kPctThreads * kPctItems];
  __shared__ float s_tot[kPctThreads / 32];
  __shared__ double s_wsum[kPctThreads / 32];
  __shared__ int s_first[3];
  __syncthreads();                                        // a previous call's readers are done
  float part = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = c[(long long)k * k_stride];
    s_v[k] = v;
    part += v;
  }
  if (threadIdx.x < 3) s_first[threadIdx.x] = K;
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_tot[threadIdx.x >> 5] = part;
  __syncthreads();
  float total = 0.f;
  for (int i = 0; i < kPctThreads / 32; ++i) total += s_tot[i];
  const float denom = total + 1.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {     // descending, ties by position
    const float v = s_v[k];
    int rank = 0;
    for (int j = 0; j < K; ++j) {
      const float u = s_v[j];
      rank += (u > v) || (u == v && j < k);
    }
    s_sorted[rank] = v;
  }
  __syncthreads();
  // thread t owns sorted ranks [t*kPctItems, (t+1)*kPctItems)
  const int i0 = threadIdx.x * kPctItems;
  double loc[kPctItems];
  double run = 0.0;
#pragma unroll
  for (int e = 0; e < kPctItems; ++e) {
    const int i = i0 + e;
    run += (i < K) ? (double)(s_sorted[i] / denom) : 0.0;
    loc[e] = run;
  }
  double incl = run;                                      // inclusive scan of the per-thread totals
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wsum[wid] = incl;
  __syncthreads();
  double base = incl - run;
  for (int w2 = 0; w2 < wid; ++w2) base += s_wsum[w2];
#pragma unroll
  for (int e = 0; e < kPctItems; ++e) {
    const int i = i0 + e;
    if (i < K) {
      const float cs = (float)(base + loc[e]);
      if (cs >= 0.1f) atomicMin(&s_first[0], i);
      if (cs >= 0.5f) atomicMin(&s_first[1], i);
      if (cs >= 0.9f) atomicMin(&s_first[2], i);
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int f = s_first[threadIdx.x];
    out3[threadIdx.x] = (f < K) ? (float)f / (float)K : NAN;
  }
}

__global__ void __launch_bounds__(kPctThreads)
usage_percentiles_small_kernel(const float* __restrict__ count, long long row_stride, long long k_stride, int K,
                               float* __restrict__ out) {
  block_usage_percentiles(count + (long long)blockIdx.x * row_stride, k_stride, K, out + blockIdx.x * 3);
}

// ------------------------------------------------------------------------------------------------
// Train-step tail (K6 + K7 + the scalar outputs of EMAVectorQuantizer.forward, model/quantizer.py:493-532) in ONE
// launch: per subspace (one block) the EMA update, the exact-count accumulation, both usage-percentile triples, the
// unused-code count, sum |weight| and the commitment MSE; the last block to finish averages them over the subspaces
// (ProductQuantizerWrapper.forward's mean, :607-608).  Replaces ~15 small launches of the eager path.
//   scratch: [M][kTailStats] floats + one unsigned counter (zero on entry, zero again on exit)
//   stats_out: [kTailStats] = total-p10/50/90, current-p10/50/90, codebook-usage, codebook-sum, commitment-loss, loss
// ------------------------------------------------------------------------------------------------
constexpr int kTailStats = 10;
__global__ void __launch_bounds__(kPctThreads)
ema_train_tail_kernel(const float* __restrict__ packed, int M, int K, int d, float decay, float alpha, float eps,
                      float k_eps, float* __restrict__ vq_count, float* __restrict__ weight_avg,
                      float* __restrict__ weight, float* __restrict__ exact_count, const double* __restrict__ sqerr,
                      double inv_nd, float beta, float* __restrict__ scratch, float* __restrict__ stats_out,
                      const float* const* __restrict__ peers, int world, float* __restrict__ packed_out,
                      float* __restrict__ zero_next) {
  const int m = blockIdx.x;
  const int ld = d + 1;
  if (zero_next != nullptr) {
    // the OTHER symmetric buffer (next step's accumulation target) is re-zeroed here: every peer finished reading it
    // before it entered the barrier that precedes this launch
    float* zn = zero_next + (long long)m * K * ld;
    for (int i = threadIdx.x; i < K * ld; i += blockDim.x) zn[i] = 0.f;
  }
  if (peers != nullptr) {
    // K5 fused in: the data-parallel sum of the packed statistics is taken straight from the peers' symmetric
    // buffers over NVLink (one pass of P2P loads, ranks added in a fixed order so every replica computes the same
    // bits) instead of a separate all-reduce; the reduced slice lands in packed_out, which the rest of the kernel
    // (and the caller, for restart / split) reads.
    const long long off = (long long)m * K * ld;
    const int cnt = K * ld;
    if ((cnt & 3) == 0) {
      for (int i = threadIdx.x * 4; i < cnt; i += blockDim.x * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
          const float4 v = __ldcv(reinterpret_cast<const float4*>(peers[r] + off + i));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(packed_out + off + i) = acc;
      }
    } else {
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < world; ++r) acc += __ldcv(peers[r] + off + i);
        packed_out[off + i] = acc;
      }
    }
    __syncthreads();
    packed = packed_out;
  }
  const float* pm = packed + (long long)m * K * ld;
  float* cnt = vq_count + (long long)m * K;
  float* ex = exact_count + (long long)m * K;
  __shared__ float s_red[kPctThreads / 32];
  __shared__ int s_unused[kPctThreads / 32];
  __shared__ float s_n, s_abs;
  __shared__ int s_unused_total;
  __shared__ bool s_last;
  float part = 0.f;
  int unused = 0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float c = pm[(long long)k * ld + d];
    float v = cnt[k] * decay;
    v = v + alpha * c;
    cnt[k] = v;
    part += v;
    ex[k] += c;
    unused += (c == 0.f);
  }
  part = warp_sum(part);
  for (int o = 16; o > 0; o >>= 1) unused += __shfl_xor_sync(0xffffffffu, unused, o);
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = part; s_unused[threadIdx.x >> 5] = unused; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f; int u = 0;
    for (int i = 0; i < kPctThreads / 32; ++i) { t += s_red[i]; u += s_unused[i]; }
    s_n = t; s_unused_total = u;
  }
  __syncthreads();
  const float n = s_n;
  const float denom_n = n + k_eps;
  float asum = 0.f;
  for (int i = threadIdx.x; i < K * d; i += blockDim.x) {
    const int k = i / d, j = i - k * d;
    const long long o = (long long)m * K * d + i;
    float a = weight_avg[o] * decay;
    a = a + alpha * pm[(long long)k * ld + j];
    weight_avg[o] = a;
    const float smoothed = (cnt[k] + eps) / denom_n * n;
    const float wv = a / smoothed;
    weight[o] = wv;
    asum += fabsf(wv);
  }
  asum = warp_sum(asum);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = asum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kPctThreads / 32; ++i) t += s_red[i];
    s_abs = t;
  }
  float* sm = scratch + (long long)m * kTailStats;
  block_usage_percentiles(ex, 1, K, sm);                          // "total"   (:496)
  block_usage_percentiles(pm + d, ld, K, sm + 3);                 // "current" (:495)
  if (threadIdx.x == 0) {
    sm[6] = (float)(K - s_unused_total) / (float)K;               // :510
    sm[7] = s_abs;                                                // :532
    sm[8] = sqerr ? (float)(sqerr[m] * inv_nd) : 0.f;             // :514
    sm[9] = 0.f;
  }
  __threadfence();
  __syncthreads();
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + (long long)M * kTailStats);
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == (unsigned)(M - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 9) {
    float t = 0.f;
    for (int i = 0; i < M; ++i) t += __ldcg(scratch + (long long)i * kTailStats + threadIdx.x);
    t /= (float)M;
    stats_out[threadIdx.x] = t;
    if (threadIdx.x == 8) stats_out[9] = beta * t;               // :526
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// codebook-side normalisation + |c|^2 in one launch (model/quantizer.py:421,426,459): one thread per code
__global__ void __launch_bounds__(256)
codebook_prepare_kernel(const float* __restrict__ cb, long long rows, int d, int mode, float* __restrict__ cbn,
                        float* __restrict__ cn2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* c = cb + i * d;
  float* o = cbn + i * d;
  const RowNorm r = row_norm_generic(mode, d, [&](int j) { return __ldg(c + j); });
  for (int j = 0; j < d; ++j) o[j] = apply_norm(__ldg(c + j), r, mode);
  cn2[i] = canonical_sumsq(d, [&](int j) { return o[j]; });
}

__global__ void cnorm2_kernel(const float* __restrict__ cb, long long rows, int d, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* c = cb + i * d;
  out[i] = canonical_sumsq(d, [&](int j) { return __ldg(c + j); });
}

// ------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------
static inline bool flat_vector_ok(const equss_zdesc* zd, int d, const void* p0, const void* p1) {
  if (zd->stride_c != 1 || zd->stride_s != zd->dim || zd->stride_b != zd->hw * (int64_t)zd->dim) return false;
  if (d % 4 != 0) return false;
  int lps = d / 4;
  if (lps > 32 || (lps & (lps - 1)) != 0) return false;
  if (((uintptr_t)p0 & 15) || ((uintptr_t)p1 & 15)) return false;
  return true;
}

#define EQUSS_DISPATCH_LPS(lps, ...)       \
  switch (lps) {                            \
    case 1: { constexpr int LPS = 1; __VA_ARGS__; break; }   \
    case 2: { constexpr int LPS = 2; __VA_ARGS__; break; }   \
    case 4: { constexpr int LPS = 4; __VA_ARGS__; break; }   \
    case 8: { constexpr int LPS = 8; __VA_ARGS__; break; }   \
    case 16: { constexpr int LPS = 16; __VA_ARGS__; break; } \
    default: { constexpr int LPS = 32; __VA_ARGS__; break; } \
  }
#define EQUSS_DISPATCH_DT(d, ...)          \
  switch (d) {                              \
    case 4: { constexpr int DT = 4; __VA_ARGS__; break; }    \
    case 8: { constexpr int DT = 8; __VA_ARGS__; break; }    \
    case 16: { constexpr int DT = 16; __VA_ARGS__; break; }  \
    case 32: { constexpr int DT = 32; __VA_ARGS__; break; }  \
    case 64: { constexpr int DT = 64; __VA_ARGS__; break; }  \
    default: { constexpr int DT = 0; __VA_ARGS__; break; }   \
  }

static int check_norm_args(int mode, const float* na, const float* nb) {
  EQUSS_REQUIRE(mode >= EQUSS_NORM_NONE && mode <= EQUSS_NORM_AFFINE, EQUSS_ERR_INVALID_ARG,
                "Unsupported normalize type %d", mode);
  EQUSS_REQUIRE(mode != EQUSS_NORM_AFFINE || (na && nb), EQUSS_ERR_INVALID_ARG,
                "EQUSS_NORM_AFFINE needs norm_a and norm_b");
  return EQUSS_OK;
}

}  // namespace equss

using namespace equss;

extern "C" int equss_pq_cnorm2(const float* codebook_norm, int M, int K, int d, float* cnorm2, void* stream) {
  EQUSS_REQUIRE(codebook_norm && cnorm2 && M > 0 && K > 0 && d > 0 && d <= kMaxD, EQUSS_ERR_INVALID_ARG,
                "equss_pq_cnorm2: bad arguments (M=%d K=%d d=%d)", M, K, d);
  long long rows = (long long)M * K;
  cnorm2_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(codebook_norm, rows, d, cnorm2);
  EQUSS_LAUNCH_OK("cnorm2_kernel");
  return EQUSS_OK;
}

extern "C" int equss_pq_gather_loss(const float* z, const equss_zdesc* zd, const float* gather_src,
                                    const int32_t* idx, int M, int K, int d, int norm_mode,
                                    const float* norm_a, const float* norm_b, float* out,
                                    float* znorm_out, double* sqerr, void* stream) {
  if (zd && zd->n_pixels == 0) return EQUSS_OK;   // empty tensors have null data pointers
  EQUSS_REQUIRE(z && zd && gather_src && idx && out && sqerr, EQUSS_ERR_INVALID_ARG,
                "equss_pq_gather_loss: null pointer");
  int rc = validate_zdesc(zd, M, d); if (rc) return rc;
  rc = check_norm_args(norm_mode, norm_a, norm_b); if (rc) return rc;
  EQUSS_REQUIRE(K > 0, EQUSS_ERR_INVALID_ARG, "equss_pq_gather_loss: K=%d", K);
  if (zd->n_pixels == 0) return EQUSS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = flat_vector_ok(zd, d, z, out) && !((uintptr_t)gather_src & 15) &&
                   (!znorm_out || !((uintptr_t)znorm_out & 15));
  if (vec) {
    int lps = d / 4;
    long long total = zd->n_pixels * (zd->dim / 4);
    int blocks = (int)((total + 255) / 256);
    int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (norm_mode == EQUSS_NORM_L2 && !znorm_out) {
      EQUSS_DISPATCH_LPS(lps, (gather_loss_flat_l2_kernel<LPS><<<blocks, 256, M * sizeof(float), st>>>(
          reinterpret_cast<const float4*>(z), zd->n_pixels, zd->dim / 4, M, K,
          reinterpret_cast<const float4*>(gather_src), idx, reinterpret_cast<float4*>(out), sqerr)));
      EQUSS_LAUNCH_OK("gather_loss_flat_l2_kernel");
      return EQUSS_OK;
    }
    EQUSS_DISPATCH_LPS(lps, (gather_loss_flat_kernel<LPS><<<blocks, 256, M * sizeof(float), st>>>(
        reinterpret_cast<const float4*>(z), zd->n_pixels, zd->dim / 4, M, K,
        reinterpret_cast<const float4*>(gather_src), idx, norm_mode, norm_a, norm_b,
        reinterpret_cast<float4*>(out), reinterpret_cast<float4*>(znorm_out), sqerr)));
    EQUSS_LAUNCH_OK("gather_loss_flat_kernel");
  } else if (norm_mode == EQUSS_NORM_L2 && !znorm_out && (d == 16 || d == 32 || d == 64) && zd->stride_s == 1 &&
             zd->stride_c == zd->hw && zd->n_pixels % zd->hw == 0 && !((uintptr_t)gather_src & 15)) {
    const int n_images = (int)(zd->n_pixels / zd->hw);
    long long tiles = (long long)n_images * ((zd->hw + 127) / 128);
    long long cap = (long long)num_sms() * 32 / M + 1;
    dim3 grid((unsigned)(tiles < cap ? tiles : cap), (unsigned)M);
    if (d == 16) gather_loss_nchw_l2_kernel<16><<<grid, 128, 0, st>>>(z, zd->hw, zd->stride_b, n_images, M, K, gather_src, idx, out, sqerr);
    else if (d == 32) gather_loss_nchw_l2_kernel<32><<<grid, 128, 0, st>>>(z, zd->hw, zd->stride_b, n_images, M, K, gather_src, idx, out, sqerr);
    else gather_loss_nchw_l2_kernel<64><<<grid, 128, 0, st>>>(z, zd->hw, zd->stride_b, n_images, M, K, gather_src, idx, out, sqerr);
    EQUSS_LAUNCH_OK("gather_loss_nchw_l2_kernel");
  } else {
    ZView zv = make_view(zd);
    long long bx = (zd->n_pixels + 127) / 128;
    long long cap = (long long)num_sms() * 16 / M + 1;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)M);
    EQUSS_DISPATCH_DT(d, (gather_loss_scalar_kernel<DT><<<grid, 128, 0, st>>>(
        z, zv, M, K, d, gather_src, idx, norm_mode, norm_a, norm_b, out, znorm_out, sqerr)));
    EQUSS_LAUNCH_OK("gather_loss_scalar_kernel");
  }
  return EQUSS_OK;
}

extern "C" int equss_pq_gather_loss_bwd(const float* z, const equss_zdesc* zd, const float* gather_src,
                                        const int32_t* idx, int M, int K, int d, int norm_mode,
                                        const float* norm_a, const float* norm_b,
                                        const float* grad_out, const float* coef, float* grad_z,
                                        const float* cb_coef, float* grad_codebook, void* stream) {
  if (zd && zd->n_pixels == 0) return EQUSS_OK;
  EQUSS_REQUIRE(z && zd && gather_src && idx, EQUSS_ERR_INVALID_ARG, "equss_pq_gather_loss_bwd: null pointer");
  EQUSS_REQUIRE(grad_z || grad_codebook, EQUSS_ERR_INVALID_ARG, "equss_pq_gather_loss_bwd: nothing to compute");
  int rc = validate_zdesc(zd, M, d); if (rc) return rc;
  rc = check_norm_args(norm_mode, norm_a, norm_b); if (rc) return rc;
  if (zd->n_pixels == 0) return EQUSS_OK;
  ZView zv = make_view(zd);
  long long bx = (zd->n_pixels + 127) / 128;
  long long cap = (long long)num_sms() * 16 / M + 1;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)M);
  cudaStream_t st = (cudaStream_t)stream;
  EQUSS_DISPATCH_DT(d, (gather_loss_bwd_kernel<DT><<<grid, 128, 0, st>>>(
      z, zv, M, K, d, gather_src, idx, norm_mode, norm_a, norm_b, grad_out, coef, grad_z, cb_coef,
      grad_codebook)));
  EQUSS_LAUNCH_OK("gather_loss_bwd_kernel");
  return EQUSS_OK;
}

extern "C" int equss_pq_accumulate(const float* z, const equss_zdesc* zd, const int32_t* idx, int M, int K,
                                   int d, int use_norm, int norm_mode, const float* norm_a,
                                   const float* norm_b, float* packed, void* stream) {
  if (zd && zd->n_pixels == 0) return EQUSS_OK;
  EQUSS_REQUIRE(z && zd && idx && packed, EQUSS_ERR_INVALID_ARG, "equss_pq_accumulate: null pointer");
  int rc = validate_zdesc(zd, M, d); if (rc) return rc;
  rc = check_norm_args(norm_mode, norm_a, norm_b); if (rc) return rc;
  EQUSS_REQUIRE(K > 0, EQUSS_ERR_INVALID_ARG, "equss_pq_accumulate: K=%d", K);
  if (zd->n_pixels == 0) return EQUSS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int bps = 8;     // blocks per SM over the whole grid (sweep on B200: 4 / 8 / 12 -> 75 / 73 / 77 us at C2)
  const bool quad = (d % 4) == 0 && (d == 4 || d == 8 || d == 16 || d == 32 || d == 64);
  // 128-bit path: [K][d+4] sums + [K] int counts; scalar path: [K][d+1] floats
  const size_t smem = quad ? (size_t)K * (d + 5) * sizeof(float) : (size_t)K * (d + 1) * sizeof(float);
  const bool smem_ok = smem <= 200 * 1024;
  // chunking: `bps` blocks per SM in total while the accumulators of several blocks fit one SM; one big block otherwise
  const bool big = smem > 56 * 1024;
  const int threads = big ? 1024 : 256;
  long long chunks = ((long long)num_sms() * (big ? 1 : bps) + M - 1) / M;
  if (chunks < 1) chunks = 1;
  long long rows_per_block = (zd->n_pixels + chunks - 1) / chunks;
  if (rows_per_block < 256) rows_per_block = 256;
  chunks = (zd->n_pixels + rows_per_block - 1) / rows_per_block;
  dim3 grid((unsigned)chunks, (unsigned)M);
  if (quad && flat_vector_ok(zd, d, z, z) && smem_ok) {
    int lps = d / 4;
    EQUSS_DISPATCH_LPS(lps, {
      auto launch = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, st>>>(z, zd->n_pixels, zd->dim, K, idx, use_norm, norm_mode, norm_a, norm_b,
                                          packed, rows_per_block);
      };
      launch(accumulate_flat_kernel<LPS, 2>);   // two rows per thread per trip (1 / 2 / 4 measured equal within noise)
    });
    EQUSS_LAUNCH_OK("accumulate_flat_kernel");
  } else if (quad && smem_ok) {
    ZView zv = make_view(zd);
    EQUSS_DISPATCH_DT(d, {
      if constexpr (DT >= 4) {
        if (smem > 48 * 1024)
          EQUSS_CUDA_OK(cudaFuncSetAttribute(accumulate_rows_kernel<DT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        accumulate_rows_kernel<DT><<<grid, (DT >= 32 && threads > 512) ? 512 : threads, smem, st>>>(z, zv, K, idx, use_norm, norm_mode, norm_a, norm_b,
                                                                packed, rows_per_block);
      }
    });
    EQUSS_LAUNCH_OK("accumulate_rows_kernel");
  } else {
    ZView zv = make_view(zd);
    EQUSS_DISPATCH_DT(d, {
      if (smem_ok && smem > 48 * 1024)
        EQUSS_CUDA_OK(cudaFuncSetAttribute(accumulate_scalar_kernel<DT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      accumulate_scalar_kernel<DT><<<grid, 128, smem_ok ? smem : 0, st>>>(
          z, zv, K, d, idx, use_norm, norm_mode, norm_a, norm_b, packed, rows_per_block, smem_ok ? 1 : 0);
    });
    EQUSS_LAUNCH_OK("accumulate_scalar_kernel");
  }
  return EQUSS_OK;
}

extern "C" int equss_ema_update(const float* packed, int M, int K, int d, double decay, double eps,
                                float* vq_count, float* weight_avg, float* weight, float* exact_count,
                                int32_t* unused_out, void* stream) {
  EQUSS_REQUIRE(packed && vq_count && weight_avg && weight, EQUSS_ERR_INVALID_ARG, "equss_ema_update: null pointer");
  EQUSS_REQUIRE(M > 0 && K > 0 && d > 0, EQUSS_ERR_INVALID_ARG, "equss_ema_update: bad shape M=%d K=%d d=%d", M, K, d);
  // Python scalars are doubles: `1 - decay` and `K * eps` are formed in double, then cast to fp32 by torch.
  ema_update_kernel<<<M, 1024, 0, (cudaStream_t)stream>>>(packed, K, d, (float)decay, (float)(1.0 - decay),
                                                          (float)eps, (float)((double)K * eps), vq_count,
                                                          weight_avg, weight, exact_count, unused_out);
  EQUSS_LAUNCH_OK("ema_update_kernel");
  return EQUSS_OK;
}

extern "C" int equss_usage_percentiles(const float* count, int64_t row_stride, int64_t k_stride, int M, int K,
                                       float* out, void* stream) {
  EQUSS_REQUIRE(M >= 0 && K > 0 && K <= 8192, EQUSS_ERR_INVALID_ARG, "equss_usage_percentiles: bad shape M=%d K=%d", M, K);
  if (M == 0) return EQUSS_OK;
  EQUSS_REQUIRE(count && out, EQUSS_ERR_INVALID_ARG, "equss_usage_percentiles: null pointer");
  if (K <= equss::kPctThreads * equss::kPctItems) {
    equss::usage_percentiles_small_kernel<<<M, equss::kPctThreads, 0, (cudaStream_t)stream>>>(count, row_stride, k_stride, K, out);
    EQUSS_LAUNCH_OK("usage_percentiles_small_kernel");
    return EQUSS_OK;
  }
  int Kp = 2;
  while (Kp < K) Kp <<= 1;
  equss::usage_percentiles_kernel<<<M, 256, Kp * sizeof(float), (cudaStream_t)stream>>>(count, row_stride, k_stride, K, Kp, out);
  EQUSS_LAUNCH_OK("usage_percentiles_kernel");
  return EQUSS_OK;
}

extern "C" int equss_pq_train_tail_scratch_floats(int M) { return M * equss::kTailStats + 4; }

extern "C" int equss_pq_train_tail(const float* packed, int M, int K, int d, double decay, double eps, float* vq_count,
                                   float* weight_avg, float* weight, float* exact_count, const double* sqerr,
                                   int64_t n_pixels, double beta, float* scratch, float* stats_out, void* stream) {
  EQUSS_REQUIRE(packed && vq_count && weight_avg && weight && exact_count && scratch && stats_out, EQUSS_ERR_INVALID_ARG,
                "equss_pq_train_tail: null pointer");
  EQUSS_REQUIRE(M > 0 && K > 0 && d > 0, EQUSS_ERR_INVALID_ARG, "equss_pq_train_tail: bad shape M=%d K=%d d=%d", M, K, d);
  EQUSS_REQUIRE(K <= equss::kPctThreads * equss::kPctItems, EQUSS_ERR_UNSUPPORTED,
                "equss_pq_train_tail: K=%d > %d; use equss_ema_update + equss_usage_percentiles", K,
                equss::kPctThreads * equss::kPctItems);
  const double inv_nd = (n_pixels > 0) ? 1.0 / ((double)n_pixels * (double)d) : 0.0;
  equss::ema_train_tail_kernel<<<M, equss::kPctThreads, 0, (cudaStream_t)stream>>>(
      packed, M, K, d, (float)decay, (float)(1.0 - decay), (float)eps, (float)((double)K * eps), vq_count, weight_avg, weight,
      exact_count, sqerr, inv_nd, (float)beta, scratch, stats_out, nullptr, 1, nullptr, nullptr);
  EQUSS_LAUNCH_OK("ema_train_tail_kernel");
  return EQUSS_OK;
}

extern "C" int equss_pq_train_tail_peers(const void* const* peer_packed, int world, float* packed_out, int M, int K, int d,
                                         double decay, double eps, float* vq_count, float* weight_avg, float* weight,
                                         float* exact_count, const double* sqerr, int64_t n_pixels, double beta,
                                         float* scratch, float* stats_out, float* zero_next, void* stream) {
  EQUSS_REQUIRE(peer_packed && packed_out && vq_count && weight_avg && weight && exact_count && scratch && stats_out,
                EQUSS_ERR_INVALID_ARG, "equss_pq_train_tail_peers: null pointer");
  EQUSS_REQUIRE(world >= 1 && world <= 64, EQUSS_ERR_INVALID_ARG, "equss_pq_train_tail_peers: world=%d", world);
  EQUSS_REQUIRE(M > 0 && K > 0 && d > 0, EQUSS_ERR_INVALID_ARG, "equss_pq_train_tail_peers: bad shape M=%d K=%d d=%d", M, K, d);
  EQUSS_REQUIRE(K <= equss::kPctThreads * equss::kPctItems, EQUSS_ERR_UNSUPPORTED, "equss_pq_train_tail_peers: K=%d > %d", K,
                equss::kPctThreads * equss::kPctItems);
  const double inv_nd = (n_pixels > 0) ? 1.0 / ((double)n_pixels * (double)d) : 0.0;
  equss::ema_train_tail_kernel<<<M, equss::kPctThreads, 0, (cudaStream_t)stream>>>(
      packed_out, M, K, d, (float)decay, (float)(1.0 - decay), (float)eps, (float)((double)K * eps), vq_count, weight_avg, weight,
      exact_count, sqerr, inv_nd, (float)beta, scratch, stats_out, (const float* const*)peer_packed, world, packed_out, zero_next);
  EQUSS_LAUNCH_OK("ema_train_tail_kernel<peers>");
  return EQUSS_OK;
}

extern "C" int equss_pq_prepare_codebook(const float* codebook, int M, int K, int d, int norm_mode, float* codebook_norm,
                                         float* cnorm2, void* stream) {
  EQUSS_REQUIRE(codebook && codebook_norm && cnorm2, EQUSS_ERR_INVALID_ARG, "equss_pq_prepare_codebook: null pointer");
  EQUSS_REQUIRE(M > 0 && K > 0 && d > 0 && d <= equss::kMaxD, EQUSS_ERR_INVALID_ARG, "equss_pq_prepare_codebook: bad shape");
  EQUSS_REQUIRE(norm_mode == EQUSS_NORM_NONE || norm_mode == EQUSS_NORM_L2 || norm_mode == EQUSS_NORM_ZNORM, EQUSS_ERR_UNSUPPORTED,
                "equss_pq_prepare_codebook: per-code modes only (none, l2, z_norm); got %d", norm_mode);
  const long long rows = (long long)M * K;
  equss::codebook_prepare_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(codebook, rows, d, norm_mode,
                                                                                                  codebook_norm, cnorm2);
  EQUSS_LAUNCH_OK("codebook_prepare_kernel");
  return EQUSS_OK;
}
