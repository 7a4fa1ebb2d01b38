// eval_probe.cu -- cluster / linear probe prediction at label resolution (K8) and the confusion
// histogram (K9).  Reference: model/evaluator.py:46-82,85-111 and model/metric.py:44-58.
//
// The reference bilinearly upsamples the (B, D, h, w) feature map to label resolution (13.4 GB at the
// cocostuff27 eval shape) and only then takes the 27 inner products per pixel.  Both steps are linear,
// and the per-pixel L2 normalisation is a positive scalar, so
//     argmax_j <normalize(interp(x)), c_j>  ==  argmax_j interp(<x, c_j>)
// (and the linear probe commutes with the interpolation exactly because the four weights sum to one).
// K8 therefore runs in two HBM-friendly steps: token-resolution logits (reads x once), then one pass
// over the label pixels that interpolates 27 logits from an L2-resident table, takes the argmax and
// feeds warp-privatised shared-memory confusion bins.
#include <cstdlib>
#include <cstring>
#include "equss_common.cuh"
#include "probe_argmax_core.cuh"

namespace equss {

constexpr int kProbeMaxHeads = 4;

// ------------------------------------------------------------------------------------------------
// step 1: logits[b*hw + s][j] = sum_c feat[b][c][s] * w[j][c] + bias[j]
//   block = 128 threads, PT pixels per thread (consecutive threads = consecutive pixels: every
//   channel read is a coalesced 512-byte row), CT output channels per block kept in registers,
//   weights staged transposed in shared memory ([DC][CT]) so each feature value meets CT/4 broadcast
//   LDS.128.
// ------------------------------------------------------------------------------------------------
template <int CT, int PT, int TB>
__global__ void __launch_bounds__(TB)
probe_logits_kernel(const float* __restrict__ feat, int D, int hw, const float* __restrict__ wmat,
                    const float* __restrict__ bias, int c_total, int c_pad, float* __restrict__ logits) {
  constexpr int DC = 128;                       // feature channels per shared-memory chunk
  __shared__ __align__(16) float s_w[DC * CT];  // [c][j]
  const int b = blockIdx.y;
  const int j0 = blockIdx.z * CT;               // first output channel of this block
  const int s0 = blockIdx.x * (TB * PT);
  const float* fb = feat + (long long)b * D * hw;
  float acc[PT][CT];
#pragma unroll
  for (int p = 0; p < PT; ++p)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[p][j] = 0.f;
  int s[PT];
  bool live[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) { s[p] = s0 + threadIdx.x + p * TB; live[p] = s[p] < hw; }
  for (int c0 = 0; c0 < D; c0 += DC) {
    const int dc = min(DC, D - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < dc * CT; i += TB) {
      int c = i / CT, j = i - c * CT;
      s_w[i] = (j0 + j < c_pad) ? __ldg(wmat + (long long)(c0 + c) * c_pad + j0 + j) : 0.f;   // K-major weights
    }
    __syncthreads();
    constexpr int U = 4;                        // feature channels in flight per thread
    for (int c = 0; c < dc; c += U) {
      float x[U][PT];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int p = 0; p < PT; ++p)
          x[u][p] = (live[p] && c + u < dc) ? __ldcs(fb + (long long)(c0 + c + u) * hw + s[p]) : 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4* w4 = reinterpret_cast<const float4*>(s_w + (c + u < dc ? c + u : c) * CT);
#pragma unroll
        for (int j4 = 0; j4 < CT / 4; ++j4) {
          float4 w = w4[j4];
#pragma unroll
          for (int p = 0; p < PT; ++p) {
            acc[p][4 * j4 + 0] = fmaf(x[u][p], w.x, acc[p][4 * j4 + 0]);
            acc[p][4 * j4 + 1] = fmaf(x[u][p], w.y, acc[p][4 * j4 + 1]);
            acc[p][4 * j4 + 2] = fmaf(x[u][p], w.z, acc[p][4 * j4 + 2]);
            acc[p][4 * j4 + 3] = fmaf(x[u][p], w.w, acc[p][4 * j4 + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PT; ++p) {
    if (!live[p]) continue;
    float* o = logits + ((long long)b * hw + s[p]) * c_pad + j0;
#pragma unroll
    for (int j4 = 0; j4 < CT / 4; ++j4) {
      if (j0 + 4 * j4 >= c_pad) break;
      float4 v;
      v.x = acc[p][4 * j4 + 0] + ((bias && j0 + 4 * j4 + 0 < c_total) ? __ldg(bias + j0 + 4 * j4 + 0) : 0.f);
      v.y = acc[p][4 * j4 + 1] + ((bias && j0 + 4 * j4 + 1 < c_total) ? __ldg(bias + j0 + 4 * j4 + 1) : 0.f);
      v.z = acc[p][4 * j4 + 2] + ((bias && j0 + 4 * j4 + 2 < c_total) ? __ldg(bias + j0 + 4 * j4 + 2) : 0.f);
      v.w = acc[p][4 * j4 + 3] + ((bias && j0 + 4 * j4 + 3 < c_total) ? __ldg(bias + j0 + 4 * j4 + 3) : 0.f);
      *reinterpret_cast<float4*>(o + 4 * j4) = v;
    }
  }
}

// K-split variant (the one used for the common head sizes): 256 threads = 4 k-groups x 64 pixels.  Each
// k-group contracts a quarter of the D feature channels for the same 64 pixels (one pixel per thread, CT
// accumulators), partial sums are combined through shared memory in a fixed order (deterministic).  Compared
// with one thread per pixel over all of D this gives 4x the warps (~16 per SM) to hide the feature-load
// latency, which is what bounds this kernel (the arithmetic is 1.5 % of fp32 peak).
template <int CT, int PT>
__global__ void __launch_bounds__(256, (PT == 1) ? 2 : 1)
probe_logits_ksplit_kernel(const float* __restrict__ feat, int D, int hw, const float* __restrict__ wmat_t,
                           const float* __restrict__ bias, int c_total, int c_pad, float* __restrict__ logits) {
  constexpr int KG = 4, PX = 64 * PT, DC = 32;
  extern __shared__ __align__(16) float s_dyn[];
  float* s_w = s_dyn;                          // [KG][DC][CT]
  float* s_red = s_dyn + KG * DC * CT;         // [KG-1][CT][PX]
  const int kg = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int b = blockIdx.y;
  int s[PT];
  bool live[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) { s[p] = blockIdx.x * PX + t + 64 * p; live[p] = s[p] < hw; }
  const int dper = (D + KG - 1) / KG;          // channels per k-group
  const int cbeg = kg * dper, cend = min(D, cbeg + dper);
  const float* fb = feat + (long long)b * D * hw;
  float acc[PT][CT];
#pragma unroll
  for (int p = 0; p < PT; ++p)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[p][j] = 0.f;
  for (int c0 = 0; c0 < dper; c0 += DC) {
    __syncthreads();
    // weights arrive K-major ([D][c_pad]): every k-group's chunk is DC contiguous rows -> 16-byte copies
    {
      const int r4 = c_pad >> 2;                               // float4 per weight row
      for (int i = threadIdx.x; i < KG * DC * r4; i += 256) {
        const int g = i / (DC * r4), r = i - g * (DC * r4);
        const int c = r / r4, j4 = r - c * r4;
        const int ch = g * dper + c0 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c < dper && ch < D) v = __ldg(reinterpret_cast<const float4*>(wmat_t + (long long)ch * c_pad) + j4);
        *reinterpret_cast<float4*>(s_w + (g * DC + c) * CT + 4 * j4) = v;
      }
      if (c0 == 0 && c_pad < CT)                                // unused accumulator columns stay zero
        for (int i = threadIdx.x; i < KG * DC * (CT - c_pad); i += 256) {
          const int row = i / (CT - c_pad), j = c_pad + i % (CT - c_pad);
          s_w[row * CT + j] = 0.f;
        }
    }
    __syncthreads();
    const float* wg = s_w + kg * DC * CT;
    constexpr int U = (PT == 1) ? 8 : 4;
#pragma unroll 1
    for (int c = 0; c < DC; c += U) {
      float x[U][PT];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ch = cbeg + c0 + c + u;
#pragma unroll
        for (int p = 0; p < PT; ++p) x[u][p] = (live[p] && ch < cend) ? __ldcs(fb + (long long)ch * hw + s[p]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4* w4 = reinterpret_cast<const float4*>(wg + (c + u) * CT);
#pragma unroll
        for (int j4 = 0; j4 < CT / 4; ++j4) {
          float4 w = w4[j4];
#pragma unroll
          for (int p = 0; p < PT; ++p) {
            acc[p][4 * j4 + 0] = fmaf(x[u][p], w.x, acc[p][4 * j4 + 0]);
            acc[p][4 * j4 + 1] = fmaf(x[u][p], w.y, acc[p][4 * j4 + 1]);
            acc[p][4 * j4 + 2] = fmaf(x[u][p], w.z, acc[p][4 * j4 + 2]);
            acc[p][4 * j4 + 3] = fmaf(x[u][p], w.w, acc[p][4 * j4 + 3]);
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < PT; ++p) {
    // fixed-order reduction over the k-groups through shared memory, one pixel slot at a time
    if (kg > 0) {
#pragma unroll
      for (int j = 0; j < CT; ++j) s_red[((kg - 1) * CT + j) * 64 + t] = acc[p][j];
    }
    __syncthreads();
    if (kg == 0 && live[p]) {
      float* o = logits + ((long long)b * hw + s[p]) * c_pad;
#pragma unroll
      for (int j4 = 0; j4 < CT / 4; ++j4) {
        if (4 * j4 >= c_pad) break;
        float r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * j4 + e;
          r[e] = ((acc[p][j] + s_red[(0 * CT + j) * 64 + t]) + s_red[(1 * CT + j) * 64 + t]) + s_red[(2 * CT + j) * 64 + t];
          if (bias && j < c_total) r[e] += __ldg(bias + j);
        }
        *reinterpret_cast<float4*>(o + 4 * j4) = make_float4(r[0], r[1], r[2], r[3]);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// warp-privatised histogram helpers
// ------------------------------------------------------------------------------------------------
// Adds one count for `bin` (or nothing if bin < 0) for every lane; equal neighbouring bins inside the
// warp are merged into a single shared-memory atomic (labels are spatially coherent).
__device__ __forceinline__ void warp_hist_add(int* hist, int bin) {
  const unsigned lane = threadIdx.x & 31;
  int prev = __shfl_up_sync(0xffffffffu, bin, 1);
  bool start = (lane == 0) || (prev != bin);
  unsigned starts = __ballot_sync(0xffffffffu, start);
  if (start && bin >= 0) {
    unsigned higher = (lane == 31) ? 0u : (starts >> (lane + 1));
    int run = higher ? (__ffs(higher)) : (32 - (int)lane);
    atomicAdd(hist + bin, run);
  }
}

struct ProbeHeads {
  int n_heads;
  int off[kProbeMaxHeads];
  int cnt[kProbeMaxHeads];
  int rows[kProbeMaxHeads];
  long long* preds[kProbeMaxHeads];
  unsigned long long* conf[kProbeMaxHeads];
  int hist_off[kProbeMaxHeads];   // offset (ints) of this head's bins inside one warp's private area
  int hist_per_warp;              // ints per warp
};

// step 2: thread per label pixel.
__global__ void __launch_bounds__(256)
probe_argmax_confusion_kernel(const float* __restrict__ logits, int B, int h, int w, int c_pad,
                              const long long* __restrict__ label, int H, int W, int C,
                              ProbeHeads heads, float scale_h, float scale_w) {
  extern __shared__ int s_hist[];   // [warps][hist_per_warp]
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < nwarps * heads.hist_per_warp; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  int* myhist = s_hist + warp * heads.hist_per_warp;
  const long long P = (long long)B * H * W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // all lanes of a warp iterate together (P rounded up to a multiple of 32 inside the loop bound)
  const long long P32 = (P + 31) & ~31LL;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P32; p += stride) {
    const bool live = p < P;
    int bins[kProbeMaxHeads];
#pragma unroll
    for (int hd = 0; hd < kProbeMaxHeads; ++hd) bins[hd] = -1;
    if (live) {
      int X = (int)(p % W);
      long long t = p / W;
      int Y = (int)(t % H);
      int b = (int)(t / H);
      // PyTorch upsample_bilinear2d, align_corners=False (area_pixel_compute_source_index)
      float sy = scale_h * ((float)Y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
      float sx = scale_w * ((float)X + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
      int y0 = (int)sy, x0 = (int)sx;
      int y1 = y0 + ((y0 < h - 1) ? 1 : 0), x1 = x0 + ((x0 < w - 1) ? 1 : 0);
      float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
      float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
      const float* base = logits + (long long)b * h * w * c_pad;
      const float* p00 = base + ((long long)y0 * w + x0) * c_pad;
      const float* p01 = base + ((long long)y0 * w + x1) * c_pad;
      const float* p10 = base + ((long long)y1 * w + x0) * c_pad;
      const float* p11 = base + ((long long)y1 * w + x1) * c_pad;
      long long lab = __ldcs(label + p);
#pragma unroll
      for (int hd = 0; hd < kProbeMaxHeads; ++hd) {
        if (hd >= heads.n_heads) break;
        float best = -INFINITY;
        int bj = 0;
        const int off = heads.off[hd], cnt = heads.cnt[hd];
        if ((off & 3) == 0) {
          // vector path: four channels per 16-byte load from each of the four neighbouring tokens
          const float4* q00 = reinterpret_cast<const float4*>(p00 + off);
          const float4* q01 = reinterpret_cast<const float4*>(p01 + off);
          const float4* q10 = reinterpret_cast<const float4*>(p10 + off);
          const float4* q11 = reinterpret_cast<const float4*>(p11 + off);
          const int n4 = (cnt + 3) >> 2;
          for (int g = 0; g < n4; ++g) {
            float4 a = __ldg(q00 + g), bq = __ldg(q01 + g), c = __ldg(q10 + g), dq = __ldg(q11 + g);
            float v0 = ly0 * (lx0 * a.x + lx1 * bq.x) + ly1 * (lx0 * c.x + lx1 * dq.x);
            float v1 = ly0 * (lx0 * a.y + lx1 * bq.y) + ly1 * (lx0 * c.y + lx1 * dq.y);
            float v2 = ly0 * (lx0 * a.z + lx1 * bq.z) + ly1 * (lx0 * c.z + lx1 * dq.z);
            float v3 = ly0 * (lx0 * a.w + lx1 * bq.w) + ly1 * (lx0 * c.w + lx1 * dq.w);
            const int j = 4 * g;
            if (v0 > best) { best = v0; bj = j; }                      // first maximal index wins (torch.argmax)
            if (j + 1 < cnt && v1 > best) { best = v1; bj = j + 1; }
            if (j + 2 < cnt && v2 > best) { best = v2; bj = j + 2; }
            if (j + 3 < cnt && v3 > best) { best = v3; bj = j + 3; }
          }
        } else {
          for (int j = 0; j < cnt; ++j) {
            float v00 = __ldg(p00 + off + j), v01 = __ldg(p01 + off + j);
            float v10 = __ldg(p10 + off + j), v11 = __ldg(p11 + off + j);
            float v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
            if (v > best) { best = v; bj = j; }     // first maximal index wins (torch.argmax)
          }
        }
        if (heads.preds[hd]) __stcs(heads.preds[hd] + p, (long long)bj);
        if (heads.conf[hd] && lab >= 0 && lab < C && bj < C) bins[hd] = bj * C + (int)lab;
      }
    }
#pragma unroll
    for (int hd = 0; hd < kProbeMaxHeads; ++hd) {
      if (hd >= heads.n_heads) break;
      if (heads.conf[hd]) warp_hist_add(myhist + heads.hist_off[hd], bins[hd]);
    }
  }
  __syncthreads();
  for (int hd = 0; hd < heads.n_heads; ++hd) {
    if (!heads.conf[hd]) continue;
    const int nb = heads.rows[hd] * C;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      int t = 0;
      for (int wv = 0; wv < nwarps; ++wv) t += s_hist[wv * heads.hist_per_warp + heads.hist_off[hd] + i];
      if (t) atomicAdd(heads.conf[hd] + i, (unsigned long long)t);
    }
  }
}

// step 2, row-block variant (the common case: every head starts at a multiple of four channels and has at most
// CMAX channels).  One block per (image, RB consecutive label rows), one thread per label column X.  The
// horizontal interpolation  Hy[c] = w0*L[y][x0][c] + w1*L[y][x1][c]  of the two token rows (y0, y1) depends only
// on X, so a thread keeps it in registers and reuses it for every label row that maps to the same token-row pair
// (8 rows at the cocostuff27 shape): per label pixel and channel that leaves  v = h0*H0[c] + h1*H1[c]  and the
// argmax update -- the association order of PyTorch's upsample_bilinear2d -- instead of four loads and seven
// multiply-adds.  Confusion bins are per block in shared memory (run-length merged per warp).
constexpr int kRowsPerBlock = 8;
constexpr int kProbeArgmaxDefaultMode = 1;   // see equss_probe_argmax_confusion
template <int CMAX, int TMAX, int MINB>
__global__ void __launch_bounds__(TMAX, MINB)
probe_argmax_rows_kernel(const float* __restrict__ logits, int h, int w, int c_pad, const long long* __restrict__ label,
                         int H, int W, int C, ProbeHeads heads, float scale_h, float scale_w, int rows_per_block,
                         int row_shift) {
  extern __shared__ int s_hist[];   // [hist_per_warp] = all heads' bins, one copy per block
  for (int i = threadIdx.x; i < heads.hist_per_warp; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int b = blockIdx.y;
  // block k covers label rows [k*RB - shift, (k+1)*RB - shift): with an integer H/h ratio and shift = RB/2 these
  // are exactly the rows that interpolate between one pair of token rows
  const int Ya = (int)blockIdx.x * rows_per_block - row_shift;
  const int Y0 = max(Ya, 0);
  const int Y1 = min(H, Ya + rows_per_block);
  const float* base = logits + (long long)b * h * w * c_pad;
  const int Wpad = (W + 31) & ~31;
  for (int X = threadIdx.x; X < Wpad; X += blockDim.x) {
    const bool live = X < W;
    float sx = scale_w * ((float)X + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    int x0 = (int)sx;
    if (x0 > w - 1) x0 = w - 1;
    const int x1 = x0 + ((x0 < w - 1) ? 1 : 0);
    const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
    // labels of this column for all rows of the block, loaded up front (one latency, RB loads in flight)
    unsigned long long labs = 0;         // one byte per row: label, or 255 = ignore (C <= 255 on this path)
#pragma unroll
    for (int r = 0; r < kRowsPerBlock; ++r) {
      const int Y = Ya + r;
      long long lab = -1;
      if (live && Y >= Y0 && Y < Y1) lab = __ldcs(label + ((long long)b * H + Y) * W + X);
      labs |= (unsigned long long)((lab >= 0 && lab < C) ? (unsigned)lab : 255u) << (8 * r);
    }
#pragma unroll 1
    for (int hd = 0; hd < heads.n_heads; ++hd) {
      const int off = heads.off[hd], cnt = heads.cnt[hd];
      long long* preds = heads.preds[hd];
      const bool want_conf = heads.conf[hd] != nullptr;
      int* hist = s_hist + heads.hist_off[hd];
      float H0[CMAX], H1[CMAX];
      int cy0 = -1;
#pragma unroll 1
      for (int r = 0; r < kRowsPerBlock; ++r) {
        const int Y = Ya + r;
        if (Y < Y0) continue;        // block-uniform
        if (Y >= Y1) break;
        // PyTorch upsample_bilinear2d, align_corners=False (area_pixel_compute_source_index)
        float sy = scale_h * ((float)Y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
        int y0 = (int)sy;
        if (y0 > h - 1) y0 = h - 1;
        const int y1 = y0 + ((y0 < h - 1) ? 1 : 0);
        const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
        if (y0 != cy0) {             // block-uniform: a new token-row pair
          cy0 = y0;
          const float4* p00 = reinterpret_cast<const float4*>(base + ((long long)y0 * w + x0) * c_pad + off);
          const float4* p01 = reinterpret_cast<const float4*>(base + ((long long)y0 * w + x1) * c_pad + off);
          const float4* p10 = reinterpret_cast<const float4*>(base + ((long long)y1 * w + x0) * c_pad + off);
          const float4* p11 = reinterpret_cast<const float4*>(base + ((long long)y1 * w + x1) * c_pad + off);
#pragma unroll
          for (int g = 0; g < CMAX / 4; ++g) {
            const float4 a = __ldg(p00 + g), bq = __ldg(p01 + g), c = __ldg(p10 + g), dq = __ldg(p11 + g);
            H0[4 * g + 0] = lx0 * a.x + lx1 * bq.x; H0[4 * g + 1] = lx0 * a.y + lx1 * bq.y;
            H0[4 * g + 2] = lx0 * a.z + lx1 * bq.z; H0[4 * g + 3] = lx0 * a.w + lx1 * bq.w;
            H1[4 * g + 0] = lx0 * c.x + lx1 * dq.x; H1[4 * g + 1] = lx0 * c.y + lx1 * dq.y;
            H1[4 * g + 2] = lx0 * c.z + lx1 * dq.z; H1[4 * g + 3] = lx0 * c.w + lx1 * dq.w;
          }
          // CMAX - 4 < cnt <= CMAX (host check): only the last group can hold channels past the head's end; they
          // are made unable to win here, so the argmax loop needs no per-channel guard
#pragma unroll
          for (int j = CMAX - 3; j < CMAX; ++j)
            if (j >= cnt) { H0[j] = -1e30f; H1[j] = -1e30f; }
        }
        float best = -INFINITY;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < CMAX; ++j) {
          const float v = ly0 * H0[j] + ly1 * H1[j];
          if (v > best) { best = v; bj = j; }       // first maximal index wins (torch.argmax)
        }
        if (live && preds) __stcs(preds + ((long long)b * H + Y) * W + X, (long long)bj);
        if (want_conf) {
          const int lab = (int)((labs >> (8 * r)) & 0xFFull);
          const int bin = (live && lab != 255 && bj < C) ? bj * C + lab : -1;
          warp_hist_add(hist, bin);
        }
      }
    }
  }
  __syncthreads();
  for (int hd = 0; hd < heads.n_heads; ++hd) {
    if (!heads.conf[hd]) continue;
    const int nb = heads.rows[hd] * C;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      const int t = s_hist[heads.hist_off[hd] + i];
      if (t) atomicAdd(heads.conf[hd] + i, (unsigned long long)t);
    }
  }
}

// step 2, row-block variant with the tournament argmax of probe_argmax_core.cuh (same interpolation arithmetic as
// probe_argmax_rows_kernel; the horizontal interpolants live in f32x2 register pairs).
//
// Work item = (image, RB consecutive label rows that share one token-row pair, slice of blockDim.x label columns).
// The grid is persistent (blocks x SMs resident CTAs): a CTA zeroes its confusion bins once, walks items
// blockIdx.x, blockIdx.x + gridDim.x, ... and flushes the bins once.  1312 items on 296 CTAs would take five rounds for
// 4.43 rounds of work, so the schedule (ProbeSchedule, built on the host) splits the items of the last, partial round
// into `parts` row ranges each: four full rounds and one half-length round at the cocostuff27 shape.
// The item's labels wait in shared memory as bytes (written and read by the same thread: no barrier inside the loop).
// RECOMP: the tournament's second pass recomputes the winning group's values instead of keeping all of them.
struct ProbeSchedule {
  int n_full;            // items [0, n_full) are processed whole
  int n_sched;           // n_full + (n_items - n_full) * parts schedule slots
  int parts;             // row ranges per item of the remainder (divides rows_per_block)
  int rowblocks;         // row blocks per image
  int slices;            // column slices per row block
};

// Items [0, n_full) fill whole rounds of the `ctas` persistent CTAs; the remainder (fewer items than CTAs) is split
// into `parts` row ranges per item -- the largest power of two that divides the rows of an item and still leaves at
// most one slot per CTA -- so the last round is 1 / parts as long and keeps parts x as many CTAs busy.
static void probe_schedule_fill(int n_items, int ctas, int rows_per_block, ProbeSchedule* sch) {
  const int rem = n_items % ctas;
  sch->parts = 1;
  while (rem > 0 && (long long)rem * sch->parts * 2 <= ctas && rows_per_block % (sch->parts * 2) == 0) sch->parts *= 2;
  sch->n_full = n_items - rem;
  sch->n_sched = sch->n_full + rem * sch->parts;
}

template <int CMAX, int TMAX, int MINB, bool RECOMP>
__global__ void __launch_bounds__(TMAX, MINB)
probe_argmax_rows_t_kernel(const float* __restrict__ logits, int h, int w, int c_pad, const long long* __restrict__ label,
                           int H, int W, int C, ProbeHeads heads, float scale_h, float scale_w, int rows_per_block,
                           int row_shift, ProbeSchedule sch) {
  using pa::f32x2;
  extern __shared__ int s_hist[];   // [hist_per_warp] = all heads' bins, one copy per block; then the label bytes
  uint8_t* s_lab = reinterpret_cast<uint8_t*>(s_hist + heads.hist_per_warp);     // [kRowsPerBlock][blockDim.x]
  for (int i = threadIdx.x; i < heads.hist_per_warp; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
#pragma unroll 1
  for (int slot = blockIdx.x; slot < sch.n_sched; slot += gridDim.x) {
    int item = slot, r_lo = 0, r_hi = rows_per_block;
    if (slot >= sch.n_full) {
      const int t = slot - sch.n_full;
      item = sch.n_full + t / sch.parts;
      const int rows = rows_per_block / sch.parts;
      r_lo = (t % sch.parts) * rows;
      r_hi = r_lo + rows;
    }
    const int slice = item % sch.slices;
    const int rbk = (item / sch.slices) % sch.rowblocks;
    const int b = item / (sch.slices * sch.rowblocks);
    const int Ya = rbk * rows_per_block - row_shift;
    const int Y0 = max(Ya + r_lo, 0);
    const int Y1 = min(H, Ya + r_hi);
    const int X = (int)(slice * blockDim.x + threadIdx.x);
    const bool live = X < W;
    // labels of this column for all rows of the item, loaded up front (one latency, RB loads in flight); one byte per
    // row: label, or 255 = ignore (C <= 255 on this path)
    {
      long long lab[kRowsPerBlock];
#pragma unroll
      for (int r = 0; r < kRowsPerBlock; ++r) {
        const int Y = Y0 + r;
        lab[r] = (live && Y < Y1) ? __ldcs(label + ((long long)b * H + Y) * W + X) : -1;
      }
#pragma unroll
      for (int r = 0; r < kRowsPerBlock; ++r)
        s_lab[r * blockDim.x + threadIdx.x] = (uint8_t)((lab[r] >= 0 && lab[r] < C) ? (unsigned)lab[r] : 255u);
    }
    const float* base = logits + (long long)b * h * w * c_pad;
    float sx = scale_w * ((float)X + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    int x0 = (int)sx;
    if (x0 > w - 1) x0 = w - 1;
    const int x1 = x0 + ((x0 < w - 1) ? 1 : 0);
    const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
#pragma unroll 1
    for (int hd = 0; hd < heads.n_heads; ++hd) {
      const int off = heads.off[hd], cnt = heads.cnt[hd];
      f32x2 H0p[CMAX / 2], H1p[CMAX / 2];
      int cy0 = -1;
#pragma unroll 1
      for (int Y = Y0; Y < Y1; ++Y) {
        // PyTorch upsample_bilinear2d, align_corners=False (area_pixel_compute_source_index)
        float sy = scale_h * ((float)Y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
        int y0 = (int)sy;
        if (y0 > h - 1) y0 = h - 1;
        const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
        if (y0 != cy0) {             // block-uniform: a new token-row pair
          cy0 = y0;
          const int y1 = y0 + ((y0 < h - 1) ? 1 : 0);
          const f32x2 lx0p = pa::pk(lx0, lx0), lx1p = pa::pk(lx1, lx1);
          const float4* p00 = reinterpret_cast<const float4*>(base + ((long long)y0 * w + x0) * c_pad + off);
          const float4* p01 = reinterpret_cast<const float4*>(base + ((long long)y0 * w + x1) * c_pad + off);
          const float4* p10 = reinterpret_cast<const float4*>(base + ((long long)y1 * w + x0) * c_pad + off);
          const float4* p11 = reinterpret_cast<const float4*>(base + ((long long)y1 * w + x1) * c_pad + off);
#pragma unroll
          for (int g = 0; g < CMAX / 4; ++g) {
            const float4 a = __ldg(p00 + g), bq = __ldg(p01 + g), c = __ldg(p10 + g), dq = __ldg(p11 + g);
            // H = fma(lx0, left, lx1 * right): the association of the row kernel
            H0p[2 * g] = pa::interp2(pa::pk(a.x, a.y), pa::pk(bq.x, bq.y), lx0p, lx1p);
            H0p[2 * g + 1] = pa::interp2(pa::pk(a.z, a.w), pa::pk(bq.z, bq.w), lx0p, lx1p);
            H1p[2 * g] = pa::interp2(pa::pk(c.x, c.y), pa::pk(dq.x, dq.y), lx0p, lx1p);
            H1p[2 * g + 1] = pa::interp2(pa::pk(c.z, c.w), pa::pk(dq.z, dq.w), lx0p, lx1p);
          }
          // CMAX - 4 < cnt <= CMAX (host check): only the last group can hold channels past the head's end; they
          // are made unable to win here, so the argmax needs no per-channel guard
          {
            float e0, e1, e2, e3, f0, f1, f2, f3;
            pa::upk(H0p[CMAX / 2 - 2], e0, e1); pa::upk(H0p[CMAX / 2 - 1], e2, e3);
            pa::upk(H1p[CMAX / 2 - 2], f0, f1); pa::upk(H1p[CMAX / 2 - 1], f2, f3);
            if (CMAX - 3 >= cnt) { e1 = -1e30f; f1 = -1e30f; }
            if (CMAX - 2 >= cnt) { e2 = -1e30f; f2 = -1e30f; }
            if (CMAX - 1 >= cnt) { e3 = -1e30f; f3 = -1e30f; }
            H0p[CMAX / 2 - 2] = pa::pk(e0, e1); H0p[CMAX / 2 - 1] = pa::pk(e2, e3);
            H1p[CMAX / 2 - 2] = pa::pk(f0, f1); H1p[CMAX / 2 - 1] = pa::pk(f2, f3);
          }
        }
        const int lab = s_lab[(Y - Y0) * blockDim.x + threadIdx.x];
        const int bj = pa::argmax_interp<CMAX, RECOMP>(H0p, H1p, ly0, ly1);
        long long* preds = heads.preds[hd];
        if (live && preds) __stcs(preds + ((long long)b * H + Y) * W + X, (long long)bj);
        if (heads.conf[hd] != nullptr) {
          const int bin = (live && lab != 255 && bj < C) ? bj * C + lab : -1;
          warp_hist_add(s_hist + heads.hist_off[hd], bin);
        }
      }
    }
  }
  __syncthreads();
  for (int hd = 0; hd < heads.n_heads; ++hd) {
    if (!heads.conf[hd]) continue;
    const int nb = heads.rows[hd] * C;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      const int t = s_hist[heads.hist_off[hd] + i];
      if (t) atomicAdd(heads.conf[hd] + i, (unsigned long long)t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K9 standalone confusion histogram (UnSegMetrics.update)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
confusion_kernel(const long long* __restrict__ preds, const long long* __restrict__ label, long long n,
                 int C, int rows, unsigned long long* __restrict__ conf, int use_smem) {
  extern __shared__ int s_hist[];   // [warps][rows*C] when use_smem
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int nb = rows * C;
  if (use_smem) {
    for (int i = threadIdx.x; i < nwarps * nb; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
  }
  int* myhist = s_hist + warp * nb;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n32 = (n + 31) & ~31LL;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n32; p += stride) {
    int bin = -1;
    if (p < n) {
      long long pr = __ldcs(preds + p), lb = __ldcs(label + p);
      // mask of model/metric.py:49 -- predictions >= num_classes are dropped even with extra classes
      if (lb >= 0 && lb < C && pr >= 0 && pr < C) bin = (int)pr * C + (int)lb;
    }
    if (use_smem) {
      warp_hist_add(myhist, bin);
    } else if (bin >= 0) {
      atomicAdd(conf + bin, 1ULL);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      int t = 0;
      for (int wv = 0; wv < nwarps; ++wv) t += s_hist[wv * nb + i];
      if (t) atomicAdd(conf + i, (unsigned long long)t);
    }
  }
}

}  // namespace equss

using namespace equss;

extern "C" int equss_probe_cpad(int c_total) { return (c_total + 3) & ~3; }

extern "C" int equss_probe_logits(const float* feat, int B, int D, int h, int w, const float* wmat,
                                  const float* bias, int c_total, float* logits, void* stream) {
  EQUSS_REQUIRE(feat && wmat && logits, EQUSS_ERR_INVALID_ARG, "equss_probe_logits: null pointer");
  EQUSS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0 && c_total > 0, EQUSS_ERR_INVALID_ARG,
                "equss_probe_logits: bad shape B=%d D=%d h=%d w=%d C=%d", B, D, h, w, c_total);
  EQUSS_REQUIRE(!((uintptr_t)logits & 15), EQUSS_ERR_INVALID_ARG, "equss_probe_logits: logits must be 16-byte aligned");
  const int hw = h * w;
  const int c_pad = equss_probe_cpad(c_total);
  cudaStream_t st = (cudaStream_t)stream;
  if (c_pad <= 56) {
    const int ptv = (getenv("EQUSS_PROBE_PT") ? atoi(getenv("EQUSS_PROBE_PT")) : 2);
    dim3 grid((hw + 64 * ptv - 1) / (64 * ptv), B, 1);
#define EQUSS_KSPLIT(CTV, PTV)                                                                                   \
    {                                                                                                         \
      const size_t smem = (size_t)(4 * 32 * CTV + 3 * CTV * 64) * sizeof(float);                              \
      EQUSS_CUDA_OK(cudaFuncSetAttribute(probe_logits_ksplit_kernel<CTV, PTV>,                                     \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
      probe_logits_ksplit_kernel<CTV, PTV><<<grid, 256, smem, st>>>(feat, D, hw, wmat, bias, c_total, c_pad, logits); \
    }
    if (ptv == 1) {
      if (c_pad <= 28) EQUSS_KSPLIT(28, 1) else if (c_pad <= 32) EQUSS_KSPLIT(32, 1) else EQUSS_KSPLIT(56, 1)
    } else {
      if (c_pad <= 28) EQUSS_KSPLIT(28, 2) else if (c_pad <= 32) EQUSS_KSPLIT(32, 2) else EQUSS_KSPLIT(56, 2)
    }
#undef EQUSS_KSPLIT
  } else {
    dim3 grid((hw + 127) / 128, B, (c_pad + 63) / 64);
    probe_logits_kernel<64, 2, 64><<<grid, 64, 0, st>>>(feat, D, hw, wmat, bias, c_total, c_pad, logits);
  }
  EQUSS_LAUNCH_OK("probe_logits_kernel");
  return EQUSS_OK;
}

extern "C" int equss_probe_argmax_schedule(int n_items, int ctas, int rows_per_block, int32_t* out3) {
  EQUSS_REQUIRE(n_items > 0 && ctas > 0 && ctas <= n_items && rows_per_block > 0 && out3, EQUSS_ERR_INVALID_ARG,
                "equss_probe_argmax_schedule: n_items=%d ctas=%d rows_per_block=%d", n_items, ctas, rows_per_block);
  ProbeSchedule sch;
  memset(&sch, 0, sizeof(sch));
  probe_schedule_fill(n_items, ctas, rows_per_block, &sch);
  out3[0] = sch.n_full; out3[1] = sch.n_sched; out3[2] = sch.parts;
  return EQUSS_OK;
}

extern "C" int equss_probe_argmax_confusion(const float* logits, int B, int h, int w, int c_total,
                                            const int64_t* label, int H, int W, int num_classes, int n_heads,
                                            const int32_t* head_off_host, const int32_t* head_cnt_host,
                                            int64_t* const* preds_out_host, int64_t* const* confusion_host,
                                            const int32_t* conf_rows_host, void* stream) {
  EQUSS_REQUIRE(logits && label && head_off_host && head_cnt_host, EQUSS_ERR_INVALID_ARG,
                "equss_probe_argmax_confusion: null pointer");
  EQUSS_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0 && c_total > 0 && num_classes > 0, EQUSS_ERR_INVALID_ARG,
                "equss_probe_argmax_confusion: bad shape");
  EQUSS_REQUIRE(n_heads >= 1 && n_heads <= kProbeMaxHeads, EQUSS_ERR_UNSUPPORTED, "n_heads=%d outside [1,%d]", n_heads,
                kProbeMaxHeads);
  ProbeHeads hd;
  memset(&hd, 0, sizeof(hd));
  hd.n_heads = n_heads;
  int per_warp = 0;
  for (int i = 0; i < n_heads; ++i) {
    hd.off[i] = head_off_host[i];
    hd.cnt[i] = head_cnt_host[i];
    EQUSS_REQUIRE(hd.off[i] >= 0 && hd.cnt[i] > 0 && hd.off[i] + hd.cnt[i] <= c_total, EQUSS_ERR_INVALID_ARG,
                  "head %d covers channels [%d,%d) outside [0,%d)", i, hd.off[i], hd.off[i] + hd.cnt[i], c_total);
    hd.preds[i] = preds_out_host ? (long long*)preds_out_host[i] : nullptr;
    hd.conf[i] = confusion_host ? (unsigned long long*)confusion_host[i] : nullptr;
    hd.rows[i] = (conf_rows_host && hd.conf[i]) ? conf_rows_host[i] : 0;
    if (hd.conf[i]) {
      EQUSS_REQUIRE(hd.rows[i] >= num_classes, EQUSS_ERR_INVALID_ARG, "confusion rows %d < num_classes %d", hd.rows[i],
                    num_classes);
      hd.hist_off[i] = per_warp;
      per_warp += hd.rows[i] * num_classes;
    }
  }
  hd.hist_per_warp = per_warp;
  const float scale_h = (float)h / (float)H, scale_w = (float)w / (float)W;
  {
    // row-block kernel: heads aligned to four channels, at most 28 / 32 channels each
    bool fast = (size_t)per_warp * sizeof(int) <= 48 * 1024 && num_classes <= 255 && getenv("EQUSS_PROBE_ARGMAX_OLD") == nullptr;
    int cmax = 0;
    for (int i = 0; i < n_heads; ++i) { fast = fast && (hd.off[i] % 4) == 0; cmax = hd.cnt[i] > cmax ? hd.cnt[i] : cmax; }
    int cmin = cmax;
    for (int i = 0; i < n_heads; ++i) cmin = hd.cnt[i] < cmin ? hd.cnt[i] : cmin;
    const int cm = (cmax + 3) & ~3;          // every head must end inside the last group of four
    if (fast && cm <= 32 && cmin > cm - 4) {
      // rows per block = label rows per token row when H/h is an integer in [2, 8]; else 8 rows, unaligned
      int rb = kRowsPerBlock, shift = 0;
      if (H % h == 0 && H / h >= 2 && H / h <= kRowsPerBlock) { rb = H / h; shift = rb / 2; }
      int threads = (W + 31) & ~31;
      if (threads > 512) threads = 256;
      dim3 grid((unsigned)((H + shift + rb - 1) / rb), (unsigned)B);
      const size_t smem = (size_t)(per_warp > 0 ? per_warp : 1) * sizeof(int);
#define EQUSS_ROWS_LAUNCH(CM, TM, MB)                                                                                 \
      probe_argmax_rows_kernel<CM, TM, MB><<<grid, threads, smem, (cudaStream_t)stream>>>(                            \
          logits, h, w, equss_probe_cpad(c_total), (const long long*)label, H, W, num_classes, hd, scale_h, scale_w, rb, shift)
#define EQUSS_ROWS_CM(TM, MB)                                                                                         \
      switch (cm) {                                                                                                   \
        case 4: EQUSS_ROWS_LAUNCH(4, TM, MB); break;   case 8: EQUSS_ROWS_LAUNCH(8, TM, MB); break;                   \
        case 12: EQUSS_ROWS_LAUNCH(12, TM, MB); break; case 16: EQUSS_ROWS_LAUNCH(16, TM, MB); break;                 \
        case 20: EQUSS_ROWS_LAUNCH(20, TM, MB); break; case 24: EQUSS_ROWS_LAUNCH(24, TM, MB); break;                 \
        case 28: EQUSS_ROWS_LAUNCH(28, TM, MB); break; default: EQUSS_ROWS_LAUNCH(32, TM, MB); break;                 \
      }
      // EQUSS_PROBE_ARGMAX_T = 0 selects the sequential-argmax kernel (one CTA per item); the default is the
      // persistent tournament kernel.  Both return identical predictions (tests/test_gpu_eval.py).  Measured on B200 at
      // the cocostuff27 shape: 80 us sequential, 70.5 us tournament (scripts/bench_probe_argmax.py; variants that were
      // slower: 160 threads x 3 CTAs with 128 registers 74-78 us, scalar FMUL / FFMA instead of f32x2 74 us, one CTA
      // per item 73 us).
      const int tmode = getenv("EQUSS_PROBE_ARGMAX_T") ? atoi(getenv("EQUSS_PROBE_ARGMAX_T")) : kProbeArgmaxDefaultMode;
#define EQUSS_ROWS_T_LAUNCH(CM, TM, MB)                                                                               \
      probe_argmax_rows_t_kernel<CM, TM, MB, true><<<tgrid, tthreads, tsmem, (cudaStream_t)stream>>>(                 \
          logits, h, w, equss_probe_cpad(c_total), (const long long*)label, H, W, num_classes, hd, scale_h, scale_w, rb, shift, sch)
#define EQUSS_ROWS_T_CM(TM, MB)                                                                                       \
      {                                                                                                               \
        const int wpad = (W + 31) & ~31;                                                                              \
        const int tthreads = wpad < TM ? wpad : TM;                                                                   \
        ProbeSchedule sch;                                                                                            \
        sch.rowblocks = (int)grid.x; sch.slices = (wpad + tthreads - 1) / tthreads;                                   \
        const long long n_items = (long long)B * sch.rowblocks * sch.slices;                                          \
        EQUSS_REQUIRE(n_items < (1ll << 28), EQUSS_ERR_UNSUPPORTED, "probe argmax: %lld work items", n_items);        \
        long long ctas = (long long)num_sms() * MB;                                                                   \
        if (ctas > n_items) ctas = n_items;                                                                           \
        probe_schedule_fill((int)n_items, (int)ctas, rb, &sch);                                                       \
        const dim3 tgrid((unsigned)ctas);                                                                             \
        const size_t tsmem = smem + (size_t)kRowsPerBlock * tthreads;                                                 \
        switch (cm) {                                                                                                 \
          case 4: EQUSS_ROWS_T_LAUNCH(4, TM, MB); break;   case 8: EQUSS_ROWS_T_LAUNCH(8, TM, MB); break;             \
          case 12: EQUSS_ROWS_T_LAUNCH(12, TM, MB); break; case 16: EQUSS_ROWS_T_LAUNCH(16, TM, MB); break;           \
          case 20: EQUSS_ROWS_T_LAUNCH(20, TM, MB); break; case 24: EQUSS_ROWS_T_LAUNCH(24, TM, MB); break;           \
          case 28: EQUSS_ROWS_T_LAUNCH(28, TM, MB); break; default: EQUSS_ROWS_T_LAUNCH(32, TM, MB); break;           \
        }                                                                                                             \
      }
      if (tmode != 0) EQUSS_ROWS_T_CM(320, 2)
      else {
        if (threads <= 320) { EQUSS_ROWS_CM(320, 2) } else { EQUSS_ROWS_CM(512, 1) }
      }
#undef EQUSS_ROWS_T_CM
#undef EQUSS_ROWS_T_LAUNCH
#undef EQUSS_ROWS_CM
#undef EQUSS_ROWS_LAUNCH
      EQUSS_LAUNCH_OK("probe_argmax_rows_kernel");
      return EQUSS_OK;
    }
  }
  const int threads = 256;
  size_t smem = (size_t)(threads / 32) * per_warp * sizeof(int);
  EQUSS_REQUIRE(smem <= 200 * 1024, EQUSS_ERR_UNSUPPORTED, "confusion bins (%d per warp) do not fit shared memory", per_warp);
  if (smem > 48 * 1024)
    EQUSS_CUDA_OK(cudaFuncSetAttribute(probe_argmax_confusion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long P = (long long)B * H * W;
  long long blocks = (P + threads - 1) / threads;
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  probe_argmax_confusion_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(
      logits, B, h, w, equss_probe_cpad(c_total), (const long long*)label, H, W, num_classes, hd, scale_h, scale_w);
  EQUSS_LAUNCH_OK("probe_argmax_confusion_kernel");
  return EQUSS_OK;
}

extern "C" int equss_confusion_update(const int64_t* preds, const int64_t* label, int64_t n, int num_classes,
                                      int rows, int64_t* confusion, void* stream) {
  EQUSS_REQUIRE(n >= 0 && num_classes > 0 && rows >= num_classes, EQUSS_ERR_INVALID_ARG,
                "equss_confusion_update: bad shape n=%lld C=%d rows=%d", (long long)n, num_classes, rows);
  if (n == 0) return EQUSS_OK;   // empty tensors have null data pointers
  EQUSS_REQUIRE(preds && label && confusion, EQUSS_ERR_INVALID_ARG, "equss_confusion_update: null pointer");
  const int threads = 256;
  size_t smem = (size_t)(threads / 32) * rows * num_classes * sizeof(int);
  int use_smem = smem <= 96 * 1024;
  if (!use_smem) smem = 0;
  if (smem > 48 * 1024)
    EQUSS_CUDA_OK(cudaFuncSetAttribute(confusion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (n + threads - 1) / threads;
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  confusion_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(
      (const long long*)preds, (const long long*)label, (long long)n, num_classes, rows,
      (unsigned long long*)confusion, use_smem);
  EQUSS_LAUNCH_OK("confusion_kernel");
  return EQUSS_OK;
}
