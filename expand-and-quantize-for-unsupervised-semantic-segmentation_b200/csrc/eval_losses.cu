// eval_losses.cu -- the two training losses of UnSegEvaluator.forward, evaluated from the TOKEN-resolution
// probe logits with the bilinear upsampling fused in, forward and backward:
//     linear_loss  = CrossEntropy(linear_probe(up(x))[mask], label[mask])            model/evaluator.py:67,72-80
//     cluster_loss = -mean_p < normalize(up(x))_p , normalize(c)_{argmax_p} >        model/evaluator.py:95-106 (alpha=None)
// The reference upsamples the (B, D, h, w) features to label resolution (13.4 GB at the cocostuff27 shape) and runs
// both probes there.  Interpolation is linear, so up(x).w_j = up(x.w_j): both losses need only the 27 + 27
// interpolated logits per label pixel -- the quantity the prediction kernel (eval_probe.cu) already forms -- plus,
// for the cosine of the cluster loss, the norm of the interpolated feature vector, which follows from the 2x2 Gram
// terms of the token grid:  |sum_t w_t x_t|^2 = sum_{t,t'} w_t w_t' <x_t, x_t'>   (token_gram_kernel).
//
// probe_losses_rows_kernel: one block per (image, label rows that share a token-row pair), one thread per label
// column; the horizontal interpolation of both token rows stays in registers for all rows of the block (as in
// probe_argmax_rows_kernel).  Backward: d loss / d (token logits) is the transpose of the interpolation applied to
// the per-pixel logit gradients (softmax - onehot, resp. -1/norm at the winning cluster).  Each thread accumulates
// its column's contribution over the block's rows in registers, lanes that share a token column take turns adding
// into a per-warp shared-memory slab (ranks from match.any: no atomics, no conflicts), and the slab is flushed to
// the global gradient table once per warp.
#include <cstring>
#include "equss_common.cuh"

namespace equss {

constexpr int kLossRows = 8;

// gram[b][y][x][0..4] = <x,x>, <x, right>, <x, down>, <x, down-right>, <x, down-left>  (0 where the neighbour is outside)
__global__ void __launch_bounds__(128)
token_gram_kernel(const float* __restrict__ feat, int D, int h, int w, float* __restrict__ gram) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const long long hw = (long long)h * w;
  const float* p = feat + (long long)b * D * hw + (long long)y * w + x;
  const bool r = x + 1 < w, dn = y + 1 < h, l = x > 0;
  float s[4][5];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int k = 0; k < 5; ++k) s[u][k] = 0.f;
  int c = 0;
  for (; c + 4 <= D; c += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* q = p + (long long)(c + u) * hw;
      const float v = __ldg(q);
      const float vr = r ? __ldg(q + 1) : 0.f;
      const float vd = dn ? __ldg(q + w) : 0.f;
      const float vdr = (dn && r) ? __ldg(q + w + 1) : 0.f;
      const float vdl = (dn && l) ? __ldg(q + w - 1) : 0.f;
      s[u][0] = fmaf(v, v, s[u][0]); s[u][1] = fmaf(v, vr, s[u][1]); s[u][2] = fmaf(v, vd, s[u][2]);
      s[u][3] = fmaf(v, vdr, s[u][3]); s[u][4] = fmaf(v, vdl, s[u][4]);
    }
  }
  for (; c < D; ++c) {
    const float* q = p + (long long)c * hw;
    const float v = __ldg(q);
    s[0][0] = fmaf(v, v, s[0][0]);
    if (r) s[0][1] = fmaf(v, __ldg(q + 1), s[0][1]);
    if (dn) s[0][2] = fmaf(v, __ldg(q + w), s[0][2]);
    if (dn && r) s[0][3] = fmaf(v, __ldg(q + w + 1), s[0][3]);
    if (dn && l) s[0][4] = fmaf(v, __ldg(q + w - 1), s[0][4]);
  }
  float* o = gram + (((long long)b * h + y) * w + x) * 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) o[k] = (s[0][k] + s[1][k]) + (s[2][k] + s[3][k]);
}

struct LossParams {
  const float* logits;     // [B*h*w][c_pad]
  const float* gram;       // [B*h*w][5]
  const long long* label;  // [B][H][W]
  int h, w, c_pad, H, W, C;
  int off_c, cnt_c, off_l, cnt_l;
  float scale_h, scale_w;
  int rows_per_block, row_shift;
  int slab_cols;           // token columns a warp can touch (+ slack); 0 = forward only
  double* sums;            // [0] sum over valid pixels of (lse - v_label), [1] sum over all pixels of picked cosine
  unsigned long long* n_valid;
  float* g_logits;         // [B*h*w][c_pad], caller-zeroed: d(sum CE)/d logits (linear channels), d(sum cos)/d logits (cluster)
};

template <int CMAX, bool GRAD>
__global__ void __launch_bounds__(320, 1)
probe_losses_rows_kernel(const LossParams p) {
  extern __shared__ float s_slab[];          // GRAD: [warps][2][slab_cols][c_pad]
  __shared__ double s_sum[2][16];
  __shared__ unsigned int s_cnt[16];
  const int h = p.h, w = p.w, c_pad = p.c_pad, H = p.H, W = p.W, C = p.C;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ya = (int)blockIdx.x * p.rows_per_block - p.row_shift;
  const int Y0 = max(Ya, 0);
  const int Y1 = min(H, Ya + p.rows_per_block);
  const float* base = p.logits + (long long)b * h * w * c_pad;
  const float* gbase = p.gram + (long long)b * h * w * 5;
  const int Wpad = (W + 31) & ~31;
  double acc_ce = 0.0, acc_cos = 0.0;
  unsigned int n_valid = 0;
  float* slab = GRAD ? s_slab + (size_t)warp * 2 * p.slab_cols * c_pad : nullptr;
  for (int X = threadIdx.x; X < Wpad; X += blockDim.x) {
    const bool live = X < W;
    float sx = p.scale_w * ((float)X + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    int x0 = (int)sx;
    if (x0 > w - 1) x0 = w - 1;
    const int x1 = x0 + ((x0 < w - 1) ? 1 : 0);
    const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
    const int xb = __shfl_sync(0xffffffffu, x0, 0);            // first token column of this warp's slab
    if constexpr (GRAD) {
      for (int i = lane; i < 2 * p.slab_cols * c_pad; i += 32) slab[i] = 0.f;
      __syncwarp();
    }
    int cy0 = -1, cy1 = -1;
#pragma unroll 1
    for (int hd = 0; hd < 2; ++hd) {
      const int off = hd == 0 ? p.off_c : p.off_l, cnt = hd == 0 ? p.cnt_c : p.cnt_l;
      float H0[CMAX], H1[CMAX];
      float G0[GRAD ? CMAX : 1], G1[GRAD ? CMAX : 1];
      float g_s00 = 0.f, g_s01 = 0.f, g_s10 = 0.f, g_s11 = 0.f, g_h0 = 0.f, g_h1 = 0.f, g_v0 = 0.f, g_v1 = 0.f, g_dd = 0.f, g_aa = 0.f;
      cy0 = -1;
      auto flush = [&]() {
        // add this thread's accumulated d/dH0, d/dH1 into the warp's slab: lanes that share a token column take turns
        if constexpr (GRAD) {
        if (cy0 < 0) return;
#pragma unroll 1
        for (int tap = 0; tap < 2; ++tap) {
          const int xt = tap == 0 ? x0 : x1;
          const float wt = tap == 0 ? lx0 : lx1;
          const unsigned peers = __match_any_sync(0xffffffffu, xt);
          const int rank = __popc(peers & ((1u << lane) - 1u));
          int rounds = __popc(peers);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
          float* s0 = slab + (size_t)(0 * p.slab_cols + (xt - xb)) * c_pad + off;
          float* s1 = slab + (size_t)(1 * p.slab_cols + (xt - xb)) * c_pad + off;
          for (int rd = 0; rd < rounds; ++rd) {
            if (rank == rd && live) {
#pragma unroll
              for (int j = 0; j < CMAX; ++j) {
                if (j < cnt) { s0[j] += wt * G0[j]; s1[j] += wt * G1[j]; }
              }
            }
            __syncwarp();
          }
        }
#pragma unroll
        for (int j = 0; j < CMAX; ++j) { G0[j] = 0.f; G1[j] = 0.f; }
        }
      };
      auto to_global = [&]() {
        // the warp's slab (rows cy0 / cy1, columns xb ..) -> global gradient table, non-zero entries only; slab cleared
        if constexpr (GRAD) {
        if (cy0 < 0) return;
        __syncwarp();
        for (int i = lane; i < 2 * p.slab_cols * c_pad; i += 32) {
          const int row = i / (p.slab_cols * c_pad), rem = i - row * p.slab_cols * c_pad;
          const int col = rem / c_pad, ch = rem - col * c_pad;
          const float vv = slab[i];
          if (vv != 0.f && xb + col < w)
            atomicAdd(p.g_logits + (((long long)b * h + (row == 0 ? cy0 : cy1)) * w + xb + col) * c_pad + ch, vv);
          slab[i] = 0.f;
        }
        __syncwarp();
        }
      };
#pragma unroll 1
      for (int r = 0; r < kLossRows; ++r) {
        const int Y = Ya + r;
        if (Y < Y0) continue;
        if (Y >= Y1) break;
        float sy = p.scale_h * ((float)Y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
        int y0 = (int)sy;
        if (y0 > h - 1) y0 = h - 1;
        const int y1 = y0 + ((y0 < h - 1) ? 1 : 0);
        const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
        if (y0 != cy0) {
          if (GRAD && cy0 >= 0) { flush(); to_global(); }   // the block spans two token-row pairs (non-integer scale)
          cy0 = y0; cy1 = y1;
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
#pragma unroll
          for (int j = CMAX - 3; j < CMAX; ++j)
            if (j >= cnt) { H0[j] = -1e30f; H1[j] = -1e30f; }
          if constexpr (GRAD) {
#pragma unroll
            for (int j = 0; j < CMAX; ++j) { G0[j] = 0.f; G1[j] = 0.f; }
          }
          if (hd == 0) {
            // Gram terms of the four taps (border: coinciding taps fall back to the self / edge terms)
            const float* g00 = gbase + ((long long)y0 * w + x0) * 5;
            const float* g01 = gbase + ((long long)y0 * w + x1) * 5;
            const float* g10 = gbase + ((long long)y1 * w + x0) * 5;
            const float* g11 = gbase + ((long long)y1 * w + x1) * 5;
            const bool dx = x1 != x0, dy = y1 != y0;
            g_s00 = __ldg(g00); g_s01 = __ldg(g01); g_s10 = __ldg(g10); g_s11 = __ldg(g11);
            g_h0 = dx ? __ldg(g00 + 1) : g_s00;
            g_h1 = dx ? __ldg(g10 + 1) : g_s10;
            g_v0 = dy ? __ldg(g00 + 2) : g_s00;
            g_v1 = dy ? __ldg(g01 + 2) : g_s01;
            g_dd = (dx && dy) ? __ldg(g00 + 3) : (dx ? g_h0 : (dy ? g_v0 : g_s00));
            g_aa = (dx && dy) ? __ldg(g01 + 4) : (dx ? g_h0 : (dy ? g_v1 : g_s01));
          }
        }
        if (!live) continue;
        float v[CMAX];
        float best = -INFINITY;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < CMAX; ++j) {
          v[j] = ly0 * H0[j] + ly1 * H1[j];
          if (v[j] > best) { best = v[j]; bj = j; }
        }
        if (hd == 0) {
          // cosine of the winning cluster: interpolated inner product over the norm of the interpolated feature
          const float a = ly0 * lx0, bq = ly0 * lx1, c = ly1 * lx0, e = ly1 * lx1;
          float n2 = a * a * g_s00 + bq * bq * g_s01 + c * c * g_s10 + e * e * g_s11 +
                     2.f * (a * bq * g_h0 + c * e * g_h1 + a * c * g_v0 + bq * e * g_v1 + a * e * g_dd + bq * c * g_aa);
          const float nrm = fmaxf(sqrtf(fmaxf(n2, 0.f)), 1e-12f);
          const float inv = 1.f / nrm;
          acc_cos += (double)(best * inv);
          if constexpr (GRAD) {
#pragma unroll
            for (int j = 0; j < CMAX; ++j) {
              const float gj = (j == bj) ? inv : 0.f;
              G0[j] = fmaf(ly0, gj, G0[j]); G1[j] = fmaf(ly1, gj, G1[j]);
            }
          }
        } else {
          const long long lab = __ldcs(p.label + ((long long)b * H + Y) * W + X);
          if (lab >= 0 && lab < C) {
            float s = 0.f, vl = 0.f;
#pragma unroll
            for (int j = 0; j < CMAX; ++j) {
              v[j] = __expf(v[j] - best);
              s += v[j];
              vl = (j == (int)lab) ? (ly0 * H0[j] + ly1 * H1[j]) : vl;
            }
            acc_ce += (double)((best + __logf(s)) - vl);
            ++n_valid;
            if constexpr (GRAD) {
              const float is = 1.f / s;
#pragma unroll
              for (int j = 0; j < CMAX; ++j) {
                const float gj = v[j] * is - ((j == (int)lab) ? 1.f : 0.f);
                G0[j] = fmaf(ly0, gj, G0[j]); G1[j] = fmaf(ly1, gj, G1[j]);
              }
            }
          }
        }
      }
      flush();
      to_global();        // heads own disjoint channels, so flushing per head costs no extra atomics
    }
  }
  acc_ce = warp_sum_d(acc_ce); acc_cos = warp_sum_d(acc_cos);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
  if (lane == 0) { s_sum[0][warp] = acc_ce; s_sum[1][warp] = acc_cos; s_cnt[warp] = n_valid; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0; unsigned int n = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s_sum[0][i]; c += s_sum[1][i]; n += s_cnt[i]; }
    if (n) { atomicAdd(p.sums, a); atomicAdd(p.n_valid, (unsigned long long)n); }
    atomicAdd(p.sums + 1, c);
  }
}

}  // namespace equss

using namespace equss;

extern "C" int equss_token_gram(const float* feat, int B, int D, int h, int w, float* gram, void* stream) {
  EQUSS_REQUIRE(feat && gram, EQUSS_ERR_INVALID_ARG, "equss_token_gram: null pointer");
  EQUSS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0 && h <= 65535 && B <= 65535, EQUSS_ERR_INVALID_ARG, "equss_token_gram: bad shape");
  dim3 grid((unsigned)((w + 127) / 128), (unsigned)h, (unsigned)B);
  token_gram_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(feat, D, h, w, gram);
  EQUSS_LAUNCH_OK("token_gram_kernel");
  return EQUSS_OK;
}

extern "C" int equss_probe_losses_supported(int h, int w, int H, int W, int c_total, int cnt_cluster, int cnt_linear,
                                            int off_cluster, int off_linear) {
  const int cm = ((cnt_cluster > cnt_linear ? cnt_cluster : cnt_linear) + 3) & ~3;
  const int cmin = cnt_cluster < cnt_linear ? cnt_cluster : cnt_linear;
  if (h <= 0 || w <= 0 || H <= 0 || W <= 0 || c_total <= 0) return 0;
  if ((off_cluster & 3) || (off_linear & 3) || cm > 32 || cmin <= cm - 4) return 0;
  const int threads = ((W + 31) & ~31) > 320 ? 320 : ((W + 31) & ~31);
  const int slab_cols = (int)(32.0 * w / W) + 3;
  const size_t smem = (size_t)(threads / 32) * 2 * slab_cols * equss_probe_cpad(c_total) * sizeof(float);
  return smem <= 160 * 1024 ? 1 : 0;
}

extern "C" int equss_probe_losses(const float* logits, const float* gram, int B, int h, int w, int c_total,
                                  const int64_t* label, int H, int W, int num_classes, int off_cluster, int cnt_cluster,
                                  int off_linear, int cnt_linear, double* sums, uint64_t* n_valid, float* grad_logits,
                                  void* stream) {
  EQUSS_REQUIRE(logits && gram && label && sums && n_valid, EQUSS_ERR_INVALID_ARG, "equss_probe_losses: null pointer");
  EQUSS_REQUIRE(B > 0 && B <= 65535, EQUSS_ERR_INVALID_ARG, "equss_probe_losses: bad batch %d", B);
  EQUSS_REQUIRE(equss_probe_losses_supported(h, w, H, W, c_total, cnt_cluster, cnt_linear, off_cluster, off_linear),
                EQUSS_ERR_UNSUPPORTED, "equss_probe_losses: heads must start at multiples of four channels, hold <= 32 channels "
                "each and end in the same group of four");
  LossParams p;
  memset(&p, 0, sizeof(p));
  p.logits = logits; p.gram = gram; p.label = (const long long*)label;
  p.h = h; p.w = w; p.c_pad = equss_probe_cpad(c_total); p.H = H; p.W = W; p.C = num_classes;
  p.off_c = off_cluster; p.cnt_c = cnt_cluster; p.off_l = off_linear; p.cnt_l = cnt_linear;
  p.scale_h = (float)h / (float)H; p.scale_w = (float)w / (float)W;
  int rb = kLossRows, shift = 0;
  if (H % h == 0 && H / h >= 2 && H / h <= kLossRows) { rb = H / h; shift = rb / 2; }
  p.rows_per_block = rb; p.row_shift = shift;
  p.sums = sums; p.n_valid = (unsigned long long*)n_valid; p.g_logits = grad_logits;
  int threads = (W + 31) & ~31;
  if (threads > 320) threads = 320;
  p.slab_cols = grad_logits ? (int)(32.0 * w / W) + 3 : 0;
  const size_t smem = grad_logits ? (size_t)(threads / 32) * 2 * p.slab_cols * p.c_pad * sizeof(float) : 0;
  dim3 grid((unsigned)((H + shift + rb - 1) / rb), (unsigned)B);
  const int cm = ((cnt_cluster > cnt_linear ? cnt_cluster : cnt_linear) + 3) & ~3;
  cudaStream_t st = (cudaStream_t)stream;
#define EQUSS_LOSS_LAUNCH(CM)                                                                                        \
  if (grad_logits) {                                                                                                 \
    if (smem > 48 * 1024)                                                                                            \
      EQUSS_CUDA_OK(cudaFuncSetAttribute(probe_losses_rows_kernel<CM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    probe_losses_rows_kernel<CM, true><<<grid, threads, smem, st>>>(p);                                              \
  } else {                                                                                                           \
    probe_losses_rows_kernel<CM, false><<<grid, threads, 0, st>>>(p);                                                \
  }
  switch (cm) {
    case 4: EQUSS_LOSS_LAUNCH(4) break;   case 8: EQUSS_LOSS_LAUNCH(8) break;
    case 12: EQUSS_LOSS_LAUNCH(12) break; case 16: EQUSS_LOSS_LAUNCH(16) break;
    case 20: EQUSS_LOSS_LAUNCH(20) break; case 24: EQUSS_LOSS_LAUNCH(24) break;
    case 28: EQUSS_LOSS_LAUNCH(28) break; default: EQUSS_LOSS_LAUNCH(32) break;
  }
#undef EQUSS_LOSS_LAUNCH
  EQUSS_LAUNCH_OK("probe_losses_rows_kernel");
  return EQUSS_OK;
}
