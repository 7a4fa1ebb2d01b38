/* equss_oracle_c.c -- plain-C restatement of the index / integer side of the EQUSS product-quantization hot path.
 *
 * TEST INFRASTRUCTURE ONLY (like oracle/equss_oracle.py): tests/ may load it as a second, independent checker of
 * the golden fixtures; the product path never links or calls it.  Built by oracle/build_c.py (gcc, no dependencies)
 * into oracle/_build/libequss_oracle_c.so.
 *
 * Every function restates, in scalar fp32 / int64 C, the arithmetic of the reference lines it cites
 * (pitlover/Expand-and-Quantize-for-Unsupervised-Semantic-Segmentation).  Float sums run in index order, so results
 * agree with the reference (PyTorch / MKL) up to fp32 summation order; the fixtures under tests/golden contain no
 * fp32 near-ties, so indices, counts and histograms must agree exactly (tests/test_oracle_c.py).
 * Parity: pinned through those fixtures (outputs of the unmodified reference run by oracle/make_golden*.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EQO_NORM_NONE 0
#define EQO_NORM_L2 1
#define EQO_NORM_ZNORM 2

/* model/quantizer.py:419-428: one row of z or of the codebook, normalised into out[d]. */
static void eqo_normalize_row(const float* x, int d, int mode, float* out) {
  if (mode == EQO_NORM_L2) {                       /* F.normalize: x / max(||x||_2, 1e-12)  (:421) */
    float s = 0.f;
    for (int j = 0; j < d; ++j) s += x[j] * x[j];
    float nrm = sqrtf(s);
    if (nrm < 1e-12f) nrm = 1e-12f;
    for (int j = 0; j < d; ++j) out[j] = x[j] / nrm;
  } else if (mode == EQO_NORM_ZNORM) {             /* (x - mean) / (unbiased std + 1e-5)  (:424-428) */
    float m = 0.f;
    for (int j = 0; j < d; ++j) m += x[j];
    m /= (float)d;
    float v = 0.f;
    for (int j = 0; j < d; ++j) v += (x[j] - m) * (x[j] - m);
    const float sd = sqrtf(v / (float)(d - 1));
    for (int j = 0; j < d; ++j) out[j] = (x[j] - m) / (sd + 1e-5f);
  } else {
    memcpy(out, x, sizeof(float) * (size_t)d);
  }
}

/* model/quantizer.py:457-476,514,534-536 for all M subspaces of flat rows z[n][D] (D = M * d):
 *   distance = sum(z_norm^2) + sum(c_norm^2) - 2 <z_norm, c_norm>,  idx = first minimal k  (:457-467)
 *   q = c_norm[idx]  (gather_raw == 0, :474 with update_norm)  or  raw codebook[idx]  (gather_raw != 0,
 *       dino_new_vq.py:403 / dino_pqgo.py:665)
 *   out = z_norm + (q - z_norm)  (:536),   sqerr[m] = sum (z_norm - q)^2  (mse = sqerr / (n * d), :514)
 * codebook: [M][K][d] raw.  idx: [M][n].  out: [n][D] or NULL.  sqerr: [M] doubles or NULL.
 * keep (optional, [M][K] bytes): pq_dropout's keep mask (dino_new_vq.py:388-391): distances run over the kept codes
 *   only and idx is the position in the kept list, which then addresses the FULL codebook in the gather. */
int eqo_pq_assign_gather(const float* z, int64_t n, int M, int K, int d, const float* codebook, int mode, int gather_raw,
                         const uint8_t* keep, int32_t* idx, float* out, double* sqerr) {
  if (n < 0 || M <= 0 || K <= 0 || d <= 0 || !z || !codebook || !idx) return -1;
  const int D = M * d;
  float* cn = (float*)malloc(sizeof(float) * (size_t)K * d);
  float* cn2 = (float*)malloc(sizeof(float) * (size_t)K);
  int* kept = (int*)malloc(sizeof(int) * (size_t)K);
  float* zn = (float*)malloc(sizeof(float) * (size_t)d);
  if (!cn || !cn2 || !kept || !zn) { free(cn); free(cn2); free(kept); free(zn); return -2; }
  for (int m = 0; m < M; ++m) {
    const float* cb = codebook + (size_t)m * K * d;
    int nk = 0;
    for (int k = 0; k < K; ++k) {
      eqo_normalize_row(cb + (size_t)k * d, d, mode, cn + (size_t)k * d);
      float s = 0.f;
      for (int j = 0; j < d; ++j) s += cn[(size_t)k * d + j] * cn[(size_t)k * d + j];
      cn2[k] = s;
      if (!keep || keep[(size_t)m * K + k]) kept[nk++] = k;
    }
    double se = 0.0;
    for (int64_t r = 0; r < n; ++r) {
      eqo_normalize_row(z + (size_t)r * D + (size_t)m * d, d, mode, zn);
      float zz = 0.f;
      for (int j = 0; j < d; ++j) zz += zn[j] * zn[j];
      float best = INFINITY;
      int bi = 0;
      for (int t = 0; t < nk; ++t) {
        const float* c = cn + (size_t)kept[t] * d;
        float dot = 0.f;
        for (int j = 0; j < d; ++j) dot += zn[j] * c[j];
        const float dist = (zz + cn2[kept[t]]) - 2.f * dot;
        if (dist < best) { best = dist; bi = t; }             /* strict <: the first minimal index wins (torch.argmin) */
      }
      idx[(size_t)m * n + r] = bi;
      const float* q = gather_raw ? cb + (size_t)bi * d : cn + (size_t)bi * d;   /* position in the kept list -> full table */
      for (int j = 0; j < d; ++j) {
        const float diff = q[j] - zn[j];
        if (out) out[(size_t)r * D + (size_t)m * d + j] = zn[j] + diff;
        se += (double)(zn[j] - q[j]) * (double)(zn[j] - q[j]);
      }
    }
    if (sqerr) sqerr[m] = se;
  }
  free(cn); free(cn2); free(kept); free(zn);
  return 0;
}

/* model/quantizer.py:485-488: per-code selection counts and sums of the RAW rows (one_hot.sum(0), one_hot^T @ z_flat).
 * count: [M][K] (exact integers in fp32), sum: [M][K][d]; both overwritten. */
int eqo_counts_sums(const float* z, int64_t n, int M, int K, int d, const int32_t* idx, float* count, float* sum) {
  if (!z || !idx || !count || !sum) return -1;
  const int D = M * d;
  memset(count, 0, sizeof(float) * (size_t)M * K);
  memset(sum, 0, sizeof(float) * (size_t)M * K * d);
  for (int m = 0; m < M; ++m)
    for (int64_t r = 0; r < n; ++r) {
      const int k = idx[(size_t)m * n + r];
      if (k < 0 || k >= K) return -3;
      count[(size_t)m * K + k] += 1.f;
      for (int j = 0; j < d; ++j) sum[((size_t)m * K + k) * d + j] += z[(size_t)r * D + (size_t)m * d + j];
    }
  return 0;
}

/* EmbeddingEMA.update, model/quantizer.py:233-254, one subspace:
 *   vq_count <- g vq_count + (1-g) count;  weight_avg <- g weight_avg + (1-g) sum;  n = sum(vq_count)
 *   weight = weight_avg / ((vq_count + eps) / (n + K eps) * n) */
int eqo_ema_update(const float* count, const float* sum, int K, int d, float decay, float eps,
                   float* vq_count, float* weight_avg, float* weight) {
  if (!count || !sum || !vq_count || !weight_avg || !weight) return -1;
  float n = 0.f;
  for (int k = 0; k < K; ++k) {
    vq_count[k] = vq_count[k] * decay + (1.f - decay) * count[k];
    n += vq_count[k];
  }
  for (int k = 0; k < K; ++k) {
    const float smoothed = (vq_count[k] + eps) / (n + (float)K * eps) * n;
    for (int j = 0; j < d; ++j) {
      weight_avg[(size_t)k * d + j] = weight_avg[(size_t)k * d + j] * decay + (1.f - decay) * sum[(size_t)k * d + j];
      weight[(size_t)k * d + j] = weight_avg[(size_t)k * d + j] / smoothed;
    }
  }
  return 0;
}

static int eqo_cmp_desc(const void* a, const void* b) {
  const float x = *(const float*)a, y = *(const float*)b;
  return (x < y) - (x > y);
}

/* get_histogram_count, model/quantizer.py:15-30: prob = count / (sum + 1), sorted descending, cumulative sum; the first
 * rank whose cumulative probability reaches 0.1 / 0.5 / 0.9, divided by K; NaN where the level is never reached. */
int eqo_usage_percentiles(const float* count, int K, float* p3) {
  if (!count || !p3 || K <= 0) return -1;
  float* prob = (float*)malloc(sizeof(float) * (size_t)K);
  if (!prob) return -2;
  float total = 0.f;
  for (int k = 0; k < K; ++k) total += count[k];
  for (int k = 0; k < K; ++k) prob[k] = count[k] / (total + 1.f);
  qsort(prob, (size_t)K, sizeof(float), eqo_cmp_desc);
  const float level[3] = {0.1f, 0.5f, 0.9f};
  for (int t = 0; t < 3; ++t) p3[t] = NAN;
  float c = 0.f;
  for (int k = 0; k < K; ++k) {
    c += prob[k];
    for (int t = 0; t < 3; ++t)
      if (isnan(p3[t]) && c >= level[t]) p3[t] = (float)k / (float)K;
  }
  free(prob);
  return 0;
}

/* UnSegMetrics.update, model/metric.py:44-58: conf[pred][label] += 1 for 0 <= label < C and 0 <= pred < C
 * (predictions >= C are dropped even when extra classes exist, :49).  conf: [C + extra][C] int64, accumulated. */
int eqo_confusion_update(const int64_t* preds, const int64_t* label, int64_t P, int C, int extra, int64_t* conf) {
  if (!preds || !label || !conf || C <= 0 || extra < 0) return -1;
  for (int64_t i = 0; i < P; ++i) {
    const int64_t l = label[i], p = preds[i];
    if (l >= 0 && l < C && p >= 0 && p < C) conf[p * C + l] += 1;
  }
  return 0;
}

/* data/precompute_knns.py:313-315: cosine similarities of L2-normalised rows, indices of the k largest per query
 * (larger similarity first; lower index first among equal similarities).  q: [nq][F], db: [n][F], idx: [nq][k]. */
int eqo_knn_topk(const float* q, int64_t nq, const float* db, int64_t n, int F, int k, int64_t* idx) {
  if (!q || !db || !idx || k <= 0 || k > n) return -1;
  float* bv = (float*)malloc(sizeof(float) * (size_t)k);
  if (!bv) return -2;
  for (int64_t r = 0; r < nq; ++r) {
    int64_t* bi = idx + (size_t)r * k;
    int have = 0;
    for (int64_t c = 0; c < n; ++c) {
      float s = 0.f;
      for (int j = 0; j < F; ++j) s += q[(size_t)r * F + j] * db[(size_t)c * F + j];
      if (have == k && !(s > bv[k - 1])) continue;              /* equal to the current k-th: the earlier column stays */
      int pos = have < k ? have : k - 1;
      while (pos > 0 && s > bv[pos - 1]) { bv[pos] = bv[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
      bv[pos] = s; bi[pos] = c;
      if (have < k) ++have;
    }
  }
  free(bv);
  return 0;
}

/* Bilinear upsampling + probe argmax, model/evaluator.py:53-54,67-70,95-106 restated at TOKEN resolution:
 *   logits[b][y][x][c] (c < C) at h x w;  pred[b][Y][X] = argmax_c of the bilinear (align_corners=False) interpolation
 *   of the logits at label pixel (Y, X) -- for the cluster probe the argmax of <normalize(interp(x)), c_j> equals the
 *   argmax of the interpolated raw inner products (a positive per-pixel scale), for the linear probe interpolation and
 *   the 1x1 convolution commute.  PyTorch's source-index arithmetic: src = scale * (dst + 0.5) - 0.5, clamped at 0. */
int eqo_probe_argmax(const float* logits, int B, int h, int w, int c_stride, int C, int H, int W, int64_t* pred) {
  if (!logits || !pred || C <= 0 || C > c_stride) return -1;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  for (int b = 0; b < B; ++b)
    for (int Y = 0; Y < H; ++Y) {
      float sy = sh * ((float)Y + 0.5f) - 0.5f;
      if (sy < 0.f) sy = 0.f;
      const int y0 = (int)sy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
      const float ly = sy - (float)y0, hy = 1.f - ly;
      for (int X = 0; X < W; ++X) {
        float sx = sw * ((float)X + 0.5f) - 0.5f;
        if (sx < 0.f) sx = 0.f;
        const int x0 = (int)sx, x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float lx = sx - (float)x0, hx = 1.f - lx;
        const float* p00 = logits + (((size_t)b * h + y0) * w + x0) * c_stride;
        const float* p01 = logits + (((size_t)b * h + y0) * w + x1) * c_stride;
        const float* p10 = logits + (((size_t)b * h + y1) * w + x0) * c_stride;
        const float* p11 = logits + (((size_t)b * h + y1) * w + x1) * c_stride;
        float best = -INFINITY;
        int bi = 0;
        for (int c = 0; c < C; ++c) {
          const float v = hy * (hx * p00[c] + lx * p01[c]) + ly * (hx * p10[c] + lx * p11[c]);
          if (v > best) { best = v; bi = c; }
        }
        pred[((size_t)b * H + Y) * W + X] = bi;
      }
    }
  return 0;
}
