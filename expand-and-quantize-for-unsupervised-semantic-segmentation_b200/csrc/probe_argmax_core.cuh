// probe_argmax_core.cuh -- first-maximal-index argmax of CMAX vertically interpolated probe logits
//     v[j] = fma(ly0, H0[j], ly1 * H1[j]),  j = 0 .. CMAX-1          (model/evaluator.py:71,106: F.interpolate + argmax)
// at ~2 ALU-pipe operations per channel instead of 3 (compare, select value, select index).
//
// The running argmax costs FSETP + FSEL + SEL per channel, all on the ALU pipe (16 lanes per SM sub-partition:
// two cycles per warp instruction), which is what bounds probe_argmax_rows_kernel.  Here
//   1. the channels are interpolated two at a time (FMUL2 / FFMA2 on f32x2 register pairs -- bit-identical to the
//      scalar mul.rn / fma.rn of the row kernel), and every group of four leaves only its maximum
//      (FMNMX3 + FMNMX: 0.5 ALU operations per channel); the CMAX/4 group maxima give `best` the same way;
//   2. a descending scan over the groups keeps the first three values and the index of the FIRST group whose
//      maximum equals `best` (1 FSETP + 4 SEL per group; the group's values are recomputed on the FMA pipe, which
//      is idle, so that only the group maxima stay live between the two passes);
//   3. three equality tests inside that group give the first maximal channel.
// NaN logits never win (fmaxf skips them; the row kernel's `v > best` does too); if no channel is above -inf
// (all NaN / -inf) the result is 0, like the row kernel.  -0.0 and +0.0 compare equal in both formulations.
//
// The header also compiles as plain C++ (tests/test_probe_argmax_core.py builds it with g++ and checks the
// tournament against the sequential loop on random, tied, NaN and signed-zero inputs): f32x2 is a struct there.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EQUSS_PA_HD __device__ __forceinline__
#else
#define EQUSS_PA_HD static inline
#endif

namespace equss {
namespace pa {

#if defined(__CUDA_ARCH__)
typedef unsigned long long f32x2;
EQUSS_PA_HD f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
EQUSS_PA_HD void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// t = h1 * l1 (rounded), v = fma(h0, l0, t): the association the row kernel compiles to
EQUSS_PA_HD f32x2 interp2(f32x2 h0, f32x2 h1, f32x2 l0, f32x2 l1) {
  f32x2 t, v;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(h1), "l"(l1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(h0), "l"(l0), "l"(t));
  return v;
}
// same values, but opaque to common-subexpression elimination: the second pass must RECOMPUTE a group instead of
// keeping all CMAX interpolated values live across the first pass (register pressure, see the header comment)
EQUSS_PA_HD f32x2 interp2_again(f32x2 h0, f32x2 h1, f32x2 l0, f32x2 l1) {
  f32x2 t, v;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(h1), "l"(l1));
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(h0), "l"(l0), "l"(t));
  return v;
}
#else
struct f32x2 { float lo, hi; };
EQUSS_PA_HD f32x2 pk(float lo, float hi) { f32x2 r; r.lo = lo; r.hi = hi; return r; }
EQUSS_PA_HD void upk(f32x2 v, float& lo, float& hi) { lo = v.lo; hi = v.hi; }
EQUSS_PA_HD float pa_mul_rn(float a, float b) { volatile float t = a * b; return t; }     // no contraction on the host
EQUSS_PA_HD f32x2 interp2(f32x2 h0, f32x2 h1, f32x2 l0, f32x2 l1) {
  return pk(fmaf(h0.lo, l0.lo, pa_mul_rn(h1.lo, l1.lo)), fmaf(h0.hi, l0.hi, pa_mul_rn(h1.hi, l1.hi)));
}
EQUSS_PA_HD f32x2 interp2_again(f32x2 h0, f32x2 h1, f32x2 l0, f32x2 l1) { return interp2(h0, h1, l0, l1); }
#endif

// H0p / H1p: CMAX/2 channel pairs (channel 2i in the low word).  Returns the first j with v[j] == max_j v[j].
template <int CMAX, bool RECOMPUTE>
EQUSS_PA_HD int argmax_interp(const f32x2 (&H0p)[CMAX / 2], const f32x2 (&H1p)[CMAX / 2], float ly0, float ly1) {
  static_assert(CMAX % 4 == 0 && CMAX >= 4, "channels come in groups of four");
  constexpr int NG = CMAX / 4;
  const f32x2 l0 = pk(ly0, ly0), l1 = pk(ly1, ly1);
  float gm[NG];
  float va[RECOMPUTE ? 1 : NG], vb[RECOMPUTE ? 1 : NG], vc[RECOMPUTE ? 1 : NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float a, b, c, d;
    upk(interp2(H0p[2 * g], H1p[2 * g], l0, l1), a, b);
    upk(interp2(H0p[2 * g + 1], H1p[2 * g + 1], l0, l1), c, d);
    gm[g] = fmaxf(fmaxf(fmaxf(a, b), c), d);
    if (!RECOMPUTE) { va[g] = a; vb[g] = b; vc[g] = c; }
  }
  float best = gm[0];
#pragma unroll
  for (int g = 1; g < NG; ++g) best = fmaxf(best, gm[g]);
  // descending scan: after it (a, b, c, gi) belong to the first group whose maximum equals best (or to the last
  // group, if none does: only when every value is NaN, which the final guard maps to 0)
  float a, b, c, d;
  int gi = NG - 1;
  if (RECOMPUTE) {
    upk(interp2_again(H0p[2 * (NG - 1)], H1p[2 * (NG - 1)], l0, l1), a, b);
    upk(interp2_again(H0p[2 * (NG - 1) + 1], H1p[2 * (NG - 1) + 1], l0, l1), c, d);
  } else {
    a = va[NG - 1]; b = vb[NG - 1]; c = vc[NG - 1];
  }
#pragma unroll
  for (int g = NG - 2; g >= 0; --g) {
    float a2, b2, c2, d2;
    if (RECOMPUTE) {
      upk(interp2_again(H0p[2 * g], H1p[2 * g], l0, l1), a2, b2);
      upk(interp2_again(H0p[2 * g + 1], H1p[2 * g + 1], l0, l1), c2, d2);
    } else {
      a2 = va[g]; b2 = vb[g]; c2 = vc[g];
    }
    const bool hit = gm[g] == best;
    a = hit ? a2 : a;
    b = hit ? b2 : b;
    c = hit ? c2 : c;
    gi = hit ? g : gi;
  }
  (void)d;
  const int l = (a == best) ? 0 : (b == best) ? 1 : (c == best) ? 2 : 3;
  return (best > -INFINITY) ? gi * 4 + l : 0;   // all NaN / -inf: nothing beats the initial -inf of the sequential loop
}

// the sequential formulation of probe_argmax_rows_kernel (the definition of the result)
template <int CMAX>
EQUSS_PA_HD int argmax_interp_sequential(const f32x2 (&H0p)[CMAX / 2], const f32x2 (&H1p)[CMAX / 2], float ly0, float ly1) {
  const f32x2 l0 = pk(ly0, ly0), l1 = pk(ly1, ly1);
  float best = -INFINITY;
  int bj = 0;
#pragma unroll
  for (int i = 0; i < CMAX / 2; ++i) {
    float x, y;
    upk(interp2(H0p[i], H1p[i], l0, l1), x, y);
    if (x > best) { best = x; bj = 2 * i; }
    if (y > best) { best = y; bj = 2 * i + 1; }
  }
  return bj;
}

}  // namespace pa
}  // namespace equss
