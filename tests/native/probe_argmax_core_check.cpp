// Host check of csrc/probe_argmax_core.cuh: the tournament argmax must return exactly what the sequential
// first-maximal-index loop of probe_argmax_rows_kernel returns, for every channel count the kernel is built for.
// Built and run by tests/test_probe_argmax_core.py with g++ (-ffp-contract=off); prints "ok <cases>" on success.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <limits>
#include "probe_argmax_core.cuh"

using namespace equss::pa;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (uint32_t)(rng_state >> 16);
}
static float urand(float a, float b) { return a + (b - a) * (float)(rnd() & 0xFFFFFF) / 16777216.f; }

// kind: 0 random, 1 few distinct levels (many exact ties), 2 with NaN / inf / signed zeros, 3 all equal,
// 4 tail channels padded with -1e30 like the kernel does, 5 every value NaN
template <int CMAX>
static long run(int iters) {
  long cases = 0;
  const float special[] = {0.f, -0.f, INFINITY, -INFINITY, NAN, 1.f, -1.f, 1e-40f, -1e-40f, 3.4e38f, -1e30f};
  for (int it = 0; it < iters; ++it) {
    const int kind = it % 6;
    float h0[CMAX], h1[CMAX];
    for (int j = 0; j < CMAX; ++j) {
      switch (kind) {
        case 0: h0[j] = urand(-8.f, 8.f); h1[j] = urand(-8.f, 8.f); break;
        case 1: h0[j] = (float)(rnd() % 3) - 1.f; h1[j] = (float)(rnd() % 3) - 1.f; break;
        case 2: h0[j] = (rnd() % 4 == 0) ? special[rnd() % 11] : urand(-2.f, 2.f);
                h1[j] = (rnd() % 4 == 0) ? special[rnd() % 11] : urand(-2.f, 2.f); break;
        case 3: h0[j] = 0.25f; h1[j] = -0.5f; break;
        case 4: h0[j] = urand(-1.f, 1.f); h1[j] = urand(-1.f, 1.f); break;
        default: h0[j] = NAN; h1[j] = NAN; break;
      }
    }
    if (kind == 4) { const int cnt = CMAX - (int)(rnd() % 4); for (int j = cnt; j < CMAX; ++j) h0[j] = h1[j] = -1e30f; }
    if (kind == 3 && (it & 8)) { const int j = rnd() % CMAX; h0[j] = 0.25f + 1e-3f; }       // one strict winner
    f32x2 H0p[CMAX / 2], H1p[CMAX / 2];
    for (int i = 0; i < CMAX / 2; ++i) { H0p[i] = pk(h0[2 * i], h0[2 * i + 1]); H1p[i] = pk(h1[2 * i], h1[2 * i + 1]); }
    // vertical weights of an 8x upsampling (and a few arbitrary ones)
    const float ly1 = (it % 7 == 0) ? urand(0.f, 1.f) : ((float)(rnd() % 8) + 0.5f) / 8.f - ((rnd() & 1) ? 0.f : 0.0625f);
    const float ly0 = 1.f - ly1;
    const int want = argmax_interp_sequential<CMAX>(H0p, H1p, ly0, ly1);
    const int got_r = argmax_interp<CMAX, true>(H0p, H1p, ly0, ly1);
    const int got_k = argmax_interp<CMAX, false>(H0p, H1p, ly0, ly1);
    if (got_r != want || got_k != want) {
      std::printf("MISMATCH CMAX=%d kind=%d it=%d: sequential %d, tournament %d (recompute) %d (keep)\n", CMAX, kind, it,
                  want, got_r, got_k);
      std::exit(1);
    }
    ++cases;
  }
  return cases;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? std::atoi(argv[1]) : 200000;
  long n = 0;
  n += run<4>(iters); n += run<8>(iters); n += run<12>(iters); n += run<16>(iters);
  n += run<20>(iters); n += run<24>(iters); n += run<28>(iters); n += run<32>(iters);
  std::printf("ok %ld\n", n);
  return 0;
}
