// Issue-rate probe: back-to-back tcgen05.mma (kind::tf32 / f16) of various N into one or two accumulators.
#include <cuda.h>
#include <cstdio>
#include "../expand-and-quantize-for-unsupervised-semantic-segmentation_b200/csrc/equss_tcgen05.cuh"
using namespace equss::ptx;
__global__ void __launch_bounds__(128, 1) k(long long* out, int N, int n_acc, int kind, int iters, int swz) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 65536 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&s_tmem);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = s_tmem;
  if (t == 0) {
    uint32_t idesc = (kind == 0) ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24))
                                 : ((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24));
    // swz = 0: K-major no swizzle (LBO 128, SBO 256); swz = 1: K-major SWIZZLE_128B (rows of 128 B, SBO 1024)
    uint32_t hi = swz ? (((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29)) : (((256u >> 4) & 0x3FFF) | (1u << 14));
    uint32_t a_lo = (smem_u32(smem) >> 4) | ((swz ? 1u : (128u >> 4)) << 16);
    uint32_t b_lo = (smem_u32(smem + 16384) >> 4) | ((swz ? 1u : (128u >> 4)) << 16);
    long long t0 = clock64();
    const uint64_t ad = desc_from(a_lo, hi), bd = desc_from(b_lo, hi);
    const uint32_t d0 = tm, d1 = tm + (uint32_t)((n_acc - 1) * N);
    if (kind == 0) {
      for (int i = 0; i < iters; i += 8) {
        umma_tf32(d0, ad, bd, idesc, 1); umma_tf32(d1, ad, bd, idesc, 1); umma_tf32(d0, ad, bd, idesc, 1); umma_tf32(d1, ad, bd, idesc, 1);
        umma_tf32(d0, ad, bd, idesc, 1); umma_tf32(d1, ad, bd, idesc, 1); umma_tf32(d0, ad, bd, idesc, 1); umma_tf32(d1, ad, bd, idesc, 1);
      }
    } else {
      for (int i = 0; i < iters; i += 8) {
        umma_f16(d0, ad, bd, idesc, 1); umma_f16(d1, ad, bd, idesc, 1); umma_f16(d0, ad, bd, idesc, 1); umma_f16(d1, ad, bd, idesc, 1);
        umma_f16(d0, ad, bd, idesc, 1); umma_f16(d1, ad, bd, idesc, 1); umma_f16(d0, ad, bd, idesc, 1); umma_f16(d1, ad, bd, idesc, 1);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  int Ns[] = {64, 128, 256};
  for (int swz = 0; swz < 2; ++swz)
  for (int kind = 0; kind < 2; ++kind)
    for (int N : Ns)
      for (int n_acc = 1; n_acc <= 2; n_acc += 1) {
        if (n_acc * N > 512) continue;
        const int iters = 2000;
        k<<<1, 128, 70000>>>(d, N, n_acc, kind, iters, swz);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        double macs = 128.0 * N * (kind == 0 ? 8 : 16);
        printf("swz=%d %s N=%3d acc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%.0f MAC/clk)\n", swz, kind == 0 ? "tf32" : "f16 ", N, n_acc,
               (double)h[0] / iters, (double)h[1] / iters, macs / ((double)h[1] / iters));
      }
  return 0;
}
