// Unit probe of tcgen05.mma kind::tf32 with an MN-major, 128B-swizzled A operand (debug tool).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include "../expand-and-quantize-for-unsupervised-semantic-segmentation_b200/csrc/equss_tcgen05.cuh"
using namespace equss::ptx;

struct Cfg { uint32_t a_lbo, a_sbo, a_layout, a_major, swz; };

__global__ void __launch_bounds__(128, 1) k(const float* A /*[128][8] m-major rows*/, const float* B /*[64][8]*/, float* Dout, Cfg c) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;            // 16 KB region
  uint8_t* sb = smem + 16384;    // 64 x 8 floats K-major no swizzle
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < 16384 / 4; i += 128) reinterpret_cast<float*>(sa)[i] = 0.f;
  __syncthreads();
  // A element (m, k): MN-major: offset = (m%32)*4 + k*128 + (m/32)*LBO   (+ swizzle)
  for (int i = t; i < 128 * 8; i += 128) {
    int m = i / 8, kk = i % 8;
    uint32_t off = (m % 32) * 4 + kk * 128 + (m / 32) * c.a_lbo;
    if (c.swz == 1) off ^= ((off >> 7) & 7) << 4;
    if (c.swz == 2) off ^= ((off >> 7) & 3) << 5;
    *reinterpret_cast<float*>(sa + off) = A[m * 8 + kk];
  }
  for (int i = t; i < 64 * 8; i += 128) {
    int n = i / 8, kk = i % 8;
    uint32_t off = (n / 8) * 256 + (kk / 4) * 128 + (n % 8) * 16 + (kk % 4) * 4;
    *reinterpret_cast<float*>(sb + off) = B[n * 8 + kk];
  }
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<64>(&s_tmem);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tmem;
  if (t == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (c.a_major << 15) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t a_hi = ((c.a_sbo >> 4) & 0x3FFF) | (1u << 14) | (c.a_layout << 29);
    uint32_t a_lo = (smem_u32(sa) >> 4) | (((c.a_lbo >> 4) & 0x3FFF) << 16);
    uint32_t b_hi = ((256u >> 4) & 0x3FFF) | (1u << 14);
    uint32_t b_lo = (smem_u32(sb) >> 4) | ((128u >> 4) << 16);
    umma_tf32(tm, desc_from(a_lo, a_hi), desc_from(b_lo, b_hi), idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 1);
  tc_fence_after();
  uint32_t v[32];
  for (int h = 0; h < 2; ++h) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + h * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) Dout[(warp * 32 + lane) * 64 + h * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tm);
}

int main() {
  std::vector<float> A(128 * 8), B(64 * 8), D(128 * 64), R(128 * 64);
  for (int m = 0; m < 128; ++m) for (int kk = 0; kk < 8; ++kk) A[m * 8 + kk] = (float)((m * 7 + kk * 3) % 11) - 5.f;
  for (int n = 0; n < 64; ++n) for (int kk = 0; kk < 8; ++kk) B[n * 8 + kk] = (float)((n * 5 + kk) % 7) - 3.f;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { float s = 0; for (int kk = 0; kk < 8; ++kk) s += A[m * 8 + kk] * B[n * 8 + kk]; R[m * 64 + n] = s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  Cfg cfgs[] = {
    {4096, 512, 1, 1, 2},    // 128B_BASE32B: LBO = 32-pixel block stride, SBO = 4-channel K-atom stride
    {512, 4096, 1, 1, 2},
    {4096, 1024, 1, 1, 2},
    {4096, 512, 1, 1, 0},
    {4096, 512, 1, 1, 1},
  };
  for (auto& c : cfgs) {
    cudaMemset(dD, 0, D.size() * 4);
    k<<<1, 128, 40000>>>(dA, dB, dD, c);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, nz = 0; for (size_t i = 0; i < D.size(); ++i) { if (D[i] != R[i]) ++bad; if (D[i] != 0) ++nz; }
    printf("lbo=%u sbo=%u layout=%u major=%u swz=%u: err=%s mismatches=%d nonzero=%d  D[0][0..3]=%g %g %g %g  ref=%g %g %g %g  D[33][1]=%g ref=%g\n",
           c.a_lbo, c.a_sbo, c.a_layout, c.a_major, c.swz, cudaGetErrorString(e), bad, nz, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3], D[33 * 64 + 1], R[33 * 64 + 1]);
  }
  return 0;
}
