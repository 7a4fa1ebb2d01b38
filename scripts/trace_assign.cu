// Pipeline timeline of the fp16-split assign kernel: per-unit clock64 stamps of CTA 0.  Debug tool, not shipped.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DEQUSS_TRACE -lcuda \
//        -o gpurun_out/trace_assign scripts/trace_assign.cu && gpurun -- gpurun_out/trace_assign
#include "../expand-and-quantize-for-unsupervised-semantic-segmentation_b200/csrc/equss_core.cu"
#include "../expand-and-quantize-for-unsupervised-semantic-segmentation_b200/csrc/pq_assign_h.cu"
#include "../expand-and-quantize-for-unsupervised-semantic-segmentation_b200/csrc/pq_assign_h_d16.cu"
#include "../expand-and-quantize-for-unsupervised-semantic-segmentation_b200/csrc/pq_assign_h_d32.cu"
#include "../expand-and-quantize-for-unsupervised-semantic-segmentation_b200/csrc/pq_assign_h_d64.cu"
#include <vector>
#include <cstdlib>
#include <cmath>
using namespace equss;
__global__ void cn2_k(const float* cb, int rows, int d, float* o) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) { float s = 0; for (int j = 0; j < d; ++j) s += cb[r * d + j] * cb[r * d + j]; o[r] = s; }
}
int main(int argc, char** argv) {
  int M = 64, K = 256, d = 16; long long N = 51200;
  if (argc > 1 && atoi(argv[1]) == 64) { M = 16; K = 512; d = 64; N = 50176; }
  int D = M * d;
  std::vector<float> hz((size_t)N * D), hc((size_t)M * K * d);
  srand(1);
  for (auto& v : hz) v = (rand() / (float)RAND_MAX) * 2 - 1;
  for (size_t r = 0; r < (size_t)M * K; ++r) { float s = 0; for (int j = 0; j < d; ++j) { float v = (rand() / (float)RAND_MAX) * 2 - 1; hc[r * d + j] = v; s += v * v; } s = 1 / sqrtf(s); for (int j = 0; j < d; ++j) hc[r * d + j] *= s; }
  float *z, *cb, *cn2; int32_t* idx; void* ws;
  cudaMalloc(&z, hz.size() * 4); cudaMalloc(&cb, hc.size() * 4); cudaMalloc(&cn2, (size_t)M * K * 4); cudaMalloc(&idx, (size_t)M * N * 4);
  cudaMemcpy(z, hz.data(), hz.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(cb, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
  cn2_k<<<(M * K + 255) / 256, 256>>>(cb, M * K, d, cn2);
  equss_zdesc zd; zd.n_pixels = N; zd.hw = N; zd.stride_b = N * D; zd.stride_s = D; zd.stride_c = 1; zd.dim = D; zd.layout = 0;
  if (argc > 3 && atoi(argv[3]) == 1) {     // NCHW: 32 images of N/32 pixels (C2: 40x40), the same buffer read as (B, D, hw)
    const long long hw = N / 32;
    zd.hw = hw; zd.stride_b = hw * D; zd.stride_s = 1; zd.stride_c = hw; zd.layout = 1;
  }
  long long wsb = assign_tch_workspace_bytes(N, M, K, d);
  cudaMalloc(&ws, wsb);
  const bool fuse = argc > 2 && atoi(argv[2]) == 1;
  float* outp; double* sq; cudaMalloc(&outp, hz.size() * 4); cudaMalloc(&sq, M * 8); cudaMemset(sq, 0, M * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int rc = fuse ? assign_tch_launch(z, &zd, cb, cn2, M, K, d, idx, ws, wsb, 0, cb, outp, sq)
                  : assign_tch_launch(z, &zd, cb, cn2, M, K, d, idx, ws, wsb, 0);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rc=%d %s  %.1f us  err=%s\n", rc, equss_last_error_string(), ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
#ifdef EQUSS_TRACE
  static long long tr[256 * 12];
  cudaMemcpyFromSymbol(tr, tch::g_trace, sizeof(tr));
  long long t0 = tr[0];
  printf("unit | conv: a_empty raw_full done | mma: loop_top issue | epi: tfull_h0 halves_done tail_done   (cycles since unit-0 convert start)\n");
  for (int i = 0; i < 40; ++i) {
    long long* r = tr + i * 12;
    printf("%3d | %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld %7lld | g: %7lld %7lld %7lld %7lld\n", i, r[0] - t0, r[1] - t0, r[2] - t0, r[7] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0, r[8] - t0, r[9] - t0, r[11] - t0, r[10] - t0);
  }
  if (!fuse) {
    int last = 0;
    for (int i = 0; i < 256; ++i) if (tr[i * 12 + 6] != 0) last = i;
    printf("units of CTA 0: %d; unit 2 -> unit %d: %lld cycles, %lld ns  => SM clock %.3f GHz\n", last + 1, last,
           tr[last * 12 + 6] - tr[2 * 12 + 6], tr[last * 12 + 8] - tr[2 * 12 + 8],
           (double)(tr[last * 12 + 6] - tr[2 * 12 + 6]) / (double)(tr[last * 12 + 8] - tr[2 * 12 + 8]));
  }
  for (int i = 150; i < 170; ++i) {
    long long* r = tr + i * 12;
    printf("%3d | %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld %7lld\n", i, r[0] - t0, r[1] - t0, r[2] - t0, r[7] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0);
  }
#endif
#ifdef EQUSS_TRACE_Q
  static long long tq[256 * 24];
  cudaMemcpyFromSymbol(tq, tch::g_traceq, sizeof(tq));
  printf("unit | per quarter: tfull_h0 rel_h0 tfull_h1 rel_h1 tail\n");
  for (int i = 150; i < 166; ++i) {
    printf("%3d |", i);
    for (int q = 0; q < 4; ++q) {
      long long* r = tq + i * 24 + q * 6;
      printf(" %7lld %5lld %5lld %5lld %5lld |", r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3]);
    }
    printf("\n");
  }
#endif
  return 0;
}
