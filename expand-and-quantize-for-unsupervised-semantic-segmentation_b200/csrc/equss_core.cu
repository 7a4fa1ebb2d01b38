// equss_core.cu -- error plumbing, device checks and bookkeeping shared by all entry points.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "equss_common.cuh"

namespace equss {

static thread_local char g_err[512] = "ok";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return EQUSS_OK;
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return EQUSS_ERR_CUDA;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int validate_zdesc(const equss_zdesc* zd, int M, int d) {
  EQUSS_REQUIRE(zd != nullptr, EQUSS_ERR_INVALID_ARG, "null equss_zdesc");
  EQUSS_REQUIRE(M > 0 && d > 0, EQUSS_ERR_INVALID_ARG, "bad subspace shape M=%d d=%d", M, d);
  EQUSS_REQUIRE(d <= kMaxD, EQUSS_ERR_UNSUPPORTED, "per-subspace dim d=%d exceeds %d", d, kMaxD);
  EQUSS_REQUIRE((long long)M * d == zd->dim, EQUSS_ERR_INVALID_ARG,
                "Embed dim %d should be divisible by #PQ %d (d=%d).", zd->dim, M, d);
  EQUSS_REQUIRE(zd->n_pixels >= 0 && zd->hw > 0, EQUSS_ERR_INVALID_ARG, "bad pixel counts n=%lld hw=%lld",
                (long long)zd->n_pixels, (long long)zd->hw);
  EQUSS_REQUIRE(zd->n_pixels % zd->hw == 0, EQUSS_ERR_INVALID_ARG, "n_pixels %lld not a multiple of hw %lld",
                (long long)zd->n_pixels, (long long)zd->hw);
  EQUSS_REQUIRE(zd->n_pixels < (1LL << 31), EQUSS_ERR_UNSUPPORTED, "n_pixels %lld >= 2^31", (long long)zd->n_pixels);
  return EQUSS_OK;
}

}  // namespace equss

using namespace equss;

extern "C" const char* equss_last_error_string(void) { return g_err; }
extern "C" int equss_version(void) { return 100; /* 0.1.0 */ }
extern "C" int64_t equss_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int equss_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error("no CUDA device visible (%s); equss_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return EQUSS_ERR_NO_DEVICE;
  }
  EQUSS_REQUIRE(device >= 0 && device < count, EQUSS_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, count);
  int major = 0, minor = 0;
  EQUSS_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  EQUSS_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
    return EQUSS_ERR_NO_DEVICE;
  }
  return EQUSS_OK;
}
