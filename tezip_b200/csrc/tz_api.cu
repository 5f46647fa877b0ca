// tezip_b200 -- library-level C ABI: version, error string, device probing, launch counter.
#include "tz_common.cuh"

#include <atomic>
#include <string.h>

namespace tz {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
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

}  // namespace tz

extern "C" {

int tz_abi_version(void) { return TZ_ABI_VERSION; }

const char *tz_last_error(void) { return tz::g_err; }

long long tz_launch_count(void) { return tz::g_launches.load(std::memory_order_relaxed); }

int tz_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    tz::set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TZ_ECUDA;
  }
  int ok = 0;
  for (int d = 0; d < n; d++) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
  }
  return ok;
}

}  // extern "C"
