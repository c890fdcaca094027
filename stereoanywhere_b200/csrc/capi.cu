// Library-level pieces of the C ABI: version, thread-local error text, device properties.
#include <stdarg.h>
#include <stdlib.h>

#include "sa_common.cuh"

namespace sa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Experiment hook: SA_B200_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity once per process.
static void apply_l2_fetch_limit() {
  static bool done = false;
  if (done) return;
  done = true;
  const char* e = getenv("SA_B200_L2_FETCH");
  if (e) {
    size_t before = 0, after = 0;
    cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
    cudaError_t rc = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
    cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
    fprintf(stderr, "[sa_b200] L2 fetch granularity %zu -> %zu (rc=%d)\n", before, after, (int)rc);
  }
}

int num_sms() {
  apply_l2_fetch_limit();
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace sa

extern "C" int sa_abi_version(void) { return SA_ABI_VERSION; }
extern "C" const char* sa_last_error(void) { return sa::g_err; }
