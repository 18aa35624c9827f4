#include "common.cuh"

#include <stdarg.h>
#include <stdio.h>

#include <atomic>

namespace wj {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what, int n_kernels) {
  g_launches.fetch_add(n_kernels, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return WJ_ERR_RUNTIME;
  }
  return WJ_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}
}  // namespace wj

extern "C" const char* wj_last_error(void) { return wj::g_err; }
extern "C" int wj_version(void) { return 1; }
extern "C" long long wj_kernel_launches(void) { return wj::g_launches.load(std::memory_order_relaxed); }
extern "C" int wj_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    wj::set_error("no CUDA device available: libwavjepa_b200 has no CPU fallback");
    return WJ_ERR_ARCH;
  }
  if (prop.major != 10) {
    wj::set_error("device %s is sm_%d%d; libwavjepa_b200 is built for sm_100a only", prop.name, prop.major, prop.minor);
    return WJ_ERR_ARCH;
  }
  return WJ_OK;
}
