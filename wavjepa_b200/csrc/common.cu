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

static void* g_det_ws = nullptr;
static size_t g_det_bytes = 0;
bool det_on() { return g_det_ws != nullptr; }
void* det_ws(size_t bytes, int* rc) {
  if (g_det_ws == nullptr) return nullptr;
  if (bytes > g_det_bytes) {
    set_error("deterministic mode: workspace of %zu bytes is too small (this call needs %zu)", g_det_bytes, bytes);
    *rc = WJ_ERR_ARG;
    return nullptr;
  }
  return g_det_ws;
}

template <typename T>
__global__ void __launch_bounds__(256) det_reduce_kernel(const T* __restrict__ ws, int nb, long long rows, int cols,
                                                         T* __restrict__ dst, long long ld) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    T acc = 0;
    for (int b = 0; b < nb; ++b) acc += ws[static_cast<long long>(b) * n + i];
    const long long r = i / cols;
    dst[r * ld + (i - r * cols)] += acc;
  }
}
int det_reduce_f32(const float* ws, int nb, long long rows, int cols, float* dst, long long ld, cudaStream_t st) {
  const long long n = rows * cols;
  if (n <= 0 || nb <= 0) return WJ_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  det_reduce_kernel<float><<<static_cast<int>(blocks), 256, 0, st>>>(ws, nb, rows, cols, dst, ld);
  return check_launch("det_reduce");
}
int det_reduce_f64(const double* ws, int nb, long long n, double* dst, cudaStream_t st) {
  if (n <= 0 || nb <= 0) return WJ_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  det_reduce_kernel<double><<<static_cast<int>(blocks), 256, 0, st>>>(ws, nb, n, 1, dst, 1);
  return check_launch("det_reduce");
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

extern "C" int wj_set_deterministic(void* workspace, size_t bytes) {
  wj::g_det_ws = bytes > 0 ? workspace : nullptr;
  wj::g_det_bytes = workspace != nullptr ? bytes : 0;
  return WJ_OK;
}
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
