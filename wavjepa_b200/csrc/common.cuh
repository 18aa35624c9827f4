// Shared host-side helpers of libwavjepa_b200.so (error reporting, launch checks).
#pragma once
#include "ptx.cuh"
#include "wavjepa_b200.h"

namespace wj {
void set_error(const char* fmt, ...);
int check_launch(const char* what, int n_kernels = 1);
int sm_count();

// Deterministic mode (wj_set_deterministic): every reduction that otherwise ends in floating-point atomics -- whose order
// varies from run to run -- writes per-block partial results into a caller-provided workspace and a second kernel adds
// them in a fixed order.  det_ws(bytes) returns the workspace when the mode is on and it is large enough, nullptr when
// the mode is off; when it is on but too small it sets the error and returns nullptr with *rc = WJ_ERR_ARG.
bool det_on();
void* det_ws(size_t bytes, int* rc);
// dst[r * ld + c] += sum_{b < nb} ws[(b * rows + r) * cols + c]   (b ascending; T = float or double)
int det_reduce_f32(const float* ws, int nb, long long rows, int cols, float* dst, long long ld, cudaStream_t st);
int det_reduce_f64(const double* ws, int nb, long long n, double* dst, cudaStream_t st);
}  // namespace wj

#define WJ_STREAM(s) reinterpret_cast<cudaStream_t>(s)
