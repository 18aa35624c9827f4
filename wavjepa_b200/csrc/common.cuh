// Shared host-side helpers of libwavjepa_b200.so (error reporting, launch checks).
#pragma once
#include "ptx.cuh"
#include "wavjepa_b200.h"

namespace wj {
void set_error(const char* fmt, ...);
int check_launch(const char* what, int n_kernels = 1);
int sm_count();
}  // namespace wj

#define WJ_STREAM(s) reinterpret_cast<cudaStream_t>(s)
