// First block of the waveform encoder: Conv1d(Cin -> C, k=10, stride=5, no bias) + GroupNorm(C groups, C channels)
// + exact GELU, fused, writing channels-last bf16 so that the next layer's implicit GEMM can TMA it directly.
// Reference: wavjepa/extractors/audio_feature_extractor.py:70 (conv), :94 (GroupNorm, eps 1e-5), :95 (GELU); block 0
// of `self.cnn` (:107-121).  The conv is 10 MAC / output and the output is 100x larger than the input, so the op is
// bound by the bf16 write (6.58 MB / instance): the conv is simply recomputed in the second pass instead of stored.
//
//   pass A (stats)   : per (instance, channel) sum / sum of squares of the bf16-rounded conv output (fp64 atomics)
//   pass B (forward) : recompute conv, normalise, affine, GELU, store [B, L_out, C] bf16
//   backward         : two passes over dY (first the two GroupNorm reductions, then dW / dgamma / dbeta)
#include "common.cuh"

namespace wj {

constexpr int kC0K = 10;      // kernel width
constexpr int kC0S = 5;       // stride
constexpr int kC0TT = 128;    // outputs per smem window
constexpr int kC0MaxCin = 2;

struct Conv0Args {
  const bf16* x;      // [B, Cin, L]
  const float* w;     // [C, Cin, 10] fp32 master weights (rounded to bf16 on load: autocast semantics)
  int B, Cin, L, L_out, C;
  double* stats;      // [B, C, 2] sum, sumsq of bf16(conv)
};

// Loads the input window of outputs [t0, t0 + kC0TT) into smem as fp32.
__device__ __forceinline__ void load_window(const Conv0Args& a, int b, int t0, float* s_x) {
  const int win = (kC0TT - 1) * kC0S + kC0K;
  for (int ci = 0; ci < a.Cin; ++ci) {
    const bf16* src = a.x + (static_cast<size_t>(b) * a.Cin + ci) * a.L;
    for (int i = threadIdx.x; i < win; i += blockDim.x) {
      const int p = t0 * kC0S + i;
      s_x[ci * win + i] = p < a.L ? __bfloat162float(src[p]) : 0.f;
    }
  }
}

template <int CIN>
__device__ __forceinline__ void load_weights(const Conv0Args& a, int c, float (&w)[CIN * kC0K]) {
#pragma unroll
  for (int i = 0; i < CIN * kC0K; ++i) w[i] = bf16_round(a.w[static_cast<size_t>(c) * CIN * kC0K + i]);
}

template <int CIN>
__device__ __forceinline__ float conv_at(const float* s_x, const float (&w)[CIN * kC0K], int tl) {
  constexpr int win = (kC0TT - 1) * kC0S + kC0K;
  float acc = 0.f;
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int j = 0; j < kC0K; ++j) acc = fmaf(s_x[ci * win + tl * kC0S + j], w[ci * kC0K + j], acc);
  return bf16_round(acc);  // conv1d output is bf16 under autocast
}

// grid (chunks, B), block C/2 threads: thread owns channels 2*tid, 2*tid+1
template <int CIN>
__global__ void __launch_bounds__(256) conv0_stats_kernel(Conv0Args a, int chunks_per_block) {
  constexpr int win = (kC0TT - 1) * kC0S + kC0K;
  __shared__ float s_x[CIN * win];
  const int b = blockIdx.y;
  const int c0 = threadIdx.x * 2;
  float w0[CIN * kC0K], w1[CIN * kC0K];
  load_weights<CIN>(a, c0, w0);
  load_weights<CIN>(a, c0 + 1, w1);
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const int t0 = (blockIdx.x * chunks_per_block + ch) * kC0TT;
    if (t0 >= a.L_out) break;
    __syncthreads();
    load_window(a, b, t0, s_x);
    __syncthreads();
    const int nt = min(kC0TT, a.L_out - t0);
    for (int tl = 0; tl < nt; ++tl) {
      const float h0 = conv_at<CIN>(s_x, w0, tl), h1 = conv_at<CIN>(s_x, w1, tl);
      s0 += h0; q0 += h0 * h0; s1 += h1; q1 += h1 * h1;
    }
  }
  double* st = a.stats + (static_cast<size_t>(b) * a.C + c0) * 2;
  atomicAdd(st + 0, static_cast<double>(s0));
  atomicAdd(st + 1, static_cast<double>(q0));
  atomicAdd(st + 2, static_cast<double>(s1));
  atomicAdd(st + 3, static_cast<double>(q1));
}

__device__ __forceinline__ void gn_stats(const double* st, int L_out, float eps, float& mean, float& rstd) {
  const double m = st[0] / L_out;
  double var = st[1] / L_out - m * m;  // biased
  if (var < 0.0) var = 0.0;
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

template <int CIN>
__global__ void __launch_bounds__(256) conv0_fwd_kernel(Conv0Args a, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        bf16* __restrict__ out, int chunks_per_block) {
  constexpr int win = (kC0TT - 1) * kC0S + kC0K;
  __shared__ float s_x[CIN * win];
  const int b = blockIdx.y;
  const int c0 = threadIdx.x * 2;
  float w0[CIN * kC0K], w1[CIN * kC0K];
  load_weights<CIN>(a, c0, w0);
  load_weights<CIN>(a, c0 + 1, w1);
  float m0, r0, m1, r1;
  const double* st = a.stats + (static_cast<size_t>(b) * a.C + c0) * 2;
  gn_stats(st, a.L_out, eps, m0, r0);
  gn_stats(st + 2, a.L_out, eps, m1, r1);
  const float g0 = gamma[c0] * r0, g1 = gamma[c0 + 1] * r1;
  const float b0 = beta[c0] - m0 * g0, b1 = beta[c0 + 1] - m1 * g1;
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const int t0 = (blockIdx.x * chunks_per_block + ch) * kC0TT;
    if (t0 >= a.L_out) break;
    __syncthreads();
    load_window(a, b, t0, s_x);
    __syncthreads();
    const int nt = min(kC0TT, a.L_out - t0);
    bf16* o = out + (static_cast<size_t>(b) * a.L_out + t0) * a.C + c0;
    for (int tl = 0; tl < nt; ++tl) {
      const float h0 = conv_at<CIN>(s_x, w0, tl), h1 = conv_at<CIN>(s_x, w1, tl);
      const float y0 = gelu_erf(fmaf(h0, g0, b0)), y1 = gelu_erf(fmaf(h1, g1, b1));
      *reinterpret_cast<uint32_t*>(o + static_cast<size_t>(tl) * a.C) = pack_bf16x2(y0, y1);
    }
  }
}

// Backward pass 1: red[b, c] = { sum_t dz, sum_t dz * hhat }, dz = dY * gelu'(z)
template <int CIN>
__global__ void __launch_bounds__(256) conv0_bwd_red_kernel(Conv0Args a, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            const bf16* __restrict__ dy, double* __restrict__ red,
                                                            int chunks_per_block) {
  constexpr int win = (kC0TT - 1) * kC0S + kC0K;
  __shared__ float s_x[CIN * win];
  const int b = blockIdx.y;
  const int c0 = threadIdx.x * 2;
  float w0[CIN * kC0K], w1[CIN * kC0K];
  load_weights<CIN>(a, c0, w0);
  load_weights<CIN>(a, c0 + 1, w1);
  float m0, r0, m1, r1;
  const double* st = a.stats + (static_cast<size_t>(b) * a.C + c0) * 2;
  gn_stats(st, a.L_out, eps, m0, r0);
  gn_stats(st + 2, a.L_out, eps, m1, r1);
  const float ga0 = gamma[c0], ga1 = gamma[c0 + 1], be0 = beta[c0], be1 = beta[c0 + 1];
  float s10 = 0.f, s20 = 0.f, s11 = 0.f, s21 = 0.f;
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const int t0 = (blockIdx.x * chunks_per_block + ch) * kC0TT;
    if (t0 >= a.L_out) break;
    __syncthreads();
    load_window(a, b, t0, s_x);
    __syncthreads();
    const int nt = min(kC0TT, a.L_out - t0);
    const bf16* d = dy + (static_cast<size_t>(b) * a.L_out + t0) * a.C + c0;
    for (int tl = 0; tl < nt; ++tl) {
      const float hh0 = (conv_at<CIN>(s_x, w0, tl) - m0) * r0, hh1 = (conv_at<CIN>(s_x, w1, tl) - m1) * r1;
      const float2 dv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d + static_cast<size_t>(tl) * a.C));
      const float dz0 = dv.x * gelu_erf_grad(fmaf(hh0, ga0, be0)), dz1 = dv.y * gelu_erf_grad(fmaf(hh1, ga1, be1));
      s10 += dz0; s20 += dz0 * hh0; s11 += dz1; s21 += dz1 * hh1;
    }
  }
  double* r = red + (static_cast<size_t>(b) * a.C + c0) * 2;
  atomicAdd(r + 0, static_cast<double>(s10));
  atomicAdd(r + 1, static_cast<double>(s20));
  atomicAdd(r + 2, static_cast<double>(s11));
  atomicAdd(r + 3, static_cast<double>(s21));
}

// Backward pass 2: dh = rstd*gamma*(dz - S1/L - hhat*S2/L);  dW[c, ci, j] += sum_t dh * x[ci, 5t+j];
// dgamma[c] += S2, dbeta[c] += S1 (added once per (b, c) by the blockIdx.x == 0 blocks).
template <int CIN>
__global__ void __launch_bounds__(256) conv0_bwd_w_kernel(Conv0Args a, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps,
                                                          const bf16* __restrict__ dy, const double* __restrict__ red,
                                                          float* __restrict__ dw, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, int chunks_per_block) {
  constexpr int win = (kC0TT - 1) * kC0S + kC0K;
  __shared__ float s_x[CIN * win];
  const int b = blockIdx.y;
  const int c0 = threadIdx.x * 2;
  float w0[CIN * kC0K], w1[CIN * kC0K];
  load_weights<CIN>(a, c0, w0);
  load_weights<CIN>(a, c0 + 1, w1);
  float m0, r0, m1, r1;
  const double* st = a.stats + (static_cast<size_t>(b) * a.C + c0) * 2;
  gn_stats(st, a.L_out, eps, m0, r0);
  gn_stats(st + 2, a.L_out, eps, m1, r1);
  const float ga0 = gamma[c0], ga1 = gamma[c0 + 1], be0 = beta[c0], be1 = beta[c0 + 1];
  const double* rd = red + (static_cast<size_t>(b) * a.C + c0) * 2;
  const float invL = 1.0f / a.L_out;
  const float S10 = static_cast<float>(rd[0]) * invL, S20 = static_cast<float>(rd[1]) * invL;
  const float S11 = static_cast<float>(rd[2]) * invL, S21 = static_cast<float>(rd[3]) * invL;
  if (blockIdx.x == 0) {
    atomicAdd(dbeta + c0, static_cast<float>(rd[0]));
    atomicAdd(dgamma + c0, static_cast<float>(rd[1]));
    atomicAdd(dbeta + c0 + 1, static_cast<float>(rd[2]));
    atomicAdd(dgamma + c0 + 1, static_cast<float>(rd[3]));
  }
  float a0[CIN * kC0K], a1[CIN * kC0K];
#pragma unroll
  for (int i = 0; i < CIN * kC0K; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const int t0 = (blockIdx.x * chunks_per_block + ch) * kC0TT;
    if (t0 >= a.L_out) break;
    __syncthreads();
    load_window(a, b, t0, s_x);
    __syncthreads();
    const int nt = min(kC0TT, a.L_out - t0);
    const bf16* d = dy + (static_cast<size_t>(b) * a.L_out + t0) * a.C + c0;
    for (int tl = 0; tl < nt; ++tl) {
      const float hh0 = (conv_at<CIN>(s_x, w0, tl) - m0) * r0, hh1 = (conv_at<CIN>(s_x, w1, tl) - m1) * r1;
      const float2 dv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d + static_cast<size_t>(tl) * a.C));
      const float dz0 = dv.x * gelu_erf_grad(fmaf(hh0, ga0, be0)), dz1 = dv.y * gelu_erf_grad(fmaf(hh1, ga1, be1));
      const float dh0 = r0 * ga0 * (dz0 - S10 - hh0 * S20), dh1 = r1 * ga1 * (dz1 - S11 - hh1 * S21);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int j = 0; j < kC0K; ++j) {
          const float xv = s_x[ci * win + tl * kC0S + j];
          a0[ci * kC0K + j] = fmaf(dh0, xv, a0[ci * kC0K + j]);
          a1[ci * kC0K + j] = fmaf(dh1, xv, a1[ci * kC0K + j]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < CIN * kC0K; ++i) {
    atomicAdd(dw + static_cast<size_t>(c0) * CIN * kC0K + i, a0[i]);
    atomicAdd(dw + static_cast<size_t>(c0 + 1) * CIN * kC0K + i, a1[i]);
  }
}

static int check_conv0(int Cin, int C, int k, int stride) {
  if (k != kC0K || stride != kC0S) { set_error("conv0: only k=10, stride=5 is built (got k=%d s=%d)", k, stride); return WJ_ERR_ARG; }
  if (Cin < 1 || Cin > kC0MaxCin) { set_error("conv0: Cin must be 1 or 2"); return WJ_ERR_ARG; }
  if (C % 2 != 0 || C / 2 > 256 || C / 2 < 32) { set_error("conv0: C must be even, 64..512"); return WJ_ERR_ARG; }
  return WJ_OK;
}

}  // namespace wj

using namespace wj;

extern "C" int wj_conv0_gn_gelu_fwd(const void* x_bf16, const float* w, const float* gamma, const float* beta, int B,
                                    int Cin, int L, int C, int k, int stride, float eps, double* stats,
                                    void* out_bf16, void* stream) {
  if (B <= 0) return WJ_OK;
  int rc = check_conv0(Cin, C, k, stride);
  if (rc) return rc;
  cudaStream_t st = WJ_STREAM(stream);
  Conv0Args a;
  a.x = reinterpret_cast<const bf16*>(x_bf16); a.w = w; a.B = B; a.Cin = Cin; a.L = L; a.C = C;
  a.L_out = (L - k) / stride + 1; a.stats = stats;
  cudaMemsetAsync(stats, 0, static_cast<size_t>(B) * C * 2 * sizeof(double), st);
  const int chunks = (a.L_out + kC0TT - 1) / kC0TT;
  const int cpb = 4;
  dim3 grid((chunks + cpb - 1) / cpb, B);
  if (Cin == 1) {
    conv0_stats_kernel<1><<<grid, C / 2, 0, st>>>(a, cpb);
    conv0_fwd_kernel<1><<<grid, C / 2, 0, st>>>(a, gamma, beta, eps, reinterpret_cast<bf16*>(out_bf16), cpb);
  } else {
    conv0_stats_kernel<2><<<grid, C / 2, 0, st>>>(a, cpb);
    conv0_fwd_kernel<2><<<grid, C / 2, 0, st>>>(a, gamma, beta, eps, reinterpret_cast<bf16*>(out_bf16), cpb);
  }
  return check_launch("conv0_gn_gelu_fwd", 2);
}

extern "C" int wj_conv0_gn_gelu_bwd(const void* x_bf16, const float* w, const float* gamma, const float* beta, int B,
                                    int Cin, int L, int C, int k, int stride, float eps, const double* stats,
                                    const void* dy_bf16, double* red_scratch, float* dw, float* dgamma, float* dbeta,
                                    void* stream) {
  if (B <= 0) return WJ_OK;
  int rc = check_conv0(Cin, C, k, stride);
  if (rc) return rc;
  cudaStream_t st = WJ_STREAM(stream);
  Conv0Args a;
  a.x = reinterpret_cast<const bf16*>(x_bf16); a.w = w; a.B = B; a.Cin = Cin; a.L = L; a.C = C;
  a.L_out = (L - k) / stride + 1; a.stats = const_cast<double*>(stats);
  cudaMemsetAsync(red_scratch, 0, static_cast<size_t>(B) * C * 2 * sizeof(double), st);
  const int chunks = (a.L_out + kC0TT - 1) / kC0TT;
  const int cpb1 = 4;
  dim3 grid1((chunks + cpb1 - 1) / cpb1, B);
  // pass 2 keeps 2*Cin*10 accumulators per thread and ends in atomics: few, long blocks
  const int cpb2 = (chunks + 1) / 2;
  dim3 grid2((chunks + cpb2 - 1) / cpb2, B);
  const bf16* dy = reinterpret_cast<const bf16*>(dy_bf16);
  if (Cin == 1) {
    conv0_bwd_red_kernel<1><<<grid1, C / 2, 0, st>>>(a, gamma, beta, eps, dy, red_scratch, cpb1);
    conv0_bwd_w_kernel<1><<<grid2, C / 2, 0, st>>>(a, gamma, beta, eps, dy, red_scratch, dw, dgamma, dbeta, cpb2);
  } else {
    conv0_bwd_red_kernel<2><<<grid1, C / 2, 0, st>>>(a, gamma, beta, eps, dy, red_scratch, cpb1);
    conv0_bwd_w_kernel<2><<<grid2, C / 2, 0, st>>>(a, gamma, beta, eps, dy, red_scratch, dw, dgamma, dbeta, cpb2);
  }
  return check_launch("conv0_gn_gelu_bwd", 2);
}
