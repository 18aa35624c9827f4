// First block of the waveform encoder: Conv1d(Cin -> C, k=10, stride=5, no bias) + GroupNorm(C groups, C channels)
// + exact GELU, fused, writing channels-last bf16 so that the next layer's implicit GEMM can TMA it directly.
// Reference: wavjepa/extractors/audio_feature_extractor.py:70 (conv), :94 (GroupNorm, eps 1e-5), :95 (GELU); block 0
// of `self.cnn` (:107-121).  The conv is 10 MAC / output and the output is 100x larger than the input (6.58 MB of bf16
// per 2 s instance), so nothing the size of the output is ever read back or stored twice:
//
//   moments   (tiny)  per instance: s[a] = sum_t x_a(t), R[a][a'] = sum_t x_a(t) x_a'(t) over the conv windows
//                     (a = ci*10 + tap, x_a(t) = x[ci, 5t + tap]); fp32 partial sums folded into fp64.
//   stats     (tiny)  GroupNorm statistics of every (instance, channel) in closed form: mean = w.s / L,
//                     E[h^2] = w^T R w / L (fp64).  The reference takes them over the bf16-ROUNDED conv output; the
//                     closed form uses the unrounded one -- a relative difference of ~1e-6 (rounding noise is zero-mean
//                     with variance 2^-18 h^2 / 3), far below the bf16 resolution of the result.
//   forward   (one pass) conv (bf16-rounded, as autocast does) -> normalise -> affine -> GELU -> [B, L_out, C] bf16
//   backward  (one pass over dY) per (instance, channel): S1 = sum dz, S2 = sum dz hhat, P[a] = sum dz x_a with
//                     dz = dY gelu'(z); then (tiny) dW[c,a] += rstd gamma (P[a] - S1/L s[a] - S2/L Q[a]),
//                     Q[a] = sum_t hhat x_a = rstd ((R w)[a] - mean s[a]); dgamma += S2, dbeta += S1.
// Thread <-> 2 adjacent channels, weights in registers, the input window of 4 consecutive outputs in registers
// (5 x LDS.128 per 4 outputs), stores / dY loads are 128 contiguous bytes per warp.
#include "common.cuh"

namespace wj {

constexpr int kC0K = 10;      // kernel width
constexpr int kC0S = 5;       // stride
constexpr int kC0TT = 128;    // outputs per smem window
constexpr int kC0MaxCin = 2;
constexpr int kC0Win = (kC0TT - 1) * kC0S + kC0K;      // 645 samples
constexpr int kC0WinPad = (kC0Win + 3 + 3) & ~3;       // 648: float4 reads may run 3 past the last tap

struct Conv0Args {
  const bf16* x;      // [B, Cin, L]
  const float* w;     // [C, Cin, 10] fp32 master weights (rounded to bf16 on load: autocast semantics)
  int B, Cin, L, L_out, C;
};

__host__ __device__ inline int conv0_moment_count(int Cin) { const int na = Cin * kC0K; return na + na * na; }

// ------------------------------------------------------------------------------------------------ moments
// grid B, block 512: thread o < NA + NA*NA owns one moment; the instance streams through smem in windows.
template <int CIN>
__global__ void __launch_bounds__(512) conv0_moments_kernel(Conv0Args a, double* __restrict__ mom) {
  constexpr int NA = CIN * kC0K;
  constexpr int NOUT = NA + NA * NA;
  constexpr int TCH = 512;                                  // outputs per window
  constexpr int WIN = (TCH - 1) * kC0S + kC0K;
  __shared__ float s_x[CIN * WIN];
  const int b = blockIdx.x;
  const int o = threadIdx.x;
  int a0 = 0, a1 = -1;
  if (o < NA) a0 = o;
  else if (o < NOUT) { a0 = (o - NA) / NA; a1 = (o - NA) % NA; }
  const int off0 = (a0 / kC0K) * WIN + (a0 % kC0K);
  const int off1 = a1 >= 0 ? (a1 / kC0K) * WIN + (a1 % kC0K) : 0;
  double total = 0.0;
  for (int t0 = 0; t0 < a.L_out; t0 += TCH) {
    const int nt = min(TCH, a.L_out - t0);
    const int win = (nt - 1) * kC0S + kC0K;
    __syncthreads();
    for (int ci = 0; ci < CIN; ++ci) {
      const bf16* src = a.x + (static_cast<size_t>(b) * CIN + ci) * a.L + static_cast<size_t>(t0) * kC0S;
      for (int i = threadIdx.x; i < win; i += blockDim.x) s_x[ci * WIN + i] = __bfloat162float(src[i]);
    }
    __syncthreads();
    if (o < NOUT) {
      float acc = 0.f;
      if (a1 < 0) {
        for (int t = 0; t < nt; ++t) acc += s_x[off0 + t * kC0S];
      } else {
        for (int t = 0; t < nt; ++t) acc = fmaf(s_x[off0 + t * kC0S], s_x[off1 + t * kC0S], acc);
      }
      total += static_cast<double>(acc);
    }
  }
  if (o < NOUT) mom[static_cast<size_t>(b) * NOUT + o] = total;
}

// grid B, block C: closed-form GroupNorm statistics -> stats[b, c] = (mean, rstd)
template <int CIN>
__global__ void conv0_stats_kernel(Conv0Args a, const double* __restrict__ mom, float eps, float* __restrict__ stats) {
  constexpr int NA = CIN * kC0K;
  constexpr int NOUT = NA + NA * NA;
  __shared__ double s_m[NOUT];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < NOUT; i += blockDim.x) s_m[i] = mom[static_cast<size_t>(b) * NOUT + i];
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= a.C) return;
  double w[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) w[i] = static_cast<double>(bf16_round(a.w[static_cast<size_t>(c) * NA + i]));
  double m = 0.0, q = 0.0;
#pragma unroll
  for (int i = 0; i < NA; ++i) m += w[i] * s_m[i];
  for (int i = 0; i < NA; ++i) {
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < NA; ++j) r += w[j] * s_m[NA + i * NA + j];
    q += w[i] * r;
  }
  const double mean = m / a.L_out;
  double var = q / a.L_out - mean * mean;   // biased
  if (var < 0.0) var = 0.0;
  stats[(static_cast<size_t>(b) * a.C + c) * 2 + 0] = static_cast<float>(mean);
  stats[(static_cast<size_t>(b) * a.C + c) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// ------------------------------------------------------------------------------------------------ main passes
// Loads the input window of outputs [t0, t0 + kC0TT) into smem as fp32 (zero past the end of the signal).
template <int CIN>
__device__ __forceinline__ void load_window(const Conv0Args& a, int b, int t0, float* s_x) {
  for (int ci = 0; ci < CIN; ++ci) {
    const bf16* src = a.x + (static_cast<size_t>(b) * CIN + ci) * a.L;
    for (int i = threadIdx.x; i < kC0WinPad; i += blockDim.x) {
      const int p = t0 * kC0S + i;
      s_x[ci * kC0WinPad + i] = (i < kC0Win && p < a.L) ? __bfloat162float(src[p]) : 0.f;
    }
  }
}

template <int CIN>
__device__ __forceinline__ void load_weights(const Conv0Args& a, int c, float (&w)[CIN * kC0K]) {
#pragma unroll
  for (int i = 0; i < CIN * kC0K; ++i) w[i] = bf16_round(a.w[static_cast<size_t>(c) * CIN * kC0K + i]);
}

// The 25 input samples (per input channel) feeding outputs tl .. tl+3 (tl a multiple of 4): 7 aligned float4 loads.
template <int CIN>
__device__ __forceinline__ void load_x4(const float* s_x, int tl, float (&xw)[CIN][28]) {
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci) {
    const float4* p = reinterpret_cast<const float4*>(s_x + ci * kC0WinPad + tl * kC0S);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float4 v = p[i];
      xw[ci][4 * i] = v.x; xw[ci][4 * i + 1] = v.y; xw[ci][4 * i + 2] = v.z; xw[ci][4 * i + 3] = v.w;
    }
  }
}

template <int CIN>
__device__ __forceinline__ float conv_at(const float (&xw)[CIN][28], const float (&w)[CIN * kC0K], int u) {
  float acc = 0.f;
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int j = 0; j < kC0K; ++j) acc = fmaf(xw[ci][u * kC0S + j], w[ci * kC0K + j], acc);
  return bf16_round(acc);  // conv1d output is bf16 under autocast
}

// grid (ceil(chunks / chunks_per_block), B), block C/2 threads: thread owns channels 2*tid, 2*tid+1
template <int CIN>
__global__ void __launch_bounds__(256) conv0_fwd_kernel(Conv0Args a, const float* __restrict__ stats,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ out,
                                                        int chunks_per_block) {
  __shared__ __align__(16) float s_x[CIN * kC0WinPad];
  const int b = blockIdx.y;
  const int c0 = threadIdx.x * 2;
  float w0[CIN * kC0K], w1[CIN * kC0K];
  load_weights<CIN>(a, c0, w0);
  load_weights<CIN>(a, c0 + 1, w1);
  const float4 st = *reinterpret_cast<const float4*>(stats + (static_cast<size_t>(b) * a.C + c0) * 2);  // m0 r0 m1 r1
  const float g0 = gamma[c0] * st.y, g1 = gamma[c0 + 1] * st.w;
  const float b0 = beta[c0] - st.x * g0, b1 = beta[c0 + 1] - st.z * g1;
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const int t0 = (blockIdx.x * chunks_per_block + ch) * kC0TT;
    if (t0 >= a.L_out) break;
    __syncthreads();
    load_window<CIN>(a, b, t0, s_x);
    __syncthreads();
    const int nt = min(kC0TT, a.L_out - t0);
    bf16* o = out + (static_cast<size_t>(b) * a.L_out + t0) * a.C + c0;
    for (int tl = 0; tl < nt; tl += 4) {
      float xw[CIN][28];
      load_x4<CIN>(s_x, tl, xw);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (tl + u < nt) {
          const float h0 = conv_at<CIN>(xw, w0, u), h1 = conv_at<CIN>(xw, w1, u);
          const float y0 = gelu_fast(fmaf(h0, g0, b0)), y1 = gelu_fast(fmaf(h1, g1, b1));
          *reinterpret_cast<uint32_t*>(o + static_cast<size_t>(tl + u) * a.C) = pack_bf16x2(y0, y1);
        }
      }
    }
  }
}

// Backward, the one pass over dY: red[b, 0, c] += S1, red[b, 1, c] += S2, red[b, 2 + a, c] += P[a]
template <int CIN>
__global__ void __launch_bounds__(256) conv0_bwd_kernel(Conv0Args a, const float* __restrict__ stats,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const bf16* __restrict__ dy,
                                                        float* __restrict__ red, int chunks_per_block) {
  constexpr int NA = CIN * kC0K;
  __shared__ __align__(16) float s_x[CIN * kC0WinPad];
  const int b = blockIdx.y;
  const int c0 = threadIdx.x * 2;
  float w0[NA], w1[NA];
  load_weights<CIN>(a, c0, w0);
  load_weights<CIN>(a, c0 + 1, w1);
  const float4 st = *reinterpret_cast<const float4*>(stats + (static_cast<size_t>(b) * a.C + c0) * 2);  // m0 r0 m1 r1
  const float ga0 = gamma[c0], ga1 = gamma[c0 + 1], be0 = beta[c0], be1 = beta[c0 + 1];
  float s10 = 0.f, s20 = 0.f, s11 = 0.f, s21 = 0.f;
  float p0[NA], p1[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) { p0[i] = 0.f; p1[i] = 0.f; }
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const int t0 = (blockIdx.x * chunks_per_block + ch) * kC0TT;
    if (t0 >= a.L_out) break;
    __syncthreads();
    load_window<CIN>(a, b, t0, s_x);
    __syncthreads();
    const int nt = min(kC0TT, a.L_out - t0);
    const bf16* d = dy + (static_cast<size_t>(b) * a.L_out + t0) * a.C + c0;
    for (int tl = 0; tl < nt; tl += 4) {
      float xw[CIN][28];
      load_x4<CIN>(s_x, tl, xw);
      uint32_t dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        dv[u] = (tl + u < nt) ? *reinterpret_cast<const uint32_t*>(d + static_cast<size_t>(tl + u) * a.C) : 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float hh0 = (conv_at<CIN>(xw, w0, u) - st.x) * st.y, hh1 = (conv_at<CIN>(xw, w1, u) - st.z) * st.w;
        const float2 dd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv[u]));   // 0 past the end
        const float dz0 = dd.x * gelu_fast_grad(fmaf(hh0, ga0, be0)), dz1 = dd.y * gelu_fast_grad(fmaf(hh1, ga1, be1));
        s10 += dz0; s20 = fmaf(dz0, hh0, s20); s11 += dz1; s21 = fmaf(dz1, hh1, s21);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int j = 0; j < kC0K; ++j) {
            const float xv = xw[ci][u * kC0S + j];
            p0[ci * kC0K + j] = fmaf(dz0, xv, p0[ci * kC0K + j]);
            p1[ci * kC0K + j] = fmaf(dz1, xv, p1[ci * kC0K + j]);
          }
      }
    }
  }
  float* r = red + static_cast<size_t>(b) * (2 + NA) * a.C + c0;
  atomicAdd(r, s10); atomicAdd(r + 1, s11);
  atomicAdd(r + a.C, s20); atomicAdd(r + a.C + 1, s21);
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    atomicAdd(r + static_cast<size_t>(2 + i) * a.C, p0[i]);
    atomicAdd(r + static_cast<size_t>(2 + i) * a.C + 1, p1[i]);
  }
}

// Backward finalize: grid C/32, block (32, 8).  Thread (cx, by) folds instances by, by+8, ... of channel c.
template <int CIN>
__global__ void __launch_bounds__(256) conv0_bwd_finalize_kernel(Conv0Args a, const double* __restrict__ mom,
                                                                 const float* __restrict__ stats,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ red, float* __restrict__ dw,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int NA = CIN * kC0K;
  constexpr int NOUT = NA + NA * NA;
  __shared__ float s_m[8][NOUT];
  __shared__ float s_acc[8][32][NA + 2];
  const int cx = threadIdx.x, by = threadIdx.y;
  const int c = blockIdx.x * 32 + cx;
  const bool c_ok = c < a.C;
  float w[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) w[i] = c_ok ? bf16_round(a.w[static_cast<size_t>(c) * NA + i]) : 0.f;
  const float ga = c_ok ? gamma[c] : 0.f;
  const float invL = 1.0f / a.L_out;
  float acc[NA + 2];
#pragma unroll
  for (int i = 0; i < NA + 2; ++i) acc[i] = 0.f;
  for (int b0 = 0; b0 < a.B; b0 += 8) {
    const int b = b0 + by;
    __syncthreads();
    if (b < a.B)
      for (int i = cx; i < NOUT; i += 32) s_m[by][i] = static_cast<float>(mom[static_cast<size_t>(b) * NOUT + i]);
    __syncthreads();
    if (b < a.B && c_ok) {
      const float mean = stats[(static_cast<size_t>(b) * a.C + c) * 2], rstd = stats[(static_cast<size_t>(b) * a.C + c) * 2 + 1];
      const float* r = red + static_cast<size_t>(b) * (2 + NA) * a.C + c;
      const float S1 = r[0], S2 = r[a.C];
      const float k1 = S1 * invL, k2 = S2 * invL * rstd, sc = rstd * ga;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        float rw = 0.f;   // (R w)[i]
#pragma unroll
        for (int j = 0; j < NA; ++j) rw = fmaf(s_m[by][NA + i * NA + j], w[j], rw);
        const float q = rw - mean * s_m[by][i];            // Q[i] / rstd
        acc[i] += sc * (r[static_cast<size_t>(2 + i) * a.C] - k1 * s_m[by][i] - k2 * q);
      }
      acc[NA] += S2;
      acc[NA + 1] += S1;
    }
  }
#pragma unroll
  for (int i = 0; i < NA + 2; ++i) s_acc[by][cx][i] = acc[i];
  __syncthreads();
  if (by == 0 && c_ok) {
#pragma unroll
    for (int i = 0; i < NA + 2; ++i) {
      float v = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) v += s_acc[y][cx][i];
      if (i < NA) dw[static_cast<size_t>(c) * NA + i] += v;
      else if (i == NA) dgamma[c] += v;
      else dbeta[c] += v;
    }
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

extern "C" int wj_conv0_moment_count(int Cin) { return conv0_moment_count(Cin); }

extern "C" int wj_conv0_gn_gelu_fwd(const void* x_bf16, const float* w, const float* gamma, const float* beta, int B,
                                    int Cin, int L, int C, int k, int stride, float eps, double* moments, float* stats,
                                    void* out_bf16, void* stream) {
  if (B <= 0) return WJ_OK;
  int rc = check_conv0(Cin, C, k, stride);
  if (rc) return rc;
  cudaStream_t st = WJ_STREAM(stream);
  Conv0Args a;
  a.x = reinterpret_cast<const bf16*>(x_bf16); a.w = w; a.B = B; a.Cin = Cin; a.L = L; a.C = C;
  a.L_out = (L - k) / stride + 1;
  const int chunks = (a.L_out + kC0TT - 1) / kC0TT;
  const int cpb = 4;
  dim3 grid((chunks + cpb - 1) / cpb, B);
  bf16* out = reinterpret_cast<bf16*>(out_bf16);
  if (Cin == 1) {
    conv0_moments_kernel<1><<<B, 512, 0, st>>>(a, moments);
    conv0_stats_kernel<1><<<B, C, 0, st>>>(a, moments, eps, stats);
    conv0_fwd_kernel<1><<<grid, C / 2, 0, st>>>(a, stats, gamma, beta, out, cpb);
  } else {
    conv0_moments_kernel<2><<<B, 512, 0, st>>>(a, moments);
    conv0_stats_kernel<2><<<B, C, 0, st>>>(a, moments, eps, stats);
    conv0_fwd_kernel<2><<<grid, C / 2, 0, st>>>(a, stats, gamma, beta, out, cpb);
  }
  return check_launch("conv0_gn_gelu_fwd", 3);
}

extern "C" int wj_conv0_gn_gelu_bwd(const void* x_bf16, const float* w, const float* gamma, const float* beta, int B,
                                    int Cin, int L, int C, int k, int stride, float eps, const double* moments,
                                    const float* stats, const void* dy_bf16, float* red_scratch, float* dw,
                                    float* dgamma, float* dbeta, void* stream) {
  if (B <= 0) return WJ_OK;
  int rc = check_conv0(Cin, C, k, stride);
  if (rc) return rc;
  (void)eps;
  cudaStream_t st = WJ_STREAM(stream);
  Conv0Args a;
  a.x = reinterpret_cast<const bf16*>(x_bf16); a.w = w; a.B = B; a.Cin = Cin; a.L = L; a.C = C;
  a.L_out = (L - k) / stride + 1;
  const int na = Cin * kC0K;
  cudaMemsetAsync(red_scratch, 0, static_cast<size_t>(B) * (2 + na) * C * sizeof(float), st);
  const int chunks = (a.L_out + kC0TT - 1) / kC0TT;
  const int cpb = 13;   // 2*(2+na) atomics per thread at the end: few, long blocks (4 per 2 s instance)
  dim3 grid((chunks + cpb - 1) / cpb, B);
  const bf16* dy = reinterpret_cast<const bf16*>(dy_bf16);
  dim3 fgrid((C + 31) / 32), fblock(32, 8);
  if (Cin == 1) {
    conv0_bwd_kernel<1><<<grid, C / 2, 0, st>>>(a, stats, gamma, beta, dy, red_scratch, cpb);
    conv0_bwd_finalize_kernel<1><<<fgrid, fblock, 0, st>>>(a, moments, stats, gamma, red_scratch, dw, dgamma, dbeta);
  } else {
    conv0_bwd_kernel<2><<<grid, C / 2, 0, st>>>(a, stats, gamma, beta, dy, red_scratch, cpb);
    conv0_bwd_finalize_kernel<2><<<fgrid, fblock, 0, st>>>(a, moments, stats, gamma, red_scratch, dw, dgamma, dbeta);
  }
  return check_launch("conv0_gn_gelu_bwd", 2);
}
