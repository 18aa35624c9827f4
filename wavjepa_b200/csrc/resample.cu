// On-GPU input pipeline of the training data module: band-limited sinc resampling with a Kaiser
// window to 16 kHz, loudness normalisation to -14 dBFS over the whole resampled clip, pad / crop to 10 s.
// Reference: data_modules/WebAudioDataModule.py:43-61 (torchaudio.transforms.Resample(lowpass_filter_width=64,
// rolloff=0.9475937167399596, resampling_method="sinc_interp_kaiser", beta=14.769656459379492)) and
// data_modules/dataset_functions.py:92-114 (normalize_audio, pre_process).  The polyphase filter table is the
// third-party algorithm of torchaudio 2.x (_get_sinc_resample_kernel / _apply_sinc_resample_kernel), restated on the
// host in wavjepa_b200/preprocess.py; this kernel is its conv1d(stride = orig) + interleave + crop:
//     y[n * new + p] = sum_k table[p][k] * xpad[n * orig + k],   xpad = x padded by (width, width + orig) zeros
// One thread per output sample, the block's input window staged in shared memory, the table stored phase-contiguous
// ([K, new]) so that a warp's taps are one coalesced line; the clip's sum of squares (for the RMS gain) is reduced per
// block and accumulated with one fp64 atomic.
#include "common.cuh"

namespace wj {

__global__ void __launch_bounds__(256) resample_sinc_kernel(const float* __restrict__ x, long long length,
                                                            const float* __restrict__ table_t, int orig, int new_,
                                                            int width, int K, long long target_length,
                                                            float* __restrict__ out, long long out_cap,
                                                            double* __restrict__ sumsq, int win_cap) {
  extern __shared__ float s_win[];
  __shared__ float s_red[8];
  const long long t0 = static_cast<long long>(blockIdx.x) * blockDim.x;
  const long long t_last = min(t0 + blockDim.x, target_length) - 1;
  float v = 0.f;
  if (t0 < target_length) {
    const long long n0 = t0 / new_, n1 = t_last / new_;
    const long long base = n0 * orig;                       // first padded index of the block's window
    const int win = static_cast<int>((n1 - n0) * orig + K);
    const bool staged = win <= win_cap;
    if (staged) {
      for (int i = threadIdx.x; i < win; i += blockDim.x) {
        const long long src = base + i - width;
        s_win[i] = (src >= 0 && src < length) ? x[src] : 0.f;
      }
    }
    __syncthreads();
    const long long t = t0 + threadIdx.x;
    if (t < target_length) {
      const long long n = t / new_;
      const int p = static_cast<int>(t - n * new_);
      const float* tp = table_t + p;
      float acc = 0.f;
      if (staged) {
        const float* w = s_win + (n - n0) * orig;
#pragma unroll 4
        for (int k = 0; k < K; ++k) acc = fmaf(__ldg(tp + static_cast<long long>(k) * new_), w[k], acc);
      } else {
        for (int k = 0; k < K; ++k) {
          const long long src = n * orig + k - width;
          const float xv = (src >= 0 && src < length) ? x[src] : 0.f;
          acc = fmaf(__ldg(tp + static_cast<long long>(k) * new_), xv, acc);
        }
      }
      if (t < out_cap) out[t] = acc;
      v = acc * acc;
    }
  } else {
    __syncthreads();
  }
  // zero padding up to the fixed clip length (pre_process pads with zeros)
  for (long long t = max(t0, target_length) + threadIdx.x; t < min(t0 + blockDim.x, out_cap); t += blockDim.x) out[t] = 0.f;
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 8) {
    float r = s_red[threadIdx.x];
    r += __shfl_xor_sync(0xffu, r, 4);
    r += __shfl_xor_sync(0xffu, r, 2);
    r += __shfl_xor_sync(0xffu, r, 1);
    if (threadIdx.x == 0 && sumsq != nullptr && r != 0.f) atomicAdd(sumsq, static_cast<double>(r));
  }
}

// clips[c, :] *= 10^((target_dbfs - 20 log10(rms_c)) / 20), rms_c = sqrt(sumsq[c] / count[c]); untouched when rms == 0
__global__ void __launch_bounds__(256) rms_gain_rows_kernel(float* __restrict__ clips, const double* __restrict__ sumsq,
                                                            const long long* __restrict__ counts, long long row_len,
                                                            float target_dbfs) {
  const int c = blockIdx.y;
  const double ss = sumsq[c];
  if (ss <= 0.0) return;
  const float rms = static_cast<float>(sqrt(ss / static_cast<double>(counts[c])));
  const float gain = powf(10.0f, (target_dbfs - 20.0f * log10f(rms)) / 20.0f);
  float* row = clips + static_cast<long long>(c) * row_len;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < row_len;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    row[i] *= gain;
}

}  // namespace wj

using namespace wj;

extern "C" int wj_resample_sinc(const float* x, int64_t length, const float* table_t, int orig, int new_, int width,
                                int64_t target_length, float* out, int64_t out_cap, double* sumsq, void* stream) {
  if (orig <= 0 || new_ <= 0 || width < 0 || length < 0 || out_cap <= 0) { set_error("wj_resample_sinc: bad arguments"); return WJ_ERR_ARG; }
  const int K = 2 * width + orig;
  const long long n_thr = target_length > out_cap ? target_length : out_cap;
  const int blocks = static_cast<int>((n_thr + 255) / 256);
  // window of a 256-output block: ceil(255 / new) + 1 frames
  const long long win = (255 / new_ + 1) * static_cast<long long>(orig) + K;
  const int win_cap = win * 4 <= 48 * 1024 ? static_cast<int>(win) : 0;
  resample_sinc_kernel<<<blocks, 256, static_cast<size_t>(win_cap) * 4, WJ_STREAM(stream)>>>(
      x, length, table_t, orig, new_, width, K, target_length, out, out_cap, sumsq, win_cap);
  return check_launch("resample_sinc");
}

extern "C" int wj_rms_gain_rows(float* clips, const double* sumsq, const int64_t* counts, int n_clips, int64_t row_len,
                                float target_dbfs, void* stream) {
  if (n_clips <= 0) return WJ_OK;
  dim3 grid(static_cast<unsigned>((row_len + 256 * 8 - 1) / (256 * 8)), n_clips);
  rms_gain_rows_kernel<<<grid, 256, 0, WJ_STREAM(stream)>>>(clips, sumsq, reinterpret_cast<const long long*>(counts), row_len,
                                                            target_dbfs);
  return check_launch("rms_gain_rows");
}
