/* libwavjepa_b200.so -- C ABI of the B200-native WavJEPA hot path.
 *
 * The reference (labhamlet/wavjepa) is pure Python/PyTorch and has NO FFI layer: its boundary for this path is the
 * Python module API (wavjepa/jepa.py JEPA, wavjepa/extractors/*, wavjepa/masking.py, hear_api/runtime.py).  This
 * header is therefore the boundary a maintainer would bind from that Python (ctypes stub shown in INTEGRATION.md);
 * every entry point cites the reference op (file:line under the reference checkout) it replaces.
 *
 * Conventions: raw device pointers + explicit sizes; the caller owns every buffer; every function is asynchronous on
 * the given cudaStream_t (passed as void*), never synchronises the device, keeps no global mutable state besides
 * immutable function attributes; returns 0 on success, <0 on error with wj_last_error() giving a thread-local message.
 * There is no CPU fallback: on a machine without an sm_100 device every compute entry point fails.
 */
#ifndef WAVJEPA_B200_H
#define WAVJEPA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WJ_OK 0
#define WJ_ERR_ARG (-1)
#define WJ_ERR_RUNTIME (-2)
#define WJ_ERR_ARCH (-3)

const char* wj_last_error(void);
int wj_version(void);
/* 0 if the current device is sm_100 (B200); WJ_ERR_ARCH otherwise. */
int wj_check_device(void);

/* ------------------------------------------------------------------------------------------------------------------
 * GEMM operand description (bf16).  A 4-D strided view, innermost dimension first:
 *   dim[0] = contiguous "virtual column" extent of one segment, dim[1] = parity/sub-row selector, dim[2] = rows per
 *   batch entry, dim[3] = batch.  stride_bytes[i] is the byte stride of dim[i+1].
 * Virtual columns vc are split in segments of seg_width (0 = one segment); segment s addresses
 *   (vc % seg_width, seg_q[s], row + seg_p[s], batch) -- this is what turns the strided Conv1d windows of
 *   wavjepa/extractors/audio_feature_extractor.py:70 into plain TMA tiles over channels-last activations.
 * Out-of-range coordinates (including negative rows) read as zero.
 */
typedef struct wj_operand {
  const void* ptr;
  int64_t dim[4];
  int64_t stride_bytes[3];
  int32_t seg_width;
  int32_t seg_q[4];
  int32_t seg_p[4];
} wj_operand_t;

/* Fused epilogue of wj_gemm_bf16, applied in this order:
 *   v = acc + bias[n]
 *   act == 1 : h = bf16(v); out2[row, n] = h (if out2); v = GELU_erf(h)          (nn.GELU, types/wavjepa_configs.py:37)
 *   act == 2 : v = v * GELU_erf'(aux[row, n])                                    (backward of the above)
 *   v += resid[row % resid_mod or row, n]                                        (residual / positional table)
 *   out[row, n] = v  (bf16, or fp32 when out_f32; fp32 reduce-add when accumulate)
 * out_rows (optional, int32 per logical row): physical row used for out/out2/resid/aux, <0 skips the row.
 */
typedef struct wj_epilogue {
  void* out;
  int64_t ld_out;
  int32_t out_f32;
  int32_t accumulate;
  void* out2;
  int64_t ld_out2;
  const float* bias;
  const void* resid;
  int32_t resid_f32;
  int32_t resid_mod;
  int64_t ld_resid;
  const void* aux;
  int64_t ld_aux;
  int32_t act;
  int32_t _pad;
  const int32_t* out_rows;
} wj_epilogue_t;

/* out[b*L + t, n] = epilogue( sum_vc A(vc; t, b) * W[n, vc] ),  W bf16 [N, K] row-major with leading dim ldw.
 * tcgen05.mma (cta_group::1, 128 x block_n x 16) fed by TMA, fp32 accumulation in TMEM.
 * Replaces: F.linear under bf16 autocast (nn.TransformerEncoderLayer / nn.MultiheadAttention projections built at
 * wavjepa/jepa.py:126-132) and nn.Conv1d(512,512,k,stride 2) (audio_feature_extractor.py:70) as implicit GEMM. */
int wj_gemm_bf16(const wj_operand_t* A, const void* W, int64_t ldw, int L, int batch, int N, int K,
                 const wj_epilogue_t* epi, int block_n, void* stream);

/* dW[m, vc] (+)= sum_{b,t} dY[b, t, m] * X(vc; t, b)   fp32 out [M, N] (leading dim ld_out); both operands are read
 * MN-major straight from the row-major activations; the token reduction is split over `splits` CTAs (0 = auto) and
 * reduced with red.global.add.  Replaces autograd's weight gradients of the Linear / Conv1d layers above. */
int wj_gemm_wgrad_bf16(const wj_operand_t* dY, const wj_operand_t* X, int L, int batch, int M, int N, float* out,
                       int64_t ld_out, int accumulate, int splits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVJEPA_B200_H */
