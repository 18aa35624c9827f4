// Integer mask kernels (warp per row), bit-exact against the reference maskers for the same seed.
//
//   wj_masks_generate : TimeInverseBlockMasker.forward (reference wavjepa/masking.py:66-128) and
//                       SpeechMasker.forward (wavjepa/masking.py:167-207), each built on compute_mask_indices
//                       (wavjepa/audio_masking.py:46-194, live path) and numpy's SeedSequence -> PCG64 ->
//                       Generator.random / Generator.choice(replace=False) chain (Floyd sampling + Lemire bounded ints).
//   wj_mask_indices   : packed token index lists + cu_seqlens for the varlen student / predictor / loss kernels
//                       (replaces the boolean-mask gathers of wavjepa/jepa.py:399,425-435).
//
// Seed contract: call `c` of attempt `a` for global row r draws from
// default_rng([base_seed, r, a*8 + c]); c = 0 context mask, 1..G target groups.
#include "common.cuh"

namespace wj {

typedef unsigned __int128 u128;

struct Pcg64 {
  u128 state, inc;
  uint32_t buf32;
  bool has32;
};

__device__ __forceinline__ u128 pcg_mult() {
  return (static_cast<u128>(0x2360ED051FC65DA4ull) << 64) | static_cast<u128>(0x4385DF649FCCF645ull);
}
__device__ __forceinline__ void pcg_step(Pcg64& g) { g.state = g.state * pcg_mult() + g.inc; }

// numpy SeedSequence(entropy = [w0, w1, w2]).generate_state(4, uint64) -> PCG64 seeding (pcg64_set_seed)
__device__ void pcg_seed(Pcg64& g, uint32_t w0, uint32_t w1, uint32_t w2) {
  uint32_t hc = 0x43b0d7e5u;
  auto hashmix = [&](uint32_t v) {
    v ^= hc;
    hc *= 0x931e8875u;
    v *= hc;
    v ^= v >> 16;
    return v;
  };
  auto mix = [](uint32_t x, uint32_t y) {
    uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
    r ^= r >> 16;
    return r;
  };
  uint32_t pool[4];
  pool[0] = hashmix(w0);
  pool[1] = hashmix(w1);
  pool[2] = hashmix(w2);
  pool[3] = hashmix(0u);
#pragma unroll
  for (int src = 0; src < 4; ++src)
#pragma unroll
    for (int dst = 0; dst < 4; ++dst)
      if (src != dst) pool[dst] = mix(pool[dst], hashmix(pool[src]));
  uint32_t hc2 = 0x8b51f9ddu;
  uint32_t o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t d = pool[i & 3] ^ hc2;
    hc2 *= 0x58f38dedu;
    d *= hc2;
    d ^= d >> 16;
    o[i] = d;
  }
  const uint64_t s0 = o[0] | (static_cast<uint64_t>(o[1]) << 32), s1 = o[2] | (static_cast<uint64_t>(o[3]) << 32);
  const uint64_t s2 = o[4] | (static_cast<uint64_t>(o[5]) << 32), s3 = o[6] | (static_cast<uint64_t>(o[7]) << 32);
  const u128 initstate = (static_cast<u128>(s0) << 64) | s1;
  const u128 initseq = (static_cast<u128>(s2) << 64) | s3;
  g.inc = (initseq << 1) | 1;
  g.state = 0;
  pcg_step(g);
  g.state += initstate;
  pcg_step(g);
  g.has32 = false;
  g.buf32 = 0;
}
__device__ __forceinline__ uint64_t pcg_next64(Pcg64& g) {
  pcg_step(g);
  const uint64_t hi = static_cast<uint64_t>(g.state >> 64), lo = static_cast<uint64_t>(g.state);
  const uint64_t x = hi ^ lo;
  const unsigned rot = static_cast<unsigned>(hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}
__device__ __forceinline__ uint32_t pcg_next32(Pcg64& g) {
  if (g.has32) {
    g.has32 = false;
    return g.buf32;
  }
  const uint64_t n = pcg_next64(g);
  g.has32 = true;
  g.buf32 = static_cast<uint32_t>(n >> 32);
  return static_cast<uint32_t>(n);
}
// numpy random_bounded_uint64(off = 0, rng = r) for r < 2^32 - 1: Lemire's method on 32-bit draws.
__device__ __forceinline__ uint32_t pcg_bounded(Pcg64& g, uint32_t r) {
  if (r == 0) return 0;
  const uint32_t excl = r + 1;
  uint64_t m = static_cast<uint64_t>(pcg_next32(g)) * excl;
  uint32_t left = static_cast<uint32_t>(m);
  if (left < excl) {
    const uint32_t thr = (0xFFFFFFFFu - r) % excl;
    while (left < thr) {
      m = static_cast<uint64_t>(pcg_next32(g)) * excl;
      left = static_cast<uint32_t>(m);
    }
  }
  return static_cast<uint32_t>(m >> 32);
}

constexpr int kMaxWords = 32;  // T' <= 1024
constexpr int kMaxCalls = 8;   // context + up to 7 target groups

struct MaskCfg {
  int kind;  // 0 = TimeInverseBlockMasker, 1 = SpeechMasker
  int T;     // T' = n_times / in_channels
  int C;     // channel expansion factor of the outputs ("(S C)" interleave), 1 = none
  int G;
  double p_ctx;
  int len_ctx;
  double p_tgt;
  int len_tgt;
  float cutoff;
  int min_ctx_len;
  uint32_t base_seed;
  uint32_t row0;
  int batch;
  int max_attempts;
};

// compute_mask_indices(shape=(1,T), None, p, len): returns false when num_mask == 0 (the reference raises there).
__device__ bool span_mask(uint32_t w0, uint32_t w1, uint32_t w2, int T, double p, int len, uint32_t* out,
                          uint32_t* chosen) {
  Pcg64 g;
  pcg_seed(g, w0, w1, w2);
  const double u = static_cast<double>(pcg_next64(g) >> 11) * (1.0 / 9007199254740992.0);
  // ((p * T) / len) + u in IEEE double without FMA contraction, truncated toward zero (audio_masking.py:82-87)
  const double x = __dadd_rn(__ddiv_rn(__dmul_rn(p, static_cast<double>(T)), static_cast<double>(len)), u);
  const int num = static_cast<int>(x);
  for (int i = 0; i < kMaxWords; ++i) { out[i] = 0u; chosen[i] = 0u; }
  if (num <= 0) return false;
  int min_len = len;
  if (T - min_len <= num) min_len = T - num - 1;
  const int n = T - min_len;
  if (n < num || n <= 0) return false;
  // Floyd's sampling (Generator.choice, replace=False); the trailing shuffle cannot change the painted set.
  for (int j = n - num; j < n; ++j) {
    uint32_t v = pcg_bounded(g, static_cast<uint32_t>(j));
    if ((chosen[v >> 5] >> (v & 31)) & 1u) v = static_cast<uint32_t>(j);
    chosen[v >> 5] |= 1u << (v & 31);
    const int e = min(static_cast<int>(v) + len, T);
    for (int t = static_cast<int>(v); t < e; ++t) out[t >> 5] |= 1u << (t & 31);
  }
  return true;
}

__global__ void __launch_bounds__(128) masks_kernel(MaskCfg cfg, uint8_t* __restrict__ ctx_hidden,
                                                    uint8_t* __restrict__ tgt, uint8_t* __restrict__ vis_hidden,
                                                    int* __restrict__ attempts, int* __restrict__ err) {
  __shared__ uint32_t s_mask[4][kMaxCalls][kMaxWords];
  __shared__ uint32_t s_scr[4][kMaxCalls][kMaxWords];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + wib;
  if (row >= cfg.batch) return;
  const int T = cfg.T, G = cfg.G;
  const int nwords = (T + 31) >> 5;
  const uint32_t grow = cfg.row0 + static_cast<uint32_t>(row);
  uint32_t ctxw = 0;
  int attempt = 0;
  bool failed = false;
  while (true) {
    // lanes 0..G each run one independent RNG stream
    bool ok = true;
    if (lane <= G && !(cfg.kind == 1 && lane == 0)) {
      const double p = lane == 0 ? cfg.p_ctx : cfg.p_tgt;
      const int len = lane == 0 ? cfg.len_ctx : cfg.len_tgt;
      ok = span_mask(cfg.base_seed, grow, static_cast<uint32_t>(attempt * 8 + lane), T, p, len, s_mask[wib][lane],
                     s_scr[wib][lane]);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (!ok) { failed = true; break; }
    __syncwarp();
    const uint32_t valid = (lane < nwords) ? ((lane == nwords - 1 && (T & 31)) ? ((1u << (T & 31)) - 1u) : 0xFFFFFFFFu) : 0u;
    uint32_t anyt = 0;
    for (int g = 1; g <= G; ++g) anyt |= s_mask[wib][g][lane];
    if (cfg.kind == 0) {
      ctxw = ~s_mask[wib][0][lane] & ~anyt & valid;  // masking.py:90-107
    } else {
      ctxw = ~anyt & valid;  // masking.py:187-188
      // filter_small_clusters (masking.py:150-165): True-runs shorter than min_ctx_len -> False
      s_scr[wib][0][lane] = ctxw;
      __syncwarp();
      if (lane == 0) {
        uint32_t* w = s_scr[wib][0];
        int run_start = -1;
        for (int t = 0; t <= T; ++t) {
          const bool bit = (t < T) && ((w[t >> 5] >> (t & 31)) & 1u);
          if (bit && run_start < 0) run_start = t;
          if (!bit && run_start >= 0) {
            if (t - run_start < cfg.min_ctx_len)
              for (int k = run_start; k < t; ++k) w[k >> 5] &= ~(1u << (k & 31));
            run_start = -1;
          }
        }
      }
      __syncwarp();
      ctxw = s_scr[wib][0][lane];
    }
    int cnt = __popc(ctxw);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    ++attempt;
    // torch: sum(int64) / n_times -> float32 true division, compared with float32(cutoff) (masking.py:108-110)
    const float ratio = __fdiv_rn(static_cast<float>(cnt), static_cast<float>(T));
    if (ratio >= cfg.cutoff) break;
    if (attempt >= cfg.max_attempts) { failed = true; break; }
    __syncwarp();
  }
  if (failed) {
    if (lane == 0) atomicExch(err, 1);
    return;
  }
  if (lane == 0 && attempts != nullptr) attempts[row] = attempt;
  // publish the accepted context word so that every lane can index arbitrary positions
  s_scr[wib][1][lane] = ctxw;
  __syncwarp();
  const int C = cfg.C, Tout = T * C;
  const uint32_t* cw = s_scr[wib][1];
  for (int i = lane; i < Tout; i += 32) {
    const int s = i / C;
    const uint32_t cbit = (cw[s >> 5] >> (s & 31)) & 1u;
    const uint8_t hidden = cbit ? 0 : 1;  // final_context_mask = ~context_positions (masking.py:115)
    ctx_hidden[static_cast<size_t>(row) * Tout + i] = hidden;
    for (int g = 0; g < G; ++g) {
      const uint8_t tb = (s_mask[wib][g + 1][s >> 5] >> (s & 31)) & 1u;
      const size_t o = (static_cast<size_t>(row) * G + g) * Tout + i;
      tgt[o] = tb;
      vis_hidden[o] = hidden ^ tb;  // logical_xor (masking.py:116)
    }
  }
}

// ---------------------------------------------------------------------------------------------- index lists
__global__ void __launch_bounds__(128) mask_count_kernel(const uint8_t* __restrict__ ctx_hidden,
                                                         const uint8_t* __restrict__ tgt,
                                                         const uint8_t* __restrict__ vis_hidden, int B, int G, int T,
                                                         int* __restrict__ n_c, int* __restrict__ n_v,
                                                         int* __restrict__ n_t, int* __restrict__ totals) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  int c = 0;
  for (int t = lane; t < T; t += 32) c += ctx_hidden[static_cast<size_t>(b) * T + t] ? 0 : 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) n_c[b] = c;
  int viol = 0;
  for (int g = 0; g < G; ++g) {
    int v = 0, tt = 0;
    const size_t base = (static_cast<size_t>(b) * G + g) * T;
    for (int t = lane; t < T; t += 32) {
      const bool vis = !vis_hidden[base + t];
      const bool tg = tgt[base + t];
      v += vis;
      tt += (tg && vis);
      viol += (tg && !vis);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v += __shfl_xor_sync(0xffffffffu, v, o);
      tt += __shfl_xor_sync(0xffffffffu, tt, o);
    }
    if (lane == 0) { n_v[b * G + g] = v; n_t[b * G + g] = tt; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) viol += __shfl_xor_sync(0xffffffffu, viol, o);
  if (lane == 0 && viol) atomicAdd(&totals[3], viol);
}

// single block: exclusive scans of n_c (B), n_v and n_t (B*G); totals = {Nc, Nv, Nt, violations, max_nc, max_nv}
__global__ void __launch_bounds__(1024) mask_scan_kernel(const int* __restrict__ n_c, const int* __restrict__ n_v,
                                                         const int* __restrict__ n_t, int B, int BG,
                                                         int* __restrict__ cu_c, int* __restrict__ cu_v,
                                                         int* __restrict__ cu_t, int* __restrict__ totals) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  __shared__ int s_max;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int which = 0; which < 3; ++which) {
    const int* src = which == 0 ? n_c : (which == 1 ? n_v : n_t);
    int* dst = which == 0 ? cu_c : (which == 1 ? cu_v : cu_t);
    const int n = which == 0 ? B : BG;
    if (tid == 0) { s_carry = 0; s_max = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
      const int i = base + tid;
      const int v = i < n ? src[i] : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) s_warp[w] = x;
      __syncthreads();
      if (w == 0) {
        int ws = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, ws, o);
          if (lane >= o) ws += y;
        }
        s_warp[lane] = ws;
      }
      __syncthreads();
      const int carry = s_carry;
      const int incl = x + (w > 0 ? s_warp[w - 1] : 0) + carry;
      if (i < n) dst[i] = incl - v;
      atomicMax(&s_max, v);
      __syncthreads();
      if (tid == 1023) s_carry = incl;
      __syncthreads();
    }
    if (tid == 0) {
      dst[n] = s_carry;
      totals[which] = s_carry;
      if (which < 2) totals[4 + which] = s_max;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) mask_fill_kernel(const uint8_t* __restrict__ ctx_hidden,
                                                        const uint8_t* __restrict__ tgt,
                                                        const uint8_t* __restrict__ vis_hidden, int B, int G, int T,
                                                        const int* __restrict__ cu_c, const int* __restrict__ cu_v,
                                                        const int* __restrict__ cu_t, int* __restrict__ ctx_rows,
                                                        int* __restrict__ vis_src, int* __restrict__ vis_pos,
                                                        int* __restrict__ tgt_vrow, int* __restrict__ tgt_trow) {
  extern __shared__ int s_idx[];  // [4][T] packed context index of every position, -1 if hidden
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + wib;
  if (b >= B) return;
  int* my = s_idx + wib * T;
  const uint32_t lt = (1u << lane) - 1u;
  int base = cu_c[b];
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool vis = t < T && !ctx_hidden[static_cast<size_t>(b) * T + t];
    const uint32_t bal = __ballot_sync(0xffffffffu, vis);
    const int r = base + __popc(bal & lt);
    if (t < T) my[t] = vis ? r : -1;
    if (vis) ctx_rows[r] = b * T + t;
    base += __popc(bal);
  }
  __syncwarp();
  for (int g = 0; g < G; ++g) {
    const int s = b * G + g;
    const size_t mb = static_cast<size_t>(s) * T;
    int vb = cu_v[s], tb = cu_t[s];
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      const bool vis = t < T && !vis_hidden[mb + t];
      const bool tg = vis && tgt[mb + t];
      const uint32_t bv = __ballot_sync(0xffffffffu, vis);
      const uint32_t bt = __ballot_sync(0xffffffffu, tg);
      const int rv = vb + __popc(bv & lt);
      if (vis) {
        vis_src[rv] = my[t];  // context feature where the context is visible, else the mask token (jepa.py:425-427)
        vis_pos[rv] = t;
      }
      if (tg) {
        const int rt = tb + __popc(bt & lt);
        tgt_vrow[rt] = rv;
        tgt_trow[rt] = b * T + t;
      }
      vb += __popc(bv);
      tb += __popc(bt);
    }
  }
}

}  // namespace wj

using namespace wj;

extern "C" int wj_masks_generate(int kind, int batch, int n_times, int in_channels, int channel_based, int n_targets,
                                 double ctx_prob, int ctx_len, double tgt_prob, int tgt_len, float cutoff,
                                 int min_context_len, uint32_t base_seed, uint32_t row0, uint8_t* ctx_hidden,
                                 uint8_t* tgt, uint8_t* vis_hidden, int* attempts, int* err_flag, void* stream) {
  if (batch <= 0) return WJ_OK;
  if (in_channels <= 0 || n_times % in_channels != 0) { set_error("wj_masks_generate: n_times %% in_channels != 0"); return WJ_ERR_ARG; }
  const int T = n_times / in_channels;
  if (T <= 0 || T > kMaxWords * 32) { set_error("wj_masks_generate: T'=%d out of range (<= 1024)", T); return WJ_ERR_ARG; }
  if (n_targets < 1 || n_targets > kMaxCalls - 1) { set_error("wj_masks_generate: n_targets must be in [1,7]"); return WJ_ERR_ARG; }
  if (tgt_len <= 0 || (kind == 0 && ctx_len <= 0)) { set_error("wj_masks_generate: span length must be positive"); return WJ_ERR_ARG; }
  if (kind != 0 && kind != 1) { set_error("wj_masks_generate: kind must be 0 (TimeInverse) or 1 (Speech)"); return WJ_ERR_ARG; }
  MaskCfg cfg;
  cfg.kind = kind; cfg.T = T; cfg.C = channel_based ? in_channels : 1; cfg.G = n_targets;
  cfg.p_ctx = ctx_prob; cfg.len_ctx = ctx_len; cfg.p_tgt = tgt_prob; cfg.len_tgt = tgt_len; cfg.cutoff = cutoff;
  cfg.min_ctx_len = min_context_len; cfg.base_seed = base_seed; cfg.row0 = row0; cfg.batch = batch;
  cfg.max_attempts = 1 << 20;
  masks_kernel<<<(batch + 3) / 4, 128, 0, WJ_STREAM(stream)>>>(cfg, ctx_hidden, tgt, vis_hidden, attempts, err_flag);
  return check_launch("masks_kernel");
}

extern "C" int wj_mask_indices(const uint8_t* ctx_hidden, const uint8_t* tgt, const uint8_t* vis_hidden, int B, int G,
                               int T, int* n_c, int* n_v, int* n_t, int* cu_c, int* cu_v, int* cu_t, int* totals,
                               int* ctx_rows, int* vis_src, int* vis_pos, int* tgt_vrow, int* tgt_trow, void* stream) {
  if (B <= 0) return WJ_OK;
  if (T > 2048) { set_error("wj_mask_indices: T too large"); return WJ_ERR_ARG; }
  cudaStream_t st = WJ_STREAM(stream);
  cudaMemsetAsync(totals, 0, 8 * sizeof(int), st);
  mask_count_kernel<<<(B + 3) / 4, 128, 0, st>>>(ctx_hidden, tgt, vis_hidden, B, G, T, n_c, n_v, n_t, totals);
  mask_scan_kernel<<<1, 1024, 0, st>>>(n_c, n_v, n_t, B, B * G, cu_c, cu_v, cu_t, totals);
  mask_fill_kernel<<<(B + 3) / 4, 128, 4 * T * sizeof(int), st>>>(ctx_hidden, tgt, vis_hidden, B, G, T, cu_c, cu_v, cu_t,
                                                                  ctx_rows, vis_src, vis_pos, tgt_vrow, tgt_trow);
  return check_launch("mask_indices", 3);
}
