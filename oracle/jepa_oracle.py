"""TEST INFRASTRUCTURE ONLY (oracle) -- plain fp32 PyTorch-CPU restatement of the WavJEPA hot path.

This file restates, op by op, what the reference (labhamlet/wavjepa) computes for one pre-training forward/backward
and for HEAR feature extraction, using nothing but dense torch tensor math on the CPU (no nn.Transformer*, no
autocast, no CUDA).  It is the checker for the CUDA path and the `cpu_baseline` / `--impl reference` arm of bench.py;
the product package (wavjepa_b200/) never imports it.

Each function cites the reference lines it follows (paths relative to the reference checkout):
  conv stack            wavjepa/extractors/audio_feature_extractor.py:52-122 (block structure), :124-138 (forward)
  per-channel extractor wavjepa/extractors/audio_channel_feature_extractor.py:154-179
  transformer layer     torch.nn.TransformerEncoderLayer semantics as configured by wavjepa/types/wavjepa_configs.py:29-47
                        (post-norm, GELU(erf), eps 1e-6, dropout 0) and built at wavjepa/jepa.py:126-130
  forward               wavjepa/jepa.py:365-419;  decoder_forward :422-440;  encoder_forward :444-454
  teacher + targets     wavjepa/jepa.py:230-270
  loss                  wavjepa/jepa.py:335-362
  EMA                   wavjepa/jepa.py:186-198
  crop + normalise      wavjepa/jepa.py:291-311
  HEAR runtime          hear_api/runtime.py:12-35, 98-155; hear_api/feature_helper.py:5-13
Pinning status: PINNED against the unmodified reference executed in the build container -- tests/golden/*.npz were
produced by tests/golden/make_golden.py from the reference's own modules; tests/test_oracle_cpu.py checks this file
against them.  (The reference itself ships no tests or golden vectors: SURVEY.md 4.)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

BASE_SPEC = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)]


class Cfg:
    def __init__(self, spec=None, in_channels=1, d_model=768, nhead=12, layers=12, d_dec=384, dec_heads=12,
                 dec_layers=12, top_k=8, seconds=2.01, sr=16000, per_channel=False, mlp_ratio=4.0):
        self.spec = list(spec or BASE_SPEC)
        self.in_channels, self.per_channel = in_channels, per_channel
        self.d_model, self.nhead, self.layers = d_model, nhead, layers
        self.d_dec, self.dec_heads, self.dec_layers = d_dec, dec_heads, dec_layers
        self.ff, self.ff_dec = int(d_model * mlp_ratio), int(d_dec * mlp_ratio)
        self.top_k = top_k
        self.target_length = int(sr * seconds)          # jepa.py:101 -> 32159 for 2.01 s
        L = self.target_length
        for (_, k, s) in self.spec:
            L = (L - k) // s + 1
        self.tokens_per_channel = L
        self.total_patches = L * (in_channels if per_channel else 1)
        self.conv_dim = self.spec[-1][0]


# ------------------------------------------------------------------------------------------------- weights
def sincos_table(n_pos: int, dim: int) -> torch.Tensor:
    """pos_embed.py:75-93 via jepa.py:163-180 (float64 math, stored fp32, [1, n_pos, dim])."""
    half = dim // 2
    omega = 1.0 / (10000.0 ** (np.arange(half, dtype=np.float64) / (dim / 2.0)))
    ang = np.arange(n_pos, dtype=np.float64)[:, None] * omega[None, :]
    return torch.from_numpy(np.concatenate([np.sin(ang), np.cos(ang)], axis=1)).float().unsqueeze(0)


def _stack_shapes(prefix: str, d: int, ff: int, n_layers: int) -> Dict[str, tuple]:
    s = {}
    for i in range(n_layers):
        b = f"{prefix}.layers.{i}."
        s[b + "self_attn.in_proj_weight"] = (3 * d, d)
        s[b + "self_attn.in_proj_bias"] = (3 * d,)
        s[b + "self_attn.out_proj.weight"] = (d, d)
        s[b + "self_attn.out_proj.bias"] = (d,)
        s[b + "linear1.weight"] = (ff, d)
        s[b + "linear1.bias"] = (ff,)
        s[b + "linear2.weight"] = (d, ff)
        s[b + "linear2.bias"] = (d,)
        for nm in ("norm1", "norm2"):
            s[b + nm + ".weight"] = (d,)
            s[b + nm + ".bias"] = (d,)
    s[f"{prefix}.norm.weight"] = (d,)
    s[f"{prefix}.norm.bias"] = (d,)
    return s


def param_shapes(cfg: Cfg) -> Dict[str, tuple]:
    """Names / shapes of the reference state_dict, in the reference's registration order."""
    s: Dict[str, tuple] = {"mask_token": (1, 1, cfg.d_dec)}
    s["pos_encoding_encoder"] = (1, cfg.total_patches, cfg.d_model)
    s["pos_encoding_decoder"] = (1, cfg.total_patches, cfg.d_dec)
    prefixes = [f"extract_audio.cnns.{c}" for c in range(cfg.in_channels)] if cfg.per_channel else ["extract_audio.cnn"]
    cin0 = 1 if cfg.per_channel else cfg.in_channels
    for p in prefixes:
        cin = cin0
        for i, (dim, k, _) in enumerate(cfg.spec):
            s[f"{p}.{i}.0.weight"] = (dim, cin, k)
            if i == 0:
                s[f"{p}.0.2.weight"] = (dim,)
                s[f"{p}.0.2.bias"] = (dim,)
            cin = dim
    s["feature_norms.weight"] = (cfg.conv_dim,)
    s["feature_norms.bias"] = (cfg.conv_dim,)
    s.update(_stack_shapes("encoder", cfg.d_model, cfg.ff, cfg.layers))
    s["post_extraction_mapper.weight"] = (cfg.d_model, cfg.conv_dim)
    s["post_extraction_mapper.bias"] = (cfg.d_model,)
    s.update(_stack_shapes("decoder", cfg.d_dec, cfg.ff_dec, cfg.dec_layers))
    s["decoder_to_encoder_mapper.weight"] = (cfg.d_model, cfg.d_dec)
    s["decoder_to_encoder_mapper.bias"] = (cfg.d_model,)
    s["encoder_to_decoder_mapper.weight"] = (cfg.d_dec, cfg.d_model)
    s["encoder_to_decoder_mapper.bias"] = (cfg.d_dec,)
    s.update(_stack_shapes("teacher_encoder", cfg.d_model, cfg.ff, cfg.layers))
    return s


def make_state_dict(cfg: Cfg, seed: int = 0, perturb: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights with the reference's init STATISTICS (jepa.py:135-161, extractor :72): conv
    kaiming-normal, linear N(0, 0.02), in_proj xavier-uniform, norms ~1/0.  With perturb=True biases / norm affine
    parameters get small random values and the teacher differs from the student, so that parity tests exercise
    every parameter (a fresh reference init has zero biases and teacher == student).  One torch CPU generator per
    tensor, seeded from (seed, index): identical on every machine."""
    sd: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(param_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if name == "pos_encoding_encoder":
            t = sincos_table(cfg.total_patches, cfg.d_model)
        elif name == "pos_encoding_decoder":
            t = sincos_table(cfg.total_patches, cfg.d_dec)
        elif name == "mask_token":
            t = torch.randn(shape, generator=g) * 0.02
        elif name.startswith("extract_audio") and name.endswith(".0.weight"):
            fan_in = shape[1] * shape[2]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif name.endswith("in_proj_weight"):
            a = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif name.endswith("weight") and len(shape) == 2:
            t = torch.randn(shape, generator=g) * 0.02
        elif name.endswith("weight"):  # norm scales
            t = torch.ones(shape) + (0.05 * torch.randn(shape, generator=g) if perturb else 0.0)
        else:  # biases
            t = 0.02 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape)
        sd[name] = t.float().contiguous()
    if not perturb:
        for k in list(sd):
            if k.startswith("teacher_encoder."):
                sd[k] = sd["encoder." + k[len("teacher_encoder."):]].clone()
    return sd


# ------------------------------------------------------------------------------------------------- ops
def conv_stack(x: torch.Tensor, sd, prefix: str, spec) -> torch.Tensor:
    """[B, Cin, L] -> [B, T, C].  Block 0: conv -> GroupNorm(C groups) -> GELU; blocks 1..: conv -> GELU; no bias."""
    for i, (dim, k, s) in enumerate(spec):
        x = F.conv1d(x, sd[f"{prefix}.{i}.0.weight"], stride=s)
        if i == 0:
            x = F.group_norm(x, dim, sd[f"{prefix}.0.2.weight"], sd[f"{prefix}.0.2.bias"], eps=1e-5)
        x = F.gelu(x)
    return x.transpose(1, 2)


def extract(x: torch.Tensor, sd, cfg: Cfg) -> torch.Tensor:
    if not cfg.per_channel:
        return conv_stack(x, sd, "extract_audio.cnn", cfg.spec)
    outs = [conv_stack(x[:, c:c + 1], sd, f"extract_audio.cnns.{c}", cfg.spec) for c in range(cfg.in_channels)]
    return torch.cat(outs, dim=1)  # channel-major tokens


def encoder_layer(x: torch.Tensor, sd, p: str, nhead: int, key_hidden: Optional[torch.Tensor], eps: float = 1e-6):
    """Post-norm layer: x = LN1(x + out_proj(MHA(x))); x = LN2(x + W2 GELU(W1 x)).  key_hidden [B, T] True = key masked."""
    B, T, d = x.shape
    dh = d // nhead
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(d, dim=-1)
    q = q.view(B, T, nhead, dh).transpose(1, 2)
    k = k.view(B, T, nhead, dh).transpose(1, 2)
    v = v.view(B, T, nhead, dh).transpose(1, 2)
    logits = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if key_hidden is not None:
        logits = logits.masked_fill(key_hidden[:, None, None, :], float("-inf"))
    att = torch.softmax(logits, dim=-1) @ v
    att = att.transpose(1, 2).reshape(B, T, d)
    x = F.layer_norm(x + F.linear(att, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"]),
                     (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    h = F.gelu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
    x = F.layer_norm(x + F.linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"]),
                     (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    return x


def run_stack(x, sd, prefix: str, n_layers: int, nhead: int, key_hidden, final_norm: bool = True, collect=None):
    d = x.shape[-1]
    for i in range(n_layers):
        x = encoder_layer(x, sd, f"{prefix}.layers.{i}.", nhead, key_hidden)
        if collect is not None:
            collect(i, x)
    if final_norm:
        x = F.layer_norm(x, (d,), sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"], 1e-5)
    return x


def local_features(audio, sd, cfg: Cfg):
    """jepa.py:391-396."""
    f = extract(audio, sd, cfg)
    f = F.layer_norm(f, (cfg.conv_dim,), sd["feature_norms.weight"], sd["feature_norms.bias"], 1e-5)
    f = F.linear(f, sd["post_extraction_mapper.weight"], sd["post_extraction_mapper.bias"])
    return f + sd["pos_encoding_encoder"]


def teacher_targets(x, sd, cfg: Cfg):
    """jepa.py:255-270 + :230-253: no mask, no final norm; top-K layer outputs, each normalised per instance over
    all (T, D) values jointly (F.instance_norm on the 4-D stack: biased variance, eps 1e-5, no affine), averaged."""
    outs: List[torch.Tensor] = []

    def collect(i, y):
        if cfg.layers - i <= cfg.top_k:
            outs.append(y)

    with torch.no_grad():
        last = run_stack(x.detach(), sd, "teacher_encoder", cfg.layers, cfg.nhead, None, final_norm=False,
                         collect=collect)
        if cfg.top_k <= 1:
            return last
        acc = torch.zeros_like(outs[0])
        for y in outs:
            mu = y.mean(dim=(1, 2), keepdim=True)
            var = y.var(dim=(1, 2), keepdim=True, unbiased=False)
            acc += (y - mu) / torch.sqrt(var + 1e-5)
        return acc / len(outs)


def forward(audio, ctx_masks, target_indices, ctx_and_target_masks, sd, cfg: Cfg) -> Dict[str, torch.Tensor]:
    """Dense, mask-as-key-padding restatement of JEPA.forward (jepa.py:365-419): same dict as the reference."""
    B = audio.shape[0]
    G = target_indices.shape[1]
    T = cfg.total_patches
    local = local_features(audio, sd, cfg)
    ctx = run_stack(local, sd, "encoder", cfg.layers, cfg.nhead, ctx_masks)
    ctx = ctx[~ctx_masks]                                                                     # [sum n_c, D]
    ctx = F.linear(ctx, sd["encoder_to_decoder_mapper.weight"], sd["encoder_to_decoder_mapper.bias"])
    tgt = sd["mask_token"].repeat(B, T, 1)
    tgt = tgt.masked_scatter((~ctx_masks)[..., None].expand(-1, -1, cfg.d_dec), ctx)            # jepa.py:425-427
    tgt = tgt + sd["pos_encoding_decoder"]
    tgt = tgt[:, None].expand(-1, G, -1, -1).reshape(B * G, T, cfg.d_dec)
    hidden = ctx_and_target_masks.reshape(B * G, T)
    dec = run_stack(tgt, sd, "decoder", cfg.dec_layers, cfg.dec_heads, hidden)
    preds = F.linear(dec, sd["decoder_to_encoder_mapper.weight"], sd["decoder_to_encoder_mapper.bias"])
    targets = teacher_targets(local, sd, cfg)
    # jepa.py:343-362
    per_t = ((preds.view(B, G, T, -1) - targets[:, None]) ** 2).mean(dim=-1)
    loss = (per_t * target_indices).sum() / (target_indices.sum() + 1e-8)
    return dict(local_features=local, contextual_features=ctx, loss=loss, preds=preds, targets=targets)


def ema_decay(step: int, r0: float = 0.999, r1: float = 0.99999, end: int = 100000) -> float:
    """jepa.py:186-191."""
    if step >= end:
        return r1
    return r1 - (r1 - r0) * (1 - step / end)


def ema_update(sd, step: int) -> None:
    """jepa.py:193-198 (in place on the teacher entries of sd)."""
    r = ema_decay(step)
    for k in list(sd):
        if k.startswith("teacher_encoder."):
            s = sd["encoder." + k[len("teacher_encoder."):]]
            sd[k] = sd[k].detach().mul(r).add((1 - r) * s.detach())


def crop_normalise(clips: torch.Tensor, starts: torch.Tensor, length: int) -> torch.Tensor:
    """jepa.py:291-311: clips [n, C, L_full], starts [n, S] -> [n*S, C, length] fp32 (caller casts to bf16)."""
    n, C, _ = clips.shape
    S = starts.shape[1]
    idx = starts[:, :, None] + torch.arange(length)
    idx = idx[:, :, None, :].expand(-1, -1, C, -1)
    crops = torch.gather(clips[:, None].expand(-1, S, -1, -1), 3, idx)
    mean = crops.mean(dim=(-2, -1), keepdim=True)
    std = crops.std(dim=(-2, -1), keepdim=True)
    return ((crops - mean) / (std + 1e-5)).flatten(0, 1)


# ------------------------------------------------------------------------------------------------- HEAR
def hear_loudness(audio: torch.Tensor) -> torch.Tensor:
    """feature_helper.py:5-13 per clip: x * 10^((-14 - 20 log10(rms)) / 20); audio [B, L]."""
    out = []
    for a in audio:
        rms = torch.sqrt(torch.mean(a ** 2))
        if rms == 0:
            out.append(a)
            continue
        gain = 10 ** ((-14.0 - 20 * torch.log10(rms)) / 20)
        out.append(a * gain)
    return torch.stack(out)


def hear_geometry(n_samples: int, unit: int, sr: int, steps: int):
    """runtime.py:107-124, 19-35: (pad, n_chunks, cut_off, pad_steps).  process_seconds floors to an int (:123)."""
    pad = unit - (n_samples % unit)
    total = n_samples + pad
    proc = unit // sr
    n_units = int((total / sr) / proc)
    total_steps = steps * n_units
    out_sr = int(steps / proc)
    pad_steps = int((pad / sr) * out_sr)
    return pad, total // unit, total_steps - pad_steps, pad_steps


def hear_timestamp_embeddings(audio: torch.Tensor, sd, cfg: Cfg, sr: int = 16000):
    """hear_api/runtime.py:98-155 (get_timestamp_embeddings) for mono [B, L] input, model in_channels == 1."""
    B, L = audio.shape
    unit, steps = cfg.target_length, cfg.total_patches
    x = hear_loudness(audio)[:, None, :]
    pad, n_chunks, cut_off, _ = hear_geometry(L, unit, sr, steps)
    x = F.pad(x, (0, pad))
    total_steps = steps * int(((L + pad) / sr) / (unit // sr))
    mask = torch.zeros(B, max(total_steps, n_chunks * steps), dtype=torch.bool)
    mask[:, cut_off:total_steps] = True
    embs = []
    for i in range(n_chunks):
        chunk = x[..., i * unit:(i + 1) * unit]
        mu = chunk.mean(dim=(-2, -1), keepdim=True)
        sdv = chunk.std(dim=(-2, -1), keepdim=True)
        chunk = (chunk - mu) / (sdv + 1e-5)
        m = mask[:, i * steps:(i + 1) * steps]
        loc = local_features(chunk, sd, cfg)
        embs.append(run_stack(loc, sd, "encoder", cfg.layers, cfg.nhead, m))
    emb = torch.cat(embs, dim=1)[:, :cut_off]
    n = emb.shape[1]
    step_ms = (L / sr) / n * 1000
    ts = torch.tensor([step_ms * i for i in range(n)]).unsqueeze(0).repeat(B, 1)
    return emb, ts


def hear_nat_feature(audio: torch.Tensor) -> torch.Tensor:
    """feature_helper.py:27-88 with in_channels = 2, for a batch [B, L] / [B, C, L]: loudness over the whole (C, L) clip,
    mono duplicated (:57-58), stereo kept (:68-69), first of four channels duplicated (:75-77)."""
    out = []
    for a in audio:
        if a.ndim == 1:
            a = a.unsqueeze(0)
        rms = torch.sqrt(torch.mean(a ** 2))
        if rms != 0:
            a = a * 10 ** ((-14.0 - 20 * torch.log10(rms)) / 20)
        if a.shape[0] == 1:
            a = torch.cat((a, a), dim=0)
        elif a.shape[0] == 4:
            a = torch.cat((a[:1], a[:1]), dim=0)
        out.append(a)
    return torch.stack(out)


def hear_nat_timestamp_embeddings(audio: torch.Tensor, sd, cfg: Cfg, sr: int = 16000):
    """hear_api/runtime_natjepa.py:102-151 (RuntimeNatJEPA.get_timestamp_embeddings) for the 2-channel model (cfg.per_channel,
    cfg.in_channels == 2): chunks normalised over (C, T) jointly (:12-16), the time mask repeated over the channel-major
    tokens (`B E -> B (C E)`, :142), per-channel embeddings averaged (`B (C S) E -> B C S E`, mean over C, :144-147)."""
    B = audio.shape[0]
    C = cfg.in_channels
    unit = cfg.target_length
    steps = cfg.total_patches // C          # tokens per channel (:83-86)
    x = hear_nat_feature(audio)
    L = x.shape[-1]
    pad, n_chunks, cut_off, _ = hear_geometry(L, unit, sr, steps)
    x = F.pad(x, (0, pad))
    total_steps = steps * int(((L + pad) / sr) / (unit // sr))
    mask = torch.zeros(B, max(total_steps, n_chunks * steps), dtype=torch.bool)
    mask[:, cut_off:total_steps] = True
    embs = []
    for i in range(n_chunks):
        chunk = x[..., i * unit:(i + 1) * unit]
        mu = chunk.mean(dim=(-2, -1), keepdim=True)
        sdv = chunk.std(dim=(-2, -1), keepdim=True)
        chunk = (chunk - mu) / (sdv + 1e-5)
        m = mask[:, i * steps:(i + 1) * steps].repeat(1, C)
        loc = local_features(chunk, sd, cfg)
        e = run_stack(loc, sd, "encoder", cfg.layers, cfg.nhead, m)
        embs.append(e.view(B, C, steps, -1).mean(dim=1))
    emb = torch.cat(embs, dim=1)[:, :cut_off]
    n = emb.shape[1]
    step_ms = (L / sr) / n * 1000
    ts = torch.tensor([step_ms * i for i in range(n)]).unsqueeze(0).repeat(B, 1)
    return emb, ts


def arch_get_embeddings(audio: torch.Tensor, sd, cfg: Cfg, sr: int = 16000) -> torch.Tensor:
    """ARCH/configs/wavjepa_wrapper.py:67-110 for one clip [L]: loudness-normalise, pad, per-chunk normalise + encode with
    the padded frames key-masked, keep the unmasked frames, mean over all kept frames -> [D]."""
    emb, _ = hear_timestamp_embeddings(audio.reshape(1, -1), sd, cfg, sr)
    return emb[0].mean(dim=0)


def data_pre_process(waveform: torch.Tensor, audio_sr: int, sr: int = 16000, seconds: int = 10) -> torch.Tensor:
    """data_modules/WebAudioDataModule.py:43-61 + dataset_functions.py:92-114: first channel, torchaudio Kaiser-sinc
    resampling (the reference's arguments), RMS normalisation to -14 dBFS over the whole resampled clip, zero-pad / crop
    to `seconds` -> [1, sr*seconds] fp32.  torchaudio is the reference's own third-party resampler (requirements.txt)."""
    import torchaudio

    audio = waveform[0, :] if waveform.ndim > 1 else waveform
    if audio_sr != sr:
        audio = torchaudio.functional.resample(audio, audio_sr, sr, lowpass_filter_width=64, rolloff=0.9475937167399596,
                                               resampling_method="sinc_interp_kaiser", beta=14.769656459379492)
    rms = torch.sqrt(torch.mean(audio ** 2))
    if rms != 0:
        audio = audio * 10 ** ((-14.0 - 20 * torch.log10(rms)) / 20)
    audio = audio.reshape(1, -1)
    padding = sr * seconds - audio.shape[1]
    if padding > 0:
        audio = F.pad(audio, (0, padding), "constant", 0)
    elif padding < 0:
        audio = audio[:, : sr * seconds]
    return audio


# ------------------------------------------------------------------------------------------------- denoiser stage
def student_features(audio, sd, cfg: Cfg):
    """extractor -> LayerNorm -> mapper -> + positions -> encoder incl. final norm on full sequences, no mask
    (wavjepa/denoiser.py:338-346 `_forward_features`; also JEPA.get_audio_representation, jepa.py:456-467)."""
    return run_stack(local_features(audio, sd, cfg), sd, "encoder", cfg.layers, cfg.nhead, None)


def denoiser_forward(generated, clean, sd_student, sd_teacher, cfg: Cfg, alpha: float) -> Dict[str, torch.Tensor]:
    """Denoiser.forward (wavjepa/denoiser.py:308-364): the student on the clean and on the generated scene, a frozen
    WavJEPA teacher on the clean scene, two dense MSE losses mixed by alpha."""
    f_clean = student_features(clean, sd_student, cfg)
    f_gen = student_features(generated, sd_student, cfg)
    with torch.no_grad():
        targets = student_features(clean, sd_teacher, Cfg()).clone()     # the teacher is always WavJEPA-base (:160-171)
    loss_clean = F.mse_loss(f_clean, targets)
    loss_dd = F.mse_loss(f_gen, targets)
    return dict(loss=alpha * loss_clean + (1 - alpha) * loss_dd, loss_clean=loss_clean, loss_denoise_dereverb=loss_dd,
                features_clean=f_clean, features_generated=f_gen, targets=targets)


def fftconvolve_full(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """torchaudio.functional.fftconvolve(mode="full") (third-party, pinned torchaudio 2.7): rfft of both inputs at
    n = len(x) + len(y) - 1, product, irfft."""
    n = x.shape[-1] + y.shape[-1] - 1
    return torch.fft.irfft(torch.fft.rfft(x, n=n) * torch.fft.rfft(y, n=n), n=n)


def scene_convolve_with_rir(waveform: torch.Tensor, rir: torch.Tensor) -> torch.Tensor:
    """generate_scenes_batch.py:12-45: waveform [B, T], rir [B, C, R] -> [B, C, T] (full convolution cut to T)."""
    T = waveform.shape[-1]
    return torch.stack([torch.stack([fftconvolve_full(waveform[b], rir[b, c]) for c in range(rir.shape[1])])
                        for b in range(waveform.shape[0])])[..., :T]


def scene_add_noise(source, noise, snr, start_idx, real_noise_length):
    """generate_scenes_batch.py:107-146: a = sqrt(||x_active||^2 / (||n_active||^2 + 1e-9) * 10^(-snr/10)) over the
    window [start, start + length); source + a * noise.  source, noise [B, 1, T]."""
    B, _, T = source.shape
    t = torch.arange(T).view(1, 1, -1)
    mask = (t >= start_idx.view(B, 1, 1)) & (t < (start_idx + real_noise_length).view(B, 1, 1))
    nx = torch.linalg.vector_norm(source * mask, ord=2, dim=-1, keepdim=True)
    nn_ = torch.linalg.vector_norm(noise * mask, ord=2, dim=-1, keepdim=True)
    a = torch.sqrt(nx ** 2 / (nn_ ** 2 + 1e-9) * 10 ** (-snr.view(B, 1, 1) / 10.0))
    return source + a * noise


def generate_scene(source_rir, noise_rirs, source, noise, real_noise_length, noise_start_idx, snr):
    """generate_scenes_batch.py:148-188, case 1 (source RIR and noise present): first RIR channel only."""
    T = source.shape[-1]
    conv = scene_convolve_with_rir(source, source_rir[:, [0], :])
    agg = torch.zeros(source.shape[0], 1, T)
    for i in range(noise_rirs.shape[1]):
        agg = agg + scene_convolve_with_rir(noise, noise_rirs[:, i, [0], :])
    return scene_add_noise(conv, agg[:, :, :T], snr, noise_start_idx, real_noise_length)


def denoiser_batch(batch, starts: torch.Tensor, perm: torch.Tensor, sr: int = 16000, target_length: int = 32159):
    """Denoiser.on_after_batch_transfer (wavjepa/denoiser.py:215-290) with the random crop starts [B, nr] and the
    shuffle permutation given: scene at 32 kHz, torchaudio Kaiser-sinc resampling of scene and clean audio, identical
    crops, per-crop normalisation -> (generated, clean) [B*nr, 1, target_length] fp32 (the caller casts to bf16)."""
    import torchaudio

    audio, source_rir, noise, noise_length, noise_start_idx, noise_rirs, snr = batch
    scene = generate_scene(source_rir, noise_rirs, audio, noise, noise_length, noise_start_idx, snr)
    clean = audio.unsqueeze(1)
    rs = lambda x: torchaudio.functional.resample(x, 32000, sr, lowpass_filter_width=64, rolloff=0.9475937167399596,
                                                  resampling_method="sinc_interp_kaiser", beta=14.769656459379492)
    if sr != 32000:
        scene, clean = rs(scene), rs(clean)
    gen = crop_normalise(scene, starts, target_length)[perm]
    cln = crop_normalise(clean, starts, target_length)[perm]
    return gen, cln
