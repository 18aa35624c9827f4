"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by tests/golden/make_golden.py, the parity tests,
__graft_entry__.smoke() and bench.py's CPU arm (N(0,1) 16 kHz noise clips, random crops, masks under
the seed contract).  Pure torch-CPU / numpy; identical on every machine."""
from __future__ import annotations

import numpy as np
import torch

from . import jepa_oracle as jo
from . import masks_oracle as mo


def subsample(t: torch.Tensor, n: int = 2048) -> np.ndarray:
    """Deterministic strided sample of a tensor (goldens store these instead of full tensors)."""
    flat = t.detach().float().reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].numpy().copy()


def training_inputs(cfg: jo.Cfg, n_clips: int, crops: int, seed: int = 1234, clip_len: int = 160000,
                    masker: str = "audioset", row0: int = 0):
    """-> dict(audio [n_clips*crops, C, L] fp32 already rounded to bf16 (what the reference's GPU hook emits,
    jepa.py:311), clips, starts, ctx_masks, target_indices, ctx_and_target_masks as torch.bool)."""
    g = torch.Generator().manual_seed(seed)
    clips = torch.randn(n_clips, cfg.in_channels, clip_len, generator=g)
    starts = torch.randint(0, clip_len - cfg.target_length + 1, (n_clips, crops), generator=g)
    audio = jo.crop_normalise(clips, starts, cfg.target_length).bfloat16().float()
    B = n_clips * crops
    T = cfg.total_patches
    if masker == "audioset":      # configs/masker/AudioSet.yaml
        c, t, v, att = mo.time_inverse_masks(seed, row0, B, T, in_channels=cfg.in_channels if cfg.per_channel else 1,
                                             n_targets=4, ctx_prob=0.65, ctx_len=10, tgt_prob=0.25, tgt_len=10,
                                             cutoff=0.1, channel_based=cfg.per_channel)
    elif masker == "librispeech":  # configs/masker/LibriSpeech.yaml
        c, t, v, att = mo.speech_masks(seed, row0, B, T, n_targets=4, tgt_prob=0.1, tgt_len=10, cutoff=0.5,
                                       min_context_len=5)
    else:
        raise ValueError(masker)
    return dict(audio=audio, clips=clips, starts=starts, ctx_masks=torch.from_numpy(c),
                target_indices=torch.from_numpy(t), ctx_and_target_masks=torch.from_numpy(v), attempts=att)


def hear_inputs(n_clips: int, n_samples: int, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n_clips, n_samples, generator=g) * 2 - 1


def denoiser_batch(B: int = 2, T32: int = 80000, rir_len: int = 4000, n_noise: int = 2, seed: int = 5):
    """Synthetic batch with the tuple layout of Denoiser.on_after_batch_transfer (wavjepa/denoiser.py:229-237):
    (audio [B, T32] @ 32 kHz, source_rir [B, 2, R], noise [B, T32], noise_length [B], noise_start_idx [B],
    noise_rirs [B, S, 2, R], snr [B]).  RIRs are exponentially decaying noise with a unit direct path."""
    g = torch.Generator().manual_seed(seed)
    audio = torch.rand(B, T32, generator=g) * 2 - 1
    decay = torch.exp(-torch.arange(rir_len) / (rir_len / 6.0))
    source_rir = torch.randn(B, 2, rir_len, generator=g) * decay * 0.1
    source_rir[..., 0] = 1.0
    noise_rirs = torch.randn(B, n_noise, 2, rir_len, generator=g) * decay * 0.1
    noise_rirs[..., 0] = 0.7
    noise_start = torch.randint(0, T32 // 4, (B,), generator=g)
    noise_len = torch.randint(T32 // 4, T32 // 2, (B,), generator=g)
    noise = torch.zeros(B, T32)
    for b in range(B):   # the loader places (and fades) the noise inside its window; zeros elsewhere
        seg = torch.randn(int(noise_len[b]), generator=g) * 0.3
        noise[b, int(noise_start[b]):int(noise_start[b]) + int(noise_len[b])] = seg
    snr = torch.tensor([5.0, 15.0, 0.0, 25.0])[:B].clone() if B <= 4 else torch.rand(B, generator=g) * 30 - 5
    return audio, source_rir, noise, noise_len, noise_start, noise_rirs, snr
